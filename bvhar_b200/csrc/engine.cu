// engine.cu — host side of the Gibbs engine and the C-ABI of include/bvhar_cuda.h.
//
// Replaces the runners of the reference (McmcRun::fit/runGibbs, triangular.h:1217-1284;
// McmcOutforecastRun::runGibbs, forecaster.h:755-790): instead of one OpenMP task per chain, every
// (window, chain) unit advances together, one fixed kernel sequence per sweep on one CUDA stream,
// captured once into a CUDA graph and replayed.  There is no CPU path: every entry point fails with
// BVHAR_ERR_CUDA when no device is usable.
#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: ranges are no-ops unless a tool (ncu --nvtx, Nsight Systems) is attached
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <atomic>
#include <vector>

#include "../../include/bvhar_cuda.h"
#include "engine.cuh"
#include "gemm_tn.cuh"
#include "kernels.h"

using namespace bvhar;

namespace {

thread_local std::string g_err;
// Several engines may be driven from different host threads (one per GPU, or several on one GPU).  Graph launches and
// stream syncs run concurrently; what is serialised is everything that uses device-wide synchronisation or stream 0
// (engine construction / destruction, one-shot helpers) and stream capture, which a device-wide sync issued by another
// thread would invalidate.
std::recursive_mutex g_api_mutex;
#define API_LOCK std::lock_guard<std::recursive_mutex> api_lock__(g_api_mutex)
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CU(expr)                                                                                         \
  do {                                                                                                   \
    cudaError_t e__ = (expr);                                                                            \
    if (e__ != cudaSuccess)                                                                              \
      return fail(BVHAR_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));                  \
  } while (0)

// Device buffers come from the device's default stream-ordered memory pool with an unlimited release
// threshold: the GBs of state / record storage freed by one fit are handed straight to the next one instead of
// going back to the driver (cudaFree + cudaMalloc of that much memory costs ~0.2 s per fit).
inline void pool_setup(int device) {
  static bool done[64] = {};
  if (device < 0 || device >= 64 || done[device]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    unsigned long long thr = ~0ULL;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  done[device] = true;
}
// Grow-only page-locked staging buffer for the chains' log-volatility inits (nullptr if it cannot be allocated).
inline double* lvol_stage_buffer(size_t count) {
  static double* buf = nullptr;
  static size_t cap = 0;
  if (count > cap) {
    if (buf) cudaFreeHost(buf);
    buf = nullptr; cap = 0;
    void* p = nullptr;
    if (cudaHostAlloc(&p, count * sizeof(double), cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    buf = static_cast<double*>(p); cap = count;
  }
  return buf;
}
// Host staging vectors of an engine's set-up are recycled through a small process-wide pool: a fresh 10 MB std::vector
// costs more in page faults than the loop that fills it (the set-up of 1 024 chains stages ~50 MB).
// NVTX range around one updater's launches: visible to `ncu --nvtx` / Nsight Systems on the eager path (set_profiling or the
// first sweep of each kind; a replayed CUDA graph carries no host ranges).  Without a tool attached the calls are no-ops.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

struct HostStagePool {
  std::mutex m;
  std::vector<std::vector<double>> free;
  // Blocks of pooled vectors are page-locked once (cudaHostRegister) and stay so while the pool owns them: the set-up's
  // uploads become true asynchronous DMAs at the link's rate (a pageable 10 MB cudaMemcpyAsync is staged through the driver's
  // bounce buffer at a third of it, and blocks the host meanwhile).  BVHAR_NO_PIN_STAGE=1 keeps them pageable.
  std::map<const double*, size_t> pinned;
  const bool pin_enabled = getenv("BVHAR_NO_PIN_STAGE") == nullptr;
  void pin(std::vector<double>& v) {
    if (!pin_enabled || v.capacity() < (1u << 16)) return;
    std::lock_guard<std::mutex> lk(m);
    if (pinned.count(v.data())) return;
    if (cudaHostRegister(v.data(), v.capacity() * sizeof(double), cudaHostRegisterPortable) == cudaSuccess) pinned[v.data()] = v.capacity();
    else cudaGetLastError();  // stays pageable
  }
  void unpin_locked(const std::vector<double>& v) {
    auto it = pinned.find(v.data());
    if (it == pinned.end()) return;
    cudaHostUnregister(const_cast<double*>(it->first));
    pinned.erase(it);
  }
  // a vector of n elements; `fill`: every element is set to `value`, otherwise the contents are unspecified (caller overwrites all)
  std::vector<double> take(size_t n, bool fill, double value = 0.0) {
    std::vector<double> v;
    {
      std::lock_guard<std::mutex> lk(m);
      size_t best = free.size();
      for (size_t i = 0; i < free.size(); ++i)
        if (free[i].capacity() >= n && (best == free.size() || free[i].capacity() < free[best].capacity())) best = i;
      if (best < free.size()) { v = std::move(free[best]); free.erase(free.begin() + best); }
    }
    if (fill) v.assign(n, value);
    else v.resize(n);
    pin(v);
    return v;
  }
  void give(std::vector<double>&& v) {
    std::lock_guard<std::mutex> lk(m);
    if (v.capacity() >= (1u << 16) && free.size() < 12) { free.push_back(std::move(v)); return; }
    unpin_locked(v);  // leaves the pool: its block is freed by the caller's vector
  }
};
// never destroyed: the blocks stay registered until the process ends (unregistering during static destruction would race
// the CUDA runtime's own teardown)
inline HostStagePool& host_stage_pool() { static HostStagePool* p = new HostStagePool; return *p; }
struct HostStageReturn {  // hands the listed vectors back when the set-up leaves scope (any exit path)
  std::vector<std::vector<double>*> vs;
  ~HostStageReturn() {
    cudaDeviceSynchronize();  // page-locked sources: no transfer may still be reading them (a no-op on the normal path)
    for (auto* v : vs) host_stage_pool().give(std::move(*v));
  }
};

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaError_t alloc(size_t count) {
    n = count;
    if (count == 0) return cudaSuccess;
    cudaError_t e = cudaMallocAsync(&p, count * sizeof(T), 0);
    if (e == cudaSuccess) e = cudaMemsetAsync(p, 0, count * sizeof(T), 0);
    return e;
  }
  cudaError_t upload(const std::vector<T>& h) {
    cudaError_t e = alloc(h.size());
    if (e != cudaSuccess || h.empty()) return e;
    return cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, 0);  // pageable source: staged before return
  }
  void release() {
    if (p) cudaFreeAsync(p, 0);
    p = nullptr;
    n = 0;
  }
  ~DevBuf() { release(); }
};

struct RecordSpec {
  std::string name;
  int cols;
  double** slot;
};

}  // namespace

struct bvhar_fit {
  Dims D{};
  DevPtrs P{};
  PriorConst PC{};
  RecPtrs R{};
  int device = 0;
  int num_iter = 0, num_burn = 0, thin = 1;
  uint32_t rec_mask = 0;
  cudaStream_t st = nullptr, s2 = nullptr, s3 = nullptr;
  cudaEvent_t evf[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evs0 = nullptr;
  cudaGraphExec_t graph[2] = {nullptr, nullptr};
  double last_ms = 0.0;
  int64_t launches = 0;
  int launches_per_sweep[2] = {0, 0};
  uint32_t sweeps_done = 0;
  int rec_filled = 0;
  size_t smem_limit = 0;
  std::vector<int> win_row0, win_n;
  std::vector<long long> win_off;
  // device storage
  DevBuf<double> X, XT, Y, XX, xnorm, XtX, XtY, prior_mean, sig_shp, sig_scl, sv_mean, sv_prec, ssvs_s1, ssvs_s2;
  DevBuf<int> d_row0, d_n, grp_idx, grp_cnt, own_flag, flags, rec_row, gKg, gMg, gz_K;
  DevBuf<long long> d_off, g1_b, g1_c, g2_b, g2_c, g3_a, g3_b, g3_y, gz_a, gz_c;
  DevBuf<uint32_t> seeds, iter;
  DevBuf<double> H, Z, Dinv, YLD, S, Cm, Acoef, alpha, cvec, prec, contem, L, chol_prec, sparse_coef, sparse_contem;
  DevBuf<double> Zn, lvol_init, lvol_sig, diag, znorm, hN[4], hG[3], hq[4], hs, vscratch, pscratch, gscratch, Pfac, hscr, Qn, Wt;
  bool coef_v2 = false, coef_big = false;
  bool direct_p = false;  // large-design SV path: G1 contracts the mixed weights and writes the P_j themselves (no S_i, no pbuild)
  std::map<std::string, std::unique_ptr<DevBuf<double>>> rec;
  std::vector<RecordSpec> rec_order;
  GemmArgs g1{}, g2{}, g3{}, gz{}, g4{};  // gz: Z'Z per unit (LDLT contemporaneous draw + state); g4: R = X'Y - X'X A (LDLT, large designs)
  DevBuf<double> XtXf, Rres;              // full X'X per window (pitch even(d)) and R [W][d][ldM]
  DevBuf<double> ws1, ws2;                // split-K slices of G1 / G2 (small batches)
  long long g1_count = 0, g2_count = 0;   // doubles of C the slices cover
  DevBuf<long long> g4_a, g4_y;
  bool ldlt_res = false;                  // the large-design LDLT sequential kernel starts from R instead of k symmetric products
  bool g3_aligned = true;
  bool timed = false;
  // per-kernel profiling (eager launches bracketed by CUDA events on the engine's stream)
  bool profiling = false;
  struct ProfEv { int id; cudaEvent_t a, b; };
  std::vector<ProfEv> prof_pending;
  std::vector<cudaEvent_t> prof_pool;
  std::vector<std::string> prof_names;
  std::vector<double> prof_ms;
  std::vector<long long> prof_cnt;
  int prof_id(const char* nm) {
    for (size_t i = 0; i < prof_names.size(); ++i) if (prof_names[i] == nm) return (int)i;
    prof_names.push_back(nm); prof_ms.push_back(0.0); prof_cnt.push_back(0);
    return (int)prof_names.size() - 1;
  }
  cudaEvent_t prof_event() {
    if (!prof_pool.empty()) { cudaEvent_t e = prof_pool.back(); prof_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
  void prof_collect() {
    for (auto& pe : prof_pending) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, pe.a, pe.b) == cudaSuccess) { prof_ms[pe.id] += ms; prof_cnt[pe.id] += 1; }
      else cudaGetLastError();
      prof_pool.push_back(pe.a); prof_pool.push_back(pe.b);
    }
    prof_pending.clear();
  }

  ~bvhar_fit() {
    cudaSetDevice(device);
    if (st) cudaStreamSynchronize(st);  // the buffers go back to the pool in stream order of stream 0
    for (auto& g : graph)
      if (g) cudaGraphExecDestroy(g);
    for (auto& pe : prof_pending) { cudaEventDestroy(pe.a); cudaEventDestroy(pe.b); }
    for (auto& e : prof_pool) cudaEventDestroy(e);
    for (auto& e : evf) if (e) cudaEventDestroy(e);
    if (s2) cudaStreamDestroy(s2);
    if (s3) cudaStreamDestroy(s3);
    if (sc) { cudaStreamSynchronize(sc); cudaStreamDestroy(sc); }
    if (evc) cudaEventDestroy(evc);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (evs0) cudaEventDestroy(evs0);
    if (st) cudaStreamDestroy(st);
  }

  // record sinks: host buffers that receive record rows while later sweeps are still running (copy engine)
  struct Sink { const RecordSpec* rec; double* host; int64_t rows; int copied; };
  std::vector<Sink> sinks;
  cudaStream_t sc = nullptr;
  cudaEvent_t evc = nullptr;
  int stream_out() {
    if (sinks.empty()) return 0;
    CU(cudaSetDevice(device));
    if (!sc) { CU(cudaStreamCreateWithFlags(&sc, cudaStreamNonBlocking)); CU(cudaEventCreateWithFlags(&evc, cudaEventDisableTiming)); }
    CU(cudaEventRecord(evc, st));
    CU(cudaStreamWaitEvent(sc, evc, 0));
    for (auto& k : sinks) {
      const int from = k.copied, to = rec_filled;
      if (to <= from) continue;
      const size_t cols = (size_t)k.rec->cols, row_b = cols * sizeof(double);
      const size_t spitch = (size_t)R.cap * row_b, dpitch = (size_t)k.rows * row_b;
      const double* dev = *k.rec->slot;
      if (thin == 1 && spitch < (1ull << 31) && dpitch < (1ull << 31)) {
        CU(cudaMemcpy2DAsync(k.host + (size_t)from * cols, dpitch, dev + (size_t)from * cols, spitch, (size_t)(to - from) * row_b, (size_t)D.U,
                             cudaMemcpyDeviceToHost, sc));
      } else {
        for (int dr = from; dr < to; ++dr) {
          if (dr % thin) continue;
          const size_t i = (size_t)(dr / thin);
          if (spitch < (1ull << 31) && dpitch < (1ull << 31)) {
            CU(cudaMemcpy2DAsync(k.host + i * cols, dpitch, dev + (size_t)dr * cols, spitch, row_b, (size_t)D.U, cudaMemcpyDeviceToHost, sc));
          } else {
            for (int u = 0; u < D.U; ++u)
              CU(cudaMemcpyAsync(k.host + ((size_t)u * k.rows + i) * cols, dev + ((size_t)u * R.cap + dr) * cols, row_b, cudaMemcpyDeviceToHost, sc));
          }
        }
      }
      k.copied = to;
    }
    return 0;
  }

  int enqueue_sweep(bool record, bool count) {
    int nl = 0;
#define LAUNCH(name, expr)                                                   \
    do {                                                                       \
      cudaEvent_t ea__ = nullptr, eb__ = nullptr;                              \
      if (profiling) { ea__ = prof_event(); eb__ = prof_event(); cudaEventRecord(ea__, st); } \
      NvtxRange nvtx__(name);  /* one range per updater launch */ \
      CU(expr);                                                                \
      if (profiling) { cudaEventRecord(eb__, st); prof_pending.push_back({prof_id(name), ea__, eb__}); } \
      ++nl;                                                                    \
    } while (0)
    // Independent updaters run on side streams (parallel branches of the captured graph): the prior hyper-parameter
    // kernels (stages 1 and 5 read only the PREVIOUS sweep's coefficients) overlap the SV weights + Gram contractions,
    // and the small G2 contraction fills the tail of G1.  Per-kernel profiling keeps everything on one stream.
    const bool fork = !profiling && s2 && s3;
    cudaStream_t sp = fork ? s2 : st, sg = fork ? s3 : st;
    LAUNCH("bump", launch_bump(P, 1, 0, st));
    if (fork) { CU(cudaEventRecord(evf[0], st)); CU(cudaStreamWaitEvent(s2, evf[0], 0)); }
    if (D.prior != 1) {
      LAUNCH("prior_coef", launch_prior(D, P, PC, 0, sp));
      LAUNCH("prior_contem", launch_prior(D, P, PC, 1, sp));
    }
    if (fork) CU(cudaEventRecord(evf[1], s2));
    if (D.sv) {
      // the normals of the log-volatility draw depend only on the sweep counter: drawn on the side branch, consumed by sv_solve
      LAUNCH("sv_normals", launch_sv_state(3, D, P, sp));
      if (fork) CU(cudaEventRecord(evf[6], s2));
      // (updateSv, stage 3, was already done at the end of the previous sweep - see below - or at create time)
      if (fork) { CU(cudaEventRecord(evf[2], st)); CU(cudaStreamWaitEvent(s3, evf[2], 0)); }
      if (direct_p) LAUNCH("sv_wmix", launch_sv_wmix(D, P, Wt.p, st));
      else {
        LAUNCH("gram_C_dmma", [&] { const cudaError_t e = launch_dgemm_tn(g2, sg); return e != cudaSuccess ? e : launch_splitk_reduce(g2, g2_count, sg); }());
        if (g2.ksplit > 1) ++nl;
      }
      if (fork) CU(cudaEventRecord(evf[3], s3));
      LAUNCH("gram_S_dmma", [&] { const cudaError_t e = launch_dgemm_tn(g1, st); return e != cudaSuccess ? e : launch_splitk_reduce(g1, g1_count, st); }());
      if (g1.ksplit > 1) ++nl;
    }
    if (coef_big) {
      if (D.sv && !direct_p) LAUNCH("coef_pbuild", launch_coef_v2(0, D, P, Pfac.p, Qn.p, smem_limit, st));
      if (fork) { CU(cudaStreamWaitEvent(st, evf[1], 0)); if (D.sv) CU(cudaStreamWaitEvent(st, evf[3], 0)); }
      LAUNCH("coef_chol_cta", launch_coef_big(1, D, P, Pfac.p, hscr.p, smem_limit, st)); ++nl;  // + invert_diag_kernel
      if (ldlt_res) LAUNCH("gram_R_dmma", launch_dgemm_tn(g4, st));
      LAUNCH("coef_seq_big", launch_coef_big(2, D, P, Pfac.p, hscr.p, smem_limit, st, ldlt_res ? Rres.p : nullptr));
    } else if (coef_v2) {
      if (D.sv) LAUNCH("coef_pbuild", launch_coef_v2(0, D, P, Pfac.p, Qn.p, smem_limit, st));
      if (fork) { CU(cudaStreamWaitEvent(st, evf[1], 0)); if (D.sv) CU(cudaStreamWaitEvent(st, evf[3], 0)); }
      LAUNCH("coef_chol", launch_coef_v2(1, D, P, Pfac.p, Qn.p, smem_limit, st));
      LAUNCH("coef_seq", launch_coef_v2(2, D, P, Pfac.p, Qn.p, smem_limit, st));
    } else {
      if (fork) { CU(cudaStreamWaitEvent(st, evf[1], 0)); if (D.sv) CU(cudaStreamWaitEvent(st, evf[3], 0)); }
      LAUNCH("coef_draw", launch_coef(D, P, vscratch.p, pscratch.p, smem_limit, st));
    }
    LAUNCH("latent_dmma", launch_dgemm_tn(g3, st));
    if (!D.sv && !ldlt_own_gram(D)) LAUNCH("gram_Z_dmma", launch_dgemm_tn(gz, st));
    if (D.k > 1) { LAUNCH("impact_draw", launch_impact(D, P, gscratch.p, st)); nl += impact_launches(D) - 1; }
    else if (!D.sv) LAUNCH("gram_Z", launch_ldlt_gram_small(D, P, gscratch.p, st));  // univariate LDLT: the state draw still reads Z'Z
    if (D.sv) {
      LAUNCH("sv_mix", launch_sv_state(0, D, P, st));
      if (fork) CU(cudaStreamWaitEvent(st, evf[6], 0));
      LAUNCH("sv_solve", launch_sv_state(1, D, P, st));
      // updateSv of the NEXT sweep (Dinv = exp(-h), YLD) depends only on h and L, both final here: it runs beside
      // sv_finish / the record copy instead of at the head of the next sweep's critical path
      if (fork) { CU(cudaEventRecord(evf[4], st)); CU(cudaStreamWaitEvent(s2, evf[4], 0)); }
      LAUNCH("sv_prep", launch_sv_prep(D, P, sp));
      if (fork) CU(cudaEventRecord(evf[5], s2));
      LAUNCH("sv_finish", launch_sv_state(2, D, P, st));
    }
    else LAUNCH("ldlt_state", launch_ldlt_state(D, P, gscratch.p, st));
    if (record) {
      LAUNCH("record", launch_record(D, P, R, st));
      LAUNCH("bump", launch_bump(P, 0, 1, st));
    }
    if (D.sv && fork) CU(cudaStreamWaitEvent(st, evf[5], 0));
#undef LAUNCH
    if (count) launches_per_sweep[record ? 1 : 0] = nl;
    return 0;
  }

  // n sweeps: the first sweep of each kind runs eagerly (sets function attributes, validates the
  // launch configuration), the sequence is then captured once and replayed.
  int step(int n, bool record) {
    CU(cudaSetDevice(device));
    const int gi = record ? 1 : 0;
    CU(cudaEventRecord(ev0, st));
    for (int s = 0; s < n; ++s) {
      if (record && rec_filled >= R.cap) return fail(BVHAR_ERR_BAD_ARG, "record buffers are full (num_iter - num_burn draws)");
      if (profiling || (!graph[gi] && launches_per_sweep[gi] == 0)) {
        int rc = enqueue_sweep(record, true);
        if (rc) return rc;
      } else {
        if (!graph[gi]) {
          API_LOCK;
          cudaGraph_t g;
          CU(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
          int rc = enqueue_sweep(record, false);
          cudaError_t ce = cudaStreamEndCapture(st, &g);
          if (rc) return rc;
          CU(ce);
          CU(cudaGraphInstantiate(&graph[gi], g, 0));
          cudaGraphDestroy(g);
        }
        CU(cudaGraphLaunch(graph[gi], st));
      }
      launches += launches_per_sweep[gi];
      ++sweeps_done;
      if (record) ++rec_filled;
    }
    CU(cudaEventRecord(ev1, st));
    timed = true;
    return 0;
  }
  int sync() {
    CU(cudaSetDevice(device));
    CU(cudaStreamSynchronize(st));
    if (sc) CU(cudaStreamSynchronize(sc));
    prof_collect();
    if (timed) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, ev0, ev1) == cudaSuccess) last_ms = ms;
      else cudaGetLastError();
    }
    std::vector<int> fl(D.U);
    CU(cudaMemcpy(fl.data(), flags.p, sizeof(int) * D.U, cudaMemcpyDeviceToHost));
    for (int u = 0; u < D.U; ++u)
      if (fl[u]) {
        char buf[160];
        snprintf(buf, sizeof buf, "non-positive-definite precision in unit %d (window %d, chain %d), flag %d", u, u / D.C, u % D.C, fl[u]);
        return fail(BVHAR_ERR_NUMERICAL, buf);
      }
    return 0;
  }
};

namespace {

int round_even(int x) { return (x + 1) & ~1; }

int build(const bvhar_fit_desc* d, bvhar_fit** out) {
  if (!d || !out) return fail(BVHAR_ERR_BAD_ARG, "null descriptor");
  API_LOCK;
  if (d->abi_version != BVHAR_CUDA_ABI_VERSION) return fail(BVHAR_ERR_BAD_ARG, "descriptor ABI version mismatch");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(BVHAR_ERR_CUDA, "no CUDA device available: the bvhar_b200 engine has no CPU fallback");
  if (d->device < 0 || d->device >= ndev) return fail(BVHAR_ERR_BAD_ARG, "device ordinal out of range");
  if (d->n_windows < 1 || d->n_chains < 1 || d->dim < 1 || d->dim_design < 1) return fail(BVHAR_ERR_BAD_ARG, "empty problem");
  if (d->prior_type < 1 || d->prior_type > 7) return fail(BVHAR_ERR_BAD_ARG, "prior_type must be 1..7");
  if (d->num_burn < 0 || d->num_iter < d->num_burn) return fail(BVHAR_ERR_BAD_ARG, "need 0 <= num_burn <= num_iter");
  if (!d->x || !d->y || !d->win_row0 || !d->win_nrow || !d->grp_mat || !d->grp_id || !d->init || !d->seed_chain || !d->sig_shape || !d->sig_scale)
    return fail(BVHAR_ERR_BAD_ARG, "missing required pointer in descriptor");
  CU(cudaSetDevice(d->device));
  pool_setup(d->device);
  // two attributes, cached per device: cudaGetDeviceProperties fills the whole property struct on every call and showed up
  // as tens of milliseconds of an engine's set-up when the device was busy
  static int cc_major[64] = {}, smem_optin[64] = {};
  const int dv = d->device;
  if (dv < 0 || dv >= 64) return fail(BVHAR_ERR_BAD_ARG, "device index out of range");
  if (!cc_major[dv]) {
    CU(cudaDeviceGetAttribute(&cc_major[dv], cudaDevAttrComputeCapabilityMajor, dv));
    CU(cudaDeviceGetAttribute(&smem_optin[dv], cudaDevAttrMaxSharedMemoryPerBlockOptin, dv));
  }
  if (cc_major[dv] < 10) return fail(BVHAR_ERR_CUDA, "bvhar_b200 is built for sm_100a (B200) only");

  const bool timing = getenv("BVHAR_TIMING") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto t_last = now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    cudaDeviceSynchronize();
    auto t = now();
    fprintf(stderr, "[bvhar_cuda_fit_create] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t - t_last).count());
    t_last = t;
  };
  std::unique_ptr<bvhar_fit> F(new bvhar_fit);
  F->device = d->device;
  F->smem_limit = (size_t)smem_optin[dv];
  F->num_iter = d->num_iter; F->num_burn = d->num_burn; F->thin = d->thin < 1 ? 1 : d->thin;
  F->rec_mask = d->record_mask;
  Dims& D = F->D;
  D.W = d->n_windows; D.C = d->n_chains; D.U = D.W * D.C;
  D.k = d->dim; D.d = d->dim_design; D.mean = d->include_mean != 0;
  D.sv = d->cov_type == BVHAR_COV_SV; D.ggl = d->is_group != 0; D.prior = d->prior_type;
  D.q = D.k * (D.k - 1) / 2;
  D.N = D.k * D.d - (D.mean ? D.k : 0);
  D.nrow = D.N / D.k;
  D.G = d->n_grp;
  D.ubase = d->unit_id_base;
  D.Pp = D.d * (D.d + 1) / 2;
  D.ldS = round_even(D.Pp);
  D.ldX = round_even(D.d);
  D.ldXX = round_even(D.Pp);
  D.T = (int)d->n_rows_total;
  D.ldT = round_even(D.T);
  D.M = D.C * D.k;
  D.ldM = round_even(D.M);
  if (D.G < 1 || D.G > 64) return fail(BVHAR_ERR_UNSUPPORTED, "number of coefficient groups must be in 1..64");
  if (D.nrow * D.k != D.N) return fail(BVHAR_ERR_BAD_ARG, "dim_design inconsistent with dim / include_mean");
  const bvhar_prior_params& pp = d->prior;
  PriorConst& PC = F->PC;
  PC.hmn_shape = pp.hmn_shape; PC.hmn_rate = pp.hmn_rate; PC.hmn_grid = pp.hmn_grid_size;
  PC.ssvs_chol_s1 = pp.ssvs_chol_s1; PC.ssvs_chol_s2 = pp.ssvs_chol_s2;
  PC.ssvs_coef_slab_shape = pp.ssvs_coef_slab_shape; PC.ssvs_coef_slab_scl = pp.ssvs_coef_slab_scl;
  PC.ssvs_chol_slab_shape = pp.ssvs_chol_slab_shape; PC.ssvs_chol_slab_scl = pp.ssvs_chol_slab_scl;
  PC.ssvs_coef_grid = pp.ssvs_coef_grid; PC.ssvs_chol_grid = pp.ssvs_chol_grid;
  PC.ng_shape_sd = pp.ng_shape_sd; PC.ng_group_shape = pp.ng_group_shape; PC.ng_group_scale = pp.ng_group_scale;
  PC.ng_global_shape = pp.ng_global_shape; PC.ng_global_scale = pp.ng_global_scale;
  PC.ng_cg_shape = pp.ng_contem_global_shape; PC.ng_cg_scale = pp.ng_contem_global_scale;
  PC.dl_grid = pp.dl_grid_size; PC.dl_shape = pp.dl_shape; PC.dl_scale = pp.dl_scale;
  PC.gdp_grid_shape = pp.gdp_grid_shape; PC.gdp_grid_rate = pp.gdp_grid_rate;
  const int maxgrid = std::max({PC.hmn_grid, PC.ssvs_coef_grid, PC.ssvs_chol_grid, PC.dl_grid, PC.gdp_grid_shape, PC.gdp_grid_rate});
  if (maxgrid > 512) return fail(BVHAR_ERR_UNSUPPORTED, "griddy-Gibbs grids are limited to 512 points");
  auto need_grid = [&](int g, const char* nm) { return g < 1 ? fail(BVHAR_ERR_BAD_ARG, std::string(nm) + " must be >= 1") : 0; };
  if (D.prior == 2 && (need_grid(PC.ssvs_coef_grid, "coef_grid") || need_grid(PC.ssvs_chol_grid, "chol_grid"))) return BVHAR_ERR_BAD_ARG;
  if (D.prior == 6 && need_grid(PC.dl_grid, "grid_size")) return BVHAR_ERR_BAD_ARG;
  if (D.prior == 7 && (need_grid(PC.gdp_grid_shape, "grid_shape") || need_grid(PC.gdp_grid_rate, "grid_rate"))) return BVHAR_ERR_BAD_ARG;

  // windows
  F->win_row0.assign(d->win_row0, d->win_row0 + D.W);
  F->win_n.assign(d->win_nrow, d->win_nrow + D.W);
  F->win_off.resize(D.W);
  long long off = 0;
  D.nmax = 0;
  for (int w = 0; w < D.W; ++w) {
    if (F->win_n[w] < 1 || F->win_row0[w] < 0 || F->win_row0[w] + F->win_n[w] > D.T)
      return fail(BVHAR_ERR_BAD_ARG, "window rows outside the design matrix");
    F->win_off[w] = off;
    off += (long long)F->win_n[w] * D.ldM;
    D.nmax = std::max(D.nmax, F->win_n[w]);
  }
  const long long tm_total = off;

  const int k = D.k, dd = D.d, N = D.N, q = D.q, G = D.G, U = D.U, nrow = D.nrow;
  const int64_t ld = d->n_rows_total;
  // ---- shared data ----
  std::vector<double> hX((size_t)D.T * D.ldX, 0.0), hXT((size_t)dd * D.ldT, 0.0), hY((size_t)D.T * k);
  for (int a = 0; a < dd; ++a)
    for (int t = 0; t < D.T; ++t) {
      const double v = d->x[(size_t)a * ld + t];
      hX[(size_t)t * D.ldX + a] = v;
      hXT[(size_t)a * D.ldT + t] = v;
    }
  for (int j = 0; j < k; ++j)
    for (int t = 0; t < D.T; ++t) hY[(size_t)t * k + j] = d->y[(size_t)j * ld + t];
  CU(F->X.upload(hX)); CU(F->XT.upload(hXT)); CU(F->Y.upload(hY));
  std::vector<double> hxn((size_t)D.W * dd, 0.0);
  for (int w = 0; w < D.W; ++w)
    for (int a = 0; a < dd; ++a) {
      double s = 0;
      const double* col = d->x + (size_t)a * ld + F->win_row0[w];
      for (int t = 0; t < F->win_n[w]; ++t) s += col[t] * col[t];
      hxn[(size_t)w * dd + a] = s;
    }
  CU(F->xnorm.upload(hxn));
  CU(F->d_row0.upload(F->win_row0)); CU(F->d_n.upload(F->win_n)); CU(F->d_off.upload(F->win_off));
  // groups
  std::vector<int> gidx(N), gcnt(G, 0), ownf(N, 0);
  std::set<int> own(d->own_id, d->own_id + d->n_own);
  for (int e = 0; e < N; ++e) {
    int g = -1;
    for (int t = 0; t < G; ++t)
      if (d->grp_id[t] == d->grp_mat[e]) { g = t; break; }
    if (g < 0) return fail(BVHAR_ERR_BAD_ARG, "grp_mat holds an id that is not in grp_id");
    gidx[e] = g;
    ++gcnt[g];
    ownf[e] = own.count(d->grp_mat[e]) ? 1 : 0;
  }
  CU(F->grp_idx.upload(gidx)); CU(F->grp_cnt.upload(gcnt)); CU(F->own_flag.upload(ownf));
  // prior mean / fixed precisions, internal (j*d + a) layout
  std::vector<double> pmean((size_t)k * dd, 0.0), pprec0((size_t)k * dd, 0.0);
  const bool minn = D.prior == 1 || D.prior == 4;
  if (minn && (!pp.minn_prior_mean || !pp.minn_prior_prec)) return fail(BVHAR_ERR_BAD_ARG, "Minnesota prior moments missing");
  for (int j = 0; j < k; ++j) {
    for (int a = 0; a < nrow; ++a)
      if (minn) { pmean[(size_t)j * dd + a] = pp.minn_prior_mean[j * nrow + a]; pprec0[(size_t)j * dd + a] = pp.minn_prior_prec[j * nrow + a]; }
    if (D.mean) {
      if (!d->mean_non) return fail(BVHAR_ERR_BAD_ARG, "mean_non missing");
      pmean[(size_t)j * dd + nrow] = d->mean_non[j];
      pprec0[(size_t)j * dd + nrow] = 1.0 / (d->sd_non * d->sd_non);
    }
  }
  CU(F->prior_mean.upload(pmean));
  CU(F->sig_shp.upload(std::vector<double>(d->sig_shape, d->sig_shape + k)));
  CU(F->sig_scl.upload(std::vector<double>(d->sig_scale, d->sig_scale + k)));
  if (D.sv) {
    if (!d->sv_init_mean || !d->sv_init_prec) return fail(BVHAR_ERR_BAD_ARG, "SV prior (initial_mean / initial_prec) missing");
    CU(F->sv_mean.upload(std::vector<double>(d->sv_init_mean, d->sv_init_mean + k)));
    std::vector<double> pr((size_t)k * k);
    for (int i = 0; i < k; ++i)
      for (int m = 0; m < k; ++m) pr[(size_t)i * k + m] = d->sv_init_prec[(size_t)m * k + i];
    CU(F->sv_prec.upload(pr));
  }
  if (D.prior == 2) {
    if (!pp.ssvs_coef_s1 || !pp.ssvs_coef_s2) return fail(BVHAR_ERR_BAD_ARG, "SSVS coef_s1 / coef_s2 missing");
    CU(F->ssvs_s1.upload(std::vector<double>(pp.ssvs_coef_s1, pp.ssvs_coef_s1 + G)));
    CU(F->ssvs_s2.upload(std::vector<double>(pp.ssvs_coef_s2, pp.ssvs_coef_s2 + G)));
  }
  std::vector<uint32_t> hseed(d->seed_chain, d->seed_chain + U);
  CU(F->seeds.upload(hseed));
  CU(F->iter.alloc(1)); CU(F->rec_row.alloc(1)); CU(F->flags.alloc(U));

  lap("shared data + uploads");
  // ---- per-unit state from the chain inits (McmcTriangular ctor tri.h:40-71 and prior ctors) ----
  HostStagePool& hpool = host_stage_pool();
  std::vector<double> hA = hpool.take((size_t)D.W * dd * D.ldM, true), halpha = hpool.take((size_t)U * N, false), hc((size_t)U * k, 0.0),
                      hprec = hpool.take((size_t)U * k * dd, false);
  std::vector<double> hcontem((size_t)U * q), hL = hpool.take((size_t)U * k * k, true), hcp((size_t)U * q, 1.0);
  std::vector<double> hH, hli, hls, hdiag;
  // SV: the n x k `lvol` blocks of the chains are staged series-major ([chain][series][t], a plain memcpy per
  // chain) and transposed into the time-major layout on the device; replicated inits are k values per unit.
  std::vector<double> hlv;     // [C][k][n0] when explicit lvol inits are given (fallback storage)
  double* hlv_stage = nullptr; // the same, page-locked
  bool lvol_explicit = false;
  if (D.sv) { hli.resize((size_t)U * k); hls.resize((size_t)U * k); }
  else hdiag.resize((size_t)U * k);
  std::vector<double> hhN[4], hhG[3], hhq[4], hhs((size_t)U * HS_SLOTS, 0.0);
  HostStageReturn hret{{&hA, &halpha, &hprec, &hL, &hhN[0], &hhN[1]}};
  // (the N-long prior vectors are materialised on the host only for the slots a prior actually initialises: a fresh
  // 10 MB vector costs more in page faults than its upload; untouched slots are zero-filled on the device)
  for (auto& v : hhG) v.assign((size_t)U * G, 0.0);
  for (auto& v : hhq) v.assign((size_t)U * q, 0.0);
  {  // which N-long slots the prior initialises (allocated up front: the unit loop below runs on several threads)
    const bvhar_chain_init& in0 = d->init[0];
    auto useN = [&](int slot, const double* src) { if (src) hhN[slot] = hpool.take((size_t)U * N, false); };  // every unit copies all N
    switch (D.prior) {
      case 2: useN(0, in0.coef_dummy); useN(1, in0.coef_slab); break;
      case 3: case 5: case 6: case 7: useN(0, in0.local_sparsity); break;
      default: break;
    }
  }
  std::mutex err_mutex;
  int err_code = 0;
  std::string err_msg;
  auto unit_fail = [&](int code, const std::string& msg) {
    std::lock_guard<std::mutex> lk(err_mutex);
    if (!err_code) { err_code = code; err_msg = msg; }
    return code;
  };
  auto init_unit = [&](int w, int c) -> int {
      const int u = w * D.C + c;
      const bvhar_chain_init& in = d->init[c];
      if (!in.init_coef || (q > 0 && !in.init_contem)) return unit_fail(BVHAR_ERR_BAD_ARG, "init_coef / init_contem missing");
      for (int j = 0; j < k; ++j)
        for (int a = 0; a < dd; ++a) {
          const double v = in.init_coef[(size_t)j * dd + a];
          hA[((size_t)w * dd + a) * D.ldM + (size_t)c * k + j] = v;
          if (a < nrow) halpha[(size_t)u * N + j * nrow + a] = v;
          else hc[(size_t)u * k + j] = v;
          hprec[(size_t)u * k * dd + j * dd + a] = pprec0[(size_t)j * dd + a];
        }
      for (int e = 0; e < q; ++e) hcontem[(size_t)u * q + e] = in.init_contem[e];
      for (int i = 0; i < k; ++i) {
        hL[((size_t)u * k + i) * k + i] = 1.0;
        for (int m = 0; m < i; ++m) hL[((size_t)u * k + i) * k + m] = in.init_contem[i * (i - 1) / 2 + m];
      }
      if (D.sv) {
        if (!in.lvol_init || !in.lvol_sig) return unit_fail(BVHAR_ERR_BAD_ARG, "lvol_init / lvol_sig missing");
        const bool rep = d->lvol_from_init || !in.lvol;
        if (!rep && F->win_n[w] != F->win_n[0]) return unit_fail(BVHAR_ERR_BAD_ARG, "lvol inits need equal window lengths (use lvol_from_init)");
        if (rep == lvol_explicit) return unit_fail(BVHAR_ERR_BAD_ARG, "either every chain or no chain carries an explicit 'lvol' init");
        for (int i = 0; i < k; ++i) {
          hli[(size_t)u * k + i] = in.lvol_init[i];
          hls[(size_t)u * k + i] = in.lvol_sig[i];
        }
        // (the lvol blocks themselves are gathered below, in parallel)
      } else {
        if (!in.init_diag) return unit_fail(BVHAR_ERR_BAD_ARG, "init_diag missing");
        for (int i = 0; i < k; ++i) hdiag[(size_t)u * k + i] = in.init_diag[i];
      }
      double* s = &hhs[(size_t)u * HS_SLOTS];
      auto cpN = [&](int slot, const double* src) {
        if (!src || hhN[slot].empty()) return;
        std::copy(src, src + N, hhN[slot].begin() + (size_t)u * N);
      };
      auto cpG = [&](int slot, const double* src) { if (src) std::copy(src, src + G, hhG[slot].begin() + (size_t)u * G); };
      auto cpq = [&](int slot, const double* src) { if (src) std::copy(src, src + q, hhq[slot].begin() + (size_t)u * q); };
      auto need = [&](const void* p, const char* nm) { return p ? 0 : unit_fail(BVHAR_ERR_BAD_ARG, std::string("param_init lacks ") + nm); };
      switch (D.prior) {
        case 4:  // tri.h:478-494
          s[HS_OWN_LAMBDA] = in.own_lambda; s[HS_CROSS_LAMBDA] = in.cross_lambda; s[HS_CONTEM_LAMBDA] = in.contem_lambda;
          for (int j = 0; j < k; ++j)
            for (int a = 0; a < nrow; ++a) hprec[(size_t)u * k * dd + j * dd + a] /= in.own_lambda;
          for (int e = 0; e < q; ++e) hcp[(size_t)u * q + e] /= in.contem_lambda;
          break;
        case 2:  // tri.h:563-577
          if (need(in.coef_dummy, "init_coef_dummy") || need(in.coef_mixture, "coef_mixture") || need(in.coef_slab, "coef_slab") ||
              (q > 0 && (need(in.chol_mixture, "chol_mixture") || need(in.contem_slab, "contem_slab")))) return BVHAR_ERR_BAD_ARG;
          cpN(0, in.coef_dummy); cpN(1, in.coef_slab); cpG(0, in.coef_mixture);
          for (int e = 0; e < q; ++e) hhq[0][(size_t)u * q + e] = 1.0;
          cpq(1, in.contem_slab); cpq(2, in.chol_mixture);
          s[HS_SPIKE] = in.coef_spike_scl; s[HS_CONTEM_SPIKE] = in.coef_spike_scl;  // tri.h:569
          break;
        case 3:  // tri.h:659-669
          if (need(in.local_sparsity, "local_sparsity") || need(in.group_sparsity, "group_sparsity") ||
              (q > 0 && need(in.contem_local_sparsity, "contem_local_sparsity"))) return BVHAR_ERR_BAD_ARG;
          cpN(0, in.local_sparsity); cpG(0, in.group_sparsity); cpq(0, in.contem_local_sparsity);
          s[HS_GLOBAL] = D.ggl ? in.global_sparsity : 1.0; s[HS_CONTEM_GLOBAL] = in.contem_global_sparsity;
          break;
        case 5:  // tri.h:757-770
          if (need(in.local_sparsity, "local_sparsity") || need(in.group_sparsity, "group_sparsity") || need(in.local_shape, "local_shape") ||
              (q > 0 && need(in.contem_local_sparsity, "contem_local_sparsity"))) return BVHAR_ERR_BAD_ARG;
          cpN(0, in.local_sparsity); cpG(0, in.group_sparsity); cpG(1, in.local_shape);
          for (int e = 0; e < q; ++e) hhq[0][(size_t)u * q + e] = in.contem_global_sparsity * in.contem_local_sparsity[e];
          s[HS_GLOBAL] = D.ggl ? in.global_sparsity : 1.0; s[HS_CONTEM_GLOBAL] = in.contem_global_sparsity; s[HS_CONTEM_SHAPE] = in.contem_shape;
          break;
        case 6:  // tri.h:852-863
          if (need(in.local_sparsity, "local_sparsity") || (q > 0 && need(in.contem_local_sparsity, "contem_local_sparsity"))) return BVHAR_ERR_BAD_ARG;
          cpN(0, in.local_sparsity); cpq(0, in.contem_local_sparsity);
          s[HS_GLOBAL] = D.ggl ? in.global_sparsity : 1.0; s[HS_CONTEM_GLOBAL] = in.contem_global_sparsity;
          break;
        case 7:  // tri.h:941-951
          if (need(in.local_sparsity, "local_sparsity") || need(in.group_rate, "group_rate") ||
              (q > 0 && (need(in.contem_local_sparsity, "contem_local_sparsity") || need(in.contem_rate, "contem_rate")))) return BVHAR_ERR_BAD_ARG;
          cpN(0, in.local_sparsity); cpG(0, in.group_rate); cpq(0, in.contem_local_sparsity); cpq(1, in.contem_rate);
          s[HS_GSHAPE] = in.gamma_shape; s[HS_GRATE] = in.gamma_rate; s[HS_CGSHAPE] = in.contem_gamma_shape; s[HS_CGRATE] = in.contem_gamma_rate;
          break;
        default: break;
      }
      return 0;
  };
  if (D.sv) {  // explicit `lvol` blocks: every chain or none (decided by chain 0, checked per unit)
    const bvhar_chain_init& in0 = d->init[0];
    lvol_explicit = !(d->lvol_from_init || !in0.lvol);
  }
  // explicit lvol inits: the chains' n x k blocks are gathered into one page-locked staging buffer by a few host threads
  // WHILE the unit loop below fills the other staging vectors (both are memory-bound host work on disjoint data)
  // and each finished slice goes to the device at once (page-locked source: a true asynchronous DMA on stream 0), so the
  // one large transfer of the set-up (164 MB at C3: 6.5 ms of PCIe time) runs under the unit loop instead of after it
  DevBuf<double> lv_stage;  // declared before the thread's guard: on an early return the guard joins first, then this is freed
  std::atomic<int> lvol_copy_err{0};
  std::thread lvol_gather;
  struct JoinGuard { std::thread& t; ~JoinGuard() { if (t.joinable()) t.join(); } } lvol_gather_guard{lvol_gather};
  if (D.sv && lvol_explicit) {
    const size_t blk = (size_t)k * F->win_n[0];
    // staged in a page-locked buffer that lives as long as the process (grow-only, engine construction is serialised):
    // a fresh pageable vector of this size costs more in page faults than the copy itself (66 ms for 164 MB at C3)
    hlv_stage = lvol_stage_buffer((size_t)D.C * blk);
    if (!hlv_stage) { hlv.resize((size_t)D.C * blk); hlv_stage = hlv.data(); }
    lv_stage.n = (size_t)D.C * blk;
    CU(cudaMallocAsync(&lv_stage.p, lv_stage.n * sizeof(double), 0));  // fully overwritten: no memset
    double* dst = hlv_stage;
    double* ddst = lv_stage.p;
    const int nchain = D.C, dev = F->device;
    const bvhar_chain_init* inits = d->init;
    std::atomic<int>* cerr = &lvol_copy_err;
    lvol_gather = std::thread([dst, ddst, blk, nchain, inits, dev, cerr] {
      const int nthr = std::max(1, std::min(8, nchain / 32));
      const int slice = 16;  // chains per transfer
      std::vector<std::thread> pool;
      for (int th = 0; th < nthr; ++th)
        pool.emplace_back([=] {
          if (cudaSetDevice(dev) != cudaSuccess) { cerr->store(1); return; }
          for (int c0 = th * slice; c0 < nchain; c0 += nthr * slice) {
            const int c1 = std::min(nchain, c0 + slice);
            for (int c = c0; c < c1; ++c)
              if (inits[c].lvol) std::memcpy(dst + (size_t)c * blk, inits[c].lvol, sizeof(double) * blk);
            if (cudaMemcpyAsync(ddst + (size_t)c0 * blk, dst + (size_t)c0 * blk, sizeof(double) * blk * (size_t)(c1 - c0), cudaMemcpyHostToDevice, 0) !=
                cudaSuccess)
              cerr->store(1);
          }
        });
      for (auto& t : pool) t.join();
    });
  }
  {
    // units write disjoint slices of the host staging vectors: a few host threads share the loop (22 ms on one thread for
    // the 1 024 chains of C3: a third of create's time)
    const int nthr = std::max(1, std::min(8, U / 64));
    if (nthr == 1) {
      for (int u = 0; u < U && !err_code; ++u) init_unit(u / D.C, u % D.C);
    } else {
      std::vector<std::thread> pool;
      for (int th = 0; th < nthr; ++th)
        pool.emplace_back([&, th] {  // contiguous unit ranges: neighbouring units share cache lines of the [a][unit, j] coefficient rows
          const int per = (U + nthr - 1) / nthr;
          for (int u = th * per; u < std::min(U, (th + 1) * per); ++u) {
            if (init_unit(u / D.C, u % D.C)) return;
          }
        });
      for (auto& t : pool) t.join();
    }
    if (err_code) return fail(err_code, err_msg);
  }
  lap("host init loops");
  if (lvol_gather.joinable()) lvol_gather.join();
  lap("lvol gather");
  if (lvol_copy_err.load()) return fail(BVHAR_ERR_CUDA, "upload of the log-volatility inits failed");
  CU(F->Acoef.upload(hA)); CU(F->alpha.upload(halpha)); CU(F->cvec.upload(hc)); CU(F->prec.upload(hprec));
  CU(F->contem.upload(hcontem)); CU(F->L.upload(hL)); CU(F->chol_prec.upload(hcp));
  CU(F->sparse_coef.alloc((size_t)U * dd * k)); CU(F->sparse_contem.alloc((size_t)U * q)); CU(F->znorm.alloc((size_t)2 * U * k));  // [U][k] column norms of Z, then [U][k] per-series sums of squared lvol increments
  CU(F->Z.alloc((size_t)tm_total));
  if (D.sv) {
    CU(F->H.alloc((size_t)tm_total)); CU(F->lvol_init.upload(hli)); CU(F->lvol_sig.upload(hls));
    {
      const int n0 = F->win_n[0];
      for (int w = 0; w < D.W; ++w) {
        if (lvol_explicit) {  // every window starts from the chains' lvol blocks (forecaster.h:867-872)
          CU(launch_scatter_lvol(lv_stage.p, F->H.p + F->win_off[w], F->win_n[w], D.M, D.ldM, n0, 0));
        } else {  // SvInits(init, num_design): lvol_init replicated over time (config.h:416-420) = rows of length 1
          CU(launch_scatter_lvol(F->lvol_init.p + (size_t)w * D.M, F->H.p + F->win_off[w], F->win_n[w], D.M, D.ldM, 0, 0));
        }
      }
      CU(cudaDeviceSynchronize());
      lv_stage.release();
    }
    lap("coef uploads + lvol upload / scatter");
    CU(F->Dinv.alloc((size_t)tm_total)); CU(F->YLD.alloc((size_t)tm_total)); CU(F->Zn.alloc((size_t)tm_total));
    {  // same decision as the coefficient-stage dispatch below: designs on the CTA-per-matrix path never touch S_i / C_i
      const bool v1 = getenv("BVHAR_COEF_V1") != nullptr, fbig = getenv("BVHAR_COEF_BIG") != nullptr;
      const bool v2 = coef_v2_fits(D, F->smem_limit) && !v1;
      F->direct_p = coef_big_fits(D, F->smem_limit) && !v1 && (D.d > 128 || !v2 || fbig) && getenv("BVHAR_NO_DIRECT_P") == nullptr;
    }
    if (F->direct_p) CU(F->Wt.alloc((size_t)tm_total));
    else { CU(F->S.alloc((size_t)U * k * D.ldS)); CU(F->Cm.alloc((size_t)U * k * D.ldX)); }
    CU(F->XX.alloc((size_t)D.T * D.ldXX));
  } else {
    CU(F->diag.upload(hdiag));
  }
  lap("big allocs");
  for (int i = 0; i < 4; ++i) {
    if (hhN[i].empty()) CU(F->hN[i].alloc((size_t)U * N));
    else CU(F->hN[i].upload(hhN[i]));
  }
  for (int i = 0; i < 3; ++i) CU(F->hG[i].upload(hhG[i]));
  for (int i = 0; i < 4; ++i) CU(F->hq[i].upload(hhq[i]));
  CU(F->hs.upload(hhs));

  lap("state uploads + allocs");
  DevPtrs& P = F->P;
  P.X = F->X.p; P.XT = F->XT.p; P.Y = F->Y.p; P.XX = F->XX.p; P.xnorm = F->xnorm.p;
  P.win_row0 = F->d_row0.p; P.win_n = F->d_n.p; P.win_off = F->d_off.p;
  P.grp_idx = F->grp_idx.p; P.grp_cnt = F->grp_cnt.p; P.own_flag = F->own_flag.p; P.prior_mean = F->prior_mean.p;
  P.sig_shp = F->sig_shp.p; P.sig_scl = F->sig_scl.p; P.sv_mean = F->sv_mean.p; P.sv_prec = F->sv_prec.p;
  P.ssvs_s1 = F->ssvs_s1.p; P.ssvs_s2 = F->ssvs_s2.p; P.seeds = F->seeds.p;
  P.H = F->H.p; P.Z = F->Z.p; P.Dinv = F->Dinv.p; P.YLD = F->YLD.p; P.Zn = F->Zn.p;
  P.S = F->S.p; P.Cm = F->Cm.p; P.Acoef = F->Acoef.p; P.alpha = F->alpha.p; P.cvec = F->cvec.p; P.prec = F->prec.p;
  P.contem = F->contem.p; P.L = F->L.p; P.chol_prec = F->chol_prec.p; P.sparse_coef = F->sparse_coef.p;
  P.sparse_contem = F->sparse_contem.p; P.lvol_init = F->lvol_init.p; P.lvol_sig = F->lvol_sig.p; P.diag = F->diag.p;
  P.znorm = F->znorm.p;
  P.hN0 = F->hN[0].p; P.hN1 = F->hN[1].p; P.hN2 = F->hN[2].p; P.hN3 = F->hN[3].p;
  P.hG0 = F->hG[0].p; P.hG1 = F->hG[1].p; P.hG2 = F->hG[2].p;
  P.hq0 = F->hq[0].p; P.hq1 = F->hq[1].p; P.hq2 = F->hq[2].p; P.hq3 = F->hq[3].p;
  P.hs = F->hs.p; P.iter = F->iter.p; P.rec_row = F->rec_row.p; P.flags = F->flags.p;

  CU(cudaStreamCreateWithFlags(&F->st, cudaStreamNonBlocking));
  CU(cudaEventCreate(&F->ev0)); CU(cudaEventCreate(&F->ev1));
  CU(cudaEventCreateWithFlags(&F->evs0, cudaEventDisableTiming));
  if (getenv("BVHAR_NO_FORK") == nullptr) {
    CU(cudaStreamCreateWithFlags(&F->s2, cudaStreamNonBlocking)); CU(cudaStreamCreateWithFlags(&F->s3, cudaStreamNonBlocking));
    for (auto& e : F->evf) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  CU(cudaDeviceSynchronize());  // uploads / memsets were ordered on stream 0; the setup kernels below run on F->st

  // ---- contraction descriptors ----
  std::vector<long long> g1b(D.W), g1c(D.W), g2b(D.W), g2c(D.W), g3a(D.W), g3b(D.W), g3y(D.W);
  for (int w = 0; w < D.W; ++w) {
    g1b[w] = (long long)F->win_row0[w] * D.ldXX; g1c[w] = (long long)w * D.M * D.ldS;
    g2b[w] = (long long)F->win_row0[w] * D.ldX;  g2c[w] = (long long)w * D.M * D.ldX;
    g3a[w] = F->win_row0[w]; g3b[w] = (long long)w * dd * D.ldM; g3y[w] = (long long)F->win_row0[w] * k;
    if (F->win_row0[w] & 1) F->g3_aligned = false;
  }
  CU(F->g1_b.upload(g1b)); CU(F->g1_c.upload(g1c)); CU(F->g2_b.upload(g2b)); CU(F->g2_c.upload(g2c));
  CU(F->g3_a.upload(g3a)); CU(F->g3_b.upload(g3b)); CU(F->g3_y.upload(g3y));
  CU(F->gKg.upload(F->win_n)); CU(F->gMg.upload(F->win_n));
  GemmArgs& g1 = F->g1;
  g1.A = P.Dinv; g1.B = P.XX; g1.C = P.S; g1.M = D.M; g1.N = D.Pp; g1.K = 0; g1.lda = D.ldM; g1.ldb = D.ldXX; g1.ldc = D.ldS;
  g1.groups = D.W; g1.a_off = F->d_off.p; g1.b_off = F->g1_b.p; g1.c_off = F->g1_c.p; g1.Kg = F->gKg.p; g1.Mg = nullptr;
  g1.epi = EPI_STORE; g1.aligned16 = 1;
  GemmArgs& g2 = F->g2;
  g2 = g1;
  g2.A = P.YLD; g2.B = P.X; g2.C = P.Cm; g2.N = dd; g2.ldb = D.ldX; g2.ldc = D.ldX; g2.b_off = F->g2_b.p; g2.c_off = F->g2_c.p;
  GemmArgs& g3 = F->g3;
  g3.A = P.XT; g3.B = P.Acoef; g3.C = P.Z; g3.M = D.nmax; g3.N = D.M; g3.K = dd; g3.lda = D.ldT; g3.ldb = D.ldM; g3.ldc = D.ldM;
  g3.groups = D.W; g3.a_off = F->g3_a.p; g3.b_off = F->g3_b.p; g3.c_off = F->d_off.p; g3.Kg = nullptr; g3.Mg = F->gMg.p;
  g3.epi = EPI_Y_MINUS; g3.Y = P.Y; g3.y_off = F->g3_y.p; g3.ldy = k; g3.ymod = k; g3.aligned16 = F->g3_aligned ? 1 : 0;

  // ---- scratch + setup kernels ----
  bool need_v = false, need_p = false;
  int ch_plan = 0;
  // coefficient stage: warp-per-matrix kernels for small designs, CTA-per-matrix blocked kernels for large ones,
  // the single-kernel path (everything inside one CTA per unit) as the last resort
  const bool force_v1 = getenv("BVHAR_COEF_V1") != nullptr, force_big = getenv("BVHAR_COEF_BIG") != nullptr;
  F->coef_v2 = coef_v2_fits(D, F->smem_limit) && !force_v1;
  F->coef_big = coef_big_fits(D, F->smem_limit) && !force_v1 && (D.d > 128 || !F->coef_v2 || force_big);
  if (F->coef_big) F->coef_v2 = false;
  if (F->coef_v2 || F->coef_big) CU(F->Pfac.alloc((size_t)U * k * D.ldS));
  if (F->direct_p) {
    if (!F->coef_big) return fail(BVHAR_ERR_CUDA, "internal: direct precision build without the large-design path");
    F->g1.A = F->Wt.p; F->g1.C = F->Pfac.p;
  }
  if (D.sv) {
    // Small batches (a few dozen chains: fewer output tiles than SMs) leave most of the GPU idle in the two time contractions
    // and each tile's K loop is a serial chain of k-tiles: cut K into slices that run as separate work items of the
    // persistent kernel and add the slices in a fixed order afterwards (deterministic: no atomics).
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, F->device);
    auto split = [&](GemmArgs& g, DevBuf<double>& ws, long long count) -> int {
      const long long tiles = (long long)((g.M + GBM - 1) / GBM) * ((g.N + GBN - 1) / GBN) * g.groups;
      const int S = g.aligned16 ? gemm_choose_ksplit(tiles, D.nmax, sms) : 1;
      if (S > 1) {
        CU(ws.alloc((size_t)S * (size_t)count));
        g.ksplit = S; g.ws = ws.p; g.ws_stride = count;
      }
      return 0;
    };
    F->g1_count = (long long)U * k * D.ldS;
    F->g2_count = (long long)U * k * D.ldX;
    if (int rc = split(F->g1, F->ws1, F->g1_count)) return rc;
    if (!F->direct_p) { if (int rc = split(F->g2, F->ws2, F->g2_count)) return rc; }
  }
  if (F->coef_big) CU(F->hscr.alloc(D.sv ? seq_time_scratch_doubles(D) : (size_t)U * k * dd));
  if (F->coef_v2 && coef_pipe_fits(D, F->smem_limit) && getenv("BVHAR_COEF_NOPIPE") == nullptr) CU(F->Qn.alloc((size_t)U * k * D.ldS));
  if (!F->coef_v2 && !F->coef_big && !coef_plan(D, F->smem_limit, &ch_plan, &need_v, &need_p))
    return fail(BVHAR_ERR_UNSUPPORTED, "dim_design too large for the shared-memory Cholesky of this build (packed d(d+1)/2 doubles must fit 227 KB)");
  if (need_v) CU(F->vscratch.alloc((size_t)U * 2 * k * dd));
  if (need_p) CU(F->pscratch.alloc((size_t)U * D.ldS));
  CU(F->gscratch.alloc((size_t)U * impact_gsize(D)));
  if (!D.sv) {  // Z'Z of every unit: one group per unit, A = B = the unit's k columns of the window's residual array
    const int kz = (k + 1) & ~1, gs = impact_gsize(D);
    std::vector<long long> za(U), zc(U);
    std::vector<int> zk(U);
    bool al = (D.ldM % 2) == 0;
    for (int u = 0; u < U; ++u) {
      const int w = u / D.C, c = u % D.C;
      za[u] = F->win_off[w] + (long long)c * k; zc[u] = (long long)u * gs; zk[u] = F->win_n[w];
      if (za[u] & 1) al = false;
    }
    CU(F->gz_a.upload(za)); CU(F->gz_c.upload(zc)); CU(F->gz_K.upload(zk));
    GemmArgs& gz = F->gz;
    gz.A = P.Z; gz.B = P.Z; gz.C = F->gscratch.p; gz.M = k; gz.N = k; gz.K = 0; gz.lda = D.ldM; gz.ldb = D.ldM; gz.ldc = kz;
    gz.groups = U; gz.a_off = F->gz_a.p; gz.b_off = F->gz_a.p; gz.c_off = F->gz_c.p; gz.Kg = F->gz_K.p; gz.Mg = nullptr;
    gz.epi = EPI_STORE; gz.aligned16 = al ? 1 : 0;
  }
  if (!impact_supported(D, F->smem_limit))
    return fail(BVHAR_ERR_UNSUPPORTED, "dim too large for the shared-memory contemporaneous draw of this build");
  // every allocation, memset and upload above was issued on stream 0, and F->st is a non-blocking stream: order
  // stream 0 before the first kernel on F->st (a staged pageable upload or a pending memset may not have landed yet)
  auto order_after_stream0 = [&]() -> int {
    CU(cudaEventRecord(F->evs0, 0));
    CU(cudaStreamWaitEvent(F->st, F->evs0, 0));
    return 0;
  };
  if (int rc = order_after_stream0()) return rc;
  if (D.sv) { CU(launch_build_xx(D, P, F->XX.p, F->st)); CU(launch_sv_prep(D, P, F->st)); }  // updateSv for the first sweep
  else {
    // X'X (packed) and X'Y per window through the same contraction kernel
    CU(F->XtX.alloc((size_t)D.W * D.ldS)); CU(F->XtY.alloc((size_t)D.W * dd * k));
    DevBuf<double>& full = F->XtXf;  // kept: the A operand of the residual contraction g4
    const int ldg = (dd + 1) & ~1;
    CU(full.alloc((size_t)D.W * dd * ldg));
    std::vector<long long> bo(D.W), co(D.W), yo(D.W), cy(D.W);
    for (int w = 0; w < D.W; ++w) { bo[w] = (long long)F->win_row0[w] * D.ldX; co[w] = (long long)w * dd * ldg; yo[w] = (long long)F->win_row0[w] * k; cy[w] = (long long)w * dd * k; }
    DevBuf<long long> dbo, dco, dyo, dcy;
    CU(dbo.upload(bo)); CU(dco.upload(co)); CU(dyo.upload(yo)); CU(dcy.upload(cy));
    if (int rc = order_after_stream0()) return rc;  // XtX / XtY / full / the offset tables live on stream 0
    GemmArgs g{};
    g.A = P.X; g.B = P.X; g.C = full.p; g.M = dd; g.N = dd; g.lda = D.ldX; g.ldb = D.ldX; g.ldc = ldg; g.groups = D.W;
    g.a_off = dbo.p; g.b_off = dbo.p; g.c_off = dco.p; g.Kg = F->gKg.p; g.epi = EPI_STORE; g.aligned16 = 1;
    CU(launch_dgemm_tn(g, F->st));
    GemmArgs gy = g;
    gy.B = P.Y; gy.b_off = dyo.p; gy.ldb = k; gy.N = k; gy.C = F->XtY.p; gy.ldc = k; gy.c_off = dcy.p; gy.aligned16 = (k % 2 == 0) ? 1 : 0;
    CU(launch_dgemm_tn(gy, F->st));
    CU(cudaStreamSynchronize(F->st));
    std::vector<double> hf((size_t)D.W * dd * ldg), hp((size_t)D.W * D.ldS, 0.0);
    CU(cudaMemcpy(hf.data(), full.p, hf.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (int w = 0; w < D.W; ++w)
      for (int a = 0; a < dd; ++a)
        for (int b = 0; b <= a; ++b) hp[(size_t)w * D.ldS + a * (a + 1) / 2 + b] = hf[((size_t)w * dd + a) * ldg + b];
    CU(cudaMemcpy(F->XtX.p, hp.data(), hp.size() * sizeof(double), cudaMemcpyHostToDevice));
    P.XtX = F->XtX.p; P.XtY = F->XtY.p;
    if (F->coef_big && getenv("BVHAR_LDLT_NO_RES") == nullptr) {
      // R = X'Y - X'X A for every chain of every window in one grouped contraction per sweep (X'X is symmetric: it is its own
      // K-major operand): the sequential kernel's prologue then needs no symmetric product at all
      CU(F->Rres.alloc((size_t)D.W * dd * D.ldM));
      std::vector<long long> ga(D.W), gy4(D.W);
      for (int w = 0; w < D.W; ++w) { ga[w] = (long long)w * dd * ldg; gy4[w] = (long long)w * dd * k; }
      CU(F->g4_a.upload(ga)); CU(F->g4_y.upload(gy4));
      if (int rc = order_after_stream0()) return rc;
      GemmArgs& g4 = F->g4;
      g4.A = full.p; g4.B = P.Acoef; g4.C = F->Rres.p; g4.M = dd; g4.N = D.M; g4.K = dd; g4.lda = ldg; g4.ldb = D.ldM; g4.ldc = D.ldM;
      g4.groups = D.W; g4.a_off = F->g4_a.p; g4.b_off = F->g3_b.p; g4.c_off = F->g3_b.p; g4.Kg = nullptr; g4.Mg = nullptr;
      g4.epi = EPI_Y_MINUS; g4.Y = P.XtY; g4.y_off = F->g4_y.p; g4.ldy = k; g4.ymod = k; g4.aligned16 = 1;
      F->ldlt_res = true;
    } else {
      F->XtXf.release();  // not needed
    }
  }
  // latent_innov = y - x * coef_mat (tri.h:57)
  CU(launch_dgemm_tn(F->g3, F->st));
  CU(cudaStreamSynchronize(F->st));

  lap("setup kernels");
  // ---- record buffers ----
  RecPtrs& R = F->R;
  std::memset(&R, 0, sizeof R);
  R.cap = F->num_iter - F->num_burn;
  const uint32_t m = F->rec_mask;
  auto add = [&](const char* nm, int cols, double** slot) -> int {
    auto buf = std::make_unique<DevBuf<double>>();
    CU(buf->alloc((size_t)U * std::max(R.cap, 1) * cols));
    *slot = buf->p;
    F->rec[nm] = std::move(buf);
    F->rec_order.push_back({nm, cols, slot});
    return 0;
  };
#define ADD(nm, cols, slot) do { int rc__ = add(nm, cols, slot); if (rc__) return rc__; } while (0)
  if (m & BVHAR_REC_COEF) {
    ADD("alpha_record", N, &R.alpha); ADD("a_record", q, &R.a);
    if (D.mean) ADD("c_record", k, &R.c);
    if (!D.sv) ADD("d_record", k, &R.d);
  }
  if (D.sv) {
    if (m & BVHAR_REC_LVOL) ADD("h_record", D.nmax * k, &R.h);
    else if (m & BVHAR_REC_LVOL_LAST) ADD("h_last_record", k, &R.h_last);
    if (m & BVHAR_REC_COEF) { ADD("h0_record", k, &R.h0); ADD("sigh_record", k, &R.sigh); }
  }
  if (m & BVHAR_REC_SPARSE) {
    ADD("alpha_sparse_record", N, &R.alpha_sparse); ADD("a_sparse_record", q, &R.a_sparse);
    if (D.mean) ADD("c_sparse_record", k, &R.c_sparse);
  }
  if (m & BVHAR_REC_PRIOR) {
    switch (D.prior) {
      case 2: ADD("gamma_record", N, &R.gamma); break;
      case 3: ADD("lambda_record", N, &R.lambda); ADD("eta_record", G, &R.eta); ADD("tau_record", 1, &R.tau); ADD("kappa_record", N, &R.kappa); break;
      case 5: ADD("lambda_record", N, &R.lambda); ADD("eta_record", G, &R.eta); ADD("tau_record", 1, &R.tau); break;
      case 6: ADD("lambda_record", N, &R.lambda); ADD("tau_record", 1, &R.tau); break;
      default: break;
    }
  }
#undef ADD
  CU(cudaDeviceSynchronize());  // allocations / uploads were ordered on stream 0, the sweeps run on F->st
  lap("record buffers");
  *out = F.release();
  return 0;
}

// ---- state views (names of the reference's members; column-major as Eigen) ----
struct View {
  enum Kind { UNIT_VEC, COEF_MAT, CHOL, TM, SQRT_SV, PREC, SCALAR } kind;
  double* base;   // device base (unit-major vector or [t][m] array)
  int len;        // per-unit length for UNIT_VEC
  int slot;       // SCALAR slot
};
bool find_view(bvhar_fit* F, const std::string& nm, View& v, int64_t& len) {
  const Dims& D = F->D;
  DevPtrs& P = F->P;
  auto uv = [&](double* b, int l) { v = {View::UNIT_VEC, b, l, 0}; len = l; return true; };
  auto sc = [&](int slot) { v = {View::SCALAR, P.hs, 1, slot}; len = 1; return true; };
  if (nm == "coef_mat") { v = {View::COEF_MAT, P.Acoef, 0, 0}; len = (int64_t)D.d * D.k; return true; }
  if (nm == "chol_lower") { v = {View::CHOL, P.L, 0, 0}; len = (int64_t)D.k * D.k; return true; }
  if (nm == "latent_innov") { v = {View::TM, P.Z, 0, 0}; len = -1; return true; }
  if (nm == "lvol_draw" && D.sv) { v = {View::TM, P.H, 0, 0}; len = -1; return true; }
  if (nm == "sqrt_sv" && D.sv) { v = {View::SQRT_SV, P.Dinv, 0, 0}; len = -1; return true; }
  if (nm == "prior_alpha_prec") { v = {View::PREC, P.prec, 0, 0}; len = (int64_t)D.k * D.d; return true; }
  if (nm == "contem_coef") return uv(P.contem, D.q);
  if (nm == "prior_chol_prec") return uv(P.chol_prec, D.q);
  if (nm == "sparse_coef") return uv(P.sparse_coef, D.d * D.k);
  if (nm == "sparse_contem") return uv(P.sparse_contem, D.q);
  if (nm == "alpha") return uv(P.alpha, D.N);
  if (nm == "diag_vec" && !D.sv) return uv(P.diag, D.k);
  if (nm == "lvol_init" && D.sv) return uv(P.lvol_init, D.k);
  if (nm == "lvol_sig" && D.sv) return uv(P.lvol_sig, D.k);
  switch (D.prior) {
    case 2:
      if (nm == "coef_dummy") return uv(P.hN0, D.N);
      if (nm == "coef_slab") return uv(P.hN1, D.N);
      if (nm == "coef_weight") return uv(P.hG0, D.G);
      if (nm == "contem_dummy") return uv(P.hq0, D.q);
      if (nm == "contem_slab") return uv(P.hq1, D.q);
      if (nm == "contem_weight") return uv(P.hq2, D.q);
      if (nm == "spike_scl") return sc(HS_SPIKE);
      if (nm == "contem_spike_scl") return sc(HS_CONTEM_SPIKE);
      break;
    case 3:
      if (nm == "local_lev") return uv(P.hN0, D.N);
      if (nm == "latent_local") return uv(P.hN1, D.N);
      if (nm == "shrink_fac") return uv(P.hN2, D.N);
      if (nm == "group_lev") return uv(P.hG0, D.G);
      if (nm == "latent_group") return uv(P.hG1, D.G);
      if (nm == "contem_local_lev") return uv(P.hq0, D.q);
      if (nm == "latent_contem_local") return uv(P.hq1, D.q);
      if (nm == "global_lev") return sc(HS_GLOBAL);
      if (nm == "latent_global") return sc(HS_LATENT_GLOBAL);
      if (nm == "contem_global_lev") return sc(HS_CONTEM_GLOBAL);
      if (nm == "latent_contem_global") return sc(HS_LATENT_CONTEM_GLOBAL);
      break;
    case 4:
      if (nm == "own_lambda") return sc(HS_OWN_LAMBDA);
      if (nm == "contem_lambda") return sc(HS_CONTEM_LAMBDA);
      break;
    case 5:
      if (nm == "local_lev") return uv(P.hN0, D.N);
      if (nm == "group_lev") return uv(P.hG0, D.G);
      if (nm == "local_shape") return uv(P.hG1, D.G);
      if (nm == "contem_fac") return uv(P.hq0, D.q);
      if (nm == "global_lev") return sc(HS_GLOBAL);
      if (nm == "contem_global_lev") return sc(HS_CONTEM_GLOBAL);
      if (nm == "contem_shape") return sc(HS_CONTEM_SHAPE);
      break;
    case 6:
      if (nm == "local_lev") return uv(P.hN0, D.N);
      if (nm == "latent_local") return uv(P.hN1, D.N);
      if (nm == "group_lev") return uv(P.hG0, D.G);
      if (nm == "contem_local_lev") return uv(P.hq0, D.q);
      if (nm == "latent_contem_local") return uv(P.hq1, D.q);
      if (nm == "global_lev") return sc(HS_GLOBAL);
      if (nm == "contem_global_lev") return sc(HS_CONTEM_GLOBAL);
      if (nm == "dir_concen") return sc(HS_DIR);
      if (nm == "contem_dir_concen") return sc(HS_CONTEM_DIR);
      break;
    case 7:
      if (nm == "local_lev") return uv(P.hN0, D.N);
      if (nm == "group_rate") return uv(P.hG0, D.G);
      if (nm == "contem_fac") return uv(P.hq0, D.q);
      if (nm == "contem_rate") return uv(P.hq1, D.q);
      if (nm == "coef_gamma_shape") return sc(HS_GSHAPE);
      if (nm == "coef_gamma_rate") return sc(HS_GRATE);
      if (nm == "contem_gamma_shape") return sc(HS_CGSHAPE);
      if (nm == "contem_gamma_rate") return sc(HS_CGRATE);
      break;
    default: break;
  }
  return false;
}

// moves one unit's view between the device layout and a column-major host vector
int view_io(bvhar_fit* F, const View& v, int w, int c, double* host, int64_t len, bool to_host) {
  const Dims& D = F->D;
  const int u = w * D.C + c, k = D.k, dd = D.d, n = F->win_n[w];
  const cudaMemcpyKind dir = to_host ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice;
  switch (v.kind) {
    case View::UNIT_VEC:
      if (len != v.len) return fail(BVHAR_ERR_BAD_ARG, "state length mismatch");
      if (v.len == 0) return 0;
      if (to_host) CU(cudaMemcpy(host, v.base + (size_t)u * v.len, sizeof(double) * v.len, dir));
      else CU(cudaMemcpy(v.base + (size_t)u * v.len, host, sizeof(double) * v.len, dir));
      return 0;
    case View::SCALAR:
      if (len != 1) return fail(BVHAR_ERR_BAD_ARG, "state length mismatch");
      if (to_host) CU(cudaMemcpy(host, v.base + (size_t)u * HS_SLOTS + v.slot, sizeof(double), dir));
      else CU(cudaMemcpy(v.base + (size_t)u * HS_SLOTS + v.slot, host, sizeof(double), dir));
      return 0;
    case View::COEF_MAT: {  // host d x k column-major <-> Acoef[w][a][c*k + j]
      if (len != (int64_t)dd * k) return fail(BVHAR_ERR_BAD_ARG, "state length mismatch");
      std::vector<double> tmp((size_t)dd * k);
      double* dev = v.base + (size_t)w * dd * D.ldM + (size_t)c * k;
      if (to_host) {
        CU(cudaMemcpy2D(tmp.data(), sizeof(double) * k, dev, sizeof(double) * D.ldM, sizeof(double) * k, dd, dir));
        for (int a = 0; a < dd; ++a) for (int j = 0; j < k; ++j) host[(size_t)j * dd + a] = tmp[(size_t)a * k + j];
      } else {
        for (int a = 0; a < dd; ++a) for (int j = 0; j < k; ++j) tmp[(size_t)a * k + j] = host[(size_t)j * dd + a];
        CU(cudaMemcpy2D(dev, sizeof(double) * D.ldM, tmp.data(), sizeof(double) * k, sizeof(double) * k, dd, dir));
        // keep coef_vec (alpha / constants) coherent with coef_mat
        std::vector<double> al(D.N), cv(k, 0.0);
        for (int j = 0; j < k; ++j) {
          for (int a = 0; a < D.nrow; ++a) al[(size_t)j * D.nrow + a] = host[(size_t)j * dd + a];
          if (D.mean) cv[j] = host[(size_t)j * dd + D.nrow];
        }
        CU(cudaMemcpy(F->P.alpha + (size_t)u * D.N, al.data(), sizeof(double) * D.N, dir));
        CU(cudaMemcpy(F->P.cvec + (size_t)u * k, cv.data(), sizeof(double) * k, dir));
      }
      return 0;
    }
    case View::CHOL: {  // host k x k column-major <-> L[u][i][m] row-major
      if (len != (int64_t)k * k) return fail(BVHAR_ERR_BAD_ARG, "state length mismatch");
      std::vector<double> tmp((size_t)k * k);
      if (to_host) {
        CU(cudaMemcpy(tmp.data(), v.base + (size_t)u * k * k, sizeof(double) * k * k, dir));
        for (int i = 0; i < k; ++i) for (int m2 = 0; m2 < k; ++m2) host[(size_t)m2 * k + i] = tmp[(size_t)i * k + m2];
      } else {
        for (int i = 0; i < k; ++i) for (int m2 = 0; m2 < k; ++m2) tmp[(size_t)i * k + m2] = host[(size_t)m2 * k + i];
        CU(cudaMemcpy(v.base + (size_t)u * k * k, tmp.data(), sizeof(double) * k * k, dir));
      }
      return 0;
    }
    case View::PREC: {  // host: head N alpha order + tail k constants <-> internal (j*d + a)
      if (len != (int64_t)k * dd) return fail(BVHAR_ERR_BAD_ARG, "state length mismatch");
      std::vector<double> tmp((size_t)k * dd);
      if (to_host) {
        CU(cudaMemcpy(tmp.data(), v.base + (size_t)u * k * dd, sizeof(double) * k * dd, dir));
        for (int j = 0; j < k; ++j) {
          for (int a = 0; a < D.nrow; ++a) host[(size_t)j * D.nrow + a] = tmp[(size_t)j * dd + a];
          if (D.mean) host[D.N + j] = tmp[(size_t)j * dd + D.nrow];
        }
      } else {
        for (int j = 0; j < k; ++j) {
          for (int a = 0; a < D.nrow; ++a) tmp[(size_t)j * dd + a] = host[(size_t)j * D.nrow + a];
          if (D.mean) tmp[(size_t)j * dd + D.nrow] = host[D.N + j];
        }
        CU(cudaMemcpy(v.base + (size_t)u * k * dd, tmp.data(), sizeof(double) * k * dd, dir));
      }
      return 0;
    }
    case View::TM:
    case View::SQRT_SV: {  // host n x k column-major <-> [t][c*k + i]
      if (len != (int64_t)n * k) return fail(BVHAR_ERR_BAD_ARG, "state length mismatch");
      std::vector<double> tmp((size_t)n * k);
      double* dev = v.base + F->win_off[w] + (size_t)c * k;
      if (to_host) {
        CU(cudaMemcpy2D(tmp.data(), sizeof(double) * k, dev, sizeof(double) * D.ldM, sizeof(double) * k, n, dir));
        for (int t = 0; t < n; ++t)
          for (int i = 0; i < k; ++i) {
            const double x = tmp[(size_t)t * k + i];
            host[(size_t)i * n + t] = v.kind == View::SQRT_SV ? 1.0 / std::sqrt(x) : x;
          }
      } else {
        for (int t = 0; t < n; ++t)
          for (int i = 0; i < k; ++i) {
            const double x = host[(size_t)i * n + t];
            tmp[(size_t)t * k + i] = v.kind == View::SQRT_SV ? 1.0 / (x * x) : x;
          }
        CU(cudaMemcpy2D(dev, sizeof(double) * D.ldM, tmp.data(), sizeof(double) * k, sizeof(double) * k, n, dir));
      }
      return 0;
    }
  }
  return fail(BVHAR_ERR_BAD_ARG, "unknown view");
}

}  // namespace

extern "C" {

const char* bvhar_cuda_last_error(void) { return g_err.c_str(); }
int bvhar_cuda_abi_version(void) { return BVHAR_CUDA_ABI_VERSION; }
int bvhar_cuda_device_count(int* count) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (count) *count = e == cudaSuccess ? n : 0;
  if (e != cudaSuccess) return fail(BVHAR_ERR_CUDA, cudaGetErrorString(e));
  return 0;
}
int bvhar_cuda_struct_sizes(int64_t* fit_desc, int64_t* chain_init, int64_t* forecast_desc, int64_t* prior_params) {
  if (fit_desc) *fit_desc = sizeof(bvhar_fit_desc);
  if (chain_init) *chain_init = sizeof(bvhar_chain_init);
  if (forecast_desc) *forecast_desc = sizeof(bvhar_forecast_desc);
  if (prior_params) *prior_params = sizeof(bvhar_prior_params);
  return 0;
}

int bvhar_cuda_outforecast_params_size(int64_t* bytes) {
  if (bytes) *bytes = sizeof(bvhar_outforecast_params);
  return 0;
}
int bvhar_cuda_host_alloc(int64_t bytes, void** out) {
  if (!out || bytes < 0) return fail(BVHAR_ERR_BAD_ARG, "bad host_alloc arguments");
  *out = nullptr;
  if (bytes == 0) return 0;
  cudaError_t e = cudaHostAlloc(out, (size_t)bytes, cudaHostAllocPortable);
  if (e != cudaSuccess) { *out = nullptr; return fail(BVHAR_ERR_CUDA, std::string("cudaHostAlloc: ") + cudaGetErrorString(e)); }
  return 0;
}
int bvhar_cuda_host_free(void* p) {
  if (!p) return 0;
  cudaError_t e = cudaFreeHost(p);
  if (e != cudaSuccess) return fail(BVHAR_ERR_CUDA, std::string("cudaFreeHost: ") + cudaGetErrorString(e));
  return 0;
}

int bvhar_cuda_fit_create(const bvhar_fit_desc* desc, bvhar_fit** out) { return build(desc, out); }

int bvhar_cuda_fit_step(bvhar_fit* fit, int n_sweeps, int record) {
  if (!fit || n_sweeps < 0) return fail(BVHAR_ERR_BAD_ARG, "bad fit handle or sweep count");
  return fit->step(n_sweeps, record != 0);
}
int bvhar_cuda_fit_sync(bvhar_fit* fit) {
  if (!fit) return fail(BVHAR_ERR_BAD_ARG, "null fit handle");
  return fit->sync();
}
int bvhar_cuda_fit_run(bvhar_fit* fit, bvhar_progress_fn progress, void* ctx) {
  if (!fit) return fail(BVHAR_ERR_BAD_ARG, "null fit handle");
  const int total = fit->num_iter, burn = fit->num_burn;
  int freq = total / 20;  // 5 %, triangular.h:1248-1251
  if (freq == 0) freq = 1;
  int done = (int)fit->sweeps_done;
  float ms_total = 0.f;
  while (done < total) {
    const bool rec = done >= burn;
    int chunk = std::min(freq - done % freq, (rec ? total : burn) - done);
    int rc = fit->step(chunk, rec);
    if (rc) return rc;
    if (rec && (rc = fit->stream_out())) return rc;
    done += chunk;
    if (progress) {
      // a non-PD flag is sticky and per unit: it is reported once, by the final sync, exactly as on the path without a
      // callback (the reference would carry NaNs in that chain and keep sampling) - only CUDA errors and the callback abort
      rc = fit->sync();
      ms_total += (float)fit->last_ms;
      if (rc && rc != BVHAR_ERR_NUMERICAL) return rc;
      if (progress(ctx, done, total)) return fail(BVHAR_ERR_INTERRUPTED, "interrupted by the progress callback");
    }
  }
  int rc = fit->sync();
  if (progress) fit->last_ms = ms_total;
  return rc;
}
int bvhar_cuda_fit_last_ms(bvhar_fit* fit, double* ms) {
  if (!fit || !ms) return fail(BVHAR_ERR_BAD_ARG, "null argument");
  *ms = fit->last_ms;
  return 0;
}
int bvhar_cuda_fit_set_profiling(bvhar_fit* fit, int on) {
  if (!fit) return fail(BVHAR_ERR_BAD_ARG, "null fit handle");
  fit->profiling = on != 0;
  if (on) { std::fill(fit->prof_ms.begin(), fit->prof_ms.end(), 0.0); std::fill(fit->prof_cnt.begin(), fit->prof_cnt.end(), 0LL); }
  return 0;
}
int bvhar_cuda_fit_profile(bvhar_fit* fit, char* buf, int64_t buflen) {
  if (!fit || !buf) return fail(BVHAR_ERR_BAD_ARG, "null argument");
  std::string s;
  char line[160];
  for (size_t i = 0; i < fit->prof_names.size(); ++i) {
    snprintf(line, sizeof line, "%s %lld %.6f\n", fit->prof_names[i].c_str(), fit->prof_cnt[i], fit->prof_ms[i]);
    s += line;
  }
  if ((int64_t)s.size() + 1 > buflen) return fail(BVHAR_ERR_BAD_ARG, "buffer too small");
  std::memcpy(buf, s.c_str(), s.size() + 1);
  return 0;
}
int bvhar_cuda_fit_launch_count(bvhar_fit* fit, int64_t* launches) {
  if (!fit || !launches) return fail(BVHAR_ERR_BAD_ARG, "null argument");
  *launches = fit->launches;
  return 0;
}

int bvhar_cuda_fit_record_dims(bvhar_fit* fit, const char* name, int64_t* rows, int64_t* cols) {
  if (!fit || !name) return fail(BVHAR_ERR_BAD_ARG, "null argument");
  for (auto& r : fit->rec_order)
    if (r.name == name) {
      if (rows) *rows = (fit->rec_filled + fit->thin - 1) / fit->thin;
      if (cols) *cols = r.cols;
      return 0;
    }
  return fail(BVHAR_ERR_BAD_ARG, std::string("no record named ") + name);
}
int bvhar_cuda_fit_record(bvhar_fit* fit, int window, int chain, const char* name, double* out, int64_t rows, int64_t cols) {
  if (!fit || !name || !out) return fail(BVHAR_ERR_BAD_ARG, "null argument");
  const Dims& D = fit->D;
  if (window < 0 || window >= D.W || chain < 0 || chain >= D.C) return fail(BVHAR_ERR_BAD_ARG, "window / chain out of range");
  CU(cudaSetDevice(fit->device));
  CU(cudaStreamSynchronize(fit->st));
  for (auto& r : fit->rec_order)
    if (r.name == name) {
      const int64_t kept = (fit->rec_filled + fit->thin - 1) / fit->thin;
      if (rows != kept || cols != r.cols) return fail(BVHAR_ERR_BAD_ARG, "record dims mismatch");
      if (kept == 0) return 0;
      const int u = window * D.C + chain;
      std::vector<double> tmp((size_t)fit->rec_filled * r.cols);
      CU(cudaMemcpy(tmp.data(), *r.slot + (size_t)u * fit->R.cap * r.cols, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
      // thin_record draw.h:1297-1310: rows 0, thin, 2 thin, ... of the kept draws
      for (int64_t j = 0; j < cols; ++j)
        for (int64_t i = 0; i < rows; ++i) out[j * rows + i] = tmp[(size_t)(i * fit->thin) * r.cols + j];
      return 0;
    }
  return fail(BVHAR_ERR_BAD_ARG, std::string("no record named ") + name);
}
// Every unit's record at once: out[unit][kept draw][column] (C order), i.e. the device layout, moved with
// one 2-D copy per unit (a single flat copy when nothing was thinned or left unfilled).
int bvhar_cuda_fit_bind_record_sink(bvhar_fit* fit, const char* name, double* host, int64_t units, int64_t rows, int64_t cols) {
  if (!fit || !name || !host) return fail(BVHAR_ERR_BAD_ARG, "null argument");
  const int64_t kept_final = ((int64_t)fit->R.cap + fit->thin - 1) / fit->thin;
  for (auto& r : fit->rec_order)
    if (r.name == name) {
      if (units != fit->D.U || rows != kept_final || cols != r.cols) return fail(BVHAR_ERR_BAD_ARG, "record sink dims mismatch");
      if (fit->rec_filled > 0) return fail(BVHAR_ERR_BAD_ARG, "record sinks must be bound before the first recorded sweep");
      for (auto& k : fit->sinks)
        if (k.rec == &r) { k.host = host; return 0; }
      fit->sinks.push_back({&r, host, rows, 0});
      return 0;
    }
  return fail(BVHAR_ERR_BAD_ARG, std::string("no record named ") + name);
}
int bvhar_cuda_fit_record_all(bvhar_fit* fit, const char* name, double* out, int64_t units, int64_t rows, int64_t cols) {
  if (!fit || !name || !out) return fail(BVHAR_ERR_BAD_ARG, "null argument");
  const Dims& D = fit->D;
  CU(cudaSetDevice(fit->device));
  CU(cudaStreamSynchronize(fit->st));
  for (auto& k : fit->sinks)  // already streamed to this very buffer while the sampler ran
    if (k.host == out && k.rec->name == name && k.copied == fit->rec_filled && rows == k.rows && units == D.U && cols == k.rec->cols) {
      if (fit->sc) CU(cudaStreamSynchronize(fit->sc));
      return 0;
    }
  for (auto& r : fit->rec_order)
    if (r.name == name) {
      const int64_t kept = (fit->rec_filled + fit->thin - 1) / fit->thin;
      if (units != D.U || rows != kept || cols != r.cols) return fail(BVHAR_ERR_BAD_ARG, "record dims mismatch");
      if (kept == 0) return 0;
      const size_t row_b = (size_t)r.cols * sizeof(double);
      if (fit->thin == 1 && fit->rec_filled == fit->R.cap) {
        CU(cudaMemcpy(out, *r.slot, (size_t)D.U * kept * row_b, cudaMemcpyDeviceToHost));
      } else {
        for (int u = 0; u < D.U; ++u)
          CU(cudaMemcpy2D(out + (size_t)u * kept * r.cols, row_b, *r.slot + (size_t)u * fit->R.cap * r.cols, row_b * fit->thin, row_b,
                          (size_t)kept, cudaMemcpyDeviceToHost));
      }
      return 0;
    }
  return fail(BVHAR_ERR_BAD_ARG, std::string("no record named ") + name);
}
// Posterior mean of every column of one record over all units and kept draws, NaNs skipped; computed on the device.
int bvhar_cuda_fit_record_mean(bvhar_fit* fit, const char* name, double* out, int64_t cols) {
  if (!fit || !name || !out) return fail(BVHAR_ERR_BAD_ARG, "null argument");
  API_LOCK;
  const Dims& D = fit->D;
  CU(cudaSetDevice(fit->device));
  CU(cudaStreamSynchronize(fit->st));
  for (auto& r : fit->rec_order)
    if (r.name == name) {
      const int64_t kept = (fit->rec_filled + fit->thin - 1) / fit->thin;
      if (cols != r.cols) return fail(BVHAR_ERR_BAD_ARG, "record dims mismatch");
      if (kept == 0) return fail(BVHAR_ERR_BAD_ARG, "no draws recorded yet");
      DevBuf<double> dm;
      CU(dm.alloc((size_t)r.cols));
      CU(launch_record_mean(*r.slot, D.U, fit->R.cap, r.cols, (int)kept, fit->thin, dm.p, 0));
      CU(cudaMemcpy(out, dm.p, sizeof(double) * r.cols, cudaMemcpyDeviceToHost));
      return 0;
    }
  return fail(BVHAR_ERR_BAD_ARG, std::string("no record named ") + name);
}
// Posterior summaries of every column of one record over all units (= chains of a one-window fit) and kept draws, computed
// on the device: out[col * 8 + {0..7}] = mean, sd, quantile(level / 2), median, quantile(1 - level / 2), split-R-hat, ESS, pip.
int bvhar_cuda_fit_record_summary(bvhar_fit* fit, const char* name, double level, double* out, int64_t cols) {
  if (!fit || !name || !out) return fail(BVHAR_ERR_BAD_ARG, "null argument");
  if (!(level > 0.0 && level < 1.0)) return fail(BVHAR_ERR_BAD_ARG, "level must be in (0, 1)");
  API_LOCK;
  const Dims& D = fit->D;
  CU(cudaSetDevice(fit->device));
  CU(cudaStreamSynchronize(fit->st));
  for (auto& r : fit->rec_order)
    if (r.name == name) {
      const int64_t kept = (fit->rec_filled + fit->thin - 1) / fit->thin;
      if (cols != r.cols) return fail(BVHAR_ERR_BAD_ARG, "record dims mismatch");
      if (kept == 0) return fail(BVHAR_ERR_BAD_ARG, "no draws recorded yet");
      DevBuf<double> dm;
      CU(dm.alloc((size_t)r.cols * 8));
      CU(launch_record_summary(*r.slot, D.U, fit->R.cap, r.cols, (int)kept, fit->thin, level, dm.p, 0));
      CU(cudaMemcpy(out, dm.p, sizeof(double) * r.cols * 8, cudaMemcpyDeviceToHost));
      return 0;
    }
  return fail(BVHAR_ERR_BAD_ARG, std::string("no record named ") + name);
}
int bvhar_cuda_fit_record_names(bvhar_fit* fit, char* buf, int64_t buflen) {
  if (!fit || !buf) return fail(BVHAR_ERR_BAD_ARG, "null argument");
  std::string s;
  for (auto& r : fit->rec_order) { s += r.name; s += '\n'; }
  if ((int64_t)s.size() + 1 > buflen) return fail(BVHAR_ERR_BAD_ARG, "buffer too small");
  std::memcpy(buf, s.c_str(), s.size() + 1);
  return 0;
}

int bvhar_cuda_fit_state_len(bvhar_fit* fit, const char* name, int64_t* len) {
  if (!fit || !name || !len) return fail(BVHAR_ERR_BAD_ARG, "null argument");
  View v;
  int64_t l = 0;
  if (!find_view(fit, name, v, l)) return fail(BVHAR_ERR_BAD_ARG, std::string("no state named ") + name);
  *len = l < 0 ? (int64_t)fit->win_n[0] * fit->D.k : l;
  return 0;
}
static int state_io(bvhar_fit* fit, int window, int chain, const char* name, double* buf, int64_t len, bool to_host) {
  if (!fit || !name || !buf) return fail(BVHAR_ERR_BAD_ARG, "null argument");
  const Dims& D = fit->D;
  if (window < 0 || window >= D.W || chain < 0 || chain >= D.C) return fail(BVHAR_ERR_BAD_ARG, "window / chain out of range");
  CU(cudaSetDevice(fit->device));
  CU(cudaStreamSynchronize(fit->st));
  View v;
  int64_t l = 0;
  if (!find_view(fit, name, v, l)) return fail(BVHAR_ERR_BAD_ARG, std::string("no state named ") + name);
  return view_io(fit, v, window, chain, buf, len, to_host);
}
int bvhar_cuda_fit_get_state(bvhar_fit* fit, int window, int chain, const char* name, double* out, int64_t len) {
  return state_io(fit, window, chain, name, out, len, true);
}
int bvhar_cuda_fit_set_state(bvhar_fit* fit, int window, int chain, const char* name, const double* in, int64_t len) {
  int rc = state_io(fit, window, chain, name, const_cast<double*>(in), len, false);
  if (rc == 0 && fit->D.sv && (!strcmp(name, "lvol_draw") || !strcmp(name, "chol_lower"))) {  // keep the cached updateSv coherent
    CU(launch_sv_prep(fit->D, fit->P, fit->st));
    CU(cudaStreamSynchronize(fit->st));
  }
  return rc;
}
int bvhar_cuda_fit_run_stage(bvhar_fit* fit, int stage, int iter) {
  if (!fit) return fail(BVHAR_ERR_BAD_ARG, "null fit handle");
  CU(cudaSetDevice(fit->device));
  const uint32_t it = (uint32_t)iter;
  CU(cudaMemcpyAsync(fit->iter.p, &it, sizeof it, cudaMemcpyHostToDevice, fit->st));
  CU(cudaStreamSynchronize(fit->st));
  const Dims& D = fit->D;
  switch (stage) {
    case 1: CU(launch_prior(D, fit->P, fit->PC, 0, fit->st)); break;
    case 2: break;  // alpha_penalty is a constant of the group structure here
    case 3: if (D.sv) CU(launch_sv_prep(D, fit->P, fit->st)); break;
    case 4:
      if (D.sv && fit->direct_p) { CU(launch_sv_wmix(D, fit->P, fit->Wt.p, fit->st)); CU(launch_dgemm_tn(fit->g1, fit->st)); }
      else if (D.sv) { CU(launch_dgemm_tn(fit->g1, fit->st)); CU(launch_dgemm_tn(fit->g2, fit->st)); }
      if (D.sv) { CU(launch_splitk_reduce(fit->g1, fit->g1_count, fit->st)); if (!fit->direct_p) CU(launch_splitk_reduce(fit->g2, fit->g2_count, fit->st)); }
      if (fit->coef_big) {
        if (D.sv && !fit->direct_p) CU(launch_coef_v2(0, D, fit->P, fit->Pfac.p, fit->Qn.p, fit->smem_limit, fit->st));
        CU(launch_coef_big(1, D, fit->P, fit->Pfac.p, fit->hscr.p, fit->smem_limit, fit->st));
        if (fit->ldlt_res) CU(launch_dgemm_tn(fit->g4, fit->st));
        CU(launch_coef_big(2, D, fit->P, fit->Pfac.p, fit->hscr.p, fit->smem_limit, fit->st, fit->ldlt_res ? fit->Rres.p : nullptr));
      } else if (fit->coef_v2) {
        CU(launch_coef_v2(0, D, fit->P, fit->Pfac.p, fit->Qn.p, fit->smem_limit, fit->st));
        CU(launch_coef_v2(1, D, fit->P, fit->Pfac.p, fit->Qn.p, fit->smem_limit, fit->st));
        CU(launch_coef_v2(2, D, fit->P, fit->Pfac.p, fit->Qn.p, fit->smem_limit, fit->st));
      } else {
        CU(launch_coef(D, fit->P, fit->vscratch.p, fit->pscratch.p, fit->smem_limit, fit->st));
      }
      break;
    case 5: CU(launch_prior(D, fit->P, fit->PC, 1, fit->st)); break;
    case 6: CU(launch_dgemm_tn(fit->g3, fit->st)); break;
    case 7:
      if (!D.sv && !ldlt_own_gram(D)) CU(launch_dgemm_tn(fit->gz, fit->st));
      if (D.k > 1) CU(launch_impact(D, fit->P, fit->gscratch.p, fit->st));
      else if (!D.sv) CU(launch_ldlt_gram_small(D, fit->P, fit->gscratch.p, fit->st));
      break;
    case 8: break;  // chol_lower is rebuilt inside the contemporaneous draw
    case 9:
      if (D.sv) { CU(launch_sv_state(3, D, fit->P, fit->st)); CU(launch_sv_state(0, D, fit->P, fit->st)); CU(launch_sv_state(1, D, fit->P, fit->st)); CU(launch_sv_state(2, D, fit->P, fit->st)); }
      else {  // the state draw reads Z'Z: recompute it for the residuals in place now
        if (ldlt_own_gram(D)) CU(launch_ldlt_gram_small(D, fit->P, fit->gscratch.p, fit->st));
        else CU(launch_dgemm_tn(fit->gz, fit->st));
        CU(launch_ldlt_state(D, fit->P, fit->gscratch.p, fit->st));
      }
      break;
    default: return fail(BVHAR_ERR_BAD_ARG, "stage must be 1..9");
  }
  if (D.sv && (stage == 7 || stage == 9)) CU(launch_sv_prep(D, fit->P, fit->st));  // keep the cached updateSv coherent with L / h
  return fit->sync();
}

void bvhar_cuda_fit_destroy(bvhar_fit* fit) {
  API_LOCK;
  delete fit;
}

// Order statistics behind RegRecords::computeActivity (config.h:747-758): boost::accumulators::tail_quantile evaluates
// n = ceil(count * prob) for the left tail and n = ceil(count * (1. - prob)) for the right tail, in double precision, with
// prob = level / 2 and 1 - level / 2 (so the two counts can differ by one, e.g. count = 1000, level = .05: 25 and 26).
static void activity_order_stats(int S, double level, int* n_lo, int* n_hi) {
  const double p_lo = level / 2, p_hi = 1 - level / 2;
  *n_lo = (int)std::ceil((double)S * p_lo);
  *n_hi = (int)std::ceil((double)S * (1. - p_hi));
}
static int activity_graph(const ForecastArgs& A, double level, DevBuf<double>& act, DevBuf<double>& scratch, cudaStream_t st) {
  int n_lo = 0, n_hi = 0;
  activity_order_stats(A.S, level, &n_lo, &n_hi);
  const int N = A.nrow * A.k, kc = A.mean ? A.k : 0;
  CU(act.alloc((size_t)A.n_units * (N + kc)));
  CU(scratch.alloc(activity_scratch_doubles(A.n_units, N + kc, A.S)));
  CU(launch_activity(A.coef, A.cst, A.n_units, N, kc, A.S, n_lo, n_hi, act.p, scratch.p, st));
  return 0;
}

// ---- density forecast -------------------------------------------------------------------
int bvhar_cuda_forecast_density(const bvhar_forecast_desc* d, double* out_density, double* out_lpl) {
  if (!d || !out_density) return fail(BVHAR_ERR_BAD_ARG, "null argument");
  API_LOCK;
  if (d->abi_version != BVHAR_CUDA_ABI_VERSION) return fail(BVHAR_ERR_BAD_ARG, "descriptor ABI version mismatch");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(BVHAR_ERR_CUDA, "no CUDA device available: no CPU fallback");
  if (d->n_draws < 1) return fail(BVHAR_ERR_BAD_ARG, "No stable MCMC draws");  // forecaster.h:286, 322
  if (d->step < 1 || d->step > 64) return fail(BVHAR_ERR_UNSUPPORTED, "forecast step must be in 1..64");
  CU(cudaSetDevice(d->device));
  const int k = d->dim, S = d->n_draws, U = d->n_units, q = k * (k - 1) / 2;
  const int ncoef = d->nrow_coef * k + (d->include_mean ? k : 0);
  ForecastArgs A{};
  A.n_units = U; A.k = k; A.nrow = d->nrow_coef; A.mean = d->include_mean; A.sv_model = d->cov_type == BVHAR_COV_SV;
  A.step = d->step; A.p = d->var_lag; A.is_vhar = d->is_vhar; A.sv = d->sv; A.get_lpl = d->get_lpl && d->valid_vec && out_lpl; A.S = S;
  A.har_r = 3 * k + (d->include_mean ? 1 : 0); A.har_c = d->var_lag * k + (d->include_mean ? 1 : 0);
  DevBuf<double> har, coef, contem, diag, sigh, last, valid, out, lpl;
  DevBuf<uint32_t> seed;
  auto up = [&](DevBuf<double>& b, const double* src, size_t n) -> cudaError_t {
    cudaError_t e = b.alloc(n);
    if (e != cudaSuccess || n == 0) return e;
    return cudaMemcpy(b.p, src, n * sizeof(double), cudaMemcpyHostToDevice);
  };
  if (d->is_vhar) { if (!d->har_trans) return fail(BVHAR_ERR_BAD_ARG, "har_trans missing"); CU(up(har, d->har_trans, (size_t)A.har_r * A.har_c)); }
  CU(up(coef, d->coef_record, (size_t)U * S * ncoef));
  CU(up(contem, d->contem_record, (size_t)U * S * q));
  CU(up(diag, d->diag_record, (size_t)U * S * k));
  if (A.sv_model) CU(up(sigh, d->sigh_record, (size_t)U * S * k));
  CU(up(last, d->last_obs, (size_t)U * d->var_lag * k));
  if (A.get_lpl) CU(up(valid, d->valid_vec, k));
  CU(seed.upload(std::vector<uint32_t>(d->seed, d->seed + U)));
  CU(out.alloc((size_t)U * S * k * d->step));
  if (A.get_lpl) CU(lpl.alloc((size_t)U * d->step));
  A.unit_base = (unsigned)d->unit_id_base;
  A.har = har.p;
  A.coef = {coef.p, (long long)S * ncoef, 1, S};  // host records: column-major n_draws x cols per unit
  A.cst = {coef.p + (size_t)d->nrow_coef * k * S, (long long)S * ncoef, 1, S};
  A.contem = {contem.p, (long long)S * q, 1, S};
  A.diag = {diag.p, (long long)S * k, 1, S};
  A.sigh = {sigh.p, (long long)S * k, 1, S};
  A.last_obs = last.p;
  A.valid = valid.p; A.seed = seed.p; A.out = out.p; A.lpl = lpl.p;
  DevBuf<double> act, act_scratch;
  A.ncoef = ncoef;
  if (d->level > 0) {  // Select forecasters (the caller's `sparse` draws are refused upstream: forecaster.h:445-447)
    if (int rc = activity_graph(A, d->level, act, act_scratch, 0)) return rc;
    A.act = act.p;
  }
  CU(launch_forecast(A, 0));
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(out_density, out.p, out.n * sizeof(double), cudaMemcpyDeviceToHost));
  if (A.get_lpl) CU(cudaMemcpy(out_lpl, lpl.p, lpl.n * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

// ---- out-of-sample refits + forecasts in one call (McmcOutforecastRun::returnForecast, forecaster.h:616-650,
// 755-806): every (window, chain) of the descriptor's window table is refitted, then forecast straight from the
// device-resident records (no host round trip of the draws).
// one batch of windows that fits in device memory (bvhar_cuda_outforecast_run below splits the window table)
static int outforecast_batch(const bvhar_fit_desc* fd, const bvhar_outforecast_params* fp, double* out_forecast, double* out_lpl) {
  bvhar_fit_desc d2 = *fd;
  d2.record_mask = BVHAR_REC_COEF | BVHAR_REC_LVOL_LAST | (fp->sparse ? BVHAR_REC_SPARSE : 0u);  // all the forecaster reads (config.h:998-1001)
  bvhar_fit* fit = nullptr;
  int rc = build(&d2, &fit);
  if (rc) return rc;
  std::unique_ptr<bvhar_fit> guard(fit);
  rc = bvhar_cuda_fit_run(fit, nullptr, nullptr);
  if (rc && rc != BVHAR_ERR_NUMERICAL) return rc;  // a non-PD flag does not abort (the reference would carry NaNs)
  API_LOCK;  // the forecast part below works on stream 0 with device-wide syncs
  const Dims& D = fit->D;
  const int k = D.k, U = D.U, q = D.q, p = fp->var_lag, step = fp->step;
  const int S = (fit->rec_filled + fit->thin - 1) / fit->thin;
  if (S < 1) return fail(BVHAR_ERR_BAD_ARG, "No stable MCMC draws");  // forecaster.h:286, 322
  const int dd_fc = p * k + (D.mean ? 1 : 0);
  if (fp->is_vhar ? (D.nrow != 3 * k) : (D.nrow != p * k)) return fail(BVHAR_ERR_BAD_ARG, "var_lag does not match the coefficient matrix");
  (void)dd_fc;
  ForecastArgs A{};
  A.n_units = U; A.k = k; A.nrow = D.nrow; A.mean = D.mean; A.sv_model = D.sv; A.step = step; A.p = p; A.is_vhar = fp->is_vhar;
  A.sv = fp->sv; A.get_lpl = fp->get_lpl && fp->valid_vec && out_lpl; A.S = S; A.unit_base = (unsigned)D.ubase;
  A.har_r = 3 * k + (D.mean ? 1 : 0); A.har_c = p * k + (D.mean ? 1 : 0);
  const RecPtrs& R = fit->R;
  const long long cap = R.cap, th = fit->thin;
  const double* coefp = fp->sparse ? R.alpha_sparse : R.alpha;
  const double* cstp = fp->sparse ? R.c_sparse : R.c;
  const double* ap = fp->sparse ? R.a_sparse : R.a;
  A.coef = {coefp, cap * D.N, th * D.N, 1};
  A.cst = {cstp, cap * k, th * k, 1};
  A.contem = {ap, cap * q, th * q, 1};
  A.diag = {D.sv ? R.h_last : R.d, cap * k, th * k, 1};
  A.sigh = {R.sigh, cap * k, th * k, 1};
  // last var_lag observations of every window: response rows of window w sit at y_full rows [p + row0, p + row0 + n)
  std::vector<double> hlast((size_t)U * p * k);
  for (int w = 0; w < D.W; ++w) {
    const int64_t end = (int64_t)p + fit->win_row0[w] + fit->win_n[w];
    if (end > fp->y_rows) return fail(BVHAR_ERR_BAD_ARG, "y_full is shorter than the window table implies");
    for (int c = 0; c < D.C; ++c)
      for (int jj = 0; jj < k; ++jj)
        for (int l = 0; l < p; ++l) hlast[((size_t)(w * D.C + c) * k + jj) * p + l] = fp->y_full[(size_t)jj * fp->y_rows + (end - p + l)];
  }
  std::vector<uint32_t> hseed((size_t)U);
  for (int u = 0; u < U; ++u) hseed[u] = fp->seed_forecast[u % D.C];  // the same seed for every window (forecaster.h:1028)
  DevBuf<double> har, last, valid, out, lpl;
  DevBuf<uint32_t> seed;
  if (fp->is_vhar) CU(har.upload(std::vector<double>(fp->har_trans, fp->har_trans + (size_t)A.har_r * A.har_c)));
  CU(last.upload(hlast)); CU(seed.upload(hseed));
  if (A.get_lpl) CU(valid.upload(std::vector<double>(fp->valid_vec, fp->valid_vec + k)));
  CU(out.alloc((size_t)U * S * k * step));
  if (A.get_lpl) CU(lpl.alloc((size_t)U * step));
  A.har = har.p; A.last_obs = last.p; A.valid = valid.p; A.seed = seed.p; A.out = out.p; A.lpl = lpl.p;
  DevBuf<double> act, act_scratch;
  A.ncoef = D.N + (D.mean ? k : 0);
  if (fp->level > 0) {  // McmcVarSelectForecaster / McmcVharSelectForecaster per refitted window (forecaster.h:1024-1030)
    CU(cudaStreamSynchronize(fit->st));
    if (int rc2 = activity_graph(A, fp->level, act, act_scratch, 0)) return rc2;
    A.act = act.p;
  }
  CU(launch_forecast(A, 0));
  CU(cudaDeviceSynchronize());
  std::vector<double> hout(out.n);
  CU(cudaMemcpy(hout.data(), out.p, out.n * sizeof(double), cudaMemcpyDeviceToHost));
  for (size_t e = 0; e < (size_t)U * S * k; ++e) out_forecast[e] = hout[e * step + (step - 1)];  // .bottomRows(1), forecaster.h:803
  if (A.get_lpl) {
    std::vector<double> hl(lpl.n);
    CU(cudaMemcpy(hl.data(), lpl.p, lpl.n * sizeof(double), cudaMemcpyDeviceToHost));
    for (int u = 0; u < U; ++u) {
      double s = 0.0;
      for (int h = 0; h < step; ++h) s += hl[(size_t)u * step + h];
      out_lpl[u] = s / step;
    }
  }
  return 0;
}

// Device bytes one (window, chain) unit of this descriptor needs, dominated by the packed Gram matrices S_i, the
// factors and the [t][m] state (an estimate with head-room; an out-of-memory batch is halved and retried anyway).
static size_t outforecast_unit_bytes(const bvhar_fit_desc* fd, int step) {
  const size_t k = fd->dim, d = fd->dim_design, q = k * (k - 1) / 2, N = k * d;
  size_t nmax = 1;
  for (int w = 0; w < fd->n_windows; ++w) nmax = std::max(nmax, (size_t)fd->win_nrow[w]);
  const size_t pp = d * (d + 1) / 2 + 1;
  const size_t kept = (size_t)std::max(1, fd->num_iter - fd->num_burn);
  size_t dbl = 3 * k * pp + k * (d + 1) + 2 * k * std::max(nmax + 3, d) + 5 * nmax * (k + 1) + k * k * k / 2 + 4 * k * k;
  dbl += kept * 2 * (N + 3 * k + q) + kept * k * (size_t)step + 40 * (N + q + k);
  return dbl * sizeof(double);
}

int bvhar_cuda_outforecast_run(const bvhar_fit_desc* fd, const bvhar_outforecast_params* fp, double* out_forecast, double* out_lpl) {
  if (!fd || !fp || !out_forecast) return fail(BVHAR_ERR_BAD_ARG, "null argument");
  if (fp->abi_version != BVHAR_CUDA_ABI_VERSION) return fail(BVHAR_ERR_BAD_ARG, "descriptor ABI version mismatch");
  if (fp->step < 1 || fp->step > 64) return fail(BVHAR_ERR_UNSUPPORTED, "forecast step must be in 1..64");
  if (!fp->y_full || !fp->seed_forecast) return fail(BVHAR_ERR_BAD_ARG, "y_full / seed_forecast missing");
  if (fp->is_vhar && !fp->har_trans) return fail(BVHAR_ERR_BAD_ARG, "har_trans missing");
  if (fp->sparse && fp->level > 0) return fail(BVHAR_ERR_BAD_ARG, "If 'level > 0', 'spare' should be false.");  // forecaster.h:445-447 (sic)
  if (fd->n_windows < 1 || fd->n_chains < 1 || !fd->win_row0 || !fd->win_nrow || !fd->seed_chain) return fail(BVHAR_ERR_BAD_ARG, "empty window table");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(BVHAR_ERR_CUDA, "no CUDA device available: no CPU fallback");
  CU(cudaSetDevice(fd->device));
  // Windows are independent (forecaster.h:628-632): the table is processed in batches sized to the free device memory
  // (2 000 windows x 4 chains of C4 would need ~130 GB of Gram matrices and factors at once).  Unit u of a batch keeps
  // its global Philox key (unit_id_base + first unit of the batch), so the result does not depend on the batch size.
  const int W = fd->n_windows, C = fd->n_chains, k = fd->dim;
  const int thin = fd->thin < 1 ? 1 : fd->thin;
  const long long S = ((long long)std::max(0, fd->num_iter - fd->num_burn) + thin - 1) / thin;
  size_t free_b = 0, total_b = 0;
  CU(cudaMemGetInfo(&free_b, &total_b));
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, fd->device) == cudaSuccess) {  // memory parked in the stream-ordered pool is reusable
    unsigned long long reserved = 0, used = 0;
    if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && reserved > used)
      free_b += (size_t)(reserved - used);
  }
  const size_t d_ = fd->dim_design;
  const size_t shared_b = (size_t)fd->n_rows_total * (d_ * (d_ + 1) / 2 + 2 * d_ + 2 * k + 8) * sizeof(double);  // X, X', XX, Y
  const size_t unit_b = outforecast_unit_bytes(fd, fp->step);
  const size_t budget = free_b > shared_b ? (size_t)(0.85 * (double)(free_b - shared_b)) : 0;
  long long batch = std::max<long long>(1, (long long)(budget / ((size_t)C * unit_b)));
  if (const char* env = getenv("BVHAR_MAX_WINDOWS_PER_BATCH")) batch = std::max(1, atoi(env));
  batch = std::min<long long>(batch, W);
  for (int w0 = 0; w0 < W;) {
    const int nw = (int)std::min<long long>(batch, W - w0);
    bvhar_fit_desc sub = *fd;
    sub.n_windows = nw;
    sub.win_row0 = fd->win_row0 + w0;
    sub.win_nrow = fd->win_nrow + w0;
    sub.seed_chain = fd->seed_chain + (size_t)w0 * C;
    sub.unit_id_base = fd->unit_id_base + w0 * C;
    const int rc = outforecast_batch(&sub, fp, out_forecast + (size_t)w0 * C * S * k, out_lpl ? out_lpl + (size_t)w0 * C : nullptr);
    if (rc == BVHAR_ERR_CUDA && nw > 1 && g_err.find("out of memory") != std::string::npos) {
      cudaGetLastError();
      cudaDeviceSynchronize();
      if (cudaDeviceGetDefaultMemPool(&pool, fd->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
      batch = (nw + 1) / 2;
      continue;  // retry the same windows in smaller batches
    }
    if (rc) return rc;
    w0 += nw;
  }
  return 0;
}

// ---- stability filter of the forecasters (subsetStable config.h:884-914, 1003-1034; is_stable draw.h:1265-1295) ----
int bvhar_cuda_is_stable(int device, int64_t n_draws, int32_t dim, int32_t nrow, int32_t lag, const double* coef, int64_t ld,
                         const double* har_trans, int64_t ldh, double threshold, int32_t* out) {
  if (!coef || !out) return fail(BVHAR_ERR_BAD_ARG, "null argument");
  if (n_draws < 0 || dim < 1 || lag < 1 || nrow < 1 || ld < (int64_t)nrow * dim || !(threshold > 0.0)) return fail(BVHAR_ERR_BAD_ARG, "bad dimensions");
  if (har_trans ? (nrow != 3 * dim || ldh < 3 * dim) : (nrow != dim * lag)) return fail(BVHAR_ERR_BAD_ARG, "coefficient rows do not match the lag structure");
  if (n_draws == 0) return 0;
  API_LOCK;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(BVHAR_ERR_CUDA, "no CUDA device available: no CPU fallback");
  CU(cudaSetDevice(device));
  pool_setup(device);
  const int m = dim * lag, ldm = (m + 1) & ~1;
  const size_t mat = (size_t)m * ldm;
  const int64_t batch = std::max<int64_t>(1, std::min<int64_t>({n_draws, (int64_t)65535, (int64_t)((3ull << 30) / (4 * mat * sizeof(double)))}));
  DevBuf<double> dcoef, dhar, R[2], Rt[2], logacc;
  DevBuf<int> state, kg, undecided;
  DevBuf<long long> off;
  CU(dcoef.alloc((size_t)batch * ld)); CU(logacc.alloc((size_t)batch)); CU(state.alloc((size_t)batch)); CU(kg.alloc((size_t)batch));
  CU(undecided.alloc(1));
  for (auto& b : R) CU(b.alloc((size_t)batch * mat));
  for (auto& b : Rt) CU(b.alloc((size_t)batch * mat));
  if (har_trans) {
    std::vector<double> h((size_t)m * 3 * dim);
    for (int c = 0; c < m; ++c)
      for (int r = 0; r < 3 * dim; ++r) h[(size_t)c * 3 * dim + r] = har_trans[(size_t)c * ldh + r];
    CU(dhar.upload(h));
  }
  {
    std::vector<long long> ho((size_t)batch);
    for (int64_t g = 0; g < batch; ++g) ho[g] = (long long)(g * mat);
    CU(off.upload(ho));
  }
  std::vector<int> hk((size_t)batch, m), hs((size_t)batch);
  for (int64_t i0 = 0; i0 < n_draws; i0 += batch) {
    const int64_t nb = std::min(batch, n_draws - i0);
    CU(cudaMemcpy(dcoef.p, coef + (size_t)i0 * ld, (size_t)nb * ld * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(kg.p, hk.data(), (size_t)nb * sizeof(int), cudaMemcpyHostToDevice));
    CU(launch_stab_build(dcoef.p, ld, nb, dim, nrow, lag, dhar.p, 3 * dim, 1.0 / threshold, R[0].p, Rt[0].p, ldm, logacc.p, state.p, 0));
    int cur = 0;
    for (int it = 0; it < STAB_MAX_SQUARINGS; ++it) {
      GemmArgs g{};
      g.M = m; g.N = m; g.K = m; g.lda = ldm; g.ldb = ldm; g.ldc = ldm; g.groups = (int)nb;
      g.a_off = off.p; g.b_off = off.p; g.c_off = off.p; g.Kg = kg.p; g.epi = EPI_STORE; g.aligned16 = 1;
      g.A = Rt[cur].p; g.B = R[cur].p; g.C = R[cur ^ 1].p;   // (B^2)[r][c]   = sum_k B[r][k] B[k][c]
      CU(launch_dgemm_tn(g, 0));
      g.A = R[cur].p; g.B = Rt[cur].p; g.C = Rt[cur ^ 1].p;  // (B^2)'[r][c]  = sum_k B[k][r] B[c][k]
      CU(launch_dgemm_tn(g, 0));
      cur ^= 1;
      CU(cudaMemsetAsync(undecided.p, 0, sizeof(int), 0));
      CU(launch_stab_norm(R[cur].p, Rt[cur].p, nb, m, ldm, logacc.p, state.p, kg.p, undecided.p, 0));
      int left = 0;
      CU(cudaMemcpy(&left, undecided.p, sizeof(int), cudaMemcpyDeviceToHost));
      if (left == 0) break;
    }
    CU(cudaMemcpy(hs.data(), state.p, (size_t)nb * sizeof(int), cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < nb; ++i) out[i0 + i] = hs[i] == 1 ? 1 : 0;
  }
  return 0;
}

// ---- spillover densities (McmcSpilloverRun::returnSpillover, spillover.h:246-260) ----
int bvhar_cuda_spillover(const bvhar_spillover_desc* d, double* connect, double* to, double* from, double* tot, double* net_pairwise) {
  if (!d || !connect || !to || !from || !tot || !net_pairwise) return fail(BVHAR_ERR_BAD_ARG, "null argument");
  if (d->abi_version != BVHAR_CUDA_ABI_VERSION) return fail(BVHAR_ERR_BAD_ARG, "descriptor ABI version mismatch");
  const int k = d->dim, p = d->lag, step = d->step, q = k * (k - 1) / 2;
  if (k < 1 || p < 1 || d->n_draws < 0 || !d->coef || !d->sv || (q > 0 && !d->contem)) return fail(BVHAR_ERR_BAD_ARG, "bad spillover descriptor");
  if (step < 2) return fail(BVHAR_ERR_BAD_ARG, "'lag_max' must larger than 0");  // lag_max = step - 1 (spillover.h:170); structural.h:10-12
  if (d->har_trans ? (d->nrow != 3 * k || d->ldh < 3 * k) : (d->nrow != k * p)) return fail(BVHAR_ERR_BAD_ARG, "coefficient rows do not match the lag structure");
  if (d->sv_kind < 0 || d->sv_kind > 3 || (d->sv_kind >= 2 && d->n_time < 1)) return fail(BVHAR_ERR_BAD_ARG, "bad sv_kind / n_time");
  if (d->n_draws == 0) return 0;
  API_LOCK;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(BVHAR_ERR_CUDA, "no CUDA device available: no CPU fallback");
  CU(cudaSetDevice(d->device));
  pool_setup(d->device);
  const int64_t S = d->n_draws;
  const size_t svw = d->sv_kind >= 2 ? (size_t)d->n_time * k : (size_t)k;
  if (d->sv_kind == 3) {  // one spillover per time point: unit = time point, the same coefficient rows for every unit
    DevBuf<double> dc, da, dsv, dh, scratch, oc, ot, of, otot, on;
    if (d->ld_coef < (int64_t)d->nrow * k || d->ld_contem < q || d->ld_sv < (int64_t)svw) return fail(BVHAR_ERR_BAD_ARG, "leading dimensions too small");
    const int64_t T = d->n_time, tb = std::max<int64_t>(1, std::min<int64_t>(T, (int64_t)((1ull << 30) / (sizeof(double) * S * (2 * (size_t)k * k + 3 * k)))));
    CU(dc.alloc((size_t)S * d->ld_coef)); CU(da.alloc((size_t)S * std::max<int64_t>(1, d->ld_contem))); CU(dsv.alloc((size_t)S * d->ld_sv));
    CU(cudaMemcpy(dc.p, d->coef, (size_t)S * d->ld_coef * sizeof(double), cudaMemcpyHostToDevice));
    if (q > 0) CU(cudaMemcpy(da.p, d->contem, (size_t)S * d->ld_contem * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dsv.p, d->sv, (size_t)S * d->ld_sv * sizeof(double), cudaMemcpyHostToDevice));
    CU(oc.alloc((size_t)tb * S * k * k)); CU(on.alloc((size_t)tb * S * k * k)); CU(ot.alloc((size_t)tb * S * k)); CU(of.alloc((size_t)tb * S * k));
    CU(otot.alloc((size_t)tb * S));
    if (d->har_trans) {
      std::vector<double> h((size_t)p * k * 3 * k);
      for (int r = 0; r < p * k; ++r)
        for (int qq = 0; qq < 3 * k; ++qq) h[(size_t)r * 3 * k + qq] = d->har_trans[(size_t)r * d->ldh + qq];
      CU(dh.upload(h));
    }
    const int n_cta = (int)std::min<int64_t>(tb * S, 592);
    SpilloverArgs A{};
    A.k = k; A.nrow = d->nrow; A.lag = p; A.step = step; A.sv_kind = 1; A.n_time = 1; A.ldh = 3 * k;
    A.ld_coef = d->ld_coef; A.ld_contem = d->ld_contem; A.ld_sv = d->ld_sv;
    A.per_unit = S; A.us_coef = 0; A.us_contem = 0; A.us_sv = k;
    A.coef = dc.p; A.contem = da.p; A.har = dh.p;
    A.slab = spillover_slab_doubles(k, p, step);
    CU(scratch.alloc((size_t)n_cta * A.slab));
    A.scratch = scratch.p; A.connect = oc.p; A.to = ot.p; A.from = of.p; A.tot = otot.p; A.net_pairwise = on.p;
    for (int64_t t0 = 0; t0 < T; t0 += tb) {
      const int64_t nt = std::min(tb, T - t0);
      A.sv = dsv.p + (size_t)t0 * k;
      A.n_draws = nt * S;
      CU(launch_spillover(A, n_cta, 0));
      CU(cudaDeviceSynchronize());
      const size_t o = (size_t)t0 * S;
      CU(cudaMemcpy(connect + o * k * k, oc.p, (size_t)nt * S * k * k * sizeof(double), cudaMemcpyDeviceToHost));
      CU(cudaMemcpy(net_pairwise + o * k * k, on.p, (size_t)nt * S * k * k * sizeof(double), cudaMemcpyDeviceToHost));
      CU(cudaMemcpy(to + o * k, ot.p, (size_t)nt * S * k * sizeof(double), cudaMemcpyDeviceToHost));
      CU(cudaMemcpy(from + o * k, of.p, (size_t)nt * S * k * sizeof(double), cudaMemcpyDeviceToHost));
      CU(cudaMemcpy(tot + o, otot.p, (size_t)nt * S * sizeof(double), cudaMemcpyDeviceToHost));
    }
    return 0;
  }
  if (d->ld_coef < (int64_t)d->nrow * k || d->ld_contem < q || d->ld_sv < (int64_t)svw) return fail(BVHAR_ERR_BAD_ARG, "leading dimensions too small");
  DevBuf<double> dc, da, dsv, dh, scratch, oc, ot, of, otot, on;
  const int64_t batch = std::min<int64_t>(S, std::max<int64_t>(1, (int64_t)((1ull << 30) / (sizeof(double) * (d->ld_coef + d->ld_contem + d->ld_sv + 2 * (size_t)k * k + 3 * k)))));
  CU(dc.alloc((size_t)batch * d->ld_coef)); CU(da.alloc((size_t)batch * std::max<int64_t>(1, d->ld_contem))); CU(dsv.alloc((size_t)batch * d->ld_sv));
  CU(oc.alloc((size_t)batch * k * k)); CU(on.alloc((size_t)batch * k * k)); CU(ot.alloc((size_t)batch * k)); CU(of.alloc((size_t)batch * k));
  CU(otot.alloc((size_t)batch));
  if (d->har_trans) {
    std::vector<double> h((size_t)p * k * 3 * k);
    for (int r = 0; r < p * k; ++r)
      for (int qq = 0; qq < 3 * k; ++qq) h[(size_t)r * 3 * k + qq] = d->har_trans[(size_t)r * d->ldh + qq];
    CU(dh.upload(h));
  }
  const int n_cta = (int)std::min<int64_t>(batch, 592);
  SpilloverArgs A{};
  A.k = k; A.nrow = d->nrow; A.lag = p; A.step = step; A.sv_kind = d->sv_kind; A.n_time = d->n_time; A.ldh = 3 * k;
  A.ld_coef = d->ld_coef; A.ld_contem = d->ld_contem; A.ld_sv = d->ld_sv;
  A.per_unit = batch; A.us_coef = 0; A.us_contem = 0; A.us_sv = 0;
  A.coef = dc.p; A.contem = da.p; A.sv = dsv.p; A.har = dh.p;
  A.slab = spillover_slab_doubles(k, p, step);
  CU(scratch.alloc((size_t)n_cta * A.slab));
  A.scratch = scratch.p; A.connect = oc.p; A.to = ot.p; A.from = of.p; A.tot = otot.p; A.net_pairwise = on.p;
  for (int64_t i0 = 0; i0 < S; i0 += batch) {
    const int64_t nb = std::min(batch, S - i0);
    CU(cudaMemcpy(dc.p, d->coef + (size_t)i0 * d->ld_coef, (size_t)nb * d->ld_coef * sizeof(double), cudaMemcpyHostToDevice));
    if (q > 0) CU(cudaMemcpy(da.p, d->contem + (size_t)i0 * d->ld_contem, (size_t)nb * d->ld_contem * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dsv.p, d->sv + (size_t)i0 * d->ld_sv, (size_t)nb * d->ld_sv * sizeof(double), cudaMemcpyHostToDevice));
    A.n_draws = nb;
    CU(launch_spillover(A, n_cta, 0));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(connect + (size_t)i0 * k * k, oc.p, (size_t)nb * k * k * sizeof(double), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(net_pairwise + (size_t)i0 * k * k, on.p, (size_t)nb * k * k * sizeof(double), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(to + (size_t)i0 * k, ot.p, (size_t)nb * k * sizeof(double), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(from + (size_t)i0 * k, of.p, (size_t)nb * k * sizeof(double), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(tot + (size_t)i0, otot.p, (size_t)nb * sizeof(double), cudaMemcpyDeviceToHost));
  }
  return 0;
}

// ---- rolling-window spillover with a refit per window (DynamicLdltSpillover::returnSpillover, spillover.h:272-443) ----
static int dynspill_batch(const bvhar_fit_desc* fd, const bvhar_dynspill_params* fp, double* to, double* from, double* tot) {
  bvhar_fit_desc d2 = *fd;
  const bool sv = fd->cov_type == BVHAR_COV_SV;
  d2.record_mask = BVHAR_REC_COEF | (sv ? BVHAR_REC_LVOL : 0u) | (fp->sparse ? BVHAR_REC_SPARSE : 0u);
  bvhar_fit* fit = nullptr;
  int rc = build(&d2, &fit);
  if (rc) return rc;
  std::unique_ptr<bvhar_fit> guard(fit);
  rc = bvhar_cuda_fit_run(fit, nullptr, nullptr);
  if (rc && rc != BVHAR_ERR_NUMERICAL) return rc;
  API_LOCK;
  const Dims& D = fit->D;
  const int k = D.k, U = D.U, q = D.q, p = fp->var_lag;
  const long long S = (fit->rec_filled + fit->thin - 1) / fit->thin, cap = fit->R.cap, th = fit->thin;
  if (S < 1) return fail(BVHAR_ERR_BAD_ARG, "no draws recorded");
  if (fp->har_trans ? (D.nrow != 3 * k) : (D.nrow != p * k)) return fail(BVHAR_ERR_BAD_ARG, "var_lag does not match the coefficient matrix");
  const RecPtrs& R = fit->R;
  DevBuf<double> dh, scratch, oc, ot, of, otot, on;
  if (fp->har_trans) {
    std::vector<double> h((size_t)p * k * 3 * k);
    for (int r = 0; r < p * k; ++r)
      for (int qq = 0; qq < 3 * k; ++qq) h[(size_t)r * 3 * k + qq] = fp->har_trans[(size_t)r * fp->ldh + qq];
    CU(dh.upload(h));
  }
  const long long nd = (long long)U * S;
  CU(oc.alloc((size_t)nd * k * k)); CU(on.alloc((size_t)nd * k * k)); CU(ot.alloc((size_t)nd * k)); CU(of.alloc((size_t)nd * k)); CU(otot.alloc((size_t)nd));
  const int n_cta = (int)std::min<long long>(nd, 592);
  SpilloverArgs A{};
  A.k = k; A.nrow = D.nrow; A.lag = p; A.step = fp->step; A.ldh = 3 * k; A.n_draws = nd; A.per_unit = S;
  A.coef = fp->sparse ? R.alpha_sparse : R.alpha; A.ld_coef = th * D.N; A.us_coef = cap * D.N;
  A.contem = fp->sparse ? R.a_sparse : R.a; A.ld_contem = th * q; A.us_contem = cap * q;
  if (sv) { A.sv = R.h; A.sv_kind = 2; A.n_time = D.nmax; A.ld_sv = th * (long long)D.nmax * k; A.us_sv = cap * (long long)D.nmax * k; }
  else { A.sv = R.d; A.sv_kind = 0; A.n_time = 0; A.ld_sv = th * k; A.us_sv = cap * k; }
  if (!A.coef || (q > 0 && !A.contem) || !A.sv) return fail(BVHAR_ERR_BAD_ARG, "records needed by the spillover were not kept");
  A.har = dh.p;
  A.slab = spillover_slab_doubles(k, p, fp->step);
  CU(scratch.alloc((size_t)n_cta * A.slab));
  A.scratch = scratch.p; A.connect = oc.p; A.to = ot.p; A.from = of.p; A.tot = otot.p; A.net_pairwise = on.p;
  CU(cudaStreamSynchronize(fit->st));
  CU(launch_spillover(A, n_cta, 0));
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(to, ot.p, (size_t)nd * k * sizeof(double), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(from, of.p, (size_t)nd * k * sizeof(double), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(tot, otot.p, (size_t)nd * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

int bvhar_cuda_dynamic_spillover(const bvhar_fit_desc* fd, const bvhar_dynspill_params* fp, double* to, double* from, double* tot) {
  if (!fd || !fp || !to || !from || !tot) return fail(BVHAR_ERR_BAD_ARG, "null argument");
  if (fp->abi_version != BVHAR_CUDA_ABI_VERSION) return fail(BVHAR_ERR_BAD_ARG, "descriptor ABI version mismatch");
  if (fp->step < 2) return fail(BVHAR_ERR_BAD_ARG, "'lag_max' must larger than 0");
  if (fd->n_windows < 1 || fd->n_chains < 1 || !fd->win_row0 || !fd->win_nrow || !fd->seed_chain) return fail(BVHAR_ERR_BAD_ARG, "empty window table");
  if (fp->har_trans && fp->ldh < 3 * fd->dim) return fail(BVHAR_ERR_BAD_ARG, "har_trans leading dimension too small");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(BVHAR_ERR_CUDA, "no CUDA device available: no CPU fallback");
  CU(cudaSetDevice(fd->device));
  const int W = fd->n_windows, C = fd->n_chains, k = fd->dim;
  const int thin = fd->thin < 1 ? 1 : fd->thin;
  const long long S = ((long long)std::max(0, fd->num_iter - fd->num_burn) + thin - 1) / thin;
  size_t free_b = 0, total_b = 0;
  CU(cudaMemGetInfo(&free_b, &total_b));
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, fd->device) == cudaSuccess) {
    unsigned long long reserved = 0, used = 0;
    if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && reserved > used)
      free_b += (size_t)(reserved - used);
  }
  size_t nmax = 1;
  for (int w = 0; w < W; ++w) nmax = std::max(nmax, (size_t)fd->win_nrow[w]);
  const size_t d_ = fd->dim_design;
  const size_t shared_b = (size_t)fd->n_rows_total * (d_ * (d_ + 1) / 2 + 2 * d_ + 2 * k + 8) * sizeof(double);
  size_t unit_b = outforecast_unit_bytes(fd, 1) + (size_t)S * (2 * (size_t)k * k + 3 * k) * sizeof(double);
  if (fd->cov_type == BVHAR_COV_SV) unit_b += (size_t)std::max(1, fd->num_iter - fd->num_burn) * nmax * k * sizeof(double);
  const size_t budget = free_b > shared_b ? (size_t)(0.85 * (double)(free_b - shared_b)) : 0;
  long long batch = std::max<long long>(1, (long long)(budget / ((size_t)C * unit_b)));
  if (const char* env = getenv("BVHAR_MAX_WINDOWS_PER_BATCH")) batch = std::max(1, atoi(env));
  batch = std::min<long long>(batch, W);
  for (int w0 = 0; w0 < W;) {
    const int nw = (int)std::min<long long>(batch, W - w0);
    bvhar_fit_desc sub = *fd;
    sub.n_windows = nw;
    sub.win_row0 = fd->win_row0 + w0;
    sub.win_nrow = fd->win_nrow + w0;
    sub.seed_chain = fd->seed_chain + (size_t)w0 * C;
    sub.unit_id_base = fd->unit_id_base + w0 * C;
    const size_t o = (size_t)w0 * C * S;
    const int rc = dynspill_batch(&sub, fp, to + o * k, from + o * k, tot + o);
    if (rc == BVHAR_ERR_CUDA && nw > 1 && g_err.find("out of memory") != std::string::npos) {
      cudaGetLastError();
      cudaDeviceSynchronize();
      if (cudaDeviceGetDefaultMemPool(&pool, fd->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
      batch = (nw + 1) / 2;
      continue;
    }
    if (rc) return rc;
    w0 += nw;
  }
  return 0;
}

// ---- building blocks ------------------------------------------------------------------------
int bvhar_cuda_dgemm_tn(int device, int64_t M, int64_t N, int64_t K, const double* A, int64_t lda, const double* B,
                        int64_t ldb, double* C, int64_t ldc) {
  API_LOCK;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(BVHAR_ERR_CUDA, "no CUDA device available: no CPU fallback");
  CU(cudaSetDevice(device));
  DevBuf<double> dA, dB, dC;
  const int64_t la = (lda + 1) & ~1LL, lb = (ldb + 1) & ~1LL;
  CU(dA.alloc((size_t)K * la)); CU(dB.alloc((size_t)K * lb)); CU(dC.alloc((size_t)M * ldc));
  CU(cudaMemcpy2D(dA.p, la * sizeof(double), A, lda * sizeof(double), M * sizeof(double), K, cudaMemcpyHostToDevice));
  CU(cudaMemcpy2D(dB.p, lb * sizeof(double), B, ldb * sizeof(double), N * sizeof(double), K, cudaMemcpyHostToDevice));
  GemmArgs g{};
  g.A = dA.p; g.B = dB.p; g.C = dC.p; g.M = (int)M; g.N = (int)N; g.K = (int)K; g.lda = (int)la; g.ldb = (int)lb; g.ldc = (int)ldc;
  g.groups = 1; g.epi = EPI_STORE; g.aligned16 = 1;
  DevBuf<double> ws;
  {  // few output tiles and a long K: the split-K path of the persistent kernel, as the engine uses it for small batches
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const long long tiles = ((M + GBM - 1) / GBM) * ((N + GBN - 1) / GBN);
    const int S = gemm_choose_ksplit(tiles, (int)K, sms);
    if (S > 1) {
      CU(ws.alloc((size_t)S * M * ldc));
      g.ksplit = S; g.ws = ws.p; g.ws_stride = (long long)M * ldc;
    }
  }
  CU(launch_dgemm_tn(g, 0));
  CU(launch_splitk_reduce(g, (long long)M * ldc, 0));
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(C, dC.p, (size_t)M * ldc * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

int bvhar_cuda_draw_coef_batched(int device, int64_t batch, int32_t d, const double* P_packed, const double* b, const double* z,
                                 double* out) {
  API_LOCK;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(BVHAR_ERR_CUDA, "no CUDA device available: no CPU fallback");
  if (batch < 1 || d < 1) return fail(BVHAR_ERR_BAD_ARG, "empty batch");
  CU(cudaSetDevice(device));
  const size_t np = (size_t)d * (d + 1) / 2;
  DevBuf<double> dP, db, dz, dout;
  CU(dP.alloc(batch * np)); CU(db.alloc((size_t)batch * d)); CU(dz.alloc((size_t)batch * d)); CU(dout.alloc((size_t)batch * d));
  CU(cudaMemcpy(dP.p, P_packed, batch * np * sizeof(double), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(db.p, b, (size_t)batch * d * sizeof(double), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(dz.p, z, (size_t)batch * d * sizeof(double), cudaMemcpyHostToDevice));
  cudaError_t e = launch_draw_coef_batched(batch, d, dP.p, db.p, dz.p, dout.p, 0);
  if (e == cudaErrorInvalidConfiguration) return fail(BVHAR_ERR_UNSUPPORTED, "d too large for the shared-memory Cholesky");
  CU(e);
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(out, dout.p, (size_t)batch * d * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

}  // extern "C"

// ---- conjugate Minnesota samplers (kernels_mniw.cu) -------------------------------------------------------------
namespace {
// lower Cholesky factor of a column-major symmetric n x n matrix (LLT, lower triangle referenced); false if not PD
bool host_chol_lower(const double* a, int n, std::vector<double>& L) {
  L.assign((size_t)n * n, 0.0);  // row-major L[i * n + j], j <= i
  for (int j = 0; j < n; ++j) {
    double s = a[(size_t)j * n + j];
    for (int m = 0; m < j; ++m) s -= L[(size_t)j * n + m] * L[(size_t)j * n + m];
    if (!(s > 0.0)) return false;
    const double piv = std::sqrt(s);
    L[(size_t)j * n + j] = piv;
    for (int i = j + 1; i < n; ++i) {
      double t = a[(size_t)j * n + i];  // A(i, j), column-major
      for (int m = 0; m < j; ++m) t -= L[(size_t)i * n + m] * L[(size_t)j * n + m];
      L[(size_t)i * n + j] = t / piv;
    }
  }
  return true;
}
// general inverse (Gauss-Jordan, partial pivoting) of a column-major n x n matrix: Eigen's inverse() of minnesota.h:427
bool host_inverse(const double* a, int n, std::vector<double>& inv) {
  std::vector<double> w((size_t)n * 2 * n, 0.0);
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) w[(size_t)i * 2 * n + j] = a[(size_t)j * n + i];
    w[(size_t)i * 2 * n + n + i] = 1.0;
  }
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int i = c + 1; i < n; ++i)
      if (std::fabs(w[(size_t)i * 2 * n + c]) > std::fabs(w[(size_t)piv * 2 * n + c])) piv = i;
    if (w[(size_t)piv * 2 * n + c] == 0.0) return false;
    if (piv != c)
      for (int j = 0; j < 2 * n; ++j) std::swap(w[(size_t)c * 2 * n + j], w[(size_t)piv * 2 * n + j]);
    const double d = 1.0 / w[(size_t)c * 2 * n + c];
    for (int j = 0; j < 2 * n; ++j) w[(size_t)c * 2 * n + j] *= d;
    for (int i = 0; i < n; ++i) {
      if (i == c) continue;
      const double f = w[(size_t)i * 2 * n + c];
      if (f == 0.0) continue;
      for (int j = 0; j < 2 * n; ++j) w[(size_t)i * 2 * n + j] -= f * w[(size_t)c * 2 * n + j];
    }
  }
  inv.assign((size_t)n * n, 0.0);  // column-major
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) inv[(size_t)j * n + i] = w[(size_t)i * 2 * n + n + j];
  return true;
}
}  // namespace

extern "C" int bvhar_cuda_mniw_run(const bvhar_mniw_desc* d, double* alpha_record, double* sigma_record, double* lambda_record,
                                   double* psi_record, uint8_t* accept_record) {
  API_LOCK;
  if (!d) return fail(BVHAR_ERR_BAD_ARG, "null descriptor");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(BVHAR_ERR_CUDA, "no CUDA device available: no CPU fallback");
  const int k = d->dim, dd = d->dim_design, C = d->num_chains;
  const int thin = d->thin < 1 ? 1 : d->thin;
  if (k < 1 || dd < 1 || C < 1 || d->num_iter < 1 || d->num_burn < 0 || d->num_burn >= d->num_iter)
    return fail(BVHAR_ERR_BAD_ARG, "mniw: dim, dim_design, num_chains >= 1 and 0 <= num_burn < num_iter are required");
  if (!d->mn_mean || !d->mn_prec || !d->iw_scale || !d->seed_chain) return fail(BVHAR_ERR_BAD_ARG, "mniw: null input");
  if (d->iw_shape <= k - 1) return fail(BVHAR_ERR_BAD_ARG, "Wrong 'shape'. shape > dim - 1 must be satisfied.");  // random.h:179-181
  if (k > 118) return fail(BVHAR_ERR_UNSUPPORTED, "mniw: dim <= 118 (two dim x dim work matrices per CTA in shared memory)");
  if (d->mh && (!d->init_par || !d->init_hess || !d->init_scale_variance)) return fail(BVHAR_ERR_BAD_ARG, "mniw: mh = 1 needs the inits");
  const int rows = (d->num_iter - d->num_burn + thin - 1) / thin;
  CU(cudaSetDevice(d->device));
  pool_setup(d->device);

  std::vector<double> Lp, Ls, R((size_t)dd * dd, 0.0);
  if (!host_chol_lower(d->mn_prec, dd, Lp)) return fail(BVHAR_ERR_NUMERICAL, "mniw: the matrix-normal precision is not positive definite");
  if (!host_chol_lower(d->iw_scale, k, Ls)) return fail(BVHAR_ERR_NUMERICAL, "mniw: the inverse-Wishart scale is not positive definite");
  for (int i = 0; i < dd; ++i)
    for (int m = i; m < dd; ++m) R[(size_t)i * dd + m] = Lp[(size_t)m * dd + i];  // U_u = L'
  DevBuf<double> dmean, dR, dLs, dalpha, dsigma, dws;
  DevBuf<uint32_t> dseed;
  DevBuf<int> dflags;
  CU(dmean.upload(std::vector<double>(d->mn_mean, d->mn_mean + (size_t)dd * k)));
  CU(dR.upload(R)); CU(dLs.upload(Ls));
  CU(dseed.upload(std::vector<uint32_t>(d->seed_chain, d->seed_chain + C)));
  CU(dflags.alloc(1));
  CU(dalpha.alloc((size_t)C * dd * k * rows)); CU(dsigma.alloc((size_t)C * k * k * rows));

  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, d->device));
  const size_t limit = 200 * 1024;
  const bool w_smem = mniw_smem(k, dd, true) <= limit;
  const size_t smem = mniw_smem(k, dd, w_smem);
  int per_sm = (int)std::min<size_t>(4, std::max<size_t>(1, (220 * 1024) / (smem + 1024)));
  const long long total = (long long)C * rows;
  const int grid = (int)std::min<long long>(total, (long long)prop.multiProcessorCount * per_sm);
  if (!w_smem) CU(dws.alloc((size_t)grid * dd * k));
  MniwArgs A{};
  A.k = k; A.d = dd; A.chains = C; A.rows = rows; A.num_burn = d->num_burn; A.thin = thin; A.shape = d->iw_shape;
  A.mean = dmean.p; A.Rfull = dR.p; A.Ls = dLs.p; A.seeds = dseed.p; A.ubase = d->unit_id_base;
  A.wscratch = w_smem ? nullptr : dws.p; A.alpha = dalpha.p; A.sigma = dsigma.p; A.flags = dflags.p;
  CU(launch_mniw(A, grid, smem, 0));

  DevBuf<double> dpar, dG, dlam, dpsi;
  DevBuf<unsigned char> dacc_all, dacc;
  if (d->mh) {
    const int np = 1 + k;
    std::vector<double> G((size_t)C * np * np, 0.0), inv, V((size_t)np * np), Lv;
    for (int c = 0; c < C; ++c) {
      if (!host_inverse(d->init_hess + (size_t)c * np * np, np, inv)) return fail(BVHAR_ERR_NUMERICAL, "mniw: the hessian of the inits is singular");
      for (size_t e = 0; e < (size_t)np * np; ++e) V[e] = d->init_scale_variance[c] * inv[e];  // gaussian_variance, minnesota.h:427
      if (!host_chol_lower(V.data(), np, Lv)) return fail(BVHAR_ERR_NUMERICAL, "mniw: scale_variance * hessian^-1 is not positive definite");
      for (int i = 0; i < np; ++i)
        for (int j = i; j < np; ++j) G[(size_t)c * np * np + (size_t)i * np + j] = Lv[(size_t)j * np + i];  // llt().matrixU()
    }
    CU(dpar.upload(std::vector<double>(d->init_par, d->init_par + (size_t)C * np)));
    CU(dG.upload(G));
    CU(dacc_all.alloc((size_t)C * d->num_iter)); CU(dacc.alloc((size_t)C * rows));
    CU(dlam.alloc((size_t)C * rows)); CU(dpsi.alloc((size_t)C * k * rows));
    MhArgs M{};
    M.k = k; M.chains = C; M.rows = rows; M.num_iter = d->num_iter; M.num_burn = d->num_burn; M.thin = thin;
    M.gamma_shp = d->gam_rate; M.gamma_rate = d->gam_shape;  // the swap of minnesota.h:426
    M.invgam_shp = d->invgam_shape; M.invgam_scl = d->invgam_scl;
    M.init_par = dpar.p; M.chol_var = dG.p; M.seeds = dseed.p; M.ubase = d->unit_id_base;
    M.accept_all = dacc_all.p; M.lambda = dlam.p; M.psi = dpsi.p; M.accept = dacc.p; M.flags = dflags.p;
    CU(launch_mh(M, 0));
  }
  CU(cudaDeviceSynchronize());
  int flags = 0;
  CU(cudaMemcpy(&flags, dflags.p, sizeof(int), cudaMemcpyDeviceToHost));
  if (alpha_record) CU(cudaMemcpy(alpha_record, dalpha.p, dalpha.n * sizeof(double), cudaMemcpyDeviceToHost));
  if (sigma_record) CU(cudaMemcpy(sigma_record, dsigma.p, dsigma.n * sizeof(double), cudaMemcpyDeviceToHost));
  if (d->mh) {
    if (lambda_record) CU(cudaMemcpy(lambda_record, dlam.p, dlam.n * sizeof(double), cudaMemcpyDeviceToHost));
    if (psi_record) CU(cudaMemcpy(psi_record, dpsi.p, dpsi.n * sizeof(double), cudaMemcpyDeviceToHost));
    if (accept_record) CU(cudaMemcpy(accept_record, dacc.p, dacc.n, cudaMemcpyDeviceToHost));
  }
  if (flags & 2) return fail(BVHAR_ERR_NUMERICAL, "'x' should be larger than 0.");  // invgamma_dens, common.h:503-505
  if (flags & 1) return fail(BVHAR_ERR_NUMERICAL, "mniw: a sampled covariance matrix was not positive definite");
  return 0;
}
