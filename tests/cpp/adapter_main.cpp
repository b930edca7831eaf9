// Drives include/bvhar_b200/mcmc_run.hpp from a problem file written by tests/test_cpp_adapter.py and writes the
// records back: the C++ host adapter is checked against the Python one on identical inputs.
// File format (both ways): repeated { int32 name_len, name bytes, int64 rows, int64 cols, rows*cols doubles }.
#include <cstdio>
#include <cstring>
#include <bvhar_b200/outforecast_run.hpp>

using namespace bvhar_b200;

static std::map<std::string, Matrix> read_all(const char* path) {
  std::map<std::string, Matrix> out;
  FILE* f = std::fopen(path, "rb");
  if (!f) throw std::runtime_error("cannot open problem file");
  int32_t nl;
  while (std::fread(&nl, sizeof nl, 1, f) == 1) {
    std::string name(nl, '\0');
    int64_t r, c;
    if (std::fread(&name[0], 1, nl, f) != (size_t)nl || std::fread(&r, sizeof r, 1, f) != 1 || std::fread(&c, sizeof c, 1, f) != 1)
      throw std::runtime_error("truncated problem file");
    Matrix m(r, c);
    if (r * c && std::fread(m.data.data(), sizeof(double), (size_t)(r * c), f) != (size_t)(r * c)) throw std::runtime_error("truncated");
    out.emplace(name, std::move(m));
  }
  std::fclose(f);
  return out;
}
static ParamList sub(const std::map<std::string, Matrix>& all, const std::string& prefix) {
  ParamList p;
  for (auto& kv : all)
    if (kv.first.compare(0, prefix.size(), prefix) == 0) p[kv.first.substr(prefix.size())] = kv.second.data;
  return p;
}
static std::vector<int32_t> ints(const Matrix& m) { return std::vector<int32_t>(m.data.begin(), m.data.end()); }

int main(int argc, char** argv) {
  if (argc != 3) { std::fprintf(stderr, "usage: adapter_main problem.bin records.bin\n"); return 2; }
  // optional "forecast" entry of the problem file: lag | (week, month), step, stable -> McmcForecastRun on the fit's records
  try {
    auto all = read_all(argv[1]);
    const Matrix& meta = all.at("meta");  // chains, iter, burn, thin, prior_type, include_mean, sv, ggl
    const int chains = (int)meta.data[0], iters = (int)meta.data[1], burn = (int)meta.data[2], thin = (int)meta.data[3];
    const int prior = (int)meta.data[4];
    const bool mean = meta.data[5] != 0, sv = meta.data[6] != 0, ggl = meta.data[7] != 0;
    std::vector<ParamList> init;
    for (int c = 0; c < chains; ++c) init.push_back(sub(all, "init" + std::to_string(c) + "."));
    std::vector<uint32_t> seeds(all.at("seed_chain").data.begin(), all.at("seed_chain").data.end());
    std::vector<Records> rec;
#define RUN(T) rec = T(chains, iters, burn, thin, all.at("x"), all.at("y"), sub(all, "cov."), sub(all, "prior."), sub(all, "icpt."), init, \
                       prior, ints(all.at("grp_id")), ints(all.at("own_id")), ints(all.at("cross_id")), ints(all.at("grp_mat")), mean, seeds, false, 1).returnRecords()
    if (sv && ggl) RUN(SvMcmc); else if (sv) RUN(SvGrpMcmc); else if (ggl) RUN(McmcLdlt); else RUN(McmcLdltGrp);
    FILE* f = std::fopen(argv[2], "wb");
    for (int c = 0; c < chains; ++c)
      for (auto& kv : rec[c]) {
        const std::string name = "c" + std::to_string(c) + "." + kv.first;
        const int32_t nl = (int32_t)name.size();
        std::fwrite(&nl, sizeof nl, 1, f); std::fwrite(name.data(), 1, nl, f);
        std::fwrite(&kv.second.rows, sizeof(int64_t), 1, f); std::fwrite(&kv.second.cols, sizeof(int64_t), 1, f);
        std::fwrite(kv.second.data.data(), sizeof(double), kv.second.data.size(), f);
      }
    if (all.count("forecast")) {  // VAR: (lag, 0, step, stable); VHAR: (week, month, step, stable)
      const Matrix& fc = all.at("forecast");
      const int a0 = (int)fc.data[0], a1 = (int)fc.data[1], step = (int)fc.data[2];
      const bool stable = fc.data[3] != 0;
      std::vector<Matrix> dens;
      if (sv) dens = a1 > 0 ? SvForecast(chains, a0, a1, step, all.at("y"), false, 0.0, rec, seeds, mean, stable, 1).returnForecast()
                            : SvForecast(chains, a0, step, all.at("y"), false, 0.0, rec, seeds, mean, stable, 1).returnForecast();
      else dens = a1 > 0 ? LdltForecast(chains, a0, a1, step, all.at("y"), false, 0.0, rec, seeds, mean, stable, 1).returnForecast()
                         : LdltForecast(chains, a0, step, all.at("y"), false, 0.0, rec, seeds, mean, stable, 1).returnForecast();
      for (int c = 0; c < chains; ++c) {
        const std::string name = "f" + std::to_string(c);
        const int32_t nl = (int32_t)name.size();
        std::fwrite(&nl, sizeof nl, 1, f); std::fwrite(name.data(), 1, nl, f);
        std::fwrite(&dens[c].rows, sizeof(int64_t), 1, f); std::fwrite(&dens[c].cols, sizeof(int64_t), 1, f);
        std::fwrite(dens[c].data.data(), sizeof(double), dens[c].data.size(), f);
      }
    }
    if (all.count("outforecast")) {
      // (vhar, expanding, lag | week, month, step, stable, level, get_lpl, sparse): the out-of-sample runners of
      // include/bvhar_b200/outforecast_run.hpp on the fit_record stored in the problem file ("rec<c>.<name>")
      const Matrix& of = all.at("outforecast");
      const bool vhar = of.data[0] != 0, expanding = of.data[1] != 0, stable = of.data[5] != 0, get_lpl = of.data[7] != 0, sparse = of.data[8] != 0;
      const int a0 = (int)of.data[2], a1 = (int)of.data[3], step = (int)of.data[4];
      const double level = of.data[6];
      std::vector<Records> fit_record(chains);
      for (int c = 0; c < chains; ++c) {
        const std::string pre = "rec" + std::to_string(c) + ".";
        for (auto& kv : all)
          if (kv.first.compare(0, pre.size(), pre) == 0) fit_record[c][kv.first.substr(pre.size())] = kv.second;
      }
      std::vector<uint32_t> sc(all.at("seed_window_chain").data.begin(), all.at("seed_window_chain").data.end());
      std::vector<uint32_t> sf(all.at("seed_forecast").data.begin(), all.at("seed_forecast").data.end());
      OutForecast res;
#define OUT_ARGS chains, iters, burn, thin, sparse, level, fit_record, sub(all, "cov."), sub(all, "prior."), sub(all, "icpt."), init, prior, \
                 ints(all.at("grp_id")), ints(all.at("own_id")), ints(all.at("cross_id")), ints(all.at("grp_mat")), mean, stable, step,     \
                 all.at("y_test"), get_lpl, sc, sf, false, 1
#define RUN_OUT(SV, GG)                                                                                                        \
      if (vhar && expanding) res = McmcVharforecastRun<Window::Expand, SV, GG>(all.at("y_train"), a0, a1, OUT_ARGS).returnForecast(); \
      else if (vhar) res = McmcVharforecastRun<Window::Roll, SV, GG>(all.at("y_train"), a0, a1, OUT_ARGS).returnForecast();           \
      else if (expanding) res = McmcVarforecastRun<Window::Expand, SV, GG>(all.at("y_train"), a0, OUT_ARGS).returnForecast();        \
      else res = McmcVarforecastRun<Window::Roll, SV, GG>(all.at("y_train"), a0, OUT_ARGS).returnForecast()
      if (sv && ggl) { RUN_OUT(true, true); } else if (sv) { RUN_OUT(true, false); } else if (ggl) { RUN_OUT(false, true); } else { RUN_OUT(false, false); }
      auto put = [&](const std::string& name, const Matrix& m) {
        const int32_t nl = (int32_t)name.size();
        std::fwrite(&nl, sizeof nl, 1, f); std::fwrite(name.data(), 1, nl, f);
        std::fwrite(&m.rows, sizeof(int64_t), 1, f); std::fwrite(&m.cols, sizeof(int64_t), 1, f);
        std::fwrite(m.data.data(), sizeof(double), m.data.size(), f);
      };
      for (size_t w = 0; w < res.forecast.size(); ++w)
        for (int c = 0; c < chains; ++c) put("o" + std::to_string(w) + "." + std::to_string(c), res.forecast[w][c]);
      if (res.has_lpl) put("lpl", res.lpl);
    }
    std::fclose(f);
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "bvhar_b200 adapter error: %s\n", e.what());
    return 3;
  }
}
