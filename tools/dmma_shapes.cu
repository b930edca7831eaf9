__global__ void k16(double* out, const double* in) {
  double a[8], b[4], c[4] = {0,0,0,0};
  for (int i = 0; i < 8; ++i) a[i] = in[threadIdx.x * 8 + i];
  for (int i = 0; i < 4; ++i) b[i] = in[1024 + threadIdx.x * 4 + i];
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
    : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
    : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  for (int i = 0; i < 4; ++i) out[threadIdx.x * 4 + i] = c[i];
}
__global__ void k8(double* out, const double* in) {
  double a[4], b[2], c[4] = {0,0,0,0};
  for (int i = 0; i < 4; ++i) a[i] = in[threadIdx.x * 8 + i];
  for (int i = 0; i < 2; ++i) b[i] = in[1024 + threadIdx.x * 4 + i];
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
    : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
    : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
  for (int i = 0; i < 4; ++i) out[threadIdx.x * 4 + i] = c[i];
}
__global__ void k4(double* out, const double* in) {
  double a[2], b[1], c[4] = {0,0,0,0};
  for (int i = 0; i < 2; ++i) a[i] = in[threadIdx.x * 8 + i];
  b[0] = in[1024 + threadIdx.x * 4];
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
    : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
    : "d"(a[0]), "d"(a[1]), "d"(b[0]));
  for (int i = 0; i < 4; ++i) out[threadIdx.x * 4 + i] = c[i];
}
