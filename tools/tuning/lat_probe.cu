// FP64 latency / throughput probe: dependent DFMA chains, varying ILP and warps per SM.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters, long long* cyc) {
  double a[ILP];
  for (int i = 0; i < ILP; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  const double m = 1.0000001, c = 1e-7;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u)
#pragma unroll
      for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], m, c);
  }
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < ILP; ++i) s += a[i];
  if (s == 1234.5) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ILP> void run(int warps) {
  double* o; long long* c; cudaMalloc(&o, 8); cudaMalloc(&c, 8);
  int iters = 2000;
  k<ILP><<<148, warps * 32>>>(o, iters, c); cudaDeviceSynchronize();
  k<ILP><<<148, warps * 32>>>(o, iters, c); cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  double per = (double)h / (iters * 16.0 * ILP);
  printf("ILP %d warps/SM %2d: %.2f cycles per DFMA per warp (chain step %.2f cyc), SM rate %.2f DFMA-warp-instr/cycle\n", ILP, warps, per, per * ILP, warps / per);
  cudaFree(o); cudaFree(c);
}
int main() {
  for (int w : {1, 4, 8, 16, 20, 32}) { run<1>(w); }
  for (int w : {4, 8, 20}) { run<2>(w); run<4>(w); }
  return 0;
}
