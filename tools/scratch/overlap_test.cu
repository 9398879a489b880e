#include <cstdio>
#include <cuda_runtime.h>
#include <chrono>
__global__ void spin(long long cycles, int smem_touch) {
  extern __shared__ double sm[];
  if (smem_touch) sm[threadIdx.x] = 1.0;
  long long t0 = clock64();
  while (clock64() - t0 < cycles) {}
}
int main(int argc, char** argv) {
  int use_smem = argc > 1 ? atoi(argv[1]) : 0;
  const size_t N = 167u << 20; const int C = 8;
  char *h, *d, *h2, *d2; cudaMallocHost(&h, N); cudaMalloc(&d, N); cudaMallocHost(&h2, 1 << 20); cudaMalloc(&d2, 1 << 20);
  cudaStream_t sa, sb; cudaStreamCreateWithFlags(&sa, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking);
  cudaEvent_t ev[C]; for (auto& e : ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  if (use_smem) cudaFuncSetAttribute(spin, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int rep = 0; rep < 3; ++rep) {
    cudaDeviceSynchronize();
    auto t0 = std::chrono::steady_clock::now();
    if (argc > 2) cudaMemcpyAsync(d2, h2, 300000, cudaMemcpyHostToDevice, sb);
    for (int c = 0; c < C; ++c) { cudaMemcpyAsync(d + c * (N / C), h + c * (N / C), N / C, cudaMemcpyHostToDevice, sa); cudaEventRecord(ev[c], sa); }
    for (int c = 0; c < C; ++c) {
      cudaStreamWaitEvent(sb, ev[c], 0);
      spin<<<148 * 5, 128, use_smem ? 40 * 1024 : 0, sb>>>(1600000, use_smem);   // ~0.8 ms
      cudaMemcpyAsync(h + c * 1024, d + c * 1024, 1024, cudaMemcpyDeviceToHost, sb);
    }
    cudaStreamSynchronize(sb);
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    printf("smem=%d rep %d total %.2f ms (copies alone ~3.5, kernels alone ~6.5)\n", use_smem, rep, ms);
  }
  return 0;
}
