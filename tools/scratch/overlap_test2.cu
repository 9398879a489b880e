#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <chrono>
__global__ void spin(long long cycles) { long long t0 = clock64(); while (clock64() - t0 < cycles) {} }
int main(int argc, char** argv) {
  int feat = argc > 1 ? atoi(argv[1]) : 0;   // bit0: timing events on sb, bit1: big D2H on third stream, bit2: small H2D first, bit3: D2H on sb
  const size_t N = 167u << 20; const int C = 8;
  char *h, *d, *h2, *d2, *h3; cudaMallocHost(&h, N); cudaMalloc(&d, N); cudaMallocHost(&h2, 1 << 20); cudaMalloc(&d2, 1 << 20); cudaMallocHost(&h3, 32 << 20);
  cudaStream_t sa, sb, sc; cudaStreamCreateWithFlags(&sa, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&sc, cudaStreamNonBlocking);
  cudaEvent_t ev[C], evd[C], tev[64]; for (auto& e : ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming); for (auto& e : evd) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  for (auto& e : tev) cudaEventCreate(&e);
  for (int rep = 0; rep < 3; ++rep) {
    cudaDeviceSynchronize();
    auto t0 = std::chrono::steady_clock::now();
    if (feat & 4) cudaMemcpyAsync(d2, h2, 300000, cudaMemcpyHostToDevice, sb);
    for (int c = 0; c < C; ++c) { cudaMemcpyAsync(d + c * (N / C), h + c * (N / C), N / C, cudaMemcpyHostToDevice, sa); cudaEventRecord(ev[c], sa); }
    for (int c = 0; c < C; ++c) {
      cudaStreamWaitEvent(sb, ev[c], 0);
      if (feat & 1) cudaEventRecord(tev[4 * c], sb);
      spin<<<148 * 5, 128, 0, sb>>>(1600000);
      if (feat & 1) cudaEventRecord(tev[4 * c + 1], sb);
      if (feat & 2) { cudaEventRecord(evd[c], sb); cudaStreamWaitEvent(sc, evd[c], 0); cudaMemcpyAsync(h3 + c * (3 << 20), d + c * (3 << 20), 3 << 20, cudaMemcpyDeviceToHost, sc); }
      if (feat & 8) cudaMemcpyAsync(h3 + c * (3 << 20), d + c * (3 << 20), 3 << 20, cudaMemcpyDeviceToHost, sb);
    }
    cudaStreamSynchronize(sb); cudaStreamSynchronize(sc);
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (rep == 2) printf("feat=%d total %.2f ms\n", feat, ms);
  }
  return 0;
}
