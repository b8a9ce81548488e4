// How many single-CTA kernels on distinct streams run concurrently? (scheduler design probe)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <chrono>
__global__ void spin(long long cycles, int* out) {
  extern __shared__ int sm[];
  long long t0 = clock64();
  int acc = 0;
  while (clock64() - t0 < cycles) acc += sm[threadIdx.x % 32] ;
  if (acc == 123456789) out[0] = acc;
}
int main(int argc, char** argv) {
  int smem = argc > 1 ? atoi(argv[1]) : 0;
  int blocks = argc > 2 ? atoi(argv[2]) : 1;
  int* d; cudaMalloc(&d, 4);
  cudaFuncSetAttribute(spin, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const long long cyc = 40000000;  // ~20 ms
  for (int n : {1, 4, 8, 16, 24, 32, 48, 64, 128}) {
    std::vector<cudaStream_t> st(n);
    for (auto& s : st) cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    cudaDeviceSynchronize();
    auto t0 = std::chrono::steady_clock::now();
    for (int k = 0; k < n; ++k) spin<<<blocks, 256, smem, st[k]>>>(cyc, d);
    cudaDeviceSynchronize();
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    printf("smem=%d blocks=%d streams=%d wall=%.1f ms (%.1fx of one)\n", smem, blocks, n, ms, ms / 20.4);
    for (auto& s : st) cudaStreamDestroy(s);
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
