// Instruction-cache probe: a loop whose body is N independent-ish integer instructions (fully unrolled,
// straight line); reports SM cycles per instruction against the body size, with 4 or 16 warps per SM on
// every SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o icache icache.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int N>
__global__ void body(int* out, int iters, int seed) {
  int a = seed + threadIdx.x, b = seed * 3 + 1, c = threadIdx.x ^ 5, d = seed - 7;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < N / 4; ++k) {
      a = max(a + (k * 7 + 1), b);
      b = (b ^ c) + (k | 3);
      c = min(c + d, a + k);
      d = d + (a & (k * 5 + 1));
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[2 * blockIdx.x] = a + b + c + d; reinterpret_cast<long long*>(out + 2 * 100000)[blockIdx.x] = t1 - t0; }
}

template <int N>
void run(int* d_out, int blocks, int threads) {
  const long long total = 3000000;   // instructions per warp
  const int iters = static_cast<int>(total / N);
  body<N><<<blocks, threads>>>(d_out, 2, 1);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  body<N><<<blocks, threads>>>(d_out, iters, 1);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long cyc[4];
  cudaMemcpy(cyc, reinterpret_cast<long long*>(d_out + 2 * 100000), sizeof(cyc), cudaMemcpyDeviceToHost);
  printf("body %6d instr (%4d KB): blocks %4d x %3d thr: %.3f ms, %.2f cycles/instr/warp (block 0)\n", N, N * 16 / 1024, blocks, threads, ms,
         double(cyc[0]) / (double(iters) * N));
}

int main() {
  int* d_out; cudaMalloc(&d_out, 4 * 300000);
  for (int cfg = 0; cfg < 2; ++cfg) {
    const int blocks = cfg == 0 ? 148 : (cfg == 1 ? 148 * 4 : 8 * 4), threads = 128;
    printf("-- %d blocks of %d threads\n", blocks, threads);
    run<768>(d_out, blocks, threads);
    run<1024>(d_out, blocks, threads);
    run<1152>(d_out, blocks, threads);
    run<1280>(d_out, blocks, threads);
    run<1408>(d_out, blocks, threads);
    run<1536>(d_out, blocks, threads);
    run<1792>(d_out, blocks, threads);
    run<2048>(d_out, blocks, threads);
    run<3072>(d_out, blocks, threads);
    run<4096>(d_out, blocks, threads);
    run<6144>(d_out, blocks, threads);
    run<8192>(d_out, blocks, threads);
  }
  return 0;
}
