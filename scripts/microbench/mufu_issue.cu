// Microbenchmark: how MUFU.EX2 shares the SM sub-partition's issue port with FMA-pipe work, for the
// softmax of the prefix kernel.  Per SMSP, W warps (1 or 2) run a loop of [1 ex2 + K independent FFMAs];
// reports cycles per ex2 instruction per SMSP.  If a MUFU blocks the issue port for its 8 cycles the cost
// is 8 + K; if it only occupies the XU pipe it is max(8, K + 1).  Also: ex2.approx.ftz.bf16x2 (two results
// per lane per instruction) and the F2FP pack.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mufu_issue mufu_issue.cu && ./mufu_issue
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

template <int K, int MODE>  // MODE 0: ex2.f32   1: ex2.bf16x2   2: ex2.f32 x2 + F2FP pack (current softmax inner step)
__global__ void __launch_bounds__(256, 1) mufu_kernel(int reps, float seed, long long* cycles, float* sink) {
  float x[8], f[8];
  uint32_t xb[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] = seed * (float)(threadIdx.x + i) * 1e-3f - 1.f;
    f[i] = seed + (float)i;
    xb[i] = 0xbf80bf80u + (uint32_t)i;  // two bf16 values near -1
  }
  const float a = 0.999f, b = 1e-4f;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      } else if (MODE == 1) {
        asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(xb[i]));
      } else {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(xb[i]) : "f"(x[i]), "f"(f[i]));
      }
#pragma unroll
      for (int k = 0; k < K; ++k) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[(i + k) & 7]) : "f"(a), "f"(b));
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i] + f[i] + __uint_as_float(xb[i]);
  if (s == 12345.678f) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int K, int MODE>
void run(int warps_per_smsp, long long* d_cyc, float* d_sink) {
  const int reps = 2000;
  mufu_kernel<K, MODE><<<148, 128 * warps_per_smsp>>>(reps, 0.5f, d_cyc, d_sink);
  mufu_kernel<K, MODE><<<148, 128 * warps_per_smsp>>>(reps, 0.5f, d_cyc, d_sink);
  long long c = 0;
  cudaMemcpy(&c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost);
  const double per = (double)c / (reps * 8.0) / warps_per_smsp;  // cycles per loop body instance per SMSP
  printf("mode %d  K=%2d  warps/SMSP=%d : %.2f cycles per [%s + %d FFMA] per SMSP (per warp: %.2f)\n", MODE, K, warps_per_smsp, per,
         MODE == 0 ? "ex2.f32" : (MODE == 1 ? "ex2.bf16x2" : "2 ex2.f32 + cvt.bf16x2"), K, per * warps_per_smsp);
}

template <int MODE>
void sweep(long long* d_cyc, float* d_sink) {
  for (int w = 1; w <= 2; ++w) {
    run<0, MODE>(w, d_cyc, d_sink);
    run<1, MODE>(w, d_cyc, d_sink);
    run<2, MODE>(w, d_cyc, d_sink);
    run<4, MODE>(w, d_cyc, d_sink);
    run<6, MODE>(w, d_cyc, d_sink);
    run<8, MODE>(w, d_cyc, d_sink);
    run<12, MODE>(w, d_cyc, d_sink);
    run<16, MODE>(w, d_cyc, d_sink);
  }
}

int main() {
  long long* d_cyc;
  float* d_sink;
  cudaMalloc(&d_cyc, 8);
  cudaMalloc(&d_sink, 4);
  sweep<0>(d_cyc, d_sink);
  sweep<1>(d_cyc, d_sink);
  sweep<2>(d_cyc, d_sink);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
