// Microbenchmark: issue rate of tcgen05.mma shapes used by the prefix kernel (operands in place, no softmax).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_rate umma_rate.cu && ./umma_rate
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra WAIT_DONE;\nbra WAIT_LOOP;\nWAIT_DONE:\n}\n" ::"r"(
          smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_ss(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %6, 0;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n}\n" ::"r"(d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 db;\nsetp.ne.b32 p, %5, 0;\nmov.b64 db, {%2, %3};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n}\n" ::"r"(d),
      "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr, uint32_t lbo) { return ((addr >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16); }
__host__ __device__ constexpr uint32_t desc_hi(uint32_t sbo) { return ((sbo >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29); }
__host__ __device__ constexpr uint32_t make_idesc(int fmt, int b_mn, int m, int n) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.b32 %0, 1, 0, P;\n}\n" : "=r"(pred));
  return pred != 0;
}

// mode: 0 SS n128 | 1 SS n64 | 2 TS n128 B MN-major | 3 SS n256 | 4 v3 block (8 SS n64 + 4 TS) | 5 v2 block (8 SS n128 + 8 TS)
//       6 TS n128 with K-major B | 7 SS n128 B MN-major
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(int mode, int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  // zero the operands (denormal / NaN payloads do not change the rate, but keep it clean)
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (warp == 1 && elect_one()) {
    constexpr uint32_t hi = desc_hi(1024);
    const uint32_t a_lo = desc_lo(smem_u32(smem), 0);                 // A tile: 128 rows x 128 cols (2 atoms of 16 KB)
    const uint32_t b_lo = desc_lo(smem_u32(smem + 64 * 1024), 0);     // B tile K-major
    const uint32_t bmn_lo = desc_lo(smem_u32(smem + 64 * 1024), 16384);  // B tile MN-major: LBO = one 64-col half (128 rows)
    const uint32_t bmn64_lo = desc_lo(smem_u32(smem + 64 * 1024), 8192); // 64-row V block: halves 8 KB apart
    const uint32_t i128 = make_idesc(1, 0, 128, 128), i64 = make_idesc(1, 0, 128, 64), i256 = make_idesc(1, 0, 128, 256);
    const uint32_t ipv = make_idesc(1, 1, 128, 128);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (mode == 0) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) { uint32_t off = ((kk / 4) * 16384 + (kk % 4) * 32) >> 4; umma_ss(tmem, a_lo + off, hi, b_lo + off, hi, i128, kk > 0); }
      } else if (mode == 1) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) { uint32_t qo = ((kk / 4) * 16384 + (kk % 4) * 32) >> 4, ko = ((kk / 4) * 8192 + (kk % 4) * 32) >> 4; umma_ss(tmem, a_lo + qo, hi, b_lo + ko, hi, i64, kk > 0); }
      } else if (mode == 2) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) umma_ts(tmem + 256, tmem + kk * 8, bmn_lo + kk * (2048 >> 4), hi, ipv, 1);
      } else if (mode == 3) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) { uint32_t off = ((kk / 4) * 16384 + (kk % 4) * 32) >> 4; umma_ss(tmem, a_lo + off, hi, b_lo + ((kk / 4) * 32768 + (kk % 4) * 32 >> 4), hi, i256, kk > 0); }
      } else if (mode == 4) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_ts(tmem + 256, tmem + 64 + kk * 8, bmn64_lo + kk * (2048 >> 4), hi, ipv, 1);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) { uint32_t qo = ((kk / 4) * 16384 + (kk % 4) * 32) >> 4, ko = ((kk / 4) * 8192 + (kk % 4) * 32) >> 4; umma_ss(tmem, a_lo + qo, hi, b_lo + ko, hi, i64, kk > 0); }
      } else if (mode == 5) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) umma_ts(tmem + 256, tmem + 128 + kk * 8, bmn_lo + kk * (2048 >> 4), hi, ipv, 1);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) { uint32_t off = ((kk / 4) * 16384 + (kk % 4) * 32) >> 4; umma_ss(tmem, a_lo + off, hi, b_lo + off, hi, i128, kk > 0); }
      } else if (mode == 6) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) { uint32_t off = ((kk / 4) * 16384 + (kk % 4) * 32) >> 4; umma_ts(tmem + 256, tmem + kk * 8, b_lo + off, hi, i128, 1); }
      } else {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) { uint32_t off = ((kk / 4) * 16384 + (kk % 4) * 32) >> 4; umma_ss(tmem + 256, a_lo + off, hi, bmn_lo + kk * (2048 >> 4), hi, ipv, 1); }
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
  int sms = 148;
  long long* d_out;
  cudaMalloc(&d_out, sms * sizeof(long long));
  const int smem_bytes = 160 * 1024 + 1024;
  cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  const char* names[] = {"SS 128x128x16 (QK v2)", "SS 128x64x16 (QK v3)", "TS 128x128x16 B MN-major (PV)", "SS 128x256x16", "v3 block: 4 TS + 8 SS n64",
                         "v2 block: 8 TS + 8 SS n128", "TS 128x128x16 B K-major", "SS 128x128x16 B MN-major"};
  const int per_rep[] = {8, 8, 8, 8, 12, 16, 8, 8};
  const double mac_per_rep[] = {8 * 128. * 128 * 16, 8 * 128. * 64 * 16, 8 * 128. * 128 * 16, 8 * 128. * 256 * 16, 4 * 128. * 128 * 16 + 8 * 128. * 64 * 16,
                                16 * 128. * 128 * 16, 8 * 128. * 128 * 16, 8 * 128. * 128 * 16};
  for (int grid : {1, 148}) {
    for (int mode = 0; mode < 8; ++mode) {
      const int reps = 2000;
      for (int it = 0; it < 2; ++it) {
        umma_rate_kernel<<<grid, 128, smem_bytes>>>(mode, reps, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
      }
      std::vector<long long> h(grid);
      cudaMemcpy(h.data(), d_out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      long long mx = 0, mn = 1LL << 62;
      for (auto v : h) { mx = v > mx ? v : mx; mn = v < mn ? v : mn; }
      printf("grid %3d mode %d %-34s cycles/MMA min %.1f max %.1f   MAC/clk/SM %.0f\n", grid, mode, names[mode], (double)mn / (reps * per_rep[mode]),
             (double)mx / (reps * per_rep[mode]), mac_per_rep[mode] * reps / (double)mx);
    }
  }
  return 0;
}
