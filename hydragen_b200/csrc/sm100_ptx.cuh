// sm_100a PTX wrappers shared by the two shared-prefix attention kernels (prefix_unit_sm100.cu: one CTA per unit;
// prefix_sm100.cu: persistent, all shared levels in one launch): mbarriers, TMA, tcgen05.mma / ld / st / commit,
// UMMA descriptors, packed fp32x2 arithmetic.
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace hg {
namespace {

// ---- PTX wrappers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// One lane of the (converged) warp, known to the compiler as such: the uniform datapath can then
// feed TMA / UMMA descriptors without a per-lane waterfall loop.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.b32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

// tcgen05.commit: the mbarrier gets one arrival when every MMA issued so far by this thread is done.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]; descriptors passed as (lo, hi) halves so that stepping the start
// address along K is one 32-bit add per operand
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "setp.ne.b32 p, %6, 0;\n"
      "mov.b64 da, {%1, %2};\n"
      "mov.b64 db, {%3, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 db;\n"
      "setp.ne.b32 p, %5, 0;\n"
      "mov.b64 db, {%2, %3};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Shared-memory matrix descriptor (sm_100 UMMA): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) |
// version=1 [46,48) | layout [61,64) with SWIZZLE_128B = 2.  lo = start | LBO, hi = SBO | version | layout.
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__host__ __device__ constexpr uint32_t desc_hi(uint32_t sbo_bytes) {
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
}

// Instruction descriptor, kind::f16: D fmt [4,6) (1 = f32) | A fmt [7,10) | B fmt [10,13) (0 = f16, 1 = bf16) |
// A major bit 15 | B major bit 16 (0 = K-major, 1 = MN-major) | N>>3 [17,23) | M>>4 [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int fmt, int b_mn_major, int m, int n) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

#define HG_R32(a, o)                                                                                                    \
  "=r"(a[o + 0]), "=r"(a[o + 1]), "=r"(a[o + 2]), "=r"(a[o + 3]), "=r"(a[o + 4]), "=r"(a[o + 5]), "=r"(a[o + 6]),        \
      "=r"(a[o + 7]), "=r"(a[o + 8]), "=r"(a[o + 9]), "=r"(a[o + 10]), "=r"(a[o + 11]), "=r"(a[o + 12]), "=r"(a[o + 13]), \
      "=r"(a[o + 14]), "=r"(a[o + 15]), "=r"(a[o + 16]), "=r"(a[o + 17]), "=r"(a[o + 18]), "=r"(a[o + 19]),               \
      "=r"(a[o + 20]), "=r"(a[o + 21]), "=r"(a[o + 22]), "=r"(a[o + 23]), "=r"(a[o + 24]), "=r"(a[o + 25]),               \
      "=r"(a[o + 26]), "=r"(a[o + 27]), "=r"(a[o + 28]), "=r"(a[o + 29]), "=r"(a[o + 30]), "=r"(a[o + 31])
#define HG_W16(a, o)                                                                                                     \
  "r"(a[o + 0]), "r"(a[o + 1]), "r"(a[o + 2]), "r"(a[o + 3]), "r"(a[o + 4]), "r"(a[o + 5]), "r"(a[o + 6]), "r"(a[o + 7]), \
      "r"(a[o + 8]), "r"(a[o + 9]), "r"(a[o + 10]), "r"(a[o + 11]), "r"(a[o + 12]), "r"(a[o + 13]), "r"(a[o + 14]),       \
      "r"(a[o + 15])
#define HG_W32(a, o) HG_W16(a, o), HG_W16(a, o + 16)

// 32 lanes x 32 consecutive 32-bit columns: thread t of the warp gets lane (warp%4)*32 + t.
#define HG_TMEM_LD32(taddr, a, o)                                                                               \
  asm volatile(                                                                                                 \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                 \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27," \
      "%28,%29,%30,%31}, [%32];"                                                                                \
      : HG_R32(a, o)                                                                                            \
      : "r"(taddr))
#define HG_TMEM_ST32(taddr, a, o)                                                                               \
  asm volatile(                                                                                                 \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "                                                          \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27," \
      "%28,%29,%30,%31};" ::HG_W32(a, o),                                                                       \
      "r"(taddr)                                                                                                \
      : "memory")
#define HG_TMEM_ST16(taddr, a, o)                                                                          \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};" ::HG_W16(a, o), \
               "r"(taddr)                                                                                  \
               : "memory")

__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <typename T>
__device__ __forceinline__ uint32_t pack2(float a, float b);
template <>
__device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));  // first source -> upper half
  return r;
}
template <>
__device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

// ---- packed fp32x2 arithmetic (sm_100: one issue slot for two lanes' worth of FMA-pipe work) -------------
__device__ __forceinline__ uint64_t pack_f2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack_f2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// smem tile (generic-proxy writes fenced by the caller) -> global through the tensor map
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(src)),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
template <int ID, int THREADS>
__device__ __forceinline__ void named_bar_sync() {
  asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(THREADS) : "memory");
}

__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// partials written by other SMs during this launch: read them at L2
__device__ __forceinline__ float4 ld_cg_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float2 ld_cg_f2(const float* p) {
  float2 r;
  asm volatile("ld.global.cg.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}


}  // namespace
}  // namespace hg
