// Shared-prefix attention on the Blackwell tensor cores: tcgen05.mma + TMEM + TMA (sm_100a).
//
// Replaces the reference's prefix branch -- hydragen/attention.py:261-338 calling
// flash_attention / flash_attention_varlen (hydragen/flash.py:284-351), i.e. flash-attn v2.3.6's
// mma.sync (Ampere) kernel, plus the LSE transposes of attention.py:276-280,333-338.
//
// Inter-sequence batching makes this a dense problem: for one (group, head) the queries of every
// sequence sharing the prefix form Q[q_per_group x d] and are multiplied against the single
// K,V[k_len x d] of that prefix.  One CTA owns TWO 128-row Q tiles (A, B) of one head and streams
// the prefix in 64-key blocks; every K/V block fetched feeds 256 query rows, and the score block of
// each tile is double buffered in TMEM so that Q K^T runs two blocks ahead of the softmax:
//
//   TMEM (512 columns)  S_A[0] S_A[1] S_B[0] S_B[1] (64 fp32 columns each) | O_A | O_B (128 each);
//                       P_t(j) (16-bit) is written back over the first 32 columns of its S buffer
//   warp 0 (1 lane)  TMA producer: Q_A, Q_B once, then a 4-deep ring whose slot u holds what MMA
//                    iteration u consumes: V_u and K_{u+2} (cp.async.bulk.tensor, SWIZZLE_128B boxes)
//   warps 1, 3       MMA issuer of tile A / B (all lanes walk the loop so descriptors stay in uniform
//                    registers; one elected lane issues).  Per key block j:  PV_t(j)  QK_t(j+2)
//                      S_t = Q_t K_j^T  (SS form, both operands K-major in smem, 128x64x16 per
//                                        instruction, fp32 accumulate in TMEM)
//                      O_t += P_t V_j   (TS form: P_t read from TMEM as the A operand, V_j straight
//                                        from its row-major smem tile as an MN-major B operand --
//                                        no transpose pass)
//   warp 2           TMEM allocator
//   warps 4-7        softmax of tile A, warps 8-11 softmax of tile B: thread t owns row t
//                    (tcgen05.ld 32x32b: lane == row, so the row max / row sum need no shuffles);
//                    software pipelined: the scores of block j+1 are fetched and reduced to their
//                    row max behind the MUFU exp2 requests of block j; scale / subtract / row sums as
//                    packed fp32x2; P_t stored to TMEM as packed 16-bit; lazy rescale of O_t (only
//                    when the running max grows by more than 2^8); epilogue O_t / l -> swizzled
//                    smem (the dead Q_t tile) -> TMA store; LSE written directly in [b, nq, hq].
//                    setmaxnreg: 56 registers for warps 0-3, 224 for the softmax warps (no spills in the loop)
//
// All producer/consumer edges are mbarriers (TMA complete_tx, tcgen05.commit, thread arrives); there
// is no __syncthreads in the main loop.  Split-KV (kv_splits > 1): the CTAs of one tile each take a
// contiguous range of key blocks and write their own partial (out, lse).
//
// Instantiations (one translation unit each, compiled in parallel: this file is also #included by
// prefix_sm100_causal.cu and prefix_sm100_split.cu):
//   <T, D, kCausal = false, kSplit = 0>  the decode hot path (hg_prefix_attn_fwd / _split_fwd)
//   <T, D, kCausal = true,  kSplit = 0>  prefill: bottom-right aligned causal mask inside every group
//                                        (hg_causal_attn_fwd): row tiles heavy-first, only the visible key
//                                        blocks are streamed, masks only in the blocks crossing the diagonal
//   <T, D, kCausal = false, kSplit = 1>  experimental split-column softmax (two warpgroups per tile); measured
//                                        slower, selected only by HYDRAGEN_B200_PREFIX_SOFTMAX=split
//   <T, D, kCausal = false, kSplit = 2>  experimental non-pipelined softmax loop (prefix_sm100_simple.cu),
//                                        HYDRAGEN_B200_PREFIX_SOFTMAX=simple
//   <T, D, kCausal = false, kSplit = 3>  alternate-block softmax: four softmax warps per sub-partition without a
//                                        per-block exchange (prefix_sm100_alt.cu, HYDRAGEN_B200_PREFIX_SOFTMAX=alt);
//                                        written after the GPU budget of round 1 was spent: NOT yet run on hardware
//
// Algorithmic work per CTA: 4 * rows * k_len * d FLOP.  Bound: tensor pipe (AI ~ 680 FLOP/B at the
// 7B config), with the MUFU exp2 rate (16/clk/SM == the 128x128x128 MMA rate) the co-limiter.
#include <cuda.h>

#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace hg {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 64;   // keys per block: S_t block = 64 TMEM columns, double buffered
constexpr int kTiles = 2;  // Q tiles per CTA (ping-pong)
constexpr int kThreads = 384;       // base: TMA / MMA warpgroup + one softmax warpgroup per tile
constexpr int kThreadsSplit = 640;  // split-column softmax: two softmax warpgroups per tile
constexpr uint32_t kTmemCols = 512;
__host__ __device__ constexpr uint32_t tmem_s(int t, int b) { return (uint32_t)t * 128u + (uint32_t)b * 64u; }  // S_t buffer b (P aliases its first 32 columns)
__host__ __device__ constexpr uint32_t tmem_o(int t) { return 256u + (uint32_t)t * 128u; }                       // O_t
constexpr float kRescaleThreshold = 8.0f;  // log2 units
#ifndef HG_PREFIX_SOFTMAX_DEFAULT
#define HG_PREFIX_SOFTMAX_DEFAULT 0  // softmax organisation used when HYDRAGEN_B200_PREFIX_SOFTMAX is not set (0 base)
#endif
#ifndef HG_PREFIX_BDELAY_DEFAULT
#define HG_PREFIX_BDELAY_DEFAULT 700  // cycles tile B's softmax starts after tile A's (0: together); long prefixes only
#endif
#ifndef HG_PREFIX_EMU_EVERY
#define HG_PREFIX_EMU_EVERY 0
#endif
// every n-th pair of exponentials is computed on the FMA pipes instead of MUFU (0: none).  Measured at cfg#2
// (B200, graph-timed): 0 -> 35.1 us, 4 -> 36.1, 3 -> 36.0, 2 -> 36.2: the warps are issue-bound, not MUFU-bound.
constexpr int kEmuEvery = HG_PREFIX_EMU_EVERY;

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
#define HG_W32(a, o)                                                                                                     \
  "r"(a[o + 0]), "r"(a[o + 1]), "r"(a[o + 2]), "r"(a[o + 3]), "r"(a[o + 4]), "r"(a[o + 5]), "r"(a[o + 6]), "r"(a[o + 7]), \
      "r"(a[o + 8]), "r"(a[o + 9]), "r"(a[o + 10]), "r"(a[o + 11]), "r"(a[o + 12]), "r"(a[o + 13]), "r"(a[o + 14]),       \
      "r"(a[o + 15]), "r"(a[o + 16]), "r"(a[o + 17]), "r"(a[o + 18]), "r"(a[o + 19]), "r"(a[o + 20]), "r"(a[o + 21]),     \
      "r"(a[o + 22]), "r"(a[o + 23]), "r"(a[o + 24]), "r"(a[o + 25]), "r"(a[o + 26]), "r"(a[o + 27]), "r"(a[o + 28]),     \
      "r"(a[o + 29]), "r"(a[o + 30]), "r"(a[o + 31])

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
__device__ __forceinline__ uint64_t fadd2_rm(uint64_t a, uint64_t b) {  // round toward -inf
  uint64_t d;
  asm("add.rm.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t fsub2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// 2^x for a pair of fp32 on the FMA / ALU pipes instead of the MUFU unit (Cody-Waite split + degree-3
// minimax polynomial, max relative error 8.8e-5 -- below the rounding of the 16-bit P it feeds):
//   r = RM(x + 1.5*2^23) keeps floor(x) in its low mantissa bits, f = x - floor(x) in [0, 1),
//   2^x = p(f) * 2^floor(x): the integer part is added straight into the exponent field of p(f).
// Inputs are clamped at -127 (2^x underflows there anyway; without it the exponent field would wrap).
__device__ __forceinline__ void exp2_poly_x2(float x0, float x1, float& p0, float& p1) {
  const uint64_t kMagic = pack_f2(12582912.f, 12582912.f);
  const uint64_t kC3 = pack_f2(0.077119089663028717041015625f, 0.077119089663028717041015625f);
  const uint64_t kC2 = pack_f2(0.227564394474029541015625f, 0.227564394474029541015625f);
  const uint64_t kC1 = pack_f2(0.695146143436431884765625f, 0.695146143436431884765625f);
  const uint64_t kOne = pack_f2(1.f, 1.f);
  const uint64_t x = pack_f2(fmaxf(x0, -127.f), fmaxf(x1, -127.f));
  const uint64_t r = fadd2_rm(x, kMagic);
  const uint64_t f = fsub2(x, fsub2(r, kMagic));
  uint64_t p = ffma2(kC3, f, kC2);
  p = ffma2(p, f, kC1);
  p = ffma2(p, f, kOne);
  float r0, r1, q0, q1;
  unpack_f2(r, r0, r1);
  unpack_f2(p, q0, q1);
  p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(r0) << 23));
  p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(r1) << 23));
}

#define HG_W16(a, o)                                                                                                     \
  "r"(a[o + 0]), "r"(a[o + 1]), "r"(a[o + 2]), "r"(a[o + 3]), "r"(a[o + 4]), "r"(a[o + 5]), "r"(a[o + 6]), "r"(a[o + 7]), \
      "r"(a[o + 8]), "r"(a[o + 9]), "r"(a[o + 10]), "r"(a[o + 11]), "r"(a[o + 12]), "r"(a[o + 13]), "r"(a[o + 14]),       \
      "r"(a[o + 15])
#define HG_TMEM_ST16(taddr, a, o)                                                                          \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};" ::HG_W16(a, o), \
               "r"(taddr)                                                                                  \
               : "memory")

// smem tile (generic-proxy writes fenced by the caller) -> global through the tensor map
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(src)),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_and_wait() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
template <int ID, int THREADS>
__device__ __forceinline__ void named_bar_sync() {
  asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(THREADS) : "memory");
}
__device__ __forceinline__ void named_bar_sync_rt(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

constexpr int kStages = 4;  // K/V ring depth

// Ring slot u (u = -2 .. n_blocks-1) holds what MMA iteration u consumes: V_u (for P V of block u)
// and K_{u+2} (for Q K^T of block u+2, issued in the same iteration); slots -2 and -1 carry only
// K_0 / K_1 for the prologue.  One full and one empty barrier per slot.
template <int D>
struct SmemLayout {
  static constexpr int kHalves = D / 64;                    // 64-element (128-byte) swizzle atoms along d
  static constexpr int kQTileBytes = BLOCK_M * D * 2;       // one Q tile (also one output staging tile)
  static constexpr int kQHalfBytes = BLOCK_M * 64 * 2;      // one Q TMA box: 128 rows x 128 B
  static constexpr int kKVTileBytes = BLOCK_N * D * 2;      // one K / V block
  static constexpr int kKVHalfBytes = BLOCK_N * 64 * 2;     // one K/V TMA box: 64 rows x 128 B
  static constexpr int kStageBytes = 2 * kKVTileBytes;      // V block then K block
  static constexpr int kQ = 0;                              // 2 tiles (A, B)
  static constexpr int kKV = kQTileBytes * kTiles;
  static constexpr int kBars = kKV + kStageBytes * kStages;
  static constexpr int kXchg = kBars + 512;  // split-column softmax: fp32 [parity][tile][half][row] row-max / row-sum exchange
  static constexpr int kTotal = kXchg + 2 * kTiles * 2 * BLOCK_M * 4;
};

struct Barriers {
  uint64_t q_full[kTiles];
  uint64_t kv_full[kStages], kv_empty[kStages];
  uint64_t s_full[kTiles][2], p_full[kTiles][2];  // per S buffer
  uint64_t pv_done[kTiles];                       // one phase per PV_t(j) (lazy-rescale path only)
  uint64_t o_full[kTiles];                        // O_t complete
  uint32_t tmem_base;
  uint32_t pad_;
  uint64_t m_ready[kTiles][4][2];  // alternate-block softmax only: reference max of a block published (per lane quarter, block parity)
};

}  // namespace

#if defined(HG_PREFIX_TRACE) && !defined(HG_PREFIX_TU_CAUSAL) && !defined(HG_PREFIX_TU_SPLIT) && !defined(HG_PREFIX_TU_SIMPLE) && !defined(HG_PREFIX_TU_ALT)
// Development aid (never compiled into the shipped library): clock64 stamps of CTA (0,0).
// Layout: [role][block j][slot]; role 0 = MMA thread, 1 = softmax warp of tile A, 2 = tile B.
__device__ long long g_trace[3 * 64 * 8];
#define HG_TRACE(role, j, slot)                                                                              \
  do {                                                                                                       \
    if (blockIdx.x == 0 && blockIdx.y == 0 && (j) < 64) g_trace[((role) * 64 + (j)) * 8 + (slot)] = clock64(); \
  } while (0)
#else
#define HG_TRACE(role, j, slot) \
  do {                          \
  } while (0)
#endif

// kCausal: bottom-right aligned causal mask inside every group (the prefill form, flash_attention(causal=True) of
// hydragen/flash.py:284-306); a separate instantiation so that the decode-path kernel is exactly the unmasked code.
//
// kSplit: softmax organisation.  0 = one warpgroup per tile (thread = one row x 64 keys of a block, software
// pipelined).  1 = TWO warpgroups per tile, each thread one row x 32 keys: four softmax warps per SM sub-partition
// instead of two hide each other's TMEM / MUFU / barrier latencies (r01e: the two-warp form keeps the MUFU unit
// only ~55 % busy); the two half-row maxima meet through shared memory and a 64-thread named barrier per block.
template <typename T, int D, bool kCausal, int kSplit>
__global__ void __launch_bounds__((kSplit == 1 || kSplit == 3) ? kThreadsSplit : kThreads, 1)
    prefix_attn_sm100_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                             const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_o,
                             T* __restrict__ out, float* __restrict__ lse, const int32_t* __restrict__ cu_seqlens_k,
                             int q_per_group, int tiles_per_group, int k_len_uniform, int hq, int hkv, float scale_log2,
                             int kv_splits, int n_q_rows, int b_delay) {
  using L = SmemLayout<D>;
  constexpr int kFmt = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;
  constexpr uint32_t kIdescQK = make_idesc(kFmt, 0, BLOCK_M, BLOCK_N);
  constexpr uint32_t kIdescPV = make_idesc(kFmt, 1, BLOCK_M, D);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  Barriers* bars = reinterpret_cast<Barriers*>(smem + L::kBars);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // blockIdx.x = (group, m-tile, kv split): the splits of one tile sit next to each other
  const int split = blockIdx.x % kv_splits;
  const int tile = blockIdx.x / kv_splits, head = blockIdx.y;
  const int grp = tile / tiles_per_group;
  // causal: later row tiles see more keys -- launch them first
  const int mt = kCausal ? tiles_per_group - 1 - tile % tiles_per_group : tile % tiles_per_group;
  const int kvh = head / (hq / hkv);
  const int q_row0 = grp * q_per_group + mt * (kTiles * BLOCK_M);
  const int rows_left = q_per_group - mt * (kTiles * BLOCK_M);  // > 0
  const bool two = rows_left > BLOCK_M;                         // tile B holds valid rows
  int k_start, k_len;
  if (cu_seqlens_k != nullptr) {
    k_start = __ldg(cu_seqlens_k + grp);
    k_len = __ldg(cu_seqlens_k + grp + 1) - k_start;
  } else {
    k_start = grp * k_len_uniform;
    k_len = k_len_uniform;
  }
  if (kv_splits > 1) {
    // split-KV: this CTA owns key blocks [split * bps, (split + 1) * bps) of its group and writes partial
    // result number `split` (rows [split * n_q_rows, ...) of out / lse); the merge is the caller's combine.
    const int bps = ((k_len + BLOCK_N - 1) / BLOCK_N + kv_splits - 1) / kv_splits;
    const int first = split * bps * BLOCK_N;
    k_start += first;
    k_len = max(0, min(k_len - first, bps * BLOCK_N));
    out += (int64_t)split * n_q_rows * hq * D;
    if (lse != nullptr) lse += (int64_t)split * n_q_rows * hq;
  }
  // causal (bottom-right aligned, flash-attn >= 2.1): row r of the group sees keys j <= r + causal_off.  The CTA
  // streams only the keys its last row can see; the rows above it are masked per element in the diagonal blocks.
  int causal_off = 0;
  if (kCausal) {
    causal_off = k_len - q_per_group;  // >= 0 (checked by the launcher)
    const int last_row = min(q_per_group, (mt + 1) * (kTiles * BLOCK_M)) - 1;
    k_len = min(k_len, last_row + causal_off + 1);
  }
  const int n_blocks = (k_len + BLOCK_N - 1) / BLOCK_N;

  if (n_blocks == 0) {  // empty prefix: out = 0, lse = -inf (uniform branch for the whole CTA)
    const int rows = min(kTiles * BLOCK_M, rows_left);
    for (int idx = threadIdx.x; idx < rows * (D / 8); idx += blockDim.x) {
      const int r = idx / (D / 8), c = idx % (D / 8);
      st_v4(out + ((int64_t)(q_row0 + r) * hq + head) * D + c * 8, make_uint4(0, 0, 0, 0));
    }
    if (lse != nullptr)
      for (int r = threadIdx.x; r < rows; r += blockDim.x) lse[(int64_t)(q_row0 + r) * hq + head] = -INFINITY;
    return;
  }

  // ---- one-time setup --------------------------------------------------------------------
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_v) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_o) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kTiles; ++i) {
      mbar_init(&bars->q_full[i], 1);
      mbar_init(&bars->pv_done[i], 1);
      mbar_init(&bars->o_full[i], 1);
      if constexpr (kSplit == 3) {
        for (int q = 0; q < 4; ++q)
          for (int b = 0; b < 2; ++b) mbar_init(&bars->m_ready[i][q][b], 32);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&bars->s_full[i][b], 1);
        mbar_init(&bars->p_full[i][b], BLOCK_M * (kSplit == 1 ? 2 : 1));
      }
    }
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&bars->kv_full[i], 1);
      mbar_init(&bars->kv_empty[i], two ? 2 : 1);  // one tcgen05.commit per MMA warp
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&bars->tmem_base);
  // The launch that follows on the stream (the fused append / suffix / combine kernel) is a programmatic
  // dependent: let it start on SMs this grid leaves idle; it waits for this grid's completion itself
  // before it reads the partial results written here.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // This launch is itself a programmatic dependent of whatever precedes it on the stream (the previous
  // layer's decode kernel, or the projection that produced q): everything above -- barrier init, TMEM
  // allocation, descriptor prefetch -- overlapped its tail; q is read and out / lse are written only
  // from here on.  (No-op when launched without the attribute.)
  asm volatile("griddepcontrol.wait;" ::: "memory");

  // Register budget (setmaxnreg must sit inside the role branch it applies to): the producer
  // warpgroup gives registers back, the two softmax warpgroups (128 live fp32 scores per thread)
  // take them: 128 x 56 + 256 x 224 = 384 x 168, the launch-time allocation (r01h: with 88 / 208 the running max, row
  // sum and loop state of the softmax threads were spilled to local memory, on the serial path between two blocks).
  // (split-column form: 640 threads x 96 at launch -> 128 x 64 + 512 x 104.)
  if (warp < 4) {
    if constexpr (kSplit == 1 || kSplit == 3) asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // =============================== TMA producer ===========================================
    if (elect_one()) {
      for (int t = 0; t < (two ? 2 : 1); ++t) {
        mbar_expect_tx(&bars->q_full[t], L::kQTileBytes);
#pragma unroll
        for (int h = 0; h < L::kHalves; ++h)
          tma_load_2d(smem + L::kQ + t * L::kQTileBytes + h * L::kQHalfBytes, &tmap_q, head * D + h * 64, q_row0 + t * BLOCK_M,
                      &bars->q_full[t]);
      }
      for (int u = -2; u < n_blocks; ++u) {
        const int st = (u + 2) % kStages;
        const bool has_v = u >= 0, has_k = u + 2 < n_blocks;
        if (!has_v && !has_k) continue;
        uint8_t* base = smem + L::kKV + st * L::kStageBytes;
        mbar_wait(&bars->kv_empty[st], (((u + 2) / kStages) & 1) ^ 1);
        mbar_expect_tx(&bars->kv_full[st], (has_v ? L::kKVTileBytes : 0) + (has_k ? L::kKVTileBytes : 0));
        if (has_k) {
#pragma unroll
          for (int h = 0; h < L::kHalves; ++h)
            tma_load_2d(base + L::kKVTileBytes + h * L::kKVHalfBytes, &tmap_k, kvh * D + h * 64, k_start + (u + 2) * BLOCK_N,
                        &bars->kv_full[st]);
        }
        if (has_v) {
#pragma unroll
          for (int h = 0; h < L::kHalves; ++h)
            tma_load_2d(base + h * L::kKVHalfBytes, &tmap_v, kvh * D + h * 64, k_start + u * BLOCK_N, &bars->kv_full[st]);
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // =============================== MMA issuers (warp 1: tile A, warp 3: tile B) ==============
    // The whole warp walks the loop and the barriers (warp-uniform, so descriptors stay in uniform
    // registers); one elected lane issues tcgen05.mma / tcgen05.commit.  Per block j and tile t:
    //   P V of block j, then Q K^T of block j+2 into the S buffer P_t(j) just vacated (same thread,
    //   same issue order: no barrier needed between them).
    const int t = warp >> 1;
    if (t == 0 || two) {
      const bool leader = elect_one();
      constexpr uint32_t kHiK = desc_hi(1024);  // SWIZZLE_128B: 8-row groups 1024 B apart
      const uint32_t q_lo = desc_lo(smem_u32(smem + L::kQ + t * L::kQTileBytes), 0);
      const uint32_t v_lo0 = desc_lo(smem_u32(smem + L::kKV), L::kKVHalfBytes);
      const uint32_t k_lo0 = desc_lo(smem_u32(smem + L::kKV + L::kKVTileBytes), 0);
      const uint32_t o_tmem = tmem + tmem_o(t);
      // S_t = Q_t K^T: D/16 instructions of 128x64x16; operand k-slice kk lives in swizzle atom kk/4
      // at byte offset (kk%4)*32 inside the 128-byte row (start-address field is in 16-byte units).
      auto issue_qk = [&](int st, int sbuf) {
        const uint32_t k_lo = k_lo0 + st * (L::kStageBytes >> 4);
        const uint32_t d_tmem = tmem + tmem_s(t, sbuf);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t q_off = ((kk / 4) * L::kQHalfBytes + (kk % 4) * 32) >> 4;
          const uint32_t k_off = ((kk / 4) * L::kKVHalfBytes + (kk % 4) * 32) >> 4;
          umma_ss(d_tmem, q_lo + q_off, kHiK, k_lo + k_off, kHiK, kIdescQK, kk > 0 ? 1u : 0u);
        }
      };
      // O_t (+)= P_t V: BLOCK_N/16 instructions of 128xDx16; A = P_t (16-bit, 8 TMEM columns per
      // k-slice), B = V tile rows [kk*16, kk*16+16) as an MN-major operand: 8-row groups 1024 B apart
      // (SBO), 64-element column halves one TMA box apart (LBO).
      auto issue_pv = [&](int st, int sbuf, bool first) {
        const uint32_t v_lo = v_lo0 + st * (L::kStageBytes >> 4);
        const uint32_t p_tmem = tmem + tmem_s(t, sbuf);
#pragma unroll
        for (int kk = 0; kk < BLOCK_N / 16; ++kk)
          umma_ts(o_tmem, p_tmem + kk * 8, v_lo + kk * (2048 >> 4), kHiK, kIdescPV, (first && kk == 0) ? 0u : 1u);
      };

      // prologue: S_t(0), S_t(1) -- the softmax warpgroup then always finds its next block ready
      mbar_wait(&bars->q_full[t], 0);
      for (int u = -2; u < 0; ++u) {
        if (u + 2 < n_blocks) {
          const int st = (u + 2) % kStages;
          mbar_wait(&bars->kv_full[st], 0);
          tc_fence_after();
          if (leader) {
            issue_qk(st, (u + 2) & 1);
            umma_commit(&bars->s_full[t][(u + 2) & 1]);
            umma_commit(&bars->kv_empty[st]);
          }
          __syncwarp();
        }
      }
      for (int j = 0; j < n_blocks; ++j) {
        const int b = j & 1;
        const int st = (j + 2) % kStages;
        const bool more = j + 2 < n_blocks;
        if (t == 0) HG_TRACE(0, j, 0);
        mbar_wait(&bars->kv_full[st], ((j + 2) / kStages) & 1);
        if (t == 0) HG_TRACE(0, j, 1);
        mbar_wait(&bars->p_full[t][b], (j >> 1) & 1);
        if (t == 0) HG_TRACE(0, j, 2);
        tc_fence_after();
        if (leader) {
          issue_pv(st, b, j == 0);
          umma_commit(j + 1 < n_blocks ? &bars->pv_done[t] : &bars->o_full[t]);
          if (more) {
            issue_qk(st, b);
            umma_commit(&bars->s_full[t][b]);
          }
          umma_commit(&bars->kv_empty[st]);
        }
        __syncwarp();
        if (t == 0) HG_TRACE(0, j, 3);
      }
    }
  }
  } else if constexpr (kSplit == 1) {
    // =============================== softmax, split-column form ===============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int sw = warp - 4;          // 0 .. 15
    const int t = sw >> 3;            // tile owned by this pair of warpgroups
    const int half = (sw >> 2) & 1;   // which 32 keys of every 64-key block (and which D/2 columns of O)
    const int rows_valid = min(BLOCK_M, rows_left - t * BLOCK_M);
    if (rows_valid > 0) {
      constexpr int DH = D / 2;
      const int wq = warp & 3;        // == sw % 4: the TMEM lane quarter this warp may access
      const int row = wq * 32 + lane;
      const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
      const uint32_t o_addr = tmem + lane_base + tmem_o(t) + (uint32_t)(half * DH);
      float* xchg = reinterpret_cast<float*>(smem + L::kXchg);
      const int pair_bar = 1 + t * 4 + wq;  // the two warps that own the same 32 rows (named barriers 1..8)
      float m_used = -INFINITY, l = 0.f;
      const int row_end = kCausal ? mt * (kTiles * BLOCK_M) + t * BLOCK_M + row + causal_off + 1 : 0x7fffffff;
      const int tile_end = kCausal ? mt * (kTiles * BLOCK_M) + t * BLOCK_M + causal_off + 1 : 0x7fffffff;
      const bool ragged = (k_len % BLOCK_N) != 0;
      // value of the partner warp (same rows, other key half) for this thread's row; double buffered by parity
      auto exchange = [&](int parity, float mine) {
        float* slot = xchg + ((parity * kTiles + t) * 2) * BLOCK_M;
        slot[half * BLOCK_M + row] = mine;
        tc_fence_before();
        named_bar_sync_rt(pair_bar, 64);
        tc_fence_after();
        return slot[(half ^ 1) * BLOCK_M + row];
      };
      for (int j = 0; j < n_blocks; ++j) {
        const int b = j & 1;
        const uint32_t s_addr = tmem + lane_base + tmem_s(t, b);
        mbar_wait(&bars->s_full[t][b], (j >> 1) & 1);
        tc_fence_after();
        uint32_t sc[32];
        HG_TMEM_LD32(s_addr + half * 32, sc, 0);
        tmem_wait_ld();
        if ((ragged && j + 1 == n_blocks) || (j + 1) * BLOCK_N > tile_end) {  // warp-uniform
          const int rem = min(k_len, row_end) - j * BLOCK_N - half * 32;
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (c >= rem) sc[c] = 0xff800000u;  // -inf
        }
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < 32; c += 8) {
          mx[0] = fmaxf(mx[0], fmaxf(__uint_as_float(sc[c + 0]), __uint_as_float(sc[c + 1])));
          mx[1] = fmaxf(mx[1], fmaxf(__uint_as_float(sc[c + 2]), __uint_as_float(sc[c + 3])));
          mx[2] = fmaxf(mx[2], fmaxf(__uint_as_float(sc[c + 4]), __uint_as_float(sc[c + 5])));
          mx[3] = fmaxf(mx[3], fmaxf(__uint_as_float(sc[c + 6]), __uint_as_float(sc[c + 7])));
        }
        const float m_half = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        // Both warps of the pair hold their scores in registers once they pass the barrier inside exchange():
        // only then may either of them overwrite the S buffer with its half of P.
        const float m_blk = fmaxf(m_half, exchange(b, m_half));
        const float m_new = fmaxf(m_used, m_blk);
        if (j == 0) {
          m_used = m_new;
        } else {
          const bool need = (m_new - m_used) * scale_log2 > kRescaleThreshold;
          if (__any_sync(0xffffffffu, need)) {  // the partner warp sees the same rows and takes the same branch
            mbar_wait(&bars->pv_done[t], (j - 1) & 1);
            tc_fence_after();
            const float alpha = need ? fast_exp2((m_used - m_new) * scale_log2) : 1.f;
            if (need) {
              m_used = m_new;
              l *= alpha;
            }
#pragma unroll
            for (int c0 = 0; c0 < DH; c0 += 32) {
              uint32_t o[32];
              HG_TMEM_LD32(o_addr + c0, o, 0);
              tmem_wait_ld();
#pragma unroll
              for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
              HG_TMEM_ST32(o_addr + c0, o, 0);
            }
          }
        }
        const float neg_mc = -m_used * scale_log2;
        const uint64_t scale2 = pack_f2(scale_log2, scale_log2), neg2 = pack_f2(neg_mc, neg_mc);
        uint64_t ps2[2] = {0ull, 0ull};
        uint32_t pk[16];
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          float x0, x1;
          unpack_f2(ffma2(pack_f2(__uint_as_float(sc[c]), __uint_as_float(sc[c + 1])), scale2, neg2), x0, x1);
          const float p0 = fast_exp2(x0), p1 = fast_exp2(x1);
          ps2[(c >> 1) & 1] = fadd2(ps2[(c >> 1) & 1], pack_f2(p0, p1));
          pk[c >> 1] = pack2<T>(p0, p1);
        }
        HG_TMEM_ST16(s_addr + half * 16, pk, 0);  // P_t(j): keys [half*32, half*32+32) -> columns [half*16, half*16+16)
        {
          float a0, a1;
          unpack_f2(fadd2(ps2[0], ps2[1]), a0, a1);
          l += a0 + a1;
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&bars->p_full[t][b]);
      }

      // ---- epilogue: this warp owns columns [half*DH, half*DH + DH) of its 32 rows of O_t ----
      mbar_wait(&bars->o_full[t], 0);
      tc_fence_after();
      l += exchange(n_blocks & 1, l);  // row sum of both key halves
      const float inv_l = (l > 0.f) ? 1.f / l : 0.f;
      const int tile_row0 = q_row0 + t * BLOCK_M;
      if (rows_valid == BLOCK_M) {
        uint8_t* stage = smem + L::kQ + t * L::kQTileBytes;
#pragma unroll
        for (int c0 = 0; c0 < DH; c0 += 32) {
          uint32_t o[32];
          HG_TMEM_LD32(o_addr + c0, o, 0);
          tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < 32; c += 8) {
            uint4 w;
            w.x = pack2<T>(__uint_as_float(o[c + 0]) * inv_l, __uint_as_float(o[c + 1]) * inv_l);
            w.y = pack2<T>(__uint_as_float(o[c + 2]) * inv_l, __uint_as_float(o[c + 3]) * inv_l);
            w.z = pack2<T>(__uint_as_float(o[c + 4]) * inv_l, __uint_as_float(o[c + 5]) * inv_l);
            w.w = pack2<T>(__uint_as_float(o[c + 6]) * inv_l, __uint_as_float(o[c + 7]) * inv_l);
            const int chunk = (half * DH + c0 + c) >> 3;  // 16-byte chunk of the row
            uint8_t* dst = stage + (chunk >> 3) * L::kQHalfBytes + row * 128 + (((chunk & 7) ^ (row & 7)) << 4);
            *reinterpret_cast<uint4*>(dst) = w;
          }
        }
        fence_proxy_async();
        named_bar_sync_rt(9 + t, 2 * BLOCK_M);  // the eight warps of this tile
        if (half == 0 && wq == 0 && lane == 0) {
#pragma unroll
          for (int h = 0; h < L::kHalves; ++h) tma_store_2d(&tmap_o, stage + h * L::kQHalfBytes, head * D + h * 64, split * n_q_rows + tile_row0);
          bulk_commit_and_wait();
        }
      } else {
        const bool row_ok = row < rows_valid;
        T* orow = out + ((int64_t)(tile_row0 + row) * hq + head) * D + half * DH;
#pragma unroll
        for (int c0 = 0; c0 < DH; c0 += 32) {
          uint32_t o[32];
          HG_TMEM_LD32(o_addr + c0, o, 0);
          tmem_wait_ld();
          if (row_ok) {
#pragma unroll
            for (int c = 0; c < 32; c += 8) {
              uint4 w;
              w.x = pack2<T>(__uint_as_float(o[c + 0]) * inv_l, __uint_as_float(o[c + 1]) * inv_l);
              w.y = pack2<T>(__uint_as_float(o[c + 2]) * inv_l, __uint_as_float(o[c + 3]) * inv_l);
              w.z = pack2<T>(__uint_as_float(o[c + 4]) * inv_l, __uint_as_float(o[c + 5]) * inv_l);
              w.w = pack2<T>(__uint_as_float(o[c + 6]) * inv_l, __uint_as_float(o[c + 7]) * inv_l);
              st_v4(orow + c0 + c, w);
            }
          }
        }
      }
      if (half == 0 && row < rows_valid && lse != nullptr)
        lse[(int64_t)(tile_row0 + row) * hq + head] = (l > 0.f) ? (m_used * scale_log2 + fast_log2(l)) * kLn2 : -INFINITY;
      tc_fence_before();
    }
  } else if constexpr (kSplit == 3) {
    // =============================== softmax, alternate-block form ============================
    // Two warpgroups per tile; the warps that own the same 32 rows take the key blocks in turn (even / odd), each with
    // its own S/P buffer (the double buffer IS the block parity), so four softmax warps share a sub-partition without
    // a per-block exchange of scores.  What the two share per row is the lazily updated reference max: the warp of
    // block j publishes the reference it used (shared memory + an mbarrier with 32 arrivals) as soon as it has the
    // row max of its block -- long before it is done with the block -- and the warp of block j+1 picks it up.  Each
    // warp keeps its own partial row sum relative to the reference it last saw; they meet once, in the epilogue.
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int sw = warp - 4;          // 0 .. 15
    const int t = sw >> 3;            // tile owned by this pair of warpgroups
    const int par = (sw >> 2) & 1;    // this warp takes key blocks j with j % 2 == par (and, in the epilogue, D/2 columns of O)
    const int rows_valid = min(BLOCK_M, rows_left - t * BLOCK_M);
    if (rows_valid > 0) {
      constexpr int DH = D / 2;
      const int half = par;
      const int wq = warp & 3;        // == sw % 4: the TMEM lane quarter this warp may access
      const int row = wq * 32 + lane;
      const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
      const uint32_t o_all = tmem + lane_base + tmem_o(t);        // every column of this row of O_t (lazy rescale)
      const uint32_t o_addr = o_all + (uint32_t)(half * DH);      // the columns this warp writes out
      const uint32_t s_addr = tmem + lane_base + tmem_s(t, par);  // this warp's S / P buffer
      float* xchg = reinterpret_cast<float*>(smem + L::kXchg);
      volatile float* m_mine = xchg + ((par * kTiles + t) * 2 + 0) * BLOCK_M + row;          // reference published by this warp
      volatile float* m_other = xchg + (((par ^ 1) * kTiles + t) * 2 + 0) * BLOCK_M + row;   // ... by its partner
      volatile float* l_mine = xchg + ((par * kTiles + t) * 2 + 1) * BLOCK_M + row;
      volatile float* l_other = xchg + (((par ^ 1) * kTiles + t) * 2 + 1) * BLOCK_M + row;
      uint64_t* ready_mine = &bars->m_ready[t][wq][par];
      uint64_t* ready_other = &bars->m_ready[t][wq][par ^ 1];
      const int pair_bar = 1 + t * 4 + wq;  // the two warps that own the same 32 rows (named barriers 1..8)
      float m_w = -INFINITY;  // reference this warp's partial row sum is expressed in
      float l = 0.f;
      const int row_end = kCausal ? mt * (kTiles * BLOCK_M) + t * BLOCK_M + row + causal_off + 1 : 0x7fffffff;
      const int tile_end = kCausal ? mt * (kTiles * BLOCK_M) + t * BLOCK_M + causal_off + 1 : 0x7fffffff;
      const bool ragged = (k_len % BLOCK_N) != 0;
      for (int j = par; j < n_blocks; j += 2) {
        mbar_wait(&bars->s_full[t][par], (j >> 1) & 1);
        tc_fence_after();
        const bool masked = (ragged && j + 1 == n_blocks) || (j + 1) * BLOCK_N > tile_end;  // warp-uniform
        const int rem = min(k_len, row_end) - j * BLOCK_N;
        uint32_t sc[32];
        // ---- pass 1: row max of the block (the scores are fetched again in pass 2: TMEM reads are cheap, registers are not)
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          HG_TMEM_LD32(s_addr + h * 32, sc, 0);
          tmem_wait_ld();
          if (masked) {
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (h * 32 + c >= rem) sc[c] = 0xff800000u;  // -inf
          }
#pragma unroll
          for (int c = 0; c < 32; c += 8) {
            mx[0] = fmaxf(mx[0], fmaxf(__uint_as_float(sc[c + 0]), __uint_as_float(sc[c + 1])));
            mx[1] = fmaxf(mx[1], fmaxf(__uint_as_float(sc[c + 2]), __uint_as_float(sc[c + 3])));
            mx[2] = fmaxf(mx[2], fmaxf(__uint_as_float(sc[c + 4]), __uint_as_float(sc[c + 5])));
            mx[3] = fmaxf(mx[3], fmaxf(__uint_as_float(sc[c + 6]), __uint_as_float(sc[c + 7])));
          }
        }
        const float m_blk = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        // ---- the reference this block is exponentiated against
        float m_used;
        if (j == 0) {
          m_used = m_blk;
        } else {
          mbar_wait(ready_other, ((j - 1) >> 1) & 1);  // block j-1's warp has published its reference
          const float m_prev = *m_other;
          const float m_new = fmaxf(m_prev, m_blk);
          const bool need = (m_new - m_prev) * scale_log2 > kRescaleThreshold;
          m_used = m_prev;
          if (__any_sync(0xffffffffu, need)) {
            // rare: O_t must be complete up to P_t(j-1) V_{j-1} before it is rescaled in place; P_t(j) has not been
            // released yet, and block j+1's warp cannot release P_t(j+1)'s rescale before P_t(j) V_j is done.
            mbar_wait(&bars->pv_done[t], (j - 1) & 1);
            tc_fence_after();
            const float alpha = need ? fast_exp2((m_prev - m_new) * scale_log2) : 1.f;
            if (need) m_used = m_new;
#pragma unroll
            for (int c0 = 0; c0 < D; c0 += 32) {
              uint32_t o[32];
              HG_TMEM_LD32(o_all + c0, o, 0);
              tmem_wait_ld();
#pragma unroll
              for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
              HG_TMEM_ST32(o_all + c0, o, 0);
            }
          }
        }
        *m_mine = m_used;
        mbar_arrive(ready_mine);  // release: the store above is visible to whoever sees this phase complete
        if (m_w != m_used) {      // bring this warp's partial row sum to the reference in force (first block: 0 * 0)
          l *= fast_exp2((m_w - m_used) * scale_log2);
          m_w = m_used;
        }
        // ---- pass 2: exponentials, row sum, P_t(j)
        const float neg_mc = -m_used * scale_log2;
        const uint64_t scale2 = pack_f2(scale_log2, scale_log2), neg2 = pack_f2(neg_mc, neg_mc);
        uint64_t ps2[2] = {0ull, 0ull};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          HG_TMEM_LD32(s_addr + h * 32, sc, 0);
          tmem_wait_ld();
          if (masked) {
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (h * 32 + c >= rem) sc[c] = 0xff800000u;
          }
          uint32_t pk[16];
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            float x0, x1;
            unpack_f2(ffma2(pack_f2(__uint_as_float(sc[c]), __uint_as_float(sc[c + 1])), scale2, neg2), x0, x1);
            const float p0 = fast_exp2(x0), p1 = fast_exp2(x1);
            ps2[(c >> 1) & 1] = fadd2(ps2[(c >> 1) & 1], pack_f2(p0, p1));
            pk[c >> 1] = pack2<T>(p0, p1);
          }
          // keys [h*32, h*32+32) -> columns [h*16, h*16+16): only columns whose scores this thread has already consumed
          HG_TMEM_ST16(s_addr + h * 16, pk, 0);
        }
        {
          float a0, a1;
          unpack_f2(fadd2(ps2[0], ps2[1]), a0, a1);
          l += a0 + a1;
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&bars->p_full[t][par]);
      }

      // ---- common reference, row-sum exchange, epilogue on this warp's D/2 columns ------------
      const int last = n_blocks - 1;
      float m_final = m_w;
      if ((last & 1) != par) {
        mbar_wait(ready_other, (last >> 1) & 1);
        m_final = *m_other;
      }
      if (m_w != m_final) l *= fast_exp2((m_w - m_final) * scale_log2);
      mbar_wait(&bars->o_full[t], 0);
      tc_fence_after();
      *l_mine = l;
      tc_fence_before();
      named_bar_sync_rt(pair_bar, 64);
      tc_fence_after();
      l += *l_other;
      const float inv_l = (l > 0.f) ? 1.f / l : 0.f;
      const int tile_row0 = q_row0 + t * BLOCK_M;
      if (rows_valid == BLOCK_M) {
        uint8_t* stage = smem + L::kQ + t * L::kQTileBytes;
#pragma unroll
        for (int c0 = 0; c0 < DH; c0 += 32) {
          uint32_t o[32];
          HG_TMEM_LD32(o_addr + c0, o, 0);
          tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < 32; c += 8) {
            uint4 w;
            w.x = pack2<T>(__uint_as_float(o[c + 0]) * inv_l, __uint_as_float(o[c + 1]) * inv_l);
            w.y = pack2<T>(__uint_as_float(o[c + 2]) * inv_l, __uint_as_float(o[c + 3]) * inv_l);
            w.z = pack2<T>(__uint_as_float(o[c + 4]) * inv_l, __uint_as_float(o[c + 5]) * inv_l);
            w.w = pack2<T>(__uint_as_float(o[c + 6]) * inv_l, __uint_as_float(o[c + 7]) * inv_l);
            const int chunk = (half * DH + c0 + c) >> 3;  // 16-byte chunk of the row
            uint8_t* dst = stage + (chunk >> 3) * L::kQHalfBytes + row * 128 + (((chunk & 7) ^ (row & 7)) << 4);
            *reinterpret_cast<uint4*>(dst) = w;
          }
        }
        fence_proxy_async();
        named_bar_sync_rt(9 + t, 2 * BLOCK_M);  // the eight warps of this tile
        if (half == 0 && wq == 0 && lane == 0) {
#pragma unroll
          for (int h = 0; h < L::kHalves; ++h) tma_store_2d(&tmap_o, stage + h * L::kQHalfBytes, head * D + h * 64, split * n_q_rows + tile_row0);
          bulk_commit_and_wait();
        }
      } else {
        const bool row_ok = row < rows_valid;
        T* orow = out + ((int64_t)(tile_row0 + row) * hq + head) * D + half * DH;
#pragma unroll
        for (int c0 = 0; c0 < DH; c0 += 32) {
          uint32_t o[32];
          HG_TMEM_LD32(o_addr + c0, o, 0);
          tmem_wait_ld();
          if (row_ok) {
#pragma unroll
            for (int c = 0; c < 32; c += 8) {
              uint4 w;
              w.x = pack2<T>(__uint_as_float(o[c + 0]) * inv_l, __uint_as_float(o[c + 1]) * inv_l);
              w.y = pack2<T>(__uint_as_float(o[c + 2]) * inv_l, __uint_as_float(o[c + 3]) * inv_l);
              w.z = pack2<T>(__uint_as_float(o[c + 4]) * inv_l, __uint_as_float(o[c + 5]) * inv_l);
              w.w = pack2<T>(__uint_as_float(o[c + 6]) * inv_l, __uint_as_float(o[c + 7]) * inv_l);
              st_v4(orow + c0 + c, w);
            }
          }
        }
      }
      if (half == 0 && row < rows_valid && lse != nullptr)
        lse[(int64_t)(tile_row0 + row) * hq + head] = (l > 0.f) ? (m_final * scale_log2 + fast_log2(l)) * kLn2 : -INFINITY;
      tc_fence_before();
    }
  } else {
    // =============================== softmax / correction / epilogue ==========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int t = (warp - 4) >> 2;     // tile owned by this warpgroup
    const int rows_valid = min(BLOCK_M, rows_left - t * BLOCK_M);
    if (rows_valid > 0) {
      const int wq = warp & 3;           // == warp % 4: the TMEM lane quarter this warp may access
      const int row = wq * 32 + lane;    // row of the tile == TMEM lane
      const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
      const uint32_t o_addr = tmem + lane_base + tmem_o(t);
      float m_used = -INFINITY;          // raw-score max the exponentials are referenced to
      float l = 0.f;

      // Software pipeline over key blocks.  Per block the warp issues 64 MUFU exp2; everything else it
      // has to do -- the scale FFMAs, the row sums and the 16-bit packing of block j, and (once S_t(j+1)
      // can have landed: its Q K^T is only issued after P_t(j-1) was consumed) fetching the scores of
      // block j+1 from TMEM and reducing them to their row max -- is written interleaved with those MUFU
      // requests in groups of 8, so that the in-order warp always has independent work behind them and
      // the two softmax warps sharing an SM sub-partition do not convoy on the MUFU unit.  Two score
      // register arrays alternate between "being exponentiated" and "being fetched".
      // first key (group-relative) this thread's row may NOT see; tile_lim: the same for the tile's first row
      const int row_end = kCausal ? mt * (kTiles * BLOCK_M) + t * BLOCK_M + row + causal_off + 1 : 0x7fffffff;
      const int tile_end = kCausal ? mt * (kTiles * BLOCK_M) + t * BLOCK_M + causal_off + 1 : 0x7fffffff;
      const bool ragged = (k_len % BLOCK_N) != 0;
      // block jb holds a key some row of this tile must not see (warp-uniform)
      auto needs_mask = [&](int jb) { return (ragged && jb + 1 == n_blocks) || (jb + 1) * BLOCK_N > tile_end; };
      if constexpr (kSplit == 2) {
        // "simple" form (experimental, HYDRAGEN_B200_PREFIX_SOFTMAX=simple): no software pipelining -- per block:
        // wait for S_t(j), fetch it, row max, (lazy rescale), exp / sum / pack, store P_t(j), arrive.  The isolated
        // instruction stream in this shape needs 1279 cycles per block pair with two warps per sub-partition
        // (scripts/microbench/softmax_stream.cu, r01r) against the 1625 of the pipelined loop inside the kernel.
        for (int j = 0; j < n_blocks; ++j) {
          const int b = j & 1;
          const uint32_t s_addr = tmem + lane_base + tmem_s(t, b);
          mbar_wait(&bars->s_full[t][b], (j >> 1) & 1);
          tc_fence_after();
          uint32_t sc[BLOCK_N];
          HG_TMEM_LD32(s_addr + 0, sc, 0);
          HG_TMEM_LD32(s_addr + 32, sc, 32);
          tmem_wait_ld();
          if (needs_mask(j)) {
            const int rem = min(k_len, row_end) - j * BLOCK_N;
#pragma unroll
            for (int c = 0; c < BLOCK_N; ++c)
              if (c >= rem) sc[c] = 0xff800000u;  // -inf
          }
          float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int c = 0; c < BLOCK_N; c += 8) {
            mx[0] = fmaxf(mx[0], fmaxf(__uint_as_float(sc[c + 0]), __uint_as_float(sc[c + 1])));
            mx[1] = fmaxf(mx[1], fmaxf(__uint_as_float(sc[c + 2]), __uint_as_float(sc[c + 3])));
            mx[2] = fmaxf(mx[2], fmaxf(__uint_as_float(sc[c + 4]), __uint_as_float(sc[c + 5])));
            mx[3] = fmaxf(mx[3], fmaxf(__uint_as_float(sc[c + 6]), __uint_as_float(sc[c + 7])));
          }
          const float m_new = fmaxf(m_used, fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])));
          if (j == 0) {
            m_used = m_new;
          } else {
            const bool need = (m_new - m_used) * scale_log2 > kRescaleThreshold;
            if (__any_sync(0xffffffffu, need)) {
              mbar_wait(&bars->pv_done[t], (j - 1) & 1);
              tc_fence_after();
              const float alpha = need ? fast_exp2((m_used - m_new) * scale_log2) : 1.f;
              if (need) {
                m_used = m_new;
                l *= alpha;
              }
#pragma unroll
              for (int c0 = 0; c0 < D; c0 += 32) {
                uint32_t o[32];
                HG_TMEM_LD32(o_addr + c0, o, 0);
                tmem_wait_ld();
#pragma unroll
                for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
                HG_TMEM_ST32(o_addr + c0, o, 0);
              }
            }
          }
          const float neg_mc = -m_used * scale_log2;
          const uint64_t scale2 = pack_f2(scale_log2, scale_log2), neg2 = pack_f2(neg_mc, neg_mc);
          uint64_t ps2[2] = {0ull, 0ull};
          uint32_t pk[BLOCK_N / 2];
#pragma unroll
          for (int c = 0; c < BLOCK_N; c += 2) {
            float x0, x1;
            unpack_f2(ffma2(pack_f2(__uint_as_float(sc[c]), __uint_as_float(sc[c + 1])), scale2, neg2), x0, x1);
            const float p0 = fast_exp2(x0), p1 = fast_exp2(x1);
            ps2[(c >> 1) & 1] = fadd2(ps2[(c >> 1) & 1], pack_f2(p0, p1));
            pk[c >> 1] = pack2<T>(p0, p1);
          }
          HG_TMEM_ST16(s_addr, pk, 0);
          HG_TMEM_ST16(s_addr + 16, pk, 16);
          {
            float a0, a1;
            unpack_f2(fadd2(ps2[0], ps2[1]), a0, a1);
            l += a0 + a1;
          }
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive(&bars->p_full[t][b]);
        }
      } else {
      uint32_t sa[BLOCK_N], sb[BLOCK_N];
      float m_blk;
      auto mask_tail = [&](uint32_t(&x)[BLOCK_N], int j) {
        const int rem = min(k_len, row_end) - j * BLOCK_N;
#pragma unroll
        for (int c = 0; c < BLOCK_N; ++c)
          if (c >= rem) x[c] = 0xff800000u;  // -inf
      };
      auto max8 = [&](float* mx, const uint32_t(&x)[BLOCK_N], int g) {
        mx[0] = fmaxf(mx[0], fmaxf(__uint_as_float(x[g * 8 + 0]), __uint_as_float(x[g * 8 + 1])));
        mx[1] = fmaxf(mx[1], fmaxf(__uint_as_float(x[g * 8 + 2]), __uint_as_float(x[g * 8 + 3])));
        mx[2] = fmaxf(mx[2], fmaxf(__uint_as_float(x[g * 8 + 4]), __uint_as_float(x[g * 8 + 5])));
        mx[3] = fmaxf(mx[3], fmaxf(__uint_as_float(x[g * 8 + 6]), __uint_as_float(x[g * 8 + 7])));
      };

      // cur: scores of block j (masked, max known in m_blk); nxt: receives block j+1.
      // kHasNext: block j+1 exists; kMaskNext: it is the ragged last block.
      auto body = [&](int j, uint32_t(&cur)[BLOCK_N], uint32_t(&nxt)[BLOCK_N], auto has_next_tag, auto mask_next_tag) {
        constexpr bool kHasNext = decltype(has_next_tag)::value;
        constexpr bool kMaskNext = decltype(mask_next_tag)::value;
        const uint32_t p_addr = tmem + lane_base + tmem_s(t, j & 1);
        if (wq == 0 && lane == 0) HG_TRACE(1 + t, j, 1);
        const float m_new = fmaxf(m_used, m_blk);
        if (j == 0) {
          m_used = m_new;
        } else {
          const bool need = (m_new - m_used) * scale_log2 > kRescaleThreshold;
          if (__any_sync(0xffffffffu, need)) {
            // rare: O_t must be complete up to P_t(j-1) V_{j-1} before it is rescaled in place; P_t(j)
            // has not been released yet, so no later MMA can be touching O_t.
            mbar_wait(&bars->pv_done[t], (j - 1) & 1);
            tc_fence_after();
            const float alpha = need ? fast_exp2((m_used - m_new) * scale_log2) : 1.f;
            if (need) {
              m_used = m_new;
              l *= alpha;
            }
#pragma unroll
            for (int c0 = 0; c0 < D; c0 += 32) {
              uint32_t o[32];
              HG_TMEM_LD32(o_addr + c0, o, 0);
              tmem_wait_ld();
#pragma unroll
              for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
              HG_TMEM_ST32(o_addr + c0, o, 0);
            }
          }
        }
        const float neg_mc = -m_used * scale_log2;
        uint32_t pk[BLOCK_N / 2];
        const uint64_t scale2 = pack_f2(scale_log2, scale_log2), neg2 = pack_f2(neg_mc, neg_mc);
        uint64_t ps2[2] = {0ull, 0ull};  // packed row sums
        auto exp8 = [&](int g) {  // in place: score -> p; every kEmuEvery-th pair goes to the FMA pipes
#pragma unroll
          for (int c = 0; c < 8; c += 2) {
            float x0, x1;
            unpack_f2(ffma2(pack_f2(__uint_as_float(cur[g * 8 + c]), __uint_as_float(cur[g * 8 + c + 1])), scale2, neg2), x0, x1);
            float p0, p1;
            if (kEmuEvery > 0 && ((c >> 1) % kEmuEvery) == kEmuEvery - 1) {
              exp2_poly_x2(x0, x1, p0, p1);
            } else {
              p0 = fast_exp2(x0);
              p1 = fast_exp2(x1);
            }
            cur[g * 8 + c] = __float_as_uint(p0);
            cur[g * 8 + c + 1] = __float_as_uint(p1);
          }
        };
        auto sum_pack8 = [&](int g) {
#pragma unroll
          for (int c = 0; c < 8; c += 2) {
            const float p0 = __uint_as_float(cur[g * 8 + c]), p1 = __uint_as_float(cur[g * 8 + c + 1]);
            ps2[(c >> 1) & 1] = fadd2(ps2[(c >> 1) & 1], pack_f2(p0, p1));
            pk[(g * 8 + c) >> 1] = pack2<T>(p0, p1);
          }
        };
#pragma unroll
        for (int g = 0; g < 6; ++g) {
          exp8(g);
          if (g >= 2) sum_pack8(g - 2);
        }
        HG_TMEM_ST16(p_addr, pk, 0);  // keys 0..31 of P
        if (wq == 0 && lane == 0) HG_TRACE(1 + t, j, 2);
        if constexpr (kHasNext) {     // S_t(j+1) has had ~3/4 of this block's MUFU time to land
          mbar_wait(&bars->s_full[t][(j + 1) & 1], ((j + 1) >> 1) & 1);
          tc_fence_after();
          const uint32_t s_addr = tmem + lane_base + tmem_s(t, (j + 1) & 1);
          HG_TMEM_LD32(s_addr + 0, nxt, 0);
          HG_TMEM_LD32(s_addr + 32, nxt, 32);
        }
        if (wq == 0 && lane == 0) HG_TRACE(1 + t, j, 3);
        exp8(6);
        sum_pack8(4);
        exp8(7);
        sum_pack8(5);
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        if constexpr (kHasNext) {
          tmem_wait_ld();
          if constexpr (kMaskNext) mask_tail(nxt, j + 1);
#pragma unroll
          for (int g = 0; g < 4; ++g) max8(mx, nxt, g);
        }
        sum_pack8(6);
        sum_pack8(7);
        HG_TMEM_ST16(p_addr + 16, pk, 16);
        if constexpr (kHasNext) {
#pragma unroll
          for (int g = 4; g < 8; ++g) max8(mx, nxt, g);
          m_blk = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        }
        {
          float a0, a1;
          unpack_f2(fadd2(ps2[0], ps2[1]), a0, a1);
          l += a0 + a1;
        }
        if (wq == 0 && lane == 0) HG_TRACE(1 + t, j, 4);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&bars->p_full[t][j & 1]);
        if (wq == 0 && lane == 0) HG_TRACE(1 + t, j, 5);
      };

      // Phase offset between the two tiles: their softmax warps share one MUFU unit per SM sub-partition, and left
      // alone they run in lock-step (both in their exp phase, then both outside it).  Starting tile B part of a
      // block later lets one tile's exponentials run behind the other's TMEM / barrier latencies.
      // (b_delay: below, after the first scores have arrived.  Measured at cfg#2, r01i: 0 -> 33.0 us, 400 -> 32.9,
      // 600 -> 32.4, 800 -> 32.2, 1000 -> 32.7, 1300 -> 33.3; no effect at B = 4096.  Short prefixes skip it.)
      {  // prologue: scores and row max of block 0
        if (wq == 0 && lane == 0) HG_TRACE(1 + t, 0, 0);
        mbar_wait(&bars->s_full[t][0], 0);
        if (t == 1 && b_delay > 0 && n_blocks >= 16) {  // counted from the moment the first scores are there
          const long long t_start = clock64();
          while (clock64() - t_start < (long long)b_delay) {
          }
        }
        tc_fence_after();
        const uint32_t s_addr = tmem + lane_base + tmem_s(t, 0);
        HG_TMEM_LD32(s_addr + 0, sa, 0);
        HG_TMEM_LD32(s_addr + 32, sa, 32);
        tmem_wait_ld();
        if (needs_mask(0)) mask_tail(sa, 0);
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int g = 0; g < 8; ++g) max8(mx, sa, g);
        m_blk = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
      }
      for (int j = 0; j < n_blocks; ++j) {
        const bool last = j + 1 == n_blocks, mask_next = !last && needs_mask(j + 1);
        if ((j & 1) == 0) {
          if (last) body(j, sa, sb, std::false_type{}, std::false_type{});
          else if (mask_next) body(j, sa, sb, std::true_type{}, std::true_type{});
          else body(j, sa, sb, std::true_type{}, std::false_type{});
        } else {
          if (last) body(j, sb, sa, std::false_type{}, std::false_type{});
          else if (mask_next) body(j, sb, sa, std::true_type{}, std::true_type{});
          else body(j, sb, sa, std::true_type{}, std::false_type{});
        }
      }
      }  // pipelined form

      // ---- epilogue --------------------------------------------------------------------------
      mbar_wait(&bars->o_full[t], 0);
      tc_fence_after();
      const float inv_l = (l > 0.f) ? 1.f / l : 0.f;
      const int tile_row0 = q_row0 + t * BLOCK_M;
      if (rows_valid == BLOCK_M) {
        // full tile: O_t / l -> the (dead) Q_t tile in the TMA 128-byte swizzle -> one bulk store per
        // 64-column half.  Row r keeps 16-byte chunk c at chunk slot c ^ (r & 7).
        uint8_t* stage = smem + L::kQ + t * L::kQTileBytes;
#pragma unroll
        for (int c0 = 0; c0 < D; c0 += 32) {
          uint32_t o[32];
          HG_TMEM_LD32(o_addr + c0, o, 0);
          tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < 32; c += 8) {
            uint4 w;
            w.x = pack2<T>(__uint_as_float(o[c + 0]) * inv_l, __uint_as_float(o[c + 1]) * inv_l);
            w.y = pack2<T>(__uint_as_float(o[c + 2]) * inv_l, __uint_as_float(o[c + 3]) * inv_l);
            w.z = pack2<T>(__uint_as_float(o[c + 4]) * inv_l, __uint_as_float(o[c + 5]) * inv_l);
            w.w = pack2<T>(__uint_as_float(o[c + 6]) * inv_l, __uint_as_float(o[c + 7]) * inv_l);
            const int chunk = (c0 + c) >> 3;  // 16-byte chunk of the row
            uint8_t* dst = stage + (chunk >> 3) * L::kQHalfBytes + row * 128 + (((chunk & 7) ^ (row & 7)) << 4);
            *reinterpret_cast<uint4*>(dst) = w;
          }
        }
        fence_proxy_async();
        if (t == 0) named_bar_sync<1, BLOCK_M>(); else named_bar_sync<2, BLOCK_M>();
        if (wq == 0 && lane == 0) {
#pragma unroll
          for (int h = 0; h < L::kHalves; ++h) tma_store_2d(&tmap_o, stage + h * L::kQHalfBytes, head * D + h * 64, split * n_q_rows + tile_row0);
          bulk_commit_and_wait();
        }
      } else {
        const bool row_ok = row < rows_valid;
        T* orow = out + ((int64_t)(tile_row0 + row) * hq + head) * D;
#pragma unroll
        for (int c0 = 0; c0 < D; c0 += 32) {
          uint32_t o[32];
          HG_TMEM_LD32(o_addr + c0, o, 0);
          tmem_wait_ld();
          if (row_ok) {
#pragma unroll
            for (int c = 0; c < 32; c += 8) {
              uint4 w;
              w.x = pack2<T>(__uint_as_float(o[c + 0]) * inv_l, __uint_as_float(o[c + 1]) * inv_l);
              w.y = pack2<T>(__uint_as_float(o[c + 2]) * inv_l, __uint_as_float(o[c + 3]) * inv_l);
              w.z = pack2<T>(__uint_as_float(o[c + 4]) * inv_l, __uint_as_float(o[c + 5]) * inv_l);
              w.w = pack2<T>(__uint_as_float(o[c + 6]) * inv_l, __uint_as_float(o[c + 7]) * inv_l);
              st_v4(orow + c0 + c, w);
            }
          }
        }
      }
      if (row < rows_valid && lse != nullptr)
        lse[(int64_t)(tile_row0 + row) * hq + head] = (l > 0.f) ? (m_used * scale_log2 + fast_log2(l)) * kLn2 : -INFINITY;
      tc_fence_before();
    }
  }

  // ---- teardown ----------------------------------------------------------------------------
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
  }
}

// ---- host side -------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D view [rows, cols] of a 16-bit tensor with row stride `row_stride` elements; boxes of box_rows rows x 64 cols,
// SWIZZLE_128B (a box row is exactly one 128-byte swizzle span), rows past `rows` read as zero / are not written.
static int make_tmap(CUtensorMap* map, const void* base, int dtype, uint64_t rows, uint64_t cols, uint64_t row_stride,
                     uint32_t box_rows) {
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(device_info().encode_tiled);
  if (fn == nullptr) return set_error(HG_ERR_NOT_INITIALIZED, "prefix: cuTensorMapEncodeTiled unavailable (call hg_init first)");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {row_stride * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, dtype == HG_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                  const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(HG_ERR_CUDA, "prefix: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
  return HG_OK;
}

// HYDRAGEN_B200_PREFIX_BDELAY (cycles, read once): start offset of tile B's softmax warps
static int prefix_b_delay() {
  static const int v = [] {
    const char* e = getenv("HYDRAGEN_B200_PREFIX_BDELAY");
    return e != nullptr ? atoi(e) : HG_PREFIX_BDELAY_DEFAULT;
  }();
  return v;
}

template <typename T, int D, bool kCausal, int kSplit>
static int launch_prefix_inst(const PrefixParams& p, int dtype, cudaStream_t s) {
  using L = SmemLayout<D>;
  const int64_t n_q_rows = (int64_t)p.n_groups * p.q_per_group;
  CUtensorMap tq, tk, tv, to;
  int rc;
  if ((rc = make_tmap(&tq, p.q, dtype, n_q_rows, (uint64_t)p.hq * D, p.q_stride_row, BLOCK_M)) != HG_OK) return rc;
  if ((rc = make_tmap(&tk, p.k, dtype, p.n_k_rows, (uint64_t)p.hkv * D, p.kv_stride_row, BLOCK_N)) != HG_OK) return rc;
  if ((rc = make_tmap(&tv, p.v, dtype, p.n_k_rows, (uint64_t)p.hkv * D, p.kv_stride_row, BLOCK_N)) != HG_OK) return rc;
  const int splits = p.kv_splits < 1 ? 1 : p.kv_splits;
  if ((rc = make_tmap(&to, p.out, dtype, n_q_rows * splits, (uint64_t)p.hq * D, (uint64_t)p.hq * D, BLOCK_M)) != HG_OK) return rc;
  const int smem_bytes = L::kTotal + 1024;
  static bool attr_set = false;  // per instantiation; idempotent, racing threads set the same value
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(prefix_attn_sm100_kernel<T, D, kCausal, kSplit>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return set_error(HG_ERR_CUDA, "prefix: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int tiles_per_group = (p.q_per_group + kTiles * BLOCK_M - 1) / (kTiles * BLOCK_M);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(p.n_groups * tiles_per_group * splits), (unsigned)p.hq, 1);
  cfg.blockDim = dim3((kSplit == 1 || kSplit == 3) ? kThreadsSplit : kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, prefix_attn_sm100_kernel<T, D, kCausal, kSplit>, tq, tk, tv, to, (T*)p.out, p.lse, p.cu_seqlens_k, p.q_per_group,
                                     tiles_per_group, p.k_len, p.hq, p.hkv, p.scale_log2, splits, (int)n_q_rows, prefix_b_delay());
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(HG_ERR_CUDA, "prefix_attn_sm100: launch failed: %s", cudaGetErrorString(e));
  }
  return check_launch("prefix_attn_sm100");
}

#if !defined(HG_PREFIX_TU_CAUSAL) && !defined(HG_PREFIX_TU_SPLIT) && !defined(HG_PREFIX_TU_SIMPLE) && !defined(HG_PREFIX_TU_ALT)
#ifdef HG_PREFIX_TRACE
extern "C" int hg_debug_read_trace(long long* host_buf, int n) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(host_buf, g_trace, sizeof(long long) * (size_t)n);
}
#endif

// Number of KV splits that brings the CTA count of a prefix launch close to the SM count without going
// below 4 key blocks per CTA: the head-parallel ranks of a tensor-parallel run own few heads each.
int suggest_prefix_splits(int n_groups, int q_per_group, int hq, int max_k_len, int max_splits) {
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  const long long base = (long long)n_groups * ((q_per_group + kTiles * BLOCK_M - 1) / (kTiles * BLOCK_M)) * hq;
  if (base <= 0) return 1;
  const int n_blocks = (max_k_len + BLOCK_N - 1) / BLOCK_N;
  int s = (int)(sms / base);
  s = min(s, n_blocks / 4);
  s = min(s, max_splits);
  return s < 1 ? 1 : s;
}

int launch_prefix_causal(const PrefixParams& p, int dtype, cudaStream_t s);  // prefix_sm100_causal.cu
int launch_prefix_split(const PrefixParams& p, int dtype, cudaStream_t s);   // prefix_sm100_split.cu
int launch_prefix_simple(const PrefixParams& p, int dtype, cudaStream_t s);  // prefix_sm100_simple.cu
int launch_prefix_alt(const PrefixParams& p, int dtype, cudaStream_t s);     // prefix_sm100_alt.cu

// HYDRAGEN_B200_PREFIX_SOFTMAX = base | split | simple | alt (read once): which softmax organisation the non-causal
// launches use (0 base: software pipelined, the default; 1 split-column; 2 simple: not pipelined; 3 alternate-block)
static int softmax_variant() {
  static const int v = [] {
    const char* e = getenv("HYDRAGEN_B200_PREFIX_SOFTMAX");
    if (e == nullptr) return HG_PREFIX_SOFTMAX_DEFAULT;
    if (e[0] == 's' && e[1] == 'p') return 1;
    if (e[0] == 's' && e[1] == 'i') return 2;
    if (e[0] == 'a') return 3;
    return 0;
  }();
  return v;
}

int launch_prefix(const PrefixParams& p, int dtype, cudaStream_t s) {
  if (p.n_groups == 0 || p.q_per_group == 0) return HG_OK;
  if (dtype != HG_F16 && dtype != HG_BF16)
    return set_error(HG_ERR_UNSUPPORTED, "prefix: the tcgen05 kernel takes f16/bf16 only (dtype %d)", dtype);
  if (p.q_stride_row % 8 != 0 || p.kv_stride_row % 8 != 0 || reinterpret_cast<uintptr_t>(p.q) % 16 != 0 ||
      reinterpret_cast<uintptr_t>(p.k) % 16 != 0 || reinterpret_cast<uintptr_t>(p.v) % 16 != 0 ||
      reinterpret_cast<uintptr_t>(p.out) % 16 != 0)
    return set_error(HG_ERR_UNSUPPORTED, "prefix: TMA needs 16-byte aligned bases and row strides");
  if (p.hq > 65535) return set_error(HG_ERR_UNSUPPORTED, "prefix: hq > 65535");
  if (p.causal) {
    if (p.cu_seqlens_k != nullptr || p.kv_splits > 1 || p.k_len < p.q_per_group)
      return set_error(HG_ERR_UNSUPPORTED, "prefix: the causal form takes uniform groups with k_len >= q rows per group and no kv split");
    return launch_prefix_causal(p, dtype, s);
  }
  if (softmax_variant() == 1) return launch_prefix_split(p, dtype, s);
  if (softmax_variant() == 2) return launch_prefix_simple(p, dtype, s);
  if (softmax_variant() == 3) return launch_prefix_alt(p, dtype, s);
  if (dtype == HG_BF16) {
    if (p.d == 128) return launch_prefix_inst<__nv_bfloat16, 128, false, 0>(p, dtype, s);
    if (p.d == 64) return launch_prefix_inst<__nv_bfloat16, 64, false, 0>(p, dtype, s);
  } else {
    if (p.d == 128) return launch_prefix_inst<__half, 128, false, 0>(p, dtype, s);
    if (p.d == 64) return launch_prefix_inst<__half, 64, false, 0>(p, dtype, s);
  }
  return set_error(HG_ERR_UNSUPPORTED, "prefix: head_dim %d not supported (64 or 128)", p.d);
}
#elif defined(HG_PREFIX_TU_CAUSAL)  // second translation unit: the causal instantiations (compiled in parallel)
int launch_prefix_causal(const PrefixParams& p, int dtype, cudaStream_t s) {
  if (dtype == HG_BF16) {
    if (p.d == 128) return launch_prefix_inst<__nv_bfloat16, 128, true, 0>(p, dtype, s);
    if (p.d == 64) return launch_prefix_inst<__nv_bfloat16, 64, true, 0>(p, dtype, s);
  } else {
    if (p.d == 128) return launch_prefix_inst<__half, 128, true, 0>(p, dtype, s);
    if (p.d == 64) return launch_prefix_inst<__half, 64, true, 0>(p, dtype, s);
  }
  return set_error(HG_ERR_UNSUPPORTED, "prefix: head_dim %d not supported (64 or 128)", p.d);
}
#elif defined(HG_PREFIX_TU_ALT)  // fifth translation unit: the alternate-block softmax instantiations
int launch_prefix_alt(const PrefixParams& p, int dtype, cudaStream_t s) {
  if (dtype == HG_BF16) {
    if (p.d == 128) return launch_prefix_inst<__nv_bfloat16, 128, false, 3>(p, dtype, s);
    if (p.d == 64) return launch_prefix_inst<__nv_bfloat16, 64, false, 3>(p, dtype, s);
  } else {
    if (p.d == 128) return launch_prefix_inst<__half, 128, false, 3>(p, dtype, s);
    if (p.d == 64) return launch_prefix_inst<__half, 64, false, 3>(p, dtype, s);
  }
  return set_error(HG_ERR_UNSUPPORTED, "prefix: head_dim %d not supported (64 or 128)", p.d);
}
#elif defined(HG_PREFIX_TU_SIMPLE)  // fourth translation unit: the non-pipelined softmax instantiations
int launch_prefix_simple(const PrefixParams& p, int dtype, cudaStream_t s) {
  if (dtype == HG_BF16) {
    if (p.d == 128) return launch_prefix_inst<__nv_bfloat16, 128, false, 2>(p, dtype, s);
    if (p.d == 64) return launch_prefix_inst<__nv_bfloat16, 64, false, 2>(p, dtype, s);
  } else {
    if (p.d == 128) return launch_prefix_inst<__half, 128, false, 2>(p, dtype, s);
    if (p.d == 64) return launch_prefix_inst<__half, 64, false, 2>(p, dtype, s);
  }
  return set_error(HG_ERR_UNSUPPORTED, "prefix: head_dim %d not supported (64 or 128)", p.d);
}
#else  // HG_PREFIX_TU_SPLIT: third translation unit: the split-column softmax instantiations
int launch_prefix_split(const PrefixParams& p, int dtype, cudaStream_t s) {
  if (dtype == HG_BF16) {
    if (p.d == 128) return launch_prefix_inst<__nv_bfloat16, 128, false, 1>(p, dtype, s);
    if (p.d == 64) return launch_prefix_inst<__nv_bfloat16, 64, false, 1>(p, dtype, s);
  } else {
    if (p.d == 128) return launch_prefix_inst<__half, 128, false, 1>(p, dtype, s);
    if (p.d == 64) return launch_prefix_inst<__half, 64, false, 1>(p, dtype, s);
  }
  return set_error(HG_ERR_UNSUPPORTED, "prefix: head_dim %d not supported (64 or 128)", p.d);
}
#endif

}  // namespace hg
