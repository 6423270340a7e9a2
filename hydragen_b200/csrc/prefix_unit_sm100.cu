// Shared-prefix attention on the Blackwell tensor cores, ONE CTA PER UNIT: tcgen05.mma + TMEM + TMA (sm_100a).
//
// Replaces the reference's prefix branch for ONE shared level -- hydragen/attention.py:261-338 calling
// flash_attention / flash_attention_varlen (hydragen/flash.py:284-351, i.e. flash-attn v2.3.6's mma.sync kernel), plus
// the LSE transposes of attention.py:276-280,333-338 -- and, as the causal instantiation, the prefill form
// flash_attention(causal=True) (flash.py:284-306).
//
// This is the round-1 kernel: a grid of one CTA per (group, pair of 128-row Q tiles, head [, key split]).  It is the
// launch used whenever ONE level is computed (every cfg#2-style decode step, prefill): the persistent stream-K kernel
// of prefix_sm100.cu runs the same main loop but measured 10-20 % slower per key block on the same box (r02l-r02s:
// 1355 cycles per block here on every SM, 1420-1800 there), so it is used only where its single launch over several
// shared levels pays (hierarchies).
//
// Inter-sequence batching makes this a dense problem: for one (group, head) the queries of every sequence sharing the
// prefix form Q[q_per_group x d] and are multiplied against the single K,V[k_len x d] of that prefix.  One CTA owns TWO
// 128-row Q tiles (A, B) of one head and streams the prefix in 64-key blocks; every K/V block fetched feeds 256 query
// rows, and the score block of each tile is double buffered in TMEM so that Q K^T runs two blocks ahead of the softmax:
//
//   TMEM (512 columns)  S_A[0] S_A[1] S_B[0] S_B[1] (64 fp32 columns each) | O_A | O_B (128 each);
//                       P_t(j) (16-bit) is written back over the first 32 columns of its S buffer
//   warp 0 (1 lane)  TMA producer: Q_A, Q_B once, then a 4-deep ring whose slot u holds what MMA iteration u
//                    consumes: V_u and K_{u+2} (cp.async.bulk.tensor, SWIZZLE_128B boxes)
//   warps 1, 3       MMA issuer of tile A / B (all lanes walk the loop so descriptors stay in uniform registers; one
//                    elected lane issues).  Per key block j:  PV_t(j)  QK_t(j+2)
//   warp 2           TMEM allocator
//   warps 4-7        softmax of tile A, warps 8-11 softmax of tile B: thread t owns row t; software pipelined (the
//                    scores of block j+1 are fetched and reduced to their row max behind the MUFU exp2 requests of block
//                    j); packed fp32x2 arithmetic; lazy rescale of O_t; epilogue O_t / l -> swizzled smem (the dead Q_t
//                    tile) -> TMA store; LSE written directly in [b, nq, hq].
//                    setmaxnreg: 56 registers for warps 0-3, 224 for the softmax warps (no spills in the loop)
//
// Split-KV (kv_splits > 1): the CTAs of one tile each take a contiguous range of key blocks and write their own partial
// (out, lse), merged by the combine that follows anyway (the head-parallel ranks of a TP run own few heads each).
//
// Instantiations: <T, D, kCausal = false> decode hot path; <T, D, kCausal = true> prefill (second translation unit,
// prefix_unit_sm100_causal.cu): row tiles heavy-first, only the visible key blocks are streamed, masks only in the
// blocks crossing the diagonal.
//
// Algorithmic work per CTA: 4 * rows * k_len * d FLOP.  Bound: tensor pipe (AI ~ 680 FLOP/B at the 7B config), with the
// MUFU exp2 rate (16/clk/SM == the 128x128x128 MMA rate) the co-limiter.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace hg {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 64;   // keys per block: S_t block = 64 TMEM columns, double buffered
constexpr int kTiles = 2;  // Q tiles per CTA (ping-pong)
constexpr int kThreads = 384;       // base: TMA / MMA warpgroup + one softmax warpgroup per tile
constexpr uint32_t kTmemCols = 512;
__host__ __device__ constexpr uint32_t tmem_s(int t, int b) { return (uint32_t)t * 128u + (uint32_t)b * 64u; }  // S_t buffer b (P aliases its first 32 columns)
__host__ __device__ constexpr uint32_t tmem_o(int t) { return 256u + (uint32_t)t * 128u; }                       // O_t
constexpr float kRescaleThreshold = 8.0f;  // log2 units
#ifndef HG_PREFIX_BDELAY_DEFAULT
#define HG_PREFIX_BDELAY_DEFAULT 700  // cycles tile B's softmax starts after tile A's (0: together); long prefixes only
#endif

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
  static constexpr int kTotal = kBars + 512;
};

struct Barriers {
  uint64_t q_full[kTiles];
  uint64_t kv_full[kStages], kv_empty[kStages];
  uint64_t s_full[kTiles][2], p_full[kTiles][2];  // per S buffer
  uint64_t pv_done[kTiles];                       // one phase per PV_t(j) (lazy-rescale path only)
  uint64_t o_full[kTiles];                        // O_t complete
  uint32_t tmem_base;
  uint32_t pad_;
};

}  // namespace

// kCausal: bottom-right aligned causal mask inside every group (the prefill form, flash_attention(causal=True) of
// hydragen/flash.py:284-306); a separate instantiation so that the decode-path kernel is exactly the unmasked code.
//
template <typename T, int D, bool kCausal>
__global__ void __launch_bounds__(kThreads, 1)
    prefix_unit_sm100_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
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
    asm volatile("griddepcontrol.wait;" ::: "memory");  // nothing is written before the preceding grid has retired
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
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
      for (int b = 0; b < 2; ++b) {
        mbar_init(&bars->s_full[i][b], 1);
        mbar_init(&bars->p_full[i][b], BLOCK_M);
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
  // This launch may be a programmatic dependent of whatever precedes it on the stream (the previous layer's decode
  // kernel, or the projection that produced q): everything above -- barrier init, TMEM allocation, descriptor prefetch --
  // overlapped its tail; q is read and out / lse are written only from here on.  Only THEN is the launch that follows
  // (the fused append / suffix / combine kernel) allowed to start on the SMs this grid leaves idle: its early work (KV
  // append, suffix walk) can then never overtake the kernel in front of this one; it waits for this grid's completion
  // itself before it reads the partial results written here.  (No-ops when launched without the attribute.)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // Register budget (setmaxnreg must sit inside the role branch it applies to): the producer
  // warpgroup gives registers back, the two softmax warpgroups (128 live fp32 scores per thread)
  // take them: 128 x 56 + 256 x 224 = 384 x 168, the launch-time allocation (r01h: with 88 / 208 the running max, row
  // sum and loop state of the softmax threads were spilled to local memory, on the serial path between two blocks).
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
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
        mbar_wait(&bars->kv_full[st], ((j + 2) / kStages) & 1);
        mbar_wait(&bars->p_full[t][b], (j >> 1) & 1);
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
      }
    }
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
        auto exp8 = [&](int g) {  // in place: score -> p
#pragma unroll
          for (int c = 0; c < 8; c += 2) {
            float x0, x1;
            unpack_f2(ffma2(pack_f2(__uint_as_float(cur[g * 8 + c]), __uint_as_float(cur[g * 8 + c + 1])), scale2, neg2), x0, x1);
            cur[g * 8 + c] = __float_as_uint(fast_exp2(x0));
            cur[g * 8 + c + 1] = __float_as_uint(fast_exp2(x1));
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
        if constexpr (kHasNext) {     // S_t(j+1) has had ~3/4 of this block's MUFU time to land
          mbar_wait(&bars->s_full[t][(j + 1) & 1], ((j + 1) >> 1) & 1);
          tc_fence_after();
          const uint32_t s_addr = tmem + lane_base + tmem_s(t, (j + 1) & 1);
          HG_TMEM_LD32(s_addr + 0, nxt, 0);
          HG_TMEM_LD32(s_addr + 32, nxt, 32);
        }
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
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&bars->p_full[t][j & 1]);
      };

      // Phase offset between the two tiles: their softmax warps share one MUFU unit per SM sub-partition, and left
      // alone they run in lock-step (both in their exp phase, then both outside it).  Starting tile B part of a
      // block later lets one tile's exponentials run behind the other's TMEM / barrier latencies.
      // (b_delay: below, after the first scores have arrived.  Measured at cfg#2, r01i: 0 -> 33.0 us, 400 -> 32.9,
      // 600 -> 32.4, 800 -> 32.2, 1000 -> 32.7, 1300 -> 33.3; no effect at B = 4096.  Short prefixes skip it.)
      {  // prologue: scores and row max of block 0
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
          bulk_commit();
          bulk_wait_all();
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

// HYDRAGEN_B200_PREFIX_BDELAY (cycles, read once): start offset of tile B's softmax warps.  (Structural replacements --
// the two tiles' MMA warps issuing strictly in turn, or in step -- measured 25 % SLOWER, r02p / r02q: the free-running
// interleave is the fast one, and this offset only nudges where it starts.)
static int prefix_b_delay() {
  static const int v = [] {
    const char* e = getenv("HYDRAGEN_B200_PREFIX_BDELAY");
    return e != nullptr ? atoi(e) : HG_PREFIX_BDELAY_DEFAULT;
  }();
  return v;
}

// One shared level (p.levels[0]) on a grid of one CTA per (group, tile pair, head, key split); kv_splits > 1 writes
// that many partial results back to back into the level's out / lse.
template <typename T, int D, bool kCausal>
static int launch_unit_inst(const PrefixParams& p, int kv_splits, int dtype, cudaStream_t s) {
  using L = SmemLayout<D>;
  const PrefixLevel& lv = p.levels[0];
  const int q_per_group = (int)(p.n_q_rows / lv.n_groups);
  CUtensorMap tq, tk, tv, to;
  int rc;
  if ((rc = make_tmap(&tq, p.q, dtype, (uint64_t)p.n_q_rows, (uint64_t)p.hq * D, (uint64_t)p.q_stride_row, BLOCK_M)) != HG_OK) return rc;
  if ((rc = make_tmap(&tk, lv.k, dtype, (uint64_t)lv.n_k_rows, (uint64_t)p.hkv * D, (uint64_t)lv.kv_stride_row, BLOCK_N)) != HG_OK) return rc;
  if ((rc = make_tmap(&tv, lv.v, dtype, (uint64_t)lv.n_k_rows, (uint64_t)p.hkv * D, (uint64_t)lv.kv_stride_row, BLOCK_N)) != HG_OK) return rc;
  if ((rc = make_tmap(&to, lv.out, dtype, (uint64_t)p.n_q_rows * kv_splits, (uint64_t)p.hq * D, (uint64_t)p.hq * D, BLOCK_M)) != HG_OK) return rc;
  const int smem_bytes = L::kTotal + 1024;
  static bool attr_set[64] = {};  // per instantiation and device; idempotent, racing threads set the same value
  const int dev = device_info().device;
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(prefix_unit_sm100_kernel<T, D, kCausal>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return set_error(HG_ERR_CUDA, "prefix: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const int tiles_per_group = (q_per_group + kTiles * BLOCK_M - 1) / (kTiles * BLOCK_M);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(lv.n_groups * tiles_per_group * kv_splits), (unsigned)p.hq, 1);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, prefix_unit_sm100_kernel<T, D, kCausal>, tq, tk, tv, to, (T*)lv.out, lv.lse, lv.cu_seqlens_k, q_per_group,
                                     tiles_per_group, lv.k_len, p.hq, p.hkv, p.scale_log2, kv_splits, (int)p.n_q_rows, prefix_b_delay());
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(HG_ERR_CUDA, "prefix_unit_sm100: launch failed: %s", cudaGetErrorString(e));
  }
  return check_launch("prefix_unit_sm100");
}

#if !defined(HG_UNIT_TU_CAUSAL)
// Number of KV splits that brings the CTA count of a one-level launch close to the SM count without going below 4 key
// blocks per CTA: the head-parallel ranks of a tensor-parallel run own few heads each.
int suggest_prefix_splits(int n_groups, int q_per_group, int hq, int max_k_len, int max_splits) {
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  const long long base = (long long)n_groups * ((q_per_group + kTiles * BLOCK_M - 1) / (kTiles * BLOCK_M)) * hq;
  if (base <= 0) return 1;
  const int n_blocks = (max_k_len + BLOCK_N - 1) / BLOCK_N;
  int s = (int)(sms / base);
  s = std::min(s, n_blocks / 4);
  s = std::min(s, max_splits);
  return s < 1 ? 1 : s;
}

int launch_prefix_unit_causal(const PrefixParams& p, int dtype, cudaStream_t s);  // prefix_unit_sm100_causal.cu

int launch_prefix_unit(const PrefixParams& p, int kv_splits, int dtype, cudaStream_t s) {
  if (p.causal) return launch_prefix_unit_causal(p, dtype, s);
  if (dtype == HG_BF16) {
    if (p.d == 128) return launch_unit_inst<__nv_bfloat16, 128, false>(p, kv_splits, dtype, s);
    return launch_unit_inst<__nv_bfloat16, 64, false>(p, kv_splits, dtype, s);
  }
  if (p.d == 128) return launch_unit_inst<__half, 128, false>(p, kv_splits, dtype, s);
  return launch_unit_inst<__half, 64, false>(p, kv_splits, dtype, s);
}
#else  // second translation unit: the causal instantiations (compiled in parallel)
int launch_prefix_unit_causal(const PrefixParams& p, int dtype, cudaStream_t s) {
  if (dtype == HG_BF16) {
    if (p.d == 128) return launch_unit_inst<__nv_bfloat16, 128, true>(p, 1, dtype, s);
    return launch_unit_inst<__nv_bfloat16, 64, true>(p, 1, dtype, s);
  }
  if (p.d == 128) return launch_unit_inst<__half, 128, true>(p, 1, dtype, s);
  return launch_unit_inst<__half, 64, true>(p, 1, dtype, s);
}
#endif

}  // namespace hg
