// Shared-prefix attention on the Blackwell tensor cores: tcgen05.mma + TMEM + TMA (sm_100a), as ONE persistent
// launch over every shared level of a hierarchy.
//
// Replaces the reference's prefix branch -- hydragen/attention.py:250-341, one flash_attention /
// flash_attention_varlen call per shared level (hydragen/flash.py:284-351, i.e. flash-attn v2.3.6's mma.sync
// kernel) plus the LSE transposes of attention.py:276-280,333-338.
//
// Inter-sequence batching makes this a dense problem: for one (level, group, head) the queries of every sequence
// sharing the prefix form Q[q_per_group x d] and are multiplied against the single K,V[k_len x d] of that prefix.
// A UNIT of work is one (level, group, pair of 128-row Q tiles, head); the launch is a grid of persistent CTAs (one
// per SM) that walk a stream-K schedule over all units of all levels (prefix_sched.h): a CTA gets a contiguous
// range of the (unit, key block) space, so a unit may be split into PIECES handled by different CTAs.  Every piece
// of a split unit leaves its unnormalised fp32 accumulator, running max and row sum in the workspace; when a CTA has
// finished its range it merges, for every split unit it took part in, its share of the rows (all pieces end at about
// the same time, so nobody waits long) and writes the final rows.  Whole units write their rows directly.
//
// Inside a piece (unchanged from round 1): the CTA streams the keys in 64-key blocks; every K/V block fetched feeds
// 256 query rows, and the score block of each tile is double buffered in TMEM so that Q K^T runs two blocks ahead of
// the softmax:
//
//   TMEM (512 columns)  S_A[0] S_A[1] S_B[0] S_B[1] (64 fp32 columns each) | O_A | O_B (128 each);
//                       P_t(j) (16-bit) is written back over the first 32 columns of its S buffer
//   warp 0 (1 lane)  TMA producer: per piece Q_A, Q_B, then a 4-deep ring whose slot u holds what MMA iteration u
//                    consumes: V_u and K_{u+2} (cp.async.bulk.tensor, SWIZZLE_128B boxes)
//   warps 1, 3       MMA issuer of tile A / B (all lanes walk the loop so descriptors stay in uniform registers; one
//                    elected lane issues).  Per key block j:  PV_t(j)  QK_t(j+2)
//                      S_t = Q_t K_j^T  (SS form, both operands K-major in smem, 128x64x16 per instruction)
//                      O_t += P_t V_j   (TS form: P_t read from TMEM as the A operand, V_j straight from its
//                                        row-major smem tile as an MN-major B operand -- no transpose pass)
//   warp 2           TMEM allocator
//   warps 4-7        softmax of tile A, warps 8-11 softmax of tile B: thread t owns row t (tcgen05.ld 32x32b: lane ==
//                    row, so the row max / row sum need no shuffles); software pipelined: the scores of block j+1
//                    are fetched and reduced to their row max behind the MUFU exp2 requests of block j; scale /
//                    subtract / row sums as packed fp32x2; lazy rescale of O_t (only when the running max grows by
//                    more than 2^8); epilogue of a whole unit: O_t / l -> swizzled smem (the dead Q_t tile) -> TMA
//                    store, LSE written directly in [b, nq, hq].
//                    setmaxnreg: 56 registers for warps 0-3, 224 for the softmax warps (no spills in the loop)
//
// All producer/consumer edges are mbarriers with phases counted across pieces (TMA complete_tx, tcgen05.commit,
// thread arrives); there is no __syncthreads between set-up and teardown.
//
// Used for launches over SEVERAL shared levels (hg_prefix_attn_grouped_fwd with n_levels >= 2); one level goes to the
// one-CTA-per-unit kernel of prefix_unit_sm100.cu, whose main loop measured 10-20 % faster per key block (launch_prefix
// below).  The kCausal template parameter is kept in the code (the mask logic is shared) but only the unmasked form is
// instantiated here.
//
// Algorithmic work per unit: 4 * rows * k_len * d FLOP.  Bound: tensor pipe (AI ~ 680 FLOP/B at the 7B config), with
// the MUFU exp2 rate (16/clk/SM == the 128x128x128 MMA rate) the co-limiter.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "common.cuh"
#include "prefix_sched.h"
#include "sm100_ptx.cuh"

namespace hg {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 64;   // keys per block: S_t block = 64 TMEM columns, double buffered
constexpr int kTiles = 2;     // Q tiles per unit (ping-pong)
constexpr int kThreads = 384;  // TMA / MMA warpgroup + one softmax warpgroup per tile
constexpr uint32_t kTmemCols = 512;
__host__ __device__ constexpr uint32_t tmem_s(int t, int b) { return (uint32_t)t * 128u + (uint32_t)b * 64u; }  // S_t buffer b (P aliases its first 32 columns)
__host__ __device__ constexpr uint32_t tmem_o(int t) { return 256u + (uint32_t)t * 128u; }                       // O_t
constexpr float kRescaleThreshold = 8.0f;  // log2 units

// ---- workspace (caller-owned, zero-initialised once; see hg_prefix_workspace_bytes) ----------------------
constexpr int kWsWordEpoch = 0;           // launches completed on this workspace
constexpr int kWsWordExit = 1;            // CTAs of the running launch that are done
constexpr int kWsWordFlags = 32;          // flag (cta, slot, tile) = word 32 + (cta * 2 + slot) * 2 + tile
constexpr int kWsFlagBytes = 4096;        // >= (32 + 4 * kMaxCtas) * 4
// one slot = the partial result of one piece: per tile, O as [16-byte chunk][row][4 floats] (a warp whose lanes are
// consecutive rows reads / writes 512 contiguous bytes), then (max in log2 units, row sum) per row
__host__ __device__ constexpr int64_t ws_slot_floats(int d) { return (int64_t)kTiles * BLOCK_M * (d + 2); }
__host__ __device__ constexpr int64_t ws_o_index(int d, int tile, int row, int chunk) { return (int64_t)tile * BLOCK_M * d + ((int64_t)chunk * BLOCK_M + row) * 4; }
__host__ __device__ constexpr int64_t ws_ml_index(int d, int tile, int row) { return (int64_t)kTiles * BLOCK_M * d + (tile * BLOCK_M + row) * 2; }
__host__ __device__ constexpr int64_t ws_bytes(int d) { return kWsFlagBytes + (int64_t)kMaxCtas * 2 * ws_slot_floats(d) * 4; }

#if defined(HG_PREFIX_TRACE)
// Development aid (never in the shipped library): %globaltimer stamps [CTA][stage] of the most recent launch.
// 0 entry | 1 set-up done | 2 griddepcontrol.wait passed | 3-5 first piece: first scores there, main loop done, epilogue
// done | 6-8 the same for the last piece | 9 merge duties done | 10 CTA done   (3-9: row 0 of tile A's softmax warpgroup)
__device__ long long g_prefix_trace[kMaxCtas * 16];
__device__ __forceinline__ void ptrace(int stage) {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  g_prefix_trace[(blockIdx.y * gridDim.x + blockIdx.x) * 16 + stage] = t;
}
#define HG_PTRACE(cond, stage) \
  do {                         \
    if (cond) ptrace(stage);   \
  } while (0)
// cycles a designated thread spends inside one kind of barrier wait, accumulated over the launch
#define HG_PWAIT(cond, slot, stmt)                                                        \
  do {                                                                                    \
    const long long hg_t0_ = clock64();                                                   \
    stmt;                                                                                 \
    if (cond) g_prefix_trace[(blockIdx.y * gridDim.x + blockIdx.x) * 16 + (slot)] += clock64() - hg_t0_;             \
  } while (0)
#else
#define HG_PWAIT(cond, slot, stmt) \
  do {                             \
    stmt;                          \
  } while (0)
#define HG_PTRACE(cond, stage) \
  do {                         \
  } while (0)
#endif

constexpr int kStages = 4;  // K/V ring depth

// Ring slot of MMA iteration u (u = -2 .. n-1 inside a piece of n key blocks) holds what that iteration consumes:
// V_u (for P V of block u) and K_{u+2} (for Q K^T of block u+2, issued in the same iteration); iterations -2 and -1
// carry only K_0 / K_1 for the prologue.  One full and one empty barrier per slot; slots and phases are counted
// across pieces.
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
  static constexpr int kSched = kBars + 512;  // copy of the launch's SchedParams (cut table included)
  static constexpr int kCtx = kSched + ((int)sizeof(SchedParams) + 15) / 16 * 16;  // WarpCtx of the 8 softmax warps
  static constexpr int kTotal = kCtx + 8 * 128;
};

struct Barriers {
  uint64_t q_full[kTiles], q_empty[kTiles];       // Q_t landed / Q_t smem (also the output staging tile) free again
  uint64_t kv_full[kStages], kv_empty[kStages];
  uint64_t s_full[kTiles][2], p_full[kTiles][2];  // per S buffer
  uint64_t pv_done[kTiles];                       // one phase per PV_t(j) (lazy-rescale path only)
  uint64_t o_full[kTiles], o_empty[kTiles];       // O_t of a piece complete / drained by the epilogue
  uint32_t tmem_base;
  uint32_t pad_;
};

// Warp-uniform bookkeeping of a softmax warp, PARKED in shared memory while its main loop runs: the loop needs every
// register (128 live scores + 32 packed P + the exp pipeline), and whatever is merely live ACROSS it gets spilled to local
// memory and re-read per key block (r02l: 2.5 LDL per block, 1440-1800 cycles per block against round 1's uniform 1355).
// What the loop needs travels in registers; what the epilogue needs is re-read from here after the loop.
struct WarpCtx {
  SchedIter it;
  SchedPiece sp;
  uint32_t c0, c1, g, np;  // uses of S buffer 0 / 1, key blocks, pieces this warp's tile has been through (barrier phases)
  int duty[2];     // units cut into more than two pieces that this CTA holds a piece of (by workspace slot)
  int first_piece;
};
static_assert(sizeof(WarpCtx) <= 128, "WarpCtx slot");

struct LevelDev {
  void* out;           // [n_q_rows, hq, D]
  float* lse;          // [n_q_rows, hq] or nullptr
  const int32_t* cu;   // cu_seqlens_k (device) or nullptr
  int k_len;           // uniform key count per group (cu == nullptr)
  int pad_;
};

// Kernel parameters.  Order matters: what the main loop touches (scale, pointers) sits at the FRONT of the constant
// bank; the 1.6 KB of tensor maps and the 1.4 KB schedule table come last.  (r02i/r02j: with the scale factor at byte
// 4148 of the bank the compiler's per-block LDC of it missed the SM's constant cache and was served by the cache the
// SMs of a GPC share -- whole GPCs ran their main loops up to 20 % slower than others, different ones every run.)
struct alignas(64) PrefixKernelParams {
  float scale_log2;
  int hkv;
  int pad0_;
  int pad_;
  uint32_t* ws_flags;  // workspace: epoch / exit counter / piece flags (nullptr: no unit is ever split)
  float* ws_part;      // workspace: partial-result slots
  LevelDev lv[kMaxLevels];
  CUtensorMap tmap_q;
  CUtensorMap tmap_k[kMaxLevels], tmap_v[kMaxLevels], tmap_o[kMaxLevels];
  SchedParams sched;
};

// Everything a role needs to know about one piece (uniform over the CTA, recomputed by every role).
struct Piece {
  int level, head, kvh, mt;
  int q_row0;      // first query row of the unit (tile A)
  int rows_left;   // query rows from q_row0 to the end of the group (> 0)
  int k_row0;      // first key row of the group in the level's K / V tensors
  int k_len;       // keys of the group this unit may see (causal: up to the horizon of its last row)
  int nb_group;    // key blocks covering k_len
  int b_lo, n;     // first key block and number of key blocks of this piece (n may be 0)
  int causal_off;
  int split, slot, unit;
};

template <bool kCausal>
__device__ __forceinline__ void resolve_piece(const PrefixKernelParams& P, const SchedParams& S, const SchedPiece& sp, Piece& pc) {
  const SchedLevel& L = S.lv[sp.level];
  const LevelDev& V = P.lv[sp.level];
  pc.level = sp.level;
  pc.head = sp.head;
  pc.kvh = sp.head / (S.hq / P.hkv);
  pc.mt = sp.mt;
  pc.q_row0 = sp.grp * L.q_per_group + sp.mt * (kTiles * BLOCK_M);
  pc.rows_left = L.q_per_group - sp.mt * (kTiles * BLOCK_M);
  if (V.cu != nullptr) {
    pc.k_row0 = __ldg(V.cu + sp.grp);
    pc.k_len = __ldg(V.cu + sp.grp + 1) - pc.k_row0;
  } else {
    pc.k_row0 = sp.grp * V.k_len;
    pc.k_len = V.k_len;
  }
  pc.causal_off = 0;
  if (kCausal) {
    // bottom-right aligned (flash-attn >= 2.1): row r of the group sees keys j <= r + causal_off.  The unit streams
    // only the keys its last row can see; the rows above it are masked per element in the diagonal blocks.
    pc.causal_off = pc.k_len - L.q_per_group;  // >= 0 (checked by the launcher)
    const int last_row = min(L.q_per_group, (sp.mt + 1) * (kTiles * BLOCK_M)) - 1;
    pc.k_len = min(pc.k_len, last_row + pc.causal_off + 1);
  }
  pc.nb_group = (pc.k_len + BLOCK_N - 1) / BLOCK_N;
  pc.b_lo = sp.b_lo;
  pc.n = max(0, min(sp.b_hi, pc.nb_group) - sp.b_lo);
  pc.split = sp.split;
  pc.slot = sp.slot;
  pc.unit = sp.unit;
}

// ---- cold paths of a split unit, kept out of line so that their registers do not weigh on the main loop -----------

// Which CTAs hold the pieces of `unit` (key order) and in which workspace slot; returns the piece count.
__device__ __noinline__ int unit_pieces(const SchedParams& S, int unit, int near, int* ctas, int* slots) {
  const int k = sched_unit_pieces(S, unit, near, ctas, slots);
  return k < kMaxUnitPieces ? k : kMaxUnitPieces;
}

// Role of this CTA's piece of a split unit: 0 = write the partial to the workspace (unit has 2 pieces, the other
// CTA owns it), 1 = owner of a 2-piece unit (other = the slot index (cta * 2 + slot) of the tail piece's partial),
// 2 = write the partial AND take part in the merge afterwards (3+ pieces).
__device__ __noinline__ int split_role(const SchedParams& S, int unit, int cta, int* other) {
  int ctas[kMaxUnitPieces], slots[kMaxUnitPieces];
  const int k = unit_pieces(S, unit, cta, ctas, slots);
  if (k > 2) return 2;
  if (k == 2 && ctas[0] == cta) {
    *other = ctas[1] * 2 + slots[1];
    return 1;
  }
  return 0;
}

// Merge of a unit cut into 3+ pieces: this CTA's slice of the rows; one thread per (row, 32-column chunk), the
// partials of two pieces in flight at a time (the loads are L2 round trips: latency, not bandwidth, is the cost).
template <typename T, int D>
__device__ __noinline__ void merge_duty(const SchedParams& S, const float* ws_part, const uint32_t* ws_flags, uint32_t epoch, int unit, int tid,
                                        int lane, T* out, float* lse, int q_row0, int rows, int head, int hq, int cta) {
  constexpr int kChunks = D / 32;
  int ctas[kMaxUnitPieces], slots[kMaxUnitPieces];
  const int k = unit_pieces(S, unit, cta, ctas, slots);
  int me = 0;
  for (int p = 0; p < k; ++p)
    if (ctas[p] == cta) me = p;
  const int r_lo = (int)((long long)me * rows / k), r_hi = (int)((long long)(me + 1) * rows / k);
  const int ntile = rows > BLOCK_M ? 2 : 1;
  for (int i = lane; i < k * ntile; i += 32) {
    const uint32_t* f = ws_flags + kWsWordFlags + (ctas[i / ntile] * 2 + slots[i / ntile]) * 2 + (i % ntile);
    while ((int32_t)(ld_acquire_gpu(f) - epoch) < 0) {
    }
  }
  __syncwarp();
  for (int item = tid; item < (r_hi - r_lo) * kChunks; item += kTiles * BLOCK_M) {
    const int r = r_lo + item / kChunks, c0 = (item % kChunks) * 32;
    const int tt = r / BLOCK_M, rr = r % BLOCK_M;
    // weights of the pieces for this row: 4 (max, sum) pairs in flight at a time
    float w[kMaxUnitPieces];
    float mx = -INFINITY, lsum = 0.f;
    for (int p0 = 0; p0 < k; p0 += 4) {
      float2 ml[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int p = min(p0 + i, k - 1);
        ml[i] = ld_cg_f2(ws_part + (int64_t)(ctas[p] * 2 + slots[p]) * ws_slot_floats(D) + ws_ml_index(D, tt, rr));
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (p0 + i < k) {
          w[p0 + i] = ml[i].y > 0.f ? ml[i].x : -INFINITY;  // max for now
          if (ml[i].y > 0.f) mx = fmaxf(mx, ml[i].x);
        }
    }
    for (int p0 = 0; p0 < k; p0 += 4) {
      float2 ml[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int p = min(p0 + i, k - 1);
        ml[i] = ld_cg_f2(ws_part + (int64_t)(ctas[p] * 2 + slots[p]) * ws_slot_floats(D) + ws_ml_index(D, tt, rr));
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (p0 + i < k) {
          const float e = ml[i].y > 0.f ? fast_exp2(w[p0 + i] - mx) : 0.f;
          w[p0 + i] = e;
          lsum += e * ml[i].y;
        }
    }
    const float inv = lsum > 0.f ? 1.f / lsum : 0.f;
    float acc[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = 0.f;
    for (int p = 0; p < k; p += 2) {
      const float w0 = w[p] * inv, w1 = p + 1 < k ? w[p + 1] * inv : 0.f;
      const float* s0 = ws_part + (int64_t)(ctas[p] * 2 + slots[p]) * ws_slot_floats(D);
      const float* s1 = ws_part + (int64_t)(ctas[min(p + 1, k - 1)] * 2 + slots[min(p + 1, k - 1)]) * ws_slot_floats(D);
      float4 a[8], b[8];
      if (w0 != 0.f) {
#pragma unroll
        for (int c = 0; c < 8; ++c) a[c] = ld_cg_f4(s0 + ws_o_index(D, tt, rr, (c0 >> 2) + c));
      }
      if (w1 != 0.f) {
#pragma unroll
        for (int c = 0; c < 8; ++c) b[c] = ld_cg_f4(s1 + ws_o_index(D, tt, rr, (c0 >> 2) + c));
      }
      if (w0 != 0.f) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          acc[4 * c + 0] += w0 * a[c].x;
          acc[4 * c + 1] += w0 * a[c].y;
          acc[4 * c + 2] += w0 * a[c].z;
          acc[4 * c + 3] += w0 * a[c].w;
        }
      }
      if (w1 != 0.f) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          acc[4 * c + 0] += w1 * b[c].x;
          acc[4 * c + 1] += w1 * b[c].y;
          acc[4 * c + 2] += w1 * b[c].z;
          acc[4 * c + 3] += w1 * b[c].w;
        }
      }
    }
    T* orow = out + ((int64_t)(q_row0 + r) * hq + head) * D + c0;
#pragma unroll
    for (int c = 0; c < 32; c += 8) {
      uint4 wv;
      wv.x = pack2<T>(acc[c + 0], acc[c + 1]);
      wv.y = pack2<T>(acc[c + 2], acc[c + 3]);
      wv.z = pack2<T>(acc[c + 4], acc[c + 5]);
      wv.w = pack2<T>(acc[c + 6], acc[c + 7]);
      st_v4(orow + c, wv);
    }
    if (c0 == 0 && lse != nullptr) lse[(int64_t)(q_row0 + r) * hq + head] = lsum > 0.f ? (mx + fast_log2(lsum)) * kLn2 : -INFINITY;
  }
}

}  // namespace

template <typename T, int D, bool kCausal>
__global__ void __launch_bounds__(kThreads, 1) prefix_attn_sm100_kernel(const __grid_constant__ PrefixKernelParams P) {
  using L = SmemLayout<D>;
  constexpr int kFmt = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;
  constexpr uint32_t kIdescQK = make_idesc(kFmt, 0, BLOCK_M, BLOCK_N);
  constexpr uint32_t kIdescPV = make_idesc(kFmt, 1, BLOCK_M, D);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  Barriers* bars = reinterpret_cast<Barriers*>(smem + L::kBars);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta_id = blockIdx.y * gridDim.x + blockIdx.x;
  // The schedule is read from shared memory: the out-of-line helpers take it by reference, and a reference into the
  // kernel parameters is a generic pointer into constant memory -- every table look-up a dependent ~0.5 us load
  // (r02g trace: 6 us per split-piece epilogue).  Copied once per CTA, before the set-up barrier.
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&P.sched);
    uint32_t* dst = reinterpret_cast<uint32_t*>(smem + L::kSched);
    for (int i = threadIdx.x; i < (int)(sizeof(SchedParams) / 4); i += kThreads) dst[i] = src[i];
  }
  const SchedParams& S = *reinterpret_cast<const SchedParams*>(smem + L::kSched);
  const int hq = P.sched.hq;
  float scale_log2 = P.scale_log2;
  asm volatile("mov.b32 %0, %0;" : "+f"(scale_log2));  // opaque: stays in a register instead of being re-read from the constant bank
  HG_PTRACE(threadIdx.x == 0, 0);
#if defined(HG_PREFIX_TRACE)
  if (threadIdx.x == 0) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    g_prefix_trace[cta_id * 16 + 11] = smid;
    g_prefix_trace[cta_id * 16 + 8] = 0;   // MMA warp A: cycles waiting for P (softmax late)
    g_prefix_trace[cta_id * 16 + 14] = 0;  // softmax A, row 0: cycles waiting for S (tensor pipe late)
    g_prefix_trace[cta_id * 16 + 15] = 0;  // MMA warp A: cycles waiting for K/V (TMA late)
  }
#endif

  // ---- one-time setup --------------------------------------------------------------------
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.tmap_q) : "memory");
    for (int l = 0; l < S.n_levels; ++l) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&P.tmap_k[l]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&P.tmap_v[l]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&P.tmap_o[l]) : "memory");
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kTiles; ++i) {
      mbar_init(&bars->q_full[i], 1);
      mbar_init(&bars->q_empty[i], 1);
      mbar_init(&bars->pv_done[i], 1);
      mbar_init(&bars->o_full[i], 1);
      mbar_init(&bars->o_empty[i], BLOCK_M);
      for (int b = 0; b < 2; ++b) {
        mbar_init(&bars->s_full[i][b], 1);
        mbar_init(&bars->p_full[i][b], BLOCK_M);
      }
    }
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&bars->kv_full[i], 1);
      mbar_init(&bars->kv_empty[i], kTiles);  // one arrival per MMA warp (a tcgen05.commit, or a plain arrive when its tile is idle)
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
  HG_PTRACE(threadIdx.x == 0, 1);
  // This launch may be a programmatic dependent of whatever precedes it on the stream (the previous layer's decode
  // kernel, or the projection that produced q): everything above overlapped its tail; q and the workspace are read,
  // and out / lse written, only from here on.  Only then is the launch that follows (the fused append / suffix /
  // combine kernel) allowed to start: its own early work (KV append, suffix walk) then never overtakes the kernel in
  // front of this one, and it waits for this grid's completion itself before it reads the results written here.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  HG_PTRACE(threadIdx.x == 0, 2);

  // Register budget (setmaxnreg must sit inside the role branch it applies to): the producer warpgroup gives registers
  // back, the two softmax warpgroups (128 live fp32 scores per thread) take them: 128 x 56 + 256 x 224 = 384 x 168,
  // the launch-time allocation.
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
      // =============================== TMA producer ===========================================
      const bool leader = elect_one();
      uint32_t ri = 0;              // ring iterations issued (all pieces)
      uint32_t nq[kTiles] = {0, 0};  // Q loads issued per tile
      SchedIter it;
      SchedPiece sp;
      sched_begin(S, cta_id, it);
      while (sched_next(S, it, sp)) {
        Piece pc;
        resolve_piece<kCausal>(P, S, sp, pc);
        if (pc.n == 0) continue;
        const int n = pc.n;
        const int key0 = pc.k_row0 + pc.b_lo * BLOCK_N;
        const int ntile = pc.rows_left > BLOCK_M ? 2 : 1;
        const CUtensorMap* tk = &P.tmap_k[pc.level];
        const CUtensorMap* tv = &P.tmap_v[pc.level];
        for (int u = -2; u < n; ++u) {
          if (u == 0) {
            // Q after the two prologue K blocks: its smem is the previous piece's output staging tile
            for (int t = 0; t < ntile; ++t) {
              mbar_wait(&bars->q_empty[t], (nq[t] & 1) ^ 1);
              if (leader) {
                mbar_expect_tx(&bars->q_full[t], L::kQTileBytes);
#pragma unroll
                for (int h = 0; h < L::kHalves; ++h)
                  tma_load_2d(smem + L::kQ + t * L::kQTileBytes + h * L::kQHalfBytes, &P.tmap_q, pc.head * D + h * 64,
                              pc.q_row0 + t * BLOCK_M, &bars->q_full[t]);
              }
              ++nq[t];
            }
          }
          const bool has_v = u >= 0, has_k = u + 2 < n;
          if (!has_v && !has_k) continue;
          const int st = ri % kStages;
          uint8_t* base = smem + L::kKV + st * L::kStageBytes;
          mbar_wait(&bars->kv_empty[st], ((ri / kStages) & 1) ^ 1);
          if (leader) {
            mbar_expect_tx(&bars->kv_full[st], (has_v ? L::kKVTileBytes : 0) + (has_k ? L::kKVTileBytes : 0));
            if (has_k) {
#pragma unroll
              for (int h = 0; h < L::kHalves; ++h)
                tma_load_2d(base + L::kKVTileBytes + h * L::kKVHalfBytes, tk, pc.kvh * D + h * 64, key0 + (u + 2) * BLOCK_N,
                            &bars->kv_full[st]);
            }
            if (has_v) {
#pragma unroll
              for (int h = 0; h < L::kHalves; ++h)
                tma_load_2d(base + h * L::kKVHalfBytes, tv, pc.kvh * D + h * 64, key0 + u * BLOCK_N, &bars->kv_full[st]);
            }
          }
          ++ri;
        }
      }
    } else if (warp == 1 || warp == 3) {
      // =============================== MMA issuers (warp 1: tile A, warp 3: tile B) ==============
      // The whole warp walks the loop and the barriers (warp-uniform, so descriptors stay in uniform registers); one
      // elected lane issues tcgen05.mma / tcgen05.commit.  Per block j and tile t: P V of block j, then Q K^T of
      // block j+2 into the S buffer P_t(j) just vacated (same thread, same issue order: no barrier between them).
      // A warp whose tile holds no rows in a piece still walks the ring and releases its slots.
      const int t = warp >> 1;
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

      uint32_t ri = 0;  // ring iterations consumed (all pieces, mirrors the producer)
      uint32_t cnt[2] = {0, 0};  // uses of S buffer 0 / 1 of tile t so far (barrier phases); block j of a piece uses buffer j & 1
      uint32_t np = 0;  // pieces in which tile t was active
      SchedIter it;
      SchedPiece sp;
      sched_begin(S, cta_id, it);
      while (sched_next(S, it, sp)) {
        Piece pc;
        resolve_piece<kCausal>(P, S, sp, pc);
        if (pc.n == 0) continue;
        const int n = pc.n;
        const bool active = t == 0 || pc.rows_left > BLOCK_M;
        // prologue: S_t(0), S_t(1) -- the softmax warpgroup then always finds its next block ready
        for (int u = -2; u < 0; ++u) {
          if (u + 2 < n) {
            const int st = ri % kStages;
            mbar_wait(&bars->kv_full[st], (ri / kStages) & 1);
            if (active) {
              if (u == -2) mbar_wait(&bars->q_full[t], np & 1);
              tc_fence_after();
              if (leader) {
                const int sb = u + 2;  // block 0 -> buffer 0, block 1 -> buffer 1
                issue_qk(st, sb);
                umma_commit(&bars->s_full[t][sb]);
                umma_commit(&bars->kv_empty[st]);
              }
            } else if (leader) {
              mbar_arrive(&bars->kv_empty[st]);
            }
            __syncwarp();
            ++ri;
          }
        }
        // O_t of the previous piece must have been drained by its epilogue before P V overwrites it
        if (active) mbar_wait(&bars->o_empty[t], (np & 1) ^ 1);
        for (int j = 0; j < n; ++j) {
          const int st = ri % kStages;
          HG_PWAIT(t == 0 && lane == 0, 15, mbar_wait(&bars->kv_full[st], (ri / kStages) & 1));
          if (active) {
            const int b = j & 1;
            HG_PWAIT(t == 0 && lane == 0, 8, mbar_wait(&bars->p_full[t][b], (cnt[b] + (uint32_t)(j >> 1)) & 1));
            tc_fence_after();
            if (leader) {
              issue_pv(st, b, j == 0);
              umma_commit(&bars->pv_done[t]);
              if (j + 1 == n) umma_commit(&bars->o_full[t]);
              if (j + 2 < n) {
                issue_qk(st, b);
                umma_commit(&bars->s_full[t][b]);
              }
              umma_commit(&bars->kv_empty[st]);
            }
          } else if (leader) {
            mbar_arrive(&bars->kv_empty[st]);
          }
          __syncwarp();
          ++ri;
        }
        if (active) {
          cnt[0] += (uint32_t)((n + 1) >> 1);
          cnt[1] += (uint32_t)(n >> 1);
          ++np;
        }
      }
    }
  } else {
    // =============================== softmax warpgroups =====================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int t = (warp - 4) >> 2;      // tile owned by this warpgroup
    const int wq = warp & 3;            // == warp % 4: the TMEM lane quarter this warp may access
    const int row = wq * 32 + lane;     // row of the tile == TMEM lane
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    const uint32_t o_addr = tmem + lane_base + tmem_o(t);
    uint32_t* const ws_flags = P.ws_flags;
    const uint32_t epoch = ws_flags != nullptr ? *reinterpret_cast<volatile uint32_t*>(ws_flags + kWsWordEpoch) + 1u : 0u;
    volatile WarpCtx* const cx = reinterpret_cast<volatile WarpCtx*>(smem + L::kCtx + (warp - 4) * 128);
    auto park_iter = [&](const SchedIter& it) {
      cx->it.cur.unit = it.cur.unit; cx->it.cur.blk = it.cur.blk; cx->it.end.unit = it.end.unit; cx->it.end.blk = it.end.blk; cx->it.first = it.first;
    };
    {
      SchedIter it0;
      sched_begin(S, cta_id, it0);
      if (lane == 0) {
        park_iter(it0);
        cx->c0 = 0; cx->c1 = 0; cx->g = 0; cx->np = 0; cx->duty[0] = -1; cx->duty[1] = -1; cx->first_piece = 1;
      }
      __syncwarp();
    }
    for (;;) {
      SchedIter it;
      it.cur.unit = cx->it.cur.unit; it.cur.blk = cx->it.cur.blk; it.end.unit = cx->it.end.unit; it.end.blk = cx->it.end.blk; it.first = cx->it.first;
      SchedPiece sp;
      if (!sched_next(S, it, sp)) break;
      __syncwarp();
      if (lane == 0) {
        park_iter(it);
        cx->sp.unit = sp.unit; cx->sp.level = sp.level; cx->sp.head = sp.head; cx->sp.grp = sp.grp; cx->sp.mt = sp.mt;
        cx->sp.b_lo = sp.b_lo; cx->sp.b_hi = sp.b_hi; cx->sp.split = sp.split; cx->sp.slot = sp.slot;
      }
      __syncwarp();
      Piece pc;
      resolve_piece<kCausal>(P, S, sp, pc);
      const int rows_valid = min(BLOCK_M, pc.rows_left - t * BLOCK_M);
      if (rows_valid <= 0) {  // tile B of a unit with at most 128 rows: nothing to compute, but the merge duty is shared
        if (pc.split) {
          int other;
          if (split_role(S, pc.unit, cta_id, &other) == 2 && lane == 0) cx->duty[pc.slot] = pc.unit;
        }
        continue;
      }
      const int n = pc.n;
      float m_used = -INFINITY;  // raw-score max the exponentials are referenced to
      float l = 0.f;

      if (n > 0) {
        // block j of the piece uses S buffer j & 1; phases: one per use of a buffer (ph[b] = uses so far), one per key block
        // for pv_done (gpar), one per piece for o_full
        const uint32_t ph0 = cx->c0 & 1, ph1 = cx->c1 & 1, gpar = cx->g & 1;
        const uint32_t o_parity = cx->np & 1;
        // Software pipeline over key blocks.  Per block the warp issues 64 MUFU exp2; everything else it has to do --
        // the scale FFMAs, the row sums and the 16-bit packing of block j, and (once S_t(j+1) can have landed: its
        // Q K^T is only issued after P_t(j-1) was consumed) fetching the scores of block j+1 from TMEM and reducing
        // them to their row max -- is written interleaved with those MUFU requests in groups of 8, so that the
        // in-order warp always has independent work behind them.  Two score register arrays alternate between
        // "being exponentiated" and "being fetched".
        // row_end: first key (group-relative) this thread's row may NOT see; tile_end: the same for the tile's first row
        // Masks: block jl of the piece (group block b_lo + jl) holds a key some row of this tile must not see from
        // jl = first_mask on (the ragged last block of the group; causal: every block from the tile's diagonal on);
        // this thread's row then sees the first rem0 - 64 * jl keys of the block.
        int first_mask, rem0;
        {
          const int row_end = kCausal ? pc.mt * (kTiles * BLOCK_M) + t * BLOCK_M + row + pc.causal_off + 1 : 0x7fffffff;
          const int tile_end = kCausal ? pc.mt * (kTiles * BLOCK_M) + t * BLOCK_M + pc.causal_off + 1 : 0x7fffffff;
          const int ragged_at = (pc.k_len % BLOCK_N) != 0 ? pc.nb_group - 1 - pc.b_lo : 0x7fffffff;
          first_mask = min(ragged_at, kCausal ? tile_end / BLOCK_N - pc.b_lo : 0x7fffffff);
          rem0 = min(pc.k_len, row_end) - pc.b_lo * BLOCK_N;
        }
        auto needs_mask = [&](int jl) { return jl >= first_mask; };
        uint32_t sa[BLOCK_N], sb[BLOCK_N];
        float m_blk;
        auto mask_tail = [&](uint32_t(&x)[BLOCK_N], int jl) {
          const int rem = rem0 - jl * BLOCK_N;
#pragma unroll
          for (int c = 0; c < BLOCK_N; ++c)
            if (c >= rem) x[c] = 0xff800000u;  // -inf
        };
        auto max8 = [&](float* mx, const uint32_t(&x)[BLOCK_N], int gq) {
          mx[0] = fmaxf(mx[0], fmaxf(__uint_as_float(x[gq * 8 + 0]), __uint_as_float(x[gq * 8 + 1])));
          mx[1] = fmaxf(mx[1], fmaxf(__uint_as_float(x[gq * 8 + 2]), __uint_as_float(x[gq * 8 + 3])));
          mx[2] = fmaxf(mx[2], fmaxf(__uint_as_float(x[gq * 8 + 4]), __uint_as_float(x[gq * 8 + 5])));
          mx[3] = fmaxf(mx[3], fmaxf(__uint_as_float(x[gq * 8 + 6]), __uint_as_float(x[gq * 8 + 7])));
        };

        // cur: scores of block j (masked, max known in m_blk); nxt: receives block j+1.
        // kHasNext: block j+1 exists; kMaskNext: it needs a mask.  (Compile-time: as warp-uniform run-time flags -- two
        // copies of the body instead of six -- every block took ~100 cycles longer, r02k.)
        auto body = [&](int j, uint32_t(&cur)[BLOCK_N], uint32_t(&nxt)[BLOCK_N], auto has_next_tag, auto mask_next_tag, auto buf_tag) {
          constexpr bool kHasNext = decltype(has_next_tag)::value;
          constexpr bool kMaskNext = decltype(mask_next_tag)::value;
          constexpr int kB = decltype(buf_tag)::value;  // == j & 1: S buffer of this block
          const uint32_t p_addr = tmem + lane_base + tmem_s(t, kB);
          const float m_new = fmaxf(m_used, m_blk);
          if (j == 0) {
            m_used = m_new;
          } else {
            const bool need = (m_new - m_used) * scale_log2 > kRescaleThreshold;
            if (__any_sync(0xffffffffu, need)) {
              // rare: O_t must be complete up to P_t(j-1) V_{j-1} before it is rescaled in place; P_t(j)
              // has not been released yet, so no later MMA can be touching O_t.
              mbar_wait(&bars->pv_done[t], gpar ^ 1u ^ (uint32_t)kB);  // block j - 1 of the piece
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
          auto exp8 = [&](int gq) {        // in place: score -> p
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
              float x0, x1;
              unpack_f2(ffma2(pack_f2(__uint_as_float(cur[gq * 8 + c]), __uint_as_float(cur[gq * 8 + c + 1])), scale2, neg2), x0, x1);
              cur[gq * 8 + c] = __float_as_uint(fast_exp2(x0));
              cur[gq * 8 + c + 1] = __float_as_uint(fast_exp2(x1));
            }
          };
          auto sum_pack8 = [&](int gq) {
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
              const float p0 = __uint_as_float(cur[gq * 8 + c]), p1 = __uint_as_float(cur[gq * 8 + c + 1]);
              ps2[(c >> 1) & 1] = fadd2(ps2[(c >> 1) & 1], pack_f2(p0, p1));
              pk[(gq * 8 + c) >> 1] = pack2<T>(p0, p1);
            }
          };
#pragma unroll
          for (int gq = 0; gq < 6; ++gq) {
            exp8(gq);
            if (gq >= 2) sum_pack8(gq - 2);
          }
          HG_TMEM_ST16(p_addr, pk, 0);  // keys 0..31 of P
          if constexpr (kHasNext) {     // S_t(j+1) has had ~3/4 of this block's MUFU time to land
            HG_PWAIT(t == 0 && row == 0, 14, mbar_wait(&bars->s_full[t][kB ^ 1], (((uint32_t)(j >> 1) + kB) & 1) ^ (kB ? ph0 : ph1)));
            tc_fence_after();
            const uint32_t s_addr = tmem + lane_base + tmem_s(t, kB ^ 1);
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
            for (int gq = 0; gq < 4; ++gq) max8(mx, nxt, gq);
          }
          sum_pack8(6);
          sum_pack8(7);
          HG_TMEM_ST16(p_addr + 16, pk, 16);
          if constexpr (kHasNext) {
#pragma unroll
            for (int gq = 4; gq < 8; ++gq) max8(mx, nxt, gq);
            m_blk = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
          }
          {
            float a0, a1;
            unpack_f2(fadd2(ps2[0], ps2[1]), a0, a1);
            l += a0 + a1;
          }
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive(&bars->p_full[t][kB]);
        };

        {  // prologue: scores and row max of the piece's first block
          mbar_wait(&bars->s_full[t][0], ph0);
          HG_PTRACE(t == 0 && row == 0, cx->first_piece ? 3 : 6);
#if defined(HG_PREFIX_TRACE)
          if (t == 0 && row == 0 && cx->first_piece) g_prefix_trace[cta_id * 16 + 12] = clock64();
#endif
          tc_fence_after();
          const uint32_t s_addr = tmem + lane_base + tmem_s(t, 0);
          HG_TMEM_LD32(s_addr + 0, sa, 0);
          HG_TMEM_LD32(s_addr + 32, sa, 32);
          tmem_wait_ld();
          if (needs_mask(0)) mask_tail(sa, 0);
          float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int gq = 0; gq < 8; ++gq) max8(mx, sa, gq);
          m_blk = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        }
        for (int j = 0; j < n; ++j) {
          const bool last = j + 1 == n, mask_next = !last && needs_mask(j + 1);
          using B0 = std::integral_constant<int, 0>;
          using B1 = std::integral_constant<int, 1>;
          if ((j & 1) == 0) {
            if (last) body(j, sa, sb, std::false_type{}, std::false_type{}, B0{});
            else if (mask_next) body(j, sa, sb, std::true_type{}, std::true_type{}, B0{});
            else body(j, sa, sb, std::true_type{}, std::false_type{}, B0{});
          } else {
            if (last) body(j, sb, sa, std::false_type{}, std::false_type{}, B1{});
            else if (mask_next) body(j, sb, sa, std::true_type{}, std::true_type{}, B1{});
            else body(j, sb, sa, std::true_type{}, std::false_type{}, B1{});
          }
        }
        mbar_wait(&bars->o_full[t], o_parity);
        tc_fence_after();
      }

      // ---- epilogue of the piece: everything about it is re-derived from the parked copy ---------------------
      SchedPiece sq;
      sq.unit = cx->sp.unit; sq.level = cx->sp.level; sq.head = cx->sp.head; sq.grp = cx->sp.grp; sq.mt = cx->sp.mt;
      sq.b_lo = cx->sp.b_lo; sq.b_hi = cx->sp.b_hi; sq.split = cx->sp.split; sq.slot = cx->sp.slot;
      Piece pe;
      resolve_piece<kCausal>(P, S, sq, pe);
      const int n_e = pe.n;
      const int rows_e = min(BLOCK_M, pe.rows_left - t * BLOCK_M);
      const bool first_piece = cx->first_piece != 0;
      HG_PTRACE(t == 0 && row == 0, first_piece ? 4 : 7);
#if defined(HG_PREFIX_TRACE)
      if (t == 0 && row == 0 && first_piece) g_prefix_trace[cta_id * 16 + 13] = clock64();
#endif
      const int tile_row0 = pe.q_row0 + t * BLOCK_M;
      T* const out = reinterpret_cast<T*>(P.lv[pe.level].out);
      float* const lse = P.lv[pe.level].lse;
      float m_log2 = n_e > 0 ? m_used * scale_log2 : -INFINITY;  // reference max of this piece's exponentials, log2 units
      // A split unit: which CTAs hold its pieces, and which one am I?
      //   2 pieces (the common case: a cut of the stream-K schedule falls inside the unit): the CTA of the HEAD piece
      //     -- its last piece -- owns the unit: it folds the tail piece's partial (written long ago: a tail piece is
      //     the first thing its CTA does) into its own accumulator, still in TMEM, and finishes the unit as if whole.
      //   more pieces (few units on many SMs: the head-parallel ranks of a TP run): every piece leaves a partial and
      //     the pieces' CTAs merge the unit together after their main work, each a slice of the rows (below).
      bool to_workspace = false;
      if (pe.split) {
        int other = 0;
        const int role = split_role(S, pe.unit, cta_id, &other);
        if (role == 1) {
          const uint32_t* flag = ws_flags + kWsWordFlags + other * 2 + t;
          while ((int32_t)(ld_acquire_gpu(flag) - epoch) < 0) {
          }
          const float* slot = P.ws_part + (int64_t)other * ws_slot_floats(D);
          const float2 ml = ld_cg_f2(slot + ws_ml_index(D, t, row));
          if (__any_sync(0xffffffffu, ml.y > 0.f)) {
            // O_t <- w_own * O_t + w_part * partial, in place in TMEM (n_e > 0 here: the head piece of a unit whose tail
            // holds keys holds keys itself); the normal epilogue below then finishes the unit
            const float m_all = fmaxf(m_log2, ml.y > 0.f ? ml.x : -INFINITY);
            const float w_own = fast_exp2(m_log2 - m_all), w_part = ml.y > 0.f ? fast_exp2(ml.x - m_all) : 0.f;
            l = w_own * l + w_part * ml.y;
            m_log2 = m_all;
            float4 pa[8], pb[8];
            auto fetch = [&](float4(&dst)[8], int c0) {
#pragma unroll
              for (int c = 0; c < 8; ++c) dst[c] = ld_cg_f4(slot + ws_o_index(D, t, row, (c0 >> 2) + c));
            };
            auto fold = [&](const float4(&src)[8], int c0) {
              uint32_t o[32];
              HG_TMEM_LD32(o_addr + c0, o, 0);
              tmem_wait_ld();
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                o[4 * c + 0] = __float_as_uint(__uint_as_float(o[4 * c + 0]) * w_own + src[c].x * w_part);
                o[4 * c + 1] = __float_as_uint(__uint_as_float(o[4 * c + 1]) * w_own + src[c].y * w_part);
                o[4 * c + 2] = __float_as_uint(__uint_as_float(o[4 * c + 2]) * w_own + src[c].z * w_part);
                o[4 * c + 3] = __float_as_uint(__uint_as_float(o[4 * c + 3]) * w_own + src[c].w * w_part);
              }
              HG_TMEM_ST32(o_addr + c0, o, 0);
            };
            fetch(pa, 0);
#pragma unroll
            for (int c0 = 0; c0 < D; c0 += 64) {
              fetch(pb, c0 + 32);
              fold(pa, c0);
              if (c0 + 64 < D) fetch(pa, c0 + 64);
              fold(pb, c0 + 32);
            }
            tmem_wait_st();
          }
        } else {
          to_workspace = true;
          if (role == 2 && lane == 0) cx->duty[pe.slot] = pe.unit;
        }
      }
      if (to_workspace) {
        // Partial result -> workspace slot of this CTA: unnormalised O row (fp32), then (max in log2 units, row sum).
        // An empty piece leaves sum = 0 and is skipped by whoever merges.
        float* slot = P.ws_part + (int64_t)(cta_id * 2 + pe.slot) * ws_slot_floats(D);
        if (n_e > 0) {
#pragma unroll
          for (int c0 = 0; c0 < D; c0 += 32) {
            uint32_t o[32];
            HG_TMEM_LD32(o_addr + c0, o, 0);
            tmem_wait_ld();
#pragma unroll
            for (int c = 0; c < 32; c += 4)
              *reinterpret_cast<uint4*>(slot + ws_o_index(D, t, row, (c0 + c) >> 2)) = make_uint4(o[c], o[c + 1], o[c + 2], o[c + 3]);
          }
          tc_fence_before();
          mbar_arrive(&bars->o_empty[t]);
          if (row == 0) mbar_arrive(&bars->q_empty[t]);  // every MMA of the piece is complete: Q_t is dead
        }
        *reinterpret_cast<float2*>(slot + ws_ml_index(D, t, row)) = make_float2(m_log2, n_e > 0 ? l : 0.f);
        __threadfence();
        if (t == 0) named_bar_sync<1, BLOCK_M>(); else named_bar_sync<2, BLOCK_M>();
        if (row == 0) st_release_gpu(ws_flags + kWsWordFlags + (cta_id * 2 + pe.slot) * 2 + t, epoch);
      } else {
        const float inv_l = (l > 0.f) ? 1.f / l : 0.f;
        auto final8 = [&](const uint32_t(&o)[32], int c) {  // columns c .. c + 8 of the chunk -> normalised 16-bit
          uint4 w;
          w.x = pack2<T>(__uint_as_float(o[c + 0]) * inv_l, __uint_as_float(o[c + 1]) * inv_l);
          w.y = pack2<T>(__uint_as_float(o[c + 2]) * inv_l, __uint_as_float(o[c + 3]) * inv_l);
          w.z = pack2<T>(__uint_as_float(o[c + 4]) * inv_l, __uint_as_float(o[c + 5]) * inv_l);
          w.w = pack2<T>(__uint_as_float(o[c + 6]) * inv_l, __uint_as_float(o[c + 7]) * inv_l);
          return w;
        };
        if (n_e > 0 && rows_e == BLOCK_M) {
          // full tile: rows -> the (dead) Q_t tile in the TMA 128-byte swizzle -> one bulk store per
          // 64-column half.  Row r keeps 16-byte chunk c at chunk slot c ^ (r & 7).
          uint8_t* stage = smem + L::kQ + t * L::kQTileBytes;
#pragma unroll
          for (int c0 = 0; c0 < D; c0 += 32) {
            uint32_t o[32];
            HG_TMEM_LD32(o_addr + c0, o, 0);
            tmem_wait_ld();
#pragma unroll
            for (int c = 0; c < 32; c += 8) {
              const int chunk = (c0 + c) >> 3;  // 16-byte chunk of the row
              uint8_t* dst = stage + (chunk >> 3) * L::kQHalfBytes + row * 128 + (((chunk & 7) ^ (row & 7)) << 4);
              *reinterpret_cast<uint4*>(dst) = final8(o, c);
            }
          }
          tc_fence_before();
          mbar_arrive(&bars->o_empty[t]);
          fence_proxy_async();
          if (t == 0) named_bar_sync<1, BLOCK_M>(); else named_bar_sync<2, BLOCK_M>();
          if (row == 0) {
#pragma unroll
            for (int h = 0; h < L::kHalves; ++h) tma_store_2d(&P.tmap_o[pe.level], stage + h * L::kQHalfBytes, pe.head * D + h * 64, tile_row0);
            bulk_commit();
            bulk_wait_read();  // the staging tile has been read: the producer may load the next Q_t over it
            mbar_arrive(&bars->q_empty[t]);
          }
        } else {
          const bool row_ok = row < rows_e;
          T* orow = out + ((int64_t)(tile_row0 + row) * hq + pe.head) * D;
          if (n_e > 0) {
#pragma unroll
            for (int c0 = 0; c0 < D; c0 += 32) {
              uint32_t o[32];
              HG_TMEM_LD32(o_addr + c0, o, 0);
              tmem_wait_ld();
              if (row_ok) {
#pragma unroll
                for (int c = 0; c < 32; c += 8) st_v4(orow + c0 + c, final8(o, c));
              }
            }
            tc_fence_before();
            mbar_arrive(&bars->o_empty[t]);
            if (row == 0) mbar_arrive(&bars->q_empty[t]);
          } else if (row_ok) {  // a group without keys: out = 0 (lse = -inf below)
#pragma unroll
            for (int c = 0; c < D; c += 8) st_v4(orow + c, make_uint4(0, 0, 0, 0));
          }
        }
        if (row < rows_e && lse != nullptr)
          lse[(int64_t)(tile_row0 + row) * hq + pe.head] = (l > 0.f) ? (m_log2 + fast_log2(l)) * kLn2 : -INFINITY;
      }
      HG_PTRACE(t == 0 && row == 0, first_piece ? 5 : 8);
      __syncwarp();
      if (lane == 0) {
        cx->first_piece = 0;
        if (n_e > 0) {
          cx->c0 = cx->c0 + (uint32_t)((n_e + 1) >> 1);
          cx->c1 = cx->c1 + (uint32_t)(n_e >> 1);
          cx->g = cx->g + (uint32_t)n_e;
          cx->np = cx->np + 1u;
        }
      }
      __syncwarp();
    }

    // ---- merge duty (units cut into more than two pieces): this CTA's share of the rows -----------------------
    // (the pieces of one unit finish at about the same time on their CTAs -- every CTA's range has the same cost -- so
    // the flags polled here are already up or about to be; nothing below depends on a CTA that is itself waiting)
    if (ws_flags != nullptr) {
      for (int dty = 0; dty < 2; ++dty) {
        const int unit = cx->duty[dty];
        if (unit < 0 || (dty == 1 && unit == cx->duty[0])) continue;
        SchedPiece su;
        sched_decode_unit(S, unit, su);
        su.b_lo = su.b_hi = su.split = su.slot = 0;
        Piece pu;
        resolve_piece<kCausal>(P, S, su, pu);
        merge_duty<T, D>(S, P.ws_part, ws_flags, epoch, unit, (warp - 4) * 32 + lane, lane, reinterpret_cast<T*>(P.lv[pu.level].out),
                         P.lv[pu.level].lse, pu.q_row0, min(kTiles * BLOCK_M, pu.rows_left), pu.head, hq, cta_id);
      }
    }
    HG_PTRACE(t == 0 && row == 0, 9);
    if (row == 0) bulk_wait_all();  // this thread's TMA stores have reached global memory before the grid retires
    tc_fence_before();
  }

  // ---- teardown ----------------------------------------------------------------------------
  __syncthreads();
  HG_PTRACE(threadIdx.x == 0, 10);
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
  }
  if (threadIdx.x == 0 && P.ws_flags != nullptr) {
    // last CTA out: the next launch on this workspace sees the next epoch (it reads it after griddepcontrol.wait,
    // i.e. after this grid has retired)
    uint32_t* f = P.ws_flags;
    const uint32_t e = *reinterpret_cast<volatile uint32_t*>(f + kWsWordEpoch);
    __threadfence();
    if (atomicAdd(f + kWsWordExit, 1u) == gridDim.x * gridDim.y - 1) {
      f[kWsWordExit] = 0u;
      __threadfence();
      f[kWsWordEpoch] = e + 1u;
    }
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

template <typename T, int D, bool kCausal>
static int launch_prefix_one(const PrefixKernelParams& kp, const SchedParams& sched, cudaStream_t s) {
  using L = SmemLayout<D>;
  const int smem_bytes = L::kTotal + 1024;
  static bool attr_set[64] = {};  // per instantiation and device; idempotent, racing threads set the same value
  const int dev = device_info().device;
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(prefix_attn_sm100_kernel<T, D, kCausal>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return set_error(HG_ERR_CUDA, "prefix: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)sched.n_ctas, 1, 1);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, prefix_attn_sm100_kernel<T, D, kCausal>, kp);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(HG_ERR_CUDA, "prefix_attn_sm100: launch failed: %s", cudaGetErrorString(e));
  }
  return check_launch("prefix_attn_sm100");
}

template <typename T, int D, bool kCausal>
static int launch_prefix_inst(const PrefixParams& p, const SchedParams& sched, int dtype, cudaStream_t s) {
  using L = SmemLayout<D>;
  PrefixKernelParams kp;
  memset(&kp, 0, sizeof(kp));
  int rc;
  if ((rc = make_tmap(&kp.tmap_q, p.q, dtype, (uint64_t)p.n_q_rows, (uint64_t)p.hq * D, (uint64_t)p.q_stride_row, BLOCK_M)) != HG_OK) return rc;
  for (int l = 0; l < p.n_levels; ++l) {
    const PrefixLevel& lv = p.levels[l];
    if ((rc = make_tmap(&kp.tmap_k[l], lv.k, dtype, (uint64_t)lv.n_k_rows, (uint64_t)p.hkv * D, (uint64_t)lv.kv_stride_row, BLOCK_N)) != HG_OK) return rc;
    if ((rc = make_tmap(&kp.tmap_v[l], lv.v, dtype, (uint64_t)lv.n_k_rows, (uint64_t)p.hkv * D, (uint64_t)lv.kv_stride_row, BLOCK_N)) != HG_OK) return rc;
    if ((rc = make_tmap(&kp.tmap_o[l], lv.out, dtype, (uint64_t)p.n_q_rows, (uint64_t)p.hq * D, (uint64_t)p.hq * D, BLOCK_M)) != HG_OK) return rc;
    kp.lv[l].out = lv.out;
    kp.lv[l].lse = lv.lse;
    kp.lv[l].cu = lv.cu_seqlens_k;
    kp.lv[l].k_len = lv.k_len;
  }
  kp.sched = sched;
  kp.hkv = p.hkv;
  kp.scale_log2 = p.scale_log2;
  if (sched.mode == 0) {
    kp.ws_flags = reinterpret_cast<uint32_t*>(p.workspace);
    kp.ws_part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(p.workspace) + kWsFlagBytes);
  }
  return launch_prefix_one<T, D, kCausal>(kp, sched, s);
}

#ifdef HG_PREFIX_TRACE
extern "C" int hg_debug_prefix_trace(long long* host_buf, int n) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(host_buf, g_prefix_trace, sizeof(long long) * (size_t)n);
}
#endif

int64_t prefix_workspace_bytes() { return ws_bytes(128); }

// HYDRAGEN_B200_PREFIX_CTAS (development knob, read once): cap on the persistent grid (default: the SM count)
static int prefix_cta_cap() {
  static const int v = [] {
    const char* e = getenv("HYDRAGEN_B200_PREFIX_CTAS");
    return e != nullptr ? atoi(e) : 0;
  }();
  return v;
}

// HYDRAGEN_B200_PREFIX_SPLIT_OVERHEAD (development knob, read once): cost of cutting units, in key blocks (default 6);
// 0 = always stream-K, a huge value = never
static long long prefix_split_overhead() {
  static const long long v = [] {
    const char* e = getenv("HYDRAGEN_B200_PREFIX_SPLIT_OVERHEAD");
    return e != nullptr ? atoll(e) : 6ll;
  }();
  return v;
}

// The schedule of one launch: levels laid end to end on the cost axis, grid sized so that every CTA gets at least one
// minimal piece and no unit is cut into more pieces than the merge handles.
int build_prefix_schedule(const PrefixParams& p, int n_sms, int split_mode, SchedParams* out) {
  SchedParams S;
  memset(&S, 0, sizeof(S));
  S.n_levels = p.n_levels;
  S.hq = p.hq;
  S.c0 = 2;         // prologue + epilogue of a unit, in key blocks (~1.6k cycles each)
  S.min_piece = 4;  // no piece shorter than this many key blocks
  long long cost = 0;
  int units = 0, w_max = 1;
  for (int l = 0; l < p.n_levels; ++l) {
    const PrefixLevel& lv = p.levels[l];
    SchedLevel& L = S.lv[l];
    if (lv.n_groups < 1 || p.n_q_rows % lv.n_groups != 0)
      return set_error(HG_ERR_INVALID_ARGUMENT, "prefix: level %d: %d groups do not divide %lld query rows", l, lv.n_groups, (long long)p.n_q_rows);
    L.n_groups = lv.n_groups;
    L.q_per_group = (int)(p.n_q_rows / lv.n_groups);
    L.tiles_per_group = (L.q_per_group + kTiles * BLOCK_M - 1) / (kTiles * BLOCK_M);
    const int kmax = lv.cu_seqlens_k != nullptr ? lv.max_k_len : lv.k_len;
    L.nb_max = (kmax + BLOCK_N - 1) / BLOCK_N;
    const long long nu = (long long)L.n_groups * L.tiles_per_group * p.hq;
    if (nu + units > 0x3fffffffLL) return set_error(HG_ERR_UNSUPPORTED, "prefix: too many work units");
    L.n_units = (int)nu;
    L.unit0 = units;
    L.cost0 = cost;
    units += L.n_units;
    cost += nu * (L.nb_max + S.c0);
    if (L.nb_max + S.c0 > w_max) w_max = L.nb_max + S.c0;
  }
  S.total_units = units;
  S.total_cost = cost;
  int cap = n_sms > 0 ? n_sms : 148;
  if (cap > kMaxCtas) cap = kMaxCtas;
  if (prefix_cta_cap() > 0 && prefix_cta_cap() < cap) cap = prefix_cta_cap();
  // Whole units dealt round-robin (mode 1) or stream-K (mode 0)?  Cutting units costs: a cut unit's tail piece writes
  // its partial to the workspace, the head piece reads it back and both pay a pipeline ramp -- measured (r02g-r02i
  // traces) at about 6 key blocks' worth of time on the CTA that finishes last.  Stream-K therefore only when the
  // whole-unit makespan exceeds the stream-K share by more than that: few units on many SMs (long prefixes on the
  // head-parallel ranks of a TP run, cfg#5) or a ragged last wave of many units; NOT the 128 units of cfg#2 on 148 SMs
  // (32 vs 27.7 + 6 blocks).
  bool split = false;
  const int n_whole = std::max(1, std::min(cap, units));
  if (split_mode == 2) split = true;
  if (split_mode == 1) {
    long long worst = 0;
    for (int c = 0; c < n_whole; ++c) {
      long long sum = 0;
      for (int l = 0; l < p.n_levels; ++l) {
        const SchedLevel& L = S.lv[l];
        // units u of the level with u % n_whole == c
        const long long first = ((c - L.unit0) % n_whole + n_whole) % n_whole;
        if (first < L.n_units) sum += ((L.n_units - 1 - first) / n_whole + 1) * (long long)(L.nb_max + S.c0);
      }
      worst = std::max(worst, sum);
    }
    const long long g0 = std::min<long long>(cap, std::max<long long>(1, cost / (S.c0 + S.min_piece)));
    split = worst > cost / g0 + prefix_split_overhead();
  }
  if (split) {
    S.mode = 0;
    long long g = cap;
    g = std::min<long long>(g, std::max<long long>(1, cost / (S.c0 + S.min_piece)));
    // pieces per unit <= w_max / (cost / g) + 2  must stay within kMaxUnitPieces
    g = std::min<long long>(g, std::max<long long>(1, (long long)(kMaxUnitPieces - 3) * cost / w_max));
    S.n_ctas = (int)std::max<long long>(1, g);
    sched_fill_bounds(S);
  } else {
    S.mode = 1;
    S.n_ctas = n_whole;
  }
  *out = S;
  return HG_OK;
}

int launch_prefix(const PrefixParams& p, int dtype, cudaStream_t s) {
  if (p.n_levels < 1 || p.n_levels > kMaxLevels) return set_error(HG_ERR_INVALID_ARGUMENT, "prefix: %d shared levels (1..%d)", p.n_levels, kMaxLevels);
  if (p.n_q_rows == 0) return HG_OK;
  if (dtype != HG_F16 && dtype != HG_BF16)
    return set_error(HG_ERR_UNSUPPORTED, "prefix: the tcgen05 kernel takes f16/bf16 only (dtype %d)", dtype);
  if (p.q_stride_row % 8 != 0 || reinterpret_cast<uintptr_t>(p.q) % 16 != 0) return set_error(HG_ERR_UNSUPPORTED, "prefix: TMA needs 16-byte aligned bases and row strides");
  for (int l = 0; l < p.n_levels; ++l) {
    const PrefixLevel& lv = p.levels[l];
    if (lv.kv_stride_row % 8 != 0 || reinterpret_cast<uintptr_t>(lv.k) % 16 != 0 || reinterpret_cast<uintptr_t>(lv.v) % 16 != 0 ||
        reinterpret_cast<uintptr_t>(lv.out) % 16 != 0)
      return set_error(HG_ERR_UNSUPPORTED, "prefix: TMA needs 16-byte aligned bases and row strides");
    if (lv.n_groups < 1 || p.n_q_rows % lv.n_groups != 0)
      return set_error(HG_ERR_INVALID_ARGUMENT, "prefix: level %d: %d groups do not divide %lld query rows", l, lv.n_groups, (long long)p.n_q_rows);
  }
  if (p.d != 64 && p.d != 128) return set_error(HG_ERR_UNSUPPORTED, "prefix: head_dim %d not supported (64 or 128)", p.d);
  const int kv_splits = p.kv_splits < 1 ? 1 : p.kv_splits;
  if (p.causal) {
    const PrefixLevel& lv = p.levels[0];
    if (p.n_levels != 1 || lv.cu_seqlens_k != nullptr || kv_splits > 1 || lv.k_len < p.n_q_rows / lv.n_groups)
      return set_error(HG_ERR_UNSUPPORTED, "prefix: the causal form takes one level of uniform groups with k_len >= q rows per group and no kv split");
  }
  if (p.hq > 65535) return set_error(HG_ERR_UNSUPPORTED, "prefix: hq > 65535");
  // ONE level: the one-CTA-per-unit kernel (prefix_unit_sm100.cu) -- per key block it is 10-20 % faster than the
  // persistent one below (r02l-r02s), whose single launch only pays when it covers several shared levels.
  // HYDRAGEN_B200_PREFIX_PERSISTENT=1 forces the persistent kernel (tests, measurements).
  static const bool force_persistent = [] {
    const char* e = getenv("HYDRAGEN_B200_PREFIX_PERSISTENT");
    return e != nullptr && e[0] == '1';
  }();
  if (p.causal || kv_splits > 1 || (p.n_levels == 1 && !force_persistent)) return launch_prefix_unit(p, kv_splits, dtype, s);
  const int allow_split = (p.workspace != nullptr && p.workspace_bytes >= ws_bytes(p.d)) ? 1 : 0;
  SchedParams sched;
  int rc = build_prefix_schedule(p, device_info().sm_count, allow_split, &sched);
  if (rc != HG_OK) return rc;
  if (dtype == HG_BF16) {
    if (p.d == 128) return launch_prefix_inst<__nv_bfloat16, 128, false>(p, sched, dtype, s);
    return launch_prefix_inst<__nv_bfloat16, 64, false>(p, sched, dtype, s);
  }
  if (p.d == 128) return launch_prefix_inst<__half, 128, false>(p, sched, dtype, s);
  return launch_prefix_inst<__half, 64, false>(p, sched, dtype, s);
}

}  // namespace hg
