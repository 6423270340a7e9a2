// Row-parallel o_proj GEMM fused with the all-reduce that follows it: the step after the attention
// hot path under head-axis tensor parallelism -- hydragen/llama.py:592-594 (o_proj of the local heads) followed by
// funcol.all_reduce in hydragen/tp.py:108-112 -- as ONE persistent kernel per rank:
//
//   phase 1 (tcgen05)  out_partial[m, n] = sum_k x[m, k] * w[n, k]   x = local attention output [M, K] (K = local heads * d),
//                      w = this rank's column slice of o_proj.weight [N, K], both K-major.  128 x 128 output tiles (128 x 256
//                      for the GEMM alone), 64-wide k-blocks through a 5- (3-) deep TMA ring, fp32 accumulators double-buffered
//                      in TMEM, epilogue TMEM -> 16-bit -> swizzled smem -> TMA store into this rank's SYMMETRIC buffer.
//                      When a tile's store has completed and a gpu-scope fence has passed, its producer raises one flag word in
//                      the memory of the rank that OWNS the tile (tile % world): "rank r's partial of tile t, call e, is in place".
//                      Warps 0-3 epilogue, 4 TMA producer, 5 MMA issuer.
//   phase 2 (NVLS)     runs BESIDE phase 1 on 2-4 more warps per CTA, each on its own: walk slices (a few rows) of the tiles
//                      this rank owns, in the order the tiles are produced; wait until all `world` flags of the slice's tile
//                      carry the call's epoch, multimem.ld_reduce the slice (the switch fetches it from every rank and adds in
//                      fp32), multimem.st the sum into every rank's buffer.  The next slice's reductions are issued before
//                      the current slice's stores (the two load opposite link directions).
//   end                the CTAs count themselves out; the last one tells every peer "my slices are written everywhere",
//                      waits for the same from them, and bumps the epoch -- the grid retires only when `out` is complete.
//
// Against a library GEMM followed by the stand-alone collective (csrc/allreduce.cu) this removes one launch, the
// collective's start barrier (the per-tile flags ARE the "inputs in place" hand-shake, at tile granularity), and lets the
// reduction of the first tiles start while later tiles are still being multiplied.  The traffic through the switch is the
// same (1 + 1/N) x message per direction, which bounds the whole at the collective's wire time.
//
// All CTAs of the grid must be resident together (phase 2 of one rank waits for phase 1 of the others): the grid is at
// most one CTA per SM, and the kernel's shared memory footprint keeps it at one CTA per SM.
//
// world == 1: phase 1 only, into an ordinary buffer (the tcgen05 GEMM on its own; used by the single-GPU parity tests).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "oproj_sched.h"
#include "sm100_ptx.cuh"

namespace hg {

namespace {

constexpr int BM = kOprojBM, BK = 64;
constexpr int kGemmWarps = 6;    // warps 0-3 epilogue (TMEM lane quarter = warp), 4 TMA, 5 MMA
constexpr int kMaxReduceWarps = 8;  // warps 6...: phase 2, running beside phase 1 (how many is a launch parameter)
constexpr int kThreads = (kGemmWarps + kMaxReduceWarps) * 32;
constexpr int kABytes = BM * BK * 2;    // one TMA box: 128 rows x 128 B
constexpr int kBoxBytes = BM * 64 * 2;  // one output box: 128 rows x 64 columns

// BN = 256: three 48 KiB ring slots; BN = 128: five 32 KiB slots.  Narrow tiles cost a third more operand traffic but
// finish in more, shorter waves -- which is what lets the reduction start early when the whole product is one wave wide.
// U = reductions in flight per lane and slice (and as many again prefetched)
template <int BN, int U>
struct Cfg {
  static constexpr int kStages = BN == 256 ? 3 : 5;
  static constexpr int kBBytes = BN * BK * 2;  // one TMA box: BN rows x 128 B
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingBytes = BM * BN * 2;  // BN / 64 output boxes
  static constexpr int kOffStaging = kStages * kStageBytes;
  static constexpr int kOffBars = kOffStaging + kStagingBytes;
  static constexpr int kSmemBytes = kOffBars + 256 + 1024;  // + alignment slack
  static constexpr uint32_t kTmemCols = 2 * BN;             // two 128 x BN fp32 accumulators
  static constexpr int kLanesPerRow = BN / 8;               // 16-byte vectors in a tile row
  static constexpr int kRowsPerInst = 32 / kLanesPerRow;    // tile rows one warp-wide reduction covers
  static constexpr int kUnitRows = U * kRowsPerInst;       // rows of a tile one warp reduces at a time
  static constexpr int kUnitsPerTile = BM / kUnitRows;
};
constexpr int kMaxStages = 5;

// flag words of a rank (uint32, zeroed once, in its own symmetric flag array; written remotely, polled at home)
constexpr int kWEpoch = 0;    // calls completed by this rank (local)
constexpr int kWDone = 1;     // CTAs of the running call that have finished phase 2 (local)
constexpr int kWOut = 32;     // + p: "rank p's slices of call e are written everywhere"
constexpr int kWReady = 128;  // + tile * world + p: "rank p's partial of tile `tile`, call e, is in place" (owner's copy only)

struct Barriers {
  uint64_t full[kMaxStages], empty[kMaxStages];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
  uint32_t last;
};

__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
template <typename T>
__device__ __forceinline__ uint4 mc_ld_reduce(const uint4* p);
template <>
__device__ __forceinline__ uint4 mc_ld_reduce<__nv_bfloat16>(const uint4* p) {
  uint4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
template <>
__device__ __forceinline__ uint4 mc_ld_reduce<__half>(const uint4* p) {
  uint4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.f16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void mc_st(uint4* p, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

#ifdef HG_OPROJ_TRACE
// Development aid: %globaltimer stamps [CTA][stage] of the most recent call.  0 entry, 1 dependency wait passed,
// 2 first accumulator complete, 3 last tile stored and signalled, 4 first slice's flags seen (reduce warp 0), 5 reduce warp 0
// done, 6 fence done, 7 (last CTA) end barrier passed, 8 first tile signalled, 9 reduce warp 0: first slice multicast,
// 10 MMA warp done
__device__ long long g_oproj_trace[160 * 16];
__device__ __forceinline__ void op_stamp(int stage) {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  g_oproj_trace[blockIdx.x * 16 + stage] = t;
}
#define HG_OSTAMP(cond, stage) \
  do {                         \
    if (cond) op_stamp(stage); \
  } while (0)
#else
#define HG_OSTAMP(cond, stage) \
  do {                         \
  } while (0)
#endif

}  // namespace

template <typename T, int BN, int kU>
__global__ void __launch_bounds__(kThreads, 1)
    oproj_allreduce_sm100_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                                 const __grid_constant__ CUtensorMap tmap_out, uint4* __restrict__ mc,
                                 uint32_t* const* __restrict__ flags, int rank, int world, int M, int N, int K, int signal_mode) {
  using C = Cfg<BN, kU>;
  constexpr int kFmt = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;
  constexpr uint32_t kIdesc = make_idesc(kFmt, 0, BM, BN);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  Barriers* bars = reinterpret_cast<Barriers*>(smem + C::kOffBars);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (M + BM - 1) / BM, n_tiles = m_tiles * ((N + BN - 1) / BN), n_kb = (K + BK - 1) / BK;

  HG_OSTAMP(threadIdx.x == 0, 0);
  if (warp == 4 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_out) : "memory");
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < C::kStages; ++i) {
      mbar_init(&bars->full[i], 1);
      mbar_init(&bars->empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->acc_full[i], 1);
      mbar_init(&bars->acc_empty[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(C::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&bars->tmem_base);
  // Possibly a programmatic dependent of the kernel that wrote x (or of the previous call on the same flags, whose last
  // CTA bumps the epoch): nothing of either is read before that grid has retired.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  HG_OSTAMP(threadIdx.x == 0, 1);
  uint32_t* mine = world > 1 ? flags[rank] : nullptr;
  const uint32_t e = world > 1 ? ld_volatile_u32(mine + kWEpoch) + 1u : 0u;  // stable for the whole call

  if (warp == 4) {
    // ====================================== phase 1: TMA producer ========================================================
    if (elect_one()) {
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int mt = tile % m_tiles, nt = tile / m_tiles;
        for (int kb = 0; kb < n_kb; ++kb, ++it) {
          const int st = it % C::kStages;
          mbar_wait(&bars->empty[st], ((it / C::kStages) & 1) ^ 1);
          mbar_expect_tx(&bars->full[st], C::kStageBytes);
          tma_load_2d(smem + st * C::kStageBytes, &tmap_x, kb * BK, mt * BM, &bars->full[st]);
          tma_load_2d(smem + st * C::kStageBytes + kABytes, &tmap_w, kb * BK, nt * BN, &bars->full[st]);
        }
      }
    }
  } else if (warp == 5) {
    // ====================================== phase 1: MMA issuer ==========================================================
    const bool leader = elect_one();
    constexpr uint32_t kHi = desc_hi(1024);  // SWIZZLE_128B, K-major: 8-row groups 1024 B apart
    const uint32_t a_lo0 = desc_lo(smem_u32(smem), 0), b_lo0 = desc_lo(smem_u32(smem + kABytes), 0);
    int it = 0, lt = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
      const int acc = lt & 1;
      mbar_wait(&bars->acc_empty[acc], ((lt >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem + (uint32_t)acc * BN;
      for (int kb = 0; kb < n_kb; ++kb, ++it) {
        const int st = it % C::kStages;
        mbar_wait(&bars->full[st], (it / C::kStages) & 1);
        tc_fence_after();
        if (leader) {
          const uint32_t a_lo = a_lo0 + st * (C::kStageBytes >> 4), b_lo = b_lo0 + st * (C::kStageBytes >> 4);
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) umma_ss(d_tmem, a_lo + kk * 2, kHi, b_lo + kk * 2, kHi, kIdesc, (kb > 0 || kk > 0) ? 1u : 0u);
          umma_commit(&bars->empty[st]);
          if (kb + 1 == n_kb) umma_commit(&bars->acc_full[acc]);
        }
        __syncwarp();
      }
    }
    HG_OSTAMP(lane == 0, 10);
  } else if (warp < 4) {
    // ====================================== phase 1: epilogue ============================================================
    uint8_t* staging = smem + C::kOffStaging;
    const int row = warp * 32 + lane;
    int lt = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
      const int mt = tile % m_tiles, nt = tile / m_tiles, acc = lt & 1;
      mbar_wait(&bars->acc_full[acc], (lt >> 1) & 1);
      tc_fence_after();
      HG_OSTAMP(threadIdx.x == 0 && lt == 0, 2);
      const uint32_t t_addr = tmem + (uint32_t)acc * BN + ((uint32_t)(warp * 32) << 16);
      // accumulator -> 16-bit -> the TMA 128-byte swizzle: row r keeps 16-byte chunk c of a 64-column box at slot c ^ (r & 7).
      // (Storing straight from the registers -- one row per lane, every thread fencing its own stores -- was measured instead,
      // r02zi: the flag leaves 1.5 us LATER and the pair is 1.5-3 us slower; the bulk store is the fast way out.)
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t o[32];
        HG_TMEM_LD32(t_addr + c0, o, 0);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 32; c += 8) {
          uint4 w;
          w.x = pack2<T>(__uint_as_float(o[c + 0]), __uint_as_float(o[c + 1]));
          w.y = pack2<T>(__uint_as_float(o[c + 2]), __uint_as_float(o[c + 3]));
          w.z = pack2<T>(__uint_as_float(o[c + 4]), __uint_as_float(o[c + 5]));
          w.w = pack2<T>(__uint_as_float(o[c + 6]), __uint_as_float(o[c + 7]));
          const int chunk = (c0 + c) >> 3;
          *reinterpret_cast<uint4*>(staging + (chunk >> 3) * kBoxBytes + row * 128 + (((chunk & 7) ^ (row & 7)) << 4)) = w;
        }
      }
      tc_fence_before();
      mbar_arrive(&bars->acc_empty[acc]);  // the MMA warp may start the tile after next in this accumulator
      fence_proxy_async();
      named_bar_sync<1, 128>();
      if (threadIdx.x == 0) {
#pragma unroll
        for (int b = 0; b < BN / 64; ++b)
          if (nt * BN + b * 64 < N) tma_store_2d(&tmap_out, staging + b * kBoxBytes, nt * BN + b * 64, mt * BM);
        bulk_commit();
        bulk_wait_read();
      }
      named_bar_sync<1, 128>();  // staging may be overwritten
      if (threadIdx.x == 0 && world > 1) {
        // The tile must be in this GPU's L2 -- the point of coherence the switch's reads go through -- before the flag
        // leaves.  Completion of the bulk store alone does not guarantee that (r02zc: stale tiles read), a gpu-scope fence
        // does; a SYSTEM-scope release is not needed for data that stays at home, and costs dearly here: it queues behind
        // the multicast stores the reduce warps next door have in flight (r02za trace: 12 us from accumulator to flag).
        bulk_wait_all();
        asm volatile("fence.proxy.async.global;" ::: "memory");
        if (signal_mode == 1) __threadfence();
        if (signal_mode == 2) __threadfence_system();
        st_relaxed_sys(flags[oproj_owner(tile, world)] + kWReady + tile * world + rank, e);
        HG_OSTAMP(lt == 0, 8);
      }
    }
    if (threadIdx.x == 0) {
      bulk_wait_all();
      HG_OSTAMP(true, 3);
    }
  } else if (world > 1) {
    // ====================================== phase 2: reduce the slices this rank owns ====================================
    // Runs BESIDE phase 1 on warps of its own: slices become reducible tile by tile, in the order the tiles are produced
    // (every rank walks the tiles in the same order), so the switch starts working while later tiles are still being
    // multiplied.  Each warp is on its own -- no CTA-wide barrier: lanes < world poll the tile's flags, the warp converges,
    // then every lane keeps kU reductions in flight and issues the next slice's before it multicasts this one's (the two
    // load opposite link directions).
    const OprojGeom geo = oproj_geom(M, N, BN, kU, world, rank);  // who reduces what: oproj_sched.h
    const int n_units = geo.n_units;
    const int64_t row_vecs = (int64_t)N / 8;  // 16-byte vectors per output row
    const int stride = oproj_unit_stride(gridDim.x, (int)(blockDim.x >> 5) - kGemmWarps);
    int polled = -1;
    // all flags of the slice's tile carry this call's epoch (">= e": a fast peer may already be in the next call)
    auto wait_ready = [&](int tile) {
      if (tile == polled) return;
      if (lane < world) {
        const uint32_t* f = mine + kWReady + tile * world + lane;
        while ((int32_t)(ld_acquire_sys(f) - e) < 0) __nanosleep(64);
      }
      __syncwarp();
      polled = tile;
    };
    // slice u: unit_rows rows x BN columns of an owned tile; instruction j of the warp covers rows_per_inst of its rows
    auto unit_addr = [&](int u, int j, bool& ok) -> uint4* {
      int r, c;
      ok = oproj_unit_vector(geo, u, j, lane, &r, &c);
      return mc + (int64_t)r * row_vecs + (c >> 3);
    };
    // warp w of CTA b starts at slice w * gridDim.x + b: the slices of the earliest tiles are spread over all SMs
    int u = oproj_first_unit(blockIdx.x, warp - kGemmWarps, gridDim.x);
    uint4 cur[kU], nxt[kU];
    if (u < n_units) {
      wait_ready(oproj_unit_tile(geo, u));
      HG_OSTAMP(warp == kGemmWarps && lane == 0, 4);
#pragma unroll
      for (int j = 0; j < kU; ++j) {
        bool ok;
        uint4* p = unit_addr(u, j, ok);
        if (ok) cur[j] = mc_ld_reduce<T>(p);
      }
    }
    while (u < n_units) {
      const int u2 = u + stride;
      if (u2 < n_units) {
        wait_ready(oproj_unit_tile(geo, u2));
#pragma unroll
        for (int j = 0; j < kU; ++j) {
          bool ok;
          uint4* p = unit_addr(u2, j, ok);
          if (ok) nxt[j] = mc_ld_reduce<T>(p);
        }
      }
#pragma unroll
      for (int j = 0; j < kU; ++j) {
        bool ok;
        uint4* p = unit_addr(u, j, ok);
        if (ok) mc_st(p, cur[j]);
      }
#pragma unroll
      for (int j = 0; j < kU; ++j) cur[j] = nxt[j];
      HG_OSTAMP(warp == kGemmWarps && lane == 0 && u < stride, 9);
      u = u2;
    }
    HG_OSTAMP(warp == kGemmWarps && lane == 0, 5);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(C::kTmemCols) : "memory");
  }
  if (world == 1) return;

  // ====================================== end: every slice of every rank has landed =======================================
  if (threadIdx.x == 0) {
    __threadfence_system();  // this CTA's multicast stores are performed (acknowledged by every replica) before it is counted
    HG_OSTAMP(true, 6);
    bars->last = (atomicAdd(mine + kWDone, 1u) == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (bars->last == 0u) return;
  if ((int)threadIdx.x < world) {
    __threadfence();
    st_relaxed_sys(flags[threadIdx.x] + kWOut + rank, e);
    while ((int32_t)(ld_acquire_sys(mine + kWOut + threadIdx.x) - e) < 0) {
    }
  }
  __syncthreads();
  HG_OSTAMP(threadIdx.x == 0, 7);
  if (threadIdx.x == 0) {
    mine[kWDone] = 0u;
    __threadfence();
    mine[kWEpoch] = e;
  }
}

#ifdef HG_OPROJ_TRACE
extern "C" int hg_debug_oproj_trace(long long* host_buf, int n) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(host_buf, g_oproj_trace, sizeof(long long) * (size_t)n);
}
#endif

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D view [rows, cols] of a 16-bit matrix, row stride in elements; boxes of box_rows x 64 columns in the 128-byte swizzle;
// reads past either extent return zero, writes past them are dropped.
static int make_tmap(CUtensorMap* map, const void* base, int dtype, uint64_t rows, uint64_t cols, uint64_t row_stride, uint32_t box_rows) {
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(device_info().encode_tiled);
  if (fn == nullptr) return set_error(HG_ERR_NOT_INITIALIZED, "oproj_allreduce: cuTensorMapEncodeTiled unavailable (call hg_init first)");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {row_stride * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, dtype == HG_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base),
                  gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(HG_ERR_CUDA, "oproj_allreduce: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
  return HG_OK;
}

// Development knobs, read at every launch (so that one process can sweep them; a captured graph keeps what it was captured with):
//   HYDRAGEN_B200_OPROJ_BN = 128 | 256   output tile width.  Default: 128 whenever there is a collective (tiles finish in more,
//                                        shorter rounds, so the reduction of the first overlaps the multiplication of the rest:
//                                        r02zf / r02zg, 3-13 % faster than 256 on 2 and on 8 GPUs), 256 for the GEMM alone
//   HYDRAGEN_B200_OPROJ_RWARPS = 1..8    reduce warps per CTA
//   HYDRAGEN_B200_OPROJ_U = 1 | 2 | 4    reductions per lane and slice; a warp has 2 U x 512 bytes in flight.  What matters is the
//                                        product CTAs x warps x 2 U = bytes of OUTPUT in flight per GPU, each of which the switch
//                                        assembles from `world` reads: enough to cover the round trip, not so much that the
//                                        multicast stores -- which load the OPPOSITE link direction -- only start when the queue of
//                                        reductions has drained.  Measured optimum (r02zf, r02zg): ~2.3 MiB on 2 GPUs, ~0.6 MiB on 8
//                                        (148 KiB: latency-bound, 1.5-3.5x slower); the default follows 4.6 MiB / world
//   HYDRAGEN_B200_OPROJ_SIGNAL = 0|1|2   fence between a tile's bulk store and its flag: none / gpu scope (default) / system scope.
//                                        Measured on 2 GPUs (r02zc, 40 checked rounds per variant): without a fence the switch reads
//                                        STALE tiles (14 bad rounds of 80); with the gpu-scope fence none, at +1 us; system scope
//                                        costs 5-8 us
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e != nullptr ? atoi(e) : dflt;
}
static int oproj_signal_mode() { return env_int("HYDRAGEN_B200_OPROJ_SIGNAL", 1); }
static int oproj_reduce_warps(int world) {
  const int w = env_int("HYDRAGEN_B200_OPROJ_RWARPS", 0);
  return (w >= 1 && w <= kMaxReduceWarps) ? w : (world <= 2 ? 4 : 2);
}
static int oproj_u(int world) {
  const int u = env_int("HYDRAGEN_B200_OPROJ_U", 0);
  return (u == 1 || u == 2 || u == 4) ? u : (world <= 4 ? 4 : 2);
}
static int pick_bn(int64_t m, int64_t n, int world) {
  const int b = env_int("HYDRAGEN_B200_OPROJ_BN", 0);
  if (b == 128 || b == 256) return b;
  (void)m;
  (void)n;
  return world > 1 ? 128 : 256;
}

int oproj_allreduce_flag_words(int64_t m, int64_t n, int world) {
  const int64_t tiles = ((m + BM - 1) / BM) * ((n + 127) / 128);  // the narrow tiling: an upper bound for either
  return (int)(kWReady + tiles * world);
}

// Host-side view of a launch (hg_oproj_allreduce_plan): the geometry launch_oproj_allreduce would choose for this shape
// and, replayed from the device formulas, how often each 16-byte vector of the output is reduced by rank `rank`.
int oproj_allreduce_plan(int64_t m, int64_t n, int world, int rank, int n_ctas, int* geometry_out, int32_t* cover_out) {
  const int n_sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  const int bn = pick_bn(m, n, world), u = world > 1 ? oproj_u(world) : 4, warps = world > 1 ? oproj_reduce_warps(world) : 0;
  const int64_t n_tiles = ((m + BM - 1) / BM) * ((n + bn - 1) / bn);
  int grid = world > 1 ? n_sms : (int)std::min<int64_t>(n_tiles, n_sms);
  if (n_ctas > 0) grid = std::min(n_ctas, n_sms);
  const OprojGeom g = oproj_geom((int)m, (int)n, bn, u, world, rank);
  if (geometry_out != nullptr) {
    const int v[8] = {bn, u, warps, grid, g.n_tiles, g.n_own, g.n_units, kWReady + g.n_tiles * world};
    for (int i = 0; i < 8; ++i) geometry_out[i] = v[i];
  }
  if (cover_out != nullptr && world > 1) {
    const int stride = oproj_unit_stride(grid, warps);
    for (int cta = 0; cta < grid; ++cta)
      for (int w = 0; w < warps; ++w)
        for (int unit = oproj_first_unit(cta, w, grid); unit < g.n_units; unit += stride)
          for (int j = 0; j < u; ++j)
            for (int lane = 0; lane < 32; ++lane) {
              int r, c;
              if (oproj_unit_vector(g, unit, j, lane, &r, &c)) cover_out[(int64_t)r * (n / 8) + c / 8] += 1;
            }
  }
  return HG_OK;
}

template <typename T, int BN, int U>
static int launch_inst(const OprojParams& p, cudaStream_t s) {
  using C = Cfg<BN, U>;
  CUtensorMap tx, tw, to;
  int rc;
  if ((rc = make_tmap(&tx, p.x, p.dtype, (uint64_t)p.m, (uint64_t)p.k, (uint64_t)p.x_stride_row, BM)) != HG_OK) return rc;
  if ((rc = make_tmap(&tw, p.w, p.dtype, (uint64_t)p.n, (uint64_t)p.k, (uint64_t)p.w_stride_row, BN)) != HG_OK) return rc;
  if ((rc = make_tmap(&to, p.out, p.dtype, (uint64_t)p.m, (uint64_t)p.n, (uint64_t)p.n, BM)) != HG_OK) return rc;
  auto kern = oproj_allreduce_sm100_kernel<T, BN, U>;
  static bool attr_set[64] = {};
  const int dev = device_info().device;
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) return set_error(HG_ERR_CUDA, "oproj_allreduce: cannot reserve %d bytes of shared memory: %s", C::kSmemBytes, cudaGetErrorString(e));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const int n_sms = device_info().sm_count;
  const int64_t n_tiles = ((p.m + BM - 1) / BM) * ((p.n + BN - 1) / BN);
  // one CTA per SM at most: all CTAs must be co-resident (phase 2 waits for the other ranks' phase 1)
  int grid = p.world > 1 ? n_sms : (int)std::min<int64_t>(n_tiles, n_sms);
  if (p.n_ctas > 0) grid = std::min(p.n_ctas, n_sms);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)((kGemmWarps + (p.world > 1 ? oproj_reduce_warps(p.world) : 0)) * 32));
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tx, tw, to, reinterpret_cast<uint4*>(p.out_mc),
                                     reinterpret_cast<uint32_t* const*>(p.flags_dev), p.rank, p.world, (int)p.m, (int)p.n, (int)p.k, oproj_signal_mode());
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(HG_ERR_CUDA, "oproj_allreduce: launch failed: %s", cudaGetErrorString(e));
  }
  return check_launch("oproj_allreduce");
}

template <typename T>
static int launch_type(const OprojParams& p, cudaStream_t s) {
  const int bn = pick_bn(p.m, p.n, p.world), u = p.world > 1 ? oproj_u(p.world) : 4;
  if (bn == 128) {
    if (u == 1) return launch_inst<T, 128, 1>(p, s);
    if (u == 2) return launch_inst<T, 128, 2>(p, s);
    return launch_inst<T, 128, 4>(p, s);
  }
  if (u == 1) return launch_inst<T, 256, 1>(p, s);
  if (u == 2) return launch_inst<T, 256, 2>(p, s);
  return launch_inst<T, 256, 4>(p, s);
}

int launch_oproj_allreduce(const OprojParams& p, cudaStream_t s) {
  if (p.dtype == HG_BF16) return launch_type<__nv_bfloat16>(p, s);
  if (p.dtype == HG_F16) return launch_type<__half>(p, s);
  return set_error(HG_ERR_UNSUPPORTED, "oproj_allreduce: 16-bit types only (dtype %d)", p.dtype);
}

}  // namespace hg
