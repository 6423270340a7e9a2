// All-reduce(sum) over NVLink / NVSwitch with in-switch reduction (NVLS): the collective that follows the
// attention hot path under head-axis tensor parallelism -- the sum of the row-parallel o_proj partials,
// funcol.all_reduce in hydragen/tp.py:108-112 of the reference -- as one small kernel per rank instead of a
// NCCL ring: the [B, hidden] bf16 message of a decode step is 4-20 MiB, i.e. latency- not bandwidth-bound.
//
// Every rank's buffer lives in symmetric memory and is also mapped at one MULTICAST address.  Rank r owns
// slice r of the buffer:
//   multimem.ld_reduce  [mc + i]  -> the switch fetches element i from every rank and adds them (fp32 accumulate)
//   multimem.st         [mc + i]  -> the switch writes the sum into every rank's buffer
// so each GPU moves 1/N of the message once in each direction, whatever N is.
//
// Cross-rank synchronisation (round 2; the round-1 form -- every CTA compare-and-swapping one flag per peer in
// remote memory, twice per call -- cost 11 us per barrier, r01n): EPOCH flags, written remotely, polled at home.
//   flag words of a rank (uint32, zero-initialised once, all in its own symmetric flag array):
//     [0] epoch      number of collectives completed by this rank (local, bumped by the last CTA of each call)
//     [1] cta_done   CTAs of the running call that have finished their stores (local)
//     [32 + p]       "rank p's input of call e is in place"        written by rank p: st.release.sys  e
//     [64 + p]       "rank p's slice of call e is written everywhere"  written by rank p: st.release.sys  e
//   start:  one thread per peer (CTA 0) raises [32 + rank] in every rank's array; every CTA polls its OWN copy
//           (ld.acquire.sys on local memory) until all `world` words hold the call's epoch.  No reset, no remote spin,
//           one one-way NVLink store on the critical path.
//   end:    CTAs count themselves out (fence.sys + atomicAdd on a local word) and EXIT; only the last one raises
//           [64 + rank] everywhere, waits until its own copy is complete, and bumps the epoch.  The grid therefore
//           retires -- and the next kernel on the stream may read the buffer -- only after every slice has landed.
// Everything is replayable from a CUDA graph (the epoch lives in device memory).
#include <cstdlib>

#include "common.cuh"

namespace hg {

namespace {

constexpr int kWordEpoch = 0, kWordDone = 1, kWordIn = 32, kWordOut = 64;

#ifdef HG_AR_TRACE
// Development aid (never in the shipped library): %globaltimer stamps [CTA][stage] of the most recent call.
// stages: 0 entry, 1 after griddepcontrol.wait, 2 start barrier passed, 3 reductions issued and stored, 4 fence.sys done,
// 5 (last CTA) peers signalled, 6 (last CTA) end barrier passed
__device__ long long g_ar_trace[256 * 8];
__device__ __forceinline__ void ar_stamp(int stage) {
  if (threadIdx.x == 0 && blockIdx.x < 256) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_ar_trace[blockIdx.x * 8 + stage] = t;
  }
}
#else
__device__ __forceinline__ void ar_stamp(int) {}
#endif

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_volatile(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <int DTYPE>
__device__ __forceinline__ uint4 mc_ld_reduce(const uint4* p);
template <>
__device__ __forceinline__ uint4 mc_ld_reduce<HG_BF16>(const uint4* p) {
  uint4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
template <>
__device__ __forceinline__ uint4 mc_ld_reduce<HG_F16>(const uint4* p) {
  uint4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.f16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
template <>
__device__ __forceinline__ uint4 mc_ld_reduce<HG_F32>(const uint4* p) {
  uint4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void mc_st(uint4* p, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// "inputs in place": returns the epoch of this call
__device__ __forceinline__ uint32_t start_barrier(uint32_t* const* flags, int rank, int world) {
  uint32_t* mine = flags[rank];
  ar_stamp(0);
  // this launch may be a programmatic dependent of the kernel that produced the input -- or of the previous
  // collective on the same flags, whose last CTA bumps the epoch: nothing is read before that grid has retired
  asm volatile("griddepcontrol.wait;" ::: "memory");
  ar_stamp(1);
  const uint32_t e = ld_volatile(mine + kWordEpoch) + 1u;  // stable for the whole call: only the last CTA to leave writes it
  // relaxed: the input was written by EARLIER kernels of this rank, which have retired -- their writes are in this
  // GPU's L2, the point of coherence the peers' (and the switch's) reads go through.  A release here would cost a
  // system-scope fence (~2 us, r02d trace) for nothing.
  if (blockIdx.x == 0 && (int)threadIdx.x < world) st_relaxed_sys(flags[threadIdx.x] + kWordIn + rank, e);
  if ((int)threadIdx.x < world) {
    // ">= e": after a one-shot call a fast peer may already have announced the NEXT call
    while ((int32_t)(ld_acquire_sys(mine + kWordIn + threadIdx.x) - e) < 0) {
    }
  }
  __syncthreads();
  ar_stamp(2);
  // the kernel that follows on the stream may start its own set-up on the SMs this grid leaves free
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  return e;
}

// CTAs that are not the last one of their rank to finish exit right away; the last one
// completes the cross-rank hand-shake for the whole rank.
__device__ __forceinline__ void finish(uint32_t* const* flags, int rank, int world, uint32_t e, bool cross_rank) {
  __shared__ uint32_t s_last;
  __syncthreads();
  ar_stamp(3);
  uint32_t* mine = flags[rank];
  if (threadIdx.x == 0) {
    __threadfence_system();  // this CTA's multicast stores are performed before it is counted
    ar_stamp(4);
    s_last = (atomicAdd(mine + kWordDone, 1u) == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last == 0u) return;
  if ((int)threadIdx.x < world && cross_rank) {
    // Every CTA of this rank drained its multicast stores (fence.sys: acknowledged by every replica) BEFORE it was
    // counted, and this CTA saw the full count: the data is in place everywhere, so the flags need no further
    // system-scope fence (r02d: a second fence.sys + st.release.sys here cost 4.4 us).
    __threadfence();
    st_relaxed_sys(flags[threadIdx.x] + kWordOut + rank, e);
    ar_stamp(5);
    while ((int32_t)(ld_acquire_sys(mine + kWordOut + threadIdx.x) - e) < 0) {
    }
  }
  __syncthreads();
  ar_stamp(6);
  if (threadIdx.x == 0) {
    mine[kWordDone] = 0u;
    __threadfence();
    mine[kWordEpoch] = e;
  }
}

}  // namespace

template <int DTYPE, int U>
__global__ void __launch_bounds__(512) allreduce_multimem_kernel(uint4* __restrict__ mc, uint32_t* const* __restrict__ flags, int rank,
                                                                 int world, int64_t n_vec) {
  const uint32_t e = start_barrier(flags, rank, world);
  const int64_t per = (n_vec + world - 1) / world;
  const int64_t begin = rank * per, end = min(n_vec, begin + per);
  // A reduction is a round trip through the switch: every thread keeps U of them in flight, and issues the NEXT
  // batch before it multicasts the current one.  The two halves load opposite NVLink directions (ld_reduce: every
  // rank's memory feeds the switch, S bytes out per GPU; multimem.st: the switch fans the sums out, S bytes in per
  // GPU), so run back to back in lock-step on all ranks (r02c trace: 18-22 us for 8 MiB on 2 GPUs = the two phases
  // one after the other at link speed) they each leave one direction idle; interleaved they overlap.
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint4 cur[U], nxt[U];
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (i + u * stride < end) cur[u] = mc_ld_reduce<DTYPE>(mc + i + u * stride);
  while (i < end) {
    const int64_t i2 = i + stride * U;
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i2 + u * stride < end) nxt[u] = mc_ld_reduce<DTYPE>(mc + i2 + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i + u * stride < end) mc_st(mc + i + u * stride, cur[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) cur[u] = nxt[u];
    i = i2;
  }
  finish(flags, rank, world, e, true);
}

// One-shot form: every rank reduces the WHOLE message through the switch into a private output buffer.  One
// cross-rank wait instead of two and no multicast store; N times the switch reductions, which is the better trade
// only for very few ranks.  Out of place by construction (peers may still be reading this rank's input); the input
// may be overwritten once a later call of either form has passed its start barrier on every rank.
template <int DTYPE, int U>
__global__ void __launch_bounds__(512) allreduce_multimem_oneshot_kernel(const uint4* __restrict__ mc, uint4* __restrict__ out,
                                                                         uint32_t* const* __restrict__ flags, int rank, int world,
                                                                         int64_t n_vec) {
  const uint32_t e = start_barrier(flags, rank, world);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride * U) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i + u * stride < n_vec) v[u] = mc_ld_reduce<DTYPE>(mc + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i + u * stride < n_vec) out[i + u * stride] = v[u];
  }
  finish(flags, rank, world, e, false);
}

#ifdef HG_AR_TRACE
extern "C" int hg_debug_ar_trace(long long* host_buf, int n) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(host_buf, g_ar_trace, sizeof(long long) * (size_t)n);
}
#endif

// HYDRAGEN_B200_AR_UNROLL = 2 | 4 | 8 (read once; development knob): reductions in flight per thread and batch
static int ar_unroll() {
  static const int v = [] {
    const char* e = getenv("HYDRAGEN_B200_AR_UNROLL");
    const int u = e != nullptr ? atoi(e) : 4;
    return (u == 2 || u == 8) ? u : 4;
  }();
  return v;
}

template <typename Kern, typename... Args>
static int launch_pdl(Kern kern, int n_blocks, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)n_blocks);
  cfg.blockDim = dim3(512);
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, args...);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(HG_ERR_CUDA, "allreduce_multimem: launch failed: %s", cudaGetErrorString(e));
  }
  return check_launch("allreduce_multimem");
}

template <int DTYPE>
static int launch_dtype(void* mc_ptr, void* out, uint32_t* const* flags, int rank, int world, int64_t n_vec, int n_blocks,
                        cudaStream_t s) {
  const int u = ar_unroll();
  if (out != nullptr) {
    if (u == 2) return launch_pdl(allreduce_multimem_oneshot_kernel<DTYPE, 2>, n_blocks, s, (const uint4*)mc_ptr, (uint4*)out, flags, rank, world, n_vec);
    if (u == 8) return launch_pdl(allreduce_multimem_oneshot_kernel<DTYPE, 8>, n_blocks, s, (const uint4*)mc_ptr, (uint4*)out, flags, rank, world, n_vec);
    return launch_pdl(allreduce_multimem_oneshot_kernel<DTYPE, 4>, n_blocks, s, (const uint4*)mc_ptr, (uint4*)out, flags, rank, world, n_vec);
  }
  if (u == 2) return launch_pdl(allreduce_multimem_kernel<DTYPE, 2>, n_blocks, s, (uint4*)mc_ptr, flags, rank, world, n_vec);
  if (u == 8) return launch_pdl(allreduce_multimem_kernel<DTYPE, 8>, n_blocks, s, (uint4*)mc_ptr, flags, rank, world, n_vec);
  return launch_pdl(allreduce_multimem_kernel<DTYPE, 4>, n_blocks, s, (uint4*)mc_ptr, flags, rank, world, n_vec);
}

int launch_allreduce_multimem(void* mc_ptr, void* out, const void* flags_dev, int rank, int world, int64_t nbytes, int dtype,
                              int n_blocks, cudaStream_t s) {
  const int64_t n_vec = nbytes / 16;
  uint32_t* const* flags = reinterpret_cast<uint32_t* const*>(flags_dev);
  switch (dtype) {
    case HG_BF16: return launch_dtype<HG_BF16>(mc_ptr, out, flags, rank, world, n_vec, n_blocks, s);
    case HG_F16: return launch_dtype<HG_F16>(mc_ptr, out, flags, rank, world, n_vec, n_blocks, s);
    case HG_F32: return launch_dtype<HG_F32>(mc_ptr, out, flags, rank, world, n_vec, n_blocks, s);
    default: return set_error(HG_ERR_INVALID_ARGUMENT, "allreduce: unknown dtype %d", dtype);
  }
}

}  // namespace hg
