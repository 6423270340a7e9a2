// All-reduce(sum) over NVLink / NVSwitch with in-switch reduction (NVLS): the collective that follows the
// attention hot path under head-axis tensor parallelism -- the sum of the row-parallel o_proj partials,
// funcol.all_reduce in hydragen/tp.py:108-112 of the reference -- as one small kernel per rank instead of a
// NCCL ring: the [B, hidden] bf16 message of a decode step is 4-20 MiB, i.e. latency- not bandwidth-bound.
//
// Every rank's buffer lives in symmetric memory and is also mapped at one MULTICAST address.  Rank r owns
// slice r of the buffer:
//   multimem.ld_reduce  [mc + i]  -> the switch fetches element i from every rank and adds them (fp32 accumulate)
//   multimem.st         [mc + i]  -> the switch writes the sum into every rank's buffer
// so each GPU moves 1/N of the message once in each direction, whatever N is.  Two cross-rank barriers
// bracket that (inputs ready / all slices written); they are per-CTA flag exchanges through peer pointers:
// CTA b of rank r raises flag (b, r) in every peer's flag array and waits for (b, p) from every peer p.
// Flags reset themselves (compare-and-swap 0->1 to raise, 1->0 to consume), so the kernel can be replayed
// from a CUDA graph with no host involvement.
#include "common.cuh"

namespace hg {

namespace {

__device__ __forceinline__ void rank_barrier(uint32_t* const* flags, int rank, int world) {
  __syncthreads();
  if ((int)threadIdx.x < world) {
    const int p = threadIdx.x;
    __threadfence_system();  // release: everything this CTA wrote (multimem.st included) before the flag
    uint32_t* theirs = flags[p] + (size_t)blockIdx.x * world + rank;
    while (atomicCAS_system(theirs, 0u, 1u) != 0u) {
    }
    uint32_t* mine = flags[rank] + (size_t)blockIdx.x * world + p;
    while (atomicCAS_system(mine, 1u, 0u) != 1u) {
    }
    __threadfence_system();  // acquire
  }
  __syncthreads();
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

}  // namespace

template <int DTYPE>
__global__ void __launch_bounds__(512) allreduce_multimem_kernel(uint4* __restrict__ mc, uint32_t* const* __restrict__ flags, int rank,
                                                                 int world, int64_t n_vec) {
  rank_barrier(flags, rank, world);  // every rank's input is in place (its earlier kernels on the stream have retired)
  const int64_t per = (n_vec + world - 1) / world;
  const int64_t begin = rank * per, end = min(n_vec, begin + per);
  // a reduction is a round trip through the switch: keep U of them in flight per thread
  constexpr int U = 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += stride * U) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i + u * stride < end) v[u] = mc_ld_reduce<DTYPE>(mc + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i + u * stride < end) mc_st(mc + i + u * stride, v[u]);
  }
  rank_barrier(flags, rank, world);  // every slice has been written into this rank's buffer
}

// One-shot form: every rank reduces the WHOLE message through the switch into a private output buffer.  One
// barrier instead of two and no multicast store; N times the switch reductions, which is the better trade
// while the message is small enough to be latency-bound.  Out of place by construction (peers may still be
// reading this rank's input); the input may be overwritten once a later call of either form has returned.
template <int DTYPE>
__global__ void __launch_bounds__(512) allreduce_multimem_oneshot_kernel(const uint4* __restrict__ mc, uint4* __restrict__ out,
                                                                         uint32_t* const* __restrict__ flags, int rank, int world,
                                                                         int64_t n_vec) {
  rank_barrier(flags, rank, world);
  constexpr int U = 4;
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
}

int launch_allreduce_multimem(void* mc_ptr, void* out, const void* flags_dev, int rank, int world, int64_t nbytes, int dtype,
                              int n_blocks, cudaStream_t s) {
  if (out != nullptr) {
    const int64_t nv = nbytes / 16;
    uint32_t* const* fl = reinterpret_cast<uint32_t* const*>(flags_dev);
    switch (dtype) {
      case HG_BF16: allreduce_multimem_oneshot_kernel<HG_BF16><<<n_blocks, 512, 0, s>>>((const uint4*)mc_ptr, (uint4*)out, fl, rank, world, nv); break;
      case HG_F16: allreduce_multimem_oneshot_kernel<HG_F16><<<n_blocks, 512, 0, s>>>((const uint4*)mc_ptr, (uint4*)out, fl, rank, world, nv); break;
      case HG_F32: allreduce_multimem_oneshot_kernel<HG_F32><<<n_blocks, 512, 0, s>>>((const uint4*)mc_ptr, (uint4*)out, fl, rank, world, nv); break;
      default: return set_error(HG_ERR_INVALID_ARGUMENT, "allreduce: unknown dtype %d", dtype);
    }
    return check_launch("allreduce_multimem_oneshot");
  }
  const int64_t n_vec = nbytes / 16;
  uint32_t* const* flags = reinterpret_cast<uint32_t* const*>(flags_dev);
  switch (dtype) {
    case HG_BF16: allreduce_multimem_kernel<HG_BF16><<<n_blocks, 512, 0, s>>>((uint4*)mc_ptr, flags, rank, world, n_vec); break;
    case HG_F16: allreduce_multimem_kernel<HG_F16><<<n_blocks, 512, 0, s>>>((uint4*)mc_ptr, flags, rank, world, n_vec); break;
    case HG_F32: allreduce_multimem_kernel<HG_F32><<<n_blocks, 512, 0, s>>>((uint4*)mc_ptr, flags, rank, world, n_vec); break;
    default: return set_error(HG_ERR_INVALID_ARGUMENT, "allreduce: unknown dtype %d", dtype);
  }
  return check_launch("allreduce_multimem");
}

}  // namespace hg
