// n-way log-sum-exp combine of partial attention results.
//
// Replaces hydragen/attention.py:21-174 (combine_lse_torch, combine_lse_kernel/combine_lse_triton,
// combine_lse) of the reference: out = sum_i w_i out_i / sum_i w_i with w_i = exp(lse_i - max lse).
// One kernel for any fan-in 1..HG_MAX_COMBINE (the reference's Triton kernel takes exactly two
// inputs and hierarchies with >= 2 shared levels fall back to ~8 eager torch launches) and any
// head_dim (the reference kernel's column mask is a global bound, so d = 63 / 129 spill into the
// next row; here rows are exact).
//
// HBM-bound: (n+1) * rows * d * sizeof(T) + n * rows * 4 bytes per call, no reuse -> no shared
// memory; 128-bit loads/stores, all n input vectors of a chunk issued before first use.
#include "common.cuh"

namespace hg {

template <typename T, int N>
__global__ void __launch_bounds__(256) combine_vec_kernel(PartialTable t, T* __restrict__ out, float* __restrict__ lse_out,
                                                          int64_t n_chunks, int chunks_per_row) {
  constexpr int VEC = Vec16<T>::VEC;
  const int n = (N > 0) ? N : t.n;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n_chunks;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = idx / chunks_per_row;
    uint4 raw[HG_MAX_COMBINE];
    float l[HG_MAX_COMBINE];
#pragma unroll
    for (int i = 0; i < HG_MAX_COMBINE; ++i) {
      if (i < n) {
        raw[i] = ld_stream_v4(reinterpret_cast<const T*>(t.outs[i]) + idx * VEC);
        l[i] = __ldg(t.lses[i] + row);
      }
    }
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < HG_MAX_COMBINE; ++i)
      if (i < n) m = fmaxf(m, l[i]);
    const float m_safe = (m == -INFINITY) ? 0.f : m;
    float acc[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
    float den = 0.f;
#pragma unroll
    for (int i = 0; i < HG_MAX_COMBINE; ++i) {
      if (i < n) {
        const float w = __expf(l[i] - m_safe);
        den += w;
        float f[VEC];
        Vec16<T>::unpack(raw[i], f);
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] = fmaf(w, f[e], acc[e]);
      }
    }
    const float inv = den > 0.f ? 1.f / den : 0.f;
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e] *= inv;
    st_v4(out + idx * VEC, Vec16<T>::pack(acc));
    if (lse_out != nullptr && idx % chunks_per_row == 0) lse_out[row] = den > 0.f ? m_safe + __logf(den) : -INFINITY;
  }
}

// Any d / any alignment: one thread per element.
template <typename T>
__global__ void __launch_bounds__(256) combine_scalar_kernel(PartialTable t, T* __restrict__ out, float* __restrict__ lse_out,
                                                             int64_t n_elems, int d) {
  const int n = t.n;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n_elems;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = idx / d;
    float m = -INFINITY;
    for (int i = 0; i < n; ++i) m = fmaxf(m, __ldg(t.lses[i] + row));
    const float m_safe = (m == -INFINITY) ? 0.f : m;
    float acc = 0.f, den = 0.f;
    for (int i = 0; i < n; ++i) {
      const float w = __expf(__ldg(t.lses[i] + row) - m_safe);
      den += w;
      acc = fmaf(w, to_f32<T>(reinterpret_cast<const T*>(t.outs[i])[idx]), acc);
    }
    out[idx] = from_f32<T>(den > 0.f ? acc / den : 0.f);
    if (lse_out != nullptr && idx % d == 0) lse_out[row] = den > 0.f ? m_safe + __logf(den) : -INFINITY;
  }
}

template <typename T>
static int launch_combine_t(const PartialTable& t, void* out, float* lse_out, int64_t rows, int d, cudaStream_t s) {
  constexpr int VEC = Vec16<T>::VEC;
  bool aligned = (d % VEC == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0);
  for (int i = 0; i < t.n; ++i) aligned = aligned && (reinterpret_cast<uintptr_t>(t.outs[i]) % 16 == 0);
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  if (aligned) {
    const int cpr = d / VEC;
    const int64_t n_chunks = rows * cpr;
    int64_t blocks = (n_chunks + 255) / 256;
    const int64_t cap = (int64_t)sms * 16;  // 16 resident 256-thread CTAs cover the SM's 64 warps twice over
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (t.n == 2)
      combine_vec_kernel<T, 2><<<(unsigned)blocks, 256, 0, s>>>(t, (T*)out, lse_out, n_chunks, cpr);
    else if (t.n == 3)
      combine_vec_kernel<T, 3><<<(unsigned)blocks, 256, 0, s>>>(t, (T*)out, lse_out, n_chunks, cpr);
    else
      combine_vec_kernel<T, 0><<<(unsigned)blocks, 256, 0, s>>>(t, (T*)out, lse_out, n_chunks, cpr);
  } else {
    const int64_t n_elems = rows * d;
    int64_t blocks = (n_elems + 255) / 256;
    const int64_t cap = (int64_t)sms * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    combine_scalar_kernel<T><<<(unsigned)blocks, 256, 0, s>>>(t, (T*)out, lse_out, n_elems, d);
  }
  return check_launch("combine_lse");
}

int launch_combine(const PartialTable& t, void* out, float* lse_out, int64_t rows, int d, int dtype, cudaStream_t s) {
  if (rows == 0) return HG_OK;
  switch (dtype) {
    case HG_F16: return launch_combine_t<__half>(t, out, lse_out, rows, d, s);
    case HG_BF16: return launch_combine_t<__nv_bfloat16>(t, out, lse_out, rows, d, s);
    case HG_F32: return launch_combine_t<float>(t, out, lse_out, rows, d, s);
    default: return set_error(HG_ERR_INVALID_ARGUMENT, "combine: unknown dtype %d", dtype);
  }
}

}  // namespace hg
