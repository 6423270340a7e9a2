// Rotary position embedding of the new q and k rows, in one launch.
//
// Replaces apply_rotary_pos_emb as the reference calls it right before the attention hot path
// (hydragen/llama.py:494-501 -> transformers 4.37.2, pinned upstream, not in tree): a table gather
// `cos[position_ids].unsqueeze(2)` and, for q and for k, `x * cos + rotate_half(x) * sin` with
// rotate_half(x) = cat(-x[d/2:], x[:d/2]) -- ten elementwise launches and as many temporaries per layer.
//
// Numerics follow that eager evaluation operation by operation so that the result is BIT-IDENTICAL to it:
// the tables are in the activation dtype (HydragenLlamaRotaryEmbedding.forward casts them,
// hydragen/llama.py:47-55), every product and the final sum are computed in fp32 from the rounded operands
// and rounded to the activation dtype (what torch's 16-bit elementwise kernels do); no FMA contraction.
//
// HBM-bound: reads and writes every q / k element once (cfg#2: 2 x 16 MiB); the table rows (one per
// sequence, shared by all its heads) stay in L1/L2.  One thread owns a 16-byte chunk of the first half of a
// head row and the matching chunk of the second half, so the update can be done in place.
#include "common.cuh"

namespace hg {

namespace {

// Round a pair of fp32 values to T and back (the rounding torch applies after every elementwise operation).
// bf16: ONE packed conversion (F2FP on the ALU/FMA side) + two bit operations -- the scalar cvt.rn.bf16.f32 is an XU
// (MUFU-pipe) instruction and made the first version of this kernel conversion-bound (r01k: XU 53 % busy, 9.0 us).
template <typename T>
__device__ __forceinline__ void rnd2(float& a, float& b);
template <>
__device__ __forceinline__ void rnd2<float>(float&, float&) {}
template <>
__device__ __forceinline__ void rnd2<__nv_bfloat16>(float& a, float& b) {
  uint32_t u;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(b), "f"(a));  // first source -> upper half
  a = __uint_as_float(u << 16);
  b = __uint_as_float(u & 0xffff0000u);
}
template <>
__device__ __forceinline__ void rnd2<__half>(float& a, float& b) {
  const float2 r = __half22float2(__floats2half2_rn(a, b));
  a = r.x;
  b = r.y;
}

template <typename T, bool I64>
__global__ void __launch_bounds__(256)
    rope_qk_kernel(const T* q, const T* k, T* q_out, T* k_out,  // q_out / k_out may alias q / k
                   const T* __restrict__ cos_t,
                   const T* __restrict__ sin_t, const void* __restrict__ positions, int64_t rows, int hq, int hkv, int d,
                   int64_t q_stride, int64_t k_stride, int64_t qo_stride, int64_t ko_stride, int64_t table_rows) {
  constexpr int VEC = Vec16<T>::VEC;
  const int half = d >> 1;
  const int chunks = half / VEC;      // 16-byte chunks per half row
  const int units = (hq + hkv) * chunks;  // threads' worth of work per token row
  // a CTA pass covers rpb whole rows (small models) or one row in several strides (units >= blockDim)
  const int rpb = units >= (int)blockDim.x ? 1 : (int)blockDim.x / units;
  const int lr = rpb == 1 ? 0 : (int)threadIdx.x / units;
  const int u0 = (int)threadIdx.x - lr * units;
  const int ustep = rpb == 1 ? (int)blockDim.x : units;
  if (lr >= rpb) return;
  for (int64_t r = (int64_t)blockIdx.x * rpb + lr; r < rows; r += (int64_t)gridDim.x * rpb)
  for (int u = u0; u < units; u += ustep) {
    const int h = u / chunks, c = u - h * chunks;
    int64_t pos = I64 ? reinterpret_cast<const int64_t*>(positions)[r] : (int64_t) reinterpret_cast<const int32_t*>(positions)[r];
    pos = pos < 0 ? 0 : (pos >= table_rows ? table_rows - 1 : pos);  // memory safety only: callers keep pos in range
    const T* src;
    T* dst;
    if (h < hq) {
      src = q + r * q_stride + (int64_t)h * d;
      dst = q_out + r * qo_stride + (int64_t)h * d;
    } else {
      src = k + r * k_stride + (int64_t)(h - hq) * d;
      dst = k_out + r * ko_stride + (int64_t)(h - hq) * d;
    }
    const int e = c * VEC;
    float xl[VEC], xh[VEC], cl[VEC], ch[VEC], sl[VEC], sh[VEC], ol[VEC], oh[VEC];
    Vec16<T>::unpack(ld_v4(src + e), xl);
    Vec16<T>::unpack(ld_v4(src + half + e), xh);
    Vec16<T>::unpack(ld_v4(cos_t + pos * d + e), cl);
    Vec16<T>::unpack(ld_v4(cos_t + pos * d + half + e), ch);
    Vec16<T>::unpack(ld_v4(sin_t + pos * d + e), sl);
    Vec16<T>::unpack(ld_v4(sin_t + pos * d + half + e), sh);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      // first half:  x_lo * cos + (-x_hi) * sin;   second half:  x_hi * cos + x_lo * sin
      float a0 = __fmul_rn(xl[i], cl[i]), a1 = __fmul_rn(-xh[i], sl[i]);
      float b0 = __fmul_rn(xh[i], ch[i]), b1 = __fmul_rn(xl[i], sh[i]);
      rnd2<T>(a0, a1);
      rnd2<T>(b0, b1);
      ol[i] = __fadd_rn(a0, a1);
      oh[i] = __fadd_rn(b0, b1);
    }
    st_v4(dst + e, Vec16<T>::pack(ol));
    st_v4(dst + half + e, Vec16<T>::pack(oh));
  }
}

template <typename T>
int launch_rope_t(const RopeParams& p, cudaStream_t s) {
  constexpr int VEC = Vec16<T>::VEC;
  const int units = (p.hq + p.hkv) * (p.d / 2 / VEC);
  if (p.rows == 0 || units == 0) return HG_OK;
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  const int rpb = units >= 256 ? 1 : 256 / units;
  const int64_t want = (p.rows + rpb - 1) / rpb;
  const int64_t cap = (int64_t)sms * 32;  // grid-stride beyond 32 CTAs per SM
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  if (p.positions_i64)
    rope_qk_kernel<T, true><<<grid, 256, 0, s>>>((const T*)p.q, (const T*)p.k, (T*)p.q_out, (T*)p.k_out, (const T*)p.cos, (const T*)p.sin,
                                                  p.positions, p.rows, p.hq, p.hkv, p.d, p.q_stride_row, p.k_stride_row,
                                                  p.q_out_stride_row, p.k_out_stride_row, p.table_rows);
  else
    rope_qk_kernel<T, false><<<grid, 256, 0, s>>>((const T*)p.q, (const T*)p.k, (T*)p.q_out, (T*)p.k_out, (const T*)p.cos, (const T*)p.sin,
                                                   p.positions, p.rows, p.hq, p.hkv, p.d, p.q_stride_row, p.k_stride_row,
                                                   p.q_out_stride_row, p.k_out_stride_row, p.table_rows);
  return check_launch("rope_qk");
}

}  // namespace

int launch_rope(const RopeParams& p, int dtype, cudaStream_t s) {
  switch (dtype) {
    case HG_F16: return launch_rope_t<__half>(p, s);
    case HG_BF16: return launch_rope_t<__nv_bfloat16>(p, s);
    case HG_F32: return launch_rope_t<float>(p, s);
  }
  return set_error(HG_ERR_INVALID_ARGUMENT, "rope: unknown dtype %d", dtype);
}

}  // namespace hg
