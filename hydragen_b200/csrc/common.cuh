// Shared device/host helpers for the hydragen_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hydragen_b200.h"

namespace hg {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// ---- error plumbing (api.cu) ------------------------------------------------------------
int set_error(int code, const char* fmt, ...);
int check_launch(const char* what);
struct DeviceInfo {
  bool ready = false;
  int device = -1;
  int sm_count = 0;
  int max_smem_optin = 0;
  void* encode_tiled = nullptr;  // cuTensorMapEncodeTiled, fetched through the runtime
};
const DeviceInfo& device_info();
bool pdl_enabled();  // programmatic dependent launch between the kernels of a decode step

// ---- 16-byte vector of T <-> fp32 --------------------------------------------------------
template <typename T>
struct Vec16;  // VEC elements of T in one 128-bit word

template <>
struct Vec16<float> {
  static constexpr int VEC = 4;
  static __device__ __forceinline__ void unpack(const uint4& u, float* f) {
    f[0] = __uint_as_float(u.x);
    f[1] = __uint_as_float(u.y);
    f[2] = __uint_as_float(u.z);
    f[3] = __uint_as_float(u.w);
  }
  static __device__ __forceinline__ uint4 pack(const float* f) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
  }
};

template <>
struct Vec16<__nv_bfloat16> {
  static constexpr int VEC = 8;
  // bf16 -> fp32 is a 16-bit shift: two ALU ops per packed pair, no conversion unit.
  static __device__ __forceinline__ void unpack(const uint4& u, float* f) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  static __device__ __forceinline__ uint4 pack(const float* f) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
};

template <>
struct Vec16<__half> {
  static constexpr int VEC = 8;
  static __device__ __forceinline__ void unpack(const uint4& u, float* f) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
  static __device__ __forceinline__ uint4 pack(const float* f) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
};

template <typename T>
__device__ __forceinline__ float to_f32(T x);
template <>
__device__ __forceinline__ float to_f32<float>(float x) { return x; }
template <>
__device__ __forceinline__ float to_f32<__half>(__half x) { return __half2float(x); }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }

template <typename T>
__device__ __forceinline__ T from_f32(float x);
template <>
__device__ __forceinline__ float from_f32<float>(float x) { return x; }
template <>
__device__ __forceinline__ __half from_f32<__half>(float x) { return __float2half_rn(x); }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

// Streaming 128-bit load: read-only path, do not allocate in L1 (each K/V row is used once).
__device__ __forceinline__ uint4 ld_stream_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ld_v4(const void* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void st_v4(void* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_log2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Pointer tables passed by value (copied from the caller's host arrays at launch).
struct PartialTable {
  const void* outs[HG_MAX_COMBINE];
  const float* lses[HG_MAX_COMBINE];
  int n;
};

// ---- launchers implemented in the .cu files ---------------------------------------------
int launch_combine(const PartialTable& t, void* out, float* lse_out, int64_t rows, int d, int dtype, cudaStream_t s);

struct RowwiseParams {
  const void* q;
  const void* k;
  const void* v;
  const void* seq_lens;
  int seq_lens_i64;
  const int32_t* cu_seqlens_k;
  int kv_group_size;
  int causal;
  void* out;
  float* lse;
  int b, nq, lk, hq, hkv, d;
  int64_t q_stride_b, q_stride_s, q_stride_h;
  int64_t kv_stride_b, kv_stride_s, kv_stride_h;
  PartialTable partials;
  float scale_log2;  // sm_scale * log2(e)
  // fused decode step (hg_decode_attn_fused): new token rows [b, hkv, d] appended at positions[b]
  const void* k_new;
  const void* v_new;
  const void* positions;
  int positions_i64;
};
int launch_rowwise(const RowwiseParams& p, int dtype, cudaStream_t s);

// One shared level of a hierarchy: its K/V, its (partial) output, its grouping of the query rows.
struct PrefixLevel {
  const void* k;
  const void* v;
  void* out;
  float* lse;
  const int32_t* cu_seqlens_k;
  int64_t n_k_rows, kv_stride_row;
  int n_groups, k_len, max_k_len;
};
struct PrefixParams {
  const void* q;
  int64_t n_q_rows, q_stride_row;
  PrefixLevel levels[4];
  int n_levels;
  int hq, hkv, d;
  float scale_log2;
  int causal;  // one level, bottom-right aligned causal mask inside every group (prefill): row r sees keys <= r + (k_len - q_per_group)
  void* workspace;  // stream-K partials and flags (hg_prefix_workspace_bytes); nullptr: whole units only
  int64_t workspace_bytes;
  int kv_splits;    // one level only: >= 1; the level's out / lse then hold kv_splits partial results back to back
};
struct SchedParams;
int launch_prefix(const PrefixParams& p, int dtype, cudaStream_t s);
int build_prefix_schedule(const PrefixParams& p, int n_sms, int split_mode, SchedParams* out);  // split_mode: 0 whole units, 1 stream-K where it pays, 2 always stream-K
int64_t prefix_workspace_bytes();
int launch_prefix_unit(const PrefixParams& p, int kv_splits, int dtype, cudaStream_t s);  // prefix_unit_sm100.cu: one level, one CTA per unit
int suggest_prefix_splits(int n_groups, int q_per_group, int hq, int max_k_len, int max_splits);

int launch_allreduce_multimem(void* mc_ptr, void* out, const void* flags_dev, int rank, int world, int64_t nbytes, int dtype,
                              int n_blocks, cudaStream_t s);

// o_proj GEMM fused with its all-reduce (oproj_allreduce.cu)
struct OprojParams {
  const void* x;  // [m, k], row stride x_stride_row
  const void* w;  // [n, k], row stride w_stride_row
  void* out;      // [m, n] contiguous: this rank's symmetric buffer (world > 1) or any buffer (world == 1)
  void* out_mc;   // multicast address of the same buffer (world > 1)
  const void* flags_dev;
  int64_t m, n, k, x_stride_row, w_stride_row;
  int rank, world, dtype, n_ctas;
};
int launch_oproj_allreduce(const OprojParams& p, cudaStream_t s);
int oproj_allreduce_flag_words(int64_t m, int64_t n, int world);
int oproj_allreduce_plan(int64_t m, int64_t n, int world, int rank, int n_ctas, int* geometry_out, int32_t* cover_out);

int launch_kv_append(const void* k_new, const void* v_new, const void* positions, int positions_i64, void* k_cache,
                     void* v_cache, int b, int nq, int lk, int hkv, int d, int dtype, cudaStream_t s);


struct RopeParams {
  const void* q;
  const void* k;
  void* q_out;
  void* k_out;
  const void* cos;
  const void* sin;
  const void* positions;
  int positions_i64;
  int64_t rows;
  int hq, hkv, d;
  int64_t q_stride_row, k_stride_row, q_out_stride_row, k_out_stride_row;
  int64_t table_rows;
};
int launch_rope(const RopeParams& p, int dtype, cudaStream_t s);

}  // namespace hg
