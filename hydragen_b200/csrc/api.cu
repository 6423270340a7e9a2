// C ABI of hydragen_b200 (see include/hydragen_b200.h): argument validation, error plumbing and
// device bring-up.  No torch types, no allocation, no synchronisation on any launch path.
#include <cstdarg>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "common.cuh"
#include "prefix_sched.h"

namespace hg {

static thread_local char g_err[512] = "";
static DeviceInfo g_info;
static std::mutex g_init_mutex;

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();  // clear the (non-sticky) launch error
    return set_error(HG_ERR_CUDA, "%s: launch failed: %s", what, cudaGetErrorString(e));
  }
  return HG_OK;
}

const DeviceInfo& device_info() { return g_info; }

// HYDRAGEN_B200_PDL=0 turns programmatic dependent launch off (debugging aid; read once)
bool pdl_enabled() {
  static const bool on = [] {
    const char* v = getenv("HYDRAGEN_B200_PDL");
    return !(v != nullptr && v[0] == '0');
  }();
  return on;
}

static bool valid_dtype(int dtype) { return dtype == HG_F16 || dtype == HG_BF16 || dtype == HG_F32; }

static int fill_partials(PartialTable& t, const void* const* outs, const float* const* lses, int n, const char* who) {
  if (n < 0 || n > HG_MAX_COMBINE) return set_error(HG_ERR_INVALID_ARGUMENT, "%s: n = %d outside [0, %d]", who, n, HG_MAX_COMBINE);
  memset(&t, 0, sizeof(t));
  t.n = n;
  if (n > 0 && (outs == nullptr || lses == nullptr)) return set_error(HG_ERR_INVALID_ARGUMENT, "%s: null pointer table", who);
  for (int i = 0; i < n; ++i) {
    if (outs[i] == nullptr || lses[i] == nullptr) return set_error(HG_ERR_INVALID_ARGUMENT, "%s: null partial %d", who, i);
    t.outs[i] = outs[i];
    t.lses[i] = lses[i];
  }
  return HG_OK;
}

}  // namespace hg

using namespace hg;

extern "C" {

int hg_abi_version(void) { return HG_ABI_VERSION; }

const char* hg_last_error(void) { return g_err; }

int hg_sm_count(void) { return g_info.sm_count; }

int hg_init(int device) {
  std::lock_guard<std::mutex> lock(g_init_mutex);
  if (g_info.ready && g_info.device == device) return HG_OK;
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return set_error(HG_ERR_CUDA, "hg_init: cudaGetDeviceProperties(%d): %s", device, cudaGetErrorString(e));
  if (prop.major != 10)
    return set_error(HG_ERR_UNSUPPORTED, "hg_init: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                     prop.major, prop.minor);
  g_info.device = device;
  g_info.sm_count = prop.multiProcessorCount;
  g_info.max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr)
    return set_error(HG_ERR_CUDA, "hg_init: cuTensorMapEncodeTiled not found: %s", cudaGetErrorString(e));
  g_info.encode_tiled = fn;
  g_info.ready = true;
  return HG_OK;
}

int hg_combine_lse(const void* const* outs_host, const float* const* lses_host, int n, void* out, float* lse_out, int64_t rows,
                   int d, int dtype, void* stream) {
  if (n < 1) return set_error(HG_ERR_INVALID_ARGUMENT, "combine: n = %d, need at least one input", n);
  if (!valid_dtype(dtype)) return set_error(HG_ERR_INVALID_ARGUMENT, "combine: unknown dtype %d", dtype);
  if (rows < 0 || d < 1) return set_error(HG_ERR_INVALID_ARGUMENT, "combine: rows = %lld, d = %d", (long long)rows, d);
  if (out == nullptr && rows > 0) return set_error(HG_ERR_INVALID_ARGUMENT, "combine: out is null");
  PartialTable t;
  int rc = fill_partials(t, outs_host, lses_host, n, "combine");
  if (rc != HG_OK) return rc;
  return launch_combine(t, out, lse_out, rows, d, dtype, (cudaStream_t)stream);
}

int hg_rowwise_attn_fwd(const void* q, const void* k, const void* v, const void* seq_lens, int seq_lens_i64,
                        const int32_t* cu_seqlens_k, int kv_group_size, int causal, void* out, float* lse, int b, int nq, int lk,
                        int hq, int hkv, int d, int64_t q_stride_b, int64_t q_stride_s, int64_t q_stride_h, int64_t kv_stride_b,
                        int64_t kv_stride_s, int64_t kv_stride_h, const void* const* partial_outs_host,
                        const float* const* partial_lses_host, int n_partials, float sm_scale, int dtype, void* stream) {
  if (!valid_dtype(dtype)) return set_error(HG_ERR_INVALID_ARGUMENT, "rowwise: unknown dtype %d", dtype);
  if (b < 0 || nq < 0 || lk < 0 || hq < 1 || hkv < 1 || d < 1)
    return set_error(HG_ERR_INVALID_ARGUMENT, "rowwise: bad sizes b=%d nq=%d lk=%d hq=%d hkv=%d d=%d", b, nq, lk, hq, hkv, d);
  if (hq % hkv != 0) return set_error(HG_ERR_INVALID_ARGUMENT, "rowwise: hq (%d) must be a multiple of hkv (%d)", hq, hkv);
  if (kv_group_size < 1 || (b % kv_group_size) != 0)
    return set_error(HG_ERR_INVALID_ARGUMENT, "rowwise: kv_group_size %d must divide b = %d", kv_group_size, b);
  if (seq_lens != nullptr && kv_group_size != 1)
    return set_error(HG_ERR_INVALID_ARGUMENT, "rowwise: seq_lens requires kv_group_size == 1");
  if (b > 0 && nq > 0 && (q == nullptr || out == nullptr)) return set_error(HG_ERR_INVALID_ARGUMENT, "rowwise: null q/out");
  if (b > 0 && lk > 0 && (k == nullptr || v == nullptr)) return set_error(HG_ERR_INVALID_ARGUMENT, "rowwise: null k/v");
  const int esz = dtype == HG_F32 ? 4 : 2;
  const int vec = 16 / esz;
  if (q_stride_b % vec || q_stride_s % vec || q_stride_h % vec || kv_stride_b % vec || kv_stride_s % vec || kv_stride_h % vec ||
      reinterpret_cast<uintptr_t>(q) % 16 || reinterpret_cast<uintptr_t>(k) % 16 || reinterpret_cast<uintptr_t>(v) % 16 ||
      reinterpret_cast<uintptr_t>(out) % 16)
    return set_error(HG_ERR_UNSUPPORTED, "rowwise: bases and strides must be 16-byte aligned");
  RowwiseParams p;
  memset(&p, 0, sizeof(p));
  int rc = fill_partials(p.partials, partial_outs_host, partial_lses_host, n_partials, "rowwise");
  if (rc != HG_OK) return rc;
  for (int i = 0; i < n_partials; ++i)
    if (reinterpret_cast<uintptr_t>(partial_outs_host[i]) % 16)
      return set_error(HG_ERR_UNSUPPORTED, "rowwise: partial outs must be 16-byte aligned");
  p.q = q; p.k = k; p.v = v;
  p.seq_lens = seq_lens; p.seq_lens_i64 = seq_lens_i64;
  p.cu_seqlens_k = cu_seqlens_k;
  p.kv_group_size = kv_group_size;
  p.causal = causal;
  p.out = out; p.lse = lse;
  p.b = b; p.nq = nq; p.lk = lk; p.hq = hq; p.hkv = hkv; p.d = d;
  p.q_stride_b = q_stride_b; p.q_stride_s = q_stride_s; p.q_stride_h = q_stride_h;
  p.kv_stride_b = kv_stride_b; p.kv_stride_s = kv_stride_s; p.kv_stride_h = kv_stride_h;
  p.scale_log2 = sm_scale * kLog2e;
  return launch_rowwise(p, dtype, (cudaStream_t)stream);
}

int hg_decode_attn_fused(const void* q, const void* k_new, const void* v_new, const void* positions, int positions_i64,
                         void* k_cache, void* v_cache, void* out, float* lse, int b, int lk, int hq, int hkv, int d,
                         int64_t q_stride_b, int64_t q_stride_h, int64_t kv_stride_b, int64_t kv_stride_s, int64_t kv_stride_h,
                         const void* const* partial_outs_host, const float* const* partial_lses_host, int n_partials,
                         float sm_scale, int dtype, void* stream) {
  if (!valid_dtype(dtype)) return set_error(HG_ERR_INVALID_ARGUMENT, "decode_attn_fused: unknown dtype %d", dtype);
  if (b < 0 || lk < 1 || hq < 1 || hkv < 1 || d < 1)
    return set_error(HG_ERR_INVALID_ARGUMENT, "decode_attn_fused: bad sizes b=%d lk=%d hq=%d hkv=%d d=%d", b, lk, hq, hkv, d);
  if (hq % hkv != 0) return set_error(HG_ERR_INVALID_ARGUMENT, "decode_attn_fused: hq (%d) must be a multiple of hkv (%d)", hq, hkv);
  if (b > 0 && (!q || !k_new || !v_new || !positions || !k_cache || !v_cache || !out))
    return set_error(HG_ERR_INVALID_ARGUMENT, "decode_attn_fused: null pointer");
  const int esz = dtype == HG_F32 ? 4 : 2;
  const int vec = 16 / esz;
  if (q_stride_b % vec || q_stride_h % vec || kv_stride_b % vec || kv_stride_s % vec || kv_stride_h % vec ||
      reinterpret_cast<uintptr_t>(q) % 16 || reinterpret_cast<uintptr_t>(k_new) % 16 || reinterpret_cast<uintptr_t>(v_new) % 16 ||
      reinterpret_cast<uintptr_t>(k_cache) % 16 || reinterpret_cast<uintptr_t>(v_cache) % 16 || reinterpret_cast<uintptr_t>(out) % 16)
    return set_error(HG_ERR_UNSUPPORTED, "decode_attn_fused: bases and strides must be 16-byte aligned");
  RowwiseParams p;
  memset(&p, 0, sizeof(p));
  int rc = fill_partials(p.partials, partial_outs_host, partial_lses_host, n_partials, "decode_attn_fused");
  if (rc != HG_OK) return rc;
  for (int i = 0; i < n_partials; ++i)
    if (reinterpret_cast<uintptr_t>(partial_outs_host[i]) % 16)
      return set_error(HG_ERR_UNSUPPORTED, "decode_attn_fused: partial outs must be 16-byte aligned");
  p.q = q; p.k = k_cache; p.v = v_cache;
  p.k_new = k_new; p.v_new = v_new; p.positions = positions; p.positions_i64 = positions_i64;
  p.kv_group_size = 1;
  p.out = out; p.lse = lse;
  p.b = b; p.nq = 1; p.lk = lk; p.hq = hq; p.hkv = hkv; p.d = d;
  p.q_stride_b = q_stride_b; p.q_stride_s = 0; p.q_stride_h = q_stride_h;
  p.kv_stride_b = kv_stride_b; p.kv_stride_s = kv_stride_s; p.kv_stride_h = kv_stride_h;
  p.scale_log2 = sm_scale * kLog2e;
  return launch_rowwise(p, dtype, (cudaStream_t)stream);
}

static int fill_prefix_levels(PrefixParams& p, const hg_prefix_level* levels_host, int n_levels, const char* who) {
  if (n_levels < 1 || n_levels > kMaxLevels) return set_error(HG_ERR_INVALID_ARGUMENT, "%s: %d shared levels (1..%d)", who, n_levels, kMaxLevels);
  if (levels_host == nullptr) return set_error(HG_ERR_INVALID_ARGUMENT, "%s: null level table", who);
  p.n_levels = n_levels;
  for (int l = 0; l < n_levels; ++l) {
    const hg_prefix_level& h = levels_host[l];
    if (h.n_groups < 1 || h.n_k_rows < 0 || h.k_len < 0 || h.max_k_len < 0)
      return set_error(HG_ERR_INVALID_ARGUMENT, "%s: level %d: bad sizes n_groups=%d n_k_rows=%lld k_len=%d max_k_len=%d", who, l, h.n_groups,
                       (long long)h.n_k_rows, h.k_len, h.max_k_len);
    if (p.n_q_rows % h.n_groups != 0)
      return set_error(HG_ERR_INVALID_ARGUMENT, "%s: level %d: %d groups do not divide %lld query rows", who, l, h.n_groups, (long long)p.n_q_rows);
    if (h.cu_seqlens_k == nullptr && (int64_t)h.n_groups * h.k_len > h.n_k_rows)
      return set_error(HG_ERR_INVALID_ARGUMENT, "%s: level %d: n_groups * k_len (%d * %d) exceeds n_k_rows (%lld)", who, l, h.n_groups, h.k_len,
                       (long long)h.n_k_rows);
    if (p.n_q_rows > 0 && (h.k == nullptr || h.v == nullptr || h.out == nullptr))
      return set_error(HG_ERR_INVALID_ARGUMENT, "%s: level %d: null tensor pointer", who, l);
    if (h.n_k_rows > 0x7fffffffLL) return set_error(HG_ERR_UNSUPPORTED, "%s: row counts must fit int32", who);
    PrefixLevel& d = p.levels[l];
    d.k = h.k; d.v = h.v; d.out = h.out; d.lse = h.lse;
    d.cu_seqlens_k = h.cu_seqlens_k;
    d.n_k_rows = h.n_k_rows; d.kv_stride_row = h.kv_stride_row;
    d.n_groups = h.n_groups; d.k_len = h.k_len;
    d.max_k_len = h.cu_seqlens_k != nullptr ? (h.max_k_len > 0 ? h.max_k_len : (int)h.n_k_rows) : h.k_len;
  }
  return HG_OK;
}

int64_t hg_prefix_workspace_bytes(void) { return prefix_workspace_bytes(); }

int hg_prefix_attn_grouped_fwd(const void* q, int64_t n_q_rows, int64_t q_stride_row, const hg_prefix_level* levels_host, int n_levels,
                               int hq, int hkv, int d, float sm_scale, int dtype, void* workspace, int64_t workspace_bytes,
                               void* stream) {
  if (!g_info.ready) return set_error(HG_ERR_NOT_INITIALIZED, "prefix: hg_init() has not been called");
  if (n_q_rows < 0 || hq < 1 || hkv < 1) return set_error(HG_ERR_INVALID_ARGUMENT, "prefix: bad sizes n_q_rows=%lld hq=%d hkv=%d", (long long)n_q_rows, hq, hkv);
  if (hq % hkv != 0) return set_error(HG_ERR_INVALID_ARGUMENT, "prefix: hq (%d) must be a multiple of hkv (%d)", hq, hkv);
  if (n_q_rows > 0 && q == nullptr) return set_error(HG_ERR_INVALID_ARGUMENT, "prefix: null tensor pointer");
  if (n_q_rows > 0x7fffffffLL) return set_error(HG_ERR_UNSUPPORTED, "prefix: row counts must fit int32");
  if (hq > 65535) return set_error(HG_ERR_UNSUPPORTED, "prefix: hq > 65535");
  if (workspace != nullptr && reinterpret_cast<uintptr_t>(workspace) % 16 != 0)
    return set_error(HG_ERR_UNSUPPORTED, "prefix: the workspace must be 16-byte aligned");
  PrefixParams p;
  memset(&p, 0, sizeof(p));
  p.q = q; p.n_q_rows = n_q_rows; p.q_stride_row = q_stride_row;
  p.hq = hq; p.hkv = hkv; p.d = d;
  p.scale_log2 = sm_scale * kLog2e;
  p.workspace = workspace; p.workspace_bytes = workspace_bytes;
  int rc = fill_prefix_levels(p, levels_host, n_levels, "prefix");
  if (rc != HG_OK) return rc;
  return launch_prefix(p, dtype, (cudaStream_t)stream);
}

int hg_prefix_attn_split_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int n_groups, int q_per_group,
                             int64_t n_k_rows, int k_len, const int32_t* cu_seqlens_k, int max_k_len, int hq, int hkv, int d,
                             int64_t q_stride_row, int64_t kv_stride_row, float sm_scale, int dtype, int kv_splits, void* stream) {
  if (kv_splits < 1 || kv_splits > HG_MAX_COMBINE)
    return set_error(HG_ERR_INVALID_ARGUMENT, "prefix: kv_splits = %d outside [1, %d]", kv_splits, HG_MAX_COMBINE);
  if (!g_info.ready) return set_error(HG_ERR_NOT_INITIALIZED, "prefix: hg_init() has not been called");
  if (n_groups < 0 || q_per_group < 0 || n_k_rows < 0 || hq < 1 || hkv < 1)
    return set_error(HG_ERR_INVALID_ARGUMENT, "prefix: bad sizes n_groups=%d q_per_group=%d n_k_rows=%lld hq=%d hkv=%d", n_groups,
                     q_per_group, (long long)n_k_rows, hq, hkv);
  if (hq % hkv != 0) return set_error(HG_ERR_INVALID_ARGUMENT, "prefix: hq (%d) must be a multiple of hkv (%d)", hq, hkv);
  if (n_groups == 0 || q_per_group == 0) return HG_OK;
  if ((int64_t)n_groups * q_per_group * kv_splits > 0x7fffffffLL) return set_error(HG_ERR_UNSUPPORTED, "prefix: row counts must fit int32");
  hg_prefix_level lv;
  memset(&lv, 0, sizeof(lv));
  lv.k = k; lv.v = v; lv.out = out; lv.lse = lse; lv.cu_seqlens_k = cu_seqlens_k;
  lv.n_k_rows = n_k_rows; lv.kv_stride_row = kv_stride_row;
  lv.n_groups = n_groups; lv.k_len = k_len; lv.max_k_len = max_k_len;
  PrefixParams p;
  memset(&p, 0, sizeof(p));
  p.q = q; p.n_q_rows = (int64_t)n_groups * q_per_group; p.q_stride_row = q_stride_row;
  p.hq = hq; p.hkv = hkv; p.d = d;
  p.scale_log2 = sm_scale * kLog2e;
  p.kv_splits = kv_splits;
  if (q == nullptr) return set_error(HG_ERR_INVALID_ARGUMENT, "prefix: null tensor pointer");
  int rc = fill_prefix_levels(p, &lv, 1, "prefix");
  if (rc != HG_OK) return rc;
  return launch_prefix(p, dtype, (cudaStream_t)stream);
}

int hg_prefix_attn_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int n_groups, int q_per_group,
                       int64_t n_k_rows, int k_len, const int32_t* cu_seqlens_k, int max_k_len, int hq, int hkv, int d,
                       int64_t q_stride_row, int64_t kv_stride_row, float sm_scale, int dtype, void* stream) {
  return hg_prefix_attn_split_fwd(q, k, v, out, lse, n_groups, q_per_group, n_k_rows, k_len, cu_seqlens_k, max_k_len, hq, hkv, d,
                                  q_stride_row, kv_stride_row, sm_scale, dtype, 1, stream);
}

int hg_prefix_suggest_splits(int n_groups, int q_per_group, int hq, int max_k_len, int max_splits) {
  if (n_groups < 1 || q_per_group < 1 || hq < 1 || max_k_len < 1 || max_splits < 1) return 1;
  return suggest_prefix_splits(n_groups, q_per_group, hq, max_k_len, max_splits > HG_MAX_COMBINE ? HG_MAX_COMBINE : max_splits);
}

int hg_prefix_schedule(const hg_prefix_level* levels_host, int n_levels, int64_t n_q_rows, int hq, int n_sms, int allow_split,
                       int32_t* pieces_out, int max_pieces, int32_t* n_ctas_out) {
  PrefixParams p;
  memset(&p, 0, sizeof(p));
  p.n_q_rows = n_q_rows; p.hq = hq; p.hkv = hq; p.d = 128;
  // the schedule depends on sizes only: pointers may be null here
  if (n_levels < 1 || n_levels > kMaxLevels || levels_host == nullptr) return set_error(HG_ERR_INVALID_ARGUMENT, "prefix_schedule: %d levels", n_levels);
  if (hq < 1 || n_q_rows < 1 || n_sms < 1) return set_error(HG_ERR_INVALID_ARGUMENT, "prefix_schedule: bad sizes");
  p.n_levels = n_levels;
  for (int l = 0; l < n_levels; ++l) {
    const hg_prefix_level& h = levels_host[l];
    PrefixLevel& d = p.levels[l];
    d.n_groups = h.n_groups; d.k_len = h.k_len;
    d.max_k_len = h.max_k_len > 0 ? h.max_k_len : h.k_len;
    d.cu_seqlens_k = h.max_k_len > 0 ? reinterpret_cast<const int32_t*>(&d) : nullptr;  // non-null = "ragged level: cost from max_k_len"
  }
  SchedParams S;
  int rc = build_prefix_schedule(p, n_sms, allow_split < 0 ? 0 : (allow_split > 2 ? 2 : allow_split), &S);
  if (rc != HG_OK) return rc;
  if (n_ctas_out != nullptr) *n_ctas_out = S.n_ctas;
  int n = 0;
  for (int c = 0; c < S.n_ctas; ++c) {
    SchedIter it;
    SchedPiece sp;
    sched_begin(S, c, it);
    while (sched_next(S, it, sp)) {
      if (pieces_out != nullptr && n < max_pieces) {
        int32_t* o = pieces_out + (int64_t)n * 10;
        o[0] = c; o[1] = sp.unit; o[2] = sp.level; o[3] = sp.head; o[4] = sp.grp; o[5] = sp.mt;
        o[6] = sp.b_lo; o[7] = sp.b_hi; o[8] = sp.split; o[9] = sp.slot;
      }
      if (sp.split) {  // cross-check the merge's view of the unit: this CTA must appear there with the same slot
        int ctas[kMaxUnitPieces], slots[kMaxUnitPieces];
        const int k = sched_unit_pieces(S, sp.unit, c, ctas, slots);
        bool found = false;
        for (int i = 0; i < k && i < kMaxUnitPieces; ++i) found = found || (ctas[i] == c && slots[i] == sp.slot);
        if (!found || k > kMaxUnitPieces)
          return set_error(HG_ERR_CUDA, "prefix_schedule: inconsistent piece table (unit %d, cta %d, %d pieces)", sp.unit, c, k);
      }
      ++n;
    }
  }
  return n;
}

int hg_causal_attn_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int b, int sq, int sk, int hq, int hkv,
                       int d, int64_t q_stride_row, int64_t kv_stride_row, float sm_scale, int dtype, void* stream) {
  if (!g_info.ready) return set_error(HG_ERR_NOT_INITIALIZED, "causal_attn: hg_init() has not been called");
  if (b < 0 || sq < 0 || sk < 0 || hq < 1 || hkv < 1)
    return set_error(HG_ERR_INVALID_ARGUMENT, "causal_attn: bad sizes b=%d sq=%d sk=%d hq=%d hkv=%d", b, sq, sk, hq, hkv);
  if (hq % hkv != 0) return set_error(HG_ERR_INVALID_ARGUMENT, "causal_attn: hq (%d) must be a multiple of hkv (%d)", hq, hkv);
  if (sk < sq) return set_error(HG_ERR_UNSUPPORTED, "causal_attn: sk (%d) < sq (%d): rows without any visible key are not supported", sk, sq);
  if (b > 0 && sq > 0 && (q == nullptr || out == nullptr || k == nullptr || v == nullptr))
    return set_error(HG_ERR_INVALID_ARGUMENT, "causal_attn: null tensor pointer");
  if ((int64_t)b * sq > 0x7fffffffLL || (int64_t)b * sk > 0x7fffffffLL)
    return set_error(HG_ERR_UNSUPPORTED, "causal_attn: row counts must fit int32");
  if (b == 0 || sq == 0) return HG_OK;
  PrefixParams p;
  memset(&p, 0, sizeof(p));
  p.q = q; p.n_q_rows = (int64_t)b * sq; p.q_stride_row = q_stride_row;
  p.n_levels = 1;
  PrefixLevel& lv = p.levels[0];
  lv.k = k; lv.v = v; lv.out = out; lv.lse = lse;
  lv.n_k_rows = (int64_t)b * sk; lv.kv_stride_row = kv_stride_row;
  lv.n_groups = b; lv.k_len = sk; lv.max_k_len = sk;
  p.hq = hq; p.hkv = hkv; p.d = d;
  p.scale_log2 = sm_scale * kLog2e;
  p.causal = 1;
  return launch_prefix(p, dtype, (cudaStream_t)stream);
}

int hg_allreduce_multimem(void* mc_ptr, void* out, const void* flags_dev, int rank, int world, int64_t nbytes, int dtype,
                          int n_blocks, void* stream) {
  if (!valid_dtype(dtype)) return set_error(HG_ERR_INVALID_ARGUMENT, "allreduce: unknown dtype %d", dtype);
  if (world < 2 || world > 32 || rank < 0 || rank >= world) return set_error(HG_ERR_INVALID_ARGUMENT, "allreduce: rank %d of %d", rank, world);
  if (mc_ptr == nullptr || flags_dev == nullptr) return set_error(HG_ERR_INVALID_ARGUMENT, "allreduce: null pointer (no multicast mapping?)");
  if (nbytes < 0 || nbytes % 16 != 0 || reinterpret_cast<uintptr_t>(mc_ptr) % 16 != 0 || reinterpret_cast<uintptr_t>(out) % 16 != 0)
    return set_error(HG_ERR_UNSUPPORTED, "allreduce: size and address must be multiples of 16 bytes");
  if (n_blocks < 1 || n_blocks > 1024) return set_error(HG_ERR_INVALID_ARGUMENT, "allreduce: n_blocks = %d", n_blocks);
  if (nbytes == 0) return HG_OK;
  return launch_allreduce_multimem(mc_ptr, out, flags_dev, rank, world, nbytes, dtype, n_blocks, (cudaStream_t)stream);
}

int hg_oproj_allreduce_flag_words(int64_t m, int64_t n, int world) {
  if (m < 0 || n < 0 || world < 1) return 0;
  return oproj_allreduce_flag_words(m, n, world);
}

int hg_oproj_allreduce_plan(int64_t m, int64_t n, int world, int rank, int n_ctas, int* geometry_out, int32_t* cover_out) {
  if (m < 1 || n < 8 || n % 8 != 0 || m > INT32_MAX || n > INT32_MAX || world < 1 || world > 32 || rank < 0 || rank >= world || n_ctas < 0)
    return set_error(HG_ERR_INVALID_ARGUMENT, "oproj_allreduce_plan: bad arguments");
  return oproj_allreduce_plan(m, n, world, rank, n_ctas, geometry_out, cover_out);
}

int hg_oproj_allreduce_fwd(const void* x, int64_t x_stride_row, const void* w, int64_t w_stride_row, void* out, void* out_mc,
                           const void* flags_dev, int64_t flag_words, int rank, int world, int64_t m, int64_t n, int64_t k, int dtype,
                           int n_ctas, void* stream) {
  if (dtype != HG_BF16 && dtype != HG_F16) return set_error(HG_ERR_UNSUPPORTED, "oproj_allreduce: 16-bit types only (dtype %d)", dtype);
  if (world < 1 || world > 32 || rank < 0 || rank >= world) return set_error(HG_ERR_INVALID_ARGUMENT, "oproj_allreduce: rank %d of %d", rank, world);
  if (m < 0 || n < 1 || k < 1 || m > INT32_MAX || n > INT32_MAX || k > INT32_MAX)
    return set_error(HG_ERR_INVALID_ARGUMENT, "oproj_allreduce: bad sizes m=%lld n=%lld k=%lld", (long long)m, (long long)n, (long long)k);
  if (n % 8 != 0 || k % 8 != 0 || x_stride_row % 8 != 0 || w_stride_row % 8 != 0 || x_stride_row < k || w_stride_row < k)
    return set_error(HG_ERR_UNSUPPORTED, "oproj_allreduce: n, k and the row strides must be multiples of 8 elements (16 bytes), strides >= k");
  if (m == 0 && world == 1) return HG_OK;
  if (x == nullptr || w == nullptr || out == nullptr) return set_error(HG_ERR_INVALID_ARGUMENT, "oproj_allreduce: null pointer");
  if (reinterpret_cast<uintptr_t>(x) % 16 || reinterpret_cast<uintptr_t>(w) % 16 || reinterpret_cast<uintptr_t>(out) % 16 ||
      reinterpret_cast<uintptr_t>(out_mc) % 16)
    return set_error(HG_ERR_UNSUPPORTED, "oproj_allreduce: pointers must be 16-byte aligned");
  if (world > 1) {
    if (out_mc == nullptr || flags_dev == nullptr) return set_error(HG_ERR_INVALID_ARGUMENT, "oproj_allreduce: null multicast / flag pointer");
    if (m == 0) return set_error(HG_ERR_INVALID_ARGUMENT, "oproj_allreduce: empty message in a collective call");
    if (flag_words < oproj_allreduce_flag_words(m, n, world))
      return set_error(HG_ERR_INVALID_ARGUMENT, "oproj_allreduce: %lld flag words, this shape needs %d", (long long)flag_words,
                       oproj_allreduce_flag_words(m, n, world));
  }
  if (n_ctas < 0) return set_error(HG_ERR_INVALID_ARGUMENT, "oproj_allreduce: n_ctas = %d", n_ctas);
  OprojParams p;
  memset(&p, 0, sizeof(p));
  p.x = x; p.w = w; p.out = out; p.out_mc = out_mc; p.flags_dev = flags_dev;
  p.m = m; p.n = n; p.k = k; p.x_stride_row = x_stride_row; p.w_stride_row = w_stride_row;
  p.rank = rank; p.world = world; p.dtype = dtype; p.n_ctas = n_ctas;
  return launch_oproj_allreduce(p, (cudaStream_t)stream);
}

int hg_kv_append(const void* k_new, const void* v_new, const void* positions, int positions_i64, void* k_cache, void* v_cache,
                 int b, int nq, int lk, int hkv, int d, int dtype, void* stream) {
  if (!valid_dtype(dtype)) return set_error(HG_ERR_INVALID_ARGUMENT, "kv_append: unknown dtype %d", dtype);
  if (b < 0 || nq < 0 || lk < 1 || hkv < 1 || d < 1) return set_error(HG_ERR_INVALID_ARGUMENT, "kv_append: bad sizes");
  if (b > 0 && nq > 0 && (!k_new || !v_new || !positions || !k_cache || !v_cache))
    return set_error(HG_ERR_INVALID_ARGUMENT, "kv_append: null pointer");
  if (reinterpret_cast<uintptr_t>(k_new) % 16 || reinterpret_cast<uintptr_t>(v_new) % 16 ||
      reinterpret_cast<uintptr_t>(k_cache) % 16 || reinterpret_cast<uintptr_t>(v_cache) % 16)
    return set_error(HG_ERR_UNSUPPORTED, "kv_append: pointers must be 16-byte aligned");
  return launch_kv_append(k_new, v_new, positions, positions_i64, k_cache, v_cache, b, nq, lk, hkv, d, dtype, (cudaStream_t)stream);
}

int hg_rope_qk(const void* q, const void* k, void* q_out, void* k_out, const void* cos_table, const void* sin_table,
               const void* positions, int positions_i64, int64_t rows, int hq, int hkv, int d, int64_t q_stride_row,
               int64_t k_stride_row, int64_t q_out_stride_row, int64_t k_out_stride_row, int64_t table_rows, int dtype,
               void* stream) {
  if (!valid_dtype(dtype)) return set_error(HG_ERR_INVALID_ARGUMENT, "rope: unknown dtype %d", dtype);
  if (rows < 0 || hq < 0 || hkv < 0 || d < 2 || table_rows < 1)
    return set_error(HG_ERR_INVALID_ARGUMENT, "rope: bad sizes rows=%lld hq=%d hkv=%d d=%d table_rows=%lld", (long long)rows, hq, hkv,
                     d, (long long)table_rows);
  const int vec = dtype == HG_F32 ? 4 : 8;
  if (d % (2 * vec) != 0) return set_error(HG_ERR_UNSUPPORTED, "rope: head_dim %d must be a multiple of %d", d, 2 * vec);
  if (rows > 0 && hq > 0 && (q == nullptr || q_out == nullptr)) return set_error(HG_ERR_INVALID_ARGUMENT, "rope: null q");
  if (rows > 0 && hkv > 0 && (k == nullptr || k_out == nullptr)) return set_error(HG_ERR_INVALID_ARGUMENT, "rope: null k");
  if (rows > 0 && (cos_table == nullptr || sin_table == nullptr || positions == nullptr))
    return set_error(HG_ERR_INVALID_ARGUMENT, "rope: null table / positions");
  if (q_stride_row % vec || k_stride_row % vec || q_out_stride_row % vec || k_out_stride_row % vec ||
      reinterpret_cast<uintptr_t>(q) % 16 || reinterpret_cast<uintptr_t>(k) % 16 || reinterpret_cast<uintptr_t>(q_out) % 16 ||
      reinterpret_cast<uintptr_t>(k_out) % 16 || reinterpret_cast<uintptr_t>(cos_table) % 16 ||
      reinterpret_cast<uintptr_t>(sin_table) % 16)
    return set_error(HG_ERR_UNSUPPORTED, "rope: bases and row strides must be 16-byte aligned");
  if ((hq > 0 && (q_stride_row < (int64_t)hq * d || q_out_stride_row < (int64_t)hq * d)) ||
      (hkv > 0 && (k_stride_row < (int64_t)hkv * d || k_out_stride_row < (int64_t)hkv * d)))
    return set_error(HG_ERR_INVALID_ARGUMENT, "rope: a row stride is smaller than heads * head_dim");
  RopeParams p;
  memset(&p, 0, sizeof(p));
  p.q = q; p.k = k; p.q_out = q_out; p.k_out = k_out;
  p.cos = cos_table; p.sin = sin_table;
  p.positions = positions; p.positions_i64 = positions_i64;
  p.rows = rows; p.hq = hq; p.hkv = hkv; p.d = d;
  p.q_stride_row = q_stride_row; p.k_stride_row = k_stride_row;
  p.q_out_stride_row = q_out_stride_row; p.k_out_stride_row = k_out_stride_row;
  p.table_rows = table_rows;
  return launch_rope(p, dtype, (cudaStream_t)stream);
}

}  // extern "C"
