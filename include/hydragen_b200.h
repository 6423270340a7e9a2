/*
 * hydragen_b200 -- C ABI of the B200 (sm_100a) shared-prefix attention hot path.
 *
 * The reference (ScalingIntelligence/hydragen) has no FFI layer: its hot path is a set of
 * Python functions that call flash-attn's CUDA extension and three Triton kernels.  Each entry
 * point below replaces one of those calls; the Python shim in hydragen_b200/{attention,flash}.py
 * keeps the reference signatures and binds these symbols with ctypes (see INTEGRATION.md for the
 * stub a maintainer of the reference would add).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless its name ends in
 *     `_host` ; the library never allocates, frees or retains device memory (the caller owns
 *     outputs and the workspace) and never synchronises: all work is enqueued on `stream`
 *     (a cudaStream_t passed as void*), so every entry point is CUDA-graph capturable.
 *   - return value: 0 on success, negative hg_status otherwise; hg_last_error() returns a
 *     thread-local message.  No C++ exception crosses the boundary.
 *   - strides are in ELEMENTS of the tensor's dtype; the innermost (head_dim) stride is 1.
 *   - log-sum-exp tensors are fp32, natural log of sum exp(scale * q.k), laid out [b, nq, hq]
 *     (row = b*nq + qi, then head) -- the layout the reference's combine consumes
 *     (hydragen/attention.py:110-126), so no transpose pass is ever needed.
 *   - a query row with no valid key yields out = 0 and lse = -inf.
 */
#ifndef HYDRAGEN_B200_H_
#define HYDRAGEN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HG_ABI_VERSION 2
#define HG_MAX_COMBINE 8 /* max partial results merged by one call (shared levels + suffix) */

typedef enum hg_dtype {
  HG_F16 = 0,
  HG_BF16 = 1,
  HG_F32 = 2 /* CUDA-core kernels only (combine, rowwise attention); not the tcgen05 prefix kernel */
} hg_dtype;

typedef enum hg_status {
  HG_OK = 0,
  HG_ERR_INVALID_ARGUMENT = -1,
  HG_ERR_UNSUPPORTED = -2, /* valid request the kernels do not cover (e.g. head_dim) */
  HG_ERR_CUDA = -3,        /* launch / driver error; message holds cudaGetErrorString */
  HG_ERR_NOT_INITIALIZED = -4
} hg_status;

/* ABI version of the loaded library (== HG_ABI_VERSION). */
int hg_abi_version(void);

/* Thread-local message of the last failing call on this thread ("" if none). */
const char* hg_last_error(void);

/* Query the device once (SM count, shared memory, driver entry point for TMA descriptors).
 * Must be called once per process after the CUDA context of `device` exists and BEFORE any
 * stream capture; later calls are no-ops.  Replaces the per-call
 * torch.cuda.get_device_properties of hydragen/flash.py:193. */
int hg_init(int device);

/* Number of SMs seen by hg_init (0 before). */
int hg_sm_count(void);

/* ---------------------------------------------------------------------------------------
 * combine: out = sum_i w_i * outs[i] / sum_i w_i,  w_i = exp(lses[i] - max_i lses[i])
 * Replaces combine_lse / combine_lse_triton / combine_lse_torch (hydragen/attention.py:21-174)
 * for ANY n in [1, HG_MAX_COMBINE] (the reference's kernel handles n == 2 only and falls back
 * to eager torch otherwise, attention.py:169-174) and any head_dim.
 *   outs_host[i] : [rows, d] contiguous, dtype `dtype`   (rows = b * nq * hq)
 *   lses_host[i] : [rows] fp32
 *   out          : [rows, d] dtype `dtype`
 *   lse_out      : [rows] fp32 merged log-sum-exp, or NULL
 * `outs_host` / `lses_host` are HOST arrays of n device pointers (copied into the launch).
 */
int hg_combine_lse(const void* const* outs_host, const float* const* lses_host, int n, void* out,
                   float* lse_out, int64_t rows, int d, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------
 * Row-wise (CUDA-core, HBM-bound) attention: one query row against ONE sequence's own keys.
 * This is the suffix branch of the decomposition.  Replaces
 *   - flash_attention_seqlen  (hydragen/flash.py:163-281: Triton split-K kernel + reduce kernel +
 *     int32 cast of seq_len), when `seq_lens` is given;
 *   - flash_attention(q, k, v, causal=True) on per-sequence K/V (hydragen/attention.py:344),
 *     when `seq_lens` is NULL and `causal` = 1 (bottom-right aligned, flash-attn >= 2.1);
 * and, when n_partials > 0, ALSO the combine that follows it (attention.py:352): the suffix
 * result is merged in registers with the already computed prefix partials and only the final
 * output is written.
 *
 *   q        [b, nq, hq, d]     strides q_stride_b, q_stride_s, q_stride_h (elements)
 *   k, v     [b_kv, lk, hkv, d] strides kv_stride_b, kv_stride_s, kv_stride_h; sequence b reads
 *            batch entry (b / kv_group_size) -- kv_group_size = 1 for the suffix branch; > 1 lets
 *            the same kernel serve as a (slow, exact) shared-prefix path for dtypes and head dims
 *            the tcgen05 kernel does not take (fp32).
 *   cu_seqlens_k  NULL, or int32 [b / kv_group_size + 1]: k, v are then packed [total, hkv, d]
 *            (kv_stride_b ignored) and group g owns rows [cu[g], cu[g+1]) (flash-attn varlen).
 *   seq_lens NULL, or [b] valid key counts (int32 if seq_lens_i64 == 0, else int64 -- the
 *            reference's decode loop passes int64, hydragen/llama.py:569); keys >= seq_lens[b]
 *            are never read (xformers_stuff.py:274-279).
 *   causal   query qi sees keys j <= qi + (len - nq).
 *   out      [b, nq, hq, d] contiguous, dtype `dtype`;  lse [b, nq, hq] fp32 or NULL.
 *   partial_outs_host / partial_lses_host : HOST arrays of n_partials device pointers,
 *            each [b, nq, hq, d] contiguous (dtype) / [b, nq, hq] fp32, merged into out/lse.
 * d must be 64, 128 or 256 (16-bit dtypes) / 64 or 128 (fp32); hq % hkv == 0.
 */
int hg_rowwise_attn_fwd(const void* q, const void* k, const void* v, const void* seq_lens,
                        int seq_lens_i64, const int32_t* cu_seqlens_k, int kv_group_size, int causal,
                        void* out, float* lse, int b, int nq, int lk, int hq, int hkv, int d,
                        int64_t q_stride_b, int64_t q_stride_s, int64_t q_stride_h,
                        int64_t kv_stride_b, int64_t kv_stride_s, int64_t kv_stride_h,
                        const void* const* partial_outs_host, const float* const* partial_lses_host,
                        int n_partials, float sm_scale, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------
 * Shared-prefix attention on the 5th-generation tensor cores (tcgen05.mma, accumulators in
 * TMEM, operands staged by TMA): all sequences that share a prefix are batched into one Q
 * matrix per head and multiplied against the ONE copy of that prefix's K/V -- for EVERY shared
 * level of a hierarchy in one persistent launch (n_levels == 1 is forwarded to hg_prefix_attn_fwd's kernel).
 * Replaces, per level, flash_attention (hydragen/flash.py:284-306 -> flash-attn _flash_attn_forward, called
 * from hydragen/attention.py:270) or flash_attention_varlen (flash.py:309-351, called from
 * attention.py:313) together with the LSE transpose that follows it (attention.py:276-280,
 * 333-338): the kernel writes each level's LSE directly in [b, nq, hq].  (The reference loops over the
 * levels in Python, attention.py:250-341: one library launch + one transpose per level.)
 *
 *   q      [n_q_rows, hq, d] with row stride q_stride_row (elements); n_q_rows = b * nq.
 *   levels_host  HOST array of n_levels (1..4) descriptors (copied into the launch).  Level i splits the query
 *          rows into n_groups contiguous groups of n_q_rows / n_groups rows ("(n s) nq -> n (s nq)",
 *          attention.py:264-268); group g attends to
 *            rows [g*k_len, (g+1)*k_len) of k, v              when cu_seqlens_k == NULL,
 *            rows [cu_seqlens_k[g], cu_seqlens_k[g+1])         otherwise (int32 DEVICE array of n_groups+1
 *          entries read by the kernel -- no host sync, unlike SharedCache.fill's .item(), hydragen/llama.py:158-163;
 *          max_k_len = an upper bound on any group's length, used to balance the schedule),
 *          k, v [n_k_rows, hkv, d] with row stride kv_stride_row (rows past n_k_rows read as zero), and writes
 *          out [n_q_rows, hq, d] contiguous (dtype), lse [n_q_rows, hq] fp32 (may be NULL): the level's
 *          partial result, in the layout hg_combine_lse / hg_decode_attn_fused consume.
 *   workspace    hg_prefix_workspace_bytes() bytes of device memory, zero-initialised ONCE by the caller and
 *          then owned by the library's launches: where whole (group, 256-row tile, head) units would leave SMs
 *          idle for a good part of the launch, the persistent CTAs (one per SM) cut the (unit, key block)
 *          space of all levels into equal ranges (stream-K), and a unit cut between CTAs leaves fp32
 *          partial accumulators and flags there, merged inside the same launch.  Launches that may run
 *          concurrently need separate workspaces; launches on one stream share one.  NULL: units are never
 *          cut (whole (group, 256-row tile, head) units are dealt round-robin) -- correct, but a launch
 *          with few units then leaves SMs idle.
 * dtype HG_F16 or HG_BF16; d 64 or 128; hq % hkv == 0.
 */
typedef struct hg_prefix_level {
  const void* k;
  const void* v;
  void* out;
  float* lse;
  const int32_t* cu_seqlens_k;
  int64_t n_k_rows;
  int64_t kv_stride_row;
  int32_t n_groups;
  int32_t k_len;     /* uniform key count per group (cu_seqlens_k == NULL) */
  int32_t max_k_len; /* ragged levels: upper bound of any group's key count (0: n_k_rows) */
  int32_t reserved;
} hg_prefix_level;

int hg_prefix_attn_grouped_fwd(const void* q, int64_t n_q_rows, int64_t q_stride_row,
                               const hg_prefix_level* levels_host, int n_levels, int hq, int hkv, int d,
                               float sm_scale, int dtype, void* workspace, int64_t workspace_bytes,
                               void* stream);

/* Bytes of workspace hg_prefix_attn_grouped_fwd wants (independent of the problem). */
int64_t hg_prefix_workspace_bytes(void);

/* One level (the non-hierarchical case): group g owns query rows [g*q_per_group, (g+1)*q_per_group).  Runs the
 * one-CTA-per-unit form of the kernel (a grid of one CTA per (group, 256-row tile, head)). */
int hg_prefix_attn_fwd(const void* q, const void* k, const void* v, void* out, float* lse,
                       int n_groups, int q_per_group, int64_t n_k_rows, int k_len,
                       const int32_t* cu_seqlens_k, int max_k_len, int hq, int hkv, int d,
                       int64_t q_stride_row, int64_t kv_stride_row, float sm_scale, int dtype,
                       void* stream);

/* Split-KV form of hg_prefix_attn_fwd for launches with few (group, tile, head) work items -- the
 * head-parallel ranks of a tensor-parallel run (hydragen/tp.py:90-112) own Hq/N heads each: the keys of
 * every group are cut into kv_splits contiguous ranges, each handled by its own CTAs, and kv_splits
 * PARTIAL results are written back to back:
 *   out [kv_splits, n_q_rows, hq, d],  lse [kv_splits, n_q_rows, hq]   (a split with no keys: out 0, lse -inf)
 * to be merged by hg_combine_lse / the n_partials of hg_rowwise_attn_fwd / hg_decode_attn_fused (this is
 * flash-attn's split-KV, whose heuristic the reference copies in hydragen/flash.py:37-73, with the reduce
 * folded into the combine that follows anyway).  kv_splits == 1 is hg_prefix_attn_fwd. */
int hg_prefix_attn_split_fwd(const void* q, const void* k, const void* v, void* out, float* lse,
                             int n_groups, int q_per_group, int64_t n_k_rows, int k_len,
                             const int32_t* cu_seqlens_k, int max_k_len, int hq, int hkv, int d,
                             int64_t q_stride_row, int64_t kv_stride_row, float sm_scale, int dtype,
                             int kv_splits, void* stream);

/* Suggested kv_splits (>= 1, <= max_splits) for a one-level launch on the device seen by hg_init:
 * fills the SMs without going below 4 key blocks per CTA.  Host-side arithmetic only. */
int hg_prefix_suggest_splits(int n_groups, int q_per_group, int hq, int max_k_len, int max_splits);

/* Host-side view of the work schedule such a launch would use on a device with n_sms SMs (no device access;
 * pointers inside levels_host are ignored, a level is ragged iff max_k_len > 0).  Writes up to max_pieces
 * records of 10 int32 {cta, unit, level, head, group, row tile, first key block, end key block, split, slot}
 * and the grid size; returns the number of pieces (negative hg_status on error).  allow_split: 0 = whole units
 * only (what a launch without workspace does), 1 = what a launch with workspace does (units are cut only where
 * that shortens the launch: few units on many SMs, or a ragged last wave), 2 = always cut.  For tests and tooling. */
int hg_prefix_schedule(const hg_prefix_level* levels_host, int n_levels, int64_t n_q_rows, int hq, int n_sms,
                       int allow_split, int32_t* pieces_out, int max_pieces, int32_t* n_ctas_out);

/* ---------------------------------------------------------------------------------------
 * Causal self-attention of a prefill chunk on the tensor cores: the same tcgen05 kernel with a bottom-right
 * aligned causal mask (row i of a sequence sees keys j <= i + (sk - sq); flash-attn >= 2.1) -- only the key
 * blocks a row tile can see are streamed, the mask is applied in the diagonal blocks.
 * Replaces flash_attention(q, k, v, causal=True) (hydragen/flash.py:284-306) as called by the prefill branches
 * (hydragen/llama.py:509, 537-542, 546).
 *   q [b * sq, hq, d] row stride q_stride_row;  k, v [b * sk, hkv, d] row stride kv_stride_row (sequence i owns
 *   rows [i*sq, (i+1)*sq) / [i*sk, (i+1)*sk));  out [b * sq, hq, d] contiguous;  lse [b * sq, hq] fp32 or NULL.
 * dtype HG_F16 / HG_BF16, d 64 or 128, sk >= sq (no row without a visible key).
 */
int hg_causal_attn_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int b, int sq,
                       int sk, int hq, int hkv, int d, int64_t q_stride_row, int64_t kv_stride_row,
                       float sm_scale, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------
 * KV-cache append for one decode step ("next" row N1 of SURVEY.md 8f): writes the new key and
 * value row of every sequence at its own position.  Replaces the two scatter_ calls with a
 * fully expanded int64 index in PerLayerKVCache.update_per_completion_kvs
 * (hydragen/llama.py:236-262).
 *   k_new, v_new [b, nq, hkv, d] contiguous;  positions [b, nq] (int32 / int64), the row of the
 *   unique cache each new token goes to;  k_cache, v_cache [b_max, lk, hkv, d] contiguous.
 */
int hg_kv_append(const void* k_new, const void* v_new, const void* positions, int positions_i64,
                 void* k_cache, void* v_cache, int b, int nq, int lk, int hkv, int d, int dtype,
                 void* stream);

/* ---------------------------------------------------------------------------------------
 * One decode step of the suffix side in ONE launch: KV append + suffix attention + combine.
 * Replaces, for nq == 1 (every decode step: hydragen/llama.py:564-587),
 *   - PerLayerKVCache.update_per_completion_kvs (llama.py:236-262: two scatter_ calls),
 *   - flash_attention_seqlen (hydragen/flash.py:163-281: cast + split-K kernel + reduce kernel),
 *   - combine_lse (hydragen/attention.py:352),
 * with seq_len[b] = positions[b] + 1 (llama.py:569) computed in the kernel.  The new token's K/V
 * row is written to row positions[b] of the caches and attended to from registers; only the
 * positions[b] older rows of the cache are read.
 *   q            [b, 1, hq, d]   strides q_stride_b, q_stride_h (elements)
 *   k_new, v_new [b, 1, hkv, d]  contiguous
 *   positions    [b] int32 / int64, 0 <= positions[b] < lk
 *   k_cache, v_cache [b_max, lk, hkv, d] with strides kv_stride_b/s/h (written at row positions[b])
 *   out [b, 1, hq, d] contiguous; lse [b, 1, hq] fp32 or NULL; partial_* as in hg_rowwise_attn_fwd.
 * hq / hkv must be 1, 2, 4 or 8; d as for hg_rowwise_attn_fwd.
 */
int hg_decode_attn_fused(const void* q, const void* k_new, const void* v_new, const void* positions,
                         int positions_i64, void* k_cache, void* v_cache, void* out, float* lse, int b,
                         int lk, int hq, int hkv, int d, int64_t q_stride_b, int64_t q_stride_h,
                         int64_t kv_stride_b, int64_t kv_stride_s, int64_t kv_stride_h,
                         const void* const* partial_outs_host, const float* const* partial_lses_host,
                         int n_partials, float sm_scale, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------
 * In-place all-reduce(sum) of a buffer that every rank holds in symmetric memory mapped at one NVLink
 * multicast address (NVLS: the NVSwitch reduces and broadcasts).  Replaces the NCCL all-reduce of the
 * row-parallel o_proj output, one per attention layer (hydragen/tp.py:108-112), for the decode-step message
 * sizes (4-20 MiB), where a ring is latency-bound.
 *   mc_ptr     multicast address of the first byte to reduce (same offset on every rank; 16-byte aligned)
 *   out        NULL: in place, two-shot (rank r reduces slice r and multicasts it back; two barriers);
 *              else a private (non-symmetric) output buffer of nbytes: one-shot (every rank reduces the whole
 *              message through the switch; one barrier) -- the input buffer may be overwritten only after a
 *              later call of either form has completed
 *   flags_dev  DEVICE array of `world` pointers: flags_dev[p] = rank p's flag array (>= 128 uint32 words,
 *              zero-initialised once, never touched by the host afterwards), reachable through NVLink peer
 *              access.  Word 0 counts the collectives a rank has completed (the epoch); words 32+p / 64+p hold
 *              the epoch at which rank p announced "input in place" / "slice written everywhere" -- peers
 *              write them with one st.release.sys each, every rank polls only its own copy
 *   nbytes     message size, a multiple of 16;  n_blocks  CTAs to use (the same on every rank, <= SM count)
 * Every rank of the group must enqueue the same call in the same order.  CUDA-graph capturable. */
int hg_allreduce_multimem(void* mc_ptr, void* out, const void* flags_dev, int rank, int world,
                          int64_t nbytes, int dtype, int n_blocks, void* stream);

/* ---------------------------------------------------------------------------------------
 * Row-parallel o_proj GEMM fused with the all-reduce that follows it ("next" row N4 of SURVEY.md 8f): replaces
 * self.o_proj(attn_output) at hydragen/llama.py:592-594 on the column slice of the weight a rank holds
 * (hydragen/tp.py:99, RowwiseParallel) plus the funcol.all_reduce of hydragen/tp.py:108-112, in one persistent launch:
 * tcgen05 GEMM tiles written to this rank's symmetric buffer, per-tile "in place" flags raised in the owning rank's
 * memory, in-switch reduction (multimem.ld_reduce / multimem.st) of each tile as soon as every rank has produced it.
 *   x [m, k]      this rank's attention output (k = local heads * head_dim), row stride x_stride_row (elements)
 *   w [n, k]      this rank's slice of o_proj.weight (torch Linear layout: out_features x in_features), row stride w_stride_row
 *   out [m, n]    contiguous; world > 1: inside this rank's symmetric allocation, at the same offset on every rank;
 *                 holds sum over ranks of x_r w_r^T on return (rounded to `dtype` once per rank and once after the sum,
 *                 accumulated in fp32 -- exactly what a library GEMM followed by hg_allreduce_multimem produces)
 *   out_mc        multicast address of `out` (world > 1; else ignored)
 *   flags_dev     DEVICE array of `world` pointers to the ranks' flag arrays for THIS entry point (uint32, zeroed once,
 *                 peer-accessible; not shared with hg_allreduce_multimem); flag_words = their length, at least
 *                 hg_oproj_allreduce_flag_words(m, n, world).  Word 0: epoch; 1: CTA count; 32+p: "rank p's slices written
 *                 everywhere"; 128 + tile*world + p: "rank p's partial of tile `tile` is in place" (used in the owner's copy)
 *   world == 1    the GEMM alone (out_mc / flags_dev ignored)
 *   n_ctas        0 = one CTA per SM; else that many (<= SM count: the CTAs of a launch must be co-resident)
 * n, k and the row strides must be multiples of 8; dtype HG_BF16 / HG_F16.  Every rank of the group must enqueue the same
 * call (same m, n, world) in the same order.  CUDA-graph capturable. */
int hg_oproj_allreduce_fwd(const void* x, int64_t x_stride_row, const void* w, int64_t w_stride_row, void* out,
                           void* out_mc, const void* flags_dev, int64_t flag_words, int rank, int world, int64_t m,
                           int64_t n, int64_t k, int dtype, int n_ctas, void* stream);
int hg_oproj_allreduce_flag_words(int64_t m, int64_t n, int world);
/* Host-side view of that launch, for tests and tooling (no device work): the geometry hg_oproj_allreduce_fwd would use --
 * geometry_out[8] = {tile width, reductions per lane and slice, reduce warps per CTA, CTAs, tiles of the product, tiles and
 * slices rank `rank` owns, flag words needed} -- and, if cover_out != NULL (m * n / 8 int32, caller-zeroed), how many times
 * rank `rank` reduces each 16-byte vector of the [m, n] output (added into cover_out: summed over the ranks it must be 1
 * everywhere). */
int hg_oproj_allreduce_plan(int64_t m, int64_t n, int world, int rank, int n_ctas, int* geometry_out,
                            int32_t* cover_out);

/* ---------------------------------------------------------------------------------------
 * Rotary position embedding of the new q and k rows in one launch ("next" row N2 of SURVEY.md 8f): the
 * step right before the hot path.  Replaces apply_rotary_pos_emb as called at hydragen/llama.py:494-501
 * (transformers 4.37.2: cos[position_ids].unsqueeze(2); x * cos + rotate_half(x) * sin for q and for k --
 * a gather and ten elementwise launches per layer).  Half-split rotation: element i pairs with i + d/2.
 * The result is bit-identical to that eager evaluation in the activation dtype (each product and the sum
 * rounded to `dtype`, no FMA contraction).
 *   q [rows, hq, d], k [rows, hkv, d]  heads dense, row strides q_stride_row / k_stride_row (elements) -- e.g.
 *                      views of one fused qkv projection output; rows = b * s
 *   q_out, k_out       same shapes, own row strides; may alias q / k (in place)
 *   cos_table, sin_table [table_rows, d] contiguous, dtype `dtype` (the cached tables of
 *                      HydragenLlamaRotaryEmbedding, hydragen/llama.py:47-55)
 *   positions [rows]   absolute position of every row (int32 / int64), 0 <= positions[r] < table_rows
 * d must be a multiple of 16 (16-bit dtypes) / 8 (fp32); hq or hkv may be 0.
 */
int hg_rope_qk(const void* q, const void* k, void* q_out, void* k_out, const void* cos_table,
               const void* sin_table, const void* positions, int positions_i64, int64_t rows, int hq,
               int hkv, int d, int64_t q_stride_row, int64_t k_stride_row, int64_t q_out_stride_row,
               int64_t k_out_stride_row, int64_t table_rows, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HYDRAGEN_B200_H_ */
