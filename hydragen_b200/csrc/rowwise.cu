// Row-wise (CUDA-core) attention: the HBM-bound suffix branch of the decomposition.
//
// Replaces, in ONE launch, the reference's decode suffix path
//   hydragen/flash.py:163-281  flash_attention_seqlen  (int32 cast of seq_len, Triton
//       _fwd_kernel_splitK at xformers_stuff.py:189-428 writing fp32 partials, Triton
//       _splitK_reduce at flash.py:76-160 re-reading them)
// and, when prefix partials are supplied, the combine that follows it
//   hydragen/attention.py:352  combine_lse(outs, lses)
// as well as the causal per-sequence call hydragen/attention.py:344 (seq_lens == NULL).
//
// Design (B200): the work unit is (sequence b, kv head): a few KB of K and V, read exactly once.
// No tensor cores (one query row per kv head in MHA decode: a GEMV), no shared-memory staging
// (no reuse).  One warp (or WPI warps for long sequences) owns a unit.  A K/V row of D elements
// is covered by LPK = D/VEC lanes with one 128-bit load each, so a warp instruction fetches
// KPS = 32/LPK whole rows: fully used 32-byte sectors at any row stride.  Each lane group keeps
// its own online softmax (m, l, acc) over the keys it visits -- nothing crosses lanes inside the
// key loop except the LPK-lane dot-product reduction -- and the groups (and warps) are merged
// once at the end.  U steps are unrolled so that 2*U independent 128-bit loads per lane are in
// flight.  Keys >= seq_len[b] are never touched (xformers_stuff.py:274-279).
//
// Algorithmic bytes per unit: 2 * len_b * D * sizeof(T) (K, V) + D * sizeof(T) * g*nq * (2 + n_partials)
// (q, out, partial outs) + 4 * g*nq * (1 + n_partials) (LSEs).  Bound: HBM.
#include <cstdlib>

#include "common.cuh"

namespace hg {

template <int VEC>
struct RowState {
  float m;  // running max, log2 domain (scores already multiplied by scale*log2e)
  float l;
  float acc[VEC];
};

template <int VEC>
__device__ __forceinline__ void merge_state(RowState<VEC>& a, float m_o, float l_o, const float* acc_o) {
  const float m_new = fmaxf(a.m, m_o);
  const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
  const float sa = fast_exp2(a.m - m_safe);
  const float so = fast_exp2(m_o - m_safe);
  a.l = a.l * sa + l_o * so;
#pragma unroll
  for (int e = 0; e < VEC; ++e) a.acc[e] = a.acc[e] * sa + acc_o[e] * so;
  a.m = m_new;
}

constexpr int kRowwiseWarps = 4;

template <typename T, int D, int WPI, int R>
__global__ void __launch_bounds__(kRowwiseWarps * 32) rowwise_attn_kernel(const RowwiseParams p) {
  constexpr int VEC = Vec16<T>::VEC;
  constexpr int LPK = D / VEC;  // lanes per key row
  static_assert(LPK >= 1 && LPK <= 32 && (LPK & (LPK - 1)) == 0, "head_dim not supported for this dtype");
  constexpr int KPS = 32 / LPK;  // key rows per warp step
  constexpr int U = 4;           // unrolled steps -> 2*U 128-bit loads in flight per lane
  constexpr int ITEMS = kRowwiseWarps / WPI;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / LPK, dl = lane % LPK;
  const int item_in_cta = warp / WPI, wi = warp % WPI;
  const int64_t item = (int64_t)blockIdx.x * ITEMS + item_in_cta;
  const int64_t n_items = (int64_t)p.b * p.hkv;
  const bool item_valid = item < n_items;
  const int b_idx = item_valid ? (int)(item / p.hkv) : 0;
  const int kvh = item_valid ? (int)(item % p.hkv) : 0;
  const int g = p.hq / p.hkv;
  const int M = p.nq * g;

  // ---- this unit's keys ---------------------------------------------------------------
  int len = 0;
  const T* kbase = reinterpret_cast<const T*>(p.k);
  const T* vbase = reinterpret_cast<const T*>(p.v);
  if (item_valid) {
    const int grp = b_idx / p.kv_group_size;
    int64_t off;
    if (p.cu_seqlens_k != nullptr) {
      const int s0 = __ldg(p.cu_seqlens_k + grp), s1 = __ldg(p.cu_seqlens_k + grp + 1);
      len = s1 - s0;
      off = (int64_t)s0 * p.kv_stride_s;
    } else {
      len = p.lk;
      off = (int64_t)grp * p.kv_stride_b;
    }
    off += (int64_t)kvh * p.kv_stride_h + dl * VEC;
    kbase += off;
    vbase += off;
    if (p.seq_lens != nullptr) {
      const int64_t sl = p.seq_lens_i64 ? reinterpret_cast<const int64_t*>(p.seq_lens)[b_idx]
                                        : (int64_t) reinterpret_cast<const int32_t*>(p.seq_lens)[b_idx];
      len = (int)max((int64_t)0, min((int64_t)len, sl));
    }
  }
  const bool single_pass = M <= R;

  __shared__ float s_merge[(WPI > 1) ? ITEMS * (WPI - 1) * R * (D + 2) : 1];

  for (int r0 = 0; r0 < M; r0 += R) {
    // ---- query rows of this pass --------------------------------------------------------
    float qf[R][VEC];
    int limit[R];
    int pass_limit = 0;
    RowState<VEC> st[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int rr = r0 + r;
      const int qi = rr / g, jh = rr % g;
      limit[r] = 0;
#pragma unroll
      for (int e = 0; e < VEC; ++e) qf[r][e] = 0.f;
      if (item_valid && rr < M) {
        limit[r] = p.causal ? max(0, min(len, qi + len - p.nq + 1)) : len;
        const T* qp = reinterpret_cast<const T*>(p.q) + (int64_t)b_idx * p.q_stride_b + (int64_t)qi * p.q_stride_s +
                      (int64_t)(kvh * g + jh) * p.q_stride_h + dl * VEC;
        Vec16<T>::unpack(ld_v4(qp), qf[r]);
      }
      pass_limit = max(pass_limit, limit[r]);
      st[r].m = -INFINITY;
      st[r].l = 0.f;
#pragma unroll
      for (int e = 0; e < VEC; ++e) st[r].acc[e] = 0.f;
    }

    // ---- key loop ----------------------------------------------------------------------
    for (int base = 0; base < pass_limit; base += U * WPI * KPS) {
      uint4 kraw[U], vraw[U];
      int key[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        key[u] = base + (u * WPI + wi) * KPS + sub;
        if (key[u] < pass_limit) {
          const int64_t o = (int64_t)key[u] * p.kv_stride_s;
          kraw[u] = single_pass ? ld_stream_v4(kbase + o) : ld_v4(kbase + o);
        } else {
          kraw[u] = make_uint4(0, 0, 0, 0);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (key[u] < pass_limit) {
          const int64_t o = (int64_t)key[u] * p.kv_stride_s;
          vraw[u] = single_pass ? ld_stream_v4(vbase + o) : ld_v4(vbase + o);
        } else {
          vraw[u] = make_uint4(0, 0, 0, 0);
        }
      }
      float s[U][R];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float kf[VEC];
        Vec16<T>::unpack(kraw[u], kf);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          float d = 0.f;
#pragma unroll
          for (int e = 0; e < VEC; ++e) d = fmaf(qf[r][e], kf[e], d);
#pragma unroll
          for (int o = LPK / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
          s[u][r] = (key[u] < limit[r]) ? d * p.scale_log2 : -INFINITY;
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float m_new = st[r].m;
#pragma unroll
        for (int u = 0; u < U; ++u) m_new = fmaxf(m_new, s[u][r]);
        const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
        const float alpha = fast_exp2(st[r].m - m_safe);
        float pu[U];
        float psum = 0.f;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          pu[u] = fast_exp2(s[u][r] - m_safe);
          psum += pu[u];
        }
        st[r].l = st[r].l * alpha + psum;
        st[r].m = m_new;
#pragma unroll
        for (int e = 0; e < VEC; ++e) st[r].acc[e] *= alpha;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          float vf[VEC];
          Vec16<T>::unpack(vraw[u], vf);
#pragma unroll
          for (int e = 0; e < VEC; ++e) st[r].acc[e] = fmaf(pu[u], vf[e], st[r].acc[e]);
        }
      }
    }

    // ---- merge the KPS lane groups of the warp -------------------------------------------
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int o = LPK; o < 32; o <<= 1) {
        const float m_o = __shfl_xor_sync(0xffffffffu, st[r].m, o);
        const float l_o = __shfl_xor_sync(0xffffffffu, st[r].l, o);
        float acc_o[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc_o[e] = __shfl_xor_sync(0xffffffffu, st[r].acc[e], o);
        merge_state<VEC>(st[r], m_o, l_o, acc_o);
      }
    }

    // ---- merge the WPI warps of the unit through shared memory ---------------------------
    if constexpr (WPI > 1) {
      __syncthreads();  // previous pass finished reading s_merge
      if (wi > 0 && sub == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          float* dst = s_merge + ((item_in_cta * (WPI - 1) + (wi - 1)) * R + r) * (D + 2);
          if (dl == 0) {
            dst[D] = st[r].m;
            dst[D + 1] = st[r].l;
          }
#pragma unroll
          for (int e = 0; e < VEC; ++e) dst[dl * VEC + e] = st[r].acc[e];
        }
      }
      __syncthreads();
      if (wi == 0 && sub == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          for (int w = 0; w < WPI - 1; ++w) {
            const float* src = s_merge + ((item_in_cta * (WPI - 1) + w) * R + r) * (D + 2);
            float acc_o[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc_o[e] = src[dl * VEC + e];
            merge_state<VEC>(st[r], src[D], src[D + 1], acc_o);
          }
        }
      }
    }

    // ---- epilogue: normalise, merge with the prefix partials, store -----------------------
    if (item_valid && wi == 0 && sub == 0) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int rr = r0 + r;
        if (rr >= M) continue;
        const int qi = rr / g, jh = rr % g;
        const int64_t orow = ((int64_t)b_idx * p.nq + qi) * p.hq + (kvh * g + jh);
        const float l = st[r].l;
        float lse = (l > 0.f) ? (st[r].m + fast_log2(l)) * kLn2 : -INFINITY;
        const float inv_l = (l > 0.f) ? 1.f / l : 0.f;
        float o[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) o[e] = st[r].acc[e] * inv_l;
        const int np = p.partials.n;
        if (np > 0) {
          float pl[HG_MAX_COMBINE];
          uint4 praw[HG_MAX_COMBINE];
          float mx = lse;
#pragma unroll
          for (int i = 0; i < HG_MAX_COMBINE; ++i) {
            if (i < np) {
              pl[i] = __ldg(p.partials.lses[i] + orow);
              praw[i] = ld_stream_v4(reinterpret_cast<const T*>(p.partials.outs[i]) + orow * D + dl * VEC);
              mx = fmaxf(mx, pl[i]);
            }
          }
          const float mx_safe = (mx == -INFINITY) ? 0.f : mx;
          const float w_s = __expf(lse - mx_safe);
          float den = w_s;
#pragma unroll
          for (int e = 0; e < VEC; ++e) o[e] *= w_s;
#pragma unroll
          for (int i = 0; i < HG_MAX_COMBINE; ++i) {
            if (i < np) {
              const float w = __expf(pl[i] - mx_safe);
              den += w;
              float f[VEC];
              Vec16<T>::unpack(praw[i], f);
#pragma unroll
              for (int e = 0; e < VEC; ++e) o[e] = fmaf(w, f[e], o[e]);
            }
          }
          const float inv = den > 0.f ? 1.f / den : 0.f;
#pragma unroll
          for (int e = 0; e < VEC; ++e) o[e] *= inv;
          lse = den > 0.f ? mx_safe + __logf(den) : -INFINITY;
        }
        st_v4(reinterpret_cast<T*>(p.out) + orow * D + dl * VEC, Vec16<T>::pack(o));
        if (p.lse != nullptr && dl == 0) p.lse[orow] = lse;
      }
    }
  }
}

// ---- decode fast path: one LPK-lane slot per (sequence, kv head) ------------------------------
// The shape that every decode step of the reference produces (hydragen/llama.py:564-587): nq == 1,
// each (sequence, kv head) unit owns `len_b` keys.  A slot of LPK = D/VEC lanes (16 for d = 128,
// 16-bit) owns one unit and walks its keys in order, U keys (2U independent 128-bit loads per lane)
// per trip; the R = Hq/Hkv query rows of the unit share every K/V row.  Adjacent slots are adjacent
// kv heads of the same cache row, so a warp load covers whole contiguous 256-byte head rows.  No
// shared memory, no cross-slot merge, 16 units per 256-thread CTA.  Full U-key trips run without
// predicates; the ragged tail is one more trip whose key slots are skipped warp-uniformly, so a
// sequence with a single key executes a single key's worth of instructions.
//
// kFused: the step's KV append (hydragen/llama.py:250-257) happens here as well.  The new token's
// K/V row is read once from k_new / v_new, stored to row positions[b] of the caches and attended to
// from registers as the last key (len_b = positions[b] + 1); the cache is only read for the
// len_b - 1 older keys.  One launch then replaces scatter_ x2 + cast + split-K + reduce + combine.
template <typename T, int D, int R, int U, int MINB, bool kFused>
__global__ void __launch_bounds__(256, MINB) decode_slot_kernel(const RowwiseParams p) {
  constexpr int VEC = Vec16<T>::VEC;
  constexpr int LPK = D / VEC;
  static_assert(LPK >= 1 && LPK <= 32 && (LPK & (LPK - 1)) == 0, "head_dim not supported for this dtype");
  constexpr int SLOTS = 256 / LPK;
  const int tid = threadIdx.x;
  const int dl = tid % LPK;
  const int64_t unit = (int64_t)blockIdx.x * SLOTS + tid / LPK;
  const int64_t n_units = (int64_t)p.b * p.hkv;
  const bool valid = unit < n_units;
  const int b_idx = valid ? (int)(unit / p.hkv) : 0;
  const int kvh = valid ? (int)(unit % p.hkv) : 0;
  // the next launch on the stream (the next layer's prefix kernel) may begin its set-up while this grid drains
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // keys read from the cache: [0, len); fused: the new token sits at row `len` and is not re-read
  int len = 0;
  if (valid) {
    if constexpr (kFused) {
      const int64_t pos = p.positions_i64 ? __ldg(reinterpret_cast<const int64_t*>(p.positions) + b_idx)
                                          : (int64_t)__ldg(reinterpret_cast<const int32_t*>(p.positions) + b_idx);
      len = (int)max((int64_t)0, min((int64_t)p.lk - 1, pos));
    } else {
      len = p.lk;
      if (p.seq_lens != nullptr) {
        const int64_t sl = p.seq_lens_i64 ? __ldg(reinterpret_cast<const int64_t*>(p.seq_lens) + b_idx)
                                          : (int64_t)__ldg(reinterpret_cast<const int32_t*>(p.seq_lens) + b_idx);
        len = (int)max((int64_t)0, min((int64_t)len, sl));
      }
    }
  }
  // first output row of the unit: rows (b, 0, kvh*R + r), r < R, are contiguous
  const int64_t orow0 = (int64_t)b_idx * p.hq + (int64_t)kvh * R;
  const int np = p.partials.n;

  // independent of len: the query rows and the new K/V row are requested before the key loop so their
  // latency overlaps it.  (The prefix partials are only touched after the key loop, behind
  // griddepcontrol.wait: this kernel is launched as the programmatic dependent of the prefix launch and
  // does its own K/V work while that one drains.)
  uint4 qraw[R];
  uint4 knew = make_uint4(0, 0, 0, 0), vnew = make_uint4(0, 0, 0, 0);
  if constexpr (kFused) {
    if (valid) {
      const int64_t o = ((int64_t)b_idx * p.hkv + kvh) * D + dl * VEC;
      knew = ld_stream_v4(reinterpret_cast<const T*>(p.k_new) + o);
      vnew = ld_stream_v4(reinterpret_cast<const T*>(p.v_new) + o);
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    qraw[r] = make_uint4(0, 0, 0, 0);
    if (valid)
      qraw[r] = ld_v4(reinterpret_cast<const T*>(p.q) + (int64_t)b_idx * p.q_stride_b + (int64_t)(kvh * R + r) * p.q_stride_h + dl * VEC);
  }
  int len_max = len, len_min = valid ? len : 0x7fffffff;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    len_max = max(len_max, __shfl_xor_sync(0xffffffffu, len_max, o));
    len_min = min(len_min, __shfl_xor_sync(0xffffffffu, len_min, o));
  }
  len_min = min(len_min, len_max);  // a warp with no valid slot: 0

  const int64_t kv_off = (int64_t)b_idx * p.kv_stride_b + (int64_t)kvh * p.kv_stride_h + dl * VEC;
  const T* kb = reinterpret_cast<const T*>(p.k) + kv_off;
  const T* vb = reinterpret_cast<const T*>(p.v) + kv_off;
  if constexpr (kFused) {
    if (valid) {  // the append: row `len` of this sequence's caches
      st_v4(const_cast<T*>(kb) + (int64_t)len * p.kv_stride_s, knew);
      st_v4(const_cast<T*>(vb) + (int64_t)len * p.kv_stride_s, vnew);
    }
  }

  float qf[R][VEC];
  RowState<VEC> st[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    Vec16<T>::unpack(qraw[r], qf[r]);
#pragma unroll
    for (int e = 0; e < VEC; ++e) qf[r][e] *= p.scale_log2;  // scores come out in the log2 domain
    st[r].m = -INFINITY;
    st[r].l = 0.f;
#pragma unroll
    for (int e = 0; e < VEC; ++e) st[r].acc[e] = 0.f;
  }

  // one key (already in registers) into the running state of every query row of the unit
  auto absorb = [&](const uint4& kraw, const uint4& vraw, bool live) {
    float kf[VEC], vf[VEC];
    Vec16<T>::unpack(kraw, kf);
    Vec16<T>::unpack(vraw, vf);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float d = 0.f;
#pragma unroll
      for (int e = 0; e < VEC; ++e) d = fmaf(qf[r][e], kf[e], d);
#pragma unroll
      for (int o = LPK / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      const float sc = live ? d : -INFINITY;
      const float m_new = fmaxf(st[r].m, sc);
      const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = fast_exp2(st[r].m - m_safe);
      const float pw = fast_exp2(sc - m_safe);
      st[r].l = st[r].l * alpha + pw;
      st[r].m = m_new;
#pragma unroll
      for (int e = 0; e < VEC; ++e) st[r].acc[e] = fmaf(pw, vf[e], st[r].acc[e] * alpha);
    }
  };

  // ---- full trips: U keys, no predicates -------------------------------------------------------
  int j0 = 0;
  for (; j0 + U <= len_min; j0 += U) {
    uint4 kraw[U], vraw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) kraw[u] = ld_stream_v4(kb + (int64_t)(j0 + u) * p.kv_stride_s);
#pragma unroll
    for (int u = 0; u < U; ++u) vraw[u] = ld_stream_v4(vb + (int64_t)(j0 + u) * p.kv_stride_s);
    float s[U][R];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float kf[VEC];
      Vec16<T>::unpack(kraw[u], kf);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float d = 0.f;
#pragma unroll
        for (int e = 0; e < VEC; ++e) d = fmaf(qf[r][e], kf[e], d);
#pragma unroll
        for (int o = LPK / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        s[u][r] = d;
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float m_new = st[r].m;
#pragma unroll
      for (int u = 0; u < U; ++u) m_new = fmaxf(m_new, s[u][r]);
      const float alpha = fast_exp2(st[r].m - m_new);  // scores are finite here: m_new > -inf
      float pu[U];
      float psum = 0.f;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        pu[u] = fast_exp2(s[u][r] - m_new);
        psum += pu[u];
      }
      st[r].l = st[r].l * alpha + psum;
      st[r].m = m_new;
#pragma unroll
      for (int e = 0; e < VEC; ++e) st[r].acc[e] *= alpha;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float vf[VEC];
        Vec16<T>::unpack(vraw[u], vf);
#pragma unroll
        for (int e = 0; e < VEC; ++e) st[r].acc[e] = fmaf(pu[u], vf[e], st[r].acc[e]);
      }
    }
  }
  // ---- ragged tail: fewer than U keys for some slot; all loads first, then key by key ------------
  for (; j0 < len_max; j0 += U) {
    uint4 kraw[U], vraw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      kraw[u] = make_uint4(0, 0, 0, 0);
      vraw[u] = make_uint4(0, 0, 0, 0);
      if (j0 + u < len) {
        kraw[u] = ld_stream_v4(kb + (int64_t)(j0 + u) * p.kv_stride_s);
        vraw[u] = ld_stream_v4(vb + (int64_t)(j0 + u) * p.kv_stride_s);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (j0 + u < len_max) absorb(kraw[u], vraw[u], j0 + u < len);  // warp-uniform skip
  }
  if constexpr (kFused) absorb(knew, vnew, valid);

  // everything above is independent of the prefix launch(es); their partial results are read from here on
  if (np > 0) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (!valid) return;
  uint4 p0raw[R];
  float p0lse[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    p0raw[r] = make_uint4(0, 0, 0, 0);
    p0lse[r] = -INFINITY;
    if (np > 0) {
      p0raw[r] = ld_v4(reinterpret_cast<const T*>(p.partials.outs[0]) + (orow0 + r) * D + dl * VEC);
      p0lse[r] = *(p.partials.lses[0] + orow0 + r);
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int64_t orow = orow0 + r;
    const float l = st[r].l;
    float lse = (l > 0.f) ? (st[r].m + fast_log2(l)) * kLn2 : -INFINITY;
    const float inv_l = (l > 0.f) ? 1.f / l : 0.f;
    float o[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) o[e] = st[r].acc[e] * inv_l;
    if (np > 0) {
      // merge with the prefix partials: first one prefetched, further shared levels one by one
      float mx = fmaxf(lse, p0lse[r]);
      for (int i = 1; i < np; ++i) mx = fmaxf(mx, *(p.partials.lses[i] + orow));
      const float mx_safe = (mx == -INFINITY) ? 0.f : mx;
      const float w_s = __expf(lse - mx_safe);
      float den = w_s;
#pragma unroll
      for (int e = 0; e < VEC; ++e) o[e] *= w_s;
      {
        const float w = __expf(p0lse[r] - mx_safe);
        den += w;
        float f[VEC];
        Vec16<T>::unpack(p0raw[r], f);
#pragma unroll
        for (int e = 0; e < VEC; ++e) o[e] = fmaf(w, f[e], o[e]);
      }
      for (int i = 1; i < np; ++i) {
        const float w = __expf(*(p.partials.lses[i] + orow) - mx_safe);
        den += w;
        float f[VEC];
        Vec16<T>::unpack(ld_v4(reinterpret_cast<const T*>(p.partials.outs[i]) + orow * D + dl * VEC), f);
#pragma unroll
        for (int e = 0; e < VEC; ++e) o[e] = fmaf(w, f[e], o[e]);
      }
      const float inv = den > 0.f ? 1.f / den : 0.f;
#pragma unroll
      for (int e = 0; e < VEC; ++e) o[e] *= inv;
      lse = den > 0.f ? mx_safe + __logf(den) : -INFINITY;
    }
    st_v4(reinterpret_cast<T*>(p.out) + orow * D + dl * VEC, Vec16<T>::pack(o));
    if (p.lse != nullptr && dl == 0) p.lse[orow] = lse;
  }
}

template <typename T, int D, int R>
static int launch_decode_slot(const RowwiseParams& p, cudaStream_t s) {
  constexpr int VEC = Vec16<T>::VEC;
  constexpr int SLOTS = 256 / (D / VEC);
  constexpr int MINB = (R == 1 ? 3 : (R <= 4 ? 2 : 1));
  const int64_t n_units = (int64_t)p.b * p.hkv;
  const int64_t blocks = (n_units + SLOTS - 1) / SLOTS;
  if (blocks > 0x7fffffffLL) return set_error(HG_ERR_UNSUPPORTED, "rowwise: too many units (%lld)", (long long)n_units);
  // With prefix partials to merge this launch is the programmatic dependent of the launch before it on the
  // stream (the prefix kernel): it may start while that one drains and blocks at griddepcontrol.wait only
  // before it reads the partials.
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)blocks);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (p.partials.n > 0 && pdl_enabled()) ? 1 : 0;
  cudaError_t e;
  // Short caches (the first decode steps, the microbenchmark): the work per unit is a handful of dependent
  // memory round trips, so what matters is how many units are resident -- 2 keys per trip, half the registers.
  const bool small = p.lk <= 32 && R <= 2;
  if (p.k_new != nullptr)
    e = small ? cudaLaunchKernelEx(&cfg, decode_slot_kernel<T, D, R, 2, (R == 1 ? 4 : 3), true>, p)
              : cudaLaunchKernelEx(&cfg, decode_slot_kernel<T, D, R, 4, MINB, true>, p);
  else
    e = small ? cudaLaunchKernelEx(&cfg, decode_slot_kernel<T, D, R, 2, (R == 1 ? 4 : 3), false>, p)
              : cudaLaunchKernelEx(&cfg, decode_slot_kernel<T, D, R, 4, MINB, false>, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(HG_ERR_CUDA, "decode_slot_attn: launch failed: %s", cudaGetErrorString(e));
  }
  return check_launch("decode_slot_attn");
}

template <typename T, int D, int WPI, int R>
static int launch_rowwise_inst(const RowwiseParams& p, cudaStream_t s) {
  constexpr int ITEMS = kRowwiseWarps / WPI;
  const int64_t n_items = (int64_t)p.b * p.hkv;
  const int64_t blocks = (n_items + ITEMS - 1) / ITEMS;
  if (blocks > 0x7fffffffLL) return set_error(HG_ERR_UNSUPPORTED, "rowwise: too many units (%lld)", (long long)n_items);
  rowwise_attn_kernel<T, D, WPI, R><<<(unsigned)blocks, kRowwiseWarps * 32, 0, s>>>(p);
  return check_launch("rowwise_attn");
}

template <typename T, int D>
static int launch_rowwise_d(const RowwiseParams& p, cudaStream_t s) {
  const int M = p.nq * (p.hq / p.hkv);
  // decode shape with enough independent units (or few keys): the slot kernel
  const int64_t n_units = (int64_t)p.b * p.hkv;
  const bool fused = p.k_new != nullptr;
  if (fused && !(p.nq == 1 && p.cu_seqlens_k == nullptr && p.kv_group_size == 1 && (M == 1 || M == 2 || M == 4 || M == 8)))
    return set_error(HG_ERR_UNSUPPORTED, "decode_attn_fused: needs nq == 1 and hq/hkv in {1, 2, 4, 8}");
  if (p.nq == 1 && p.cu_seqlens_k == nullptr && p.kv_group_size == 1 && (fused || n_units >= 2048 || p.lk <= 64)) {
    switch (M) {
      case 1: return launch_decode_slot<T, D, 1>(p, s);
      case 2: return launch_decode_slot<T, D, 2>(p, s);
      case 4: return launch_decode_slot<T, D, 4>(p, s);
      case 8: return launch_decode_slot<T, D, 8>(p, s);
      default: break;
    }
  }
  // long per-sequence KV: split the keys of one unit over the 4 warps of the CTA
  const bool wide = p.lk > 256;
  if (M == 1) return wide ? launch_rowwise_inst<T, D, 4, 1>(p, s) : launch_rowwise_inst<T, D, 1, 1>(p, s);
  return wide ? launch_rowwise_inst<T, D, 4, 4>(p, s) : launch_rowwise_inst<T, D, 1, 4>(p, s);
}

template <typename T>
static int launch_rowwise_t(const RowwiseParams& p, cudaStream_t s) {
  constexpr int VEC = Vec16<T>::VEC;
  switch (p.d) {
    case 64: return launch_rowwise_d<T, 64>(p, s);
    case 128: return launch_rowwise_d<T, 128>(p, s);
    case 256:
      if constexpr (256 / VEC <= 32) return launch_rowwise_d<T, 256>(p, s);
      // fallthrough
    default: return set_error(HG_ERR_UNSUPPORTED, "rowwise: head_dim %d not supported for this dtype", p.d);
  }
}

int launch_rowwise(const RowwiseParams& p, int dtype, cudaStream_t s) {
  if (p.b == 0 || p.nq == 0) return HG_OK;
  switch (dtype) {
    case HG_F16: return launch_rowwise_t<__half>(p, s);
    case HG_BF16: return launch_rowwise_t<__nv_bfloat16>(p, s);
    case HG_F32: return launch_rowwise_t<float>(p, s);
    default: return set_error(HG_ERR_INVALID_ARGUMENT, "rowwise: unknown dtype %d", dtype);
  }
}

// ---- KV append ----------------------------------------------------------
// Replaces the scatter_ pair of hydragen/llama.py:250-257: one 128-bit copy per thread, the
// destination row taken from positions[b, qi]; no index tensor is materialised.
template <int ELT>
__global__ void __launch_bounds__(256) kv_append_kernel(const uint4* __restrict__ k_new, const uint4* __restrict__ v_new,
                                                        const void* __restrict__ positions, int positions_i64,
                                                        uint4* __restrict__ k_cache, uint4* __restrict__ v_cache,
                                                        int64_t n_tokens, int nq, int lk, int row_vecs) {
  const int64_t total = n_tokens * row_vecs;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t tok = idx / row_vecs;
    const int c = (int)(idx % row_vecs);
    const int64_t b = tok / nq;
    const int64_t pos = positions_i64 ? reinterpret_cast<const int64_t*>(positions)[tok]
                                      : (int64_t) reinterpret_cast<const int32_t*>(positions)[tok];
    if (pos < 0 || pos >= lk) continue;
    const int64_t dst = (b * lk + pos) * row_vecs + c;
    k_cache[dst] = k_new[idx];
    v_cache[dst] = v_new[idx];
  }
}

int launch_kv_append(const void* k_new, const void* v_new, const void* positions, int positions_i64, void* k_cache,
                     void* v_cache, int b, int nq, int lk, int hkv, int d, int dtype, cudaStream_t s) {
  const int esz = (dtype == HG_F32) ? 4 : 2;
  const int64_t row_bytes = (int64_t)hkv * d * esz;
  if (row_bytes % 16 != 0) return set_error(HG_ERR_UNSUPPORTED, "kv_append: hkv*d*sizeof must be a multiple of 16");
  const int row_vecs = (int)(row_bytes / 16);
  const int64_t n_tokens = (int64_t)b * nq;
  if (n_tokens == 0) return HG_OK;
  const int64_t total = n_tokens * row_vecs;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)(device_info().sm_count > 0 ? device_info().sm_count : 148) * 8;
  if (blocks > cap) blocks = cap;
  kv_append_kernel<2><<<(unsigned)blocks, 256, 0, s>>>((const uint4*)k_new, (const uint4*)v_new, positions, positions_i64,
                                                       (uint4*)k_cache, (uint4*)v_cache, n_tokens, nq, lk, row_vecs);
  return check_launch("kv_append");
}

}  // namespace hg
