"""Generate tests/golden/hydragen_golden.npz by running THE REFERENCE'S OWN operator code.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

What runs: /root/reference/hydragen/attention.py's ``hydragen_attention`` and
``combine_lse_torch`` -- unmodified, imported from where they lie -- on the case list of
/root/reference/tests/test_attention.py:26-32 (plus BASELINE.json's cfg#1 toy shape and a
few extra shapes the reference never tests).  The only substitution: the three attention
primitives the reference gets from CUDA-only third-party code (flash-attn v2.3.6 and its
xformers-derived Triton kernel: hydragen/flash.py:284-351, 163-281) are replaced, in the
reference module's namespace, by the fp64 CPU primitives of oracle/hydragen_oracle.py, and
the 2-input Triton combine (attention.py:105-151) by the reference's own
``combine_lse_torch`` (the equality the reference's tests/test_combine_lse.py pins).

So the goldens pin the in-tree logic (inter-sequence batching order, LSE layouts,
varlen bookkeeping, early return, n-way combine) against the reference itself; the
softmax-attention primitive is pinned by its mathematical definition.

Inputs are NOT stored (the 16384-long prefix case alone would be 64 MB): they are regenerated
from ``oracle.hydragen_oracle.build_case`` (CPU generator, fixed seed); a float64 checksum of
the inputs is stored beside every output to catch generator drift.
"""

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")

from oracle import hydragen_oracle as O  # noqa: E402

import hydragen.attention as ref  # noqa: E402  (the reference, unmodified)

F64 = torch.float64


def _flash_attention(q, k, v, causal=False):
    out, lse = O.flash_attention(q, k, v, causal=causal, compute_dtype=F64)
    return out, lse


def _flash_attention_varlen(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, causal=False):
    return O.flash_attention_varlen(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, causal=causal, compute_dtype=F64)


def _flash_attention_seqlen(q, k, v, seq_len=None):
    return O.flash_attention_seqlen(q, k, v, seq_len=seq_len, compute_dtype=F64)


ref.flash_attention = _flash_attention
ref.flash_attention_varlen = _flash_attention_varlen
ref.flash_attention_seqlen = _flash_attention_seqlen
ref.combine_lse_triton = lambda o1, l1, o2, l2: ref.combine_lse_torch([o1, o2], [l1, l2])


from make_golden_cases import case_list, checksum, DT  # noqa: E402


def main():
    out = {}
    for name, sizes, hq, hkv, d, dt, seed, nq in case_list():
        c = O.build_case(sizes, hq, hkv, d, dtype=DT[dt], seed=seed, nq=nq)
        # the reference's own operator (attention.py:177-354), fp64 primitives
        args = dict(c)
        ref_out = ref.hydragen_attention(
            q=args["q"].double(), k=args["k"].double(), v=args["v"].double(),
            shared_ks=[x.double() for x in args["shared_ks"]], shared_vs=[x.double() for x in args["shared_vs"]],
            shared_cu_seq_lens=args["shared_cu_seq_lens"], shared_max_seq_lens=args["shared_max_seq_lens"],
            use_varlens=args["use_varlens"], seq_lens=args["seq_lens"],
        )
        # the reference test's ground truth (test_attention.py:132-178)
        truth = O.concat_attention(c["q"], c["k"], c["v"], c["shared_ks"], c["shared_vs"], c["shared_cu_seq_lens"], c["use_varlens"], c["seq_lens"])
        err = (ref_out - truth).abs().max().item()
        assert err < 1e-9, (name, err)
        out[name + "/out"] = ref_out.numpy().astype(np.float32)
        out[name + "/checksum"] = np.array(checksum(c), dtype=np.float64)
        print(f"{name:28s} out {tuple(ref_out.shape)}  |decomposed - concatenated| = {err:.2e}")

    # combine: reference combine_lse_torch on the grid of tests/test_combine_lse.py:11-24 (seeded) + 3-way
    g = torch.Generator().manual_seed(1234)
    for n in (2, 3, 5):
        for (bs, s, h, d) in [(1, 1, 1, 63), (2, 3, 2, 64), (3, 2, 3, 128), (3, 3, 3, 129), (5, 1, 32, 128)]:
            outs = [torch.rand(bs, s, h, d, generator=g) for _ in range(n)]
            lses = [torch.rand(bs, s, h, generator=g) * 8 - 4 for _ in range(n)]
            r = ref.combine_lse_torch(outs, lses)
            key = f"combine_n{n}_{bs}x{s}x{h}x{d}"
            out[key + "/out"] = r.numpy()
            for i in range(n):
                out[key + f"/o{i}"] = outs[i].numpy()
                out[key + f"/l{i}"] = lses[i].numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hydragen_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
