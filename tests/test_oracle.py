"""CPU tests: the oracle against the committed golden vectors (which were produced by the
REFERENCE's own hydragen_attention / combine_lse_torch code, see tests/golden/make_golden.py),
and the oracle's internal consistency (decomposed == concatenated, tests/test_attention.py:132-187
of the reference)."""

import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))

from oracle import hydragen_oracle as O
import make_golden_cases as MG

DT = {"float16": torch.float16, "bfloat16": torch.bfloat16, "float32": torch.float32}


@pytest.mark.parametrize("case", MG.case_list(), ids=lambda c: c[0])
def test_oracle_matches_reference_golden(case, golden):
    name, sizes, hq, hkv, d, dt, seed, nq = case
    c = O.build_case(sizes, hq, hkv, d, dtype=DT[dt], seed=seed, nq=nq)
    # generator drift guard: the regenerated inputs are the ones the golden was made from
    assert abs(MG.checksum(c) - float(golden[name + "/checksum"])) < 1e-6 * max(1.0, abs(float(golden[name + "/checksum"])))
    out = O.hydragen_attention(**c)
    ref = torch.from_numpy(golden[name + "/out"]).double()
    assert out.shape == ref.shape
    assert (out - ref).abs().max().item() < 2e-6  # golden is stored as float32
    # and the reference test's own criterion (decomposed == attention over the concatenation)
    truth = O.concat_attention(c["q"], c["k"], c["v"], c["shared_ks"], c["shared_vs"], c["shared_cu_seq_lens"], c["use_varlens"], c["seq_lens"])
    assert (out - truth).abs().max().item() < 1e-9


def test_combine_matches_reference_golden(golden):
    keys = sorted({k.split("/")[0] for k in golden.files if k.startswith("combine_")})
    assert len(keys) == 15
    for key in keys:
        n = int(key.split("_")[1][1:])
        outs = [torch.from_numpy(golden[f"{key}/o{i}"]) for i in range(n)]
        lses = [torch.from_numpy(golden[f"{key}/l{i}"]) for i in range(n)]
        got = O.combine_lse_torch(outs, lses)
        ref = torch.from_numpy(golden[key + "/out"])
        assert torch.allclose(got, ref, rtol=1e-6, atol=1e-7), key


def test_oracle_primitive_layouts():
    g = torch.Generator().manual_seed(7)
    q = torch.randn(3, 2, 8, 64, generator=g)
    k = torch.randn(3, 5, 2, 64, generator=g)
    v = torch.randn(3, 5, 2, 64, generator=g)
    out, lse = O.flash_attention(q, k, v)
    assert out.shape == (3, 2, 8, 64) and lse.shape == (3, 8, 2)
    # against torch's own SDPA (independent implementation of the same definition), GQA by repeat
    kk = k.repeat_interleave(4, dim=2)
    vv = v.repeat_interleave(4, dim=2)
    sd = torch.nn.functional.scaled_dot_product_attention(q.permute(0, 2, 1, 3).double(), kk.permute(0, 2, 1, 3).double(), vv.permute(0, 2, 1, 3).double())
    assert (out - sd.permute(0, 2, 1, 3)).abs().max() < 1e-12
    # lse definition
    s = torch.einsum("bqhd,bkhd->bhqk", q.double(), kk.double()) * 64**-0.5
    assert (lse - torch.logsumexp(s, -1)).abs().max() < 1e-12
    # seqlen masking + [b, q, h] layout
    sl = torch.tensor([5, 1, 3])
    o2, l2 = O.flash_attention_seqlen(q, k, v, seq_len=sl)
    assert l2.shape == (3, 2, 8)
    o1, _ = O.flash_attention(q[1:2], k[1:2, :1], v[1:2, :1])
    assert (o2[1:2] - o1).abs().max() < 1e-12
    # causal, bottom-right aligned: last query sees everything, first sees sk - sq + 1 keys
    oc, _ = O.flash_attention(q, k, v, causal=True)
    of, _ = O.flash_attention(q[:, 1:], k, v)
    assert (oc[:, 1:] - of).abs().max() < 1e-12
    o4, _ = O.flash_attention(q[:, :1], k[:, :4], v[:, :4])
    assert (oc[:, :1] - o4).abs().max() < 1e-12


def test_oracle_empty_rows():
    q = torch.randn(2, 1, 2, 64)
    k = torch.randn(2, 4, 2, 64)
    v = torch.randn(2, 4, 2, 64)
    out, lse = O.flash_attention_seqlen(q, k, v, seq_len=torch.tensor([0, 2]))
    assert torch.all(out[0] == 0) and torch.all(torch.isinf(lse[0])) and torch.all(lse[0] < 0)
    merged = O.combine_lse_torch([out, out], [lse, lse])
    assert torch.isfinite(merged).all()


def test_early_return_without_unique_keys():
    c = O.build_case([[12], [0, 0, 0]], 4, 2, 64, dtype=torch.float32, seed=3)
    assert c["k"].shape[1] == 0
    out = O.hydragen_attention(**c)
    so, _ = O.flash_attention(c["q"].reshape(1, 3, 4, 64), c["shared_ks"][0], c["shared_vs"][0])
    assert (out - so.reshape(3, 1, 4, 64)).abs().max() < 1e-12


def test_rope_oracle_matches_transformers_golden():
    """oracle.apply_rotary_pos_emb / rotary_tables vs the outputs of transformers' own apply_rotary_pos_emb
    (tests/golden/make_golden_rope.py) -- bit-exact in every dtype."""
    import make_golden_rope as MR

    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "rope_golden.npz"))
    for name, b, s, hq, hkv, d, dtype, max_pos, seed in MR.CASES:
        q, k, pos = MR.make_inputs(b, s, hq, hkv, d, dtype, max_pos, seed)
        assert abs(MR.checksum(q, k, pos) - float(gold[name + "/checksum"])) < 1e-6
        cos, sin = O.rotary_tables(d, max_pos, 10000.0, DT[dtype])
        qe, ke = O.apply_rotary_pos_emb(q, k, cos, sin, pos, unsqueeze_dim=2)
        for got, key in ((qe, "/q"), (ke, "/k")):
            ref = torch.from_numpy(gold[name + key])
            got = got.contiguous().view(torch.int16) if DT[dtype] != torch.float32 else got
            assert torch.equal(got, ref), f"{name}{key}"
