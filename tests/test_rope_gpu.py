"""GPU parity tests of the RoPE prologue kernel (hg_rope_qk) -- BIT-EXACT:
against the golden vectors produced by transformers' own apply_rotary_pos_emb
(tests/golden/make_golden_rope.py), against the CPU oracle on seeded inputs, in place, on strided views of
a fused qkv buffer, with int32 / int64 positions, and at the full cfg#2 size through size-independent
properties (norm preservation of every rotated pair, position 0 = identity, q and k agree on shared rows)."""

import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))

from oracle import hydragen_oracle as O
import make_golden_rope as MR

pytestmark = pytest.mark.gpu

DT = {"float16": torch.float16, "bfloat16": torch.bfloat16, "float32": torch.float32}


def _bits(t):
    t = t.contiguous()
    return t.view(torch.int16) if t.dtype != torch.float32 else t.view(torch.int32)


def _same_bits(a, b, what=""):
    a, b = _bits(a.cpu()), _bits(b.cpu())
    n = (a != b).sum().item()
    assert n == 0, f"{what}: {n} of {a.numel()} elements differ"


@pytest.mark.parametrize("case", MR.CASES, ids=lambda c: c[0])
def test_matches_transformers_golden(case):
    from hydragen_b200.rope import apply_rotary_pos_emb

    name, b, s, hq, hkv, d, dtype, max_pos, seed = case
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "rope_golden.npz"))
    q, k, pos = MR.make_inputs(b, s, hq, hkv, d, dtype, max_pos, seed)
    assert abs(MR.checksum(q, k, pos) - float(gold[name + "/checksum"])) < 1e-6
    cos, sin = O.rotary_tables(d, max_pos, 10000.0, DT[dtype])
    qe, ke = apply_rotary_pos_emb(q.cuda(), k.cuda(), cos.cuda(), sin.cuda(), pos.cuda(), unsqueeze_dim=2)
    for got, key in ((qe, "/q"), (ke, "/k")):
        ref = torch.from_numpy(gold[name + key])
        got = got.cpu().contiguous()
        got = got.view(torch.int16) if DT[dtype] != torch.float32 else got
        assert torch.equal(got, ref), f"{name}{key}"


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("pos_dtype", [torch.int64, torch.int32])
def test_oracle_inplace_and_fused_qkv_views(dtype, pos_dtype):
    """q / k as strided views of ONE fused projection output [b, s, (hq + 2 hkv) d], rotated in place; v untouched."""
    from hydragen_b200.rope import apply_rotary_pos_emb

    g = torch.Generator().manual_seed(11)
    b, s, hq, hkv, d, max_pos = 7, 3, 8, 2, 128, 777
    qkv = torch.randn(b, s, (hq + 2 * hkv) * d, generator=g).to(dtype)
    pos = torch.randint(0, max_pos, (b, s), generator=g)
    cos, sin = O.rotary_tables(d, max_pos, 10000.0, dtype)
    qs, ks = slice(0, hq * d), slice(hq * d, (hq + hkv) * d)
    ref_q, ref_k = O.apply_rotary_pos_emb(qkv[..., qs].reshape(b, s, hq, d), qkv[..., ks].reshape(b, s, hkv, d), cos, sin, pos)
    dev = qkv.cuda()
    before_v = dev[..., (hq + hkv) * d :].clone()
    qv = dev[..., qs].unflatten(-1, (hq, d))
    kv = dev[..., ks].unflatten(-1, (hkv, d))
    oq, ok = apply_rotary_pos_emb(qv, kv, cos.cuda(), sin.cuda(), pos.cuda().to(pos_dtype), inplace=True)
    assert oq.data_ptr() == qv.data_ptr() and ok.data_ptr() == kv.data_ptr()
    _same_bits(qv, ref_q, "q")
    _same_bits(kv, ref_k, "k")
    assert torch.equal(dev[..., (hq + hkv) * d :], before_v)  # v part of the fused buffer is not touched


def test_out_of_place_leaves_inputs_and_handles_edge_shapes():
    from hydragen_b200.rope import apply_rotary_pos_emb

    g = torch.Generator().manual_seed(3)
    for b, s, hq, hkv, d in [(1, 1, 1, 1, 16), (2, 1, 40, 40, 128), (3, 130, 2, 1, 64), (0, 1, 4, 4, 128), (5, 1, 300, 4, 16)]:
        q = torch.randn(b, s, hq, d, generator=g).to(torch.bfloat16)
        k = torch.randn(b, s, hkv, d, generator=g).to(torch.bfloat16)
        pos = torch.randint(0, 64, (b, s), generator=g)
        cos, sin = O.rotary_tables(d, 64, 500000.0, torch.bfloat16)
        qd, kd = q.cuda(), k.cuda()
        oq, ok = apply_rotary_pos_emb(qd, kd, cos.cuda(), sin.cuda(), pos.cuda())
        assert torch.equal(qd.cpu(), q) and torch.equal(kd.cpu(), k)
        rq, rk = O.apply_rotary_pos_emb(q, k, cos, sin, pos)
        _same_bits(oq, rq, f"q {b, s, hq, hkv, d}")
        _same_bits(ok, rk, f"k {b, s, hq, hkv, d}")


def test_full_size_properties():
    """cfg#2 size (B=1024, 32+32 heads, d=128): position 0 is the identity, rows at the same position and with
    the same content rotate identically, and every (i, i + d/2) pair keeps its norm to bf16 accuracy."""
    from hydragen_b200.rope import apply_rotary_pos_emb

    g = torch.Generator().manual_seed(5)
    b, h, d, max_pos = 1024, 32, 128, 4096
    q = torch.randn(b, 1, h, d, generator=g).to(torch.bfloat16).cuda()
    k = q.clone()
    pos = torch.randint(1, max_pos, (b, 1), generator=g).cuda()
    pos[:17] = 0
    cos, sin = (t.cuda() for t in O.rotary_tables(d, max_pos, 10000.0, torch.bfloat16))
    oq, ok = apply_rotary_pos_emb(q, k, cos, sin, pos)
    assert torch.equal(oq, ok)  # the q and the k code path are the same arithmetic
    assert torch.equal(oq[:17], q[:17])  # cos = 1, sin = 0
    n_in = q.float()[..., : d // 2] ** 2 + q.float()[..., d // 2 :] ** 2
    n_out = oq.float()[..., : d // 2] ** 2 + oq.float()[..., d // 2 :] ** 2
    assert ((n_in - n_out).abs() <= 0.04 * n_in + 1e-3).all()
    # and the whole thing against the eager statement on the device (same ops, same dtype): bit-exact
    rq, _ = O.apply_rotary_pos_emb(q, k, cos, sin, pos)
    assert torch.equal(oq, rq)


def test_argument_errors():
    from hydragen_b200.rope import apply_rotary_pos_emb

    q = torch.zeros(2, 1, 4, 128, dtype=torch.bfloat16, device="cuda")
    k = torch.zeros(2, 1, 2, 128, dtype=torch.bfloat16, device="cuda")
    cos = torch.zeros(16, 128, dtype=torch.bfloat16, device="cuda")
    pos = torch.zeros(2, 1, dtype=torch.long, device="cuda")
    with pytest.raises(ValueError):
        apply_rotary_pos_emb(q, k, cos.float(), cos.float(), pos)
    with pytest.raises(ValueError):
        apply_rotary_pos_emb(q, k, cos[:, :64], cos[:, :64], pos)
    with pytest.raises(ValueError):
        apply_rotary_pos_emb(q, k, cos, cos, pos[:1])
    with pytest.raises(ValueError):
        apply_rotary_pos_emb(q, k, cos, cos, pos.float())
