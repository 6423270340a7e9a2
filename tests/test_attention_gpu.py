"""GPU parity tests (run on the B200 with ``-m gpu``): the CUDA path, called through the
reference-shaped Python surface -> ctypes -> C ABI, against (a) the committed golden vectors made
by the reference's own operator code, (b) the CPU oracle on the same seeded inputs, and (c) at
BASELINE.json's full sizes, size-independent properties (decomposed == attention over the
explicit concatenation computed by an independent kernel; two independent prefix kernels agree).

Tolerances: the reference's own for fp16 -- every |diff| <= 2e-3 and mean rdiff <= 5e-3
(tests/test_attention.py:36-38,185 of the reference); bf16 is untested upstream: 8x the fp16
atol (3 fewer mantissa bits) = 1.6e-2 and mean rdiff <= 2e-2; fp32 1e-4 / 1e-4.
"""

import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))

from oracle import hydragen_oracle as O
import make_golden_cases as MG

pytestmark = pytest.mark.gpu

TOL = {torch.float16: (2e-3, 5e-3), torch.bfloat16: (1.6e-2, 2e-2), torch.float32: (1e-4, 1e-4)}
DT = MG.DT


def _dev(x, device="cuda:0"):
    if x is None:
        return None
    if isinstance(x, list):
        return [_dev(t, device) for t in x]
    if isinstance(x, torch.Tensor):
        return x.to(device)
    return x


def _assert_close(got, ref, dtype, what=""):
    atol, rtol = TOL[dtype]
    got = got.double().cpu()
    ref = ref.double().cpu()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert torch.isfinite(got).all(), what
    ad = (got - ref).abs()
    rd = O.rdiff(got, ref).mean().item()
    assert ad.max().item() <= atol and rd <= rtol, f"{what}: max abs {ad.max().item():.3e} (atol {atol}), mean rdiff {rd:.3e} (rtol {rtol})"


@pytest.fixture(autouse=True)
def _backend_env():
    old = os.environ.get("HYDRAGEN_B200_PREFIX_BACKEND")
    yield
    if old is None:
        os.environ.pop("HYDRAGEN_B200_PREFIX_BACKEND", None)
    else:
        os.environ["HYDRAGEN_B200_PREFIX_BACKEND"] = old


@pytest.mark.parametrize("backend", ["auto", "rowwise"])
@pytest.mark.parametrize("case", MG.case_list(), ids=lambda c: c[0])
def test_operator_matches_reference_golden(case, backend, golden):
    """hydragen_attention on the GPU vs the golden produced by the reference's operator code."""
    from hydragen_b200.attention import hydragen_attention

    name, sizes, hq, hkv, d, dt, seed, nq = case
    os.environ["HYDRAGEN_B200_PREFIX_BACKEND"] = backend
    c = O.build_case(sizes, hq, hkv, d, dtype=DT[dt], seed=seed, nq=nq)
    out = hydragen_attention(**{k: _dev(v) for k, v in c.items()})
    torch.cuda.synchronize()
    assert out.dtype == DT[dt] and out.shape == c["q"].shape
    _assert_close(out, torch.from_numpy(golden[name + "/out"]), DT[dt], f"{name}/{backend}")


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("d", [64, 128])
@pytest.mark.parametrize("shape", [(1, 1, 1, 8, 8), (2, 130, 257, 8, 2), (3, 128, 128, 4, 4), (1, 300, 1000, 2, 1)])
def test_flash_attention_tcgen05(dtype, d, shape):
    """The prefix primitive (tcgen05) alone, out AND lse, incl. ragged tile edges
    (q rows % 128 != 0, keys % 128 != 0) and GQA."""
    from hydragen_b200.flash import flash_attention

    os.environ["HYDRAGEN_B200_PREFIX_BACKEND"] = "tcgen05"
    b, sq, sk, hq, hkv = shape
    g = torch.Generator().manual_seed(b * 1000 + sq + sk + d)
    q = torch.randn(b, sq, hq, d, generator=g).to(dtype)
    k = torch.randn(b, sk, hkv, d, generator=g).to(dtype)
    v = torch.randn(b, sk, hkv, d, generator=g).to(dtype)
    out, lse = flash_attention(q.cuda(), k.cuda(), v.cuda())
    torch.cuda.synchronize()
    ro, rl = O.flash_attention(q, k, v)
    assert lse.shape == (b, hq, sq) and lse.dtype == torch.float32
    _assert_close(out, ro, dtype, "out")
    assert (lse.double().cpu() - rl).abs().max().item() < 5e-3


def test_flash_attention_large_scores_rescale():
    """Forces the lazy-rescale path: the row max grows by far more than 2^8 between key blocks."""
    from hydragen_b200.flash import flash_attention

    os.environ["HYDRAGEN_B200_PREFIX_BACKEND"] = "tcgen05"
    g = torch.Generator().manual_seed(11)
    b, sq, sk, h, d = 1, 256, 1024, 2, 128
    q = torch.randn(b, sq, h, d, generator=g)
    k = torch.randn(b, sk, h, d, generator=g)
    v = torch.randn(b, sk, h, d, generator=g)
    # scale blocks of keys so that later blocks dominate (scores up to ~ +-60 after scaling)
    ramp = torch.linspace(0.2, 6.0, sk).reshape(1, sk, 1, 1)
    k = (k * ramp).to(torch.bfloat16)
    q, v = q.to(torch.bfloat16), v.to(torch.bfloat16)
    out, lse = flash_attention(q.cuda(), k.cuda(), v.cuda())
    ro, rl = O.flash_attention(q, k, v)
    _assert_close(out, ro, torch.bfloat16, "out")
    assert (lse.double().cpu() - rl).abs().max().item() < 2e-2


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
def test_flash_attention_varlen(dtype):
    from hydragen_b200.flash import flash_attention_varlen

    g = torch.Generator().manual_seed(5)
    lens = [129, 2, 300, 64]
    n, qps, hq, hkv, d = len(lens), 6, 8, 2, 128
    q = torch.randn(n * qps, hq, d, generator=g).to(dtype)
    k = torch.randn(sum(lens), hkv, d, generator=g).to(dtype)
    v = torch.randn(sum(lens), hkv, d, generator=g).to(dtype)
    cu_k = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32)
    cu_q = torch.arange(0, n + 1, dtype=torch.int32) * qps
    out, lse = flash_attention_varlen(q.cuda(), k.cuda(), v.cuda(), cu_q.cuda(), cu_k.cuda(), qps, max(lens))
    ro, rl = O.flash_attention_varlen(q, k, v, cu_q, cu_k, qps, max(lens))
    assert lse.shape == (n, hq, qps)
    _assert_close(out, ro, dtype, "out")
    assert (lse.double().cpu() - rl).abs().max().item() < 5e-3


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_flash_attention_varlen_ragged_queries_and_causal(dtype, causal):
    """The general form of the primitive (hydragen/flash.py:309-351 -> flash_attn_varlen_func): ragged query groups, an empty
    group, a group without keys, and the bottom-right aligned causal mask per sequence; lse [n, hq, max_seqlen_q], -inf past a
    sequence's rows.  Never produced by the Hydragen path, kept for surface parity (one dense call per sequence)."""
    from hydragen_b200.flash import flash_attention_varlen

    g = torch.Generator().manual_seed(17)
    qlens, klens = [1, 5, 64, 130, 0, 3], [7, 64, 64, 200, 9, 0]
    n, hq, hkv, d = len(qlens), 4, 2, 128
    q = torch.randn(sum(qlens), hq, d, generator=g).to(dtype)
    k = torch.randn(sum(klens), hkv, d, generator=g).to(dtype)
    v = torch.randn(sum(klens), hkv, d, generator=g).to(dtype)
    cu_q = torch.tensor([0] + list(torch.tensor(qlens).cumsum(0)), dtype=torch.int32)
    cu_k = torch.tensor([0] + list(torch.tensor(klens).cumsum(0)), dtype=torch.int32)
    out, lse = flash_attention_varlen(q.cuda(), k.cuda(), v.cuda(), cu_q.cuda(), cu_k.cuda(), max(qlens), max(klens), causal=causal)
    assert out.shape == q.shape and lse.shape == (n, hq, max(qlens))
    # the oracle sequence by sequence (its own varlen walks every group, including the ones without keys, through softmax)
    for i in range(n):
        qs, qe, ks, ke = int(cu_q[i]), int(cu_q[i + 1]), int(cu_k[i]), int(cu_k[i + 1])
        if qe == qs:
            continue
        if ke == ks:
            assert (out[qs:qe] == 0).all() and torch.isneginf(lse[i, :, : qe - qs]).all()
            continue
        ro, rl = O.flash_attention(q[qs:qe][None], k[ks:ke][None], v[ks:ke][None], causal=causal)
        _assert_close(out[qs:qe][None], ro, dtype, f"out[{i}]")
        assert (lse[i, :, : qe - qs].double().cpu() - rl[0]).abs().max().item() < 2e-2
        assert torch.isneginf(lse[i, :, qe - qs :]).all()


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("cfg", [(4, 1, 8, 8, 128, 40), (3, 1, 8, 1, 128, 300), (5, 2, 4, 2, 64, 17), (2, 1, 32, 32, 128, 700), (130, 1, 2, 2, 128, 9)])
def test_flash_attention_seqlen(dtype, cfg):
    """The suffix primitive: keys < seq_len[b] only; int32 and int64 lengths; GQA fold; lse [b,q,h];
    a zero-length row gives out = 0 / lse = -inf; long caches take the 4-warps-per-sequence path."""
    from hydragen_b200.flash import flash_attention_seqlen

    b, nq, hq, hkv, d, lk = cfg
    if dtype == torch.float32 and d == 256:
        pytest.skip("unsupported")
    g = torch.Generator().manual_seed(lk)
    q = torch.randn(b, nq, hq, d, generator=g).to(dtype)
    k = torch.randn(b, lk, hkv, d, generator=g).to(dtype)
    v = torch.randn(b, lk, hkv, d, generator=g).to(dtype)
    sl = torch.randint(1, lk + 1, (b,), generator=g)
    sl[0] = lk
    if b > 1:
        sl[1] = 0
    for sl_dtype in (torch.int32, torch.int64):
        out, lse = flash_attention_seqlen(q.cuda(), k.cuda(), v.cuda(), sl.to(sl_dtype).cuda())
        ro, rl = O.flash_attention_seqlen(q, k, v, sl)
        assert lse.shape == (b, nq, hq)
        _assert_close(out, ro, dtype, "out")
        fin = torch.isfinite(rl)
        assert (lse.double().cpu()[fin] - rl[fin]).abs().max().item() < 5e-3
        assert torch.equal(torch.isinf(lse.cpu()), ~fin)
        if b > 1:
            assert torch.all(out[1] == 0)


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_causal_suffix_prefill(dtype):
    """flash_attention(causal=True) with sq != sk: bottom-right aligned (hydragen/attention.py:344,
    hydragen/llama.py:537-542)."""
    from hydragen_b200.flash import flash_attention

    g = torch.Generator().manual_seed(3)
    q = torch.randn(2, 7, 8, 128, generator=g).to(dtype)
    k = torch.randn(2, 19, 4, 128, generator=g).to(dtype)
    v = torch.randn(2, 19, 4, 128, generator=g).to(dtype)
    out, lse = flash_attention(q.cuda(), k.cuda(), v.cuda(), causal=True)
    ro, rl = O.flash_attention(q, k, v, causal=True)
    _assert_close(out, ro, dtype, "out")
    assert (lse.double().cpu() - rl).abs().max().item() < 5e-3


CAUSAL_TC_CASES = [
    # b, sq, sk, hq, hkv, d   (tile = 2 x 128 query rows, key block = 64)
    (1, 128, 128, 4, 4, 128),     # one full tile A, diagonal inside the first two blocks
    (2, 300, 300, 4, 2, 128),     # two row tiles per sequence, the second partial; GQA
    (3, 70, 200, 8, 8, 64),       # sq < sk: bottom-right alignment, d = 64
    (1, 257, 1000, 2, 1, 128),    # tile B holds a single row; ragged last key block; MQA
    (2, 16, 16, 4, 4, 128),       # a chunk shorter than one key block (forced onto the tensor cores here)
    (1, 17, 17, 4, 2, 64),        # the tiny-model prefill shape of tests/test_llama_gpu.py (d = 64, GQA)
    (1, 640, 640, 2, 2, 128),     # three row tiles: tile order reversed (heavy first)
]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("case", CAUSAL_TC_CASES, ids=lambda c: "x".join(map(str, c)))
def test_causal_prefill_tcgen05(case, dtype, monkeypatch):
    """flash_attention(causal=True) on the masked instantiation of the tcgen05 kernel (hg_causal_attn_fwd) vs the
    fp64 oracle, and vs the CUDA-core kernel on the same inputs."""
    from hydragen_b200.flash import flash_attention

    b, sq, sk, hq, hkv, d = case
    g = torch.Generator().manual_seed(sq * 7 + sk)
    q = torch.randn(b, sq, hq, d, generator=g).to(dtype)
    k = torch.randn(b, sk, hkv, d, generator=g).to(dtype)
    v = torch.randn(b, sk, hkv, d, generator=g).to(dtype)
    ro, rl = O.flash_attention(q, k, v, causal=True)
    monkeypatch.setenv("HYDRAGEN_B200_CAUSAL_BACKEND", "tcgen05")
    out, lse = flash_attention(q.cuda(), k.cuda(), v.cuda(), causal=True)
    assert lse.shape == (b, hq, sq)
    _assert_close(out, ro, dtype, "tcgen05 causal out")
    assert (lse.double().cpu() - rl).abs().max().item() < 5e-3
    monkeypatch.setenv("HYDRAGEN_B200_CAUSAL_BACKEND", "rowwise")
    out2, lse2 = flash_attention(q.cuda(), k.cuda(), v.cuda(), causal=True)
    _assert_close(out2, ro, dtype, "rowwise causal out")
    assert (lse2.double().cpu() - rl).abs().max().item() < 5e-3


def test_causal_prefill_full_size_property():
    """Llama-2-7B prefill shape (2048 tokens, 32 heads): causality as a size-independent property -- changing keys
    and values after position p must not change any output row <= p -- plus row 0 == v[0] and a sampled check
    against the oracle on a few heads."""
    from hydragen_b200.flash import flash_attention

    g = torch.Generator().manual_seed(17)
    s, h, d, p = 2048, 32, 128, 1337
    q = torch.randn(1, s, h, d, generator=g).to(torch.bfloat16).cuda()
    k = torch.randn(1, s, h, d, generator=g).to(torch.bfloat16).cuda()
    v = torch.randn(1, s, h, d, generator=g).to(torch.bfloat16).cuda()
    out, lse = flash_attention(q, k, v, causal=True)
    k2, v2 = k.clone(), v.clone()
    k2[:, p + 1 :] = torch.randn(1, s - p - 1, h, d, generator=g).to(torch.bfloat16).cuda()
    v2[:, p + 1 :] = 5.0
    out2, lse2 = flash_attention(q, k2, v2, causal=True)
    assert torch.equal(out[:, : p + 1], out2[:, : p + 1]) and torch.equal(lse[..., : p + 1], lse2[..., : p + 1])
    assert not torch.equal(out[:, p + 1 :], out2[:, p + 1 :])
    assert torch.equal(out[:, 0], v[:, 0])  # the first token attends to itself only
    hs = [0, 13, 31]
    ro, rl = O.flash_attention(q[:, :, hs].cpu(), k[:, :, hs].cpu(), v[:, :, hs].cpu(), causal=True)
    _assert_close(out[:, :, hs], ro, torch.bfloat16, "full-size causal")
    assert (lse[:, hs].double().cpu() - rl).abs().max().item() < 5e-3


@pytest.mark.parametrize("split", ["1", "0"])
@pytest.mark.parametrize("ctas", [None, "5", "37"])
def test_persistent_schedule_variants(split, ctas, monkeypatch):
    """The persistent prefix kernel (used by default for hierarchies of >= 2 shared levels; forced here for single levels
    too with HYDRAGEN_B200_PREFIX_PERSISTENT=1) gives the oracle's result whatever the schedule: units cut between CTAs
    and merged through the workspace (stream-K) or dealt whole (HYDRAGEN_B200_PREFIX_SPLIT=0), on the full grid or on a
    grid of 5 / 37 CTAs (HYDRAGEN_B200_PREFIX_CTAS; the switches are read once per process -> subprocess): many pieces
    per CTA, many units per CTA, hierarchies of 2-3 levels incl. ragged ones in ONE launch, few long units cut into many
    pieces (the 3+-piece merge), a ragged level with an EMPTY group (out 0, lse -inf)."""
    import subprocess

    code = (
        "import torch, sys; sys.path.insert(0, %r)\n"
        "from oracle import hydragen_oracle as O\n"
        "from hydragen_b200.attention import hydragen_attention\n"
        "for sizes, hq, hkv, d, dt in [([[300], [17, 130], [3] * 24], 8, 4, 128, torch.bfloat16), ([[1000], [4] * 260], 4, 4, 64, torch.float16),\n"
        "                              ([[70, 200, 5], [2] * 12], 8, 1, 128, torch.bfloat16), ([[2100], [1] * 300], 2, 2, 128, torch.bfloat16)]:\n"
        "    c = O.build_case(sizes, hq, hkv, d, dtype=dt, seed=4)\n"
        "    dev = lambda x: None if x is None else ([dev(t) for t in x] if isinstance(x, list) else (x.cuda() if isinstance(x, torch.Tensor) else x))\n"
        "    for rep in range(3):\n"
        "        out = hydragen_attention(**{k: dev(v) for k, v in c.items()})\n"
        "        err = (out.double().cpu() - O.hydragen_attention(**c)).abs().max().item()\n"
        "        assert err <= (1.6e-2 if dt == torch.bfloat16 else 2e-3), (sizes, rep, err)\n"
        "from hydragen_b200.flash import flash_attention_varlen\n"
        "for dt, qps in [(torch.float16, 40), (torch.bfloat16, 300)]:\n"
        "    g = torch.Generator().manual_seed(qps)\n"
        "    lens = [700, 65, 0, 130, 1]; n, hq, hkv, d = len(lens), 4, 2, 128\n"
        "    q = torch.randn(n * qps, hq, d, generator=g).to(dt); k = torch.randn(sum(lens), hkv, d, generator=g).to(dt); v = torch.randn(sum(lens), hkv, d, generator=g).to(dt)\n"
        "    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32); cu_q = torch.arange(0, n + 1, dtype=torch.int32) * qps\n"
        "    out, lse = flash_attention_varlen(q.cuda(), k.cuda(), v.cuda(), cu_q.cuda(), cu.cuda(), qps, max(lens))\n"
        "    ro, rl = O.flash_attention_varlen(q, k, v, cu_q, cu, qps, max(lens))\n"
        "    assert (out[2 * qps:3 * qps] == 0).all() and torch.isinf(lse[2]).all()\n"
        "    assert (out.double().cpu() - ro).abs().max().item() <= (1.6e-2 if dt == torch.bfloat16 else 2e-3)\n"
        "    fin = torch.isfinite(rl); assert (lse.double().cpu()[fin] - rl[fin]).abs().max().item() < 5e-3\n"
        "print('schedule ok')\n"
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, HYDRAGEN_B200_PREFIX_SPLIT=split, HYDRAGEN_B200_PREFIX_PERSISTENT="1", HYDRAGEN_B200_PREFIX_SPLIT_OVERHEAD="0")
    if ctas is not None:
        env["HYDRAGEN_B200_PREFIX_CTAS"] = ctas
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "schedule ok" in r.stdout, r.stdout + r.stderr


def test_no_unique_keys_early_return():
    from hydragen_b200.attention import hydragen_attention_nopad

    c = O.build_case([[40], [0, 0, 0, 0]], 8, 8, 128, dtype=torch.bfloat16, seed=9)
    out = hydragen_attention_nopad(_dev(c["q"]), _dev(c["k"]), _dev(c["v"]), _dev(c["shared_ks"]), _dev(c["shared_vs"]))
    ref = O.hydragen_attention(**c)
    _assert_close(out, ref, torch.bfloat16)


def test_strided_inputs():
    """q as a slice of a fused qkv projection, K/V as slices of a larger cache (views, not copies)."""
    from hydragen_b200.attention import hydragen_attention_nopad

    g = torch.Generator().manual_seed(21)
    b, hq, hkv, d, ls, lu = 16, 8, 8, 128, 200, 24
    qkv = torch.randn(b, 1, 3 * hq * d, generator=g).to(torch.bfloat16).cuda()
    q = qkv[..., : hq * d].view(b, 1, hq, d)
    cache = torch.randn(b + 3, lu + 8, hkv, d, generator=g).to(torch.bfloat16).cuda()
    cache_v = torch.randn(b + 3, lu + 8, hkv, d, generator=g).to(torch.bfloat16).cuda()
    k, v = cache[:b], cache_v[:b]
    flat = torch.randn(2 * ls + 50, hkv, d, generator=g).to(torch.bfloat16).cuda()
    flat_v = torch.randn(2 * ls + 50, hkv, d, generator=g).to(torch.bfloat16).cuda()
    sk, sv = flat[: 2 * ls].view(2, ls, hkv, d), flat_v[: 2 * ls].view(2, ls, hkv, d)
    sl = torch.randint(1, lu + 1, (b,), generator=g).cuda()
    out = hydragen_attention_nopad(q, k, v, [sk], [sv], seq_len=sl)
    ref = O.hydragen_attention_nopad(q.cpu(), k.cpu(), v.cpu(), [sk.cpu()], [sv.cpu()], seq_len=sl.cpu())
    _assert_close(out, ref, torch.bfloat16)


def test_cuda_graph_capture_and_replay():
    """The operator is capturable (no sync / host reads on the launch path) and replays correctly after
    the inputs change in place -- the way hydragen/llama.py:781-866 uses it."""
    from hydragen_b200.attention import hydragen_attention_nopad

    g = torch.Generator().manual_seed(2)
    b, h, d, ls, lu = 64, 8, 128, 384, 16
    mk = lambda *s: torch.randn(*s, generator=g).to(torch.bfloat16).cuda()
    q, k, v, sk, sv = mk(b, 1, h, d), mk(b, lu, h, d), mk(b, lu, h, d), mk(1, ls, h, d), mk(1, ls, h, d)
    sl = torch.full((b,), 3, dtype=torch.int64).cuda()
    hydragen_attention_nopad(q, k, v, [sk], [sv], seq_len=sl)  # warm-up (lazy init outside capture)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            hydragen_attention_nopad(q, k, v, [sk], [sv], seq_len=sl)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = hydragen_attention_nopad(q, k, v, [sk], [sv], seq_len=sl)
    for step in range(3):
        q.copy_(mk(b, 1, h, d))
        sl.fill_(5 + 4 * step)
        graph.replay()
        torch.cuda.synchronize()
        ref = O.hydragen_attention_nopad(q.cpu(), k.cpu(), v.cpu(), [sk.cpu()], [sv.cpu()], seq_len=sl.cpu())
        _assert_close(out, ref, torch.bfloat16, f"replay {step}")


def test_full_size_microbenchmark_config_properties():
    """BASELINE.json configs[1] at full size (B=1024, prefix 2048, 32 heads, d=128, bf16), beyond what
    the CPU oracle finishes quickly.  Size-independent checks:
      1. the tcgen05 prefix kernel and the CUDA-core row-wise kernel (independent code) agree on
         out and lse for every (sequence, head);
      2. decomposed attention == attention over the explicit concatenation [prefix ; suffix]
         computed by the row-wise kernel in one pass (the reference test's criterion,
         tests/test_attention.py:132-187), on a batch subset that fits memory;
      3. the CPU oracle (fp32 on the host cores) on ALL 1024 sequences and heads."""
    from hydragen_b200.attention import hydragen_attention_nopad
    from hydragen_b200.flash import flash_attention_seqlen, prefix_attention_grouped

    torch.manual_seed(0)
    B, Ls, Lu, H, D = 1024, 2048, 32, 32, 128
    dev = "cuda:0"
    q = torch.randn(B, 1, H, D, device=dev, dtype=torch.bfloat16)
    k = torch.randn(B, Lu, H, D, device=dev, dtype=torch.bfloat16)
    v = torch.randn(B, Lu, H, D, device=dev, dtype=torch.bfloat16)
    sk = torch.randn(1, Ls, H, D, device=dev, dtype=torch.bfloat16)
    sv = torch.randn(1, Ls, H, D, device=dev, dtype=torch.bfloat16)
    sl = torch.randint(1, Lu + 1, (B,), device=dev)

    os.environ["HYDRAGEN_B200_PREFIX_BACKEND"] = "tcgen05"
    o_tc, l_tc = prefix_attention_grouped(q, sk, sv, n_groups=1)
    os.environ["HYDRAGEN_B200_PREFIX_BACKEND"] = "rowwise"
    o_rw, l_rw = prefix_attention_grouped(q, sk, sv, n_groups=1)
    os.environ["HYDRAGEN_B200_PREFIX_BACKEND"] = "auto"
    torch.cuda.synchronize()
    assert (l_tc - l_rw).abs().max().item() < 5e-3
    assert (o_tc.float() - o_rw.float()).abs().max().item() < 4e-3  # |out| <~ 0.1 here; 1 bf16 ulp

    out = hydragen_attention_nopad(q, k, v, [sk], [sv], seq_len=sl)
    nb = 128
    kc = torch.cat([sk.expand(nb, -1, -1, -1), k[:nb]], dim=1)
    vc = torch.cat([sv.expand(nb, -1, -1, -1), v[:nb]], dim=1)
    o_cat, _ = flash_attention_seqlen(q[:nb], kc, vc, seq_len=sl[:nb] + Ls)
    torch.cuda.synchronize()
    _assert_close(out[:nb], o_cat, torch.bfloat16, "decomposed vs concatenated")

    ref = O.hydragen_attention_nopad(q.cpu(), k.cpu(), v.cpu(), [sk.cpu()], [sv.cpu()], seq_len=sl.cpu(), compute_dtype=torch.float32)
    _assert_close(out, ref, torch.bfloat16, "cfg#2 full size vs oracle")


def test_full_size_decode_config_ragged_suffix():
    """BASELINE.json configs[2]'s decode shape at full size: B = 1024 sequences at DIFFERENT points of their completion
    (suffix 1 .. 127 rows in a 128-row unique cache), prefix 2048, 32 heads, d = 128, bf16 -- one fused decode step
    (KV append at positions[b] + suffix + combine behind the persistent prefix launch) against the CPU oracle on every
    sequence, and the caches afterwards against scatter_."""
    from hydragen_b200.attention import hydragen_attention_decode

    g = torch.Generator().manual_seed(5)
    B, Ls, Lu, H, D = 1024, 2048, 128, 32, 128
    dt = torch.bfloat16
    mk = lambda *s: torch.randn(*s, generator=g).to(dt)
    q, kn, vn = mk(B, 1, H, D), mk(B, 1, H, D), mk(B, 1, H, D)
    kc, vc = mk(B, Lu, H, D), mk(B, Lu, H, D)
    sk, sv = mk(1, Ls, H, D), mk(1, Ls, H, D)
    pos = torch.randint(0, Lu - 1, (B,), generator=g)
    pos[0], pos[1] = 0, Lu - 2  # both ends: no older rows at all / 127 rows
    kcd, vcd = kc.cuda(), vc.cuda()
    out = hydragen_attention_decode(q.cuda(), kn.cuda(), vn.cuda(), pos.cuda(), kcd, vcd, [sk.cuda()], [sv.cuda()])
    torch.cuda.synchronize()
    idx = pos.view(B, 1, 1, 1).expand(B, 1, H, D)
    kc.scatter_(1, idx, kn)
    vc.scatter_(1, idx, vn)
    assert torch.equal(kcd.cpu(), kc) and torch.equal(vcd.cpu(), vc)
    ref = O.hydragen_attention_nopad(q, kc, vc, [sk], [sv], seq_len=pos + 1, compute_dtype=torch.float32)
    _assert_close(out, ref, dt, "cfg#3 decode step, ragged suffix, full size")


def test_full_size_two_level_hierarchy():
    """BASELINE.json configs[3] at full size: 1 prefix of 1024 tokens -> 32 second-level prompts of 64 tokens -> 32
    completions each (B = 1024), 32 heads, d = 128, bf16.  Both shared levels run in ONE persistent prefix launch
    (level 2: 32 groups x 32 rows x 64 keys), merged 3-way with the suffix by the fused launch; checked against the CPU
    oracle on every sequence and against the level-by-level result (one launch per level)."""
    from hydragen_b200 import _lib
    from hydragen_b200.attention import hydragen_attention_nopad
    from hydragen_b200.flash import prefix_attention_grouped, prefix_attention_levels

    g = torch.Generator().manual_seed(6)
    B, H, D, Lu = 1024, 32, 128, 16
    dt = torch.bfloat16
    mk = lambda *s: torch.randn(*s, generator=g).to(dt)
    q, k, v = mk(B, 1, H, D), mk(B, Lu, H, D), mk(B, Lu, H, D)
    s1k, s1v, s2k, s2v = mk(1, 1024, H, D), mk(1, 1024, H, D), mk(32, 64, H, D), mk(32, 64, H, D)
    sl = torch.randint(1, Lu + 1, (B,), generator=g)
    n_ctas, pieces = _lib.prefix_schedule([(1, 1024, 0), (32, 64, 0)], B, H, n_sms=_lib.load().hg_sm_count() or 148)
    assert {p[2] for p in pieces} == {0, 1}  # one launch holds units of both levels
    qd, dev = q.cuda(), (lambda *ts: [t.cuda() for t in ts])
    out = hydragen_attention_nopad(qd, k.cuda(), v.cuda(), dev(s1k, s2k), dev(s1v, s2v), seq_len=sl.cuda())
    outs, lses = prefix_attention_levels(qd, dev(s1k, s2k), dev(s1v, s2v), [1, 32], [None, None], [None, None])
    o1, l1 = prefix_attention_grouped(qd, s1k.cuda(), s1v.cuda(), n_groups=1)
    o2, l2 = prefix_attention_grouped(qd, s2k.cuda(), s2v.cuda(), n_groups=32)
    torch.cuda.synchronize()
    # the grouped launch schedules differently from the single-level ones: same maths, so (nearly) the same bits
    assert (outs[0].float() - o1.float()).abs().max().item() < 4e-3 and (lses[0] - l1).abs().max().item() < 5e-3
    assert (outs[1].float() - o2.float()).abs().max().item() < 2e-2 and (lses[1] - l2).abs().max().item() < 5e-3
    ref = O.hydragen_attention_nopad(q, k, v, [s1k, s2k], [s1v, s2v], seq_len=sl, compute_dtype=torch.float32)
    _assert_close(out, ref, dt, "cfg#4 two-level hierarchy, full size")


def test_kv_append():
    from hydragen_b200 import _lib

    g = torch.Generator().manual_seed(4)
    b, nq, hkv, d, lk = 37, 1, 8, 128, 48
    kc = torch.zeros(b + 2, lk, hkv, d, dtype=torch.bfloat16).cuda()
    vc = torch.zeros_like(kc)
    kn = torch.randn(b, nq, hkv, d, generator=g).to(torch.bfloat16).cuda()
    vn = torch.randn(b, nq, hkv, d, generator=g).to(torch.bfloat16).cuda()
    pos = torch.randint(0, lk, (b, nq), generator=g).cuda()
    _lib.kv_append(kn, vn, pos, kc, vc)
    # the reference's formulation: scatter_ with a fully expanded index (hydragen/llama.py:250-257)
    rk = torch.zeros_like(kc)
    rv = torch.zeros_like(vc)
    idx = pos.view(b, nq, 1, 1).expand(b, nq, hkv, d)
    rk[:b].scatter_(1, idx, kn)
    rv[:b].scatter_(1, idx, vn)
    assert torch.equal(kc, rk) and torch.equal(vc, rv)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("cfg", [(37, 8, 8, 128, 48), (64, 8, 1, 128, 16), (5, 4, 2, 64, 33), (130, 32, 32, 128, 16), (3, 16, 4, 128, 200)])
def test_decode_attention_fused(dtype, cfg):
    """KV append + suffix attention in one launch (no prefix partials): cache rows written exactly where
    the reference's scatter_ puts them (hydragen/llama.py:250-257), output == flash_attention_seqlen with
    seq_len = position + 1 (llama.py:569) on the updated cache."""
    from hydragen_b200.flash import decode_attention_fused

    b, hq, hkv, d, lk = cfg
    g = torch.Generator().manual_seed(b * 7 + hq + lk)
    q = torch.randn(b, 1, hq, d, generator=g).to(dtype)
    kn = torch.randn(b, 1, hkv, d, generator=g).to(dtype)
    vn = torch.randn(b, 1, hkv, d, generator=g).to(dtype)
    kc = torch.randn(b + 1, lk, hkv, d, generator=g).to(dtype)
    vc = torch.randn(b + 1, lk, hkv, d, generator=g).to(dtype)
    pos = torch.randint(0, lk, (b, 1), generator=g)
    pos[0, 0], pos[-1, 0] = 0, lk - 1  # first and last row of the cache
    kcd, vcd = kc.cuda(), vc.cuda()
    out, lse = decode_attention_fused(q.cuda(), kn.cuda(), vn.cuda(), pos.cuda(), kcd[:b], vcd[:b])
    torch.cuda.synchronize()
    rk, rv = kc.clone(), vc.clone()
    idx = pos.view(b, 1, 1, 1).expand(b, 1, hkv, d)
    rk[:b].scatter_(1, idx, kn)
    rv[:b].scatter_(1, idx, vn)
    assert torch.equal(kcd.cpu(), rk) and torch.equal(vcd.cpu(), rv)  # append is a bit-exact copy; nothing else touched
    ro, rl = O.flash_attention_seqlen(q, rk[:b], rv[:b], seq_len=pos[:, 0] + 1)
    _assert_close(out, ro, dtype, "out")
    assert (lse.double().cpu() - rl).abs().max().item() < 5e-3


@pytest.mark.parametrize("sizes", [[[1], [96]], [[2], [5, 5]], [[3, 3], [7, 9, 11, 4, 6, 2]]])
def test_hydragen_attention_decode_matches_unfused(sizes):
    """hydragen_attention_decode (prefix + fused append/suffix/combine) == kv_append followed by
    hydragen_attention(seq_lens = pos + 1), and both match the oracle on the updated cache."""
    from hydragen_b200 import _lib
    from hydragen_b200.attention import hydragen_attention, hydragen_attention_decode

    hq, hkv, d, dtype = 8, 2, 128, torch.bfloat16
    c = O.build_case(sizes, hq, hkv, d, dtype=dtype, seed=3)
    b = c["q"].shape[0]
    g = torch.Generator().manual_seed(17)
    lk = c["k"].shape[1] + 3
    kc = torch.randn(b, lk, hkv, d, generator=g).to(dtype)
    vc = torch.randn(b, lk, hkv, d, generator=g).to(dtype)
    kn = torch.randn(b, 1, hkv, d, generator=g).to(dtype)
    vn = torch.randn(b, 1, hkv, d, generator=g).to(dtype)
    pos = torch.randint(0, lk, (b, 1), generator=g)
    shared = {k: _dev(c[k]) for k in ("shared_ks", "shared_vs", "shared_cu_seq_lens", "shared_max_seq_lens", "use_varlens")}
    k1, v1 = kc.cuda(), vc.cuda()
    out1 = hydragen_attention_decode(c["q"].cuda(), kn.cuda(), vn.cuda(), pos.cuda(), k1, v1, **shared)
    k2, v2 = kc.cuda(), vc.cuda()
    _lib.kv_append(kn.cuda(), vn.cuda(), pos.cuda(), k2, v2)
    out2 = hydragen_attention(c["q"].cuda(), k2, v2, seq_lens=(pos[:, 0] + 1).cuda(), **shared)
    torch.cuda.synchronize()
    assert torch.equal(k1, k2) and torch.equal(v1, v2)
    _assert_close(out1, out2, dtype, "fused vs unfused")
    ref = O.hydragen_attention(c["q"], k2.cpu(), v2.cpu(), c["shared_ks"], c["shared_vs"], c["shared_cu_seq_lens"], c["shared_max_seq_lens"],
                               c["use_varlens"], seq_lens=pos[:, 0] + 1)
    _assert_close(out1, ref, dtype, "fused vs oracle")


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("splits", [2, 3, 8])
def test_prefix_split_kv_partials_merge_to_full(dtype, splits):
    """Split-KV prefix launch (hg_prefix_attn_split_fwd): the merged partials == the unsplit result == oracle,
    for uniform and ragged (varlen) groups, incl. splits that receive no keys (out 0, lse -inf)."""
    from hydragen_b200 import _lib
    from hydragen_b200.attention import combine_lse_cuda

    g = torch.Generator().manual_seed(splits)
    lens = [700, 65, 130, 1]  # ragged groups: the short ones leave later splits empty
    n, qps, hq, hkv, d = len(lens), 40, 4, 2, 128
    q = torch.randn(n * qps, 1, hq, d, generator=g).to(dtype)
    k = torch.randn(sum(lens), hkv, d, generator=g).to(dtype)
    v = torch.randn(sum(lens), hkv, d, generator=g).to(dtype)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32)
    qd, kd, vd, cud = q.cuda(), k.cuda(), v.cuda(), cu.cuda()
    out = torch.empty(splits, n * qps, 1, hq, d, device="cuda", dtype=dtype)
    lse = torch.empty(splits, n * qps, 1, hq, device="cuda", dtype=torch.float32)
    _lib.prefix_attn_fwd(qd, kd, vd, out, lse, n, qps, kd.shape[0], 0, cud, max(lens), hq, hkv, d, hq * d, hkv * d, d**-0.5, kv_splits=splits)
    merged, mlse = combine_lse_cuda([out[i] for i in range(splits)], [lse[i] for i in range(splits)], return_lse=True)
    torch.cuda.synchronize()
    assert torch.isinf(lse).any() or splits == 2  # some (group, split) pairs are empty
    cu_q = torch.arange(0, n + 1, dtype=torch.int32) * qps
    ro, rl = O.flash_attention_varlen(q.view(n * qps, hq, d), k, v, cu_q, cu, qps, max(lens))
    _assert_close(merged.view(n * qps, hq, d), ro, dtype, f"split-KV x{splits}")
    rl = rl.permute(0, 2, 1).reshape(n * qps, hq)  # [n, h, qps] -> rows
    assert (mlse.view(n * qps, hq).double().cpu() - rl).abs().max().item() < 5e-3


def test_operator_uses_split_kv_when_few_heads():
    """A head-parallel rank's shape (few local heads, long prefix): the operator picks kv_splits > 1 by itself and
    the result still matches the oracle."""
    from hydragen_b200 import _lib
    from hydragen_b200.attention import hydragen_attention_nopad

    g = torch.Generator().manual_seed(9)
    b, hq, hkv, d, ls, lu = 256, 2, 1, 128, 1500, 8
    mk = lambda *s: torch.randn(*s, generator=g).to(torch.bfloat16)
    q, k, v, sk, sv = mk(b, 1, hq, d), mk(b, lu, hkv, d), mk(b, lu, hkv, d), mk(1, ls, hkv, d), mk(1, ls, hkv, d)
    sl = torch.randint(1, lu + 1, (b,), generator=g)
    assert _lib.prefix_suggest_splits(torch.device("cuda:0"), 1, b, hq, ls, 8) > 1
    out = hydragen_attention_nopad(q.cuda(), k.cuda(), v.cuda(), [sk.cuda()], [sv.cuda()], seq_len=sl.cuda())
    ref = O.hydragen_attention_nopad(q, k, v, [sk], [sv], seq_len=sl)
    _assert_close(out, ref, torch.bfloat16, "auto split-KV")


def test_host_decode_pipeline_matches_direct_calls():
    """The host-buffer entry (pinned host q/k/v -> H2D -> kernels -> D2H on three streams) gives, layer by layer and
    step after step, exactly what the device-resident call gives."""
    from hydragen_b200.attention import hydragen_attention_decode
    from hydragen_b200.host import HostDecodeLayer, HostDecodePipeline

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(21)
    L, b, hq, hkv, d, ls, lk = 3, 96, 8, 4, 128, 200, 16
    mk = lambda *s: torch.randn(*s, generator=g).to(torch.bfloat16)
    layers, ref_state = [], []
    for _ in range(L):
        sk, sv = mk(1, ls, hkv, d).cuda(), mk(1, ls, hkv, d).cuda()
        kc, vc = mk(b, lk, hkv, d), mk(b, lk, hkv, d)
        layers.append(HostDecodeLayer(mk(b, 1, hq, d).pin_memory(), mk(b, 1, hkv, d).pin_memory(), mk(b, 1, hkv, d).pin_memory(),
                                      torch.empty(b, 1, hq, d, dtype=torch.bfloat16).pin_memory(),
                                      torch.empty(b, 1, hq, d, dtype=torch.bfloat16, device=dev), torch.empty(b, 1, hkv, d, dtype=torch.bfloat16, device=dev),
                                      torch.empty(b, 1, hkv, d, dtype=torch.bfloat16, device=dev), kc.cuda(), vc.cuda(), [sk], [sv]))
        ref_state.append((kc.cuda(), vc.cuda(), sk, sv))
    pipe = HostDecodePipeline(dev)
    for step in range(3):
        pos = torch.full((b,), step, dtype=torch.int64, device=dev)
        for ly in layers:  # new host inputs every step
            ly.q_host.copy_(mk(b, 1, hq, d))
            ly.k_host.copy_(mk(b, 1, hkv, d))
            ly.v_host.copy_(mk(b, 1, hkv, d))
        pipe.step(layers, pos)
        pipe.synchronize()
        for ly, (kc, vc, sk, sv) in zip(layers, ref_state):
            ref = hydragen_attention_decode(ly.q_host.cuda(), ly.k_host.cuda(), ly.v_host.cuda(), pos, kc, vc, [sk], [sv])
            torch.cuda.synchronize()
            assert torch.equal(ly.out_host, ref.cpu())
            assert torch.equal(ly.k_cache, kc) and torch.equal(ly.v_cache, vc)
    # the same step captured once as a CUDA graph (copies + kernels of every layer on the three streams) and replayed
    pos = torch.full((b,), 3, dtype=torch.int64, device=dev)  # static device tensor: updated in place between replays
    graph = pipe.capture(layers, pos)
    for step in range(3, 6):
        pos.fill_(step)
        for ly in layers:
            ly.q_host.copy_(mk(b, 1, hq, d))
            ly.k_host.copy_(mk(b, 1, hkv, d))
            ly.v_host.copy_(mk(b, 1, hkv, d))
        graph.replay()
        torch.cuda.synchronize()
        for ly, (kc, vc, sk, sv) in zip(layers, ref_state):
            ref = hydragen_attention_decode(ly.q_host.cuda(), ly.k_host.cuda(), ly.v_host.cuda(), pos, kc, vc, [sk], [sv])
            torch.cuda.synchronize()
            assert torch.equal(ly.out_host, ref.cpu()), f"graphed step {step}"
            assert torch.equal(ly.k_cache, kc) and torch.equal(ly.v_cache, vc)
