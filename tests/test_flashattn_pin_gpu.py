"""Pins the ONE piece of the oracle that the reference's own tests never pin: the isolated softmax-attention
primitive ``(out, lse)``.  Its arithmetic lives in flash-attn (pinned v2.3.6, requirements.txt:7 of the reference;
call sites hydragen/flash.py:295-304, 336-349), which is not in the reference tree -- but a flash-attn build IS
installed on the GPU box.  Here its public API (``flash_attn_func(..., return_attn_probs=True)`` -> out, lse;
``flash_attn_varlen_func``) is run on the shapes the reference's test makes it see (tests/test_attention.py:26-32:
every shared level and the unique level of the five ``sizes_list`` cases, kvheads 1 and 8, fp16, d 128) plus the
causal form, and BOTH the CPU oracle (fp64) and this repo's kernels are compared with it at the reference's
tolerances (test_attention.py:36-38: atol 2e-3, mean rdiff 5e-3; LSE 5e-3).  Skipped when flash-attn cannot be imported."""

import pytest
import torch

from oracle import hydragen_oracle as O

pytestmark = pytest.mark.gpu

fa = pytest.importorskip("flash_attn")
ATOL, RTOL, LSE_TOL = 2e-3, 5e-3, 5e-3


def _close(got, ref, what):
    got, ref = got.double().cpu(), ref.double().cpu()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    ad = (got - ref).abs().max().item()
    rd = O.rdiff(got, ref).mean().item()
    assert ad <= ATOL and rd <= RTOL, f"{what}: max abs {ad:.3e}, mean rdiff {rd:.3e}"


def _shapes():
    """(b, sq, sk) of every flash_attention call hydragen_attention makes on the reference's test cases: level i of a
    case batches final_batch / n_i queries per shared sequence (sq) against its keys; ragged levels are padded here
    to their maximum (the varlen form is pinned separately below)."""
    seen = []
    for sizes in O.REFERENCE_SIZES_LIST:
        fb = len(sizes[-1])
        for lens in sizes[:-1]:
            n = len(lens)
            if len(set(lens)) == 1:
                seen.append((n, fb // n, lens[0]))
        seen.append((fb, 1, max(sizes[-1])))
    return sorted(set(seen))


@pytest.mark.parametrize("kvh", [1, 8])
@pytest.mark.parametrize("shape", _shapes(), ids=lambda s: "x".join(map(str, s)))
def test_primitive_matches_flash_attn(shape, kvh):
    from hydragen_b200.flash import flash_attention

    b, sq, sk = shape
    g = torch.Generator().manual_seed(b * 1000 + sq * 10 + kvh)
    q = torch.randn(b, sq, 8, 128, generator=g).half()
    k = torch.randn(b, sk, kvh, 128, generator=g).half()
    v = torch.randn(b, sk, kvh, 128, generator=g).half()
    fo, fl, _ = fa.flash_attn_func(q.cuda(), k.cuda(), v.cuda(), softmax_scale=128**-0.5, causal=False, return_attn_probs=True)
    ro, rl = O.flash_attention(q, k, v)           # CPU oracle, fp64
    oo, ol = flash_attention(q.cuda(), k.cuda(), v.cuda())  # this repo's kernels
    _close(ro, fo, f"oracle vs flash-attn {fa.__version__}")
    _close(oo, fo, f"kernel vs flash-attn {fa.__version__}")
    assert (rl - fl.double().cpu()).abs().max().item() < LSE_TOL, "oracle lse vs flash-attn"  # both [b, h, sq]
    assert (ol.double().cpu() - fl.double().cpu()).abs().max().item() < LSE_TOL, "kernel lse vs flash-attn"


@pytest.mark.parametrize("shape", [(2, 7, 19), (1, 128, 128), (2, 300, 300), (1, 70, 200)], ids=lambda s: "x".join(map(str, s)))
def test_causal_primitive_matches_flash_attn(shape):
    """flash_attention(causal=True): bottom-right aligned when sq != sk (flash-attn >= 2.1), as the reference's prefill
    and suffix branches rely on (hydragen/attention.py:344, hydragen/llama.py:509, 537-542)."""
    from hydragen_b200.flash import flash_attention

    b, sq, sk = shape
    g = torch.Generator().manual_seed(sq)
    q = torch.randn(b, sq, 8, 128, generator=g).half()
    k = torch.randn(b, sk, 4, 128, generator=g).half()
    v = torch.randn(b, sk, 4, 128, generator=g).half()
    fo, fl, _ = fa.flash_attn_func(q.cuda(), k.cuda(), v.cuda(), softmax_scale=128**-0.5, causal=True, return_attn_probs=True)
    ro, rl = O.flash_attention(q, k, v, causal=True)
    oo, ol = flash_attention(q.cuda(), k.cuda(), v.cuda(), causal=True)
    _close(ro, fo, "oracle vs flash-attn (causal)")
    _close(oo, fo, "kernel vs flash-attn (causal)")
    assert (rl - fl.double().cpu()).abs().max().item() < LSE_TOL
    assert (ol.double().cpu() - fl.double().cpu()).abs().max().item() < LSE_TOL


def test_varlen_primitive_matches_flash_attn():
    """flash_attention_varlen on the ragged level of the reference's 4th case ([9, 10, 11, 4] keys, 2 queries each):
    installed flash-attn returns the LSE as [h, total_q] (2.8) where v2.3.6 returned [n, h, max_q] -- compared per row."""
    from hydragen_b200.flash import flash_attention_varlen

    lens, qps = [9, 10, 11, 4], 2
    n = len(lens)
    g = torch.Generator().manual_seed(11)
    q = torch.randn(n * qps, 8, 128, generator=g).half()
    k = torch.randn(sum(lens), 8, 128, generator=g).half()
    v = torch.randn(sum(lens), 8, 128, generator=g).half()
    cu_q = (torch.arange(0, n + 1, dtype=torch.int32) * qps)
    cu_k = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32)
    res = fa.flash_attn_varlen_func(q.cuda(), k.cuda(), v.cuda(), cu_q.cuda(), cu_k.cuda(), qps, max(lens), softmax_scale=128**-0.5,
                                    causal=False, return_attn_probs=True)
    fo, fl = res[0], res[1]
    ro, rl = O.flash_attention_varlen(q, k, v, cu_q, cu_k, qps, max(lens))
    oo, ol = flash_attention_varlen(q.cuda(), k.cuda(), v.cuda(), cu_q.cuda(), cu_k.cuda(), qps, max(lens))
    _close(ro, fo, "oracle vs flash-attn (varlen)")
    _close(oo, fo, "kernel vs flash-attn (varlen)")
    fl = fl.double().cpu()
    if fl.ndim == 2:  # [h, total_q] -> [n, h, max_q]
        fl = fl.view(8, n, qps).permute(1, 0, 2)
    assert (rl - fl).abs().max().item() < LSE_TOL
    assert (ol.double().cpu() - fl).abs().max().item() < LSE_TOL
