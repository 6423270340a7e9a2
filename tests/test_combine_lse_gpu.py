"""GPU parity tests of the combine kernel: the reference's own test grid
(tests/test_combine_lse.py:9-29: bs, seq, heads in {1,2,3}, hdim in {63,64,128,129}, fp32) against
the golden produced by the reference's combine_lse_torch and against the torch statement on the
device; n-way fan-in; 16-bit dtypes; -inf partials."""

from itertools import product

import pytest
import torch

from oracle import hydragen_oracle as O

pytestmark = pytest.mark.gpu


def test_combine_matches_reference_golden(golden):
    from hydragen_b200.attention import combine_lse

    keys = sorted({k.split("/")[0] for k in golden.files if k.startswith("combine_")})
    for key in keys:
        n = int(key.split("_")[1][1:])
        outs = [torch.from_numpy(golden[f"{key}/o{i}"]).cuda() for i in range(n)]
        lses = [torch.from_numpy(golden[f"{key}/l{i}"]).cuda() for i in range(n)]
        got = combine_lse(outs, lses)
        ref = torch.from_numpy(golden[key + "/out"])
        assert torch.allclose(got.cpu(), ref, rtol=2e-5, atol=2e-6), key


def test_reference_grid():
    """The reference's sweep and criterion (mean rdiff < 0.1) -- and a much tighter one."""
    from hydragen_b200.attention import combine_lse_torch, combine_lse_triton

    torch.manual_seed(0)
    for bs, s, h, d in product([1, 2, 3], [1, 2, 3], [1, 2, 3], [63, 64, 128, 129]):
        o1 = torch.rand(bs, s, h, d, device="cuda")
        o2 = torch.rand_like(o1)
        l1 = torch.rand(bs, s, h, device="cuda")
        l2 = torch.rand_like(l1)
        a = combine_lse_torch([o1, o2], [l1, l2])
        b = combine_lse_triton(o1, l1, o2, l2)
        assert O.rdiff(a, b).mean().item() < 0.1
        assert torch.allclose(a, b, rtol=2e-5, atol=2e-6), (bs, s, h, d)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("n", [1, 2, 3, 5, 8])
def test_nway_and_dtypes(dtype, n):
    from hydragen_b200.attention import combine_lse_cuda

    g = torch.Generator().manual_seed(n)
    shape = (33, 1, 32, 128)
    outs = [torch.randn(*shape, generator=g).to(dtype) for _ in range(n)]
    lses = [torch.randn(*shape[:-1], generator=g) * 6 for _ in range(n)]
    if n >= 3:
        lses[1][0] = float("-inf")  # an empty partial contributes nothing
    got, lse = combine_lse_cuda([o.cuda() for o in outs], [l.cuda() for l in lses], return_lse=True)
    ref = O.combine_lse_torch([o.double() for o in outs], [l.double() for l in lses])
    tol = {torch.float16: 2e-3, torch.bfloat16: 1.6e-2, torch.float32: 1e-5}[dtype]
    assert (got.double().cpu() - ref).abs().max().item() <= tol
    ref_lse = torch.logsumexp(torch.stack(lses).double(), 0)
    assert (lse.double().cpu() - ref_lse).abs().max().item() < 1e-4


def test_all_partials_empty():
    from hydragen_b200.attention import combine_lse_cuda

    o = torch.zeros(2, 1, 4, 64, device="cuda")
    l = torch.full((2, 1, 4), float("-inf"), device="cuda")
    got, lse = combine_lse_cuda([o, o], [l, l], return_lse=True)
    assert torch.all(got == 0) and torch.all(torch.isinf(lse))


def test_argument_errors():
    from hydragen_b200.attention import combine_lse_cuda

    o = torch.zeros(2, 1, 4, 64, device="cuda")
    l = torch.zeros(2, 1, 4, device="cuda")
    with pytest.raises(ValueError):
        combine_lse_cuda([], [])
    with pytest.raises(ValueError):
        combine_lse_cuda([o] * 9, [l] * 9)
    with pytest.raises(ValueError):
        combine_lse_cuda([o, o], [l, l[:1]])
