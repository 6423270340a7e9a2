"""Host-side logic of the fused o_proj + all-reduce launch (csrc/oproj_sched.h, the formulas phase 2 of the kernel runs on the device,
replayed on the host by hg_oproj_allreduce_plan): for any world size, shape and launch geometry every 16-byte vector of the [m, n]
output is reduced exactly once, by exactly one rank; ownership is balanced; the flag array is large enough.  No GPU needed."""
import numpy as np
import pytest

from hydragen_b200 import _lib

SHAPES = [(1024, 4096), (2048, 5120), (200, 1032), (1, 8), (130, 264), (4096, 4096), (8, 11008)]


@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("m,n", SHAPES)
def test_every_vector_reduced_exactly_once(m, n, world):
    cover = np.zeros(m * n // 8, dtype=np.int32)
    owned = []
    for rank in range(world):
        mine = np.zeros_like(cover)
        g = _lib.oproj_allreduce_plan(m, n, world, rank, cover=mine)
        assert mine.max(initial=0) <= 1
        cover += mine
        owned.append(g["tiles_owned"])
        assert g["flag_words"] <= _lib.oproj_allreduce_flag_words(m, n, world)  # the bound the caller allocates by
        assert g["slices_owned"] == g["tiles_owned"] * (128 // (g["u"] * (32 // (g["bn"] // 8))))
    assert (cover == 1).all(), f"{int((cover != 1).sum())} of {cover.size} vectors not reduced exactly once"
    assert sum(owned) == g["tiles"] and max(owned) - min(owned) <= 1  # tiles dealt round-robin


@pytest.mark.parametrize("n_ctas", [1, 5, 37, 148])
def test_any_grid_size_covers_the_output(n_ctas):
    m, n, world = 1000, 2056, 4
    cover = np.zeros(m * n // 8, dtype=np.int32)
    for rank in range(world):
        g = _lib.oproj_allreduce_plan(m, n, world, rank, n_ctas=n_ctas, cover=cover)
        assert g["ctas"] == n_ctas
    assert (cover == 1).all()


def test_knobs_change_the_geometry_not_the_coverage(monkeypatch):
    m, n, world = 1024, 4096, 8
    for bn in ("128", "256"):
        for u in ("1", "2", "4"):
            for warps in ("1", "3", "8"):
                monkeypatch.setenv("HYDRAGEN_B200_OPROJ_BN", bn)
                monkeypatch.setenv("HYDRAGEN_B200_OPROJ_U", u)
                monkeypatch.setenv("HYDRAGEN_B200_OPROJ_RWARPS", warps)
                cover = np.zeros(m * n // 8, dtype=np.int32)
                for rank in range(world):
                    g = _lib.oproj_allreduce_plan(m, n, world, rank, cover=cover)
                assert (g["bn"], g["u"], g["reduce_warps"]) == (int(bn), int(u), int(warps))
                assert (cover == 1).all()


def test_single_rank_has_no_reduction_and_bad_arguments_are_refused():
    g = _lib.oproj_allreduce_plan(1024, 4096, 1, 0)
    assert g["bn"] == 256 and g["reduce_warps"] == 0 and g["ctas"] == 128  # the GEMM alone: one CTA per 128 x 256 tile
    with pytest.raises(ValueError):
        _lib.oproj_allreduce_plan(1024, 4100, 2, 0)  # n not a multiple of 8
    with pytest.raises(ValueError):
        _lib.oproj_allreduce_plan(1024, 4096, 2, 2)
