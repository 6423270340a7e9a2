"""CPU tests of the persistent prefix kernel's work schedule (csrc/prefix_sched.h through hg_prefix_schedule: the same
integer arithmetic the device runs).  For every configuration -- BASELINE.json's cfg#2 / cfg#4 / cfg#5 per-GPU shapes,
head-parallel ranks with few heads, ragged levels, tiny problems -- the pieces handed to the CTAs must tile the
(unit, key block) space exactly once, respect the minimum piece length, use at most two workspace slots per CTA,
and the merge's view of a split unit (which CTAs hold its pieces, in which slot) must agree with the CTAs' own view
(checked inside hg_prefix_schedule, which fails if not)."""

import collections

import pytest

from hydragen_b200 import _lib

CASES = {
    "cfg2": ([(1, 2048, 0)], 1024, 32),
    "cfg2_b4096": ([(1, 2048, 0)], 4096, 32),
    "cfg4_two_levels": ([(1, 1024, 0), (32, 64, 0)], 1024, 32),
    "cfg5_per_gpu": ([(1, 16384, 0)], 2048, 5),
    "tp8_rank_of_cfg2": ([(1, 2048, 0)], 1024, 4),
    "one_head": ([(1, 2048, 0)], 1024, 1),
    "ragged_level": ([(5, 0, 700)], 200, 4),
    "three_levels": ([(1, 300, 0), (2, 0, 130), (24, 3, 0)], 72, 8),
    "tiny": ([(1, 1, 0)], 1, 1),
    "short_keys_many_units": ([(64, 17, 0)], 64 * 300, 8),
}


def _units(levels, n_q_rows, hq):
    out = []
    for ng, kl, mk in levels:
        tiles = (n_q_rows // ng + 255) // 256
        nb = ((mk if mk > 0 else kl) + 63) // 64
        out += [nb] * (ng * tiles * hq)
    return out


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("n_sms", [148, 7, 1])
def test_stream_k_schedule_tiles_every_unit_exactly_once(name, n_sms, built_lib):
    levels, n_q_rows, hq = CASES[name]
    n_ctas, pieces = _lib.prefix_schedule(levels, n_q_rows, hq, n_sms=n_sms, allow_split=2)
    nbs = _units(levels, n_q_rows, hq)
    assert 1 <= n_ctas <= n_sms
    covered = collections.defaultdict(list)
    per_cta = collections.defaultdict(list)
    for cta, unit, level, head, grp, mt, b_lo, b_hi, split, slot in pieces:
        assert 0 <= cta < n_ctas and 0 <= unit < len(nbs) and 0 <= head < hq
        assert 0 <= b_lo < b_hi <= nbs[unit] or (nbs[unit] == 0 and b_lo == b_hi == 0)
        assert split == (0 if (b_lo == 0 and b_hi == nbs[unit]) else 1)
        if split:
            assert b_hi - b_lo >= 4, "no piece shorter than min_piece key blocks"
        covered[unit].append((b_lo, b_hi))
        per_cta[cta].append((unit, b_lo, split, slot))
    assert sorted(covered) == list(range(len(nbs)))
    for unit, segs in covered.items():
        segs.sort()
        assert segs[0][0] == 0 and segs[-1][1] == nbs[unit]
        assert all(a[1] == b[0] for a, b in zip(segs, segs[1:])), f"unit {unit}: {segs}"
        assert len(segs) <= 24
    for cta, ps in per_cta.items():
        assert ps == sorted(ps), "a CTA walks its range in (unit, block) order"
        split_slots = [slot for _, _, split, slot in ps if split]
        assert len(split_slots) <= 2 and len(set(split_slots)) == len(split_slots)
        assert all(slot == (0 if i == 0 else 1) for i, (_, _, _, slot) in enumerate(ps))
    # balance: no CTA carries more than its share plus one unit's fixed cost plus the snapping slack
    cost = lambda c: sum((b_hi - b_lo) + (2 if b_lo == 0 else 0) for cc, u, *_r, b_lo, b_hi, _s, _t in pieces if cc == c)
    if n_ctas > 1:
        total = sum(nb + 2 for nb in nbs)
        assert max(cost(c) for c in range(n_ctas)) <= total / n_ctas + 2 + 2 * 4 + 1


@pytest.mark.parametrize("name", ["cfg2", "cfg4_two_levels", "short_keys_many_units"])
def test_whole_unit_schedule(name, built_lib):
    levels, n_q_rows, hq = CASES[name]
    n_ctas, pieces = _lib.prefix_schedule(levels, n_q_rows, hq, allow_split=False)
    nbs = _units(levels, n_q_rows, hq)
    assert n_ctas == min(148, len(nbs))
    assert sorted(p[1] for p in pieces) == list(range(len(nbs)))
    assert all(p[8] == 0 and p[6] == 0 and p[7] == nbs[p[1]] and p[0] == p[1] % n_ctas for p in pieces)


def test_stream_k_balances_cfg2_over_every_sm(built_lib):
    """Forced stream-K at cfg#2: all 148 SMs get equal shares of the 128 x 32 key blocks."""
    n_ctas, pieces = _lib.prefix_schedule(*CASES["cfg2"], allow_split=2)
    assert n_ctas == 148
    blocks = collections.Counter()
    for p in pieces:
        blocks[p[0]] += p[7] - p[6]
    assert max(blocks.values()) - min(blocks.values()) <= 8 and sum(blocks.values()) == 128 * 32


def test_units_are_cut_only_where_it_pays(built_lib):
    """The launch's own choice (allow_split=1): cutting units costs about 6 key blocks on the critical path (measured), so
    cfg#2's 128 units on 148 SMs stay whole (32 blocks vs 27.7 + 6), while few long units (cfg#5 per GPU, the ranks of a
    TP run) and a ragged last wave (B = 4096: 3.46 waves) are cut."""
    split = lambda name: any(p[8] for p in _lib.prefix_schedule(*CASES[name], allow_split=1)[1])
    assert not split("cfg2") and not split("cfg4_two_levels")
    assert split("cfg5_per_gpu") and split("tp8_rank_of_cfg2") and split("cfg2_b4096") and split("one_head")
