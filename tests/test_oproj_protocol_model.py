"""CPU model of the cross-rank protocol of the fused o_proj + all-reduce launch (csrc/oproj_allreduce.cu), run under random
interleavings of every agent of every rank.  It restates the protocol, not the kernel's arithmetic: what is checked is the ORDER of the
events the flags allow --

  * a slice is reduced only when every rank's partial OF THIS CALL is in place (never a stale partial, never an earlier sum);
  * a rank's grid retires only when every slice of its buffer holds this call's sum (what the next kernel on the stream reads);
  * ranks run free of each other (one may be a whole call ahead), calls repeat on the same buffer, nothing deadlocks;

with the flag words used exactly as in the kernel: epoch read once at grid start, `ready[tile][src]` written with the epoch by the
producer of a tile into the OWNER's array and compared ">= epoch" (signed), CTAs counted out on a local word, the last one exchanging
`out[src]` with the peers, resetting the count and bumping the epoch.  Two negative controls show that the model can fail: without the
end-of-call exchange, and with the tile flag raised before the tile is written, violations are found.  The tiling / ownership
formulas themselves are checked against the real host/device header in tests/test_oproj_plan.py; the real kernel against matmul + NCCL
on 2 / 4 / 8 GPUs in tests/test_multigpu_gpu.py.
"""
import random

import pytest


class Violation(Exception):
    pass


class Rank:
    def __init__(self, world, n_tiles, slices):
        self.epoch, self.done = 0, 0
        self.ready = [[0] * world for _ in range(n_tiles)]  # ready[tile][src], meaningful in the owner's copy
        self.out = [0] * world
        # buffer content per (tile, slice): ("partial", call) written by this rank's GEMM, ("sum", call) by the tile's owner
        self.buf = {(t, s): ("sum", 0) for t in range(n_tiles) for s in range(slices)}
        self.retired = 0  # calls whose grid has retired on this rank


def run(world, n_tiles, ctas, slices, calls, seed, end_barrier=True, flag_before_store=False, max_steps=200000):
    rng = random.Random(seed)
    ranks = [Rank(world, n_tiles, slices) for _ in range(world)]

    def grid(r):
        """Generator-based agents of one call on rank r; yields after every externally visible event; yields 'blocked' while spinning."""
        me = ranks[r]
        e = me.epoch + 1  # read once, after the previous grid of this rank has retired
        agents = []

        def gemm(c):
            for tile in range(c, n_tiles, ctas):
                owner = ranks[tile % world]
                if flag_before_store:  # negative control: the flag overtakes the data
                    owner.ready[tile][r] = e
                    yield
                for s in range(slices):
                    me.buf[(tile, s)] = ("partial", e)
                yield  # bulk store complete + gpu-scope fence
                owner.ready[tile][r] = e
                yield

        def reducer(c):
            owned = [t for t in range(n_tiles) if t % world == r]
            units = [(t, s) for t in owned for s in range(slices)]
            for tile, s in units[c::ctas]:
                while any(me.ready[tile][src] - e < 0 for src in range(world)):
                    yield "blocked"
                for src in range(world):  # multimem.ld_reduce: the switch reads the slice from every rank
                    if ranks[src].buf[(tile, s)] != ("partial", e):
                        raise Violation(f"rank {r} call {e}: slice {(tile, s)} of rank {src} holds {ranks[src].buf[(tile, s)]}")
                yield
                for dst in range(world):  # multimem.st
                    ranks[dst].buf[(tile, s)] = ("sum", e)
                yield

        state = {"left": ctas}

        def cta(c):
            # the GEMM warps and the reduce warps of a CTA run beside each other
            subs = [gemm(c), reducer(c)]
            while subs:
                g = rng.choice(subs)
                try:
                    v = next(g)
                    yield v
                except StopIteration:
                    subs.remove(g)
            # count out (fence.sys + local atomic); only the last CTA goes on
            me.done += 1
            last = me.done == ctas
            yield
            if last:
                if end_barrier:
                    for p in range(world):
                        ranks[p].out[r] = e
                    yield
                    while any(me.out[p] - e < 0 for p in range(world)):
                        yield "blocked"
                me.done = 0
                yield
                me.epoch = e
            state["left"] -= 1

        agents = [cta(c) for c in range(ctas)]
        while agents:
            a = rng.choice(agents)
            try:
                yield next(a)
            except StopIteration:
                agents.remove(a)
        # grid retired: the next kernel on this rank's stream reads the buffer
        for key, val in me.buf.items():
            if val != ("sum", e):
                raise Violation(f"rank {r} retired call {e} with slice {key} = {val}")
        me.retired = e

    streams = [iter(()) for _ in range(world)]
    issued = [0] * world
    steps = blocked_in_a_row = 0
    while True:
        live = []
        for r in range(world):
            live.append(r)
        r = rng.choice(live)
        try:
            v = next(streams[r])
        except StopIteration:
            if issued[r] == calls:
                if all(issued[p] == calls and ranks[p].retired == calls for p in range(world)):
                    return ranks
                v = "blocked"
            else:
                issued[r] += 1
                streams[r] = grid(r)
                v = None
        blocked_in_a_row = blocked_in_a_row + 1 if v == "blocked" else 0
        steps += 1
        if blocked_in_a_row > 20000 or steps > max_steps * calls:
            raise Violation("no progress (deadlock or livelock)")


@pytest.mark.parametrize("world,n_tiles,ctas,slices", [(2, 4, 2, 2), (2, 5, 3, 1), (3, 7, 2, 2), (4, 8, 3, 2), (8, 16, 2, 1), (4, 3, 5, 2)])
def test_protocol_holds_under_random_interleavings(world, n_tiles, ctas, slices):
    for seed in range(40):
        ranks = run(world, n_tiles, ctas, slices, calls=3, seed=seed)
        assert all(rk.epoch == 3 and rk.done == 0 for rk in ranks)


def test_model_detects_a_missing_end_barrier():
    """Without the exchange of `out` flags a fast rank starts the next call and overwrites partials a slow owner has yet to reduce."""
    found = 0
    for seed in range(60):
        try:
            run(2, 4, 2, 2, calls=3, seed=seed, end_barrier=False)
        except Violation:
            found += 1
    assert found > 0


def test_model_detects_a_flag_that_overtakes_its_tile():
    """The measured failure of r02zc (flag sent with no fence after the bulk store), in the model: the owner reduces a stale tile."""
    found = 0
    for seed in range(60):
        try:
            run(2, 4, 2, 2, calls=2, seed=seed, flag_before_store=True)
        except Violation:
            found += 1
    assert found > 0
