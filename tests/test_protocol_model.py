"""CPU model of the synchronisation protocol of the alternate-block softmax variant of the prefix kernel
(csrc/prefix_sm100.cu, ``kSplit == 3``: HYDRAGEN_B200_PREFIX_SOFTMAX=alt), which has not been run on hardware yet.

The model keeps exactly the kernel's barrier set for one Q tile -- ``s_full[2]``, ``p_full[2]``, ``pv_done``, ``o_full``
and the per-lane-quarter ``m_ready[4][2]`` -- with mbarrier phase/parity semantics (``try_wait.parity(P)`` succeeds once the
phase with parity P has completed), the kernel's own parity expressions (``(j >> 1) & 1``, ``((j - 1) >> 1) & 1``,
``(j - 1) & 1`` ...) and an in-order tensor pipe whose completions lag their issue, and runs the MMA warp and the eight
softmax warps under random schedules.  It asserts that no wait ever passes before the completion it stands for has
happened (a parity expression that is off by one phase shows up here) and that every schedule terminates."""

import random

import pytest


class Bar:
    def __init__(self, count=1): self.count=count; self.arr=0; self.phase=0
    def arrive(self):
        self.arr+=1
        if self.arr==self.count: self.arr=0; self.phase+=1
    def test(self, parity): return (self.phase & 1) != parity   # try_wait.parity: the phase with this parity has completed

def run(n_blocks, seed, rescale_prob):
    rnd=random.Random(seed)
    s_full=[Bar(),Bar()]; p_full=[Bar(4),Bar(4)]   # 4 warps per (tile, parity) warpgroup stand for the 128 threads
    pv_done=Bar(); o_full=Bar(); m_ready=[[Bar() for _ in range(2)] for _ in range(4)]  # [quarter][parity], 1 arrival = the warp
    pending=[]   # tensor-pipe completions in issue order: list of callables
    log=[]
    def mma():
        for u in (-2,-1):
            if u+2<n_blocks:
                b=(u+2)&1
                pending.append(lambda b=b: s_full[b].arrive())
            yield
        for j in range(n_blocks):
            b=j&1
            while not p_full[b].test((j>>1)&1): yield
            assert p_full[b].phase >= (j>>1)+1, ("p_full early", j)
            last = j+1==n_blocks
            pending.append((lambda: o_full.arrive()) if last else (lambda: pv_done.arrive()))
            if j+2<n_blocks: pending.append(lambda b=b: s_full[b].arrive())
            yield
    def softmax(q, par):
        m_w=None
        for j in range(par, n_blocks, 2):
            while not s_full[par].test((j>>1)&1): yield
            assert s_full[par].phase >= (j>>1)+1, ("s_full early", j)
            yield  # pass 1
            if j>0:
                while not m_ready[q][par^1].test(((j-1)>>1)&1): yield
                assert m_ready[q][par^1].phase >= ((j-1)>>1)+1, ("m_ready early", j)
                if rnd.random()<rescale_prob:
                    while not pv_done.test((j-1)&1): yield
                    assert pv_done.phase >= j, ("pv_done early", j, pv_done.phase)
            m_ready[q][par].arrive()
            yield  # pass 2
            p_full[par].arrive()
            yield
        last=n_blocks-1
        if (last&1)!=par:
            while not m_ready[q][par^1].test((last>>1)&1): yield
            assert m_ready[q][par^1].phase >= (last>>1)+1, ("final m_ready early",)
        while not o_full.test(0): yield
        log.append(("done",q,par))
    procs=[mma()]+[softmax(q,par) for q in range(4) for par in range(2)]
    alive=list(procs)
    steps=0
    while alive:
        steps+=1
        if steps>200000: raise RuntimeError(("deadlock", n_blocks, seed))
        # tensor pipe completes its oldest op now and then
        if pending and rnd.random()<0.3: pending.pop(0)()
        p=rnd.choice(alive)
        try: next(p)
        except StopIteration: alive.remove(p)
    assert len(log)==8 and not pending or all(True for _ in pending)


@pytest.mark.parametrize("n_blocks", list(range(1, 12)))
def test_alt_softmax_protocol_model(n_blocks):
    for seed in range(40):
        run(n_blocks, seed, rescale_prob=0.3)
