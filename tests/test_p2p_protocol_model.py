"""A model of the peer-memory exchange protocol (csrc/p2p_dev.cuh, csrc/p2p.cu) run under random schedules on the CPU.

What the CUDA kernels rely on, restated as a small state machine per CTA and checked by brute force:

* a region holds two data slots and two flag sets chosen by the parity of the rank's call counter;
* CTA c of call k stages its chunk in its own region, writes k into flag[parity][rank][c] of EVERY region, waits until
  flag[parity][p][c] == k for every p in its own region, then reads chunk c of every region's slot;
* the last CTA of a call to finish advances the rank's call counter; a rank's next kernel starts only after that
  (stream order); CTAs of one kernel run in any interleaving and never wait for each other;
* calls may differ in size and in their number of CTAs (so chunk boundaries move from call to call).

The invariant: every element a CTA reads carries the tag (call k, owner p) -- never a value staged for another call, even
though ranks drift apart by up to one call and slots are reused every second call.  A variant WITHOUT the parity
alternation must be caught by the same checker (so the checker is known to be able to fail).
"""
import random

import pytest


class Region:
    def __init__(self, world, max_ctas, slot_elems):
        self.call = 0
        self.done = 0
        self.flag = [[[0] * max_ctas for _ in range(world)] for _ in range(2)]
        self.slot = [[None] * slot_elems for _ in range(2)]


def chunks(n, G):
    per = (n + G - 1) // G
    return [(min(n, c * per), min(n, c * per + per)) for c in range(G)]


def run(world, calls, seed, alternate=True, max_steps=2_000_000):
    """calls: list of (n elements, G CTAs).  Returns None, or a description of the first violation."""
    rng = random.Random(seed)
    max_ctas = max(G for _, G in calls)
    slot_elems = max(n for n, _ in calls)
    regs = [Region(world, max_ctas, slot_elems) for _ in range(world)]
    # per rank: index of the kernel in flight and the program counter of each of its CTAs
    # pc: 0 begin, 1 stage, 2 signal, 3 wait, 4 read first half, 5 read second half, 6 close, 7 finished
    kern = [0] * world
    pcs = [[0] * calls[0][1] for _ in range(world)]
    seen = [[None] * calls[0][1] for _ in range(world)]          # the call number each CTA read at `begin`
    steps = 0
    while any(k < len(calls) for k in kern):
        steps += 1
        if steps > max_steps:
            return "no progress (deadlock?)"
        r = rng.randrange(world)
        if kern[r] >= len(calls):
            continue
        n, G = calls[kern[r]]
        c = rng.randrange(G)
        pc = pcs[r][c]
        me = regs[r]
        if pc == 7:
            continue
        if pc == 0:
            seen[r][c] = me.call + 1
        k = seen[r][c]
        par = (k & 1) if alternate else 0
        lo, hi = chunks(n, G)[c]
        if pc == 1:
            for e in range(lo, hi):
                me.slot[par][e] = (k, r)
        elif pc == 2:
            for p in range(world):
                regs[p].flag[par][r][c] = k
        elif pc == 3:
            if any(me.flag[par][p][c] != k for p in range(world)):
                continue                                         # keep spinning
        elif pc in (4, 5):
            mid = (lo + hi) // 2
            a, b = (lo, mid) if pc == 4 else (mid, hi)
            for p in range(world):
                for e in range(a, b):
                    if regs[p].slot[par][e] != (k, p):
                        return "rank %d call %d CTA %d read %r from rank %d element %d" % (r, k, c, regs[p].slot[par][e], p, e)
        elif pc == 6:
            me.done += 1
            if me.done == G:
                me.done = 0
                me.call = k
        pcs[r][c] = pc + 1
        if all(x == 7 for x in pcs[r]):                          # kernel finished: the next one of this rank may start
            kern[r] += 1
            if kern[r] < len(calls):
                G2 = calls[kern[r]][1]
                pcs[r] = [0] * G2
                seen[r] = [None] * G2
    return None


STEP = [(202, 13), (200, 25), (1349, 64), (1444, 64)]            # the four exchanges of a COGMEN step (sizes scaled down)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_slot_and_flag_discipline_holds_under_random_schedules(world):
    rng = random.Random(world)
    for seed in range(12):
        calls = STEP * 3 if seed % 2 == 0 else [(rng.randrange(1, 300), rng.randrange(1, 9)) for _ in range(10)]
        assert run(world, calls, seed) is None


def test_the_checker_catches_a_protocol_without_slot_alternation():
    """Same machine, one slot: a fast rank re-stages a chunk for call k+1 while a slow peer still reads call k."""
    hits = [run(2, STEP * 3, seed, alternate=False) for seed in range(20)]
    assert any(h is not None and "read" in h for h in hits), hits
