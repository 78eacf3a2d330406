"""Executable model of the paired send's tile order and per-tile handshake (no GPU, no library).

The migration fused into a tile pass pairs tile t of one rank with tile t ^ (1 << cbit) of its
partner.  Three index rules carry the protocol; they are restated here from the sources and checked exhaustively:

* fused pass (rustqip_b200/csrc/jit_codegen.cpp, QIP_PAIRED): launch position b walks tile swapbits(b, 0, cbit);
* stand-in kernel (rustqip_b200/csrc/dist.cu, k_paired_send): CTA j handles tile swapbits(2 j + give, 0, cbit);
* a tile is given away when its counter bit cbit equals `give`; the two ranks of a pair have opposite `give`.

and a small discrete simulation runs both sides with a BOUNDED number of resident CTAs (the hardware limit that makes
the order matter): whatever the mix of fused passes and stand-in kernels, every tile completes -- no deadlock."""
import itertools

import pytest


def swapbits(x, a, b):
    if a == b:
        return x
    ba, bb = (x >> a) & 1, (x >> b) & 1
    return (x & ~((1 << a) | (1 << b))) | (bb << a) | (ba << b)


def fused_order(k, cbit):
    """tile counter values in launch order of the fused pass over 2^k tiles"""
    return [swapbits(b, 0, cbit) for b in range(1 << k)]


def standin_order(k, cbit, give):
    """tiles of the give-half in CTA order of the stand-in kernel"""
    return [swapbits(2 * j + give, 0, cbit) for j in range(1 << (k - 1))]


@pytest.mark.parametrize("k", [1, 2, 5, 9])
def test_orders_are_permutations_and_agree(k):
    for cbit in range(k):
        order = fused_order(k, cbit)
        assert sorted(order) == list(range(1 << k))  # every tile exactly once
        for give in (0, 1):
            gives = [t for t in order if (t >> cbit) & 1 == give]
            assert gives == standin_order(k, cbit, give)  # the j-th tile given away is the same tile on both code paths
            # the partner (opposite give) reaches the PAIRED tile at the same position j
            partner = [t for t in order if (t >> cbit) & 1 == 1 - give]
            assert [t ^ (1 << cbit) for t in gives] == partner
        if cbit:  # kept and given tiles alternate: HBM and NVLink traffic overlap
            assert all(((order[2 * j] >> cbit) & 1) != ((order[2 * j + 1] >> cbit) & 1) for j in range(1 << (k - 1)))


class Side:
    """One rank: a queue of CTAs in launch order, at most `slots` resident at a time.  A CTA of a given tile: load ->
    announce to the partner -> (compute) -> wait for the partner's announcement of the paired tile -> store -> exit.
    CTAs of kept tiles (fused pass only) never wait."""

    def __init__(self, k, cbit, give, fused, slots):
        self.cbit, self.give = cbit, give
        self.queue = fused_order(k, cbit) if fused else standin_order(k, cbit, give)
        self.resident = []  # [tile, announced]
        self.slots = slots
        self.flags = set()  # announcements received from the partner (tile indices of THIS side)
        self.done = 0

    def gives(self, t):
        return (t >> self.cbit) & 1 == self.give

    def step(self, partner):
        progressed = False
        while self.queue and len(self.resident) < self.slots:  # the block scheduler fills free slots in launch order
            self.resident.append([self.queue.pop(0), False])
            progressed = True
        for cta in list(self.resident):
            t = cta[0]
            if not self.gives(t):
                self.resident.remove(cta)  # kept tile: load, compute, local store
                self.done += 1
                progressed = True
                continue
            if not cta[1]:
                partner.flags.add(t ^ (1 << self.cbit))  # "my copy of t is in shared memory"
                cta[1] = True
                progressed = True
            if t in self.flags:  # the partner has loaded the tile this one overwrites
                self.resident.remove(cta)
                self.done += 1
                progressed = True
        return progressed


@pytest.mark.parametrize("fused_a,fused_b", list(itertools.product([True, False], repeat=2)))
@pytest.mark.parametrize("slots_a,slots_b", [(1, 1), (2, 5), (4, 4), (7, 3)])
def test_bounded_resident_sets_never_deadlock(fused_a, fused_b, slots_a, slots_b):
    k = 7
    for cbit in range(k):
        a = Side(k, cbit, 1, fused_a, slots_a)
        b = Side(k, cbit, 0, fused_b, slots_b)
        total_a, total_b = len(a.queue), len(b.queue)
        # an adversarial scheduler: the two GPUs advance at very different rates (a runs `ra` steps per step of b)
        for ra, rb in [(1, 1), (5, 1), (1, 5)]:
            a = Side(k, cbit, 1, fused_a, slots_a)
            b = Side(k, cbit, 0, fused_b, slots_b)
            for _ in range(100000):
                moved = False
                for _ in range(ra):
                    moved = a.step(b) or moved
                for _ in range(rb):
                    moved = b.step(a) or moved
                if a.done == total_a and b.done == total_b:
                    break
                assert moved, "deadlock: cbit %d, a %d/%d b %d/%d" % (cbit, a.done, total_a, b.done, total_b)
            assert a.done == total_a and b.done == total_b


def test_mismatched_orders_can_deadlock():
    """Negative control: a stand-in kernel walking the give-half in plain counter order against a fused pass that walks
    its tiles in the exchanged-bits order deadlocks with small resident sets -- why k_paired_send uses the same order."""
    k, cbit = 7, 6

    class Plain(Side):
        def __init__(self, *a):
            super().__init__(*a)
            self.queue = [t for t in range(1 << k) if self.gives(t)]

    a = Side(k, cbit, 1, True, 2)
    b = Plain(k, cbit, 0, False, 2)
    stuck = False
    for _ in range(10000):
        moved = a.step(b)
        moved = b.step(a) or moved
        if not moved:
            stuck = True
            break
        if a.done == 1 << k and b.done == 1 << (k - 1):
            break
    assert stuck
