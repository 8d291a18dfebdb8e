"""K1b tile-buffer rings (csrc/nls_pass_tiled.cuh, nt_ring_pos): a parity-only barrier model under random and
adversarial schedules.  mbarrier waits see one bit of the phase count, so the hand-off is only sound if no
waiter is ever two phases away from its barrier.  The per-consumer rings guarantee that; one ring shared by
independent consumers does not (shown here with a slow consumer), which is why the kernel does not use it."""
import random

import pytest

NCONS, NPROD = 4, 3
NP = NCONS * NPROD


def ring_pos_per_consumer(q, r, nbuf):
    c, j = divmod(q, NPROD)
    rb, rx = divmod(nbuf, NCONS)
    R = rb + (1 if c < rx else 0)
    base = c * rb + min(c, rx)
    Lc = r * NPROD + j
    return base + Lc % R, Lc // R


def ring_pos_shared(q, r, nbuf):
    L = r * NP + q
    return L % nbuf, L // nbuf


class Sim:
    """Producers and consumers as step functions; barriers keep the true phase count but waits test parity only,
    like mbarrier.try_wait.parity.  `violations` counts waits that pass while the true phase is not the awaited one."""

    def __init__(self, pos, nbuf, rounds, weight):
        self.pos, self.nbuf, self.rounds, self.weight = pos, nbuf, rounds, weight
        self.full = [0] * nbuf   # completed phases
        self.empty = [0] * nbuf
        self.owner = [None] * nbuf  # slab currently held in the buffer
        self.prod_r = [0] * NP
        self.cons_k = [0] * NCONS   # index into the consumer's sequence (r, j)
        self.violations = 0
        self.corrupt = 0

    @staticmethod
    def parity_wait_passes(completed, parity):
        # waiting on parity P passes when the phase in progress has the other parity
        return (completed & 1) != parity

    def step_producer(self, q):
        r = self.prod_r[q]
        if r >= self.rounds:
            return False
        buf, use = self.pos(q, r, self.nbuf)
        if not self.parity_wait_passes(self.empty[buf], (use & 1) ^ 1):
            return False
        if self.empty[buf] != use:          # passed although `use` consumptions have not happened
            self.violations += 1
        if self.owner[buf] is not None:
            self.corrupt += 1               # overwrote a tile that was never consumed
        self.owner[buf] = (q, r)
        self.full[buf] += 1
        self.prod_r[q] = r + 1
        return True

    def step_consumer(self, c):
        k = self.cons_k[c]
        r, j = divmod(k, NPROD)
        if r >= self.rounds:
            return False
        q = c * NPROD + j
        buf, use = self.pos(q, r, self.nbuf)
        if not self.parity_wait_passes(self.full[buf], use & 1):
            return False
        if self.full[buf] != use + 1:
            self.violations += 1
        if self.owner[buf] != (q, r):
            self.corrupt += 1
        self.owner[buf] = None
        self.empty[buf] += 1
        self.cons_k[c] = k + 1
        return True

    def run(self, rng, max_steps=200000):
        actors = [("p", q) for q in range(NP)] + [("c", c) for c in range(NCONS)]
        w = [self.weight(a) for a in actors]
        idle = 0
        for _ in range(max_steps):
            kind, i = rng.choices(actors, weights=w)[0]
            moved = self.step_producer(i) if kind == "p" else self.step_consumer(i)
            idle = 0 if moved else idle + 1
            if all(r >= self.rounds for r in self.prod_r) and all(k >= self.rounds * NPROD for k in self.cons_k):
                return True
            if idle > 5000:
                return False  # stuck
        return False


@pytest.mark.parametrize("nbuf", [12, 13, 14, 15, 16, 19, 24])
def test_per_consumer_rings_never_alias(nbuf):
    for seed in range(3):
        rng = random.Random(seed)
        slow = rng.randrange(NCONS)
        # one consumer 50x slower than the rest, random producer speeds: the adversarial case
        weight = lambda a, slow=slow, rng=rng: (0.02 if a == ("c", slow) else 1.0) * (0.2 + rng.random())
        sim = Sim(ring_pos_per_consumer, nbuf, rounds=40, weight=weight)
        assert sim.run(rng), "hand-off stalled"
        assert sim.violations == 0 and sim.corrupt == 0


def test_ring_shared_by_independent_consumers_can_alias():
    """the layout the kernel had for a short while (L mod 15 over all producers): a lagging consumer lets a
    producer of another consumer pass a parity wait two phases early"""
    hits = 0
    for seed in range(20):
        rng = random.Random(seed)
        weight = lambda a: 0.002 if a == ("c", 0) else 1.0
        sim = Sim(ring_pos_shared, 15, rounds=40, weight=weight)
        sim.run(rng, max_steps=60000)
        hits += sim.violations + sim.corrupt
    assert hits > 0


def test_every_slab_gets_a_distinct_buffer_within_a_ring_window():
    for nbuf in (12, 15, 17):
        for q in range(NP):
            c = q // NPROD
            rb, rx = divmod(nbuf, NCONS)
            R = rb + (1 if c < rx else 0)
            seen = {}
            for r in range(50):
                buf, use = ring_pos_per_consumer(q, r, nbuf)
                base = c * rb + min(c, rx)
                assert base <= buf < base + R
                seen.setdefault(buf, []).append(use)
            for uses in seen.values():
                assert uses == sorted(uses)
