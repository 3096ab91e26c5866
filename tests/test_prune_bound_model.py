"""Model check of the per-query bound of work-bounding hit limits (csrc/fmb_scheme.cuh prune_insert / prune_bound): the n smallest
discovery-order keys of a query are kept in n slots; a new key is bubbled in by one atomicMin per slot whose displaced (larger) value
is carried to the next slot.  The kernel reads slot n-1 as "no later row may exceed this key" WHILE other lanes are inserting, so the
claim that matters is about every intermediate state, not only the state at rest:

    whenever slot n-1 holds a finite x, at least n of the keys whose insertion has started are <= x

(then dropping a subtree whose smallest reachable key is > x can never drop one of the query's first n rows).  The model below runs the
insertions of many "lanes" under random interleavings, one atomicMin per step, and checks that claim after every step, plus the state at
rest: the slots hold exactly the n smallest keys.  This is a restatement in Python of the device code, kept next to it as the argument
for its comment -- the device code itself is exercised by tests/test_gpu_search_n.py."""
import random

import pytest

INF = (1 << 64) - 1


class Lane:
    """one prune_insert(key, copies) call, advanced one atomicMin at a time"""

    def __init__(self, key, copies, n):
        self.key, self.copies, self.n = key, min(copies, n), n
        self.c, self.i, self.k = 0, 0, key           # copy being inserted, slot it is at, value it carries

    def done(self):
        return self.c >= self.copies

    def step(self, slots):
        old = slots[self.i]
        slots[self.i] = min(old, self.k)             # atomicMin
        self.k = max(old, self.k)                    # the displaced (or unchanged) larger value goes on
        self.i += 1
        if self.i >= self.n or self.k == INF:
            self.c += 1
            self.i, self.k = self.c, self.key         # the next copy starts at slot c: slots 0 .. c-1 already hold values <= key


@pytest.mark.parametrize("n", [1, 2, 3, 8])
def test_bound_is_valid_at_every_moment(n):
    rng = random.Random(1234 + n)
    for trial in range(300):
        n_hits = rng.randint(1, 12)
        keys = rng.sample(range(1, 1000), n_hits)
        lanes = [Lane(k, rng.choice([1, 1, 1, 2, 5, 20]), n) for k in keys]
        slots = [INF] * n
        started = []                                  # keys (with multiplicity min(copies, n)) whose insertion has started
        order = list(range(len(lanes)))
        while any(not l.done() for l in lanes):
            li = rng.choice([j for j in order if not lanes[j].done()])
            lane = lanes[li]
            if lane.c == 0 and lane.i == 0 and lane.k == lane.key and (lane.key, li) not in [(s[0], s[1]) for s in started]:
                started += [(lane.key, li)] * lane.copies
            lane.step(slots)
            if slots[n - 1] != INF:
                assert sum(1 for s in started if s[0] <= slots[n - 1]) >= n, (trial, slots, started)
        everything = sorted(s[0] for s in started)
        assert slots == (everything + [INF] * n)[:n], (trial, slots, everything)


def test_equal_keys_of_one_wide_hit_fill_the_slots():
    slots = [INF] * 4
    lane = Lane(77, 9, 4)                             # one hit of nine rows against a limit of four
    while not lane.done():
        lane.step(slots)
    assert slots == [77, 77, 77, 77]
