"""CPU: a small executable model of the peer-memory exchange protocol (visgeom_b200/csrc/vg_peer.cuh) under random
interleavings -- N ranks, two slot parities, words tagged with the exchange number, the rule "collect e-1 before
posting e" -- checking what the CUDA code relies on: every rank reads exactly the words of the exchange it waits for
(never an older or a newer one), nobody waits for ever, every rank forms the same rank-ordered sum.  The negative case
shows the rule is what makes two parities enough.  (The CUDA implementation itself is exercised on two GPUs by
tests/test_peer_exchange_gpu.py; this is the argument of vg_peer.cuh's header, made runnable.)"""
import random

import pytest


class Rank:
    def __init__(self, r, n, epochs, obey_rule):
        self.r, self.n, self.epochs = r, n, epochs
        self.inbox = [[(0, None)] * n for _ in range(2)]        # [parity][source] = (tag, value); zeroed: tag 0
        self.sums = {}
        # program: a list of micro-steps; a store to one destination / a poll of one source is one step
        self.prog = []
        for e in range(1, epochs + 1):
            if e > 1 and obey_rule:
                self.prog += [("poll", e - 1, s) for s in range(n)]
            self.prog += [("store", e, d) for d in range(n)]
        if obey_rule:
            self.prog += [("poll", epochs, s) for s in range(n)]
        else:                                                    # a sloppy rank: posts everything, collects afterwards
            for e in range(1, epochs + 1):
                self.prog += [("poll", e, s) for s in range(n)]
        self.pc = 0
        self.acc = {}

    def done(self):
        return self.pc >= len(self.prog)


def value(r, e):
    return (r + 1) * 1000 + e


def run(n, epochs, seed, obey_rule=True, max_steps=200000):
    rng = random.Random(seed)
    ranks = [Rank(r, n, epochs, obey_rule) for r in range(n)]
    steps = 0
    while not all(k.done() for k in ranks):
        steps += 1
        if steps > max_steps:
            return "hang", ranks
        k = rng.choice([k for k in ranks if not k.done()])
        op, e, x = k.prog[k.pc]
        if op == "store":
            ranks[x].inbox[e & 1][k.r] = (e, value(k.r, e))     # one atomic tagged word into the destination's inbox
            k.pc += 1
        else:
            tag, val = k.inbox[e & 1][x]
            if tag == e:
                k.acc.setdefault(e, []).append(val)
                k.pc += 1
                if len(k.acc[e]) == n:
                    k.sums[e] = sum(k.acc[e])                   # sources polled in rank order
            elif tag > e:
                return "overwritten", ranks                     # the real kernel would poll for ever (bounded: NaN)
            # tag < e: not there yet, poll again later
    return "ok", ranks


@pytest.mark.parametrize("n", [2, 3, 8])
def test_two_parities_suffice_under_the_rule(n):
    for seed in range(60):
        status, ranks = run(n, epochs=7, seed=seed)
        assert status == "ok", (n, seed, status)
        for e in range(1, 8):
            want = sum(value(r, e) for r in range(n))
            assert all(k.sums[e] == want for k in ranks)


def test_without_the_rule_a_slot_can_be_overwritten():
    """A rank that posts exchange e without having collected e-1 lets a peer's words of e+1 land on words of e-1 it has
    not read yet: some interleaving loses an exchange."""
    outcomes = {run(3, epochs=5, seed=s, obey_rule=False)[0] for s in range(200)}
    assert "overwritten" in outcomes
