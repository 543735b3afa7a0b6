"""Host model of the peer-memory exchange protocol of laghos_b200/csrc/device/p2p.cuh (no GPU).

Every rank runs the same stream of exchanges k = 1, 2, ...; exchange k = publish (write my value into every
peer's slot [k & 1][my rank], then raise my flag there to k) followed by consume (wait until every peer's flag in MY
buffer is >= k, then read the slots of parity k & 1).  Slots and flags are double-buffered by the parity of k only.
The model interleaves the ranks' steps at random and checks the claim the kernels rely on: when a rank consumes
exchange k, every slot it reads still holds the value of exchange k (a peer can be at most one exchange ahead, so it
never overwrites a slot that has not been read), and nobody deadlocks."""
import random

import pytest


class Rank:
    def __init__(self, r, n):
        self.r, self.n = r, n
        self.slot = [[None] * n, [None] * n]      # [parity][source rank] -> (k, value)
        self.flag = [[0] * n, [0] * n]            # [parity][source rank] -> sequence number
        self.k = 1                                # exchange in progress
        self.stage = "publish"
        self.results = []


def step(ranks, i, value_of):
    me = ranks[i]
    par = me.k & 1
    if me.stage == "publish":
        for peer in ranks:                        # data first, then the flag (release ordering)
            peer.slot[par][me.r] = (me.k, value_of(me.r, me.k))
        for peer in ranks:
            peer.flag[par][me.r] = me.k
        me.stage = "consume"
        return True
    if all(me.flag[par][s] >= me.k for s in range(me.n)):
        vals = []
        for s in range(me.n):
            kk, v = me.slot[par][s]
            assert kk == me.k, f"rank {me.r} exchange {me.k}: slot of rank {s} holds exchange {kk}"
            vals.append(v)
        me.results.append(sum(vals))              # ascending rank order: identical on every rank
        me.k += 1
        me.stage = "publish"
        return True
    return False                                  # spinning on a flag


@pytest.mark.parametrize("nranks", [2, 4, 8])
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_parity_double_buffering_is_enough(nranks, seed):
    rng = random.Random(seed)
    ranks = [Rank(r, nranks) for r in range(nranks)]
    nexch = 40

    def value_of(r, k):
        return (r + 1) * 1000 + k

    idle = 0
    while any(rk.k <= nexch for rk in ranks):
        # adversarial scheduling: mostly favour one rank so that it runs as far ahead as the protocol allows
        fav = rng.randrange(nranks)
        i = fav if rng.random() < 0.7 else rng.randrange(nranks)
        if ranks[i].k > nexch:
            continue
        progressed = step(ranks, i, value_of)
        idle = 0 if progressed else idle + 1
        assert idle < 10000 * nranks, "deadlock"
        lead = max(rk.k for rk in ranks) - min(rk.k for rk in ranks)
        assert lead <= 1, "a rank ran more than one exchange ahead of a peer"
    expect = [sum(value_of(r, k) for r in range(nranks)) for k in range(1, nexch + 1)]
    for rk in ranks:
        assert rk.results == expect
