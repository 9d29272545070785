"""Host logic of bench.py's pipelined graphs (configs 1 / 2): graph b finds the moments of batch b already computed -- by graph
b - 1's side branch -- unless the run of consecutive steps starts at b or another leg overwrote the buffers; then, and only then,
an eager DSNT pass primes them.  No GPU: the engine and the graphs are stand-ins that record the calls."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench


class _Graph:
    def __init__(self, log, b):
        self.log, self.b = log, b

    def replay(self):
        self.log.append(("replay", self.b))


class _Engine:
    def __init__(self, log):
        self.log = log

    def dsnt_into(self, ws, logits, mu, cov):
        self.log.append(("prime", logits, mu))

    def run_into(self, ws, logits, seed=0):
        self.log.append(("eager", logits, seed))


def _cfg(rank, pipelined=True, graphs=True):
    log = []
    c = bench.LogitsPSM.__new__(bench.LogitsPSM)
    c.rank, c.ws, c.eng = rank, object(), _Engine(log)
    c.pool = [f"batch{b}" for b in range(c.POOL)]
    c.moments = [("mu0", "cov0"), ("mu1", "cov1")]
    c.gs = [_Graph(log, b) for b in range(c.POOL)] if graphs else None
    c.pipelined, c.primed_for = pipelined, None
    return c, log


def test_consecutive_steps_prime_once():
    c, log = _cfg(rank=3)
    for i in range(2 * c.POOL + 3):
        c.step(i)
    assert log[0] == ("prime", "batch3", "mu1")                       # batch (0 + rank) % POOL into the pair of its parity
    assert [e for e in log if e[0] == "prime"] == [log[0]]
    assert [e[1] for e in log if e[0] == "replay"] == [(i + 3) % c.POOL for i in range(2 * c.POOL + 3)]


def test_a_break_in_the_sequence_primes_again():
    c, log = _cfg(rank=0)
    c.step(0); c.step(1)
    c.primed_for = None                                               # what step_evented does: it overwrites ws.mu / ws.cov
    c.step(2)
    c.step(3)
    c.step(9)                                                          # not the batch graph 3 prefetched
    primes = [e for e in log if e[0] == "prime"]
    assert primes == [("prime", "batch0", "mu0"), ("prime", "batch2", "mu0"), ("prime", "batch1", "mu1")]
    assert log[-1] == ("replay", 1) and c.primed_for == 2


def test_plain_graphs_and_eager_steps_never_prime():
    c, log = _cfg(rank=0, pipelined=False)
    for i in range(5):
        c.step(i)
    assert [e[0] for e in log] == ["replay"] * 5
    c, log = _cfg(rank=1, graphs=False)
    c.step(4)
    assert log == [("eager", "batch5", 1004)]


def test_pool_is_even_so_the_moment_pairs_alternate_across_the_wrap():
    assert bench.LogitsPSM.POOL % 2 == 0
