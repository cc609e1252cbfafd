"""Host logic of the probe source (no GPU): the default probes are the reference's draws, src/sgvamp.py:326 -
np.random.binomial(p=1/2, n=1, size=M) * 2 - 1 from numpy's legacy global RNG, one call per cohort and iteration in
loop order; with n_probes > 1 the further probes of a (cohort, iteration) follow its first one immediately."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sgvamp-py_b200"))


def _order(iterations, mine, n_probes):
    return [key for it in range(iterations) for k in mine for key in [(it, k)] + [(it, k, p) for p in range(1, n_probes)]]


def test_probe_source_draws_in_reference_order():
    import sgvamp
    M = 257
    for n_probes, mine in [(1, [0]), (1, [0, 1, 2]), (3, [0, 1]), (2, [1])]:
        order = _order(4, mine, n_probes)
        np.random.seed(99)
        src = sgvamp._ProbeSource(order, M)
        got = [src.get(key) for key in order]
        np.random.seed(99)
        exp = [(np.random.binomial(p=1 / 2, n=1, size=M) * 2 - 1).astype(np.int8) for _ in order]
        assert all(np.array_equal(a, b) for a, b in zip(got, exp))
        assert all(set(np.unique(u)) <= {-1, 1} and u.dtype == np.int8 for u in got)
        # the state recorded before the first draw of an iteration restarts the sequence there (checkpoint / resume)
        np.random.set_state(src.states[(2, mine[0])])
        again = (np.random.binomial(p=1 / 2, n=1, size=M) * 2 - 1).astype(np.int8)
        assert np.array_equal(again, got[order.index((2, mine[0]))])


def test_cli_seeded_probes_are_indexed_by_iteration_and_probe():
    import main as cli
    M = 64
    one = cli._SeededProbes(7, 2)
    multi = cli._SeededProbes(7, 2, n_probes=3)
    rs = np.random.RandomState(7 + 1)
    draws = [(rs.binomial(p=1 / 2, n=1, size=M) * 2 - 1).astype(np.int8) for _ in range(9)]
    assert np.array_equal(one(1, 2, M), draws[2]) and np.array_equal(one(1, 0, M), draws[0])
    assert np.array_equal(multi(1, 1, M, 2), draws[5]) and np.array_equal(multi(1, 0, M), draws[0])
    assert np.array_equal(multi(1, 2, M, 1), draws[7])
