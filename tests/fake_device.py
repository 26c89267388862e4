"""Stand-ins for the device in tests of the HOST logic of the search (no GPU): an engine whose "rollouts" are a
deterministic function of the leaf, and a simulator that only carries the time axis.  Used by
tests/test_host_forest.py, tests/test_forest_gloo.py and oracle/gen_golden.py (the master-policy fixture)."""
import random
import zlib

import numpy as np

from iboat_pmcts_b200 import _native as N
from iboat_pmcts_b200 import forest as ft


class FakeEngine:
    """Stands in for engine.Engine: rollout_forest files `replicas` rewards per leaf into bins chosen from
    the leaf's first action, its depth and the scenario (so the trees have something to prefer)."""
    precision = N.PREC_F64

    def __init__(self):
        self.calls = []

    def set_precision(self, mode):
        self.precision = mode

    def rollout_forest(self, scens, n_leaf, replicas, root, dest, tmin, prefix, prefix_len, prefix_stride,
                       prefix_shared=True, z=None, seed=None, stream=0, id0=0, out=None, flags=0):
        S = len(scens)
        pre = np.asarray(prefix, np.float64).reshape(S, n_leaf, prefix_stride)
        plen = np.asarray(prefix_len).reshape(S, n_leaf)
        bins = out["leaf_bins"].reshape(S, n_leaf, 12)
        bins[:] = 0
        for s in range(S):
            for l in range(n_leaf):
                path = tuple(int(a) for a in pre[s, l, :plen[s, l]])
                h = zlib.crc32(repr((int(scens[s]), path, int(seed), int(stream) + s)).encode())
                centre = 2 + (path[0] // 45 + 3 * int(scens[s])) % 8
                rng = np.random.default_rng(h)
                draws = np.clip(np.rint(rng.normal(centre, 1.2, replicas)), 0, 11).astype(int)
                bins[s, l] = np.bincount(draws, minlength=12)
        self.calls.append((S, n_leaf, replicas, int(stream)))


class FakeSim:
    times = np.arange(0, 3 + 6 / 24, 6 * (1 / 24))
    _engine = None

    def __init__(self, slot):
        self.slot = slot

    def scenario(self):
        return FakeSim._engine, self.slot

    def __deepcopy__(self, memo):
        return self


def run(host, n_scen=3, budget=150, leaves=16, replicas=24, seed=4, group=False, pipeline=1):
    FakeSim._engine = FakeEngine()
    random.seed(10)
    np.random.seed(10)
    forest = ft.Forest(listsimulators=[FakeSim(i) for i in range(n_scen)], destination=[46.7, 355.0], timemin=2.0,
                       budget=budget)
    master = forest.launch_search([0, 44.0, 355.0], 10, mode="batched", leaves=leaves, replicas=replicas, seed=seed,
                                  host=host, group=group, pipeline=pipeline)
    return forest, master, FakeSim._engine.calls
