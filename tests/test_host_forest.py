"""The C++ host runtime of the batched search (csrc/pmcts_forest_host.cpp, pmcts_forest_*) against the Python
classes it restates (worker.Tree.select_leaves / back_up_leaves, forest._master_add): same leaves, wave after
wave, same counters in every worker-tree node and every master node, same best policy.  No GPU: the rollout
launch is replaced by a deterministic stand-in that fills the per-leaf histograms from the leaf's path."""
import random
import zlib

import numpy as np
import pytest

from iboat_pmcts_b200 import _native as N
from iboat_pmcts_b200 import forest as ft
from iboat_pmcts_b200 import master_node as mn
from iboat_pmcts_b200 import worker
from iboat_pmcts_b200.master import MasterTree


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


def test_native_forest_equals_python_forest():
    fp, mp, calls_p = run("python")
    fn, mn_, calls_n = run("native")
    assert calls_p == calls_n and len(calls_p) == 10            # 9 waves of 16 + one of 6
    # master dictionary: same keys in the same order, same counters, same wiring
    assert list(mp) == list(mn_)
    for k in mp:
        a, b = mp[k], mn_[k]
        assert np.array_equal(a._table, b._table), k
        assert a.arm == b.arm and (a.parentNode.hash if a.parentNode else None) == (b.parentNode.hash if b.parentNode else None)
    # worker trees: same nodes in the same order, same counters, same untried actions
    for i in range(3):
        tp, tn = fp.workers[i], fn.workers[i]
        assert len(tp.Nodes) == len(tn.Nodes) == 151 and tp.ite == tn.ite == 150 and tp.depth == tn.depth
        for a, b in zip(tp.Nodes, tn.Nodes):
            assert [int(x) for x in a.origins] == [int(x) for x in b.origins]
            assert np.array_equal(a._table, b._table)
            assert [int(x) for x in a.actions] == [int(x) for x in b.actions]
            assert len(a.children) == len(b.children) and a.depth == b.depth
        assert int(tn.rootNode.visits()) == 150 * 24
        assert np.array_equal(np.array([h.h for h in tn.Nodes[5].Values]), tn.table[5])
    # and the decision taken from it
    pols = []
    for forest, master in ((fp, mp), (fn, mn_)):
        mt = MasterTree([w.Simulator for w in forest.workers.values()], [46.7, 355.0], nodes=mn.deepcopy_dict(master))
        mt.get_best_policy(verbose=False)
        pols.append({k: [int(a) for a in v] for k, v in mt.best_policy.items()})
    assert pols[0] == pols[1] and len(pols[0][-1]) >= 1


def test_two_waves_in_flight():
    """pipeline=2: wave w + 1 is selected while wave w is out (its leaves hold virtual visits).  Every rollout is
    still filed exactly once, the trees have one node per iteration, the search is reproducible, and it is not the
    synchronous search."""
    f1, m1, calls1 = run("native", pipeline=1)
    f2, m2, calls2 = run("native", pipeline=2)
    f3, m3, _ = run("native", pipeline=2)
    assert [c[:3] for c in calls1] == [c[:3] for c in calls2] and len(calls2) == 10
    assert list(m2) == list(m3) and all(np.array_equal(m2[k]._table, m3[k]._table) for k in m2)
    assert list(m1) != list(m2)
    root = m2[ft.ROOT_HASH]._table
    assert [int(root[s].sum()) for s in range(3)] == [150 * 24] * 3
    for i in range(3):
        t = f2.workers[i]
        assert len(t.Nodes) == 151 and int(t.rootNode.visits()) == 150 * 24
        # a node's counters are the sum of its children's plus what was filed at the node itself (as a leaf)
        for node in t.Nodes:
            below = sum(int(c._table.sum()) for c in node.children)
            assert int(node._table.sum()) >= below
    assert f2.last_search_report["waves"] == 10


def test_native_forest_reads_the_module_coefficients(monkeypatch):
    base = run("native", budget=64, leaves=8)[1]
    monkeypatch.setattr(worker, "UCT_COEFF", 5.0)   # much more exploration: another tree
    monkeypatch.setattr(worker, "RHO", 0.9)
    other = run("native", budget=64, leaves=8)[1]
    check = run("python", budget=64, leaves=8)[1]
    assert list(other) == list(check) and all(np.array_equal(other[k]._table, check[k]._table) for k in other)
    assert list(other) != list(base)


def test_native_forest_errors():
    import ctypes as C
    lib = N.lib()
    h = C.c_void_p()
    ids = np.zeros(1, np.int32)
    prob = np.ones(1)
    perms = np.tile(np.arange(8, dtype=np.uint8), (4, 1))
    N.check_forest(lib.pmcts_forest_create(1, 1, ids.ctypes.data, 12, prob.ctypes.data, 0.1, 0.3, perms.ctypes.data, 4,
                                           C.byref(h)))
    try:
        leaf, plen, path = np.empty(8, np.int32), np.empty(8, np.int32), np.empty((8, 32), np.int8)
        with pytest.raises(N.PmctsError):   # 8 leaves need 9 nodes, capacity is 4
            N.check_forest(lib.pmcts_forest_select(h, 8, 32, leaf.ctypes.data, plen.ctypes.data, path.ctypes.data))
        with pytest.raises(N.PmctsError):   # path buffer shorter than the deepest possible leaf
            N.check_forest(lib.pmcts_forest_select(h, 1, 4, leaf.ctypes.data, plen.ctypes.data, path.ctypes.data))
    finally:
        lib.pmcts_forest_destroy(h)


def test_native_shuffles_are_cpythons():
    """pmcts_python_shuffles == random.Random(seed).shuffle(list(range(n))) repeated, for seeds of one and two
    32-bit limbs: the native search pops the same untried actions as worker.Node.__init__ (worker.py:50-51)."""
    import random
    from iboat_pmcts_b200 import _native as N
    L = N.lib()
    for seed, n_items in ((0, 8), (5, 8), (1000003 * 3 + 7, 8), (2 ** 32 + 17, 8), (2 ** 63 + 5, 8), (99, 13)):
        out = np.empty((500, n_items), np.uint8)
        N.check_forest(L.pmcts_python_shuffles(seed, 500, n_items, out.ctypes.data))
        rng = random.Random(seed)
        ref = []
        for _ in range(500):
            order = list(range(n_items))
            rng.shuffle(order)
            ref.append(order)
        assert np.array_equal(out, np.array(ref, np.uint8)), seed
