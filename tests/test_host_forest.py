"""The C++ host runtime of the batched search (csrc/pmcts_forest_host.cpp, pmcts_forest_*) against the Python
classes it restates (worker.Tree.select_leaves / back_up_leaves, forest._master_add): same leaves, wave after
wave, same counters in every worker-tree node and every master node, same best policy.  No GPU: the rollout
launch is replaced by a deterministic stand-in that fills the per-leaf histograms from the leaf's path."""
import numpy as np
import pytest

from iboat_pmcts_b200 import _native as N
from iboat_pmcts_b200 import forest as ft
from iboat_pmcts_b200 import master_node as mn
from iboat_pmcts_b200 import worker


from fake_device import FakeEngine, FakeSim, run  # noqa: E402,F401


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
    # (the decision taken from it -- MasterTree, reduced on the device -- is compared in tests/test_master_policy_gpu.py)
    # deepcopy_dict of the array-backed dictionary: same nodes, wired, the source emptied -- without building a
    # node object until one is read
    copy = mn.deepcopy_dict(mn_)
    assert len(mn_) == 0 and len(copy) == len(mp) and not copy._filled
    root = copy[ft.ROOT_HASH]
    assert root.depth == 0 and len(root.children) == 8 and all(c.parentNode is root and c.depth == 1 for c in root.children)
    assert list(copy) == list(mp) and all(np.array_equal(copy[k]._table, mp[k]._table) for k in mp)


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
    # the batched form (one generator per worker tree, filled on the runtime's threads) == one call per seed
    seeds = np.asarray([1000003 * 4 + i for i in range(23)] + [2 ** 40 + 3], np.uint64)
    batch = np.empty((seeds.size, 322, 8), np.uint8)
    N.check_forest(L.pmcts_python_shuffles_batch(seeds.ctypes.data, seeds.size, 322, 8, batch.ctypes.data))
    for i, seed in enumerate(seeds):
        one = np.empty((322, 8), np.uint8)
        N.check_forest(L.pmcts_python_shuffles(int(seed), 322, 8, one.ctypes.data))
        assert np.array_equal(batch[i], one), int(seed)


def test_native_generators_are_the_interpreters():
    """pmcts_python_gauss / pmcts_numpy_randint8 continue random.getstate() / np.random.get_state() exactly as
    random.gauss(0, 1) / np.random.randint(8) do -- values bit for bit, states afterwards included (odd counts leave
    a cached second normal behind) -- which is what lets the native sequential search grow the reference's tree."""
    import random
    L = N.lib()
    for seed, count in ((1, 7), (12345, 1000), (2 ** 40 + 3, 625 * 2 + 1)):
        random.seed(seed)
        random.random()
        if seed == 12345:
            random.gauss(0.0, 1.0)  # start with a cached normal
        st = N.MtState.from_python(random.getstate())
        out = np.empty(count)
        N.check_forest(L.pmcts_python_gauss(st, count, out.ctypes.data))
        want = np.array([random.gauss(0.0, 1.0) for _ in range(count)])
        assert np.array_equal(out, want)
        assert st.to_python() == random.getstate()
        np.random.seed(seed % (2 ** 32))
        np.random.randint(100, size=3)
        like = np.random.get_state()
        ns = N.MtState.from_numpy(like)
        got = np.empty(count, np.uint8)
        N.check_forest(L.pmcts_numpy_randint8(ns, count, got.ctypes.data))
        assert np.array_equal(got, np.array([np.random.randint(8) for _ in range(count)], np.uint8))
        back, now = ns.to_numpy(like), np.random.get_state()
        assert np.array_equal(back[1], now[1]) and back[2] == now[2]


def _master_digest(master):
    import hashlib
    h = hashlib.sha256()
    for k in master:
        h.update(str(k).encode())
        h.update(np.ascontiguousarray(master[k]._table).tobytes())
    return h.hexdigest()


def test_search_does_not_depend_on_the_host_threads():
    """The host runtime's pool (pmcts_forest_host.cpp: tasks claimed from one counter, workers that poll before they
    sleep) only decides WHO walks a tree or files a scenario's block: one thread, three threads that sleep at once and
    the default pool return the same master, search after search, with one and with two waves in flight."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import fake_device as fd, test_host_forest as t\n"
            "for rep in range(2):\n"
            "    for pipe in (1, 2):\n"
            "        f, m, _ = fd.run('native', n_scen=10, budget=65, leaves=32, replicas=4, seed=6, pipeline=pipe)\n"
            "        print(pipe, t._master_digest(m))\n") % (os.path.dirname(here), here)
    outs = []
    for env_extra in ({}, {"PMCTS_HOST_THREADS": "1"}, {"PMCTS_HOST_THREADS": "3", "PMCTS_POOL_SPIN_US": "0"},
                      {"PMCTS_HOST_THREADS": "12", "PMCTS_POOL_SPIN_US": "400"}):
        env = dict(os.environ, **env_extra)
        res = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                             text=True, timeout=300)
        assert res.returncode == 0, res.stderr[-2000:]
        outs.append(res.stdout.split())
    assert len(outs[0]) == 8 and len(set(outs[0][1::4])) == 1 and len(set(outs[0][3::4])) == 1
    for other in outs[1:]:
        assert other == outs[0]


@pytest.mark.parametrize("n_scen,budget,leaves,replicas,seed", [
    (1, 40, 64, 8, 1),       # one tree, the wave larger than the budget: a single short wave
    (2, 9, 1, 3, 2),         # one leaf per wave: the sequential order, wave by wave
    (5, 77, 10, 300, 3),     # counts above 255: the 16-bit update block; a ragged last wave
    (4, 128, 32, 255, 4),    # the largest count one byte holds
    (12, 33, 11, 2, 5),      # more trees than a typical pool has threads
])
def test_native_forest_equals_python_forest_over_shapes(n_scen, budget, leaves, replicas, seed):
    """The equality of the first test over the corner shapes of a search: same master (keys, order, counters), same
    worker trees node by node, same launches."""
    fp, mp, calls_p = run("python", n_scen=n_scen, budget=budget, leaves=leaves, replicas=replicas, seed=seed)
    fn, mn_, calls_n = run("native", n_scen=n_scen, budget=budget, leaves=leaves, replicas=replicas, seed=seed)
    assert calls_p == calls_n and len(calls_n) == -(-budget // leaves)
    assert list(mp) == list(mn_)
    for k in mp:
        assert np.array_equal(mp[k]._table, mn_[k]._table), k
    for i in range(n_scen):
        tp, tn = fp.workers[i], fn.workers[i]
        assert len(tp.Nodes) == len(tn.Nodes) == budget + 1
        for a, b in zip(tp.Nodes, tn.Nodes):
            assert [int(x) for x in a.origins] == [int(x) for x in b.origins]
            assert np.array_equal(a._table, b._table)
        assert int(tn.rootNode.visits()) == budget * replicas
    assert int(mn_[ft.ROOT_HASH]._table.sum()) == n_scen * budget * replicas


def test_forest_stays_searchable_after_a_failed_search():
    """A search that fails between a selection and its back-up (here: the rollout stand-in refuses the third wave, with
    one and with two waves in flight) takes back the virtual visits of the waves it had selected and drops them: the
    same forest handle is searched again, files every rollout of the completed waves exactly once, and has no
    selection left waiting for a back-up."""
    import ctypes as C
    lib = N.lib()
    S, B, K, D = 3, 8, 5, 12
    for pipeline in (1, 2):
        h = C.c_void_p()
        ids = np.arange(S, dtype=np.int32)
        prob = np.full(S, 1.0 / S)
        max_nodes = 200
        perms = np.empty((S, max_nodes, 8), np.uint8)
        seeds = np.arange(1, S + 1, dtype=np.uint64)
        N.check_forest(lib.pmcts_python_shuffles_batch(seeds.ctypes.data, S, max_nodes, 8, perms.ctypes.data))
        N.check_forest(lib.pmcts_forest_create(S, S, ids.ctypes.data, D, prob.ctypes.data, 0.1, 0.3, perms.ctypes.data,
                                               max_nodes, C.byref(h)))
        try:
            state = dict(fail_at=2, waves=0)

            def hook(_user, wave, n_local, _ids, b, stride, plen_p, path_p, bins_p):
                if wave == state["fail_at"]:
                    return 7
                bins = np.ctypeslib.as_array(C.cast(bins_p, C.POINTER(C.c_uint32)), shape=(n_local * b, 12))
                bins[:] = 0
                bins[:, 3 + wave % 4] = K
                state["waves"] += 1
                return 0

            cb = N.ROLLOUT_HOOK_FN(hook)
            slots = np.zeros(S * 10 * B, np.int8)

            def search(budget):
                p = N.SearchParams()
                p.n_leaves, p.replicas, p.budget, p.seed, p.pipeline, p.world, p.rank = B, K, budget, 1, pipeline, 1, 0
                p.slots_w, p.slots_m, p.rollout_hook = slots.ctypes.data, slots.ctypes.data, cb
                rep = N.SearchReport()
                root, dest = np.array([0, 44.0, 355.0]), np.array([46.7, 355.0])
                return lib.pmcts_forest_search(None, h, None, root.ctypes.data, dest.ctypes.data, 2.0, C.byref(p),
                                               C.byref(rep)), rep.as_dict()

            rc, rep = search(5 * B)
            assert rc != 0 and "rollout_hook returned 7" in lib.pmcts_last_error().decode()
            filed = rep["waves"]
            assert filed == (2 if pipeline == 1 else 1)   # (two in flight: wave 1 was out when wave 2 was refused)
            state["fail_at"] = -1
            rc, rep2 = search(4 * B)
            assert rc == 0 and rep2["waves"] == 4
            # every completed wave is in the trees and in the master exactly once; nothing waits for a back-up
            for i in range(S):
                n = int(lib.pmcts_forest_tree_size(h, i))
                table = np.empty((n, 8, 12), np.int64)
                N.check_forest(lib.pmcts_forest_export_tree(h, i, None, None, table.ctypes.data, None, None))
                assert int(table[0].sum()) == (filed + 4) * B * K
            leaf_bins, sl = np.zeros((S * B, 12), np.uint32), np.zeros(S * B, np.int8)
            assert lib.pmcts_forest_backup(h, B, leaf_bins.ctypes.data, sl.ctypes.data) != 0
            assert "oldest selection" in lib.pmcts_forest_last_error().decode()
        finally:
            lib.pmcts_forest_destroy(h)


def test_native_sequential_search_replays_the_recorded_reference_run(golden_dir):
    """pmcts::forest_sequential -- the native Tree.uct_search (worker.py:147-253, 348-378) -- against the run recorded
    from the UNMODIFIED reference (tests/golden/tree_replay_c1.json: random.seed(1), np.random.seed(1), 300 iterations,
    master update every 10), on the CPU: the rollouts are handed to a callback that returns what the reference's
    rollout of that iteration returned (the reward's bin; the transitions made, tree_replay_c1_steps.json), everything
    else -- tree policy with the master's share, shuffles of the untried actions at node creation and the normals of
    the transitions from Python's generator, action slots from NumPy's -- runs natively.  Every iteration must pick the
    reference's leaf; all node histograms and the master's must be the reference's; and the two generators must end
    where the interpreter's end after the same draws."""
    import ctypes as C
    import json
    import os
    import random
    from iboat_pmcts_b200.utils import Hist
    with open(os.path.join(golden_dir, "tree_replay_c1.json")) as f:
        rec = json.load(f)
    with open(os.path.join(golden_dir, "tree_replay_c1_steps.json")) as f:
        steps = json.load(f)["steps"]
    its = rec["iterations"]
    lib = N.lib()
    h = C.c_void_p()
    ids, prob = np.zeros(1, np.int32), np.ones(1)
    N.check_forest(lib.pmcts_forest_create(1, 1, ids.ctypes.data, 12, prob.ctypes.data, rec["uct_coeff"], rec["rho"], None,
                                           len(its) + 8, C.byref(h)))
    try:
        random.seed(1)
        np.random.seed(1)
        py, npst = N.MtState.from_python(random.getstate()), N.MtState.from_numpy(np.random.get_state())
        seen = []

        def hook(_user, iteration, scen, path, length, z, n_instants, bin_p, steps_p):
            seen.append([45 * int(path[j]) for j in range(length)])
            if seen[-1] != its[iteration]["origins"] or n_instants != 13 or scen != 0:
                return 1
            bin_p[0] = Hist.bin_of(its[iteration]["reward"])
            steps_p[0] = steps[iteration]
            return 0

        cb = N.SEQ_ROLLOUT_HOOK_FN(hook)
        rep = N.SearchReport()
        rc = lib.pmcts_forest_search_sequential_hook(h, 0, rec["budget"], rec["frequency"], C.byref(py), C.byref(npst), cb,
                                                     None, C.byref(rep))
        assert rc == 0, (lib.pmcts_forest_last_error().decode(), len(seen), seen[-1:], its[len(seen) - 1]["origins"])
        assert seen == [it["origins"] for it in its]
        assert rep.rollouts == 300 and rep.transitions == sum(steps) and rep.waves == 30
        # the worker tree: nodes in creation order, their action paths and 8 x 12 histograms
        n = int(lib.pmcts_forest_tree_size(h, 0))
        assert n == len(rec["nodes"]) == 301
        parent, arm, table = np.empty(n, np.int32), np.empty(n, np.int8), np.empty((n, 8, 12), np.int64)
        N.check_forest(lib.pmcts_forest_export_tree(h, 0, parent.ctypes.data, arm.ctypes.data, table.ctypes.data, None, None))
        paths = []
        for i in range(n):
            paths.append([] if parent[i] < 0 else paths[parent[i]] + [45 * int(arm[i])])
            assert paths[i] == rec["nodes"][i]["origins"], i
            assert np.array_equal(table[i], np.array(rec["nodes"][i]["hist"])), i
        # the master: same nodes (by the hash of their path), same histograms
        m = int(lib.pmcts_forest_master_size(h))
        mparent, marm, mtable = np.empty(m, np.int32), np.empty(m, np.int8), np.empty((m, 1, 8, 12), np.uint32)
        N.check_forest(lib.pmcts_forest_export_master(h, mparent.ctypes.data, marm.ctypes.data, mtable.ctypes.data))
        assert m == len(rec["master_nodes"])
        mpaths = []
        for i in range(m):
            mpaths.append(() if mparent[i] < 0 else mpaths[mparent[i]] + (45 * int(marm[i]),))
            assert np.array_equal(mtable[i, 0], np.array(rec["master_nodes"][str(hash(mpaths[i]))])), mpaths[i]
        # the generators: where the interpreter's are after the same draws (301 shuffles interleaved with the normals of
        # 3017 transitions; 300 + 300 action slots)
        slots_drawn = 300 + 300
        np.random.seed(1)
        np.random.randint(8, size=slots_drawn)
        assert npst.to_numpy(np.random.get_state())[2] == np.random.get_state()[2]
        assert np.array_equal(npst.to_numpy(np.random.get_state())[1], np.random.get_state()[1])
        random.seed(1)
        random.shuffle(list(range(8)))             # the root's untried actions (worker.py:51)
        for k in steps:
            random.shuffle(list(range(8)))         # the node the iteration expands
            for _ in range(k):
                random.gauss(0.0, 1.0)             # one normal per transition (simulatorTLKT.py:515)
        assert py.to_python() == random.getstate()
    finally:
        lib.pmcts_forest_destroy(h)
