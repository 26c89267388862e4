"""N4: a search result written by MasterTree.save_tree(reference_format=True) is read by the UNMODIFIED reference's own
MasterTree.load_tree, whose own post-processing then runs on it (code/solver/master.py:57-185, 390-410).  Live test:
runs where the reference is mounted (/root/reference: the build container).  No GPU: the tree comes from the native
search driven by the host-logic stand-in of tests/fake_device.py."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from fake_device import run
from iboat_pmcts_b200 import master_node as mn
from iboat_pmcts_b200.master import MasterTree
from oracle import refshim

pytestmark = pytest.mark.skipif(not refshim.reference_available(), reason="reference not mounted")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_loads_the_tree_and_decides_on_it(tmp_path):
    forest, master, _ = run("native", n_scen=4, budget=90)
    nodes = mn.deepcopy_dict(master)
    from iboat_pmcts_b200 import synth
    from iboat_pmcts_b200.forest import create_simulators
    from iboat_pmcts_b200.weatherTLKT import Weather
    weathers = [Weather(*[synth.wind_fields(s)[i] for i in (1, 2, 0, 3, 4)]) for s in range(4)]
    sims = create_simulators(weathers, 4, simtimestep=6, stateinit=(0, 44.0, 355.0), ndaysim=3)  # (no device needed)
    mt = MasterTree(sims, [46.7, 355.0], nodes=nodes)
    mt.save_tree("tree", directory=str(tmp_path) + "/", reference_format=True)
    assert not nodes._filled  # written from the arrays: no node object of the mirror classes was built
    parent, arm, table = nodes.arrays()
    np.savez(tmp_path / "want.npz", parent=parent, arm=arm, table=table)
    script = textwrap.dedent("""
        import io, contextlib, pickle, sys
        import numpy as np
        sys.path.insert(0, %r)
        from oracle import refshim
        R = refshim.load_reference()
        with open(%r, "rb") as f:
            tree = pickle.load(f)                      # what MasterTree.load_tree does (master.py:400-410)
        assert type(tree) is R.master.MasterTree and type(tree.nodes[hash(())]) is R.master_node.MasterNode
        want = np.load(%r)
        # same nodes, same histograms, wired like deepcopy_dict's result
        paths = {hash(()): ()}
        root = tree.nodes[hash(())]
        assert root.depth == 0 and root.parentNode is None
        seen = 0
        stack = [(root, 0)]
        index = {}
        parent, arm, table = want["parent"], want["arm"], want["table"]
        kids = {}
        for i in range(1, parent.size):
            kids.setdefault(int(parent[i]), []).append(i)
        while stack:
            node, i = stack.pop()
            seen += 1
            got = np.array([[h.h for h in row] for row in node.rewards])
            assert np.array_equal(got, table[i]), i
            assert type(node.rewards[0, 0]) is R.utils.Hist
            assert [int(c.arm) for c in node.children] == [45 * int(arm[j]) for j in kids.get(i, [])]
            for c, j in zip(node.children, kids.get(i, [])):
                assert c.parentNode is node and c.depth == node.depth + 1
                stack.append((c, j))
        assert seen == parent.size == len(tree.nodes)
        sim = tree.Simulators[2]                       # a reference Simulator with working SciPy interpolators
        assert type(sim) is R.simulatorTLKT.Simulator and len(sim.times) == 13
        assert np.isfinite(sim.uWindAvg([0.3, 44.2, 355.3])).all()
        with contextlib.redirect_stdout(io.StringIO()):
            tree.get_best_policy()                     # the reference's own reductions on the loaded tree
        print({k: [int(a) for a in v] for k, v in tree.best_policy.items()})
        """) % (ROOT, str(tmp_path / "tree.pickle"), str(tmp_path / "want.npz"))
    res = subprocess.run([sys.executable, "-W", "ignore", "-c", script], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                         text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    policy = eval(res.stdout.strip().splitlines()[-1])
    assert set(policy) == {-1, 0, 1, 2, 3} and all(len(v) >= 1 for v in policy.values())


def test_tree_saved_by_the_reference_loads_into_this_package(tmp_path):
    """The other direction: a seeded search of the UNMODIFIED reference (Tree.uct_search on two scenarios into one
    master dictionary, deepcopy_dict, MasterTree, save_tree) is read by this package's MasterTree.load_tree -- no
    reference code imported on this side -- into its own classes: same nodes in the same order, same histograms,
    children wired in the reference's order, simulators whose weather grids are the reference's, the stored policies."""
    script = textwrap.dedent("""
        import contextlib, io, json, random, sys
        import numpy as np
        sys.path.insert(0, %r)
        from oracle import refshim, gen_golden as G
        R = refshim.load_reference()
        wrk, mnode, mstr = R.worker, R.master_node, R.master
        sims = G.make_sims(R, [0, 1])
        random.seed(3); np.random.seed(3)
        master = {hash(()): mnode.MasterNode(2, nodehash=hash(()))}
        with contextlib.redirect_stdout(io.StringIO()):
            for w in range(2):
                tree = wrk.Tree(workerid=w, nscenario=2, budget=60, simulator=sims[w], destination=[46.7, 355.0],
                                TimeMin=2.0, buffer=[])
                tree.uct_search([0, 44.0, 355.0], 10, master)
        class ProxyLike(dict):       # (Manager().dict().items() is a list copy; deepcopy_dict deletes while it iterates)
            def items(self):
                return list(dict.items(self))
        mt = mstr.MasterTree(sims, [46.7, 355.0], nodes=mnode.deepcopy_dict(ProxyLike(master)), proba=[0.25, 0.75])
        with contextlib.redirect_stdout(io.StringIO()):
            mt.get_best_policy()
        mt.save_tree("ref_tree")          # the reference's fixed place: ../results/
        out = dict(keys=[str(k) for k in mt.nodes],
                   hist={str(k): [[[int(x) for x in h.h] for h in row] for row in nd.rewards] for k, nd in mt.nodes.items()},
                   kids={str(k): [str(c.hash) for c in nd.children] for k, nd in mt.nodes.items()},
                   depth={str(k): nd.depth for k, nd in mt.nodes.items()},
                   arm={str(k): None if nd.arm is None else int(nd.arm) for k, nd in mt.nodes.items()},
                   policy={str(k): [int(a) for a in v] for k, v in mt.best_policy.items()},
                   wind=[float(sims[1].uWindAvg([0.3, 44.2, 355.3])[0]), float(sims[1].vWindAvg([0.3, 44.2, 355.3])[0])],
                   points=[list(map(float, p)) for p in mt.get_points(mt.nodes[hash(())], [], mt.probability, idscenario=1)])
        print(json.dumps(out))
        """) % ROOT
    (tmp_path / "code").mkdir()
    (tmp_path / "results").mkdir()
    res = subprocess.run([sys.executable, "-W", "ignore", "-c", script], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                         text=True, timeout=600, cwd=str(tmp_path / "code"))
    assert res.returncode == 0, res.stderr[-2000:]
    import json
    want = json.loads(res.stdout.strip().splitlines()[-1])
    from iboat_pmcts_b200 import master_node, simulatorTLKT, weatherTLKT
    mt = MasterTree.load_tree("ref_tree", directory=str(tmp_path / "results") + "/")
    assert type(mt) is MasterTree and mt.numScenarios == 2 and list(mt.probability) == [0.25, 0.75]
    assert [str(k) for k in mt.nodes] == want["keys"] and len(mt.nodes) > 60
    for k, nd in mt.nodes.items():
        assert type(nd) is master_node.MasterNode and nd.hash == k
        assert np.array_equal(nd._table, np.array(want["hist"][str(k)]))
        assert [str(c.hash) for c in nd.children] == want["kids"][str(k)]
        assert all(c.parentNode is nd for c in nd.children)
        assert nd.depth == want["depth"][str(k)] and (None if nd.arm is None else int(nd.arm)) == want["arm"][str(k)]
    assert {str(k): [int(a) for a in v] for k, v in mt.best_policy.items()} == want["policy"]
    assert all(mt.nodes[nd.hash] is nd for v in mt.best_nodes_policy.values() for nd in v)
    sim = mt.Simulators[1]
    assert type(sim) is simulatorTLKT.Simulator and type(sim._weather) is weatherTLKT.Weather and len(sim.times) == 13
    from iboat_pmcts_b200 import synth
    t, la, lo, u, v = synth.wind_fields(1)
    assert np.array_equal(sim._weather.u, u) and np.array_equal(sim._weather.lat, la) and np.array_equal(sim._weather.time, t)
    got = mt.get_points(mt.nodes[hash(())], [], mt.probability, idscenario=1)
    assert [list(map(float, p)) for p in got] == want["points"]
    # and a tree saved by this package in its own format still comes back through the same door
    mt.save_tree("own", directory=str(tmp_path) + "/")
    again = MasterTree.load_tree("own", directory=str(tmp_path) + "/")
    assert type(again) is MasterTree and [str(k) for k in again.nodes] == want["keys"]


def test_native_sequential_search_of_three_workers_is_the_references(tmp_path, golden_dir):
    """Tree.get_master_uct across scenarios (worker.py:302-346: scenarios that never visited a node skipped, the rest
    weighted by their probability without renormalising) and its share in Tree.best_child (worker.py:224-253), pinned to
    the reference itself: the UNMODIFIED reference searches three scenarios one after the other into ONE master
    dictionary (skewed probabilities, the Tutorial's RHO / UCT_COEFF, seeded generators), recording the leaf, reward
    and transitions of every iteration; the native sequential search then runs the same three searches on the CPU with
    the rollouts answered from the recording.  It must pick the reference's leaf at every one of the 3 x 120 iterations
    -- the later trees steer by what the earlier ones left in the master -- and end with the reference's trees, master
    and generator states."""
    import ctypes as C
    import json
    import random
    from iboat_pmcts_b200 import _native as N
    from iboat_pmcts_b200.utils import Hist
    g = np.load(os.path.join(golden_dir, "rollouts_c1.npz"))
    dest, tmin = [float(x) for x in g["destination"]], float(g["timemin"])
    budget, frequency, prob = int(os.environ.get("PMCTS_REPLAY_BUDGET", "120")), 10, [0.5, 0.2, 0.3]
    script = textwrap.dedent("""
        import contextlib, io, json, random, sys
        import numpy as np
        sys.path.insert(0, %r)
        from oracle import refshim, gen_golden as G
        R = refshim.load_reference()
        simc, wrk, mnode = R.simulatorTLKT, R.worker, R.master_node
        sims = G.make_sims(R, [0, 1, 2])
        wrk.RHO, wrk.UCT_COEFF = 0.5, 1 / 2 ** 0.5
        random.seed(4); np.random.seed(4)
        made = {"n": 0}
        orig_step, orig_backup = simc.Simulator.doStep, wrk.Node.back_up
        def counting_step(self, action):
            made["n"] += 1
            return orig_step(self, action)
        its = []
        def rec_backup(self, reward):
            its.append(dict(origins=[int(a) for a in self.origins], reward=float(reward), steps=made["n"]))
            made["n"] = 0
            return orig_backup(self, reward)
        simc.Simulator.doStep, wrk.Node.back_up = counting_step, rec_backup
        master = {hash(()): mnode.MasterNode(3, nodehash=hash(()))}
        out = dict(workers=[])
        with contextlib.redirect_stdout(io.StringIO()):
            for w in range(3):
                its = []
                tree = wrk.Tree(workerid=w, nscenario=3, probability=%r, budget=%d, simulator=sims[w], destination=%r,
                                TimeMin=%r, buffer=[])
                tree.uct_search([0, 44.0, 355.0], %d, master)
                out["workers"].append(dict(iterations=its, nodes=[dict(origins=[int(a) for a in nd.origins],
                                                                       hist=[[int(x) for x in h.h] for h in nd.Values])
                                                                  for nd in tree.Nodes]))
        out["master"] = {str(k): [[[int(x) for x in h.h] for h in row] for row in nd.rewards] for k, nd in master.items()}
        st = random.getstate()
        out["py_state"] = [st[0], list(st[1]), st[2]]
        ns = np.random.get_state()
        out["np_key"], out["np_pos"] = [int(x) for x in ns[1]], int(ns[2])
        print(json.dumps(out))
        """) % (ROOT, prob, budget, dest, tmin, frequency)
    res = subprocess.run([sys.executable, "-W", "ignore", "-c", script], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                         text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]
    rec = json.loads(res.stdout.strip().splitlines()[-1])

    lib = N.lib()
    h = C.c_void_p()
    ids, p = np.arange(3, dtype=np.int32), np.asarray(prob)
    N.check_forest(lib.pmcts_forest_create(3, 3, ids.ctypes.data, 12, p.ctypes.data, 1 / 2 ** 0.5, 0.5, None, budget + 8,
                                           C.byref(h)))
    try:
        random.seed(4)
        np.random.seed(4)
        py, npst = N.MtState.from_python(random.getstate()), N.MtState.from_numpy(np.random.get_state())
        for w in range(3):
            its = rec["workers"][w]["iterations"]
            assert len(its) == budget
            miss = []

            def hook(_user, iteration, scen, path, length, z, n_instants, bin_p, steps_p, its=its, w=w, miss=miss):
                got = [45 * int(path[j]) for j in range(length)]
                if got != its[iteration]["origins"] or scen != w:
                    miss.append((iteration, got, its[iteration]["origins"]))
                    return 1
                bin_p[0] = Hist.bin_of(its[iteration]["reward"])
                steps_p[0] = its[iteration]["steps"]
                return 0

            cb = N.SEQ_ROLLOUT_HOOK_FN(hook)
            rc = lib.pmcts_forest_search_sequential_hook(h, w, budget, frequency, C.byref(py), C.byref(npst), cb, None, None)
            assert rc == 0 and not miss, (w, miss, lib.pmcts_forest_last_error().decode())
            n = int(lib.pmcts_forest_tree_size(h, w))
            nodes = rec["workers"][w]["nodes"]
            assert n == len(nodes) == budget + 1
            parent, arm, table = np.empty(n, np.int32), np.empty(n, np.int8), np.empty((n, 8, 12), np.int64)
            N.check_forest(lib.pmcts_forest_export_tree(h, w, parent.ctypes.data, arm.ctypes.data, table.ctypes.data, None, None))
            paths = []
            for i in range(n):
                paths.append([] if parent[i] < 0 else paths[parent[i]] + [45 * int(arm[i])])
                assert paths[i] == nodes[i]["origins"] and np.array_equal(table[i], np.array(nodes[i]["hist"])), (w, i)
        m = int(lib.pmcts_forest_master_size(h))
        assert m == len(rec["master"])
        mparent, marm, mtable = np.empty(m, np.int32), np.empty(m, np.int8), np.empty((m, 3, 8, 12), np.uint32)
        N.check_forest(lib.pmcts_forest_export_master(h, mparent.ctypes.data, marm.ctypes.data, mtable.ctypes.data))
        mpaths = []
        for i in range(m):
            mpaths.append(() if mparent[i] < 0 else mpaths[mparent[i]] + (45 * int(marm[i]),))
            assert np.array_equal(mtable[i], np.array(rec["master"][str(hash(mpaths[i]))])), mpaths[i]
        shared = sum(1 for i in range(m) if (mtable[i].reshape(3, -1).sum(1) > 0).sum() >= 2)
        assert shared > 10   # nodes more than one scenario visited: where the master's value mixes scenarios
        want_py = (rec["py_state"][0], tuple(rec["py_state"][1]), rec["py_state"][2])
        assert py.to_python() == want_py
        assert list(npst.key) == rec["np_key"] and npst.pos == rec["np_pos"]
    finally:
        lib.pmcts_forest_destroy(h)

    # The same three searches through the PYTHON classes of the mirror (worker.Tree.uct_search: tree_policy, expand,
    # best_child, get_master_uct, Node.back_up, integrate_buffer; master_node.MasterNode), the rollout answered from the
    # recording and drawing the normals its transitions drew: the reference's leaves, trees, master and generator states.
    from iboat_pmcts_b200 import master_node as mn, synth, worker
    from iboat_pmcts_b200.forest import create_simulators
    from iboat_pmcts_b200.weatherTLKT import Weather
    weathers = [Weather(*[synth.wind_fields(s)[i] for i in (1, 2, 0, 3, 4)]) for s in range(3)]
    sims = create_simulators(weathers, 3, simtimestep=6, stateinit=(0, 44.0, 355.0), ndaysim=3)
    saved = (worker.RHO, worker.UCT_COEFF)
    worker.RHO, worker.UCT_COEFF = 0.5, 1 / 2 ** 0.5
    try:
        random.seed(4)
        np.random.seed(4)
        master = {hash(()): mn.MasterNode(3, nodehash=hash(()))}
        for w in range(3):
            its = rec["workers"][w]["iterations"]
            tree = worker.Tree(workerid=w, nscenario=3, probability=prob, budget=budget, simulator=sims[w],
                               destination=dest, TimeMin=tmin, buffer=[])

            def recorded_rollout(leaf, tree=tree, its=its):
                it = its[tree.ite]
                assert [int(a) for a in leaf.origins] == it["origins"], (tree.id, tree.ite)
                for _ in range(it["steps"]):
                    random.gauss(0.0, 1.0)
                return it["reward"]

            tree.default_policy = recorded_rollout
            import contextlib
            import io
            with contextlib.redirect_stdout(io.StringIO()):
                tree.uct_search([0, 44.0, 355.0], frequency, master)
            nodes = rec["workers"][w]["nodes"]
            assert len(tree.Nodes) == len(nodes)
            for node, want in zip(tree.Nodes, nodes):
                assert [int(a) for a in node.origins] == want["origins"]
                assert np.array_equal(np.array([hh.h for hh in node.Values]), np.array(want["hist"]))
        assert {str(k) for k in master} == set(rec["master"])
        for k, node in master.items():
            assert np.array_equal(node._table, np.array(rec["master"][str(k)])), k
        st = random.getstate()
        assert (st[0], st[1], st[2]) == want_py
        ns = np.random.get_state()
        assert [int(x) for x in ns[1]] == rec["np_key"] and int(ns[2]) == rec["np_pos"]
    finally:
        worker.RHO, worker.UCT_COEFF = saved
