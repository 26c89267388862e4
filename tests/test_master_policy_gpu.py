"""MasterTree's reductions on the device (pmcts_master_policy, csrc/pmcts_master.cu) against the reference's own
post-processing (code/solver/master.py:57-185):

* tests/golden/master_policy_s10.npz -- a 10-scenario master tree run through the UNMODIFIED reference MasterTree in
  the build container (oracle/gen_golden.py --only-master): best policy of every scenario and of the global tree for
  uniform and for skewed scenario probabilities, the utilities the reference returned along those policies, and every
  guessed reward its global descent left on the nodes;
* tests/golden/tree_replay_c1.json -- the policy of the recorded reference search;
* the array-backed dictionary of the native search and a plain dictionary of MasterNode objects give the same
  answers; pickle round trip.
"""
import json
import os

import numpy as np
import pytest

from fake_device import run
from iboat_pmcts_b200 import master_node as mn
from iboat_pmcts_b200 import worker
from iboat_pmcts_b200.forest import ROOT_HASH
from iboat_pmcts_b200.master import MasterTree
from iboat_pmcts_b200.simulatorTLKT import ACTIONS

pytestmark = pytest.mark.gpu


class _Sim:
    times = np.arange(0, 3 + 6 / 24, 6 * (1 / 24))


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "master_policy_s10.npz"))


def _tree(gold, tag, plain):
    parent, arm, table = gold["parent"], gold["arm"], gold["table"].astype(int)
    nodes = mn.MasterDict(parent.copy(), arm.copy(), table.copy(), wired=True)
    hashes = [nodes.node_at(i).hash for i in range(parent.size)]  # fixture index -> key
    if plain:  # a dictionary of node objects, as the sequential search / an unpickled tree hands over
        nodes = dict(nodes.items())
    mt = MasterTree([_Sim()] * 10, [46.7, 355.0], nodes=nodes,
                    proba=[] if tag == "uniform" else list(gold["probability_" + tag]))
    return mt, hashes


@pytest.mark.parametrize("plain", [False, True])
@pytest.mark.parametrize("tag", ["uniform", "skewed"])
def test_policies_utilities_and_guesses_are_the_references(gold, tag, plain, monkeypatch):
    import iboat_pmcts_b200.master as master_mod
    monkeypatch.setattr(master_mod, "UCT_COEFF", float(gold["uct_coeff"]))
    mt, hashes = _tree(gold, tag, plain)
    mt.get_best_policy(verbose=False)
    want = gold["policy_" + tag]
    for sid in range(-1, 10):
        assert [int(a) for a in mt.best_policy[sid]] == [int(a) for a in want[sid + 1] if a >= 0], sid
        path = mt.best_nodes_policy[sid]
        assert path[0].parentNode is None and len(path) == len(mt.best_policy[sid]) + 1
        assert [int(n.arm) for n in path[1:]] == [int(a) for a in mt.best_policy[sid]]
        assert all(child.parentNode is par for par, child in zip(path, path[1:]))
    # utilities along the policies: exploitation exact up to summation order, exploration to libm's last bits
    for idx, sid, value, expl in gold["utility_" + tag]:
        got = mt.get_utility(mt.nodes[hashes[int(idx)]], int(sid))
        assert got[0] == pytest.approx(value, rel=1e-13, abs=0) and got[1] == pytest.approx(expl, rel=1e-13, abs=0)
    # every guess the reference's global descent computed
    assert len(gold["guess_" + tag]) > 50
    for idx, s, g in gold["guess_" + tag]:
        assert mt.guess_reward(mt.nodes[hashes[int(idx)]], int(s)) == pytest.approx(g, rel=1e-12, abs=0)
    # get_best_child agrees with the policy step by step, and is (None, None) where a scenario expanded no child
    for sid in range(-1, 10):
        node = mt.best_nodes_policy[sid][0]
        for action in mt.best_policy[sid]:
            child, act = mt.get_best_child(node, sid)
            assert act == action and child.parentNode is node
            node = child
        if sid >= 0 and node.children:
            assert mt.get_best_child(node, sid) == (None, None)


def test_recorded_reference_search_policy(golden_dir, monkeypatch):
    """The master dictionary of the reference's own recorded uct_search (replayed on the mirror classes, plain
    dictionary of node objects) gives the reference's best policy."""
    import test_host_tree as tht
    with open(os.path.join(golden_dir, "tree_replay_c1.json")) as f:
        replay = json.load(f)
    _, master = tht._replay_tree(replay, monkeypatch)
    import iboat_pmcts_b200.master as master_mod
    # MasterTree binds UCT_COEFF at import (quirk Q4); the recorded run had set worker.UCT_COEFF *after* import,
    # so the reference's post-processing used the import-time value: keep it
    nodes = mn.deepcopy_dict(master)
    mt = MasterTree([_Sim()], [46.7, 355.0], nodes=nodes)
    mt.get_best_policy(verbose=False)
    for key, want in replay["best_policy"].items():
        assert [int(a) for a in mt.best_policy[int(key)]] == want
    assert master_mod.UCT_COEFF == 1 / 5 * 1 / 2 ** 0.5
    # pickle round trip (master.py:390-410): histograms, wiring and policy survive
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        mt.Simulators = []
        mt.save_tree("tree", directory=d + "/")
        back = MasterTree.load_tree("tree", directory=d + "/")
    assert set(back.nodes) == set(mt.nodes) and back.best_policy == mt.best_policy
    for k, node in mt.nodes.items():
        other = back.nodes[k]
        assert np.array_equal(other._table, node._table) and other.depth == node.depth
        other.rewards[0, 0].add(0.3)           # the Hist views alias the unpickled table again
        assert other._table[0, 0].sum() == node._table[0, 0].sum() + 1


def test_native_and_python_hosts_decide_alike():
    """The same search on the Python classes (plain dictionary) and in the native runtime (array-backed dictionary,
    never turned into node objects): same policies for every scenario and the global tree; the array-backed
    dictionary stays unread."""
    pols = []
    for host in ("python", "native"):
        forest, master, _ = run(host, n_scen=4, budget=100)
        nodes = mn.deepcopy_dict(master)
        mt = MasterTree([w.Simulator for w in forest.workers.values()], [46.7, 355.0], nodes=nodes)
        mt.get_best_policy(verbose=False)
        pols.append({k: [int(a) for a in v] for k, v in mt.best_policy.items()})
        if host == "native":
            assert isinstance(nodes, mn.MasterDict) and not nodes._filled  # no dictionary-wide object build
            assert len(nodes) > 100 and sum(x is not None for x in nodes._nodes) < 40
            # a saved tree is a plain pickle of plain nodes (what the reference's tools read)
            import pickle
            back = pickle.loads(pickle.dumps(mt))
            assert type(back.nodes) is dict and len(back.nodes) == len(nodes)
            root = back.nodes[ROOT_HASH]
            assert len(root.children) == 8 and all(c.parentNode is root for c in root.children)
    assert pols[0] == pols[1] and len(pols[0][-1]) >= 1
