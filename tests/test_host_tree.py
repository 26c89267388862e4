"""Host side of the search (no GPU): reward histograms, worker-tree back-up, master updates and the
master-tree post-processing, replayed on the leaf / slot / reward sequence recorded from a seeded run of
the reference's own ``Tree.uct_search`` (tests/golden/tree_replay_c1.json, oracle/gen_golden.py).
Visit counts must be bit-exact; the best policy must be the reference's.
"""
import json
import os

import numpy as np
import pytest

from iboat_pmcts_b200 import master_node as mn
from iboat_pmcts_b200 import worker
from iboat_pmcts_b200.forest import ROOT_HASH, _master_add
from iboat_pmcts_b200.utils import Hist


@pytest.fixture(scope="module")
def replay(golden_dir):
    with open(os.path.join(golden_dir, "tree_replay_c1.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def kat(golden_dir):
    with open(os.path.join(golden_dir, "kat.json")) as f:
        return json.load(f)


class _Sim:
    """What Tree needs from a Simulator when no rollout is launched."""
    times = np.arange(0, 3 + 6 / 24, 6 * (1 / 24))


def test_hist_bins_and_means(kat):
    for value, want in kat["hist_bin"]:
        assert Hist.bin_of(value) == want
        h = Hist()
        h.add(value)
        assert h.h[want] == 1 and h.h.sum() == 1
    assert Hist().get_mean() == 0 and Hist().is_empty()
    h = Hist([0, 2, 0, 1, 0, 0, 0, 0, 0, 0, 0, 3])
    assert h.get_mean() == np.dot(h.h, Hist.MEANS) / 6
    assert np.array_equal(Hist.MEANS, (np.arange(12) + 0.5) * 0.125)
    # a Hist can be a live view of a table row
    table = np.zeros((3, 12), dtype=int)
    v = Hist(store=table[1])
    v.add(0.5)
    assert table[1, 3] == 1


def _replay_tree(replay, monkeypatch):
    """Feeds the recorded iterations into a fresh Tree + master dictionary."""
    monkeypatch.setattr(worker, "RHO", replay["rho"])
    monkeypatch.setattr(worker, "UCT_COEFF", replay["uct_coeff"])
    tree = worker.Tree(workerid=0, nscenario=1, budget=replay["budget"], simulator=_Sim(), destination=[46.7, 355.0],
                       TimeMin=2.0)
    tree._make_root([0, 44.0, 355.0])
    master = {ROOT_HASH: mn.MasterNode(1, nodehash=ROOT_HASH)}
    by_path = {(): tree.rootNode}
    slots = iter([])
    feed = {"queue": []}
    monkeypatch.setattr(np.random, "randint", lambda *a, **k: feed["queue"].pop(0))
    mslots = list(replay["master_slots"])
    freq = replay["frequency"]
    for i, it in enumerate(replay["iterations"]):
        origins = it["origins"]
        parent = by_path[tuple(origins[:-1])]
        # the expansion the reference's tree policy made at this iteration
        assert origins[-1] in [int(a) for a in parent.actions]
        parent.actions.remove(origins[-1])
        row = tree._new_row()
        leaf = worker.Node(parent=parent, origins=origins, depth=parent.depth + 1, store=row)
        parent.children.append(leaf)
        tree.Nodes.append(leaf)
        tree.depth = max(tree.depth, leaf.depth)
        by_path[tuple(origins)] = leaf
        feed["queue"] = [it["slot"]]
        leaf.back_up(it["reward"])
        tree.buffer.append([0, hash(tuple(origins)), hash(tuple(origins[:-1])), origins[-1], it["reward"]])
        if (i + 1) % freq == 0:
            feed["queue"] = [mslots.pop(0) for _ in range(freq)]
            tree.integrate_buffer(master)
            tree.buffer = []
            assert not feed["queue"]
    del slots
    return tree, master


def test_replayed_search_visit_counts_are_bit_exact(replay, monkeypatch):
    tree, master = _replay_tree(replay, monkeypatch)
    assert len(tree.Nodes) == len(replay["nodes"]) and tree.depth == replay["depth"]
    for node, want in zip(tree.Nodes, replay["nodes"]):
        assert [int(a) for a in node.origins] == want["origins"]
        assert np.array_equal(np.array([h.h for h in node.Values]), np.array(want["hist"]))
    # the table behind the nodes is the same numbers
    assert np.array_equal(tree.table[:len(tree.Nodes)], np.array([n["hist"] for n in replay["nodes"]]))
    assert tree.rootNode.visits() == replay["budget"]
    # master dictionary: same keys (CPython tuple hashes), same histograms
    assert {str(k) for k in master} == set(replay["master_nodes"])
    for k, node in master.items():
        assert np.array_equal(np.array([h.h for h in node.rewards[0]]), np.array(replay["master_nodes"][str(k)]))


def test_deepcopy_dict_rewires_the_replayed_master(replay, monkeypatch):
    """(MasterTree's policy on this dictionary: tests/test_master_policy_gpu.py -- the reductions run on the device.)"""
    _, master = _replay_tree(replay, monkeypatch)
    nodes = mn.deepcopy_dict(master)
    assert len(master) == 0  # emptied, like the reference's
    root = nodes[ROOT_HASH]
    assert root.depth == 0 and root.parentNode is None and len(root.children) == 8
    assert max(n.depth for n in nodes.values()) == replay["depth"]


def test_batched_master_update_equals_integrate_buffer(replay, monkeypatch):
    """forest._master_add with one-reward histograms == Tree.integrate_buffer on the same updates."""
    _, master = _replay_tree(replay, monkeypatch)
    batched = {ROOT_HASH: mn.MasterNode(1, nodehash=ROOT_HASH)}
    mslots = list(replay["master_slots"])
    for it, slot in zip(replay["iterations"], mslots):
        bins = np.zeros(12, np.uint32)
        bins[Hist.bin_of(it["reward"])] = 1
        _master_add(batched, 1, 0, it["origins"], slot, bins)
    assert set(batched) == set(master)
    for k in master:
        assert np.array_equal(batched[k]._table, master[k]._table)
        assert batched[k].arm == master[k].arm
        assert (batched[k].parentNode.hash if batched[k].parentNode else None) == \
               (master[k].parentNode.hash if master[k].parentNode else None)


def test_leaf_batched_selection_spreads_and_backs_up(monkeypatch):
    """select_leaves puts virtual visits on the paths of a wave (so the wave spreads over the tree) and
    back_up_leaves removes them while adding whole histograms."""
    import random
    random.seed(3)
    np.random.seed(3)
    tree = worker.Tree(workerid=0, nscenario=1, budget=100, simulator=_Sim(), destination=[46.7, 355.0], TimeMin=2.0)
    tree._make_root([0, 44.0, 355.0])
    master = {ROOT_HASH: mn.MasterNode(1, nodehash=ROOT_HASH)}
    for wave in range(6):
        leaves = tree.select_leaves(16, master)
        assert len({tuple(l.origins) for l in leaves}) == 16  # all distinct
        assert tree.rootNode.pending == 16
        bins = np.zeros((16, 12), np.uint32)
        bins[np.arange(16), np.random.randint(12, size=16)] = 32
        tree.back_up_leaves(leaves, bins, np.random.randint(8, size=16))
        assert all(n.pending == 0 for n in tree.Nodes)
    assert tree.ite == 96 and tree.rootNode.visits() == 96 * 32
    assert tree.depth >= 2
    # every node holds its own wave's rollouts (in a random slot, quirk Q2) plus those of all nodes below it
    for node in tree.Nodes[1:]:
        below = sum(1 for n in tree.Nodes if n.origins[:len(node.origins)] == node.origins)
        assert node._table.sum() == below * 32


def test_hist_bin_on_and_around_every_threshold():
    """Hist.add (utils.py:34-47) counts the interior thresholds STRICTLY below the value: on a threshold the value stays
    in the lower bin, one ulp above it moves up.  The mirror's Hist, the oracle's orc_hist_bin (what the kernels' bins
    are compared with) and NumPy's own statement of the rule agree on every threshold, its two neighbours, values far
    outside the range and 10^5 random rewards."""
    from oracle import cport
    thresholds = np.asarray(Hist.THRESH[1:-1], np.float64)
    assert thresholds.size == 11 and np.array_equal(thresholds, 0.125 * np.arange(1, 12))
    values = [-1e300, -0.1, -0.0, 0.0, 5e-324, 1.5, 1.5000001, 2.0, 1e300]
    for t in thresholds:
        values += [float(np.nextafter(t, -np.inf)), float(t), float(np.nextafter(t, np.inf))]
    values += np.random.default_rng(3).uniform(-0.2, 1.8, 100000).tolist()
    values = np.asarray(values)
    want = (values[:, None] > thresholds[None, :]).sum(1)
    got_oracle = np.array([cport.hist_bin(float(v)) for v in values])
    assert np.array_equal(got_oracle, want)
    head = values[:200]
    assert [Hist.bin_of(float(v)) for v in head] == want[:200].tolist()
    for k, t in enumerate(thresholds):
        assert Hist.bin_of(float(t)) == k and Hist.bin_of(float(np.nextafter(t, np.inf))) == k + 1
    h = Hist()
    for v in head:
        h.add(float(v))
    assert np.array_equal(h.h, np.bincount(want[:200], minlength=12))
