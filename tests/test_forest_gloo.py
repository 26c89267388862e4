"""Multi-rank host logic of the sharded search on the CPU (gloo, world_size 2): the per-wave update blocks
are all-gathered and applied in (rank, scenario, leaf) order, so every rank ends with the same master
dictionary as a single process that owns all scenarios.  No GPU: the rollout results are synthetic.
"""
import os
import pickle
import socket
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from iboat_pmcts_b200 import master_node as mn
from iboat_pmcts_b200.forest import ROOT_HASH, _all_gather_blocks, _master_add

MAX_DEPTH = 12  # width of the path block: len(times) - 1 of the 6-hourly, 3-day axis

S, B, WAVES = 4, 6, 5


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _wave_block(scen_ids, wave):
    """Deterministic synthetic wave for the given scenarios: leaf paths, master slots, per-leaf histograms."""
    paths = np.full((len(scen_ids), B, MAX_DEPTH), -1, np.int32)
    slots = np.zeros((len(scen_ids), B), np.int32)
    bins = np.zeros((len(scen_ids), B, 12), np.uint32)
    for i, s in enumerate(scen_ids):
        rng = np.random.default_rng(1000 * wave + s)
        for l in range(B):
            depth = 1 + (wave + l) % 3
            paths[i, l, :depth] = 45 * rng.integers(0, 8, depth)
            slots[i, l] = rng.integers(0, 8)
            bins[i, l] = rng.multinomial(32, np.full(12, 1 / 12))
    return dict(ids=np.asarray(scen_ids, np.int32), paths=paths, slots=slots, bins=bins)


def _apply(master, blocks):
    for blk in blocks:
        for si, scen_id in enumerate(blk["ids"]):
            for li in range(B):
                path = [a for a in blk["paths"][si, li] if a >= 0]
                _master_add(master, S, int(scen_id), path, int(blk["slots"][si, li]), blk["bins"][si, li])


def _tables(master):
    return {k: (n._table.copy(), n.arm, n.parentNode.hash if n.parentNode else None) for k, n in master.items()}


def _rank_main(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        local = list(range(rank * (S // world), (rank + 1) * (S // world)))
        master = {ROOT_HASH: mn.MasterNode(S, nodehash=ROOT_HASH)}
        for wave in range(WAVES):
            blocks = _all_gather_blocks(_wave_block(local, wave), dist, None, world)
            assert [list(b["ids"]) for b in blocks] == [list(range(r * (S // world), (r + 1) * (S // world))) for r in range(world)]
            _apply(master, blocks)
        with open(os.path.join(out_dir, "rank%d.pkl" % rank), "wb") as f:
            pickle.dump(_tables(master), f)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_ranks_build_the_same_master_as_one_process():
    world = 2
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_rank_main, args=(world, _free_port(), d), nprocs=world, join=True)
        got = []
        for r in range(world):
            with open(os.path.join(d, "rank%d.pkl" % r), "rb") as f:
                got.append(pickle.load(f))
    # single process, same (rank, scenario, leaf) order
    master = {ROOT_HASH: mn.MasterNode(S, nodehash=ROOT_HASH)}
    for wave in range(WAVES):
        _apply(master, [_wave_block(list(range(r * (S // world), (r + 1) * (S // world))), wave) for r in range(world)])
    want = _tables(master)
    for g in got:
        assert set(g) == set(want)
        for k in want:
            assert np.array_equal(g[k][0], want[k][0]) and g[k][1:] == want[k][1:]
    # every scenario's root row holds all its rollouts
    root = want[ROOT_HASH][0]
    assert [int(root[s].sum()) for s in range(S)] == [WAVES * B * 32] * S
    # integer adds commute: any order of the blocks gives the same counts
    master2 = {ROOT_HASH: mn.MasterNode(S, nodehash=ROOT_HASH)}
    for wave in reversed(range(WAVES)):
        _apply(master2, [_wave_block(list(range(r * (S // world), (r + 1) * (S // world))), wave) for r in reversed(range(world))])
    for k, v in _tables(master2).items():
        assert np.array_equal(v[0], want[k][0])


def test_single_rank_gather_is_identity():
    blk = _wave_block([0, 1], 0)
    assert _all_gather_blocks(blk, None, None, 1) == [blk]
    assert torch.distributed.is_available()


# ---- the native search driver (pmcts_forest_search) at world_size 2 ---------------------------------------
def _driver_rank_main(rank, world, port, out_dir, pipeline):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), here):
        if p not in sys.path:
            sys.path.insert(0, p)
    import test_host_forest as thf
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        forest, master, calls = thf.run("native", n_scen=4, budget=70, leaves=16, replicas=24, group=None,
                                        pipeline=pipeline)
        assert all(c[0] == 2 for c in calls)  # every launch covers this rank's two scenarios
        local = range(rank * 2, rank * 2 + 2)
        trees = {i: (forest.workers[i].table[:71].copy(), [list(map(int, n.origins)) for n in forest.workers[i].Nodes])
                 for i in local}
        with open(os.path.join(out_dir, "rank%d.pkl" % rank), "wb") as f:
            pickle.dump((_tables(master), trees, forest.last_search_report), f)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("pipeline", [1, 2])
def test_native_driver_two_ranks_equal_one_process(pipeline):
    """pmcts_forest_search with the scenarios split over two gloo ranks (update blocks exchanged through the host
    all-gather callback): both ranks end with the master of the single-process search, and each rank's worker
    trees are the single-process trees of its scenarios -- for the synchronous search and for two waves in flight."""
    import test_host_forest as thf
    world = 2
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_driver_rank_main, args=(world, _free_port(), d, pipeline), nprocs=world, join=True)
        got = []
        for r in range(world):
            with open(os.path.join(d, "rank%d.pkl" % r), "rb") as f:
                got.append(pickle.load(f))
    forest, master, _ = thf.run("native", n_scen=4, budget=70, leaves=16, replicas=24, group=False, pipeline=pipeline)
    want = _tables(master)
    for tables, trees, report in got:
        assert list(tables) == list(want)  # same nodes, created in the same order
        for k in want:
            assert np.array_equal(tables[k][0], want[k][0]) and tables[k][1:] == want[k][1:]
        for i, (table, origins) in trees.items():
            assert np.array_equal(table, forest.workers[i].table[:71])
            assert origins == [list(map(int, n.origins)) for n in forest.workers[i].Nodes]
        assert report["waves"] == 5 and report["leaves"] == 2 * 70 and report["rollouts"] == 2 * 70 * 24
    root = want[ROOT_HASH][0]
    assert [int(root[s].sum()) for s in range(4)] == [70 * 24] * 4


# ---- plan evaluation (estimate_perfomance_plan, seed=int) with the scenarios split over two ranks ----------------
class _StatsEngine:
    """Stands in for the engine in the plan evaluation: the fused statistics of a launch are a deterministic function
    of (scenario slot, noise stream, seed) -- what K4 returns per scenario: runs, sum t, sum t^2, arrived, transitions."""
    precision = 0

    def set_precision(self, mode):
        self.precision = mode

    def rollout_forest(self, scens, n_leaf, replicas, root, dest, tmin, prefix, prefix_len, prefix_stride,
                       prefix_shared=True, z=None, seed=None, stream=0, id0=0, out=None, flags=0):
        st = out["stats"].reshape(len(scens), 8)
        for j, slot in enumerate(scens):
            t = np.random.default_rng([int(seed), int(stream) + j, int(slot)]).uniform(2.0, 3.0, replicas)
            st[j] = [replicas, t.sum(), (t * t).sum(), replicas - 1, 9 * replicas, replicas, 0, 0]


class _StatsSim:
    times = np.arange(0, 3 + 6 / 24, 6 * (1 / 24))

    def __init__(self, eng, slot):
        self._eng, self._slot = eng, slot

    def scenario(self):
        return self._eng, self._slot


def _plan_rank_main(rank, world, port, out_dir):
    from iboat_pmcts_b200.isochrones import estimate_perfomance_plan
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        eng = _StatsEngine()
        sims = [_StatsSim(eng, 100 + s) for s in range(rank * 3, rank * 3 + 3)]  # three scenarios per rank, in rank order
        res = estimate_perfomance_plan(sims, 40, [0, 44.0, 355.0], [46.7, 355.0], plan=[0, 45], verbose=False, seed=9,
                                       group=dist.group.WORLD)
        with open(os.path.join(out_dir, "rank%d.pkl" % rank), "wb") as f:
            pickle.dump(res, f)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_plan_evaluation_two_ranks_equal_one_process():
    """Config 5 sharded like the search: each rank evaluates the plan on its scenarios, the per-scenario (runs, sum t,
    sum t^2) are all-gathered, and every rank returns the means and variances over all six scenarios that one process
    holding them all returns -- bit for bit (the noise streams are numbered by global scenario index)."""
    from iboat_pmcts_b200.isochrones import estimate_perfomance_plan
    world = 2
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_plan_rank_main, args=(world, _free_port(), d), nprocs=world, join=True)
        got = []
        for r in range(world):
            with open(os.path.join(d, "rank%d.pkl" % r), "rb") as f:
                got.append(pickle.load(f))
    eng = _StatsEngine()
    want = estimate_perfomance_plan([_StatsSim(eng, 100 + s) for s in range(6)], 40, [0, 44.0, 355.0], [46.7, 355.0],
                                    plan=[0, 45], verbose=False, seed=9)
    assert len(want[0]) == 7 and len(want[1]) == 7
    assert got[0] == want and got[1] == want
    with pytest.raises(ValueError):
        estimate_perfomance_plan([_StatsSim(eng, 100)], 4, [0, 44.0, 355.0], [46.7, 355.0], seed=None, group=object())
