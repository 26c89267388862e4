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
from iboat_pmcts_b200.forest import MAX_DEPTH, ROOT_HASH, _all_gather_blocks, _master_add

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
