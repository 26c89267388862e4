"""Sharded forest search on N GPUs (torchrun): every rank must return the same master dictionary, and the
same one as a single-rank run (the search does not depend on the number of ranks).

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/forest_multi_gpu.py
"""
import hashlib
import os
import random
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from iboat_pmcts_b200 import forest as ft, master_node as mn, synth  # noqa: E402
from iboat_pmcts_b200.master import MasterTree  # noqa: E402
from iboat_pmcts_b200.weatherTLKT import Weather  # noqa: E402

STATE0 = [0, 44.0, 355.0]
DEST, TMIN = [46.77751004666887, 355.0], 2.022360548788055


def digest(master):
    h = hashlib.sha256()
    for k in sorted(master):
        h.update(str(k).encode())
        h.update(np.ascontiguousarray(master[k]._table).tobytes())
    return h.hexdigest()


def search(n_scen, budget, group_world):
    ws = []
    for s in range(n_scen):
        t, la, lo, u, v = synth.wind_fields(s)
        ws.append(Weather(la, lo, t, u, v))
    sims = ft.create_simulators(ws, n_scen, simtimestep=6, stateinit=tuple(STATE0), ndaysim=3)
    random.seed(7)
    np.random.seed(7)
    forest = ft.Forest(listsimulators=sims, destination=DEST, timemin=TMIN, budget=budget)
    t0 = time.perf_counter()
    master = forest.launch_search(STATE0, 10, mode="batched", leaves=64, replicas=64, seed=5,
                                  group=None if group_world else False)
    dt = time.perf_counter() - t0
    return sims, master, dt


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_scen, budget = 8, 512
    sims, master, dt = search(n_scen, budget, True)
    d = digest(master)
    if world > 1:
        all_d = [None] * world
        dist.all_gather_object(all_d, d)
        assert len(set(all_d)) == 1, all_d
    root = master[ft.ROOT_HASH]
    assert [int(root._table[s].sum()) for s in range(n_scen)] == [budget * 64] * n_scen
    if rank == 0:
        mt = MasterTree(sims, DEST, nodes=mn.deepcopy_dict(master))
        mt.get_best_policy(verbose=False)
        from iboat_pmcts_b200.engine import default_engine
        peers = [v[0] for v in default_engine().__dict__.get("_peers", {}).values()]
        how = "one rank" if world == 1 else ("peer windows" if any(p is not None for p in peers) else "library all-gather")
        print("world %d (%s): %d scenarios x %d leaves x 64 replicas in %.2f s; master digest %s; global policy %s"
              % (world, how, n_scen, budget, dt, d[:16], [int(a) for a in mt.best_policy[-1]]), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
