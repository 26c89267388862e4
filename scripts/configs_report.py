"""Runs BASELINE.json configs 1, 2 and 5 through the reference-API mirror on one GPU and prints one JSON
line per config (configs 3 and 4 are bench.py --workload c3 / c4).  CPU figures: bench.py's cpu_baseline and
--impl reference (the Python reference itself cannot travel to the GPU box; its measured numbers are in
BASELINE.md)."""
import json
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from iboat_pmcts_b200 import forest as ft, isochrones as iso, master_node as mn, synth, worker  # noqa: E402
from iboat_pmcts_b200.master import MasterTree  # noqa: E402
from iboat_pmcts_b200.weatherTLKT import Weather  # noqa: E402

STATE0 = [0, 44.0, 355.0]


def sims_for(ids):
    ws = []
    for s in ids:
        t, la, lo, u, v = synth.wind_fields(s)
        ws.append(Weather(la, lo, t, u, v))
    return ft.create_simulators(ws, len(ws), simtimestep=6, stateinit=tuple(STATE0), ndaysim=3)


def main():
    import io
    import contextlib
    cores = os.cpu_count()
    # ---- config 1: Tutorial-shaped, 1 scenario, 1 tree, 1000 rollouts --------------------------------
    sims = sims_for([0])
    random.seed(1); np.random.seed(1)
    dest, tmin = ft.initialize_simulators(sims, 50, STATE0, 0)
    worker.RHO, worker.UCT_COEFF = 0.5, 1 / 2 ** 0.5
    f = ft.Forest(listsimulators=sims, destination=dest, timemin=tmin, budget=1000)
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        m = f.launch_search(STATE0, 10, mode="sequential")
    dt_seq = time.perf_counter() - t0
    mt = MasterTree(sims, dest, nodes=mn.deepcopy_dict(m)); mt.get_best_policy(verbose=False)
    pol_seq = [int(a) for a in mt.best_policy[-1]]
    f = ft.Forest(listsimulators=sims, destination=dest, timemin=tmin, budget=1000)
    t0 = time.perf_counter()
    m = f.launch_search(STATE0, 10, mode="batched", leaves=50, replicas=1, seed=1)
    dt_b = time.perf_counter() - t0
    mt = MasterTree(sims, dest, nodes=mn.deepcopy_dict(m)); mt.get_best_policy(verbose=False)
    print(json.dumps(dict(config="c1 tutorial: 1 scenario, 1 tree, 1000 rollouts", destination=dest, timemin=tmin,
                          sequential_reference_order=dict(seconds=dt_seq, rollouts_per_s=1000 / dt_seq, policy=pol_seq),
                          batched_50_leaves=dict(seconds=dt_b, rollouts_per_s=1000 / dt_b,
                                                 policy=[int(a) for a in mt.best_policy[-1]]),
                          reference_python_cpu="8.1 s, 124 rollouts/s (BASELINE.md, build container)")), flush=True)
    # ---- config 2: forest, 10 scenarios x 10^4 rollouts per tree ---------------------------------------
    sims = sims_for(range(10))
    random.seed(2); np.random.seed(2)
    dest, tmin = ft.initialize_simulators(sims, 50, STATE0, 0, seed=2)
    for leaves, replicas, budget in ((256, 1, 10000), (64, 32, 320)):
        f = ft.Forest(listsimulators=sims, destination=dest, timemin=tmin, budget=budget)
        t0 = time.perf_counter()
        m = f.launch_search(STATE0, 10, mode="batched", leaves=leaves, replicas=replicas, seed=3)
        dt = time.perf_counter() - t0
        mt = MasterTree(sims, dest, nodes=mn.deepcopy_dict(m)); mt.get_best_policy(verbose=False)
        tot = 10 * budget * replicas
        print(json.dumps(dict(config="c2 forest: 10 scenarios x %d leaves x %d replicas per tree" % (budget, replicas),
                              seconds=dt, rollouts_per_s=tot / dt, global_policy=[int(a) for a in mt.best_policy[-1]],
                              reference_python_cpu="3 scenarios x 300: 63 rollouts/s with the Manager dict (BASELINE.md)")),
              flush=True)
    # ---- config 5: plan evaluation, 10^5 draws per scenario, 10 scenarios ---------------------------------
    plan = [0, 0, 315, 0, 45, 0]
    iso.estimate_perfomance_plan(sims, 100000, STATE0, dest, plan=plan, verbose=False, seed=4)  # warm-up at full size (scratch allocation)
    t0 = time.perf_counter()
    mean, var = iso.estimate_perfomance_plan(sims, 100000, STATE0, dest, plan=plan, verbose=False, seed=5)
    dt = time.perf_counter() - t0
    print(json.dumps(dict(config="c5 policy evaluation: plan %s, 10^5 draws x 10 scenarios" % plan, seconds=dt,
                          rollouts_per_s=1e6 / dt, global_mean_days=float(mean[0]), global_var=float(var[0]),
                          cpu="bench.py's cpu_baseline / --impl reference time the CPU restatement of the same rollouts")),
          flush=True)

if __name__ == "__main__":
    main()
