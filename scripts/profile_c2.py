"""cProfile of BASELINE config 2 (10 scenarios x 10^4 leaves per tree) through Forest.launch_search(mode="batched")."""
import cProfile
import os
import pstats
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import configs_report as cr  # noqa: E402
from iboat_pmcts_b200 import forest as ft  # noqa: E402


def main():
    leaves = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    replicas = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    budget = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
    sims = cr.sims_for(range(10))
    random.seed(2)
    np.random.seed(2)
    dest, tmin = ft.initialize_simulators(sims, 50, cr.STATE0, 0, seed=2)
    f = ft.Forest(listsimulators=sims, destination=dest, timemin=tmin, budget=budget)
    pipeline = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    f.launch_search(cr.STATE0, 10, mode="batched", leaves=leaves, replicas=replicas, seed=3, pipeline=pipeline)  # warm-up
    for _ in range(3):  # plain timings first (the profiler's own overhead is of the order of the search)
        f = ft.Forest(listsimulators=sims, destination=dest, timemin=tmin, budget=budget)
        t0 = time.perf_counter()
        f.launch_search(cr.STATE0, 10, mode="batched", leaves=leaves, replicas=replicas, seed=3, pipeline=pipeline)
        dt = time.perf_counter() - t0
        rep = f.last_search_report
        print("leaves %d replicas %d budget %d pipeline %d: %.2f ms, %.3g rollouts/s | native %s" % (
            leaves, replicas, budget, pipeline, dt * 1e3, rep["rollouts"] / dt,
            {k: (round(v, 3) if isinstance(v, float) else v) for k, v in rep.items()}))
    f = ft.Forest(listsimulators=sims, destination=dest, timemin=tmin, budget=budget)
    pr = cProfile.Profile()
    pr.enable()
    f.launch_search(cr.STATE0, 10, mode="batched", leaves=leaves, replicas=replicas, seed=3, pipeline=pipeline)
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(16)


if __name__ == "__main__":
    main()
