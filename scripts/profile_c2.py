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
    f.launch_search(cr.STATE0, 10, mode="batched", leaves=leaves, replicas=replicas, seed=3)  # warm-up
    f = ft.Forest(listsimulators=sims, destination=dest, timemin=tmin, budget=budget)
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    f.launch_search(cr.STATE0, 10, mode="batched", leaves=leaves, replicas=replicas, seed=3)
    pr.disable()
    print("leaves %d replicas %d budget %d: %.3f s" % (leaves, replicas, budget, time.perf_counter() - t0))
    pstats.Stats(pr).sort_stats("cumulative").print_stats(22)


if __name__ == "__main__":
    main()
