"""Decision parity of the searches: which first heading does MasterTree.get_best_policy return?

For scenarios 0-9 and 20 seeds each, three searches of one scenario's tree from the same root to the same destination:
  seq      Forest.launch_search(mode="sequential"): the reference's own loop and random streams (one rollout per
           iteration, 1000 iterations -- Tutorial.py's budget);
  batch1k  mode="batched" with the same 1000 rollouts (20 waves of 50 leaves, 1 replica);
  batch10k mode="batched" at BASELINE config 2's budget (5 waves of 64 leaves, 32 replicas: 10^4 rollouts).
MCTS is a randomised algorithm: two reference-order searches with different seeds do not always agree either, so the
table reports, per scenario, the modal heading of each kind of search, how often each kind returns the modal heading of
the reference-order searches, and the pairwise agreement seq-vs-seq as the yardstick.

  python scripts/decision_parity.py [n_seeds] > profiles/r02_decision_parity.json
"""
import contextlib
import io
import json
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from iboat_pmcts_b200 import forest as ft, master_node as mn, synth, worker  # noqa: E402
from iboat_pmcts_b200.master import MasterTree  # noqa: E402
from iboat_pmcts_b200.weatherTLKT import Weather  # noqa: E402

STATE0 = [0, 44.0, 355.0]


def first_heading(sims, dest, master):
    mt = MasterTree(sims, dest, nodes=mn.deepcopy_dict(master))
    mt.get_best_policy(verbose=False)
    return int(mt.best_policy[-1][0])


def study(scenarios=range(10), n_seeds=20, seq_budget=1000, out=sys.stdout):
    worker.RHO, worker.UCT_COEFF = 0.5, 1 / 2 ** 0.5  # Tutorial.py:61-62
    rows = []
    t_start = time.perf_counter()
    for s in scenarios:
        t, la, lo, u, v = synth.wind_fields(s)
        sims = ft.create_simulators([Weather(la, lo, t, u, v)], 1, simtimestep=6, stateinit=tuple(STATE0), ndaysim=3)
        random.seed(100 + s)
        np.random.seed(100 + s)
        dest, tmin = ft.initialize_simulators(sims, 50, STATE0, 0)
        got = dict(seq=[], batch1k=[], batch10k=[])
        for seed in range(n_seeds):
            random.seed(seed)
            np.random.seed(seed)
            f = ft.Forest(listsimulators=sims, destination=dest, timemin=tmin, budget=seq_budget)
            with contextlib.redirect_stdout(io.StringIO()):
                m = f.launch_search(STATE0, 10, mode="sequential")
            got["seq"].append(first_heading(sims, dest, m))
            np.random.seed(seed)
            f = ft.Forest(listsimulators=sims, destination=dest, timemin=tmin, budget=seq_budget)
            m = f.launch_search(STATE0, 10, mode="batched", leaves=50, replicas=1, seed=seed)
            got["batch1k"].append(first_heading(sims, dest, m))
            np.random.seed(seed)
            f = ft.Forest(listsimulators=sims, destination=dest, timemin=tmin, budget=320)
            m = f.launch_search(STATE0, 10, mode="batched", leaves=64, replicas=32, seed=seed)
            got["batch10k"].append(first_heading(sims, dest, m))
        modal = max(set(got["seq"]), key=got["seq"].count)
        seq = got["seq"]
        pairs = [(a == b) for i, a in enumerate(seq) for b in seq[i + 1:]]
        row = dict(scenario=s, destination=[float(x) for x in dest], timemin=float(tmin), modal_seq=modal,
                   headings={k: {str(h): v.count(h) for h in sorted(set(v))} for k, v in got.items()},
                   share_of_modal={k: v.count(modal) / len(v) for k, v in got.items()},
                   seq_vs_seq_pairwise=float(np.mean(pairs)) if pairs else 1.0,
                   batch1k_vs_seq_same_seed=float(np.mean([a == b for a, b in zip(got["batch1k"], seq)])),
                   batch10k_vs_seq_same_seed=float(np.mean([a == b for a, b in zip(got["batch10k"], seq)])))
        rows.append(row)
        print(json.dumps(row), file=out, flush=True)
    total = dict(summary=True, scenarios=len(rows), seeds=n_seeds,
                 share_of_modal={k: float(np.mean([r["share_of_modal"][k] for r in rows])) for k in ("seq", "batch1k", "batch10k")},
                 seq_vs_seq_pairwise=float(np.mean([r["seq_vs_seq_pairwise"] for r in rows])),
                 batch1k_vs_seq_same_seed=float(np.mean([r["batch1k_vs_seq_same_seed"] for r in rows])),
                 batch10k_vs_seq_same_seed=float(np.mean([r["batch10k_vs_seq_same_seed"] for r in rows])),
                 seconds=time.perf_counter() - t_start)
    print(json.dumps(total), file=out, flush=True)
    return rows, total


if __name__ == "__main__":
    study(n_seeds=int(sys.argv[1]) if len(sys.argv) > 1 else 20)
