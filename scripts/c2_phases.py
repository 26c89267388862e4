"""Where the time of one BASELINE config-2 step goes (Forest.launch_search + MasterTree.get_best_policy on 10 scenarios,
320 leaves x 32 replicas per tree): medians over 20 steps of each phase, the native driver's own report, and a
cProfile of one step."""
import cProfile
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import bench as B  # noqa: E402
from iboat_pmcts_b200 import forest as ft, master_node as mn, synth  # noqa: E402
from iboat_pmcts_b200.master import MasterTree  # noqa: E402
from iboat_pmcts_b200.weatherTLKT import Weather  # noqa: E402

ws = []
for s in range(10):
    t, la, lo, u, v = synth.wind_fields(s)
    ws.append(Weather(la, lo, t, u.astype(np.float32), v.astype(np.float32)))
sims = ft.create_simulators(ws, 10, simtimestep=6, stateinit=tuple(B.STATE0), ndaysim=3)
forests = [ft.Forest(listsimulators=sims, destination=list(B.DEST), timemin=B.TMIN, budget=320) for _ in range(30)]


def step(f, i, rec=None):
    np.random.seed(i)
    t0 = time.perf_counter()
    m = f.launch_search(list(B.STATE0), 10, mode="batched", leaves=64, replicas=32, seed=i)
    t1 = time.perf_counter()
    nodes = mn.deepcopy_dict(m)
    mt = MasterTree(sims, list(B.DEST), nodes=nodes)
    t2 = time.perf_counter()
    mt.get_best_policy(verbose=False)
    t3 = time.perf_counter()
    if rec is not None:
        r = f.last_search_report
        rec.append((t1 - t0, t2 - t1, t3 - t2, r["ms_total"] * 1e-3, r["ms_select"] * 1e-3, r["ms_wait"] * 1e-3,
                    r["ms_backup"] * 1e-3, r["ms_master"] * 1e-3))
    return mt


for i in range(4):
    step(forests[i], i)
rec = []
for i in range(4, 24):
    step(forests[i], i, rec)
r = np.median(np.array(rec), 0) * 1e3
print("medians over 20 steps: launch_search %.3f ms (native driver %.3f: select %.3f, GPU wait %.3f, back-up %.3f, master "
      "update %.3f), deepcopy + MasterTree %.3f, get_best_policy %.3f; whole step %.3f (min %.3f, max %.3f)"
      % (r[0], r[3], r[4], r[5], r[6], r[7], r[1], r[2], float(np.median(np.array(rec)[:, :3].sum(1))) * 1e3,
         float(np.array(rec)[:, :3].sum(1).min()) * 1e3, float(np.array(rec)[:, :3].sum(1).max()) * 1e3))
print("last native report:", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in forests[23].last_search_report.items()})
pr = cProfile.Profile()
pr.enable()
step(forests[24], 24)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
