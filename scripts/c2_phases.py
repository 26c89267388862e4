import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
import bench as B
from iboat_pmcts_b200 import forest as ft, master_node as mn, synth
from iboat_pmcts_b200.master import MasterTree
from iboat_pmcts_b200.weatherTLKT import Weather
ws = []
for s in range(10):
    t, la, lo, u, v = synth.wind_fields(s)
    ws.append(Weather(la, lo, t, u.astype(np.float32), v.astype(np.float32)))
sims = ft.create_simulators(ws, 10, simtimestep=6, stateinit=tuple(B.STATE0), ndaysim=3)
for i in range(8):
    f = ft.Forest(listsimulators=sims, destination=list(B.DEST), timemin=B.TMIN, budget=320)
    np.random.seed(i)
    t0 = time.perf_counter()
    m = f.launch_search(list(B.STATE0), 10, mode="batched", leaves=64, replicas=32, seed=i)
    t1 = time.perf_counter()
    nodes = mn.deepcopy_dict(m)
    t2 = time.perf_counter()
    mt = MasterTree(sims, list(B.DEST), nodes=nodes)
    t3 = time.perf_counter()
    ta = time.perf_counter()
    arrs = mt._arrays()
    tb = time.perf_counter()
    from iboat_pmcts_b200.engine import default_engine
    from iboat_pmcts_b200.master import UCT_COEFF
    out = default_engine().master_policy(arrs[0], arrs[1], arrs[2], arrs[3], mt.probability, UCT_COEFF, 12)
    tc = time.perf_counter()
    mt.get_best_policy(verbose=False)
    t4 = time.perf_counter()
    print("   arrays %.3f kernel-call %.3f full get_best_policy %.3f ms; rows %d nodes %d" % (1e3*(tb-ta), 1e3*(tc-tb), 1e3*(t4-tc), arrs[3].shape[0], arrs[0].size))
    print("search %.3f copy %.3f tree %.3f policy %.3f ms | native %.3f" % (1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t4-t3), f.last_search_report["ms_total"]), mt.best_policy[-1])
