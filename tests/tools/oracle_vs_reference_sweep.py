"""The C oracle against the UNMODIFIED Python reference on all 64 scenarios of BASELINE config 4 (live: needs
/root/reference; CPU only; TEST INFRASTRUCTURE).

tests/golden/ pins the oracle to the reference bit for bit on scenarios 0-2 (rollouts) and 0 (raw steps), and
tests/test_oracle_vs_reference.py on random inputs of scenario 3.  The multi-GPU bench rolls out scenarios 0-63, whose
parity rests on the oracle (tests/tools/fastmodel_sweep.py, the GPU tests), so this script extends the pin: for every
scenario the reference's own ``Tree.default_policy`` (code/solver/worker.py:255-300) runs ROLLOUTS leaves of mixed depth
with an injected noise feed, on the destination / Tmin of bench.py, and the oracle is held to its reward, its
transitions, its arrival flag and fraction, its final state and every position on the way -- all bit for bit.  Raw
``Simulator.doStep`` (simulatorTLKT.py:181-216) is compared the same way on the config-3 axis: 6 boats x 150 steps of
864 s per scenario, every position of every step.

    python tests/tools/oracle_vs_reference_sweep.py [--scenarios 64] [--rollouts 64] > profiles/r02_oracle_vs_reference_sweep.jsonl
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import types
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

STATE0, DEST, TMIN = [0, 44.0, 355.0], [46.77751005, 355.0], 2.022360548788055  # bench.py


def one_scenario(args):
    s, n = args
    import pmcts_helpers as H
    from oracle import gen_golden as G, refshim
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        R = refshim.load_reference()
    simc, wrk = R.simulatorTLKT, R.worker
    sim = G.make_sims(R, [s])[0]
    nT = len(sim.times)
    tree = wrk.Tree(workerid=0, nscenario=1, simulator=sim, destination=list(DEST), TimeMin=TMIN, buffer=[])
    tree.rootNode = types.SimpleNamespace(state=tuple(STATE0))
    actions = [int(a) for a in simc.ACTIONS]
    z = np.random.default_rng(9000 + s).standard_normal((n, nT))
    feeder = G.ZFeeder(z)
    simc.rand = feeder
    prefix = np.zeros((n, 3))
    plen = np.zeros(n, np.int32)
    reward, steps = np.zeros(n), np.zeros(n, np.int32)
    atdest, frac = np.zeros(n, bool), np.full(n, np.nan)
    tl, to = np.full((nT, n), np.nan), np.full((nT, n), np.nan)
    for r in range(n):
        origins = [actions[r % 8]] + [actions[(r + 3) % 8], actions[(5 * r + 1) % 8]][: (r // 8) % 3]
        prefix[r, :len(origins)], plen[r] = origins, len(origins)
        feeder.begin(r)
        log = []
        orig = sim.doStep

        def rec(action, _orig=orig, _log=log):
            st = _orig(action)
            _log.append((st[1], st[2]))
            return st

        sim.doStep = rec
        reward[r] = tree.default_policy(types.SimpleNamespace(origins=origins))
        del sim.doStep
        steps[r] = len(log)
        for j, (x, y) in enumerate(log):
            tl[j, r], to[j, r] = x, y
        res = wrk.Tree.is_state_at_dest(list(DEST), sim.prevState, sim.state)
        atdest[r] = bool(res[0])
        if res[0]:
            frac[r] = res[1]
    o = H.oracle_sim(s)
    got = o.rollout_batch(n, tuple(STATE0), list(DEST), TMIN, prefix=prefix, prefix_len=plen,
                          z=np.ascontiguousarray(z.T), traj_steps=nT)
    same = dict(reward=np.array_equal(got["reward"], reward), steps=np.array_equal(got["steps"], steps),
                arrived=np.array_equal(got["status"] == 1, atdest), frac=np.array_equal(got["frac"], frac, equal_nan=True),
                traj=np.array_equal(got["traj_lat"], tl, equal_nan=True) and np.array_equal(got["traj_lon"], to, equal_nan=True))
    # raw doStep (simulatorTLKT.py:181-216) on the config-3 axis: a few boats of bench.py's fleet, octagon headings
    from iboat_pmcts_b200 import synth
    from oracle import cport
    nb, nsteps = 6, 150
    t, la, lo, u, v = synth.wind_fields(s)
    times3 = np.linspace(0, 5, 501)
    sim3 = simc.Simulator(times3, la, lo, R.weatherTLKT.Weather(la, lo, t, u, v), [0, 44.0, 355.0])
    lat0, lon0, h0 = synth.octagon_fleet(nb, seed=7 + s)
    head = synth.octagon_headings(h0, nsteps)
    z3 = np.random.default_rng(7000 + s).standard_normal((nb, nsteps))
    feeder3 = G.ZFeeder(z3)
    simc.rand = feeder3
    tl3, to3 = np.empty((nsteps, nb)), np.empty((nsteps, nb))
    for b in range(nb):
        feeder3.begin(b)
        sim3.reset([0, lat0[b], lon0[b]])
        for j in range(nsteps):
            st = sim3.doStep(float(head[j // 25, b]))
            tl3[j, b], to3[j, b] = st[1], st[2]
    o3 = cport.OracleSim(t, la, lo, u, v, times3, la, lo)
    r3 = o3.step_batch(0, lat0, lon0, head, nsteps=nsteps, seg_len=25, z=np.ascontiguousarray(z3.T), traj=True)
    same["raw_steps"] = np.array_equal(r3["traj_lat"], tl3) and np.array_equal(r3["traj_lon"], to3)
    import random
    simc.rand = random
    return dict(scenario=s, rollouts=n, transitions=int(steps.sum()), arrived=int(atdest.sum()), raw_steps=nb * nsteps,
                bit_exact={k: bool(v) for k, v in same.items()}, all_bit_exact=bool(all(same.values())))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenarios", type=int, default=64)
    ap.add_argument("--rollouts", type=int, default=64)
    a = ap.parse_args()
    with mp.get_context("spawn").Pool(min(os.cpu_count() or 1, 16)) as pool:
        lines = pool.map(one_scenario, [(s, a.rollouts) for s in range(a.scenarios)], chunksize=1)
    for line in lines:
        print(json.dumps(line))
    tot = dict(scenarios=len(lines), rollouts=sum(l["rollouts"] for l in lines),
               transitions=sum(l["transitions"] for l in lines), arrived=sum(l["arrived"] for l in lines),
               raw_steps=sum(l["raw_steps"] for l in lines),
               scenarios_bit_exact=sum(l["all_bit_exact"] for l in lines))
    print(json.dumps(dict(total=tot)))
    return 0 if tot["scenarios_bit_exact"] == len(lines) else 1


if __name__ == "__main__":
    sys.exit(main())
