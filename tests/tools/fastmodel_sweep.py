"""Decisions of the mixed-precision arithmetic on ALL 64 scenarios of BASELINE config 4, on the CPU.

The 8-GPU bench rolls out scenarios 0-63 (8 per GPU); the GPU parity tests cover scenarios 0-7 at full size.  This script
runs the host build of the kernels' own header (tests/fastmodel: csrc/pmcts_fast.cuh compiled with g++, the "turbo"
variant the bench launches) on every one of the 64 scenarios against the fp64 oracle, same root / destination / Tmin as
bench.py, same injected normals for both, and counts per scenario: rollouts the fp32 path leaves to the fp64 finisher
(inside a decision margin) and rollouts it decides UNLIKE the oracle (transitions, status or histogram bin) -- the second
count must be zero for the histograms of the search to be the oracle's.  One JSON line per scenario + a total.

    python tests/tools/fastmodel_sweep.py [--scenarios 64] [--leaves 512] [--replicas 256] > profiles/r02_fastmodel_sweep.jsonl

``--k1 BOATS`` does the same for BASELINE config 3 (raw doStep: bench.py's fleet of BOATS boats walking octagons for 500
steps of 864 s on scenario 0), in chunks of 10^5 boats: boats handed to the literal kernel, status / step count equal to
the oracle's, worst final position.

TEST INFRASTRUCTURE (loads oracle/ and tests/fastmodel); nothing of the product runs here.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fastmodel  # noqa: E402
import pmcts_helpers as H  # noqa: E402
from iboat_pmcts_b200 import synth  # noqa: E402

STATE0, DEST, TMIN = (0, 44.0, 355.0), [46.77751005, 355.0], 2.022360548788055  # bench.py


def k1(boats, seed):
    t, la, lo, u, v = synth.wind_fields(0)
    times = np.linspace(0, 5, 501)
    f = fastmodel.FastModel(t, la, lo, u, v, times, la, lo)
    from oracle import cport
    o = cport.OracleSim(t, la, lo, u, v, times, la, lo)
    lat0, lon0, h0 = synth.octagon_fleet(boats, seed=7)
    head = synth.octagon_headings(h0, 500)
    tot = dict(boats=0, handed_over=0, status_or_steps_differ=0, transitions=0, worst_position_rel=0.0)
    t0 = time.time()
    for c0 in range(0, boats, 100000):
        c1 = min(boats, c0 + 100000)
        z = np.random.default_rng(seed + c0).standard_normal((500, c1 - c0))
        hd = np.ascontiguousarray(head[:, c0:c1])
        got = f.step_batch(0, lat0[c0:c1], lon0[c0:c1], hd, nsteps=500, seg_len=25, z=z)
        want = o.step_batch(0, lat0[c0:c1], lon0[c0:c1], hd, nsteps=500, seg_len=25, z=z, nthreads=os.cpu_count() or 8)
        mine = got["needs_literal"] == 0
        differ = mine & ((got["status"] != want["status"]) | (got["k"] != want["k"]))
        ok = mine & (want["status"] == 0)
        rel = max(float(np.max(np.abs(got[k][ok] - want[k][ok]) / np.abs(want[k][ok]))) for k in ("lat", "lon"))
        line = dict(chunk=[c0, c1], boats=c1 - c0, handed_over=int((~mine).sum()), status_or_steps_differ=int(differ.sum()),
                    transitions=int(want["k"].sum()), worst_position_rel=rel)
        for k in ("boats", "handed_over", "status_or_steps_differ", "transitions"):
            tot[k] += line[k]
        tot["worst_position_rel"] = max(tot["worst_position_rel"], rel)
        print(json.dumps(line), flush=True)
    tot.update(seconds=round(time.time() - t0, 1),
               note="config 3 on the host build of csrc/pmcts_fast.cuh (K1 arithmetic) against oracle/pmcts_oracle.c, "
                    "injected normals; positions after 500 fused steps")
    print(json.dumps(dict(total=tot)), flush=True)
    return 0 if tot["status_or_steps_differ"] == 0 else 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k1", type=int, default=0)
    ap.add_argument("--scenarios", type=int, default=64)
    ap.add_argument("--leaves", type=int, default=512)
    ap.add_argument("--replicas", type=int, default=256)
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--philox", action="store_true",
                    help="noise from Philox counters on both sides (the bench's mode): the model draws its normals with "
                         "the kernels' fp32 Box-Muller, the oracle with its fp64 one")
    ap.add_argument("--mode", choices=("tree", "full-prefix", "plan"), default="tree",
                    help="tree: the reference's replay of the first action only (quirk Q1); full-prefix: "
                         "PMCTS_ROLL_REPLAY_FULL_PREFIX; plan: plan evaluation (PMCTS_ROLL_PLAN_EVAL, Tmin 0)")
    a = ap.parse_args()
    if a.k1:
        return k1(a.k1, a.seed)
    n_leaf, rep = a.leaves, a.replicas
    n = n_leaf * rep
    tot = dict(rollouts=0, undecided=0, differently=0, transitions=0, arrived=0)
    t0 = time.time()
    for s in range(a.scenarios):
        t, la, lo, u, v = synth.wind_fields(s)
        times, lats, lons = H.sim_axes(la, lo)
        f = fastmodel.FastModel(t, la, lo, u, v, times, lats, lons)
        o = H.oracle_sim(s)
        rng = np.random.default_rng(a.seed + s)
        prefix = 45.0 * rng.integers(0, 8, (n_leaf, 4))
        plen = rng.integers(1, 5, n_leaf).astype(np.int32)
        if a.philox:
            noise = dict(seed=a.seed, stream=100 + s)
        else:
            noise = dict(z=rng.standard_normal((times.size, n)))
        flags = {"tree": 0, "full-prefix": 2, "plan": 4}[a.mode]
        tmin = 0.0 if a.mode == "plan" else TMIN
        got = f.rollout_batch(n_leaf, rep, STATE0, DEST, tmin, prefix=prefix, prefix_len=plen, flags=flags, **noise)
        assert f.last_variant() == (1 if a.mode == "plan" else 2), "not the variant the product launches"
        want = o.rollout_batch(n, STATE0, DEST, tmin, prefix=np.repeat(prefix, rep, 0), prefix_len=np.repeat(plen, rep),
                               replay_full=a.mode == "full-prefix", mode=1 if a.mode == "plan" else 0,
                               nthreads=os.cpu_count() or 8, **noise)
        decided = got["uncertain"] == 0
        bad = (got["steps"] != want["steps"]) | (got["status"] != want["status"]) | (got["bin"] != want["bin"])
        rel = max(float(np.max(np.abs(got[k][decided] - want[k][decided]) / np.abs(want[k][decided])))
                  for k in ("final_lat", "final_lon"))
        line = dict(scenario=s, rollouts=n, undecided=int((~decided).sum()), differently=int((bad & decided).sum()),
                    transitions=int(want["steps"].sum()), arrived=int((want["status"] == 1).sum()),
                    worst_position_rel=rel)
        for k in tot:
            tot[k] += line[k]
        print(json.dumps(line), flush=True)
    tot.update(scenarios=a.scenarios, mode=a.mode, undecided_share=tot["undecided"] / tot["rollouts"], seconds=round(time.time() - t0, 1),
               noise="philox (fp32 Box-Muller in the model, fp64 in the oracle)" if a.philox else "injected normals",
               note="host build of csrc/pmcts_fast.cuh (the variant the product launches in this mode) against oracle/pmcts_oracle.c, "
                    "root / destination / Tmin of bench.py; 'differently' counts rollouts the fp32 path decided itself")
    print(json.dumps(dict(total=tot)), flush=True)
    return 0 if tot["differently"] == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
