"""Decisions of the mixed-precision arithmetic on ALL 64 scenarios of BASELINE config 4, on the CPU.

The 8-GPU bench rolls out scenarios 0-63 (8 per GPU); the GPU parity tests cover scenarios 0-7 at full size.  This script
runs the host build of the kernels' own header (tests/fastmodel: csrc/pmcts_fast.cuh compiled with g++, the "turbo"
variant the bench launches) on every one of the 64 scenarios against the fp64 oracle, same root / destination / Tmin as
bench.py, same injected normals for both, and counts per scenario: rollouts the fp32 path leaves to the fp64 finisher
(inside a decision margin) and rollouts it decides UNLIKE the oracle (transitions, status or histogram bin) -- the second
count must be zero for the histograms of the search to be the oracle's.  One JSON line per scenario + a total.

    python scripts/fastmodel_sweep.py [--scenarios 64] [--leaves 512] [--replicas 256] > profiles/r02_fastmodel_sweep.jsonl

TEST INFRASTRUCTURE (loads oracle/ and tests/fastmodel); nothing of the product runs here.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fastmodel  # noqa: E402
import pmcts_helpers as H  # noqa: E402
from iboat_pmcts_b200 import synth  # noqa: E402

STATE0, DEST, TMIN = (0, 44.0, 355.0), [46.77751005, 355.0], 2.022360548788055  # bench.py


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenarios", type=int, default=64)
    ap.add_argument("--leaves", type=int, default=512)
    ap.add_argument("--replicas", type=int, default=256)
    ap.add_argument("--seed", type=int, default=2026)
    a = ap.parse_args()
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
        z = rng.standard_normal((times.size, n))
        got = f.rollout_batch(n_leaf, rep, STATE0, DEST, TMIN, prefix=prefix, prefix_len=plen, z=z)
        assert f.last_variant() == 2, "not the turbo variant"
        want = o.rollout_batch(n, STATE0, DEST, TMIN, prefix=np.repeat(prefix, rep, 0), prefix_len=np.repeat(plen, rep), z=z,
                               nthreads=os.cpu_count() or 8)
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
    tot.update(scenarios=a.scenarios, undecided_share=tot["undecided"] / tot["rollouts"], seconds=round(time.time() - t0, 1),
               note="host build of csrc/pmcts_fast.cuh (turbo variant) against oracle/pmcts_oracle.c, injected normals, "
                    "root / destination / Tmin of bench.py; 'differently' counts rollouts the fp32 path decided itself")
    print(json.dumps(dict(total=tot)), flush=True)
    return 0 if tot["differently"] == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
