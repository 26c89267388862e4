"""The fp64 finisher (rollout_batch_kernel on the work list of the mixed-precision launch) and the literal fp64
rollout kernel on their own: per-launch durations from the library's own event pairs, launches NOT overlapped.

  python tests/tools/finisher_probe.py            # (PMCTS_B200_LIB=<other build> for A/B comparisons)

Three shapes: the c4 shard (8 scenarios x 4096 leaves x 256 replicas), one c2 search wave (10 scenarios x 64 leaves
x 32 replicas), and a literal-mode launch of 131 072 rollouts.  Also prints a digest of the histograms so that two
builds can be compared for equal results.
"""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import pmcts_helpers as H  # noqa: E402
from iboat_pmcts_b200 import _native as N, synth  # noqa: E402
from iboat_pmcts_b200.engine import Engine  # noqa: E402

DEST, TMIN = [46.77751005, 355.0], 2.0223605


def main():
    dev = torch.device("cuda:0")
    eng = Engine(0)
    n_scen = 10
    for s in range(n_scen):
        t, la, lo, u, v = synth.wind_fields(s)
        times, lats, lons = H.sim_axes(la, lo)
        eng.upload_scenario(s, t, la, lo, u.astype(np.float32), v.astype(np.float32), times,
                            [lats[0], lats[-1], lons[0], lons[-1]])
    print("library:", N.LIB_PATH)
    shapes = [("c4 shard", N.PREC_FAST, 8, 4096, 256), ("c2 wave", N.PREC_FAST, 10, 64, 32),
              ("literal", N.PREC_F64, 1, 4096, 32)]
    for name, prec, S, n_leaf, rep in shapes:
        eng.set_precision(prec)
        rng = np.random.default_rng(7)
        prefix = torch.tensor(45.0 * rng.integers(0, 8, (n_leaf, 4)), device=dev).reshape(-1)
        plen = torch.tensor(rng.integers(1, 5, n_leaf).astype(np.int32), device=dev)
        bins = torch.zeros(S * n_leaf * 12, dtype=torch.int32, device=dev)
        stats = torch.zeros(S * 8, dtype=torch.float64, device=dev)
        reps = 6
        eng.enable_timing(False)
        for it in range(reps + 2):
            if it == 2:
                eng.synchronize()
                eng.enable_timing(True)
                eng.kernel_ms(reset=True)
            bins.zero_()
            torch.cuda.synchronize()
            eng.rollout_forest(list(range(S)), n_leaf, rep, H.STATE0, DEST, TMIN, prefix, plen, 4, prefix_shared=True,
                               seed=11, stream=3, out=dict(leaf_bins=bins, stats=stats))
            eng.synchronize()
        main_ms = eng.kernel_ms(reset=False, kind=N.KERNEL_ROLLOUT) / reps
        fin_ms = eng.kernel_ms(reset=True, kind=N.KERNEL_FINISHER) / reps
        eng.enable_timing(False)
        redo = eng.last_redo_count() if prec == N.PREC_FAST else 0
        digest = hashlib.sha1(bins.cpu().numpy().tobytes()).hexdigest()[:12]
        st = stats.cpu().numpy().reshape(S, 8)
        print("%-9s %2d x %4d x %3d rollouts: main kernel %.4f ms, fp64 finisher %.4f ms (%d rollouts re-run), "
              "arrived %d transitions %d bins %s" % (name, S, n_leaf, rep, main_ms, fin_ms, redo, int(st[:, 3].sum()),
                                                     int(st[:, 4].sum()), digest))


main()
