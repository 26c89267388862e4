"""Quick device-resident timing of K1 / K2 (not the bench; used while optimising)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np, torch
from iboat_pmcts_b200 import synth, _native as N
from iboat_pmcts_b200.engine import Engine
import pmcts_helpers as H

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    modes = [int(m) for m in sys.argv[2].split(",")] if len(sys.argv) > 2 else [N.PREC_F64, N.PREC_FAST]
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    nsteps = 500
    dev = torch.device("cuda:0")
    eng = Engine(0)
    eng.use_torch_stream()
    t, la, lo, u, v = synth.wind_fields(0)
    times = np.linspace(0, 5, nsteps + 1)
    eng.upload_scenario(0, t, la, lo, u, v, times, [la[0], la[-1], lo[0], lo[-1]])
    times2, lats, lons = H.sim_axes(la, lo)
    eng.upload_scenario(1, t, la, lo, u, v, times2, [lats[0], lats[-1], lons[0], lons[-1]])
    lat0, lon0, h0 = synth.octagon_fleet(n)
    head = torch.tensor(synth.octagon_headings(h0, nsteps), device=dev)
    for mode in modes:
        eng.set_precision(mode)
        for it in range(reps):
            k = torch.zeros(n, dtype=torch.int32, device=dev)
            lat = torch.tensor(lat0, device=dev); lon = torch.tensor(lon0, device=dev)
            st = torch.zeros(n, dtype=torch.uint8, device=dev)
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.step_batch(0, k, lat, lon, head.reshape(-1), nsteps=nsteps, seg_len=25, status=st, seed=42)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
        ok = int((st == 0).sum())
        print(f"K1 mode={mode} n={n} steps={nsteps}: {ms:.2f} ms -> {n*nsteps/ms*1e3:.3e} transitions/s (ok boats {ok})")
        # rollouts
        n_leaf, rep = 4096, 256
        nr = n_leaf * rep
        prefix = torch.tensor(45.0 * (np.arange(n_leaf) % 8), device=dev)
        plen = torch.ones(n_leaf, dtype=torch.int32, device=dev)
        out = dict(steps=torch.empty(nr, dtype=torch.int32, device=dev),
                   leaf_bins=torch.zeros(n_leaf * 12, dtype=torch.int32, device=dev))
        for it in range(reps):
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.rollout_batch(1, n_leaf, rep, H.STATE0, [46.77751005, 355.0], 2.0223605, prefix, plen, 1, seed=1, out=out)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
        tr = int(out["steps"].sum())
        print(f"   redo count {eng.last_redo_count()} of {nr}")
        print(f"K2 mode={mode} rollouts={nr}: {ms:.2f} ms -> {nr/ms*1e3:.3e} rollouts/s, {tr/ms*1e3:.3e} transitions/s ({tr/nr:.2f} steps/rollout)")

main()
