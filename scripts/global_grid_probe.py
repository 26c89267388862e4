"""K1 on a grid that does not fit L1 (SURVEY.md 8d "large-grid variant"): near-global 0.5-degree, 3-hourly wind
(301 x 720 x 41 points; 2.1 MB per simulator instant after the time blend), boats spread over the whole of it, so
the four corner gathers of a step come from L2 rather than L1.  Prints the throughput.

  python scripts/global_grid_probe.py [n_boats] [nsteps]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

from iboat_pmcts_b200 import _native as N
from iboat_pmcts_b200 import synth
from iboat_pmcts_b200.engine import Engine


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    dev = torch.device("cuda:0")
    t, la, lo, u, v = synth.wind_fields(0, lat=(-75.0, 75.0, 0.5), lon=(0.0, 359.5, 0.5))
    times = np.linspace(0, 1, nsteps + 1)  # 864 s steps when nsteps = 100
    eng = Engine(0)
    eng.use_torch_stream()
    eng.set_precision(N.PREC_FAST)
    t0 = time.perf_counter()
    eng.upload_scenario(0, t, la, lo, u.astype(np.float32), v.astype(np.float32), times, [la[0], la[-1], lo[0], lo[-1]])
    torch.cuda.synchronize()
    print("grid %d x %d x %d, %d simulator instants: upload + time blend %.2f s, fp32 slabs %.0f MB"
          % (t.size, la.size, lo.size, times.size, time.perf_counter() - t0, times.size * la.size * lo.size * 8 / 1e6))
    rng = np.random.default_rng(11)
    lat0 = rng.uniform(-60.0, 60.0, n)
    lon0 = rng.uniform(10.0, 350.0, n)
    nseg = (nsteps + 24) // 25
    head_h = 45.0 * rng.integers(0, 8, (nseg, n)).astype(np.float64)
    head = torch.tensor(head_h, device=dev).reshape(-1)
    lat0_d, lon0_d = torch.tensor(lat0, device=dev), torch.tensor(lon0, device=dev)
    k = torch.zeros(n, dtype=torch.int32, device=dev)
    lat, lon = lat0_d.clone(), lon0_d.clone()
    st = torch.zeros(n, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    best = None
    for it in range(5):
        k.zero_(); st.zero_(); lat.copy_(lat0_d); lon.copy_(lon0_d)
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.step_batch(0, k, lat, lon, head, nsteps=nsteps, seg_len=25, status=st, seed=42, stream=0)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None or ms < best else best
    tr = int(k.sum().item())
    print("K1 variant %d: %d boats x %d steps: %.3f ms -> %.3e transitions/s (%d boats stopped early)"
          % (eng.last_kernel_variant(), n, nsteps, best, tr / (best * 1e-3), int((st != 0).sum().item())))
    print("algorithmic gather: %.1f GB per launch (4 corners x 8 B per transition) -> %.2f TB/s"
          % (tr * 32 / 1e9, tr * 32 / (best * 1e-3) / 1e12))
    # (parity on a grid of this kind: tests/test_gpu_parity.py::test_large_grid_scattered_boats)

if __name__ == "__main__":
    main()
