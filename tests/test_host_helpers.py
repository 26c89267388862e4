"""The scalar helpers of the API that stay on the host, as in the reference: geometry of
``Simulator``, ``Boat.getDeterDyn``, ``Weather.returnPolarVel``, ``Tree.is_state_at_dest / is_state_terminal``, the axes
``create_simulators`` builds.  Bit for bit against the KATs recorded from the reference (tests/golden/kat.json) and, where
the reference is mounted, against the reference itself on random inputs.  No GPU: a mirror ``Simulator`` only touches the
device when it steps or interpolates."""
import json
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from iboat_pmcts_b200 import simulatorTLKT as SimC, synth
from iboat_pmcts_b200 import worker
from iboat_pmcts_b200.forest import create_simulators
from iboat_pmcts_b200.weatherTLKT import Weather
from oracle import refshim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sim(scenario=0, **kw):
    t, la, lo, u, v = synth.wind_fields(scenario)
    args = dict(simtimestep=6, stateinit=(0, 44.0, 355.0), ndaysim=3)
    args.update(kw)
    return create_simulators([Weather(la, lo, t + 5.0, u, v)], 1, **args)[0]


def test_scalar_helpers_reproduce_the_recorded_reference_values(golden_dir):
    with open(os.path.join(golden_dir, "kat.json")) as f:
        kat = json.load(f)
    sim = _sim()
    for d, b, la, lo, want in kat["getDestination"]:
        assert sim.getDestination(d, b, [la, lo]) == want
    for a, b, want in kat["getDistAndBearing"]:
        assert list(sim.getDistAndBearing(a, b)) == want
    for p, w, want in kat["getDeterDyn"]:
        assert float(SimC.Boat.getDeterDyn(p, w, SimC.Boat.FIT_VELOCITY)) == want
    for la, lo, want in kat["fromGeoToCartesian"]:
        assert list(SimC.Simulator.fromGeoToCartesian([la, lo])) == want
    for u, v, want in kat["returnPolarVel"]:
        assert list(Weather.returnPolarVel(u, v)) == want
    for dest, a, b, hit, frac in kat["is_state_at_dest"]:
        got = worker.Tree.is_state_at_dest(dest, [0] + a, [1] + b)
        assert got[0] == hit and (got[1] is None if frac is None else float(got[1]) == frac)
    # the terminal test: last instant, or outside [lats[0], lats[-1]] x [lons[0], lons[-1]] (worker.py:417-436, quirk Q8)
    assert not worker.Tree.is_state_terminal(sim, [3, 44.0, 355.0])
    assert worker.Tree.is_state_terminal(sim, [len(sim.times) - 1, 44.0, 355.0])
    assert worker.Tree.is_state_terminal(sim, [3, float(sim.lats[-1]) + 1e-9, 355.0])
    assert worker.Tree.is_state_terminal(sim, [3, 44.0, float(sim.lons[0]) - 1e-9])
    assert not worker.Tree.is_state_terminal(sim, [3, float(sim.lats[-1]), float(sim.lons[0])])


@pytest.mark.skipif(not refshim.reference_available(), reason="reference not mounted")
def test_scalar_helpers_and_simulator_axes_against_the_live_reference(tmp_path):
    script = textwrap.dedent("""
        import json, sys
        import numpy as np
        sys.path.insert(0, %r)
        from oracle import refshim
        from iboat_pmcts_b200 import synth
        R = refshim.load_reference()
        simc, wrk = R.simulatorTLKT, R.worker
        out = dict(axes=[], dest=[], brg=[], dyn=[], cart=[], at=[], polar=[])
        t, la, lo, u, v = synth.wind_fields(0)
        for kw in (dict(simtimestep=6, ndaysim=3), dict(simtimestep=3, ndaysim=8), dict(simtimestep=1, ndaysim=2, deltalatlon=0.25)):
            w = R.weatherTLKT.Weather(la, lo, t + 5.0, u, v)
            s = R.forest.create_simulators([w], 1, stateinit=(0, 44.0, 355.0), **kw)[0]
            out["axes"].append([list(map(float, s.times)), list(map(float, s.lats)), list(map(float, s.lons)), list(map(float, w.time))])
        rng = np.random.default_rng(11)
        B = simc.Boat
        for _ in range(400):
            a = [float(rng.uniform(-70, 70)), float(rng.uniform(0, 360))]
            b = [a[0] + float(rng.uniform(-3, 3)), a[1] + float(rng.uniform(-4, 4))]
            d, brg = float(rng.uniform(-2e4, 2e5)), float(rng.uniform(0, 360))
            p, wm = float(rng.uniform(0, 360)), float(rng.uniform(0, 30))
            out["dest"].append([d, brg, a, s.getDestination(d, brg, a)])
            out["brg"].append([a, b, list(s.getDistAndBearing(a, b))])
            out["dyn"].append([p, wm, float(B.getDeterDyn(p, wm, B.FIT_VELOCITY))])
            out["cart"].append([a, list(simc.Simulator.fromGeoToCartesian(a))])
            uu, vv = float(rng.normal(0, 8)), float(rng.normal(0, 8))
            out["polar"].append([uu, vv, list(R.weatherTLKT.Weather.returnPolarVel(uu, vv))])
            dst = [46.7 + float(rng.normal(0, 0.2)), 355.0 + float(rng.normal(0, 0.2))]
            pa = [0, dst[0] - float(rng.uniform(0.05, 0.5)), dst[1] + float(rng.normal(0, 0.1))]
            bearing = s.getDistAndBearing(pa[1:], dst)[1] + float(rng.normal(0, 0.15))
            pb = [1] + s.getDestination(float(rng.uniform(1e4, 9e4)), bearing, pa[1:])
            r = wrk.Tree.is_state_at_dest(dst, pa, pb)
            out["at"].append([dst, pa, pb, bool(r[0]), None if r[1] is None else float(r[1])])
        print(json.dumps(out))
        """) % ROOT
    res = subprocess.run([sys.executable, "-W", "ignore", "-c", script], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                         text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    want = json.loads(res.stdout.strip().splitlines()[-1])
    for kw, (times, lats, lons, wtime) in zip((dict(simtimestep=6, ndaysim=3), dict(simtimestep=3, ndaysim=8),
                                               dict(simtimestep=1, ndaysim=2, deltalatlon=0.25)), want["axes"]):
        t, la, lo, u, v = synth.wind_fields(0)
        w = Weather(la, lo, t + 5.0, u, v)
        s = create_simulators([w], 1, stateinit=(0, 44.0, 355.0), **kw)[0]
        assert list(map(float, s.times)) == times and list(map(float, s.lats)) == lats and list(map(float, s.lons)) == lons
        assert list(map(float, w.time)) == wtime and w.time[0] == 0.0   # shifted in place, like the reference (forest.py:179)
    sim = _sim()
    for d, brg, a, dest in want["dest"]:
        assert sim.getDestination(d, brg, a) == dest
    for a, b, r in want["brg"]:
        assert list(sim.getDistAndBearing(a, b)) == r
    for p, wm, r in want["dyn"]:
        assert float(SimC.Boat.getDeterDyn(p, wm, SimC.Boat.FIT_VELOCITY)) == r
    for a, r in want["cart"]:
        assert list(SimC.Simulator.fromGeoToCartesian(a)) == r
    for uu, vv, r in want["polar"]:
        assert list(Weather.returnPolarVel(uu, vv)) == r
    hits = 0
    for dst, pa, pb, hit, frac in want["at"]:
        got = worker.Tree.is_state_at_dest(dst, pa, pb)
        assert got[0] == hit and (got[1] is None if frac is None else float(got[1]) == frac)
        hits += hit
    assert 10 < hits < 390
