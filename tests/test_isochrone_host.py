"""N2 on the CPU: the host logic of ``isochrones.Isochrone`` (sectors, fronts, pruning, the last leg, the plan read back
from the parents -- isochrones.py:24-464) against the run recorded from the reference (tests/golden/isochrone_c1.json,
Tutorial.py:81-92).  The batched ``doStep`` launches the solver issues are answered here by the fp64 oracle (bit-exact
with the reference's ``doStep``), so the fronts must come out EQUAL, not close; on the GPU the same test runs with the
literal kernel in their place (tests/test_api_gpu.py::test_isochrone_solver_matches_the_reference)."""
import json
import os
import random

import numpy as np

import pmcts_helpers as H
from iboat_pmcts_b200 import isochrones as iso
from iboat_pmcts_b200 import simulatorTLKT as SimC
from iboat_pmcts_b200 import synth
from iboat_pmcts_b200.forest import create_simulators
from iboat_pmcts_b200.weatherTLKT import Weather
from oracle_engine import on_oracle


def test_isochrone_fronts_are_the_references(golden_dir, monkeypatch):
    with open(os.path.join(golden_dir, "isochrone_c1.json")) as f:
        want = json.load(f)
    t, la, lo, u, v = synth.wind_fields(0)
    sim = create_simulators([Weather(la, lo, t, u, v)], 1, simtimestep=6, stateinit=H.STATE0, ndaysim=3)[0]
    monkeypatch.setattr(SimC.Boat, "UNCERTAINTY_COEFF", 0)   # Tutorial.py:81
    fake = on_oracle([sim], [H.oracle_sim(0)], monkeypatch)
    random.seed(3)
    solver = iso.Isochrone(sim, H.STATE0, want["destination"], delta_cap=5, increment_cap=9, nb_secteur=300, resolution=100)
    temps, actions, caps_fin, positions = solver.isochrone_methode()
    assert solver.C0 == want["C0"]
    assert [len(f) for f in solver.isochrone_stock] == [len(f) for f in want["isochrone_stock"]]
    for got_front, want_front in zip(solver.isochrone_stock, want["isochrone_stock"]):
        assert np.array_equal(np.array(got_front, dtype=np.float64), np.array(want_front, dtype=np.float64))
    assert temps == want["temps_transit"]
    assert list(map(float, actions)) == want["liste_actions"]
    assert list(map(float, caps_fin)) == want["liste_caps_fin"]
    assert np.array_equal(np.array(positions, dtype=np.float64), np.array(want["liste_positions"], dtype=np.float64))
    # one launch per isochrone for the fronts; only the last leg steps boat by boat
    fronts = len(want["isochrone_stock"])
    assert fronts <= fake.launches < fronts + 2000
