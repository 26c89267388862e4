"""Live check of the C restatement (oracle/) against the UNMODIFIED Python reference, on random inputs.
Runs only where the reference is mounted (/root/reference: the build container); the committed fixtures
under tests/golden/ carry the same guarantee to the GPU box.  Bit for bit unless stated."""
import random

import numpy as np
import pytest

from oracle import refshim

pytestmark = pytest.mark.skipif(not refshim.reference_available(), reason="reference not mounted")


@pytest.fixture(scope="module")
def world():
    import warnings
    from iboat_pmcts_b200 import synth
    import pmcts_helpers as H
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        R = refshim.load_reference()
    t, la, lo, u, v = synth.wind_fields(3)
    w = R.weatherTLKT.Weather(la, lo, t, u, v)
    sim = R.forest.create_simulators([w], 1, simtimestep=6, stateinit=(0, 44.0, 355.0), ndaysim=3)[0]
    return R, sim, H.oracle_sim(3)


def test_random_transitions_bit_exact(world):
    R, sim, o = world
    rng = np.random.default_rng(0)
    random.seed(5)
    for _ in range(300):
        k = int(rng.integers(0, 12))
        lat, lon, action = rng.uniform(41, 49), rng.uniform(346, 359), float(rng.choice(R.simulatorTLKT.ACTIONS))
        state = random.getstate()
        z = random.gauss(0.0, 1.0)
        random.setstate(state)
        sim.reset([k, lat, lon])
        got = list(sim.doStep(action))
        want = o.step_batch(k, [lat], [lon], [[action]], z=[[z]])
        assert got == [int(want["k"][0]), float(want["lat"][0]), float(want["lon"][0])]


def test_scalar_functions_bit_exact(world):
    R, sim, o = world
    rng = np.random.default_rng(1)
    B = R.simulatorTLKT.Boat
    for _ in range(500):
        p, w = rng.uniform(0, 360), rng.uniform(0, 25)
        assert float(B.getDeterDyn(p, w, B.FIT_VELOCITY)) == o.deter_dyn(p, w)
        a, b = [rng.uniform(40, 50), rng.uniform(345, 360)], [rng.uniform(40, 50), rng.uniform(345, 360)]
        assert list(sim.getDistAndBearing(a, b)) == list(o.dist_bearing(a[0], a[1], b[0], b[1]))
        d, brg = rng.uniform(-1e4, 1e5), rng.uniform(0, 360)
        assert sim.getDestination(d, brg, a) == o.get_destination(d, brg, a[0], a[1])
        t = rng.uniform(0, 5)
        assert (float(sim.uWindAvg([t, a[0], a[1]])[0]), float(sim.vWindAvg([t, a[0], a[1]])[0])) == o.interp(t, a[0], a[1])


def test_arrival_test_and_rollout_bit_exact(world):
    R, sim, o = world
    Tree = R.worker.Tree
    rng = np.random.default_rng(2)
    dest = [46.77751004666887, 355.0]
    hits = 0
    for _ in range(300):
        a = [0, rng.uniform(46.3, 46.8), rng.uniform(354.8, 355.2)]
        brg = sim.getDistAndBearing(a[1:], dest)[1] + rng.normal(0, 0.2)
        b = [1] + sim.getDestination(rng.uniform(2e4, 6e4), brg, a[1:])
        got = Tree.is_state_at_dest(dest, a, b)
        want = o.is_at_dest(dest, a[1:], b[1:])
        assert got[0] == want[0] and (got[1] is None and want[1] is None or float(got[1]) == want[1])
        hits += bool(got[0])
    assert 5 < hits < 295


@pytest.mark.parametrize("scenario", [20, 41, 63])
def test_rollouts_and_raw_steps_of_other_scenarios_bit_exact(scenario):
    """A slice of tests/tools/oracle_vs_reference_sweep.py (which holds the oracle to the reference on ALL 64 scenarios of
    config 4: profiles/r02_oracle_vs_reference_sweep.jsonl): the reference's own Tree.default_policy on 24 leaves of
    mixed depth and 6 x 150 raw doStep on the config-3 axis, rewards / arrival fractions / every position equal."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools", "oracle_vs_reference_sweep.py")
    spec = importlib.util.spec_from_file_location("oracle_vs_reference_sweep", path)
    sweep = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sweep)
    line = sweep.one_scenario((scenario, 24))
    assert line["all_bit_exact"], line
    assert line["transitions"] > 24 * 6 and line["raw_steps"] == 900
