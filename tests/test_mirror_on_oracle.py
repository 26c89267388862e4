"""The host logic of the mirror classes against the fixtures recorded from the reference, on the CPU: every launch the
classes issue is answered by the fp64 oracle (tests/oracle_engine.py), which is bit-exact with the reference's own
``doStep`` / ``default_policy`` -- so what these tests pin is everything AROUND the launches (the order in which
``random.gauss`` is consumed, the bookkeeping of ``Simulator`` states, the means / minima / variances, the exceptions, the
search loop), and they pin it EXACTLY.  The same scenarios run on the GPU with the kernels in the oracle's place
(tests/test_api_gpu.py, tolerances 1e-11)."""
import json
import os
import random

import numpy as np
import pytest

import pmcts_helpers as H
from iboat_pmcts_b200 import forest as ft
from iboat_pmcts_b200 import isochrones as iso
from iboat_pmcts_b200 import master_node as mn
from iboat_pmcts_b200 import synth, worker
from iboat_pmcts_b200.weatherTLKT import Weather
from oracle_engine import on_oracle

STATE0 = list(H.STATE0)


class Feeder:
    """Stands in for the ``random`` module: gauss(mu, s) = mu + z[row, j] * s, one row per trajectory (a row starts at
    every getstate(), which the mirror calls once per trajectory)."""

    def __init__(self, z):
        self.z, self.i, self.j = np.asarray(z), -1, 0

    def getstate(self):
        self.i += 1
        self.j = 0
        return (self.i, 0)

    def setstate(self, st):
        self.i, self.j = st

    def gauss(self, mu, sigma):
        zz = float(self.z[self.i, self.j])
        self.j += 1
        return mu + zz * sigma


def make(scenarios, monkeypatch):
    ws = []
    for s in scenarios:
        t, la, lo, u, v = synth.wind_fields(s)
        ws.append(Weather(la, lo, t, u, v))
    sims = ft.create_simulators(ws, len(ws), simtimestep=6, stateinit=STATE0, ndaysim=3)
    eng = on_oracle(sims, [H.oracle_sim(s) for s in scenarios], monkeypatch)
    return sims, eng


def test_initialize_simulators_and_plan_evaluation_are_the_references(golden_dir, monkeypatch):
    """forest.initialize_simulators (forest.py:210-331) and isochrones.estimate_perfomance_plan (isochrones.py:469-580)
    in the reference's noise order: destination, Tmin, means and variances EQUAL to the recorded ones."""
    g = np.load(os.path.join(golden_dir, "rollouts_c1.npz"))
    pe = np.load(os.path.join(golden_dir, "plan_eval_c5.npz"))
    sims, eng = make([int(s) for s in g["scenarios"]], monkeypatch)
    assert np.array_equal(sims[0].times, g["times"]) and np.array_equal(sims[0].lats, g["lats"])
    monkeypatch.setattr(ft, "rand", Feeder(g["zinit"]))
    destination, timemin = ft.initialize_simulators(sims, int(g["ntra"]), STATE0, 0)
    assert list(destination) == g["destination"].tolist() and timemin == float(g["timemin"])
    assert eng.launches == 2 * len(sims) * int(g["ntra"])
    for name in ("plan", "empty"):
        monkeypatch.setattr(ft, "rand", Feeder(pe[name + "_z"]))
        mean, var = iso.estimate_perfomance_plan(sims, int(pe[name + "_ntra"]), STATE0, g["destination"].tolist(),
                                                 plan=pe[name + "_plan"].tolist(), verbose=False)
        assert [float(x) for x in mean] == pe[name + "_mean"].tolist()
        assert [float(x) for x in var] == pe[name + "_var"].tolist()


def test_seeded_uct_search_is_the_references(golden_dir, monkeypatch):
    """random.seed(1); np.random.seed(1); Tree.uct_search (worker.py:147-186) with the rollouts of Tree.default_policy
    answered by the oracle: the leaves in the reference's order, its rewards bit for bit, its histograms, its master --
    and the generators left where the reference's run leaves them (the draws of every transition consumed)."""
    with open(os.path.join(golden_dir, "tree_replay_c1.json")) as f:
        replay = json.load(f)
    with open(os.path.join(golden_dir, "tree_replay_c1_steps.json")) as f:
        steps = json.load(f)["steps"]
    g = np.load(os.path.join(golden_dir, "rollouts_c1.npz"))
    sims, eng = make([0], monkeypatch)
    random.seed(1)
    np.random.seed(1)
    monkeypatch.setattr(worker, "RHO", replay["rho"])
    monkeypatch.setattr(worker, "UCT_COEFF", replay["uct_coeff"])
    tree = worker.Tree(workerid=0, nscenario=1, budget=replay["budget"], simulator=sims[0],
                       destination=[float(x) for x in g["destination"]], TimeMin=float(g["timemin"]))
    master = {ft.ROOT_HASH: mn.MasterNode(1, nodehash=ft.ROOT_HASH)}
    seen = []
    orig = worker.Node.back_up

    def rec(self, reward):
        seen.append(([int(a) for a in self.origins], float(reward)))
        return orig(self, reward)

    monkeypatch.setattr(worker.Node, "back_up", rec)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        tree.uct_search(STATE0, replay["frequency"], master)
    assert [s[0] for s in seen] == [it["origins"] for it in replay["iterations"]]
    assert [s[1] for s in seen] == [it["reward"] for it in replay["iterations"]]
    assert tree.transitions == sum(steps) and eng.launches == replay["budget"]
    for node, want in zip(tree.Nodes, replay["nodes"]):
        assert np.array_equal(np.array([h.h for h in node.Values]), np.array(want["hist"]))
    for k, node in master.items():
        assert np.array_equal(np.array([h.h for h in node.rewards[0]]), np.array(replay["master_nodes"][str(k)]))
    # the generators: 301 shuffles interleaved with one normal per transition; 300 + 300 action slots
    end_py, end_np = random.getstate(), np.random.get_state()
    random.seed(1)
    random.shuffle(list(range(8)))
    for k in steps:
        random.shuffle(list(range(8)))
        for _ in range(k):
            random.gauss(0.0, 1.0)
    assert end_py == random.getstate()
    np.random.seed(1)
    np.random.randint(8, size=600)
    assert end_np[2] == np.random.get_state()[2] and np.array_equal(end_np[1], np.random.get_state()[1])
    # quirk Q1: the scalar path replays one action only, unless asked to replay the whole prefix
    tree.get_sim_to_estimate_state(tree.Nodes[-1])
    assert sims[0].state[0] == 1
    tree.replay_full_prefix = True
    deep = max(tree.Nodes, key=lambda n: n.depth)
    tree.get_sim_to_estimate_state(deep)
    assert sims[0].state[0] == deep.depth


def test_do_step_bookkeeping_and_exceptions(monkeypatch):
    """Simulator.doStep (simulatorTLKT.py:181-216): returns its own state list, keeps prevState, draws one normal per
    call, raises IndexError past the last instant and ValueError outside the weather grid."""
    sims, eng = make([0], monkeypatch)
    sim, o = sims[0], H.oracle_sim(0)
    random.seed(11)
    state = random.getstate()
    zs = [random.gauss(0.0, 1.0) for _ in range(12)]
    random.setstate(state)
    sim.reset(STATE0)
    k, la, lo = 0, 44.0, 355.0
    for j, action in enumerate([0, 45, 90, 315, 0, 0, 270, 0, 45, 0, 0, 0]):
        want = o.step_batch(k, [la], [lo], [[float(action)]], z=[[zs[j]]])
        prev = list(sim.state)
        st = sim.doStep(action)
        assert st is sim.state and sim.prevState == prev
        k, la, lo = int(want["k"][0]), float(want["lat"][0]), float(want["lon"][0])
        assert st == [k, la, lo] and isinstance(st[0], int) and isinstance(st[1], float)
    assert random.gauss(0.0, 1.0) != zs[0] and eng.launches == 12   # twelve draws were consumed, one per call
    with pytest.raises(IndexError):
        sim.doStep(0)
    sim.reset([0, 49.9, 359.9])
    with pytest.raises(ValueError):
        for _ in range(12):
            sim.doStep(45)
