"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/* by running the UNMODIFIED Python reference.

Run in the build container (needs /root/reference):  python -m oracle.gen_golden
The fixtures pin the C restatement (oracle/pmcts_oracle.c) and, through it, the CUDA path, to
outputs of the reference itself on the synthetic scenarios of iboat_pmcts_b200/synth.py.

Noise: the reference draws one ``random.gauss`` per transition (simulatorTLKT.py:515).  Here the
``random`` module seen by simulatorTLKT is replaced by a feeder that returns ``mu + z*sigma`` for an
explicit z[rollout, step] (the same formula random.gauss ends with), so both sides of a parity test
consume identical standard normals.
"""
import json
import os
import random
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402
from iboat_pmcts_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


class ZFeeder:
    """Stands in for the ``random`` module inside simulatorTLKT: gauss(mu, s) = mu + z*s."""

    def __init__(self, z):
        self.z = np.asarray(z, dtype=np.float64)
        self.i = 0
        self.j = 0

    def begin(self, i):
        self.i, self.j = i, 0

    def gauss(self, mu, sigma):
        zz = float(self.z[self.i, self.j])
        self.j += 1
        return mu + zz * sigma


def make_sims(R, scenarios, simtimestep=6, ndaysim=3, stateinit=(0, 44.0, 355.0)):
    ws = []
    for s in scenarios:
        t, la, lo, u, v = synth.wind_fields(s)
        ws.append(R.weatherTLKT.Weather(la, lo, t, u, v))
    return R.forest.create_simulators(ws, len(ws), simtimestep=simtimestep, stateinit=stateinit,
                                      ndaysim=ndaysim)


def kat(R):
    simc, wth, wrk, rutils = R.simulatorTLKT, R.weatherTLKT, R.worker, R.utils
    sim = make_sims(R, [0])[0]
    B = simc.Boat
    out = {}
    out["getDeterDyn"] = [[p, w, float(B.getDeterDyn(p, w, B.FIT_VELOCITY))] for p, w in
                          [(0, 10), (20, 10), (35, 10), (45, 5), (90, 10), (135, 15), (160, 10), (170, 10),
                           (180, 10), (200, 10), (90, 0), (90, 19.1), (90, 25), (60, 2), (34.999, 7.5),
                           (160.001, 3.25), (359.0, 12.0), (181.0, 18.0)]]
    out["getDestination"] = [[d, b, la, lo, sim.getDestination(d, b, [la, lo])] for d, b, la, lo in
                             [(1e5, 45, 44, 355), (-5000, 270, 44, 355), (49000.5, 123.4, 47.25, 351.125),
                              (0.0, 10, 45, 350), (2.2e5, 359.9, 41.0, 359.5)]]
    out["getDistAndBearing"] = [[a, b, list(sim.getDistAndBearing(a, b))] for a, b in
                                [([44, 355], [46.5, 355]), ([44, 355], [45, 357]), ([46.4, 355.2], [46.5, 355.0]),
                                 ([47.0, 352.0], [43.5, 358.75])]]
    out["returnPolarVel"] = [[u, v, list(wth.Weather.returnPolarVel(u, v))] for u, v in
                             [(3, 4), (-3, -4), (0.0, 5.5), (-7.25, 0.125), (1e-3, -1e-3)]]
    out["fromGeoToCartesian"] = [[la, lo, simc.Simulator.fromGeoToCartesian([la, lo])] for la, lo in
                                 [(44, 355), (46.5, 355.0), (40.0, 359.5)]]
    cases = [([46.5, 355], [0, 46, 355], [1, 47, 355.0001]), ([46.5, 355], [0, 46, 355], [1, 46.4, 355.0001]),
             ([46.5, 355], [0, 46, 355], [1, 47, 355.5]), ([46.5, 355.0], [3, 46.2, 354.9], [4, 46.8, 355.1])]
    out["is_state_at_dest"] = []
    for d, a, b in cases:
        r = wrk.Tree.is_state_at_dest(d, a, b)
        out["is_state_at_dest"].append([d, a[1:], b[1:], bool(r[0]), None if r[1] is None else float(r[1])])
    bins = []
    for val in [0, 0.125, 0.1250001, 0.5, 0.99, 1.0, 1.2, 1.5, 2.0, -0.1, 0.62500000001, 1.375, 1.3750001]:
        h = rutils.Hist()
        h.add(val)
        bins.append([val, int(np.argmax(h.h))])
    out["hist_bin"] = bins
    out["hist_means"] = [float(x) for x in rutils.Hist.MEANS]
    tmax, tmin = 3.0, 2.4144127151444956
    from math import exp
    out["reward"] = [[tmax, tmin, t, (exp((tmax * 1.001 - t) / (tmax * 1.001 - tmin)) - 1) / (exp(1) - 1)]
                     for t in [2.0, tmin, 2.75, 3.0]]
    out["constants"] = dict(UCT_COEFF=wrk.UCT_COEFF, RHO=wrk.RHO, DESTINATION_ANGLE=simc.DESTINATION_ANGLE,
                            EARTH_RADIUS=simc.EARTH_RADIUS, ACTIONS=[int(a) for a in simc.ACTIONS],
                            root_hash=hash(tuple([])), hash_0_45=hash((0, 45)))
    rng = np.random.default_rng(11)
    pts = np.stack([rng.uniform(0, 5, 40), rng.uniform(40, 50, 40), rng.uniform(345, 360, 40)], 1)
    pts[0] = [0.0, 40.0, 345.0]
    pts[1] = [5.0, 50.0, 360.0]
    pts[2] = [0.125, 44.5, 355.0]
    W = make_sims(R, [0])[0]
    out["getWind_scenario0"] = [[float(t), float(a), float(o), float(W.uWindAvg([t, a, o])[0]),
                                 float(W.vWindAvg([t, a, o])[0])] for t, a, o in pts]
    return out


def raw_steps(R, nboat=24, nsteps=500):
    """BASELINE config 3 shape: octagon fleet, 500 steps of 864 s on scenario 0."""
    simc = R.simulatorTLKT
    t, la, lo, u, v = synth.wind_fields(0)
    W = R.weatherTLKT.Weather(la, lo, t, u, v)
    times = np.linspace(0, 5, nsteps + 1)
    sim = simc.Simulator(times, la, lo, W, [0, 44.0, 355.0])
    lat0, lon0, h0 = synth.octagon_fleet(nboat)
    head = synth.octagon_headings(h0, nsteps)
    z = np.random.default_rng(42).standard_normal((nboat, nsteps))
    feeder = ZFeeder(z)
    simc.rand = feeder
    tl = np.empty((nsteps, nboat))
    to = np.empty((nsteps, nboat))
    for b in range(nboat):
        feeder.begin(b)
        sim.reset([0, lat0[b], lon0[b]])
        for j in range(nsteps):
            st = sim.doStep(float(head[j // 25, b]))
            tl[j, b], to[j, b] = st[1], st[2]
    simc.rand = random
    return dict(lat0=lat0, lon0=lon0, h0=h0, z=z.T.copy(), traj_lat=tl, traj_lon=to, times=times)


def init_and_rollouts(R, scenarios=(0, 1, 2), ntra=50, n_per_action=6):
    """Tutorial-shaped (config 1/2) initialisation + default-policy rollouts with injected noise."""
    simc, wrk, frst = R.simulatorTLKT, R.worker, R.forest
    sims = make_sims(R, list(scenarios))
    nT = len(sims[0].times)
    state0 = [0, 44.0, 355.0]
    # --- initialize_simulators (forest.py:210-299) with a recorded noise feed ------------------
    S = len(sims)
    zinit = np.random.default_rng(5).standard_normal((2 * S * ntra, nT))
    feeder = ZFeeder(zinit)
    simc.rand = feeder
    counter = {"i": -1}
    orig_reset = simc.Simulator.reset

    def counting_reset(self, st):
        counter["i"] += 1
        feeder.begin(counter["i"])
        return orig_reset(self, st)

    simc.Simulator.reset = counting_reset
    destination, timemin = frst.initialize_simulators(sims, ntra, state0, 0)
    simc.Simulator.reset = orig_reset
    assert counter["i"] == 2 * S * ntra - 1
    # --- default-policy rollouts (worker.py:255-300) --------------------------------------------
    actions = [int(a) for a in simc.ACTIONS]
    recs = []
    n = len(actions) * n_per_action
    zroll = np.random.default_rng(6).standard_normal((S, n, nT))
    traj_lat = np.full((S, nT, n), np.nan)
    traj_lon = np.full((S, nT, n), np.nan)
    reward = np.zeros((S, n))
    steps = np.zeros((S, n), np.int32)
    atdest = np.zeros((S, n), np.int8)
    frac = np.full((S, n), np.nan)
    kfin = np.zeros((S, n), np.int32)
    prefix = np.zeros((n, 3))
    prefix_len = np.zeros(n, np.int32)
    for si, sim in enumerate(sims):
        tree = wrk.Tree(workerid=si, nscenario=S, simulator=sim, destination=destination, TimeMin=timemin,
                        buffer=[])
        tree.rootNode = types.SimpleNamespace(state=tuple(state0))
        feeder = ZFeeder(zroll[si])
        simc.rand = feeder
        for r in range(n):
            a = actions[r % 8]
            # deeper origins exercise quirk Q1 (only origins[0] is replayed, worker.py:297-298)
            origins = [a] + [actions[(r + 3) % 8], actions[(r + 5) % 8]][: (r // 8) % 3]
            prefix[r, :len(origins)] = origins
            prefix_len[r] = len(origins)
            node = types.SimpleNamespace(origins=origins)
            feeder.begin(r)
            log = []
            orig = sim.doStep

            def rec(action, _orig=orig, _log=log):
                st = _orig(action)
                _log.append((st[1], st[2]))
                return st

            sim.doStep = rec
            reward[si, r] = tree.default_policy(node)
            del sim.doStep
            steps[si, r] = len(log)
            for j, (x, y) in enumerate(log):
                traj_lat[si, j, r], traj_lon[si, j, r] = x, y
            res = wrk.Tree.is_state_at_dest(destination, sim.prevState, sim.state)
            atdest[si, r] = bool(res[0])
            if res[0]:
                frac[si, r] = res[1]
            kfin[si, r] = sim.state[0]
    simc.rand = random
    return dict(scenarios=np.array(scenarios), ntra=ntra, zinit=zinit, destination=np.array(destination),
                timemin=timemin, zroll=zroll, prefix=prefix, prefix_len=prefix_len, traj_lat=traj_lat,
                traj_lon=traj_lon, reward=reward, steps=steps, atdest=atdest, frac=frac, kfin=kfin,
                times=np.asarray(sims[0].times), lats=np.asarray(sims[0].lats), lons=np.asarray(sims[0].lons))


def plan_eval(R, destination, scenarios=(0, 1, 2), ntra=20, plan=(0, 0, 315, 0, 45, 0)):
    """estimate_perfomance_plan (isochrones.py:469-580) with an injected noise feed."""
    simc, iso = R.simulatorTLKT, R.isochrones
    sims = make_sims(R, list(scenarios))
    nT = len(sims[0].times)
    S = len(sims)
    out = {}
    for name, pl in (("plan", list(plan)), ("empty", [])):
        z = np.random.default_rng(8).standard_normal((S * ntra, nT))
        feeder = ZFeeder(z)
        simc.rand = feeder
        counter = {"i": -1}
        orig_reset = simc.Simulator.reset

        def counting_reset(self, st):
            counter["i"] += 1
            feeder.begin(counter["i"])
            return orig_reset(self, st)

        simc.Simulator.reset = counting_reset
        mean, var = iso.estimate_perfomance_plan(sims, ntra, [0, 44.0, 355.0], list(destination), pl,
                                                 plot=False, verbose=False)
        simc.Simulator.reset = orig_reset
        simc.rand = random
        out[name] = dict(plan=pl, ntra=ntra, z=z, mean=[float(x) for x in mean], var=[float(x) for x in var])
    return out


def tree_replay(R, destination, timemin, budget=300, frequency=10):
    """One seeded reference uct_search (worker.py:147-186) on scenario 0 with a plain-dict master:
    records, per iteration, the expanded leaf, the random slot of Node.back_up (worker.py:62), the
    random slot of MasterNode.add_reward (master_node.py:42) and the reward, then every node's
    histograms and the master's best policy."""
    simc, wrk, mnode, mstr = R.simulatorTLKT, R.worker, R.master_node, R.master
    sims = make_sims(R, [0])
    random.seed(1)
    np.random.seed(1)
    saved = (wrk.RHO, wrk.UCT_COEFF)
    wrk.RHO, wrk.UCT_COEFF = 0.5, 1 / 2 ** 0.5  # Tutorial.py:61-62
    slots = []
    orig_randint = np.random.randint

    def rec_randint(*a, **k):
        v = orig_randint(*a, **k)
        slots.append(int(v))
        return v

    np.random.randint = rec_randint
    tree = wrk.Tree(workerid=0, nscenario=1, budget=budget, simulator=sims[0], destination=destination,
                    TimeMin=timemin, buffer=[])
    master = {hash(tuple([])): mnode.MasterNode(1, nodehash=hash(tuple([])))}
    iters = []
    orig_backup = wrk.Node.back_up

    # (transitions per iteration -- the normals an iteration draws from Python's generator -- for the replay of the
    #  NATIVE sequential search on the CPU: kept out of the iteration records, see --only-replay-steps)
    made = {"n": 0}
    steps_per_iteration = []
    orig_step = simc.Simulator.doStep

    def counting_step(self, action):
        made["n"] += 1
        return orig_step(self, action)

    simc.Simulator.doStep = counting_step

    def rec_backup(self, reward):
        iters.append(dict(origins=[int(a) for a in self.origins], reward=float(reward), n_slots_before=len(slots)))
        steps_per_iteration.append(made["n"])
        made["n"] = 0
        return orig_backup(self, reward)

    wrk.Node.back_up = rec_backup
    import io
    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        tree.uct_search([0, 44.0, 355.0], frequency, master)
    wrk.Node.back_up = orig_backup
    simc.Simulator.doStep = orig_step
    tree_replay.steps_per_iteration = steps_per_iteration
    np.random.randint = orig_randint
    # slot stream: per iteration one worker slot; every `frequency` iterations `frequency` master slots
    for it in iters:
        it["slot"] = slots[it.pop("n_slots_before")]
    master_slots = []
    pos = 0
    for i in range(len(iters)):
        pos += 1
        if (i + 1) % frequency == 0:
            master_slots.extend(slots[pos:pos + frequency])
            pos += frequency
    assert pos == len(slots)
    nodes = [dict(origins=[int(a) for a in nd.origins], hist=[[int(x) for x in h.h] for h in nd.Values])
             for nd in tree.Nodes]
    mnodes = {str(k): [[int(x) for x in h.h] for h in nd.rewards[0]] for k, nd in master.items()}
    class _ProxyLike(dict):  # Manager().dict().items() returns a list copy; deepcopy_dict deletes while iterating
        def items(self):
            return list(dict.items(self))

    new_dict = mnode.deepcopy_dict(_ProxyLike(master))
    mt = mstr.MasterTree(sims, destination, nodes=new_dict)
    with contextlib.redirect_stdout(io.StringIO()):
        mt.get_best_policy()
    wrk.RHO, wrk.UCT_COEFF = saved
    return dict(budget=budget, frequency=frequency, rho=0.5, uct_coeff=1 / 2 ** 0.5, iterations=iters,
                master_slots=master_slots, nodes=nodes, master_nodes=mnodes,
                best_policy={str(k): [int(a) for a in v] for k, v in mt.best_policy.items()},
                depth=int(tree.depth))


def isochrone_fixture(R, destination):
    """The reference's deterministic isochrone solver (isochrones.py:24-464) on scenario 0 with
    Boat.UNCERTAINTY_COEFF = 0 (Tutorial.py:81): every isochrone it keeps, and its solution."""
    simc, iso = R.simulatorTLKT, R.isochrones
    sims = make_sims(R, [0])
    saved = simc.Boat.UNCERTAINTY_COEFF
    simc.Boat.UNCERTAINTY_COEFF = 0
    random.seed(3)
    try:
        import io
        import contextlib
        with contextlib.redirect_stdout(io.StringIO()):
            solver = iso.Isochrone(sims[0], [0, 44.0, 355.0], list(destination), delta_cap=5, increment_cap=9,
                                   nb_secteur=300, resolution=100)  # Tutorial.py:90-91
            temps, actions, caps_fin, positions = solver.isochrone_methode()
    finally:
        simc.Boat.UNCERTAINTY_COEFF = saved
    return dict(destination=[float(x) for x in destination], temps_transit=float(temps),
                liste_actions=[float(a) for a in actions], liste_caps_fin=[float(a) for a in caps_fin],
                liste_positions=[[float(p[0]), float(p[1])] for p in positions], C0=float(solver.C0),
                isochrone_stock=[[[int(st[0]), float(st[1]), float(st[2])] for st in front]
                                 for front in solver.isochrone_stock])


def master_policy_fixture(R, n_scen=10, budget=120, leaves=16, replicas=24):
    """The reference's own MasterTree post-processing (master.py:57-185) on a 10-scenario master tree: best policy
    of every scenario and of the global tree (uniform and skewed scenario probabilities), and along every policy
    the utilities / guessed rewards the reference computed on the way.  The tree itself comes from this repo's host
    search runtime driven by a deterministic stand-in for the rollouts (tests/fake_device.py: no GPU here); what is
    pinned is the reduction of a given table, so any table with shallow nodes shared by all scenarios and deep
    nodes visited by one will do."""
    import contextlib
    import io
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fake_device
    mnode, mstr = R.master_node, R.master
    _, master, _ = fake_device.run("native", n_scen=n_scen, budget=budget, leaves=leaves, replicas=replicas)
    parent, arm, table = master.arrays()
    n = parent.size
    actions = [int(a) for a in R.simulatorTLKT.ACTIONS]

    def build():
        paths, nodes = [], {}
        objs = []
        for i in range(n):
            path = [] if parent[i] < 0 else paths[parent[i]] + [actions[arm[i]]]
            paths.append(path)
            node = mnode.MasterNode(n_scen, nodehash=hash(tuple(path)), parentNode=objs[parent[i]] if parent[i] >= 0 else None,
                                    action=actions[arm[i]] if parent[i] >= 0 else None)
            for s in range(n_scen):
                for a in range(8):
                    node.rewards[s, a].h = np.array(table[i, s, a], dtype=int)
            objs.append(node)
            nodes[node.hash] = node

        class _ProxyLike(dict):  # Manager().dict().items() returns a list copy; deepcopy_dict deletes while iterating
            def items(self):
                return list(dict.items(self))
        return mnode.deepcopy_dict(_ProxyLike(nodes)), paths

    out = dict(parent=parent, arm=arm, table=table.astype(np.uint16), uct_coeff=float(mstr.UCT_COEFF))
    rng = np.random.default_rng(11)
    skew = rng.uniform(0.2, 1.0, n_scen)
    skew[3] = 12.0  # one scenario carries most of the weight: another global policy than the uniform one
    skew /= skew.sum()
    for tag, proba in (("uniform", []), ("skewed", list(skew))):
        nodes, paths = build()
        index = {hash(tuple(p)): i for i, p in enumerate(paths)}
        mt = mstr.MasterTree([None] * n_scen, [46.7, 355.0], nodes=nodes, proba=proba)
        with contextlib.redirect_stdout(io.StringIO()):
            mt.get_best_policy()
        pol = np.full((n_scen + 1, 12), -1, np.int32)
        for sid, acts in mt.best_policy.items():
            pol[sid + 1, :len(acts)] = acts
        out["policy_" + tag] = pol
        out["probability_" + tag] = np.asarray(mt.probability, np.float64)
        # utilities along the policies, as the reference's own calls return them
        rec = []
        for sid, path_nodes in mt.best_nodes_policy.items():
            for node in path_nodes[1:]:
                value, expl = mt.get_utility(node, sid)
                rec.append((index[node.hash], sid, value, expl))
        out["utility_" + tag] = np.asarray(rec, np.float64)
        # every guess the global descent left on the nodes (master.py:165-167)
        gs = [(index[h], s, g) for h, node in nodes.items() for s, g in node.guessed_rewards.items()]
        out["guess_" + tag] = np.asarray(gs, np.float64)
    return out


def main():
    R = refshim.load_reference()
    os.makedirs(GOLD, exist_ok=True)
    if "--only-master" in sys.argv:
        fx = master_policy_fixture(R)
        np.savez_compressed(os.path.join(GOLD, "master_policy_s10.npz"), **fx)
        print("master policy: nodes", fx["parent"].size, "global policy", fx["policy_uniform"][0], "skewed",
              fx["policy_skewed"][0], "guesses", fx["guess_uniform"].shape, "utilities", fx["utility_uniform"].shape)
        return
    if "--only-replay-steps" in sys.argv:
        # the recorded search run once more, counting the transitions of every iteration; written next to the recorded
        # run only if it IS that run (same leaves, same rewards, same histograms)
        g = np.load(os.path.join(GOLD, "rollouts_c1.npz"))
        tr = tree_replay(R, [float(x) for x in g["destination"]], float(g["timemin"]))
        with open(os.path.join(GOLD, "tree_replay_c1.json")) as f:
            old = json.load(f)
        assert tr == old, "the re-run is not the recorded search"
        steps = tree_replay.steps_per_iteration
        assert len(steps) == len(old["iterations"]) and min(steps) >= 1
        with open(os.path.join(GOLD, "tree_replay_c1_steps.json"), "w") as f:
            json.dump(dict(steps=steps, note="transitions (doStep calls, one normal each) of every iteration of "
                                             "tree_replay_c1.json"), f)
        print("replay steps:", len(steps), "iterations,", sum(steps), "transitions")
        return
    if "--only-isochrone" in sys.argv:
        g = np.load(os.path.join(GOLD, "rollouts_c1.npz"))
        fx = isochrone_fixture(R, g["destination"])
        with open(os.path.join(GOLD, "isochrone_c1.json"), "w") as f:
            json.dump(fx, f)
        print("isochrone: transit", fx["temps_transit"], "actions", fx["liste_actions"], "fronts",
              [len(x) for x in fx["isochrone_stock"]])
        return
    with open(os.path.join(GOLD, "kat.json"), "w") as f:
        json.dump(kat(R), f, indent=1)
    np.savez(os.path.join(GOLD, "raw_steps_c3.npz"), **raw_steps(R))
    g = init_and_rollouts(R)
    np.savez(os.path.join(GOLD, "rollouts_c1.npz"), **g)
    pe = plan_eval(R, g["destination"])
    np.savez(os.path.join(GOLD, "plan_eval_c5.npz"),
             **{f"{k}_{kk}": np.asarray(vv) for k, v in pe.items() for kk, vv in v.items()})
    tr = tree_replay(R, [float(x) for x in g["destination"]], float(g["timemin"]))
    with open(os.path.join(GOLD, "tree_replay_c1.json"), "w") as f:
        json.dump(tr, f)
    with open(os.path.join(GOLD, "isochrone_c1.json"), "w") as f:
        json.dump(isochrone_fixture(R, g["destination"]), f)
    np.savez_compressed(os.path.join(GOLD, "master_policy_s10.npz"), **master_policy_fixture(R))
    print("destination", g["destination"], "timemin", g["timemin"])
    print("best_policy", tr["best_policy"], "depth", tr["depth"])
    for fn in sorted(os.listdir(GOLD)):
        print(fn, os.path.getsize(os.path.join(GOLD, fn)))


if __name__ == "__main__":
    main()
