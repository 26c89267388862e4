"""Mirror of ``code/isochrones/isochrones.py`` (reference): the policy-evaluation entry point
``estimate_perfomance_plan`` (isochrones.py:469-580) and the deterministic isochrone solver ``Isochrone``
(isochrones.py:24-464), whose front expansion is one batched ``doStep`` launch per
isochrone.  Plotting functions are out of scope.
"""
import numpy as np

from . import _native as N
from .forest import _raise_unless, _reference_noise
from .utils import out_of_scope


def estimate_perfomance_plan(sims, ntra, stateinit, destination, plan=list(), plot=False, verbose=True, seed=None,
                             group=None):
    """Mean and variance of the arrival time when every boat first applies the headings of ``plan`` (no
    arrival / terminal test in between) and then sails straight to ``destination``; ``ntra`` noisy runs per
    scenario.  A run that misses the destination counts ``times[-1]``.

    Returns ``[global_mean] + per_scenario_means, [global_variance] + per_scenario_variances`` (population
    variances), like the reference.

    ``seed=None``: noise from ``random.gauss`` in the reference's order (one launch per run), so seeded
    scripts reproduce the reference.  ``seed=int``: Philox noise, ONE launch per scenario (K4: the plan-
    evaluation mode of the rollout kernel with its fused sum / sum-of-squares reduction).

    ``group`` (with ``seed``): a ``torch.distributed`` process group whose ranks each hold a share of the scenarios, in
    rank order (the sharding of ``Forest.launch_search``).  Every rank evaluates its own ``sims``; the per-scenario
    ``(runs, sum t, sum t^2)`` are all-gathered -- 64 bytes per scenario, the one exchange of this path -- and every
    rank returns the result over ALL scenarios, which is what one process holding them all returns (the noise stream
    of a scenario is numbered by its global index).
    """
    if plot:
        raise NotImplementedError("plotting is outside the scope of pmcts_b200")
    k0, lat0, lon0 = int(stateinit[0]), float(stateinit[1]), float(stateinit[2])
    plan = [float(a) for a in plan]
    if seed is not None:
        return _plan_eval_fused(sims, int(ntra), [k0, lat0, lon0], destination, plan, int(seed), verbose, group)
    if group is not None:
        raise ValueError("estimate_perfomance_plan: a process group needs seed=int (the reference's own noise order "
                         "is one generator walked scenario after scenario: it does not shard)")
    mean_arrival_times, var_arrival_times, all_arrival_times = [], [], []
    for si, sim in enumerate(sims):
        eng, scen = sim.scenario()
        nT = len(sim.times)
        prefix = np.array([plan], np.float64) if plan else None
        prec = eng.precision
        eng.set_precision(N.PREC_F64)
        try:
            arrivaltimes = []
            for _ in range(ntra):
                z, consume = _reference_noise(nT)
                r = eng.rollout_batch_host(scen, 1, 1, [k0, lat0, lon0], destination, 0.0, prefix=prefix,
                                           z=z.reshape(nT, 1), flags=N.ROLL_PLAN_EVAL,
                                           want=("t_arr", "steps", "status"))
                consume(int(r["steps"][0]))
                _raise_unless(r["status"], (N.ST_ARRIVED, N.ST_HORIZON, N.ST_OUT_OF_BOX))
                arrivaltimes.append(float(r["t_arr"][0]))
        finally:
            eng.set_precision(prec)
        all_arrival_times.extend(arrivaltimes)
        average_scenario = np.mean(arrivaltimes)
        mean_arrival_times.append(average_scenario)
        variance_scenario = 0
        for value in arrivaltimes:
            variance_scenario += (average_scenario - value) ** 2
        var_arrival_times.append(variance_scenario / ntra)
    global_mean_time = np.mean(all_arrival_times)
    variance_globale = 0
    for value in all_arrival_times:
        variance_globale += (global_mean_time - value) ** 2
    variance_globale = variance_globale / len(all_arrival_times)
    if verbose:
        for nb in range(len(sims)):
            print("arrival time, scenario", nb, "=", mean_arrival_times[nb], " variance =", var_arrival_times[nb])
        print("mean arrival time =", global_mean_time, " global variance =", variance_globale)
    return [global_mean_time] + mean_arrival_times, [variance_globale] + var_arrival_times


def _plan_eval_fused(sims, ntra, root, destination, plan, seed, verbose, group=None):
    """K4: every scenario's ``ntra`` runs in one launch per group of scenarios, with the count / sum /
    sum of squares of the arrival times reduced in the kernel (no per-run output leaves the GPU)."""
    eng = None
    slots = []
    for sim in sims:
        eng, scen = sim.scenario()
        slots.append(scen)
    prefix = np.array(plan, np.float64) if plan else None
    plen = np.array([len(plan)], np.int32) if plan else None
    S = len(sims)
    rank, world = 0, 1
    if group is not None:
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    first = rank * S  # global index of this rank's first scenario: numbers its noise streams
    stats = np.zeros((S, 8))
    prec = eng.precision
    eng.set_precision(N.PREC_FAST)
    try:
        for lo in range(0, S, 16):
            chunk = slots[lo:lo + 16]
            try:
                out = dict(stats=stats[lo:lo + len(chunk)].reshape(-1))
                eng.rollout_forest(chunk, 1, ntra, root, destination, 0.0, prefix, plen, len(plan), prefix_shared=True,
                                   seed=seed, stream=first + lo, out=out, flags=N.ROLL_PLAN_EVAL)
                stats[lo:lo + len(chunk)] = out["stats"].reshape(len(chunk), 8)
            except N.PmctsError:  # scenarios with different axes: one launch each
                for j, scen in enumerate(chunk):
                    out = dict(stats=np.zeros(8))
                    eng.rollout_batch(scen, 1, ntra, root, destination, 0.0, prefix, plen, len(plan), seed=seed,
                                      stream=first + lo + j, out=out, flags=N.ROLL_PLAN_EVAL)
                    stats[lo + j] = out["stats"]
    finally:
        eng.set_precision(prec)
    local = stats
    if world > 1:  # (before anything may raise: every rank takes part in the exchange)
        stats = _all_gather_stats(local, group, world, getattr(eng, "device", None))
    if (local[:, 0] != ntra).any():
        # some run ended in an error state (left the weather grid ...): find it and raise what the reference raises
        for si, sim in enumerate(sims):
            e, scen = sim.scenario()
            pre = np.tile(prefix, (ntra, 1)) if prefix is not None else None
            r = e.rollout_batch_host(scen, ntra, 1, root, destination, 0.0, prefix=pre, seed=seed, stream=first + si,
                                     flags=N.ROLL_PLAN_EVAL, want=("status",))
            _raise_unless(r["status"], (N.ST_ARRIVED, N.ST_HORIZON, N.ST_OUT_OF_BOX))
    if (stats[:, 0] != ntra).any():
        raise RuntimeError("estimate_perfomance_plan: runs ended in an error state (scenarios %s%s)"
                           % (np.flatnonzero(stats[:, 0] != ntra).tolist(), ", other ranks' among them" if world > 1 else ""))
    S = stats.shape[0]
    estimate_perfomance_plan.last_stats = stats  # per scenario: runs, sum t, sum t^2, arrived, transitions, rollouts
    means = stats[:, 1] / ntra
    variances = np.maximum(stats[:, 2] / ntra - means ** 2, 0.0)
    gmean = stats[:, 1].sum() / (S * ntra)
    gvar = max(stats[:, 2].sum() / (S * ntra) - gmean ** 2, 0.0)
    if verbose:
        for nb in range(S):
            print("arrival time, scenario", nb, "=", means[nb], " variance =", variances[nb])
        print("mean arrival time =", gmean, " global variance =", gvar)
    return [gmean] + means.tolist(), [gvar] + variances.tolist()


def _all_gather_stats(stats, group, world, device):
    """[S_local][8] of every rank -> [world * S_local][8] in rank order (on the device over NCCL, on the host otherwise)."""
    import torch
    import torch.distributed as dist
    mine = torch.from_numpy(np.ascontiguousarray(stats))
    if dist.get_backend(group) == "nccl":
        mine = mine.to(torch.device("cuda", device))
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return torch.cat(out).cpu().numpy()


# ---- the deterministic isochrone solver (isochrones.py:24-464): a consumer of batched doStep --------------
class Node():
    """A point of an isochrone: state ``[time, lat, lon]``, parent node of the previous isochrone (``pere``),
    heading followed from it (``act``), bearing ``C`` and distance ``D`` from the starting point."""

    def __init__(self, time, lat, lon, parent=None, action=None, C=None, D=None):
        self.time = time
        self.lat = lat
        self.lon = lon
        self.pere = parent
        self.act = action
        self.C = C
        self.D = D

    def give_state(self):
        return [self.time, self.lat, self.lon]

    def __str__(self):
        return '({},{}) à {}\nC={},D={}'.format(self.lat, self.lon, self.time, self.C, self.D)


class Secteur():
    """Angular sector seen from the starting point; keeps the nodes that fall in it and their distances."""

    def __init__(self, cap_sup, cap_inf):
        self.cap_sup = cap_sup
        self.cap_inf = cap_inf
        self.liste_noeud = []
        self.liste_distance = []

    def recherche_meilleur_noeud(self):
        if not self.liste_distance:
            return None
        return self.liste_noeud[self.liste_distance.index(max(self.liste_distance))]

    def __str__(self):
        return '({},{}) à {}'.format(self.cap_inf, self.cap_sup, self.liste_distance)


class Isochrone():
    """Hagiwara's isochrone method with deterministic dynamics, same attributes and methods as the
    reference's class.  Each isochrone is expanded by ONE batched ``doStep`` launch (every node of the front x
    every allowed heading; 600 x 19 transitions per isochrone with the tutorial's parameters) instead of
    one scalar ``doStep`` per pair; the sector bookkeeping stays on the host.  Like the reference it draws one
    ``random.gauss`` per transition, so with ``Boat.UNCERTAINTY_COEFF != 0`` a seeded run sees the same noise.
    """

    def __init__(self, simulateur, coord_depart, coord_arrivee, delta_cap=10, increment_cap=9, nb_secteur=10,
                 resolution=200, temps=0):
        self.sim = simulateur
        self.dep = coord_depart[1:]
        self.arr = coord_arrivee
        self.temps_dep = temps
        noeuddep = Node(temps, coord_depart[1], coord_depart[2])
        self.isochrone_future = []
        self.distance_moy_iso = 0
        self.reso = resolution
        self.p = nb_secteur
        self.constante = np.pi / (60 * 180)
        self.delta_t = (self.sim.times[noeuddep.time + 1] - self.sim.times[noeuddep.time])
        self.liste_actions = [l * delta_cap for l in range(-increment_cap, increment_cap + 1)]
        self.liste_positions = []
        self.isochrone_stock = []
        self.temps_transit = 0
        D0, C0 = self.sim.getDistAndBearing(self.dep, self.arr)
        self.C0 = C0
        self.isochrone_actuelle = self._expand([noeuddep], [C0])
        self.isochrone_stock.append([n.give_state() for n in self.isochrone_actuelle])

    def recentrage_cap(self, cap):
        return cap % 360

    def reset(self, coord_depart=None, coord_arrivee=None):
        if coord_depart is not None:
            self.dep = coord_depart
        if coord_arrivee is not None:
            self.arr = coord_arrivee
        noeuddep = Node(0, self.dep[0], self.dep[1])
        self.isochrone_future = []
        self.distance_moy_iso = 0
        self.liste_positions = []
        self.isochrone_stock = []
        self.temps_transit = 0
        D0, C0 = self.sim.getDistAndBearing(self.dep, self.arr)
        self.C0 = C0
        self.isochrone_actuelle = self._expand([noeuddep], [C0])
        return None

    def _expand(self, nodes, centres):
        """Children of every node for every allowed heading around ``centres[i]`` -- one launch of the
        literal fp64 ``doStep`` kernel for all (node, heading) pairs, in the reference's loop order."""
        import random as rand
        from .simulatorTLKT import raise_for_status
        eng, scen = self.sim.scenario()
        caps = [[self.recentrage_cap(c + a) for a in self.liste_actions] for c in centres]
        n = len(nodes) * len(self.liste_actions)
        k = np.repeat(np.array([nd.time for nd in nodes], np.int32), len(self.liste_actions))
        lat = np.repeat(np.array([nd.lat for nd in nodes], np.float64), len(self.liste_actions))
        lon = np.repeat(np.array([nd.lon for nd in nodes], np.float64), len(self.liste_actions))
        head = np.array(caps, np.float64).reshape(1, n)
        z = np.array([rand.gauss(0.0, 1.0) for _ in range(n)], np.float64).reshape(1, n)
        st = np.zeros(n, np.uint8)
        prec = eng.precision
        eng.set_precision(N.PREC_F64)
        try:
            eng.step_batch(scen, k, lat, lon, head, nsteps=1, status=st, z=z)
        finally:
            eng.set_precision(prec)
        for s in st:
            if s != N.ST_OK:
                raise_for_status(int(s))  # IndexError past the horizon, ValueError outside the grid
        out = []
        per = len(self.liste_actions)
        for i in range(n):
            Dij, Cij = self.sim.getDistAndBearing(self.dep, [float(lat[i]), float(lon[i])])
            out.append(Node(int(k[i]), float(lat[i]), float(lon[i]), nodes[i // per], float(head[0, i]), Cij, Dij))
        return out

    def isochrone_brouillon(self):
        """All the nodes reachable from the current isochrone with the allowed headings."""
        self.isochrone_future = self._expand(self.isochrone_actuelle, [x.C for x in self.isochrone_actuelle])
        total = 0
        for node in self.isochrone_future:
            total += node.D
        self.distance_moy_iso = total / len(self.isochrone_future)
        return None

    def secteur_liste(self):
        """2p sectors of angular width atan(resolution / mean distance) around the bearing to the arrival."""
        from math import atan2
        delta_S = atan2(self.reso, self.distance_moy_iso) * 180 / np.pi
        liste_S = []
        for k in range(1, 2 * self.p + 1):
            cap_sup = self.recentrage_cap(self.C0 + (k - self.p) * delta_S)
            cap_inf = self.recentrage_cap(self.C0 + (k - self.p - 1) * delta_S)
            liste_S.append(Secteur(cap_sup, cap_inf))
        return liste_S, delta_S

    def associer_xij_a_S(self, liste_S, delta_S):
        """Puts every reachable node in the sector its bearing falls in (nodes outside the fan are dropped)."""
        borne_sup = liste_S[-1].cap_sup
        borne_inf = liste_S[0].cap_inf
        for xij in self.isochrone_future:
            Cij = xij.C
            if borne_sup > borne_inf:
                inside = not (Cij < borne_inf or Cij > borne_sup)
            else:  # the fan straddles north
                inside = (0 <= Cij < borne_sup or 360 > Cij > borne_inf)
            if not inside:
                continue
            diff = Cij - self.C0
            if diff <= -180:
                diff = diff + 360
            elif diff >= 180:
                diff = diff - 360
            indice_S = int(diff / delta_S + self.p)
            liste_S[indice_S].liste_noeud.append(xij)
            liste_S[indice_S].liste_distance.append(xij.D)
        return liste_S

    def nouvelle_isochrone_propre(self, liste_S):
        """Keeps the farthest node of every sector: the new current isochrone."""
        self.isochrone_actuelle = []
        liste_etats = []
        for sect in liste_S:
            keep = sect.recherche_meilleur_noeud()
            if keep is not None:
                self.isochrone_actuelle.append(keep)
                liste_etats.append(keep.give_state())
        self.isochrone_stock.append(liste_etats)
        return None

    def isochrone_proche_arrivee(self):
        """Nodes of the current isochrone within 20 km of the arrival."""
        top = []
        for xi in self.isochrone_actuelle:
            Ddest, _ = self.sim.getDistAndBearing([xi.lat, xi.lon], self.arr)
            if Ddest <= 20000:
                top.append(xi)
        return len(top) > 0, top

    def aller_point_arrivee(self, Top_noeud):
        """Sends every node close to the arrival straight to it; returns the node that arrives first, the
        total travel time and the headings of that last leg.  Raises IndexError when none arrives in time."""
        from .worker import Tree
        times, caps = [], []
        for noeud in Top_noeud:
            atDest, frac = False, 0
            finish_caps = []
            self.sim.reset([noeud.time, noeud.lat, noeud.lon])
            try:
                while not atDest:
                    _, cap = self.sim.getDistAndBearing(self.sim.state[1:3], self.arr)
                    finish_caps.append(cap)
                    self.sim.doStep(cap)
                    atDest, frac = Tree.is_state_at_dest(self.arr, self.sim.prevState, self.sim.state)
                times.append(self.sim.times[self.sim.state[0]] - (1 - frac) * self.delta_t
                             - self.sim.times[self.temps_dep])
                caps.append(finish_caps)
            except IndexError:
                pass
        if not times:
            raise IndexError
        best = times.index(min(times))
        # like the reference (isochrones.py:396-398) the node is taken from Top_noeud at the index of the
        # best *successful* run, which is another node whenever an earlier one ran out of time
        return Top_noeud[best], times[best], caps[best]

    def isochrone_methode(self):
        """Runs the method; returns (travel time in days, headings, headings of the last leg, positions), or
        Nones when no solution exists within the time horizon."""
        liste_caps_fin = None
        try:
            arrive = False
            while not arrive:
                self.isochrone_brouillon()
                liste_S, delta_S = self.secteur_liste()
                liste_S = self.associer_xij_a_S(liste_S, delta_S)
                self.nouvelle_isochrone_propre(liste_S)
                arrive, Top_noeud = self.isochrone_proche_arrivee()
            node, temps_total, liste_caps_fin = self.aller_point_arrivee(Top_noeud)
            passage, caps = [], []
            while node.pere is not None:
                passage.append([node.lat, node.lon])
                caps.append(node.act)
                node = node.pere
            passage.append([node.lat, node.lon])
            self.liste_positions = passage[::-1]
            self.liste_positions.append(self.arr)
            self.liste_actions = caps[::-1]
            self.temps_transit = temps_total
        except IndexError:
            print('Pas de solution trouvée dans le temps imparti.\nVeuillez raffiner vos paramètres de recherche.')
            self.temps_transit = None
            self.liste_actions = None
            liste_caps_fin = None
            self.liste_positions = None
        return self.temps_transit, self.liste_actions, liste_caps_fin, self.liste_positions

    def positions_to_states(self):
        return [np.array([i, p[0], p[1]]) for i, p in enumerate(self.liste_positions)]


for _name in ("plot_trajectory", "plot_comparision", "plot_comparision_percent", "plot_mean_squared_error",
              "plot_risk_probability"):
    globals()[_name] = out_of_scope("isochrones." + _name)
