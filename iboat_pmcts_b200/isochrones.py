"""Mirror of the policy-evaluation entry point of ``code/isochrones/isochrones.py`` (reference):
``estimate_perfomance_plan`` (isochrones.py:469-580).  The isochrone solver itself is a consumer of
batched ``doStep`` that SURVEY.md section 8f ranks "next" (N2); it is not part of this package yet.
"""
import numpy as np

from . import _native as N
from .forest import _raise_unless, _reference_noise


def estimate_perfomance_plan(sims, ntra, stateinit, destination, plan=list(), plot=False, verbose=True, seed=None):
    """Mean and variance of the arrival time when every boat first applies the headings of ``plan`` (no
    arrival / terminal test in between) and then sails straight to ``destination``; ``ntra`` noisy runs per
    scenario.  A run that misses the destination counts ``times[-1]``.

    Returns ``[global_mean] + per_scenario_means, [global_variance] + per_scenario_variances`` (population
    variances), like the reference.

    ``seed=None``: noise from ``random.gauss`` in the reference's order (one launch per run), so seeded
    scripts reproduce the reference.  ``seed=int``: Philox noise, ONE launch per scenario (K4: the plan-
    evaluation mode of the rollout kernel with its fused sum / sum-of-squares reduction).
    """
    if plot:
        raise NotImplementedError("plotting is outside the scope of pmcts_b200")
    k0, lat0, lon0 = int(stateinit[0]), float(stateinit[1]), float(stateinit[2])
    plan = [float(a) for a in plan]
    if seed is not None:
        return _plan_eval_fused(sims, int(ntra), [k0, lat0, lon0], destination, plan, int(seed), verbose)
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


def _plan_eval_fused(sims, ntra, root, destination, plan, seed, verbose):
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
    stats = np.zeros((S, 8))
    prec = eng.precision
    eng.set_precision(N.PREC_FAST)
    try:
        for lo in range(0, S, 16):
            chunk = slots[lo:lo + 16]
            try:
                out = dict(stats=stats[lo:lo + len(chunk)].reshape(-1))
                eng.rollout_forest(chunk, 1, ntra, root, destination, 0.0, prefix, plen, len(plan), prefix_shared=True,
                                   seed=seed, stream=lo, out=out, flags=N.ROLL_PLAN_EVAL)
                stats[lo:lo + len(chunk)] = out["stats"].reshape(len(chunk), 8)
            except N.PmctsError:  # scenarios with different axes: one launch each
                for j, scen in enumerate(chunk):
                    out = dict(stats=np.zeros(8))
                    eng.rollout_batch(scen, 1, ntra, root, destination, 0.0, prefix, plen, len(plan), seed=seed,
                                      stream=lo + j, out=out, flags=N.ROLL_PLAN_EVAL)
                    stats[lo + j] = out["stats"]
    finally:
        eng.set_precision(prec)
    if (stats[:, 0] != ntra).any():
        # some run ended in an error state (left the weather grid ...): find it and raise what the reference raises
        for si, sim in enumerate(sims):
            e, scen = sim.scenario()
            pre = np.tile(prefix, (ntra, 1)) if prefix is not None else None
            r = e.rollout_batch_host(scen, ntra, 1, root, destination, 0.0, prefix=pre, seed=seed, stream=si,
                                     flags=N.ROLL_PLAN_EVAL, want=("status",))
            _raise_unless(r["status"], (N.ST_ARRIVED, N.ST_HORIZON, N.ST_OUT_OF_BOX))
    means = stats[:, 1] / ntra
    variances = np.maximum(stats[:, 2] / ntra - means ** 2, 0.0)
    gmean = stats[:, 1].sum() / (S * ntra)
    gvar = max(stats[:, 2].sum() / (S * ntra) - gmean ** 2, 0.0)
    if verbose:
        for nb in range(S):
            print("arrival time, scenario", nb, "=", means[nb], " variance =", variances[nb])
        print("mean arrival time =", gmean, " global variance =", gvar)
    return [gmean] + means.tolist(), [gvar] + variances.tolist()
