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
    mean_arrival_times, var_arrival_times, all_arrival_times = [], [], []
    for si, sim in enumerate(sims):
        eng, scen = sim.scenario()
        nT = len(sim.times)
        prefix = np.array([plan], np.float64) if plan else None
        prec = eng.precision
        eng.set_precision(N.PREC_F64 if seed is None else N.PREC_FAST)
        try:
            if seed is None:
                arrivaltimes = []
                for _ in range(ntra):
                    z, consume = _reference_noise(nT)
                    r = eng.rollout_batch_host(scen, 1, 1, [k0, lat0, lon0], destination, 0.0, prefix=prefix,
                                               z=z.reshape(nT, 1), flags=N.ROLL_PLAN_EVAL,
                                               want=("t_arr", "steps", "status"))
                    consume(int(r["steps"][0]))
                    _raise_unless(r["status"], (N.ST_ARRIVED, N.ST_HORIZON, N.ST_OUT_OF_BOX))
                    arrivaltimes.append(float(r["t_arr"][0]))
            else:
                pre = np.tile(prefix, (ntra, 1)) if prefix is not None else None
                r = eng.rollout_batch_host(scen, ntra, 1, [k0, lat0, lon0], destination, 0.0, prefix=pre, seed=seed,
                                           stream=si, flags=N.ROLL_PLAN_EVAL, want=("t_arr", "status"))
                _raise_unless(r["status"], (N.ST_ARRIVED, N.ST_HORIZON, N.ST_OUT_OF_BOX))
                arrivaltimes = r["t_arr"].tolist()
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
