"""Shared test helpers: scenario construction for the oracle and oracle-side drivers of the
reference's Monte-Carlo initialisation / plan evaluation (used to pin the oracle to the goldens and,
in the gpu tests, as the checker of the CUDA path)."""
import numpy as np

from iboat_pmcts_b200 import synth
from oracle import cport

STATE0 = (0, 44.0, 355.0)


def sim_axes(wlat, wlon, simtimestep=6, ndaysim=3, deltalatlon=0.5):
    """Axes built by create_simulators (code/solver/forest.py:182-184)."""
    times = np.arange(0, ndaysim + simtimestep / 24, simtimestep * (1 / 24))
    lats = np.arange(wlat[0], wlat[-1], deltalatlon)
    lons = np.arange(wlon[0], wlon[-1], deltalatlon)
    return times, lats, lons


def oracle_sim(scenario, times=None, sigma_coeff=0.2):
    t, la, lo, u, v = synth.wind_fields(scenario)
    if times is None:
        times, lats, lons = sim_axes(la, lo)
    else:
        lats, lons = la, lo
    return cport.OracleSim(t, la, lo, u, v, times, lats, lons, sigma_coeff=sigma_coeff)


def oracle_initialize(osims, ntra, stateinit, heading, zinit):
    """initialize_simulators (code/solver/forest.py:210-299) on oracle batches.
    zinit[(phase*S + s)*ntra + i, step] is the injected noise of trajectory i of scenario s."""
    S = len(osims)
    nT = osims[0].times.size
    means = []
    for s, o in enumerate(osims):
        z = zinit[s * ntra:(s + 1) * ntra].T  # [step][traj]
        r = o.step_batch(stateinit[0], np.full(ntra, stateinit[1]), np.full(ntra, stateinit[2]),
                         np.full((1, ntra), float(heading)), nsteps=nT, seg_len=nT, z=z,
                         flags=1)
        d = [o.dist_bearing(stateinit[1], stateinit[2], a, b)[0] for a, b in zip(r["lat"], r["lon"])]
        means.append(np.mean(d))
    mindist = np.min(means)
    dest = osims[-1].get_destination(mindist, heading, stateinit[1], stateinit[2])
    tmins = []
    for s, o in enumerate(osims):
        z = zinit[(S + s) * ntra:(S + s + 1) * ntra].T
        r = o.rollout_batch(ntra, stateinit, dest, 0.0, mode=1, z=z)
        arr = r["t_arr"][r["status"] == 1]
        if arr.size:
            tmins.append(arr.min())
    return dest, min(tmins)


def oracle_plan_eval(osims, ntra, stateinit, dest, plan, z):
    """estimate_perfomance_plan (code/isochrones/isochrones.py:469-580) on oracle batches."""
    means, vars_, all_t = [], [], []
    for s, o in enumerate(osims):
        zz = z[s * ntra:(s + 1) * ntra].T
        pre = np.tile(np.asarray(plan, dtype=np.float64), (ntra, 1)) if len(plan) else None
        r = o.rollout_batch(ntra, stateinit, dest, 0.0, prefix=pre, mode=1, z=zz)
        t = r["t_arr"]
        m = np.mean(t)
        means.append(m)
        var = 0.0
        for val in t:
            var += (m - val) ** 2
        vars_.append(var / ntra)
        all_t.extend(t.tolist())
    gm = np.mean(all_t)
    gv = 0.0
    for val in all_t:
        gv += (gm - val) ** 2
    gv /= len(all_t)
    return [gm] + means, [gv] + vars_


def southern_world(seed=0):
    """A different corner of the globe: southern hemisphere, negative longitudes, 0.25-degree grid, hourly
    weather, 2-hour simulator steps, destination to the south-west."""
    rng = np.random.default_rng(seed)
    wt = np.arange(0, 2.0001, 1 / 24)
    wlat = np.arange(-42.0, -33.99, 0.25)
    wlon = np.arange(-60.0, -49.99, 0.25)
    T, LAT, LON = np.meshgrid(wt, wlat, wlon, indexing="ij")
    u = (7 * np.sin(0.4 * LAT + 2 * T) + 4 * np.cos(0.3 * LON) + rng.normal(0, 1, T.shape)).astype(np.float32).astype(np.float64)
    v = (6 * np.cos(0.35 * LON - T) - 3 * np.sin(0.2 * LAT) + rng.normal(0, 1, T.shape)).astype(np.float32).astype(np.float64)
    times = np.arange(0, 2.0001, 2 / 24)
    lats, lons = wlat[:-1], wlon[:-1]
    return wt, wlat, wlon, u, v, times, lats, lons
