"""Arithmetic of the mixed-precision kernels (csrc/pmcts_fast.cuh) against the fp64 oracle, on the CPU.

tests/fastmodel compiles the very header the CUDA kernels are built from with g++ (MUFU approximations
replaced by IEEE operations), so these tests pin the *formulas* -- the difference-form geodesy, the
dot-product point of sail, the centred polar polynomial, the decision margins -- without a GPU.  The
kernels themselves are compared with the oracle in test_gpu_parity.py.
"""
import ctypes as C
import os

import numpy as np
import pytest

import pmcts_helpers as H
from iboat_pmcts_b200 import synth
from oracle import cport
import fastmodel

DEST, TMIN = [46.77751005, 355.0], 2.0223605


def model(scen, times=None):
    t, la, lo, u, v = synth.wind_fields(scen)
    if times is None:
        times, lats, lons = H.sim_axes(la, lo)
    else:
        lats, lons = la, lo
    return fastmodel.FastModel(t, la, lo, u, v, times, lats, lons)


def test_boat_speed_matches_get_deter_dyn():
    """Weather.returnPolarVel + point of sail + Boat.getDeterDyn: absolute error below 1e-6 m/s
    over the whole (wind, heading) domain, including the tack / gybe cosine extensions, calm wind and
    the wind clamp."""
    f, o = model(0), H.oracle_sim(0)
    rng = np.random.default_rng(0)
    worst = 0.0
    for _ in range(20000):
        mag, hd, wd = rng.uniform(0.0, 25.0), rng.uniform(0, 360), rng.uniform(0, 360)
        u = float(np.float32(mag * np.sin(np.radians(wd))))
        v = float(np.float32(mag * np.cos(np.radians(wd))))
        m, ang = cport.polar_vel(u, v)
        want = o.deter_dyn(abs((ang + 180) % 360 - hd), m)
        worst = max(worst, abs(f.boat_speed(u, v, hd) - want))
    assert worst < 1e-6
    # a wind of exactly zero has no direction; the fp32 path does not give it one (the reference's atan2(0, 0) = 0 is the
    # literal kernel's business): the speed is not a number, and a transition in such a wind is handed over, not taken
    for hd in (0.0, 45.0, 200.0, 359.0):
        assert np.isnan(f.boat_speed(0.0, 0.0, hd))
    t, la, lo, u, v = synth.wind_fields(0)
    times, lats, lons = H.sim_axes(la, lo)
    calm = fastmodel.FastModel(t, la, lo, np.zeros_like(u), np.zeros_like(v), times, lats, lons)
    n = 8
    r = calm.step_batch(0, np.full(n, 44.0), np.full(n, 355.0), np.full((1, n), 45.0), nsteps=3, seg_len=3,
                        z=np.zeros((3, n)))
    assert r["needs_literal"].all() and (r["k"] == 0).all()
    got = calm.rollout_batch(2, 4, H.STATE0, DEST, TMIN, prefix=np.array([[0.0], [45.0]]), prefix_len=np.ones(2, np.int32),
                             z=np.zeros((13, 8)))
    assert (got["uncertain"] != 0).all()


def test_move_and_bearing_building_blocks():
    L = fastmodel.lib()
    o = H.oracle_sim(0)
    rng = np.random.default_rng(1)
    e_lat = e_lon = e_brg = 0.0
    for _ in range(20000):
        lat, lon, hd = rng.uniform(-70, 70), rng.uniform(0, 360), rng.uniform(0, 360)
        dist = rng.uniform(-10e3, 60e3)
        a, b = C.c_double(), C.c_double()
        L.fm_move(C.c_float(dist / 6371e3), C.c_double(hd), C.c_double(lat), C.c_double(lon), C.byref(a), C.byref(b))
        want = o.get_destination(dist, hd, lat, lon)
        step_deg = abs(dist) / 6371e3 * 180 / np.pi
        e_lat = max(e_lat, abs(a.value - want[0]) / max(step_deg, 1e-3))
        e_lon = max(e_lon, abs(b.value - want[1]) * np.cos(np.radians(lat)) / max(step_deg, 1e-3))
        dl, do = lat + rng.uniform(-8, 8), lon + rng.uniform(-12, 12)
        sh, ch = C.c_float(), C.c_float()
        L.fm_bearing(C.c_double(lat), C.c_double(lon), C.c_double(dl), C.c_double(do), C.byref(sh), C.byref(ch))
        brg = o.dist_bearing(lat, lon, dl, do)[1]
        e = np.arctan2(sh.value, ch.value) - np.radians(brg)
        e_brg = max(e_brg, abs((e + np.pi) % (2 * np.pi) - np.pi))
    # increments are exact to fp32 relative accuracy *of the increment*
    assert e_lat < 1e-6 and e_lon < 1e-6 and e_brg < 1e-6


def test_raw_steps_golden_and_seeded():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "raw_steps_c3.npz"))
    f = model(0, times=g["times"])
    nsteps, n = g["traj_lat"].shape
    head = synth.octagon_headings(g["h0"], nsteps)
    r = f.step_batch(0, g["lat0"], g["lon0"], head, nsteps=nsteps, seg_len=25, z=g["z"], traj=True)
    assert (r["status"] == 0).all() and (r["k"] == nsteps).all() and not r["needs_literal"].any()
    assert np.max(np.abs(r["traj_lat"] - g["traj_lat"]) / g["traj_lat"]) < 2e-8
    assert np.max(np.abs(r["traj_lon"] - g["traj_lon"]) / g["traj_lon"]) < 2e-8
    # what the fp32 formulas do not cover is handed over, not approximated
    head2 = head.copy()
    head2[2, 0] = -45.0
    z2 = g["z"].copy()
    z2[60, 1] = 5000.0
    r = f.step_batch(0, g["lat0"], g["lon0"], head2, nsteps=nsteps, seg_len=25, z=z2)
    assert r["needs_literal"][0] == 1 and r["k"][0] == 50
    assert r["needs_literal"][1] == 1 and r["k"][1] == 60
    assert not r["needs_literal"][2:].any()


@pytest.mark.parametrize("scen,flags", [(4, 0), (0, 0), (7, 2), (37, 0), (63, 2)])  # (63: the last scenario of config 4)
def test_rollout_decisions_outside_margins_agree(scen, flags):
    """Every rollout the fp32 path decides itself takes the decisions of the oracle (steps, status,
    histogram bin), its trajectory stays within 2e-7 relative, and few rollouts are left undecided."""
    f, o = model(scen), H.oracle_sim(scen)
    rng = np.random.default_rng(100 + scen)
    n_leaf, rep = 256, 128
    n = n_leaf * rep
    prefix = 45.0 * rng.integers(0, 8, (n_leaf, 4))
    plen = rng.integers(1, 5, n_leaf).astype(np.int32)
    z = rng.standard_normal((13, n))
    got = f.rollout_batch(n_leaf, rep, H.STATE0, DEST, TMIN, prefix=prefix, prefix_len=plen, z=z, flags=flags,
                          traj_steps=13)
    want = o.rollout_batch(n, H.STATE0, DEST, TMIN, prefix=np.repeat(prefix, rep, 0),
                           prefix_len=np.repeat(plen, rep), replay_full=bool(flags & 2), z=z, nthreads=8,
                           traj_steps=13)
    ok = got["uncertain"] == 0
    assert 0.0 < (~ok).mean() < 0.01
    for key in ("steps", "status", "bin"):
        assert np.array_equal(got[key][ok], want[key][ok]), key
    el = np.abs(got["traj_lat"] - want["traj_lat"])[:, ok] / np.abs(want["traj_lat"][:, ok])
    eo = np.abs(got["traj_lon"] - want["traj_lon"])[:, ok] / np.abs(want["traj_lon"][:, ok])
    assert np.nanmax(el) < 2e-7 and np.nanmax(eo) < 2e-7
    assert np.array_equal(np.isnan(got["traj_lat"][:, ok]), np.isnan(want["traj_lat"][:, ok]))
    arr = ok & (want["status"] == 1)
    assert np.allclose(got["t_arr"][arr], want["t_arr"][arr], rtol=1e-6)
    assert np.allclose(got["reward"][ok], want["reward"][ok], rtol=1e-6, atol=3e-6)
    # the same launch without the trajectories is the "turbo" variant (position relative to the destination in radians,
    # cell and weights of the wind grid from the fp32 index): it answers to the oracle directly, not through the
    # generic code
    lean = f.rollout_batch(n_leaf, rep, H.STATE0, DEST, TMIN, prefix=prefix, prefix_len=plen, z=z, flags=flags)
    assert f.last_variant() == 2
    okl = lean["uncertain"] == 0
    assert 0.0 < (~okl).mean() < 0.01
    for key in ("steps", "status", "bin"):
        assert np.array_equal(lean[key][okl], want[key][okl]), key
    assert np.allclose(lean["final_lat"][okl], want["final_lat"][okl], rtol=3e-7, atol=0)
    assert np.allclose(lean["final_lon"][okl], want["final_lon"][okl], rtol=3e-7, atol=0)
    arr = okl & (want["status"] == 1)
    assert np.allclose(lean["t_arr"][arr], want["t_arr"][arr], rtol=1e-6)
    assert np.allclose(lean["reward"][okl], want["reward"][okl], rtol=1e-6, atol=3e-6)
    # with the margins closed the fp32 path would still be right almost always: the margins are slack
    f.set_margins(angle=0.0, along=0.0, near_miss2=0.0, box=0.0, bin=0.0)
    for kw in (dict(traj_steps=13), dict()):
        raw = f.rollout_batch(n_leaf, rep, H.STATE0, DEST, TMIN, prefix=prefix, prefix_len=plen, z=z, flags=flags, **kw)
        bad = (raw["steps"] != want["steps"]) | (raw["status"] != want["status"]) | (raw["bin"] != want["bin"])
        assert bad[raw["uncertain"] == 0].sum() <= 2


def test_specialised_variants_equal_the_generic_one():
    """rollout_one's compile-time variants (the reference's boat as immediates; "turbo": range checks, bounds
    test, step-length load and sin / cos refresh compiled out where the host proved them redundant) decide and
    move like the generic code: same steps, status, bin, hand-overs; positions bit-identical, except that the turbo
    variant carries its position in radians inside the loop (equal to fp32 rounding, 5e-8 relative)."""
    f = model(2)
    rng = np.random.default_rng(77)
    n_leaf, rep = 128, 64
    prefix = 45.0 * rng.integers(0, 8, (n_leaf, 4))
    plen = rng.integers(1, 5, n_leaf).astype(np.int32)
    z = rng.standard_normal((13, n_leaf * rep))
    for flags, want_variant in ((0, 2), (2, 2), (4, 1)):   # tree, full-prefix replay: turbo; plan evaluation: default boat
        f.force_generic(False)
        a = f.rollout_batch(n_leaf, rep, H.STATE0, DEST, TMIN, prefix=prefix, prefix_len=plen, z=z, flags=flags)
        assert f.last_variant() == want_variant
        f.force_generic(True)
        b = f.rollout_batch(n_leaf, rep, H.STATE0, DEST, TMIN, prefix=prefix, prefix_len=plen, z=z, flags=flags)
        assert f.last_variant() == 0
        # (the turbo variant rounds differently -- radians relative to the destination, fp32 cell split -- so a rollout on
        #  the edge of a margin may be decided by one variant and left to the literal kernel by the other: decisions are
        #  compared where both decided, and such rollouts must be rare)
        both = (a["uncertain"] == 0) & (b["uncertain"] == 0)
        assert (a["uncertain"] != b["uncertain"]).mean() < (1e-3 if want_variant == 2 else 1e-12)
        for key in ("steps", "status", "bin"):
            assert np.array_equal(a[key][both], b[key][both]), (flags, key)
        for key in ("final_lat", "final_lon", "reward"):
            if want_variant == 2:
                tol = 3e-7 if key != "reward" else 1e-6
                assert np.allclose(a[key][both], b[key][both], rtol=tol, atol=0 if key != "reward" else 3e-6, equal_nan=True), (flags, key)
            else:
                assert np.array_equal(a[key], b[key], equal_nan=True), (flags, key)
    # a root that cannot start a step, or sits outside the grid, is not turbo: the generic code reports it
    f.force_generic(False)
    r = f.rollout_batch(4, 2, (12, 44.0, 355.0), DEST, TMIN, prefix=prefix[:4], prefix_len=plen[:4], z=z[:, :8])
    assert f.last_variant() == 1 and (r["status"] == 6).all()      # PMCTS_ST_PAST_HORIZON
    r = f.rollout_batch(4, 2, (0, 39.0, 355.0), DEST, TMIN, prefix=prefix[:4], prefix_len=plen[:4], z=z[:, :8])
    assert f.last_variant() == 1 and (r["status"] == 4).all()      # PMCTS_ST_OUT_OF_GRID


def test_plan_evaluation_mode_and_golden_rollouts(golden_dir):
    g = np.load(os.path.join(golden_dir, "rollouts_c1.npz"))
    pe = np.load(os.path.join(golden_dir, "plan_eval_c5.npz"))
    dest, tmin = g["destination"].tolist(), float(g["timemin"])
    S, nT, n = g["traj_lat"].shape
    for si, scen in enumerate(g["scenarios"]):
        f = model(int(scen))
        z = np.ascontiguousarray(g["zroll"][si].T)
        r = f.rollout_batch(n, 1, H.STATE0, dest, tmin, prefix=g["prefix"], prefix_len=g["prefix_len"], z=z,
                            traj_steps=nT)
        ok = r["uncertain"] == 0
        assert ok.mean() > 0.95
        assert np.array_equal(r["steps"][ok], g["steps"][si][ok])
        assert np.array_equal(r["status"][ok] == 1, g["atdest"][si][ok] == 1)
        assert np.nanmax(np.abs(r["traj_lat"] - g["traj_lat"][si])[:, ok] / g["traj_lat"][si][:, ok]) < 2e-7
        # plan evaluation: arrival time or times[-1]
        nt2 = int(pe["plan_ntra"])
        plan = pe["plan_plan"].astype(np.float64)
        zz = np.ascontiguousarray(pe["plan_z"][si * nt2:(si + 1) * nt2].T)
        o = H.oracle_sim(int(scen))
        want = o.rollout_batch(nt2, H.STATE0, dest, 0.0, prefix=np.tile(plan, (nt2, 1)), mode=1, z=zz)
        got = f.rollout_batch(nt2, 1, H.STATE0, dest, 0.0, prefix=np.tile(plan, (nt2, 1)), z=zz, flags=4)
        ok = got["uncertain"] == 0
        assert np.array_equal(got["status"][ok], want["status"][ok])
        assert np.allclose(got["t_arr"][ok], want["t_arr"][ok], rtol=1e-6)


def test_other_hemisphere_and_grid():
    wt, wlat, wlon, u, v, times, lats, lons = H.southern_world()
    f = fastmodel.FastModel(wt, wlat, wlon, u, v, times, lats, lons)
    o = cport.OracleSim(wt, wlat, wlon, u, v, times, lats, lons)
    root, dest, tmin = (0, -36.0, -52.0), [-36.7, -52.9], 0.6
    rng = np.random.default_rng(5)
    n_leaf, rep = 64, 64
    n = n_leaf * rep
    prefix = 45.0 * rng.integers(0, 8, (n_leaf, 3))
    plen = rng.integers(1, 4, n_leaf).astype(np.int32)
    z = rng.standard_normal((times.size, n))
    for flags in (0, 2, 4):
        got = f.rollout_batch(n_leaf, rep, root, dest, tmin, prefix=prefix, prefix_len=plen, z=z, flags=flags,
                              traj_steps=times.size)
        want = o.rollout_batch(n, root, dest, tmin, prefix=np.repeat(prefix, rep, 0), prefix_len=np.repeat(plen, rep),
                               replay_full=bool(flags & 2), mode=1 if flags & 4 else 0, z=z, nthreads=8,
                               traj_steps=times.size)
        ok = got["uncertain"] == 0
        assert ok.mean() > 0.97
        for key in ("steps", "status", "bin"):
            assert np.array_equal(got[key][ok], want[key][ok]), (flags, key)
        el = np.abs(got["traj_lat"] - want["traj_lat"])[:, ok] / np.abs(want["traj_lat"][:, ok])
        eo = np.abs(got["traj_lon"] - want["traj_lon"])[:, ok] / np.abs(want["traj_lon"][:, ok])
        assert np.nanmax(el) < 3e-7 and np.nanmax(eo) < 3e-7
        assert (want["status"] == 1).mean() > 0.2   # a good share arrives, the rest runs out of time or leaves the box
    # raw steps with arbitrary (non multiple-of-45) headings
    nb = 512
    lat0, lon0 = rng.uniform(-40, -36, nb), rng.uniform(-58, -52, nb)
    head = rng.uniform(0, 360, (3, nb))
    zz = rng.standard_normal((24, nb))
    w = o.step_batch(0, lat0, lon0, head, nsteps=24, seg_len=8, z=zz, nthreads=8)
    g = f.step_batch(0, lat0, lon0, head, nsteps=24, seg_len=8, z=zz)
    same = w["status"] == 0
    assert np.array_equal(g["status"], w["status"]) and np.array_equal(g["k"], w["k"]) and same.sum() > nb // 2
    assert np.max(np.abs(g["lat"][same] - w["lat"][same]) / np.abs(w["lat"][same])) < 1e-7
    assert np.max(np.abs(g["lon"][same] - w["lon"][same]) / np.abs(w["lon"][same])) < 1e-7
