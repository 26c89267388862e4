"""GPU parity tests: the CUDA path (through the C ABI) against the CPU restatement (oracle/) on the
same seeded inputs, and against the golden fixtures generated from the reference itself.

Tolerances (BASELINE.json north_star): per-step lat / lon within 1e-6 relative, arrival to the step,
frac / arrival time within 1e-6; integer outputs (steps, status, bins, histograms) bit-exact.
The literal fp64 mode is additionally held to 1e-11 (it differs from the reference only by libm ulps).
"""
import json
import os

import numpy as np
import pytest

import pmcts_helpers as H
from iboat_pmcts_b200 import _native as N
from iboat_pmcts_b200 import synth
from oracle import cport

pytestmark = pytest.mark.gpu

RTOL = 1e-6
MODES = [N.PREC_F64, N.PREC_FAST]


def rtol_for(mode):
    return 1e-11 if mode == N.PREC_F64 else RTOL


def tol_frac(mode):
    """frac is a fraction of one step; the arrival time it feeds is held to 1e-6 relative below."""
    return dict(rtol=1e-6, atol=1e-9) if mode == N.PREC_F64 else dict(rtol=0, atol=1e-5)


def tol_reward(mode):
    """Rewards go to zero near the horizon, so the fast mode is held to an absolute tolerance."""
    return dict(rtol=1e-6, atol=1e-12) if mode == N.PREC_F64 else dict(rtol=1e-6, atol=3e-6)


@pytest.fixture(scope="module")
def eng():
    from iboat_pmcts_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def upload(eng, slot, scenario, times=None):
    t, la, lo, u, v = synth.wind_fields(scenario)
    if times is None:
        times, lats, lons = H.sim_axes(la, lo)
    else:
        lats, lons = la, lo
    eng.upload_scenario(slot, t, la, lo, u, v, times, [lats[0], lats[-1], lons[0], lons[-1]])
    return times


def relerr(a, b):
    return np.nanmax(np.abs(a - b) / np.abs(b))


def test_get_wind_matches_reference(eng, golden_dir):
    with open(os.path.join(golden_dir, "kat.json")) as f:
        kat = json.load(f)
    t, la, lo, u, v = synth.wind_fields(0)
    pts = np.array(kat["getWind_scenario0"])
    # one simulator instant per query time, so the time-blended slabs are exercised at arbitrary times
    order = np.argsort(pts[:, 0])
    pts = pts[order]
    eng.upload_scenario(7, t, la, lo, u, v, pts[:, 0], [40, 50, 345, 360])
    uu, vv, st = eng.get_wind(7, np.arange(len(pts)), pts[:, 1], pts[:, 2])
    assert (st == 0).all()
    assert np.allclose(uu, pts[:, 3], rtol=1e-12, atol=1e-13)
    assert np.allclose(vv, pts[:, 4], rtol=1e-12, atol=1e-13)
    _, _, st = eng.get_wind(7, [0, 0], [39.99, 45.0], [350.0, 360.01])
    assert (st == N.ST_OUT_OF_GRID).all()


@pytest.mark.parametrize("mode", MODES)
def test_raw_steps_config3_golden(eng, golden_dir, mode):
    """BASELINE config 3 shape against trajectories recorded from the reference (24 boats x 500 steps)."""
    g = np.load(os.path.join(golden_dir, "raw_steps_c3.npz"))
    eng.set_precision(mode)
    upload(eng, 3, 0, times=g["times"])
    nsteps, n = g["traj_lat"].shape
    head = synth.octagon_headings(g["h0"], nsteps)
    k = np.zeros(n, np.int32)
    lat, lon = g["lat0"].copy(), g["lon0"].copy()
    pl, po = lat.copy(), lon.copy()
    st = np.zeros(n, np.uint8)
    tl, to = np.empty((nsteps, n)), np.empty((nsteps, n))
    eng.step_batch(3, k, lat, lon, head, nsteps=nsteps, seg_len=25, prev_lat=pl, prev_lon=po, status=st,
                   z=np.ascontiguousarray(g["z"]), traj_lat=tl, traj_lon=to)
    assert (st == 0).all() and (k == nsteps).all()
    assert relerr(tl, g["traj_lat"]) < rtol_for(mode)
    assert relerr(to, g["traj_lon"]) < rtol_for(mode)
    assert np.array_equal(lat, tl[-1]) and np.array_equal(pl, tl[-2])
    # stepping at the last time index is the reference's IndexError (simulatorTLKT.py:208)
    eng.step_batch(3, k, lat, lon, head[:1], nsteps=1, status=st, z=np.zeros((1, n)))
    assert (st == N.ST_PAST_HORIZON).all() and (k == nsteps).all()


@pytest.mark.parametrize("mode", MODES)
def test_raw_steps_vs_oracle_seeded(eng, mode):
    """Larger seeded sweep against the oracle, one launch per step and fused, Philox and injected noise."""
    eng.set_precision(mode)
    nsteps, n = 200, 4096
    times = np.linspace(0, 5, 501)
    upload(eng, 3, 1, times=times)
    o = H.oracle_sim(1, times=times)
    lat0, lon0, h0 = synth.octagon_fleet(n, seed=9)
    head = synth.octagon_headings(h0, nsteps)
    want = o.step_batch(0, lat0, lon0, head, nsteps=nsteps, seg_len=25, seed=42, stream=3, id0=100, nthreads=8)
    k = np.zeros(n, np.int32)
    lat, lon = lat0.copy(), lon0.copy()
    st = np.zeros(n, np.uint8)
    eng.step_batch(3, k, lat, lon, head, nsteps=nsteps, seg_len=25, status=st, seed=42, stream=3, id0=100)
    # in-kernel Philox: integers identical, Box-Muller in fp32 (fast) or fp64
    tol = 1e-11 if mode == N.PREC_F64 else RTOL
    assert (st == 0).all() and np.array_equal(k, want["k"])
    assert relerr(lat, want["lat"]) < tol and relerr(lon, want["lon"]) < tol
    # one launch per step == fused
    z = np.random.default_rng(3).standard_normal((20, n))
    k1 = np.zeros(n, np.int32)
    a1, o1 = lat0.copy(), lon0.copy()
    eng.step_batch(3, k1, a1, o1, head, nsteps=20, seg_len=25, z=z)
    k2 = np.zeros(n, np.int32)
    a2, o2 = lat0.copy(), lon0.copy()
    for j in range(20):
        eng.step_batch(3, k2, a2, o2, head[0], nsteps=1, z=z[j:j + 1])
    assert np.array_equal(k1, k2)
    if mode == N.PREC_F64:
        assert np.array_equal(a1, a2) and np.array_equal(o1, o2)
    else:  # the fused launch carries sin / cos(lat) between steps (refreshed every 16): last-bit differences
        assert relerr(a1, a2) < 1e-9 and relerr(o1, o2) < 1e-9
    w = o.step_batch(0, lat0, lon0, head, nsteps=20, seg_len=25, z=z, nthreads=8)
    assert relerr(a1, w["lat"]) < rtol_for(mode) and relerr(o1, w["lon"]) < rtol_for(mode)


@pytest.mark.parametrize("mode", MODES)
def test_rollouts_golden(eng, golden_dir, mode):
    """Default-policy rollouts recorded from the reference's Tree.default_policy (3 scenarios)."""
    g = np.load(os.path.join(golden_dir, "rollouts_c1.npz"))
    eng.set_precision(mode)
    dest, tmin = g["destination"].tolist(), float(g["timemin"])
    S, nT, n = g["traj_lat"].shape
    for si, scen in enumerate(g["scenarios"]):
        upload(eng, si, int(scen))
        z = np.ascontiguousarray(g["zroll"][si].T)
        r = eng.rollout_batch_host(si, n, 1, H.STATE0, dest, tmin, prefix=g["prefix"], prefix_len=g["prefix_len"],
                                   z=z, traj_steps=nT)
        assert np.array_equal(r["steps"], g["steps"][si])
        assert np.array_equal(r["status"] == N.ST_ARRIVED, g["atdest"][si] == 1)
        assert relerr(r["traj_lat"], g["traj_lat"][si]) < rtol_for(mode)
        assert relerr(r["traj_lon"], g["traj_lon"][si]) < rtol_for(mode)
        assert np.array_equal(np.isnan(r["traj_lat"]), np.isnan(g["traj_lat"][si]))
        arrived = g["atdest"][si] == 1
        assert arrived.any()
        assert np.allclose(r["frac"][arrived], g["frac"][si][arrived], **tol_frac(mode))
        assert np.allclose(r["reward"], g["reward"][si], **tol_reward(mode))


@pytest.mark.parametrize("mode", MODES)
def test_rollouts_vs_oracle_and_leaf_reduction(eng, mode):
    eng.set_precision(mode)
    upload(eng, 0, 4)
    o = H.oracle_sim(4)
    dest, tmin = [46.77751005, 355.0], 2.0223605
    n_leaf, rep = 37, 48  # ragged: leaves straddle warps
    n = n_leaf * rep
    rng = np.random.default_rng(1)
    prefix = 45.0 * rng.integers(0, 8, (n_leaf, 4))
    plen = rng.integers(1, 5, n_leaf).astype(np.int32)
    z = rng.standard_normal((13, n))
    for flags, full in ((0, False), (N.ROLL_REPLAY_FULL_PREFIX, True)):
        got = eng.rollout_batch_host(0, n_leaf, rep, H.STATE0, dest, tmin, prefix=prefix, prefix_len=plen, z=z,
                                     flags=flags)
        want = o.rollout_batch(n, H.STATE0, dest, tmin, prefix=np.repeat(prefix, rep, 0),
                               prefix_len=np.repeat(plen, rep), replay_full=full, z=z, nthreads=8)
        assert np.array_equal(got["steps"], want["steps"])
        assert np.array_equal(got["status"], want["status"])
        assert np.array_equal(got["bin"], want["bin"])
        assert relerr(got["final_lat"], want["final_lat"]) < rtol_for(mode)
        assert relerr(got["final_lon"], want["final_lon"]) < rtol_for(mode)
        assert np.allclose(got["reward"], want["reward"], **tol_reward(mode))
        arr = want["status"] == 1
        assert np.allclose(got["t_arr"][arr], want["t_arr"][arr], rtol=1e-6)
        assert np.isnan(got["t_arr"][~arr]).all()
        bins = np.zeros((n_leaf, 12), np.uint32)
        np.add.at(bins, (np.repeat(np.arange(n_leaf), rep), want["bin"]), 1)
        assert np.array_equal(got["leaf_bins"], bins)
        assert got["stats"][0] == arr.sum() and got["stats"][3] == arr.sum()
        assert np.isclose(got["stats"][1], want["t_arr"][arr].sum(), rtol=1e-9 if mode == N.PREC_F64 else 1e-6)


@pytest.mark.parametrize("mode", MODES)
def test_initialisation_and_plan_eval_golden(eng, golden_dir, mode):
    """initialize_simulators (forest.py:210-299) and estimate_perfomance_plan (isochrones.py:469-580)
    rebuilt from K1 / K2 launches, against values recorded from the reference."""
    g = np.load(os.path.join(golden_dir, "rollouts_c1.npz"))
    pe = np.load(os.path.join(golden_dir, "plan_eval_c5.npz"))
    eng.set_precision(mode)
    ntra = int(g["ntra"])
    S = len(g["scenarios"])
    nT = g["times"].size
    means = []
    osim = H.oracle_sim(0)
    for s, scen in enumerate(g["scenarios"]):
        upload(eng, s, int(scen))
        z = np.ascontiguousarray(g["zinit"][s * ntra:(s + 1) * ntra].T)
        k = np.zeros(ntra, np.int32)
        lat, lon = np.full(ntra, 44.0), np.full(ntra, 355.0)
        st = np.zeros(ntra, np.uint8)
        eng.step_batch(s, k, lat, lon, np.zeros((1, ntra)), nsteps=nT, seg_len=nT, status=st, z=z,
                       flags=N.STEP_STOP_BEFORE_TERMINAL)
        assert (st == N.ST_HORIZON).all() and (k == nT - 2).all()
        means.append(np.mean([osim.dist_bearing(44.0, 355.0, a, b)[0] for a, b in zip(lat, lon)]))
    dest = osim.get_destination(np.min(means), 0, 44.0, 355.0)
    assert np.allclose(dest, g["destination"], rtol=1e-6)
    tmins = []
    for s in range(S):
        z = np.ascontiguousarray(g["zinit"][(S + s) * ntra:(S + s + 1) * ntra].T)
        r = eng.rollout_batch_host(s, ntra, 1, H.STATE0, g["destination"].tolist(), 0.0, z=z, flags=N.ROLL_PLAN_EVAL)
        tmins.append(r["t_arr"][r["status"] == N.ST_ARRIVED].min())
    assert np.isclose(min(tmins), float(g["timemin"]), rtol=1e-6)
    for name in ("plan", "empty"):
        nt2 = int(pe[name + "_ntra"])
        plan = pe[name + "_plan"].astype(np.float64)
        means, all_t = [], []
        for s in range(S):
            z = np.ascontiguousarray(pe[name + "_z"][s * nt2:(s + 1) * nt2].T)
            pre = np.tile(plan, (nt2, 1)) if plan.size else None
            r = eng.rollout_batch_host(s, nt2, 1, H.STATE0, g["destination"].tolist(), 0.0, prefix=pre, z=z,
                                       flags=N.ROLL_PLAN_EVAL)
            means.append(r["t_arr"].mean())
            all_t.append(r["t_arr"])
            assert np.isclose(r["stats"][1] / r["stats"][0], means[-1], rtol=1e-12)
        all_t = np.concatenate(all_t)
        assert np.allclose([all_t.mean()] + means, pe[name + "_mean"], rtol=1e-6)
        v = [np.mean((all_t - all_t.mean()) ** 2)] + [np.mean((t - t.mean()) ** 2) for t in np.split(all_t, S)]
        assert np.allclose(v, pe[name + "_var"], rtol=1e-4, atol=1e-12)


def test_fast_mode_large_batch_decisions_and_redo(eng):
    """10^5 rollouts in the mixed-precision mode: every integer output equals the oracle's (the decisions
    the fp32 arithmetic cannot take are re-taken by the fp64 kernel), only a small fraction needs that,
    and with margins forced wide open the whole batch goes through the fp64 kernel and reproduces the
    literal mode bit for bit."""
    upload(eng, 0, 6)
    o = H.oracle_sim(6)
    dest, tmin = [46.77751005, 355.0], 2.0223605
    n_leaf, rep = 512, 200
    n = n_leaf * rep
    rng = np.random.default_rng(21)
    prefix = 45.0 * rng.integers(0, 8, (n_leaf, 3))
    plen = rng.integers(1, 4, n_leaf).astype(np.int32)
    z = rng.standard_normal((13, n))
    want = o.rollout_batch(n, H.STATE0, dest, tmin, prefix=np.repeat(prefix, rep, 0),
                           prefix_len=np.repeat(plen, rep), z=z, nthreads=8)
    bins = np.zeros((n_leaf, 12), np.uint32)
    np.add.at(bins, (np.repeat(np.arange(n_leaf), rep), want["bin"]), 1)
    eng.set_precision(N.PREC_F64)
    lit = eng.rollout_batch_host(0, n_leaf, rep, H.STATE0, dest, tmin, prefix=prefix, prefix_len=plen, z=z)
    eng.set_precision(N.PREC_FAST)
    got = eng.rollout_batch_host(0, n_leaf, rep, H.STATE0, dest, tmin, prefix=prefix, prefix_len=plen, z=z)
    redo = eng.last_redo_count()
    assert 0 < redo < 0.02 * n
    for key in ("steps", "status", "bin"):
        assert np.array_equal(got[key], want[key]), key
    assert np.array_equal(got["leaf_bins"], bins)
    assert relerr(got["final_lat"], want["final_lat"]) < RTOL and relerr(got["final_lon"], want["final_lon"]) < RTOL
    arr = want["status"] == 1
    assert np.allclose(got["t_arr"][arr], want["t_arr"][arr], rtol=1e-6)
    assert got["stats"][5] == n and got["stats"][4] == want["steps"].sum() and got["stats"][3] == arr.sum()
    # all margins wide open: every rollout is handed to the fp64 kernel
    eng.set_fast_margins(angle=1e30, along=1e30, near_miss2=1e30, box=1e30, bin=1e30)
    try:
        allredo = eng.rollout_batch_host(0, n_leaf, rep, H.STATE0, dest, tmin, prefix=prefix, prefix_len=plen, z=z)
        assert eng.last_redo_count() == n
    finally:
        eng.set_fast_margins()
    for key in ("steps", "status", "bin", "reward", "t_arr", "final_lat", "final_lon", "leaf_bins"):
        assert np.array_equal(allredo[key], lit[key], equal_nan=True), key
    assert allredo["stats"][5] == n
    eng.set_precision(N.PREC_F64)


def test_kernel_variants_agree_with_each_other_and_the_oracle(eng):
    """The throughput launches (histogram / statistics outputs only) pick a compile-time variant of K2:
    TURBO for tree-mode launches on the reference's boat, DEFAULT_BOAT for plan evaluation, GENERIC for any
    other pmcts_set_params, FULL_OUTPUT when per-rollout arrays are requested.  Same leaf histograms and
    statistics from all of them, injected and Philox noise, equal to the oracle's bins."""
    upload(eng, 0, 5)
    o = H.oracle_sim(5)
    dest, tmin = [46.77751005, 355.0], 2.0223605
    n_leaf, rep = 256, 96
    n = n_leaf * rep
    rng = np.random.default_rng(33)
    prefix = 45.0 * rng.integers(0, 8, (n_leaf, 3))
    plen = rng.integers(1, 4, n_leaf).astype(np.int32)
    z = rng.standard_normal((13, n))
    eng.set_precision(N.PREC_FAST)
    lean = ("leaf_bins", "stats")
    try:
        for flags, fast_variant in ((0, N.VARIANT_TURBO), (N.ROLL_REPLAY_FULL_PREFIX, N.VARIANT_TURBO),
                                    (N.ROLL_PLAN_EVAL, N.VARIANT_DEFAULT_BOAT)):
            want = o.rollout_batch(n, H.STATE0, dest, tmin, prefix=np.repeat(prefix, rep, 0),
                                   prefix_len=np.repeat(plen, rep), replay_full=bool(flags & N.ROLL_REPLAY_FULL_PREFIX),
                                   mode=1 if flags & N.ROLL_PLAN_EVAL else 0, z=z, nthreads=8)
            bins = np.zeros((n_leaf, 12), np.uint32)
            np.add.at(bins, (np.repeat(np.arange(n_leaf), rep), want["bin"]), 1)
            res = {}
            for noise in ("z", "philox"):
                kw = dict(z=z) if noise == "z" else dict(seed=99, stream=3)
                a = eng.rollout_batch_host(0, n_leaf, rep, H.STATE0, dest, tmin, prefix=prefix, prefix_len=plen,
                                           flags=flags, want=lean, **kw)
                assert eng.last_kernel_variant() == fast_variant
                b = eng.rollout_batch_host(0, n_leaf, rep, H.STATE0, dest, tmin, prefix=prefix, prefix_len=plen,
                                           flags=flags, **kw)
                assert eng.last_kernel_variant() == N.VARIANT_FULL_OUTPUT
                # any non-default tunable (here a slightly wider margin): generic kernel
                eng.set_fast_margins(angle=2.5e-4)
                c = eng.rollout_batch_host(0, n_leaf, rep, H.STATE0, dest, tmin, prefix=prefix, prefix_len=plen,
                                           flags=flags, want=lean, **kw)
                assert eng.last_kernel_variant() == N.VARIANT_GENERIC
                eng.set_fast_margins()
                others = [b, c]
                if fast_variant == N.VARIANT_TURBO:
                    # two rollouts per thread on the packed FP32 pipe (PMCTS_OPT_PACKED_FP32): odd and even batch sizes
                    eng.set_option(N.OPT_PACKED_FP32, 1)
                    try:
                        d = eng.rollout_batch_host(0, n_leaf, rep, H.STATE0, dest, tmin, prefix=prefix, prefix_len=plen,
                                                   flags=flags, want=lean, **kw)
                        assert eng.last_kernel_variant() == N.VARIANT_TURBO_PACKED
                        others.append(d)
                        if noise == "philox":
                            odd = [eng.rollout_batch_host(0, 37, 3, H.STATE0, dest, tmin, prefix=prefix[:37],
                                                          prefix_len=plen[:37], flags=flags, want=lean, **kw)
                                   for _ in (eng.set_option(N.OPT_PACKED_FP32, 1), eng.set_option(N.OPT_PACKED_FP32, 0))]
                            assert np.array_equal(odd[0]["leaf_bins"], odd[1]["leaf_bins"]) and odd[0]["stats"][5] == 111
                    finally:
                        eng.set_option(N.OPT_PACKED_FP32, 0)
                for other in others:
                    assert np.array_equal(a["leaf_bins"], other["leaf_bins"])
                    assert np.array_equal(a["stats"][[0, 3, 4, 5]], other["stats"][[0, 3, 4, 5]])
                    assert np.allclose(a["stats"][1:3], other["stats"][1:3], rtol=1e-7, equal_nan=True)  # arrival times: 1e-6 each
                res[noise] = a
            assert np.array_equal(res["z"]["leaf_bins"], bins)
            assert res["z"]["stats"][4] == want["steps"].sum() and res["z"]["stats"][3] == (want["status"] == 1).sum()
        eng.set_precision(N.PREC_F64)
        eng.rollout_batch_host(0, n_leaf, rep, H.STATE0, dest, tmin, prefix=prefix, prefix_len=plen, z=z, want=lean)
        assert eng.last_kernel_variant() == N.VARIANT_LITERAL
    finally:
        eng.set_fast_margins()
        eng.set_precision(N.PREC_F64)


@pytest.mark.parametrize("mode", MODES)
def test_action_index_headings(eng, mode):
    """PMCTS_STEP_HEADING_ACTIONS: headings given as uint8 indices into the reference's ACTIONS (45 degrees each)
    move the boats exactly like the same headings in fp64 degrees -- host arrays and device tensors."""
    import torch
    eng.set_precision(mode)
    times = np.linspace(0, 5, 101)
    upload(eng, 3, 1, times=times)
    n, nsteps = 3000, 60
    lat0, lon0, h0 = synth.octagon_fleet(n, seed=3)
    acts = ((h0[None, :] + np.arange(6)[:, None]) % 8).astype(np.uint8)       # [nseg][n]
    res = []
    for head in (45.0 * acts.astype(np.float64), acts):
        k, st, lat, lon = np.zeros(n, np.int32), np.zeros(n, np.uint8), lat0.copy(), lon0.copy()
        eng.step_batch(3, k, lat, lon, head, nsteps=nsteps, seg_len=10, status=st, seed=9, stream=1)
        res.append((k, st, lat, lon))
    for a, b in zip(*res):
        assert np.array_equal(a, b)
    dev = torch.device("cuda:0")
    eng.use_torch_stream()
    k, st = torch.zeros(n, dtype=torch.int32, device=dev), torch.zeros(n, dtype=torch.uint8, device=dev)
    lat, lon = torch.tensor(lat0, device=dev), torch.tensor(lon0, device=dev)
    eng.step_batch(3, k, lat, lon, torch.tensor(acts, device=dev).reshape(-1), nsteps=nsteps, seg_len=10, status=st,
                   seed=9, stream=1)
    torch.cuda.synchronize()
    eng.use_own_stream()
    assert np.array_equal(lat.cpu().numpy(), res[0][2]) and np.array_equal(lon.cpu().numpy(), res[0][3])
    assert (res[0][0] == nsteps).all()
    eng.set_precision(N.PREC_F64)


def test_fast_mode_hands_over_what_it_does_not_cover(eng):
    """K1 in the mixed-precision mode: headings outside [0, 360) and steps longer than its series allow
    are finished by the fp64 kernel from the step they stopped at -- same result as the literal mode."""
    times = np.linspace(0, 5, 41)
    upload(eng, 3, 2, times=times)
    n, nsteps = 1000, 40
    lat0, lon0, h0 = synth.octagon_fleet(n, seed=4)
    head = synth.octagon_headings(h0, nsteps, seg_len=5)
    # 360.0 is outside [0, 360) (the fast path declines it) yet physically sane: the reference's fold of
    # |wind_from - 360| gives the same point of sail as heading 0
    head[3:, ::7] = 360.0
    z = np.random.default_rng(8).standard_normal((nsteps, n))
    z[17, ::11] = 4000.0      # one absurd gust: a step of ~0.5 rad
    res = {}
    for mode in MODES:
        eng.set_precision(mode)
        k = np.zeros(n, np.int32)
        lat, lon = lat0.copy(), lon0.copy()
        st = np.zeros(n, np.uint8)
        eng.step_batch(3, k, lat, lon, head, nsteps=nsteps, seg_len=5, status=st, z=z)
        res[mode] = (k, lat, lon, st)
    kf, laf, lof, stf = res[N.PREC_FAST]
    kl, lal, lol, stl = res[N.PREC_F64]
    assert np.array_equal(kf, kl) and np.array_equal(stf, stl)
    assert (stf <= N.ST_PAST_HORIZON).all()
    ok = stl == 0
    assert ok.sum() > n // 2
    assert relerr(laf[ok], lal[ok]) < RTOL and relerr(lof[ok], lol[ok]) < RTOL
    eng.set_precision(N.PREC_F64)


def test_calm_wind_is_left_to_the_literal_kernel(eng):
    """A wind of exactly zero has no direction: the reference goes by atan2(0, 0) = 0 (weatherTLKT.py:161-162), the
    mixed-precision kernels give it none -- their speed is then not a number and the step's length test hands the
    rollout / boat to the fp64 kernel.  A band of dead calm across the route (exact zeros, so the interpolated wind is
    exactly zero inside it): K2 and K1 in the mixed-precision mode decide and move like the oracle."""
    from oracle import cport
    t, la, lo, u, v = synth.wind_fields(3)
    u, v = u.copy(), v.copy()
    band = (la >= 44.9) & (la <= 46.1)
    u[:, band, :] = 0.0
    v[:, band, :] = 0.0
    times, lats, lons = H.sim_axes(la, lo)
    eng.upload_scenario(7, t, la, lo, u, v, times, [lats[0], lats[-1], lons[0], lons[-1]])
    o = cport.OracleSim(t, la, lo, u, v, times, lats, lons, sigma_coeff=0.2)
    dest, tmin = [46.77751005, 355.0], 2.0223605
    n_leaf, rep = 16, 64
    n = n_leaf * rep
    rng = np.random.default_rng(5)
    prefix = 45.0 * rng.integers(0, 8, (n_leaf, 2))
    plen = rng.integers(1, 3, n_leaf).astype(np.int32)
    z = rng.standard_normal((13, n))
    want = o.rollout_batch(n, H.STATE0, dest, tmin, prefix=np.repeat(prefix, rep, 0), prefix_len=np.repeat(plen, rep), z=z,
                           nthreads=8)
    bins = np.zeros((n_leaf, 12), np.uint32)
    np.add.at(bins, (np.repeat(np.arange(n_leaf), rep), want["bin"]), 1)
    eng.set_precision(N.PREC_FAST)
    try:
        full = eng.rollout_batch_host(7, n_leaf, rep, H.STATE0, dest, tmin, prefix=prefix, prefix_len=plen, z=z)
        redo = eng.last_redo_count()
        lean = eng.rollout_batch_host(7, n_leaf, rep, H.STATE0, dest, tmin, prefix=prefix, prefix_len=plen, z=z,
                                      want=("leaf_bins", "stats"))
        assert eng.last_kernel_variant() == N.VARIANT_TURBO
        assert redo > 50                                   # (139 of the 1024 reach the band: the wind dies down towards it)
        for key in ("steps", "status", "bin"):
            assert np.array_equal(full[key], want[key]), key
        assert relerr(full["final_lat"], want["final_lat"]) < RTOL and relerr(full["final_lon"], want["final_lon"]) < RTOL
        assert np.array_equal(full["leaf_bins"], bins) and np.array_equal(lean["leaf_bins"], bins)
        assert lean["stats"][4] == want["steps"].sum()
        # K1: boats started inside the band and south of it, raw steps
        nb, nsteps = 256, 8
        lat0 = np.where(np.arange(nb) % 2 == 0, 45.5, 44.2) + 1e-3 * np.arange(nb)
        lon0 = np.full(nb, 355.0) + 2e-3 * np.arange(nb)
        head = np.tile(45.0 * (np.arange(nb) % 8), (1, 1)).astype(np.float64)
        zz = rng.standard_normal((nsteps, nb))
        ref = o.step_batch(0, lat0, lon0, head, nsteps=nsteps, seg_len=nsteps, z=zz)
        k = np.zeros(nb, np.int32)
        lat, lon = lat0.copy(), lon0.copy()
        st = np.zeros(nb, np.uint8)
        eng.step_batch(7, k, lat, lon, head, nsteps=nsteps, seg_len=nsteps, status=st, z=zz)
        assert np.array_equal(k, ref["k"]) and np.array_equal(st, ref["status"])
        assert relerr(lat, ref["lat"]) < RTOL and relerr(lon, ref["lon"]) < RTOL
    finally:
        eng.set_precision(N.PREC_F64)


@pytest.mark.parametrize("mode", MODES)
def test_rollout_forest_equals_per_scenario_launches(eng, mode):
    """One launch over several scenarios (pmcts_rollout_forest) == one pmcts_rollout_batch per scenario, with
    shared and with per-scenario leaf sets, Philox streams stream + s."""
    eng.set_precision(mode)
    scen_ids = [2, 5, 9]
    for slot, sid in enumerate(scen_ids):
        upload(eng, 10 + slot, sid)
    dest, tmin = [46.77751005, 355.0], 2.0223605
    n_leaf, rep = 40, 50
    n = n_leaf * rep
    S = len(scen_ids)
    rng = np.random.default_rng(5)
    for shared in (True, False):
        P = 1 if shared else S
        prefix = 45.0 * rng.integers(0, 8, (P, n_leaf, 2))
        plen = rng.integers(1, 3, (P, n_leaf)).astype(np.int32)
        out = dict(steps=np.empty(S * n, np.int32), status=np.empty(S * n, np.uint8), reward=np.empty(S * n),
                   leaf_bins=np.empty(S * n_leaf * 12, np.uint32), stats=np.empty(S * 8))
        eng.rollout_forest([10, 11, 12], n_leaf, rep, H.STATE0, dest, tmin, prefix.reshape(-1), plen.reshape(-1), 2,
                           prefix_shared=shared, seed=77, stream=5, id0=1000, out=out)
        for s in range(S):
            one = eng.rollout_batch_host(10 + s, n_leaf, rep, H.STATE0, dest, tmin, prefix=prefix[0 if shared else s],
                                         prefix_len=plen[0 if shared else s], seed=77, stream=5 + s, id0=1000)
            assert np.array_equal(out["steps"][s * n:(s + 1) * n], one["steps"])
            assert np.array_equal(out["status"][s * n:(s + 1) * n], one["status"])
            assert np.array_equal(out["reward"][s * n:(s + 1) * n], one["reward"])
            assert np.array_equal(out["leaf_bins"].reshape(S, n_leaf, 12)[s], one["leaf_bins"])
            assert np.array_equal(out["stats"].reshape(S, 8)[s][[0, 3, 4, 5]], one["stats"][[0, 3, 4, 5]])
            assert np.isclose(out["stats"].reshape(S, 8)[s][1], one["stats"][1], rtol=1e-12)
        # against the oracle: integer outputs exact
        o = H.oracle_sim(scen_ids[1])
        want = o.rollout_batch(n, H.STATE0, dest, tmin, prefix=np.repeat(prefix[0 if shared else 1], rep, 0),
                               prefix_len=np.repeat(plen[0 if shared else 1], rep), seed=77, stream=6, id0=1000)
        if mode == N.PREC_F64:
            assert np.array_equal(out["steps"][n:2 * n], want["steps"])
            assert np.array_equal(out["status"][n:2 * n], want["status"])
    # scenarios with different simulator axes cannot share a launch
    upload(eng, 13, 1, times=np.linspace(0, 5, 41))
    with pytest.raises(N.PmctsError):
        eng.rollout_forest([10, 13], n_leaf, rep, H.STATE0, dest, tmin, out=dict(stats=np.empty(16)))
    eng.set_precision(N.PREC_F64)


def test_deferred_finisher_gives_the_same_results(eng):
    """PMCTS_ROLL_DEFER_REDO: the fp64 finisher runs on a side stream and overlaps the next launch; after
    pmcts_join the outputs equal those of the ordinary (in-order) call, launch after launch."""
    import torch
    eng.set_precision(N.PREC_FAST)
    for slot, sid in enumerate((1, 4)):
        upload(eng, 20 + slot, sid)
    dev = torch.device("cuda:0")
    eng.use_torch_stream()
    try:
        dest, tmin = [46.77751005, 355.0], 2.0223605
        n_leaf, rep, S = 256, 128, 2
        prefix = torch.tensor(45.0 * (np.arange(n_leaf) % 8).astype(np.float64), device=dev)
        plen = torch.ones(n_leaf, dtype=torch.int32, device=dev)

        def call(seed, flags, bins, stats):
            eng.rollout_forest([20, 21], n_leaf, rep, H.STATE0, dest, tmin, prefix, plen, 1, prefix_shared=True,
                               seed=seed, out=dict(leaf_bins=bins, stats=stats), flags=flags)

        want = []
        for seed in range(5):
            b = torch.zeros(S * n_leaf * 12, dtype=torch.int32, device=dev)
            st = torch.zeros(S * 8, dtype=torch.float64, device=dev)
            call(seed, 0, b, st)
            torch.cuda.synchronize()
            assert eng.last_redo_count() > 0
            want.append((b.cpu().numpy(), st.cpu().numpy()))
        got = []
        for seed in range(5):   # five deferred launches back to back: the two work areas are reused
            b = torch.zeros(S * n_leaf * 12, dtype=torch.int32, device=dev)
            st = torch.zeros(S * 8, dtype=torch.float64, device=dev)
            call(seed, N.ROLL_DEFER_REDO, b, st)
            eng.join(keep_last=1)
            got.append((b, st))
        eng.join()
        snap = [(b.cpu().numpy(), st.cpu().numpy()) for b, st in got]  # on torch's stream, after the join
        for (b, st), (wb, wst) in zip(snap, want):
            assert np.array_equal(b, wb) and int(b.sum()) == S * n_leaf * rep
            assert np.array_equal(st[[0, 3, 4, 5, 8, 11, 12, 13]], wst[[0, 3, 4, 5, 8, 11, 12, 13]])
            assert np.allclose(st, wst, rtol=1e-12)
    finally:
        eng.use_own_stream()
        eng.set_precision(N.PREC_F64)


def test_deferred_host_arrays_and_wait(eng):
    """PMCTS_MEM_HOST | PMCTS_ROLL_DEFER_REDO: the call returns while its finisher and the copies back to the caller's
    (pinned) arrays still run on the side stream; after pmcts_wait the arrays hold what the synchronous call returns,
    with two calls in flight and the slots reused."""
    import torch
    eng.set_precision(N.PREC_FAST)
    for slot, sid in enumerate((1, 4)):
        upload(eng, 20 + slot, sid)
    try:
        dest, tmin = [46.77751005, 355.0], 2.0223605
        n_leaf, rep, S = 256, 128, 2
        prefix = 45.0 * (np.arange(n_leaf) % 8).astype(np.float64)
        plen = np.ones(n_leaf, np.int32)

        def pinned(n, dt):
            return torch.empty(n, dtype=dt, pin_memory=True).numpy()

        def call(seed, flags, out):
            eng.rollout_forest([20, 21], n_leaf, rep, H.STATE0, dest, tmin, prefix, plen, 1, prefix_shared=True,
                               seed=seed, out=out, flags=flags)

        want = []
        for seed in range(5):
            out = dict(leaf_bins=np.empty(S * n_leaf * 12, np.uint32), stats=np.empty(S * 8))
            call(seed, 0, out)
            want.append(out)
        outs = [dict(leaf_bins=pinned(S * n_leaf * 12, torch.int32).view(np.uint32), stats=pinned(S * 8, torch.float64))
                for _ in range(5)]
        for seed in range(5):
            outs[seed]["leaf_bins"][:] = 0xFFFFFFFF
            call(seed, N.ROLL_DEFER_REDO, outs[seed])
            eng.wait(keep_last=1)
            if seed:  # the previous call is complete, this one may still be running
                assert np.array_equal(outs[seed - 1]["leaf_bins"], want[seed - 1]["leaf_bins"])
        eng.wait()
        for got, w in zip(outs, want):
            assert np.array_equal(got["leaf_bins"], w["leaf_bins"]) and int(got["leaf_bins"].sum()) == S * n_leaf * rep
            assert np.array_equal(got["stats"][[0, 3, 4, 5, 8, 11, 12, 13]], w["stats"][[0, 3, 4, 5, 8, 11, 12, 13]])
            assert np.allclose(got["stats"], w["stats"], rtol=1e-12)
        # a synchronous host call right after deferred ones still works (its own staging buffer)
        out = dict(leaf_bins=np.empty(S * n_leaf * 12, np.uint32), stats=np.empty(S * 8))
        call(2, N.ROLL_DEFER_REDO, outs[0])
        call(3, 0, out)
        eng.wait()
        assert np.array_equal(out["leaf_bins"], want[3]["leaf_bins"]) and np.array_equal(outs[0]["leaf_bins"], want[2]["leaf_bins"])
    finally:
        eng.set_precision(N.PREC_F64)


def test_deferred_step_batch_pipeline(eng):
    """PMCTS_STEP_DEFER: copies in, kernel and copies out of consecutive host-array calls overlap; after pmcts_wait
    every call's arrays hold what the synchronous call returns (five calls, two in flight, both staging buffers
    reused), for fp64 and action-index headings."""
    import torch
    eng.set_precision(N.PREC_FAST)
    times = np.linspace(0, 5, 101)
    upload(eng, 3, 1, times=times)
    n, nsteps = 20000, 80
    lat0, lon0, h0 = synth.octagon_fleet(n, seed=5)
    acts = ((h0[None, :] + np.arange(4)[:, None]) % 8).astype(np.uint8)

    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True).numpy()
        t[...] = a
        return t

    try:
        for head in (45.0 * acts.astype(np.float64), acts):
            want = []
            for c in range(5):
                k, st, lat, lon = np.zeros(n, np.int32), np.zeros(n, np.uint8), lat0 + 0.01 * c, lon0.copy()
                eng.step_batch(3, k, lat, lon, head, nsteps=nsteps, seg_len=20, status=st, seed=9, stream=c)
                want.append((k, st, lat, lon))
            got, hp = [], pinned(head)
            for c in range(5):
                arrs = (pinned(np.zeros(n, np.int32)), pinned(np.zeros(n, np.uint8)), pinned(lat0 + 0.01 * c), pinned(lon0))
                eng.step_batch(3, arrs[0], arrs[2], arrs[3], hp, nsteps=nsteps, seg_len=20, status=arrs[1], seed=9,
                               stream=c, flags=N.STEP_DEFER)
                got.append(arrs)
                eng.wait(keep_last=1)
                if c:
                    for a, b in zip(got[c - 1], want[c - 1]):
                        assert np.array_equal(a, b)
            eng.wait()
            for g, w in zip(got, want):
                for a, b in zip(g, w):
                    assert np.array_equal(a, b)
            assert (want[0][0] == nsteps).all()
    finally:
        eng.set_precision(N.PREC_F64)


def test_deferred_step_batch_with_trajectories(eng):
    """PMCTS_STEP_DEFER with trajectory outputs over six pipelined calls: the NaN fill of call i + 2 must not land in
    a staging buffer whose copy back for call i is still running (each call's trajectories equal the synchronous
    call's, NaN pattern included)."""
    import torch
    eng.set_precision(N.PREC_FAST)
    times = np.linspace(0, 5, 101)
    upload(eng, 3, 1, times=times)
    n, nsteps = 60000, 40
    lat0, lon0, h0 = synth.octagon_fleet(n, seed=6)
    head = 45.0 * ((h0[None, :] + np.arange(2)[:, None]) % 8).astype(np.float64)

    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True).numpy()
        t[...] = a
        return t

    try:
        want = []
        for c in range(6):
            k, st, lat, lon = np.full(n, 70, np.int32), np.zeros(n, np.uint8), lat0 + 0.01 * c, lon0.copy()
            tl, to = np.empty((nsteps, n)), np.empty((nsteps, n))
            eng.step_batch(3, k, lat, lon, head, nsteps=nsteps, seg_len=20, status=st, seed=9, stream=c,
                           traj_lat=tl, traj_lon=to)
            want.append((k, st, lat, lon, tl, to))
        assert np.isnan(want[0][4][-1]).all() and not np.isnan(want[0][4][0]).any()  # boats stop at the horizon (k = 100)
        got, hp = [], pinned(head)
        for c in range(6):
            arrs = (pinned(np.full(n, 70, np.int32)), pinned(np.zeros(n, np.uint8)), pinned(lat0 + 0.01 * c),
                    pinned(lon0), pinned(np.zeros((nsteps, n))), pinned(np.zeros((nsteps, n))))
            eng.step_batch(3, arrs[0], arrs[2], arrs[3], hp, nsteps=nsteps, seg_len=20, status=arrs[1], seed=9,
                           stream=c, traj_lat=arrs[4], traj_lon=arrs[5], flags=N.STEP_DEFER)
            got.append(arrs)
        eng.wait()
        for g, w in zip(got, want):
            for a, b in zip(g, w):
                assert np.array_equal(a, b, equal_nan=True)
    finally:
        eng.set_precision(N.PREC_F64)


def test_injected_noise_must_cover_every_instant(eng):
    """z is [len(times)][n]: a shorter array (sized to the steps made, say) is refused instead of being read past
    its end."""
    upload(eng, 0, 0)
    nT = len(H.sim_axes(*synth.wind_fields(0)[1:3])[0])
    z = np.zeros((nT - 1, 8))
    with pytest.raises(ValueError):
        eng.rollout_batch_host(0, 8, 1, H.STATE0, [46.7, 355.0], 2.0, prefix=np.zeros((8, 1)), z=z)


@pytest.mark.parametrize("mode", MODES)
def test_other_hemisphere_and_grid(eng, mode):
    """Southern hemisphere, negative longitudes, 0.25-degree grid, hourly weather, 2-hour steps, arbitrary
    headings: the kernels against the oracle, all three rollout modes."""
    from oracle import cport
    eng.set_precision(mode)
    wt, wlat, wlon, u, v, times, lats, lons = H.southern_world()
    eng.upload_scenario(30, wt, wlat, wlon, u, v, times, [lats[0], lats[-1], lons[0], lons[-1]])
    o = cport.OracleSim(wt, wlat, wlon, u, v, times, lats, lons)
    root, dest, tmin = (0, -36.0, -52.0), [-36.7, -52.9], 0.6
    rng = np.random.default_rng(5)
    n_leaf, rep = 64, 64
    n = n_leaf * rep
    prefix = 45.0 * rng.integers(0, 8, (n_leaf, 3))
    plen = rng.integers(1, 4, n_leaf).astype(np.int32)
    z = rng.standard_normal((times.size, n))
    for flags in (0, N.ROLL_REPLAY_FULL_PREFIX, N.ROLL_PLAN_EVAL):
        got = eng.rollout_batch_host(30, n_leaf, rep, root, dest, tmin, prefix=prefix, prefix_len=plen, z=z, flags=flags)
        want = o.rollout_batch(n, root, dest, tmin, prefix=np.repeat(prefix, rep, 0), prefix_len=np.repeat(plen, rep),
                               replay_full=bool(flags & N.ROLL_REPLAY_FULL_PREFIX),
                               mode=1 if flags & N.ROLL_PLAN_EVAL else 0, z=z, nthreads=8)
        for key in ("steps", "status", "bin"):
            assert np.array_equal(got[key], want[key]), (flags, key)
        assert relerr(got["final_lat"], want["final_lat"]) < rtol_for(mode)
        assert relerr(got["final_lon"], want["final_lon"]) < rtol_for(mode)
        assert (want["status"] == 1).mean() > 0.2
    nb = 512
    lat0, lon0 = rng.uniform(-40, -36, nb), rng.uniform(-58, -52, nb)
    head = rng.uniform(0, 360, (3, nb))
    zz = rng.standard_normal((24, nb))
    w = o.step_batch(0, lat0, lon0, head, nsteps=24, seg_len=8, z=zz, nthreads=8)
    k = np.zeros(nb, np.int32)
    lat, lon = lat0.copy(), lon0.copy()
    st = np.zeros(nb, np.uint8)
    eng.step_batch(30, k, lat, lon, head, nsteps=24, seg_len=8, status=st, z=zz)
    same = w["status"] == 0
    assert np.array_equal(st, w["status"]) and np.array_equal(k, w["k"]) and same.sum() > nb // 2
    assert relerr(lat[same], w["lat"][same]) < rtol_for(mode) and relerr(lon[same], w["lon"][same]) < rtol_for(mode)
    eng.set_precision(N.PREC_F64)


def test_staged_slabs_give_the_same_boats(eng):
    """K1 with the step's wind slab staged in shared memory (TMA bulk copy, blocks whose boats share the time
    index) == K1 reading every gather through L1 / L2, bit for bit; blocks with mixed time indices, frozen
    boats and boats that run off the time axis take the same answers either way."""
    eng.set_precision(N.PREC_FAST)
    nsteps, n = 60, 5000
    times = np.linspace(0, 5, 101)
    upload(eng, 3, 1, times=times)
    lat0, lon0, h0 = synth.octagon_fleet(n, seed=12)
    head = synth.octagon_headings(h0, nsteps, seg_len=7)
    res = {}
    for mixed in (False, True):
        for stage in (1, 0):
            eng.set_option(N.OPT_STAGE_SLABS, stage)
            k = np.zeros(n, np.int32)
            st = np.zeros(n, np.uint8)
            if mixed:
                k[:] = np.random.default_rng(1).integers(0, 60, n)   # most blocks: no common index; some boats hit the end
                k[:256] = 17                                          # two blocks with a common index ...
                st[100:140] = N.ST_OUT_OF_BOX                         # ... part of whose boats are frozen
                k[100:140] = 3
            lat, lon = lat0.copy(), lon0.copy()
            eng.step_batch(3, k, lat, lon, head, nsteps=nsteps, seg_len=7, status=st, seed=21, stream=2)
            res[(mixed, stage)] = (k, lat, lon, st)
        for a, b in zip(res[(mixed, 1)], res[(mixed, 0)]):
            assert np.array_equal(a, b)
    eng.set_option(N.OPT_STAGE_SLABS, 0)
    k, lat, lon, st = res[(False, 1)]
    assert (k == nsteps).all() and (st == 0).all()
    o = H.oracle_sim(1, times=times)
    want = o.step_batch(0, lat0, lon0, head, nsteps=nsteps, seg_len=7, seed=21, stream=2, nthreads=8)
    assert relerr(lat, want["lat"]) < RTOL and relerr(lon, want["lon"]) < RTOL
    k, lat, lon, st = res[(True, 1)]
    assert (st[100:140] == N.ST_OUT_OF_BOX).all() and (k[100:140] == 3).all()
    assert (st == N.ST_PAST_HORIZON).sum() > 0 and (k[st == N.ST_PAST_HORIZON] == 100).all()
    eng.set_precision(N.PREC_F64)


@pytest.mark.parametrize("mode", MODES)
def test_large_grid_scattered_boats(eng, mode):
    """A near-global 1-degree grid (151 x 360 nodes, odd row length, slabs far larger than L1) with boats spread
    over all of it and both hemispheres: K1 against the oracle, same statuses and instants, positions to 1e-6."""
    t, la, lo, u, v = synth.wind_fields(3, lat=(-75.0, 75.0, 1.0), lon=(0.0, 359.0, 1.0))
    nsteps, n = 40, 20000
    times = np.linspace(0, 1.5, nsteps + 1)
    eng.upload_scenario(5, t, la, lo, u.astype(np.float32), v.astype(np.float32), times, [la[0], la[-1], lo[0], lo[-1]])
    eng.set_precision(mode)
    rng = np.random.default_rng(8)
    lat0, lon0 = rng.uniform(-74.0, 74.0, n), rng.uniform(0.5, 358.5, n)   # some start next to the edges and leave
    head = 360.0 * rng.random((2, n))
    k = np.zeros(n, np.int32)
    st = np.zeros(n, np.uint8)
    lat, lon = lat0.copy(), lon0.copy()
    eng.step_batch(5, k, lat, lon, head, nsteps=nsteps, seg_len=20, status=st, seed=17, stream=4)
    o = cport.OracleSim(t, la, lo, u, v, times, la, lo)
    want = o.step_batch(0, lat0, lon0, head, nsteps=nsteps, seg_len=20, seed=17, stream=4, nthreads=8)
    assert np.array_equal(st, want["status"]) and np.array_equal(k, want["k"])
    assert (st == N.ST_OUT_OF_GRID).sum() > 0 and (st == 0).sum() > n // 2
    tol = rtol_for(mode)
    # (coordinates pass through zero on this grid: relative to max(|x|, 1 degree))
    assert np.max(np.abs(lat - want["lat"]) / np.maximum(np.abs(want["lat"]), 1.0)) < tol
    assert np.max(np.abs(lon - want["lon"]) / np.maximum(np.abs(want["lon"]), 1.0)) < tol
    eng.free_scenario(5)
    eng.set_precision(N.PREC_F64)


def test_edge_cases(eng):
    eng.set_precision(N.PREC_F64)
    times = upload(eng, 0, 2)
    nT = times.size
    # empty batches are no-ops
    eng.step_batch(0, np.zeros(0, np.int32), np.zeros(0), np.zeros(0), np.zeros((1, 0)))
    r = eng.rollout_batch_host(0, 0, 1, H.STATE0, [46.0, 355.0], 2.0)
    assert r["reward"].size == 0
    # a boat leaving the weather grid freezes with OUT_OF_GRID (SciPy ValueError) at its last valid state
    k = np.zeros(2, np.int32)
    lat, lon = np.array([49.9, 44.0]), np.array([359.9, 355.0])
    st = np.zeros(2, np.uint8)
    eng.step_batch(0, k, lat, lon, np.full((1, 2), 45.0), nsteps=nT - 1, seg_len=nT, status=st)
    assert st[0] == N.ST_OUT_OF_GRID and 0 < k[0] < nT - 1 and (lat[0] > 50.0 or lon[0] > 360.0)
    assert st[1] in (N.ST_OK, N.ST_OUT_OF_GRID)
    # unknown scenario slot, mixed host/device arguments
    with pytest.raises(N.PmctsError):
        eng.step_batch(99, k, lat, lon, np.zeros((1, 2)))
    # rollout from a root on the horizon: the first prefix step is the IndexError case
    r = eng.rollout_batch_host(0, 1, 1, (nT - 1, 44.0, 355.0), [46.0, 355.0], 2.0, prefix=[[0.0]])
    assert r["status"][0] == N.ST_PAST_HORIZON


def test_device_pointer_path_and_backup(eng):
    """Device-resident arrays (torch tensors): same results as the host-buffer path; K3 back-up and
    histogram statistics against a NumPy restatement."""
    import torch
    eng.set_precision(N.PREC_F64)
    upload(eng, 0, 5)
    dev = torch.device("cuda:0")
    n_leaf, rep = 64, 32
    n = n_leaf * rep
    dest, tmin = [46.77751005, 355.0], 2.0223605
    prefix = 45.0 * (np.arange(n_leaf) % 8).reshape(n_leaf, 1).astype(np.float64)
    host = eng.rollout_batch_host(0, n_leaf, rep, H.STATE0, dest, tmin, prefix=prefix, seed=5, stream=1)
    eng.use_torch_stream()
    out = dict(reward=torch.empty(n, dtype=torch.float64, device=dev),
               status=torch.empty(n, dtype=torch.uint8, device=dev),
               leaf_bins=torch.zeros(n_leaf * 12, dtype=torch.int32, device=dev))
    eng.rollout_batch(0, n_leaf, rep, H.STATE0, dest, tmin, torch.tensor(prefix, device=dev).reshape(-1),
                      torch.ones(n_leaf, dtype=torch.int32, device=dev), 1, seed=5, stream=1, out=out)
    torch.cuda.synchronize()
    assert np.array_equal(out["reward"].cpu().numpy(), host["reward"])
    assert np.array_equal(out["status"].cpu().numpy(), host["status"])
    assert np.array_equal(out["leaf_bins"].cpu().numpy().astype(np.uint32).reshape(n_leaf, 12), host["leaf_bins"])
    # a small tree: node 0 root, nodes 1..8 children of the root, leaves = children of node 1..8
    n_nodes = 1 + 8 + n_leaf
    parent = np.full(n_nodes, -1, np.int32)
    arm = np.zeros(n_nodes, np.int8)
    parent[1:9] = 0
    arm[1:9] = np.arange(8)
    leaf_node = np.arange(9, 9 + n_leaf, dtype=np.int32)
    parent[9:] = 1 + (np.arange(n_leaf) % 8)
    arm[9:] = (np.arange(n_leaf) // 8) % 8
    slot = np.random.default_rng(0).integers(0, 8, n_leaf).astype(np.int8)
    table = torch.zeros(n_nodes * 96, dtype=torch.int32, device=dev)
    t = lambda a: torch.as_tensor(a, device=dev)  # noqa: E731
    eng.backup(table, n_nodes, t(parent), t(arm), t(leaf_node), t(slot), out["leaf_bins"])
    want = np.zeros((n_nodes, 8, 12), np.int64)
    for l in range(n_leaf):
        node = leaf_node[l]
        want[node, slot[l]] += host["leaf_bins"][l]
        while parent[node] >= 0:
            want[parent[node], arm[node]] += host["leaf_bins"][l]
            node = parent[node]
    torch.cuda.synchronize()
    got = table.cpu().numpy().reshape(n_nodes, 8, 12)
    assert np.array_equal(got, want)
    assert got[0].sum() == n
    # two scenarios at once, histograms at an address that is not 16-byte aligned (scalar-load variant of K3)
    big = torch.zeros(1 + 2 * n_leaf * 12, dtype=torch.int32, device=dev)
    big[1:1 + n_leaf * 12] = out["leaf_bins"]
    big[1 + n_leaf * 12:] = 2 * out["leaf_bins"]
    table2 = torch.zeros(2 * n_nodes * 96, dtype=torch.int32, device=dev)
    eng.backup(table2, n_nodes, t(parent), t(arm), t(np.tile(leaf_node, 2)), t(np.tile(slot, 2)), big[1:], n_scen=2)
    torch.cuda.synchronize()
    got2 = table2.cpu().numpy().reshape(2, n_nodes, 8, 12)
    assert np.array_equal(got2[0], want) and np.array_equal(got2[1], 2 * want)
    # the form histograms travel in between GPUs: one- and two-byte counts (pmcts_pack_bins), backed up as such, and
    # the library's own all-gather of them (a communicator of one rank: recv == send)
    import ctypes as C
    for dtype in (torch.uint8, torch.int16):
        packed = torch.zeros(n_leaf * 12, dtype=dtype, device=dev)
        eng.pack_bins(out["leaf_bins"], packed)
        assert np.array_equal(packed.cpu().numpy().astype(np.int64), out["leaf_bins"].cpu().numpy())
        table3 = torch.zeros(n_nodes * 96, dtype=torch.int32, device=dev)
        eng.backup(table3, n_nodes, t(parent), t(arm), t(leaf_node), t(slot), packed)
        torch.cuda.synchronize()
        assert np.array_equal(table3.cpu().numpy().reshape(n_nodes, 8, 12), want)
    lib = N.lib()
    uid = np.zeros(128, np.uint8)
    N.check(lib.pmcts_comm_unique_id(uid.ctypes.data))
    comm = C.c_void_p()
    N.check(lib.pmcts_comm_create(eng._h, 1, 0, uid.ctypes.data, C.byref(comm)))
    recv = torch.zeros_like(packed)
    N.check(lib.pmcts_allgather_updates(eng._h, comm, packed.data_ptr(), packed.numel() * 2, recv.data_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(recv, packed)
    lib.pmcts_comm_destroy(comm)
    mean = torch.empty(n_nodes * 8, dtype=torch.float64, device=dev)
    visits = torch.empty(n_nodes, dtype=torch.int32, device=dev)
    eng.hist_stats(table, n_nodes, mean, visits)
    torch.cuda.synchronize()
    centres = (np.arange(12) + 0.5) * 0.125
    tot = want.sum(-1)
    wmean = np.where(tot > 0, (want * centres).sum(-1) / np.maximum(tot, 1), 0.0)
    assert np.array_equal(mean.cpu().numpy().reshape(n_nodes, 8), wmean)
    assert np.array_equal(visits.cpu().numpy(), want.sum((1, 2)))
    eng.use_own_stream()


@pytest.mark.parametrize("mode", MODES)
def test_no_writes_outside_the_callers_arrays(eng, mode):
    """compute-sanitizer is closed on this pool, so bounds are checked the hard way: every device array is
    carved out of a larger buffer filled with a sentinel, sizes are awkward (not multiples of the warp /
    block size), and the guard zones must come back untouched."""
    import torch
    eng.set_precision(mode)
    times = upload(eng, 0, 8)
    nT = times.size
    dev = torch.device("cuda:0")
    eng.use_torch_stream()
    G = 64  # guard elements on each side

    def guarded(n, dtype, fill):
        buf = torch.full((n + 2 * G,), fill, dtype=dtype, device=dev)
        return buf, buf[G:G + n]

    def check(buf, n, fill, name):
        b = buf.cpu()
        if b.dtype.is_floating_point:
            assert bool((b[:G] == fill).all()) and bool((b[G + n:] == fill).all()), name
        else:
            assert bool((b[:G] == fill).all()) and bool((b[G + n:] == fill).all()), name

    try:
        for n_leaf, rep in ((1, 1), (3, 37), (7, 129), (33, 31)):
            n = n_leaf * rep
            stride = 3
            rng = np.random.default_rng(n)
            prefix = torch.tensor(45.0 * rng.integers(0, 8, (n_leaf, stride)), device=dev).reshape(-1)
            plen = torch.tensor(rng.integers(1, stride + 1, n_leaf).astype(np.int32), device=dev)
            bufs = {}
            out = {}
            for name, dt, size, fill in (("reward", torch.float64, n, -7.0), ("t_arr", torch.float64, n, -7.0),
                                         ("steps", torch.int32, n, -7), ("status", torch.uint8, n, 201),
                                         ("bin", torch.int32, n, -7), ("frac", torch.float64, n, -7.0),
                                         ("final_lat", torch.float64, n, -7.0), ("final_lon", torch.float64, n, -7.0),
                                         ("traj_lat", torch.float64, 5 * n, -7.0), ("traj_lon", torch.float64, 5 * n, -7.0),
                                         ("leaf_bins", torch.int32, n_leaf * 12, -7), ("stats", torch.float64, 8, -7.0)):
                bufs[name], out[name] = guarded(size, dt, fill)
                bufs[name + "_meta"] = (size, fill)
            eng.rollout_batch(0, n_leaf, rep, H.STATE0, [46.77751005, 355.0], 2.0223605, prefix, plen, stride,
                              seed=3, stream=n, out=out, traj_stride=5,
                              flags=N.ROLL_REPLAY_FULL_PREFIX)
            torch.cuda.synchronize()
            for name in out:
                size, fill = bufs[name + "_meta"]
                check(bufs[name], size, fill, "%s n_leaf=%d rep=%d" % (name, n_leaf, rep))
            assert int(out["leaf_bins"].sum()) == n and float(out["stats"][5]) == n
            assert bool((out["status"] >= 1).all()) and bool((out["status"] <= 3).all())
        for n in (1, 31, 129, 1000):
            lat0, lon0, h0 = synth.octagon_fleet(n, seed=n)
            head = torch.tensor(synth.octagon_headings(h0, 12, seg_len=5), device=dev).reshape(-1)
            kb, k = guarded(n, torch.int32, -7)
            k.zero_()
            lab, lat = guarded(n, torch.float64, -7.0)
            lob, lon = guarded(n, torch.float64, -7.0)
            lat.copy_(torch.tensor(lat0, device=dev))
            lon.copy_(torch.tensor(lon0, device=dev))
            stb, st = guarded(n, torch.uint8, 201)
            st.zero_()
            plb, pl = guarded(n, torch.float64, -7.0)
            tlb, tl = guarded(12 * n, torch.float64, -7.0)
            eng.step_batch(0, k, lat, lon, head, nsteps=12, seg_len=5, prev_lat=pl, status=st, seed=9, traj_lat=tl)
            torch.cuda.synchronize()
            for buf, size, fill, name in ((kb, n, -7, "k"), (lab, n, -7.0, "lat"), (lob, n, -7.0, "lon"),
                                          (stb, n, 201, "status"), (plb, n, -7.0, "prev_lat"), (tlb, 12 * n, -7.0, "traj")):
                check(buf, size, fill, "%s n=%d" % (name, n))
            assert bool((k == 12).all())
    finally:
        eng.use_own_stream()
        eng.set_precision(N.PREC_F64)
