"""TEST INFRASTRUCTURE ONLY -- ctypes front end of the CPU fp64 restatement (pmcts_oracle.c).

May be imported only by tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle_pmcts.so")
_OVERRIDE = os.environ.get("PMCTS_ORACLE_LIB")  # another build of the checker itself (tests/tools/sanitize_host.sh)

c_dp = C.POINTER(C.c_double)


class OrcWeather(C.Structure):
    _fields_ = [("nt", C.c_int), ("nlat", C.c_int), ("nlon", C.c_int),
                ("time", c_dp), ("lat", c_dp), ("lon", c_dp), ("u", c_dp), ("v", c_dp)]


class OrcSim(C.Structure):
    _fields_ = [("w", OrcWeather), ("nT", C.c_int), ("times", c_dp),
                ("lat_lo", C.c_double), ("lat_hi", C.c_double),
                ("lon_lo", C.c_double), ("lon_hi", C.c_double),
                ("polar", C.c_double * 36),
                ("wind_max", C.c_double), ("pofsail_min", C.c_double), ("pofsail_max", C.c_double),
                ("sigma_coeff", C.c_double), ("dest_angle", C.c_double), ("earth_radius", C.c_double)]


class OrcNoise(C.Structure):
    _fields_ = [("z", c_dp), ("seed", C.c_uint64), ("stream", C.c_uint64), ("id0", C.c_uint64)]


def build(force=False):
    """Compiles the restatement (gcc only).  Building the checker is not using it."""
    src = os.path.join(HERE, "pmcts_oracle.c")
    if (not force and os.path.exists(LIB_PATH)
            and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(src),
                                                  os.path.getmtime(os.path.join(HERE, "pmcts_oracle.h")))):
        return LIB_PATH
    subprocess.run(["make", "-C", HERE, "-B", "liboracle_pmcts.so"], check=True,
                   stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not _OVERRIDE and not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(_OVERRIDE or LIB_PATH)
        L.orc_deter_dyn.restype = C.c_double
        L.orc_deter_dyn.argtypes = [C.POINTER(OrcSim), C.c_double, C.c_double]
        L.orc_reward.restype = C.c_double
        L.orc_reward.argtypes = [C.c_double] * 3
        L.orc_hist_bin.argtypes = [C.c_double]
        L.orc_philox_normal.restype = C.c_double
        L.orc_philox_normal.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32]
        L.orc_interp.argtypes = [C.POINTER(OrcWeather), C.c_double, C.c_double, C.c_double, c_dp, c_dp]
        L.orc_polar_vel.argtypes = [C.c_double, C.c_double, c_dp, c_dp]
        L.orc_get_destination.argtypes = [C.POINTER(OrcSim)] + [C.c_double] * 4 + [c_dp, c_dp]
        L.orc_dist_bearing.argtypes = [C.POINTER(OrcSim)] + [C.c_double] * 4 + [c_dp, c_dp]
        L.orc_geo_to_cart.argtypes = [C.c_double, C.c_double, c_dp]
        L.orc_is_at_dest.argtypes = [C.POINTER(OrcSim), c_dp, c_dp, c_dp, c_dp]
        L.orc_is_terminal.argtypes = [C.POINTER(OrcSim), C.c_int, C.c_double, C.c_double]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(c_dp) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class OracleSim:
    """One weather scenario + simulator axes, mirroring ``Simulator(times, lats, lons, Weather, ...)``
    (code/model/simulatorTLKT.py:75-88)."""

    def __init__(self, wtime, wlat, wlon, u, v, times, lats, lons, sigma_coeff=0.2, polar=None):
        self._keep = [_f64(wtime), _f64(wlat), _f64(wlon), _f64(u), _f64(v), _f64(times)]
        wt, wl, wo, uu, vv, tt = self._keep
        assert uu.shape == (wt.size, wl.size, wo.size) and vv.shape == uu.shape
        self.s = OrcSim()
        lib().orc_default_params(C.byref(self.s))
        self.s.w = OrcWeather(wt.size, wl.size, wo.size, _dp(wt), _dp(wl), _dp(wo), _dp(uu), _dp(vv))
        self.s.nT = tt.size
        self.s.times = _dp(tt)
        self.s.lat_lo, self.s.lat_hi = float(lats[0]), float(lats[-1])
        self.s.lon_lo, self.s.lon_hi = float(lons[0]), float(lons[-1])
        self.s.sigma_coeff = sigma_coeff
        if polar is not None:
            self.s.polar = (C.c_double * 36)(*np.asarray(polar, dtype=np.float64).ravel())
        self.times = tt

    # ---- scalar building blocks ------------------------------------------------------------
    def interp(self, t, lat, lon):
        u, v = C.c_double(), C.c_double()
        rc = lib().orc_interp(C.byref(self.s.w), t, lat, lon, C.byref(u), C.byref(v))
        if rc:
            raise ValueError("out of bounds")
        return u.value, v.value

    def deter_dyn(self, p, w):
        return lib().orc_deter_dyn(C.byref(self.s), p, w)

    def get_destination(self, dist, bearing, lat, lon):
        a, b = C.c_double(), C.c_double()
        lib().orc_get_destination(C.byref(self.s), dist, bearing, lat, lon, C.byref(a), C.byref(b))
        return [a.value, b.value]

    def dist_bearing(self, lat, lon, dlat, dlon):
        a, b = C.c_double(), C.c_double()
        lib().orc_dist_bearing(C.byref(self.s), lat, lon, dlat, dlon, C.byref(a), C.byref(b))
        return a.value, b.value

    def is_at_dest(self, dest, a, b):
        fr = C.c_double()
        d = (C.c_double * 2)(*dest)
        aa = (C.c_double * 2)(*a)
        bb = (C.c_double * 2)(*b)
        rc = lib().orc_is_at_dest(C.byref(self.s), d, aa, bb, C.byref(fr))
        if rc < 0:
            raise ZeroDivisionError("float division by zero")
        return [True, fr.value] if rc else [False, None]

    def is_terminal(self, k, lat, lon):
        return bool(lib().orc_is_terminal(C.byref(self.s), k, lat, lon))

    # ---- batched drivers ---------------------------------------------------------------------
    @staticmethod
    def _noise(z, seed, stream, id0):
        nz = OrcNoise()
        if z is not None:
            z = _f64(z)
            nz.z = _dp(z)
        nz.seed, nz.stream, nz.id0 = seed, stream, id0
        return nz, z

    def step_batch(self, k, lat, lon, heading, nsteps=1, seg_len=1, z=None, seed=0, stream=0, id0=0,
                   status=None, traj=False, flags=0, nthreads=1):
        """Returns dict(k, lat, lon, prev_lat, prev_lon, status[, traj_lat, traj_lon])."""
        lat = _f64(lat).copy()
        lon = _f64(lon).copy()
        n = lat.size
        k = np.ascontiguousarray(np.broadcast_to(np.asarray(k, dtype=np.int32), (n,))).copy()
        heading = _f64(heading).reshape(-1, n)
        assert heading.shape[0] * seg_len >= nsteps
        prev_lat, prev_lon = lat.copy(), lon.copy()
        status = np.zeros(n, np.uint8) if status is None else np.ascontiguousarray(status, np.uint8).copy()
        tl = np.full((nsteps, n), np.nan) if traj else None
        to = np.full((nsteps, n), np.nan) if traj else None
        nz, zkeep = self._noise(z, seed, stream, id0)
        lib().orc_step_batch(C.byref(self.s), C.c_int64(n), C.c_int(nsteps),
                             k.ctypes.data_as(C.POINTER(C.c_int32)), _dp(lat), _dp(lon),
                             _dp(prev_lat), _dp(prev_lon), _dp(heading), C.c_int(seg_len),
                             C.byref(nz), status.ctypes.data_as(C.POINTER(C.c_uint8)),
                             _dp(tl), _dp(to), C.c_int(flags), C.c_int(nthreads))
        out = dict(k=k, lat=lat, lon=lon, prev_lat=prev_lat, prev_lon=prev_lon, status=status)
        if traj:
            out.update(traj_lat=tl, traj_lon=to)
        return out

    def rollout_batch(self, n, root, dest, tmin, prefix=None, prefix_len=None, replay_full=False,
                      mode=0, z=None, seed=0, stream=0, id0=0, traj_steps=0, nthreads=1):
        """Returns dict(reward, t_arr, steps, status, bin, frac, final_lat, final_lon[, traj_*])."""
        root = _f64(root)
        dest = _f64(dest)
        if prefix is not None:
            prefix = _f64(prefix).reshape(n, -1)
            stride = prefix.shape[1]
            if prefix_len is None:
                prefix_len = np.full(n, stride, np.int32)
            prefix_len = np.ascontiguousarray(prefix_len, np.int32)
        else:
            stride = 0
            prefix_len = None
        reward = np.empty(n)
        t_arr = np.empty(n)
        frac = np.empty(n)
        flat = np.empty(n)
        flon = np.empty(n)
        steps = np.empty(n, np.int32)
        bins = np.empty(n, np.int32)
        status = np.empty(n, np.uint8)
        tl = np.full((traj_steps, n), np.nan) if traj_steps else None
        to = np.full((traj_steps, n), np.nan) if traj_steps else None
        nz, zkeep = self._noise(z, seed, stream, id0)
        i32p = C.POINTER(C.c_int32)
        lib().orc_rollout_batch(C.byref(self.s), C.c_int64(n), _dp(root), _dp(dest), C.c_double(tmin),
                                _dp(prefix), prefix_len.ctypes.data_as(i32p) if prefix_len is not None else None,
                                C.c_int(stride), C.c_int(int(replay_full)), C.c_int(mode), C.byref(nz),
                                _dp(reward), _dp(t_arr), steps.ctypes.data_as(i32p),
                                status.ctypes.data_as(C.POINTER(C.c_uint8)), bins.ctypes.data_as(i32p),
                                _dp(frac), _dp(flat), _dp(flon), _dp(tl), _dp(to), C.c_int(traj_steps),
                                C.c_int(nthreads))
        out = dict(reward=reward, t_arr=t_arr, steps=steps, status=status, bin=bins, frac=frac,
                   final_lat=flat, final_lon=flon)
        if traj_steps:
            out.update(traj_lat=tl, traj_lon=to)
        return out


def polar_vel(u, v):
    a, b = C.c_double(), C.c_double()
    lib().orc_polar_vel(u, v, C.byref(a), C.byref(b))
    return a.value, b.value


def geo_to_cart(lat, lon):
    out = (C.c_double * 3)()
    lib().orc_geo_to_cart(lat, lon, out)
    return list(out)


def hist_bin(r):
    return lib().orc_hist_bin(r)


def reward(tmax, tmin, t):
    return lib().orc_reward(tmax, tmin, t)


def philox4x32(ctr, key):
    out = (C.c_uint32 * 4)()
    lib().orc_philox4x32((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
    return list(out)


def philox_normal(seed, stream, boat, step):
    return lib().orc_philox_normal(seed, stream, boat, step)
