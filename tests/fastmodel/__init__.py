"""TEST INFRASTRUCTURE ONLY -- ctypes front end of the host build of csrc/pmcts_fast.cuh.

Lets the CPU suite measure the arithmetic of the mixed-precision kernels against the oracle without
a GPU.  The product package never imports this module (the kernels themselves run on the GPU only).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "libfastmodel.so")
SRC = [os.path.join(HERE, "fastmodel.cpp")] + [os.path.join(ROOT, "iboat_pmcts_b200", "csrc", f)
                                               for f in ("pmcts_fast.cuh", "pmcts_types.h", "pmcts_host.h")]

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in SRC):
            subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-w", "-o", LIB,
                            SRC[0]], check=True)
        L = C.CDLL(LIB)
        L.fm_create.restype = C.c_void_p
        L.fm_boat_speed.restype = C.c_float
        L.fm_boat_speed.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_double]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, np.float64)


class FastModel:
    def __init__(self, wtime, wlat, wlon, u, v, times, lats, lons, sigma_coeff=0.2):
        self._keep = [_f64(x) for x in (wtime, wlat, wlon, u, v, times)]
        wt, wl, wo, uu, vv, tt = self._keep
        box = _f64([lats[0], lats[-1], lons[0], lons[-1]])
        self.h = C.c_void_p(lib().fm_create(_p(wt), wt.size, _p(wl), wl.size, _p(wo), wo.size, _p(uu), _p(vv),
                                            _p(tt), tt.size, _p(box), C.c_double(sigma_coeff)))
        self.times = tt

    def __del__(self):
        if getattr(self, "h", None):
            lib().fm_destroy(self.h)
            self.h = None

    def force_generic(self, on=True):
        """Run the generic variant of rollout_one even where rollout_impl would pick a specialised one."""
        lib().fm_force_generic(self.h, 1 if on else 0)

    def last_variant(self):
        """0 generic, 1 default boat (constants as immediates), 2 turbo -- of the last rollout_batch."""
        return int(lib().fm_last_variant(self.h))

    def set_margins(self, angle=4e-3, along=2e-4, near_miss2=1e-5, box=3e-5, bin=2e-5):
        lib().fm_set_margins(self.h, C.c_float(angle), C.c_float(along), C.c_float(near_miss2), C.c_float(box),
                             C.c_double(bin))

    def boat_speed(self, u, v, heading):
        return float(lib().fm_boat_speed(self.h, u, v, heading))

    def rollout_batch(self, n_leaf, replicas, root, dest, tmin, prefix=None, prefix_len=None, z=None, seed=None,
                      stream=0, id0=0, flags=0, traj_steps=0):
        n = n_leaf * replicas
        root, dest = _f64(root), _f64(dest)
        stride = 0
        if prefix is not None:
            prefix = _f64(prefix).reshape(n_leaf, -1)
            stride = prefix.shape[1]
            if prefix_len is None:
                prefix_len = np.full(n_leaf, stride, np.int32)
            prefix_len = np.ascontiguousarray(prefix_len, np.int32)
        mode = 0 if z is not None else (1 if seed is not None else 2)
        if z is not None:
            z = _f64(z)
        out = dict(reward=np.empty(n), t_arr=np.empty(n), steps=np.empty(n, np.int32), status=np.empty(n, np.uint8),
                   bin=np.empty(n, np.int32), frac=np.empty(n), final_lat=np.empty(n), final_lon=np.empty(n),
                   uncertain=np.empty(n, np.uint8))
        tl = np.full((traj_steps, n), np.nan) if traj_steps else None
        to = np.full((traj_steps, n), np.nan) if traj_steps else None
        lib().fm_rollout_batch(self.h, C.c_int64(n_leaf), C.c_int(replicas), _p(root), _p(dest), C.c_double(tmin),
                               _p(prefix), _p(prefix_len), C.c_int(stride), _p(z), C.c_uint64(seed or 0),
                               C.c_uint64(stream), C.c_uint64(id0), C.c_int(mode), C.c_int(flags),
                               _p(out["reward"]), _p(out["t_arr"]), _p(out["steps"]), _p(out["status"]),
                               _p(out["bin"]), _p(out["frac"]), _p(out["final_lat"]), _p(out["final_lon"]),
                               _p(tl), _p(to), C.c_int(traj_steps), _p(out["uncertain"]))
        if traj_steps:
            out.update(traj_lat=tl, traj_lon=to)
        return out

    def step_batch(self, k, lat, lon, heading, nsteps=1, seg_len=1, z=None, seed=None, stream=0, id0=0, traj=False):
        lat, lon = _f64(lat).copy(), _f64(lon).copy()
        n = lat.size
        k = np.ascontiguousarray(np.broadcast_to(np.asarray(k, np.int32), (n,))).copy()
        heading = _f64(heading).reshape(-1, n)
        status = np.zeros(n, np.uint8)
        needs = np.zeros(n, np.uint8)
        mode = 0 if z is not None else (1 if seed is not None else 2)
        if z is not None:
            z = _f64(z)
        tl = np.full((nsteps, n), np.nan) if traj else None
        to = np.full((nsteps, n), np.nan) if traj else None
        lib().fm_step_batch(self.h, C.c_int64(n), C.c_int(nsteps), _p(k), _p(lat), _p(lon), _p(heading),
                            C.c_int(seg_len), _p(z), C.c_uint64(seed or 0), C.c_uint64(stream), C.c_uint64(id0),
                            C.c_int(mode), _p(status), _p(tl), _p(to), _p(needs))
        out = dict(k=k, lat=lat, lon=lon, status=status, needs_literal=needs)
        if traj:
            out.update(traj_lat=tl, traj_lon=to)
        return out
