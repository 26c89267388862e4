"""Engine: one GPU context of the rollout engine (thin object layer over include/pmcts_b200.h).

Array arguments are either NumPy arrays (host memory: the library copies them to the device and
the results back inside the call -- ``PMCTS_MEM_HOST``) or CUDA ``torch.Tensor``s (device memory:
the call is asynchronous on the ctx stream).  A call must not mix the two kinds.
"""
import ctypes as C

import numpy as np

from . import _native as N

try:  # torch is plumbing only (device buffers / streams); the host path needs NumPy alone
    import torch
except Exception:  # pragma: no cover
    torch = None

_NP = {"f64": np.float64, "i32": np.int32, "u8": np.uint8, "u32": np.uint32, "i8": np.int8, "i16": np.int16}


def _is_tensor(x):
    return torch is not None and isinstance(x, torch.Tensor)


class _Args:
    """Collects array arguments, checks host/device consistency, keeps them alive for the call."""

    def __init__(self):
        self.kind = None
        self.keep = []

    def _set_kind(self, kind):
        if self.kind is None:
            self.kind = kind
        elif self.kind != kind:
            raise TypeError("pmcts_b200: a call takes either NumPy (host) or CUDA tensor (device) arrays, not both")

    def ptr(self, x, dtype, size=None, writable=False):
        if x is None:
            return None
        if _is_tensor(x):
            if not x.is_cuda:
                raise TypeError("pmcts_b200: torch tensors must live on a CUDA device")
            want = {"f64": torch.float64, "i32": torch.int32, "u8": torch.uint8, "u32": torch.int32,
                    "i8": torch.int8, "i16": torch.int16}[dtype]
            if x.dtype != want or not x.is_contiguous():
                raise TypeError("pmcts_b200: expected contiguous %s tensor, got %s" % (want, x.dtype))
            if size is not None and x.numel() < size:
                raise ValueError("pmcts_b200: tensor too small (%d < %d)" % (x.numel(), size))
            self._set_kind("device")
            self.keep.append(x)
            return x.data_ptr()
        if not isinstance(x, np.ndarray) or x.dtype != _NP[dtype] or not x.flags.c_contiguous:
            if writable:
                raise TypeError("pmcts_b200: output arrays must be C-contiguous %s ndarrays" % dtype)
            x = np.ascontiguousarray(x, dtype=_NP[dtype])
        if size is not None and x.size < size:
            raise ValueError("pmcts_b200: array too small (%d < %d)" % (x.size, size))
        self._set_kind("host")
        self.keep.append(x)
        return x.ctypes.data

    @property
    def mem_flag(self):
        return N.MEM_HOST if self.kind != "device" else 0


def make_noise(args, z=None, seed=None, stream=0, id0=0, size=None):
    nz = N.Noise()
    if z is not None:
        nz.mode = N.NOISE_INJECTED
        nz.z = args.ptr(z, "f64", size)
    elif seed is not None:
        nz.mode = N.NOISE_PHILOX
        nz.seed, nz.stream, nz.id0 = int(seed), int(stream), int(id0)
    else:
        nz.mode = N.NOISE_NONE
    return nz


class Engine:
    """One ``pmcts_ctx``: owns the uploaded scenarios of one GPU."""

    def __init__(self, device=0, precision=N.PREC_F64):
        self._lib = N.lib()
        h = C.c_void_p()
        N.check(self._lib.pmcts_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)
        self._nT = {}  # simulator instants per scenario slot (size of an injected-noise array)
        self.set_precision(precision)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pmcts_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- context ----------------------------------------------------------------------------
    def set_precision(self, mode):
        N.check(self._lib.pmcts_set_precision(self._h, int(mode)))
        self.precision = int(mode)

    def set_fast_margins(self, angle=-1.0, along=-1.0, near_miss2=-1.0, box=-1.0, bin=-1.0):
        """Decision margins of the mixed-precision kernels (negative = keep; all negative = defaults)."""
        N.check(self._lib.pmcts_set_fast_margins(self._h, float(angle), float(along), float(near_miss2),
                                                 float(box), float(bin)))

    def last_redo_count(self):
        """Rollouts the last FAST launch handed to the fp64 kernel (synchronises)."""
        return int(self._lib.pmcts_last_redo_count(self._h))

    def use_torch_stream(self):
        """Orders the engine's work on torch's current stream of this device."""
        s = torch.cuda.current_stream(self.device).cuda_stream
        N.check(self._lib.pmcts_set_stream(self._h, C.c_void_p(s)))

    def use_own_stream(self):
        N.check(self._lib.pmcts_reset_stream(self._h))

    def synchronize(self):
        N.check(self._lib.pmcts_synchronize(self._h))

    def probe_fp32_peak(self):
        """Measured FP32 FMA throughput of this device, TFLOP/s."""
        v = C.c_double(0.0)
        N.check(self._lib.pmcts_probe_fp32_peak(self._h, C.byref(v)))
        return v.value

    def last_kernel_variant(self):
        """N.VARIANT_* of the last step / rollout launch."""
        return int(self._lib.pmcts_last_kernel_variant(self._h))

    def set_option(self, option, value):
        N.check(self._lib.pmcts_set_option(self._h, int(option), int(value)))

    def join(self, keep_last=0):
        """Orders the ctx stream after the deferred finishers of all but the ``keep_last`` latest launches."""
        N.check(self._lib.pmcts_join(self._h, int(keep_last)))

    def wait(self, keep_last=0):
        """Blocks until the deferred launches of all but the ``keep_last`` latest calls are complete (host arrays
        filled)."""
        N.check(self._lib.pmcts_wait(self._h, int(keep_last)))

    def launch_count(self):
        return int(self._lib.pmcts_launch_count(self._h))

    def enable_timing(self, on=True):
        N.check(self._lib.pmcts_enable_timing(self._h, int(bool(on))))

    def kernel_ms(self, reset=True, kind=None):
        """Device milliseconds of the launches timed since the last reset (all kernels or one kind)."""
        if kind is None:
            return float(self._lib.pmcts_kernel_ms_total(self._h, int(bool(reset))))
        return float(self._lib.pmcts_kernel_ms(self._h, int(kind), int(bool(reset))))

    def set_params(self, polar, wind_max, pofsail_min, pofsail_max, sigma_coeff, dest_angle, earth_radius):
        pol = np.ascontiguousarray(polar, dtype=np.float64).reshape(36)
        N.check(self._lib.pmcts_set_params(self._h, pol.ctypes.data, float(wind_max), float(pofsail_min),
                                           float(pofsail_max), float(sigma_coeff), float(dest_angle),
                                           float(earth_radius)))

    # ---- scenarios ----------------------------------------------------------------------------
    def upload_scenario(self, scen, wtime, wlat, wlon, u, v, sim_times=None, box=None):
        wtime = np.ascontiguousarray(wtime, np.float64)
        wlat = np.ascontiguousarray(wlat, np.float64)
        wlon = np.ascontiguousarray(wlon, np.float64)
        u = np.asarray(u)
        v = np.asarray(v)
        dt = np.float32 if (u.dtype == np.float32 and v.dtype == np.float32) else np.float64
        u = np.ascontiguousarray(u, dt)
        v = np.ascontiguousarray(v, dt)
        if u.shape != (wtime.size, wlat.size, wlon.size) or v.shape != u.shape:
            raise ValueError("u, v must have shape (ntime, nlat, nlon)")
        if sim_times is None:  # weather only: serves interp_wind
            N.check(self._lib.pmcts_scenario_upload(
                self._h, int(scen), wtime.ctypes.data, wtime.size, wlat.ctypes.data, wlat.size,
                wlon.ctypes.data, wlon.size, u.ctypes.data, v.ctypes.data, 1 if dt == np.float32 else 0,
                None, 0, None))
            self._nT[int(scen)] = 0
            return
        sim_times = np.ascontiguousarray(sim_times, np.float64)
        box = np.ascontiguousarray(box, np.float64).reshape(4)
        N.check(self._lib.pmcts_scenario_upload(
            self._h, int(scen), wtime.ctypes.data, wtime.size, wlat.ctypes.data, wlat.size,
            wlon.ctypes.data, wlon.size, u.ctypes.data, v.ctypes.data, 1 if dt == np.float32 else 0,
            sim_times.ctypes.data, sim_times.size, box.ctypes.data))
        self._nT[int(scen)] = int(sim_times.size)

    def free_scenario(self, scen):
        N.check(self._lib.pmcts_scenario_free(self._h, int(scen)))
        self._nT.pop(int(scen), None)

    def _z_size(self, scen, n_scen, n):
        """Elements the library reads from an injected-noise array: z[n_scen][nT][n] (one row per simulator
        instant, not per step made)."""
        return n_scen * self._nT.get(int(scen), 0) * n

    def get_wind(self, scen, k, lat, lon):
        """Simulator.getWind for arrays of query points (host arrays in, host arrays out)."""
        lat = np.ascontiguousarray(lat, np.float64).ravel()
        lon = np.ascontiguousarray(lon, np.float64).ravel()
        n = lat.size
        k = np.ascontiguousarray(np.broadcast_to(np.asarray(k, np.int32), (n,)))
        u, v, st = np.empty(n), np.empty(n), np.empty(n, np.uint8)
        N.check(self._lib.pmcts_get_wind(self._h, int(scen), n, k.ctypes.data, lat.ctypes.data, lon.ctypes.data,
                                         u.ctypes.data, v.ctypes.data, st.ctypes.data, N.MEM_HOST))
        return u, v, st

    def interp_wind(self, scen, t, lat, lon):
        """SciPy-exact trilinear interpolation of the raw grid at arbitrary (t, lat, lon) points."""
        lat = np.ascontiguousarray(lat, np.float64).ravel()
        lon = np.ascontiguousarray(lon, np.float64).ravel()
        t = np.ascontiguousarray(np.broadcast_to(np.asarray(t, np.float64), lat.shape))
        n = lat.size
        u, v, st = np.empty(n), np.empty(n), np.empty(n, np.uint8)
        N.check(self._lib.pmcts_interp_wind(self._h, int(scen), n, t.ctypes.data, lat.ctypes.data, lon.ctypes.data,
                                            u.ctypes.data, v.ctypes.data, st.ctypes.data, N.MEM_HOST))
        return u, v, st

    # ---- K1 ---------------------------------------------------------------------------------------
    def step_batch(self, scen, k, lat, lon, heading, nsteps=1, seg_len=1, prev_lat=None, prev_lon=None,
                   status=None, z=None, seed=None, stream=0, id0=0, traj_lat=None, traj_lon=None, flags=0):
        """In-place batched Simulator.doStep.  k/lat/lon/(prev_*)/(status) are updated in place.  ``heading`` is
        [nseg][n] degrees (float64), or -- dtype uint8 -- indices into the reference's ACTIONS (45 degrees each)."""
        a = _Args()
        n = lat.numel() if _is_tensor(lat) else lat.size
        nseg = (nsteps + seg_len - 1) // seg_len
        pk = a.ptr(k, "i32", n, True)
        plat, plon = a.ptr(lat, "f64", n, True), a.ptr(lon, "f64", n, True)
        ppl, ppo = a.ptr(prev_lat, "f64", n, True), a.ptr(prev_lon, "f64", n, True)
        if getattr(heading, "dtype", None) == (torch.uint8 if _is_tensor(heading) else np.uint8):
            ph = a.ptr(heading, "u8", nseg * n)
            flags = int(flags) | N.STEP_HEADING_ACTIONS
        else:
            ph = a.ptr(heading, "f64", nseg * n)
        pst = a.ptr(status, "u8", n, True)
        ptl, pto = a.ptr(traj_lat, "f64", nsteps * n, True), a.ptr(traj_lon, "f64", nsteps * n, True)
        nz = make_noise(a, z, seed, stream, id0, nsteps * n)
        N.check(self._lib.pmcts_step_batch(self._h, int(scen), n, int(nsteps), pk, plat, plon, ppl, ppo, ph,
                                           int(seg_len), C.byref(nz), pst, ptl, pto, int(flags) | a.mem_flag))

    # ---- K2 / K4 / K5 -----------------------------------------------------------------------------
    def rollout_batch(self, scen, n_leaf, replicas, root, dest, tmin, prefix=None, prefix_len=None,
                      prefix_stride=0, z=None, seed=None, stream=0, id0=0, out=None, traj_stride=0, flags=0):
        """Batched default-policy rollouts.  ``out`` maps output names (reward, t_arr, steps, status,
        bin, frac, final_lat, final_lon, traj_lat, traj_lon, leaf_bins, stats) to preallocated arrays.
        Injected noise ``z`` has shape [len(times)][n] (a row per simulator instant; ValueError if shorter)."""
        out = out or {}
        a = _Args()
        n = int(n_leaf) * int(replicas)
        root = np.ascontiguousarray(root, np.float64).reshape(3)
        dest = np.ascontiguousarray(dest, np.float64).reshape(2)
        ppre = a.ptr(prefix, "f64", n_leaf * prefix_stride) if prefix is not None else None
        plen = a.ptr(prefix_len, "i32", n_leaf) if prefix_len is not None else None
        g = lambda name, dt, size: a.ptr(out.get(name), dt, size, True)  # noqa: E731
        ptrs = [g("reward", "f64", n), g("t_arr", "f64", n), g("steps", "i32", n), g("status", "u8", n),
                g("bin", "i32", n), g("frac", "f64", n), g("final_lat", "f64", n), g("final_lon", "f64", n),
                g("traj_lat", "f64", traj_stride * n), g("traj_lon", "f64", traj_stride * n)]
        plb = g("leaf_bins", "u32", n_leaf * 12)
        pstats = g("stats", "f64", 8)
        nz = make_noise(a, z, seed, stream, id0, self._z_size(scen, 1, n))
        N.check(self._lib.pmcts_rollout_batch(self._h, int(scen), int(n_leaf), int(replicas), root.ctypes.data,
                                              dest.ctypes.data, float(tmin), ppre, plen, int(prefix_stride),
                                              C.byref(nz), *ptrs, int(traj_stride), plb, pstats,
                                              int(flags) | a.mem_flag))

    def rollout_forest(self, scens, n_leaf, replicas, root, dest, tmin, prefix=None, prefix_len=None,
                       prefix_stride=0, prefix_shared=True, z=None, seed=None, stream=0, id0=0, out=None, flags=0):
        """K2 for the scenario slots ``scens`` in one launch (``pmcts_rollout_forest``).  ``out`` maps
        reward / t_arr / steps / status / bin ([S][n]), leaf_bins ([S][n_leaf][12]) and stats ([S][8]) to
        preallocated arrays."""
        out = out or {}
        a = _Args()
        scens = np.ascontiguousarray(scens, np.int32)
        S = scens.size
        n = int(n_leaf) * int(replicas)
        P = 1 if prefix_shared else S
        root = np.ascontiguousarray(root, np.float64).reshape(3)
        dest = np.ascontiguousarray(dest, np.float64).reshape(2)
        ppre = a.ptr(prefix, "f64", P * n_leaf * prefix_stride) if prefix is not None else None
        plen = a.ptr(prefix_len, "i32", P * n_leaf) if prefix_len is not None else None
        g = lambda name, dt, size: a.ptr(out.get(name), dt, size, True)  # noqa: E731
        ptrs = [g("reward", "f64", S * n), g("t_arr", "f64", S * n), g("steps", "i32", S * n),
                g("status", "u8", S * n), g("bin", "i32", S * n), g("leaf_bins", "u32", S * n_leaf * 12),
                g("stats", "f64", S * 8)]
        nz = make_noise(a, z, seed, stream, id0, self._z_size(scens[0], S, n) if S else 0)
        N.check(self._lib.pmcts_rollout_forest(self._h, S, scens.ctypes.data, int(n_leaf), int(replicas),
                                               root.ctypes.data, dest.ctypes.data, float(tmin), ppre, plen,
                                               int(prefix_stride), int(bool(prefix_shared)), C.byref(nz), *ptrs,
                                               int(flags) | a.mem_flag))

    def rollout_batch_host(self, scen, n_leaf, replicas, root, dest, tmin, prefix=None, prefix_len=None,
                           z=None, seed=None, stream=0, id0=0, traj_steps=0, flags=0, want=None):
        """Convenience wrapper: allocates NumPy outputs and returns them as a dict."""
        n = int(n_leaf) * int(replicas)
        names = want or ("reward", "t_arr", "steps", "status", "bin", "frac", "final_lat", "final_lon",
                         "leaf_bins", "stats")
        spec = dict(reward=(np.float64, n), t_arr=(np.float64, n), steps=(np.int32, n), status=(np.uint8, n),
                    bin=(np.int32, n), frac=(np.float64, n), final_lat=(np.float64, n),
                    final_lon=(np.float64, n), leaf_bins=(np.uint32, n_leaf * 12), stats=(np.float64, 8))
        out = {nm: np.empty(spec[nm][1], spec[nm][0]) for nm in names}
        if traj_steps:
            out["traj_lat"] = np.full((traj_steps, n), np.nan)
            out["traj_lon"] = np.full((traj_steps, n), np.nan)
        stride = 0
        if prefix is not None:
            prefix = np.ascontiguousarray(prefix, np.float64).reshape(n_leaf, -1)
            stride = prefix.shape[1]
            if prefix_len is None:
                prefix_len = np.full(n_leaf, stride, np.int32)
            prefix_len = np.ascontiguousarray(prefix_len, np.int32)
        if z is not None:
            z = np.ascontiguousarray(z, np.float64)
        self.rollout_batch(scen, n_leaf, replicas, root, dest, tmin, prefix, prefix_len, stride, z, seed,
                           stream, id0, out, traj_steps, flags)
        if "leaf_bins" in out:
            out["leaf_bins"] = out["leaf_bins"].reshape(n_leaf, 12)
        return out

    # ---- K3 / statistics --------------------------------------------------------------------------
    def backup(self, table, n_nodes, parent, arm, leaf_node, leaf_slot, leaf_bins, n_scen=1):
        """table[n_scen][n_nodes][8][12] += per-leaf bins along every ancestor chain (device arrays).  ``leaf_bins``
        may be the uint32 counters K2 produced or their packed form (uint8 / int16 tensors from ``pack_bins``)."""
        a = _Args()
        n_leaf = leaf_node.numel() // n_scen
        flag, kind = 0, "u32"
        if _is_tensor(leaf_bins) and leaf_bins.dtype == torch.uint8:
            flag, kind = N.BACKUP_BINS_U8, "u8"
        elif _is_tensor(leaf_bins) and leaf_bins.dtype == torch.int16:
            flag, kind = N.BACKUP_BINS_U16, "i16"
        N.check(self._lib.pmcts_backup(self._h, a.ptr(table, "u32", n_scen * n_nodes * 96), int(n_nodes),
                                       int(n_scen), a.ptr(parent, "i32", n_nodes), a.ptr(arm, "i8", n_nodes),
                                       n_leaf, a.ptr(leaf_node, "i32", n_scen * n_leaf),
                                       a.ptr(leaf_slot, "i8", n_scen * n_leaf),
                                       a.ptr(leaf_bins, kind, n_scen * n_leaf * 12), a.mem_flag | flag))

    def pack_bins(self, bins, out):
        """Narrows uint32 histogram counters (int32 tensor) into ``out`` (uint8 or int16 tensor of the same length)."""
        width = 1 if out.dtype == torch.uint8 else 2
        N.check(self._lib.pmcts_pack_bins(self._h, bins.data_ptr(), int(bins.numel()), out.data_ptr(), width))

    def hist_stats(self, table, n_nodes, mean=None, visits=None):
        a = _Args()
        N.check(self._lib.pmcts_hist_stats(self._h, a.ptr(table, "u32", n_nodes * 96), int(n_nodes),
                                           a.ptr(mean, "f64", n_nodes * 8, True),
                                           a.ptr(visits, "u32", n_nodes, True), a.mem_flag))


    def master_policy(self, parent, arm, row_index, rows, probability, uct_coeff, max_depth,
                      want=("policy", "policy_nodes")):
        """``pmcts_master_policy`` on host arrays: returns a dict with the outputs named in ``want`` (value,
        explore, guess, best_child, policy, policy_nodes)."""
        parent = np.ascontiguousarray(parent, np.int32)
        arm = np.ascontiguousarray(arm, np.int8)
        row_index = np.ascontiguousarray(row_index, np.int32)
        rows = np.ascontiguousarray(rows, np.uint32)
        prob = np.ascontiguousarray(probability, np.float64)
        n, S = int(parent.size), int(prob.size)
        n_rows = int(rows.size // 96)
        if row_index.size != n * S:
            raise ValueError("row_index must have shape (nodes, scenarios)")
        spec = dict(value=((n, S + 1), np.float64), explore=((n, S + 1), np.float64), guess=((n, S), np.float64),
                    best_child=((n, S + 1), np.int32), policy=((S + 1, max_depth), np.int8),
                    policy_nodes=((S + 1, max_depth + 1), np.int32))
        out = {k: np.empty(*spec[k]) for k in want}
        ptr = lambda k: out[k].ctypes.data if k in out else None  # noqa: E731
        N.check(self._lib.pmcts_master_policy(self._h, n, S, parent.ctypes.data, arm.ctypes.data, row_index.ctypes.data,
                                              n_rows, rows.ctypes.data, prob.ctypes.data, float(uct_coeff),
                                              int(max_depth), ptr("value"), ptr("explore"), ptr("guess"),
                                              ptr("best_child"), ptr("policy"), ptr("policy_nodes"), N.MEM_HOST))
        return out


    def forest_master_policy(self, native, probability, uct_coeff, max_depth, want=("policy", "policy_nodes")):
        """``pmcts_forest_master_policy``: the reductions of ``master_policy`` on the master of a native forest."""
        prob = np.ascontiguousarray(probability, np.float64)
        n, S = native.master_size(), int(prob.size)
        spec = dict(value=((n, S + 1), np.float64), explore=((n, S + 1), np.float64), guess=((n, S), np.float64),
                    best_child=((n, S + 1), np.int32), policy=((S + 1, max_depth), np.int8),
                    policy_nodes=((S + 1, max_depth + 1), np.int32))
        out = {k: np.empty(*spec[k]) for k in want}
        ptr = lambda k: out[k].ctypes.data if k in out else None  # noqa: E731
        N.check(self._lib.pmcts_forest_master_policy(self._h, native.handle, prob.ctypes.data, float(uct_coeff),
                                                     int(max_depth), ptr("value"), ptr("explore"), ptr("guess"),
                                                     ptr("best_child"), ptr("policy"), ptr("policy_nodes")))
        return out


_default = {}
_default_device = None


def set_default_device(device):
    """Device used by the reference-API mirror (Weather / Simulator / Tree / Forest ...); by default
    LOCAL_RANK when launched by torchrun, else 0."""
    global _default_device
    _default_device = int(device)


def default_engine(device=None):
    """Process-wide engine per device, created on first use."""
    import os
    if device is None:
        device = _default_device if _default_device is not None else int(os.environ.get("LOCAL_RANK", "0"))
    if device not in _default:
        _default[device] = Engine(device)
    return _default[device]


class Slot:
    """A scenario slot of an engine, freed when the last Python object that uses it goes away."""
    _next = {}

    def __init__(self, engine):
        self.engine = engine
        key = id(engine)
        Slot._next[key] = Slot._next.get(key, 1 << 20) + 1
        self.id = Slot._next[key]

    def __del__(self):
        try:
            if self.engine._h:
                self.engine.free_scenario(self.id)
        except Exception:
            pass
