"""Drop-in mirror of ``code/model/weatherTLKT.py`` (reference) for the rollout path.

``Weather`` keeps the reference's attributes (``lat, lon, time, u, v``) and methods on the path:
``Interpolators`` (weatherTLKT.py:369-378), ``returnPolarVel`` (:148-163), ``getPolarVel`` (:135-146), ``crop``
(:165-218), ``load`` (:51-78).  The two
interpolators are callables with SciPy's ``RegularGridInterpolator`` call convention, evaluated on the GPU
by ``pmcts_interp_wind`` (same evaluation order as SciPy, bit for bit).  Download and plotting methods of
the reference are out of scope.
"""
import math

import numpy as np

from . import _native as N
from .engine import Slot, default_engine
from .utils import out_of_scope


class _GridInterpolator:
    """``scipy.interpolate.RegularGridInterpolator((time, lat, lon), values)`` look-alike, linear,
    ``bounds_error=True``; evaluated by the CUDA engine on the weather's device-resident grid."""

    def __init__(self, weather, component):
        self._weather = weather
        self._component = component
        self.grid = (weather.time, weather.lat, weather.lon)
        self.values = weather.u if component == 0 else weather.v
        self.method = "linear"
        self.bounds_error = True

    def __call__(self, xi):
        xi = np.asarray(xi, dtype=np.float64)
        if xi.shape[-1] != 3:
            raise ValueError("The requested sample points xi have dimension %d, but this "
                             "RegularGridInterpolator has dimension 3" % xi.shape[-1])
        pts = xi.reshape(-1, 3)
        slot = self._weather._weather_slot()
        u, v, st = slot.engine.interp_wind(slot.id, pts[:, 0], pts[:, 1], pts[:, 2])
        if (st != N.ST_OK).any():
            bad = int(np.flatnonzero(st != N.ST_OK)[0])
            axis = 0
            for d, g in enumerate(self.grid):
                if not (g[0] <= pts[bad, d] <= g[-1]):
                    axis = d
                    break
            raise ValueError("One of the requested xi is out of bounds in dimension %d" % axis)
        out = u if self._component == 0 else v
        return out.reshape(xi.shape[:-1]) if xi.ndim > 1 else out.reshape(1)


class Weather:
    """Wind forecast container: ``time`` in days, ``lat`` / ``lon`` in degrees (ascending), ``u`` / ``v`` of
    shape ``(ntime, nlat, nlon)`` in m/s (weatherTLKT.py:40-49)."""

    def __init__(self, lat=None, lon=None, time=None, u=None, v=None):
        self.lat = lat
        self.lon = lon
        self.time = time
        self.u = u
        self.v = v

    @classmethod
    def load(cls, path, latBound=[-90, 90], lonBound=[0, 360], timeSteps=[0, 64]):
        """Loads a pickled ``Weather`` (as saved by the reference's ``Weather.download``,
        weatherTLKT.py:51-78, 128-133) and crops it.  Files written by the reference name the class
        ``weatherTLKT.Weather``: that name is read as this class, whether or not the module alias of INTEGRATION.md
        is in place."""
        import pickle

        class _Unpickler(pickle.Unpickler):
            def find_class(self, module, name):
                if (module, name) == ("weatherTLKT", "Weather"):
                    return cls
                return super().find_class(module, name)

        with open(path, 'rb') as f:
            obj = _Unpickler(f).load()
        return obj.crop(latBound, lonBound, timeSteps)

    # the device copies are caches, not state: they are dropped on copy / pickle
    def __getstate__(self):
        d = dict(self.__dict__)
        for k in ("_wslot", "_wkey", "uInterpolator", "vInterpolator"):
            d.pop(k, None)
        return d

    def _fingerprint(self):
        """Key of the device copy: identity and shape of the fields, the time axis ends, and a checksum of a
        strided sample of u / v (at most 4096 values each), so that an in-place edit of the fields is seen
        without hashing a global grid on every call."""
        u, v = np.asarray(self.u), np.asarray(self.v)
        su, sv = u.reshape(-1), v.reshape(-1)
        step = max(1, su.size // 4096)
        check = (float(np.ma.filled(su[::step], 0.0).sum(dtype=np.float64)),
                 float(np.ma.filled(sv[::step], 0.0).sum(dtype=np.float64)))
        return (id(self.u), id(self.v), float(np.asarray(self.time)[0]), float(np.asarray(self.time)[-1]),
                np.shape(self.u), check)

    def _weather_slot(self):
        key = self._fingerprint()
        if getattr(self, "_wslot", None) is None or self._wkey != key:
            eng = default_engine()
            slot = Slot(eng)
            eng.upload_scenario(slot.id, self.time, self.lat, self.lon, self.u, self.v)
            self._wslot, self._wkey = slot, key
        return self._wslot

    def Interpolators(self):
        """Adds ``uInterpolator`` and ``vInterpolator`` (``u = self.uInterpolator([t, lat, lon])``)."""
        self._wslot = None
        self.uInterpolator = _GridInterpolator(self, 0)
        self.vInterpolator = _GridInterpolator(self, 1)

    def getPolarVel(self):
        """Adds ``wMag`` / ``wAng``: magnitude and direction (degrees, where the wind blows to) of the whole field
        (weatherTLKT.py:135-146).  Element by element the reference's expressions -- ``math.atan2`` rather than NumPy's
        vector form, whose last digit may differ."""
        u, v = np.ma.getdata(self.u), np.ma.getdata(self.v)
        self.wMag = np.array((u ** 2 + v ** 2) ** 0.5, dtype=np.float64)  # (in the fields' own precision, as there)
        atan2 = np.frompyfunc(math.atan2, 2, 1)
        self.wAng = np.mod(180 / math.pi * atan2(u.astype(np.float64), v.astype(np.float64)).astype(np.float64), 360.0)

    @staticmethod
    def returnPolarVel(u, v):
        """Wind magnitude and direction (degrees, clockwise from north, where the wind blows to)."""
        mag = (u ** 2 + v ** 2) ** 0.5
        ang = (180 / math.pi * math.atan2(u, v)) % 360
        return mag, ang

    def crop(self, latBound=[-90, 90], lonBound=[0, 360], timeSteps=[0, 64]):
        """Sub-grid strictly inside the bounds and the given range of time steps (weatherTLKT.py:165-218)."""
        t0, t1 = timeSteps
        if latBound != [-90, 90] or lonBound != [0, 360]:
            lat, lon = np.asarray(self.lat), np.asarray(self.lon)
            ia = np.flatnonzero((lat > latBound[0]) & (lat < latBound[1]))
            io = np.flatnonzero((lon > lonBound[0]) & (lon < lonBound[1]))
            out = Weather()
            out.time = self.time[t0:t1]
            out.lat, out.lon = lat[ia], lon[io]
            out.u = np.array(np.asarray(self.u)[t0:t1][:, ia][:, :, io], dtype=np.float64)
            out.v = np.array(np.asarray(self.v)[t0:t1][:, ia][:, :, io], dtype=np.float64)
            return out
        if timeSteps != [0, 64]:
            out = Weather()
            out.lat, out.lon = self.lat, self.lon
            out.time = self.time[t0:t1]
            out.u = self.u[t0:t1]
            out.v = self.v[t0:t1]
            return out
        return self


for _name in ("download", "plotQuiver", "plotColorQuiver", "plotMultipleQuiver", "animateQuiver"):
    setattr(Weather, _name, staticmethod(out_of_scope("Weather." + _name)))
