"""Drop-in mirror of ``code/model/simulatorTLKT.py`` (reference): ``Simulator``, ``Boat`` and the module
constants.  Transitions run on the GPU through the C ABI (``pmcts_step_batch``); there is no CPU
fallback.  Plotting methods of the reference are out of scope.

Scalar calls (``Simulator.doStep``) use the literal fp64 kernel and draw their noise from Python's
``random.gauss`` exactly as ``Boat.addUncertainty`` does (simulatorTLKT.py:504-517), so a seeded script
sees the same random stream as with the reference.  The batched entry points (``Simulator.doStepBatch``,
``Simulator.rollouts``) are the throughput path.
"""
import math
import random as rand
from math import radians as rad

import numpy as np

from . import _native as N
from .engine import Slot, default_engine
from .utils import out_of_scope
from .weatherTLKT import Weather  # noqa: F401  (the reference module exposes it too)

#: Headings the boat can follow (degrees), sorted (simulatorTLKT.py:29).
ACTIONS = tuple(np.arange(0, 360, 45))
A_DICT = {a: i for i, a in enumerate(ACTIONS)}
DAY_TO_SEC = 24 * 60 * 60
SEC_TO_DAYS = 1 / DAY_TO_SEC
EARTH_RADIUS = 6371e3
DESTINATION_ANGLE = rad(0.005)
HOURS_TO_DAY = 1 / 24

_STATUS_ERRORS = {
    N.ST_OUT_OF_GRID: (ValueError, "One of the requested xi is out of bounds"),
    N.ST_PAST_HORIZON: (IndexError, "index out of bounds: no simulator instant after the current one"),
    N.ST_DEGENERATE: (ZeroDivisionError, "float division by zero"),
}


def raise_for_status(status):
    """Turns a per-boat status byte into the exception the reference raises at that point."""
    if status in _STATUS_ERRORS:
        exc, msg = _STATUS_ERRORS[status]
        raise exc(msg)


def sync_params(engine):
    """Pushes the module globals the reference reads at call time (``Boat.*``, ``DESTINATION_ANGLE``,
    ``EARTH_RADIUS``) to the engine when they changed, so monkey-patching them keeps working."""
    key = (tuple(np.asarray(Boat.FIT_VELOCITY, dtype=np.float64).ravel()), float(Boat.POLAR_MAX_WIND_MAG),
           float(Boat.POLAR_MIN_POFSAIL), float(Boat.POLAR_MAX_POFSAIL), float(Boat.UNCERTAINTY_COEFF),
           float(DESTINATION_ANGLE), float(EARTH_RADIUS))
    if getattr(engine, "_param_key", None) != key:
        engine.set_params(key[0], *key[1:])
        engine._param_key = key


class Simulator:
    """Boat + weather interaction on one scenario (simulatorTLKT.py:49-88).  State is
    ``[time index, lat, lon]``; ``times`` in days; ``lats`` / ``lons`` bound the zone of interest."""

    def __init__(self, times, lats, lons, WeatherAvg, stateInit):
        self.times = times
        self.lats = lats
        self.lons = lons
        self.state = list(stateInit)
        self.prevState = list(stateInit)
        WeatherAvg.Interpolators()
        self.uWindAvg = WeatherAvg.uInterpolator
        self.vWindAvg = WeatherAvg.vInterpolator
        self._weather = WeatherAvg
        # the device copy of the scenario, shared with every deep copy of this simulator (Forest.__init__ copies one
        # per worker before anything has been uploaded): whichever copy needs it first uploads it for all
        self._dev = {"slot": None, "key": None}

    # ---- device scenario ---------------------------------------------------------------------
    def __deepcopy__(self, memo):
        # Forest.__init__ deep-copies one Simulator per worker (forest.py:43); the device scenario is
        # immutable, so copies share it
        import copy
        new = Simulator.__new__(Simulator)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ("_dev", "_weather", "uWindAvg", "vWindAvg"):
                setattr(new, k, v)
            else:
                setattr(new, k, copy.deepcopy(v, memo))
        return new

    def _axis_key(self):
        return (np.asarray(self.times, np.float64).tobytes(), float(self.lats[0]), float(self.lats[-1]),
                float(self.lons[0]), float(self.lons[-1]), self._weather._fingerprint())

    def scenario(self):
        """(engine, slot id) of this simulator's scenario on the device, uploaded on first use (and again when the
        axes or the weather change; the check costs a few strided sums, so a caller that loops over many launches
        holds on to the pair)."""
        key = self._axis_key()
        dev = self._dev
        if dev["slot"] is None or dev["key"] != key:
            eng = default_engine()
            slot = Slot(eng)
            w = self._weather
            eng.upload_scenario(slot.id, w.time, w.lat, w.lon, w.u, w.v, self.times,
                                [self.lats[0], self.lats[-1], self.lons[0], self.lons[-1]])
            dev["slot"], dev["key"] = slot, key
        sync_params(dev["slot"].engine)
        return dev["slot"].engine, dev["slot"].id

    # ---- reference API -------------------------------------------------------------------------
    def reset(self, stateInit):
        self.state = list(stateInit)
        self.prevState = list(stateInit)

    def getDistAndBearing(self, position, destination):
        """Great-circle distance (m, haversine) and initial bearing (degrees in [0, 360)) from
        ``position`` to ``destination`` (both ``[lat, lon]`` in degrees); simulatorTLKT.py:100-130."""
        lat_d, lon_d = rad(destination[0]), rad(destination[1])
        lat_p, lon_p = rad(position[0]), rad(position[1])
        a = (math.sin((lat_d - lat_p) / 2)) ** 2 + math.cos(lat_p) * math.cos(lat_d) * (math.sin((lon_d - lon_p) / 2)) ** 2
        distance = 2 * EARTH_RADIUS * math.atan2(a ** 0.5, (1 - a) ** 0.5)
        x = math.cos(lat_d) * math.sin(lon_d - lon_p)
        y = math.cos(lat_p) * math.sin(lat_d) - math.sin(lat_p) * math.cos(lat_d) * math.cos(lon_d - lon_p)
        bearing = (math.atan2(x, y) * 180 / math.pi + 360) % 360
        return distance, bearing

    def getDestination(self, distance, bearing, departure):
        """Point reached from ``departure`` after ``distance`` metres on the great circle of initial
        ``bearing`` (no longitude wrap); simulatorTLKT.py:132-166."""
        lat0, lon0 = rad(departure[0]), rad(departure[1])
        brg = rad(bearing)
        delta = distance / EARTH_RADIUS
        lat1 = math.asin(math.sin(lat0) * math.cos(delta) + math.cos(lat0) * math.sin(delta) * math.cos(brg))
        lon1 = lon0 + math.atan2(math.sin(brg) * math.sin(delta) * math.cos(lat0),
                                 math.cos(delta) - math.sin(lat0) * math.sin(lat1))
        return [lat1 * 180 / math.pi, lon1 * 180 / math.pi]

    def getWind(self):
        """Wind at the current state: (u toward east, v toward north) in m/s, as SciPy returns them."""
        p = [self.times[self.state[0]], self.state[1], self.state[2]]
        return self.uWindAvg(p), self.vWindAvg(p)

    def doStep(self, action):
        """One transition with heading ``action`` (degrees).  Returns a reference to ``self.state``.
        Raises what the reference raises: ``ValueError`` outside the weather grid, ``IndexError`` at the
        last time index."""
        eng, scen = self.scenario()
        self.prevState = list(self.state)
        z = rand.gauss(0.0, 1.0)  # Boat.addUncertainty: gauss(v, v * c) == v + z * (v * c)
        k = np.array([self.state[0]], np.int32)
        lat = np.array([self.state[1]], np.float64)
        lon = np.array([self.state[2]], np.float64)
        st = np.zeros(1, np.uint8)
        prec = eng.precision
        eng.set_precision(N.PREC_F64)
        try:
            eng.step_batch(scen, k, lat, lon, np.array([[float(action)]]), nsteps=1, status=st,
                           z=np.array([[z]], np.float64))
        finally:
            eng.set_precision(prec)
        raise_for_status(int(st[0]))
        self.state[0], self.state[1], self.state[2] = int(k[0]), float(lat[0]), float(lon[0])
        return self.state

    @staticmethod
    def fromGeoToCartesian(coordinates):
        """The reference's mapping (simulatorTLKT.py:218-238), latitude used as a colatitude."""
        lat, lon = rad(coordinates[0]), rad(coordinates[1])
        return [math.sin(lat) * math.cos(lon), math.sin(lat) * math.sin(lon), math.cos(lat)]

    # ---- batched entry points (no counterpart in the reference; what its loops turn into) ----------
    def doStepBatch(self, k, lat, lon, headings, nsteps=1, seg_len=1, z=None, seed=None, stream=0, id0=0,
                    status=None, flags=0):
        """``doStep`` for arrays of boats, ``nsteps`` fused; arrays are updated in place (NumPy arrays:
        host round trip; CUDA tensors: asynchronous).  ``headings[(j // seg_len), i]`` in degrees."""
        eng, scen = self.scenario()
        eng.step_batch(scen, k, lat, lon, headings, nsteps=nsteps, seg_len=seg_len, status=status, z=z, seed=seed,
                       stream=stream, id0=id0, flags=flags)

    def rollouts(self, n, root, destination, timemin, prefix=None, prefix_len=None, z=None, seed=None, stream=0,
                 id0=0, replicas=1, flags=0, want=None):
        """``n * replicas`` default-policy rollouts from ``root`` (see ``Engine.rollout_batch_host``)."""
        eng, scen = self.scenario()
        return eng.rollout_batch_host(scen, n, replicas, root, destination, timemin, prefix=prefix,
                                      prefix_len=prefix_len, z=z, seed=seed, stream=stream, id0=id0, flags=flags,
                                      want=want)


class Boat:
    """Boat dynamics (simulatorTLKT.py:447-517).  The class attributes are read at every launch."""

    #: 5th-order polynomial fit of the velocity polar: ``FIT_VELOCITY[i][j]`` multiplies pOfSail^i * wind^j.
    FIT_VELOCITY = ((-0.0310198207067136, -0.0600881995286002, 0.0286695485969272, -0.00684406929296715,
                     0.000379636836557256, -6.77704610076153e-06),
                    (0.0106968590293653, 0.00665508747173206, -4.03686836415123e-05, 2.38962919033178e-05,
                     -6.16724919464073e-07, 0),
                    (-0.000370260554113750, -7.68003222256045e-05, -2.44425618051996e-06, -2.16961875441841e-08, 0, 0),
                    (4.94175336099537e-06, 5.68757177397705e-07, 1.17646387245609e-08, 0, 0, 0),
                    (-2.91900968067076e-08, -1.67287818401469e-09, 0, 0, 0, 0),
                    (6.34463723894578e-11, 0, 0, 0, 0, 0))
    POLAR_MAX_WIND_MAG = 19.1
    POLAR_MIN_POFSAIL = 35
    POLAR_MAX_POFSAIL = 160
    UNCERTAINTY_COEFF = 0.2

    @staticmethod
    def getDeterDyn(pOfSail, windMag, fitCoeffs):
        """Deterministic boat speed (m/s) for a point of sail (degrees) and a wind magnitude (m/s)."""
        polyval2d = np.polynomial.polynomial.polyval2d
        p = 360 - pOfSail if pOfSail > 180 else pOfSail
        w = min(windMag, Boat.POLAR_MAX_WIND_MAG)
        for limit, beyond in ((Boat.POLAR_MIN_POFSAIL, p < Boat.POLAR_MIN_POFSAIL),
                              (Boat.POLAR_MAX_POFSAIL, p > Boat.POLAR_MAX_POFSAIL)):
            if beyond:  # tacking / gybing: the speed made good along the limit angle
                at_limit = math.cos(math.pi * limit / 180) * polyval2d(limit, w, fitCoeffs)
                return at_limit / (math.cos(math.pi * p / 180))
        return polyval2d(p, w, fitCoeffs)

    @staticmethod
    def addUncertainty(boatSpeed):
        """Noisy boat speed: one draw from Python's global Mersenne Twister, as in the reference."""
        return rand.gauss(boatSpeed, boatSpeed * Boat.UNCERTAINTY_COEFF)


for _name in ("prepareBaseMap", "plotTraj", "animateTraj", "play_scenario"):
    setattr(Simulator, _name, staticmethod(out_of_scope("Simulator." + _name)))
