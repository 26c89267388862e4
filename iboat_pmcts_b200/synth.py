"""Synthetic "GFS-ensemble-shaped" weather scenarios.

The reference ships no weather data (its ``code/data/.gitignore`` ignores everything), so every
test and benchmark input is generated here, deterministically, from a scenario id.  The grid is the
one the reference's tutorial crops to (``code/Tutorial.py:30,35``: lat [40,50], lon [345,360]) at
the GEFS resolution the reference downloads (0.5 degree, 3-hourly, ``code/model/weatherTLKT.py:80-133``).

Values are rounded to float32 (what the NOMADS download yields) and widened back to float64 (what
``Weather.crop`` produces, ``weatherTLKT.py:192-193``), so a float32 device copy is loss-free.
"""
import numpy as np

GRID_LAT = (40.0, 50.0, 0.5)
GRID_LON = (345.0, 360.0, 0.5)
GRID_TIME = (0.0, 5.0, 0.125)


def axes(lat=GRID_LAT, lon=GRID_LON, time=GRID_TIME):
    """Returns (time, lat, lon) axes; end points included."""
    def ax(spec):
        a, b, d = spec
        return np.arange(a, b + d * 0.02, d)
    return ax(time), ax(lat), ax(lon)


def wind_fields(scenario, lat=GRID_LAT, lon=GRID_LON, time=GRID_TIME):
    """Returns (time[nt], lat[nlat], lon[nlon], u[nt,nlat,nlon], v[...]) as float64 arrays holding
    float32-representable wind values in m/s."""
    t, la, lo = axes(lat, lon, time)
    # Ensemble-shaped: every member perturbs the phases / amplitudes of one shared synoptic
    # pattern (a GEFS member is a perturbed run of the same forecast), plus white noise per grid
    # point so the interpolation is not trivially smooth.
    base = np.random.default_rng(1001).uniform(0.0, 2.0 * np.pi, 4)
    member = np.random.default_rng(3000 + int(scenario))
    ph = base + 0.25 * member.standard_normal(4)
    amp = 1.0 + 0.1 * member.standard_normal(2)
    T, LAT, LON = np.meshgrid(t, la, lo, indexing="ij")
    u = amp[0] * 8.0 * np.sin(0.3 * LAT + ph[0] + T) + 3.0 * np.cos(0.2 * LON + ph[1])
    v = amp[1] * 8.0 * np.cos(0.25 * LON + ph[2] - 0.5 * T) + 3.0 * np.sin(0.2 * LAT + ph[3])
    noise = np.random.default_rng(2000 + int(scenario)).normal(0.0, 1.0, (2,) + u.shape)
    u = (u + noise[0]).astype(np.float32).astype(np.float64)
    v = (v + noise[1]).astype(np.float32).astype(np.float64)
    return t, la, lo, u, v


def octagon_fleet(n, seed=7):
    """Start states and heading offsets of BASELINE config 3 (raw ``doStep`` sweep): boats start in
    lat U(44,46), lon U(351.5,353.5) and walk a noisy octagon -- heading 45 deg * ((h0 + j // 25) % 8)
    at step j -- so nobody leaves the weather grid in 500 steps of 864 s."""
    rng = np.random.default_rng(seed)
    lat0 = rng.uniform(44.0, 46.0, n)
    lon0 = rng.uniform(351.5, 353.5, n)
    h0 = rng.integers(0, 8, n)
    return lat0, lon0, h0


def octagon_headings(h0, nsteps, seg_len=25):
    """Heading segments [nseg, n] (degrees, float64) for :func:`octagon_fleet`."""
    nseg = (nsteps + seg_len - 1) // seg_len
    seg = np.arange(nseg)[:, None]
    return (45.0 * ((h0[None, :] + seg) % 8)).astype(np.float64)
