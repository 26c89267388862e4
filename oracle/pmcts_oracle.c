/* TEST INFRASTRUCTURE ONLY -- see pmcts_oracle.h.  CPU fp64 restatement of the IBoat-PMCTS
 * rollout hot path, written to follow the reference's evaluation order operation by operation so
 * that it can be compared bit-for-bit with the Python reference run in the build container.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fno-builtin-pow -pthread -shared -fPIC (oracle/Makefile).
 * -ffp-contract=off matters: CPython never fuses a*b+c.
 *
 * Citations are file:line in the reference checkout (PBarde/IBoat-PMCTS).
 */
#include "pmcts_oracle.h"

#include <math.h>
#include <stddef.h>
#include <string.h>
#include <pthread.h>

#define ORC_PI 3.141592653589793 /* math.pi */

/* Boat.FIT_VELOCITY and Boat.* constants, code/model/simulatorTLKT.py:454-471;
 * EARTH_RADIUS :40, DESTINATION_ANGLE = radians(0.005) :43. */
static const double FIT_VELOCITY[36] = {
    -0.0310198207067136, -0.0600881995286002, 0.0286695485969272, -0.00684406929296715, 0.000379636836557256, -6.77704610076153e-06,
    0.0106968590293653, 0.00665508747173206, -4.03686836415123e-05, 2.38962919033178e-05, -6.16724919464073e-07, 0,
    -0.000370260554113750, -7.68003222256045e-05, -2.44425618051996e-06, -2.16961875441841e-08, 0, 0,
    4.94175336099537e-06, 5.68757177397705e-07, 1.17646387245609e-08, 0, 0, 0,
    -2.91900968067076e-08, -1.67287818401469e-09, 0, 0, 0, 0,
    6.34463723894578e-11, 0, 0, 0, 0, 0};

void orc_default_params(orc_sim *s) {
    memcpy(s->polar, FIT_VELOCITY, sizeof(FIT_VELOCITY));
    s->wind_max = 19.1;
    s->pofsail_min = 35;
    s->pofsail_max = 160;
    s->sigma_coeff = 0.2;
    s->earth_radius = 6371e3;
    s->dest_angle = 0.005 * (ORC_PI / 180.0); /* math.radians(0.005) */
}

/* math.radians: CPython multiplies by the constant pi/180. */
static inline double py_rad(double x) { return x * (ORC_PI / 180.0); }

/* Python float %: fmod, then move the result to the sign of the divisor. */
static inline double py_mod(double a, double b) {
    double m = fmod(a, b);
    if (m != 0.0) {
        if ((b < 0) != (m < 0)) m += b;
    } else {
        m = copysign(0.0, b);
    }
    return m;
}

/* SciPy RegularGridInterpolator.find_indices for one axis (compiled _rgi_cython.find_indices,
 * behaviour characterised in SURVEY.md section 8 a3): i = clip(searchsorted_right(g, x) - 1, 0, n-2),
 * weight (x - g[i]) / (g[i+1] - g[i]). */
static inline int axis_cell(const double *g, int n, double x, double *w) {
    int lo = 0, hi = n; /* first index with g[idx] > x */
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (g[mid] <= x) lo = mid + 1; else hi = mid;
    }
    int i = lo - 1;
    if (i < 0) i = 0;
    if (i > n - 2) i = n - 2;
    *w = (x - g[i]) / (g[i + 1] - g[i]);
    return i;
}

/* Simulator.getWind (simulatorTLKT.py:168-179) -> two SciPy linear RegularGridInterpolator calls
 * built at weatherTLKT.py:377-378 with bounds_error=True.  _evaluate_linear (scipy/interpolate/
 * _rgi.py:520-549): corners in itertools.product order (time slowest, lon fastest), weight
 * ((1*wt)*wlat)*wlon, accumulated from 0.0.  Returns -1 where SciPy raises ValueError. */
int orc_interp(const orc_weather *w, double t, double lat, double lon, double *u, double *v) {
    if (!(w->time[0] <= t && t <= w->time[w->nt - 1])) return -1;
    if (!(w->lat[0] <= lat && lat <= w->lat[w->nlat - 1])) return -1;
    if (!(w->lon[0] <= lon && lon <= w->lon[w->nlon - 1])) return -1;
    double yt, ya, yo;
    int it = axis_cell(w->time, w->nt, t, &yt);
    int ia = axis_cell(w->lat, w->nlat, lat, &ya);
    int io = axis_cell(w->lon, w->nlon, lon, &yo);
    double wt[2] = {1 - yt, yt}, wa[2] = {1 - ya, ya}, wo[2] = {1 - yo, yo};
    double su = 0.0, sv = 0.0;
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b)
            for (int c = 0; c < 2; ++c) {
                double weight = 1.0;
                weight = weight * wt[a];
                weight = weight * wa[b];
                weight = weight * wo[c];
                size_t idx = ((size_t)(it + a) * w->nlat + (ia + b)) * w->nlon + (io + c);
                su = su + w->u[idx] * weight;
                sv = sv + w->v[idx] * weight;
            }
    *u = su;
    *v = sv;
    return 0;
}

/* Weather.returnPolarVel, code/model/weatherTLKT.py:148-163 (CPython ** is libm pow). */
void orc_polar_vel(double u, double v, double *mag, double *ang) {
    *mag = pow(pow(u, 2.0) + pow(v, 2.0), 0.5);
    *ang = py_mod(180 / ORC_PI * atan2(u, v), 360.0);
}

/* numpy.polynomial.polynomial.polyval2d(p, w, c): Horner in p down the rows for every column,
 * then Horner in w (numpy polyval starts with c[-1] + x*0). */
static double polar_poly(const double *c, double p, double w) {
    double col[6];
    for (int j = 0; j < 6; ++j) {
        double acc = c[5 * 6 + j] + p * 0;
        for (int i = 4; i >= 0; --i) acc = c[i * 6 + j] + acc * p;
        col[j] = acc;
    }
    double acc = col[5] + w * 0;
    for (int j = 4; j >= 0; --j) acc = col[j] + acc * w;
    return acc;
}

/* Boat.getDeterDyn, code/model/simulatorTLKT.py:473-502. */
double orc_deter_dyn(const orc_sim *s, double p, double w) {
    if (p > 180) p = 360 - p;
    if (w > s->wind_max) w = s->wind_max;
    if (p < s->pofsail_min) {
        double v = cos(ORC_PI * s->pofsail_min / 180) * polar_poly(s->polar, s->pofsail_min, w);
        return v / cos(ORC_PI * p / 180);
    } else if (p > s->pofsail_max) {
        double v = cos(ORC_PI * s->pofsail_max / 180) * polar_poly(s->polar, s->pofsail_max, w);
        return v / cos(ORC_PI * p / 180);
    }
    return polar_poly(s->polar, p, w);
}

/* Simulator.getDestination, code/model/simulatorTLKT.py:132-166 (no longitude wrap). */
void orc_get_destination(const orc_sim *s, double dist, double bearing, double lat, double lon,
                         double *olat, double *olon) {
    double latDep = py_rad(lat), lonDep = py_rad(lon);
    double brg = py_rad(bearing);
    double dr = dist / s->earth_radius;
    double latDest = asin(sin(latDep) * cos(dr) + cos(latDep) * sin(dr) * cos(brg));
    double lonDest = lonDep + atan2(sin(brg) * sin(dr) * cos(latDep), cos(dr) - sin(latDep) * sin(latDest));
    *olat = latDest * 180 / ORC_PI;
    *olon = lonDest * 180 / ORC_PI;
}

/* Simulator.getDistAndBearing, code/model/simulatorTLKT.py:100-130. */
void orc_dist_bearing(const orc_sim *s, double lat, double lon, double dlat, double dlon,
                      double *dist, double *bearing) {
    double latDest = py_rad(dlat), lonDest = py_rad(dlon);
    double latPos = py_rad(lat), lonPos = py_rad(lon);
    double a = pow(sin((latDest - latPos) / 2), 2.0) +
               cos(latPos) * cos(latDest) * pow(sin((lonDest - lonPos) / 2), 2.0);
    if (dist) *dist = 2 * s->earth_radius * atan2(pow(a, 0.5), pow(1 - a, 0.5));
    double x = cos(latDest) * sin(lonDest - lonPos);
    double y = cos(latPos) * sin(latDest) - sin(latPos) * cos(latDest) * cos(lonDest - lonPos);
    *bearing = py_mod(atan2(x, y) * 180 / ORC_PI + 360, 360.0);
}

/* Simulator.fromGeoToCartesian, code/model/simulatorTLKT.py:218-238 (latitude used as
 * colatitude -- quirk Q6, reproduced on purpose). */
void orc_geo_to_cart(double lat, double lon, double xyz[3]) {
    lat = py_rad(lat);
    lon = py_rad(lon);
    xyz[0] = sin(lat) * cos(lon);
    xyz[1] = sin(lat) * sin(lon);
    xyz[2] = cos(lat);
}

/* numpy.dot of two length-3 float64 vectors as evaluated by the OpenBLAS (0.3.30, Haswell kernel)
 * that NumPy 2.3.5 bundles: x0*y0 rounded, then two fused multiply-adds.  Measured against np.dot
 * on 20 000 random vectors in the build container: 20 000 / 20 000 bit-equal (a plain
 * left-to-right sum matches only 67 %).  The difference is <= 1 ulp either way. */
static inline double dot3(const double *x, const double *y) {
    return fma(x[2], y[2], fma(x[1], y[1], x[0] * y[0]));
}

/* Tree.is_state_at_dest, code/solver/worker.py:380-415.  Returns 1 / 0, or -1 where CPython
 * raises ZeroDivisionError (ya == 0 or a degenerate plane). */
int orc_is_at_dest(const orc_sim *s, const double dest[2], const double a[2], const double b[2],
                   double *frac) {
    double A[3], B[3], D[3];
    orc_geo_to_cart(a[0], a[1], A);
    orc_geo_to_cart(b[0], b[1], B);
    orc_geo_to_cart(dest[0], dest[1], D);
    if (A[1] == 0.0) return -1;
    double den = B[2] - B[1] / A[1] * A[2];
    if (den == 0.0) return -1;
    double c = (B[1] / A[1] * A[0] - B[0]) / den;
    double bb = -(A[0] + c * A[2]) / A[1];
    double d = fabs(D[0] + bb * D[1] + c * D[2]) / sqrt(1 + pow(bb, 2.0) + pow(c, 2.0));
    double alpha = asin(d);
    if (alpha > s->dest_angle) return 0;
    double vad[3], vdb[3], vab[3];
    for (int i = 0; i < 3; ++i) {
        vad[i] = D[i] - A[i];
        vdb[i] = B[i] - D[i];
        vab[i] = B[i] - A[i];
    }
    double p = dot3(vad, vdb);
    if (p < 0) return 0;
    *frac = dot3(vad, vab) / dot3(vab, vab);
    return 1;
}

/* Tree.is_state_terminal, code/solver/worker.py:417-436. */
int orc_is_terminal(const orc_sim *s, int k, double lat, double lon) {
    if (s->times[k] == s->times[s->nT - 1]) return 1;
    if (lat < s->lat_lo || lat > s->lat_hi) return 1;
    if (lon < s->lon_lo || lon > s->lon_hi) return 1;
    return 0;
}

/* Simulator.doStep, code/model/simulatorTLKT.py:181-216, with Boat.addUncertainty (:504-517)
 * written as random.gauss(v, s) = v + z*s for an explicit standard normal z.
 * Returns ORC_OK, ORC_OUT_OF_GRID (SciPy ValueError) or ORC_PAST_HORIZON (IndexError at :208;
 * note the reference evaluates the wind first, so an out-of-grid query wins). */
int orc_do_step(const orc_sim *s, int *k, double *lat, double *lon, double action, double z) {
    double u, v, mag, ang;
    if (*k < 0 || *k >= s->nT) return ORC_PAST_HORIZON;
    if (orc_interp(&s->w, s->times[*k], *lat, *lon, &u, &v)) return ORC_OUT_OF_GRID;
    orc_polar_vel(u, v, &mag, &ang);
    double pofsail = fabs(py_mod(ang + 180, 360.0) - action);
    double speed = orc_deter_dyn(s, pofsail, mag);
    speed = speed + z * (speed * s->sigma_coeff);
    if (*k + 1 >= s->nT) return ORC_PAST_HORIZON;
    double dt = (s->times[*k + 1] - s->times[*k]) * (24 * 60 * 60);
    double dl = speed * dt;
    double nlat, nlon;
    orc_get_destination(s, dl, action, *lat, *lon, &nlat, &nlon);
    *k += 1;
    *lat = nlat;
    *lon = nlon;
    return ORC_OK;
}

/* Hist.add, code/solver/utils.py:34-47 with THRESH = arange(0, 1.625, 0.125) (:23): the bin is the
 * number of interior thresholds 0.125 .. 1.375 strictly below the value. */
int orc_hist_bin(double reward) {
    int i = 0;
    for (int j = 1; j <= 11; ++j) {
        if (reward > j * 0.125) ++i; else break;
    }
    return i;
}

/* Reward model of Tree.default_policy, code/solver/worker.py:276-277. */
double orc_reward(double tmax, double tmin, double t_arr) {
    return (exp((tmax * 1.001 - t_arr) / (tmax * 1.001 - tmin)) - 1) / (exp(1) - 1);
}

/* ---- Philox4x32-10 (Salmon et al. 2011) and the per-(boat, step) normal draw ---------------- */
void orc_philox4x32(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Counter layout: ctr = (boat_lo, boat_hi, step / 4, stream), key = (seed_lo, seed_hi).  One
 * Philox block feeds four consecutive steps: words (0,1) -> Box-Muller pair for steps 4q, 4q+1,
 * words (2,3) -> steps 4q+2, 4q+3.  u = (x + 0.5) * 2^-32 in (0, 1). */
double orc_philox_normal(uint64_t seed, uint64_t stream, uint64_t boat, uint32_t step) {
    uint32_t ctr[4] = {(uint32_t)boat, (uint32_t)(boat >> 32), step >> 2, (uint32_t)stream};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t x[4];
    orc_philox4x32(ctr, key, x);
    int pair = (step >> 1) & 1;
    double u0 = ((double)x[2 * pair] + 0.5) * (1.0 / 4294967296.0);
    double u1 = ((double)x[2 * pair + 1] + 0.5) * (1.0 / 4294967296.0);
    double r = sqrt(-2.0 * log(u0));
    double th = 2.0 * ORC_PI * u1;
    return (step & 1) ? r * sin(th) : r * cos(th);
}

static inline double draw(const orc_noise *nz, int64_t n, int64_t i, int step) {
    if (nz->z) return nz->z[(size_t)step * n + i];
    return orc_philox_normal(nz->seed, nz->stream, nz->id0 + (uint64_t)i, (uint32_t)step);
}

/* ---- batched drivers: a static pthread split over boats / rollouts -------------------------- */
typedef void (*item_fn)(void *ctx, int64_t i);
typedef struct { item_fn fn; void *ctx; int64_t lo, hi; } span_t;

static void *span_run(void *p) {
    span_t *s = (span_t *)p;
    for (int64_t i = s->lo; i < s->hi; ++i) s->fn(s->ctx, i);
    return NULL;
}

static void parallel_for(int64_t n, int nthreads, item_fn fn, void *ctx) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if ((int64_t)nthreads > n) nthreads = n > 0 ? (int)n : 1;
    span_t spans[256];
    pthread_t tid[256];
    int64_t chunk = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; ++t) {
        spans[t].fn = fn; spans[t].ctx = ctx;
        spans[t].lo = t * chunk;
        spans[t].hi = (t + 1) * chunk < n ? (t + 1) * chunk : n;
    }
    for (int t = 1; t < nthreads; ++t) pthread_create(&tid[t], NULL, span_run, &spans[t]);
    span_run(&spans[0]);
    for (int t = 1; t < nthreads; ++t) pthread_join(tid[t], NULL);
}

typedef struct {
    const orc_sim *s; int64_t n; int nsteps; int32_t *k; double *lat, *lon, *prev_lat, *prev_lon;
    const double *heading; int seg_len; const orc_noise *noise; uint8_t *status;
    double *traj_lat, *traj_lon; int flags;
} step_ctx;

static void step_one(void *p, int64_t i) {
    step_ctx *c = (step_ctx *)p;
    const int64_t n = c->n;
    int kk = c->k[i];
    double la = c->lat[i], lo = c->lon[i];
    double pla = c->prev_lat ? c->prev_lat[i] : la, plo = c->prev_lon ? c->prev_lon[i] : lo;
    int st = c->status ? c->status[i] : 0;
    for (int j = 0; j < c->nsteps && st == ORC_OK; ++j) {
        double h = c->heading[(size_t)(j / c->seg_len) * n + i];
        double la0 = la, lo0 = lo;
        if (c->flags & ORC_STEP_STOP_BEFORE_TERMINAL) {
            /* forest.py:239: stop while the *next* state would be terminal */
            if (kk + 1 >= c->s->nT) { st = ORC_PAST_HORIZON; break; }
            if (orc_is_terminal(c->s, kk + 1, la, lo)) {
                st = (c->s->times[kk + 1] == c->s->times[c->s->nT - 1]) ? ORC_HORIZON : ORC_OUT_OF_BOX;
                break;
            }
        }
        st = orc_do_step(c->s, &kk, &la, &lo, h, draw(c->noise, n, i, j));
        if (st == ORC_OK) {
            pla = la0; plo = lo0; /* doStep copies state into prevState first (:196) */
            if (c->traj_lat) c->traj_lat[(size_t)j * n + i] = la;
            if (c->traj_lon) c->traj_lon[(size_t)j * n + i] = lo;
        }
    }
    c->k[i] = kk; c->lat[i] = la; c->lon[i] = lo;
    if (c->prev_lat) c->prev_lat[i] = pla;
    if (c->prev_lon) c->prev_lon[i] = plo;
    if (c->status) c->status[i] = (uint8_t)st;
}

void orc_step_batch(const orc_sim *s, int64_t n, int nsteps, int32_t *k, double *lat, double *lon,
                    double *prev_lat, double *prev_lon, const double *heading, int seg_len,
                    const orc_noise *noise, uint8_t *status, double *traj_lat, double *traj_lon,
                    int flags, int nthreads) {
    step_ctx c = {s, n, nsteps, k, lat, lon, prev_lat, prev_lon, heading, seg_len, noise, status,
                  traj_lat, traj_lon, flags};
    parallel_for(n, nthreads, step_one, &c);
}

typedef struct {
    const orc_sim *s; int64_t n; const double *root, *dest; double tmin; const double *prefix;
    const int32_t *prefix_len; int prefix_stride, replay_full, mode; const orc_noise *noise;
    double *reward, *t_arr; int32_t *steps; uint8_t *status; int32_t *bin;
    double *frac_out, *final_lat, *final_lon, *traj_lat, *traj_lon; int traj_stride;
} roll_ctx;

static void rollout_one(void *p, int64_t i) {
    roll_ctx *c = (roll_ctx *)p;
    const orc_sim *s = c->s;
    const int64_t n = c->n;
    const double *dest = c->dest;
    const double tmax = s->times[s->nT - 1];
    int k = (int)c->root[0];
    double A[2] = {c->root[1], c->root[2]}, B[2] = {c->root[1], c->root[2]};
    int st = ORC_OK, nstep = 0, at = 0;
    double frac = 0.0, brg = 0.0;
    const double *pre = c->prefix ? c->prefix + (size_t)i * c->prefix_stride : NULL;
    int plen = c->prefix_len ? c->prefix_len[i] : 0;
#define STEP(action)                                                                        \
    do {                                                                                    \
        double la_ = B[0], lo_ = B[1];                                                      \
        st = orc_do_step(s, &k, &B[0], &B[1], (action), draw(c->noise, n, i, nstep));       \
        if (st == ORC_OK) {                                                                 \
            A[0] = la_; A[1] = lo_;                                                         \
            if (c->traj_lat && nstep < c->traj_stride) c->traj_lat[(size_t)nstep * n + i] = B[0]; \
            if (c->traj_lon && nstep < c->traj_stride) c->traj_lon[(size_t)nstep * n + i] = B[1]; \
            ++nstep;                                                                        \
        }                                                                                   \
    } while (0)
#define ATDEST()                                                                            \
    do {                                                                                    \
        int r_ = orc_is_at_dest(s, dest, A, B, &frac);                                      \
        if (r_ < 0) st = ORC_DEGENERATE; else at = r_;                                      \
    } while (0)
    if (c->mode == 0) {
        /* Tree.get_sim_to_estimate_state, worker.py:283-300: first action unconditionally; the
         * remaining ones only with replay_full (in the reference the loop guard negates a
         * non-empty list and never runs -- quirk Q1). */
        if (plen > 0) STEP(pre[0]);
        if (c->replay_full) {
            for (int a = 1; a < plen && st == ORC_OK; ++a) {
                if (orc_is_terminal(s, k, B[0], B[1])) break;
                ATDEST();
                if (st != ORC_OK || at) break;
                STEP(pre[a]);
            }
        }
    } else {
        /* estimate_perfomance_plan, isochrones.py:498-511: every plan action, unchecked; an empty
         * plan takes one step towards the destination (also forest.py:271-272). */
        for (int a = 0; a < plen && st == ORC_OK; ++a) STEP(pre[a]);
        if (plen == 0) {
            orc_dist_bearing(s, B[0], B[1], dest[0], dest[1], NULL, &brg);
            STEP(brg);
        }
    }
    /* Tree.default_policy, worker.py:264-271 (same loop at isochrones.py:513-520, forest.py:275-282). */
    if (st == ORC_OK) {
        orc_dist_bearing(s, B[0], B[1], dest[0], dest[1], NULL, &brg);
        ATDEST();
        while (st == ORC_OK && !at && !orc_is_terminal(s, k, B[0], B[1])) {
            STEP(brg);
            if (st != ORC_OK) break;
            orc_dist_bearing(s, B[0], B[1], dest[0], dest[1], NULL, &brg);
            ATDEST();
        }
    }
#undef STEP
#undef ATDEST
    double ta = NAN, rw = 0.0;
    if (st == ORC_OK) {
        if (at) {
            ta = s->times[k] - (1 - frac) * (s->times[k] - s->times[k - 1]); /* worker.py:274-275 */
            rw = orc_reward(tmax, c->tmin, ta);
            st = ORC_ARRIVED;
        } else {
            st = (s->times[k] == tmax) ? ORC_HORIZON : ORC_OUT_OF_BOX;
            if (c->mode == 1) ta = tmax; /* isochrones.py:527-530 */
        }
    }
    if (c->reward) c->reward[i] = rw;
    if (c->t_arr) c->t_arr[i] = ta;
    if (c->steps) c->steps[i] = nstep;
    if (c->status) c->status[i] = (uint8_t)st;
    if (c->bin) c->bin[i] = orc_hist_bin(rw);
    if (c->frac_out) c->frac_out[i] = (st == ORC_ARRIVED) ? frac : NAN;
    if (c->final_lat) c->final_lat[i] = B[0];
    if (c->final_lon) c->final_lon[i] = B[1];
}

void orc_rollout_batch(const orc_sim *s, int64_t n, const double root[3], const double dest[2],
                       double tmin, const double *prefix, const int32_t *prefix_len,
                       int prefix_stride, int replay_full, int mode, const orc_noise *noise,
                       double *reward, double *t_arr, int32_t *steps, uint8_t *status,
                       int32_t *bin, double *frac_out, double *final_lat, double *final_lon,
                       double *traj_lat, double *traj_lon, int traj_stride, int nthreads) {
    roll_ctx c = {s, n, root, dest, tmin, prefix, prefix_len, prefix_stride, replay_full, mode,
                  noise, reward, t_arr, steps, status, bin, frac_out, final_lat, final_lon,
                  traj_lat, traj_lon, traj_stride};
    parallel_for(n, nthreads, rollout_one, &c);
}
