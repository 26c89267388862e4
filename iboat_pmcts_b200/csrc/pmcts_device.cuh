// pmcts_b200 device-side simulator: one sailboat transition and its helpers.
//
// Everything here restates, for one CUDA thread, what the reference computes per Simulator.doStep
// call (code/model/simulatorTLKT.py:181-216) and per default-policy iteration
// (code/solver/worker.py:255-300).  Two arithmetic modes:
//   Literal  -- fp64, formula by formula as the reference writes them (the parity anchor);
//   Fast     -- same mathematics re-derived for the GPU: fp32 wind / polar, fp64 geodesy carried as
//               unit vectors so that no full-range transcendental is left in the step.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pmcts_types.h"

namespace pmcts {

// ---- Philox4x32-10 (Salmon et al. 2011); counter layout documented in oracle/pmcts_oracle.c ----
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Per-thread noise stream: one Philox block feeds four consecutive steps.
struct NoiseStream {
    double cache[4];
    int cached_block;
    __device__ NoiseStream() : cached_block(-1) {}

    template <bool kFast>
    __device__ __forceinline__ double draw(const NoiseView &nz, int64_t n, int64_t i, int step) {
        if (nz.mode == PMCTS_NOISE_INJECTED) return nz.z[(size_t)step * n + i];
        if (nz.mode == PMCTS_NOISE_NONE) return 0.0;
        const int block = step >> 2;
        if (block != cached_block) {
            const uint64_t boat = nz.id0 + (uint64_t)i;
            uint32_t x[4];
            philox4x32_10((uint32_t)boat, (uint32_t)(boat >> 32), (uint32_t)block, (uint32_t)nz.stream,
                          (uint32_t)nz.seed, (uint32_t)(nz.seed >> 32), x);
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                if (kFast) {
                    const float u0 = ((float)x[2 * p] + 0.5f) * 2.3283064365386963e-10f;
                    const float u1 = ((float)x[2 * p + 1] + 0.5f) * 2.3283064365386963e-10f;
                    const float r = sqrtf(-2.0f * __logf(fminf(u0, 0.99999994f)));
                    float s, c;
                    __sincosf(6.2831853071795865f * u1, &s, &c);
                    cache[2 * p] = r * c;
                    cache[2 * p + 1] = r * s;
                } else {
                    const double u0 = ((double)x[2 * p] + 0.5) * (1.0 / 4294967296.0);
                    const double u1 = ((double)x[2 * p + 1] + 0.5) * (1.0 / 4294967296.0);
                    const double r = sqrt(-2.0 * log(u0));
                    double s, c;
                    sincos(2.0 * kPi * u1, &s, &c);
                    cache[2 * p] = r * c;
                    cache[2 * p + 1] = r * s;
                }
            }
            cached_block = block;
        }
        return cache[step & 3];
    }
};

// Python's float % for a positive modulus.
__device__ __forceinline__ double py_mod(double a, double m) {
    double r = fmod(a, m);
    if (r < 0.0) r += m;
    return r;
}

// Python's float % 360 on the angles this path takes it of.  In (-360, 720) fmod has nothing to divide: the argument
// comes back as it is, or less 360 -- a subtraction that loses no bit -- and Python's sign fix is the one rounded add
// py_mod does.  Same results bit for bit as py_mod(a, 360.0), without the library's long division (three per literal
// transition); anything else, NaN included, still goes there.
__device__ __forceinline__ double py_mod360(double a) {
    if (a > -360.0 && a < 720.0) {
        if (a < 0.0) return a + 360.0;
        return a >= 360.0 ? a - 360.0 : a;
    }
    return py_mod(a, 360.0);
}

// sin / cos of a position's latitude and longitude (radians = degrees * kDegToRad, as every formula below forms them).
// The literal rollout needs them of the same position up to four times -- getDestination, getDistAndBearing and twice
// fromGeoToCartesian, as B and then as A of the arrival test -- and the same function of the same number is the same
// number: it takes them once (rollout_batch_kernel).
struct PosTrig {
    double sl, cl, so, co;
    bool lon_valid;
};
__device__ __forceinline__ void trig_lat(double lat, PosTrig &t) {
    sincos(lat * kDegToRad, &t.sl, &t.cl);
    t.lon_valid = false;
}
__device__ __forceinline__ void trig_lon(double lon, PosTrig &t) {
    if (!t.lon_valid) sincos(lon * kDegToRad, &t.so, &t.co);
    t.lon_valid = true;
}

// ---- wind: Simulator.getWind (simulatorTLKT.py:168-179) on the time-blended slab of instant k ----
// SciPy's linear RegularGridInterpolator semantics (bounds_error=True; cell index clipped to n-2 so
// the last grid node is reached with weight 1).
__device__ __forceinline__ bool wind_cell(const ScenarioView &sc, double lat, double lon, int &ia,
                                          int &io, double &wa, double &wo) {
    if (!(lat >= sc.lat0 && lat <= sc.lat_last && lon >= sc.lon0 && lon <= sc.lon_last)) return false;
    const double fa = (lat - sc.lat0) * sc.inv_dlat;
    const double fo = (lon - sc.lon0) * sc.inv_dlon;
    ia = min((int)fa, sc.nlat - 2);
    io = min((int)fo, sc.nlon - 2);
    wa = fa - (double)ia;
    wo = fo - (double)io;
    return true;
}

__device__ __forceinline__ bool wind_f64(const ScenarioView &sc, int k, double lat, double lon,
                                         double &u, double &v) {
    int ia, io;
    double wa, wo;
    if (!wind_cell(sc, lat, lon, ia, io, wa, wo) || !sc.k_valid[k]) return false;
    const double2 *row0 = sc.slab64 + ((size_t)k * sc.nlat + ia) * sc.nlon + io;
    const double2 *row1 = row0 + sc.nlon;
    const double2 c00 = __ldg(row0), c01 = __ldg(row0 + 1), c10 = __ldg(row1), c11 = __ldg(row1 + 1);
    const double w00 = (1.0 - wa) * (1.0 - wo), w01 = (1.0 - wa) * wo, w10 = wa * (1.0 - wo), w11 = wa * wo;
    u = ((c00.x * w00 + c01.x * w01) + c10.x * w10) + c11.x * w11;
    v = ((c00.y * w00 + c01.y * w01) + c10.y * w10) + c11.y * w11;
    return true;
}

// ---- Literal fp64 formulas --------------------------------------------------------------------
// numpy polyval2d order: Horner in point-of-sail per column, then Horner in wind.
__device__ __forceinline__ double polar_poly(const double *c, double p, double w) {
    double col[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double acc = c[30 + j];
#pragma unroll
        for (int i = 4; i >= 0; --i) acc = fma(acc, p, c[i * 6 + j]);
        col[j] = acc;
    }
    double acc = col[5];
#pragma unroll
    for (int j = 4; j >= 0; --j) acc = fma(acc, w, col[j]);
    return acc;
}

// Boat.getDeterDyn (simulatorTLKT.py:473-502)
__device__ __forceinline__ double deter_dyn(const BoatParams &bp, double p, double w) {
    if (p > 180.0) p = 360.0 - p;
    w = fmin(w, bp.wind_max);
    if (p < bp.pofsail_min) return bp.cos_pmin * polar_poly(bp.polar, bp.pofsail_min, w) / cos(kPi * p / 180.0);
    if (p > bp.pofsail_max) return bp.cos_pmax * polar_poly(bp.polar, bp.pofsail_max, w) / cos(kPi * p / 180.0);
    return polar_poly(bp.polar, p, w);
}

// Simulator.getDestination (simulatorTLKT.py:132-166)
// (lat_trig: sin / cos of lat * kDegToRad when the caller has them)
__device__ __forceinline__ void get_destination(const BoatParams &bp, double dist, double bearing,
                                                double lat, double lon, double &olat, double &olon,
                                                const PosTrig *lat_trig = nullptr) {
    const double latDep = lat * kDegToRad, lonDep = lon * kDegToRad, brg = bearing * kDegToRad;
    const double dr = dist / bp.earth_radius;
    double sl, cl, sd, cd, sb, cb;
    if (lat_trig) {
        sl = lat_trig->sl;
        cl = lat_trig->cl;
    } else {
        sincos(latDep, &sl, &cl);
    }
    sincos(dr, &sd, &cd);
    sincos(brg, &sb, &cb);
    const double s2 = sl * cd + cl * sd * cb;
    const double latDest = asin(s2);
    const double lonDest = lonDep + atan2(sb * sd * cl, cd - sl * s2);
    olat = latDest * 180.0 / kPi;
    olon = lonDest * 180.0 / kPi;
}

// Bearing part of Simulator.getDistAndBearing (simulatorTLKT.py:124-128); the default policy never
// uses the distance (worker.py:264, 270).
__device__ __forceinline__ double bearing_to(double lat, double lon, double dlat, double dlon) {
    const double latDest = dlat * kDegToRad, latPos = lat * kDegToRad;
    const double dl = dlon * kDegToRad - lon * kDegToRad;
    double sD, cD, sP, cP, sL, cL;
    sincos(latDest, &sD, &cD);
    sincos(latPos, &sP, &cP);
    sincos(dl, &sL, &cL);
    const double x = cD * sL;
    const double y = cP * sD - sP * cD * cL;
    return py_mod360(atan2(x, y) * 180.0 / kPi + 360.0);
}
// The same with sin / cos of both latitudes given (the destination's once per rollout, the position's from its PosTrig).
__device__ __forceinline__ double bearing_to(const PosTrig &pos, double lon, double sD, double cD, double dlon) {
    const double sP = pos.sl, cP = pos.cl;
    const double dl = dlon * kDegToRad - lon * kDegToRad;
    double sL, cL;
    sincos(dl, &sL, &cL);
    const double x = cD * sL;
    const double y = cP * sD - sP * cD * cL;
    return py_mod360(atan2(x, y) * 180.0 / kPi + 360.0);
}

// Haversine distance of Simulator.getDistAndBearing (simulatorTLKT.py:117-122).
__device__ __forceinline__ double distance_to(const BoatParams &bp, double lat, double lon, double dlat,
                                              double dlon) {
    const double latDest = dlat * kDegToRad, latPos = lat * kDegToRad;
    const double s1 = sin((latDest - latPos) / 2), s2 = sin((dlon * kDegToRad - lon * kDegToRad) / 2);
    const double a = s1 * s1 + cos(latPos) * cos(latDest) * s2 * s2;
    return 2 * bp.earth_radius * atan2(sqrt(a), sqrt(1 - a));
}

// Simulator.fromGeoToCartesian (simulatorTLKT.py:218-238): latitude used as colatitude (quirk Q6).
__device__ __forceinline__ void geo_to_cart(double lat, double lon, double &x, double &y, double &z) {
    double sl, cl, so, co;
    sincos(lat * kDegToRad, &sl, &cl);
    sincos(lon * kDegToRad, &so, &co);
    x = sl * co;
    y = sl * so;
    z = cl;
}
__device__ __forceinline__ void geo_to_cart(const PosTrig &t, double &x, double &y, double &z) {
    x = t.sl * t.co;
    y = t.sl * t.so;
    z = t.cl;
}

// Tree.is_state_at_dest (worker.py:380-415) on cartesian points.  Returns 1 / 0, -1 = ZeroDivisionError.
__device__ __forceinline__ int at_dest_cart(const BoatParams &bp, double xa, double ya, double za,
                                            double xb, double yb, double zb, double xd, double yd,
                                            double zd, double &frac) {
    if (ya == 0.0) return -1;
    const double q = yb / ya;
    const double den = zb - q * za;
    if (den == 0.0) return -1;
    const double c = (q * xa - xb) / den;
    const double b = -(xa + c * za) / ya;
    const double d = fabs(xd + b * yd + c * zd) / sqrt(1.0 + b * b + c * c);
    if (asin(d) > bp.dest_angle) return 0;
    const double ax = xd - xa, ay = yd - ya, az = zd - za;  // A -> D
    const double bx = xb - xd, by = yb - yd, bz = zb - zd;  // D -> B
    const double cx = xb - xa, cy = yb - ya, cz = zb - za;  // A -> B
    if (ax * bx + ay * by + az * bz < 0.0) return 0;
    frac = (ax * cx + ay * cy + az * cz) / (cx * cx + cy * cy + cz * cz);
    return 1;
}

// Tree.is_state_terminal (worker.py:417-436)
__device__ __forceinline__ bool is_terminal(const ScenarioView &sc, int k, double lat, double lon) {
    return k >= sc.nT - 1 || lat < sc.box_lat_lo || lat > sc.box_lat_hi || lon < sc.box_lon_lo ||
           lon > sc.box_lon_hi;
}

// Hist.add bin (utils.py:34-47): number of interior thresholds j/8, j = 1..11, strictly below r.
__device__ __forceinline__ int hist_bin(double r) {
    int b = (int)ceil(r * 8.0) - 1;  // r*8 is exact
    return max(0, min(11, b));
}

// Reward of Tree.default_policy (worker.py:276-277).
__device__ __forceinline__ double reward_of(double tmax, double tmin, double t_arr) {
    return (exp((tmax * 1.001 - t_arr) / (tmax * 1.001 - tmin)) - 1.0) / (exp(1.0) - 1.0);
}

// Simulator.doStep (simulatorTLKT.py:181-216), literal fp64.  z is the standard normal of
// Boat.addUncertainty (:504-517).
__device__ __forceinline__ int do_step_literal(const ScenarioView &sc, const BoatParams &bp, int &k,
                                               double &lat, double &lon, double heading, double z,
                                               const PosTrig *lat_trig = nullptr) {
    if (k < 0 || k >= sc.nT) return PMCTS_ST_PAST_HORIZON;
    double u, v;
    if (!wind_f64(sc, k, lat, lon, u, v)) return PMCTS_ST_OUT_OF_GRID;
    const double mag = sqrt(u * u + v * v);                       // Weather.returnPolarVel
    const double ang = py_mod360(180.0 / kPi * atan2(u, v));  // weatherTLKT.py:161-162
    const double pofsail = fabs(py_mod360(ang + 180.0) - heading);
    double speed = deter_dyn(bp, pofsail, mag);
    speed = speed + z * (speed * bp.sigma_coeff);
    if (k + 1 >= sc.nT) return PMCTS_ST_PAST_HORIZON;
    const double dl = speed * sc.dt_sec[k];
    double nlat, nlon;
    get_destination(bp, dl, heading, lat, lon, nlat, nlon, lat_trig);
    k += 1;
    lat = nlat;
    lon = nlon;
    return PMCTS_ST_OK;
}

}  // namespace pmcts
