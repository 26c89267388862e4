// TEST INFRASTRUCTURE ONLY -- host build of the mixed-precision simulator (csrc/pmcts_fast.cuh).
//
// The product runs pmcts_fast.cuh on the GPU only.  This file compiles the very same header with
// g++ (PMCTS_HD expands to `inline`, the MUFU approximations to their IEEE stand-ins) so that the
// CPU test-suite can measure the arithmetic error of the re-derived formulas against the fp64 oracle
// and check that every decision the fp32 path takes outside its margins agrees with it.
// Nothing in iboat_pmcts_b200/ loads this library.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../iboat_pmcts_b200/csrc/pmcts_fast.cuh"
#include "../../iboat_pmcts_b200/csrc/pmcts_host.h"

using namespace pmcts;

namespace {

const double kFit[36] = {
    -0.0310198207067136, -0.0600881995286002, 0.0286695485969272, -0.00684406929296715, 0.000379636836557256, -6.77704610076153e-06,
    0.0106968590293653, 0.00665508747173206, -4.03686836415123e-05, 2.38962919033178e-05, -6.16724919464073e-07, 0,
    -0.000370260554113750, -7.68003222256045e-05, -2.44425618051996e-06, -2.16961875441841e-08, 0, 0,
    4.94175336099537e-06, 5.68757177397705e-07, 1.17646387245609e-08, 0, 0, 0,
    -2.91900968067076e-08, -1.67287818401469e-09, 0, 0, 0, 0,
    6.34463723894578e-11, 0, 0, 0, 0, 0};

struct Model {
    ScenarioView sc{};
    BoatParams bp{};
    std::vector<float2> slab32;
    std::vector<double> times;
    TimeCells tc;
    bool force_generic = false;  // fm_force_generic: run the generic variant even for the default boat
    int last_variant = 0;        // 0 generic, 1 default boat, 2 turbo
};

}  // namespace

extern "C" {

void *fm_create(const double *wtime, int nt, const double *wlat, int nlat, const double *wlon, int nlon,
                const double *u, const double *v, const double *sim_times, int nT, const double box[4],
                double sigma) {
    Model *m = new Model();
    m->tc = time_cells(wtime, nt, sim_times, nT);
    m->times.assign(sim_times, sim_times + nT);
    const size_t plane = (size_t)nlat * nlon;
    const size_t stride = (plane + 1) & ~(size_t)1;
    m->slab32.resize(stride * nT);
    for (int k = 0; k < nT; ++k)
        for (size_t p = 0; p < plane; ++p) {
            const size_t lo = (size_t)m->tc.it[k] * plane + p, hi = lo + plane;
            const double w = m->tc.wt[k];
            m->slab32[k * stride + p].x = (float)(u[lo] * (1.0 - w) + u[hi] * w);
            m->slab32[k * stride + p].y = (float)(v[lo] * (1.0 - w) + v[hi] * w);
        }
    ScenarioView &s = m->sc;
    s.slab32 = m->slab32.data();
    s.slab32_stride = (int)stride;
    s.times = m->times.data();
    s.dt_sec = m->tc.dt.data();
    s.dt_sec32 = m->tc.dt32.data();
    s.k_valid = m->tc.valid.data();
    s.kv_lo = m->tc.kv_lo;
    s.kv_hi = m->tc.kv_hi;
    s.nT = nT;
    s.nlat = nlat;
    s.nlon = nlon;
    s.lat0 = wlat[0];
    s.lon0 = wlon[0];
    s.inv_dlat = 1.0 / ((wlat[nlat - 1] - wlat[0]) / (nlat - 1));
    s.inv_dlon = 1.0 / ((wlon[nlon - 1] - wlon[0]) / (nlon - 1));
    s.lat_last = wlat[nlat - 1];
    s.lon_last = wlon[nlon - 1];
    s.box_lat_lo = box[0]; s.box_lat_hi = box[1]; s.box_lon_lo = box[2]; s.box_lon_hi = box[3];
    s.box_fa_lo = (float)((box[0] - s.lat0) * s.inv_dlat);
    s.box_fa_hi = (float)((box[1] - s.lat0) * s.inv_dlat);
    s.box_fo_lo = (float)((box[2] - s.lon0) * s.inv_dlon);
    s.box_fo_hi = (float)((box[3] - s.lon0) * s.inv_dlon);
    s.fa_c0 = -s.lat0 * s.inv_dlat;
    s.fo_c0 = -s.lon0 * s.inv_dlon;
    s.ks_lo = m->tc.kv_lo > 0 ? m->tc.kv_lo : 0;
    s.ks_hi = m->tc.kv_hi < nT - 2 ? m->tc.kv_hi : nT - 2;
    finish_fast_view(s, m->tc, box);
    s.tmax = sim_times[nT - 1];
    BoatParams &p = m->bp;
    std::memcpy(p.polar, kFit, sizeof kFit);
    p.wind_max = 19.1; p.pofsail_min = 35.0; p.pofsail_max = 160.0; p.sigma_coeff = sigma;
    p.dest_angle = 0.005 * kDegToRad; p.earth_radius = 6371e3;
    p.cos_pmin = std::cos(kPi * 35.0 / 180.0);
    p.cos_pmax = std::cos(kPi * 160.0 / 180.0);
    derive_fast_params(p);
    default_fast_margins(p);
    return m;
}

void fm_destroy(void *h) { delete (Model *)h; }

void fm_force_generic(void *h, int on) { ((Model *)h)->force_generic = on != 0; }
int fm_last_variant(void *h) { return ((Model *)h)->last_variant; }

void fm_set_margins(void *h, float angle, float along, float near_miss2, float box, double bin) {
    BoatParams &p = ((Model *)h)->bp;
    p.margin_angle = angle; p.margin_along = along; p.near_miss2 = near_miss2; p.margin_box = box; p.margin_bin = bin;
}

// Boat.getDeterDyn through the fast path, for unit checks: wind (u, v), heading degrees.
float fm_boat_speed(void *h, float u, float v, double heading_deg) {
    Model *m = (Model *)h;
    float sh, ch;
    fast::heading_unit(heading_deg, sh, ch);
    return fast::boat_speed(m->bp, u, v, sh, ch);
}

void fm_rollout_batch(void *h, int64_t n_leaf, int replicas, const double root[3], const double dest[2], double tmin,
                      const double *prefix, const int32_t *prefix_len, int prefix_stride, const double *z,
                      uint64_t seed, uint64_t stream, uint64_t id0, int noise_mode, int flags, double *reward,
                      double *t_arr, int32_t *steps, uint8_t *status, int32_t *bin, double *frac, double *final_lat,
                      double *final_lon, double *traj_lat, double *traj_lon, int traj_stride, uint8_t *uncertain) {
    Model *m = (Model *)h;
    RolloutArgs a{};
    a.n_leaf = n_leaf; a.replicas = replicas;
    a.root_k = root[0]; a.root_lat = root[1]; a.root_lon = root[2];
    a.dest_lat = dest[0]; a.dest_lon = dest[1]; a.tmin = tmin;
    a.prefix = prefix; a.prefix_len = prefix_len; a.prefix_stride = prefix_stride;
    a.traj_lat = traj_lat; a.traj_lon = traj_lon; a.traj_stride = traj_stride;
    a.flags = flags;
    fast::prepare_rollout(m->sc, a);
    NoiseView nz{};
    nz.mode = noise_mode; nz.z = z; nz.seed = seed; nz.stream = stream; nz.id0 = id0;
    const int64_t n = n_leaf * replicas;
    const bool defboat = fast::is_default_boat(m->bp) && !m->force_generic;
    const bool turbo = defboat && fast::turbo_ok(m->sc, a);
    m->last_variant = turbo ? 2 : defboat ? 1 : 0;
    for (int64_t r = 0; r < n; ++r) {
        fast::RollOut o;
        // the same variant selection as rollout_impl (pmcts_b200.cu)
        if (turbo && !(flags & PMCTS_ROLL_REPLAY_FULL_PREFIX))
            fast::rollout_one<-1, false, true, true, true>(m->sc, m->bp, nz, a, n, r, r / replicas, o);
        else if (turbo) fast::rollout_one<-1, false, true, true>(m->sc, m->bp, nz, a, n, r, r / replicas, o);
        else if (defboat) fast::rollout_one<-1, false, true, false>(m->sc, m->bp, nz, a, n, r, r / replicas, o);
        else fast::rollout_one<-1, false>(m->sc, m->bp, nz, a, n, r, r / replicas, o);
        if (reward) reward[r] = o.rw;
        if (t_arr) t_arr[r] = o.ta;
        if (steps) steps[r] = o.nstep;
        if (status) status[r] = (uint8_t)o.st;
        if (bin) bin[r] = o.bin;
        if (frac) frac[r] = (o.st == PMCTS_ST_ARRIVED) ? o.frac : NAN;
        if (final_lat) final_lat[r] = o.lat;
        if (final_lon) final_lon[r] = o.lon;
        if (uncertain) uncertain[r] = (uint8_t)o.why;
    }
}

// K1 through the fast path: same contract as pmcts_step_batch on host arrays; needs_literal[i] is set
// (and the boat frozen with status 0) where the fast path hands over to the fp64 kernel.
void fm_step_batch(void *h, int64_t n, int nsteps, int32_t *k, double *lat, double *lon, const double *heading,
                   int seg_len, const double *z, uint64_t seed, uint64_t stream, uint64_t id0, int noise_mode,
                   uint8_t *status, double *traj_lat, double *traj_lon, uint8_t *needs_literal) {
    Model *m = (Model *)h;
    NoiseView nz{};
    nz.mode = noise_mode; nz.z = z; nz.seed = seed; nz.stream = stream; nz.id0 = id0;
    for (int64_t i = 0; i < n; ++i) {
        int kk = k[i];
        double la = lat[i], lo = lon[i];
        int st = status ? status[i] : 0;
        fast::Noise ns;
        float sh = 0.f, ch = 1.f;
        bool hok = true;
        if (needs_literal) needs_literal[i] = 0;
        float s = 0.f, c = 1.f;
        for (int j = 0; j < nsteps && st == PMCTS_ST_OK; ++j) {
            if (j % seg_len == 0) hok = fast::heading_unit(heading[(size_t)(j / seg_len) * n + i], sh, ch);
            if (!hok) { if (needs_literal) needs_literal[i] = 1; break; }
            if (j % 16 == 0) fast::sincos_lat(la, s, c);  // as step_fast_kernel: refreshed every 16 steps
            const fast::GridPos g = fast::grid_pos(m->sc, la, lo);
            const float zz = ns.draw<-1>(nz, n, i, j);
            fast::Move mv;
            st = fast::fast_step(m->sc, m->bp, kk, la, lo, g, sh, ch, s, c, zz, mv);
            if (st == fast::kStNeedsLiteral) { st = 0; if (needs_literal) needs_literal[i] = 1; break; }
            if (st == PMCTS_ST_OK) {
                s = mv.s_new;
                c = mv.c_new;
                if (traj_lat) traj_lat[(size_t)j * n + i] = la;
                if (traj_lon) traj_lon[(size_t)j * n + i] = lo;
            }
        }
        k[i] = kk; lat[i] = la; lon[i] = lo;
        if (status) status[i] = (uint8_t)st;
    }
}

}  // extern "C"

// ---- unit probes of the building blocks ----------------------------------------------------------
extern "C" {
// Simulator.getDestination through the increments: d = distance / earth radius (rad).
void fm_move(float d, double heading_deg, double lat, double lon, double *olat, double *olon) {
    float sh, ch, s, c;
    fast::heading_unit(heading_deg, sh, ch);
    fast::sincos_lat(lat, s, c);
    const fast::Move m = fast::move_increments(d, sh, ch, s, c);
    *olat = fma((double)m.dlat, 180.0 / kPi, lat);
    *olon = fma((double)m.dlon, 180.0 / kPi, lon);
}
// Unit vector of the bearing from (lat, lon) to (dlat, dlon).
void fm_bearing(double lat, double lon, double dlat, double dlon, float *sh, float *ch) {
    float s, c;
    fast::sincos_lat(lat, s, c);
    const fast::DestRel r = fast::dest_rel(lat, lon, dlat, dlon);
    fast::bearing_unit(r, s, (float)std::sin(dlat * kDegToRad), (float)std::cos(dlat * kDegToRad), *sh, *ch);
}
}
