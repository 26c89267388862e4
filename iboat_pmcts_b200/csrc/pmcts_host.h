// pmcts_b200 host-side derivations shared by the C ABI (pmcts_b200.cu) and the host-compiled
// numerics model of the mixed-precision kernels (tests/fastmodel): parameters derived once per
// pmcts_set_params call and the per-instant time cells of pmcts_scenario_upload.
#pragma once
#include <cmath>
#include <vector>

#include "pmcts_types.h"

namespace pmcts {

// Fills the mixed-precision members of BoatParams from its fp64 members.
inline void derive_fast_params(BoatParams &p) {
    const long double pm = 0.5L * ((long double)p.pofsail_min + p.pofsail_max);
    const long double ph = 0.5L * ((long double)p.pofsail_max - p.pofsail_min);
    const long double wh = 0.5L * (long double)p.wind_max;
    long double binom[6][6] = {};
    for (int i = 0; i < 6; ++i) {
        binom[i][0] = 1;
        for (int a = 1; a <= i; ++a) binom[i][a] = binom[i - 1][a - 1] + (a <= i - 1 ? binom[i - 1][a] : 0);
    }
    long double bound = 0;
    bool tri = true;
    for (int a = 0; a < 6; ++a)
        for (int b = 0; b < 6; ++b) {
            long double d = 0;
            for (int i = a; i < 6; ++i)
                for (int j = b; j < 6; ++j)
                    d += (long double)p.polar[i * 6 + j] * binom[i][a] * powl(pm, i - a) * powl(ph, a) *
                         powl(wh, j) * binom[j][b];
            p.pc[a * 6 + b] = (float)d;
            bound += fabsl(d);
            if (a + b > 5 && p.pc[a * 6 + b] != 0.0f) tri = false;
        }
    p.polar_tri = tri ? 1 : 0;
    p.p_mid = (float)pm;
    p.p_ihalf = (float)(1.0L / ph);
    p.w_ihalf = (float)(1.0L / wh);
    p.wind_max_f = (float)p.wind_max;
    p.pmin_f = (float)p.pofsail_min;
    p.pmax_f = (float)p.pofsail_max;
    p.cos_pmin_f = (float)p.cos_pmin;
    p.cos_pmax_f = (float)p.cos_pmax;
    p.sigma_f = (float)p.sigma_coeff;
    p.inv_radius_f = (float)(1.0 / p.earth_radius);
    const double sa = std::sin(p.dest_angle);
    p.sin2_dest_angle = (float)(sa * sa);
    p.speed_bound = (float)bound;
}

// Default decision margins (calibrated in tests/test_fastmodel.py: >= 30x the largest error seen).
inline void default_fast_margins(BoatParams &p);  // (defined in pmcts_fast.cuh's terms below)

// Time cell / weight of every simulator instant on the weather time axis (SciPy find_indices
// semantics: i = clip(searchsorted_right(t) - 1, 0, nt - 2)), and the validity range of k.
struct TimeCells {
    std::vector<int> it;
    std::vector<double> wt, dt;
    std::vector<float> dt32;
    std::vector<uint8_t> valid;
    int kv_lo = 0, kv_hi = -1;
    bool contiguous = true;  // the valid instants form one run (always, for an ascending simulator axis)
    double dt_max = 0.0;
};

inline TimeCells time_cells(const double *wtime, int nt, const double *sim_times, int nT) {
    TimeCells c;
    c.it.resize(nT);
    c.wt.resize(nT);
    c.dt.assign(nT, 0.0);
    c.dt32.assign(nT, 0.0f);
    c.valid.resize(nT);
    bool seen = false;
    for (int k = 0; k < nT; ++k) {
        const double t = sim_times[k];
        c.valid[k] = (t >= wtime[0] && t <= wtime[nt - 1]) ? 1 : 0;
        int lo = 0, hi = nt;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (wtime[mid] <= t) lo = mid + 1; else hi = mid;
        }
        int i = lo - 1;
        if (i < 0) i = 0;
        if (i > nt - 2) i = nt - 2;
        c.it[k] = i;
        c.wt[k] = c.valid[k] ? (t - wtime[i]) / (wtime[i + 1] - wtime[i]) : 0.0;
        if (k + 1 < nT) {
            c.dt[k] = (sim_times[k + 1] - sim_times[k]) * 86400.0;
            c.dt32[k] = (float)c.dt[k];
            if (std::fabs(c.dt[k]) > c.dt_max) c.dt_max = std::fabs(c.dt[k]);
        }
        if (c.valid[k]) {
            if (!seen) c.kv_lo = k;
            c.kv_hi = k;
            seen = true;
        }
    }
    for (int k = c.kv_lo; k <= c.kv_hi; ++k)
        if (!c.valid[k]) c.contiguous = false;
    return c;
}

// Members of the mixed-precision view that follow from the time cells and the terminal box (after ks_lo / ks_hi,
// the axes and nT have been set).
inline void finish_fast_view(ScenarioView &vw, const TimeCells &tc, const double box[4]) {
    const int nT = vw.nT;
    // (the box ends short of the last grid node -- create_simulators' does by a whole cell -- so a position inside it
    // lies in a cell that has an upper neighbour: no clamp of the cell index)
    const double eps_lat = 1e-3 / vw.inv_dlat, eps_lon = 1e-3 / vw.inv_dlon;
    vw.roll_safe = (vw.ks_lo == 0 && vw.ks_hi == nT - 2 && box && box[0] >= vw.lat0 && box[1] <= vw.lat_last - eps_lat &&
                    box[2] >= vw.lon0 && box[3] <= vw.lon_last - eps_lon) ? 1 : 0;
    vw.inv_dlat_rad = vw.inv_dlat * (180.0 / kPi);
    vw.inv_dlon_rad = vw.inv_dlon * (180.0 / kPi);
    vw.dt_uniform32 = nT > 1 ? tc.dt32[0] : 0.0f;
    for (int k = 1; k + 1 < nT; ++k)
        if (tc.dt32[k] != tc.dt32[0]) vw.dt_uniform32 = 0.0f;
    vw.box_margin_scale = 1.0f + 0.01f * (float)(vw.nlat > vw.nlon ? vw.nlat : vw.nlon);
}

}  // namespace pmcts
