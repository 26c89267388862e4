// pmcts_b200 shared plain-data types: device views of a scenario, boat parameters, noise and launch
// arguments.  Included by the literal fp64 kernels (pmcts_device.cuh), the mixed-precision kernels
// (pmcts_fast.cuh) and the host-compiled numerics model of the latter (tests/fastmodel).
#pragma once
#include <stdint.h>

#include "../../include/pmcts_b200.h"

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define PMCTS_HD __host__ __device__ __forceinline__
#else
#define PMCTS_HD inline
struct double2 { double x, y; };
struct float2 { float x, y; };
#endif

namespace pmcts {

constexpr double kPi = 3.141592653589793;
constexpr double kDegToRad = kPi / 180.0;

// Device view of one uploaded scenario (see pmcts_scenario_upload): one time-blended (u, v) slab per
// simulator instant, so a transition gathers 4 grid corners instead of 8 and needs no time weight.
struct ScenarioView {
    const double2 *slab64;  // [nT][nlat][nlon] (u, v) fp64
    const float2 *slab32;   // same, rounded to fp32 (Fast mode); slab k starts at slab32 + k * slab32_stride
    const double *times;    // [nT] simulator instants, days
    const double *dt_sec;   // [nT] (times[k+1]-times[k])*86400, last entry unused
    const uint8_t *k_valid; // [nT] 1 if times[k] lies inside the weather time axis
    int nT, nlat, nlon;
    double lat0, lon0, inv_dlat, inv_dlon, lat_last, lon_last;
    double box_lat_lo, box_lat_hi, box_lon_lo, box_lon_hi;
    double tmax;
    // ---- mixed-precision ("fast") view ----
    const float *dt_sec32;  // [nT] dt_sec rounded to fp32
    int kv_lo, kv_hi;       // k_valid[k] == 1 exactly for kv_lo <= k <= kv_hi
    int fast_ok;            // 0: this scenario must run on the literal fp64 kernels
    float max_abs_lat;      // max(|lat0|, |lat_last|), degrees
    double dt_max;          // max_k dt_sec[k]
    float box_fa_lo, box_fa_hi, box_fo_lo, box_fo_hi;  // terminal box in grid-index units
    double fa_c0, fo_c0;    // -lat0 * inv_dlat, -lon0 * inv_dlon: fractional index = fma(lat, inv_dlat, fa_c0)
    int ks_lo, ks_hi;       // instants a transition can start from: max(0, kv_lo) .. min(nT - 2, kv_hi)
    int slab32_stride;      // cells between fp32 slabs: nlat * nlon rounded up to even, so every slab is 16-byte
                            // aligned (TMA bulk copies into shared memory need that)
    float dt_uniform32;     // dt_sec32[k] when it is the same for every k, else 0
    float box_margin_scale; // 1 + 0.01 * max(nlat, nlon): scales BoatParams::margin_box to the fp32 spacing of the
                            // largest fractional grid index
    double inv_dlat_rad, inv_dlon_rad;  // inv_dlat, inv_dlon per radian (the turbo rollout keeps its position in radians)
    int roll_safe;          // 1: every instant 0 .. nT - 2 can start a step and the terminal box lies inside the
                            // grid, so a position that passed Tree.is_state_terminal needs no further range check
};

// Boat.* / module constants (pmcts_set_params), plus values derived once on the host.
struct BoatParams {
    double polar[36];
    double wind_max, pofsail_min, pofsail_max, sigma_coeff;
    double dest_angle, earth_radius;
    double cos_pmin, cos_pmax;  // cos(pi * POLAR_{MIN,MAX}_POFSAIL / 180)
    // ---- mixed-precision ("fast") copy, derived in pmcts_set_params ----
    // The same polynomial re-expanded about the centre of its domain, P(p, w) = sum pc[i][j] x^i y^j
    // with x = (p - p_mid) * p_ihalf in [-1, 1] over [POLAR_MIN, POLAR_MAX] and y = w * w_ihalf - 1
    // in [-1, 1] over [0, wind_max]: O(1) coefficients, so an fp32 Horner loses no digits to
    // cancellation (the raw monomial form cancels from ~30 down to ~1 m/s).
    float pc[36];
    int polar_tri;  // 1 if pc[i][j] == 0 for i + j > 5 (the reference's fit): 20-FMA triangular Horner
    float p_mid, p_ihalf, w_ihalf;
    float wind_max_f, pmin_f, pmax_f, cos_pmin_f, cos_pmax_f, sigma_f, inv_radius_f;
    float sin2_dest_angle;  // sin(DESTINATION_ANGLE)^2
    float speed_bound;      // upper bound of |P| over the domain
    // decision margins of the mixed-precision path: a test closer than its margin is not decided in
    // fp32, the rollout is re-run by the literal kernel (pmcts_fast.cuh)
    float margin_angle;  // relative, on sin^2(angle to the plane) vs sin^2(DESTINATION_ANGLE)
    float margin_along;  // position of D along A -> B, in units of |B - A|
    float near_miss2;    // (|D - B| / |B - A|)^2 below which the next bearing is ill-conditioned
    float margin_box;    // terminal box edges, grid-index units
    double margin_bin;   // reward * 8 against the integer bin edges
};

struct NoiseView {
    int mode;
    const double *z;
    uint64_t seed, stream, id0;
    // Philox4x32's two multipliers, passed as values rather than written as literals: with the multiplier in a
    // (uniform) register the 32 x 32 -> 64 product is one IMAD.WIDE; as an immediate with the top bit set the
    // compiler emits the product plus a correction add (3 extra instructions per round)
    uint32_t m0 = 0xD2511F53u, m1 = 0xCD9E8D57u;
};

// What every rollout of a launch derives from the root state and the destination (fast::prepare_root, on
// the host with the device's own formulas): sin / cos(root lat), the destination seen from the root
// (fast::DestRel) and the root's place on the wind grid (fast::GridPos).
struct RootPre {
    float s, c;
    float sin_dlat, sin_dlon, h_dlon, x_dlat;
    float wa, wo, fa, fo;
    int ia, io, inside;
};

// Arguments of one rollout launch (pmcts_rollout_batch).
struct RolloutArgs {
    int64_t n_leaf;
    int replicas;
    double root_k, root_lat, root_lon, dest_lat, dest_lon, tmin;
    const double *prefix;
    const int32_t *prefix_len;
    int prefix_stride;
    double *reward, *t_arr;
    int32_t *steps;
    uint8_t *status;
    int32_t *bin;
    double *frac, *final_lat, *final_lon, *traj_lat, *traj_lon;
    int traj_stride;
    uint32_t *leaf_bins;
    double *stats;
    int flags;
    // mixed-precision path: destination trigonometry (host, fp64 -> fp32) and the work list of
    // rollouts whose decisions fell inside the error margins (re-run by the literal fp64 kernel)
    float sin_dest_lat, cos_dest_lat;
    double root_lat_rad, root_lon_rad, dest_lat_rad, dest_lon_rad;  // the same positions in radians
    double fa_c0_rel, fo_c0_rel;  // grid index of the destination: a position held relative to it maps with these offsets
    double rw_t0;  // 1.001 * TimeMax
    float rw_inv;  // 1 / (1.001 * TimeMax - TimeMin)
    RootPre root_pre;
    uint32_t *redo_list;
    uint32_t *redo_count;
    const uint32_t *work_list;   // literal kernel: rollout ids to run (NULL = all)
    const uint32_t *work_count;
};

}  // namespace pmcts
