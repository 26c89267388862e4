// pmcts_b200 -- kernels and C ABI (include/pmcts_b200.h) of the B200-native rollout engine.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
//        (see __graft_entry__.build()).  Depends on the CUDA runtime only.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "pmcts_device.cuh"
#include "pmcts_fast.cuh"
#include "pmcts_fast2.cuh"
#include "pmcts_host.h"

namespace pmcts {

// ================================================================================================
// Kernels
// ================================================================================================

constexpr int kBlock = 128;

// Scenario upload: blends the weather grid in time once per simulator instant.
//   slab[k][a][o] = (1 - wt[k]) * W[it[k]][a][o] + wt[k] * W[it[k] + 1][a][o]
// SciPy applies the time weight inside the 8-corner sum (_rgi.py:520-549); blending first is the
// same polynomial and differs only in rounding (~1e-16 relative).
template <typename T>
__global__ void blend_time_slabs_kernel(const T *__restrict__ u, const T *__restrict__ v,
                                        const int *__restrict__ it, const double *__restrict__ wt,
                                        int nT, int plane, int stride32, double2 *__restrict__ slab64,
                                        float2 *__restrict__ slab32) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)nT * plane) return;
    const int k = (int)(idx / plane), p = (int)(idx % plane);
    const size_t lo = (size_t)it[k] * plane + p, hi = lo + plane;
    const double w = wt[k];
    const double uu = (double)u[lo] * (1.0 - w) + (double)u[hi] * w;
    const double vv = (double)v[lo] * (1.0 - w) + (double)v[hi] * w;
    slab64[idx] = make_double2(uu, vv);
    slab32[(size_t)k * stride32 + p] = make_float2((float)uu, (float)vv);
}

// Weather.uInterpolator / vInterpolator (weatherTLKT.py:369-378): SciPy's linear RegularGridInterpolator
// at arbitrary (t, lat, lon), on the raw grid kept from the upload.  Evaluation order of
// scipy/interpolate/_rgi.py:520-549 -- corners in product order (time slowest, lon fastest), weight
// ((1 * wt) * wlat) * wlon, accumulated from 0 -- with explicitly unfused multiplies and adds, so the
// result equals SciPy's bit for bit.
__device__ __forceinline__ int axis_cell(const double *__restrict__ g, int n, double x, double &w) {
    int lo = 0, hi = n;  // first index with g[idx] > x
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (g[mid] <= x) lo = mid + 1; else hi = mid;
    }
    int i = min(max(lo - 1, 0), n - 2);
    w = __ddiv_rn(__dsub_rn(x, g[i]), __dsub_rn(g[i + 1], g[i]));
    return i;
}

template <typename T>
__global__ void interp_wind_kernel(const T *__restrict__ u, const T *__restrict__ v, const double *__restrict__ gt,
                                   int nt, const double *__restrict__ ga, int nlat, const double *__restrict__ go,
                                   int nlon, int64_t n, const double *__restrict__ t, const double *__restrict__ lat,
                                   const double *__restrict__ lon, double *__restrict__ ou, double *__restrict__ ov,
                                   uint8_t *__restrict__ status) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double tt = t[i], la = lat[i], lo = lon[i];
    const bool ok = gt[0] <= tt && tt <= gt[nt - 1] && ga[0] <= la && la <= ga[nlat - 1] && go[0] <= lo &&
                    lo <= go[nlon - 1];
    double su = NAN, sv = NAN;
    if (ok) {
        double w[3][2];
        double y;
        const int it = axis_cell(gt, nt, tt, y);
        w[0][0] = __dsub_rn(1.0, y); w[0][1] = y;
        const int ia = axis_cell(ga, nlat, la, y);
        w[1][0] = __dsub_rn(1.0, y); w[1][1] = y;
        const int io = axis_cell(go, nlon, lo, y);
        w[2][0] = __dsub_rn(1.0, y); w[2][1] = y;
        su = 0.0;
        sv = 0.0;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const double weight = __dmul_rn(__dmul_rn(__dmul_rn(1.0, w[0][a]), w[1][b]), w[2][c]);
                    const size_t idx = ((size_t)(it + a) * nlat + (ia + b)) * nlon + (io + c);
                    su = __dadd_rn(su, __dmul_rn((double)u[idx], weight));
                    sv = __dadd_rn(sv, __dmul_rn((double)v[idx], weight));
                }
    }
    ou[i] = su;
    ov[i] = sv;
    if (status) status[i] = ok ? PMCTS_ST_OK : PMCTS_ST_OUT_OF_GRID;
}

__global__ void get_wind_kernel(ScenarioView sc, int64_t n, const int32_t *__restrict__ k,
                                const double *__restrict__ lat, const double *__restrict__ lon,
                                double *__restrict__ u, double *__restrict__ v,
                                uint8_t *__restrict__ status) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double uu = NAN, vv = NAN;
    const int kk = k[i];
    const bool ok = kk >= 0 && kk < sc.nT && wind_f64(sc, kk, lat[i], lon[i], uu, vv);
    u[i] = uu;
    v[i] = vv;
    if (status) status[i] = ok ? PMCTS_ST_OK : PMCTS_ST_OUT_OF_GRID;
}

// K1 -- Simulator.doStep for n boats, nsteps transitions fused in one launch.
// One thread per boat; state lives in registers across the steps; headings are re-read once per
// segment (coalesced, [segment][boat]).
__global__ void __launch_bounds__(kBlock)
step_batch_kernel(ScenarioView sc, BoatParams bp, int64_t n, int nsteps, int32_t *__restrict__ k_io,
                  double *__restrict__ lat_io, double *__restrict__ lon_io,
                  double *__restrict__ prev_lat, double *__restrict__ prev_lon,
                  const double *__restrict__ heading, int seg_len, NoiseView nz,
                  uint8_t *__restrict__ status_io, double *__restrict__ traj_lat,
                  double *__restrict__ traj_lon, int flags, const int32_t *__restrict__ j_resume,
                  const uint32_t *__restrict__ any_resume) {
    // Resume mode (j_resume != NULL): only the boats the mixed-precision kernel handed over
    // (status kStNeedsLiteral) run, each from the step it stopped at.
    if (any_resume && *any_resume == 0) return;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int st = status_io ? status_io[i] : PMCTS_ST_OK;
    int j0 = 0;
    if (j_resume) {
        if (st != fast::kStNeedsLiteral) return;
        st = PMCTS_ST_OK;
        j0 = j_resume[i];
    }
    int k = k_io[i];
    double lat = lat_io[i], lon = lon_io[i];
    double pla = prev_lat ? prev_lat[i] : lat, plo = prev_lon ? prev_lon[i] : lon;
    NoiseStream ns;
    double h = 0.0;
    for (int j = j0; j < nsteps && st == PMCTS_ST_OK; ++j) {
        if (j % seg_len == 0 || j == j0) h = heading[(size_t)(j / seg_len) * n + i];
        if (flags & PMCTS_STEP_STOP_BEFORE_TERMINAL) {  // forest.py:239
            if (k + 1 >= sc.nT) { st = PMCTS_ST_PAST_HORIZON; break; }
            if (is_terminal(sc, k + 1, lat, lon)) {
                st = (k + 1 >= sc.nT - 1) ? PMCTS_ST_HORIZON : PMCTS_ST_OUT_OF_BOX;
                break;
            }
        }
        const double z = ns.draw<false>(nz, n, i, j);
        const double la0 = lat, lo0 = lon;
        st = do_step_literal(sc, bp, k, lat, lon, h, z);
        if (st == PMCTS_ST_OK) {
            pla = la0;
            plo = lo0;
            if (traj_lat) traj_lat[(size_t)j * n + i] = lat;
            if (traj_lon) traj_lon[(size_t)j * n + i] = lon;
        }
    }
    k_io[i] = k;
    lat_io[i] = lat;
    lon_io[i] = lon;
    if (prev_lat) prev_lat[i] = pla;
    if (prev_lon) prev_lon[i] = plo;
    if (status_io) status_io[i] = (uint8_t)st;
}

// Several scenarios in one launch (pmcts_rollout_forest): the scenarios share their geometry (axes,
// box, simulator instants); only the wind slabs differ.  select_scenario turns the launch-wide
// arguments into those of scenario s: slabs, noise stream, and every per-scenario array offset.
constexpr int kMaxForest = 16;
// Statistics are accumulated into kStatsCopies sets of partial sums per scenario (picked by block index)
// and folded into the caller's stats[8] by stats_finish_kernel: a single set of six addresses hit by every
// warp of the grid serialises in one L2 slice and stalls every other access that maps to it.
constexpr int kStatsCopies = 64;
struct ForestSlabs {
    const float2 *s32[kMaxForest];
    const double2 *s64[kMaxForest];
    int n_scen, prefix_shared;
};

template <bool kLean = false>
__device__ __forceinline__ void select_scenario(int s, int64_t n, const ForestSlabs &fs, ScenarioView &sc,
                                                NoiseView &nz, RolloutArgs &a) {
    sc.slab32 = fs.s32[s];
    sc.slab64 = fs.s64[s];
    nz.stream += (uint64_t)s;
    if (nz.z) nz.z += (size_t)s * sc.nT * n;
    const size_t on = (size_t)s * n, ol = (size_t)s * a.n_leaf;
    if (!fs.prefix_shared) {
        if (a.prefix) a.prefix += ol * a.prefix_stride;
        if (a.prefix_len) a.prefix_len += ol;
    }
    if (a.leaf_bins) a.leaf_bins += ol * 12;
    if (a.stats) a.stats += (size_t)s * (kStatsCopies * 8);  // a.stats is the launch's partial-sum area
    if (kLean) return;  // a lean launch has no per-rollout outputs
    if (a.reward) a.reward += on;
    if (a.t_arr) a.t_arr += on;
    if (a.steps) a.steps += on;
    if (a.status) a.status += on;
    if (a.bin) a.bin += on;
    if (a.frac) a.frac += on;
    if (a.final_lat) a.final_lat += on;
    if (a.final_lon) a.final_lon += on;
    if (a.traj_lat) a.traj_lat += on * a.traj_stride;
    if (a.traj_lon) a.traj_lon += on * a.traj_stride;
}

// Work-list entry of the mixed-precision kernels: scenario index in the top 4 bits, rollout id below.
constexpr int kRedoShift = 28;

// Per-leaf reduction of the reward bins and the plan-evaluation statistics, shared by the literal and
// the mixed-precision rollout kernels.  Lanes holding the same (leaf, bin) find each other with one
// MATCH; the lowest lane of each group adds the group's size to leaf_bins[leaf][bin].
// `a` is the lane's own (scenario-offset) argument block; with kUniform every lane of the warp belongs
// to the same scenario (the stats of a warp go to one place), otherwise lanes add their own.
// Every lane of the warp calls this (converged) -- lanes without a rollout pass counted = false -- so the warp
// primitives run with the full mask, as PTX requires of every lane that executes them.
template <bool kUniform>
__device__ __forceinline__ void reduce_rollouts(const RolloutArgs &a, bool counted, int scen,
                                                int64_t leaf, int bin, int st, int nstep, double ta) {
    constexpr unsigned warp_mask = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    if (a.leaf_bins) {
        const unsigned long long key =
            counted ? ((unsigned long long)(scen * a.n_leaf + leaf) << 4) | (unsigned)bin : ~0ull - (unsigned)lane;
        const unsigned peers = __match_any_sync(warp_mask, key);
        if (counted && lane == __ffs(peers) - 1)
            atomicAdd(&a.leaf_bins[leaf * 12 + bin], (uint32_t)__popc(peers));
    }
    double *stats = a.stats ? a.stats + ((blockIdx.x + (threadIdx.x >> 5)) & (kStatsCopies - 1)) * 8 : nullptr;
    if (stats && !kUniform) {
        if (counted) {
            if (!isnan(ta)) {
                atomicAdd(&stats[0], 1.0);
                atomicAdd(&stats[1], ta);
                atomicAdd(&stats[2], ta * ta);
            }
            if (st == PMCTS_ST_ARRIVED) atomicAdd(&stats[3], 1.0);
            atomicAdd(&stats[4], (double)nstep);
            atomicAdd(&stats[5], 1.0);
        }
    } else if (stats) {
        const bool has_t = counted && !isnan(ta);
        const int c_t = __reduce_add_sync(warp_mask, has_t ? 1 : 0);
        const int c_arr = __reduce_add_sync(warp_mask, (counted && st == PMCTS_ST_ARRIVED) ? 1 : 0);
        const int c_steps = __reduce_add_sync(warp_mask, counted ? nstep : 0);
        const int c_roll = __reduce_add_sync(warp_mask, counted ? 1 : 0);
        double s1 = has_t ? ta : 0.0, s2 = has_t ? ta * ta : 0.0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            s1 += __shfl_xor_sync(warp_mask, s1, off);
            s2 += __shfl_xor_sync(warp_mask, s2, off);
        }
        if (lane == 0) {
            if (c_t) {
                atomicAdd(&stats[0], (double)c_t);
                atomicAdd(&stats[1], s1);
                atomicAdd(&stats[2], s2);
            }
            if (c_arr) atomicAdd(&stats[3], (double)c_arr);
            atomicAdd(&stats[4], (double)c_steps);
            atomicAdd(&stats[5], (double)c_roll);
        }
    }
}

// Folds the partial sums of a launch into the caller's stats[n_scen][8] (always adds: the caller's array
// was zeroed first unless PMCTS_ROLL_ACCUMULATE).  One block per scenario, one thread per copy.
__global__ void stats_finish_kernel(const double *__restrict__ part, double *__restrict__ stats) {
    const int s = blockIdx.x, c = threadIdx.x;
    const double *p = part + ((size_t)s * kStatsCopies + c) * 8;
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        double v = p[q];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if ((c & 31) == 0 && v != 0.0) atomicAdd(&stats[s * 8 + q], v);
    }
}

// K2 / K4 / K5, literal fp64 -- one default-policy rollout per thread (worker.py:255-300), reward
// binned and reduced per leaf, optional (count, sum t, sum t^2, arrived) reduction for plan
// evaluation (isochrones.py:537-550).  With a work list (a.work_list) the kernel runs only the listed
// rollouts, grid-striding over *a.work_count: that is how the rollouts the mixed-precision kernel
// could not decide are finished.
__global__ void __launch_bounds__(kBlock)
rollout_batch_kernel(ScenarioView sc0, BoatParams bp, NoiseView nz0, RolloutArgs a0, ForestSlabs fs) {
    const int64_t n = a0.n_leaf * a0.replicas;
    const int64_t total = a0.work_list ? (int64_t)*a0.work_count : n;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < total; base += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = base + threadIdx.x;
    const bool live = t < total;
    int64_t r = live ? t : 0;
    int scen = blockIdx.y;
    if (live && a0.work_list) {
        const uint32_t e = a0.work_list[t];
        scen = (int)(e >> kRedoShift);
        r = (int64_t)(e & ((1u << kRedoShift) - 1));
    }
    ScenarioView sc = sc0;
    NoiseView nz = nz0;
    RolloutArgs a = a0;
    select_scenario(scen, n, fs, sc, nz, a);
    int bin = 0, st = PMCTS_ST_OK, nstep_out = 0;
    int64_t leaf = 0;
    double ta = NAN, rw = 0.0;
    if (live) {
        leaf = r / a.replicas;
        int k = (int)a.root_k;
        double alat = a.root_lat, alon = a.root_lon, blat = a.root_lat, blon = a.root_lon;
        double xd, yd, zd;
        geo_to_cart(a.dest_lat, a.dest_lon, xd, yd, zd);
        int nstep = 0, at = 0;
        double frac = 0.0, brg = 0.0;
        NoiseStream ns;
        const double *pre = a.prefix ? a.prefix + (size_t)leaf * a.prefix_stride : nullptr;
        const int plen = a.prefix_len ? a.prefix_len[leaf] : 0;
        // sin / cos of the current position B and of the one before it, A: taken once per position (PosTrig); the
        // destination's latitude once per rollout
        PosTrig tb, ta_;
        trig_lat(blat, tb);
        ta_ = tb;
        double sD, cD;
        sincos(a.dest_lat * kDegToRad, &sD, &cD);

        auto step = [&](double action) {
            const double z = ns.draw<false>(nz, n, r, nstep);
            const double la0 = blat, lo0 = blon;
            st = do_step_literal(sc, bp, k, blat, blon, action, z, &tb);
            if (st == PMCTS_ST_OK) {
                alat = la0;
                alon = lo0;
                ta_ = tb;
                trig_lat(blat, tb);
                if (a.traj_lat && nstep < a.traj_stride) a.traj_lat[(size_t)nstep * n + r] = blat;
                if (a.traj_lon && nstep < a.traj_stride) a.traj_lon[(size_t)nstep * n + r] = blon;
                ++nstep;
            }
        };
        auto at_dest = [&]() {
            double xa, ya, za, xb, yb, zb;
            trig_lon(alon, ta_);
            trig_lon(blon, tb);
            geo_to_cart(ta_, xa, ya, za);
            geo_to_cart(tb, xb, yb, zb);
            const int res = at_dest_cart(bp, xa, ya, za, xb, yb, zb, xd, yd, zd, frac);
            if (res < 0) st = PMCTS_ST_DEGENERATE; else at = res;
        };
        auto bearing = [&]() { return bearing_to(tb, blon, sD, cD, a.dest_lon); };

        if (!(a.flags & PMCTS_ROLL_PLAN_EVAL)) {
            // Tree.get_sim_to_estimate_state (worker.py:283-300): the first action always; the rest
            // only when asked to (the reference's guard never lets its loop run -- quirk Q1).
            if (plen > 0) step(pre[0]);
            if (a.flags & PMCTS_ROLL_REPLAY_FULL_PREFIX) {
                for (int j = 1; j < plen && st == PMCTS_ST_OK; ++j) {
                    if (is_terminal(sc, k, blat, blon)) break;
                    at_dest();
                    if (st != PMCTS_ST_OK || at) break;
                    step(pre[j]);
                }
            }
        } else {
            // estimate_perfomance_plan (isochrones.py:498-511): whole plan, unchecked.
            for (int j = 0; j < plen && st == PMCTS_ST_OK; ++j) step(pre[j]);
            if (plen == 0) step(bearing());
        }
        // Tree.default_policy (worker.py:264-271)
        if (st == PMCTS_ST_OK) {
            brg = bearing();
            at_dest();
            while (st == PMCTS_ST_OK && !at && !is_terminal(sc, k, blat, blon)) {
                step(brg);
                if (st != PMCTS_ST_OK) break;
                brg = bearing();
                at_dest();
            }
        }
        if (st == PMCTS_ST_OK) {
            if (at) {
                const double tk = sc.times[k];
                ta = tk - (1.0 - frac) * (tk - sc.times[k - 1]);  // worker.py:274-275
                rw = reward_of(sc.tmax, a.tmin, ta);
                st = PMCTS_ST_ARRIVED;
            } else {
                st = (k >= sc.nT - 1) ? PMCTS_ST_HORIZON : PMCTS_ST_OUT_OF_BOX;
                if (a.flags & PMCTS_ROLL_PLAN_EVAL) ta = sc.tmax;  // isochrones.py:527-530
            }
        }
        bin = hist_bin(rw);
        nstep_out = nstep;
        if (a.reward) a.reward[r] = rw;
        if (a.t_arr) a.t_arr[r] = ta;
        if (a.steps) a.steps[r] = nstep;
        if (a.status) a.status[r] = (uint8_t)st;
        if (a.bin) a.bin[r] = bin;
        if (a.frac) a.frac[r] = (st == PMCTS_ST_ARRIVED) ? frac : NAN;
        if (a.final_lat) a.final_lat[r] = blat;
        if (a.final_lon) a.final_lon[r] = blon;
    }
    if (a0.work_list) reduce_rollouts<false>(a, live, scen, leaf, bin, st, nstep_out, ta);
    else reduce_rollouts<true>(a, live, scen, leaf, bin, st, nstep_out, ta);
    }
}

// K2 / K4 / K5, mixed precision -- same contract; the arithmetic is pmcts_fast.cuh.  A rollout whose
// decisions fell inside the error margins writes nothing: its id goes to a.redo_list and the literal
// kernel above finishes it (same stream, next launch).  kNoise fixes the noise source at compile time;
// kLean is the throughput form: per-leaf histograms and statistics only, no per-rollout outputs.
template <int kNoise, bool kLean, bool kDefBoat = false, bool kTurbo = false, bool kFirstOnly = false>
// (turbo: a cap of 73 registers lets ptxas keep the pinned constants below; it uses 64, so 8 blocks per SM stay resident)
__global__ void __launch_bounds__(kBlock, kLean ? (kTurbo ? 7 : 8) : 6)  // lean: 64 registers, 8 blocks per SM
rollout_fast_kernel(ScenarioView sc, BoatParams bp, NoiseView nz, RolloutArgs a, ForestSlabs fs) {
    const int64_t n = a.n_leaf * a.replicas;
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int scen = blockIdx.y;
    select_scenario<kLean>(scen, n, fs, sc, nz, a);
    const bool live = r < n;
    fast::RollOut o;
    o.st = PMCTS_ST_OK; o.nstep = 0; o.bin = 0; o.why = 0; o.ta = NAN;
    bool counted = false;
    // n < 2^28 (the work-list entries hold the rollout id in 28 bits): a 32-bit division, done once
    const int64_t leaf = (int64_t)((uint32_t)r / (uint32_t)a.replicas);
    if (kTurbo) {
        // The loop's scalar constants as opaque register values: ptxas otherwise re-loads them from the constant bank
        // at every step (17 of the 316 instructions of a step).  Twelve registers; the kernel still fits 64.
#define PIN_I(x) asm volatile("" : "+r"(x))
#define PIN_F(x) asm volatile("" : "+f"(x))
        PIN_I(sc.nlon); PIN_I(sc.slab32_stride); PIN_I(sc.nT);
        PIN_F(sc.dt_uniform32); PIN_F(sc.box_margin_scale);
        PIN_F(sc.box_fa_lo); PIN_F(sc.box_fa_hi); PIN_F(sc.box_fo_lo); PIN_F(sc.box_fo_hi);
        PIN_F(bp.sigma_f); PIN_F(a.sin_dest_lat); PIN_F(a.cos_dest_lat);
        // (the scenario's slab pointer comes out of an indexed constant-bank load: once, not once per step)
        asm volatile("" : "+l"(sc.slab32));
#undef PIN_I
#undef PIN_F
    }
    if (live) {
        fast::rollout_one<kNoise, kLean, kDefBoat, kTurbo, kFirstOnly>(sc, bp, nz, a, n, r, leaf, o);
        if (o.why) {
            a.redo_list[atomicAdd(a.redo_count, 1u)] = ((uint32_t)scen << kRedoShift) | (uint32_t)r;
        } else {
            counted = true;
            if (!kLean) {
                if (a.reward) a.reward[r] = o.rw;
                if (a.t_arr) a.t_arr[r] = o.ta;
                if (a.steps) a.steps[r] = o.nstep;
                if (a.status) a.status[r] = (uint8_t)o.st;
                if (a.bin) a.bin[r] = o.bin;
                if (a.frac) a.frac[r] = (o.st == PMCTS_ST_ARRIVED) ? o.frac : NAN;
                if (a.final_lat) a.final_lat[r] = o.lat;
                if (a.final_lon) a.final_lon[r] = o.lon;
            }
        }
    }
    reduce_rollouts<true>(a, counted, scen, leaf, o.bin, o.st, o.nstep, o.ta);
}

// K2, turbo variant, two rollouts per thread on the packed FP32 pipe (pmcts_fast2.cuh): thread t carries the
// rollouts 2 t and 2 t + 1 of its scenario.  Same contract and same results as rollout_fast_kernel<kNoise, true, true,
// true>: histograms / statistics only, undecided rollouts queued for the fp64 kernel.
template <int kNoise, int kMinBlocks = 4>
__global__ void __launch_bounds__(kBlock, kMinBlocks)
rollout_fast2_kernel(ScenarioView sc, BoatParams bp, NoiseView nz, RolloutArgs a, ForestSlabs fs) {
    const int64_t n = a.n_leaf * a.replicas;
    const int64_t r0 = 2 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
    const int scen = blockIdx.y;
    select_scenario<true>(scen, n, fs, sc, nz, a);
    const bool live[2] = {r0 < n, r0 + 1 < n};
    const int64_t leaf[2] = {(int64_t)((uint32_t)r0 / (uint32_t)a.replicas),
                             (int64_t)((uint32_t)(r0 + 1) / (uint32_t)a.replicas)};
#define PIN_I(x) asm volatile("" : "+r"(x))
#define PIN_F(x) asm volatile("" : "+f"(x))
    PIN_I(sc.nlon); PIN_I(sc.slab32_stride); PIN_I(sc.nT);
    PIN_F(sc.dt_uniform32); PIN_F(sc.box_margin_scale);
    PIN_F(sc.box_fa_lo); PIN_F(sc.box_fa_hi); PIN_F(sc.box_fo_lo); PIN_F(sc.box_fo_hi);
    PIN_F(bp.sigma_f); PIN_F(a.sin_dest_lat); PIN_F(a.cos_dest_lat);
#undef PIN_I
#undef PIN_F
    fast2::Pair o;
    fast2::rollout_pair<kNoise>(sc, bp, nz, a, n, r0, leaf, live, o);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        // epilogue of fast::rollout_one, per half
        int st = o.st[h], why = o.why[h];
        double ta = NAN, rw = 0.0;
        if (live[h] && st == PMCTS_ST_OK && !why) {
            const int k = o.k[h];
            if (o.at[h]) {
                const double tk = sc.times[k];
                ta = tk - (1.0 - (double)o.frac[h]) * (tk - sc.times[k - 1]);  // worker.py:274-275
                rw = (double)((fast::exp_approx((float)(a.rw_t0 - ta) * a.rw_inv) - 1.0f) * 0.58197670686932642f);
                st = PMCTS_ST_ARRIVED;
            } else {
                st = (k >= sc.nT - 1) ? PMCTS_ST_HORIZON : PMCTS_ST_OUT_OF_BOX;
            }
        }
        const int bin = fast::hist_bin_checked(bp, rw, why);
        bool counted = false;
        if (live[h]) {
            if (why) a.redo_list[atomicAdd(a.redo_count, 1u)] = ((uint32_t)scen << kRedoShift) | (uint32_t)(r0 + h);
            else counted = true;
        }
        reduce_rollouts<true>(a, counted, scen, leaf[h], bin, st, o.nstep[h], ta);
    }
}

// K1, mixed precision -- Simulator.doStep for n boats, nsteps fused (same contract as
// step_batch_kernel).  A boat the fp32 formulas do not cover (heading outside [0, 360), a step longer
// than fast::kMaxStepAngle) stops with the internal status kStNeedsLiteral and its step index in
// j_resume; step_batch_kernel then finishes it.  kLean: no trajectory / previous-state output and no
// stop-before-terminal mode -- the raw sweep of BASELINE config 3.
using fast::kRefreshTrig;

template <int kNoise, bool kLean, bool kDefBoat = false>
__global__ void __launch_bounds__(kBlock)
step_fast_kernel(ScenarioView sc, BoatParams bp, int64_t n, int nsteps, int32_t *__restrict__ k_io,
                 double *__restrict__ lat_io, double *__restrict__ lon_io, double *__restrict__ prev_lat,
                 double *__restrict__ prev_lon, const double *__restrict__ heading, int seg_len, NoiseView nz,
                 uint8_t *__restrict__ status_io, double *__restrict__ traj_lat, double *__restrict__ traj_lon,
                 int flags, int32_t *__restrict__ j_resume, uint32_t *__restrict__ any_resume) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int k = k_io[i];
    double lat = lat_io[i], lon = lon_io[i];
    double pla = lat, plo = lon;
    if (!kLean) {
        if (prev_lat) pla = prev_lat[i];
        if (prev_lon) plo = prev_lon[i];
    }
    int st = status_io[i];
    fast::Noise ns;
    float sh = 0.0f, ch = 1.0f, s = 0.0f, c = 1.0f;
    int seg_left = 0, trig_left = 0;
    const double *hp = heading + i;
    for (int j = 0; j < nsteps && st == PMCTS_ST_OK; ++j) {
        if (seg_left == 0) {
            seg_left = seg_len;
            const bool ok = fast::heading_unit(*hp, sh, ch);
            hp += n;
            if (!ok) { st = fast::kStNeedsLiteral; j_resume[i] = j; break; }
        }
        --seg_left;
        if (!kLean && (flags & PMCTS_STEP_STOP_BEFORE_TERMINAL)) {  // forest.py:239
            if (k + 1 >= sc.nT) { st = PMCTS_ST_PAST_HORIZON; break; }
            if (is_terminal(sc, k + 1, lat, lon)) {
                st = (k + 1 >= sc.nT - 1) ? PMCTS_ST_HORIZON : PMCTS_ST_OUT_OF_BOX;
                break;
            }
        }
        // sin / cos(lat): exact from the latitude every kRefreshTrig steps, carried by the step's own
        // identities (s' = s cos d + c a, c' = H) in between -- they only scale increments, and the drift of
        // a few fp32 roundings is re-set before it matters
        if (trig_left == 0) {
            trig_left = kRefreshTrig;
            fast::sincos_lat(lat, s, c);
        }
        --trig_left;
        const fast::GridPos g = fast::grid_pos(sc, lat, lon);
        const float z = ns.draw<kNoise, true>(nz, n, i, j);
        const double la0 = lat, lo0 = lon;
        fast::Move m;
        st = fast::fast_step<false, kDefBoat, false, kDefBoat>(sc, bp, k, lat, lon, g, sh, ch, s, c, z, m);
        if (st == fast::kStNeedsLiteral) { j_resume[i] = j; break; }
        if (st == PMCTS_ST_OK) {
            s = m.s_new;
            c = m.c_new;
            if (!kLean) {
                pla = la0;
                plo = lo0;
                if (traj_lat) traj_lat[(size_t)j * n + i] = lat;
                if (traj_lon) traj_lon[(size_t)j * n + i] = lon;
            }
        }
    }
    if (st == fast::kStNeedsLiteral) *any_resume = 1u;
    k_io[i] = k;
    lat_io[i] = lat;
    lon_io[i] = lon;
    if (!kLean) {
        if (prev_lat) prev_lat[i] = pla;
        if (prev_lon) prev_lon[i] = plo;
    }
    status_io[i] = (uint8_t)st;
}

// ---- TMA bulk copy + mbarrier (PTX, sm_90+): global -> shared without passing through registers -------
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}

// K1, mixed precision, wind slab staged in shared memory -- the raw sweep when every running boat of a block
// stands at the same time index (BASELINE config 3: the whole fleet starts at k = 0).  At step j all of them
// interpolate on slab k0 + j: one thread brings that slab (nlat * nlon * 8 bytes; 5.2 KB on the 0.5-degree
// grid) into shared memory with ONE TMA bulk copy, two steps ahead, into a two-slot ring guarded by
// mbarriers; the four corner gathers of every boat are then shared-memory loads.  Blocks whose boats do not
// share a time index, and grids whose slab does not fit the ring, take the L1 / L2 path of step_fast_kernel.
template <int kNoise>
__global__ void __launch_bounds__(kBlock)
step_fast_staged_kernel(ScenarioView sc, BoatParams bp, int64_t n, int nsteps, int32_t *__restrict__ k_io,
                        double *__restrict__ lat_io, double *__restrict__ lon_io,
                        const double *__restrict__ heading, int seg_len, NoiseView nz,
                        uint8_t *__restrict__ status_io, int32_t *__restrict__ j_resume,
                        uint32_t *__restrict__ any_resume) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ int s_kb;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool mine = i < n;
    int k = 0, st = PMCTS_ST_HORIZON;
    double lat = 0.0, lon = 0.0;
    if (mine) {
        k = k_io[i];
        lat = lat_io[i];
        lon = lon_io[i];
        st = status_io[i];
    }
    // common time index of the running boats of this block
    if (threadIdx.x == 0) s_kb = 0x7fffffff;
    __syncthreads();
    if (st == PMCTS_ST_OK) atomicMin(&s_kb, k);
    __syncthreads();
    const int kb = s_kb;
    if (kb == 0x7fffffff) return;  // nobody runs
    const bool staged = __syncthreads_and(st != PMCTS_ST_OK || k == kb) != 0;

    fast::Noise ns;
    float sh = 0.0f, ch = 1.0f, s = 0.0f, c = 1.0f;
    int seg_left = 0, trig_left = 0;
    const double *hp = heading + (mine ? i : 0);

    // one transition of this thread's boat (lean form of step_fast_kernel's loop body)
    auto advance = [&](int j, const float2 *slab) {
        if (seg_left == 0) {
            seg_left = seg_len;
            const bool ok = fast::heading_unit(*hp, sh, ch);
            hp += n;
            if (!ok) { st = fast::kStNeedsLiteral; j_resume[i] = j; return; }
        }
        --seg_left;
        if (trig_left == 0) {
            trig_left = kRefreshTrig;
            fast::sincos_lat(lat, s, c);
        }
        --trig_left;
        const fast::GridPos g = fast::grid_pos(sc, lat, lon);
        const float z = ns.draw<kNoise, true>(nz, n, i, j);
        fast::Move m;
        st = slab ? fast::fast_step<true>(sc, bp, k, lat, lon, g, sh, ch, s, c, z, m, slab)
                  : fast::fast_step<false>(sc, bp, k, lat, lon, g, sh, ch, s, c, z, m);
        if (st == fast::kStNeedsLiteral) { j_resume[i] = j; return; }
        if (st == PMCTS_ST_OK) {
            s = m.s_new;
            c = m.c_new;
        }
    };

    if (!staged) {
        for (int j = 0; j < nsteps && st == PMCTS_ST_OK; ++j) advance(j, nullptr);
    } else {
        const int cells = sc.slab32_stride;
        const uint32_t bytes = (uint32_t)cells * sizeof(float2);
        float2 *const ring = reinterpret_cast<float2 *>(smem_raw);  // slot b at ring + b * cells
        auto issue = [&](int j) {  // slab of step j into slot j & 1 (thread 0 only)
            const int kj = kb + j;
            if (j < nsteps && kj >= 0 && kj < sc.nT) {
                mbar_expect_tx(&bars[j & 1], bytes);
                tma_bulk_load(ring + (j & 1) * cells, sc.slab32 + (size_t)kj * cells, bytes, &bars[j & 1]);
            }
        };
        if (threadIdx.x == 0) {
            mbar_init(&bars[0], 1);
            mbar_init(&bars[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            issue(0);
            issue(1);
        }
        __syncthreads();
        for (int j = 0; j < nsteps; ++j) {
            const int kj = kb + j;
            if (kj >= 0 && kj < sc.nT) mbar_wait(&bars[j & 1], (uint32_t)(j >> 1) & 1u);
            if (st == PMCTS_ST_OK) advance(j, ring + (j & 1) * cells);
            // everybody is done with slot j & 1 (and learns whether anybody still runs) before it is refilled
            if (!__syncthreads_or(st == PMCTS_ST_OK)) break;
            if (threadIdx.x == 0) issue(j + 2);
        }
    }
    if (!mine) return;
    if (st == fast::kStNeedsLiteral) *any_resume = 1u;
    k_io[i] = k;
    lat_io[i] = lat;
    lon_io[i] = lon;
    status_io[i] = (uint8_t)st;
}

// PMCTS_STEP_HEADING_ACTIONS: headings handed over as indices into the reference's ACTIONS (simulatorTLKT.py:29,
// 0 .. 7 -> 0, 45, ... 315 degrees; any uint8 a means 45 a degrees) -- an eighth of the bytes a host caller has to
// send -- are expanded to the fp64 degrees the step kernels read.
__global__ void expand_actions_kernel(const uint8_t *__restrict__ a, double *__restrict__ h, size_t count) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) h[i] = 45.0 * (double)a[i];
}

// PMCTS_ROLL_PREFIX_ACTIONS: the leaves' action paths as int8 indices into ACTIONS (-1 = padding, never read).
__global__ void expand_prefix_kernel(const int8_t *__restrict__ a, double *__restrict__ h, size_t count) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) h[i] = 45.0 * (double)a[i];
}

// Narrowing of a wave's histograms for the exchange (pmcts_pack_bins): a count is at most `replicas`.
template <typename T>
__global__ void pack_bins_kernel(const uint32_t *__restrict__ bins, T *__restrict__ out, size_t count) {
    // (a few blocks striding, for the same reason as backup_kernel)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        out[i] = (T)bins[i];
}

// Roofline denominator: what the FP32 FMA pipes of this chip sustain (pmcts_probe_fp32_peak).  16 independent
// FFMA chains per thread, nothing else in the loop.
__global__ void __launch_bounds__(256) fp32_peak_kernel(float *__restrict__ out, int iters, float a, float b) {
    float x[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) x[q] = (float)(threadIdx.x + q);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 16; ++q) x[q] = fmaf(x[q], a, b);
    }
    float s = 0.0f;
#pragma unroll
    for (int q = 0; q < 16; ++q) s += x[q];
    if (s == 123.456f) out[0] = s;  // keeps the chains alive; never true for the a, b the host passes
}

// K3 -- Node.back_up (worker.py:55-70) for n_scen tables at once: one thread per (scenario, leaf).
// kBin: the width the histograms arrive in (uint32 from K2; uint8 / uint16 from an exchanged update block).
// The thread reads its leaf's 12 counts in three vector loads (a warp reads 32 leaves' worth of consecutive bytes),
// walks the chain of ancestors ONCE and adds the non-empty bins (usually two to four of the twelve) to every row on
// the way with fire-and-forget reductions -- the first version walked the chain once per (leaf, bin), twelve times the
// dependent parent / arm loads, and was bound by their latency (long scoreboard 34 cycles per issue).
// A few fat blocks striding over the leaves rather than one thread each: the back-up of a wave runs on a side stream
// UNDER the next wave's rollout kernel, whose blocks hold every block slot of the chip; a grid of thousands of tiny
// blocks then waits for a freed slot thousands of times, a grid of two blocks per SM waits once.
template <typename kBin>
__device__ __forceinline__ void load_counts(const kBin *__restrict__ p, bool vector, uint32_t (&c)[12]) {
    if (!vector) {
#pragma unroll
        for (int b = 0; b < 12; ++b) c[b] = (uint32_t)p[b];
    } else if (sizeof(kBin) == 4) {
        const uint4 *q = reinterpret_cast<const uint4 *>(p);
        const uint4 q0 = q[0], q1 = q[1], q2 = q[2];
        c[0] = q0.x, c[1] = q0.y, c[2] = q0.z, c[3] = q0.w, c[4] = q1.x, c[5] = q1.y, c[6] = q1.z, c[7] = q1.w;
        c[8] = q2.x, c[9] = q2.y, c[10] = q2.z, c[11] = q2.w;
    } else if (sizeof(kBin) == 2) {
        const uint2 *q = reinterpret_cast<const uint2 *>(p);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const uint2 v = q[k];
            c[4 * k] = v.x & 65535u, c[4 * k + 1] = v.x >> 16, c[4 * k + 2] = v.y & 65535u, c[4 * k + 3] = v.y >> 16;
        }
    } else {
        const uint32_t *q = reinterpret_cast<const uint32_t *>(p);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const uint32_t v = q[k];
            c[4 * k] = v & 255u, c[4 * k + 1] = (v >> 8) & 255u, c[4 * k + 2] = (v >> 16) & 255u, c[4 * k + 3] = v >> 24;
        }
    }
}

template <typename kBin>
__global__ void __launch_bounds__(256)
backup_kernel(uint32_t *__restrict__ table, int64_t n_nodes, const int32_t *__restrict__ parent,
              const int8_t *__restrict__ arm, int64_t n_leaf, int n_scen, const int32_t *__restrict__ leaf_node,
              const int8_t *__restrict__ leaf_slot, const kBin *__restrict__ leaf_bins, bool vector) {
    const int64_t total = n_leaf * n_scen;
    const int lane = threadIdx.x & 31;
    // (the trip count is the warp's: every lane stays in the loop, `walking` says whether it still carries counts)
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < total;
         base += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = base + lane;
        uint32_t c[12];
        uint32_t live = 0;
        uint32_t *tab = nullptr;
        int node = 0, slot = 0;
        if (i < total) {
            load_counts<kBin>(leaf_bins + i * 12, vector, c);
#pragma unroll
            for (int b = 0; b < 12; ++b) live |= (c[b] != 0u) << b;
            tab = table + (size_t)(i / n_leaf) * n_nodes * 96;
            node = leaf_node[i];
            slot = leaf_slot[i];
        }
        bool walking = live != 0;
        // Up the chain.  Leaves of a wave share their upper ancestors (all of them the root): lanes that have reached
        // the same row add their counts together, one of them files the sum and carries it on, the others are done --
        // a warp of 32 sibling leaves ends with one reduction per bin at the root instead of 32 on one address.
        while (true) {
            const unsigned mask = __ballot_sync(0xffffffffu, walking);
            if (!mask) break;
            if (walking) {
                uint32_t *row = tab + ((size_t)node * 8 + slot) * 12;
                const unsigned peers = __match_any_sync(mask, (unsigned long long)(uintptr_t)row);
                if (peers != (1u << lane)) {
                    live = __reduce_or_sync(peers, live);
#pragma unroll
                    for (int b = 0; b < 12; ++b)
                        if (live >> b & 1u) c[b] = __reduce_add_sync(peers, c[b]);
                    walking = (int)__ffs(peers) - 1 == lane;
                }
                if (walking) {
#pragma unroll
                    for (int b = 0; b < 12; ++b)
                        if (live >> b & 1u) atomicAdd(row + b, c[b]);
                    const int p = parent[node];
                    if (p < 0) walking = false;
                    slot = arm[node];
                    node = p;
                }
            }
        }
    }
}

// Hist.get_mean per (node, action) + visit total per node (utils.py:49-59, worker.py:91-95).
// Bin centres (b + 0.5) / 8 and counts are exact in fp64, so the dot product is exact in any order.
__global__ void hist_stats_kernel(const uint32_t *__restrict__ table, int64_t n_nodes,
                                  double *__restrict__ mean, uint32_t *__restrict__ visits) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // (node, action)
    const bool live = t < n_nodes * 8;
    uint32_t tot = 0;
    if (live) {
        const uint4 *row = reinterpret_cast<const uint4 *>(table + t * 12);
        const uint4 q0 = __ldg(row), q1 = __ldg(row + 1), q2 = __ldg(row + 2);
        const uint32_t h[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
        double dot = 0.0;
#pragma unroll
        for (int b = 0; b < 12; ++b) {
            tot += h[b];
            dot += (double)h[b] * ((b + 0.5) * 0.125);
        }
        if (mean) mean[t] = tot ? dot / (double)tot : 0.0;
    }
    if (visits) {
        // 8 consecutive lanes hold the 8 actions of one node
        uint32_t s = tot;
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (live && (t & 7) == 0) visits[t >> 3] = s;
    }
}

// ================================================================================================
// Host side
// ================================================================================================

static thread_local std::string g_last_error;

static int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (expr);                                                                \
        if (e_ != cudaSuccess)                                                                  \
            return fail(PMCTS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                    \
    } while (0)

struct Scenario {
    ScenarioView view{};
    void *blob = nullptr;  // one allocation holding slabs + axes
    uint64_t times_hash = 0;  // FNV-1a of the simulator instants (launches over several scenarios need equal axes)
    // raw weather grid kept for Weather.*Interpolator queries at arbitrary times
    const void *raw_u = nullptr, *raw_v = nullptr;
    const double *wtime = nullptr, *wlat = nullptr, *wlon = nullptr;
    int nt = 0, raw_dtype = 0;
};

}  // namespace pmcts

using namespace pmcts;

struct pmcts_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    int precision = PMCTS_PREC_F64;
    BoatParams params{};
    std::map<int, Scenario> scenarios;
    // grow-only device scratch for PMCTS_MEM_HOST calls (deferred launches stage in their slot's own buffer)
    struct StageBuf {
        char *p = nullptr;
        size_t cap = 0, used = 0;
    } stage;
    // grow-only device work area of the mixed-precision kernels: [0] redo / resume counter,
    // then the rollout ids (uint32) or per-boat resume steps (int32) and a status byte per boat
    char *work = nullptr;
    size_t work_cap = 0;
    double *head_buf = nullptr;  // PMCTS_STEP_HEADING_ACTIONS: the action indices expanded to degrees
    size_t head_cap = 0;
    bool margins_set = false;
    int packed_fp32 = 0;       // K2 turbo: two rollouts per thread on FFMA2 / FMUL2 / FADD2 (pmcts_fast2.cuh); measured
                               // slower than one per thread (4 warps per scheduler instead of 8: profiles/README.md), so off
    bool stage_slabs = false;  // K1: stage the step's wind slab in shared memory (pmcts_set_option); measured
                               // 21 % slower than the L1 / L2 gathers on config 3 (profiles/README.md), so off
    // PMCTS_ROLL_DEFER_REDO: the fp64 finisher of a rollout launch runs on a side stream so that it overlaps
    // the next launch; two work areas alternate, each guarded by the events of the launch that used it last
    cudaStream_t side_stream = nullptr;
    struct Deferred {
        char *work = nullptr;
        size_t cap = 0;
        StageBuf stage;  // PMCTS_MEM_HOST | PMCTS_ROLL_DEFER_REDO: device copies of this launch's host arrays
        cudaEvent_t main_done = nullptr, redo_done = nullptr;
        bool pending = false;  // not yet joined by the caller (pmcts_join)
        bool used = false;     // redo_done has been recorded at least once
        uint64_t seq = 0;
    } defer[2];
    uint64_t defer_seq = 0;
    // PMCTS_STEP_DEFER (host arrays): copies in on their own stream, kernel on the ctx stream, copies out on the side
    // stream -- the three stages of consecutive calls overlap; two staging buffers alternate
    cudaStream_t in_stream = nullptr;
    struct StepDeferred {
        StageBuf stage;
        cudaEvent_t in_done = nullptr, main_done = nullptr, out_done = nullptr;
        bool pending = false, used = false;
        uint64_t seq = 0;
    } step_defer[2];
    uint64_t step_seq = 0;
    int64_t launches = 0;
    int last_variant = PMCTS_VARIANT_LITERAL;  // of the last K1 / K2 launch (pmcts_last_kernel_variant)
    bool timing = false;
    struct Timed { cudaEvent_t a, b; int kind; };
    std::vector<Timed> timed;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> event_pool;
    double timed_ms[5] = {0, 0, 0, 0, 0};  // per PMCTS_KERNEL_* kind
    // Every entry point that takes a ctx holds this lock for its duration: calls on one ctx from several host
    // threads are serialised (they are stream-ordered anyway); recursive because upload frees the slot it replaces.
    std::recursive_mutex mu;
    void *attachment[2] = {nullptr, nullptr};  // pmcts::ctx_attachment
    void (*attachment_free[2])(void *) = {nullptr, nullptr};
    Deferred *last_fast = nullptr;  // work area of the last mixed-precision rollout launch when it was a deferred one
};

namespace {

const double kFitVelocity[36] = {
    -0.0310198207067136, -0.0600881995286002, 0.0286695485969272, -0.00684406929296715, 0.000379636836557256, -6.77704610076153e-06,
    0.0106968590293653, 0.00665508747173206, -4.03686836415123e-05, 2.38962919033178e-05, -6.16724919464073e-07, 0,
    -0.000370260554113750, -7.68003222256045e-05, -2.44425618051996e-06, -2.16961875441841e-08, 0, 0,
    4.94175336099537e-06, 5.68757177397705e-07, 1.17646387245609e-08, 0, 0, 0,
    -2.91900968067076e-08, -1.67287818401469e-09, 0, 0, 0, 0,
    6.34463723894578e-11, 0, 0, 0, 0, 0};

struct DeviceGuard {
    int prev = -1;
    std::unique_lock<std::recursive_mutex> lock;
    explicit DeviceGuard(pmcts_ctx *ctx) : DeviceGuard(ctx->device) { lock = std::unique_lock<std::recursive_mutex>(ctx->mu); }
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        int cur = -1;
        cudaGetDevice(&cur);
        if (prev >= 0 && cur != prev) cudaSetDevice(prev);
    }
};

// Bump allocator over the ctx scratch buffer (PMCTS_MEM_HOST staging).
struct Stager {
    pmcts_ctx *ctx;
    bool host;
    pmcts_ctx::StageBuf *buf;
    struct Out { void *host_ptr; void *dev_ptr; size_t bytes; };
    std::vector<Out> outs;

    cudaStream_t in_stream;  // where the copies in are enqueued (the ctx stream unless the caller pipelines them)

    Stager(pmcts_ctx *c, bool h, pmcts_ctx::StageBuf *b = nullptr, cudaStream_t ins = nullptr)
        : ctx(c), host(h), buf(b ? b : &c->stage), in_stream(ins ? ins : c->stream) {}

    static size_t align(size_t x) { return (x + 255) & ~(size_t)255; }

    int reserve(size_t total) {
        if (!host) return PMCTS_OK;
        total += 4096;
        if (total > buf->cap) {
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            if (ctx->side_stream) CUDA_TRY(cudaStreamSynchronize(ctx->side_stream));
            if (ctx->in_stream) CUDA_TRY(cudaStreamSynchronize(ctx->in_stream));
            if (buf->p) CUDA_TRY(cudaFree(buf->p));
            buf->p = nullptr;
            buf->cap = 0;
            size_t cap = total + total / 4;
            CUDA_TRY(cudaMalloc((void **)&buf->p, cap));
            buf->cap = cap;
        }
        buf->used = 0;
        return PMCTS_OK;
    }

    void *carve(size_t bytes) {
        void *p = buf->p + buf->used;
        buf->used += align(bytes);
        return p;
    }

    // input (and optionally output) array
    template <typename T>
    int in(T *&ptr, size_t count, bool also_out = false) {
        if (!host || ptr == nullptr) return PMCTS_OK;
        const size_t bytes = count * sizeof(T);
        void *d = carve(bytes);
        CUDA_TRY(cudaMemcpyAsync(d, (const void *)ptr, bytes, cudaMemcpyHostToDevice, in_stream));
        if (also_out) outs.push_back({(void *)ptr, d, bytes});
        ptr = (T *)d;
        return PMCTS_OK;
    }
    // fill: -1 leave as is, 0 zero, 0xFF all-ones (a quiet NaN for fp64)
    template <typename T>
    int out(T *&ptr, size_t count, int fill = -1) {
        if (ptr == nullptr) return PMCTS_OK;
        const size_t bytes = count * sizeof(T);
        if (host) {
            void *d = carve(bytes);
            outs.push_back({(void *)ptr, d, bytes});
            ptr = (T *)d;
        }
        if (fill >= 0) CUDA_TRY(cudaMemsetAsync(ptr, fill, bytes, ctx->stream));
        return PMCTS_OK;
    }
    // results back to the caller's arrays on `stream` (no host synchronisation)
    int copy_back(cudaStream_t stream) {
        if (!host) return PMCTS_OK;
        for (auto &o : outs)
            CUDA_TRY(cudaMemcpyAsync(o.host_ptr, o.dev_ptr, o.bytes, cudaMemcpyDeviceToHost, stream));
        return PMCTS_OK;
    }
    int finish() {
        if (!host) return PMCTS_OK;
        int rc = copy_back(ctx->stream);
        if (rc) return rc;
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return PMCTS_OK;
    }
};

size_t al(size_t bytes) { return Stager::align(bytes); }

struct LaunchTimer {
    pmcts_ctx *ctx;
    cudaEvent_t a = nullptr, b = nullptr;
    int kind;
    cudaStream_t stream;
    explicit LaunchTimer(pmcts_ctx *c, int kind_ = PMCTS_KERNEL_OTHER, cudaStream_t s = nullptr)
        : ctx(c), kind(kind_), stream(s ? s : c->stream) {
        ctx->launches++;
        if (!ctx->timing) return;
        if (!ctx->event_pool.empty()) {
            a = ctx->event_pool.back().first;
            b = ctx->event_pool.back().second;
            ctx->event_pool.pop_back();
        } else {
            cudaEventCreate(&a);
            cudaEventCreate(&b);
        }
        cudaEventRecord(a, stream);
    }
    ~LaunchTimer() {
        if (!a) return;
        cudaEventRecord(b, stream);
        ctx->timed.push_back({a, b, kind});
    }
};

int grid_for(int64_t n) { return (int)((n + kBlock - 1) / kBlock); }

// Work area of the mixed-precision kernels: 256 B header (counter) + `bytes` payload.
int ensure_work(pmcts_ctx *ctx, size_t bytes) {
    const size_t need = 256 + bytes;
    if (need > ctx->work_cap) {
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (ctx->work) CUDA_TRY(cudaFree(ctx->work));
        ctx->work = nullptr;
        ctx->work_cap = 0;
        const size_t cap = need + need / 4;
        CUDA_TRY(cudaMalloc((void **)&ctx->work, cap));
        ctx->work_cap = cap;
    }
    return PMCTS_OK;
}

// Can this launch run on the fp32 kernels?  (Otherwise the literal fp64 kernels run: still the GPU.)
bool fast_eligible(const pmcts_ctx *ctx, const Scenario *sc) {
    const BoatParams &p = ctx->params;
    return ctx->precision == PMCTS_PREC_FAST && sc->view.fast_ok && p.pofsail_min >= 0.0 &&
           p.pofsail_min <= 90.0 && p.pofsail_max >= 90.0 && p.pofsail_max <= 180.0 && p.wind_max > 0.0;
}

int find_scenario(pmcts_ctx *ctx, int scen, Scenario **out, bool need_axis = true) {
    auto it = ctx->scenarios.find(scen);
    if (it == ctx->scenarios.end()) return fail(PMCTS_ERR_SCENARIO, "scenario slot %d not uploaded", scen);
    if (need_axis && it->second.view.nT < 1)
        return fail(PMCTS_ERR_SCENARIO, "scenario slot %d holds a weather grid only (uploaded without simulator instants)", scen);
    *out = &it->second;
    return PMCTS_OK;
}

NoiseView noise_view(const pmcts_noise *nz) {
    NoiseView v{};
    if (!nz) {
        v.mode = PMCTS_NOISE_NONE;
        return v;
    }
    v.mode = nz->mode;
    v.z = nz->z;
    v.seed = nz->seed;
    v.stream = nz->stream;
    v.id0 = nz->id0;
    return v;
}

bool uniform_axis(const double *g, int n, double *inv_step) {
    if (n < 2) return false;
    const double step = (g[n - 1] - g[0]) / (n - 1);
    if (!(step > 0)) return false;
    for (int i = 0; i < n; ++i)
        if (std::fabs(g[i] - (g[0] + i * step)) > 1e-9 * std::fabs(step) + 1e-12 * std::fabs(g[i])) return false;
    *inv_step = 1.0 / step;
    return true;
}

}  // namespace

// (for the other translation units of the library -- pmcts_search.cu: they report through pmcts_last_error too)
namespace pmcts {
void set_last_error(const char *msg) { g_last_error = msg ? msg : ""; }
void count_launches(pmcts_ctx *ctx, int n) { ctx->launches += n; }
bool ctx_timing(const pmcts_ctx *ctx) { return ctx->timing; }
std::unique_lock<std::recursive_mutex> ctx_lock(pmcts_ctx *ctx) { return std::unique_lock<std::recursive_mutex>(ctx->mu); }
void **ctx_attachment(pmcts_ctx *ctx, void (*free_fn)(void *), int key) {
    ctx->attachment_free[key] = free_fn;
    return &ctx->attachment[key];
}
}  // namespace pmcts

extern "C" {

const char *pmcts_last_error(void) { return g_last_error.c_str(); }
const char *pmcts_version(void) { return "pmcts_b200 0.1 (sm_100a)"; }

int pmcts_create(int device, pmcts_ctx **out) {
    if (!out) return fail(PMCTS_ERR_ARG, "out is NULL");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(PMCTS_ERR_NO_DEVICE, "no CUDA device (%s); pmcts_b200 has no CPU path",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(PMCTS_ERR_ARG, "device %d out of range [0,%d)", device, count);
    DeviceGuard g(device);
    pmcts_ctx *ctx = new pmcts_ctx();
    ctx->device = device;
    e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete ctx;
        return fail(PMCTS_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e));
    }
    ctx->stream = ctx->own_stream;
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (ctx->sm_count < 1) ctx->sm_count = 148;
    pmcts_set_params(ctx, kFitVelocity, 19.1, 35.0, 160.0, 0.2, 0.005 * kDegToRad, 6371e3);
    *out = ctx;
    return PMCTS_OK;
}

void pmcts_destroy(pmcts_ctx *ctx) {
    if (!ctx) return;
    DeviceGuard g(ctx->device);  // (no lock: the caller must not use a ctx it is destroying)
    cudaStreamSynchronize(ctx->stream);
    for (int key = 0; key < 2; ++key) {
        if (ctx->attachment[key] && ctx->attachment_free[key]) ctx->attachment_free[key](ctx->attachment[key]);
        ctx->attachment[key] = nullptr;
    }
    for (auto &kv : ctx->scenarios) cudaFree(kv.second.blob);
    if (ctx->side_stream) cudaStreamSynchronize(ctx->side_stream);
    if (ctx->stage.p) cudaFree(ctx->stage.p);
    for (auto &d : ctx->defer)
        if (d.stage.p) cudaFree(d.stage.p);
    if (ctx->in_stream) cudaStreamSynchronize(ctx->in_stream);
    for (auto &d : ctx->step_defer) {
        if (d.stage.p) cudaFree(d.stage.p);
        if (d.in_done) cudaEventDestroy(d.in_done);
        if (d.main_done) cudaEventDestroy(d.main_done);
        if (d.out_done) cudaEventDestroy(d.out_done);
    }
    if (ctx->in_stream) cudaStreamDestroy(ctx->in_stream);
    if (ctx->work) cudaFree(ctx->work);
    if (ctx->head_buf) cudaFree(ctx->head_buf);
    for (auto &d : ctx->defer) {
        if (d.work) cudaFree(d.work);
        if (d.main_done) cudaEventDestroy(d.main_done);
        if (d.redo_done) cudaEventDestroy(d.redo_done);
    }
    if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
    for (auto &p : ctx->timed) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    for (auto &p : ctx->event_pool) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int pmcts_set_stream(pmcts_ctx *ctx, void *cuda_stream) {
    if (!ctx) return fail(PMCTS_ERR_ARG, "ctx is NULL");
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    ctx->stream = (cudaStream_t)cuda_stream;
    return PMCTS_OK;
}

void *pmcts_get_stream(pmcts_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
int pmcts_get_device(pmcts_ctx *ctx) { return ctx ? ctx->device : PMCTS_ERR_ARG; }

int pmcts_reset_stream(pmcts_ctx *ctx) {
    if (!ctx) return fail(PMCTS_ERR_ARG, "ctx is NULL");
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    ctx->stream = ctx->own_stream;
    return PMCTS_OK;
}

int pmcts_synchronize(pmcts_ctx *ctx) {
    if (!ctx) return fail(PMCTS_ERR_ARG, "ctx is NULL");
    DeviceGuard g(ctx);
    if (ctx->in_stream) CUDA_TRY(cudaStreamSynchronize(ctx->in_stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (ctx->side_stream) CUDA_TRY(cudaStreamSynchronize(ctx->side_stream));
    return PMCTS_OK;
}

int pmcts_set_precision(pmcts_ctx *ctx, int mode) {
    if (!ctx) return fail(PMCTS_ERR_ARG, "ctx is NULL");
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (mode != PMCTS_PREC_F64 && mode != PMCTS_PREC_FAST) return fail(PMCTS_ERR_ARG, "unknown precision mode %d", mode);
    ctx->precision = mode;
    return PMCTS_OK;
}

int64_t pmcts_launch_count(pmcts_ctx *ctx) { return ctx ? ctx->launches : 0; }

int pmcts_last_kernel_variant(pmcts_ctx *ctx) { return ctx ? ctx->last_variant : PMCTS_ERR_ARG; }

int pmcts_set_option(pmcts_ctx *ctx, int option, int value) {
    if (!ctx) return fail(PMCTS_ERR_ARG, "ctx is NULL");
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (option == PMCTS_OPT_STAGE_SLABS) {
        ctx->stage_slabs = value != 0;
        return PMCTS_OK;
    }
    if (option == PMCTS_OPT_PACKED_FP32) {
        ctx->packed_fp32 = value;
        return PMCTS_OK;
    }
    return fail(PMCTS_ERR_ARG, "unknown option %d", option);
}

int pmcts_probe_fp32_peak(pmcts_ctx *ctx, double *tflops) {
    if (!ctx || !tflops) return fail(PMCTS_ERR_ARG, "NULL argument");
    DeviceGuard g(ctx);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, ctx->device));
    float *out = nullptr;
    CUDA_TRY(cudaMalloc((void **)&out, sizeof(float)));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 8192;
    const double flop = 2.0 * 16.0 * (double)iters * (double)blocks * (double)threads;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {  // the first launch warms up
        CUDA_TRY(cudaEventRecord(e0, ctx->stream));
        fp32_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(out, iters, 0.999f, 0.001f);
        CUDA_TRY(cudaEventRecord(e1, ctx->stream));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms > 0.f) best = std::max(best, flop / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    CUDA_TRY(cudaGetLastError());
    *tflops = best;
    return PMCTS_OK;
}

int pmcts_enable_timing(pmcts_ctx *ctx, int on) {
    if (!ctx) return fail(PMCTS_ERR_ARG, "ctx is NULL");
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    ctx->timing = on != 0;
    return PMCTS_OK;
}

static void drain_timers(pmcts_ctx *ctx) {
    cudaStreamSynchronize(ctx->stream);
    if (ctx->side_stream) cudaStreamSynchronize(ctx->side_stream);
    for (auto &p : ctx->timed) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) ctx->timed_ms[p.kind < 5 ? p.kind : 3] += ms;
        ctx->event_pool.push_back({p.a, p.b});
    }
    ctx->timed.clear();
}

double pmcts_kernel_ms(pmcts_ctx *ctx, int kind, int reset) {
    if (!ctx || kind < 0 || kind > 4) return -1.0;
    DeviceGuard g(ctx);
    drain_timers(ctx);
    const double r = ctx->timed_ms[kind];
    if (reset) ctx->timed_ms[kind] = 0.0;
    return r;
}

double pmcts_kernel_ms_total(pmcts_ctx *ctx, int reset) {
    if (!ctx) return -1.0;
    double r = 0.0;
    for (int kind = 0; kind < 5; ++kind) r += pmcts_kernel_ms(ctx, kind, reset);
    return r;
}

int pmcts_set_params(pmcts_ctx *ctx, const double polar[36], double wind_max, double pofsail_min,
                     double pofsail_max, double sigma_coeff, double dest_angle, double earth_radius) {
    if (!ctx || !polar) return fail(PMCTS_ERR_ARG, "NULL argument");
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    BoatParams &p = ctx->params;
    for (int i = 0; i < 36; ++i) p.polar[i] = polar[i];
    p.wind_max = wind_max;
    p.pofsail_min = pofsail_min;
    p.pofsail_max = pofsail_max;
    p.sigma_coeff = sigma_coeff;
    p.dest_angle = dest_angle;
    p.earth_radius = earth_radius;
    p.cos_pmin = std::cos(kPi * pofsail_min / 180.0);
    p.cos_pmax = std::cos(kPi * pofsail_max / 180.0);
    derive_fast_params(p);
    if (!ctx->margins_set) default_fast_margins(p);
    return PMCTS_OK;
}

int pmcts_set_fast_margins(pmcts_ctx *ctx, double angle, double along, double near_miss2, double box, double bin) {
    if (!ctx) return fail(PMCTS_ERR_ARG, "ctx is NULL");
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    BoatParams &p = ctx->params;
    if (angle < 0 && along < 0 && near_miss2 < 0 && box < 0 && bin < 0) {
        default_fast_margins(p);
        ctx->margins_set = false;
        return PMCTS_OK;
    }
    if (angle >= 0) p.margin_angle = (float)angle;
    if (along >= 0) p.margin_along = (float)along;
    if (near_miss2 >= 0) p.near_miss2 = (float)near_miss2;
    if (box >= 0) p.margin_box = (float)box;
    if (bin >= 0) p.margin_bin = bin;
    ctx->margins_set = true;
    return PMCTS_OK;
}

int64_t pmcts_last_redo_count(pmcts_ctx *ctx) {
    if (!ctx) return 0;
    DeviceGuard g(ctx);
    const char *work = ctx->last_fast ? ctx->last_fast->work : ctx->work;  // a deferred launch counts in its slot
    if (!work) return 0;
    uint32_t c = 0;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return -1;
    if (ctx->side_stream && cudaStreamSynchronize(ctx->side_stream) != cudaSuccess) return -1;
    if (cudaMemcpy(&c, work, sizeof c, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (int64_t)c;
}

int pmcts_scenario_free(pmcts_ctx *ctx, int scen) {
    if (!ctx) return fail(PMCTS_ERR_ARG, "ctx is NULL");
    DeviceGuard g(ctx);
    auto it = ctx->scenarios.find(scen);
    if (it == ctx->scenarios.end()) return PMCTS_OK;
    // deferred finishers (side stream) and pipelined copies (in stream) may still read the slabs
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (ctx->side_stream) CUDA_TRY(cudaStreamSynchronize(ctx->side_stream));
    if (ctx->in_stream) CUDA_TRY(cudaStreamSynchronize(ctx->in_stream));
    CUDA_TRY(cudaFree(it->second.blob));
    ctx->scenarios.erase(it);
    return PMCTS_OK;
}

int pmcts_scenario_upload(pmcts_ctx *ctx, int scen, const double *wtime, int nt, const double *wlat,
                          int nlat, const double *wlon, int nlon, const void *u, const void *v,
                          int dtype, const double *sim_times, int nT, const double box[4]) {
    if (!ctx || !wtime || !wlat || !wlon || !u || !v) return fail(PMCTS_ERR_ARG, "NULL argument");
    if (nT < 0 || (nT > 0 && (!sim_times || !box))) return fail(PMCTS_ERR_ARG, "a simulator axis needs sim_times and box");
    if (nt < 2 || nlat < 2 || nlon < 2) return fail(PMCTS_ERR_ARG, "axes too short");
    static const double no_box[4] = {0, 0, 0, 0};
    if (nT == 0) box = no_box;
    if (dtype != 0 && dtype != 1) return fail(PMCTS_ERR_ARG, "dtype must be 0 (f64) or 1 (f32)");
    double inv_dlat, inv_dlon;
    if (!uniform_axis(wlat, nlat, &inv_dlat) || !uniform_axis(wlon, nlon, &inv_dlon))
        return fail(PMCTS_ERR_AXIS, "latitude / longitude axes must be ascending and uniformly spaced");
    for (int i = 1; i < nt; ++i)
        if (!(wtime[i] > wtime[i - 1])) return fail(PMCTS_ERR_AXIS, "weather time axis must be strictly ascending");
    DeviceGuard g(ctx);
    int rc = pmcts_scenario_free(ctx, scen);
    if (rc) return rc;

    // per-instant time cell and weight (SciPy find_indices semantics on the time axis)
    const TimeCells tc = time_cells(wtime, nt, sim_times, nT);
    const std::vector<int> &it = tc.it;
    const std::vector<double> &wt = tc.wt, &dt = tc.dt;
    const std::vector<uint8_t> &valid = tc.valid;

    const size_t plane = (size_t)nlat * nlon, cells = plane * nT;
    const size_t esz = dtype == 0 ? 8 : 4;
    const size_t o_slab64 = 0;
    const size_t o_slab32 = o_slab64 + al(cells * sizeof(double2));
    const size_t stride32 = (plane + 1) & ~(size_t)1;  // every fp32 slab 16-byte aligned
    const size_t o_times = o_slab32 + al(stride32 * nT * sizeof(float2));
    const size_t o_dt = o_times + al(nT * sizeof(double));
    const size_t o_valid = o_dt + al(nT * sizeof(double));
    const size_t o_dt32 = o_valid + al(nT);
    const size_t raw = (size_t)nt * plane * esz;
    const size_t o_u = o_dt32 + al(nT * sizeof(float));
    const size_t o_v = o_u + al(raw);
    const size_t o_wtime = o_v + al(raw);
    const size_t o_wlat = o_wtime + al(nt * sizeof(double));
    const size_t o_wlon = o_wlat + al(nlat * sizeof(double));
    const size_t total = o_wlon + al(nlon * sizeof(double));
    // temporaries: time cells
    const size_t t_it = 0, t_wt = al(nT * sizeof(int));
    const size_t tmp_total = t_wt + al(nT * sizeof(double)) + 256;

    Scenario sc;
    char *tmp = nullptr;
    CUDA_TRY(cudaMalloc(&sc.blob, total));
    cudaError_t e = cudaMalloc((void **)&tmp, tmp_total);
    if (e != cudaSuccess) {
        cudaFree(sc.blob);
        return fail(PMCTS_ERR_CUDA, "cudaMalloc(%zu) failed: %s", tmp_total, cudaGetErrorString(e));
    }
    char *base = (char *)sc.blob;
    cudaStream_t s = ctx->stream;
    auto cleanup = [&]() { cudaFree(tmp); cudaFree(sc.blob); };
#define UP_TRY(expr)                                                                              \
    do {                                                                                          \
        cudaError_t e2_ = (expr);                                                                 \
        if (e2_ != cudaSuccess) {                                                                 \
            cleanup();                                                                            \
            return fail(PMCTS_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e2_));         \
        }                                                                                         \
    } while (0)
    UP_TRY(cudaMemsetAsync(base + o_slab32, 0, stride32 * nT * sizeof(float2), s));  // (the padding cell of odd planes)
    UP_TRY(cudaMemcpyAsync(base + o_u, u, raw, cudaMemcpyHostToDevice, s));
    UP_TRY(cudaMemcpyAsync(base + o_v, v, raw, cudaMemcpyHostToDevice, s));
    UP_TRY(cudaMemcpyAsync(base + o_wtime, wtime, nt * sizeof(double), cudaMemcpyHostToDevice, s));
    UP_TRY(cudaMemcpyAsync(base + o_wlat, wlat, nlat * sizeof(double), cudaMemcpyHostToDevice, s));
    UP_TRY(cudaMemcpyAsync(base + o_wlon, wlon, nlon * sizeof(double), cudaMemcpyHostToDevice, s));
    if (nT > 0) {
        UP_TRY(cudaMemcpyAsync(tmp + t_it, it.data(), nT * sizeof(int), cudaMemcpyHostToDevice, s));
        UP_TRY(cudaMemcpyAsync(tmp + t_wt, wt.data(), nT * sizeof(double), cudaMemcpyHostToDevice, s));
        UP_TRY(cudaMemcpyAsync(base + o_times, sim_times, nT * sizeof(double), cudaMemcpyHostToDevice, s));
        UP_TRY(cudaMemcpyAsync(base + o_dt, dt.data(), nT * sizeof(double), cudaMemcpyHostToDevice, s));
        UP_TRY(cudaMemcpyAsync(base + o_valid, valid.data(), nT, cudaMemcpyHostToDevice, s));
        UP_TRY(cudaMemcpyAsync(base + o_dt32, tc.dt32.data(), nT * sizeof(float), cudaMemcpyHostToDevice, s));
    }
    if (nT > 0) {
        LaunchTimer lt(ctx);
        const int grid = (int)((cells + 255) / 256);
        if (dtype == 0)
            blend_time_slabs_kernel<double><<<grid, 256, 0, s>>>(
                (const double *)(base + o_u), (const double *)(base + o_v), (const int *)(tmp + t_it),
                (const double *)(tmp + t_wt), nT, (int)plane, (int)stride32, (double2 *)(base + o_slab64),
                (float2 *)(base + o_slab32));
        else
            blend_time_slabs_kernel<float><<<grid, 256, 0, s>>>(
                (const float *)(base + o_u), (const float *)(base + o_v), (const int *)(tmp + t_it),
                (const double *)(tmp + t_wt), nT, (int)plane, (int)stride32, (double2 *)(base + o_slab64),
                (float2 *)(base + o_slab32));
    }
    UP_TRY(cudaGetLastError());
    UP_TRY(cudaStreamSynchronize(s));  // host vectors and tmp die here
#undef UP_TRY
    cudaFree(tmp);

    ScenarioView &vw = sc.view;
    vw.slab64 = (const double2 *)(base + o_slab64);
    vw.slab32 = (const float2 *)(base + o_slab32);
    vw.times = (const double *)(base + o_times);
    vw.dt_sec = (const double *)(base + o_dt);
    vw.k_valid = (const uint8_t *)(base + o_valid);
    vw.nT = nT;
    vw.nlat = nlat;
    vw.nlon = nlon;
    vw.lat0 = wlat[0];
    vw.lon0 = wlon[0];
    vw.inv_dlat = inv_dlat;
    vw.inv_dlon = inv_dlon;
    vw.lat_last = wlat[nlat - 1];
    vw.lon_last = wlon[nlon - 1];
    vw.box_lat_lo = box[0];
    vw.box_lat_hi = box[1];
    vw.box_lon_lo = box[2];
    vw.box_lon_hi = box[3];
    vw.tmax = nT > 0 ? sim_times[nT - 1] : 0.0;
    sc.raw_u = base + o_u;
    sc.raw_v = base + o_v;
    sc.wtime = (const double *)(base + o_wtime);
    sc.wlat = (const double *)(base + o_wlat);
    sc.wlon = (const double *)(base + o_wlon);
    sc.nt = nt;
    sc.raw_dtype = dtype;
    vw.dt_sec32 = (const float *)(base + o_dt32);
    vw.kv_lo = tc.kv_lo;
    vw.kv_hi = tc.kv_hi;
    vw.max_abs_lat = (float)std::fmax(std::fabs(wlat[0]), std::fabs(wlat[nlat - 1]));
    vw.dt_max = tc.dt_max;
    vw.box_fa_lo = (float)((box[0] - vw.lat0) * inv_dlat);
    vw.box_fa_hi = (float)((box[1] - vw.lat0) * inv_dlat);
    vw.box_fo_lo = (float)((box[2] - vw.lon0) * inv_dlon);
    vw.box_fo_hi = (float)((box[3] - vw.lon0) * inv_dlon);
    // the fp32 kernels assume |lat| <= 90 on the grid (their sin / cos kernels cover [-pi/2, pi/2])
    // and one contiguous run of valid instants
    vw.slab32_stride = (int)stride32;
    vw.fa_c0 = -vw.lat0 * inv_dlat;
    vw.fo_c0 = -vw.lon0 * inv_dlon;
    vw.ks_lo = std::max(0, tc.kv_lo);
    vw.ks_hi = std::min(nT - 2, tc.kv_hi);
    finish_fast_view(vw, tc, box);
    vw.fast_ok = (nT > 0 && tc.contiguous && vw.max_abs_lat <= 90.0f && stride32 * nT < ((size_t)1 << 31)) ? 1 : 0;
    {
        uint64_t h = 1469598103934665603ull;
        const unsigned char *bytes = (const unsigned char *)sim_times;
        for (size_t i = 0; i < (size_t)nT * sizeof(double); ++i) h = (h ^ bytes[i]) * 1099511628211ull;
        sc.times_hash = h;
    }
    ctx->scenarios[scen] = sc;
    return PMCTS_OK;
}

int pmcts_get_wind(pmcts_ctx *ctx, int scen, int64_t n, const int32_t *k, const double *lat,
                   const double *lon, double *u, double *v, uint8_t *status, int flags) {
    if (!ctx || !k || !lat || !lon || !u || !v) return fail(PMCTS_ERR_ARG, "NULL argument");
    if (n <= 0) return PMCTS_OK;
    DeviceGuard g(ctx);
    Scenario *sc;
    int rc = find_scenario(ctx, scen, &sc);
    if (rc) return rc;
    Stager st(ctx, flags & PMCTS_MEM_HOST);
    rc = st.reserve(al(n * 4) + 4 * al(n * 8) + al(n));
    if (rc) return rc;
    if ((rc = st.in(k, n)) || (rc = st.in(lat, n)) || (rc = st.in(lon, n)) || (rc = st.out(u, n)) ||
        (rc = st.out(v, n)) || (rc = st.out(status, n)))
        return rc;
    {
        LaunchTimer lt(ctx);
        get_wind_kernel<<<grid_for(n), kBlock, 0, ctx->stream>>>(sc->view, n, k, lat, lon, u, v, status);
    }
    CUDA_TRY(cudaGetLastError());
    return st.finish();
}

int pmcts_interp_wind(pmcts_ctx *ctx, int scen, int64_t n, const double *t, const double *lat, const double *lon,
                      double *u, double *v, uint8_t *status, int flags) {
    if (!ctx || !t || !lat || !lon || !u || !v) return fail(PMCTS_ERR_ARG, "NULL argument");
    if (n <= 0) return PMCTS_OK;
    DeviceGuard g(ctx);
    Scenario *sc;
    int rc = find_scenario(ctx, scen, &sc, false);
    if (rc) return rc;
    Stager st(ctx, flags & PMCTS_MEM_HOST);
    rc = st.reserve(5 * al(n * 8) + al(n));
    if (rc) return rc;
    if ((rc = st.in(t, n)) || (rc = st.in(lat, n)) || (rc = st.in(lon, n)) || (rc = st.out(u, n)) ||
        (rc = st.out(v, n)) || (rc = st.out(status, n)))
        return rc;
    {
        LaunchTimer lt(ctx);
        const ScenarioView &vw = sc->view;
        if (sc->raw_dtype == 0)
            interp_wind_kernel<double><<<grid_for(n), kBlock, 0, ctx->stream>>>(
                (const double *)sc->raw_u, (const double *)sc->raw_v, sc->wtime, sc->nt, sc->wlat, vw.nlat, sc->wlon,
                vw.nlon, n, t, lat, lon, u, v, status);
        else
            interp_wind_kernel<float><<<grid_for(n), kBlock, 0, ctx->stream>>>(
                (const float *)sc->raw_u, (const float *)sc->raw_v, sc->wtime, sc->nt, sc->wlat, vw.nlat, sc->wlon,
                vw.nlon, n, t, lat, lon, u, v, status);
    }
    CUDA_TRY(cudaGetLastError());
    return st.finish();
}

int pmcts_step_batch(pmcts_ctx *ctx, int scen, int64_t n, int nsteps, int32_t *k, double *lat,
                     double *lon, double *prev_lat, double *prev_lon, const double *heading,
                     int seg_len, const pmcts_noise *noise, uint8_t *status, double *traj_lat,
                     double *traj_lon, int flags) {
    if (!ctx || !k || !lat || !lon || !heading) return fail(PMCTS_ERR_ARG, "NULL argument");
    if (n < 0 || nsteps < 0 || seg_len < 1) return fail(PMCTS_ERR_ARG, "bad sizes");
    if (n == 0 || nsteps == 0) return PMCTS_OK;
    DeviceGuard g(ctx);
    Scenario *sc;
    int rc = find_scenario(ctx, scen, &sc);
    if (rc) return rc;
    NoiseView nz = noise_view(noise);
    if (nz.mode == PMCTS_NOISE_INJECTED && !nz.z) return fail(PMCTS_ERR_ARG, "injected noise needs z");
    const size_t nseg = ((size_t)nsteps + seg_len - 1) / seg_len;
    const bool actions = (flags & PMCTS_STEP_HEADING_ACTIONS) != 0;
    const uint8_t *heading8 = actions ? reinterpret_cast<const uint8_t *>(heading) : nullptr;
    // PMCTS_STEP_DEFER with host arrays: three-stage pipeline over consecutive calls (see pmcts_ctx::StepDeferred)
    const bool defer = (flags & PMCTS_STEP_DEFER) && (flags & PMCTS_MEM_HOST);
    pmcts_ctx::StepDeferred *slot = nullptr;
    if (defer) {
        slot = &ctx->step_defer[ctx->step_seq & 1];
        if (!ctx->in_stream) CUDA_TRY(cudaStreamCreateWithFlags(&ctx->in_stream, cudaStreamNonBlocking));
        if (!ctx->side_stream) {
            int lo = 0, hi = 0;
            CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            CUDA_TRY(cudaStreamCreateWithPriority(&ctx->side_stream, cudaStreamNonBlocking, hi));
        }
        if (!slot->in_done) {
            CUDA_TRY(cudaEventCreateWithFlags(&slot->in_done, cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&slot->main_done, cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&slot->out_done, cudaEventDisableTiming));
        }
        // the call that staged in this buffer two calls ago must have copied its results out before anything of this
        // call touches the buffer: the copies in (in_stream) and the fills of the trajectory outputs, which the
        // Stager enqueues on the ctx stream
        if (slot->used) {
            CUDA_TRY(cudaStreamWaitEvent(ctx->in_stream, slot->out_done, 0));
            CUDA_TRY(cudaStreamWaitEvent(ctx->stream, slot->out_done, 0));
        }
    }
    Stager st(ctx, flags & PMCTS_MEM_HOST, slot ? &slot->stage : nullptr, slot ? ctx->in_stream : nullptr);
    rc = st.reserve(al(n * 4) + 4 * al(n * 8) + al(nseg * n * (actions ? 1 : 8)) + al(n) +
                    (nz.mode == PMCTS_NOISE_INJECTED ? al((size_t)nsteps * n * 8) : 0) +
                    (traj_lat ? al((size_t)nsteps * n * 8) : 0) + (traj_lon ? al((size_t)nsteps * n * 8) : 0));
    if (rc) return rc;
    if ((rc = st.in(k, n, true)) || (rc = st.in(lat, n, true)) || (rc = st.in(lon, n, true)) ||
        (rc = st.in(prev_lat, n, true)) || (rc = st.in(prev_lon, n, true)) ||
        (rc = actions ? st.in(heading8, nseg * n) : st.in(heading, nseg * n)) || (rc = st.in(status, n, true)) ||
        (rc = st.out(traj_lat, (size_t)nsteps * n, 0xFF)) || (rc = st.out(traj_lon, (size_t)nsteps * n, 0xFF)))
        return rc;
    if (nz.mode == PMCTS_NOISE_INJECTED && (rc = st.in(nz.z, (size_t)nsteps * n))) return rc;
    if (defer) {
        CUDA_TRY(cudaEventRecord(slot->in_done, ctx->in_stream));
        CUDA_TRY(cudaStreamWaitEvent(ctx->stream, slot->in_done, 0));
    }
    if (actions) {
        const size_t count = nseg * (size_t)n;
        if (count > ctx->head_cap) {
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            if (ctx->head_buf) CUDA_TRY(cudaFree(ctx->head_buf));
            ctx->head_buf = nullptr;
            ctx->head_cap = 0;
            CUDA_TRY(cudaMalloc((void **)&ctx->head_buf, (count + count / 4) * sizeof(double)));
            ctx->head_cap = count + count / 4;
        }
        LaunchTimer lt(ctx);
        expand_actions_kernel<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(heading8, ctx->head_buf, count);
        CUDA_TRY(cudaGetLastError());
        heading = ctx->head_buf;
    }
    if (fast_eligible(ctx, sc)) {
        // work area: counter | j_resume[n] | status[n] (when the caller passes no status array)
        const size_t o_status = al(n * sizeof(int32_t));
        if ((rc = ensure_work(ctx, o_status + al(n)))) return rc;
        uint32_t *any_resume = (uint32_t *)ctx->work;
        int32_t *j_resume = (int32_t *)(ctx->work + 256);
        ctx->last_fast = nullptr;
        uint8_t *status_dev = status;
        CUDA_TRY(cudaMemsetAsync(any_resume, 0, sizeof(uint32_t), ctx->stream));
        if (!status_dev) {
            status_dev = (uint8_t *)(ctx->work + 256 + o_status);
            CUDA_TRY(cudaMemsetAsync(status_dev, 0, n, ctx->stream));
        }
        {
            LaunchTimer lt(ctx, PMCTS_KERNEL_STEP);
            const bool lean = !prev_lat && !prev_lon && !traj_lat && !traj_lon && !(flags & PMCTS_STEP_STOP_BEFORE_TERMINAL);
#define PMCTS_LAUNCH_STEP(NOISE, LEAN, ...)                                                                 \
            step_fast_kernel<NOISE, LEAN, ##__VA_ARGS__><<<grid_for(n), kBlock, 0, ctx->stream>>>(          \
                sc->view, ctx->params, n, nsteps, k, lat, lon, prev_lat, prev_lon, heading, seg_len, nz,    \
                status_dev, traj_lat, traj_lon, flags, j_resume, any_resume)
            // two slabs of the fp32 grid in shared memory: worth it when a block runs several steps and the ring
            // leaves room for >= 8 blocks per SM
            const size_t ring = 2 * (size_t)sc->view.slab32_stride * sizeof(float2);
            const bool stage = lean && ctx->stage_slabs && nsteps >= 4 && ring <= 24 * 1024;
            // (K1's specialised variant also takes the step length as a constant)
            const bool defboat = fast::is_default_boat(ctx->params) && sc->view.dt_uniform32 > 0.0f;
            ctx->last_variant = !lean ? PMCTS_VARIANT_FULL_OUTPUT : stage ? PMCTS_VARIANT_STAGED
                                : defboat ? PMCTS_VARIANT_DEFAULT_BOAT : PMCTS_VARIANT_GENERIC;
#define PMCTS_LAUNCH_STAGED(NOISE)                                                                          \
            step_fast_staged_kernel<NOISE><<<grid_for(n), kBlock, ring, ctx->stream>>>(                      \
                sc->view, ctx->params, n, nsteps, k, lat, lon, heading, seg_len, nz, status_dev, j_resume,   \
                any_resume)
            if (stage && nz.mode == PMCTS_NOISE_PHILOX) PMCTS_LAUNCH_STAGED(PMCTS_NOISE_PHILOX);
            else if (stage && nz.mode == PMCTS_NOISE_INJECTED) PMCTS_LAUNCH_STAGED(PMCTS_NOISE_INJECTED);
            else if (stage) PMCTS_LAUNCH_STAGED(PMCTS_NOISE_NONE);
            else if (lean && defboat && nz.mode == PMCTS_NOISE_PHILOX) PMCTS_LAUNCH_STEP(PMCTS_NOISE_PHILOX, true, true);
            else if (lean && defboat && nz.mode == PMCTS_NOISE_INJECTED) PMCTS_LAUNCH_STEP(PMCTS_NOISE_INJECTED, true, true);
            else if (lean && defboat) PMCTS_LAUNCH_STEP(PMCTS_NOISE_NONE, true, true);
            else if (lean && nz.mode == PMCTS_NOISE_PHILOX) PMCTS_LAUNCH_STEP(PMCTS_NOISE_PHILOX, true);
            else if (lean && nz.mode == PMCTS_NOISE_INJECTED) PMCTS_LAUNCH_STEP(PMCTS_NOISE_INJECTED, true);
            else if (lean) PMCTS_LAUNCH_STEP(PMCTS_NOISE_NONE, true);
            else if (nz.mode == PMCTS_NOISE_PHILOX) PMCTS_LAUNCH_STEP(PMCTS_NOISE_PHILOX, false);
            else if (nz.mode == PMCTS_NOISE_INJECTED) PMCTS_LAUNCH_STEP(PMCTS_NOISE_INJECTED, false);
            else PMCTS_LAUNCH_STEP(PMCTS_NOISE_NONE, false);
#undef PMCTS_LAUNCH_STEP
#undef PMCTS_LAUNCH_STAGED
        }
        CUDA_TRY(cudaGetLastError());
        {
            LaunchTimer lt(ctx, PMCTS_KERNEL_STEP);
            step_batch_kernel<<<grid_for(n), kBlock, 0, ctx->stream>>>(
                sc->view, ctx->params, n, nsteps, k, lat, lon, prev_lat, prev_lon, heading, seg_len, nz,
                status_dev, traj_lat, traj_lon, flags, j_resume, any_resume);
        }
    } else {
        ctx->last_variant = PMCTS_VARIANT_LITERAL;
        LaunchTimer lt(ctx, PMCTS_KERNEL_STEP);
        step_batch_kernel<<<grid_for(n), kBlock, 0, ctx->stream>>>(
            sc->view, ctx->params, n, nsteps, k, lat, lon, prev_lat, prev_lon, heading, seg_len, nz, status,
            traj_lat, traj_lon, flags, nullptr, nullptr);
    }
    CUDA_TRY(cudaGetLastError());
    if (defer) {
        CUDA_TRY(cudaEventRecord(slot->main_done, ctx->stream));
        CUDA_TRY(cudaStreamWaitEvent(ctx->side_stream, slot->main_done, 0));
        if ((rc = st.copy_back(ctx->side_stream))) return rc;
        CUDA_TRY(cudaEventRecord(slot->out_done, ctx->side_stream));
        slot->pending = true;
        slot->used = true;
        slot->seq = ctx->step_seq++;
        return PMCTS_OK;
    }
    return st.finish();
}

static int rollout_impl(pmcts_ctx *ctx, int n_scen, const int32_t *scens, int prefix_shared, int64_t n_leaf,
                        int replicas, const double root[3], const double dest[2], double tmin,
                        const double *prefix, const int32_t *prefix_len, int prefix_stride,
                        const pmcts_noise *noise, double *reward, double *t_arr, int32_t *steps, uint8_t *status,
                        int32_t *bin, double *frac, double *final_lat, double *final_lon, double *traj_lat,
                        double *traj_lon, int traj_stride, uint32_t *leaf_bins, double *stats, int flags) {
    if (!ctx || !root || !dest || !scens) return fail(PMCTS_ERR_ARG, "NULL argument");
    if (n_scen < 1 || n_scen > kMaxForest) return fail(PMCTS_ERR_ARG, "1 .. %d scenarios per launch", kMaxForest);
    if (n_leaf < 0 || replicas < 1 || prefix_stride < 0 || traj_stride < 0) return fail(PMCTS_ERR_ARG, "bad sizes");
    if ((prefix == nullptr) != (prefix_len == nullptr)) return fail(PMCTS_ERR_ARG, "prefix and prefix_len go together");
    if (n_leaf == 0) return PMCTS_OK;
    DeviceGuard g(ctx);
    Scenario *sc;
    int rc = find_scenario(ctx, scens[0], &sc);
    if (rc) return rc;
    ForestSlabs fs{};
    fs.n_scen = n_scen;
    fs.prefix_shared = prefix_shared;
    bool all_fast = true;
    for (int i = 0; i < n_scen; ++i) {
        Scenario *si;
        if ((rc = find_scenario(ctx, scens[i], &si))) return rc;
        const ScenarioView &x = si->view, &y = sc->view;
        if (x.nT != y.nT || x.nlat != y.nlat || x.nlon != y.nlon || x.lat0 != y.lat0 || x.lon0 != y.lon0 ||
            x.inv_dlat != y.inv_dlat || x.inv_dlon != y.inv_dlon || x.box_lat_lo != y.box_lat_lo ||
            x.box_lat_hi != y.box_lat_hi || x.box_lon_lo != y.box_lon_lo || x.box_lon_hi != y.box_lon_hi ||
            x.tmax != y.tmax || x.kv_lo != y.kv_lo || x.kv_hi != y.kv_hi || si->times_hash != sc->times_hash)
            return fail(PMCTS_ERR_SCENARIO, "scenarios of one launch must share their axes, box and simulator instants "
                                            "(slot %d differs from slot %d)", scens[i], scens[0]);
        fs.s32[i] = x.slab32;
        fs.s64[i] = x.slab64;
        all_fast = all_fast && x.fast_ok;
    }
    const int64_t n = n_leaf * replicas;
    const size_t S = (size_t)n_scen, P = prefix_shared ? 1 : S;
    NoiseView nz = noise_view(noise);
    const int max_steps = sc->view.nT;
    if (nz.mode == PMCTS_NOISE_INJECTED && !nz.z) return fail(PMCTS_ERR_ARG, "injected noise needs z");
    const bool acc = (flags & PMCTS_ROLL_ACCUMULATE) && !(flags & PMCTS_MEM_HOST);
    // the fp32 kernels also need the destination within 90 degrees of every grid point (their
    // sin / cos kernels cover [-pi/2, pi/2] for latD - lat and lonD - lon)
    const ScenarioView &vw = sc->view;
    const bool dest_ok = std::fabs(dest[0]) <= 90.0 && std::fabs(dest[0] - vw.lat0) <= 89.0 &&
                         std::fabs(dest[0] - vw.lat_last) <= 89.0 && std::fabs(dest[1] - vw.lon0) <= 89.0 &&
                         std::fabs(dest[1] - vw.lon_last) <= 89.0;
    const bool fast = fast_eligible(ctx, sc) && all_fast && dest_ok && n < ((int64_t)1 << kRedoShift);
    // PMCTS_ROLL_DEFER_REDO: finisher, fold and -- with host arrays -- the copies back to the caller run on the side
    // stream under the caller's next launch; two slots (work area, staging buffer, events) alternate
    const bool defer = fast && (flags & PMCTS_ROLL_DEFER_REDO);
    pmcts_ctx::Deferred *slot = nullptr;
    if (defer) {
        slot = &ctx->defer[ctx->defer_seq & 1];
        if (!ctx->side_stream) {
            int lo = 0, hi = 0;
            CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            CUDA_TRY(cudaStreamCreateWithPriority(&ctx->side_stream, cudaStreamNonBlocking, hi));
        }
        if (!slot->main_done) {
            CUDA_TRY(cudaEventCreateWithFlags(&slot->main_done, cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&slot->redo_done, cudaEventDisableTiming));
        }
        // the launch that used this slot two calls ago must have finished with it (the caller may have
        // joined it on another stream -- pmcts_set_stream between calls -- so this does not look at `pending`)
        if (slot->used) CUDA_TRY(cudaStreamWaitEvent(ctx->stream, slot->redo_done, 0));
    }
    Stager st(ctx, flags & PMCTS_MEM_HOST, slot ? &slot->stage : nullptr);
    const bool pre_actions = (flags & PMCTS_ROLL_PREFIX_ACTIONS) && prefix;
    const int8_t *prefix8 = pre_actions ? reinterpret_cast<const int8_t *>(prefix) : nullptr;
    const size_t pre_count = P * (size_t)n_leaf * prefix_stride;
    const size_t traj_bytes = al(S * traj_stride * n * 8);
    rc = st.reserve(al(P * n_leaf * prefix_stride * 8) + al(P * n_leaf * 4) + 6 * al(S * n * 8) + 2 * al(S * n * 4) +
                    al(S * n) + (traj_lat ? traj_bytes : 0) + (traj_lon ? traj_bytes : 0) + al(S * n_leaf * 48) +
                    al(S * 64) + (nz.mode == PMCTS_NOISE_INJECTED ? al(S * max_steps * n * 8) : 0));
    if (rc) return rc;
    if ((rc = pre_actions ? st.in(prefix8, pre_count) : st.in(prefix, pre_count)) || (rc = st.in(prefix_len, P * n_leaf)) ||
        (rc = st.out(reward, S * n)) || (rc = st.out(t_arr, S * n)) || (rc = st.out(steps, S * n)) ||
        (rc = st.out(status, S * n)) || (rc = st.out(bin, S * n)) || (rc = st.out(frac, S * n)) ||
        (rc = st.out(final_lat, S * n)) || (rc = st.out(final_lon, S * n)) ||
        (rc = st.out(traj_lat, S * traj_stride * n, 0xFF)) || (rc = st.out(traj_lon, S * traj_stride * n, 0xFF)) ||
        (rc = st.out(leaf_bins, S * n_leaf * 12, acc ? -1 : 0)) || (rc = st.out(stats, S * 8, acc ? -1 : 0)))
        return rc;
    if (nz.mode == PMCTS_NOISE_INJECTED && (rc = st.in(nz.z, S * max_steps * n))) return rc;
    RolloutArgs a{};
    a.n_leaf = n_leaf;
    a.replicas = replicas;
    a.root_k = root[0];
    a.root_lat = root[1];
    a.root_lon = root[2];
    a.dest_lat = dest[0];
    a.dest_lon = dest[1];
    a.tmin = tmin;
    a.prefix = prefix;
    a.prefix_len = prefix_len;
    a.prefix_stride = prefix_stride;
    a.reward = reward;
    a.t_arr = t_arr;
    a.steps = steps;
    a.status = status;
    a.bin = bin;
    a.frac = frac;
    a.final_lat = final_lat;
    a.final_lon = final_lon;
    a.traj_lat = traj_lat;
    a.traj_lon = traj_lon;
    a.traj_stride = traj_stride;
    a.leaf_bins = leaf_bins;
    a.flags = flags;
    // work area: counter | partial statistics [n_scen][kStatsCopies][8] | redo list | action paths expanded to degrees
    const size_t part_bytes = al(S * kStatsCopies * 8 * sizeof(double));
    const size_t redo_bytes = al(S * n * sizeof(uint32_t));
    const size_t work_bytes = part_bytes + redo_bytes + (pre_actions ? al(pre_count * sizeof(double)) : 0);
    char *work = nullptr;
    if (defer) {
        if (256 + work_bytes > slot->cap) {
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            CUDA_TRY(cudaStreamSynchronize(ctx->side_stream));
            if (slot->work) CUDA_TRY(cudaFree(slot->work));
            slot->work = nullptr;
            slot->cap = 0;
            const size_t cap = 256 + work_bytes + work_bytes / 4;
            CUDA_TRY(cudaMalloc((void **)&slot->work, cap));
            slot->cap = cap;
        }
        work = slot->work;
    } else {
        if ((rc = ensure_work(ctx, work_bytes))) return rc;
        work = ctx->work;
    }
    if (pre_actions) {
        // (in the launch's own work area: the deferred fp64 finisher of this launch reads the paths too, after the next
        // launch has been enqueued)
        double *expanded = (double *)(work + 256 + part_bytes + redo_bytes);
        LaunchTimer lt(ctx);
        expand_prefix_kernel<<<(unsigned)((pre_count + 255) / 256), 256, 0, ctx->stream>>>(prefix8, expanded, pre_count);
        CUDA_TRY(cudaGetLastError());
        a.prefix = expanded;
    }
    double *stats_part = (double *)(work + 256);
    // (counter and partial sums are neighbours in the work area: one memset clears both for a mixed-precision launch)
    const size_t stats_part_bytes = S * kStatsCopies * 8 * sizeof(double);
    if (stats) {
        a.stats = stats_part;
        if (fast) CUDA_TRY(cudaMemsetAsync(work, 0, 256 + stats_part_bytes, ctx->stream));
        else CUDA_TRY(cudaMemsetAsync(stats_part, 0, stats_part_bytes, ctx->stream));
    }
    const dim3 grid_all(grid_for(n), n_scen);
    cudaStream_t tail_stream = ctx->stream;  // where the fp64 finisher and the statistics fold run
    if (fast) {
        fast::prepare_rollout(vw, a);
        a.redo_count = (uint32_t *)work;
        a.redo_list = (uint32_t *)(work + 256 + part_bytes);
        ctx->last_fast = slot;  // (NULL for a launch that counts in ctx->work)
        if (!stats) CUDA_TRY(cudaMemsetAsync(a.redo_count, 0, sizeof(uint32_t), ctx->stream));
        {
            LaunchTimer lt(ctx, PMCTS_KERNEL_ROLLOUT);
            const bool lean = !reward && !t_arr && !steps && !status && !bin && !frac && !final_lat && !final_lon &&
                              !traj_lat && !traj_lon;
            // the reference's own boat: its 29 constants as immediates (fast::defboat)
            const bool defboat = fast::is_default_boat(ctx->params);
#define PMCTS_LAUNCH_ROLL(NOISE, LEAN, ...) \
            rollout_fast_kernel<NOISE, LEAN, ##__VA_ARGS__><<<grid_all, kBlock, 0, ctx->stream>>>(vw, ctx->params, nz, a, fs)
            // turbo: tree mode from a root that can start a step, on a scenario whose terminal box implies every
            // range check of the loop (fast::turbo_ok)
            const bool turbo = defboat && fast::turbo_ok(vw, a);
            const bool first_only = !(flags & PMCTS_ROLL_REPLAY_FULL_PREFIX);
            ctx->last_variant = !lean ? PMCTS_VARIANT_FULL_OUTPUT : turbo ? PMCTS_VARIANT_TURBO
                                : defboat ? PMCTS_VARIANT_DEFAULT_BOAT : PMCTS_VARIANT_GENERIC;
            // turbo launches run two rollouts per thread on the packed FP32 pipe (PMCTS_OPT_PACKED_FP32, default on)
            const dim3 grid_pair(grid_for((n + 1) / 2), n_scen);
#define PMCTS_LAUNCH_PAIR(NOISE) rollout_fast2_kernel<NOISE><<<grid_pair, kBlock, 0, ctx->stream>>>(vw, ctx->params, nz, a, fs)
            const bool packed = lean && turbo && ctx->packed_fp32;
            if (packed) ctx->last_variant = PMCTS_VARIANT_TURBO_PACKED;
            if (packed && nz.mode == PMCTS_NOISE_PHILOX && ctx->packed_fp32 == 5)
                rollout_fast2_kernel<PMCTS_NOISE_PHILOX, 5><<<grid_pair, kBlock, 0, ctx->stream>>>(vw, ctx->params, nz, a, fs);
            else if (packed && nz.mode == PMCTS_NOISE_PHILOX && ctx->packed_fp32 == 6)
                rollout_fast2_kernel<PMCTS_NOISE_PHILOX, 6><<<grid_pair, kBlock, 0, ctx->stream>>>(vw, ctx->params, nz, a, fs);
            else
            if (packed && nz.mode == PMCTS_NOISE_PHILOX) PMCTS_LAUNCH_PAIR(PMCTS_NOISE_PHILOX);
            else if (packed && nz.mode == PMCTS_NOISE_INJECTED) PMCTS_LAUNCH_PAIR(PMCTS_NOISE_INJECTED);
            else if (packed) PMCTS_LAUNCH_PAIR(PMCTS_NOISE_NONE);
            else
#undef PMCTS_LAUNCH_PAIR
            // (turbo, only the first action replayed -- the reference's behaviour: the loop without the prefix test)
            if (lean && turbo && first_only && nz.mode == PMCTS_NOISE_PHILOX) PMCTS_LAUNCH_ROLL(PMCTS_NOISE_PHILOX, true, true, true, true);
            else if (lean && turbo && first_only && nz.mode == PMCTS_NOISE_INJECTED) PMCTS_LAUNCH_ROLL(PMCTS_NOISE_INJECTED, true, true, true, true);
            else if (lean && turbo && first_only) PMCTS_LAUNCH_ROLL(PMCTS_NOISE_NONE, true, true, true, true);
            else if (lean && turbo && nz.mode == PMCTS_NOISE_PHILOX) PMCTS_LAUNCH_ROLL(PMCTS_NOISE_PHILOX, true, true, true);
            else if (lean && turbo && nz.mode == PMCTS_NOISE_INJECTED) PMCTS_LAUNCH_ROLL(PMCTS_NOISE_INJECTED, true, true, true);
            else if (lean && turbo) PMCTS_LAUNCH_ROLL(PMCTS_NOISE_NONE, true, true, true);
            else if (lean && defboat && nz.mode == PMCTS_NOISE_PHILOX) PMCTS_LAUNCH_ROLL(PMCTS_NOISE_PHILOX, true, true);
            else if (lean && defboat && nz.mode == PMCTS_NOISE_INJECTED) PMCTS_LAUNCH_ROLL(PMCTS_NOISE_INJECTED, true, true);
            else if (lean && defboat) PMCTS_LAUNCH_ROLL(PMCTS_NOISE_NONE, true, true);
            else if (lean && nz.mode == PMCTS_NOISE_PHILOX) PMCTS_LAUNCH_ROLL(PMCTS_NOISE_PHILOX, true);
            else if (lean && nz.mode == PMCTS_NOISE_INJECTED) PMCTS_LAUNCH_ROLL(PMCTS_NOISE_INJECTED, true);
            else if (lean) PMCTS_LAUNCH_ROLL(PMCTS_NOISE_NONE, true);
            else if (nz.mode == PMCTS_NOISE_PHILOX) PMCTS_LAUNCH_ROLL(PMCTS_NOISE_PHILOX, false);
            else if (nz.mode == PMCTS_NOISE_INJECTED) PMCTS_LAUNCH_ROLL(PMCTS_NOISE_INJECTED, false);
            else PMCTS_LAUNCH_ROLL(PMCTS_NOISE_NONE, false);
#undef PMCTS_LAUNCH_ROLL
        }
        CUDA_TRY(cudaGetLastError());
        if (defer) {
            CUDA_TRY(cudaEventRecord(slot->main_done, ctx->stream));
            CUDA_TRY(cudaStreamWaitEvent(ctx->side_stream, slot->main_done, 0));
            tail_stream = ctx->side_stream;
        }
        a.work_list = a.redo_list;
        a.work_count = a.redo_count;
        {
            LaunchTimer lt(ctx, PMCTS_KERNEL_FINISHER, tail_stream);
            const int grid = (int)std::min<int64_t>((int64_t)grid_for(n) * n_scen, 2 * 148);
            rollout_batch_kernel<<<grid, kBlock, 0, tail_stream>>>(vw, ctx->params, nz, a, fs);
        }
    } else {
        ctx->last_variant = PMCTS_VARIANT_LITERAL;
        LaunchTimer lt(ctx, PMCTS_KERNEL_ROLLOUT);
        rollout_batch_kernel<<<grid_all, kBlock, 0, ctx->stream>>>(vw, ctx->params, nz, a, fs);
    }
    CUDA_TRY(cudaGetLastError());
    if (stats) {
        LaunchTimer lt(ctx, PMCTS_KERNEL_OTHER, tail_stream);
        stats_finish_kernel<<<n_scen, kStatsCopies, 0, tail_stream>>>(stats_part, stats);
        CUDA_TRY(cudaGetLastError());
    }
    if (defer) {
        // host arrays: the copies back follow the finisher and the fold on the side stream; the call returns without
        // waiting for them (pmcts_wait)
        if ((rc = st.copy_back(ctx->side_stream))) return rc;
        CUDA_TRY(cudaEventRecord(slot->redo_done, ctx->side_stream));
        slot->pending = true;
        slot->used = true;
        slot->seq = ctx->defer_seq++;
        return PMCTS_OK;
    }
    return st.finish();
}

int pmcts_wait(pmcts_ctx *ctx, int keep_last) {
    if (!ctx) return fail(PMCTS_ERR_ARG, "ctx is NULL");
    if (keep_last < 0) keep_last = 0;
    DeviceGuard g(ctx);
    for (auto &d : ctx->defer) {
        if (d.pending && d.seq + (uint64_t)keep_last < ctx->defer_seq) {
            CUDA_TRY(cudaEventSynchronize(d.redo_done));
            d.pending = false;
        }
    }
    for (auto &d : ctx->step_defer) {
        if (d.pending && d.seq + (uint64_t)keep_last < ctx->step_seq) {
            CUDA_TRY(cudaEventSynchronize(d.out_done));
            d.pending = false;
        }
    }
    return PMCTS_OK;
}

int pmcts_join(pmcts_ctx *ctx, int keep_last) {
    if (!ctx) return fail(PMCTS_ERR_ARG, "ctx is NULL");
    if (keep_last < 0) keep_last = 0;
    DeviceGuard g(ctx);
    for (auto &d : ctx->defer) {
        if (d.pending && d.seq + (uint64_t)keep_last < ctx->defer_seq) {
            CUDA_TRY(cudaStreamWaitEvent(ctx->stream, d.redo_done, 0));
            d.pending = false;
        }
    }
    return PMCTS_OK;
}

int pmcts_rollout_batch(pmcts_ctx *ctx, int scen, int64_t n_leaf, int replicas, const double root[3],
                        const double dest[2], double tmin, const double *prefix,
                        const int32_t *prefix_len, int prefix_stride, const pmcts_noise *noise,
                        double *reward, double *t_arr, int32_t *steps, uint8_t *status, int32_t *bin,
                        double *frac, double *final_lat, double *final_lon, double *traj_lat,
                        double *traj_lon, int traj_stride, uint32_t *leaf_bins, double *stats,
                        int flags) {
    const int32_t one = scen;
    return rollout_impl(ctx, 1, &one, 1, n_leaf, replicas, root, dest, tmin, prefix, prefix_len, prefix_stride, noise,
                        reward, t_arr, steps, status, bin, frac, final_lat, final_lon, traj_lat, traj_lon, traj_stride,
                        leaf_bins, stats, flags);
}

int pmcts_rollout_forest(pmcts_ctx *ctx, int n_scen, const int32_t *scen, int64_t n_leaf, int replicas,
                         const double root[3], const double dest[2], double tmin, const double *prefix,
                         const int32_t *prefix_len, int prefix_stride, int prefix_shared,
                         const pmcts_noise *noise, double *reward, double *t_arr, int32_t *steps,
                         uint8_t *status, int32_t *bin, uint32_t *leaf_bins, double *stats, int flags) {
    return rollout_impl(ctx, n_scen, scen, prefix_shared ? 1 : 0, n_leaf, replicas, root, dest, tmin, prefix,
                        prefix_len, prefix_stride, noise, reward, t_arr, steps, status, bin, nullptr, nullptr, nullptr,
                        nullptr, nullptr, 0, leaf_bins, stats, flags);
}

int pmcts_backup(pmcts_ctx *ctx, uint32_t *table, int64_t n_nodes, int n_scen, const int32_t *parent,
                 const int8_t *arm, int64_t n_leaf, const int32_t *leaf_node, const int8_t *leaf_slot,
                 const uint32_t *leaf_bins, int flags) {
    if (!ctx || !table || !parent || !arm || !leaf_node || !leaf_slot || !leaf_bins)
        return fail(PMCTS_ERR_ARG, "NULL argument");
    if (flags & PMCTS_MEM_HOST) return fail(PMCTS_ERR_ARG, "pmcts_backup works on a device-resident table");
    if (n_scen < 1 || n_scen > 65535 || n_nodes < 1) return fail(PMCTS_ERR_ARG, "bad sizes");
    if (n_leaf <= 0) return PMCTS_OK;
    DeviceGuard g(ctx);
    {
        LaunchTimer lt(ctx, PMCTS_KERNEL_BACKUP);
        const int64_t total = n_leaf * n_scen;
        const unsigned grid = (unsigned)std::min<int64_t>((total + 255) / 256, 2 * ctx->sm_count);
        // (a leaf's counts in three vector loads when the array starts on a 16-byte boundary -- whatever the width,
        //  12 counts then start on a multiple of their vector)
        const bool vector = (uintptr_t)leaf_bins % 16 == 0;
        if (flags & PMCTS_BACKUP_BINS_U8)
            backup_kernel<uint8_t><<<grid, 256, 0, ctx->stream>>>(table, n_nodes, parent, arm, n_leaf, n_scen, leaf_node,
                                                                 leaf_slot, (const uint8_t *)leaf_bins, vector);
        else if (flags & PMCTS_BACKUP_BINS_U16)
            backup_kernel<uint16_t><<<grid, 256, 0, ctx->stream>>>(table, n_nodes, parent, arm, n_leaf, n_scen, leaf_node,
                                                                  leaf_slot, (const uint16_t *)leaf_bins, vector);
        else
            backup_kernel<uint32_t><<<grid, 256, 0, ctx->stream>>>(table, n_nodes, parent, arm, n_leaf, n_scen, leaf_node,
                                                                  leaf_slot, leaf_bins, vector);
    }
    CUDA_TRY(cudaGetLastError());
    return PMCTS_OK;
}

int pmcts_pack_bins(pmcts_ctx *ctx, const uint32_t *bins, int64_t count, void *out, int width) {
    if (!ctx || !bins || !out) return fail(PMCTS_ERR_ARG, "NULL argument");
    if (width != 1 && width != 2) return fail(PMCTS_ERR_ARG, "width must be 1 or 2 bytes");
    if (count <= 0) return PMCTS_OK;
    DeviceGuard g(ctx);
    {
        LaunchTimer lt(ctx);
        const unsigned grid = (unsigned)std::min<int64_t>((count + 255) / 256, 2 * ctx->sm_count);
        if (width == 1) pack_bins_kernel<uint8_t><<<grid, 256, 0, ctx->stream>>>(bins, (uint8_t *)out, (size_t)count);
        else pack_bins_kernel<uint16_t><<<grid, 256, 0, ctx->stream>>>(bins, (uint16_t *)out, (size_t)count);
    }
    CUDA_TRY(cudaGetLastError());
    return PMCTS_OK;
}

int pmcts_hist_stats(pmcts_ctx *ctx, const uint32_t *table, int64_t n_nodes, double *mean,
                     uint32_t *visits, int flags) {
    if (!ctx || !table) return fail(PMCTS_ERR_ARG, "NULL argument");
    if (n_nodes <= 0) return PMCTS_OK;
    DeviceGuard g(ctx);
    Stager st(ctx, flags & PMCTS_MEM_HOST);
    int rc = st.reserve(al(n_nodes * 96 * 4) + al(n_nodes * 8 * 8) + al(n_nodes * 4));
    if (rc) return rc;
    if ((rc = st.in(table, (size_t)n_nodes * 96)) || (rc = st.out(mean, (size_t)n_nodes * 8)) ||
        (rc = st.out(visits, n_nodes)))
        return rc;
    {
        LaunchTimer lt(ctx);
        hist_stats_kernel<<<grid_for(n_nodes * 8), kBlock, 0, ctx->stream>>>(table, n_nodes, mean, visits);
    }
    CUDA_TRY(cudaGetLastError());
    return st.finish();
}

}  // extern "C"
