// pmcts_b200 mixed-precision rollout, two rollouts per thread on Blackwell's packed FP32 pipe.
//
// The throughput form of K2 (rollout_fast_kernel<.., lean, default boat, turbo>) is bound by instruction issue, and
// more than half of what it issues are FP32 multiplies and adds (pmcts_fast.cuh's polynomials, dot products and
// identities).  sm_100 executes them two at a time: FFMA2 / FMUL2 / FADD2 take a 64-bit register pair per operand --
// or one register or an immediate broadcast to both halves -- and issue as ONE instruction.  Here a thread carries
// two rollouts (2 t and 2 t + 1 of the launch) and every FP32 quantity of pmcts_fast.cuh becomes a float2 holding
// the value of both: same formulas, same constants, same order of operations per rollout -- the halves never mix.
// What has no packed form runs per half as before: MUFU (rsqrt, rcp, lg2, sin, cos), min / max / select, the fp64
// position and its conversions, the grid gathers and Philox.
//
// Control flow: the two rollouts step together.  A rollout that has ended (arrived, terminal, or left to the fp64
// kernel) stays in its half with its step length forced to zero -- move_increments(0) leaves the position bit for
// bit where it was -- and its decisions masked, until the other one ends too.
//
// Device-only (packed intrinsics); decisions and histograms are checked against the scalar kernels and the oracle
// in tests/test_gpu_parity.py.  Only the turbo variant exists in this form (see fast::turbo_ok): tree mode, the
// reference's boat as immediates, equal steps, terminal box inside the grid.
#pragma once
#include "pmcts_fast.cuh"

#if defined(__CUDACC__)
namespace pmcts {
namespace fast2 {

using fast::defboat::cos_pmax;
using fast::defboat::cos_pmin;

typedef float2 f2;

__device__ __forceinline__ f2 mk(float x, float y) { return make_float2(x, y); }
__device__ __forceinline__ f2 bc(float c) { return make_float2(c, c); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, float c) { return __ffma2_rn(a, b, bc(c)); }
__device__ __forceinline__ f2 fma2(f2 a, float b, float c) { return __ffma2_rn(a, bc(b), bc(c)); }
__device__ __forceinline__ f2 fma2(f2 a, float b, f2 c) { return __ffma2_rn(a, bc(b), c); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 mul2(f2 a, float b) { return __fmul2_rn(a, bc(b)); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 neg2(f2 a) { return mk(-a.x, -a.y); }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { return __fadd2_rn(a, neg2(b)); }
__device__ __forceinline__ f2 rsqrt2(f2 a) { return mk(fast::rsqrt_approx(a.x), fast::rsqrt_approx(a.y)); }
__device__ __forceinline__ f2 rcp2(f2 a) { return mk(fast::rcp_approx(a.x), fast::rcp_approx(a.y)); }

// ---- the polynomial kernels of pmcts_fast.cuh, both halves at once ---------------------------------------------
__device__ __forceinline__ f2 sinc_poly(f2 y) {
    f2 p = fma2(bc(2.605224902e-06f), y, -1.980907528e-04f);
    p = fma2(p, y, 8.333051063e-03f);
    p = fma2(p, y, -1.666665799e-01f);
    return fma2(p, y, 1.0f);
}
__device__ __forceinline__ f2 cosc_poly(f2 y) {
    f2 p = fma2(bc(2.633458343e-07f), y, -2.477660893e-05f);
    p = fma2(p, y, 1.388868840e-03f);
    p = fma2(p, y, -4.166666179e-02f);
    return fma2(p, y, 0.5f);
}
__device__ __forceinline__ f2 sinc_small(f2 y) { return fma2(fma2(bc(8.3333333e-03f), y, -1.6666667e-01f), y, 1.0f); }
__device__ __forceinline__ f2 cosc_small(f2 y) { return fma2(fma2(bc(1.3888889e-03f), y, -4.1666667e-02f), y, 0.5f); }
__device__ __forceinline__ f2 asinc_small(f2 y) {
    return fma2(fma2(fma2(bc(4.4642857e-02f), y, 7.5e-02f), y, 1.6666667e-01f), y, 1.0f);
}
// fast::atan2_pos
__device__ __forceinline__ f2 atan2_pos(f2 s, f2 c) {
    const f2 ac = mk(fabsf(c.x), fabsf(c.y));
    const f2 mx = mk(fmaxf(s.x, ac.x), fmaxf(s.y, ac.y)), mn = mk(fminf(s.x, ac.x), fminf(s.y, ac.y));
    const f2 t = mul2(mn, rcp2(mx));
    const f2 y = mul2(t, t);
    f2 p = fma2(bc(2.903537903e-03f), y, -1.628295011e-02f);
    p = fma2(p, y, 4.303927486e-02f);
    p = fma2(p, y, -7.533668294e-02f);
    p = fma2(p, y, 1.065467381e-01f);
    p = fma2(p, y, -1.420713276e-01f);
    p = fma2(p, y, 1.999305400e-01f);
    p = fma2(p, y, -3.333309395e-01f);
    p = fma2(p, y, 1.0f);
    f2 r = mul2(t, p);
    if (s.x > ac.x) r.x = 1.57079632679f - r.x;
    if (s.y > ac.y) r.y = 1.57079632679f - r.y;
    if (c.x < 0.0f) r.x = 3.14159265359f - r.x;
    if (c.y < 0.0f) r.y = 3.14159265359f - r.y;
    return r;
}

// fast::polar_centred<true>: the reference's polar about the centre of its domain, 20 packed FMAs
__device__ __forceinline__ f2 polar_centred(f2 x, f2 y) {
    namespace d = fast::defboat;
    f2 c0 = fma2(bc(d::pc30), x, d::pc24);
    c0 = fma2(c0, x, d::pc18); c0 = fma2(c0, x, d::pc12); c0 = fma2(c0, x, d::pc6); c0 = fma2(c0, x, d::pc0);
    f2 c1 = fma2(bc(d::pc25), x, d::pc19);
    c1 = fma2(c1, x, d::pc13); c1 = fma2(c1, x, d::pc7); c1 = fma2(c1, x, d::pc1);
    f2 c2 = fma2(bc(d::pc20), x, d::pc14);
    c2 = fma2(c2, x, d::pc8); c2 = fma2(c2, x, d::pc2);
    f2 c3 = fma2(bc(d::pc15), x, d::pc9);
    c3 = fma2(c3, x, d::pc3);
    const f2 c4 = fma2(bc(d::pc10), x, d::pc4);
    f2 acc = fma2(bc(d::pc5), y, c4);
    acc = fma2(acc, y, c3); acc = fma2(acc, y, c2); acc = fma2(acc, y, c1);
    return fma2(acc, y, c0);
}

// fast::boat_speed<true>
__device__ __forceinline__ f2 boat_speed(f2 u, f2 v, f2 sh, f2 ch) {
    namespace d = fast::defboat;
    const f2 m2 = fma2(u, u, mul2(v, v));
    const f2 rm = rsqrt2(m2);
    f2 mag = mul2(m2, rm), du = mul2(u, rm), dv = mul2(v, rm);
    if (!(m2.x > 0.0f)) { du.x = 0.0f; dv.x = 1.0f; mag.x = 0.0f; }  // atan2(0, 0) = 0: a calm wind "blows north"
    if (!(m2.y > 0.0f)) { du.y = 0.0f; dv.y = 1.0f; mag.y = 0.0f; }
    const f2 cosp = neg2(fma2(du, sh, mul2(dv, ch)));
    const f2 cr = fma2(du, ch, neg2(mul2(dv, sh)));
    const f2 sinp = mk(fabsf(cr.x), fabsf(cr.y));
    const f2 p = mul2(atan2_pos(sinp, cosp), 57.2957795131f);
    const f2 w = mk(fminf(mag.x, d::wind_max), fminf(mag.y, d::wind_max));
    const f2 pe = mk(fminf(fmaxf(p.x, d::pmin), d::pmax), fminf(fmaxf(p.y, d::pmin), d::pmax));
    f2 sp = polar_centred(mul2(add2(pe, bc(-d::p_mid)), d::p_ihalf), fma2(w, d::w_ihalf, -1.0f));
    if (p.x < d::pmin) sp.x *= cos_pmin * fast::rcp_approx(cosp.x);
    else if (p.x > d::pmax) sp.x *= cos_pmax * fast::rcp_approx(cosp.x);
    if (p.y < d::pmin) sp.y *= cos_pmin * fast::rcp_approx(cosp.y);
    else if (p.y > d::pmax) sp.y *= cos_pmax * fast::rcp_approx(cosp.y);
    return sp;
}

struct Move2 {
    f2 sin_dlat, sin_dlon, h_dlon, cm1_dlat, dlat, dlon, s_new, c_new;
};

// fast::move_increments
__device__ __forceinline__ Move2 move_increments(f2 d, f2 sh, f2 ch, f2 s, f2 c) {
    Move2 m;
    const f2 d2 = mul2(d, d);
    const f2 sd = mul2(d, sinc_small(d2));
    const f2 e = mul2(d2, cosc_small(d2));  // 1 - cos(d)
    const f2 a = mul2(sd, ch), b = mul2(sd, sh);
    const f2 cd = sub2(bc(1.0f), e);
    const f2 A = fma2(c, cd, neg2(mul2(s, a)));
    const f2 b2 = mul2(b, b);
    const f2 n2 = fma2(A, A, b2);
    const f2 rH = rsqrt2(n2);
    const f2 H = mul2(n2, rH);
    const f2 g = mul2(b2, rcp2(add2(H, A)));
    m.sin_dlon = mul2(b, rH);
    m.s_new = fma2(s, cd, mul2(c, a));
    m.c_new = H;
    m.h_dlon = mul2(g, rH);
    m.sin_dlat = fma2(neg2(s), g, a);
    m.cm1_dlat = fma2(c, g, neg2(e));
    m.dlat = mul2(m.sin_dlat, asinc_small(mul2(m.sin_dlat, m.sin_dlat)));
    m.dlon = mul2(m.sin_dlon, asinc_small(mul2(m.sin_dlon, m.sin_dlon)));
    return m;
}

struct DestRel2 {
    f2 sin_dlat, sin_dlon, h_dlon, x_dlat;
};

// fast::dest_rel: the differences in fp64 per half, the series packed
__device__ __forceinline__ DestRel2 dest_rel(double lat0, double lon0, double lat1, double lon1, double dlat, double dlon) {
    DestRel2 r;
    const f2 xa = mk((float)((dlat - lat0) * kDegToRad), (float)((dlat - lat1) * kDegToRad));
    const f2 xo = mk((float)((dlon - lon0) * kDegToRad), (float)((dlon - lon1) * kDegToRad));
    const f2 ya = mul2(xa, xa), yo = mul2(xo, xo);
    r.sin_dlat = mul2(xa, sinc_poly(ya));
    r.sin_dlon = mul2(xo, sinc_poly(yo));
    r.h_dlon = mul2(yo, cosc_poly(yo));
    r.x_dlat = xa;
    return r;
}

// fast::bearing_unit
__device__ __forceinline__ void bearing_unit(const DestRel2 &r, f2 s, float cD, f2 &sh, f2 &ch) {
    const f2 x = mul2(r.sin_dlon, cD);
    const f2 y = fma2(mul2(s, cD), r.h_dlon, r.sin_dlat);
    const f2 n2 = fma2(x, x, mul2(y, y));
    const f2 rn = rsqrt2(n2);
    sh = mul2(x, rn);
    ch = mul2(y, rn);
    if (!(n2.x > 0.0f)) { sh.x = 0.0f; ch.x = 1.0f; }  // atan2(0, 0) = 0
    if (!(n2.y > 0.0f)) { sh.y = 0.0f; ch.y = 1.0f; }
}

// Simulator.getWind on the time-blended fp32 slab (fast::wind_at): the two longitude blends of a rollout run on its
// (u, v) pair as loaded; the latitude blend lands in the (rollout 0, rollout 1) pairs everything else works on.
struct Cell {
    int ia, io;
};
__device__ __forceinline__ void wind_row_pair(const ScenarioView &sc, int k, Cell g, float wo, f2 &t0, f2 &t1) {
    const int cell = k * sc.slab32_stride + (g.ia * sc.nlon + g.io);
    const float2 *row0 = sc.slab32 + cell;
    const float2 *row1 = sc.slab32 + (cell + sc.nlon);
    const float2 c00 = __ldg(row0), c01 = __ldg(row0 + 1), c10 = __ldg(row1), c11 = __ldg(row1 + 1);
    t0 = fma2(sub2(c01, c00), wo, c00);
    t1 = fma2(sub2(c11, c10), wo, c10);
}

// Result of the pair: what rollout_fast_kernel's epilogue needs per half.
struct Pair {
    int st[2], nstep[2], why[2], at[2], k[2];
    float frac[2];
};

// fast::rollout_one<kNoise, lean, default boat, turbo> for the rollouts r0 = 2 t and r0 + 1 (live[h]: the half holds a
// rollout of the launch).  leaf[h] = r / replicas.
template <int kNoise>
__device__ __forceinline__ void rollout_pair(const ScenarioView &sc, const BoatParams &bp, const NoiseView &nz,
                                             const RolloutArgs &a, int64_t n, int64_t r0, const int64_t leaf[2],
                                             const bool live[2], Pair &o) {
    namespace d = fast::defboat;
    using namespace fast;
    const float sD = a.sin_dest_lat, cD = a.cos_dest_lat;
    const RootPre &rp = a.root_pre;
    bool act[2];
    int plen[2], n_fixed[2];
    const double *pre[2];
    double fixed_heading[2];
    double lat[2], lon[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        pre[h] = a.prefix ? a.prefix + (size_t)leaf[h] * a.prefix_stride : nullptr;
        fixed_heading[h] = (live[h] && pre[h] && a.prefix_stride > 0) ? pre[h][0] : 0.0;
        plen[h] = (live[h] && a.prefix_len) ? a.prefix_len[leaf[h]] : 0;
        n_fixed[h] = (a.flags & PMCTS_ROLL_REPLAY_FULL_PREFIX) ? plen[h] : imin(plen[h], 1);
        o.st[h] = PMCTS_ST_OK;
        o.nstep[h] = 0;
        o.why[h] = 0;
        o.at[h] = 0;
        o.frac[h] = 0.0f;
        o.k[h] = (int)a.root_k;
        lat[h] = a.root_lat;
        lon[h] = a.root_lon;
        act[h] = live[h];
        if (live[h] && plen[h] == 0) {
            o.st[h] = PMCTS_ST_DEGENERATE;  // worker.py:264-265 on prevState == state (ZeroDivisionError)
            act[h] = false;
        } else if (live[h] && (a.root_lat == 0.0 || a.root_lon == 0.0)) {
            o.why[h] = kWhyDegenerate;  // ya == 0 in worker.py:396: left to the literal kernel
            act[h] = false;
        }
    }
    f2 sA = bc(rp.s), cA = bc(rp.c);
    DestRel2 rd;
    rd.sin_dlat = bc(rp.sin_dlat); rd.sin_dlon = bc(rp.sin_dlon); rd.h_dlon = bc(rp.h_dlon); rd.x_dlat = bc(rp.x_dlat);
    Cell g[2] = {{rp.ia, rp.io}, {rp.ia, rp.io}};
    f2 wa = bc(rp.wa), wo = bc(rp.wo), fa = bc(rp.fa), fo = bc(rp.fo);
    // noise: the four normals of each half's current Philox block, as (half 0, half 1) pairs in a shift register
    f2 z0 = bc(0.0f), z1 = z0, z2 = z0, z3 = z0;
    const int n_fixed_max = imax(n_fixed[0], n_fixed[1]);
    const float box_m = d::margin_box * sc.box_margin_scale;
    f2 t00 = bc(0.0f), t01 = t00, t10 = t00, t11 = t00;
    int j = 0;
    while (act[0] || act[1]) {
        // ---- heading: the leaf's own actions first, then the bearing to the destination -----------------------
        f2 sh, ch;
        bearing_unit(rd, sA, cD, sh, ch);
        if (j < n_fixed_max) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (act[h] && j < n_fixed[h]) {
                    float s1, c1;
                    if (!heading_unit(fixed_heading[h], s1, c1)) {
                        o.why[h] |= kWhyStep;
                        o.nstep[h] = j;
                        act[h] = false;
                    } else {
                        if (h == 0) { sh.x = s1; ch.x = c1; } else { sh.y = s1; ch.y = c1; }
                        if (j + 1 < n_fixed[h]) fixed_heading[h] = pre[h][j + 1];
                    }
                }
            }
        }
        // ---- noise ------------------------------------------------------------------------------------------------
        f2 z;
        if (kNoise == PMCTS_NOISE_INJECTED) {
            z = mk(live[0] ? (float)nz.z[(size_t)j * n + r0] : 0.0f, live[1] ? (float)nz.z[(size_t)j * n + r0 + 1] : 0.0f);
        } else if (kNoise == PMCTS_NOISE_NONE) {
            z = bc(0.0f);
        } else {
            if ((j & 3) == 0) {
                uint32_t x[2][4];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint64_t boat = nz.id0 + (uint64_t)(r0 + h);
                    fast::philox4x32_10((uint32_t)boat, (uint32_t)(boat >> 32), (uint32_t)(j >> 2), (uint32_t)nz.stream,
                                        (uint32_t)nz.seed, (uint32_t)(nz.seed >> 32), x[h]);
                }
                float a0, a1, b0, b1;
                box_muller(x[0][0], x[0][1], a0, a1);
                box_muller(x[1][0], x[1][1], b0, b1);
                z0 = mk(a0, b0); z1 = mk(a1, b1);
                box_muller(x[0][2], x[0][3], a0, a1);
                box_muller(x[1][2], x[1][3], b0, b1);
                z2 = mk(a0, b0); z3 = mk(a1, b1);
            }
            z = z0;
            z0 = z1; z1 = z2; z2 = z3;
        }
        // ---- one transition (fast::fast_step<.., turbo>) --------------------------------------------------------
        // (an ended half may stand outside the grid -- the terminal box ends a rollout after the step that left it --
        // so only a half that is still stepping gathers; the other keeps whatever it read last)
        if (act[0]) wind_row_pair(sc, o.k[0], g[0], wo.x, t00, t01);
        if (act[1]) wind_row_pair(sc, o.k[1], g[1], wo.y, t10, t11);
        const f2 u = mk(fma_(wa.x, t01.x - t00.x, t00.x), fma_(wa.y, t11.x - t10.x, t10.x));
        const f2 v = mk(fma_(wa.x, t01.y - t00.y, t00.y), fma_(wa.y, t11.y - t10.y, t10.y));
        f2 speed = boat_speed(u, v, sh, ch);
        speed = fma2(z, mul2(speed, bp.sigma_f), speed);  // Boat.addUncertainty
        f2 dstep = mul2(mul2(speed, sc.dt_uniform32), d::inv_radius);
        bool step_ok[2] = {fabsf(dstep.x) <= kMaxStepAngle, fabsf(dstep.y) <= kMaxStepAngle};
        // a half that is not stepping (ended, or about to be handed over) moves by exactly nothing
        if (!(act[0] && step_ok[0])) dstep.x = 0.0f;
        if (!(act[1] && step_ok[1])) dstep.y = 0.0f;
        const Move2 m = move_increments(dstep, sh, ch, sA, cA);
        step_ok[0] = step_ok[0] && fabsf(m.sin_dlon.x) <= kMaxStepAngle;  // high latitude: dlon = b / cos(lat)
        step_ok[1] = step_ok[1] && fabsf(m.sin_dlon.y) <= kMaxStepAngle;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (act[h] && !step_ok[h]) {  // fast::kStNeedsLiteral: the rollout continues in the fp64 kernel
                o.why[h] |= kWhyStep;
                o.nstep[h] = j;
                act[h] = false;
            }
        }
        // (a half whose longitude increment was refused after the move keeps its old position: its increments are
        // only applied when it is still active)
        if (act[0]) {
            lat[0] = fma((double)m.dlat.x, 180.0 / kPi, lat[0]);
            lon[0] = fma((double)m.dlon.x, 180.0 / kPi, lon[0]);
            o.k[0] += 1;
        }
        if (act[1]) {
            lat[1] = fma((double)m.dlat.y, 180.0 / kPi, lat[1]);
            lon[1] = fma((double)m.dlon.y, 180.0 / kPi, lon[1]);
            o.k[1] += 1;
        }
        ++j;
        const f2 sB = m.s_new, cB = m.c_new;
        // ---- fast::grid_pos<false> per half ---------------------------------------------------------------------------
        {
            const double fa0 = fma(lat[0], sc.inv_dlat, sc.fa_c0), fo0 = fma(lon[0], sc.inv_dlon, sc.fo_c0);
            const double fa1 = fma(lat[1], sc.inv_dlat, sc.fa_c0), fo1 = fma(lon[1], sc.inv_dlon, sc.fo_c0);
            g[0].ia = imin((int)fa0, sc.nlat - 2); g[0].io = imin((int)fo0, sc.nlon - 2);
            g[1].ia = imin((int)fa1, sc.nlat - 2); g[1].io = imin((int)fo1, sc.nlon - 2);
            wa = mk((float)(fa0 - (double)g[0].ia), (float)(fa1 - (double)g[1].ia));
            wo = mk((float)(fo0 - (double)g[0].io), (float)(fo1 - (double)g[1].io));
            fa = mk((float)fa0, (float)fa1);
            fo = mk((float)fo0, (float)fo1);
        }
        // ---- fast::fast_at_dest<true> ---------------------------------------------------------------------------------
        // reach test first (see fast::fast_at_dest): a half whose destination is more than a step and a half away is
        // not at it
        const f2 ba2_half = fma2(mul2(sA, sB), m.h_dlon, neg2(m.cm1_dlat));
        const f2 da2_low = fma2(mul2(mul2(sA, sD), rd.h_dlon), 2.0f, mul2(rd.sin_dlat, rd.sin_dlat));
        const bool reach0 = act[0] && !(ba2_half.x > 0.0f && da2_low.x > 4.0f * ba2_half.x);
        const bool reach1 = act[1] && !(ba2_half.y > 0.0f && da2_low.y > 4.0f * ba2_half.y);
        if (reach0 || reach1) {
            const bool reach[2] = {reach0, reach1};
            const f2 dBy = mul2(sB, m.sin_dlon);
            const f2 q = fma2(neg2(mul2(cA, sB)), m.h_dlon, m.sin_dlat);
            const f2 r = fma2(mul2(cA, sD), rd.h_dlon, neg2(rd.sin_dlat));
            const f2 dDy = mul2(rd.sin_dlon, sD);
            const f2 T = fma2(dBy, r, mul2(dDy, q));
            const f2 N2 = fma2(dBy, dBy, mul2(q, q));
            const f2 lhs = mul2(T, T), rhs = mul2(N2, d::sin2_dest_angle);
            const f2 gap = sub2(lhs, rhs), tol = mul2(rhs, d::margin_angle);
            bool near[2];
            const float N2h[2] = {N2.x, N2.y}, gaph[2] = {gap.x, gap.y}, tolh[2] = {tol.x, tol.y};
            const float lhsh[2] = {lhs.x, lhs.y}, rhsh[2] = {rhs.x, rhs.y};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                near[h] = false;
                if (reach[h]) {
                    if (!(N2h[h] > 0.0f)) {  // A == B
                        o.why[h] |= kWhyDegenerate;
                    } else {
                        if (fabsf(gaph[h]) <= tolh[h]) o.why[h] |= kWhyAngle;
                        near[h] = !(lhsh[h] > rhsh[h]);
                    }
                }
            }
            if (near[0] || near[1]) {  // (once per rollout: the step that crosses the destination)
                const f2 dBx = fma2(sA, m.cm1_dlat, fma2(cA, m.sin_dlat, neg2(mul2(sB, m.h_dlon))));
                const f2 dBz = fma2(cA, m.cm1_dlat, neg2(mul2(sA, m.sin_dlat)));
                const f2 xx = mul2(rd.x_dlat, rd.x_dlat);
                const f2 cm1D = mul2(neg2(xx), cosc_poly(xx));
                const f2 dDx = fma2(sA, cm1D, fma2(cA, rd.sin_dlat, neg2(mul2(rd.h_dlon, sD))));
                const f2 dDz = fma2(cA, cm1D, neg2(mul2(sA, rd.sin_dlat)));
                const f2 ab = fma2(dDx, dBx, fma2(dDy, dBy, mul2(dDz, dBz)));
                const f2 dd = fma2(dDx, dDx, fma2(dDy, dDy, mul2(dDz, dDz)));
                const f2 bb = fma2(dBx, dBx, fma2(dBy, dBy, mul2(dBz, dBz)));
                const f2 p = sub2(ab, dd);  // (D - A).(B - D)
                const f2 along = mul2(bb, d::margin_along), miss = mul2(bb, d::near_miss2);
                const f2 gapB = add2(fma2(ab, -2.0f, dd), bb);
                const f2 fr = mul2(ab, rcp2(bb));
                const float ph[2] = {p.x, p.y}, alongh[2] = {along.x, along.y}, missh[2] = {miss.x, miss.y};
                const float gapBh[2] = {gapB.x, gapB.y}, frh[2] = {fr.x, fr.y};
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (near[h]) {
                        if (fabsf(ph[h]) <= alongh[h]) o.why[h] |= kWhyAlong;
                        if (ph[h] < 0.0f) {
                            if (gapBh[h] < missh[h]) o.why[h] |= kWhyNearMiss;
                        } else {
                            o.frac[h] = frh[h];
                            o.at[h] = 1;
                        }
                    }
                }
            }
        }
        // ---- fast::fast_terminal<true> (only for a half still undecided) ---------------------------------------------
        {
            const f2 da0 = add2(fa, bc(-sc.box_fa_lo)), da1 = sub2(bc(sc.box_fa_hi), fa);
            const f2 do0 = add2(fo, bc(-sc.box_fo_lo)), do1 = sub2(bc(sc.box_fo_hi), fo);
            const float lo[2] = {fminf(fminf(da0.x, da1.x), fminf(do0.x, do1.x)), fminf(fminf(da0.y, da1.y), fminf(do0.y, do1.y))};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (act[h]) {
                    bool ended = o.at[h] || o.why[h];
                    if (!ended) {
                        if (o.k[h] >= sc.nT - 1) {
                            ended = true;
                        } else {
                            if (fabsf(lo[h]) <= box_m) o.why[h] |= kWhyBox;  // only the binding edge can flip
                            ended = lo[h] < 0.0f || o.why[h];
                        }
                    }
                    if (ended) {
                        o.nstep[h] = j;
                        act[h] = false;
                    }
                }
            }
        }
        sA = sB;
        cA = cB;
        rd = dest_rel(lat[0], lon[0], lat[1], lon[1], a.dest_lat, a.dest_lon);
    }
}

}  // namespace fast2
}  // namespace pmcts
#endif
