// pmcts_b200 mixed-precision ("fast") simulator: the same transition as pmcts_device.cuh, re-derived
// so that an fp32 datapath is enough.
//
// The reference (code/model/simulatorTLKT.py:132-216, code/solver/worker.py:255-300, 380-415) works
// on absolute latitudes / longitudes through full-range fp64 transcendentals.  Evaluated in fp32 those
// formulas lose the answer to cancellation (a step moves the boat by 1e-4 .. 1e-2 rad on top of a
// position of ~1 rad).  Here every quantity that is small is computed *as a small quantity*:
//
//   * the position is kept in fp64 degrees, but a step only produces the increments
//         sin(dlat) = a - s g,   sin(dlon) = b / H,     with a = sin(d) cos(hdg), b = sin(d) sin(hdg),
//         A = c cos(d) - s a,  H = sqrt(A^2 + b^2) = cos(lat'),  g = b^2 / (H + A),  (s, c) = sin/cos(lat)
//     which is Simulator.getDestination (simulatorTLKT.py:132-166) rewritten exactly (no series in d);
//     (s, c) only scale small numbers, so fp32 is enough for them;
//   * the bearing to the destination (simulatorTLKT.py:124-128) and the arrival test of
//     Tree.is_state_at_dest (worker.py:380-415, including its latitude-as-colatitude mapping,
//     simulatorTLKT.py:233-238) are written in differences to the current position:
//         y = sin(latD - lat) + s cD (1 - cos(lonD - lon)),  x = cD sin(lonD - lon)
//         D.(A x B) = dBy r + dDy q,  |A x B|^2 = dBy^2 + q^2        (see fast_at_dest)
//   * the point of sail is the angle between the heading and the wind vector (a dot / cross
//     product), so the wind direction never goes through atan2 / mod 360 / abs;
//   * the polar polynomial is evaluated in variables centred on its domain (BoatParams::pc).
//
// What fp32 cannot guarantee is the *decision* of a test whose two sides are closer than the
// rounding noise (arrival angle against DESTINATION_ANGLE, overshoot test, terminal box, histogram
// bin edges).  Every such test carries a margin; a rollout that lands inside one is not decided here:
// it is queued (RolloutArgs::redo_list) and re-run by the literal fp64 kernel.
//
// Everything is PMCTS_HD so that tests/fastmodel can compile the same code for the host and measure
// its error against the oracle without a GPU (the product never runs the host build).
#pragma once
#include <math.h>

#include "pmcts_types.h"

namespace pmcts {
namespace fast {

// ---- hardware approximations (MUFU) with host stand-ins ------------------------------------------
PMCTS_HD float rcp_approx(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}
PMCTS_HD float rsqrt_approx(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / sqrtf(x);
#endif
}
PMCTS_HD float exp_approx(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
    return r;
#else
    return expf(x);
#endif
}
PMCTS_HD float fma_(float a, float b, float c) { return fmaf(a, b, c); }

// ---- packed FP32 (sm_100: FFMA2 / FMUL2 / FADD2 work on a 64-bit register pair and issue as ONE instruction; an
// operand may also be one register or an immediate broadcast to both halves).  The mixed-precision kernels are bound
// by instruction issue, so wherever one rollout evaluates the same expression on two quantities -- the (u, v)
// components of the wind, the latitude / longitude legs of a polynomial -- the two go through as a pair.  Each half
// is the same IEEE operation as its scalar form (round to nearest, fused where the scalar form is fused), so results
// do not change; the host build does them one by one.
PMCTS_HD float2 pk(float x, float y) {
    float2 r;
    r.x = x;
    r.y = y;
    return r;
}
PMCTS_HD float2 pk_fma(float2 a, float2 b, float2 c) {
#if defined(__CUDA_ARCH__)
    return __ffma2_rn(a, b, c);
#else
    return pk(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
PMCTS_HD float2 pk_mul(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
    return __fmul2_rn(a, b);
#else
    return pk(a.x * b.x, a.y * b.y);
#endif
}
PMCTS_HD float2 pk_sub(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
    return __fadd2_rn(a, pk(-b.x, -b.y));
#else
    return pk(a.x - b.x, a.y - b.y);
#endif
}
PMCTS_HD float2 pk_fma(float2 a, float2 b, float c) { return pk_fma(a, b, pk(c, c)); }
PMCTS_HD float2 pk_fma(float2 a, float b, float2 c) { return pk_fma(a, pk(b, b), c); }
PMCTS_HD float2 pk_mul(float2 a, float b) { return pk_mul(a, pk(b, b)); }
PMCTS_HD int imin(int a, int b) { return a < b ? a : b; }
PMCTS_HD int imax(int a, int b) { return a > b ? a : b; }

// ---- polynomial kernels (minimax fits, coefficients rounded to fp32; max abs error after rounding:
// sinc 8e-9, cosc 3e-9 on |x| <= pi/2, atan 4e-8 on [0, 1]) -----------------------------------------
// sin(x) = x * sinc_poly(x^2), 1 - cos(x) = x^2 * cosc_poly(x^2), |x| <= pi/2
PMCTS_HD float sinc_poly(float y) {
    float p = 2.605224902e-06f;
    p = fma_(p, y, -1.980907528e-04f);
    p = fma_(p, y, 8.333051063e-03f);
    p = fma_(p, y, -1.666665799e-01f);
    return fma_(p, y, 1.0f);
}
PMCTS_HD float cosc_poly(float y) {
    float p = 2.633458343e-07f;
    p = fma_(p, y, -2.477660893e-05f);
    p = fma_(p, y, 1.388868840e-03f);
    p = fma_(p, y, -4.166666179e-02f);
    return fma_(p, y, 0.5f);
}
// (sinc_poly of two arguments at once)
PMCTS_HD float2 sinc_poly2(float2 y) {
    float2 p = pk_fma(pk(2.605224902e-06f, 2.605224902e-06f), y, -1.980907528e-04f);
    p = pk_fma(p, y, 8.333051063e-03f);
    p = pk_fma(p, y, -1.666665799e-01f);
    return pk_fma(p, y, 1.0f);
}
// short forms for a step length |x| <= 0.1 rad (next terms: 2e-10 and 3e-11 relative)
PMCTS_HD float sinc_small(float y) { return fma_(fma_(8.3333333e-03f, y, -1.6666667e-01f), y, 1.0f); }
PMCTS_HD float cosc_small(float y) { return fma_(fma_(1.3888889e-03f, y, -4.1666667e-02f), y, 0.5f); }
// asin(x) = x * asinc_small(x^2), |x| <= 0.12 (next term 3e-10 relative)
PMCTS_HD float asinc_small(float y) {
    return fma_(fma_(fma_(4.4642857e-02f, y, 7.5e-02f), y, 1.6666667e-01f), y, 1.0f);
}
PMCTS_HD float2 asinc_small2(float2 y) {
    return pk_fma(pk_fma(pk_fma(pk(4.4642857e-02f, 4.4642857e-02f), y, 7.5e-02f), y, 1.6666667e-01f), y, 1.0f);
}
// atan2(s, c) for s >= 0, result in [0, pi]
PMCTS_HD float atan2_pos(float s, float c) {
    const float ac = fabsf(c);
    const float mx = fmaxf(s, ac), mn = fminf(s, ac);
    const float t = mn * rcp_approx(mx);
    const float y = t * t;
    float p = 2.903537903e-03f;
    p = fma_(p, y, -1.628295011e-02f);
    p = fma_(p, y, 4.303927486e-02f);
    p = fma_(p, y, -7.533668294e-02f);
    p = fma_(p, y, 1.065467381e-01f);
    p = fma_(p, y, -1.420713276e-01f);
    p = fma_(p, y, 1.999305400e-01f);
    p = fma_(p, y, -3.333309395e-01f);
    p = fma_(p, y, 1.0f);
    float r = t * p;
    if (s > ac) r = 1.57079632679f - r;
    if (c < 0.0f) r = 3.14159265359f - r;
    return r;
}

// ---- noise: Philox4x32-10 + Box-Muller in fp32 (same counter layout as oracle/pmcts_oracle.c) ------
PMCTS_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                            uint32_t out[4], uint32_t m0 = 0xD2511F53u, uint32_t m1 = 0xCD9E8D57u) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#if defined(__CUDA_ARCH__)
        // (one IMAD.WIDE per product: written as a 64-bit C multiply the compiler adds the -- zero -- cross term of the
        // widened multiplier back in, two more instructions per round)
        uint64_t p0, p1;
        asm("mul.wide.u32 %0, %1, %2;" : "=l"(p0) : "r"(m0), "r"(c0));
        asm("mul.wide.u32 %0, %1, %2;" : "=l"(p1) : "r"(m1), "r"(c2));
#else
        const uint64_t p0 = (uint64_t)m0 * c0, p1 = (uint64_t)m1 * c2;
#endif
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

PMCTS_HD void box_muller(uint32_t x0, uint32_t x1, float &z0, float &z1) {
    // u = (x + 0.5) * 2^-32 in (0, 1); the angle is taken about pi so the hardware sin / cos stay in
    // their accurate range: cos(2 pi u) = -cos(2 pi u - pi)
    const float u0 = fminf(fma_((float)x0, 2.3283064365386963e-10f, 1.1641532182693481e-10f), 0.99999994f);
    const float th = fma_((float)x1, 1.4629180792671596e-09f, -3.14159265359f + 7.3145904e-10f);
#if defined(__CUDA_ARCH__)
    float l2;  // (u0 >= 2^-33 is a normal number: no denormal scaling around the hardware logarithm)
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u0));
    const float t = -1.3862943611198906f * l2;  // -2 ln(u0) > 0
    const float r = t * rsqrt_approx(t);
    float s, c;
    __sincosf(th, &s, &c);
#else
    const float r = sqrtf(-2.0f * logf(u0));
    const float s = sinf(th), c = cosf(th);
#endif
    z0 = -r * c;
    z1 = -r * s;
}

struct Noise {
    // The four normals of a Philox block sit in a shift register: consecutive steps starting at a multiple of
    // four (step 0 in every kernel; the only way they draw) take c0 and move the rest down, so the common path
    // has no indexed selection.
    float c0, c1, c2, c3;
    int cached_block;
    PMCTS_HD Noise() : c0(0.0f), c1(0.0f), c2(0.0f), c3(0.0f), cached_block(-1) {}
    // kMode: PMCTS_NOISE_* known at compile time, or -1 to read nz.mode
    // kParamMul: take Philox's multipliers from nz (see NoiseView) -- measured faster in K1 (1.21e11 ->
    // 1.24e11 transitions/s) and slower in K2 (1.119 -> 1.137 ms), so each kernel picks its own
    template <int kMode, bool kParamMul = false>
    PMCTS_HD float draw(const NoiseView &nz, int64_t n, int64_t i, int step) {
        const int mode = kMode >= 0 ? kMode : nz.mode;
        if (mode == PMCTS_NOISE_INJECTED) return (float)nz.z[(size_t)step * n + i];
        if (mode == PMCTS_NOISE_NONE) return 0.0f;
        const int block = step >> 2;
        if (block != cached_block) {
            const uint64_t boat = nz.id0 + (uint64_t)i;
            uint32_t x[4];
            philox4x32_10((uint32_t)boat, (uint32_t)(boat >> 32), (uint32_t)block, (uint32_t)nz.stream,
                          (uint32_t)nz.seed, (uint32_t)(nz.seed >> 32), x, kParamMul ? nz.m0 : 0xD2511F53u,
                          kParamMul ? nz.m1 : 0xCD9E8D57u);
            box_muller(x[0], x[1], c0, c1);
            box_muller(x[2], x[3], c2, c3);
            cached_block = block;
        }
        const float z = c0;
        c0 = c1; c1 = c2; c2 = c3;
        return z;
    }
};

// ---- position on the wind grid -------------------------------------------------------------------
// Cell index and weights of SciPy's linear RegularGridInterpolator (weatherTLKT.py:377-378; cell
// clipped to n-2 so that the last node is reached with weight 1); the index is taken in fp64, the
// in-cell weight is an fp32 number in [0, 1].
struct GridPos {
    int ia, io;
    float wa, wo;
    float fa, fo;  // fractional index (terminal-box test)
    bool inside;   // SciPy bounds_error=True
};

// kInside = false: the caller holds the position against the terminal box of a roll_safe scenario (box
// inside the grid), which implies the bounds test.
// kRad (with kInside = false only): the position is in radians and known to lie inside the terminal box of a
// roll_safe scenario, which ends before the last grid node: no clamp either.
// kRad also takes the offsets of the index (fa_c0 / fo_c0 of the scenario for an absolute position in radians; the
// turbo loop holds its position relative to the destination and passes the destination's index).
template <bool kInside = true, bool kRad = false>
PMCTS_HD GridPos grid_pos(const ScenarioView &sc, double lat, double lon, double c0a = 0.0, double c0o = 0.0) {
    GridPos g;
    g.inside = !kInside || (lat >= sc.lat0 && lat <= sc.lat_last && lon >= sc.lon0 && lon <= sc.lon_last);
    // (lat - lat0) / dlat as one fma; inside the grid it is >= -1e-13, so truncation needs no lower clamp
    const double fa = fma(lat, kRad ? sc.inv_dlat_rad : sc.inv_dlat, kRad ? c0a : sc.fa_c0);
    const double fo = fma(lon, kRad ? sc.inv_dlon_rad : sc.inv_dlon, kRad ? c0o : sc.fo_c0);
    g.fa = (float)fa;
    g.fo = (float)fo;
    if (kRad) {
        // turbo (grids of at most 128 cells a side, see turbo_ok): cell and in-cell weight from the fp32 index.  The
        // weight is then good to half an ulp of the index (<= 4e-6 of a cell); the interpolant is continuous across
        // cells, so a cell taken one off at a boundary changes nothing.  Spares the fp64 round trip of the exact
        // split: 2 F2I.F64 + 2 I2F.F64 + 2 DADD + 2 F2F per step -> 2 F2I + 2 I2F + 1 FADD2 (K2 0.923 -> 0.907 ms);
        // measured against the oracle (tests/test_fastmodel.py): final positions 5e-8 relative at worst, as before.
        g.ia = (int)g.fa;
        g.io = (int)g.fo;
        const float2 w = pk_sub(pk(g.fa, g.fo), pk((float)g.ia, (float)g.io));
        g.wa = w.x;
        g.wo = w.y;
        return g;
    }
    g.ia = kRad ? (int)fa : imin((int)fa, sc.nlat - 2);
    g.io = kRad ? (int)fo : imin((int)fo, sc.nlon - 2);
    g.wa = (float)(fa - (double)g.ia);
    g.wo = (float)(fo - (double)g.io);
    return g;
}

// Simulator.getWind (simulatorTLKT.py:168-179) on the time-blended fp32 slab of instant k.  kShared: the
// slab has been staged in shared memory (`slab` points there); otherwise it is read through L1 / L2.
template <bool kShared = false>
PMCTS_HD void wind_at(const ScenarioView &sc, int k, const GridPos &g, float &u, float &v,
                      const float2 *slab = nullptr) {
    // one 32-bit cell index (< 2^31 cells, checked at upload), then one widening multiply-add per row
    const int cell = (kShared ? 0 : k * sc.slab32_stride) + (g.ia * sc.nlon + g.io);
    const float2 *base = kShared ? slab : sc.slab32;
    const float2 *row0 = base + cell;
    const float2 *row1 = base + (cell + sc.nlon);
#if defined(__CUDA_ARCH__)
    float2 c00, c01, c10, c11;
    if (kShared) {
        c00 = row0[0]; c01 = row0[1]; c10 = row1[0]; c11 = row1[1];
    } else {
        c00 = __ldg(row0); c01 = __ldg(row0 + 1); c10 = __ldg(row1); c11 = __ldg(row1 + 1);
    }
#else
    const float2 c00 = row0[0], c01 = row0[1], c10 = row1[0], c11 = row1[1];
#endif
    // the three blends of (u, v) as pairs: 6 packed operations instead of 12
    const float2 r0 = pk_fma(pk_sub(c01, c00), g.wo, c00), r1 = pk_fma(pk_sub(c11, c10), g.wo, c10);
    const float2 uv = pk_fma(pk_sub(r1, r0), g.wa, r0);
    u = uv.x;
    v = uv.y;
}

// ---- Boat.getDeterDyn (simulatorTLKT.py:473-502) ----------------------------------------------------
// The reference's own boat (Boat.FIT_VELOCITY and the module constants, simulatorTLKT.py:454-471) as
// compile-time constants: derive_fast_params' output for pmcts_create's defaults.  sm_100 takes constants
// from uniform registers, not from the constant bank, and the rollout loop needs more of them than there are;
// with kDefBoat these 29 -- and the six tunables below -- become instruction immediates (the noise coefficient stays
// a run-time value: Tutorial.py sets it).  Kernels are instantiated both ways and the host
// picks the specialised one when is_default_boat(params) (tests/test_abi.py pins the table to
// derive_fast_params).
namespace defboat {
// (scalars: device code may use a constexpr scalar, not an element of a constexpr array)
constexpr float pc0 = 1.373763084e+00f, pc1 = 1.079237461e-01f, pc2 = -6.408512592e-01f, pc3 = 1.135649323e+00f, pc4 = -3.408502787e-02f, pc5 = -5.383408070e-01f;
constexpr float pc6 = 7.206667066e-01f, pc7 = 8.731120229e-01f, pc8 = 2.533754408e-01f, pc9 = -2.119406760e-01f, pc10 = -3.206158578e-01f, pc11 = 0.0f;
constexpr float pc12 = 5.637606978e-02f, pc13 = 2.704189122e-01f, pc14 = 1.337060332e-01f, pc15 = -7.381652296e-02f, pc16 = 0.0f, pc17 = 0.0f;
constexpr float pc18 = -3.346018121e-02f, pc19 = 3.288391531e-01f, pc20 = 2.619542181e-01f, pc21 = 0.0f, pc22 = 0.0f, pc23 = 0.0f;
constexpr float pc24 = -2.172237635e-01f, pc25 = -2.437742054e-01f, pc26 = 0.0f, pc27 = 0.0f, pc28 = 0.0f, pc29 = 0.0f;
constexpr float pc30 = 6.050717458e-02f, pc31 = 0.0f, pc32 = 0.0f, pc33 = 0.0f, pc34 = 0.0f, pc35 = 0.0f;
constexpr float pc[36] = {pc0, pc1, pc2, pc3, pc4, pc5, pc6, pc7, pc8, pc9, pc10, pc11, pc12, pc13, pc14, pc15, pc16, pc17, pc18, pc19, pc20, pc21, pc22, pc23, pc24, pc25, pc26, pc27, pc28, pc29, pc30, pc31, pc32, pc33, pc34, pc35};  // host side (is_default_boat)
constexpr float p_mid = 9.75e+01f, p_ihalf = 1.600000076e-02f, w_ihalf = 1.047120392e-01f;
constexpr float wind_max = 1.910000038e+01f, pmin = 35.0f, pmax = 160.0f;
constexpr float cos_pmin = 8.191520572e-01f, cos_pmax = -9.396926165e-01f;
// DESTINATION_ANGLE (worker.py:20) and EARTH_RADIUS (simulatorTLKT.py:26) of the reference, and the default
// decision margins (pmcts_host.h: default_fast_margins takes them from here)
constexpr float sin2_dest_angle = 7.615435393e-09f, inv_radius = 1.569612351e-07f;
// (near_miss2: a boat that stops within 1/1000 of a step short of the destination.  Round 1 used 1/316; the measured
//  position error -- 2e-7 relative, 1.6e-5 of a step -- turns into a bearing error of a degree at 1/1000, which moves the
//  reward of the step that follows by 5e-6, less than half of the margin the histogram's bin edges carry.  Over 2^20 rollouts of 8 scenarios the
//  fp32 path decides like the oracle with the margin at 1e-5, 1e-6, 1e-7 and closed; 1e-6 leaves 0.12 % of the rollouts
//  to the fp64 finisher instead of 0.32 %.)
constexpr float margin_angle = 2e-4f, margin_along = 2e-5f, near_miss2 = 1e-6f, margin_box = 3e-5f;
constexpr double margin_bin = 1e-4;  // the fp32 reward is good to ~1e-5 in these units
}  // namespace defboat

inline bool is_default_boat(const BoatParams &bp) {
    for (int i = 0; i < 36; ++i)
        if (bp.pc[i] != defboat::pc[i]) return false;
    return bp.polar_tri == 1 && bp.p_mid == defboat::p_mid && bp.p_ihalf == defboat::p_ihalf &&
           bp.w_ihalf == defboat::w_ihalf && bp.wind_max_f == defboat::wind_max && bp.pmin_f == defboat::pmin &&
           bp.pmax_f == defboat::pmax && bp.cos_pmin_f == defboat::cos_pmin && bp.cos_pmax_f == defboat::cos_pmax &&
           bp.sin2_dest_angle == defboat::sin2_dest_angle && bp.inv_radius_f == defboat::inv_radius &&
           bp.margin_angle == defboat::margin_angle && bp.margin_along == defboat::margin_along &&
           bp.near_miss2 == defboat::near_miss2 && bp.margin_box == defboat::margin_box;
}

template <bool kDefBoat = false>
PMCTS_HD float polar_centred(const BoatParams &bp, float x, float y) {
#define PMCTS_PC(i) (kDefBoat ? defboat::pc##i : bp.pc[i])
    if (kDefBoat || bp.polar_tri) {
        // column j (power of y) has degree 5 - j in x
        float c0 = PMCTS_PC(30);
        c0 = fma_(c0, x, PMCTS_PC(24)); c0 = fma_(c0, x, PMCTS_PC(18)); c0 = fma_(c0, x, PMCTS_PC(12));
        c0 = fma_(c0, x, PMCTS_PC(6)); c0 = fma_(c0, x, PMCTS_PC(0));
        float c1 = PMCTS_PC(25);
        c1 = fma_(c1, x, PMCTS_PC(19)); c1 = fma_(c1, x, PMCTS_PC(13)); c1 = fma_(c1, x, PMCTS_PC(7));
        c1 = fma_(c1, x, PMCTS_PC(1));
        float c2 = PMCTS_PC(20);
        c2 = fma_(c2, x, PMCTS_PC(14)); c2 = fma_(c2, x, PMCTS_PC(8)); c2 = fma_(c2, x, PMCTS_PC(2));
        float c3 = PMCTS_PC(15);
        c3 = fma_(c3, x, PMCTS_PC(9)); c3 = fma_(c3, x, PMCTS_PC(3));
        float c4 = fma_(PMCTS_PC(10), x, PMCTS_PC(4));
        float acc = PMCTS_PC(5);
        acc = fma_(acc, y, c4); acc = fma_(acc, y, c3); acc = fma_(acc, y, c2); acc = fma_(acc, y, c1);
        return fma_(acc, y, c0);
    }
#undef PMCTS_PC
    const float *c = bp.pc;
    float acc = 0.0f;
#pragma unroll
    for (int j = 5; j >= 0; --j) {
        float col = c[30 + j];
#pragma unroll
        for (int i = 4; i >= 0; --i) col = fma_(col, x, c[i * 6 + j]);
        acc = fma_(acc, y, col);
    }
    return acc;
}

// Boat speed for wind (u, v) and the unit heading vector (sh, ch) = (sin, cos)(heading):
// Weather.returnPolarVel (weatherTLKT.py:148-163) + point of sail (simulatorTLKT.py:203) +
// Boat.getDeterDyn.  The wind blows *to* (u, v); the point of sail is the angle between the heading
// and the direction the wind comes *from*, folded to [0, 180] -- i.e. the plain angle between the
// two unit vectors (the reference's abs / 360 - p fold is that angle for headings in [0, 360)).
template <bool kDefBoat = false>
PMCTS_HD float boat_speed(const BoatParams &bp, float u, float v, float sh, float ch) {
    const float wind_max = kDefBoat ? defboat::wind_max : bp.wind_max_f;
    const float pmin = kDefBoat ? defboat::pmin : bp.pmin_f, pmax = kDefBoat ? defboat::pmax : bp.pmax_f;
    const float p_mid = kDefBoat ? defboat::p_mid : bp.p_mid, p_ihalf = kDefBoat ? defboat::p_ihalf : bp.p_ihalf;
    const float w_ihalf = kDefBoat ? defboat::w_ihalf : bp.w_ihalf;
    const float cos_pmin = kDefBoat ? defboat::cos_pmin : bp.cos_pmin_f;
    const float cos_pmax = kDefBoat ? defboat::cos_pmax : bp.cos_pmax_f;
    const float m2 = fma_(u, u, v * v);
    // A wind of exactly zero has no direction (the reference goes by atan2(0, 0) = 0: a calm wind "blows north").
    // It is not given one here: 0 * rsqrt(0) is not a number, neither is anything computed from it -- the tack /
    // gybe factor below carries it into the speed whatever minimum and maximum make of the angle -- and the step's
    // length test (fast_step) hands what is not a number to the literal kernel, which has the reference's rule.  Four
    // instructions of the step loop (the test, three defaults) for a case interpolated winds do not produce.
    const float rm = rsqrt_approx(m2);
    const float mag = m2 * rm, du = u * rm, dv = v * rm;
    const float cosp = -fma_(du, sh, dv * ch);
    const float sinp = fabsf(fma_(du, ch, -dv * sh));
    const float p = atan2_pos(sinp, cosp) * 57.2957795131f;
    const float w = fminf(mag, wind_max);
    const float pe = fminf(fmaxf(p, pmin), pmax);
    float sp = polar_centred<kDefBoat>(bp, (pe - p_mid) * p_ihalf, fma_(w, w_ihalf, -1.0f));
    if (p != pe) sp *= (p < pmin ? cos_pmin : cos_pmax) * rcp_approx(cosp);  // tacking / gybing (rare): one branch
    return sp;
}

// ---- Simulator.getDestination (simulatorTLKT.py:132-166) in increments ---------------------------
struct Move {
    float sin_dlat;  // sin(lat' - lat)
    float sin_dlon;  // sin(lon' - lon)
    float h_dlon;    // 1 - cos(lon' - lon)
    float cm1_dlat;  // cos(lat' - lat) - 1
    float dlat, dlon;  // radians
    float s_new, c_new;  // sin / cos of the new latitude (exact identities: s cd + c a, H)
};

// d = angular distance (rad, may be negative), (sh, ch) heading, (s, c) = sin/cos(lat).
PMCTS_HD Move move_increments(float d, float sh, float ch, float s, float c) {
    Move m;
    const float d2 = d * d;
    const float sd = d * sinc_small(d2);
    const float e = d2 * cosc_small(d2);  // 1 - cos(d)
    const float2 ba = pk_mul(pk(sh, ch), sd);
    const float b = ba.x, a = ba.y;
    const float cd = 1.0f - e;
    const float A = fma_(c, cd, -s * a);
    const float b2 = b * b;
    const float n2 = fma_(A, A, b2);
    const float rH = rsqrt_approx(n2);
    const float H = n2 * rH;
    const float g = b2 * rcp_approx(H + A);
    m.sin_dlon = b * rH;
    m.s_new = fma_(s, cd, c * a);
    m.c_new = H;
    m.h_dlon = g * rH;
    m.sin_dlat = fma_(-s, g, a);
    m.cm1_dlat = fma_(c, g, -e);
    const float2 sn = pk(m.sin_dlat, m.sin_dlon);
    const float2 an = pk_mul(sn, asinc_small2(pk_mul(sn, sn)));
    m.dlat = an.x;
    m.dlon = an.y;
    return m;
}

// sin / cos of a latitude given in fp64 degrees
PMCTS_HD void sincos_lat(double lat_deg, float &s, float &c) {
    const float x = (float)(lat_deg * kDegToRad);
    const float y = x * x;
    s = x * sinc_poly(y);
    c = fma_(-y, cosc_poly(y), 1.0f);
}

// Destination seen from the current position: the three numbers the bearing and the arrival test need.
struct DestRel {
    float sin_dlat;  // sin(latD - lat)
    float sin_dlon;  // sin(lonD - lon)
    float h_dlon;    // 1 - cos(lonD - lon)
    float x_dlat;    // latD - lat, radians (for 1 - cos(latD - lat) in the arrival branch)
};

// kRad: the position is in radians AND relative to the destination (dlat, dlon unused: the difference is minus the
// position -- written as a negation, which `0.0 - lat` is not for the compiler: the two differ in the sign of zero).
template <bool kRad = false>
PMCTS_HD DestRel dest_rel(double lat, double lon, double dlat, double dlon) {
    DestRel r;
    const float2 x = kRad ? pk(-(float)lat, -(float)lon)
                          : pk((float)((dlat - lat) * kDegToRad), (float)((dlon - lon) * kDegToRad));
    const float2 y = pk_mul(x, x);
    const float2 sn = pk_mul(x, sinc_poly2(y));
    r.sin_dlat = sn.x;
    r.sin_dlon = sn.y;
    r.h_dlon = y.y * cosc_poly(y.y);
    r.x_dlat = x.x;
    return r;
}

// Unit vector of the initial bearing to the destination (simulatorTLKT.py:124-128):
// (x, y) = (cD sin(dlon), c sD - s cD cos(dlon)) with y rewritten as sin(latD - lat) + s cD (1 - cos(dlon)).
PMCTS_HD void bearing_unit(const DestRel &r, float s, float sD, float cD, float &sh, float &ch) {
    (void)sD;
    const float x = cD * r.sin_dlon;
    const float y = fma_(s * cD, r.h_dlon, r.sin_dlat);
    const float n2 = fma_(x, x, y * y);
    // (a boat exactly on the destination -- atan2(0, 0) = 0 in the reference -- gets no heading here: 0 * rsqrt(0) is
    //  not a number and fast_step's length tests leave the boat to the literal kernel, see boat_speed)
    const float2 h = pk_mul(pk(x, y), rsqrt_approx(n2));
    sh = h.x;
    ch = h.y;
}

// ---- Tree.is_state_at_dest (worker.py:380-415) -----------------------------------------------------
// Points are mapped by Simulator.fromGeoToCartesian (simulatorTLKT.py:233-238), which uses the
// latitude as a colatitude: P = (s cos(lon), s sin(lon), c) with (s, c) = sin/cos(lat).  Rotating
// about the z axis so that A sits at longitude 0 leaves the test unchanged and gives, with B and D
// written relative to A,
//     dB = (sA cm1B + cA sdB - sB hB,  sB sin(dlonB),  cA cm1B - sA sdB)     (B - A)
//     dD = (sA cm1D + cA sdD - sD hD,  sD sin(dlonD),  cA cm1D - sA sdD)     (D - A)
//     D . (A x B) = dBy r + dDy q,   |A x B|^2 = dBy^2 + q^2,
//     q = sdB - cA sB hB,   r = -sdD + cA sD hD
// (sdX = sin(latX - latA), cm1X = cos(latX - latA) - 1, hX = 1 - cos(lonX - lonA)).
// The distance of D to the plane (O, A, B) is |D . (A x B)| / |A x B|; the reference takes its asin
// and compares with DESTINATION_ANGLE, then asks for (D - A).(B - D) >= 0 and reports
// frac = (D - A).(B - A) / |B - A|^2.
// Returns 1 arrived, 0 not arrived; *uncertain is set when a comparison is closer than its margin
// (or the plane is degenerate: the reference raises ZeroDivisionError there).
// Why a rollout was left to the literal kernel (RollOut::why bits).
enum : int { kWhyAngle = 1, kWhyAlong = 2, kWhyBox = 4, kWhyBin = 8, kWhyDegenerate = 16, kWhyStep = 32, kWhyNearMiss = 64 };

// The margins live in BoatParams (margin_*; defaults in pmcts_host.h, pmcts_set_fast_margins overrides).

template <bool kDefBoat = false>
PMCTS_HD int fast_at_dest(const BoatParams &bp, const DestRel &rd, const Move &m, float sA, float cA, float sB,
                          float sD, float &frac, int &why) {
    const float sin2_dest_angle = kDefBoat ? defboat::sin2_dest_angle : bp.sin2_dest_angle;
    const float margin_angle = kDefBoat ? defboat::margin_angle : bp.margin_angle;
    const float margin_along = kDefBoat ? defboat::margin_along : bp.margin_along;
    const float near_miss2 = kDefBoat ? defboat::near_miss2 : bp.near_miss2;
    // Reach test.  (D - A).(B - D) >= 0 puts D in the ball whose diameter is AB, so D can only be "between" A and B --
    // or close enough to that for one of the margins below to matter -- when |D - A| <= |B - A| (1 + 1e-2).  In the
    // reference's coordinates |Q - P|^2 = 2 (1 - cos(latQ - latP)) + 2 sP sQ (1 - cos(lonQ - lonP)), which the step and
    // the destination hold term by term (all terms >= 0 on one side of the equator: no cancellation);
    // 2 (1 - cos x) >= sin^2 x bounds |D - A|^2 from below with what is at hand.  A boat still more than a step and a
    // half from the destination -- every step of a rollout but its last one or two -- is decided here.
    const float ba2_half = fma_(sA * sB, m.h_dlon, -m.cm1_dlat);
    const float da2_low = fma_((sA * sD) * rd.h_dlon, 2.0f, rd.sin_dlat * rd.sin_dlat);
    if (ba2_half > 0.0f && da2_low > 4.0f * ba2_half) return 0;  // |D - A|^2 > 2 |B - A|^2
    const float dBy = sB * m.sin_dlon;
    const float q = fma_(-cA * sB, m.h_dlon, m.sin_dlat);
    const float r = fma_(cA * sD, rd.h_dlon, -rd.sin_dlat);
    const float dDy = sD * rd.sin_dlon;
    const float T = fma_(dBy, r, dDy * q);
    const float N2 = fma_(dBy, dBy, q * q);
    const float lhs = T * T, rhs = sin2_dest_angle * N2;
    if (!(N2 > 0.0f)) {  // A == B
        why |= kWhyDegenerate;
        return 0;
    }
    if (fabsf(lhs - rhs) <= margin_angle * rhs) why |= kWhyAngle;
    if (lhs > rhs) return 0;
    const float dBx = fma_(sA, m.cm1_dlat, fma_(cA, m.sin_dlat, -sB * m.h_dlon));
    const float dBz = fma_(cA, m.cm1_dlat, -sA * m.sin_dlat);
    const float cm1D = -(rd.x_dlat * rd.x_dlat) * cosc_poly(rd.x_dlat * rd.x_dlat);
    const float dDx = fma_(sA, cm1D, fma_(cA, rd.sin_dlat, -sD * rd.h_dlon));
    const float dDz = fma_(cA, cm1D, -sA * rd.sin_dlat);
    const float ab = fma_(dDx, dBx, fma_(dDy, dBy, dDz * dBz));
    const float dd = fma_(dDx, dDx, fma_(dDy, dDy, dDz * dDz));
    const float bb = fma_(dBx, dBx, fma_(dBy, dBy, dBz * dBz));
    const float p = ab - dd;  // (D - A).(B - D)
    if (fabsf(p) <= margin_along * bb) why |= kWhyAlong;
    if (p < 0.0f) {
        // B stopped short of D by less than kNearMiss of a step: the next bearing is taken almost on top
        // of the destination and amplifies the (tiny) position error by step / distance
        if (dd - 2.0f * ab + bb < near_miss2 * bb) why |= kWhyNearMiss;
        return 0;
    }
    frac = ab * rcp_approx(bb);
    // frac feeds the arrival time; the approximate reciprocal costs ~1e-7 relative
    return 1;
}

// ---- one transition -------------------------------------------------------------------------------
// The short series (sinc_small, cosc_small, asinc_small) hold to 3e-10 for angles below 0.12 rad; a
// longer step (|speed * dt / R| or the longitude increment above kMaxStepAngle) is not taken here.
constexpr float kMaxStepAngle = 0.12f;
constexpr int kStNeedsLiteral = 255;  // internal status: this boat / rollout continues in the fp64 kernel
// Simulator.doStep (simulatorTLKT.py:181-216).  g = grid position of (lat, lon); (sh, ch) heading;
// (s, c) = sin/cos(lat); z standard normal.  On PMCTS_ST_OK the state is advanced and m holds the
// increments; otherwise nothing changes.
// kTurbo: the caller guarantees that k can start a step and that the position is on the grid (a roll_safe
// scenario whose state passed the terminal test, or the host-checked root), and that all steps have the same
// length sc.dt_uniform32: no range checks, no load of the step length (kUniformDt alone: only the latter).
template <bool kShared = false, bool kDefBoat = false, bool kTurbo = false, bool kUniformDt = kTurbo, bool kRad = false>
PMCTS_HD int fast_step(const ScenarioView &sc, const BoatParams &bp, int &k, double &lat, double &lon,
                       const GridPos &g, float sh, float ch, float s, float c, float z, Move &m,
                       const float2 *slab = nullptr) {
    // instants a step can start from: inside the weather time axis and not the last one
    if (!kTurbo &&
        ((unsigned)(k - sc.ks_lo) > (unsigned)(sc.ks_hi - sc.ks_lo) || sc.ks_hi < sc.ks_lo || !g.inside)) {
        // which error the reference raises: IndexError before anything for a bad index, ValueError from the
        // wind query (simulatorTLKT.py:176-177), IndexError from times[k + 1] (:208)
        if (k < 0 || k >= sc.nT) return PMCTS_ST_PAST_HORIZON;
        if (!g.inside || k < sc.kv_lo || k > sc.kv_hi) return PMCTS_ST_OUT_OF_GRID;
        return PMCTS_ST_PAST_HORIZON;
    }
    float u, v;
    wind_at<kShared>(sc, k, g, u, v, slab);
    float speed = boat_speed<kDefBoat>(bp, u, v, sh, ch);
    speed = fma_(z, speed * bp.sigma_f, speed);  // Boat.addUncertainty (:504-517)
#if defined(__CUDA_ARCH__)
    const float dt = kUniformDt ? sc.dt_uniform32 : __ldg(sc.dt_sec32 + k);
#else
    const float dt = kUniformDt ? sc.dt_uniform32 : sc.dt_sec32[k];
#endif
    const float d = speed * dt * (kDefBoat ? defboat::inv_radius : bp.inv_radius_f);
    m = move_increments(d, sh, ch, s, c);
    // one exit for both limits (a step too long for the short series; at high latitude dlon = b / cos(lat) too large):
    // what is not a number fails either test -- see boat_speed and bearing_unit
    if (!(fabsf(d) <= kMaxStepAngle) | !(fabsf(m.sin_dlon) <= kMaxStepAngle)) return kStNeedsLiteral;
    if (kRad) {  // (position in radians)
        lat += (double)m.dlat;
        lon += (double)m.dlon;
    } else {
        lat = fma((double)m.dlat, 180.0 / kPi, lat);
        lon = fma((double)m.dlon, 180.0 / kPi, lon);
    }
    k += 1;
    return PMCTS_ST_OK;
}

// Heading given in degrees (an action of the tree, a plan, a K1 heading segment).  The dot-product
// point of sail equals the reference's abs / fold only for headings in [0, 360).
PMCTS_HD bool heading_unit(double deg, float &sh, float &ch) {
    if (!(deg >= 0.0 && deg < 360.0)) return false;
    // reduce to [-90, 90] exactly (multiples of 90 are exact in fp64), then the lat kernels
    double a = deg;
    int quad = 0;
    if (a > 315.0) { a -= 360.0; }
    else if (a > 225.0) { a -= 270.0; quad = 3; }
    else if (a > 135.0) { a -= 180.0; quad = 2; }
    else if (a > 45.0) { a -= 90.0; quad = 1; }
    float s, c;
    sincos_lat(a, s, c);
    switch (quad) {
        case 0: sh = s; ch = c; break;
        case 1: sh = c; ch = -s; break;
        case 2: sh = -s; ch = -c; break;
        default: sh = -c; ch = s; break;
    }
    return true;
}

// Hist.add bin (utils.py:34-47) of a reward, flagging values closer than `margin` to a bin edge.
PMCTS_HD int hist_bin_checked(const BoatParams &bp, double r, int &why) {
    const double r8 = r * 8.0;
    const double c = ceil(r8);
    if (r8 > 0.0 && r8 < 11.5 && (c - r8 < bp.margin_bin || r8 - (c - 1.0) < bp.margin_bin)) why |= kWhyBin;
    const int b = (int)c - 1;
    return b < 0 ? 0 : (b > 11 ? 11 : b);
}


// Tree.is_state_terminal (worker.py:417-436) on the fp32 grid coordinates of the position.
template <bool kDefBoat = false>
PMCTS_HD bool fast_terminal(const ScenarioView &sc, const BoatParams &bp, int k, const GridPos &g, int &why) {
    // (the horizon test is folded into the result rather than taken as an early exit: a branch around a dozen
    //  straight-line instructions costs a convergence barrier pair per step -- K2 0.940 -> 0.929 ms)
    const bool horizon = k >= sc.nT - 1;
    const float m = (kDefBoat ? defboat::margin_box : bp.margin_box) * sc.box_margin_scale;  // 1 + 0.01 * (largest grid index): fp32 spacing of fa / fo
    const float2 f = pk(g.fa, g.fo);
    const float2 d0 = pk_sub(f, pk(sc.box_fa_lo, sc.box_fo_lo)), d1 = pk_sub(pk(sc.box_fa_hi, sc.box_fo_hi), f);
    const float lo = fminf(fminf(d0.x, d1.x), fminf(d0.y, d1.y));
    why |= (!horizon && fabsf(lo) <= m) ? kWhyBox : 0;  // only the binding edge can flip
    return horizon || lo < 0.0f;
}

// Result of one rollout.
struct RollOut {
    int st, nstep, bin;
    int why;  // 0, or the kWhy* reasons this rollout is left to the literal kernel
    double ta, rw, frac, lat, lon;
};

// Fills the members of RolloutArgs every rollout of a launch would otherwise derive for itself: destination
// trigonometry, reward constants, and the root state seen through sincos_lat / dest_rel / grid_pos (the
// device's own fp32 formulas, evaluated once on the host -- they contain no hardware approximation, so the
// host and the device agree bit for bit).
inline void prepare_rollout(const ScenarioView &sc, RolloutArgs &a) {
    a.sin_dest_lat = (float)sin(a.dest_lat * kDegToRad);
    a.cos_dest_lat = (float)cos(a.dest_lat * kDegToRad);
    a.root_lat_rad = a.root_lat * kDegToRad;
    a.root_lon_rad = a.root_lon * kDegToRad;
    a.dest_lat_rad = a.dest_lat * kDegToRad;
    a.dest_lon_rad = a.dest_lon * kDegToRad;
    a.fa_c0_rel = fma(a.dest_lat_rad, sc.inv_dlat_rad, sc.fa_c0);
    a.fo_c0_rel = fma(a.dest_lon_rad, sc.inv_dlon_rad, sc.fo_c0);
    a.rw_t0 = sc.tmax * 1.001;
    a.rw_inv = (float)(1.0 / (sc.tmax * 1.001 - a.tmin));
    RootPre &p = a.root_pre;
    sincos_lat(a.root_lat, p.s, p.c);
    const DestRel rd = dest_rel(a.root_lat, a.root_lon, a.dest_lat, a.dest_lon);
    p.sin_dlat = rd.sin_dlat; p.sin_dlon = rd.sin_dlon; p.h_dlon = rd.h_dlon; p.x_dlat = rd.x_dlat;
    const GridPos g = grid_pos(sc, a.root_lat, a.root_lon);
    p.wa = g.wa; p.wo = g.wo; p.fa = g.fa; p.fo = g.fo;
    p.ia = g.ia; p.io = g.io; p.inside = g.inside ? 1 : 0;
}

// Whether a launch may use the kTurbo variant (see rollout_one).  Call after prepare_rollout.
inline bool turbo_ok(const ScenarioView &sc, const RolloutArgs &a);

// Steps between re-evaluations of sin / cos(lat) from the fp64 latitude; in between they are carried by the
// step's own identities (s' = s cos d + c a, c' = H).  They only scale increments, and their drift over 16
// steps stays below 2e-6 relative (tests/test_fastmodel.py).
constexpr int kRefreshTrig = 16;

inline bool turbo_ok(const ScenarioView &sc, const RolloutArgs &a) {
    const int k0 = (int)a.root_k;
    // (grids of at most 128 cells a side: the turbo loop splits the fp32 grid index into cell and weight, which is good
    //  to half an ulp of the index -- 4e-6 of a cell below 128, the level of the loop's other fp32 roundings)
    return sc.roll_safe && sc.dt_uniform32 > 0.0f && sc.nT - 1 <= kRefreshTrig && !(a.flags & PMCTS_ROLL_PLAN_EVAL) &&
           (double)k0 == a.root_k && k0 >= 0 && k0 <= sc.nT - 2 && a.root_pre.inside && sc.nlat <= 129 && sc.nlon <= 129;
}

// Tree.get_sim_to_estimate_state + Tree.default_policy (worker.py:255-300), estimate_perfomance_plan's
// inner loop (isochrones.py:498-530).  One loop serves every mode:
//   step j uses heading prefix[j] while j < n_fixed, the bearing to the destination afterwards;
//   the arrival / terminal tests run after step j once j + 1 >= n_unchecked.
// kNoise: PMCTS_NOISE_* or -1 (runtime); kLean: no trajectory output; kDefBoat: the reference's boat as
// immediates.  leaf = r / a.replicas.
// kTurbo (turbo_ok below, decided on the host): tree mode on a roll_safe scenario with equal steps, at most
// kRefreshTrig of them, and a root that can start a step -- every state the loop steps from has passed the
// terminal test (or is the root), so the step's range checks, the bounds test of the grid position, the load
// of the step length and the sin / cos refresh are compiled out.
// kFirstOnly (with kTurbo; the host passes it when PMCTS_ROLL_REPLAY_FULL_PREFIX is not set): at most the FIRST step has
// a fixed heading -- the reference's own behaviour, quirk Q1 -- so inside the loop the next heading is always the
// bearing to the destination and the loop carries no test of the step number.
template <int kNoise, bool kLean, bool kDefBoat = false, bool kTurbo = false, bool kFirstOnly = false>
PMCTS_HD void rollout_one(const ScenarioView &sc, const BoatParams &bp, const NoiseView &nz, const RolloutArgs &a,
                          int64_t n, int64_t r, int64_t leaf, RollOut &o) {
    // kTurbo: the position lives in radians inside the loop -- nothing but the histogram leaves a turbo launch, and
    // the differences to the destination and the grid index need no conversion then
    constexpr bool kRad = kTurbo;
    int k = (int)a.root_k;
    // ... and relative to the destination: the differences the arrival test and the next bearing start from are then
    // the state itself (two fp64 subtractions per step less)
    double lat = kRad ? a.root_lat_rad - a.dest_lat_rad : a.root_lat, lon = kRad ? a.root_lon_rad - a.dest_lon_rad : a.root_lon;
    const double dest_lat = kRad ? 0.0 : a.dest_lat, dest_lon = kRad ? 0.0 : a.dest_lon;
    const double *pre = a.prefix ? a.prefix + (size_t)leaf * a.prefix_stride : nullptr;
    // both loads are issued back to back: the first heading does not wait for the prefix length
    double fixed_heading = (pre && a.prefix_stride > 0) ? pre[0] : 0.0;
    const int plen = a.prefix_len ? a.prefix_len[leaf] : 0;
    const bool plan = (a.flags & PMCTS_ROLL_PLAN_EVAL) != 0;
    const int n_fixed = plan ? plen : ((a.flags & PMCTS_ROLL_REPLAY_FULL_PREFIX) ? plen : imin(plen, 1));
    const int n_unchecked = plan ? imax(plen, 1) : 1;
    const float sD = a.sin_dest_lat, cD = a.cos_dest_lat;
    o.st = PMCTS_ST_OK;
    o.nstep = 0;
    o.why = 0;
    o.frac = 0.0;
    int at = 0;
    float frac = 0.0f;
    if (!plan && plen == 0) {
        // worker.py:264-265 on prevState == state: the plane is degenerate (ZeroDivisionError)
        o.st = PMCTS_ST_DEGENERATE;
    } else if (a.root_lat == 0.0 || a.root_lon == 0.0) {
        o.why = kWhyDegenerate;  // ya == 0 in worker.py:396: left to the literal kernel
    } else {
        Noise ns;
        const RootPre &rp = a.root_pre;
        float sA = rp.s, cA = rp.c;
        DestRel rd;
        rd.sin_dlat = rp.sin_dlat; rd.sin_dlon = rp.sin_dlon; rd.h_dlon = rp.h_dlon; rd.x_dlat = rp.x_dlat;
        GridPos g;
        g.ia = rp.ia; g.io = rp.io; g.wa = rp.wa; g.wo = rp.wo; g.fa = rp.fa; g.fo = rp.fo; g.inside = rp.inside != 0;
        int trig_left = kRefreshTrig;
        // (not unrolled: two or four steps per pass spare the hand-over moves and the Philox block test, but measured
        //  slower -- 0.947-0.955 ms against 0.903 -- at 64 as at 72 registers)
        // The heading of the step about to be taken (step o.nstep): a fixed one while the prefix lasts, the bearing to
        // the destination afterwards.  It is worked out at the END of the step before -- next to the destination's
        // relative position it starts from -- so that the loop's head has no branch on the step number.
        float sh = 0.0f, ch = 1.0f;
        auto heading = [&](bool first) -> bool {
            if ((!kFirstOnly || first) && o.nstep < n_fixed) {
                if (!heading_unit(fixed_heading, sh, ch)) return false;
                if (o.nstep + 1 < n_fixed) fixed_heading = pre[o.nstep + 1];
            } else {
                bearing_unit(rd, sA, sD, cD, sh, ch);
            }
            return true;
        };
        if (!heading(true)) o.why |= kWhyStep;
        else for (;;) {
            const float z = ns.draw<kNoise>(nz, n, r, o.nstep);
            Move m;
            o.st = fast_step<false, kDefBoat, kTurbo, kTurbo, kRad>(sc, bp, k, lat, lon, g, sh, ch, sA, cA, z, m);
            if (o.st == kStNeedsLiteral) {
                o.st = PMCTS_ST_OK;
                o.why |= kWhyStep;
                break;
            }
            if (o.st != PMCTS_ST_OK) break;
            if (!kLean) {
                if (a.traj_lat && o.nstep < a.traj_stride) a.traj_lat[(size_t)o.nstep * n + r] = kRad ? (lat + a.dest_lat_rad) * (180.0 / kPi) : lat;
                if (a.traj_lon && o.nstep < a.traj_stride) a.traj_lon[(size_t)o.nstep * n + r] = kRad ? (lon + a.dest_lon_rad) * (180.0 / kPi) : lon;
            }
            ++o.nstep;
            float sB = m.s_new, cB = m.c_new;
            if (!kTurbo && --trig_left == 0) {  // (a turbo rollout has at most kRefreshTrig steps)
                trig_left = kRefreshTrig;
                sincos_lat(lat, sB, cB);
            }
            g = grid_pos<!kTurbo, kRad>(sc, lat, lon, a.fa_c0_rel, a.fo_c0_rel);
            if (kTurbo || o.nstep >= n_unchecked) {  // (turbo: tree mode, every step is checked)
                at = fast_at_dest<kDefBoat>(bp, rd, m, sA, cA, sB, sD, frac, o.why);
                if (at || o.why) break;
                if (fast_terminal<kDefBoat>(sc, bp, k, g, o.why)) break;
                if (o.why) break;
            }
            sA = sB;
            cA = cB;
            rd = dest_rel<kRad>(lat, lon, dest_lat, dest_lon);
            if (!heading(false)) {
                o.why |= kWhyStep;
                break;
            }
        }
    }
    o.ta = NAN;
    o.rw = 0.0;
    if (o.st == PMCTS_ST_OK && !o.why) {
        if (at) {
            const double tk = sc.times[k];
            o.frac = (double)frac;
            o.ta = tk - (1.0 - o.frac) * (tk - sc.times[k - 1]);  // worker.py:274-275
            // reward (worker.py:276-277): the difference of times in fp64, the exponential in fp32
            o.rw = (double)((exp_approx((float)(a.rw_t0 - o.ta) * a.rw_inv) - 1.0f) * 0.58197670686932642f);
            o.st = PMCTS_ST_ARRIVED;
        } else {
            o.st = (k >= sc.nT - 1) ? PMCTS_ST_HORIZON : PMCTS_ST_OUT_OF_BOX;
            if (plan) o.ta = sc.tmax;  // isochrones.py:527-530
        }
    }
    o.bin = hist_bin_checked(bp, o.rw, o.why);
    o.lat = kRad ? (lat + a.dest_lat_rad) * (180.0 / kPi) : lat;
    o.lon = kRad ? (lon + a.dest_lon_rad) * (180.0 / kPi) : lon;
}

}  // namespace fast

// Default decision margins (calibrated in tests/test_fastmodel.py: >= 30x the largest error seen).
inline void default_fast_margins(BoatParams &p) {
    p.margin_angle = fast::defboat::margin_angle;
    p.margin_along = fast::defboat::margin_along;
    p.near_miss2 = fast::defboat::near_miss2;
    p.margin_box = fast::defboat::margin_box;
    p.margin_bin = fast::defboat::margin_bin;
}

}  // namespace pmcts
