// lbl_core.cuh -- arithmetic shared by the sm_100a kernels (and by the CPU emulation
// harness under tests/emu, which compiles this header as plain C++ to debug kernel logic
// on machines without a GPU; the product never runs that build).
//
// Reference citations are relative to /root/reference/pyLBL/c_lib/.
#pragma once

#include <math.h>
#include <stdint.h>
// Opt-in index checks at the places where a kernel indexes shared memory, the record arrays or
// the spectrum from computed ranges (build with -DLBL_DEBUG_BOUNDS; compute-sanitizer is not
// available on every pool).  A failing check stops the kernel with a device-side assert.
#ifdef LBL_DEBUG_BOUNDS
#include <assert.h>
#define LBL_CHECK(cond) assert(cond)
#else
#define LBL_CHECK(cond) ((void)0)
#endif

#if defined(__CUDACC__)
#define LBL_HD __host__ __device__ __forceinline__
#define LBL_HD_NOINLINE __host__ __device__ __noinline__
#else
#define LBL_HD inline
#define LBL_HD_NOINLINE
#endif

#if !defined(__CUDACC__)
// Plain-C++ stand-ins for the CUDA vector types (CPU emulation build only).
struct alignas(16) double2 { double x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) float4 { float x, y, z, w; };
#endif

namespace lbl
{

// ---- constants (same doubles the reference obtains at run time) --------------------
constexpr double kPi = 0x1.921fb54442d18p+1;      // M_PI
constexpr double kRsqrPi = 0x1.20dd750429b6dp-1;  // 1./sqrt(M_PI)           voigt.c:7
constexpr double kSqrLn2 = 0x1.aa4499161cd47p-1;  // sqrt(log(2.))           voigt.c:8
constexpr double kR2 = 0x1.683271f84129fp+13;     // 2*log(2)*8314.472       spectra.c:14
constexpr double kVlight = 2.99792458e8;          //                         spectra.c:12
constexpr double kPaToAtm = 9.86923e-6;           //                         spectra.c:13
constexpr double kC2 = 1.4387752;                 //                         spectra.c:15
constexpr double kBig = 1.0e150;                  // "infinite" denominator: masks a term
constexpr double kLorentzY = 70.55;               // voigt.c:17
constexpr int kIdxClamp = 1 << 30;

// ---- records produced by the scaling kernel, consumed by the summation kernels ------
// All arrays are [layer][line] with lines in ascending order of unshifted centre.
struct alignas(16) FarAB   // far-wing (Lorentz) form: term = 1/((v*a + b)^2 + c)
{
    double a, b;
};
struct alignas(16) LineChk  // integer bookkeeping of one (layer, line)
{
    int cb;    // floor(nu') - v0 : cell of the shifted centre (decides the window, spectra.c:48-62)
    int nlo;   // grid indices [nlo, nhi]: points that must take the full Humlicek path
    int nhi;
    int pad;
};
struct alignas(16) LineGen  // operands of the full profile (voigt.c:13-15,34-43,188)
{
    double nu;      // shifted centre
    double repwid;  // sqrt(ln2)/alpha
    double y;       // repwid*gamma
    double cof;     // sw*rsqrpi*repwid
    double xlim0;   // sqrt(15100 + y*(40 - 3.6*y))                  voigt.c:34
    double xlim1;   // y >= 8.425 ? 0 : sqrt(164 - y*(4.3 + 1.8*y))  voigt.c:36-43
};

struct alignas(16) Far32  // FP32-mode operands of one (layer, line): term = amp/((v-nu')^2 + g2)
{
    float cbf;   // floor(nu') - v0 as a float (exact below 2^24)
    float frac;  // nu' - floor(nu') in [0, 1)
    float g2;    // gamma^2
    float amp;   // A * 2^shift(layer), A = sw'*gamma/pi   (0 masks the line)
};

struct LayerIn
{
    double pressure;     // [Pa]
    double temperature;  // [K]
    double vmr;          // [mol mol-1]
    double slack;        // >= max |nu' - nu| over this layer's lines [cm-1]
    double kappa;        // near-zone half width <= kappa * nu   [dimensionless]
    double pad;
};

// ---- exact (non-contracted) IEEE operations where bit parity is required ------------
LBL_HD double mul_rn(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
LBL_HD double add_rn(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}

// Grid point i of the reference grid: v[i] = v0 + i*dv, dv = 1./n_per_v (absorption.c:33-39).
LBL_HD double grid_point(int v0, double dv, int i)
{
    return add_rn((double)v0, mul_rn((double)i, dv));
}

// ---- reciprocal seed ------------------------------------------------------------------
// rcp.approx.ftz.f64 (SASS MUFU.RCP64H): uses only the upper 32 bits of the operand and
// returns an approximation whose lower 32 bits are zero.  One Newton step in the caller
// squares its relative error.
LBL_HD double rcp_seed(double q)
{
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(q));
    return r;
#else
    // Emulation: operand and result truncated to their high words (20 mantissa bits).
    union { double d; uint64_t u; } in, out;
    in.d = q;
    in.u &= 0xFFFFFFFF00000000ull;
    out.d = 1.0 / in.d;
    out.u &= 0xFFFFFFFF00000000ull;
    return out.d;
#endif
}

// Seed whose low word is borrowed from `dead` (any double that is no longer needed).  The
// hardware instruction writes only the high word; rcp_seed() has to zero the low word with an
// extra MOV per evaluation, whereas arbitrary low bits merely perturb the seed by < 2^-20,
// which the Newton step squares away exactly like the seed's own error.
LBL_HD double rcp_seed_lo(double q, double dead)
{
#if defined(__CUDA_ARCH__)
    double r;
    asm("{\n\t"
        ".reg .b32 lo, hi, zlo, zhi;\n\t"
        ".reg .f64 t;\n\t"
        "rcp.approx.ftz.f64 t, %1;\n\t"
        "mov.b64 {zlo, zhi}, t;\n\t"
        "mov.b64 {lo, hi}, %2;\n\t"
        "mov.b64 %0, {lo, zhi};\n\t"
        "}"
        : "=d"(r)
        : "d"(q), "d"(dead));
    return r;
#else
    union { double d; uint64_t u; } in, out, low;
    in.d = q;
    in.u &= 0xFFFFFFFF00000000ull;
    out.d = 1.0 / in.d;
    low.d = dead;
    out.u = (out.u & 0xFFFFFFFF00000000ull) | (low.u & 0x00000000FFFFFFFFull);
    return out.d;
#endif
}

LBL_HD double fma_(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}

// One far-wing evaluation, 4 FP64-pipe instructions + 1 MUFU:
//   d = v*a + b ; q = d*d + c ; r0 ~ 1/q ; acc += r0*(2 - q*r0)
// With a = 1/sqrt(A), b = -nu'*a, c = gamma^2/A this adds A/((v-nu')^2 + gamma^2), i.e.
// the reference's Lorentz/region-0 term (voigt.c:24, voigt.c:82 with :188) to ~1e-12.
LBL_HD double far_term(double v, double a, double b, double c, double acc)
{
    double d = fma_(v, a, b);
    double q = fma_(d, d, c);
    double r0 = rcp_seed(q);
    double e2 = fma_(-q, r0, 2.0);
    return fma_(r0, e2, acc);
}

// The same term with the seed's low word borrowed from d, as far_terms() forms it: the near-zone
// kernels take the Lorentz form back with THIS function at points where the summation kernel
// added it through far_terms(), so that both sides hold the same bits.  (Close to a line centre
// at low pressure the Lorentz form is 1/(sqrt(pi)*y) times the true profile, 1e4 and more: a
// difference of 1e-12 between two ways of forming it would be 1e-8 of the result.)
LBL_HD double far_term_lo(double v, double a, double b, double c, double acc)
{
    double d = fma_(v, a, b);
    double q = fma_(d, d, c);
    double r0 = rcp_seed_lo(q, d);
    double e2 = fma_(-q, r0, 2.0);
    return fma_(r0, e2, acc);
}

// The same arithmetic for P points at once, written stage by stage so that the P
// independent dependency chains are interleaved in the instruction stream (the in-order
// warp scheduler cannot do that by itself): all d, then all q, all seeds, all corrections.
template <int P>
LBL_HD void far_terms(const double (&v)[P], double a, double b, double c, double (&acc)[P])
{
    double d[P], q[P], r[P];
#pragma unroll
    for (int p = 0; p < P; ++p)
    {
        d[p] = fma_(v[p], a, b);
    }
#pragma unroll
    for (int p = 0; p < P; ++p)
    {
        q[p] = fma_(d[p], d[p], c);
    }
#pragma unroll
    for (int p = 0; p < P; ++p)
    {
        r[p] = rcp_seed_lo(q[p], d[p]);
    }
#pragma unroll
    for (int p = 0; p < P; ++p)
    {
        q[p] = fma_(-q[p], r[p], 2.0);
    }
#pragma unroll
    for (int p = 0; p < P; ++p)
    {
        acc[p] = fma_(r[p], q[p], acc[p]);
    }
}

// Reciprocal to ~1e-12: hardware seed (2^-20) and one Newton step.  Used by the full profile
// below, where the reference divides (voigt.c:82,95,113,145,159-183); three orders of
// magnitude inside the 1e-9 parity target.
LBL_HD double rcp_newton2(double q)
{
    const double r = rcp_seed(q);
    return r * fma_(-q, r, 2.0);
}

// Two lines at once for P points: 1/q1 + 1/q2 = (q1 + q2)/(q1*q2) needs ONE reciprocal.
// 9 FP64-pipe instructions and 1 MUFU per two evaluations (instead of 8 and 2): the MUFU
// pipe, which a one-reciprocal-per-evaluation loop saturates together with the FP64 pipe,
// is relieved at the price of half an FMA per evaluation.
template <int P>
LBL_HD void far_terms_pair(const double (&v)[P], double a1, double b1, double c1, double a2,
                           double b2, double c2, double (&acc)[P])
{
    double d[P], s[P], m[P], r[P];
#pragma unroll
    for (int p = 0; p < P; ++p)
    {
        d[p] = fma_(v[p], a1, b1);
        m[p] = fma_(v[p], a2, b2);
    }
#pragma unroll
    for (int p = 0; p < P; ++p)
    {
        d[p] = fma_(d[p], d[p], c1);   // q1
        m[p] = fma_(m[p], m[p], c2);   // q2
    }
#pragma unroll
    for (int p = 0; p < P; ++p)
    {
        s[p] = d[p] + m[p];
        m[p] = d[p] * m[p];
    }
#pragma unroll
    for (int p = 0; p < P; ++p)
    {
        r[p] = rcp_seed_lo(m[p], d[p]);
    }
#pragma unroll
    for (int p = 0; p < P; ++p)
    {
        m[p] = fma_(-m[p], r[p], 2.0);
        s[p] = s[p] * r[p];
    }
#pragma unroll
    for (int p = 0; p < P; ++p)
    {
        acc[p] = fma_(s[p], m[p], acc[p]);
    }
}

// ---- opt-in FP32 far-wing arithmetic ----------------------------------------------------------
// A wavenumber difference cannot be formed in FP32 from absolute wavenumbers (ulp(5000) =
// 5e-4 cm-1), so positions are split into an integer cell and a fraction: for a point in
// cell `cellf` at fraction `fp`, and a line in cell `cbf` at fraction `frac`,
//     v - nu' = ((cellf - cbf) - frac) + fp
// where the first bracket is exact (small integers) and the absolute error of the whole is
// <= 2e-6 cm-1 -- against a far-zone distance of at least ~0.1 cm-1.  Amplitudes carry a
// per-layer power-of-two scale so that they sit mid-range in FP32.
LBL_HD float rcp_f32(float x)
{
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}

LBL_HD float fmaf_(float a, float b, float c)
{
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}

// Two lines at P points: amp1/q1 + amp2/q2 = (amp1*q2 + amp2*q1)/(q1*q2): 8 FP32 + 1 MUFU.
template <int P>
LBL_HD void far32_pair(const float (&fp)[P], float t1, float g1, float a1, float t2, float g2,
                       float a2, float (&acc)[P])
{
#pragma unroll
    for (int p = 0; p < P; ++p)
    {
        const float u1 = t1 + fp[p];
        const float u2 = t2 + fp[p];
        const float q1 = fmaf_(u1, u1, g1);
        const float q2 = fmaf_(u2, u2, g2);
        const float m = q1 * q2;
        const float n = fmaf_(a2, q1, a1 * q2);
        acc[p] = fmaf_(n, rcp_f32(m), acc[p]);
    }
}

template <int P>
LBL_HD void far32_one(const float (&fp)[P], float t1, float g1, float a1, float (&acc)[P])
{
#pragma unroll
    for (int p = 0; p < P; ++p)
    {
        const float u1 = t1 + fp[p];
        acc[p] = fmaf_(a1, rcp_f32(fmaf_(u1, u1, g1)), acc[p]);
    }
}

// ---- full Humlicek W4 / CPF12 profile for one point (voigt.c:74-188) -------------------
// Only reached for y < 70.55 (the y >= 70.55 branch is pure Lorentz and has an empty near
// zone).  The region limits that need a square root are computed once per (layer, line) by
// the scaling kernel.  The profile is split at xlim1 into
//   outer : region 0 (Lorentz form) and Humlicek W4 region 1  -- ~98 % of the near zone, cheap
//   core  : W4 regions 2, 3 and CPF12                          -- a few points per line, long
// so that the near-zone kernel can run the long part with the lanes packed.

// voigt.c:48-53: for y <= 1e-6 the W4 regions 1 and 2 are switched off.
LBL_HD double voigt_outer_limit(double y, double xlim0, double xlim1)
{
    return (y <= 0.000001) ? xlim0 : xlim1;
}

// Regions 0 and 1 (voigt.c:79-97); requires abx >= voigt_outer_limit().  Returns K(x,y).
LBL_HD double voigt_outer(double abx, double xq, double y, double xlim0)
{
    const double yq = y * y;
    if (abx >= xlim0)
    {
        return (y * kRsqrPi) * rcp_newton2(xq + yq);
    }
    const double a0 = yq + 0.5;
    const double d0 = a0 * a0;
    const double d2 = yq + yq - 1.;
    const double d = kRsqrPi * rcp_newton2(d0 + xq * (d2 + xq));
    return d * y * (a0 + xq);
}

// W4 region 2 (voigt.c:98-115): |x| in [xlim2, xlim1).  Returns K(x,y).
LBL_HD double voigt_region2(double xq, double y)
{
    const double yq = y * y;
    const double h0 = 0.5625 + yq * (4.5 + yq * (10.5 + yq * (6.0 + yq)));
    const double h2 = -4.5 + yq * (9.0 + yq * (6.0 + yq * 4.0));
    const double h4 = 10.5 - yq * (6.0 - yq * 6.0);
    const double h6 = -6.0 + yq * 4.0;
    const double e0 = 1.875 + yq * (8.25 + yq * (5.5 + yq));
    const double e2 = 5.25 + yq * (1.0 + yq * 3.0);
    const double e4 = 0.75 * h6;
    const double d = kRsqrPi * rcp_newton2(h0 + xq * (h2 + xq * (h4 + xq * (h6 + xq))));
    return d * y * (e0 + xq * (e2 + xq * (e4 + xq)));
}

// voigt.c:44,48-53: lower limit of region 2 (regions 1 and 2 are off for y <= 1e-6).
LBL_HD double voigt_region2_limit(double y, double xlim0)
{
    return (y <= 0.000001) ? xlim0 : 6.8 - y;
}

// W4 region 3 (voigt.c:116-147): |x| < 2.4*y.  Returns K(x,y).
LBL_HD double voigt_region3(double xq, double y)
{
    const double z0 = 272.1014 + y * (1280.829 + y * (2802.870 + y * (3764.966
                      + y * (3447.629 + y * (2256.981 + y * (1074.409 + y * (369.1989
                      + y * (88.26741 + y * (13.39880 + y)))))))));
    const double z2 = 211.678 + y * (902.3066 + y * (1758.336 + y * (2037.310
                      + y * (1549.675 + y * (793.4273 + y * (266.2987
                      + y * (53.59518 + y * 5.0)))))));
    const double z4 = 78.86585 + y * (308.1852 + y * (497.3014 + y * (479.2576
                      + y * (269.2916 + y * (80.39278 + y * 10.0)))));
    const double z6 = 22.03523 + y * (55.02933 + y * (92.75679 + y * (53.59518
                      + y * 10.0)));
    const double z8 = 1.496460 + y * (13.39880 + y * 5.0);
    const double p0 = 153.5168 + y * (549.3954 + y * (919.4955 + y * (946.8970
                      + y * (662.8097 + y * (328.2151 + y * (115.3772 + y * (27.93941
                      + y * (4.264678 + y * 0.3183291))))))));
    const double p2 = -34.16955 + y * (-1.322256 + y * (124.5975 + y * (189.7730
                      + y * (139.4665 + y * (56.81652 + y * (12.79458
                      + y * 1.2733163))))));
    const double p4 = 2.584042 + y * (10.46332 + y * (24.01655 + y * (29.81482
                      + y * (12.79568 + y * 1.9099744))));
    const double p6 = -0.07272979 + y * (0.9377051 + y * (4.266322 + y * 1.273316));
    const double p8 = 0.0005480304 + y * 0.3183291;
    // The reference writes sqrt(pi) as the 8-digit literal here (voigt.c:145).
    const double d = 1.7724538 * rcp_newton2(z0 + xq * (z2 + xq * (z4 + xq * (z6 + xq * (z8 + xq)))));
    return d * (p0 + xq * (p2 + xq * (p4 + xq * (p6 + xq * p8))));
}

// CPF12 (voigt.c:148-186): 2.4*y <= |x| < xlim2; kSubI: |x| <= 18.1*y + 1.65 (sub-region i),
// else sub-region ii with the exp(-x^2) term.  Returns K(x,y).
template <bool kSubI>
LBL_HD double voigt_cpf12(double xi, double y)
{
    const double tc[6] = {0.31424038, 0.94778839, 1.5976826, 2.2795071, 3.0206370, 3.8897249};
    const double cc[6] = {1.0117281, -0.75197147, 0.012557727,
                          0.010022008, -0.00024206814, 0.00000050084806};
    const double sc[6] = {1.393237, 0.23115241, -0.15535147,
                          0.0062183662, 0.000091908299, -0.00000062752596};
    const double y0 = 1.5;
    const double ypy0 = y + y0;
    const double ypy0q = ypy0 * ypy0;
    const double yf = y + (y0 + y0);
    const double y0q = y0 * y0;
    double buf = 0.;
#pragma unroll
    for (int j = 0; j < 6; ++j)
    {
        const double dm = xi - tc[j];
        const double mq = dm * dm;
        const double mf = rcp_newton2(mq + ypy0q);
        const double xm = mf * dm;
        const double ym = mf * ypy0;
        const double dp = xi + tc[j];
        const double pq = dp * dp;
        const double pf = rcp_newton2(pq + ypy0q);
        const double xp = pf * dp;
        const double yp = pf * ypy0;
        if (kSubI)
        {
            buf += cc[j] * (ym + yp) - sc[j] * (xm - xp);
        }
        else
        {
            buf += (cc[j] * (mq * mf - y0 * ym) + sc[j] * yf * xm) * rcp_newton2(mq + y0q)
                 + (cc[j] * (pq * pf - y0 * yp) - sc[j] * yf * xp) * rcp_newton2(pq + y0q);
        }
    }
    if (!kSubI)
    {
        buf = y * buf + exp(-(xi * xi));
    }
    return buf;
}

// Which of the three the point takes: 0 region 3, 1 CPF12 sub-region i, 2 sub-region ii.
LBL_HD int voigt_inner_kind(double abx, double y)
{
    if (abx < 2.4 * y) return 0;
    return (abx <= 18.1 * y + 1.65) ? 1 : 2;
}

// W4 region 3 and CPF12 (voigt.c:116-186): |x| < xlim2.  Returns K(x,y).
static LBL_HD_NOINLINE double voigt_inner(double xi, double y)
{
    const double abx = fabs(xi);
    const int kind = voigt_inner_kind(abx, y);
    if (kind == 0) return voigt_region3(abx * abx, y);
    if (kind == 1) return voigt_cpf12<true>(xi, y);
    return voigt_cpf12<false>(xi, y);
}

// Regions 2, 3 and CPF12 (voigt.c:98-186); requires abx < voigt_outer_limit().  Returns K(x,y).
LBL_HD double voigt_core(double xi, double y, double xlim0)
{
    const double abx = fabs(xi);
    if (abx >= voigt_region2_limit(y, xlim0))
    {
        return voigt_region2(abx * abx, y);
    }
    return voigt_inner(xi, y);
}

// The whole profile for one point: cof*K(x,y)  (voigt.c:76-188).
LBL_HD double voigt_general(double v, double nu, double repwid, double y, double cof,
                            double xlim0, double xlim1)
{
    const double xi = (v - nu) * repwid;
    const double abx = fabs(xi);
    if (abx >= voigt_outer_limit(y, xlim0, xlim1))
    {
        return cof * voigt_outer(abx, abx * abx, y, xlim0);
    }
    return cof * voigt_core(xi, y, xlim0);
}

// The same as a real call: for kernels that need it at a rare point only (the pedestal sums:
// a node inside a line's near zone) and would otherwise carry its registers through their loops.
static LBL_HD_NOINLINE double voigt_general_call(double v, double nu, double repwid, double y, double cof,
                                                 double xlim0, double xlim1)
{
    return voigt_general(v, nu, repwid, y, cof, xlim0, xlim1);
}

// ---- per-(layer, line) scaling: spectra.c:17-45 plus the derived records -----------------
struct LineIn
{
    double nu, sw, gamma_air, gamma_self, n_air, elower, delta_air, mass;
};

// Linear interpolation on the 1-K TIPS axis (spectral_database.c:97-104).
LBL_HD double tips_interp(const double* t, const double* q, double temperature)
{
    const int i = (int)floor(temperature) - (int)t[0];
    return q[i] + (q[i + 1] - q[i]) * (temperature - t[i]) / (t[i + 1] - t[i]);
}

LBL_HD void scale_line(const LineIn& ln, const LayerIn& ly, double q_ref, double q_t,
                       int v0, int n_per_v,
                       FarAB& ab, double& cc, LineChk& chk, LineGen& gen)
{
    const double p = mul_rn(ly.pressure, kPaToAtm);                 // spectra.c:17
    const double pp = p * ly.vmr;                                    // :18
    const double tfact = 296. / ly.temperature;                      // :19
    const double nu = add_rn(ln.nu, mul_rn(p, ln.delta_air));        // :22 (bit-exact)
    const double gamma = (ln.gamma_air * (p - pp) + ln.gamma_self * pp) * pow(tfact, ln.n_air);
    const double alpha = (ln.nu / kVlight) * sqrt(kR2 * ly.temperature / ln.mass);   // :29
    const double sb = exp(ln.elower * kC2 * (ly.temperature - 296.) / (ly.temperature * 296.));
    const double g = exp((-kC2 * ln.nu) / ly.temperature);
    const double gref = exp((-kC2 * ln.nu) / 296.);
    const double se = (1. - g) / (1. - gref);
    const double sq = q_ref / q_t;
    const double sw = ln.sw * sb * se * sq * 0.01 * 0.01;            // :45

    const double repwid = kSqrLn2 / alpha;                           // voigt.c:13
    const double y = repwid * gamma;                                 // voigt.c:14
    const double cof = sw * kRsqrPi * repwid;                        // voigt.c:188
    gen.nu = nu;
    gen.repwid = repwid;
    gen.y = y;
    gen.cof = cof;
    gen.xlim0 = (y < kLorentzY) ? sqrt(15100. + y * (40. - y * 3.6)) : 0.;
    gen.xlim1 = (y >= 8.425) ? 0. : sqrt(164. - y * (4.3 + y * 1.8));

    // Window cell (spectra.c:48): floor(nu') relative to the first grid wavenumber.
    double cbd = floor(nu) - (double)v0;
    cbd = fmin(fmax(cbd, -(double)kIdxClamp), (double)kIdxClamp);
    chk.cb = (int)cbd;
    chk.pad = 0;

    // Lorentz amplitude in x-space: term = ax/(x^2+y^2), ax = cof*y*rsqrpi (voigt.c:82,:188;
    // the y >= 70.55 branch voigt.c:24 is the same quantity).
    // The summation multiplies two denominators q = (v*a+b)^2 + c (far_terms_pair), so q
    // must stay below ~1e150: lines so weak that a > 1e70 (amplitude below ~1e-130 m2, a
    // hundred orders of magnitude under any HITRAN intensity) are dropped.
    const double ax = cof * (y * kRsqrPi);
    bool usable = ax >= 1.0e-200 && ax <= 1.0e100 && repwid <= 1.0e50 && y <= 1.0e50;
    if (usable)
    {
        const double inv = 1. / ax;
        ab.a = repwid * sqrt(inv);
        ab.b = -nu * ab.a;
        cc = (y * y) * inv;
        usable = ab.a <= 1.0e70 && cc <= 1.0e140 && fabs(ab.b) <= 1.0e74;
    }
    if (!usable)
    {
        ab.a = 0.;
        ab.b = 0.;
        cc = kBig;
    }

    // Near zone: every grid point with |x| < xlim0 must take the full profile (voigt.c:79).
    // Enclose it in integer grid indices with one point of margin either side.
    if (y >= kLorentzY || !(repwid > 0.) || !(y == y))
    {
        chk.nlo = 0x7fffffff;
        chk.nhi = -0x7fffffff;
    }
    else
    {
        const double half = (gen.xlim0 / repwid) * (1. + 0x1p-30);
        double lo = floor((nu - half - (double)v0) * (double)n_per_v) - 1.;
        double hi = ceil((nu + half - (double)v0) * (double)n_per_v) + 1.;
        lo = fmin(fmax(lo, -(double)kIdxClamp), (double)kIdxClamp);
        hi = fmin(fmax(hi, -(double)kIdxClamp), (double)kIdxClamp);
        chk.nlo = (int)lo;
        chk.nhi = (int)hi;
    }
}

// Contribution of one line to one grid point, choosing the path exactly as the summation
// kernel does (used by the pedestal recurrence so that both see the same numbers).
LBL_HD double line_point(double v, int i, const FarAB& ab, double cc, const LineChk& chk,
                         const LineGen& gen)
{
    if (i >= chk.nlo && i <= chk.nhi)
    {
        return voigt_general(v, gen.nu, gen.repwid, gen.y, gen.cof, gen.xlim0, gen.xlim1);
    }
    return far_term(v, ab.a, ab.b, cc, 0.0);
}

// First index j in [0, n) with x[j] >= key (n if none).
LBL_HD int lower_bound(const double* x, int n, double key)
{
    int lo = 0, hi = n;
    while (lo < hi)
    {
        const int mid = (lo + hi) >> 1;
        if (x[mid] < key)
        {
            lo = mid + 1;
        }
        else
        {
            hi = mid;
        }
    }
    return lo;
}

// Line ranges a group of threads covering grid points [first, last] must visit.
//   [j0,j1) checked | [j1,j2) plain | [j2,j3) checked | [j3,j4) plain | [j4,j5) checked
// "plain" lines are active for every point of the group and far from all of them.
struct Segments
{
    int j[6];
};

LBL_HD Segments find_segments(const double* nu_sorted, int n_lines, int v0, int n_per_v,
                              double dv, int cut_off, int first, int last, double slack,
                              double kappa)
{
    Segments s;
    const int c_lo = first / n_per_v;
    const int c_hi = last / n_per_v;
    const double base = (double)v0;
    // Superset of lines whose window can touch any point of the group.
    // (The extra cell cb == c-cut-1 that reaches only a cell's first point is K2b's.)
    s.j[0] = lower_bound(nu_sorted, n_lines, base + (double)(c_lo - cut_off) - slack);
    s.j[5] = lower_bound(nu_sorted, n_lines, base + (double)(c_hi + cut_off + 1) + slack);
    // Lines certainly inside every point's window.
    int core_lo = lower_bound(nu_sorted, n_lines, base + (double)(c_hi - cut_off) + slack);
    int core_hi = lower_bound(nu_sorted, n_lines, base + (double)(c_lo + cut_off + 1) - slack);
    if (core_lo < s.j[0]) core_lo = s.j[0];
    if (core_hi > s.j[5]) core_hi = s.j[5];
    if (core_lo >= core_hi)
    {
        // No common core (tiny cut-off or a group spanning many cells): check everything.
        s.j[1] = s.j[2] = s.j[3] = s.j[4] = s.j[5];
        return s;
    }
    // Lines that can be "near" some point of the group.
    const double v_first = base + (double)first * dv;
    const double v_last = base + (double)last * dv;
    // A line at nu is near the group only if nu - kappa*nu <= v_last, i.e. within
    // kappa*v_last/(1-kappa) above it (and less than that below it).
    const double reach = (kappa < 0.5)
        ? (kappa * fabs(v_last) / (1.0 - kappa)) * (1.0 + 0x1p-20) + slack + 3.0 * dv
        : 1.0e300;
    int near_lo = lower_bound(nu_sorted, n_lines, v_first - reach);
    int near_hi = lower_bound(nu_sorted, n_lines, v_last + reach);
    if (near_lo < core_lo) near_lo = core_lo;
    if (near_lo > core_hi) near_lo = core_hi;
    if (near_hi < near_lo) near_hi = near_lo;
    if (near_hi > core_hi) near_hi = core_hi;
    s.j[1] = core_lo;
    s.j[2] = near_lo;
    s.j[3] = near_hi;
    s.j[4] = core_hi;
    return s;
}

}  // namespace lbl
