// lbl_threads.cuh -- per-thread / per-warp bodies of the kernels, written so that the same
// code runs inside the sm_100a kernels (lbl_kernels.cuh) and inside the CPU emulation
// harness used to debug kernel logic without a GPU (tests/emu/emu.cpp).
#pragma once

#include "lbl_core.cuh"

namespace lbl
{

#if defined(__CUDA_ARCH__)
#define LBL_LDG(p) __ldg(p)
#else
#define LBL_LDG(p) (*(p))
#endif

// Packed line list on the device, ascending unshifted centre.
struct LinesView
{
    int n;  // active lines (the reference's DB-order prefix before its early break)
    const double* nu;
    const double* sw;
    const double* gamma_air;
    const double* gamma_self;
    const double* n_air;
    const double* elower;
    const double* delta_air;
    const double* mass;
    const int* iso;          // local_iso_id - 1 (TIPS block)
    const int* db_to_sorted; // sorted position of DB row r (nullptr = identity)
    // Per-wavenumber index of the sorted list: cell_first[k] = number of lines with
    // nu < cell_w0 + k, k in [0, cell_n); every line lies in [cell_w0, cell_w0 + cell_n - 1].
    // nullptr = none (plain binary search).
    const int* cell_first = nullptr;
    int cell_w0 = 0, cell_n = 0;
};

// Index of the first line with nu >= x (= lower_bound(nu, n, x)): the per-wavenumber index
// brackets it to the lines of one cm-1, so the search takes 3-4 probes instead of ~17.
LBL_HD int first_line_at(const LinesView& l, double x)
{
    int lo = 0, hi = l.n;
    if (l.cell_first)
    {
        const double k = floor(x) - (double)l.cell_w0;
        if (!(k >= 0.))
        {
            return 0;
        }
        if (k >= (double)(l.cell_n - 1))
        {
            return l.n;
        }
        lo = LBL_LDG(l.cell_first + (int)k);
        hi = LBL_LDG(l.cell_first + (int)k + 1);
    }
    while (lo < hi)
    {
        const int mid = (lo + hi) >> 1;
        if (LBL_LDG(l.nu + mid) < x)
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

struct TipsView
{
    int num_iso, num_t;
    const double* t;  // [num_iso*num_t]
    const double* q;
};

struct Records
{
    FarAB* ab;     // [layer][line]
    double* cc;    // [layer][line]
    LineChk* chk;  // [layer][line]
    LineGen* gen;  // [layer][line]
    Far32* f32;    // [layer][line], FP32 mode only (else nullptr)
    unsigned long long* amp_max;  // [layer] bit pattern of the largest amplitude A (FP32 mode)
};

struct GridSpec
{
    int v0, vn, n_per_v, cut_off;
    int n;      // (vn - v0)*n_per_v output points per layer
    int ncell;  // vn - v0
    double dv;  // 1./n_per_v
    // Spectral band of this call: only the cells [cell_lo, cell_hi) of the grid are computed
    // (0, ncell = the whole grid).  Windows, the active-line prefix and the pedestal
    // recurrence are always those of the WHOLE grid (spectra.c:48-62, absorption.c:80-83), so
    // the bands of a grid, computed separately, are the slices of the whole computation.
    int cell_lo, cell_hi;
};

// First cell of the far-field kernel's block that holds the band's first cell (blocks of
// `block_cells` consecutive cells at absolute positions of the grid).
LBL_HD int cell_block_base(const GridSpec& g, int block_cells) { return g.cell_lo - g.cell_lo % block_cells; }
LBL_HD int band_first_point(const GridSpec& g) { return g.cell_lo * g.n_per_v; }
LBL_HD int band_end_point(const GridSpec& g) { return g.cell_hi * g.n_per_v; }   // exclusive

// Lorentz amplitude A = sw'*gamma/pi of a record (0 for a dropped line).
LBL_HD double record_amplitude(const FarAB& ab)
{
    return (ab.a > 0.) ? 1. / (ab.a * ab.a) : 0.;
}

// ---------------------------------------------------------------------------------------
// K1: scaling of one (layer, line).  Returns the line's window size (e - s + 1, the
// reference's count of grid evaluations, spectra.c:48-62) for the eval counter.
// ---------------------------------------------------------------------------------------
LBL_HD long long scale_thread(const LinesView& ln, const TipsView& tips, const LayerIn* layers,
                              const GridSpec& g, const Records& rec, int layer, int j,
                              double& amplitude)
{
    const LayerIn ly = layers[layer];
    LineIn in;
    in.nu = ln.nu[j];
    in.sw = ln.sw[j];
    in.gamma_air = ln.gamma_air[j];
    in.gamma_self = ln.gamma_self[j];
    in.n_air = ln.n_air[j];
    in.elower = ln.elower[j];
    in.delta_air = ln.delta_air[j];
    in.mass = ln.mass[j];
    const int iso = ln.iso[j];
    const double* tt = tips.t + (size_t)iso * tips.num_t;
    const double* qq = tips.q + (size_t)iso * tips.num_t;
    const double q_ref = tips_interp(tt, qq, 296.);
    const double q_t = tips_interp(tt, qq, ly.temperature);

    FarAB ab;
    double cc;
    LineChk chk;
    LineGen gen;
    scale_line(in, ly, q_ref, q_t, g.v0, g.n_per_v, ab, cc, chk, gen);
    const size_t o = (size_t)layer * ln.n + j;
    rec.ab[o] = ab;
    rec.cc[o] = cc;
    rec.chk[o] = chk;
    rec.gen[o] = gen;
    amplitude = record_amplitude(ab);

    // Reference window (spectra.c:48-62) for the evaluation count.
    long long s = (long long)(chk.cb - g.cut_off) * g.n_per_v;
    if (s >= g.n)
    {
        return 0;
    }
    if (s < 0)
    {
        s = 0;
    }
    long long e = (long long)(chk.cb + g.cut_off + 1) * g.n_per_v;
    if (e >= g.n)
    {
        e = g.n - 1;
    }
    // the band's share of the window
    if (s < (long long)band_first_point(g)) s = band_first_point(g);
    if (e > (long long)band_end_point(g) - 1) e = (long long)band_end_point(g) - 1;
    return (e >= s) ? (e - s + 1) : 0;
}

// Per-layer power-of-two scale of the FP32 amplitudes: the largest amplitude maps to ~2^40.
LBL_HD int amp_shift(unsigned long long amp_max_bits)
{
    union { unsigned long long u; double d; } x;
    x.u = amp_max_bits;
    if (!(x.d > 0.))
    {
        return 0;
    }
    int e;
    frexp(x.d, &e);
    return 40 - e;
}

// K1f: FP32 operands of one (layer, line) from its FP64 records.
LBL_HD void far32_thread(const Records& rec, int n_lines, int layer, int j)
{
    const size_t o = (size_t)layer * n_lines + j;
    const FarAB ab = rec.ab[o];
    const double amp = record_amplitude(ab);
    Far32 f;
    if (amp > 0.)
    {
        const double nu = rec.gen[o].nu;
        const int shift = amp_shift(rec.amp_max[layer]);
        f.cbf = (float)rec.chk[o].cb;
        f.frac = (float)(nu - floor(nu));
        f.g2 = (float)(rec.cc[o] * amp);
        f.amp = (float)ldexp(amp, shift);
        if (!(f.g2 > 0.f)) f.g2 = 1.0e-30f;
    }
    else
    {
        f.cbf = 0.f;
        f.frac = 0.f;
        f.g2 = 1.f;
        f.amp = 0.f;
    }
    rec.f32[o] = f;
}

// ---------------------------------------------------------------------------------------
// K2: gather summation of the far-wing (Lorentz) terms.  One thread owns P consecutive grid
// points that lie in one integer-wavenumber cell (P divides n_per_v); the 32 threads of a
// warp own 32*P consecutive points and walk the same line ranges.
//
// Every active (line, point) pair belongs to exactly one of three classes:
//   near  : point inside the line's near zone [nlo, nhi]          -> K2b (full Humlicek)
//   far   : otherwise, line cell cb in [cell-cut, cell+cut]        -> here
//   node  : otherwise, cb == cell-cut-1 and the point is the cell's first (r == 0),
//           because that line's inclusive end index e lands on it  -> K2b
// (window membership: SURVEY section 8(a) Q3, derived from spectra.c:48-62).
// ---------------------------------------------------------------------------------------
struct SumArgs
{
    LinesView lines;
    Records rec;
    const LayerIn* layers;
    GridSpec grid;
    double* out;  // [layer][n]
    int n_layers;
    int tpw;      // K2: threads of one warp that share a layer (power of two); a warp covers
                  // tpw*P consecutive points of 32/tpw consecutive layers
    int near_masked;  // 1: the summation kernel left near-zone points out (K2); 0: it added the
                      // Lorentz form there too (K2c) and K2b must add (profile - Lorentz)
    int tile0 = 0;    // K2, K2b: first tile / span of this launch (band calls; tiles are always
                      // those of the whole grid, so that a band reproduces its slice bit for bit)
    int layer0 = 0;   // K2c and K2b<32>: first layer of this launch (grid.y counts from here;
                      // n_layers stays the end bound), so that a group of layers can be summed,
                      // corrected and copied out while the next group computes
};

template <int P>
LBL_HD void plain_range(const FarAB* __restrict__ ab, const double* __restrict__ cc, int jb, int je,
                        const double (&v)[P], double (&acc)[P])
{
    // Lines are taken two at a time so that each pair shares one reciprocal (far_terms_pair).
    int j = jb;
    for (; j + 1 < je; j += 2)
    {
        const double2 l1 = LBL_LDG(reinterpret_cast<const double2*>(ab + j));
        const double2 l2 = LBL_LDG(reinterpret_cast<const double2*>(ab + j + 1));
        const double c1 = LBL_LDG(cc + j);
        const double c2 = LBL_LDG(cc + j + 1);
        far_terms_pair<P>(v, l1.x, l1.y, c1, l2.x, l2.y, c2, acc);
    }
    if (j < je)
    {
        const double2 l = LBL_LDG(reinterpret_cast<const double2*>(ab + j));
        const double c = LBL_LDG(cc + j);
        far_terms<P>(v, l.x, l.y, c, acc);
    }
}

// Lines [jb, je) one at a time (far_terms: one reciprocal each).
template <int P>
LBL_HD void single_range(const FarAB* __restrict__ ab, const double* __restrict__ cc, int jb, int je,
                         const double (&v)[P], double (&acc)[P])
{
    for (int j = jb; j < je; ++j)
    {
        const double2 l = LBL_LDG(reinterpret_cast<const double2*>(ab + j));
        far_terms<P>(v, l.x, l.y, LBL_LDG(cc + j), acc);
    }
}

template <int P>
LBL_HD void masked_range(const FarAB* __restrict__ ab, const double* __restrict__ cc,
                         const LineChk* __restrict__ chk, int jb, int je, int i_first, int cell,
                         int cut_off, const double (&v)[P], double (&acc)[P])
{
    const int cmin = cell - cut_off;
    const int cmax = cell + cut_off;
    const int i_last = i_first + P - 1;
    for (int j = jb; j < je; ++j)
    {
        const int4 ck = LBL_LDG(reinterpret_cast<const int4*>(chk + j));  // cb, nlo, nhi
        if (ck.x < cmin || ck.x > cmax)
        {
            continue;
        }
        const double2 l = LBL_LDG(reinterpret_cast<const double2*>(ab + j));
        const double c = LBL_LDG(cc + j);
        if (ck.y > i_last || ck.z < i_first)
        {
            far_terms<P>(v, l.x, l.y, c, acc);
        }
        else
        {
            // Some of this thread's points are in the near zone: they are K2b's.
#pragma unroll
            for (int p = 0; p < P; ++p)
            {
                const int i = i_first + p;
                const double cp = (i >= ck.y && i <= ck.z) ? kBig : c;
                acc[p] = far_term(v[p], l.x, l.y, cp, acc[p]);
            }
        }
    }
}

// Which layer and which points a lane of K2 owns.  On fine grids a warp is 32 threads of one
// layer (tpw = 32).  On coarse grids (few points per cell) that warp would span many cells
// and most of each line window would be edge; there the warp is folded: tpw threads along
// the points, 32/tpw consecutive layers.
struct SumLane
{
    int layer;
    int i_first;
    int group_first, group_last;  // points of this lane's layer that the warp covers
    bool valid;                   // owns real points (stores its result)
    bool any;                     // false: the whole warp is past the end of the grid
};

// Lines [jb, je) at P consecutive points of one cell, Lorentz form everywhere (window test only).
template <int P>
LBL_HD void window_range(const FarAB* __restrict__ ab, const double* __restrict__ cc,
                         const LineChk* __restrict__ chk, int jb, int je, int cell, int cut_off,
                         const double (&v)[P], double (&acc)[P])
{
    const int cmin = cell - cut_off;
    const int cmax = cell + cut_off;
    for (int j = jb; j < je; ++j)
    {
        const int cb = LBL_LDG(reinterpret_cast<const int4*>(chk + j)).x;
        if (cb < cmin || cb > cmax)
        {
            continue;
        }
        const double2 l = LBL_LDG(reinterpret_cast<const double2*>(ab + j));
        far_terms<P>(v, l.x, l.y, LBL_LDG(cc + j), acc);
    }
}

template <int P>
LBL_HD SumLane sum_lane(const SumArgs& a, int layer_group, int tile, int lane)
{
    const GridSpec& g = a.grid;
    SumLane s;
    const int lp = 32 / a.tpw;
    s.layer = layer_group * lp + lane / a.tpw;
    const bool layer_ok = s.layer < a.n_layers;
    if (!layer_ok) s.layer = a.n_layers - 1;
    s.group_first = (tile + a.tile0) * a.tpw * P;
    s.any = s.group_first < g.n;
    s.group_last = s.group_first + a.tpw * P - 1;
    if (s.group_last > g.n - 1) s.group_last = g.n - 1;
    s.i_first = s.group_first + (lane % a.tpw) * P;
    s.valid = layer_ok && s.i_first < g.n;
    if (s.i_first >= g.n) s.i_first = g.n - P;  // idle lanes shadow a real thread, store nothing
    return s;
}

template <int P>
LBL_HD void sum_thread(const SumArgs& a, int layer_group, int tile, int lane)
{
    const GridSpec& g = a.grid;
    const SumLane sl = sum_lane<P>(a, layer_group, tile, lane);
    if (!sl.any)
    {
        return;
    }
    const int layer = sl.layer;
    const int i_first = sl.i_first;
    const bool valid = sl.valid;
    const LayerIn ly = a.layers[layer];
    const Segments seg = find_segments(a.lines.nu, a.lines.n, g.v0, g.n_per_v, g.dv, g.cut_off,
                                       sl.group_first, sl.group_last, ly.slack, ly.kappa);
    const int cell = i_first / g.n_per_v;

    double v[P], acc[P];
#pragma unroll
    for (int p = 0; p < P; ++p)
    {
        v[p] = grid_point(g.v0, g.dv, i_first + p);
        acc[p] = 0.;
    }
    const size_t off = (size_t)layer * a.lines.n;
    const FarAB* ab = a.rec.ab + off;
    const double* cc = a.rec.cc + off;
    const LineChk* chk = a.rec.chk + off;

    masked_range<P>(ab, cc, chk, seg.j[0], seg.j[1], i_first, cell, g.cut_off, v, acc);
    plain_range<P>(ab, cc, seg.j[1], seg.j[2], v, acc);
    masked_range<P>(ab, cc, chk, seg.j[2], seg.j[3], i_first, cell, g.cut_off, v, acc);
    plain_range<P>(ab, cc, seg.j[3], seg.j[4], v, acc);
    masked_range<P>(ab, cc, chk, seg.j[4], seg.j[5], i_first, cell, g.cut_off, v, acc);

    if (valid)
    {
        double* o = a.out + (size_t)layer * g.n + i_first;
#pragma unroll
        for (int p = 0; p < P; ++p)
        {
            o[p] = acc[p];
        }
    }
}

// ---------------------------------------------------------------------------------------
// K2 (FP32 mode): same ranges, same masks, same ownership as sum_thread; the far-wing terms
// are evaluated in FP32 (far32_pair) and flushed into FP64 accumulators every kF32Flush
// lines so that the FP32 running sums stay short.  Near zone, node terms and the pedestal
// remain FP64 (K2b, K3).  Stated tolerance of the mode: 1e-4 (tests/helpers.py: FP32_TOL).
// ---------------------------------------------------------------------------------------
constexpr int kF32Flush = 64;

template <int P>
LBL_HD void flush32(float (&acc32)[P], double (&acc)[P])
{
#pragma unroll
    for (int p = 0; p < P; ++p)
    {
        acc[p] += (double)acc32[p];
        acc32[p] = 0.f;
    }
}

template <int P>
LBL_HD void plain_range32(const Far32* __restrict__ f32, int jb, int je, float cellf,
                          const float (&fp)[P], float (&acc32)[P], double (&acc)[P])
{
    int j = jb;
    while (j < je)
    {
        const int stop = (je - j > kF32Flush) ? j + kF32Flush : je;
        for (; j + 1 < stop; j += 2)
        {
            const float4 l1 = LBL_LDG(reinterpret_cast<const float4*>(f32 + j));
            const float4 l2 = LBL_LDG(reinterpret_cast<const float4*>(f32 + j + 1));
            far32_pair<P>(fp, (cellf - l1.x) - l1.y, l1.z, l1.w, (cellf - l2.x) - l2.y, l2.z, l2.w,
                          acc32);
        }
        if (j < stop)
        {
            const float4 l1 = LBL_LDG(reinterpret_cast<const float4*>(f32 + j));
            far32_one<P>(fp, (cellf - l1.x) - l1.y, l1.z, l1.w, acc32);
            ++j;
        }
        flush32<P>(acc32, acc);
    }
}

template <int P>
LBL_HD void masked_range32(const Far32* __restrict__ f32, const LineChk* __restrict__ chk, int jb,
                           int je, int i_first, int cell, int cut_off, float cellf,
                           const float (&fp)[P], float (&acc32)[P], double (&acc)[P])
{
    const int cmin = cell - cut_off;
    const int cmax = cell + cut_off;
    const int i_last = i_first + P - 1;
    int since_flush = 0;
    for (int j = jb; j < je; ++j)
    {
        const int4 ck = LBL_LDG(reinterpret_cast<const int4*>(chk + j));
        if (ck.x < cmin || ck.x > cmax)
        {
            continue;
        }
        const float4 l = LBL_LDG(reinterpret_cast<const float4*>(f32 + j));
        const float t = (cellf - l.x) - l.y;
        if (ck.y > i_last || ck.z < i_first)
        {
            far32_one<P>(fp, t, l.z, l.w, acc32);
        }
        else
        {
#pragma unroll
            for (int p = 0; p < P; ++p)
            {
                const int i = i_first + p;
                const float ap = (i >= ck.y && i <= ck.z) ? 0.f : l.w;
                const float u = t + fp[p];
                acc32[p] = fmaf_(ap, rcp_f32(fmaf_(u, u, l.z)), acc32[p]);
            }
        }
        if (++since_flush == kF32Flush)
        {
            flush32<P>(acc32, acc);
            since_flush = 0;
        }
    }
    flush32<P>(acc32, acc);
}

template <int P>
LBL_HD void sum32_thread(const SumArgs& a, int layer_group, int tile, int lane)
{
    const GridSpec& g = a.grid;
    const SumLane sl = sum_lane<P>(a, layer_group, tile, lane);
    if (!sl.any)
    {
        return;
    }
    const int layer = sl.layer;
    const int i_first = sl.i_first;
    const bool valid = sl.valid;
    const LayerIn ly = a.layers[layer];
    const Segments seg = find_segments(a.lines.nu, a.lines.n, g.v0, g.n_per_v, g.dv, g.cut_off,
                                       sl.group_first, sl.group_last, ly.slack, ly.kappa);
    const int cell = i_first / g.n_per_v;
    const float cellf = (float)cell;
    const double cell_origin = (double)g.v0 + (double)cell;

    float fp[P], acc32[P];
    double acc[P];
#pragma unroll
    for (int p = 0; p < P; ++p)
    {
        fp[p] = (float)(grid_point(g.v0, g.dv, i_first + p) - cell_origin);
        acc32[p] = 0.f;
        acc[p] = 0.;
    }
    const size_t off = (size_t)layer * a.lines.n;
    const Far32* f32 = a.rec.f32 + off;
    const LineChk* chk = a.rec.chk + off;

    masked_range32<P>(f32, chk, seg.j[0], seg.j[1], i_first, cell, g.cut_off, cellf, fp, acc32, acc);
    plain_range32<P>(f32, seg.j[1], seg.j[2], cellf, fp, acc32, acc);
    masked_range32<P>(f32, chk, seg.j[2], seg.j[3], i_first, cell, g.cut_off, cellf, fp, acc32, acc);
    plain_range32<P>(f32, seg.j[3], seg.j[4], cellf, fp, acc32, acc);
    masked_range32<P>(f32, chk, seg.j[4], seg.j[5], i_first, cell, g.cut_off, cellf, fp, acc32, acc);

    if (valid)
    {
        const double unscale = ldexp(1.0, -amp_shift(a.rec.amp_max[layer]));
        double* o = a.out + (size_t)layer * g.n + i_first;
#pragma unroll
        for (int p = 0; p < P; ++p)
        {
            o[p] = acc[p] * unscale;
        }
    }
}

// ---------------------------------------------------------------------------------------
// K2c: cell-tiled summation with a polynomial far field (fine grids, n_per_v >= 64).
//
// A warp owns a group of G consecutive integer-wavenumber cells of one layer.  All points
// (r > 0) of a cell share the line window [cell-cut, cell+cut], so there are no window
// edges inside a cell.  Lines are split by the distance d of their centre beyond the cells'
// edges.  With h the half-length of the cell interval (~0.5 cm-1), a line at distance d has its
// poles outside the Bernstein ellipse rho = c + sqrt(c^2 - 1), c = (d + h)/h, of that interval,
// and the Chebyshev interpolant of the lines' sum through n nodes is exact to ~rho^-n of a
// line's size on the cell:
//   direct   d < kFarMin = 0.25.  Evaluated at every grid point (kCellP consecutive points per
//            thread).  Unlike K2, this kernel adds the Lorentz form at a line's near-zone
//            points too -- a smooth function that the interpolation handles like any other --
//            and K2b adds (profile - Lorentz) there.
//   mid      kFarMin <= d < kVeryFar = 1.0: kNodes = 32 nodes (rho >= 2.6, rho^-32 = 4e-14).
//            Lane k evaluates the mid lines at node k of each cell.
//   very far kVeryFar <= d < kFar8 = 6.0: kNodes16 = 16 nodes (rho >= 5.9, rho^-16 = 5e-13).  A
//            half-warp takes a cell's 16 nodes (G = 2), or the two half-warps split the lines
//            of the one cell (G = 1): one evaluation per lane per line.
//   far-8    d >= kFar8, three quarters of the window: kNodes8 = 8 nodes (rho >= 26,
//            rho^-8 = 5e-12), the lanes of a cell splitting the lines 2 or 4 ways.
// The boundaries are set so that the far field as a whole stays within 1e-11 of the direct sum
// (tests/test_emulated_kernels.py::test_far_field_kernel_against_direct_kernel), a hundredth
// of the parity budget.  The three node sums are turned into Chebyshev coefficients (32x32,
// 16x16 and 8x8 transforms, lbl_cheb.h), the coefficients are added -- the series live on the
// same interval -- and every point of the cell receives the sum by Clenshaw's recurrence.
// ---------------------------------------------------------------------------------------
constexpr int kNodes = 32;
constexpr int kNodes16 = 16;
constexpr int kNodes8 = 8;
constexpr int kCellP = 4;         // points per thread in the direct part
constexpr double kFarMin = 0.25;  // cm-1 beyond the cell edges where the 32-node field starts
constexpr double kVeryFar = 1.0;  // cm-1 beyond the cell edges where the 16-node field starts
constexpr double kFar8 = 6.0;     // cm-1 beyond the cell edges where the 8-node field starts

struct CellArgs
{
    SumArgs sum;
    const double* node_offset;     // [kNodes] node position relative to the cell origin v0+cell
    const double* transform;       // [kNodes][kNodes] node sums -> Chebyshev coefficients (k-major)
    const double* node_offset16;   // [kNodes16]
    const double* transform16;     // [kNodes16][kNodes16]
    const double* node_offset8;    // [kNodes8]
    const double* transform8;      // [kNodes8][kNodes8]
    unsigned long long* executed;  // statistics: evaluations actually performed (or nullptr)
    // The range searches of every (layer, cell group), done ahead by cell_keys_kernel:
    // keys[(layer_in_launch_chunk * key_groups + group) * kKeyStride + which] (nullptr: search here)
    const int* keys = nullptr;
    int key_groups = 0;
    int key_layer0 = 0;   // chunk-local layer index of the launch's first layer
};
constexpr int kKeyStride = 16;   // ints per (layer, cell group): ten keys, padded to 64 bytes

// Line ranges of a cell group, as indices into the nu-sorted line list, from below:
//   [j0,j1) window edge, tested per cell (16 nodes) | [j1,j2) 8-node | [j2,j3) 16-node |
//   [j3,j4) 32-node | [j4,j5) direct | [j5,j6) 32-node | [j6,j7) 16-node | [j7,j8) 8-node |
//   [j8,j9) window edge, tested per cell (16 nodes)
struct CellSegments
{
    int j[10];
};
constexpr int kCellKeys = 10;

// The search keys of those ranges for the group of `cells` consecutive cells starting at
// `cell` (one search each; the kernel gives one key to each of ten lanes).
LBL_HD double cell_search_key(const GridSpec& g, const LayerIn& ly, int cell, int cells, int which)
{
    const double lo = (double)g.v0 + (double)cell;                       // first point of the group
    const double lo_last = lo + (double)(cells - 1);                     // first point of its last cell
    const double hi = lo_last + (double)(g.n_per_v - 1) * g.dv;          // last point of the group
    const double reach = kFarMin + ly.slack;
    switch (which)
    {
        case 0: return lo - (double)g.cut_off - ly.slack;            // first line in any cell's window
        case 1: return lo_last - (double)g.cut_off + ly.slack;       // first line certainly in all of them
        case 2: return lo - kFar8 - ly.slack;                        // end of the lower 8-node range
        case 3: return lo - kVeryFar - ly.slack;                     // 32-node range
        case 4: return lo - reach;                                   // direct range
        case 5: return hi + reach;
        case 6: return hi + kVeryFar + ly.slack;
        case 7: return hi + kFar8 + ly.slack;                        // start of the upper 8-node range
        case 8: return lo + (double)(g.cut_off + 1) - ly.slack;      // end of the certain part
        default: return lo_last + (double)(g.cut_off + 1) + ly.slack; // end of the last window
    }
}

LBL_HD int clamp_int(int x, int lo, int hi)
{
    return x < lo ? lo : (x > hi ? hi : x);
}

LBL_HD CellSegments cell_segments_from(const int (&found)[kCellKeys])
{
    CellSegments s;
    const int w_lo = found[0];
    const int w_hi = found[9] > w_lo ? found[9] : w_lo;
    // direct range, clamped to the window (it may reach beyond it when cut_off is tiny)
    const int d_lo = clamp_int(found[4], w_lo, w_hi);
    const int d_hi = clamp_int(found[5], d_lo, w_hi);
    // 32-node range around it
    const int m_lo = clamp_int(found[3], w_lo, d_lo);
    const int m_hi = clamp_int(found[6], d_hi, w_hi);
    // lines certainly inside every cell's window
    const int c_lo = clamp_int(found[1], w_lo, m_lo);
    const int c_hi = clamp_int(found[8], m_hi, w_hi);
    // 8-node ranges: the outer part of the certain lines
    const int e_lo = clamp_int(found[2], c_lo, m_lo);
    const int e_hi = clamp_int(found[7], m_hi, c_hi);
    s.j[0] = w_lo;
    s.j[1] = c_lo;
    s.j[2] = e_lo;
    s.j[3] = m_lo;
    s.j[4] = d_lo;
    s.j[5] = d_hi;
    s.j[6] = m_hi;
    s.j[7] = e_hi;
    s.j[8] = c_hi;
    s.j[9] = w_hi;
    return s;
}

LBL_HD CellSegments cell_segments(const LinesView& lines, const GridSpec& g, const LayerIn& ly,
                                  int cell, int cells)
{
    int found[kCellKeys];
    for (int which = 0; which < kCellKeys; ++which)
    {
        found[which] = first_line_at(lines, cell_search_key(g, ly, cell, cells, which));
    }
    return cell_segments_from(found);
}

// Mid lines of [jb, je) at this lane's node in each of the G cells.  Pairs of lines share a
// reciprocal; U*G independent chains per lane keep the FP64 pipe busy.
template <int G>
LBL_HD void node_plain(const FarAB* __restrict__ ab, const double* __restrict__ cc, int jb, int je,
                       const double (&v)[G], double (&sum)[G])
{
    constexpr int U = (G >= 4) ? 1 : 4 / G;
    double acc[U][G];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int c = 0; c < G; ++c) acc[u][c] = 0.;
    int j = jb;
    for (; j + 2 * U - 1 < je; j += 2 * U)
    {
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            const double2 l1 = LBL_LDG(reinterpret_cast<const double2*>(ab + j + 2 * u));
            const double2 l2 = LBL_LDG(reinterpret_cast<const double2*>(ab + j + 2 * u + 1));
            const double c1 = LBL_LDG(cc + j + 2 * u);
            const double c2 = LBL_LDG(cc + j + 2 * u + 1);
            far_terms_pair<G>(v, l1.x, l1.y, c1, l2.x, l2.y, c2, acc[u]);
        }
    }
    for (; j < je; ++j)
    {
        const double2 l = LBL_LDG(reinterpret_cast<const double2*>(ab + j));
        far_terms<G>(v, l.x, l.y, LBL_LDG(cc + j), acc[0]);
    }
#pragma unroll
    for (int c = 0; c < G; ++c)
    {
        double t = 0.;
#pragma unroll
        for (int u = 0; u < U; ++u) t += acc[u][c];
        sum[c] += t;
    }
}

// Very far lines of [jb, je) at ONE point per lane (a 16-node point of the lane's cell).
// `first`/`stride`: which lines of the range this lane takes (all of them when a half-warp
// owns a cell; every other pair when the two half-warps share one cell).
LBL_HD double node16_plain(const FarAB* __restrict__ ab, const double* __restrict__ cc, int jb,
                           int je, int first, int stride, double v)
{
    double acc[4] = {0., 0., 0., 0.};
    const double vv[1] = {v};
    // pairs (jb + 2p, jb + 2p + 1), p = first, first + stride, ...
    const int pairs = (je - jb) >> 1;
    int p = first;
    for (; p + 3 * stride < pairs; p += 4 * stride)
    {
#pragma unroll
        for (int u = 0; u < 4; ++u)
        {
            const int j = jb + 2 * (p + u * stride);
            const double2 l1 = LBL_LDG(reinterpret_cast<const double2*>(ab + j));
            const double2 l2 = LBL_LDG(reinterpret_cast<const double2*>(ab + j + 1));
            double one[1] = {acc[u]};
            far_terms_pair<1>(vv, l1.x, l1.y, LBL_LDG(cc + j), l2.x, l2.y, LBL_LDG(cc + j + 1), one);
            acc[u] = one[0];
        }
    }
    for (; p < pairs; p += stride)
    {
        const int j = jb + 2 * p;
        const double2 l1 = LBL_LDG(reinterpret_cast<const double2*>(ab + j));
        const double2 l2 = LBL_LDG(reinterpret_cast<const double2*>(ab + j + 1));
        double one[1] = {acc[0]};
        far_terms_pair<1>(vv, l1.x, l1.y, LBL_LDG(cc + j), l2.x, l2.y, LBL_LDG(cc + j + 1), one);
        acc[0] = one[0];
    }
    if (((je - jb) & 1) && first == 0)
    {
        const int j = je - 1;   // odd line out
        const double2 l = LBL_LDG(reinterpret_cast<const double2*>(ab + j));
        acc[1] = far_term(v, l.x, l.y, LBL_LDG(cc + j), acc[1]);
    }
    return (acc[0] + acc[1]) + (acc[2] + acc[3]);
}

// Very far lines next to a window edge: the lane's cell's own window test decides.
LBL_HD double node16_tested(const FarAB* __restrict__ ab, const double* __restrict__ cc,
                            const LineChk* __restrict__ chk, int jb, int je, int first, int stride,
                            int cell, int cut_off, double v)
{
    double acc = 0.;
    for (int j = jb + first; j < je; j += stride)
    {
        const int cb = LBL_LDG(reinterpret_cast<const int4*>(chk + j)).x;
        if (cb < cell - cut_off || cb > cell + cut_off)
        {
            continue;
        }
        const double2 l = LBL_LDG(reinterpret_cast<const double2*>(ab + j));
        acc = far_term(v, l.x, l.y, LBL_LDG(cc + j), acc);
    }
    return acc;
}

// How the 32 lanes map onto 16-node points: (cell offset, node, first line share, stride).
template <int G>
struct Lane16
{
    int cell_off, node, first, stride;
};
template <int G>
LBL_HD Lane16<G> lane16(int lane)
{
    Lane16<G> m;
    m.node = lane & 15;
    if (G >= 2)
    {
        m.cell_off = lane >> 4;   // a half-warp per cell (G == 2)
        m.first = 0;
        m.stride = 1;
    }
    else
    {
        m.cell_off = 0;           // the half-warps split the lines of the one cell
        m.first = lane >> 4;
        m.stride = 2;
    }
    return m;
}

// How the 32 lanes map onto 8-node points: each cell's lanes split its lines 2 ways (G == 2,
// a half-warp per cell) or 4 ways (G == 1); the shares are added up by shuffle.
template <int G>
LBL_HD Lane16<G> lane8(int lane)
{
    Lane16<G> m;
    m.node = lane & 7;
    if (G >= 2)
    {
        m.cell_off = lane >> 4;
        m.first = (lane >> 3) & 1;
        m.stride = 2;
    }
    else
    {
        m.cell_off = 0;
        m.first = lane >> 3;
        m.stride = 4;
    }
    return m;
}

// Phase 1, lane = node: sums of the mid lines at this lane's 32-node point of each cell (f32)
// and of the very far lines at this lane's 16-node point (f16; for G == 1 the two half-warps
// hold partial sums of the same 16 points).
template <int G>
LBL_HD void cell_far_lane(const CellArgs& a, int layer, int cell, int lane, const CellSegments& seg,
                          double (&f32)[G], double& f16, double& f8)
{
    static_assert(G == 1 || G == 2, "a warp holds the 16-node points of one or two cells");
    const GridSpec& g = a.sum.grid;
    const size_t off = (size_t)layer * a.sum.lines.n;
    const FarAB* ab = a.sum.rec.ab + off;
    const double* cc = a.sum.rec.cc + off;
    const LineChk* chk = a.sum.rec.chk + off;
    double v[G];
#pragma unroll
    for (int q = 0; q < G; ++q)
    {
        v[q] = ((double)g.v0 + (double)(cell + q)) + a.node_offset[lane];
        f32[q] = 0.;
    }
    node_plain<G>(ab, cc, seg.j[3], seg.j[4], v, f32);
    node_plain<G>(ab, cc, seg.j[5], seg.j[6], v, f32);
    const Lane16<G> m = lane16<G>(lane);
    const int my_cell = cell + m.cell_off;
    const double v16 = ((double)g.v0 + (double)my_cell) + a.node_offset16[m.node];
    f16 = node16_tested(ab, cc, chk, seg.j[0], seg.j[1], m.first, m.stride, my_cell, g.cut_off, v16);
    f16 += node16_plain(ab, cc, seg.j[2], seg.j[3], m.first, m.stride, v16);
    f16 += node16_plain(ab, cc, seg.j[6], seg.j[7], m.first, m.stride, v16);
    f16 += node16_tested(ab, cc, chk, seg.j[8], seg.j[9], m.first, m.stride, my_cell, g.cut_off, v16);
    const Lane16<G> m8 = lane8<G>(lane);
    const double v8 = ((double)g.v0 + (double)(cell + m8.cell_off)) + a.node_offset8[m8.node];
    f8 = node16_plain(ab, cc, seg.j[1], seg.j[2], m8.first, m8.stride, v8);
    f8 += node16_plain(ab, cc, seg.j[7], seg.j[8], m8.first, m8.stride, v8);
}

// Phase 2, lane = kCellP consecutive points of chunk `chunk` of the cell: the direct lines.
LBL_HD void cell_direct_lane(const CellArgs& a, int layer, int cell, int chunk, int lane,
                             const CellSegments& seg)
{
    const GridSpec& g = a.sum.grid;
    constexpr int P = kCellP;
    int r_first = (chunk * 32 + lane) * P;          // offset inside the cell
    const bool valid = r_first < g.n_per_v;
    if (!valid) r_first = 0;
    int count = g.n_per_v - r_first;                // points this thread really owns
    if (count > P) count = P;
    const int i_first = cell * g.n_per_v + r_first;
    double v[P], acc[P];
#pragma unroll
    for (int p = 0; p < P; ++p)
    {
        // a thread at the end of the cell shadows the cell's last point with its spare slots
        const int r = (p < count) ? r_first + p : g.n_per_v - 1;
        v[p] = grid_point(g.v0, g.dv, cell * g.n_per_v + r);
        acc[p] = 0.;
    }
    const size_t off = (size_t)layer * a.sum.lines.n;
    LBL_CHECK(!valid || (i_first >= 0 && i_first + count <= g.n));
    LBL_CHECK(seg.j[4] >= 0 && seg.j[4] <= seg.j[5] && seg.j[5] <= a.sum.lines.n);
    if (g.cut_off >= 4)
    {
        // Direct lines lie within kFarMin of the (at most 2-cell) group: their window cell is
        // within 2 of this cell, inside any window with cut_off >= 4 -- no per-line test.
        // One reciprocal per line here, not one per pair: K2b takes these terms back at the
        // points next to the line centre and must be able to form the very same bits
        // (far_term_lo); a pair's combined reciprocal cannot be reproduced line by line.
        single_range<P>(a.sum.rec.ab + off, a.sum.rec.cc + off, seg.j[4], seg.j[5], v, acc);
    }
    else
    {
        window_range<P>(a.sum.rec.ab + off, a.sum.rec.cc + off, a.sum.rec.chk + off, seg.j[4],
                        seg.j[5], cell, g.cut_off, v, acc);
    }
    if (valid)
    {
        double* o = a.sum.out + (size_t)layer * g.n + i_first;
#pragma unroll
        for (int p = 0; p < P; ++p)
        {
            if (p < count) o[p] = acc[p];
        }
    }
}

// Phase 3a, lane = coefficient index: node sums -> Chebyshev coefficients of the interpolant
// (c_j = sum_k M[k][j] F_k, lbl_cheb.h).  `nodes` is 32 or 16; lanes >= nodes return 0.
LBL_HD double cell_coefficient(const double* transform, const double* field, int nodes, int lane)
{
    double c = 0.;
    if (lane < nodes)
    {
        for (int k = 0; k < nodes; ++k)
        {
            c = fma_(LBL_LDG(transform + (size_t)k * nodes + lane), field[k], c);
        }
    }
    return c;
}

// Phase 3b, lane = points lane, lane+32, ... of the cell: add the interpolated far field,
// evaluated from its Chebyshev coefficients by Clenshaw's recurrence
//   b_j = c_j + 2 s b_(j+1) - b_(j+2),   p(s) = c_0 + s b_1 - b_2,
// with s in [-1, 1] the point's position on the cell interval.  `coef` holds the SUM of the
// three fields' coefficients (they are Chebyshev series on the same interval: kNodes terms,
// the 16- and 8-node fields contributing to the first 16 and 8).  (A stored interpolation
// matrix would cost one cache read per point and node -- on this grid more L1 traffic than
// the rest of the kernel together; the recurrence costs two FP64 operations per point and
// coefficient and reads only the coefficients, by broadcast.)
LBL_HD void cell_field_lane(const CellArgs& a, int layer, int cell, int lane, int nlanes,
                            const double* coef)
{
    const GridSpec& g = a.sum.grid;
    double* o = a.sum.out + (size_t)layer * g.n + (size_t)cell * g.n_per_v;
    constexpr int R = 4;   // points per lane in flight: independent recurrences hide the latency
    const double to_s = 2.0 / (double)(g.n_per_v - 1);
    for (int r0 = lane; r0 < g.n_per_v; r0 += R * nlanes)
    {
        double s2[R], b1[R], b2[R];
#pragma unroll
        for (int u = 0; u < R; ++u)
        {
            int r = r0 + u * nlanes;
            if (r >= g.n_per_v) r = g.n_per_v - 1;   // spare slots shadow the last point
            s2[u] = 2.0 * fma_((double)r, to_s, -1.0);
            b1[u] = b2[u] = 0.;
        }
#pragma unroll 2
        for (int j = kNodes - 1; j >= 1; --j)
        {
            const double cj = coef[j];
#pragma unroll
            for (int u = 0; u < R; ++u)
            {
                const double t = fma_(s2[u], b1[u], cj - b2[u]);
                b2[u] = b1[u];
                b1[u] = t;
            }
        }
        const double c0 = coef[0];
#pragma unroll
        for (int u = 0; u < R; ++u)
        {
            const int r = r0 + u * nlanes;
            if (r < g.n_per_v)
            {
                o[r] += fma_(0.5 * s2[u], b1[u], c0) - b2[u];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// K2b: near-zone and node terms, one grid point per lane (dense in the near zone, where K2's
// P-points-per-thread layout would leave most lanes idle).  A warp covers T consecutive
// points of 32/T consecutive layers; T shrinks with the grid resolution so that a tile stays
// about as wide as a near zone.  Adds into the spectrum K2 wrote (same stream, no atomics:
// each point still has exactly one owner).
// ---------------------------------------------------------------------------------------
// Candidate range of K2b's near loop for grid points [t_first, t_last] (same reach as
// find_segments uses for its checked middle segment).
LBL_HD void near_candidates(const LinesView& lines, const GridSpec& g, const LayerIn& ly,
                            int t_first, int t_last, int& jlo, int& jhi)
{
    const double base = (double)g.v0;
    const double v_first = base + (double)t_first * g.dv;
    const double v_last = base + (double)t_last * g.dv;
    const double reach = (ly.kappa < 0.5)
        ? (ly.kappa * fabs(v_last) / (1.0 - ly.kappa)) * (1.0 + 0x1p-20) + ly.slack + 3.0 * g.dv
        : 1.0e300;
    jlo = first_line_at(lines, v_first - reach);
    jhi = first_line_at(lines, v_last + reach);
}

// What K2b needs of one (layer, line) to evaluate its near zone, derived once per tile.
// mode 0: the summation kernel added the Lorentz form ax/(x^2+y^2) at these points (K2c adds
// it everywhere), so K2b adds profile - Lorentz; mode 1: the summation kernel added nothing
// (K2 masks the near zone; lines too weak for the folded operands are dropped): K2b adds the
// profile.
struct NearLine
{
    int nlo, nhi, cb, tag;        // tag = (sorted line index << 1) | mode
    double nu, repwid, lim_outer, xlim0;
    double ax, d0, d2, n0;
    double yq, lim_r2, y, cof;
    double a, b, c;               // the summation kernel's operands (far_term)
    int c_lo, c_hi;               // near_block_kernel: grid indices that hold the points with
                                  // |x| < lim_outer (regions 2, 3, CPF12), clipped to the span
                                  // (144 B: a multiple of 16, and as a shared-memory stride free of
                                  // bank conflicts for the 16-byte stores of 8 lanes; 128 B is not)
};

LBL_HD NearLine near_line(const int4& ck, int j, const LineGen& gen, const FarAB& ab, double cc,
                          bool near_masked)
{
    const bool lorentz_added = !near_masked && cc != kBig;
    NearLine nl;
    nl.nlo = ck.y;
    nl.nhi = ck.z;
    nl.cb = ck.x;
    nl.tag = (j << 1) | (lorentz_added ? 0 : 1);
    nl.nu = gen.nu;
    nl.repwid = gen.repwid;
    nl.lim_outer = voigt_outer_limit(gen.y, gen.xlim0, gen.xlim1);
    nl.xlim0 = gen.xlim0;
    nl.yq = gen.y * gen.y;
    const double a0 = nl.yq + 0.5;
    nl.ax = gen.cof * (gen.y * kRsqrPi);
    nl.d0 = a0 * a0;
    nl.d2 = nl.yq + nl.yq - 1.;
    nl.n0 = -0.5 * nl.yq - 0.25;
    nl.lim_r2 = voigt_region2_limit(gen.y, gen.xlim0);
    nl.y = gen.y;
    nl.cof = gen.cof;
    nl.a = ab.a;
    nl.b = ab.b;
    nl.c = cc;
    nl.c_lo = 0;
    nl.c_hi = -1;
    return nl;
}

// Value K2b adds at wavenumber v for this line, without the region-3/CPF12 part: when `core`
// comes back true the caller still owes cof*voigt_inner(x, y) (queued and lane-packed on the
// device).  In mode 0 the W4 region-1 rational minus the Lorentz form is taken in closed form,
//   y/sqrt(pi) [ (a0+x^2)/(d0 + x^2 (d2+x^2)) - 1/(x^2+y^2) ]
//     = y/sqrt(pi) (1.5 x^2 - 0.5 y^2 - 0.25) / ((d0 + x^2 (d2+x^2)) (x^2+y^2)),
// (a0 = y^2+0.5, d0 = a0^2, d2 = 2y^2-1: voigt.c:84-97), one reciprocal and no cancellation;
// in region 0 (voigt.c:79-83) profile and Lorentz form are the same number.  Closer to the
// centre the Lorentz form is taken back exactly as the summation kernel formed it
// (d = v*a + b cancels there, and the same rounding must cancel on both sides).
LBL_HD double near_point(const NearLine& nl, double v, bool& core)
{
    const double abx = fabs((v - nl.nu) * nl.repwid);
    const double xq = abx * abx;
    const bool lorentz_added = (nl.tag & 1) == 0;
    core = false;
    if (abx >= nl.lim_outer)
    {
        if (lorentz_added)
        {
            if (abx >= nl.xlim0)
            {
                return 0.;
            }
            const double den = fma_(xq, nl.d2 + xq, nl.d0) * (xq + nl.yq);
            return (nl.ax * fma_(1.5, xq, nl.n0)) * rcp_newton2(den);
        }
        return nl.cof * voigt_outer(abx, xq, nl.y, nl.xlim0);
    }
    const double back = lorentz_added ? -far_term_lo(v, nl.a, nl.b, nl.c, 0.) : 0.;
    if (abx >= nl.lim_r2)
    {
        return back + nl.cof * voigt_region2(xq, nl.y);   // short rational: on the spot
    }
    core = true;
    return back;
}

template <int T>
LBL_HD void fixup_thread(const SumArgs& a, int tile, int layer_group, int lane)
{
    const GridSpec& g = a.grid;
    constexpr int LP = 32 / T;
    int layer = layer_group * LP + lane / T;
    tile += a.tile0;
    int i = tile * T + lane % T;
    const bool valid = (layer < a.n_layers) && (i < g.n) && i >= band_first_point(g) &&
                       i < band_end_point(g);
    if (layer >= a.n_layers) layer = a.n_layers - 1;
    if (i >= g.n) i = g.n - 1;
    const int t_first = tile * T;
    int t_last = t_first + T - 1;
    if (t_last > g.n - 1) t_last = g.n - 1;

    const LayerIn ly = a.layers[layer];
    const size_t off = (size_t)layer * a.lines.n;
    const LineChk* chk = a.rec.chk + off;
    const LineGen* gen = a.rec.gen + off;
    const double v = grid_point(g.v0, g.dv, i);
    const int cell = i / g.n_per_v;
    const bool is_node = (i - cell * g.n_per_v) == 0;
    double acc = 0.;

    // (1) near zone: every line whose [nlo, nhi] contains i.
    {
        int jlo, jhi;
        near_candidates(a.lines, g, ly, t_first, t_last, jlo, jhi);
        for (int j = jlo; j < jhi; ++j)
        {
            const int4 ck = LBL_LDG(reinterpret_cast<const int4*>(chk + j));
            if (i < ck.y || i > ck.z)
            {
                continue;
            }
            // inside the line's window? s <= i <= e, spectra.c:48-62 (unclamped form)
            const long long s = (long long)(ck.x - g.cut_off) * g.n_per_v;
            const long long e = (long long)(ck.x + g.cut_off + 1) * g.n_per_v;
            if ((long long)i < s || (long long)i > e)
            {
                continue;
            }
            const NearLine nl = near_line(ck, j, gen[j], a.rec.ab[off + j], a.rec.cc[off + j],
                                          a.near_masked != 0);
            bool core;
            acc += near_point(nl, v, core);
            if (core)
            {
                acc += nl.cof * voigt_inner((v - nl.nu) * nl.repwid, nl.y);
            }
        }
    }
    // (2) node terms: lines with cb == cell-cut-1 reach exactly the cell's first point.
    if (is_node)
    {
        const double key = (double)g.v0 + (double)(cell - g.cut_off - 1);
        const int jlo = first_line_at(a.lines, key - ly.slack);
        const int jhi = first_line_at(a.lines, key + 1.0 + ly.slack);
        const FarAB* ab = a.rec.ab + off;
        const double* cc = a.rec.cc + off;
        for (int j = jlo; j < jhi; ++j)
        {
            const int4 ck = LBL_LDG(reinterpret_cast<const int4*>(chk + j));
            if (ck.x != cell - g.cut_off - 1 || (i >= ck.y && i <= ck.z))
            {
                continue;  // other cell, or already taken by the near loop above
            }
            const double2 l = LBL_LDG(reinterpret_cast<const double2*>(ab + j));
            acc = far_term(v, l.x, l.y, LBL_LDG(cc + j), acc);
        }
    }
    if (valid)
    {
        a.out[(size_t)layer * g.n + i] += acc;
    }
}

// ---------------------------------------------------------------------------------------
// K3: pedestal recurrence (spectra.c:66-78).
//
// The reference subtracts, after each line, min(k[s], k[e]) of the ACCUMULATED spectrum
// from k[s..e], so the pedestal of line l depends on all earlier lines.
// s and e are multiples of n_per_v (or the clamps 0 and n-1), hence only the running
// spectrum at integer-wavenumber nodes -- plus the last grid point n-1 -- is needed.
// Lines are walked in DATABASE order.  A line with window cell cb touches "slots"
//   t = 0 .. 2*cut+1  -> node cb-cut+t      (valid if inside [max(cb-cut,0), min(cb+cut+1, ncell-1)])
//   t = 2*cut+2       -> grid point n-1     (valid if e is clamped to n-1 and n_per_v > 1)
// and each line's pedestal is accumulated into pedbin[cb + cut + 1].
// ---------------------------------------------------------------------------------------
struct PedArgs
{
    LinesView lines;
    Records rec;
    GridSpec grid;
    double* pedbin;  // [layer][ncell + 2*cut + 2]
    int n_rows = 0;  // database rows the recurrence walks (<= lines.n; fewer for a band call on
                     // a nu-sorted database: rows past the band's last window cannot reach it)
};

struct PedWindow
{
    int base;     // node of slot 0 (= cb - cut, may be negative)
    int s_node;   // first node inside the window
    int e_node;   // last node inside the window
    int s_slot;   // slot of k[s]
    int e_slot;   // slot of k[e]
    bool tail;    // k[e] is grid point n-1, which is not a node
    bool skip;    // the reference does not process this line on this grid
};

LBL_HD PedWindow ped_window(int cb, const GridSpec& g)
{
    PedWindow w;
    const int e_raw = cb + g.cut_off + 1;
    // s >= n: spectra.c:49-53.  e < 0: undefined behaviour in the reference.
    w.skip = (cb - g.cut_off >= g.ncell) || (e_raw < 0);
    w.base = cb - g.cut_off;
    w.s_node = w.base > 0 ? w.base : 0;
    const bool clamped = e_raw >= g.ncell;
    w.tail = clamped && g.n_per_v > 1;
    w.e_node = clamped ? g.ncell - 1 : e_raw;
    w.s_slot = w.s_node - w.base;
    w.e_slot = w.tail ? 2 * g.cut_off + 2 : w.e_node - w.base;
    return w;
}

// Storage index (in the ncell+1 node array) of slot t, or -1 if the slot is outside the window.
LBL_HD int ped_slot_index(const PedWindow& w, const GridSpec& g, int t)
{
    if (t == 2 * g.cut_off + 2)
    {
        return w.tail ? g.ncell : -1;
    }
    if (t > 2 * g.cut_off + 2)
    {
        return -1;
    }
    const int node = w.base + t;
    return (node >= w.s_node && node <= w.e_node) ? node : -1;
}

// K3a: contribution of database row r to slot t (0 outside the window).
LBL_HD double pedestal_term(const PedArgs& a, int layer, int r, int t)
{
    const GridSpec& g = a.grid;
    const int j = a.lines.db_to_sorted ? a.lines.db_to_sorted[r] : r;
    const size_t o = (size_t)layer * a.lines.n + j;
    const LineChk chk = a.rec.chk[o];
    const PedWindow w = ped_window(chk.cb, g);
    if (w.skip)
    {
        return 0.;
    }
    const int idx = ped_slot_index(w, g, t);
    if (idx < 0)
    {
        return 0.;
    }
    const int i = (idx == g.ncell) ? g.n - 1 : idx * g.n_per_v;
    return line_point(grid_point(g.v0, g.dv, i), i, a.rec.ab[o], a.rec.cc[o], chk, a.rec.gen[o]);
}

// K3a, warp form.  A warp takes one TILE of 32 consecutive database rows; its 32 lanes take
// the slots lane, lane+32, ... of every row, so the per-line work (record loads, window) is
// warp-uniform and the stores are coalesced.  What is stored for row l is not the bare term
// but the RUN-LOCAL PREFIX: the sum of the terms of rows l', l'+1, ..., l that precede it in
// the same run (maximal stretch of equal window cell cb inside the tile).  K3b then needs
// one row per run -- the last -- for its node update.  The bare terms at the two nodes that
// decide the pedestal, f[s] and f[e], go to the two spare slots 2*cut+3 and 2*cut+4.
constexpr int kPedTileRows = 32;

template <int K>
LBL_HD void pedestal_terms_tile(const PedArgs& a, int layer, int tile, int lane, double* rows)
{
    const GridSpec& g = a.grid;
    constexpr int wpad = 32 * K;
    const int spare = 2 * g.cut_off + 3;
    const int first = tile * kPedTileRows;
    const int cnt = (a.n_rows - first < kPedTileRows) ? a.n_rows - first : kPedTileRows;
    double run_sum[K];
    int prev_cb = 0;
#if defined(__CUDA_ARCH__)
    // Lane m fetches the records of the tile's row m once (independent, coalesced loads); the
    // row loop below gets them by shuffle instead of waiting on memory row after row.
    int my_j = 0;
    int4 my_ck = make_int4(0, 0, 0, 0);
    double2 my_l = make_double2(0., 0.);
    double my_c = 0.;
    if (lane < cnt)
    {
        my_j = a.lines.db_to_sorted ? LBL_LDG(a.lines.db_to_sorted + first + lane) : first + lane;
        const size_t mo = (size_t)layer * a.lines.n + my_j;
        my_ck = LBL_LDG(reinterpret_cast<const int4*>(a.rec.chk + mo));
        my_l = LBL_LDG(reinterpret_cast<const double2*>(a.rec.ab + mo));
        my_c = LBL_LDG(a.rec.cc + mo);
    }
#endif
    for (int m = 0; m < cnt; ++m)
    {
#if defined(__CUDA_ARCH__)
        const int j = __shfl_sync(0xffffffffu, my_j, m);
        const size_t o = (size_t)layer * a.lines.n + j;
        int4 ck;
        ck.x = __shfl_sync(0xffffffffu, my_ck.x, m);
        ck.y = __shfl_sync(0xffffffffu, my_ck.y, m);
        ck.z = __shfl_sync(0xffffffffu, my_ck.z, m);
        ck.w = 0;
        double2 l;
        l.x = __shfl_sync(0xffffffffu, my_l.x, m);
        l.y = __shfl_sync(0xffffffffu, my_l.y, m);
        const double c = __shfl_sync(0xffffffffu, my_c, m);
#else
        const int r = first + m;
        const int j = a.lines.db_to_sorted ? LBL_LDG(a.lines.db_to_sorted + r) : r;
        const size_t o = (size_t)layer * a.lines.n + j;
        const int4 ck = LBL_LDG(reinterpret_cast<const int4*>(a.rec.chk + o));
        const double2 l = LBL_LDG(reinterpret_cast<const double2*>(a.rec.ab + o));
        const double c = LBL_LDG(a.rec.cc + o);
#endif
        if (m == 0 || ck.x != prev_cb)
        {
#pragma unroll
            for (int k = 0; k < K; ++k) run_sum[k] = 0.;
        }
        prev_cb = ck.x;
        double* row = rows + (size_t)m * wpad;
        const PedWindow w = ped_window(ck.x, g);
        if (!w.skip)
        {
#pragma unroll
            for (int k = 0; k < K; ++k)
            {
                const int t = lane + 32 * k;
                const int idx = ped_slot_index(w, g, t);
                if (idx >= 0)
                {
                    const int i = (idx == g.ncell) ? g.n - 1 : idx * g.n_per_v;
                    const double v = grid_point(g.v0, g.dv, i);
                    double val;
                    if (i >= ck.y && i <= ck.z)
                    {
                        const LineGen gen = a.rec.gen[o];
                        val = voigt_general(v, gen.nu, gen.repwid, gen.y, gen.cof, gen.xlim0, gen.xlim1);
                    }
                    else
                    {
                        val = far_term(v, l.x, l.y, c, 0.);
                    }
                    run_sum[k] += val;
                    if (t == w.s_slot) row[spare] = val;
                    if (t == w.e_slot) row[spare + 1] = val;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k)
        {
            const int t = lane + 32 * k;
            if (t != spare && t != spare + 1) row[t] = run_sum[k];
        }
    }
}

// K3b: the sequential chain, organised in RUNS of consecutive lines that share one window
// cell cb (in a nu-sorted database a run is all lines of one integer-wavenumber bin).
// Inside a run the window does not move, and after the first line one of the two deciding
// nodes k[s], k[e] is exactly zero (the pedestal just removed it), so with d = k[s] - k[e]
//     k[s] = max(d, 0),  k[e] = max(-d, 0),  d_l = d_(l-1) + (f_l[s] - f_l[e]):
// the recurrence over the run is a prefix sum.  Lanes take one line each for the pedestals
// (warp scan), then one slot each for the node updates
//     node += sum_l f_l[slot] - sum_l pedestal_l.
// Sums are re-associated relative to the reference's line-by-line order (rounding-level
// differences, ~1e-16 of the node values).
//
// Node storage.  Tracked points have storage indices 0..ncell-1 (nodes) and ncell (the last
// grid point).  Index c always lives in the same register of the same lane:
// lane = c mod 32, register k = (c div 32) mod K -- a ring of 32*K indices that follows the
// window.  When the window moves a lane swaps, per register, the index that left the ring
// for the one that entered, through a backing array; because an index is only ever touched
// by its own (lane, register), no synchronisation is needed.
template <int K>
struct PedLane
{
    int cb;
    bool have;
    PedWindow w;
    int idx[K];      // storage index currently held in own[k] (-1: none)
    int slot[K];     // K3a row slot of idx[k] in the current window (-1: outside the window)
    double own[K];   // running spectrum at idx[k]
    double ks, ke;   // k[s], k[e] of the current window (same value in every lane)
    double binsum;   // pedestal accumulated for the current cb (same value in every lane)
};

template <int K>
LBL_HD void ped_lane_init(PedLane<K>& st)
{
    st.have = false;
    st.cb = 0;
    st.binsum = 0.;
    st.ks = st.ke = 0.;
#pragma unroll
    for (int k = 0; k < K; ++k)
    {
        st.own[k] = 0.;
        st.idx[k] = -1;
        st.slot[k] = -1;
    }
}

// Storage indices of the window's two deciding points k[s], k[e].
LBL_HD int ped_s_index(const PedWindow& w)
{
    return w.s_node;
}
LBL_HD int ped_e_index(const PedWindow& w, const GridSpec& g)
{
    return w.tail ? g.ncell : w.e_node;
}

// Window move, lane-local part: re-map the ring onto the new window, swapping values with
// the backing array where the held index changes.
template <int K>
LBL_HD void ped_lane_move(PedLane<K>& st, const GridSpec& g, int lane, int cb, const PedWindow& w,
                          double* backing, double* bins)
{
    if (st.have && lane == 0)
    {
        bins[st.cb + g.cut_off + 1] += st.binsum;
    }
    st.cb = cb;
    st.have = true;
    st.w = w;
    st.binsum = 0.;
    const int lo = ped_s_index(w);
    const int hi = ped_e_index(w, g);
#pragma unroll
    for (int k = 0; k < K; ++k)
    {
        // the unique index in [lo, lo + 32K) congruent to lane + 32k
        const int c = lo + ((lane + 32 * k - lo) & (32 * K - 1));
        if (c != st.idx[k])
        {
            if (st.idx[k] >= 0 && st.idx[k] <= g.ncell) backing[st.idx[k]] = st.own[k];
            st.own[k] = (c <= g.ncell) ? backing[c] : 0.;
            st.idx[k] = c;
        }
        st.slot[k] = (c > hi) ? -1 : ((c == g.ncell && w.tail) ? 2 * g.cut_off + 2 : c - w.base);
    }
}

// The value this lane holds for storage index c (0 if it does not hold it).
template <int K>
LBL_HD double ped_lane_value(const PedLane<K>& st, int c)
{
    double v = 0.;
#pragma unroll
    for (int k = 0; k < K; ++k)
    {
        if (st.idx[k] == c) v = st.own[k];
    }
    return v;
}

// One line: k[s] += f[s]; k[e] += f[e]; pedestal = min(k[s], k[e]); both -= pedestal
// (spectra.c:65-77 restricted to the two nodes that decide the pedestal).
LBL_HD double ped_line_value(double ks_prev, double ke_prev, double fs, double fe, double& ks_new,
                             double& ke_new)
{
    const double ks1 = ks_prev + fs;
    const double ke1 = ke_prev + fe;
    const double pedestal = (ke1 < ks1) ? ke1 : ks1;  // spectra.c:68-72
    ks_new = ks1 - pedestal;
    ke_new = ke1 - pedestal;
    return pedestal;
}

// Node update of one run: own += (sum of the run's terms, = K3a's prefix in the run's last
// row) - (sum of the run's pedestals).
template <int K>
LBL_HD void ped_lane_slots(PedLane<K>& st, const double* last_row, double pedsum)
{
#pragma unroll
    for (int k = 0; k < K; ++k)
    {
        if (st.slot[k] >= 0) st.own[k] += last_row[st.slot[k]] - pedsum;
    }
    st.binsum += pedsum;
}

// End of the layer: the last bin's pedestal.
template <int K>
LBL_HD void ped_lane_finish(PedLane<K>& st, const GridSpec& g, int lane, double* bins)
{
    if (st.have && lane == 0)
    {
        bins[st.cb + g.cut_off + 1] += st.binsum;
    }
}

// Generic (slow) form of the whole recurrence for one layer, used when the window is wider
// than the register-resident kernel supports, and by the CPU emulation harness.
template <class Sync>
LBL_HD void pedestal_layer(const PedArgs& a, int layer, int lane, int nlanes, double* nodes,
                           Sync sync)
{
    const GridSpec& g = a.grid;
    const int nb = g.ncell + 2 * g.cut_off + 2;
    const int nslots = 2 * g.cut_off + 3;
    double* bins = a.pedbin + (size_t)layer * nb;
    for (int c = lane; c <= g.ncell; c += nlanes)
    {
        nodes[c] = 0.;
    }
    for (int b = lane; b < nb; b += nlanes)
    {
        bins[b] = 0.;
    }
    sync();
    for (int r = 0; r < a.n_rows; ++r)
    {
        const int j = a.lines.db_to_sorted ? a.lines.db_to_sorted[r] : r;
        const int cb = a.rec.chk[(size_t)layer * a.lines.n + j].cb;
        const PedWindow w = ped_window(cb, g);
        if (w.skip)
        {
            continue;
        }
        for (int t = lane; t < nslots; t += nlanes)
        {
            const int idx = ped_slot_index(w, g, t);
            if (idx >= 0) nodes[idx] += pedestal_term(a, layer, r, t);
        }
        sync();
        const double ks = nodes[ped_slot_index(w, g, w.s_slot)];
        const double ke = nodes[ped_slot_index(w, g, w.e_slot)];
        const double pedestal = (ke < ks) ? ke : ks;  // spectra.c:68-72
        sync();
        for (int t = lane; t < nslots; t += nlanes)
        {
            const int idx = ped_slot_index(w, g, t);
            if (idx >= 0) nodes[idx] -= pedestal;
        }
        if (lane == 0)
        {
            bins[cb + g.cut_off + 1] += pedestal;
        }
        sync();
    }
}

// ---------------------------------------------------------------------------------------
// K3 for nu-sorted databases: the recurrence in closed form per RUN, without node state.
//
// A run is a maximal stretch of consecutive database rows with the same window cell cb (all
// lines of one integer-wavenumber bin, or a fragment of it where pressure shifts put two
// neighbouring lines' cells out of order).  All lines of a run share the window [s, e].
// With k[] the reference's running spectrum, d = k[s] - k[e]: every line adds (f[s], f[e]) and
// then removes min(k[s], k[e]) from both, so after it k[s] = max(d, 0), k[e] = max(-d, 0) and
// d moves by f[s] - f[e] -- the clamping never feeds back into d.  Over a whole run:
//     d_end = (k[s] - k[e])_before + sum(f[s]) - sum(f[e])
//     sum of the run's pedestals = sum(f[e]) + k[e]_before   if d_end > 0   (k[e] ends at zero)
//                                = sum(f[s]) + k[s]_before   otherwise      (telescoping)
// and the values before the run are sums over the EARLIER rows whose windows cover the point,
//     k[x]_before = sum_l f_l(x)  -  sum_l pedestal_l ,
// the first a plain gather (ped_run_sums: every run of every layer in parallel, like the
// summation kernel but at one point per run), the second a sum over the <= 2*cut+2 pedestal
// bins (index cb + cut + 1) of the cells whose windows cover x -- the only sequential part
// (ped_chain_run: a handful of dependent operations per run).  "Earlier" is by database row,
// and the coverage test is on each line's own cell, so out-of-order cells are reproduced.
// Rows of a sorted database that precede a run are the rows before its first row; unsorted
// databases keep the slot-ring kernels above.
// ---------------------------------------------------------------------------------------
struct PedRunArgs
{
    LinesView lines;
    Records rec;
    GridSpec grid;
    const LayerIn* layers;
    int n_rows;          // rows the recurrence walks (<= lines.n)
    int* run_row;        // [layer][n_rows + 1] first row of each run, then the sentinel n_rows
    int* n_runs;         // [layer]
    int* run_cb;         // [layer][n_rows][4] per run: own pedestal bin (-1: skipped), first bins of
                         //   the ranges covering k[s] and k[e], length of the latter
    double* run_sums;    // [layer][n_rows][4] per run: sum f[s], sum f[e] over its lines;
                         //   sum over earlier covering rows of f(s-point), of f(e-point)
    double* pedbin;      // [layer][ncell + 2*cut + 2]
};

// The two grid points that decide a run's pedestals, and for each the range of pedestal bins
// (index cb' + cut + 1) of the cells whose windows cover it:  a node c*n_per_v is covered by
// cb' in [c-cut-1, c+cut]; the last grid point n-1 (when it is not a node) by
// cb' in [ncell-cut-1, ncell-1+cut]  (spectra.c:48-62).
struct PedPoints
{
    bool skip;
    int i_s, i_e;      // grid indices of k[s], k[e]
    int bs, ns;        // bins [bs, bs+ns) cover k[s]
    int be, ne;        // bins [be, be+ne) cover k[e]
};

LBL_HD PedPoints ped_points(int cb, const GridSpec& g)
{
    const PedWindow w = ped_window(cb, g);
    PedPoints p;
    p.skip = w.skip;
    p.i_s = w.s_node * g.n_per_v;
    p.i_e = w.tail ? g.n - 1 : w.e_node * g.n_per_v;
    p.bs = w.s_node;
    p.ns = 2 * g.cut_off + 2;
    p.be = w.tail ? g.ncell : w.e_node;
    p.ne = w.tail ? 2 * g.cut_off + 1 : 2 * g.cut_off + 2;
    return p;
}

// One line at one grid point, by the path the summation kernels take there.
LBL_HD double ped_line_at(const PedRunArgs& a, size_t o, const int4& ck, double v, int i)
{
    if (i >= ck.y && i <= ck.z)
    {
        const LineGen gen = a.rec.gen[o];
        return voigt_general_call(v, gen.nu, gen.repwid, gen.y, gen.cof, gen.xlim0, gen.xlim1);
    }
    const double2 l = LBL_LDG(reinterpret_cast<const double2*>(a.rec.ab + o));
    return far_term(v, l.x, l.y, LBL_LDG(a.rec.cc + o), 0.);
}

// This lane's share (rows row_lo + lane, + nlanes, ...) of the sums of a run's own lines at its
// two points.
LBL_HD void ped_sum_own(const PedRunArgs& a, size_t off, int row_lo, int row_hi, const PedPoints& pp,
                        int lane, int nlanes, double& at_s, double& at_e)
{
    const GridSpec& g = a.grid;
    const double v_s = grid_point(g.v0, g.dv, pp.i_s);
    const double v_e = grid_point(g.v0, g.dv, pp.i_e);
    LBL_CHECK(row_lo >= 0 && row_hi <= a.lines.n && pp.i_s >= 0 && pp.i_e < g.n);
    for (int j = row_lo + lane; j < row_hi; j += nlanes)
    {
        const int4 ck = LBL_LDG(reinterpret_cast<const int4*>(a.rec.chk + off + j));
        at_s += ped_line_at(a, off + j, ck, v_s, pp.i_s);
        at_e += ped_line_at(a, off + j, ck, v_e, pp.i_e);
    }
}

// First earlier row that can cover a point whose covering bins start at `first_bin`: cell >=
// first covering cell, i.e. shifted centre >= v0 + that cell, i.e. unshifted centre within
// `slack` of it.
LBL_HD int ped_first_covering_row(const PedRunArgs& a, int layer, int first_bin)
{
    const GridSpec& g = a.grid;
    return first_line_at(a.lines, (double)g.v0 + (double)(first_bin - g.cut_off - 1) - a.layers[layer].slack);
}

// This lane's share of the sum, at grid point i, of the rows before row_lo whose pedestal bin
// lies in [first_bin, first_bin + n_bins).
LBL_HD double ped_sum_before(const PedRunArgs& a, int layer, size_t off, int row_lo, int first_bin, int n_bins,
                             int i, int lane, int nlanes)
{
    const GridSpec& g = a.grid;
    const double v = grid_point(g.v0, g.dv, i);
    double sum = 0.;
    LBL_CHECK(ped_first_covering_row(a, layer, first_bin) >= 0 && row_lo <= a.lines.n);
    for (int j = ped_first_covering_row(a, layer, first_bin) + lane; j < row_lo; j += nlanes)
    {
        const int4 ck = LBL_LDG(reinterpret_cast<const int4*>(a.rec.chk + off + j));
        if ((unsigned)(ck.x + g.cut_off + 1 - first_bin) < (unsigned)n_bins)
        {
            sum += ped_line_at(a, off + j, ck, v, i);
        }
    }
    return sum;
}

// This lane's share of the four sums of the run [row_lo, row_hi).
LBL_HD void ped_run_sums(const PedRunArgs& a, int layer, int row_lo, int row_hi, int lane, int nlanes,
                         double (&out)[4])
{
    const size_t off = (size_t)layer * a.lines.n;
    out[0] = out[1] = out[2] = out[3] = 0.;
    const int cb = LBL_LDG(reinterpret_cast<const int4*>(a.rec.chk + off + row_lo)).x;
    const PedPoints pp = ped_points(cb, a.grid);
    if (pp.skip)
    {
        return;
    }
    ped_sum_own(a, off, row_lo, row_hi, pp, lane, nlanes, out[0], out[1]);
    out[2] = ped_sum_before(a, layer, off, row_lo, pp.bs, pp.ns, pp.i_s, lane, nlanes);
    out[3] = ped_sum_before(a, layer, off, row_lo, pp.be, pp.ne, pp.i_e, lane, nlanes);
}

// The sequential step: the run's pedestal sum from its four gathered sums and the pedestals
// already binned (sum_ps, sum_pe: over the bins that cover k[s], k[e]).
LBL_HD double ped_chain_run(const double (&sums)[4], double sum_ps, double sum_pe)
{
    const double ks0 = sums[2] - sum_ps;
    const double ke0 = sums[3] - sum_pe;
    const double d_end = (ks0 - ke0) + (sums[0] - sums[1]);
    // The pedestals telescope on either side: sum = sum f[s] + k[s]_before - k[s]_after
    //                                             = sum f[e] + k[e]_before - k[e]_after,
    // and one of k[s]_after = max(d_end, 0), k[e]_after = max(-d_end, 0) is exactly zero: taking
    // that side keeps the sum free of cancellation (k[s] next to a strong line's core can be
    // 1e10 times k[e]; the reference's min() returns the small one exactly).
    return d_end > 0. ? sums[1] + ke0 : sums[0] + ks0;
}

// ---------------------------------------------------------------------------------------
// The chain over runs as a (min,+) scan.  Call a stretch of consecutive runs REGULAR when their
// pedestal bins are strictly increasing and above every bin touched so far, none is skipped,
// every run's k[e] range is still empty (no earlier line reaches its k[e]), and the first
// run's bin lies inside every run's k[s] range.  (In a nu-sorted database all runs are like
// that except where pressure shifts put neighbouring lines' cells out of order, and the last
// 2*cut+2 cells, whose k[e] is the clamped end of the grid.)  For such a stretch, with
//     alpha_i = sum f[e] + B_i                       (k[e] before the run is B_i: no pedestal reaches it)
//     h_i     = sum f[s] + A_i - W_i ,   W_i = the pedestals already binned inside the run's k[s] range
//     q_i     = sum of the pedestals of runs 0..i of the stretch,
// ped_chain_run reads  P_i = min(alpha_i, h_i - q_(i-1)),  i.e.  q_i = min(q_(i-1) + alpha_i, h_i):
// a Lindley-type recursion whose solution is a prefix minimum,
//     q_i = S_i + min(0, min_(k<=i) (h_k - S_k)) ,   S_i = alpha_0 + ... + alpha_i ,
// two warp scans for up to 32 runs instead of 32 dependent steps.  All sums stay local to the
// stretch and the 2*cut+2 bins before it (no spectrum-wide prefix sums, so a distant strong band
// cannot cancel against local values).  Irregular runs take the sequential step (ped_chain_run
// with explicit sums over the bins).
// ---------------------------------------------------------------------------------------
struct PedRunInfo
{
    int bin;       // own pedestal bin (cb + cut + 1), -1: the reference skips these lines
    int bs;        // first bin of the k[s] range [bs, bs + 2*cut + 2)
    int be, ne;    // k[e] range [be, be + ne)
    double sums[4];
};

// Is run `cur` the continuation of a regular stretch that started at bin `first_bin`, after a
// run (or, for the first run, after everything so far) whose highest bin is `prev_top`?
LBL_HD bool ped_run_regular(const PedRunInfo& cur, int first_bin, int prev_top)
{
    return cur.bin >= 0 && cur.bin > prev_top && cur.be > prev_top && cur.bs <= first_bin;
}

// W_i: pedestals binned in [bs, first_bin) (sequential sum: few terms, deterministic).
LBL_HD double ped_prior_window(const double* bins, int bs, int first_bin)
{
    double w = 0.;
    for (int b = bs; b < first_bin; ++b) w += bins[b];
    return w;
}

// K4a: pedestal seen by the points of one cell: corr[0] for r > 0, corr[1] for r == 0.
LBL_HD void pedestal_cell(const double* bins, int cell, int cut_off, double* corr2)
{
    // Lines with cb in [cell-cut, cell+cut] cover every point of the cell; the cell's first
    // point is also covered by cb == cell-cut-1.  bins index = cb + cut + 1.
    double sum = 0.;
    for (int b = cell + 1; b <= cell + 2 * cut_off + 1; ++b)
    {
        sum += bins[b];
    }
    corr2[0] = sum;
    corr2[1] = bins[cell] + sum;
}

}  // namespace lbl
