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
};

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
};

struct GridSpec
{
    int v0, vn, n_per_v, cut_off;
    int n;      // (vn - v0)*n_per_v output points per layer
    int ncell;  // vn - v0
    double dv;  // 1./n_per_v
};

// ---------------------------------------------------------------------------------------
// K1: scaling of one (layer, line).  Returns the line's window size (e - s + 1, the
// reference's count of grid evaluations, spectra.c:48-62) for the eval counter.
// ---------------------------------------------------------------------------------------
LBL_HD long long scale_thread(const LinesView& ln, const TipsView& tips, const LayerIn* layers,
                              const GridSpec& g, const Records& rec, int layer, int j)
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
    return (e >= s) ? (e - s + 1) : 0;
}

// ---------------------------------------------------------------------------------------
// K2: gather summation.  One thread owns P consecutive grid points that lie in one
// integer-wavenumber cell (P divides n_per_v); the 32 threads of a warp own 32*P
// consecutive points and walk the same line ranges.
// ---------------------------------------------------------------------------------------
struct SumArgs
{
    LinesView lines;
    Records rec;
    const LayerIn* layers;
    GridSpec grid;
    double* out;  // [layer][n]
};

template <int P>
LBL_HD void plain_range(const FarAB* __restrict__ ab, const double* __restrict__ cc, int jb, int je,
                        const double (&v)[P], double (&acc)[P])
{
#pragma unroll 2
    for (int j = jb; j < je; ++j)
    {
        const double2 l = LBL_LDG(reinterpret_cast<const double2*>(ab + j));
        const double c = LBL_LDG(cc + j);
#pragma unroll
        for (int p = 0; p < P; ++p)
        {
            acc[p] = far_term(v[p], l.x, l.y, c, acc[p]);
        }
    }
}

template <int P>
LBL_HD void checked_range(const FarAB* __restrict__ ab, const double* __restrict__ cc,
                          const LineChk* __restrict__ chk, const LineGen* __restrict__ gen,
                          int jb, int je, int i_first, int cell, bool owns_node, int cut_off,
                          const double (&v)[P], double (&acc)[P])
{
    const int cmin = cell - cut_off;
    const int cmax = cell + cut_off;
    const int i_last = i_first + P - 1;
    for (int j = jb; j < je; ++j)
    {
        const int4 ck = LBL_LDG(reinterpret_cast<const int4*>(chk + j));  // cb, nlo, nhi
        // Window membership (from spectra.c:48-62): all points of
        // cell c see lines with c-cut <= cb <= c+cut; the cell's first point (r == 0) also
        // sees cb == c-cut-1, because that line's inclusive end index e lands on it.
        const bool core = (ck.x >= cmin) && (ck.x <= cmax);
        const bool node = core || (owns_node && ck.x == cmin - 1);
        if (!node)
        {
            continue;
        }
        const double2 l = LBL_LDG(reinterpret_cast<const double2*>(ab + j));
        const double c = LBL_LDG(cc + j);
        const bool near = (ck.y <= i_last) && (ck.z >= i_first);
        if (!near)
        {
            acc[0] = far_term(v[0], l.x, l.y, c, acc[0]);
            if (core)
            {
#pragma unroll
                for (int p = 1; p < P; ++p)
                {
                    acc[p] = far_term(v[p], l.x, l.y, c, acc[p]);
                }
            }
        }
        else
        {
            const double2 g0 = LBL_LDG(reinterpret_cast<const double2*>(gen + j));
            const double2 g1 = LBL_LDG(reinterpret_cast<const double2*>(gen + j) + 1);
#pragma unroll
            for (int p = 0; p < P; ++p)
            {
                if (p == 0 || core)
                {
                    const int i = i_first + p;
                    if (i >= ck.y && i <= ck.z)
                    {
                        acc[p] += voigt_general(v[p], g0.x, g0.y, g1.x, g1.y);
                    }
                    else
                    {
                        acc[p] = far_term(v[p], l.x, l.y, c, acc[p]);
                    }
                }
            }
        }
    }
}

template <int P>
LBL_HD void sum_thread(const SumArgs& a, int layer, int tid)
{
    const GridSpec& g = a.grid;
    const int warp_first = (tid & ~31) * P;
    if (warp_first >= g.n)
    {
        return;
    }
    int warp_last = warp_first + 32 * P - 1;
    if (warp_last > g.n - 1)
    {
        warp_last = g.n - 1;
    }
    const LayerIn ly = a.layers[layer];
    const Segments seg = find_segments(a.lines.nu, a.lines.n, g.v0, g.n_per_v, g.dv, g.cut_off,
                                       warp_first, warp_last, ly.slack, ly.kappa);
    int i_first = tid * P;
    const bool valid = i_first < g.n;
    if (!valid)
    {
        i_first = g.n - P;  // idle lanes of the last warp shadow a real thread, store nothing
    }
    const int cell = i_first / g.n_per_v;
    const bool owns_node = (i_first - cell * g.n_per_v) == 0;

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
    const LineGen* gen = a.rec.gen + off;

    checked_range<P>(ab, cc, chk, gen, seg.j[0], seg.j[1], i_first, cell, owns_node, g.cut_off, v, acc);
    plain_range<P>(ab, cc, seg.j[1], seg.j[2], v, acc);
    checked_range<P>(ab, cc, chk, gen, seg.j[2], seg.j[3], i_first, cell, owns_node, g.cut_off, v, acc);
    plain_range<P>(ab, cc, seg.j[3], seg.j[4], v, acc);
    checked_range<P>(ab, cc, chk, gen, seg.j[4], seg.j[5], i_first, cell, owns_node, g.cut_off, v, acc);

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
// K3: pedestal recurrence (spectra.c:66-78), one warp per layer.
//
// The reference subtracts, after each line, min(k[s], k[e]) of the ACCUMULATED spectrum
// from k[s..e], so the pedestal of line l depends on all earlier lines.
// s and e are multiples of n_per_v (or the clamps 0 and n-1), hence only the running
// spectrum at integer-wavenumber nodes -- plus the last grid point n-1 -- is needed.
// This walks the lines in DATABASE order, keeps those tracked points in `nodes`
// (ncell + 1 doubles), and accumulates each line's pedestal into `pedbin[cb + cut + 1]`.
// ---------------------------------------------------------------------------------------
struct PedArgs
{
    LinesView lines;
    Records rec;
    GridSpec grid;
    double* pedbin;  // [layer][ncell + 2*cut + 2]
};

template <class Sync>
LBL_HD void pedestal_layer(const PedArgs& a, int layer, int lane, int nlanes, double* nodes,
                           Sync sync)
{
    const GridSpec& g = a.grid;
    const int nb = g.ncell + 2 * g.cut_off + 2;
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
    const size_t off = (size_t)layer * a.lines.n;
    const bool tail_point = g.n_per_v > 1;  // is grid point n-1 distinct from the last node?
    for (int r = 0; r < a.lines.n; ++r)
    {
        const int j = a.lines.db_to_sorted ? a.lines.db_to_sorted[r] : r;
        const LineChk chk = a.rec.chk[off + j];
        const int cb = chk.cb;
        if (cb - g.cut_off >= g.ncell)
        {
            continue;  // s >= n, spectra.c:49-53
        }
        const int e_raw = cb + g.cut_off + 1;
        if (e_raw < 0)
        {
            continue;  // e < 0: undefined behaviour in the reference; contributes nothing
        }
        const int s_c = (cb - g.cut_off > 0) ? cb - g.cut_off : 0;
        const bool e_clamped = e_raw >= g.ncell;
        const int e_node = e_clamped ? g.ncell - 1 : e_raw;
        const bool use_tail = e_clamped && tail_point;
        const FarAB ab = a.rec.ab[off + j];
        const double cc = a.rec.cc[off + j];
        const LineGen gen = a.rec.gen[off + j];
        for (int c = s_c + lane; c <= e_node; c += nlanes)
        {
            const int i = c * g.n_per_v;
            nodes[c] += line_point(grid_point(g.v0, g.dv, i), i, ab, cc, chk, gen);
        }
        if (use_tail && lane == nlanes - 1)
        {
            const int i = g.n - 1;
            nodes[g.ncell] += line_point(grid_point(g.v0, g.dv, i), i, ab, cc, chk, gen);
        }
        sync();
        const double ks = nodes[s_c];
        const double ke = use_tail ? nodes[g.ncell] : nodes[e_node];
        const double pedestal = (ke < ks) ? ke : ks;  // spectra.c:68-72
        sync();
        for (int c = s_c + lane; c <= e_node; c += nlanes)
        {
            nodes[c] -= pedestal;
        }
        if (use_tail && lane == nlanes - 1)
        {
            nodes[g.ncell] -= pedestal;
        }
        if (lane == 0)
        {
            bins[cb + g.cut_off + 1] += pedestal;
        }
        sync();
    }
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
