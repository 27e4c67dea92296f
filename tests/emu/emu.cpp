// emu.cpp -- CPU emulation of the kernels' per-thread bodies (TEST INFRASTRUCTURE ONLY).
//
// Compiles pylbl_b200/csrc/lbl_threads.cuh as plain C++ and runs the exact per-thread code
// of K1/K2/K3/K4 in sequential loops over (layer, thread).  It lets the CPU-only test suite
// check the gather logic (window membership, segment search, near/far split, pedestal
// recurrence) against the oracle without a GPU.  The product never loads this file; the
// MUFU reciprocal seed is emulated (see rcp_seed in lbl_core.cuh).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../pylbl_b200/csrc/lbl_cheb.h"
#include "../../pylbl_b200/csrc/lbl_threads.cuh"
#include "../../pylbl_b200/csrc/lbl_bands.h"

using namespace lbl;

template <int P>
static void run_sum(const SumArgs& a, int n_layers, bool fp32)
{
    const int tiles = (a.grid.n + a.tpw * P - 1) / (a.tpw * P) + 3;   // + idle warps of the last block
    const int lp = 32 / a.tpw;
    for (int group = 0; group < (n_layers + lp - 1) / lp; ++group)
    {
        for (int tile = 0; tile < tiles; ++tile)
        {
            for (int lane = 0; lane < 32; ++lane)
            {
                if (fp32)
                {
                    sum32_thread<P>(a, group, tile, lane);
                }
                else
                {
                    sum_thread<P>(a, group, tile, lane);
                }
            }
        }
    }
}

template <int G>
static void run_cells(const CellArgs& ca, const LinesView& ln, const GridSpec& g,
                      const std::vector<LayerIn>& layers, int n_layers)
{
    std::vector<double> fields((size_t)G * kNodes), fields16((size_t)G * kNodes16, 0.);
    std::vector<double> fields8((size_t)G * kNodes8, 0.);
    const int chunks = (g.n_per_v + 32 * kCellP - 1) / (32 * kCellP);
    for (int layer = 0; layer < n_layers; ++layer)
    {
        for (int cell = 0; cell < g.ncell; cell += G)
        {
            const CellSegments seg = cell_segments(ln, g, layers[layer], cell, G);
            const int cells = std::min(G, g.ncell - cell);
            std::fill(fields16.begin(), fields16.end(), 0.);
            std::fill(fields8.begin(), fields8.end(), 0.);
            for (int lane = 0; lane < 32; ++lane)
            {
                double f32[G], f16, f8;
                cell_far_lane<G>(ca, layer, cell, lane, seg, f32, f16, f8);
                for (int q = 0; q < G; ++q) fields[(size_t)q * kNodes + lane] = f32[q];
                const Lane16<G> m = lane16<G>(lane);
                fields16[(size_t)m.cell_off * kNodes16 + m.node] += f16;   // G == 1: two partial sums
                const Lane16<G> m8 = lane8<G>(lane);
                fields8[(size_t)m8.cell_off * kNodes8 + m8.node] += f8;    // 2 or 4 partial sums
            }
            for (int q = 0; q < cells; ++q)
                for (int chunk = 0; chunk < chunks; ++chunk)
                    for (int lane = 0; lane < 32; ++lane)
                        cell_direct_lane(ca, layer, cell + q, chunk, lane, seg);
            std::vector<double> coef((size_t)G * kNodes);
            for (int q = 0; q < G; ++q)
                for (int lane = 0; lane < 32; ++lane)
                    coef[(size_t)q * kNodes + lane] =
                        cell_coefficient(ca.transform, fields.data() + (size_t)q * kNodes, kNodes, lane) +
                        cell_coefficient(ca.transform16, fields16.data() + (size_t)q * kNodes16, kNodes16, lane) +
                        cell_coefficient(ca.transform8, fields8.data() + (size_t)q * kNodes8, kNodes8, lane);
            for (int q = 0; q < cells; ++q)
                for (int lane = 0; lane < 32; ++lane)
                    cell_field_lane(ca, layer, cell + q, lane, 32, coef.data() + (size_t)q * kNodes);
        }
    }
}

struct NoSync
{
    void operator()() const {}
};

template <int T>
static void run_fixup(const SumArgs& a, int n_layers)
{
    const int tiles = (a.grid.n + T - 1) / T;
    const int lp = 32 / T;
    const int groups = (n_layers + lp - 1) / lp;
    for (int group = 0; group < groups; ++group)
    {
        for (int tile = 0; tile < tiles; ++tile)
        {
            for (int lane = 0; lane < 32; ++lane)
            {
                fixup_thread<T>(a, tile, group, lane);
            }
        }
    }
}

// The run-organised chain of K3b: 32 emulated lanes, warp collectives done sequentially.
template <int K>
static void run_chain(const PedArgs& pa, int layer, double* nodes)
{
    const GridSpec& g = pa.grid;
    const int wpad = 32 * K;
    const int nb = g.ncell + 2 * g.cut_off + 2;
    double* bins = pa.pedbin + (size_t)layer * nb;
    for (int c = 0; c <= g.ncell; ++c) nodes[c] = 0.;
    for (int b = 0; b < nb; ++b) bins[b] = 0.;
    std::vector<PedLane<K>> lanes(32);
    for (auto& st : lanes) ped_lane_init(st);
    const int n = pa.lines.n;
    std::vector<int> cells(n);
    for (int r = 0; r < n; ++r)
    {
        const int j = pa.lines.db_to_sorted ? pa.lines.db_to_sorted[r] : r;
        cells[r] = pa.rec.chk[(size_t)layer * n + j].cb;
    }
    std::vector<double> rows((size_t)32 * wpad);
    for (int first = 0; first < n; first += 32)   // tiles of 32 lines, as staged by cp.async
    {
        const int cnt = std::min(32, n - first);
        for (int lane = 0; lane < 32; ++lane)
            pedestal_terms_tile<K>(pa, layer, first / 32, lane, rows.data());
        int l = 0;
        while (l < cnt)
        {
            const int cb = cells[first + l];
            int run = 1;
            while (l + run < cnt && cells[first + l + run] == cb) ++run;
            if (!lanes[0].have || cb != lanes[0].cb)
            {
                const PedWindow w = ped_window(cb, g);
                if (w.skip) { l += run; continue; }
                for (int lane = 0; lane < 32; ++lane) ped_lane_move(lanes[lane], g, lane, cb, w, nodes, bins);
                const int cs = ped_s_index(w), ce = ped_e_index(w, g);
                const double ks0 = ped_lane_value(lanes[cs & 31], cs);
                const double ke0 = ped_lane_value(lanes[ce & 31], ce);
                for (int lane = 0; lane < 32; ++lane) { lanes[lane].ks = ks0; lanes[lane].ke = ke0; }
            }
            const double* row0 = rows.data() + (size_t)l * wpad;
            const int spare = 2 * g.cut_off + 3;
            PedLane<K>& u = lanes[0];
            double pedsum = 0., ks = u.ks, ke = u.ke;
            if (run <= 2)
            {
                for (int m = 0; m < run; ++m)
                    pedsum += ped_line_value(ks, ke, row0[(size_t)m * wpad + spare],
                                             row0[(size_t)m * wpad + spare + 1], ks, ke);
            }
            else
            {
                double scan = 0., ks_end = ks, ke_end = ke;
                for (int m = 0; m < run; ++m)   // lane m
                {
                    const double fs = row0[(size_t)m * wpad + spare];
                    const double fe = row0[(size_t)m * wpad + spare + 1];
                    const double d_prev = (ks - ke) + scan;   // exclusive prefix
                    const double ks_prev = (m == 0) ? ks : fmax(d_prev, 0.);
                    const double ke_prev = (m == 0) ? ke : fmax(-d_prev, 0.);
                    pedsum += ped_line_value(ks_prev, ke_prev, fs, fe, ks_end, ke_end);
                    scan += fs - fe;
                }
                ks = ks_end;
                ke = ke_end;
            }
            for (int lane = 0; lane < 32; ++lane)
            {
                lanes[lane].ks = ks;
                lanes[lane].ke = ke;
                ped_lane_slots(lanes[lane], row0 + (size_t)(run - 1) * wpad, pedsum);
            }
            l += run;
        }
    }
    for (int lane = 0; lane < 32; ++lane) ped_lane_finish(lanes[lane], g, lane, bins);
}

// The run-based recurrence for nu-sorted rows (PedRunArgs): what ped_run_count/scatter,
// ped_nodes_kernel and ped_chain_runs_kernel do, one emulated lane.
static void run_chain_sorted(const PedRunArgs& ra, int layer, const LayerIn* layers)
{
    const GridSpec& g = ra.grid;
    const int nb = g.ncell + 2 * g.cut_off + 2;
    double* bins = ra.pedbin + (size_t)layer * nb;
    for (int b = 0; b < nb; ++b) bins[b] = 0.;
    const size_t off = (size_t)layer * ra.lines.n;
    std::vector<int> starts;
    for (int j = 0; j < ra.n_rows; ++j)
    {
        if (j == 0 || ra.rec.chk[off + j].cb != ra.rec.chk[off + j - 1].cb) starts.push_back(j);
    }
    starts.push_back(ra.n_rows);
    // the runs' gathered sums and windows (ped_nodes_kernel)
    const int n_runs = (int)starts.size() - 1;
    std::vector<PedRunInfo> runs((size_t)std::max(n_runs, 0));
    for (int r = 0; r < n_runs; ++r)
    {
        ped_run_sums(ra, layer, starts[r], starts[r + 1], 0, 1, runs[r].sums);
        const int cb = ra.rec.chk[off + starts[r]].cb;
        const PedPoints pp = ped_points(cb, g);
        runs[r].bin = pp.skip ? -1 : cb + g.cut_off + 1;
        runs[r].bs = pp.bs;
        runs[r].be = pp.be;
        runs[r].ne = pp.ne;
    }
    // the chain (ped_chain_runs_kernel): regular stretches of up to 32 runs by the (min,+) scan,
    // everything else by the sequential step
    const bool scan = getenv("EMU_PED_SEQUENTIAL") == nullptr;
    const int ns = 2 * g.cut_off + 2;
    int top = -1;
    int r0 = 0;
    while (r0 < n_runs)
    {
        int len = 0;
        if (scan)
        {
            int prev_top = top;
            while (len < 32 && r0 + len < n_runs && ped_run_regular(runs[r0 + len], runs[r0].bin, prev_top))
            {
                prev_top = runs[r0 + len].bin;
                ++len;
            }
        }
        if (len >= 2)
        {
            double prefix = 0., running_min = 0., q_prev = 0.;
            std::vector<double> ped(len);
            for (int i = 0; i < len; ++i)
            {
                const PedRunInfo& ri = runs[r0 + i];
                const double alpha = ri.sums[1] + ri.sums[3];
                const double h = (ri.sums[0] + ri.sums[2]) - ped_prior_window(bins, ri.bs, runs[r0].bin);
                prefix += alpha;                                   // S_i
                running_min = std::min(running_min, h - prefix);   // min(0, min_k (h_k - S_k))
                const double q = prefix + running_min;
                ped[i] = std::min(alpha, h - q_prev);
                q_prev = q;
            }
            for (int i = 0; i < len; ++i) bins[runs[r0 + i].bin] += ped[i];
            top = runs[r0 + len - 1].bin;
            r0 += len;
            continue;
        }
        const PedRunInfo& ri = runs[r0];
        ++r0;
        if (ri.bin < 0) continue;
        double ps = 0., pe = 0.;
        for (int k = 0; k < ns; ++k) ps += bins[ri.bs + k];
        if (ri.be <= top)
        {
            for (int k = 0; k < ri.ne; ++k) pe += bins[ri.be + k];
        }
        bins[ri.bin] += ped_chain_run(ri.sums, ps, pe);
        top = std::max(top, ri.bin);
    }
}

extern "C" int emu_absorption(int n_layers, const double* pressure, const double* temperature,
                              const double* vmr, int v0, int vn, int n_per_v, double* k,
                              int n_lines, const double* nu, const double* sw,
                              const double* gamma_air, const double* gamma_self,
                              const double* n_air, const double* elower, const double* delta_air,
                              const int* local_iso_id, const double* iso_mass, int num_iso,
                              int num_t, const double* tips_t, const double* tips_q, int cut_off,
                              int remove_pedestal, int points_per_thread, int precision,
                              long long* n_evals)
{
    GridSpec g;
    g.v0 = v0;
    g.vn = vn;
    g.n_per_v = n_per_v;
    g.cut_off = cut_off;
    g.n = (vn - v0) * n_per_v;
    g.ncell = vn - v0;
    g.cell_lo = 0;
    g.cell_hi = vn - v0;
    g.dv = 1. / n_per_v;
    std::memset(k, 0, sizeof(double) * (size_t)n_layers * g.n);
    *n_evals = 0;

    // Early-break prefix in database order (absorption.c:80-83).
    int na = 0;
    for (; na < n_lines; ++na)
    {
        if (nu[na] > vn + cut_off + 1 || nu[na] < v0 - (cut_off + 1)) break;
    }
    if (na == 0) return 0;
    // stable sort by unshifted centre
    std::vector<int> order(na), inv(na);
    for (int i = 0; i < na; ++i) order[i] = i;
    for (int i = 1; i < na; ++i)  // insertion sort keeps it stable; inputs are near-sorted
    {
        int x = order[i], j = i - 1;
        while (j >= 0 && nu[order[j]] > nu[x]) { order[j + 1] = order[j]; --j; }
        order[j + 1] = x;
    }
    std::vector<double> c_nu(na), c_sw(na), c_ga(na), c_gs(na), c_na(na), c_el(na), c_da(na), c_m(na);
    std::vector<int> c_iso(na);
    double max_delta = 0., min_mass = 0.;
    for (int j = 0; j < na; ++j)
    {
        const int r = order[j];
        inv[r] = j;
        c_nu[j] = nu[r]; c_sw[j] = sw[r]; c_ga[j] = gamma_air[r]; c_gs[j] = gamma_self[r];
        c_na[j] = n_air[r]; c_el[j] = elower[r]; c_da[j] = delta_air[r];
        int iso = local_iso_id[r];
        if (iso == 0) iso = 10;
        c_iso[j] = iso - 1;
        c_m[j] = iso_mass[iso - 1];
        if (fabs(c_da[j]) > max_delta) max_delta = fabs(c_da[j]);
        if (c_m[j] > 0. && (min_mass == 0. || c_m[j] < min_mass)) min_mass = c_m[j];
    }
    (void)num_iso;
    LinesView ln;
    ln.n = na;
    ln.nu = c_nu.data(); ln.sw = c_sw.data(); ln.gamma_air = c_ga.data();
    ln.gamma_self = c_gs.data(); ln.n_air = c_na.data(); ln.elower = c_el.data();
    ln.delta_air = c_da.data(); ln.mass = c_m.data(); ln.iso = c_iso.data();
    bool identity = true;
    for (int j = 0; j < na; ++j) identity = identity && order[j] == j;
    ln.db_to_sorted = identity ? nullptr : inv.data();   // as the library: no permutation for sorted rows
    // per-wavenumber index, as make_plan builds it
    std::vector<int> cell_first((size_t)((vn - v0) + 2 * cut_off + 7));
    ln.cell_w0 = v0 - cut_off - 3;
    ln.cell_n = (int)cell_first.size();
    for (int k = 0, j = 0; k < ln.cell_n; ++k)
    {
        while (j < na && c_nu[j] < (double)ln.cell_w0 + (double)k) ++j;
        cell_first[k] = j;
    }
    ln.cell_first = cell_first.data();
    TipsView tips{num_iso, num_t, tips_t, tips_q};

    std::vector<LayerIn> layers(n_layers);
    for (int l = 0; l < n_layers; ++l)
    {
        layers[l].pressure = pressure[l];
        layers[l].temperature = temperature[l];
        layers[l].vmr = vmr[l];
        layers[l].slack = fabs(pressure[l] * kPaToAtm) * max_delta * (1. + 1e-9) + 1e-9;
        layers[l].kappa = 148.3 * sqrt(kR2 * fabs(temperature[l]) / (min_mass > 0. ? min_mass : 1.)) / kVlight;
        layers[l].pad = 0.;
    }
    std::vector<FarAB> ab((size_t)na * n_layers);
    std::vector<double> cc((size_t)na * n_layers);
    std::vector<LineChk> chk((size_t)na * n_layers);
    std::vector<LineGen> gen((size_t)na * n_layers);
    const bool fp32 = precision == 1;
    std::vector<Far32> f32(fp32 ? (size_t)na * n_layers : 0);
    std::vector<unsigned long long> amp_max(n_layers, 0ull);
    Records rec{ab.data(), cc.data(), chk.data(), gen.data(), fp32 ? f32.data() : nullptr,
                fp32 ? amp_max.data() : nullptr};
    for (int l = 0; l < n_layers; ++l)
    {
        for (int j = 0; j < na; ++j)
        {
            double amp = 0.;
            *n_evals += scale_thread(ln, tips, layers.data(), g, rec, l, j, amp);
            union { double d; unsigned long long u; } bits;
            bits.d = amp;
            if (amp > 0. && bits.u > amp_max[l]) amp_max[l] = bits.u;
        }
    }
    if (fp32)
    {
        for (int l = 0; l < n_layers; ++l)
            for (int j = 0; j < na; ++j) far32_thread(rec, na, l, j);
    }
    SumArgs sa;
    sa.lines = ln;
    sa.rec = rec;
    sa.layers = layers.data();
    sa.grid = g;
    sa.out = k;
    sa.n_layers = n_layers;
    sa.near_masked = (points_per_thread == 0) ? 0 : 1;
    {
        // same rule as pick_threads_per_layer() in lbl_api.cu
        int tpw = 32;
        while (tpw > 1 && tpw * points_per_thread > 2 * n_per_v) tpw >>= 1;
        while (tpw < 32 && (32 / tpw) > n_layers) tpw <<= 1;
        sa.tpw = tpw;
    }
    if (points_per_thread == 0)
    {
        // K2c: cell-tiled summation with the two-level Chebyshev far field, one emulated warp
        // per group of cells (the library uses 2 cells up to n_per_v = 256, else 1)
        std::vector<double> nodes, weights, nodes16, weights16, nodes8, weights8;
        build_cheb_nodes(kNodes, n_per_v, nodes);
        build_cheb_nodes(kNodes16, n_per_v, nodes16);
        build_cheb_nodes(kNodes8, n_per_v, nodes8);
        build_cheb_transform(kNodes, weights);
        build_cheb_transform(kNodes16, weights16);
        build_cheb_transform(kNodes8, weights8);
        CellArgs ca;
        ca.sum = sa;
        ca.node_offset = nodes.data();
        ca.transform = weights.data();
        ca.node_offset16 = nodes16.data();
        ca.transform16 = weights16.data();
        ca.node_offset8 = nodes8.data();
        ca.transform8 = weights8.data();
        ca.executed = nullptr;
        if (n_per_v <= 256) run_cells<2>(ca, ln, g, layers, n_layers);
        else run_cells<1>(ca, ln, g, layers, n_layers);
    }
    else switch (points_per_thread)
    {
        case 10: run_sum<10>(sa, n_layers, fp32); break;
        case 8: run_sum<8>(sa, n_layers, fp32); break;
        case 5: run_sum<5>(sa, n_layers, fp32); break;
        case 4: run_sum<4>(sa, n_layers, fp32); break;
        case 2: run_sum<2>(sa, n_layers, fp32); break;
        case 1: run_sum<1>(sa, n_layers, fp32); break;
        default: return 1;
    }
    // K2b: same tile rule as pick_fixup_tile() in lbl_api.cu.
    if (n_per_v >= 64) run_fixup<32>(sa, n_layers);
    else if (n_per_v >= 32) run_fixup<16>(sa, n_layers);
    else if (n_per_v >= 16) run_fixup<8>(sa, n_layers);
    else run_fixup<4>(sa, n_layers);
    if (remove_pedestal)
    {
        const int nb = g.ncell + 2 * cut_off + 2;
        std::vector<double> pedbin((size_t)nb * n_layers), nodes(g.ncell + 1);
        std::vector<double> corr((size_t)2 * g.ncell * n_layers);
        PedArgs pa;
        pa.lines = ln;
        pa.rec = rec;
        pa.grid = g;
        pa.pedbin = pedbin.data();
        pa.n_rows = ln.n;
        PedRunArgs ra;
        ra.lines = ln;
        ra.rec = rec;
        ra.grid = g;
        ra.layers = layers.data();
        ra.n_rows = ln.n;
        ra.run_row = nullptr;
        ra.n_runs = nullptr;
        ra.run_cb = nullptr;
        ra.run_sums = nullptr;
        ra.pedbin = pedbin.data();
        // nu-sorted rows: the run-based recurrence; otherwise (or on request) the slot ring
        const bool runs_path = ln.db_to_sorted == nullptr && getenv("EMU_PED_SLOTS") == nullptr;
        for (int l = 0; l < n_layers; ++l)
        {
            if (runs_path)
            {
                run_chain_sorted(ra, l, layers.data());
            }
            else
            switch ((2 * cut_off + 5 + 31) / 32)
            {
                case 1: run_chain<1>(pa, l, nodes.data()); break;
                case 2: run_chain<2>(pa, l, nodes.data()); break;
                case 3:
                case 4: run_chain<4>(pa, l, nodes.data()); break;
                default: pedestal_layer(pa, l, 0, 1, nodes.data(), NoSync()); break;
            }
            for (int c = 0; c < g.ncell; ++c)
            {
                pedestal_cell(pedbin.data() + (size_t)l * nb, c, cut_off,
                              corr.data() + 2 * ((size_t)l * g.ncell + c));
            }
            for (int i = 0; i < g.n; ++i)
            {
                const int c = i / n_per_v;
                const int r = i - c * n_per_v;
                k[(size_t)l * g.n + i] -= corr[2 * ((size_t)l * g.ncell + c) + (r == 0 ? 1 : 0)];
            }
        }
    }
    return 0;
}

// The band partition of lbl_gas_band_edges (pylbl_b200/csrc/lbl_bands.h) on given costs:
// cell_cost[ncell], prefix_cost[ncell + 1] (what a band ending at that cell pays on top).
extern "C" int emu_partition_bands(const double* cell_cost, const double* prefix_cost, int ncell,
                                   int n_bands, int* edges)
{
    std::vector<double> cum((size_t)ncell + 1, 0.), prefix(prefix_cost, prefix_cost + ncell + 1);
    for (int c = 0; c < ncell; ++c) cum[c + 1] = cum[c] + cell_cost[c];
    lbl::partition_bands(cum, prefix, n_bands, edges);
    return 0;
}
