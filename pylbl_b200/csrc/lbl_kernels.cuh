// lbl_kernels.cuh -- sm_100a kernels of the line-by-line path.
//
//   K1  scale_kernel                per-(layer, line) scaling           replaces spectra.c:17-45
//   K2  sum_kernel<P>               direct far-wing sum (coarse grids)  replaces spectra.c:48-65 + voigt.c:79-83
//   K2c sum_cell_kernel<G>          cell-tiled sum with the Chebyshev far field (fine grids), same rows
//   K2b near_block_kernel           near zone (Humlicek regions 1-3, CPF12) and node terms, line-major
//       fixup_kernel<T>             the same, point-major (coarse grids, tiny cut-offs)   voigt.c:84-187
//   K3a pedestal_terms_kernel<K>    line values at the tracked points   spectra.c:66-78
//   K3b pedestal_chain_kernel<K>    the accumulated-pedestal recurrence (pedestal_kernel: generic fallback)
//   K4  pedestal_cells/apply        pedestal correction per point
//
// None of these uses atomics on the spectrum: every output point has one owner, and where
// several warps contribute to a point (near_block_kernel) they do so through private stripes
// that are added in a fixed order.
#pragma once

#include <cuda_runtime.h>

#include "lbl_threads.cuh"

namespace lbl
{

constexpr int kSumBlock = 128;
constexpr int kScaleBlock = 256;

// K1.  grid = (ceil(n_lines/256), layers).  The only atomic in the library is the
// per-layer evaluation COUNTER below (a statistic, not part of the spectrum).
__global__ void __launch_bounds__(kScaleBlock)
scale_kernel(LinesView ln, TipsView tips, const LayerIn* __restrict__ layers, GridSpec g,
             Records rec, unsigned long long* __restrict__ evals)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int layer = blockIdx.y;
    long long w = 0;
    double amp = 0.;
    if (j < ln.n)
    {
        w = scale_thread(ln, tips, layers, g, rec, layer, j, amp);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        w += __shfl_down_sync(0xffffffffu, w, o);
        amp = fmax(amp, __shfl_down_sync(0xffffffffu, amp, o));
    }
    if ((threadIdx.x & 31) == 0)
    {
        if (w != 0) atomicAdd(evals + layer, (unsigned long long)w);
        // FP32 mode: largest amplitude of the layer (non-negative doubles order like integers).
        if (rec.amp_max && amp > 0.)
        {
            atomicMax(rec.amp_max + layer, (unsigned long long)__double_as_longlong(amp));
        }
    }
}

// K1f (FP32 mode).  Thread per (layer, line): FP32 operands from the FP64 records.
__global__ void __launch_bounds__(kScaleBlock)
far32_kernel(Records rec, int n_lines)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_lines)
    {
        far32_thread(rec, n_lines, blockIdx.y, j);
    }
}

// K2.  block = 128 threads = 4 independent warps; warp w of the grid's x dimension covers
// tpw*P points (tile w) of the 32/tpw layers of layer group blockIdx.y.
template <int P>
__global__ void __launch_bounds__(kSumBlock)
sum_kernel(const SumArgs a)
{
    sum_thread<P>(a, blockIdx.y, (blockIdx.x * blockDim.x + threadIdx.x) >> 5, threadIdx.x & 31);
}

// K2c.  Cell-tiled summation with the Chebyshev far field (see lbl_threads.cuh).
// block = 128 threads = 4 independent warps; warp = G consecutive cells of layer blockIdx.y.
//
// Phase 1  lane = node: the far lines at this lane's node of each cell (cell_far_lane).
// Phase 2  lane = kCellP consecutive points: the direct lines in Lorentz form at every point
//          (cell_direct_lane); spectra stored.
// Phase 3  lane = points lane, lane+32, ...: + interpolated far field (cell_field_lane).
// The near zone (profile - Lorentz) and the node terms are added afterwards by K2b.
// ---- TMA bulk copy + mbarrier (PTX; SASS: UBLKCP / SYNCS) ------------------------------------
__device__ __forceinline__ unsigned smem_addr(const void* p)
{
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, unsigned bytes,
                                              unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LBL_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LBL_DONE;\n"
        "bra LBL_WAIT;\n"
        "LBL_DONE:\n"
        "}\n" ::"r"(smem_addr(bar)), "r"(parity)
        : "memory");
}

// Staging ring of the far-field kernel: lines per chunk (8 KB of (a,b) + 4 KB of c) and chunks in
// flight.  Measured on config 2 (K2c per step): 128x6 16.3 ms, 256x3 15.2, 384x2 14.7, 512x2 14.6,
// 640x2 14.9, 512x3 15.5 (shared memory then caps the resident blocks) -- fewer, larger chunks
// mean fewer block-wide barriers and fewer passes over the range table.
#ifndef LBL_STAGE_LINES
#define LBL_STAGE_LINES 512
#endif
#ifndef LBL_STAGES
#define LBL_STAGES 2
#endif
// Warps (cells, or cell pairs) per block of the far-field kernel: they share one staging ring.
#ifndef LBL_CELL_WARPS
#define LBL_CELL_WARPS 4
#endif
constexpr int kCellBlock = 32 * LBL_CELL_WARPS;
constexpr int kStageLines = LBL_STAGE_LINES;
constexpr int kStages = LBL_STAGES;

// Mid lines [jb, je) of one staged chunk, operands in shared memory (index j - base).
template <int G>
__device__ __forceinline__ void node_plain_staged(const double2* __restrict__ s_ab,
                                                  const double* __restrict__ s_cc, int base, int jb,
                                                  int je, const double (&v)[G], double (&sum)[G])
{
    constexpr int U = (G >= 4) ? 1 : 4 / G;
    double acc[U][G];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int c = 0; c < G; ++c) acc[u][c] = 0.;
    int j = jb - base;
    const int stop = je - base;
    LBL_CHECK(j >= 0 && stop <= kStageLines);
    for (; j + 2 * U - 1 < stop; j += 2 * U)
    {
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            const double2 l1 = s_ab[j + 2 * u];
            const double2 l2 = s_ab[j + 2 * u + 1];
            far_terms_pair<G>(v, l1.x, l1.y, s_cc[j + 2 * u], l2.x, l2.y, s_cc[j + 2 * u + 1], acc[u]);
        }
    }
    for (; j < stop; ++j)
    {
        const double2 l = s_ab[j];
        far_terms<G>(v, l.x, l.y, s_cc[j], acc[0]);
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

// Very far lines [jb, je) of one staged chunk at one point per lane (see node16_plain).
__device__ __forceinline__ double node16_plain_staged(const double2* __restrict__ s_ab,
                                                      const double* __restrict__ s_cc, int base,
                                                      int jb, int je, int first, int stride, double v)
{
    double acc[4] = {0., 0., 0., 0.};
    const double vv[1] = {v};
    const int b = jb - base;
    const int pairs = (je - jb) >> 1;
    LBL_CHECK(b >= 0 && je - base <= kStageLines);
    int p = first;
    for (; p + 3 * stride < pairs; p += 4 * stride)
    {
#pragma unroll
        for (int u = 0; u < 4; ++u)
        {
            const int j = b + 2 * (p + u * stride);
            const double2 l1 = s_ab[j];
            const double2 l2 = s_ab[j + 1];
            double one[1] = {acc[u]};
            far_terms_pair<1>(vv, l1.x, l1.y, s_cc[j], l2.x, l2.y, s_cc[j + 1], one);
            acc[u] = one[0];
        }
    }
    for (; p < pairs; p += stride)
    {
        const int j = b + 2 * p;
        const double2 l1 = s_ab[j];
        const double2 l2 = s_ab[j + 1];
        double one[1] = {acc[0]};
        far_terms_pair<1>(vv, l1.x, l1.y, s_cc[j], l2.x, l2.y, s_cc[j + 1], one);
        acc[0] = one[0];
    }
    if (((je - jb) & 1) && first == 0)
    {
        const int j = je - 1 - base;
        const double2 l = s_ab[j];
        acc[1] = far_term(v, l.x, l.y, s_cc[j], acc[1]);
    }
    return (acc[0] + acc[1]) + (acc[2] + acc[3]);
}

// K2c prologue, hoisted out of the summation kernel: the ten range searches of every
// (layer, cell group) -- 4-7 dependent loads each -- are done here, one search per thread, so
// that a warp of the summation kernel starts from one coalesced 40-byte load instead of a chain
// of dependent ones.  grid = (ceil(groups*16/256), layers of the chunk).
__global__ void __launch_bounds__(256)
cell_keys_kernel(LinesView lines, GridSpec g, const LayerIn* __restrict__ layers, int cells_per_group,
                 int first_cell, int groups, int* __restrict__ keys)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int group = idx / kKeyStride;
    const int which = idx - group * kKeyStride;
    if (group >= groups || which >= kCellKeys)
    {
        return;
    }
    const int layer = blockIdx.y;
    const int cell0 = first_cell + group * cells_per_group;
    keys[((size_t)layer * groups + group) * kKeyStride + which] =
        first_line_at(lines, cell_search_key(g, layers[layer], cell0, cells_per_group, which));
}

// M: how the far-range loops are instantiated (2: one generic copy, 1: one copy per node count
// with literal strides, 0: one per range).
template <int G, int M>
#ifndef LBL_CELL_RESIDENT
#define LBL_CELL_RESIDENT (1024 / kCellBlock)
#endif
__global__ void __launch_bounds__(kCellBlock, LBL_CELL_RESIDENT)
sum_cell_kernel(const CellArgs a)
{
    constexpr int kWarps = kCellBlock / 32;
    __shared__ double fields[kWarps][G][kNodes];
    __shared__ double fields16[kWarps][G][kNodes16];
    __shared__ double fields8[kWarps][G][kNodes8];
    __shared__ alignas(16) double2 s_ab[kStages][kStageLines];
    __shared__ alignas(16) double s_cc[kStages][kStageLines];
    __shared__ alignas(8) unsigned long long full[kStages];
    __shared__ int s_range[kWarps][2];
    __shared__ int s_seg[kWarps][kCellKeys];
    const GridSpec& g = a.sum.grid;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int layer = blockIdx.y + a.sum.layer0;
    // A block takes kWarps * G consecutive cells at an ABSOLUTE position of the grid: a spectral
    // band that starts inside such a group finds the same blocks as the whole grid does.  Cells
    // of the group outside the band still give their line ranges to the block's staging range
    // -- the chunks are then cut where the whole-grid call cuts them and every sum is formed in
    // the same order, to the same bits -- but compute nothing.
    const int cell0 = cell_block_base(g, kWarps * G) + (blockIdx.x * kWarps + warp) * G;
    const bool exists = cell0 < g.ncell;      // idle warps of the last block still join the barriers
    const int q_lo = max(0, g.cell_lo - cell0);
    const int q_hi = min(G, g.cell_hi - cell0);
    const bool active = exists && q_lo < q_hi;
    // the ten range boundaries, one per lane, shared by shuffle
    int mine = 0;
    if (exists && lane < kCellKeys)
    {
        if (a.keys)
        {
            const size_t row = (size_t)(blockIdx.y + a.key_layer0) * a.key_groups + (blockIdx.x * kWarps + warp);
            mine = __ldg(a.keys + row * kKeyStride + lane);
        }
        else
        {
            mine = first_line_at(a.sum.lines, cell_search_key(g, a.sum.layers[layer], cell0, G, lane));
        }
    }
    int found[kCellKeys];
#pragma unroll
    for (int which = 0; which < kCellKeys; ++which)
    {
        found[which] = __shfl_sync(0xffffffffu, mine, which);
    }
    const CellSegments seg = cell_segments_from(found);

    // ---- phase 1: far fields at the nodes ------------------------------------------------
    // The four warps of a block own neighbouring cell groups, so their far-line ranges
    // [j1, j4) u [j5, j8) overlap almost entirely.  The block walks the union in chunks that
    // one thread stages into shared memory with TMA bulk copies (ring of kStages, mbarrier per
    // stage); each warp takes from a chunk what lies inside its own ranges.
    if (lane == 0)
    {
        s_range[warp][0] = exists ? seg.j[1] : 0x7fffffff;
        s_range[warp][1] = exists ? seg.j[8] : 0;
    }
    if (threadIdx.x == 0)
    {
        for (int st = 0; st < kStages; ++st) mbar_init(&full[st], 1);
    }
    __syncthreads();
    int lo_all = s_range[0][0], hi_all = s_range[0][1];
#pragma unroll
    for (int w = 1; w < kWarps; ++w)
    {
        lo_all = min(lo_all, s_range[w][0]);
        hi_all = max(hi_all, s_range[w][1]);
    }
    const size_t off = (size_t)layer * a.sum.lines.n;
    const FarAB* ab = a.sum.rec.ab + off;
    const double* cc = a.sum.rec.cc + off;
    const LineChk* chk = a.sum.rec.chk + off;
    // 16-byte alignment of the c[] source (8-byte elements): the chunk must start at an even
    // ABSOLUTE element index; at worst this stages one element of the previous layer.
    lo_all -= (int)((off + (size_t)lo_all) & 1);
    const int n_chunks = (hi_all > lo_all) ? (hi_all - lo_all + kStageLines - 1) / kStageLines : 0;
    auto issue = [&](int t) {
        const int stage = t % kStages;
        const int first = lo_all + t * kStageLines;
        int cnt = hi_all - first;
        if (cnt > kStageLines) cnt = kStageLines;
        cnt = (cnt + 1) & ~1;   // the copies move multiples of 16 bytes (buffers carry slack)
        mbar_expect_tx(&full[stage], (unsigned)(cnt * 24));
        bulk_copy_g2s(&s_ab[stage][0], ab + first, (unsigned)(cnt * 16), &full[stage]);
        bulk_copy_g2s(&s_cc[stage][0], cc + first, (unsigned)(cnt * 8), &full[stage]);
    };
    if (threadIdx.x == 0)
    {
        for (int t = 0; t < kStages && t < n_chunks; ++t) issue(t);
    }
    const Lane16<G> m16 = lane16<G>(lane);
    const int my_cell = cell0 + m16.cell_off;
    const double v16 = ((double)g.v0 + (double)my_cell) + a.node_offset16[m16.node];
    const Lane16<G> m8 = lane8<G>(lane);
    const double v8 = ((double)g.v0 + (double)(cell0 + m8.cell_off)) + a.node_offset8[m8.node];
    double v[G], f[G];
    double f16 = 0., f8 = 0.;
#pragma unroll
    for (int q = 0; q < G; ++q)
    {
        v[q] = ((double)g.v0 + (double)(cell0 + q)) + a.node_offset[lane];
        f[q] = 0.;
    }
    // One copy of each loop for all the ranges that use it (`unroll 1`; the range bounds are read
    // from shared memory by index): with every range's loop inlined separately the kernel was 60 KB
    // of code and its warps -- each in a different range -- stalled on instruction fetch more
    // than on anything else.
    if (lane == 0)
    {
#pragma unroll
        for (int q = 0; q < kCellKeys; ++q) s_seg[warp][q] = seg.j[q];
    }
    __syncwarp();
    if (active)
    {
#pragma unroll 1
        for (int side = 0; side < 2; ++side)
        {
            f16 += node16_tested(ab, cc, chk, s_seg[warp][8 * side], s_seg[warp][8 * side + 1], m16.first,
                                 m16.stride, my_cell, g.cut_off, v16);
        }
    }
    for (int t = 0; t < n_chunks; ++t)
    {
        const int stage = t % kStages;
        mbar_wait(&full[stage], (unsigned)((t / kStages) & 1));
        const int first = lo_all + t * kStageLines;
        const int last = min(first + kStageLines, hi_all);
        if (active)
        {
            if (G == 1 && M == 2)
            {
                // ranges j1..j4 and j5..j8: 8-, 16-, 32-node | 32-, 16-, 8-node lines; every kind
                // is "pairs of lines at one point per lane", shared among 4, 2 or 1 groups of lanes
#pragma unroll 1
                for (int k = 0; k < 6; ++k)
                {
                    const int q = (k < 3) ? k + 1 : k + 2;
                    const int b = max(first, s_seg[warp][q]);
                    const int e = min(last, s_seg[warp][q + 1]);
                    if (b < e)
                    {
                        const int kind = (k < 3) ? k : 5 - k;     // 0: 8 nodes, 1: 16 nodes, 2: 32 nodes
                        const double vk = (kind == 0) ? v8 : ((kind == 1) ? v16 : v[0]);
                        const int first_k = (kind == 0) ? m8.first : ((kind == 1) ? m16.first : 0);
                        const int stride_k = (kind == 0) ? m8.stride : ((kind == 1) ? m16.stride : 1);
                        const double r = node16_plain_staged(s_ab[stage], s_cc[stage], first, b, e, first_k,
                                                             stride_k, vk);
                        f8 += (kind == 0) ? r : 0.;
                        f16 += (kind == 1) ? r : 0.;
                        f[0] += (kind == 2) ? r : 0.;
                    }
                }
            }
            else if (M == 1)
            {
#pragma unroll 1
                for (int k = 0; k < 6; ++k)
                {
                    const int q = (k < 3) ? k + 1 : k + 2;
                    const int b = max(first, s_seg[warp][q]);
                    const int e = min(last, s_seg[warp][q + 1]);
                    if (b < e)
                    {
                        // (the stride is a literal in each call: three copies of the loop, not six)
                        const int kind = (k < 3) ? k : 5 - k;     // 0: 8 nodes, 1: 16 nodes, 2: 32 nodes
                        if (kind == 0)
                            f8 += node16_plain_staged(s_ab[stage], s_cc[stage], first, b, e, m8.first, 4 / G, v8);
                        else if (kind == 1)
                            f16 += node16_plain_staged(s_ab[stage], s_cc[stage], first, b, e, m16.first, 2 / G, v16);
                        else if (G == 1)
                            f[0] += node16_plain_staged(s_ab[stage], s_cc[stage], first, b, e, 0, 1, v[0]);
                        else
                            node_plain_staged<G>(s_ab[stage], s_cc[stage], first, b, e, v, f);
                    }
                }
            }
            else
            {
                int b = max(first, seg.j[1]), e = min(last, seg.j[2]);
                if (b < e) f8 += node16_plain_staged(s_ab[stage], s_cc[stage], first, b, e, m8.first, m8.stride, v8);
                b = max(first, seg.j[2]); e = min(last, seg.j[3]);
                if (b < e) f16 += node16_plain_staged(s_ab[stage], s_cc[stage], first, b, e, m16.first, m16.stride, v16);
                b = max(first, seg.j[3]); e = min(last, seg.j[4]);
                if (b < e) node_plain_staged<G>(s_ab[stage], s_cc[stage], first, b, e, v, f);
                b = max(first, seg.j[5]); e = min(last, seg.j[6]);
                if (b < e) node_plain_staged<G>(s_ab[stage], s_cc[stage], first, b, e, v, f);
                b = max(first, seg.j[6]); e = min(last, seg.j[7]);
                if (b < e) f16 += node16_plain_staged(s_ab[stage], s_cc[stage], first, b, e, m16.first, m16.stride, v16);
                b = max(first, seg.j[7]); e = min(last, seg.j[8]);
                if (b < e) f8 += node16_plain_staged(s_ab[stage], s_cc[stage], first, b, e, m8.first, m8.stride, v8);
            }
        }
        __syncthreads();   // every warp is done with this stage
        if (threadIdx.x == 0 && t + kStages < n_chunks) issue(t + kStages);
    }
    if (!active)
    {
        return;
    }
    if (G == 1)
    {
        f16 += __shfl_xor_sync(0xffffffffu, f16, 16);   // the half-warps held partial sums
    }
    // the lanes that shared the lines of an 8-node point add up
    f8 += __shfl_xor_sync(0xffffffffu, f8, 8);
    if (G == 1)
    {
        f8 += __shfl_xor_sync(0xffffffffu, f8, 16);
    }
    double (*field)[kNodes] = fields[warp];
    double (*field16)[kNodes16] = fields16[warp];
    double (*field8)[kNodes8] = fields8[warp];
#pragma unroll
    for (int q = 0; q < G; ++q) field[q][lane] = f[q];
    if (G >= 2 || lane < 16) field16[m16.cell_off][m16.node] = f16;
    if ((lane & 8) == 0 && (G >= 2 || lane < 8)) field8[m8.cell_off][m8.node] = f8;

    // ---- phase 2: direct lines, Lorentz form ------------------------------------------------
    // (Running this phase first, while the first bulk copies are in flight, was measured 2 %
    // slower: the warps of a block then reach the chunk loop out of step.)
    const int chunks = (g.n_per_v + 32 * kCellP - 1) / (32 * kCellP);
    for (int q = q_lo; q < q_hi; ++q)
    {
        for (int chunk = 0; chunk < chunks; ++chunk)
        {
            cell_direct_lane(a, layer, cell0 + q, chunk, lane, seg);
        }
    }
    __syncwarp();   // spectra stored, node sums in shared memory

    // ---- phase 3: + interpolated far fields ---------------------------------------------------
    // node sums -> Chebyshev coefficients (lane = coefficient index); the three series live
    // on the same interval, so their coefficients add up to one series of kNodes terms
    {
        double c[G];
#pragma unroll
        for (int q = 0; q < G; ++q)
        {
            c[q] = cell_coefficient(a.transform, field[q], kNodes, lane) +
                   cell_coefficient(a.transform16, field16[q], kNodes16, lane) +
                   cell_coefficient(a.transform8, field8[q], kNodes8, lane);
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < G; ++q) field[q][lane] = c[q];
        __syncwarp();
    }
    for (int q = q_lo; q < q_hi; ++q)
    {
        cell_field_lane(a, layer, cell0 + q, lane, 32, field[q]);
    }
    if (a.executed && lane == 0)
    {
        // statistics only: Lorentz evaluations this warp performed (nodes + direct slots)
        const unsigned long long mid = (unsigned long long)((seg.j[4] - seg.j[3]) + (seg.j[6] - seg.j[5]));
        const unsigned long long vfar = (unsigned long long)((seg.j[1] - seg.j[0]) + (seg.j[3] - seg.j[2]) +
                                                             (seg.j[7] - seg.j[6]) + (seg.j[9] - seg.j[8]));
        const unsigned long long far8 = (unsigned long long)((seg.j[2] - seg.j[1]) + (seg.j[8] - seg.j[7]));
        const unsigned long long direct = (unsigned long long)(seg.j[5] - seg.j[4]);
        atomicAdd(a.executed, mid * (kNodes * G) + vfar * (kNodes16 * G) + far8 * (kNodes8 * G) +
                              direct * (unsigned long long)((q_hi - q_lo) * chunks * 32 * kCellP));
    }
}

// K2, FP32 mode (opt-in).
template <int P>
__global__ void __launch_bounds__(kSumBlock)
sum32_kernel(const SumArgs a)
{
    sum32_thread<P>(a, blockIdx.y, (blockIdx.x * blockDim.x + threadIdx.x) >> 5, threadIdx.x & 31);
}

// K2b.  One warp per tile of T points x 32/T layers; grid = (ceil(tiles/4), ceil(layers/(32/T))).
//
// Warp-level form of fixup_thread() (lbl_threads.cuh), same terms, different schedule: the
// candidate lines are walked warp-uniformly; lanes whose point falls in a line's outer near
// zone (regions 0/1, the bulk) evaluate it on the spot; the few points per line that need
// the long core branches (regions 2/3, CPF12) are queued as (line, lane) pairs in shared
// memory and evaluated 32 at a time with the lanes packed, instead of once per line with
// two or three lanes active.
constexpr int kFixQueue = 64;

template <int T>
__device__ __forceinline__ double fixup_drain(const SumArgs& a, const int* queue, int cnt, int lane,
                                              int layer, double v, double acc)
{
    const int entry = queue[lane < cnt ? lane : 0];
    const int src = entry & 31;
    const int jj = entry >> 5;
    const double v_src = __shfl_sync(0xffffffffu, v, src);
    const int layer_src = (T == 32) ? layer : __shfl_sync(0xffffffffu, layer, src);
    double val = 0.;
    if (lane < cnt)
    {
        const LineGen* gp = a.rec.gen + (size_t)layer_src * a.lines.n + jj;
        const double2 g0 = __ldg(reinterpret_cast<const double2*>(gp));
        const double2 g1 = __ldg(reinterpret_cast<const double2*>(gp) + 1);
        const double2 g2 = __ldg(reinterpret_cast<const double2*>(gp) + 2);
        val = g1.y * voigt_inner((v_src - g0.x) * g0.y, g1.x);
    }
    for (int e = 0; e < cnt; ++e)
    {
        const double val_e = __shfl_sync(0xffffffffu, val, e);
        const int src_e = __shfl_sync(0xffffffffu, src, e);
        if (lane == src_e) acc += val_e;
    }
    return acc;
}

template <int T>
__device__ __forceinline__ void fixup_warp(const SumArgs& a, int tile, int layer_group, int lane,
                                           int* queue)
{
    const GridSpec& g = a.grid;
    constexpr int LP = 32 / T;
    int layer = layer_group * LP + lane / T;
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
    const int cb_min = cell - g.cut_off - (is_node ? 1 : 0);
    const int cb_max = cell + g.cut_off;
    double acc = 0.;

    int jlo, jhi;
    near_candidates(a.lines, g, ly, t_first, t_last, jlo, jhi);
    int wlo = jlo, whi = jhi;
    if (T != 32)
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            wlo = min(wlo, __shfl_xor_sync(0xffffffffu, wlo, o));
            whi = max(whi, __shfl_xor_sync(0xffffffffu, whi, o));
        }
    }
    int qn = 0;
    for (int j = wlo; j < whi; ++j)
    {
        bool core = false;
        if (j >= jlo && j < jhi)
        {
            const int4 ck = __ldg(reinterpret_cast<const int4*>(chk + j));
            if (i >= ck.y && i <= ck.z)
            {
                // inside the line's window?  s <= i <= e of spectra.c:48-62 in cell form:
                // cell-cut <= cb <= cell+cut, plus cb == cell-cut-1 for a cell's first point
                if (ck.x >= cb_min && ck.x <= cb_max)
                {
                    const double2 g0 = __ldg(reinterpret_cast<const double2*>(gen + j));
                    const double2 g1 = __ldg(reinterpret_cast<const double2*>(gen + j) + 1);
                    const double2 g2 = __ldg(reinterpret_cast<const double2*>(gen + j) + 2);
                    const double abx = fabs((v - g0.x) * g0.y);
                    if (!a.near_masked)
                    {
                        // the summation kernel added the Lorentz form here: take it back
                        const double2 l = __ldg(reinterpret_cast<const double2*>(a.rec.ab + off + j));
                        acc -= far_term_lo(v, l.x, l.y, __ldg(a.rec.cc + off + j), 0.);
                    }
                    if (abx >= voigt_outer_limit(g1.x, g2.x, g2.y))
                    {
                        acc += g1.y * voigt_outer(abx, abx * abx, g1.x, g2.x);
                    }
                    else if (abx >= voigt_region2_limit(g1.x, g2.x))
                    {
                        acc += g1.y * voigt_region2(abx * abx, g1.x);   // short rational: on the spot
                    }
                    else
                    {
                        core = true;   // region 3 / CPF12: queued and lane-packed
                    }
                }
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, core);
        if (m)
        {
            if (core)
            {
                queue[qn + __popc(m & ((1u << lane) - 1u))] = (j << 5) | lane;
            }
            qn += __popc(m);
            __syncwarp();
            if (qn >= 32)
            {
                acc = fixup_drain<T>(a, queue, 32, lane, layer, v, acc);
                __syncwarp();
                const int keep = qn - 32;
                const int moved = (lane < keep) ? queue[32 + lane] : 0;
                __syncwarp();
                if (lane < keep) queue[lane] = moved;
                qn = keep;
                __syncwarp();
            }
        }
    }
    if (qn > 0)
    {
        acc = fixup_drain<T>(a, queue, qn, lane, layer, v, acc);
    }

    // node terms: lines with cb == cell-cut-1 reach exactly the cell's first point.
    if (is_node)
    {
        const double key = (double)g.v0 + (double)(cell - g.cut_off - 1);
        const int nlo = first_line_at(a.lines, key - ly.slack);
        const int nhi = first_line_at(a.lines, key + 1.0 + ly.slack);
        const FarAB* ab = a.rec.ab + off;
        const double* cc = a.rec.cc + off;
        for (int j = nlo; j < nhi; ++j)
        {
            const int4 ck = __ldg(reinterpret_cast<const int4*>(chk + j));
            if (ck.x != cell - g.cut_off - 1 || (i >= ck.y && i <= ck.z))
            {
                continue;  // other cell, or already taken by the near loop above
            }
            const double2 l = __ldg(reinterpret_cast<const double2*>(ab + j));
            acc = far_term(v, l.x, l.y, __ldg(cc + j), acc);
        }
    }
    if (valid)
    {
        a.out[(size_t)layer * g.n + i] += acc;
    }
}

// The T = 32 form (one layer per warp; every grid with at least 64 points per cm-1): the
// candidate lines are the same for all lanes, so the warp first derives their NearLine
// records lane-parallel -- lane m tests line jlo+m against the tile and, if it can touch it,
// loads its records and writes the derived constants to shared memory -- and then walks the
// compacted list with broadcast reads.  No global-memory latency and no per-line setup sits
// in the evaluation loop.
__device__ __forceinline__ void fixup_warp_staged(const SumArgs& a, int tile, int layer, int lane,
                                                  int* queue, NearLine* slots, unsigned* masks)
{
    const GridSpec& g = a.grid;
    int i = tile * 32 + lane;
    const bool valid = i < g.n && i >= band_first_point(g) && i < band_end_point(g);
    if (i >= g.n) i = g.n - 1;
    const int t_first = tile * 32;
    int t_last = t_first + 31;
    if (t_last > g.n - 1) t_last = g.n - 1;

    const LayerIn ly = a.layers[layer];
    const size_t off = (size_t)layer * a.lines.n;
    const LineChk* chk = a.rec.chk + off;
    const LineGen* gen = a.rec.gen + off;
    const double v = grid_point(g.v0, g.dv, i);
    const int cell = i / g.n_per_v;
    const bool is_node = (i - cell * g.n_per_v) == 0;
    // The tile spans at most two cells (n_per_v >= 32): lanes [0, split) lie in cell_a, the
    // rest in cell_a + 1; lane 0 (if t_first is a node) and lane `split` are first points.
    const int cell_a = t_first / g.n_per_v;
    const int split = min((cell_a + 1) * g.n_per_v - t_first, 32);
    const unsigned lanes_a = split >= 32 ? 0xffffffffu : ((1u << split) - 1u);
    const bool first_is_node = cell_a * g.n_per_v == t_first;
    const unsigned below = (1u << lane) - 1u;
    double acc = 0.;

    int jlo, jhi;
    near_candidates(a.lines, g, ly, t_first, t_last, jlo, jhi);
    int qn = 0;
    for (int base = jlo; base < jhi; base += 32)
    {
        // Lane m: which points of the tile does line base+m touch?  Bit l of `mine` = point
        // t_first+l lies in the line's near zone [nlo, nhi] and inside its window
        // (cell-cut <= cb <= cell+cut, plus cb == cell-cut-1 for a cell's first point:
        // s <= i <= e of spectra.c:48-62 in cell form).
        const int j = base + lane;
        unsigned mine = 0;
        int4 ck = make_int4(0, 0, 0, 0);
        if (j < jhi)
        {
            ck = __ldg(reinterpret_cast<const int4*>(chk + j));
            const int lo = max(ck.y - t_first, 0);
            const int hi = min(ck.z - t_first, t_last - t_first);
            if (lo <= hi)
            {
                const unsigned span = (0xffffffffu >> (31 - (hi - lo))) << lo;
                const int da = ck.x - cell_a;
                unsigned win = 0;
                if (da >= -g.cut_off && da <= g.cut_off) win |= lanes_a;
                if (da - 1 >= -g.cut_off && da - 1 <= g.cut_off) win |= ~lanes_a;
                if (first_is_node && da == -g.cut_off - 1) win |= 1u;
                if (split < 32 && da - 1 == -g.cut_off - 1) win |= 1u << split;
                mine = span & win;
            }
        }
        const unsigned listed = __ballot_sync(0xffffffffu, mine != 0);
        if (listed == 0)
        {
            continue;
        }
        if (mine != 0)
        {
            LineGen gn;
            const double2 g0 = __ldg(reinterpret_cast<const double2*>(gen + j));
            const double2 g1 = __ldg(reinterpret_cast<const double2*>(gen + j) + 1);
            const double2 g2 = __ldg(reinterpret_cast<const double2*>(gen + j) + 2);
            gn.nu = g0.x; gn.repwid = g0.y; gn.y = g1.x; gn.cof = g1.y; gn.xlim0 = g2.x; gn.xlim1 = g2.y;
            const double2 l = __ldg(reinterpret_cast<const double2*>(a.rec.ab + off + j));
            FarAB ab;
            ab.a = l.x; ab.b = l.y;
            const int slot = __popc(listed & below);
            slots[slot] = near_line(ck, j, gn, ab, __ldg(a.rec.cc + off + j), a.near_masked != 0);
            masks[slot] = mine;
        }
        __syncwarp();
        const int n_listed = __popc(listed);
        unsigned cores = 0;   // bit s: slot s leaves this lane a region-3 / CPF12 evaluation
        for (int s = 0; s < n_listed; ++s)
        {
            if ((masks[s] >> lane) & 1u)
            {
                bool core;
                acc += near_point(slots[s], v, core);
                if (core) cores |= 1u << s;
            }
        }
        // queue the left-over evaluations, lane-packed
        while (__any_sync(0xffffffffu, cores != 0))
        {
            const bool has = cores != 0;
            const int s = has ? __ffs(cores) - 1 : 0;
            cores &= cores - 1u;
            const unsigned m = __ballot_sync(0xffffffffu, has);
            if (has)
            {
                queue[qn + __popc(m & below)] = ((slots[s].tag >> 1) << 5) | lane;
            }
            qn += __popc(m);
            __syncwarp();
            if (qn >= 32)
            {
                acc = fixup_drain<32>(a, queue, 32, lane, layer, v, acc);
                __syncwarp();
                const int keep = qn - 32;
                const int moved = (lane < keep) ? queue[32 + lane] : 0;
                __syncwarp();
                if (lane < keep) queue[lane] = moved;
                qn = keep;
                __syncwarp();
            }
        }
        __syncwarp();   // the slots are rewritten by the next batch
    }
    if (qn > 0)
    {
        acc = fixup_drain<32>(a, queue, qn, lane, layer, v, acc);
    }

    // node terms: lines with cb == cell-cut-1 reach exactly the cell's first point.  The lanes
    // share the lines of that cell and the owner of the point collects the sum.
    unsigned node_lanes = __ballot_sync(0xffffffffu, is_node);
    while (node_lanes)
    {
        const int src = __ffs(node_lanes) - 1;
        node_lanes &= node_lanes - 1;
        const int cb_node = __shfl_sync(0xffffffffu, cell, src) - g.cut_off - 1;
        const int i_node = __shfl_sync(0xffffffffu, i, src);
        const double v_node = __shfl_sync(0xffffffffu, v, src);
        const double key = (double)g.v0 + (double)cb_node;
        const int nlo = first_line_at(a.lines, key - ly.slack);
        const int nhi = first_line_at(a.lines, key + 1.0 + ly.slack);
        const FarAB* ab = a.rec.ab + off;
        const double* cc = a.rec.cc + off;
        double part = 0.;
        for (int j = nlo + lane; j < nhi; j += 32)
        {
            const int4 ck = __ldg(reinterpret_cast<const int4*>(chk + j));
            if (ck.x == cb_node && !(i_node >= ck.y && i_node <= ck.z))   // else: the near loop's
            {
                const double2 l = __ldg(reinterpret_cast<const double2*>(ab + j));
                part = far_term(v_node, l.x, l.y, __ldg(cc + j), part);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            part += __shfl_xor_sync(0xffffffffu, part, o);
        }
        if (lane == src) acc += part;
    }
    if (valid)
    {
        a.out[(size_t)layer * g.n + i] += acc;
    }
}

// ---------------------------------------------------------------------------------------
// K2b, line-major form (fine grids).  A block owns kNbSpan consecutive points of one layer
// and keeps four private accumulator stripes in shared memory, one per warp.  The lines whose
// near zone meets the span are staged kNbBatch at a time (NearLine records, as in the T = 32
// form); warp w then takes the staged lines w, w+4, ... and walks each line's zone 32 points
// at a time, so every lane holds a point INSIDE the zone (a tile-major warp finds 19 of its 32
// points inside on average), and the per-tile search and set-up is paid once per span.
// Region 3 / CPF12 points are queued per warp and evaluated 32 at a time.  At the end the
// stripes are added in a fixed order -- no atomics, the result does not depend on timing.
// Requires cut_off >= reach + 1 cm-1 (checked by the host): then a point inside a line's near
// zone is always inside the line's window, and no (line, point) pair needs a window test.
// ---------------------------------------------------------------------------------------
constexpr int kNbSpan = 512;
constexpr int kNbBatch = 64;
constexpr int kNbQueue = 64;   // entries per queue: at most 31 left over + 32 new

// Drains `cnt` (<= 32) queued (slot, point) pairs with the lanes packed: lane e evaluates entry
// e and adds it to its point.  kInner == false: points around |x| = xlim1 -- W4 region 2, or
// region 0/1 for the boundary points the index range took along; kInner == true: W4 region 3 /
// CPF12.  Inside xlim1 the Lorentz form is taken back where the summation kernel added it.
template <bool kInner>
__device__ __forceinline__ void near_block_drain(const GridSpec& g, const NearLine* slots, const int* q,
                                                 int cnt, int lane, int p0, double* mine)
{
    const int entry = q[lane < cnt ? lane : 0];
    const int k = entry & 0xffff;          // point, relative to the span
    double val = 0.;
    bool pending = lane < cnt;
    if (pending)
    {
        const NearLine& nl = slots[entry >> 16];
        const double v = grid_point(g.v0, g.dv, p0 + k);
        const double xi = (v - nl.nu) * nl.repwid;
        const bool lorentz_added = (nl.tag & 1) == 0;
        if (kInner)
        {
            val = nl.cof * voigt_inner(xi, nl.y);
            if (lorentz_added) val -= far_term_lo(v, nl.a, nl.b, nl.c, 0.);   // the bits K2c added
        }
        else
        {
            bool core;
            val = near_point(nl, v, core);   // regions 0, 1, 2 (never `core`: sorted at the push)
        }
    }
    // Two entries may name the same point (two lines' cores overlapping): the lowest lane of
    // each group of equal points goes first, the others in later rounds.
    while (true)
    {
        const unsigned todo = __ballot_sync(0xffffffffu, pending);
        if (todo == 0)
        {
            break;
        }
        if (pending)
        {
            const unsigned peers = __match_any_sync(todo, k);
            if (lane == __ffs(peers) - 1)
            {
                LBL_CHECK(k >= 0 && k < kNbSpan);
                mine[k] += val;
                pending = false;
            }
        }
        __syncwarp();
    }
}

// Appends the lanes flagged `push` to a queue (lane-compacted) and drains a full warp's worth.
template <bool kInner>
__device__ __forceinline__ void near_block_push(const GridSpec& g, const NearLine* slots, int* q, int& qn,
                                                bool push, int entry, int lane, unsigned below, int p0,
                                                double* mine)
{
    const unsigned m = __ballot_sync(0xffffffffu, push);
    if (m == 0)
    {
        return;
    }
    if (push)
    {
        LBL_CHECK(qn + __popc(m & below) < kNbQueue);
        q[qn + __popc(m & below)] = entry;
    }
    qn += __popc(m);
    __syncwarp();
    if (qn >= 32)
    {
        near_block_drain<kInner>(g, slots, q, 32, lane, p0, mine);
        __syncwarp();
        const int keep = qn - 32;
        const int moved = (lane < keep) ? q[32 + lane] : 0;
        __syncwarp();
        if (lane < keep) q[lane] = moved;
        qn = keep;
        __syncwarp();
    }
}

// Core index range of a staged line inside the span [p0, p1]: the points with |x| < lim_outer
// (W4 regions 2, 3 and CPF12, voigt.c:98-186) lie in [c_lo, c_hi].  One index of margin either
// side (the bounds are exact to ~1e-10 of an index): the range may hold a few region-0/1
// points, which the queue evaluates as such, but no core point is ever outside it.
__device__ __forceinline__ void near_core_range(const NearLine& nl, const GridSpec& g, int& c_lo, int& c_hi)
{
    const double half = nl.lim_outer / nl.repwid;
    const double lo = ceil((nl.nu - half - (double)g.v0) * (double)g.n_per_v) - 1.;
    const double hi = floor((nl.nu + half - (double)g.v0) * (double)g.n_per_v) + 1.;
    c_lo = (int)fmin(fmax(lo, (double)nl.nlo), (double)nl.nhi + 1.);
    c_hi = (int)fmax(fmin(hi, (double)nl.nhi), (double)c_lo - 1.);
}

#ifndef LBL_NEAR_RESIDENT
#define LBL_NEAR_RESIDENT 8
#endif
__global__ void __launch_bounds__(128, LBL_NEAR_RESIDENT)
near_block_kernel(const SumArgs a)
{
    __shared__ double acc[4][kNbSpan];
    __shared__ __align__(16) NearLine slots[kNbBatch];
    __shared__ int queues[4][2][kNbQueue];
    __shared__ int batch_count[2];
    const GridSpec& g = a.grid;
    const int layer = blockIdx.y + a.layer0;
    const int p0 = (a.tile0 + blockIdx.x) * kNbSpan;   // spans of the whole grid, also for a band
    const int p1 = min(p0 + kNbSpan, g.n) - 1;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const unsigned below = (1u << lane) - 1u;
    for (int k = threadIdx.x; k < 4 * kNbSpan; k += 128) (&acc[0][0])[k] = 0.;
    __syncthreads();

    const LayerIn ly = a.layers[layer];
    const size_t off = (size_t)layer * a.lines.n;
    const LineChk* chk = a.rec.chk + off;
    const LineGen* gen = a.rec.gen + off;
    double* mine = acc[warp];
    int* q_mid = queues[warp][0];     // around |x| = xlim1: W4 region 2 and boundary points
    int* q_in = queues[warp][1];      // W4 region 3 / CPF12: long
    int n_mid = 0, n_in = 0;

    int jlo, jhi;
    near_candidates(a.lines, g, ly, p0, p1, jlo, jhi);
    for (int jb = jlo; jb < jhi; jb += kNbBatch)
    {
        __syncthreads();   // the previous batch is consumed
        const int j = jb + threadIdx.x;
        int4 ck = make_int4(0, 0, 0, 0);
        bool touches = false;
        if (threadIdx.x < kNbBatch && j < jhi)
        {
            ck = __ldg(reinterpret_cast<const int4*>(chk + j));
            touches = ck.z >= p0 && ck.y <= p1;
        }
        const unsigned listed = __ballot_sync(0xffffffffu, touches);
        if (lane == 0 && warp < 2) batch_count[warp] = __popc(listed);
        __syncthreads();
        if (touches)
        {
            LineGen gn;
            const double2 g0 = __ldg(reinterpret_cast<const double2*>(gen + j));
            const double2 g1 = __ldg(reinterpret_cast<const double2*>(gen + j) + 1);
            const double2 g2 = __ldg(reinterpret_cast<const double2*>(gen + j) + 2);
            gn.nu = g0.x; gn.repwid = g0.y; gn.y = g1.x; gn.cof = g1.y; gn.xlim0 = g2.x; gn.xlim1 = g2.y;
            const double2 l = __ldg(reinterpret_cast<const double2*>(a.rec.ab + off + j));
            FarAB ab;
            ab.a = l.x; ab.b = l.y;
            NearLine nl = near_line(ck, j, gn, ab, __ldg(a.rec.cc + off + j), a.near_masked != 0);
            nl.nlo = max(ck.y, p0);   // the zone, clipped to the span
            nl.nhi = min(ck.z, p1);
            near_core_range(nl, g, nl.c_lo, nl.c_hi);
            LBL_CHECK(__popc(listed & below) + (warp == 1 ? batch_count[0] : 0) < kNbBatch);
            slots[__popc(listed & below) + (warp == 1 ? batch_count[0] : 0)] = nl;
        }
        __syncthreads();
        const int n_listed = batch_count[0] + batch_count[1];
        for (int s = warp; s < n_listed; s += 4)
        {
            const NearLine& nl = slots[s];
            const int zlo = nl.nlo, zhi = nl.nhi;
            const int c_lo = nl.c_lo, c_hi = nl.c_hi;
            // ---- the few points inside |x| < xlim1 (and its boundary): sorted by kind into the
            // two queues, evaluated 32 at a time -- no branch of the profile ever runs with two
            // or three lanes active
            for (int i0 = c_lo; i0 <= c_hi; i0 += 32)
            {
                const int i = i0 + lane;
                const bool inside = i <= c_hi;
                const double v = grid_point(g.v0, g.dv, i);
                const double abx = fabs((v - nl.nu) * nl.repwid);
                const bool inner = inside && abx < nl.lim_r2;
                const int entry = (s << 16) | (i - p0);
                near_block_push<false>(g, slots, q_mid, n_mid, inside && !inner, entry, lane, below, p0, mine);
                near_block_push<true>(g, slots, q_in, n_in, inner, entry, lane, below, p0, mine);
            }
            // ---- the zone outside it, W4 regions 0 and 1 (voigt.c:79-97), left part then right
            // part, lanes packed across the gap.  Where the Lorentz form is already in the spectrum
            // region 0 adds nothing and region 1 the closed-form difference (near_point).
            const int n_left = c_lo - zlo;
            const int n_out = n_left + (zhi - c_hi);
            const bool lorentz_added = (nl.tag & 1) == 0;
            const double nu = nl.nu, repwid = nl.repwid, xlim0 = nl.xlim0;
            const double ax = nl.ax, d0 = nl.d0, d2 = nl.d2, n0 = nl.n0, yq = nl.yq;
            for (int t0 = 0; t0 < n_out; t0 += 32)
            {
                const int t = t0 + lane;
                const bool inside = t < n_out;
                const int i = (t < n_left) ? zlo + t : c_hi + 1 + (t - n_left);
                const double v = grid_point(g.v0, g.dv, i);
                const double abx = fabs((v - nu) * repwid);
                const double xq = abx * abx;
                double add = 0.;
                if (lorentz_added)
                {
                    const double den = fma_(xq, d2 + xq, d0) * (xq + yq);
                    add = (abx < xlim0) ? (ax * fma_(1.5, xq, n0)) * rcp_newton2(den) : 0.;
                }
                else
                {
                    add = nl.cof * voigt_outer(abx, xq, nl.y, xlim0);
                }
                LBL_CHECK(!inside || (i >= p0 && i <= p1));
                if (inside) mine[i - p0] += add;
            }
        }
        // the queues name slots of this batch: empty them before the slots are rewritten
        if (n_mid > 0) near_block_drain<false>(g, slots, q_mid, n_mid, lane, p0, mine);
        if (n_in > 0) near_block_drain<true>(g, slots, q_in, n_in, lane, p0, mine);
        n_mid = n_in = 0;
        __syncwarp();
    }

    // node terms: lines with cb == cell-cut-1 reach exactly the cell's first point.  Warp w
    // takes every fourth node of the span; its lanes share the lines of that cell.
    {
        const int c_first = (p0 + g.n_per_v - 1) / g.n_per_v;
        const FarAB* ab = a.rec.ab + off;
        const double* cc = a.rec.cc + off;
        for (int c = c_first + warp; (long long)c * g.n_per_v <= p1; c += 4)
        {
            const int i_node = c * g.n_per_v;
            const int cb_node = c - g.cut_off - 1;
            const double v_node = grid_point(g.v0, g.dv, i_node);
            const double key = (double)g.v0 + (double)cb_node;
            const int nlo = first_line_at(a.lines, key - ly.slack);
            const int nhi = first_line_at(a.lines, key + 1.0 + ly.slack);
            double part = 0.;
            for (int j = nlo + lane; j < nhi; j += 32)
            {
                const int4 ck = __ldg(reinterpret_cast<const int4*>(chk + j));
                if (ck.x == cb_node)
                {
                    if (i_node >= ck.y && i_node <= ck.z)   // near zone 26 cm-1 out: never with
                    {                                       // the host's cut_off condition
                        const LineGen gn = gen[j];
                        part += voigt_general(v_node, gn.nu, gn.repwid, gn.y, gn.cof, gn.xlim0, gn.xlim1);
                    }
                    else
                    {
                        const double2 l = __ldg(reinterpret_cast<const double2*>(ab + j));
                        part = far_term(v_node, l.x, l.y, __ldg(cc + j), part);
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
            {
                part += __shfl_xor_sync(0xffffffffu, part, o);
            }
            LBL_CHECK(i_node >= p0 && i_node <= p1);
            if (lane == 0) mine[i_node - p0] += part;
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; p0 + k <= p1; k += 128)
    {
        const double total = (acc[0][k] + acc[1][k]) + (acc[2][k] + acc[3][k]);
        if (total != 0. && p0 + k >= band_first_point(g) && p0 + k < band_end_point(g))
        {
            a.out[(size_t)layer * g.n + p0 + k] += total;
        }
    }
}

template <int T>
__global__ void __launch_bounds__(128, 10)
fixup_kernel(const SumArgs a)
{
    __shared__ int queues[4][kFixQueue];
    const int tile = a.tile0 + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (T == 32)
    {
        __shared__ __align__(16) NearLine slots[4][32];
        __shared__ unsigned masks[4][32];
        if (tile * 32 >= a.grid.n)
        {
            return;
        }
        fixup_warp_staged(a, tile, blockIdx.y + a.layer0, threadIdx.x & 31, queues[threadIdx.x >> 5],
                          slots[threadIdx.x >> 5], masks[threadIdx.x >> 5]);
        return;
    }
    if (tile * T >= a.grid.n)
    {
        return;
    }
    fixup_warp<T>(a, tile, blockIdx.y, threadIdx.x & 31, queues[threadIdx.x >> 5]);
}

struct WarpSync
{
    __device__ __forceinline__ void operator()() const { __syncwarp(); }
};

// K3 (generic fallback).  grid = layers, block = 32.  `scratch` == nullptr: nodes live in
// dynamic shared memory; otherwise in global memory ([layer][ncell+1]).
__global__ void __launch_bounds__(32)
pedestal_kernel(const PedArgs a, double* scratch)
{
    extern __shared__ double smem_nodes[];
    double* nodes = scratch ? scratch + (size_t)blockIdx.x * (a.grid.ncell + 1) : smem_nodes;
    pedestal_layer(a, blockIdx.x, threadIdx.x, 32, nodes, WarpSync());
}

// One database row of a K3a tile, prepared by the lane that owns the row.
struct __align__(16) PedRow
{
    double a, b, c;            // Lorentz operands (far_term)
    double near_v0, near_v1;   // full-profile values at the (up to two) tracked points inside
                               // the line's near zone
    int base, t_lo, t_hi;      // node of slot 0; slots of the first and last node in the window
    int flags;                 // 1: k[e] is grid point n-1 (slot 2*cut+2); 2: more than two
                               // tracked points in the near zone (generic path for this row)
    int s_slot, e_slot;        // slots of k[s], k[e]
    int near_t0, near_t1;      // slots of near_v0, near_v1 (-1: none)
    int cb, j, nlo, nhi;
};

// K3a.  terms[layer][row r][slot t], 32*K slots per row; one warp per tile of 32 rows.
// grid = (ceil(tiles / 4), layers), block = 128.
// Row m of the tile is prepared by lane m: window, slots, and the full profile at the tracked
// points that fall inside the line's near zone (about every second line has one) -- dense over
// the rows -- and parked in shared memory; the row loop then reads it by broadcast and every
// lane evaluates the Lorentz form at its K slots.
template <int K>
__global__ void __launch_bounds__(128, 6)
pedestal_terms_kernel(const PedArgs a, double* __restrict__ terms)
{
    __shared__ PedRow tile_rows[4][kPedTileRows];
    const GridSpec& g = a.grid;
    const int layer = blockIdx.y;
    const int tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (tile * kPedTileRows >= a.n_rows)
    {
        return;
    }
    constexpr int wpad = 32 * K;
    double* rows = terms + ((size_t)layer * a.lines.n + (size_t)tile * kPedTileRows) * wpad;
    PedRow* mine = tile_rows[threadIdx.x >> 5];
    const int spare = 2 * g.cut_off + 3;
    const int tail_slot = 2 * g.cut_off + 2;
    const int first = tile * kPedTileRows;
    const int cnt = min(a.n_rows - first, kPedTileRows);

    if (lane < cnt)
    {
        PedRow pr;
        pr.j = a.lines.db_to_sorted ? __ldg(a.lines.db_to_sorted + first + lane) : first + lane;
        const size_t o = (size_t)layer * a.lines.n + pr.j;
        const int4 ck = __ldg(reinterpret_cast<const int4*>(a.rec.chk + o));
        const double2 l = __ldg(reinterpret_cast<const double2*>(a.rec.ab + o));
        pr.a = l.x;
        pr.b = l.y;
        pr.c = __ldg(a.rec.cc + o);
        pr.cb = ck.x;
        pr.nlo = ck.y;
        pr.nhi = ck.z;
        const PedWindow w = ped_window(ck.x, g);
        pr.base = w.base;
        pr.t_lo = w.skip ? 1 : w.s_node - w.base;
        pr.t_hi = w.skip ? 0 : w.e_node - w.base;
        pr.flags = (!w.skip && w.tail) ? 1 : 0;
        pr.s_slot = w.skip ? -1 : w.s_slot;
        pr.e_slot = w.skip ? -1 : w.e_slot;
        pr.near_t0 = pr.near_t1 = -1;
        pr.near_v0 = pr.near_v1 = 0.;
        if (!w.skip && ck.y <= ck.z)
        {
            // tracked points inside [nlo, nhi]: nodes idx*n_per_v, and grid point n-1 (tail)
            int lo_idx = ck.y <= 0 ? 0 : (ck.y + g.n_per_v - 1) / g.n_per_v;
            int hi_idx = ck.z < 0 ? -1 : ck.z / g.n_per_v;
            lo_idx = max(lo_idx, w.s_node);
            hi_idx = min(hi_idx, w.e_node);
            const bool tail_near = w.tail && g.n - 1 >= ck.y && g.n - 1 <= ck.z;
            const int n_near = max(hi_idx - lo_idx + 1, 0) + (tail_near ? 1 : 0);
            if (n_near > 2)
            {
                pr.flags |= 2;
            }
            else if (n_near > 0)
            {
                const LineGen gen = a.rec.gen[o];
                int filled = 0;
                for (int idx = lo_idx; idx <= hi_idx; ++idx)
                {
                    const double val = voigt_general(grid_point(g.v0, g.dv, idx * g.n_per_v), gen.nu,
                                                     gen.repwid, gen.y, gen.cof, gen.xlim0, gen.xlim1);
                    if (filled == 0) { pr.near_t0 = idx - w.base; pr.near_v0 = val; }
                    else { pr.near_t1 = idx - w.base; pr.near_v1 = val; }
                    ++filled;
                }
                if (tail_near)
                {
                    const double val = voigt_general(grid_point(g.v0, g.dv, g.n - 1), gen.nu, gen.repwid,
                                                     gen.y, gen.cof, gen.xlim0, gen.xlim1);
                    if (filled == 0) { pr.near_t0 = tail_slot; pr.near_v0 = val; }
                    else { pr.near_t1 = tail_slot; pr.near_v1 = val; }
                }
            }
        }
        mine[lane] = pr;
    }
    __syncwarp();

    double run_sum[K];
    int prev_cb = 0;
    for (int m = 0; m < cnt; ++m)
    {
        const PedRow& pr = mine[m];
        const int cb = pr.cb;
        if (m == 0 || cb != prev_cb)
        {
#pragma unroll
            for (int k = 0; k < K; ++k) run_sum[k] = 0.;
        }
        prev_cb = cb;
        double* row = rows + (size_t)m * wpad;
        const int flags = pr.flags;
        const int t_lo = pr.t_lo, t_hi = pr.t_hi;
        if (t_lo <= t_hi)   // else: the reference does not process this line on this grid
        {
            const double la = pr.a, lb = pr.b, lc = pr.c;
            const int base = pr.base;
#pragma unroll
            for (int k = 0; k < K; ++k)
            {
                const int t = lane + 32 * k;
                const bool is_tail = (flags & 1) && t == tail_slot;
                if ((t >= t_lo && t <= t_hi) || is_tail)
                {
                    const int i = is_tail ? g.n - 1 : (base + t) * g.n_per_v;
                    const double v = grid_point(g.v0, g.dv, i);
                    double val = far_term(v, la, lb, lc, 0.);
                    if (flags & 2)
                    {
                        if (i >= pr.nlo && i <= pr.nhi)
                        {
                            const LineGen gen = a.rec.gen[(size_t)layer * a.lines.n + pr.j];
                            val = voigt_general(v, gen.nu, gen.repwid, gen.y, gen.cof, gen.xlim0, gen.xlim1);
                        }
                    }
                    else
                    {
                        if (t == pr.near_t0) val = pr.near_v0;
                        if (t == pr.near_t1) val = pr.near_v1;
                    }
                    run_sum[k] += val;
                    if (t == pr.s_slot) row[spare] = val;
                    if (t == pr.e_slot) row[spare + 1] = val;
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

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit()
{
    asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

constexpr int kPedTile = kPedTileRows;    // lines per staged tile (= K3a's tile)
constexpr int kPedStages = 4;   // cp.async ring depth

__device__ __forceinline__ double shfl_up_f64(double x, int delta)
{
    return __shfl_up_sync(0xffffffffu, x, delta);
}

// K3b.  grid = layers, block = one warp.  Walks the database-ordered lines of one layer in
// runs of equal window cell (see PedLane).  The per-line terms (K3a) and window cells stream
// in through a 4-stage cp.async ring so that no global-memory latency sits on the chain.
// K = slots per lane (32*K >= 2*cut+3), wpad = 32*K.
// Dynamic shared memory: ring | cells | nodes (ncell+1) | bins (nb)   -- the last two move to
// global memory (`scratch`, a.pedbin) when the grid is too wide.
template <int K>
__global__ void __launch_bounds__(32)
pedestal_chain_kernel(const PedArgs a, const double* __restrict__ terms, double* scratch)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ alignas(8) unsigned long long full[kPedStages];
    const GridSpec& g = a.grid;
    constexpr int wpad = 32 * K;
    const int layer = blockIdx.x;
    const int lane = threadIdx.x;
    const int n = a.n_rows;           // rows walked
    const int stride = a.lines.n;     // rows per layer in the record and term arrays
    const int nb = g.ncell + 2 * g.cut_off + 2;
    double* ring = reinterpret_cast<double*>(smem_raw);
    int4* hdr = reinterpret_cast<int4*>(ring + kPedStages * kPedTile * wpad);
    double* smem_tail = reinterpret_cast<double*>(hdr + kPedStages * kPedTile);
    double* out_bins = a.pedbin + (size_t)layer * nb;
    double* nodes = scratch ? scratch + (size_t)layer * (g.ncell + 1) : smem_tail;
    double* bins = scratch ? out_bins : smem_tail + (g.ncell + 1);
    for (int c = lane; c <= g.ncell; c += 32) nodes[c] = 0.;
    for (int b = lane; b < nb; b += 32) bins[b] = 0.;
    __syncwarp();

    const double* src = terms + (size_t)layer * stride * wpad;
    const LineChk* chk = a.rec.chk + (size_t)layer * stride;
    const int ntiles = (n + kPedTile - 1) / kPedTile;

    auto issue = [&](int t) {
        if (t < ntiles)
        {
            const int stage = t % kPedStages;
            const int first = t * kPedTile;
            const int cnt = (n - first < kPedTile) ? n - first : kPedTile;
            const char* gsrc = reinterpret_cast<const char*>(src + (size_t)first * wpad);
            char* sdst = reinterpret_cast<char*>(ring + (size_t)stage * kPedTile * wpad);
            const int bytes = cnt * wpad * 8;
            if (lane == 0)
            {
                // one TMA bulk copy per tile (16 KB of terms) instead of 32 cp.async per lane:
                // this warp is a serial chain, every instruction it issues is on the clock
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&full[stage], (unsigned)bytes);
                bulk_copy_g2s(sdst, gsrc, (unsigned)bytes, &full[stage]);
            }
            if (lane < cnt)
            {
                const int r = first + lane;
                const int j = a.lines.db_to_sorted ? a.lines.db_to_sorted[r] : r;
                cp_async16(hdr + stage * kPedTile + lane, chk + j);
            }
        }
        cp_async_commit();
    };
    if (lane == 0)
    {
        for (int st = 0; st < kPedStages; ++st) mbar_init(&full[st], 1);
    }
    __syncwarp();
    for (int t = 0; t < kPedStages - 1; ++t) issue(t);

    PedLane<K> st;
    ped_lane_init(st);
    for (int t = 0; t < ntiles; ++t)
    {
        issue(t + kPedStages - 1);
        cp_async_wait<kPedStages - 1>();
        const int stage = t % kPedStages;
        mbar_wait(&full[stage], (unsigned)((t / kPedStages) & 1));
        __syncwarp();
        const int first = t * kPedTile;
        const int cnt = (n - first < kPedTile) ? n - first : kPedTile;
        const double* rows = ring + (size_t)stage * kPedTile * wpad;
        const int4* cells = hdr + stage * kPedTile;
        const int my_cb = cells[lane < cnt ? lane : cnt - 1].x;
        int l = 0;
        while (l < cnt)
        {
            // Run = maximal stretch of lines l, l+1, ... with the same window cell.
            const int cb = __shfl_sync(0xffffffffu, my_cb, l);
            const unsigned differ = __ballot_sync(0xffffffffu, lane >= l && (lane >= cnt || my_cb != cb));
            const int run = (differ ? __ffs(differ) - 1 : 32) - l;
            if (!st.have || cb != st.cb)
            {
                const PedWindow w = ped_window(cb, g);
                if (w.skip)
                {
                    l += run;
                    continue;
                }
                ped_lane_move(st, g, lane, cb, w, nodes, bins);
                // k[s], k[e]: broadcast from the lanes that hold them.
                const int cs = ped_s_index(w), ce = ped_e_index(w, g);
                st.ks = __shfl_sync(0xffffffffu, ped_lane_value(st, cs), cs & 31);
                st.ke = __shfl_sync(0xffffffffu, ped_lane_value(st, ce), ce & 31);
            }
            const double* row0 = rows + (size_t)l * wpad;
            const int spare = 2 * g.cut_off + 3;   // f[s], f[e] of each line (K3a)
            double pedsum = 0.;
            if (run <= 2)
            {
                // Short run: every lane does the same two-node update (no shuffles).
                for (int m = 0; m < run; ++m)
                {
                    const double fs = row0[(size_t)m * wpad + spare];
                    const double fe = row0[(size_t)m * wpad + spare + 1];
                    pedsum += ped_line_value(st.ks, st.ke, fs, fe, st.ks, st.ke);
                }
            }
            else
            {
                // Lane m takes line m of the run: d before it = d0 + sum_{j<m} (fs_j - fe_j),
                // and K3a's rows already hold those sums: row j, slot t = sum over the run's
                // lines 0..j of f[t].  No scan is needed.
                const bool mine = lane < run;
                const double fs = mine ? row0[(size_t)lane * wpad + spare] : 0.;
                const double fe = mine ? row0[(size_t)lane * wpad + spare + 1] : 0.;
                const bool later = mine && lane > 0;
                const double ps = later ? row0[(size_t)(lane - 1) * wpad + st.w.s_slot] : 0.;
                const double pe = later ? row0[(size_t)(lane - 1) * wpad + st.w.e_slot] : 0.;
                const double d_prev = (st.ks - st.ke) + (ps - pe);
                const double ks_prev = (lane == 0) ? st.ks : fmax(d_prev, 0.);
                const double ke_prev = (lane == 0) ? st.ke : fmax(-d_prev, 0.);
                double ks_new, ke_new;
                (void)ped_line_value(ks_prev, ke_prev, fs, fe, ks_new, ke_new);
                // The pedestals telescope: ped_l = ks_prev_l + fs_l - ks_new_l and
                // ks_prev_(l+1) = ks_new_l, so their sum over the run is
                // sum(fs) + ks_before - ks_after, and sum(fs) is the last row's slot s.
                const double ks_before = st.ks;
                st.ks = __shfl_sync(0xffffffffu, ks_new, run - 1);
                st.ke = __shfl_sync(0xffffffffu, ke_new, run - 1);
                pedsum = (row0[(size_t)(run - 1) * wpad + st.w.s_slot] + ks_before) - st.ks;
            }
            ped_lane_slots(st, row0 + (size_t)(run - 1) * wpad, pedsum);
            l += run;
        }
        __syncwarp();
    }
    ped_lane_finish(st, g, lane, bins);
    __syncwarp();
    if (!scratch)
    {
        for (int b = lane; b < nb; b += 32) out_bins[b] = bins[b];
    }
}

// ---------------------------------------------------------------------------------------
// K3 for nu-sorted databases (see PedRunArgs in lbl_threads.cuh): runs -> node sums -> chain.
// ---------------------------------------------------------------------------------------
constexpr int kRunTile = 256;

// K3r-1.  Run starts per tile of kRunTile rows.  grid = (tiles, layers).
__global__ void __launch_bounds__(kRunTile)
ped_run_count_kernel(const LineChk* __restrict__ chk, int stride, int n_rows, int* __restrict__ tile_count)
{
    const int j = blockIdx.x * kRunTile + threadIdx.x;
    const LineChk* c = chk + (size_t)blockIdx.y * stride;
    const bool start = j < n_rows && (j == 0 || c[j].cb != c[j - 1].cb);
    const int count = __syncthreads_count(start);
    if (threadIdx.x == 0) tile_count[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = count;
}

// K3r-2.  Compacts the run starts: run_row[layer][r] = first row of run r (ascending), then the
// sentinel n_rows; n_runs[layer].  grid = (tiles, layers).
__global__ void __launch_bounds__(kRunTile)
ped_run_scatter_kernel(const LineChk* __restrict__ chk, int stride, int n_rows,
                       const int* __restrict__ tile_count, int* __restrict__ run_row,
                       int* __restrict__ n_runs)
{
    __shared__ int warp_count[kRunTile / 32];
    __shared__ int base_s;
    const int layer = blockIdx.y;
    const int tiles = gridDim.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * kRunTile + threadIdx.x;
    const LineChk* c = chk + (size_t)layer * stride;
    const bool start = j < n_rows && (j == 0 || c[j].cb != c[j - 1].cb);
    const unsigned m = __ballot_sync(0xffffffffu, start);
    if (lane == 0) warp_count[warp] = __popc(m);
    if (warp == 0)
    {
        // runs in the tiles before this one
        int before = 0;
        for (int t = lane; t < (int)blockIdx.x; t += 32) before += tile_count[(size_t)layer * tiles + t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
        if (lane == 0) base_s = before;
    }
    __syncthreads();
    int rank = base_s + __popc(m & ((1u << lane) - 1u));
    for (int w = 0; w < warp; ++w) rank += warp_count[w];
    int* rows = run_row + (size_t)layer * (n_rows + 1);
    if (start) rows[rank] = j;
    if (blockIdx.x == tiles - 1 && threadIdx.x == 0)
    {
        int total = base_s;
        for (int w = 0; w < kRunTile / 32; ++w) total += warp_count[w];
        n_runs[layer] = total;
        rows[total] = n_rows;
    }
}

// K3n.  The four gathered sums of every run (the quantities ped_run_sums defines).  The large one
// is the sum, at the run's k[s] point, of the earlier rows whose windows cover it: the lines of
// the 2*cut+1 cells below the run.  Consecutive runs need nearly the same rows, so a warp takes
// kNodeRuns consecutive runs: every lane loads a row ONCE and evaluates it at the k[s] points of
// all of them (the run descriptors are uniform reads from shared memory), then the lanes'
// partial sums are reduced per run.  (A warp per run, the first form of this kernel, read every
// row through L2 once per run that needs it, 2*cut+1 = 51 times: 150 GB per call on a
// million-line list, L2-bound.)  Row j goes to lane j mod 32 whatever the tile holds, so a run's
// sums do not depend on which other runs the call covers (spectral bands stop at different rows).
// The small sums -- the run's own lines at its two points, and earlier rows at k[e], which only
// out-of-order cells produce -- are taken by the eight runs side by side, four lanes each.
// grid = (blocks, layers), the warps of a layer striding over its tiles of kNodeRuns runs.
// (The lane layout of the second half fixes kNodeRuns at 8; 16 runs per warp was measured slower
// with the first half alone.)
#ifndef LBL_NODE_RUNS
#define LBL_NODE_RUNS 8
#endif
// (Resident blocks of 256 threads per SM: 2 / 3 / 4 = 128 / 80 / 64 registers.  The kernel waits on
// loads more than on anything else: 3 and 4 both take 0.8 ms off the benchmark step, 3 spills less;
// 16 runs per warp was slower at either.)
#ifndef LBL_NODES_RESIDENT
#define LBL_NODES_RESIDENT 3
#endif
constexpr int kNodeRuns = LBL_NODE_RUNS;

struct NodeRun
{
    int row_lo;       // rows before this one are "earlier" (0: run absent or skipped)
    int first_bin;    // bins [first_bin, first_bin + 2*cut+2) cover k[s]
    int i_s;
    int pad;
    double v_s;
};

__global__ void __launch_bounds__(256, LBL_NODES_RESIDENT)
ped_nodes_kernel(const PedRunArgs a)
{
    __shared__ NodeRun s_runs[8][kNodeRuns];
    const GridSpec& g = a.grid;
    const int layer = blockIdx.y;
    const int lane = threadIdx.x & 31;
    NodeRun* tile_runs = s_runs[threadIdx.x >> 5];
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int n_runs = a.n_runs[layer];
    const int* rows = a.run_row + (size_t)layer * (a.n_rows + 1);
    const size_t off = (size_t)layer * a.lines.n;
    const unsigned ns = 2 * g.cut_off + 2;
    for (int tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; tile * kNodeRuns < n_runs; tile += warps)
    {
        // lanes 0..kNodeRuns-1 describe the runs of the tile
        const int r = tile * kNodeRuns + lane;
        const bool have = lane < kNodeRuns && r < n_runs;
        int row_lo = 0, row_hi = 0, cb = 0;
        PedPoints pp;
        pp.skip = true;
        pp.i_s = pp.i_e = pp.bs = pp.be = pp.ns = pp.ne = 0;
        if (have)
        {
            row_lo = rows[r];
            row_hi = rows[r + 1];
            cb = LBL_LDG(reinterpret_cast<const int4*>(a.rec.chk + off + row_lo)).x;
            pp = ped_points(cb, g);
        }
        const bool live = have && !pp.skip;
        int j_first = live ? ped_first_covering_row(a, layer, pp.bs) : 0x7fffffff;
        int j_end = live ? row_lo : 0;
        if (lane < kNodeRuns)
        {
            NodeRun nr;
            nr.row_lo = live ? row_lo : 0;
            nr.first_bin = pp.bs;
            nr.i_s = pp.i_s;
            nr.pad = 0;
            nr.v_s = grid_point(g.v0, g.dv, pp.i_s);
            tile_runs[lane] = nr;
        }
        // what all runs of the tile have in common (the fast path below): rows before the first
        // run, in the bins every run's k[s] range holds, with no k[s] point in their near zone
        int row_lo_min = live ? row_lo : 0x7fffffff;
        int bs_max = live ? pp.bs : -0x40000000, bs_min = live ? pp.bs : 0x40000000;
        int is_min = live ? pp.i_s : 0x7fffffff, is_max = live ? pp.i_s : -1;
        const bool all_live = __ballot_sync(0xffffffffu, live) == (1u << kNodeRuns) - 1u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            j_first = min(j_first, __shfl_xor_sync(0xffffffffu, j_first, o));
            j_end = max(j_end, __shfl_xor_sync(0xffffffffu, j_end, o));
            row_lo_min = min(row_lo_min, __shfl_xor_sync(0xffffffffu, row_lo_min, o));
            bs_max = max(bs_max, __shfl_xor_sync(0xffffffffu, bs_max, o));
            bs_min = min(bs_min, __shfl_xor_sync(0xffffffffu, bs_min, o));
            is_min = min(is_min, __shfl_xor_sync(0xffffffffu, is_min, o));
            is_max = max(is_max, __shfl_xor_sync(0xffffffffu, is_max, o));
        }
        __syncwarp();
        LBL_CHECK(j_end <= a.lines.n && (j_first >= 0 || j_end == 0));
        double before_s[kNodeRuns];
#pragma unroll
        for (int q = 0; q < kNodeRuns; ++q) before_s[q] = 0.;
        for (int base = j_first & ~31; base < j_end; base += 32)
        {
            const int j = base + lane;
            const bool in = j < j_end;
            int4 ck = make_int4(-0x40000000, 0, -1, 0);
            double2 l = make_double2(0., 0.);
            double c = 1.;
            if (in)
            {
                ck = LBL_LDG(reinterpret_cast<const int4*>(a.rec.chk + off + j));
                l = LBL_LDG(reinterpret_cast<const double2*>(a.rec.ab + off + j));
                c = LBL_LDG(a.rec.cc + off + j);
            }
            const int bin = ck.x + g.cut_off + 1;
            // Three quarters of the chunks lie where every row counts for every run of the tile and
            // no run's point is near a row's centre: no test per (row, run) then, the term goes
            // straight into the sum.  (Both paths add a term by the same fused operation, so a
            // run's sum does not depend on which path its tile's chunks took.)
            const bool plain = in && j < row_lo_min && bin >= bs_max && (unsigned)(bin - bs_min) < ns &&
                               !(ck.y <= is_max && ck.z >= is_min);
            if (all_live && __all_sync(0xffffffffu, plain))
            {
#pragma unroll
                for (int q = 0; q < kNodeRuns; ++q)
                {
                    before_s[q] = far_term(tile_runs[q].v_s, l.x, l.y, c, before_s[q]);
                }
                continue;
            }
#pragma unroll
            for (int q = 0; q < kNodeRuns; ++q)
            {
                const NodeRun nr = tile_runs[q];
                const bool cover = j < nr.row_lo && (unsigned)(bin - nr.first_bin) < ns;
                double t = far_term(nr.v_s, l.x, l.y, c, before_s[q]);
                if (cover && nr.i_s >= ck.y && nr.i_s <= ck.z)
                {
                    // inside the line's near zone the summation kernels hold the full profile
                    const LineGen gen = a.rec.gen[off + j];
                    t = before_s[q] +
                        voigt_general_call(nr.v_s, gen.nu, gen.repwid, gen.y, gen.cof, gen.xlim0, gen.xlim1);
                }
                before_s[q] = cover ? t : before_s[q];
            }
        }
        // The eight partial sums of every lane -> one total per run, landing in the four lanes
        // lane >> 2 == run: each exchange halves the values a lane still carries (7 exchanges
        // instead of 8 x 5), the last two add up the four lanes of a group.
        static_assert(kNodeRuns == 8, "the reduction below pairs 8 runs with 8 groups of 4 lanes");
        {
            const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0, up4 = (lane & 4) != 0;
            double four[4], two[2];
#pragma unroll
            for (int i = 0; i < 4; ++i)
            {
                const double keep = up16 ? before_s[i + 4] : before_s[i];
                const double give = up16 ? before_s[i] : before_s[i + 4];
                four[i] = keep + __shfl_xor_sync(0xffffffffu, give, 16);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
            {
                const double keep = up8 ? four[i + 2] : four[i];
                const double give = up8 ? four[i] : four[i + 2];
                two[i] = keep + __shfl_xor_sync(0xffffffffu, give, 8);
            }
            const double keep = up4 ? two[1] : two[0];
            const double give = up4 ? two[0] : two[1];
            double one = keep + __shfl_xor_sync(0xffffffffu, give, 4);
            one += __shfl_xor_sync(0xffffffffu, one, 2);
            one += __shfl_xor_sync(0xffffffffu, one, 1);
            before_s[0] = one;      // run (lane >> 2): bit 4 chose runs 4-7, bit 3 the upper pair, bit 2 the odd one
        }
        // The small sums, the eight runs side by side: group q = lanes 4q .. 4q+3 takes run q (its
        // own rows at its two points; earlier rows at k[e]), so the eight chains of dependent
        // loads -- two range searches and a row loop each -- overlap instead of following one
        // another.
        const int q = lane >> 2, sub = lane & 3;
        const int q_lo = __shfl_sync(0xffffffffu, row_lo, q);
        const int q_hi = __shfl_sync(0xffffffffu, row_hi, q);
        const int q_cb = __shfl_sync(0xffffffffu, cb, q);
        const bool q_have = __shfl_sync(0xffffffffu, (int)have, q) != 0;
        const bool q_live = __shfl_sync(0xffffffffu, (int)live, q) != 0;
        PedPoints qp = ped_points(q_cb, g);
        double sums[4] = {0., 0., before_s[0], 0.};
        if (q_live)
        {
            ped_sum_own(a, off, q_lo, q_hi, qp, sub, 4, sums[0], sums[1]);
            sums[3] = ped_sum_before(a, layer, off, q_lo, qp.be, qp.ne, qp.i_e, sub, 4);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            if (k != 2)
            {
                sums[k] += __shfl_xor_sync(0xffffffffu, sums[k], 2);
                sums[k] += __shfl_xor_sync(0xffffffffu, sums[k], 1);
            }
        }
        if (q_have && sub == 0)
        {
            // with the sums, what the chain needs of the run's window (ped_points): its own bin
            // and the first bin of the two ranges that cover k[s] and k[e], packed into one int4
            const size_t o = (size_t)layer * a.n_rows + tile * kNodeRuns + q;
            int4 w;
            w.x = qp.skip ? -1 : q_cb + g.cut_off + 1;      // own bin (-1: not processed)
            w.y = qp.bs;
            w.z = qp.be;
            w.w = qp.ne;
            reinterpret_cast<int4*>(a.run_cb)[o] = w;
            double2* dst = reinterpret_cast<double2*>(a.run_sums + 4 * o);
            dst[0] = make_double2(sums[0], sums[1]);
            dst[1] = make_double2(sums[2], sums[3]);
        }
        __syncwarp();
    }
}

// K3c.  The chain over the runs of one layer: one warp per layer; the pedestal bins live in
// shared memory (or, for very wide grids, in the global pedbin array itself).  Lane i holds run
// r0 + i; the longest regular prefix of those 32 runs (ped_run_regular; all of them, nearly
// always) is resolved at once by the (min,+) scan described in lbl_threads.cuh -- a prefix sum
// and a prefix minimum across the lanes -- and an irregular run by the sequential step
// (ped_chain_run with explicit sums over the bins).
// grid = layers, block = 32, dynamic shared memory = nb doubles (0: bins in global memory).
__global__ void __launch_bounds__(32)
ped_chain_runs_kernel(const PedRunArgs a, int bins_in_smem, int use_scan)
{
    extern __shared__ double smem_bins[];
    const GridSpec& g = a.grid;
    const int layer = blockIdx.x;
    const int lane = threadIdx.x;
    const int nb = g.ncell + 2 * g.cut_off + 2;
    double* out_bins = a.pedbin + (size_t)layer * nb;
    double* bins = bins_in_smem ? smem_bins : out_bins;
    for (int b = lane; b < nb; b += 32) bins[b] = 0.;
    __syncwarp();
    const int n_runs = a.n_runs[layer];
    const int4* wins = reinterpret_cast<const int4*>(a.run_cb) + (size_t)layer * a.n_rows;
    const double2* sums = reinterpret_cast<const double2*>(a.run_sums + 4 * (size_t)layer * a.n_rows);
    const int ns = 2 * g.cut_off + 2;
    int top = -1;     // highest bin written so far: ranges above it sum to zero
    int r0 = 0;
    while (r0 < n_runs)
    {
        const int r = r0 + lane;
        PedRunInfo me;
        me.bin = -1;
        me.bs = me.be = me.ne = 0;
        me.sums[0] = me.sums[1] = me.sums[2] = me.sums[3] = 0.;
        if (r < n_runs)
        {
            const int4 w = wins[r];
            const double2 f = sums[2 * r], k = sums[2 * r + 1];
            me.bin = w.x; me.bs = w.y; me.be = w.z; me.ne = w.w;
            me.sums[0] = f.x; me.sums[1] = f.y; me.sums[2] = k.x; me.sums[3] = k.y;
        }
        const int first_bin = __shfl_sync(0xffffffffu, me.bin, 0);
        int prev_top = __shfl_up_sync(0xffffffffu, me.bin, 1);
        if (lane == 0) prev_top = top;
        const bool regular = use_scan && r < n_runs && ped_run_regular(me, first_bin, prev_top);
        const unsigned mask = __ballot_sync(0xffffffffu, regular);
        const int len = (mask == 0xffffffffu) ? 32 : __ffs(~mask) - 1;
        if (len >= 2)
        {
            const bool in = lane < len;
            const double alpha = in ? me.sums[1] + me.sums[3] : 0.;
            const double h = in ? (me.sums[0] + me.sums[2]) - ped_prior_window(bins, me.bs, first_bin)
                                : 1.0e300;
            double prefix = alpha;          // S_i = alpha_0 + ... + alpha_i
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const double t = __shfl_up_sync(0xffffffffu, prefix, o);
                if (lane >= o) prefix += t;
            }
            double low = h - prefix;        // min over k <= i of (h_k - S_k)
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const double t = __shfl_up_sync(0xffffffffu, low, o);
                if (lane >= o) low = fmin(low, t);
            }
            const double q = prefix + fmin(low, 0.);
            double q_prev = __shfl_up_sync(0xffffffffu, q, 1);
            if (lane == 0) q_prev = 0.;
            LBL_CHECK(!in || (me.bin >= 0 && me.bin < nb));
            if (in) bins[me.bin] += fmin(alpha, h - q_prev);
            top = __shfl_sync(0xffffffffu, me.bin, len - 1);
            r0 += len;
            __syncwarp();
            continue;
        }
        // the sequential step for run r0 (lane 0 holds it)
        r0 += 1;
        const int bin = first_bin;
        if (bin < 0)
        {
            continue;   // the reference does not process these lines on this grid
        }
        const int bs = __shfl_sync(0xffffffffu, me.bs, 0);
        const int be = __shfl_sync(0xffffffffu, me.be, 0);
        const int ne = __shfl_sync(0xffffffffu, me.ne, 0);
        double s4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) s4[q] = __shfl_sync(0xffffffffu, me.sums[q], 0);
        double ps = 0., pe = 0.;
        LBL_CHECK(bs >= 0 && bs + ns <= nb);
        for (int k = lane; k < ns; k += 32) ps += bins[bs + k];
        const bool e_side = be <= top;    // else every bin of the range is still zero
        if (e_side)
        {
            LBL_CHECK(be >= 0 && be + ne <= nb);
            for (int k = lane; k < ne; k += 32) pe += bins[be + k];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            ps += __shfl_xor_sync(0xffffffffu, ps, o);
            if (e_side) pe += __shfl_xor_sync(0xffffffffu, pe, o);
        }
        const double pedestal = ped_chain_run(s4, ps, pe);
        LBL_CHECK(bin >= 0 && bin < nb);
        if (lane == 0) bins[bin] += pedestal;
        top = max(top, bin);
        __syncwarp();
    }
    if (bins_in_smem)
    {
        for (int b = lane; b < nb; b += 32) out_bins[b] = bins[b];
    }
}

// K4a.  One thread per (layer, cell): the pedestal seen by the cell's points.
__global__ void pedestal_cells_kernel(const double* __restrict__ pedbin, GridSpec g,
                                      int n_layers, double* __restrict__ corr)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_layers * g.ncell)
    {
        return;
    }
    const int layer = idx / g.ncell;
    const int cell = idx - layer * g.ncell;
    const int nb = g.ncell + 2 * g.cut_off + 2;
    pedestal_cell(pedbin + (size_t)layer * nb, cell, g.cut_off, corr + 2 * (size_t)idx);
}

// K4b.  k[layer][i] -= pedestal(cell(i), i is the cell's first point), over the band's points.
// With an accumulator (the device-side gas sum, spectroscopy.py:181-191,225-234) the corrected
// value is not stored but added, scaled, into acc[layer][i - band start]:
//   acc += scale[layer] * (k - pedestal).
template <bool kPedestal, bool kMix>
__global__ void apply_kernel(double* __restrict__ out, const double* __restrict__ corr, GridSpec g,
                             int n_layers, double* __restrict__ acc, const double* __restrict__ scale)
{
    const int p_lo = band_first_point(g);
    const int width = band_end_point(g) - p_lo;
    const size_t total = (size_t)n_layers * width;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x)
    {
        const int layer = (int)(idx / width);
        const int i = p_lo + (int)(idx - (size_t)layer * width);
        const size_t o = (size_t)layer * g.n + i;
        double k = out[o];
        if (kPedestal)
        {
            const int cell = i / g.n_per_v;
            const int r = i - cell * g.n_per_v;
            k -= corr[2 * ((size_t)layer * g.ncell + cell) + (r == 0 ? 1 : 0)];
        }
        if (kMix)
        {
            acc[idx] = fma(scale[layer], k, acc[idx]);
        }
        else
        {
            out[o] = k;
        }
    }
}

// Gas-sum epilogue: acc[layer][i] += scale[layer] * k[layer][i]  (spectroscopy.py:181-191).
__global__ void mix_add_kernel(double* __restrict__ acc, const double* __restrict__ k,
                               const double* __restrict__ scale, int n, int n_layers)
{
    const size_t total = (size_t)n_layers * n;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x)
    {
        acc[idx] = fma(scale[idx / n], k[idx], acc[idx]);
    }
}

// FP64 peak probe: 8 independent DFMA chains per thread, nothing else in the loop.
__global__ void __launch_bounds__(256) dfma_probe_kernel(double* out, int iters, double a, double b)
{
    double x[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) x[c] = threadIdx.x * 1e-3 + c;
    for (int i = 0; i < iters; ++i)
    {
#pragma unroll
        for (int c = 0; c < 8; ++c) x[c] = fma(x[c], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) s += x[c];
    if (s == 12345.678) out[0] = s;
}

}  // namespace lbl
