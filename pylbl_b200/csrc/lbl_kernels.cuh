// lbl_kernels.cuh -- sm_100a kernels of the line-by-line path.
//
//   K1 scale_kernel          per-(layer, line) scaling      replaces spectra.c:17-45
//   K2 sum_kernel<P>         gather Voigt summation         replaces spectra.c:48-65 + voigt.c
//   K3 pedestal_kernel       pedestal recurrence            replaces spectra.c:66-78
//   K4 pedestal_cells/apply  pedestal correction per point
//
// None of these uses atomics on the spectrum: every output point is owned by one thread.
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
    if (j < ln.n)
    {
        w = scale_thread(ln, tips, layers, g, rec, layer, j);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        w += __shfl_down_sync(0xffffffffu, w, o);
    }
    if ((threadIdx.x & 31) == 0 && w != 0)
    {
        atomicAdd(evals + layer, (unsigned long long)w);
    }
}

// K2.  grid = (ceil(n/(P*128)), layers), block = 128 threads = 4 independent warps.
template <int P>
__global__ void __launch_bounds__(kSumBlock)
sum_kernel(const SumArgs a)
{
    sum_thread<P>(a, blockIdx.y, blockIdx.x * blockDim.x + threadIdx.x);
}

// K2b.  One warp per tile of T points x 32/T layers; grid = (ceil(tiles/4), ceil(layers/(32/T))).
template <int T>
__global__ void __launch_bounds__(128)
fixup_kernel(const SumArgs a)
{
    const int tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (tile * T >= a.grid.n)
    {
        return;
    }
    fixup_thread<T>(a, tile, blockIdx.y, threadIdx.x & 31);
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

// K3a.  terms[layer][row r][slot t], wpad slots per row (zero beyond the window).
__global__ void __launch_bounds__(256)
pedestal_terms_kernel(const PedArgs a, int wpad, double* __restrict__ terms)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int layer = blockIdx.y;
    if (idx >= (long long)a.lines.n * wpad)
    {
        return;
    }
    const int r = (int)(idx / wpad);
    const int t = (int)(idx - (long long)r * wpad);
    terms[(size_t)layer * a.lines.n * wpad + idx] = pedestal_term(a, layer, r, t);
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

constexpr int kPedTile = 32;    // lines per staged tile
constexpr int kPedStages = 4;   // cp.async ring depth

// K3b.  grid = layers, block = one warp.  The sequential chain of the recurrence: per line
// three dependent FP64 operations on register-resident nodes; the per-line terms (K3a) and
// window cells stream in through a 4-stage cp.async ring so that no global-memory latency
// sits on the chain.  K = slots per lane (32*K >= 2*cut+3), wpad = 32*K.
template <int K>
__global__ void __launch_bounds__(32)
pedestal_chain_kernel(const PedArgs a, const double* __restrict__ terms, double* scratch)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const GridSpec& g = a.grid;
    constexpr int wpad = 32 * K;
    const int layer = blockIdx.x;
    const int lane = threadIdx.x;
    const int n = a.lines.n;
    double* ring = reinterpret_cast<double*>(smem_raw);
    int4* hdr = reinterpret_cast<int4*>(ring + kPedStages * kPedTile * wpad);
    double* nodes = scratch ? scratch + (size_t)layer * (g.ncell + 1)
                            : reinterpret_cast<double*>(hdr + kPedStages * kPedTile);
    const int nb = g.ncell + 2 * g.cut_off + 2;
    double* bins = a.pedbin + (size_t)layer * nb;
    for (int c = lane; c <= g.ncell; c += 32) nodes[c] = 0.;
    for (int b = lane; b < nb; b += 32) bins[b] = 0.;
    __syncwarp();

    const double* src = terms + (size_t)layer * n * wpad;
    const LineChk* chk = a.rec.chk + (size_t)layer * n;
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
            for (int o = lane * 16; o < bytes; o += 32 * 16)
            {
                cp_async16(sdst + o, gsrc + o);
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
    for (int t = 0; t < kPedStages - 1; ++t) issue(t);

    PedLane<K> st;
    ped_lane_init(st);
    for (int t = 0; t < ntiles; ++t)
    {
        issue(t + kPedStages - 1);
        cp_async_wait<kPedStages - 1>();
        __syncwarp();
        const int stage = t % kPedStages;
        const int first = t * kPedTile;
        const int cnt = (n - first < kPedTile) ? n - first : kPedTile;
        const double* rows = ring + (size_t)stage * kPedTile * wpad;
        const int4* cells = hdr + stage * kPedTile;
        for (int l = 0; l < cnt; ++l)
        {
            const int cb = cells[l].x;
            if (!st.have || cb != st.cb)
            {
                const PedWindow w = ped_window(cb, g);
                if (w.skip)
                {
                    continue;
                }
                ped_lane_flush(st, g, lane, nodes, bins);
                __syncwarp();
                ped_lane_reload(st, g, lane, cb, w, nodes);
                __syncwarp();
            }
            ped_lane_line(st, lane, rows + (size_t)l * wpad);
        }
        __syncwarp();
    }
    ped_lane_flush(st, g, lane, nodes, bins);
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

// K4b.  k[layer][i] -= pedestal(cell(i), i is the cell's first point).
__global__ void pedestal_apply_kernel(double* __restrict__ out, const double* __restrict__ corr,
                                      GridSpec g, int n_layers)
{
    const size_t total = (size_t)n_layers * g.n;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x)
    {
        const int layer = (int)(idx / g.n);
        const int i = (int)(idx - (size_t)layer * g.n);
        const int cell = i / g.n_per_v;
        const int r = i - cell * g.n_per_v;
        out[idx] -= corr[2 * ((size_t)layer * g.ncell + cell) + (r == 0 ? 1 : 0)];
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
