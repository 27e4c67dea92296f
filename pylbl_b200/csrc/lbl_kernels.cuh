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

struct WarpSync
{
    __device__ __forceinline__ void operator()() const { __syncwarp(); }
};

// K3.  grid = layers, block = 32 (one warp walks the line list of one layer in DB order).
// `scratch` == nullptr: nodes live in dynamic shared memory; otherwise in global memory
// ([layer][ncell+1]) for grids too wide for shared memory.
__global__ void __launch_bounds__(32)
pedestal_kernel(const PedArgs a, double* scratch)
{
    extern __shared__ double smem_nodes[];
    double* nodes = scratch ? scratch + (size_t)blockIdx.x * (a.grid.ncell + 1) : smem_nodes;
    pedestal_layer(a, blockIdx.x, threadIdx.x, 32, nodes, WarpSync());
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

}  // namespace lbl
