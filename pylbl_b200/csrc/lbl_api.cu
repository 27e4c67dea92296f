// lbl_api.cu -- C ABI (include/pylbl_b200.h), device packing and launch orchestration.
//
// Reference behaviour replaced here: pyLBL/c_lib/absorption.c:19-99 (driver of one
// (gas, layer) call) -- batched over layers, with the database read hoisted into
// lbl_gas_open().  Citations are relative to /root/reference/pyLBL/c_lib/.
#include <cuda_runtime.h>
#include <sys/stat.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/pylbl_b200.h"
#include "lbl_cheb.h"
#include "lbl_bands.h"
#include "lbl_db.h"
#include "lbl_kernels.cuh"

namespace lbl
{
namespace
{

thread_local std::string g_last_error;
int g_chunk_layers = 0;

int fail(const std::string& msg)
{
    g_last_error = msg;
    fprintf(stderr, "%s\n", msg.c_str());
    return 1;
}

#define LBL_CUDA(call)                                                                     \
    do                                                                                     \
    {                                                                                      \
        cudaError_t err__ = (call);                                                        \
        if (err__ != cudaSuccess)                                                          \
        {                                                                                  \
            return fail(std::string("Error: CUDA: ") + cudaGetErrorString(err__) + " at " + \
                        #call);                                                            \
        }                                                                                  \
    } while (0)

// Grow-only device buffer.
struct DevBuf
{
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

// A packed, nu-sorted line list on the device.
struct DeviceLines
{
    int n = 0;
    DevBuf nu, sw, gamma_air, gamma_self, n_air, elower, delta_air, mass, iso, db_to_sorted;
    bool has_perm = false;
    size_t bytes = 0;
    void release()
    {
        for (DevBuf* b : {&nu, &sw, &gamma_air, &gamma_self, &n_air, &elower, &delta_air, &mass,
                          &iso, &db_to_sorted})
        {
            b->release();
        }
    }
    LinesView view(int n_active) const
    {
        LinesView v;
        v.n = n_active;
        v.nu = nu.as<double>();
        v.sw = sw.as<double>();
        v.gamma_air = gamma_air.as<double>();
        v.gamma_self = gamma_self.as<double>();
        v.n_air = n_air.as<double>();
        v.elower = elower.as<double>();
        v.delta_air = delta_air.as<double>();
        v.mass = mass.as<double>();
        v.iso = iso.as<int>();
        v.db_to_sorted = has_perm ? db_to_sorted.as<int>() : nullptr;
        return v;
    }
};

struct Plan  // which lines a given (v0, vn, cut_off) sees, and where they are on the device
{
    int v0 = 0, vn = 0, cut_off = 0;
    int n_active = 0;
    bool valid = false;
    DeviceLines own;               // used only when the database rows are not nu-sorted
    std::vector<int> sorted_to_db; // host copy of the permutation (empty = identity)
    DevBuf cell_first;             // per-wavenumber index of the sorted active lines (LinesView)
    int cell_w0 = 0, cell_n = 0;
};


}  // namespace
}  // namespace lbl

namespace lbl
{
namespace
{
struct DeviceStreams;
}
}  // namespace lbl

using namespace lbl;

struct lbl_gas
{
    int device = 0;
    std::string database, formula;
    MoleculeData mol;
    DeviceLines base;  // all rows, valid when mol.sorted
    DevBuf tips_t, tips_q;
    Plan plan;
    size_t open_h2d = 0;
    bool open_h2d_reported = false;

    // the device's shared streams (DeviceStreams); s_compute = early
    cudaStream_t s_compute = nullptr, s_late = nullptr, s_copy = nullptr, s_main = nullptr;
    DeviceStreams* streams = nullptr;
    size_t group_budget = (size_t)4 << 30;   // bytes per layer group, see lbl_gas_submit
    int copy_groups = 0;                     // lbl_gas_set_copy_groups (0 = automatic)
    DevBuf rec_ab, rec_cc, rec_chk, rec_gen, layers_dev, evals_dev, pedbin, pedcorr, pednodes,
        pedterms, ped_tiles, ped_run_row, ped_n_runs, ped_run_cb, ped_run_sums, rec_f32, amp_max, cell_keys, cheb_nodes, cheb_weights, cheb_nodes16, cheb_weights16, cheb_nodes8, cheb_weights8,
        executed_dev;
    int cheb_npv = 0;
    unsigned long long* executed_host = nullptr;  // pinned
    DevBuf out[2], mix_scale_dev;
    double* mix_scale_host = nullptr;  // pinned
    size_t mix_scale_host_cap = 0;
    LayerIn* layers_host = nullptr;  // pinned
    size_t layers_host_cap = 0;
    unsigned long long* evals_host = nullptr;  // pinned
    size_t evals_host_cap = 0;

    // per-call state
    std::vector<cudaEvent_t> ev_pool;
    int ev_used = 0;
    struct ChunkEvents
    {
        cudaEvent_t k1_begin, k1_end, ped_begin, ped_end, applied;
        bool pedestal;
    };
    struct SumEvents   // one per launch of the summation kernel + K2b
    {
        cudaEvent_t k2_begin, k2_end, k2b_end;
    };
    std::vector<SumEvents> sum_events;
    std::vector<ChunkEvents> chunk_events;
    cudaEvent_t ev_call_begin = nullptr, ev_call_end = nullptr, ev_compute_end = nullptr;
    cudaEvent_t ev_out_ready[2] = {nullptr, nullptr}, ev_out_free[2] = {nullptr, nullptr};
    bool out_busy[2] = {false, false};
    bool pending = false;
    bool copies_in_call = false;
    bool farfield_last = false;
    lbl_stats stats{};
    // what is resident after the last call (for windows/scaled/device_result)
    GridSpec last_grid{};
    int last_chunk_first = 0, last_chunk_layers = 0, last_slot = 0;
    std::vector<LayerIn> last_layers;
};

namespace lbl
{
namespace
{

int upload(DevBuf& b, const void* src, size_t bytes, cudaStream_t s, size_t& counter)
{
    LBL_CUDA(b.reserve(std::max<size_t>(bytes, 16)));
    if (bytes)
    {
        LBL_CUDA(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, s));
    }
    counter += bytes;
    return 0;
}

// Uploads rows order[0..n) of the molecule (order empty = identity).
int pack_lines(lbl_gas* g, DeviceLines& d, const std::vector<int>& order, int n)
{
    const MoleculeData& m = g->mol;
    auto gather_d = [&](const std::vector<double>& src, std::vector<double>& tmp) {
        if (order.empty()) return src.data();
        tmp.resize(n);
        for (int i = 0; i < n; ++i) tmp[i] = src[order[i]];
        return (const double*)tmp.data();
    };
    std::vector<double> tmp;
    size_t bytes = 0;
    const std::vector<double>* cols[8] = {&m.nu, &m.sw, &m.gamma_air, &m.gamma_self,
                                          &m.n_air, &m.elower, &m.delta_air, &m.mass};
    DevBuf* bufs[8] = {&d.nu, &d.sw, &d.gamma_air, &d.gamma_self,
                       &d.n_air, &d.elower, &d.delta_air, &d.mass};
    for (int c = 0; c < 8; ++c)
    {
        const double* src = gather_d(*cols[c], tmp);
        if (upload(*bufs[c], src, sizeof(double) * n, g->s_compute, bytes)) return 1;
        LBL_CUDA(cudaStreamSynchronize(g->s_compute));  // tmp is reused
    }
    std::vector<int> iso0(n);
    for (int i = 0; i < n; ++i) iso0[i] = m.iso[order.empty() ? i : order[i]] - 1;
    if (upload(d.iso, iso0.data(), sizeof(int) * n, g->s_compute, bytes)) return 1;
    d.has_perm = !order.empty();
    if (d.has_perm)
    {
        std::vector<int> inv(n);
        for (int i = 0; i < n; ++i) inv[order[i]] = i;
        if (upload(d.db_to_sorted, inv.data(), sizeof(int) * n, g->s_compute, bytes)) return 1;
    }
    LBL_CUDA(cudaStreamSynchronize(g->s_compute));
    d.n = n;
    d.bytes = bytes;
    g->open_h2d += bytes;
    return 0;
}

// The reference walks rows in database order and STOPS at the first row outside
// [v0-(cut+1), vn+cut+1] (absorption.c:80-83); rows after it are never seen.
int active_prefix(const MoleculeData& m, int v0, int vn, int cut_off)
{
    const int n = (int)m.nu.size();
    const double hi = (double)(vn + cut_off + 1);
    const double lo = (double)(v0 - (cut_off + 1));
    if (m.sorted)
    {
        if (n == 0 || m.nu[0] < lo) return 0;
        return (int)(std::upper_bound(m.nu.begin(), m.nu.end(), hi) - m.nu.begin());
    }
    int r = 0;
    for (; r < n; ++r)
    {
        if (m.nu[r] > hi || m.nu[r] < lo) break;
    }
    return r;
}

int make_plan(lbl_gas* g, int v0, int vn, int cut_off)
{
    Plan& p = g->plan;
    if (p.valid && p.v0 == v0 && p.vn == vn && p.cut_off == cut_off) return 0;
    p.valid = false;
    p.v0 = v0;
    p.vn = vn;
    p.cut_off = cut_off;
    p.n_active = active_prefix(g->mol, v0, vn, cut_off);
    if (g->mol.has_tips)
    {
        // Only rows the reference would touch are held to this (it indexes the TIPS block
        // local_iso_id-1 without a check, spectra.c:41-42).
        for (int r = 0; r < p.n_active; ++r)
        {
            if (g->mol.iso[r] > g->mol.num_iso)
            {
                return fail("Error: line refers to an isotopologue without TIPS data.");
            }
        }
    }
    p.sorted_to_db.clear();
    if (!g->mol.sorted && p.n_active > 0)
    {
        p.sorted_to_db.resize(p.n_active);
        std::iota(p.sorted_to_db.begin(), p.sorted_to_db.end(), 0);
        const std::vector<double>& nu = g->mol.nu;
        std::stable_sort(p.sorted_to_db.begin(), p.sorted_to_db.end(),
                         [&](int a, int b) { return nu[a] < nu[b]; });
        if (pack_lines(g, p.own, p.sorted_to_db, p.n_active)) return 1;
    }
    // Per-wavenumber index (LinesView::cell_first).  Every active line lies within
    // [v0-(cut+1), vn+cut+1] (active_prefix), inside the table's [w0, w0+n-1].
    {
        const std::vector<double>& nu = g->mol.nu;
        p.cell_w0 = v0 - cut_off - 3;
        p.cell_n = (vn - v0) + 2 * cut_off + 7;
        std::vector<int> first((size_t)p.cell_n);
        int j = 0;
        for (int k = 0; k < p.cell_n; ++k)
        {
            const double w = (double)p.cell_w0 + (double)k;
            while (j < p.n_active && nu[p.sorted_to_db.empty() ? j : p.sorted_to_db[j]] < w) ++j;
            first[k] = j;
        }
        size_t bytes = 0;
        if (upload(p.cell_first, first.data(), sizeof(int) * first.size(), g->s_compute, bytes)) return 1;
        LBL_CUDA(cudaStreamSynchronize(g->s_compute));   // `first` goes out of scope
        g->open_h2d += bytes;
    }
    p.valid = true;
    return 0;
}

cudaEvent_t next_event(lbl_gas* g)
{
    if (g->ev_used == (int)g->ev_pool.size())
    {
        cudaEvent_t e;
        cudaEventCreate(&e);
        g->ev_pool.push_back(e);
    }
    return g->ev_pool[g->ev_used++];
}

int pick_points_per_thread(int n_per_v)
{
    // PYLBL_B200_POINTS overrides the choice (tuning experiments); it must divide n_per_v.
    if (const char* env = getenv("PYLBL_B200_POINTS"))
    {
        const int p = atoi(env);
        if ((p == 10 || p == 8 || p == 5 || p == 4 || p == 2 || p == 1) && n_per_v % p == 0) return p;
    }
    const int candidates[] = {5, 4, 8, 10, 2, 1};
    for (int p : candidates)
    {
        if (n_per_v % p == 0) return p;
    }
    return 1;
}

// Threads of a warp that share a layer: as many as keep the warp within about two
// integer-wavenumber cells (a wider warp wastes work on window edges), but no more layers per
// warp than there are layers.
int pick_threads_per_layer(int n_per_v, int P, int n_layers)
{
    int tpw = 32;
    if (const char* env = getenv("PYLBL_B200_TPW"))   // tuning experiments: 1, 2, 4, 8, 16 or 32
    {
        const int t = atoi(env);
        if (t >= 1 && t <= 32 && (t & (t - 1)) == 0)
        {
            tpw = t;
            while (tpw < 32 && (32 / tpw) > n_layers) tpw <<= 1;
            return tpw;
        }
    }
    while (tpw > 1 && tpw * P > 2 * n_per_v) tpw >>= 1;
    while (tpw < 32 && (32 / tpw) > n_layers) tpw <<= 1;
    return tpw;
}

template <int P>
void launch_sum(SumArgs a, int n_layers, bool fp32, cudaStream_t s)
{
    // one warp per tile of tpw*P points; a band launches the tiles (of the whole grid) it meets
    const int tile_points = a.tpw * P;
    a.tile0 = band_first_point(a.grid) / tile_points;
    const int tiles = (band_end_point(a.grid) + tile_points - 1) / tile_points - a.tile0;
    const int lp = 32 / a.tpw;
    dim3 grid((tiles + kSumBlock / 32 - 1) / (kSumBlock / 32), (n_layers + lp - 1) / lp);
    if (fp32)
    {
        sum32_kernel<P><<<grid, kSumBlock, 0, s>>>(a);
    }
    else
    {
        sum_kernel<P><<<grid, kSumBlock, 0, s>>>(a);
    }
}

void launch_sum_dispatch(int P, const SumArgs& a, int n_layers, bool fp32, cudaStream_t s)
{
    switch (P)
    {
        case 10: launch_sum<10>(a, n_layers, fp32, s); break;
        case 8: launch_sum<8>(a, n_layers, fp32, s); break;
        case 5: launch_sum<5>(a, n_layers, fp32, s); break;
        case 4: launch_sum<4>(a, n_layers, fp32, s); break;
        case 2: launch_sum<2>(a, n_layers, fp32, s); break;
        default: launch_sum<1>(a, n_layers, fp32, s); break;
    }
}

// K2b tile width: about one near zone (~0.3 cm-1) of grid points, a power of two in [4, 32].
int pick_fixup_tile(int n_per_v)
{
    if (n_per_v >= 64) return 32;
    if (n_per_v >= 32) return 16;
    if (n_per_v >= 16) return 8;
    return 4;
}

// K2b, line-major form: valid when a point inside a line's near zone is always inside the
// line's window, i.e. cut_off >= reach + 1 cm-1 (reach: as near_candidates() computes it).
bool near_block_applies(const SumArgs& a, const std::vector<LayerIn>& layers, int first, int count)
{
    if (a.grid.n_per_v < 64) return false;
    if (const char* env = getenv("PYLBL_B200_NEARBLOCK")) { if (atoi(env) == 0) return false; }
    const double v_abs = std::max(std::fabs((double)a.grid.v0), std::fabs((double)a.grid.vn));
    for (int l = first; l < first + count; ++l)
    {
        const LayerIn& ly = layers[l];
        if (!(ly.kappa < 0.5)) return false;
        const double reach = (ly.kappa * v_abs / (1.0 - ly.kappa)) * (1.0 + 0x1p-20) + ly.slack +
                             3.0 * a.grid.dv;
        if (!((double)a.grid.cut_off >= reach + 1.0)) return false;
    }
    return true;
}

void launch_near_block(SumArgs a, int n_layers, cudaStream_t s)
{
    cudaFuncSetAttribute(near_block_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);   // 8 blocks x 28 KB per SM
    a.tile0 = band_first_point(a.grid) / kNbSpan;
    dim3 grid((band_end_point(a.grid) + kNbSpan - 1) / kNbSpan - a.tile0, n_layers);
    near_block_kernel<<<grid, 128, 0, s>>>(a);
}

void launch_fixup_dispatch(int T, SumArgs a, int n_layers, cudaStream_t s)
{
    a.tile0 = band_first_point(a.grid) / T;
    const int tiles = (band_end_point(a.grid) + T - 1) / T - a.tile0;
    const int lp = 32 / T;
    dim3 grid((tiles + 3) / 4, (n_layers + lp - 1) / lp);
    switch (T)
    {
        case 32:
            cudaFuncSetAttribute(fixup_kernel<32>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared);   // 10 blocks x 21 KB per SM
            fixup_kernel<32><<<grid, 128, 0, s>>>(a);
            break;
        case 16: fixup_kernel<16><<<grid, 128, 0, s>>>(a); break;
        case 8: fixup_kernel<8><<<grid, 128, 0, s>>>(a); break;
        default: fixup_kernel<4><<<grid, 128, 0, s>>>(a); break;
    }
}

template <int K>
cudaError_t launch_chain(const PedArgs& pa, double* terms, double* scratch, int n_layers,
                         size_t smem, cudaStream_t s)
{
    cudaError_t e = cudaSuccess;
    {
        const int tiles = (pa.n_rows + kPedTileRows - 1) / kPedTileRows;
        dim3 tg((tiles + 3) / 4, n_layers);
        pedestal_terms_kernel<K><<<tg, 128, 0, s>>>(pa, terms);
    }
    if (smem > 48 * 1024)
    {
        e = cudaFuncSetAttribute(pedestal_chain_kernel<K>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    pedestal_chain_kernel<K><<<n_layers, 32, smem, s>>>(pa, terms, scratch);
    return cudaGetLastError();
}

// All handles of a device share one small set of streams.  A GPU context has few hardware
// work queues (8 by default, CUDA_DEVICE_MAX_CONNECTIONS); with three private streams per
// handle, seven gases in flight alias their streams onto those queues, and an operation then
// waits for unrelated work that happens to share its queue (measured: the same step took
// 31 ms or 70 ms from one process to the next).  The shared set, in submission order on each:
//   early  scaling kernels (and the uploads) of every call: short, prioritised, never waits
//   side[] pedestal terms + chain, round-robin over kSideStreams streams: long, thin,
//          prioritised, so that several gases' chains run side by side
//   main   the summation kernels (K2/K2c + K2b), low priority, gas after gas: each gas
//          finishes, and starts copying out, while the next one computes
//   late   pedestal apply and the small statistics copies: waits for main and side
//   copy   the device-to-host copies of the spectra
constexpr int kSideStreams = 4;
struct DeviceStreams
{
    cudaStream_t main = nullptr, early = nullptr, late = nullptr, copy = nullptr;
    cudaStream_t side[kSideStreams] = {};
    unsigned next_side = 0;
    bool ready = false;
};
std::mutex g_streams_mu;
DeviceStreams g_streams[64];

int device_streams(int device, DeviceStreams** out)
{
    if (device < 0 || device >= 64) return fail("Error: device index out of range.");
    std::lock_guard<std::mutex> lock(g_streams_mu);
    DeviceStreams& d = g_streams[device];
    if (!d.ready)
    {
        int least = 0, greatest = 0;
        LBL_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        LBL_CUDA(cudaStreamCreateWithPriority(&d.main, cudaStreamNonBlocking, least));
        LBL_CUDA(cudaStreamCreateWithPriority(&d.early, cudaStreamNonBlocking, greatest));
        LBL_CUDA(cudaStreamCreateWithPriority(&d.late, cudaStreamNonBlocking, greatest));
        LBL_CUDA(cudaStreamCreateWithFlags(&d.copy, cudaStreamNonBlocking));
        for (int i = 0; i < kSideStreams; ++i)
        {
            LBL_CUDA(cudaStreamCreateWithPriority(&d.side[i], cudaStreamNonBlocking, greatest));
        }
        d.ready = true;
    }
    *out = &d;
    return 0;
}

cudaStream_t next_side_stream(DeviceStreams* d)
{
    std::lock_guard<std::mutex> lock(g_streams_mu);
    return d->side[d->next_side++ % kSideStreams];
}

// Interpolation tables of K2c for one grid resolution: for each of the three far fields
// (32, 16 and 8 nodes) the Chebyshev nodes of the first kind on the cell interval
// [0, (n_per_v-1)/n_per_v] and the node-values -> coefficients transform (lbl_cheb.h).
int ensure_cheb_tables(lbl_gas* g, int n_per_v)
{
    if (g->cheb_npv == n_per_v) return 0;
    struct Field { int n; DevBuf* nodes; DevBuf* transform; };
    const Field fields[3] = {{kNodes, &g->cheb_nodes, &g->cheb_weights},
                             {kNodes16, &g->cheb_nodes16, &g->cheb_weights16},
                             {kNodes8, &g->cheb_nodes8, &g->cheb_weights8}};
    size_t bytes = 0;
    for (const Field& f : fields)
    {
        std::vector<double> nodes, transform;
        build_cheb_nodes(f.n, n_per_v, nodes);
        build_cheb_transform(f.n, transform);
        if (upload(*f.nodes, nodes.data(), sizeof(double) * nodes.size(), g->s_compute, bytes)) return 1;
        if (upload(*f.transform, transform.data(), sizeof(double) * transform.size(), g->s_compute, bytes))
            return 1;
        LBL_CUDA(cudaStreamSynchronize(g->s_compute));   // the host vectors go out of scope
    }
    g->open_h2d += bytes;
    g->cheb_npv = n_per_v;
    return 0;
}

int set_device(lbl_gas* g)
{
    LBL_CUDA(cudaSetDevice(g->device));
    return 0;
}

}  // namespace
}  // namespace lbl

// =========================================================================================
extern "C" {

const char* lbl_last_error(void)
{
    return g_last_error.c_str();
}

int lbl_version(void)
{
    return 200;
}

int lbl_set_chunk_layers(int layers)
{
    g_chunk_layers = layers < 0 ? 0 : layers;
    return 0;
}

int lbl_device_count(int* count)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver)
    {
        // A machine without a usable GPU has zero devices; that is an answer, not a failure.
        cudaGetLastError();
        g_last_error = std::string("Error: CUDA: ") + cudaGetErrorString(e);
        *count = 0;
        return 0;
    }
    if (e != cudaSuccess)
    {
        *count = 0;
        return fail(std::string("Error: CUDA: ") + cudaGetErrorString(e));
    }
    *count = n;
    return 0;
}

namespace
{
struct DeviceTimer
{
    cudaStream_t stream = nullptr;
    cudaEvent_t begin = nullptr, end = nullptr;
};
std::mutex g_timer_mu;
std::map<int, DeviceTimer> g_timers;

int get_timer(int device, DeviceTimer** out)
{
    std::lock_guard<std::mutex> lock(g_timer_mu);
    DeviceTimer& t = g_timers[device];
    if (!t.stream)
    {
        LBL_CUDA(cudaSetDevice(device));
        LBL_CUDA(cudaStreamCreateWithFlags(&t.stream, cudaStreamNonBlocking));
        LBL_CUDA(cudaEventCreate(&t.begin));
        LBL_CUDA(cudaEventCreate(&t.end));
    }
    *out = &t;
    return 0;
}
}  // namespace

int lbl_timer_start(int device)
{
    DeviceTimer* t = nullptr;
    if (get_timer(device, &t)) return 1;
    LBL_CUDA(cudaSetDevice(device));
    LBL_CUDA(cudaEventRecord(t->begin, t->stream));
    return 0;
}

int lbl_timer_join(lbl_gas* g)
{
    if (!g) return fail("Error: null handle.");
    DeviceTimer* t = nullptr;
    if (get_timer(g->device, &t)) return 1;
    LBL_CUDA(cudaSetDevice(g->device));
    if (g->pending)
    {
        LBL_CUDA(cudaStreamWaitEvent(t->stream, g->ev_call_end, 0));
    }
    return 0;
}

int lbl_timer_stop(int device, float* ms)
{
    DeviceTimer* t = nullptr;
    if (get_timer(device, &t)) return 1;
    LBL_CUDA(cudaSetDevice(device));
    LBL_CUDA(cudaEventRecord(t->end, t->stream));
    LBL_CUDA(cudaEventSynchronize(t->end));
    LBL_CUDA(cudaEventElapsedTime(ms, t->begin, t->end));
    return 0;
}

int lbl_measure_fp64_peak(int device, double* tflops)
{
    *tflops = 0.;
    LBL_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    LBL_CUDA(cudaGetDeviceProperties(&prop, device));
    double* out = nullptr;
    LBL_CUDA(cudaMalloc(&out, 64));
    cudaEvent_t a, b;
    LBL_CUDA(cudaEventCreate(&a));
    LBL_CUDA(cudaEventCreate(&b));
    const int blocks = prop.multiProcessorCount * 8, iters = 20000;
    float best = 1e30f;
    for (int r = 0; r < 8; ++r)
    {
        LBL_CUDA(cudaEventRecord(a, 0));
        dfma_probe_kernel<<<blocks, 256>>>(out, iters, 1.0000001, 1e-9);
        LBL_CUDA(cudaEventRecord(b, 0));
        LBL_CUDA(cudaEventSynchronize(b));
        float ms = 0.f;
        LBL_CUDA(cudaEventElapsedTime(&ms, a, b));
        if (r >= 2 && ms < best) best = ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(out);
    *tflops = 2.0 * (double)blocks * 256 * 8 * iters / (best * 1e-3) / 1e12;
    return 0;
}

int lbl_host_alloc(size_t bytes, void** ptr)
{
    *ptr = nullptr;
    LBL_CUDA(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable));
    return 0;
}

int lbl_host_free(void* ptr)
{
    if (ptr) LBL_CUDA(cudaFreeHost(ptr));
    return 0;
}

// Size and modification time of a file (0, 0 when it cannot be examined).
static void file_stamp(const char* path, long long* size, long long* mtime)
{
    struct stat st;
    *size = *mtime = 0;
    if (stat(path, &st) == 0)
    {
        *size = (long long)st.st_size;
        *mtime = (long long)st.st_mtime;
    }
}

int lbl_pack_database(const char* database, const char* formula, const char* pack_path)
{
    if (!database || !formula || !pack_path) return fail("Error: null argument.");
    MoleculeData mol;
    std::string err;
    if (read_molecule(database, formula, mol, err)) return fail(err);
    long long size = 0, mtime = 0;
    file_stamp(database, &size, &mtime);
    if (write_pack(pack_path, formula, mol, size, mtime, err)) return fail(err);
    return 0;
}

int lbl_pack_info(const char* pack_path, char* formula, int formula_capacity, long long* n_lines,
                  int* num_iso, int* num_t, int* sorted, long long* source_size,
                  long long* source_mtime)
{
    if (!pack_path) return fail("Error: null argument.");
    MoleculeData unused;
    PackInfo info;
    std::string err;
    if (read_pack(pack_path, unused, info, true, err)) return fail(err);
    if (formula && formula_capacity > 0)
    {
        strncpy(formula, info.formula.c_str(), (size_t)formula_capacity - 1);
        formula[formula_capacity - 1] = 0;
    }
    if (n_lines) *n_lines = info.n_lines;
    if (num_iso) *num_iso = info.num_iso;
    if (num_t) *num_t = info.num_t;
    if (sorted) *sorted = info.sorted ? 1 : 0;
    if (source_size) *source_size = info.source_size;
    if (source_mtime) *source_mtime = info.source_mtime;
    return 0;
}

static int open_handle(std::unique_ptr<lbl_gas>& g, int device, lbl_gas** out);

int lbl_gas_open(const char* database, const char* formula, int device, lbl_gas** out)
{
    *out = nullptr;
    if (!database || !formula) return fail("Error: null argument.");
    std::unique_ptr<lbl_gas> g(new lbl_gas);
    g->database = database;
    g->formula = formula;
    std::string err;
    if (read_molecule(database, formula, g->mol, err)) return fail(err);
    return open_handle(g, device, out);
}

int lbl_gas_open_pack(const char* pack_path, int device, lbl_gas** out)
{
    *out = nullptr;
    if (!pack_path) return fail("Error: null argument.");
    std::unique_ptr<lbl_gas> g(new lbl_gas);
    PackInfo info;
    std::string err;
    if (read_pack(pack_path, g->mol, info, false, err)) return fail(err);
    g->database = pack_path;
    g->formula = info.formula;
    return open_handle(g, device, out);
}

// The device side of opening: streams, events, TIPS table and (for sorted rows) the lines.
static int open_handle(std::unique_ptr<lbl_gas>& g, int device, lbl_gas** out)
{
    int ndev = 0;
    if (lbl_device_count(&ndev)) return 1;
    if (ndev == 0)
    {
        return fail("Error: no CUDA device (" + g_last_error + "); this library has no CPU path.");
    }
    if (device < 0 || device >= ndev) return fail("Error: CUDA device index out of range.");
    g->device = device;
    if (set_device(g.get())) return 1;
    {
        size_t free_bytes = 0, total_bytes = 0;
        if (cudaMemGetInfo(&free_bytes, &total_bytes) == cudaSuccess)
        {
            // A share of what is free now: a column's gases each hold a handle on the device (and
            // each handle up to three such groups: records and two output slabs), and every handle
            // keeps what it has grown to.
            g->group_budget = std::min<size_t>(std::max<size_t>(free_bytes / 20, (size_t)2 << 30),
                                               (size_t)8 << 30);
        }
    }
    if (device_streams(device, &g->streams)) return 1;
    g->s_compute = g->streams->early;
    g->s_late = g->streams->late;
    g->s_copy = g->streams->copy;
    g->s_main = g->streams->main;
    LBL_CUDA(cudaEventCreate(&g->ev_call_begin));
    LBL_CUDA(cudaEventCreate(&g->ev_call_end));
    LBL_CUDA(cudaEventCreate(&g->ev_compute_end));
    for (int i = 0; i < 2; ++i)
    {
        LBL_CUDA(cudaEventCreateWithFlags(&g->ev_out_ready[i], cudaEventDisableTiming));
        LBL_CUDA(cudaEventCreateWithFlags(&g->ev_out_free[i], cudaEventDisableTiming));
    }
    if (g->mol.has_tips)
    {
        size_t bytes = 0;
        if (upload(g->tips_t, g->mol.tips_t.data(), sizeof(double) * g->mol.tips_t.size(),
                   g->s_compute, bytes)) return 1;
        if (upload(g->tips_q, g->mol.tips_q.data(), sizeof(double) * g->mol.tips_q.size(),
                   g->s_compute, bytes)) return 1;
        g->open_h2d += bytes;
        if (g->mol.sorted)
        {
            if (pack_lines(g.get(), g->base, std::vector<int>(), (int)g->mol.nu.size())) return 1;
        }
        LBL_CUDA(cudaStreamSynchronize(g->s_compute));
    }
    g->stats.n_lines = (int)g->mol.nu.size();
    *out = g.release();
    return 0;
}

int lbl_gas_close(lbl_gas* g)
{
    if (!g) return 0;
    cudaSetDevice(g->device);
    if (g->pending) lbl_gas_wait(g);
    g->base.release();
    g->plan.own.release();
    g->plan.cell_first.release();
    for (DevBuf* b : {&g->tips_t, &g->tips_q, &g->rec_ab, &g->rec_cc, &g->rec_chk, &g->rec_gen,
                      &g->layers_dev, &g->evals_dev, &g->pedbin, &g->pedcorr, &g->pednodes,
                      &g->pedterms, &g->ped_tiles, &g->ped_run_row, &g->ped_n_runs, &g->ped_run_cb,
                      &g->ped_run_sums, &g->rec_f32, &g->amp_max, &g->cell_keys, &g->cheb_nodes, &g->cheb_weights,
                      &g->cheb_nodes16, &g->cheb_weights16, &g->cheb_nodes8, &g->cheb_weights8,
                      &g->executed_dev,
                      &g->out[0], &g->out[1]})
    {
        b->release();
    }
    if (g->layers_host) cudaFreeHost(g->layers_host);
    if (g->mix_scale_host) cudaFreeHost(g->mix_scale_host);
    g->mix_scale_dev.release();
    if (g->evals_host) cudaFreeHost(g->evals_host);
    if (g->executed_host) cudaFreeHost(g->executed_host);
    for (cudaEvent_t e : g->ev_pool) cudaEventDestroy(e);
    if (g->ev_call_begin) cudaEventDestroy(g->ev_call_begin);
    if (g->ev_call_end) cudaEventDestroy(g->ev_call_end);
    if (g->ev_compute_end) cudaEventDestroy(g->ev_compute_end);
    for (int i = 0; i < 2; ++i)
    {
        if (g->ev_out_ready[i]) cudaEventDestroy(g->ev_out_ready[i]);
        if (g->ev_out_free[i]) cudaEventDestroy(g->ev_out_free[i]);
    }
    delete g;
    return 0;
}

}  // extern "C"

// Accumulator of the device-side gas sum.  All additions run on the device's `late` stream, in
// submission order: no atomics, and the sum does not depend on timing.
struct lbl_mix
{
    int device = 0;
    int n_layers = 0, n = 0;
    DeviceStreams* streams = nullptr;
    cudaEvent_t ev_added = nullptr;    // after the last enqueued addition (late stream)
    cudaEvent_t ev_copied = nullptr;   // after the last enqueued copy to the host (copy stream)
    bool copies_pending = false;
    DevBuf acc;
};

namespace
{
// Everything one call may ask for (the extern "C" entry points fill this in).
struct CallSpec
{
    bool blocking = false;
    int n_layers = 0;
    const double* pressure = nullptr;
    const double* temperature = nullptr;
    const double* vmr = nullptr;
    int v0 = 0, vn = 0, n_per_v = 0, cut_off = 0, remove_pedestal = 0, precision = 0;
    int band_lo = 0, band_hi = 0;      // cells [band_lo, band_hi) of the grid; 0, 0 = all of it
    double* k_host = nullptr;          // [n_layers][band points], or NULL
    long long k_pitch = 0;             // doubles between the starts of host rows (0 = dense)
    lbl_mix* mix = nullptr;            // accumulate scale[L]*k[L][:] into rows mix_row0 + L
    int mix_row0 = 0;
    const double* mix_scale = nullptr; // [n_layers], host
    double* mix_host = nullptr;        // non-NULL: copy each finished layer group of the
                                       // accumulator to mix_host (this is the sum's last gas)
};

// The accumulator rows [row0, row0 + rows) go to the host once the additions enqueued so far
// have run.
int mix_rows_to_host(lbl_mix* m, int row0, int rows, double* host)
{
    LBL_CUDA(cudaEventRecord(m->ev_added, m->streams->late));
    LBL_CUDA(cudaStreamWaitEvent(m->streams->copy, m->ev_added, 0));
    LBL_CUDA(cudaMemcpyAsync(host + (size_t)row0 * m->n, m->acc.as<double>() + (size_t)row0 * m->n,
                             sizeof(double) * (size_t)rows * m->n, cudaMemcpyDeviceToHost,
                             m->streams->copy));
    LBL_CUDA(cudaEventRecord(m->ev_copied, m->streams->copy));
    m->copies_pending = true;
    return 0;
}
}  // namespace

static int submit_call(lbl_gas* g, const CallSpec& call)
{
    const bool blocking = call.blocking;
    const int n_layers = call.n_layers;
    const double* pressure = call.pressure;
    const double* temperature = call.temperature;
    const double* vmr = call.vmr;
    const int v0 = call.v0, vn = call.vn, n_per_v = call.n_per_v, cut_off = call.cut_off;
    const int remove_pedestal = call.remove_pedestal, precision = call.precision;
    double* k_host = call.k_host;
    lbl_mix* mix = call.mix;
    if (!g) return fail("Error: null handle.");
    if (g->pending && lbl_gas_wait(g)) return 1;
    if (mix && k_host) return fail("Error: a call feeds either the host array or the accumulator.");
    if (precision != LBL_PRECISION_FP64 && precision != LBL_PRECISION_FP32)
    {
        return fail("Error: unsupported precision mode.");
    }
    const bool fp32_requested = precision == LBL_PRECISION_FP32;
    if (n_layers < 0 || n_per_v < 1 || vn <= v0 || cut_off < 0)
    {
        return fail("Error: invalid grid or layer count.");
    }
    const long long n_ll = (long long)(vn - v0) * n_per_v;
    if (n_ll > (1ll << 30) || (long long)(vn - v0 + 2 * cut_off + 4) * n_per_v > (1ll << 30))
    {
        return fail("Error: spectral grid too large for 32-bit indices.");
    }
    if (set_device(g)) return 1;

    GridSpec grid;
    grid.v0 = v0;
    grid.vn = vn;
    grid.n_per_v = n_per_v;
    grid.cut_off = cut_off;
    grid.n = (int)n_ll;
    grid.ncell = vn - v0;
    grid.dv = 1. / n_per_v;  // absorption.c:33
    grid.cell_lo = 0;
    grid.cell_hi = grid.ncell;
    if (call.band_lo != 0 || call.band_hi != 0)
    {
        if (call.band_lo < 0 || call.band_hi > grid.ncell || call.band_lo >= call.band_hi)
        {
            return fail("Error: band is not a non-empty cell range of the grid.");
        }
        grid.cell_lo = call.band_lo;
        grid.cell_hi = call.band_hi;
    }
    const bool whole_grid = grid.cell_lo == 0 && grid.cell_hi == grid.ncell;
    const int band_n = (grid.cell_hi - grid.cell_lo) * n_per_v;   // output points per layer
    const int band_p0 = grid.cell_lo * n_per_v;
    if (mix)
    {
        if (mix->device != g->device) return fail("Error: gas and accumulator live on different devices.");
        if (mix->n != band_n || call.mix_row0 < 0 || call.mix_row0 + n_layers > mix->n_layers)
        {
            return fail("Error: accumulator shape does not fit this call.");
        }
        if (!call.mix_scale) return fail("Error: null scale array.");
    }

    lbl_stats& st = g->stats;
    st = lbl_stats{};
    st.n_lines = (int)g->mol.nu.size();
    st.n_layers = n_layers;
    st.n_points = band_n;
    g->chunk_events.clear();
    g->sum_events.clear();
    g->ev_used = 0;
    g->last_chunk_layers = 0;
    g->last_grid = grid;
    if (!g->open_h2d_reported)
    {
        st.h2d_bytes += (long long)g->open_h2d;
        g->open_h2d_reported = true;
    }
    if (n_layers == 0) return 0;

    // No TIPS rows: "lines can't be calculated", absorption.c:53-59 -> zeros, success.
    if (call.k_pitch != 0 && call.k_pitch < band_n) return fail("Error: host row pitch shorter than a row.");
    auto zero_host = [&]() {
        if (!k_host) return;
        const size_t pitch = call.k_pitch > 0 ? (size_t)call.k_pitch : (size_t)band_n;
        for (int l = 0; l < n_layers; ++l) std::memset(k_host + (size_t)l * pitch, 0, sizeof(double) * band_n);
    };
    if (!g->mol.has_tips)
    {
        zero_host();
        if (mix && call.mix_host) return mix_rows_to_host(mix, call.mix_row0, n_layers, call.mix_host);
        return 0;
    }
    const size_t h2d_before = g->open_h2d;
    if (make_plan(g, v0, vn, cut_off)) return 1;
    st.h2d_bytes += (long long)(g->open_h2d - h2d_before);
    const Plan& plan = g->plan;
    st.n_active = plan.n_active;
    // Fine grids use the cell-tiled kernel with the polynomial far field (K2c); coarse grids
    // (a cell holds fewer points than the 32 interpolation nodes would cost) use K2.
    // K2c pays ~2.3 ns of set-up per (cell, layer) whatever the number of lines, K2 ~2 ns per
    // (line, layer) on a 0.01 cm-1 grid: below ~0.8 lines per cm-1 (measured cross-over on
    // config 2: CO, O2) the direct kernel is the faster one.  On finer grids K2's cost per line
    // grows with n_per_v, K2c's set-up only in part, so the cross-over density falls.
    const double sparse = 0.8 * std::min(1.0, 300.0 / (double)n_per_v);
    bool farfield = n_per_v >= 64 && (double)plan.n_active >= sparse * (double)grid.ncell;
    if (const char* env = getenv("PYLBL_B200_FARFIELD"))
    {
        farfield = n_per_v >= 64 && atoi(env) != 0 && (atoi(env) == 2 || farfield);   // 2 = always
    }
    // The FP32 mode is a mode of the direct kernel K2.  Where K2c runs it is already faster
    // in FP64 than K2 is in FP32, so the request is honoured with FP64 arithmetic there.
    const bool fp32 = fp32_requested && !farfield;
    st.fp32_used = fp32 ? 1 : 0;
    // cells per warp of K2c: two cells share the loads of the 32-node lines, but one cell per
    // warp needs fewer registers (8 resident blocks per SM) and has no per-cell window edges
    // inside the group; measured faster on every BASELINE grid
    int cells_per_warp = 0;
    if (farfield)
    {
        cells_per_warp = 1;
        if (const char* env = getenv("PYLBL_B200_CELLS")) cells_per_warp = atoi(env) == 2 ? 2 : 1;
    }
    // K2c's blocks sit at absolute cell positions (cell_block_base): the groups run from the block
    // that holds the band's first cell to the end of the block that holds its last
    int cell_groups = 0, cell_first = 0;
    if (farfield)
    {
        const int block_cells = (kCellBlock / 32) * cells_per_warp;
        cell_first = cell_block_base(grid, block_cells);
        const int cell_end = std::min(grid.ncell, (grid.cell_hi + block_cells - 1) / block_cells * block_cells);
        cell_groups = (cell_end - cell_first + cells_per_warp - 1) / cells_per_warp;
    }
    bool hoisted_keys = farfield;
    if (const char* env = getenv("PYLBL_B200_KEYS")) hoisted_keys = farfield && atoi(env) != 0;
    const int P = farfield ? kCellP : pick_points_per_thread(n_per_v);
    st.points_per_thread = P;

    // The reference indexes the TIPS table without a bounds check
    // (spectral_database.c:102-103); outside the table that is undefined behaviour, so it
    // is an error here.
    // Checked per isotopologue block, each against its own first temperature, as the kernel
    // (and the reference) index it.
    auto tips_ok = [&](double t) {
        if (!(t == t)) return false;
        for (int iso = 0; iso < g->mol.num_iso; ++iso)
        {
            const double t0 = g->mol.tips_t[(size_t)iso * g->mol.num_t];
            const long long i = (long long)std::floor(t) - (long long)(int)t0;
            if (!(i >= 0 && i + 1 < g->mol.num_t)) return false;
        }
        return true;
    };
    if (!tips_ok(296.)) return fail("Error: TIPS table does not cover 296 K.");
    for (int l = 0; l < n_layers; ++l)
    {
        if (!tips_ok(temperature[l]))
        {
            return fail("Error: layer temperature outside the TIPS table.");
        }
    }

    if (plan.n_active == 0)
    {
        // Every row is past the early break: the reference returns the zeroed k.
        zero_host();
        if (mix && call.mix_host) return mix_rows_to_host(mix, call.mix_row0, n_layers, call.mix_host);
        return 0;
    }
    const DeviceLines& dl = g->mol.sorted ? g->base : plan.own;
    LinesView lines = dl.view(plan.n_active);
    lines.cell_first = plan.cell_first.as<int>();
    lines.cell_w0 = plan.cell_w0;
    lines.cell_n = plan.cell_n;
    TipsView tips;
    tips.num_iso = g->mol.num_iso;
    tips.num_t = g->mol.num_t;
    tips.t = g->tips_t.as<double>();
    tips.q = g->tips_q.as<double>();

    // ---- chunking over layers ------------------------------------------------------
    // Pedestal chain: K slots per lane cover the 2*cut+3 tracked points of a line window.
    // (+2 spare slots per row for the bare f[s], f[e] of each line.)
    int ped_k = (2 * cut_off + 5 + 31) / 32;
    if (ped_k == 3) ped_k = 4;   // the node ring is indexed with a power-of-two mask
    // nu-sorted databases (HITRAN order): the run-based recurrence (PedRunArgs); otherwise the
    // slot-ring kernels, which make no assumption on the row order.
    bool ped_runs = remove_pedestal && g->mol.sorted;
    if (const char* env = getenv("PYLBL_B200_PEDRUNS")) ped_runs = ped_runs && atoi(env) != 0;
    const bool ped_chain = remove_pedestal && !ped_runs && ped_k <= 4;
    const int ped_wpad = 32 * ped_k;
    const size_t rec_per_layer = (size_t)plan.n_active *
        (sizeof(FarAB) + sizeof(double) + sizeof(LineChk) + sizeof(LineGen) +
         (ped_chain ? sizeof(double) * ped_wpad : 0) + (ped_runs ? 4 * sizeof(double) + 5 * sizeof(int) : 0) +
         (fp32 ? sizeof(Far32) : 0));
    const size_t out_per_layer = sizeof(double) * (size_t)grid.n;
    // Memory budget of one layer group (records + pedestal buffers, and one output slab): a
    // twentieth of what was free on the device when the handle was opened, between 2 and 8 GB.
    // Fewer, larger groups matter with the pedestal on: the chain takes as long for 10 layers
    // as for 60.  (Not queried per call: cudaMemGetInfo stalls the submitting thread for tens
    // of milliseconds while the GPU is busy.)
    const size_t budget = g->group_budget;
    long long chunk = std::min<long long>(n_layers,
                                          std::max<long long>(1, (long long)(budget / rec_per_layer)));
    chunk = std::min<long long>(chunk, std::max<long long>(1, (long long)(budget / out_per_layer)));
    chunk = std::min<long long>(chunk, 65535);   // layers are gridDim.y of the kernels
    int chunk_override = g_chunk_layers;
    if (const char* env = getenv("PYLBL_B200_CHUNK_LAYERS")) chunk_override = atoi(env);
    if (chunk_override > 0)
    {
        chunk = std::min<long long>(chunk, chunk_override);
    }
    const int n_chunks = (int)((n_layers + chunk - 1) / chunk);

    // ---- buffers -------------------------------------------------------------------
    LBL_CUDA(g->rec_ab.reserve(sizeof(FarAB) * (size_t)plan.n_active * chunk + 64));
    LBL_CUDA(g->rec_cc.reserve(sizeof(double) * (size_t)plan.n_active * chunk + 64));
    LBL_CUDA(g->rec_chk.reserve(sizeof(LineChk) * (size_t)plan.n_active * chunk));
    LBL_CUDA(g->rec_gen.reserve(sizeof(LineGen) * (size_t)plan.n_active * chunk));
    if (fp32)
    {
        LBL_CUDA(g->rec_f32.reserve(sizeof(Far32) * (size_t)plan.n_active * chunk));
        LBL_CUDA(g->amp_max.reserve(sizeof(unsigned long long) * (size_t)n_layers));
    }
    if (hoisted_keys)
    {
        LBL_CUDA(g->cell_keys.reserve(sizeof(int) * kKeyStride * (size_t)cell_groups * chunk));
    }
    LBL_CUDA(g->layers_dev.reserve(sizeof(LayerIn) * (size_t)n_layers));
    LBL_CUDA(g->evals_dev.reserve(sizeof(unsigned long long) * (size_t)n_layers));
    LBL_CUDA(g->out[0].reserve(out_per_layer * chunk));
    if (n_chunks > 1) LBL_CUDA(g->out[1].reserve(out_per_layer * chunk));
    const int nb = grid.ncell + 2 * cut_off + 2;
    size_t ped_smem = 0;
    bool ped_nodes_in_smem = true;
    if (remove_pedestal)
    {
        LBL_CUDA(g->pedbin.reserve(sizeof(double) * (size_t)nb * chunk));
        LBL_CUDA(g->pedcorr.reserve(sizeof(double) * 2 * (size_t)grid.ncell * chunk));
        const size_t node_bytes = sizeof(double) * (size_t)(grid.ncell + 1);
        const size_t ring_bytes = ped_chain
            ? (size_t)kPedStages * kPedTile * (sizeof(double) * ped_wpad + sizeof(int4)) : 0;
        // The chain kernel also keeps the per-cell pedestal bins in shared memory.
        const size_t bins_bytes = ped_chain ? sizeof(double) * (size_t)nb : 0;
        ped_smem = ring_bytes + node_bytes + bins_bytes;
        ped_nodes_in_smem = true;
        // With many layers in flight the chain kernel keeps its node and bin arrays in global
        // memory (they are touched only when the window moves, by the lane that owns the index,
        // and stay in L1): a one-warp block holding ~150 KB of shared memory would evict the
        // summation kernel's blocks from its SM for the whole length of the chain.  With few
        // layers (scalar calls) latency matters more and the arrays stay in shared memory.
        if ((ped_chain && chunk > 16) || ped_smem > 200 * 1024)
        {
            // Grid too wide for shared memory: the node array goes to global memory.
            LBL_CUDA(g->pednodes.reserve(node_bytes * chunk));
            ped_smem = ring_bytes;
            ped_nodes_in_smem = false;
        }
        if (ped_runs)
        {
            // sized for the worst case (every row its own run); only the runs' share is touched
            const size_t rows = (size_t)plan.n_active;
            const size_t tiles = (rows + kRunTile - 1) / kRunTile;
            LBL_CUDA(g->ped_tiles.reserve(sizeof(int) * tiles * chunk));
            LBL_CUDA(g->ped_run_row.reserve(sizeof(int) * (rows + 1) * chunk));
            LBL_CUDA(g->ped_n_runs.reserve(sizeof(int) * (size_t)chunk));
            LBL_CUDA(g->ped_run_cb.reserve(sizeof(int) * 4 * rows * chunk));
            LBL_CUDA(g->ped_run_sums.reserve(sizeof(double) * 4 * rows * chunk));
            ped_smem = sizeof(double) * (size_t)nb <= 100 * 1024 ? sizeof(double) * (size_t)nb : 0;
            if (ped_smem > 48 * 1024)
            {
                LBL_CUDA(cudaFuncSetAttribute(ped_chain_runs_kernel,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ped_smem));
            }
        }
        else if (ped_chain)
        {
            LBL_CUDA(g->pedterms.reserve(sizeof(double) * ped_wpad * (size_t)plan.n_active * chunk));
        }
        else if (ped_smem > 48 * 1024)
        {
            LBL_CUDA(cudaFuncSetAttribute(pedestal_kernel,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)ped_smem));
        }
    }
    if (g->layers_host_cap < (size_t)n_layers)
    {
        if (g->layers_host) cudaFreeHost(g->layers_host);
        g->layers_host = nullptr;
        LBL_CUDA(cudaHostAlloc((void**)&g->layers_host, sizeof(LayerIn) * n_layers,
                               cudaHostAllocDefault));
        g->layers_host_cap = n_layers;
    }
    if (farfield)
    {
        const size_t h2d0 = g->open_h2d;
        if (ensure_cheb_tables(g, n_per_v)) return 1;
        st.h2d_bytes += (long long)(g->open_h2d - h2d0);
        LBL_CUDA(g->executed_dev.reserve(sizeof(unsigned long long)));
        if (!g->executed_host)
        {
            LBL_CUDA(cudaHostAlloc((void**)&g->executed_host, sizeof(unsigned long long),
                                   cudaHostAllocDefault));
        }
    }
    if (g->evals_host_cap < (size_t)n_layers)
    {
        if (g->evals_host) cudaFreeHost(g->evals_host);
        g->evals_host = nullptr;
        LBL_CUDA(cudaHostAlloc((void**)&g->evals_host, sizeof(unsigned long long) * n_layers,
                               cudaHostAllocDefault));
        g->evals_host_cap = n_layers;
    }

    // ---- per-layer inputs ------------------------------------------------------------
    g->last_layers.resize(n_layers);
    for (int l = 0; l < n_layers; ++l)
    {
        LayerIn& ly = g->layers_host[l];
        ly.pressure = pressure[l];
        ly.temperature = temperature[l];
        ly.vmr = vmr[l];
        const double p_atm = std::fabs(pressure[l] * kPaToAtm);
        ly.slack = p_atm * g->mol.max_abs_delta * (1. + 1e-9) + 1e-9;
        // Near-zone half width D = xlim0*alpha/sqrt(ln2) <= 123.4/0.8325 * alpha, and
        // alpha = nu*sqrt(r2*T/mass)/vlight (spectra.c:29, voigt.c:13,34).
        const double mm = g->mol.min_mass > 0. ? g->mol.min_mass : 1.;
        ly.kappa = 148.3 * std::sqrt(kR2 * std::fabs(temperature[l]) / mm) / kVlight;
        ly.pad = 0.;
        g->last_layers[l] = ly;
    }

    // Rows the pedestal recurrence must walk.  It runs in database order over the whole grid's
    // active lines; for a band of a nu-sorted database the rows past the band's last window
    // (window cell > cell_hi-1+cut, i.e. nu > v0+cell_hi+cut+slack) come later in the order
    // and cannot reach the band's bins.
    int ped_rows = plan.n_active;
    if (remove_pedestal && !whole_grid && g->mol.sorted)
    {
        double slack_max = 0.;
        for (int l = 0; l < n_layers; ++l) slack_max = std::max(slack_max, g->layers_host[l].slack);
        const double limit = (double)v0 + (double)grid.cell_hi + (double)cut_off + slack_max;
        ped_rows = (int)(std::upper_bound(g->mol.nu.begin(), g->mol.nu.begin() + plan.n_active, limit) -
                         g->mol.nu.begin());
    }
    if (mix)
    {
        // the scales ride behind the layer states in the same pinned buffer / device buffer
        LBL_CUDA(g->mix_scale_dev.reserve(sizeof(double) * (size_t)n_layers));
        if (g->mix_scale_host_cap < (size_t)n_layers)
        {
            if (g->mix_scale_host) cudaFreeHost(g->mix_scale_host);
            g->mix_scale_host = nullptr;
            LBL_CUDA(cudaHostAlloc((void**)&g->mix_scale_host, sizeof(double) * n_layers,
                                   cudaHostAllocDefault));
            g->mix_scale_host_cap = n_layers;
        }
        std::memcpy(g->mix_scale_host, call.mix_scale, sizeof(double) * (size_t)n_layers);
    }

    cudaStream_t sc = g->s_compute;               // early: uploads and scaling kernels
    cudaStream_t sm = g->s_main;                  // summation kernels
    cudaStream_t sl = g->s_late;                  // apply, statistics
    cudaStream_t ss = next_side_stream(g->streams);   // pedestal terms + chain of this call
    cudaEvent_t previous_applied = nullptr;
    LBL_CUDA(cudaEventRecord(g->ev_call_begin, sc));
    LBL_CUDA(cudaMemcpyAsync(g->layers_dev.p, g->layers_host, sizeof(LayerIn) * n_layers,
                             cudaMemcpyHostToDevice, sc));
    LBL_CUDA(cudaMemsetAsync(g->evals_dev.p, 0, sizeof(unsigned long long) * n_layers, sc));
    if (mix)
    {
        LBL_CUDA(cudaMemcpyAsync(g->mix_scale_dev.p, g->mix_scale_host, sizeof(double) * n_layers,
                                 cudaMemcpyHostToDevice, sc));
        st.h2d_bytes += (long long)(sizeof(double) * n_layers);
    }
    if (farfield)
    {
        LBL_CUDA(cudaMemsetAsync(g->executed_dev.p, 0, sizeof(unsigned long long), sc));
    }
    g->farfield_last = farfield;
    if (fp32)
    {
        LBL_CUDA(cudaMemsetAsync(g->amp_max.p, 0, sizeof(unsigned long long) * n_layers, sc));
    }
    st.h2d_bytes += (long long)(sizeof(LayerIn) * n_layers);

    Records rec;
    rec.ab = g->rec_ab.as<FarAB>();
    rec.cc = g->rec_cc.as<double>();
    rec.chk = g->rec_chk.as<LineChk>();
    rec.gen = g->rec_gen.as<LineGen>();
    rec.f32 = fp32 ? g->rec_f32.as<Far32>() : nullptr;
    rec.amp_max = nullptr;

    for (int c = 0; c < n_chunks; ++c)
    {
        const int first = (int)(c * chunk);
        const int nl = (int)std::min<long long>(chunk, n_layers - first);
        const int slot = c & 1;
        const LayerIn* layers_c = g->layers_dev.as<LayerIn>() + first;
        lbl_gas::ChunkEvents ev;
        ev.pedestal = remove_pedestal != 0;
        ev.k1_begin = next_event(g);
        ev.k1_end = next_event(g);
        ev.ped_begin = next_event(g);
        ev.ped_end = next_event(g);
        ev.applied = next_event(g);

        if (g->out_busy[slot])
        {
            // The previous copy out of this buffer must have drained.
            LBL_CUDA(cudaStreamWaitEvent(sm, g->ev_out_free[slot], 0));
            g->out_busy[slot] = false;
        }
        if (previous_applied)
        {
            // The records, terms and corrections of the previous layer group are rewritten.
            LBL_CUDA(cudaStreamWaitEvent(sc, previous_applied, 0));
        }

        // K1 (+ K1f in FP32 mode)
        LBL_CUDA(cudaEventRecord(ev.k1_begin, sc));
        {
            rec.amp_max = fp32 ? g->amp_max.as<unsigned long long>() + first : nullptr;
            dim3 grid1((plan.n_active + kScaleBlock - 1) / kScaleBlock, nl);
            scale_kernel<<<grid1, kScaleBlock, 0, sc>>>(lines, tips, layers_c, grid, rec,
                                                        g->evals_dev.as<unsigned long long>() + first);
            st.total_launches++;
            if (fp32)
            {
                far32_kernel<<<grid1, kScaleBlock, 0, sc>>>(rec, plan.n_active);
                st.total_launches++;
            }
            if (hoisted_keys)
            {
                dim3 gridk((cell_groups * kKeyStride + 255) / 256, nl);
                cell_keys_kernel<<<gridk, 256, 0, sc>>>(lines, grid, layers_c, cells_per_warp, cell_first, cell_groups,
                                                        g->cell_keys.as<int>());
                st.total_launches++;
            }
        }
        LBL_CUDA(cudaEventRecord(ev.k1_end, sc));

        // K3 + K4a on the side stream, concurrent with K2.
        if (remove_pedestal)
        {
            LBL_CUDA(cudaStreamWaitEvent(ss, ev.k1_end, 0));
            LBL_CUDA(cudaEventRecord(ev.ped_begin, ss));
            PedArgs pa;
            pa.lines = lines;
            pa.rec = rec;
            pa.grid = grid;
            pa.pedbin = g->pedbin.as<double>();
            pa.n_rows = ped_rows;
            double* scratch = ped_nodes_in_smem ? nullptr : g->pednodes.as<double>();
            if (ped_runs)
            {
                PedRunArgs ra;
                ra.lines = lines;
                ra.rec = rec;
                ra.grid = grid;
                ra.layers = layers_c;
                ra.n_rows = ped_rows;
                ra.run_row = g->ped_run_row.as<int>();
                ra.n_runs = g->ped_n_runs.as<int>();
                ra.run_cb = g->ped_run_cb.as<int>();
                ra.run_sums = g->ped_run_sums.as<double>();
                ra.pedbin = g->pedbin.as<double>();
                const int tiles = (ped_rows + kRunTile - 1) / kRunTile;
                if (tiles > 0)
                {
                    dim3 gt(tiles, nl);
                    ped_run_count_kernel<<<gt, kRunTile, 0, ss>>>(rec.chk, lines.n, ped_rows,
                                                                 g->ped_tiles.as<int>());
                    ped_run_scatter_kernel<<<gt, kRunTile, 0, ss>>>(rec.chk, lines.n, ped_rows,
                                                                   g->ped_tiles.as<int>(), ra.run_row, ra.n_runs);
                    // about one run per occupied cell: enough warps to take them in a few rounds
                    // (with few layers -- the scalar plugin call -- more warps per layer)
                    const int runs_guess = std::min(ped_rows, grid.ncell + 2 * cut_off + 8);
                    const int per_layer = std::max(64, 1184 / nl);
                    dim3 gn(std::max(1, std::min(per_layer, (runs_guess / kNodeRuns + 7) / 8)), nl);
                    ped_nodes_kernel<<<gn, 256, 0, ss>>>(ra);
                }
                else
                {
                    LBL_CUDA(cudaMemsetAsync(ra.n_runs, 0, sizeof(int) * nl, ss));
                }
                {
                    int use_scan = 1;
                    if (const char* env = getenv("PYLBL_B200_PEDSCAN")) use_scan = atoi(env) != 0;
                    ped_chain_runs_kernel<<<nl, 32, ped_smem, ss>>>(ra, ped_smem > 0 ? 1 : 0, use_scan);
                }
                st.total_launches += 4;
            }
            else if (ped_chain)
            {
                cudaError_t ce = cudaSuccess;
                switch (ped_k)
                {
                    case 1: ce = launch_chain<1>(pa, g->pedterms.as<double>(), scratch, nl, ped_smem, ss); break;
                    case 2: ce = launch_chain<2>(pa, g->pedterms.as<double>(), scratch, nl, ped_smem, ss); break;
                    default: ce = launch_chain<4>(pa, g->pedterms.as<double>(), scratch, nl, ped_smem, ss); break;
                }
                LBL_CUDA(ce);
                st.total_launches++;
            }
            else
            {
                pedestal_kernel<<<nl, 32, ped_smem, ss>>>(pa, scratch);
            }
            const int cells = nl * grid.ncell;
            pedestal_cells_kernel<<<(cells + 127) / 128, 128, 0, ss>>>(
                g->pedbin.as<double>(), grid, nl, g->pedcorr.as<double>());
            st.total_launches += 2;
            LBL_CUDA(cudaEventRecord(ev.ped_end, ss));
        }

        // K2
        SumArgs sa;
        sa.lines = lines;
        sa.rec = rec;
        sa.layers = layers_c;
        sa.grid = grid;
        sa.out = g->out[slot].as<double>();
        sa.tpw = pick_threads_per_layer(n_per_v, P, nl);
        sa.near_masked = farfield ? 0 : 1;
        CellArgs ca;
        if (farfield)
        {
            ca.node_offset = g->cheb_nodes.as<double>();
            ca.transform = g->cheb_weights.as<double>();
            ca.node_offset16 = g->cheb_nodes16.as<double>();
            ca.transform16 = g->cheb_weights16.as<double>();
            ca.node_offset8 = g->cheb_nodes8.as<double>();
            ca.transform8 = g->cheb_weights8.as<double>();
            ca.executed = g->executed_dev.as<unsigned long long>();
            ca.keys = hoisted_keys ? g->cell_keys.as<int>() : nullptr;
            ca.key_groups = cell_groups;
            st.cells_per_warp = cells_per_warp;
        }
        // Layer groups of this chunk: each is summed (main stream), corrected and copied out
        // while the next one computes, so that only the last group's copy is exposed.  The
        // records and the pedestal chain are per chunk (the chain takes as long for 15 layers
        // as for 60).
        int n_groups = 1;
        if (farfield && (k_host || (mix && call.mix_host)))
        {
            // Splitting costs a kernel tail per group.  It pays when the caller blocks on this
            // call (nothing else would hide the copy); with several calls in flight the copy of
            // one gas hides behind the kernels of the next and one group is best.
            n_groups = (blocking && nl >= 16) ? 2 : 1;
            // the last gas of a device-side sum: only its last group's rows are copied exposed
            if (mix && call.mix_host && nl >= 8) n_groups = std::min(4, nl / 4);
            if (g->copy_groups > 0) n_groups = std::min(g->copy_groups, nl);
            if (const char* env = getenv("PYLBL_B200_COPY_GROUPS")) n_groups = std::max(1, std::min(atoi(env), nl));
        }
        LBL_CUDA(cudaStreamWaitEvent(sm, ev.k1_end, 0));
        for (int q = 0; q < n_groups; ++q)
        {
            const int q0 = (int)((long long)nl * q / n_groups);
            const int q1 = (int)((long long)nl * (q + 1) / n_groups);
            lbl_gas::SumEvents se;
            se.k2_begin = next_event(g);
            se.k2_end = next_event(g);
            se.k2b_end = next_event(g);
            sa.layer0 = q0;
            sa.n_layers = q1;
            LBL_CUDA(cudaEventRecord(se.k2_begin, sm));
            if (farfield)
            {
                ca.sum = sa;
                ca.key_layer0 = q0;
                const int groups = cell_groups;
                dim3 gridc((groups + kCellBlock / 32 - 1) / (kCellBlock / 32), q1 - q0);
                // 23 KB of static shared memory per block: ask for the large carve-out so that
                // shared memory does not cap the resident blocks below the register limit.
                int far_mode = 1;
                if (const char* env = getenv("PYLBL_B200_FARLOOP")) far_mode = atoi(env);
                auto launch = [&](auto kernel) -> cudaError_t {
                    cudaError_t ce = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                                          cudaSharedmemCarveoutMaxShared);
                    if (ce != cudaSuccess) return ce;
                    kernel<<<gridc, kCellBlock, 0, sm>>>(ca);
                    return cudaSuccess;
                };
                if (cells_per_warp == 2 && far_mode == 0) LBL_CUDA(launch(sum_cell_kernel<2, 0>));
                else if (cells_per_warp == 2) LBL_CUDA(launch(sum_cell_kernel<2, 1>));
                else if (far_mode == 2) LBL_CUDA(launch(sum_cell_kernel<1, 2>));
                else if (far_mode == 0) LBL_CUDA(launch(sum_cell_kernel<1, 0>));
                else LBL_CUDA(launch(sum_cell_kernel<1, 1>));
            }
            else
            {
                launch_sum_dispatch(P, sa, nl, fp32, sm);
            }
            LBL_CUDA(cudaEventRecord(se.k2_end, sm));
            if (near_block_applies(sa, g->last_layers, first, nl))
            {
                launch_near_block(sa, q1 - q0, sm);
            }
            else
            {
                launch_fixup_dispatch(pick_fixup_tile(n_per_v), sa, q1 - q0, sm);
            }
            LBL_CUDA(cudaEventRecord(se.k2b_end, sm));
            st.sum_launches++;
            st.total_launches += 2;
            g->sum_events.push_back(se);
            // The late stream takes over: apply the pedestal corrections, hand the group to the
            // copy stream.
            LBL_CUDA(cudaStreamWaitEvent(sl, se.k2b_end, 0));

            if (remove_pedestal && q == 0) LBL_CUDA(cudaStreamWaitEvent(sl, ev.ped_end, 0));
            if (remove_pedestal || mix)
            {
                const size_t total = (size_t)(q1 - q0) * band_n;
                const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
                double* out_q = g->out[slot].as<double>() + (size_t)q0 * grid.n;
                const double* corr_q = remove_pedestal
                    ? g->pedcorr.as<double>() + 2 * (size_t)q0 * grid.ncell : nullptr;
                double* acc_q = mix ? mix->acc.as<double>() +
                                          (size_t)(call.mix_row0 + first + q0) * band_n : nullptr;
                const double* scale_q = mix ? g->mix_scale_dev.as<double>() + first + q0 : nullptr;
                if (mix && remove_pedestal)
                    apply_kernel<true, true><<<blocks, 256, 0, sl>>>(out_q, corr_q, grid, q1 - q0, acc_q, scale_q);
                else if (mix)
                    apply_kernel<false, true><<<blocks, 256, 0, sl>>>(out_q, corr_q, grid, q1 - q0, acc_q, scale_q);
                else
                    apply_kernel<true, false><<<blocks, 256, 0, sl>>>(out_q, corr_q, grid, q1 - q0, acc_q, scale_q);
                st.total_launches++;
            }
            LBL_CUDA(cudaGetLastError());

            if (k_host)
            {
                LBL_CUDA(cudaEventRecord(g->ev_out_ready[slot], sl));
                LBL_CUDA(cudaStreamWaitEvent(g->s_copy, g->ev_out_ready[slot], 0));
                const size_t pitch = call.k_pitch > 0 ? (size_t)call.k_pitch : (size_t)band_n;
                if (whole_grid && pitch == (size_t)grid.n)
                {
                    LBL_CUDA(cudaMemcpyAsync(k_host + (size_t)(first + q0) * grid.n,
                                             g->out[slot].as<double>() + (size_t)q0 * grid.n,
                                             out_per_layer * (q1 - q0), cudaMemcpyDeviceToHost, g->s_copy));
                }
                else
                {
                    // the band's columns of the whole-grid rows
                    LBL_CUDA(cudaMemcpy2DAsync(k_host + (size_t)(first + q0) * pitch,
                                               sizeof(double) * pitch,
                                               g->out[slot].as<double>() + (size_t)q0 * grid.n + band_p0,
                                               sizeof(double) * grid.n, sizeof(double) * band_n,
                                               q1 - q0, cudaMemcpyDeviceToHost, g->s_copy));
                }
                st.d2h_bytes += (long long)(sizeof(double) * (size_t)band_n * (q1 - q0));
            }
            if (mix && call.mix_host)
            {
                // last gas of the sum: these rows of the accumulator are final
                if (mix_rows_to_host(mix, call.mix_row0 + first + q0, q1 - q0, call.mix_host)) return 1;
                st.d2h_bytes += (long long)(sizeof(double) * (size_t)band_n * (q1 - q0));
            }
        }
        LBL_CUDA(cudaEventRecord(ev.applied, sl));
        previous_applied = ev.applied;
        g->chunk_events.push_back(ev);
        if (k_host)
        {
            LBL_CUDA(cudaEventRecord(g->ev_out_free[slot], g->s_copy));
            g->out_busy[slot] = true;
        }
        g->last_chunk_first = first;
        g->last_chunk_layers = nl;
        g->last_slot = slot;
    }
    if (mix)
    {
        LBL_CUDA(cudaEventRecord(mix->ev_added, sl));
    }
    // The small statistics copies go where they cannot hold anything up: all device-to-host
    // copies share one DMA queue, and on the late stream they would sit behind the bulk copies
    // in flight -- this call's, or those of another gas of the column (the previous layer group
    // of a gas sum) -- and with them every later gas's correction kernels.  They ride the copy
    // stream, after the kernels' end mark; the call ends when the last copy has landed.
    g->copies_in_call = k_host || (mix && call.mix_host);
    LBL_CUDA(cudaEventRecord(g->ev_compute_end, sl));      // the last kernel of the call
    cudaStream_t s_end = g->s_copy;
    LBL_CUDA(cudaStreamWaitEvent(g->s_copy, g->ev_compute_end, 0));
    LBL_CUDA(cudaMemcpyAsync(g->evals_host, g->evals_dev.p, sizeof(unsigned long long) * n_layers,
                             cudaMemcpyDeviceToHost, s_end));
    if (farfield)
    {
        LBL_CUDA(cudaMemcpyAsync(g->executed_host, g->executed_dev.p, sizeof(unsigned long long),
                                 cudaMemcpyDeviceToHost, s_end));
    }
    LBL_CUDA(cudaEventRecord(g->ev_call_end, s_end));
    g->pending = true;
    if (mix) g->last_chunk_layers = 0;   // the spectra went into the accumulator, uncorrected here
    return 0;
}

extern "C" {

int lbl_gas_wait(lbl_gas* g)
{
    if (!g) return fail("Error: null handle.");
    if (!g->pending) return 0;
    if (set_device(g)) return 1;
    g->pending = false;
    // ev_call_end closes the call: it follows the last kernel, the last copy (the compute
    // stream waited for it) and, through the pedestal events, the side stream.
    LBL_CUDA(cudaEventSynchronize(g->ev_call_end));
    g->out_busy[0] = g->out_busy[1] = false;
    lbl_stats& st = g->stats;
    for (int l = 0; l < st.n_layers; ++l) st.evals += (long long)g->evals_host[l];
    st.executed = g->farfield_last ? (long long)*g->executed_host : st.evals;
    float ms = 0.f;
    for (const lbl_gas::ChunkEvents& ev : g->chunk_events)
    {
        LBL_CUDA(cudaEventElapsedTime(&ms, ev.k1_begin, ev.k1_end));
        st.scale_ms += ms;
        if (ev.pedestal)
        {
            LBL_CUDA(cudaEventElapsedTime(&ms, ev.ped_begin, ev.ped_end));
            st.pedestal_ms += ms;
        }
    }
    for (const lbl_gas::SumEvents& se : g->sum_events)
    {
        LBL_CUDA(cudaEventElapsedTime(&ms, se.k2_begin, se.k2_end));
        st.sum_ms += ms;
        LBL_CUDA(cudaEventElapsedTime(&ms, se.k2_end, se.k2b_end));
        st.fixup_ms += ms;
    }
    LBL_CUDA(cudaEventElapsedTime(&ms, g->ev_call_begin, g->ev_call_end));
    st.total_ms = ms;
    if (g->copies_in_call)
    {
        LBL_CUDA(cudaEventElapsedTime(&ms, g->ev_compute_end, g->ev_call_end));
        st.copy_tail_ms = ms;
    }
    if (getenv("PYLBL_B200_TIMELINE"))
    {
        // debugging aid: when this call's kernels and copies ended, relative to lbl_timer_start
        DeviceTimer* t = nullptr;
        if (get_timer(g->device, &t) == 0 && !g->sum_events.empty())
        {
            float a = 0.f, b = 0.f, c = 0.f;
            cudaEventElapsedTime(&a, t->begin, g->ev_call_begin);
            cudaEventElapsedTime(&b, t->begin, g->sum_events.back().k2b_end);
            cudaEventElapsedTime(&c, t->begin, g->ev_call_end);
            fprintf(stderr, "timeline %-4s begin %7.2f  sum+k2b end %7.2f  call end %7.2f ms\n",
                    g->formula.c_str(), a, b, c);
        }
    }
    return 0;
}

static CallSpec basic_call(bool blocking, int n_layers, const double* pressure,
                           const double* temperature, const double* vmr, int v0, int vn, int n_per_v,
                           int cut_off, int remove_pedestal, int precision, double* k_host)
{
    CallSpec c;
    c.blocking = blocking;
    c.n_layers = n_layers;
    c.pressure = pressure;
    c.temperature = temperature;
    c.vmr = vmr;
    c.v0 = v0;
    c.vn = vn;
    c.n_per_v = n_per_v;
    c.cut_off = cut_off;
    c.remove_pedestal = remove_pedestal;
    c.precision = precision;
    c.k_host = k_host;
    return c;
}

int lbl_gas_submit(lbl_gas* g, int n_layers, const double* pressure, const double* temperature,
                   const double* vmr, int v0, int vn, int n_per_v, int cut_off,
                   int remove_pedestal, int precision, double* k_host)
{
    return submit_call(g, basic_call(false, n_layers, pressure, temperature, vmr, v0, vn, n_per_v,
                                     cut_off, remove_pedestal, precision, k_host));
}

int lbl_gas_compute(lbl_gas* g, int n_layers, const double* pressure, const double* temperature,
                    const double* vmr, int v0, int vn, int n_per_v, int cut_off,
                    int remove_pedestal, int precision, double* k_host)
{
    if (submit_call(g, basic_call(true, n_layers, pressure, temperature, vmr, v0, vn, n_per_v,
                                  cut_off, remove_pedestal, precision, k_host)))
    {
        return 1;
    }
    return lbl_gas_wait(g);
}

int lbl_gas_submit_band(lbl_gas* g, int n_layers, const double* pressure, const double* temperature,
                        const double* vmr, int v0, int vn, int n_per_v, int cut_off,
                        int remove_pedestal, int precision, int band_lo, int band_hi, double* k_host,
                        long long k_pitch)
{
    CallSpec c = basic_call(false, n_layers, pressure, temperature, vmr, v0, vn, n_per_v, cut_off,
                            remove_pedestal, precision, k_host);
    if (band_lo == 0 && band_hi == 0) return fail("Error: empty band.");
    c.band_lo = band_lo;
    c.band_hi = band_hi;
    c.k_pitch = k_pitch;
    return submit_call(g, c);
}

int lbl_gas_submit_mix(lbl_gas* g, int n_layers, const double* pressure, const double* temperature,
                       const double* vmr, int v0, int vn, int n_per_v, int cut_off,
                       int remove_pedestal, int precision, lbl_mix* mix, int row0,
                       const double* scale, double* total_host)
{
    if (!mix) return fail("Error: null accumulator.");
    CallSpec c = basic_call(false, n_layers, pressure, temperature, vmr, v0, vn, n_per_v, cut_off,
                            remove_pedestal, precision, nullptr);
    c.mix = mix;
    c.mix_row0 = row0;
    c.mix_scale = scale;
    c.mix_host = total_host;
    return submit_call(g, c);
}

int lbl_gas_band_edges(lbl_gas* g, int v0, int vn, int n_per_v, int cut_off, int n_bands, int* edges)
{
    if (!g || !edges) return fail("Error: null argument.");
    const int ncell = vn - v0;
    if (ncell <= 0 || n_bands < 1 || n_per_v < 1 || cut_off < 0) return fail("Error: invalid grid or band count.");
    // Cost model, in units of one far-line node evaluation, fitted to the kernel times of equal
    // and of unequal bands of BASELINE configs[3] (tools/band_cost.py; all 16 within 0.3 ms of
    // 10-34 ms).  A cell costs a fixed part (range boundaries, transforms, interpolation), the
    // lines of its window (a few node evaluations each), the lines next to it (evaluated at
    // every one of its n_per_v points) and their near zones, whose width grows with the
    // wavenumber (Doppler width: 2*kappa*nu*n_per_v points).  A BAND costs its cells plus the
    // pedestal recurrence over every active row up to its last window (the prefix of the whole
    // grid) -- about 55 evaluations per row, run beside the band's own kernels.
    const int na = active_prefix(g->mol, v0, vn, cut_off);
    std::vector<double> nu(g->mol.nu.begin(), g->mol.nu.begin() + na);
    if (!g->mol.sorted) std::sort(nu.begin(), nu.end());
    auto count = [&](double lo, double hi) {
        return (double)(std::lower_bound(nu.begin(), nu.end(), hi) - std::lower_bound(nu.begin(), nu.end(), lo));
    };
    const double mass = g->mol.min_mass > 0. ? g->mol.min_mass : 30.;
    const double kappa = 148.3 * std::sqrt(kR2 * 250. / mass) / kVlight;
    std::vector<double> cum((size_t)ncell + 1, 0.), prefix((size_t)ncell + 1, 0.);
    for (int c = 0; c < ncell; ++c)
    {
        const double w = (double)v0 + c;
        const double beside = count(w - 0.5, w + 1.5);
        const double zone_points = 2. * kappa * std::fabs(w + 0.5) * n_per_v;
        const double cost = 400. + 0.3 * n_per_v + count(w - cut_off, w + cut_off + 1.) +
                            0.125 * n_per_v * beside + 0.41 * zone_points * beside;
        cum[c + 1] = cum[c] + cost;
        prefix[c + 1] = 55. * count(-1e300, w + 1. + cut_off + 1.);   // rows a band ending here walks
    }
    partition_bands(cum, prefix, n_bands, edges);
    return 0;
}

int lbl_gas_set_copy_groups(lbl_gas* g, int groups)
{
    if (!g) return fail("Error: null handle.");
    if (groups < 0) return fail("Error: negative group count.");
    g->copy_groups = groups;
    return 0;
}

int lbl_gas_stats(lbl_gas* g, lbl_stats* out)
{
    if (!g || !out) return fail("Error: null argument.");
    *out = g->stats;
    return 0;
}

int lbl_gas_device_result(lbl_gas* g, double** device_ptr, long long* count)
{
    if (!g) return fail("Error: null handle.");
    if (g->pending && lbl_gas_wait(g)) return 1;
    if (g->last_chunk_layers == 0) return fail("Error: no spectra resident on the device.");
    *device_ptr = g->out[g->last_slot].as<double>();
    *count = (long long)g->last_chunk_layers * g->last_grid.n;
    return 0;
}

static int resident_layer(lbl_gas* g, int layer, int capacity, int* local)
{
    if (!g) return fail("Error: null handle.");
    if (g->pending && lbl_gas_wait(g)) return 1;
    *local = 0;
    if (g->stats.n_active == 0 && layer >= 0 && layer < g->stats.n_layers)
    {
        return 0;  // nothing was processed (early break on the first row, or no TIPS data)
    }
    if (layer < g->last_chunk_first || layer >= g->last_chunk_first + g->last_chunk_layers)
    {
        return fail("Error: that layer's line records are no longer resident.");
    }
    if (capacity < g->stats.n_active) return fail("Error: output capacity too small.");
    *local = layer - g->last_chunk_first;
    return set_device(g);
}

int lbl_gas_windows(lbl_gas* g, int layer, int* s_out, int* e_out, int capacity)
{
    int local = 0;
    if (resident_layer(g, layer, capacity, &local)) return 1;
    const int na = g->stats.n_active;
    if (na == 0) return 0;
    std::vector<LineChk> chk(na);
    LBL_CUDA(cudaMemcpy(chk.data(), g->rec_chk.as<LineChk>() + (size_t)local * na,
                        sizeof(LineChk) * na, cudaMemcpyDeviceToHost));
    const GridSpec& gr = g->last_grid;
    const std::vector<int>& order = g->plan.sorted_to_db;
    for (int j = 0; j < na; ++j)
    {
        const int r = order.empty() ? j : order[j];
        // spectra.c:48-62 from the bit-exact cell of the shifted centre.
        long long s = (long long)(chk[j].cb - gr.cut_off) * gr.n_per_v;
        if (s >= gr.n)
        {
            s_out[r] = -1;
            e_out[r] = -1;
            continue;
        }
        if (s < 0) s = 0;
        long long e = (long long)(chk[j].cb + gr.cut_off + 1) * gr.n_per_v;
        if (e >= gr.n) e = gr.n - 1;
        s_out[r] = (int)s;
        e_out[r] = (int)e;
    }
    return 0;
}

int lbl_gas_scaled(lbl_gas* g, int layer, double* out, int capacity)
{
    int local = 0;
    if (resident_layer(g, layer, capacity, &local)) return 1;
    const int na = g->stats.n_active;
    if (na == 0) return 0;
    std::vector<LineGen> gen(na);
    LBL_CUDA(cudaMemcpy(gen.data(), g->rec_gen.as<LineGen>() + (size_t)local * na,
                        sizeof(LineGen) * na, cudaMemcpyDeviceToHost));
    const std::vector<int>& order = g->plan.sorted_to_db;
    for (int j = 0; j < na; ++j)
    {
        const int r = order.empty() ? j : order[j];
        out[4 * r + 0] = gen[j].nu;
        out[4 * r + 1] = kSqrLn2 / gen[j].repwid;                 // alpha
        out[4 * r + 2] = gen[j].y / gen[j].repwid;                // gamma
        out[4 * r + 3] = gen[j].cof / (kRsqrPi * gen[j].repwid);  // sw'
    }
    return 0;
}

// ---- gas-summed absorption on the device ----------------------------------------------------
int lbl_mix_open(int device, int n_layers, int n_points, lbl_mix** out)
{
    *out = nullptr;
    if (n_layers <= 0 || n_points <= 0) return fail("Error: invalid accumulator shape.");
    int ndev = 0;
    if (lbl_device_count(&ndev)) return 1;
    if (device < 0 || device >= ndev) return fail("Error: CUDA device index out of range.");
    std::unique_ptr<lbl_mix> m(new lbl_mix);
    m->device = device;
    m->n_layers = n_layers;
    m->n = n_points;
    LBL_CUDA(cudaSetDevice(device));
    if (device_streams(device, &m->streams)) return 1;
    LBL_CUDA(cudaEventCreateWithFlags(&m->ev_added, cudaEventDisableTiming));
    LBL_CUDA(cudaEventCreateWithFlags(&m->ev_copied, cudaEventDisableTiming));
    LBL_CUDA(m->acc.reserve(sizeof(double) * (size_t)n_layers * n_points));
    lbl_mix* raw = m.release();
    *out = raw;
    return lbl_mix_reset(raw);
}

int lbl_mix_reset(lbl_mix* m)
{
    if (!m) return fail("Error: null handle.");
    LBL_CUDA(cudaSetDevice(m->device));
    // on the late stream, where the additions run: ordered with them, and behind any copy to
    // the host still reading the previous sum
    if (m->copies_pending)
    {
        LBL_CUDA(cudaStreamWaitEvent(m->streams->late, m->ev_copied, 0));
    }
    LBL_CUDA(cudaMemsetAsync(m->acc.p, 0, sizeof(double) * (size_t)m->n_layers * m->n,
                             m->streams->late));
    LBL_CUDA(cudaEventRecord(m->ev_added, m->streams->late));
    return 0;
}

int lbl_mix_add(lbl_mix* m, lbl_gas* g, const double* scale)
{
    if (!m || !g || !scale) return fail("Error: null argument.");
    if (g->device != m->device) return fail("Error: gas and accumulator live on different devices.");
    LBL_CUDA(cudaSetDevice(m->device));
    if (g->pending && lbl_gas_wait(g)) return 1;
    if (g->stats.n_layers != m->n_layers || g->stats.n_points != m->n)
    {
        return fail("Error: accumulator shape differs from the gas's last call.");
    }
    if (g->stats.n_active == 0)
    {
        return 0;   // the gas contributed an all-zero spectrum
    }
    if (g->last_chunk_first != 0 || g->last_chunk_layers != m->n_layers)
    {
        return fail("Error: the gas's spectra are not resident on the device "
                    "(compute with k_host == NULL and a single layer group, or use lbl_gas_submit_mix).");
    }
    DevBuf scale_dev;
    LBL_CUDA(scale_dev.reserve(sizeof(double) * (size_t)m->n_layers));
    cudaStream_t sl = m->streams->late;
    LBL_CUDA(cudaMemcpyAsync(scale_dev.p, scale, sizeof(double) * m->n_layers, cudaMemcpyHostToDevice, sl));
    const size_t total = (size_t)m->n_layers * m->n;
    const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
    apply_kernel<false, true><<<blocks, 256, 0, sl>>>(g->out[g->last_slot].as<double>(), nullptr,
                                                      g->last_grid, m->n_layers, m->acc.as<double>(),
                                                      scale_dev.as<double>());
    LBL_CUDA(cudaGetLastError());
    LBL_CUDA(cudaEventRecord(m->ev_added, sl));
    // the gas's spectra and the staged scales are consumed before anything reuses them
    LBL_CUDA(cudaStreamSynchronize(sl));
    scale_dev.release();
    return 0;
}

int lbl_mix_wait(lbl_mix* m)
{
    if (!m) return fail("Error: null handle.");
    LBL_CUDA(cudaSetDevice(m->device));
    LBL_CUDA(cudaEventSynchronize(m->ev_added));
    if (m->copies_pending)
    {
        LBL_CUDA(cudaEventSynchronize(m->ev_copied));
        m->copies_pending = false;
    }
    return 0;
}

int lbl_mix_download(lbl_mix* m, double* host)
{
    if (!m || !host) return fail("Error: null argument.");
    LBL_CUDA(cudaSetDevice(m->device));
    LBL_CUDA(cudaStreamWaitEvent(m->streams->copy, m->ev_added, 0));
    LBL_CUDA(cudaMemcpyAsync(host, m->acc.p, sizeof(double) * (size_t)m->n_layers * m->n,
                             cudaMemcpyDeviceToHost, m->streams->copy));
    LBL_CUDA(cudaEventRecord(m->ev_copied, m->streams->copy));
    LBL_CUDA(cudaEventSynchronize(m->ev_copied));
    m->copies_pending = false;
    return 0;
}

int lbl_mix_device_result(lbl_mix* m, double** device_ptr, long long* count)
{
    if (!m || !device_ptr || !count) return fail("Error: null argument.");
    *device_ptr = m->acc.as<double>();
    *count = (long long)m->n_layers * m->n;
    return 0;
}

int lbl_mix_close(lbl_mix* m)
{
    if (!m) return 0;
    cudaSetDevice(m->device);
    lbl_mix_wait(m);
    m->acc.release();
    if (m->ev_added) cudaEventDestroy(m->ev_added);
    if (m->ev_copied) cudaEventDestroy(m->ev_copied);
    delete m;
    return 0;
}

// ---- MT-CKD continua on the device ----------------------------------------------------------
}  // extern "C"

#include "lbl_continuum.cuh"

struct lbl_continuum
{
    int device = 0;
    DeviceStreams* streams = nullptr;
    struct Spectrum
    {
        double lower = 0., upper = 0., resolution = 0.;
        std::vector<double> data;
    };
    std::map<std::string, Spectrum> table;
    std::vector<std::unique_ptr<DevBuf>> arrays;   // coefficient arrays on the device
    std::map<std::string, ContinuumView> continua;
    DevBuf layers_dev, values_dev, out_dev;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};   // before K5a, between, after K5b
    bool timed = false;
    bool finalized = false;
};

namespace
{
// wavenumbers of a table variable, as utils.py:138-144 forms them
std::vector<double> wavenumbers(const lbl_continuum::Spectrum& s)
{
    std::vector<double> w(s.data.size());
    for (size_t i = 0; i < w.size(); ++i) w[i] = s.lower + (double)i * s.resolution;
    return w;
}

// `sub` laid over the grid of `host` (utils.py:64-81): first and last host index it covers.
int subgrid_bounds(const lbl_continuum::Spectrum& host, const lbl_continuum::Spectrum& sub, int* lower,
                   int* upper)
{
    if (host.resolution != sub.resolution) return fail("Error: grid and subgrid have different resolutions.");
    if (host.lower > sub.lower || host.upper < sub.upper) return fail("Error: subgrid not contained in grid.");
    *lower = (int)((sub.lower - host.lower) / host.resolution);
    *upper = (int)((sub.upper - host.lower) / host.resolution);
    if (*upper - *lower + 1 != (int)sub.data.size() || *upper >= (int)host.data.size())
    {
        return fail("Error: subgrid does not fit its grid.");
    }
    return 0;
}

int continuum_upload(lbl_continuum* c, const std::vector<double>& v, const double** out)
{
    c->arrays.emplace_back(new DevBuf);
    DevBuf& b = *c->arrays.back();
    LBL_CUDA(b.reserve(sizeof(double) * std::max<size_t>(v.size(), 1)));
    LBL_CUDA(cudaMemcpy(b.p, v.data(), sizeof(double) * v.size(), cudaMemcpyHostToDevice));
    *out = b.as<double>();
    return 0;
}

int continuum_band(lbl_continuum* c, ContinuumView& cv, int kind, double lower, double resolution,
                   std::initializer_list<const std::vector<double>*> coefficients)
{
    if (cv.n_bands >= kMaxBands) return fail("Error: too many bands.");
    BandView& b = cv.band[cv.n_bands++];
    b.kind = kind;
    b.n = (int)(*coefficients.begin())->size();
    b.lower = lower;
    b.resolution = resolution;
    b.inv_resolution = 1. / resolution;
    b.x_last = lower + (double)(b.n - 1) * resolution;
    b.value_offset = cv.row;
    cv.row += b.n;
    int k = 0;
    for (int q = 0; q < 4; ++q) b.c[q] = nullptr;
    for (const std::vector<double>* v : coefficients)
    {
        if ((int)v->size() != b.n) return fail("Error: coefficient arrays of a band differ in length.");
        if (continuum_upload(c, *v, &b.c[k++])) return 1;
    }
    return 0;
}
}  // namespace

extern "C" {

int lbl_continuum_create(int device, lbl_continuum** out)
{
    *out = nullptr;
    int ndev = 0;
    if (lbl_device_count(&ndev)) return 1;
    if (ndev == 0) return fail("Error: no CUDA device (" + g_last_error + "); this library has no CPU path.");
    if (device < 0 || device >= ndev) return fail("Error: CUDA device index out of range.");
    std::unique_ptr<lbl_continuum> c(new lbl_continuum);
    c->device = device;
    LBL_CUDA(cudaSetDevice(device));
    if (device_streams(device, &c->streams)) return 1;
    for (cudaEvent_t& e : c->ev) LBL_CUDA(cudaEventCreate(&e));
    *out = c.release();
    return 0;
}

int lbl_continuum_last_ms(lbl_continuum* c, float* bands_ms, float* apply_ms)
{
    if (!c || !bands_ms || !apply_ms) return fail("Error: null argument.");
    if (!c->timed) return fail("Error: no continuum call to report.");
    LBL_CUDA(cudaSetDevice(c->device));
    LBL_CUDA(cudaEventSynchronize(c->ev[2]));
    LBL_CUDA(cudaEventElapsedTime(bands_ms, c->ev[0], c->ev[1]));
    LBL_CUDA(cudaEventElapsedTime(apply_ms, c->ev[1], c->ev[2]));
    return 0;
}

int lbl_continuum_set_spectrum(lbl_continuum* c, const char* name, double lower, double upper,
                               double resolution, int count, const double* data)
{
    if (!c || !name || !data || count < 1) return fail("Error: invalid continuum spectrum.");
    if (c->finalized) return fail("Error: continuum table already finalized.");
    lbl_continuum::Spectrum& s = c->table[name];
    s.lower = lower;
    s.upper = upper;
    s.resolution = resolution;
    s.data.assign(data, data + count);
    return 0;
}

int lbl_continuum_finalize(lbl_continuum* c)
{
    if (!c) return fail("Error: null handle.");
    if (c->finalized) return 0;
    LBL_CUDA(cudaSetDevice(c->device));
    auto need = [&](const char* name, const lbl_continuum::Spectrum** out) {
        auto it = c->table.find(name);
        if (it == c->table.end()) return fail(std::string("Error: continuum table lacks ") + name + ".");
        *out = &it->second;
        return 0;
    };
    auto simple = [&](ContinuumView& cv, int kind, std::initializer_list<const char*> names) {
        std::vector<const lbl_continuum::Spectrum*> sp;
        for (const char* n : names)
        {
            const lbl_continuum::Spectrum* s = nullptr;
            if (need(n, &s)) return 1;
            sp.push_back(s);
        }
        std::vector<const std::vector<double>*> co;
        for (const lbl_continuum::Spectrum* s : sp) co.push_back(&s->data);
        if (cv.n_bands >= kMaxBands) return fail("Error: too many bands.");
        // initializer_list cannot be built at run time: add the band by hand
        BandView& b = cv.band[cv.n_bands++];
        b.kind = kind;
        b.n = (int)sp[0]->data.size();
        b.lower = sp[0]->lower;
        b.resolution = sp[0]->resolution;
        b.inv_resolution = 1. / b.resolution;
        b.x_last = b.lower + (double)(b.n - 1) * b.resolution;
        b.value_offset = cv.row;
        cv.row += b.n;
        for (int q = 0; q < 4; ++q) b.c[q] = nullptr;
        for (size_t k = 0; k < co.size(); ++k)
        {
            if ((int)co[k]->size() != b.n) return fail("Error: coefficient arrays of a band differ in length.");
            if (continuum_upload(c, *co[k], &b.c[k])) return 1;
        }
        return 0;
    };
    const lbl_continuum::Spectrum* s = nullptr;
    const lbl_continuum::Spectrum* x = nullptr;

    {   // CO2, carbon_dioxide.py:9-45
        ContinuumView cv{};
        if (need("bfco2", &s)) return 1;
        std::vector<double> tcorr(s->data.size(), 1.), xfac(s->data.size(), 1.);
        int lo = 0, hi = 0;
        if (need("tdep_bandhead", &x) || subgrid_bounds(*s, *x, &lo, &hi)) return 1;
        std::copy(x->data.begin(), x->data.end(), tcorr.begin() + lo);
        if (need("x_factor_co2", &x) || subgrid_bounds(*s, *x, &lo, &hi)) return 1;
        std::copy(x->data.begin(), x->data.end(), xfac.begin() + lo);
        if (continuum_band(c, cv, kCo2Hartmann, s->lower, s->resolution, {&s->data, &xfac, &tcorr})) return 1;
        c->continua["CO2"] = cv;
    }
    {   // H2O self, water_vapor.py:7-35
        ContinuumView cv{};
        if (simple(cv, kH2oSelf, {"bs296", "bs260"})) return 1;
        c->continua["H2OSelf"] = cv;
    }
    {   // H2O foreign, water_vapor.py:38-79
        ContinuumView cv{};
        if (need("bfh2o", &s) || need("xfac_rhu", &x)) return 1;
        int lo = 0, hi = 0;
        if (subgrid_bounds(*s, *x, &lo, &hi)) return 1;
        std::vector<double> scale(s->data.size(), 0.);
        for (int i = 1; i < (int)x->data.size(); ++i) scale[lo + i] = x->data[i];
        scale[lo] = scale[lo + 1];
        const std::vector<double> w = wavenumbers(*s);
        for (size_t i = (size_t)hi + 1; i < scale.size(); ++i)
        {
            const double vdelsq1 = (w[i] - 255.67) * (w[i] - 255.67);
            const double vf1 = std::pow((w[i] - 255.67) / 57.83, 8);
            const double vdelmsq1 = (w[i] + 255.67) * (w[i] + 255.67);
            const double vmf1 = std::pow((w[i] + 255.67) / 57.83, 8);
            const double vf2 = std::pow(w[i] / 630., 8);
            scale[i] = 1. + (0.06 - 0.42 * ((57600. / (vdelsq1 + 57600. + vf1)) +
                                            (57600. / (vdelmsq1 + 57600. + vmf1)))) / (1. + 0.3 * vf2);
        }
        if (continuum_band(c, cv, kH2oForeign, s->lower, s->resolution, {&s->data, &scale})) return 1;
        c->continua["H2OForeign"] = cv;
    }
    {   // N2, nitrogen.py:7-79
        ContinuumView cv{};
        if (simple(cv, kN2Rotation, {"ct_296", "ct_220", "sf_296", "sf_220"})) return 1;
        if (simple(cv, kN2Fundamental, {"xn2_272", "xn2_228", "a_h2o"})) return 1;
        if (simple(cv, kN2Overtone, {"xn2"})) return 1;
        c->continua["N2"] = cv;
    }
    {   // O2, oxygen.py:7-148
        ContinuumView cv{};
        if (simple(cv, kO2Fundamental, {"o2_f", "o2_t"})) return 1;
        if (simple(cv, kO2Nir, {"o2_inf1"})) return 1;
        {   // oxygen.py:57-71: analytic, on arange(9100., 11002., 2.)
            std::vector<double> data;
            const double hw1 = 58.96, hw2 = 45.04;
            for (int i = 0; 9100. + 2. * i < 11002.; ++i)
            {
                const double g = 9100. + 2. * i;
                const double dv1 = g - 9375., dv2 = g - 9439.;
                const double damp1 = dv1 < 0. ? std::exp(dv1 / 176.1) : 1.;
                const double damp2 = dv2 < 0. ? std::exp(dv2 / 176.1) : 1.;
                const double o2inf = 0.31831 * (((1.166e-04 * damp1 / hw1) / (1. + (dv1 / hw1) * (dv1 / hw1))) +
                                                ((3.086e-05 * damp2 / hw2) / (1. + (dv2 / hw2) * (dv2 / hw2)))) * 1.054;
                data.push_back(o2inf / g);
            }
            if (continuum_band(c, cv, kO2Nir2, 9100., 2., {&data})) return 1;
        }
        if (simple(cv, kO2Nir3, {"o2_inf3"})) return 1;
        if (simple(cv, kO2Visible, {"o2_invis"})) return 1;
        {   // oxygen.py:114-126: analytic, on arange(36000., 100010., 10.)
            std::vector<double> data;
            for (int i = 0; 36000. + 10. * i < 100010.; ++i)
            {
                const double g = 36000. + 10. * i;
                if (g <= 36000.)
                {
                    data.push_back(0.);
                    continue;
                }
                const double corr = g <= 40000. ? ((40000. - g) / 4000.) * 7.917e-7 : 0.;
                const double yratio = g / 48811.0;
                data.push_back(6.884e-4 * yratio * std::exp(-69.738 * std::pow(std::log(yratio), 2)) - corr);
            }
            if (continuum_band(c, cv, kO2Herzberg, 36000., 10., {&data})) return 1;
        }
        if (simple(cv, kO2Uv, {"o2_infuv"})) return 1;
        c->continua["O2"] = cv;
    }
    {   // O3, ozone.py:5-70
        ContinuumView cv{};
        if (simple(cv, kO3ChappuisWulf, {"x_o3", "y_o3", "z_o3"})) return 1;
        if (simple(cv, kO3HartleyHuggins, {"o3_hh0", "o3_hh1", "o3_hh2"})) return 1;
        if (simple(cv, kO3Uv, {"o3_huv"})) return 1;
        c->continua["O3"] = cv;
    }
    c->finalized = true;
    return 0;
}

int lbl_continuum_compute(lbl_continuum* c, const char* name, int n_layers, const double* temperature,
                          const double* pressure, const double* vmr6, int v0, int vn, int n_per_v,
                          lbl_mix* mix, int row0, double* k_host)
{
    if (!c || !name || !temperature || !pressure || !vmr6) return fail("Error: null argument.");
    if (!c->finalized) return fail("Error: continuum table not finalized.");
    // one name, or several separated by commas: their bands are summed in one pass over the output
    ContinuumView merged{};
    {
        std::string list(name);
        size_t pos = 0;
        while (pos <= list.size())
        {
            const size_t comma = std::min(list.find(',', pos), list.size());
            const std::string one = list.substr(pos, comma - pos);
            auto it = c->continua.find(one);
            if (it == c->continua.end()) return fail("Error: no continuum named " + one + ".");
            for (int b = 0; b < it->second.n_bands; ++b)
            {
                if (merged.n_bands >= kMaxBands) return fail("Error: too many bands in one continuum call.");
                BandView band = it->second.band[b];
                // the part of the band's own grid that the call's grid [v0, vn) can reach; a band
                // wholly outside it contributes zeros and is left out
                const double first = (double)v0, last = (double)vn;
                const double j_first = std::floor((first - band.lower) / band.resolution) - 1.;
                const double j_last = std::ceil((last - band.lower) / band.resolution) + 1.;
                if (j_last < 0. || j_first > (double)(band.n - 1)) continue;
                band.j_lo = (int)std::max(j_first, 0.);
                band.j_hi = (int)std::min(j_last, (double)(band.n - 1));
                band.value_offset = merged.row;
                merged.row += band.n;
                merged.band[merged.n_bands++] = band;
            }
            pos = comma + 1;
        }
    }
    if (n_layers < 1 || n_per_v < 1 || vn <= v0) return fail("Error: invalid grid or layer count.");
    const long long n_ll = (long long)(vn - v0) * n_per_v;
    if (n_ll > (1ll << 30)) return fail("Error: spectral grid too large for 32-bit indices.");
    const int n = (int)n_ll;
    if (mix && k_host) return fail("Error: a call feeds either the host array or the accumulator.");
    if (!mix && !k_host) return fail("Error: nowhere to put the continuum.");
    if (mix && (mix->device != c->device || mix->n != n || row0 < 0 || row0 + n_layers > mix->n_layers))
    {
        return fail("Error: accumulator shape does not fit this call.");
    }
    LBL_CUDA(cudaSetDevice(c->device));
    const ContinuumView& cv = merged;
    std::vector<ContinuumLayer> layers((size_t)n_layers);
    for (int l = 0; l < n_layers; ++l)
    {
        const double* x = vmr6 + 6 * (size_t)l;
        layers[l] = ContinuumLayer{temperature[l], pressure[l], x[0], x[1], x[2], x[3], x[4], x[5]};
    }
    cudaStream_t sl = c->streams->late;    // where the accumulator's additions are ordered
    LBL_CUDA(cudaStreamSynchronize(sl));   // the previous call's buffers are free
    LBL_CUDA(c->layers_dev.reserve(sizeof(ContinuumLayer) * (size_t)n_layers));
    LBL_CUDA(c->values_dev.reserve(sizeof(double) * 2 * (size_t)std::max(cv.row, 1) * n_layers));
    LBL_CUDA(cudaMemcpyAsync(c->layers_dev.p, layers.data(), sizeof(ContinuumLayer) * n_layers,
                             cudaMemcpyHostToDevice, sl));
    dim3 gb((std::max(cv.row, 1) + 127) / 128, n_layers);
    LBL_CUDA(cudaEventRecord(c->ev[0], sl));
    continuum_bands_kernel<<<gb, 128, 0, sl>>>(cv, c->layers_dev.as<ContinuumLayer>(), c->values_dev.as<double>());
    continuum_slopes_kernel<<<gb, 128, 0, sl>>>(cv, c->values_dev.as<double>());
    LBL_CUDA(cudaEventRecord(c->ev[1], sl));
    c->timed = true;
    const size_t total = (size_t)n_layers * n;
    if (n_layers > 65535) return fail("Error: more than 65535 layers in one continuum call.");
    const dim3 ga((n + 256 * kContinuumPoints - 1) / (256 * kContinuumPoints), n_layers);
    if (mix)
    {
        continuum_apply_kernel<true><<<ga, 256, 0, sl>>>(
            cv, c->values_dev.as<double>(), v0, 1. / n_per_v, 0, n,
            mix->acc.as<double>() + (size_t)row0 * n);
        LBL_CUDA(cudaGetLastError());
        LBL_CUDA(cudaEventRecord(c->ev[2], sl));
        LBL_CUDA(cudaEventRecord(mix->ev_added, sl));
        LBL_CUDA(cudaStreamSynchronize(sl));   // `layers` goes out of scope
        return 0;
    }
    LBL_CUDA(c->out_dev.reserve(sizeof(double) * total));
    continuum_apply_kernel<false><<<ga, 256, 0, sl>>>(cv, c->values_dev.as<double>(), v0, 1. / n_per_v, 0, n,
                                                      c->out_dev.as<double>());
    LBL_CUDA(cudaGetLastError());
    LBL_CUDA(cudaEventRecord(c->ev[2], sl));
    LBL_CUDA(cudaMemcpyAsync(k_host, c->out_dev.p, sizeof(double) * total, cudaMemcpyDeviceToHost, sl));
    LBL_CUDA(cudaStreamSynchronize(sl));
    return 0;
}

int lbl_continuum_close(lbl_continuum* c)
{
    if (!c) return 0;
    cudaSetDevice(c->device);
    for (auto& b : c->arrays) b->release();
    c->layers_dev.release();
    c->values_dev.release();
    c->out_dev.release();
    for (cudaEvent_t e : c->ev) if (e) cudaEventDestroy(e);
    delete c;
    return 0;
}

// ---- the reference's own entry point (absorption.c:19-30) ---------------------------------
int absorption(double pressure, double temperature, double volume_mixing_ratio, int v0, int vn,
               int n_per_v, double* k, char* database, char* formula, int cut_off,
               int remove_pedestal)
{
    static std::mutex mu;
    static std::map<std::string, lbl_gas*> cache;
    std::lock_guard<std::mutex> lock(mu);
    int device = 0;
    if (const char* env = getenv("PYLBL_B200_DEVICE")) device = atoi(env);
    const std::string key = std::string(database) + "\n" + formula + "\n" + std::to_string(device);
    lbl_gas* g = nullptr;
    auto it = cache.find(key);
    if (it == cache.end())
    {
        if (lbl_gas_open(database, formula, device, &g)) return 1;
        cache[key] = g;
    }
    else
    {
        g = it->second;
    }
    return lbl_gas_compute(g, 1, &pressure, &temperature, &volume_mixing_ratio, v0, vn, n_per_v,
                           cut_off, remove_pedestal, LBL_PRECISION_FP64, k);
}

}  // extern "C"
