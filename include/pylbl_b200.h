/* pylbl_b200.h -- C ABI of the B200 line-by-line absorption library (libpylbl_b200.so).
 *
 * This is the drop-in boundary for pyLBL's lines backend.  The reference reaches its C
 * library through ctypes (pyLBL/c_lib/gas_optics.py:11-12,68-91) and that library exports
 * exactly one entry point, absorption() (pyLBL/c_lib/absorption.c:19-30).  This header
 * declares
 *   (1) that same symbol with that same signature, and
 *   (2) a handle-based, layer-batched form of it, which is what the Python `Gas` class of
 *       this repository (pylbl_b200/gas_optics.py) binds.
 * Only plain C types cross the boundary.  Every function returns 0 on success and 1 on
 * error (the reference's convention, absorption.c:11-15, spectral_database.c:11-16); the
 * message is printed to stderr like the reference does and kept for lbl_last_error().
 * No function throws.  There is no CPU fallback: without a CUDA device every compute
 * entry point fails with 1.
 */
#ifndef PYLBL_B200_H_
#define PYLBL_B200_H_

#include <stddef.h>

#if defined(__GNUC__)
#define LBL_API __attribute__((visibility("default")))
#else
#define LBL_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define LBL_PRECISION_FP64 0
#define LBL_PRECISION_FP32 1

/* One molecule of one spectral database, packed into device memory. */
typedef struct lbl_gas lbl_gas;

/* Statistics of the most recent lbl_gas_compute()/lbl_gas_wait() on a handle. */
typedef struct lbl_stats
{
    long long evals;        /* sum over layers and processed lines of (e-s+1), spectra.c:48-62 */
    long long executed;     /* far-wing evaluations the summation kernel actually performed (the
                               far-field interpolation makes this smaller than `evals`) */
    long long h2d_bytes;    /* bytes copied host->device by the call (layer states, first-use packing) */
    long long d2h_bytes;    /* bytes copied device->host by the call (spectra) */
    int n_lines;            /* transition rows of the molecule */
    int n_active;           /* rows before the reference's early break (absorption.c:80-83) */
    int n_layers;
    int n_points;           /* output points per layer: (vn - v0)*n_per_v, or the band's share */
    int points_per_thread;
    int cells_per_warp;     /* > 0: the cell-tiled summation kernel with the far-field interpolation ran */
    int sum_launches;       /* launches of the summation kernel */
    int total_launches;     /* all kernel launches */
    int fp32_used;          /* 1: the far wings were summed in FP32 (LBL_PRECISION_FP32 on the direct
                               kernel); 0: in FP64 (always on the far-field kernel) */
    float scale_ms;         /* K1, CUDA events on the launching stream */
    float sum_ms;           /* K2 (dominant kernel), summed over launches */
    float fixup_ms;         /* K2b (near-zone and node terms) */
    float pedestal_ms;      /* K3 + K4 */
    float total_ms;         /* first enqueue -> last copy complete */
    float copy_tail_ms;     /* last kernel of the call done -> last copy to the host complete: the
                               part of the device->host copy that nothing of this call hides */
} lbl_stats;

/* ---- (1) the reference's own entry point ---------------------------------------------
 * Replaces: pyLBL/c_lib/absorption.c:19-30 (same argument order and meaning).
 * k is caller-allocated, (vn-v0)*n_per_v doubles, overwritten (absorption.c:41).
 * Handles are cached per (database, formula) inside the library, so the sqlite read and
 * the device upload happen on the first call only.  Uses CUDA device 0 unless the
 * environment variable PYLBL_B200_DEVICE says otherwise.
 */
LBL_API int absorption(double pressure, double temperature, double volume_mixing_ratio,
               int v0, int vn, int n_per_v, double* k, char* database, char* formula,
               int cut_off, int remove_pedestal);

/* ---- (2) handle-based, batched form --------------------------------------------------
 * lbl_gas_open replaces the per-call open_database/molecule_id/tips_data/mass_data/
 * line_parameters sequence (absorption.c:45-79, spectral_database.c:19-180): it runs the
 * same four queries once and packs the rows into sorted structure-of-arrays buffers on
 * CUDA device `device`.  A molecule without TIPS rows opens successfully and computes
 * all-zero spectra (absorption.c:53-59).  Unknown molecule -> 1 (spectral_database.c:152).
 */
LBL_API int lbl_gas_open(const char* database, const char* formula, int device, lbl_gas** out);
LBL_API int lbl_gas_close(lbl_gas* gas);

/* Packed line-list cache.  The reference re-reads the sqlite file on
 * every absorption() call (absorption.c:45-79); lbl_gas_open reads it once per handle;
 * lbl_pack_database writes what that read produced -- TIPS table, isotopologue masses and the
 * transition rows in DATABASE ROW ORDER (the early break absorption.c:80-83 and the pedestal
 * spectra.c:66-78 depend on it) -- to one flat, checksummed binary file, and
 * lbl_gas_open_pack opens a handle from such a file without touching sqlite.  Spectra from
 * either kind of handle are bit-identical.  lbl_pack_database and lbl_pack_info need no GPU.
 * lbl_pack_info: any output pointer may be NULL; source_size/source_mtime are the size and
 * modification time of the sqlite file the pack was made from (for staleness checks).
 */
LBL_API int lbl_pack_database(const char* database, const char* formula, const char* pack_path);
LBL_API int lbl_pack_info(const char* pack_path, char* formula, int formula_capacity,
                          long long* n_lines, int* num_iso, int* num_t, int* sorted,
                          long long* source_size, long long* source_mtime);
LBL_API int lbl_gas_open_pack(const char* pack_path, int device, lbl_gas** out);

/* Layer-batched absorption(): for layer L, k_host[L*n .. L*n+n) receives what the
 * reference's absorption(pressure[L], temperature[L], vmr[L], v0, vn, n_per_v, ...) writes
 * to k.  n = (vn-v0)*n_per_v.  k_host may be pageable or pinned (lbl_host_alloc) memory; it
 * may be NULL, in which case the spectra stay on the device (lbl_gas_device_result).
 * precision: LBL_PRECISION_FP64 (parity target 1e-9) or LBL_PRECISION_FP32 (opt-in, 1e-4).
 * lbl_gas_compute = lbl_gas_submit + lbl_gas_wait.  After lbl_gas_submit returns, work is
 * enqueued on the handle's CUDA streams; k_host must stay valid until lbl_gas_wait.
 */
LBL_API int lbl_gas_compute(lbl_gas* gas, int n_layers, const double* pressure,
                    const double* temperature, const double* volume_mixing_ratio,
                    int v0, int vn, int n_per_v, int cut_off, int remove_pedestal,
                    int precision, double* k_host);
LBL_API int lbl_gas_submit(lbl_gas* gas, int n_layers, const double* pressure,
                   const double* temperature, const double* volume_mixing_ratio,
                   int v0, int vn, int n_per_v, int cut_off, int remove_pedestal,
                   int precision, double* k_host);
LBL_API int lbl_gas_wait(lbl_gas* gas);

/* Spectral band of a grid (band sharding of one spectrum over several GPUs).
 * Computes the output cells [band_lo, band_hi) -- integer wavenumbers v0+band_lo .. v0+band_hi --
 * of the grid (v0, vn, n_per_v): k_host[L*m .. L*m+m), m = (band_hi-band_lo)*n_per_v, receives
 * points [band_lo*n_per_v, band_hi*n_per_v) of what lbl_gas_submit would write for layer L, bit
 * for bit.  The line windows (spectra.c:48-62), the active-line prefix before the reference's
 * early break (absorption.c:80-83) and the accumulated pedestal (spectra.c:66-78) are those of
 * the WHOLE grid: a separate reference call on the band's own (v0, vn) would see a different
 * prefix (typically none) and different pedestals.  Bands of one grid are
 * independent: one handle per device, no communication.  lbl_stats.evals counts the band's
 * share of every line window. */
LBL_API int lbl_gas_submit_band(lbl_gas* gas, int n_layers, const double* pressure,
                   const double* temperature, const double* volume_mixing_ratio,
                   int v0, int vn, int n_per_v, int cut_off, int remove_pedestal,
                   int precision, int band_lo, int band_hi, double* k_host,
                   long long k_pitch);
/* k_pitch: doubles between the starts of consecutive host rows (0 = dense, m); with
 * k_pitch = (vn-v0)*n_per_v and k_host pointing at column band_lo*n_per_v of a whole-grid array,
 * the bands of several devices land side by side in one array.
 * lbl_gas_band_edges proposes n_bands contiguous bands for this molecule such that the costliest
 * band is as cheap as possible (a band costs its cells -- window lines, direct lines and near
 * zones, which widen with the wavenumber -- plus the pedestal recurrence over the rows before
 * its end; the model assumes the pedestal is removed): edges[0..n_bands] are cell indices,
 * band b = [edges[b], edges[b+1]), non-empty while there are cells.  Needs no GPU work. */
LBL_API int lbl_gas_band_edges(lbl_gas* gas, int v0, int vn, int n_per_v, int cut_off, int n_bands,
                   int* edges);

/* Layer groups for the copy back to the host (fine grids, k_host != NULL): the layers of a
 * call are summed, corrected and copied out in `groups` consecutive groups, so that all but
 * the last group's copy overlaps later kernels.  Every extra group costs a kernel tail; it
 * pays for a call nothing is queued behind (a blocking call, or the last gas of a column).
 * 0 = automatic: two groups for lbl_gas_compute, one for lbl_gas_submit.
 */
LBL_API int lbl_gas_set_copy_groups(lbl_gas* gas, int groups);

/* Results of the last call. */
LBL_API int lbl_gas_stats(lbl_gas* gas, lbl_stats* out);
/* Device pointer to the spectra of the last chunk of layers ([layers][n] doubles). */
LBL_API int lbl_gas_device_result(lbl_gas* gas, double** device_ptr, long long* count);
/* Window indices (s, e) of spectra.c:48-62 for every processed line of `layer`, in
 * database row order; -1,-1 for lines the reference skips (s >= n).  capacity >= n_active. */
LBL_API int lbl_gas_windows(lbl_gas* gas, int layer, int* s, int* e, int capacity);
/* Scaled line parameters of `layer` (spectra.c:17-45) in database row order:
 * out[4*r + {0,1,2,3}] = nu', alpha, gamma, sw'.  capacity >= n_active. */
LBL_API int lbl_gas_scaled(lbl_gas* gas, int layer, double* out, int capacity);

/* ---- gas-summed absorption on the device ("next" step after the path) -------------------
 * Replaces the host-side  beta = n * k[:grid.size]  and the sum over gases of
 * pyLBL/spectroscopy.py:181-191,225-234 (output_format="total", lines mechanism): the
 * spectra of several gases stay on the device, are scaled per layer and summed there, and
 * one array comes back instead of one per gas.
 *   lbl_mix_open    allocates a zeroed accumulator of n_layers*n_points doubles on `device`
 *   lbl_mix_reset   zeroes it again (asynchronous)
 *   lbl_gas_submit_mix  like lbl_gas_submit with k_host == NULL, and then, on the device,
 *                   acc[row0+L][i] += scale[L] * k[L][i]   (scale = number density p*x/(kB*T)
 *                   gives beta in m-1, spectroscopy.py:18-29); nothing waits on the host, so the
 *                   gases of a column are all in flight together; the additions of successive
 *                   calls run in submission order (no atomics).  total_host != NULL marks the
 *                   LAST gas of the sum: each layer group of the accumulator is copied to
 *                   total_host[(row0+L)*n_points ..] as soon as this gas has been added to it,
 *                   while later groups still compute
 *   lbl_mix_wait    blocks until every addition and copy enqueued so far has finished
 *   lbl_mix_add     adds the spectra a gas's last call left resident on the device (k_host ==
 *                   NULL, one layer group); blocking
 *   lbl_mix_download  copies the whole accumulator to host memory (blocking)
 *   lbl_mix_device_result  device pointer to the accumulator */
typedef struct lbl_mix lbl_mix;
LBL_API int lbl_mix_open(int device, int n_layers, int n_points, lbl_mix** out);
LBL_API int lbl_mix_reset(lbl_mix* mix);
LBL_API int lbl_gas_submit_mix(lbl_gas* gas, int n_layers, const double* pressure,
                   const double* temperature, const double* volume_mixing_ratio,
                   int v0, int vn, int n_per_v, int cut_off, int remove_pedestal,
                   int precision, lbl_mix* mix, int row0, const double* scale,
                   double* total_host);
LBL_API int lbl_mix_wait(lbl_mix* mix);
LBL_API int lbl_mix_add(lbl_mix* mix, lbl_gas* gas, const double* scale);
LBL_API int lbl_mix_download(lbl_mix* mix, double* host);
LBL_API int lbl_mix_device_result(lbl_mix* mix, double** device_ptr, long long* count);
LBL_API int lbl_mix_close(lbl_mix* mix);

/* ---- MT-CKD continua on the device (the other mechanism the driver adds per gas) -------------
 * Replaces the host-side plugin pyLBL/mt_ckd, i.e. what spectroscopy.py:194-198 calls per
 * (gas, layer): BandedContinuum.spectra(T, p, vmr, grid) (mt_ckd/utils.py:157-174) -- every
 * band's formula (mt_ckd/carbon_dioxide.py, water_vapor.py, nitrogen.py, oxygen.py, ozone.py) on
 * the band's own coarse grid, numpy.interp onto the caller's grid (zero outside the band), times
 * 100 -- for all layers at once, on the grid (v0, vn, n_per_v) of the lines calls.
 *   lbl_continuum_create / _set_spectrum / _finalize: the coefficient table, variable by
 *       variable as the reference's file names them (bfco2, bs296, ..., with the wavenumber_*
 *       attributes of each), then the band tables are built on the device.  No file format of
 *       its own: the host side reads pylbl_b200/data/mt_ckd.npz (tools/convert_mt_ckd.py).
 *   lbl_continuum_compute: continuum `name` ("CO2", "H2OForeign", "H2OSelf", "N2", "O2", "O3";
 *       or several, comma-separated: their sum, in one pass over the output)
 *       in m-1 for n_layers states; vmr6[L*6 ..] = mole fractions of H2O, CO2, O3, N2, O2 and
 *       the sum over ALL gases of the atmosphere (utils.py:18-30); pressure in Pa.  Either
 *       k_host[L*n ..] receives it, or (mix != NULL) it is added to rows row0+L of the
 *       device-side accumulator.  Blocking. */
typedef struct lbl_continuum lbl_continuum;
LBL_API int lbl_continuum_create(int device, lbl_continuum** out);
LBL_API int lbl_continuum_set_spectrum(lbl_continuum* c, const char* name, double lower_bound,
                   double upper_bound, double resolution, int count, const double* data);
LBL_API int lbl_continuum_finalize(lbl_continuum* c);
LBL_API int lbl_continuum_compute(lbl_continuum* c, const char* name, int n_layers,
                   const double* temperature, const double* pressure, const double* vmr6,
                   int v0, int vn, int n_per_v, lbl_mix* mix, int row0, double* k_host);
/* CUDA-event durations of the last call's two kernels: the band formulas (K5a) and the
 * interpolate-and-add pass over the output (K5b, HBM-bound). */
LBL_API int lbl_continuum_last_ms(lbl_continuum* c, float* bands_ms, float* apply_ms);
LBL_API int lbl_continuum_close(lbl_continuum* c);

/* Pinned host memory for k_host (lets the device->host copy run asynchronously). */
LBL_API int lbl_host_alloc(size_t bytes, void** ptr);
LBL_API int lbl_host_free(void* ptr);

/* Device-side stopwatch over several handles (one per CUDA device): lbl_timer_start marks
 * "now" on the device, lbl_timer_join makes the stopwatch wait for the end of the last call
 * submitted on `gas`, lbl_timer_stop returns the elapsed milliseconds between the two marks
 * as measured by CUDA events (it blocks until the joined work has finished). */
LBL_API int lbl_timer_start(int device);
LBL_API int lbl_timer_join(lbl_gas* gas);
LBL_API int lbl_timer_stop(int device, float* ms);
/* Measures this device's FP64 FMA throughput with independent DFMA chains (the roofline
 * denominator of the summation kernel); result in TFLOP/s (2 flop per FMA). */
LBL_API int lbl_measure_fp64_peak(int device, double* tflops);

LBL_API int lbl_device_count(int* count);
/* Layers per launch group: 0 = automatic. */
LBL_API int lbl_set_chunk_layers(int layers);
LBL_API const char* lbl_last_error(void);
LBL_API int lbl_version(void);

#ifdef __cplusplus
}
#endif

#endif /* PYLBL_B200_H_ */
