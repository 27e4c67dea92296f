// lbl_continuum.cuh -- MT-CKD continua on the device.
//
// Replaces the host-side plugin pyLBL/mt_ckd (reference paths relative to /root/reference/pyLBL):
//   mt_ckd/utils.py:18-61      number densities and the radiation term
//   mt_ckd/utils.py:157-174    BandedContinuum.spectra: every band's spectrum on the band's own
//                              coarse grid, numpy.interp onto the caller's grid (0 outside), x100
//   mt_ckd/{carbon_dioxide,water_vapor,nitrogen,oxygen,ozone}.py   the band formulas
// called once per (gas, layer) by the driver (spectroscopy.py:194-198).  Here: one small kernel
// evaluates all bands of a continuum for all layers on their own grids (K5a), one HBM-bound
// kernel interpolates and adds them into the output (K5b: 8 bytes read + 8 written per point,
// or 16 + 8 into the gas-sum accumulator).
// Included by lbl_api.cu (same translation unit: it needs lbl_mix and the device streams).
#pragma once

namespace lbl
{

constexpr double kLoschmidt = 2.6867775e19;   // utils.py:7
constexpr double kP0 = 1013.25;               // utils.py:8  [mb]
constexpr double kT0 = 296.;                  // utils.py:10
constexpr double kT273 = 273.15;              // utils.py:11
constexpr int kMaxBands = 20;   // all six continua together: 1 + 1 + 1 + 3 + 7 + 3 = 16

enum BandKind
{
    kCo2Hartmann = 0, kH2oSelf, kH2oForeign, kN2Rotation, kN2Fundamental, kN2Overtone,
    kO2Fundamental, kO2Nir, kO2Nir2, kO2Nir3, kO2Visible, kO2Herzberg, kO2Uv,
    kO3ChappuisWulf, kO3HartleyHuggins, kO3Uv
};

struct BandView
{
    int kind, n;
    double lower, resolution;     // band grid w_j = lower + j*resolution (utils.py:138-144)
    double inv_resolution, x_last;   // x_last = lower + (n-1)*resolution
    int j_lo, j_hi;               // part of the band grid the call's wavenumber grid can reach
    const double* c[4];           // coefficient arrays on the band grid
    int value_offset;             // of this band in the per-layer value row
};

struct ContinuumView
{
    int n_bands;
    int row;                      // doubles per layer in the value array (all bands)
    BandView band[kMaxBands];
};

// Per-layer state: temperature [K], pressure [Pa], mole fractions H2O, CO2, O3, N2, O2 and the sum
// over every gas of the atmosphere (air_number_density sums them all, utils.py:18-30).
struct ContinuumLayer
{
    double t, p, h2o, co2, o3, n2, o2, sum_x;
};

__device__ __forceinline__ double radiation_term(double w, double t)   // utils.py:47-61
{
    const double x = w / (t / kC2);
    if (x <= 10.)
    {
        const double e = exp(-x);
        return w * (1. - e) / (1. + e);
    }
    return w;
}

// One band of one layer at its grid point j (the band modules' spectra(), pressure in mb).
__device__ double band_value(const BandView& b, const ContinuumLayer& ly, int j)
{
    const double t = ly.t;
    const double p = ly.p * 0.01;                                            // utils.py:13,172
    const double w = b.lower + (double)j * b.resolution;
    const double ndry = kLoschmidt * (p / kP0) * (kT273 / t) * (1. - ly.h2o);   // utils.py:33-44
    const double nair = ndry * ly.sum_x;                                     // utils.py:18-30
    const double rad = radiation_term(w, t);
    const double c0 = b.c[0][j];
    switch (b.kind)
    {
        case kCo2Hartmann:      // carbon_dioxide.py:36-42
        {
            const double n = ndry * ly.co2;
            return n * 1.e-20 * (p / kP0) * (kT0 / t) * rad * b.c[1][j] * pow(t / 246., b.c[2][j]) * c0;
        }
        case kH2oSelf:          // water_vapor.py:22-31
        {
            const double tf = (t - kT0) / (260. - kT0);
            const double nh2o = ndry * ly.h2o;
            return nh2o * (nh2o / nair) * (p / kP0) * (kT0 / t) * 1.e-20 * rad * c0 * pow(b.c[1][j] / c0, tf);
        }
        case kH2oForeign:       // water_vapor.py:69-76
        {
            const double nh2o = ndry * ly.h2o;
            return (1. - (nh2o / nair)) * (p / kP0) * (kT0 / t) * 1.e-20 * nh2o * rad * b.c[1][j] * c0;
        }
        case kN2Rotation:       // nitrogen.py:20-35
        {
            const double tau = ((ndry * ly.n2) / kLoschmidt) * (p / kP0) * (kT273 / t);
            const double f = (t - kT0) / (220. - kT0);
            const double c = c0 * pow(b.c[1][j] / c0, f);
            const double s = b.c[2][j] * pow(b.c[3][j] / b.c[2][j], f);
            const double fo2 = (s - 1.) * ly.n2 / ly.o2;
            return tau * rad * c * (ly.n2 + fo2 * ly.o2 + ly.h2o);
        }
        case kN2Fundamental:    // nitrogen.py:46-60
        {
            const double tau = ((ndry * ly.n2) / kLoschmidt) * (p / kP0) * (kT273 / t);
            const double xt = (1. / t - 1. / 272.) / (1. / 228. - 1. / 272.);
            const double ao2 = 1.294 - 0.4545 * t / kT0;
            double k0 = 0.;
            if (j > 0 && j < b.n - 1) k0 = c0 * pow(b.c[1][j] / c0, xt);
            k0 = k0 / w;
            const double k1 = ao2 * k0;
            const double k2 = (9. / 7.) * b.c[2][j] * k0;
            return tau * rad * (k0 * ly.n2 + ly.o2 * k1 + ly.h2o * k2);
        }
        case kN2Overtone:       // nitrogen.py:70-76
        {
            const double tau = ((ndry * ly.n2) / kLoschmidt) * (p / kP0) * (kT273 / t) * (ly.n2 + ly.o2 + ly.h2o);
            return tau * rad * c0 / w;
        }
        case kO2Fundamental:    // oxygen.py:23-32
        {
            const double tau = (ndry * ly.o2) * 1.e-20 * (p / kP0) * (kT273 / t);
            const double xkt = (1. / kT0) - (1. / t);
            return tau * rad * (1.e20 / kLoschmidt) * c0 * exp(b.c[1][j] * xkt) / w;
        }
        case kO2Nir:            // oxygen.py:42-51
        {
            const double tau = ((ndry * ly.o2) / kLoschmidt) * (p / kP0) * (kT273 / t) *
                               ((1. / 0.446) * ly.o2 + (0.3 / 0.446) * ly.n2 + ly.h2o);
            return tau * rad * c0 / w;
        }
        case kO2Nir2:           // oxygen.py:73-78
        {
            const double no2 = ndry * ly.o2;
            const double adj = (no2 / nair) * (1. / ly.o2) * no2 * 1.e-20 * (p / kP0) * (kT0 / t);
            return adj * rad * c0;
        }
        case kO2Nir3:           // oxygen.py:88-92
        {
            const double tau = ((ndry * ly.o2) / kLoschmidt) * (p / kP0) * (kT273 / t);
            return tau * rad * c0 / w;
        }
        case kO2Visible:        // oxygen.py:102-108
        {
            const double no2 = ndry * ly.o2;
            const double adj = (no2 / nair) * no2 * 1.e-20 * (p / kP0) * (kT273 / t);
            const double factor = 1. / (kLoschmidt * 1.e-20 * (55. * kT273 / kT0) * (55. * kT273 / kT0) * 89.5);
            return adj * rad * factor * c0 / w;
        }
        case kO2Herzberg:       // oxygen.py:128-132
        {
            const double factor = 1. + 0.83 * (p / kP0) * (kT273 / t);
            return 1.e-20 * (ndry * ly.o2) * rad * factor * c0 / w;
        }
        case kO2Uv:             // oxygen.py:142-145
            return 1.e-20 * (ndry * ly.o2) * rad * c0 / w;
        case kO3ChappuisWulf:   // ozone.py:22-27
        {
            const double dt = t - kT273;
            return 1.e-20 * (ndry * ly.o3) * rad * (c0 + b.c[1][j] * dt + b.c[2][j] * dt * dt) / w;
        }
        case kO3HartleyHuggins: // ozone.py:43-50
        {
            const double dt = t - kT273;
            return 1.e-20 * (ndry * ly.o3) * rad * (c0 / w) * (1. + b.c[1][j] * dt + b.c[2][j] * dt * dt);
        }
        default:                // kO3Uv, ozone.py:64-67
            return (ndry * ly.o3) * rad * c0 / w;
    }
}

// K5a.  values[layer][band offset + j] (2*row doubles per layer: values, then slopes), for the
// part [j_lo, j_hi] of each band that the call's grid can reach.  grid = (ceil(row/128), layers).
__global__ void __launch_bounds__(128)
continuum_bands_kernel(const ContinuumView cv, const ContinuumLayer* __restrict__ layers,
                       double* __restrict__ values)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= cv.row)
    {
        return;
    }
    int b = 0;
    while (b + 1 < cv.n_bands && idx >= cv.band[b + 1].value_offset) ++b;
    const int j = idx - cv.band[b].value_offset;
    if (j < cv.band[b].j_lo || j > cv.band[b].j_hi)
    {
        return;
    }
    values[(size_t)blockIdx.y * 2 * cv.row + idx] = band_value(cv.band[b], layers[blockIdx.y], j);
}

// K5a'.  slopes[layer][band offset + j] = (f[j+1] - f[j]) / (x[j+1] - x[j]), the slope numpy.interp
// uses on [x_j, x_j+1) (numpy precomputes them the same way); stored behind the values of the
// layer.  grid = (ceil(row/128), layers).
__global__ void __launch_bounds__(128)
continuum_slopes_kernel(const ContinuumView cv, double* __restrict__ values)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= cv.row)
    {
        return;
    }
    int b = 0;
    while (b + 1 < cv.n_bands && idx >= cv.band[b + 1].value_offset) ++b;
    const BandView& band = cv.band[b];
    const int j = idx - band.value_offset;
    if (j < band.j_lo || j >= band.j_hi)
    {
        return;
    }
    double* f = values + (size_t)blockIdx.y * 2 * cv.row;
    const double xj = band.lower + (double)j * band.resolution;
    const double xn = band.lower + (double)(j + 1) * band.resolution;
    f[cv.row + idx] = __ddiv_rn(__dsub_rn(f[idx + 1], f[idx]), __dsub_rn(xn, xj));
}

// The continuum is piecewise linear in the wavenumber: between two consecutive nodes of the
// bands' own grids, sum_b 100*interp_b(x) = alpha + beta*(x - origin).  `Segment` holds that line
// (origin = the wavenumber it was set up at, so that nothing large cancels) and the wavenumber
// up to which it is valid.
struct Segment
{
    double alpha, beta, origin, valid_below;
};

// numpy.interp(x, xp, fp, left=0, right=0) semantics for xp[j] = lower + j*res, accumulated over
// the bands: exact node hits and the last node return fp[j] (a constant piece), outside a band 0.
__device__ __forceinline__ Segment continuum_segment(const ContinuumView& cv, const double* __restrict__ f,
                                                     double x)
{
    Segment s;
    s.alpha = 0.;
    s.beta = 0.;
    s.origin = x;
    s.valid_below = 1.0e300;
    for (int b = 0; b < cv.n_bands; ++b)
    {
        const BandView& band = cv.band[b];
        const int last = band.n - 1;
        if (!(x >= band.lower))
        {
            s.valid_below = fmin(s.valid_below, band.lower);      // the band starts further up
            continue;
        }
        if (!(x <= band.x_last))
        {
            continue;
        }
        int j = (int)((x - band.lower) * band.inv_resolution);
        j = j < 0 ? 0 : (j > last ? last : j);
        // largest j with xp[j] <= x (the guess above can be one off next to a grid point)
        while (j < last && band.lower + (double)(j + 1) * band.resolution <= x) ++j;
        while (j > 0 && band.lower + (double)j * band.resolution > x) --j;
        const double xj = band.lower + (double)j * band.resolution;
        const double* fb = f + band.value_offset;
        if (j == last || xj == x)
        {
            // a node: this very point takes fp[j]; the next grid point starts a new segment
            s.alpha += 100. * fb[j];
            s.valid_below = x;
            continue;
        }
        const double slope = fb[cv.row + j];
        s.alpha += 100. * __dadd_rn(__dmul_rn(slope, __dsub_rn(x, xj)), fb[j]);   // numpy's own value at x
        s.beta += 100. * slope;
        s.valid_below = fmin(s.valid_below, band.lower + (double)(j + 1) * band.resolution);
    }
    return s;
}

// K5b.  dst[layer][k] (+)= sum over the bands of 100*interp(v) (utils.py:168-174) for the grid
// points p_lo + k, k < width, of layer blockIdx.y; kPoints consecutive points per thread, which
// mostly share one segment.  kAdd: into the gas-sum accumulator.  HBM-bound: 8 bytes written per
// point, and 8 read when adding.
constexpr int kContinuumPoints = 4;

template <bool kAdd>
__global__ void __launch_bounds__(256)
continuum_apply_kernel(const ContinuumView cv, const double* __restrict__ values, int v0, double dv,
                       int p_lo, int width, double* __restrict__ dst)
{
    constexpr int P = kContinuumPoints;
    const double* f = values + (size_t)blockIdx.y * 2 * cv.row;
    double* out = dst + (size_t)blockIdx.y * width;
    const bool vector = (width % 2 == 0) && ((reinterpret_cast<size_t>(out) & 15) == 0);
    for (int k0 = (blockIdx.x * blockDim.x + threadIdx.x) * P; k0 < width; k0 += gridDim.x * blockDim.x * P)
    {
        double val[P];
        Segment seg;
        seg.valid_below = -1.0e300;
        seg.alpha = seg.beta = seg.origin = 0.;
#pragma unroll
        for (int q = 0; q < P; ++q)
        {
            const double v = grid_point(v0, dv, p_lo + min(k0 + q, width - 1));
            if (!(v < seg.valid_below))
            {
                seg = continuum_segment(cv, f, v);
            }
            val[q] = fma(seg.beta, v - seg.origin, seg.alpha);
        }
        if (vector && k0 + P <= width)
        {
            double2* o2 = reinterpret_cast<double2*>(out + k0);
#pragma unroll
            for (int q = 0; q < P; q += 2)
            {
                double2 w = kAdd ? o2[q / 2] : make_double2(0., 0.);
                w.x += val[q];
                w.y += val[q + 1];
                o2[q / 2] = w;
            }
        }
        else
        {
#pragma unroll
            for (int q = 0; q < P; ++q)
            {
                if (k0 + q < width) out[k0 + q] = kAdd ? out[k0 + q] + val[q] : val[q];
            }
        }
    }
}

}  // namespace lbl
