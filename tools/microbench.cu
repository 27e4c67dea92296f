// microbench.cu -- measures the B200 denominators this path is judged against:
//   * FP64 DFMA peak (independent FMA chains, CUDA events)            -> roofline "peak"
//   * MUFU.RCP64H (rcp.approx.ftz.f64) throughput and accuracy        -> seed of far_term()
//   * far_term() issue rate with operands in registers                -> kernel ceiling
// MEASURED_PEAKS.json carries HBM and bf16 numbers only; the FP64 figure has to be taken
// on the box.  Prints one JSON object per line.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/microbench tools/microbench.cu
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../pylbl_b200/csrc/lbl_core.cuh"

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int kChains = 8;

__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b)
{
    double x[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c) x[c] = threadIdx.x * 1e-3 + c;
    for (int i = 0; i < iters; ++i)
    {
#pragma unroll
        for (int c = 0; c < kChains; ++c) x[c] = fma(x[c], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < kChains; ++c) s += x[c];
    if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) rcp_kernel(double* out, int iters)
{
    double x[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c) x[c] = 1.5 + threadIdx.x * 1e-3 + c;
    for (int i = 0; i < iters; ++i)
    {
#pragma unroll
        for (int c = 0; c < kChains; ++c) x[c] = lbl::rcp_seed(x[c]);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < kChains; ++c) s += x[c];
    if (s == 12345.678) out[0] = s;
}

// far_term with P accumulators per thread and per-"line" operands read through L1 like K2.
template <int P, int STAGED>
__global__ void __launch_bounds__(128) far_kernel(const double2* __restrict__ ab,
                                                  const double* __restrict__ cc, int n_lines,
                                                  double* out)
{
    double v[P], acc[P];
#pragma unroll
    for (int p = 0; p < P; ++p)
    {
        v[p] = 1000.0 + (blockIdx.x * 128 + threadIdx.x) * P * 0.01 + p * 0.01;
        acc[p] = 0.;
    }
    if (STAGED == 2)
    {
        for (int j = 0; j + 1 < n_lines; j += 2)
        {
            const double2 l1 = __ldg(ab + j);
            const double c1 = __ldg(cc + j);
            const double2 l2 = __ldg(ab + j + 1);
            const double c2 = __ldg(cc + j + 1);
            lbl::far_terms_pair<P>(v, l1.x, l1.y, c1, l2.x, l2.y, c2, acc);
        }
    }
    else
#pragma unroll 2
    for (int j = 0; j < n_lines; ++j)
    {
        const double2 l = __ldg(ab + j);
        const double c = __ldg(cc + j);
        if (STAGED)
        {
            lbl::far_terms<P>(v, l.x, l.y, c, acc);
        }
        else
        {
#pragma unroll
            for (int p = 0; p < P; ++p) acc[p] = lbl::far_term(v[p], l.x, l.y, c, acc[p]);
        }
    }
    double s = 0;
#pragma unroll
    for (int p = 0; p < P; ++p) s += acc[p];
    out[blockIdx.x * 128 + threadIdx.x] = s;
}

__global__ void rcp_accuracy_kernel(const double* q, int n, double* err_seed, double* err_newton)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = q[i];
    const double r0 = lbl::rcp_seed(x);
    const double exact = 1.0 / x;
    const double r1 = r0 * fma(-x, r0, 2.0);
    err_seed[i] = fabs(r0 - exact) / exact;
    err_newton[i] = fabs(r1 - exact) / exact;
}

template <class F>
static float time_ms(F launch, int reps)
{
    cudaEvent_t a, b;
    CHECK(cudaEventCreate(&a));
    CHECK(cudaEventCreate(&b));
    launch();
    launch();
    CHECK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r)
    {
        CHECK(cudaEventRecord(a));
        launch();
        CHECK(cudaEventRecord(b));
        CHECK(cudaEventSynchronize(b));
        float ms;
        CHECK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    return best;
}

template <int P, int STAGED>
static void bench_far(const double2* ab, const double* cc, int n_lines, double* out, int sms)
{
    const int blocks = sms * 16;
    float ms = time_ms([&] { far_kernel<P, STAGED><<<blocks, 128>>>(ab, cc, n_lines, out); }, 5);
    const double evals = (double)blocks * 128 * P * n_lines;
    printf("{\"bench\": \"far_term\", \"points_per_thread\": %d, \"staged\": %d, \"ms\": %.4f, \"evals_per_s\": %.4e, "
           "\"dfma_per_s\": %.4e}\n", P, (int)STAGED, ms, evals / (ms * 1e-3), 4.0 * evals / (ms * 1e-3));
}

int main()
{
    cudaDeviceProp prop;
    CHECK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int clock_khz = 0;
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
    printf("{\"bench\": \"device\", \"name\": \"%s\", \"sms\": %d, \"max_clock_mhz\": %.0f}\n",
           prop.name, sms, clock_khz / 1000.0);
    double* out;
    CHECK(cudaMalloc(&out, sizeof(double) * sms * 16 * 128));

    {
        const int blocks = sms * 8, iters = 20000;
        float ms = time_ms([&] { dfma_kernel<<<blocks, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        const double fma = (double)blocks * 256 * kChains * iters;
        printf("{\"bench\": \"dfma_peak\", \"ms\": %.4f, \"dfma_per_s\": %.4e, \"fp64_tflops\": %.3f, "
               "\"dfma_per_clk_per_sm_at_max_clock\": %.2f}\n",
               ms, fma / (ms * 1e-3), 2.0 * fma / (ms * 1e-3) / 1e12,
               fma / (ms * 1e-3) / (clock_khz * 1e3) / sms);
    }
    {
        const int blocks = sms * 8, iters = 5000;
        float ms = time_ms([&] { rcp_kernel<<<blocks, 256>>>(out, iters); }, 5);
        const double ops = (double)blocks * 256 * kChains * iters;
        printf("{\"bench\": \"rcp64h\", \"ms\": %.4f, \"ops_per_s\": %.4e, "
               "\"ops_per_clk_per_sm_at_max_clock\": %.2f}\n",
               ms, ops / (ms * 1e-3), ops / (ms * 1e-3) / (clock_khz * 1e3) / sms);
    }
    {
        const int n_lines = 4096;
        std::vector<double2> ab(n_lines);
        std::vector<double> cc(n_lines);
        for (int j = 0; j < n_lines; ++j)
        {
            const double A = 1e-30 * (1 + j % 7), nu = 900.0 + 0.05 * j, g = 0.05;
            ab[j].x = 1.0 / sqrt(A);
            ab[j].y = -nu * ab[j].x;
            cc[j] = g * g / A;
        }
        double2* d_ab;
        double* d_cc;
        CHECK(cudaMalloc(&d_ab, sizeof(double2) * n_lines));
        CHECK(cudaMalloc(&d_cc, sizeof(double) * n_lines));
        CHECK(cudaMemcpy(d_ab, ab.data(), sizeof(double2) * n_lines, cudaMemcpyHostToDevice));
        CHECK(cudaMemcpy(d_cc, cc.data(), sizeof(double) * n_lines, cudaMemcpyHostToDevice));
        bench_far<1, 0>(d_ab, d_cc, n_lines, out, sms);
        bench_far<1, 1>(d_ab, d_cc, n_lines, out, sms);
        bench_far<1, 2>(d_ab, d_cc, n_lines, out, sms);
        bench_far<2, 0>(d_ab, d_cc, n_lines, out, sms);
        bench_far<2, 1>(d_ab, d_cc, n_lines, out, sms);
        bench_far<2, 2>(d_ab, d_cc, n_lines, out, sms);
        bench_far<4, 0>(d_ab, d_cc, n_lines, out, sms);
        bench_far<4, 1>(d_ab, d_cc, n_lines, out, sms);
        bench_far<4, 2>(d_ab, d_cc, n_lines, out, sms);
        bench_far<5, 0>(d_ab, d_cc, n_lines, out, sms);
        bench_far<5, 1>(d_ab, d_cc, n_lines, out, sms);
        bench_far<5, 2>(d_ab, d_cc, n_lines, out, sms);
        bench_far<8, 0>(d_ab, d_cc, n_lines, out, sms);
        bench_far<8, 1>(d_ab, d_cc, n_lines, out, sms);
        bench_far<8, 2>(d_ab, d_cc, n_lines, out, sms);
        bench_far<10, 0>(d_ab, d_cc, n_lines, out, sms);
        bench_far<10, 1>(d_ab, d_cc, n_lines, out, sms);
        bench_far<10, 2>(d_ab, d_cc, n_lines, out, sms);
    }
    {
        const int n = 1 << 20;
        std::vector<double> q(n);
        srand(1);
        for (int i = 0; i < n; ++i)
        {
            const double m = 1.0 + (double)rand() / RAND_MAX;
            const int e = (rand() % 400) - 200;
            q[i] = ldexp(m, e);
        }
        double *d_q, *d_e0, *d_e1;
        CHECK(cudaMalloc(&d_q, sizeof(double) * n));
        CHECK(cudaMalloc(&d_e0, sizeof(double) * n));
        CHECK(cudaMalloc(&d_e1, sizeof(double) * n));
        CHECK(cudaMemcpy(d_q, q.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
        rcp_accuracy_kernel<<<(n + 255) / 256, 256>>>(d_q, n, d_e0, d_e1);
        std::vector<double> e0(n), e1(n);
        CHECK(cudaMemcpy(e0.data(), d_e0, sizeof(double) * n, cudaMemcpyDeviceToHost));
        CHECK(cudaMemcpy(e1.data(), d_e1, sizeof(double) * n, cudaMemcpyDeviceToHost));
        double m0 = 0, m1 = 0;
        for (int i = 0; i < n; ++i)
        {
            if (e0[i] > m0) m0 = e0[i];
            if (e1[i] > m1) m1 = e1[i];
        }
        printf("{\"bench\": \"rcp64h_accuracy\", \"samples\": %d, \"max_rel_err_seed\": %.3e, "
               "\"max_rel_err_after_newton\": %.3e}\n", n, m0, m1);
    }
    return 0;
}
