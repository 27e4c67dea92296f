/* lbl_oracle.c -- CPU restatement of pyLBL's c_lib line-by-line path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under pylbl_b200/ may import, link or call this
 * file; it exists so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg can check the CUDA path against an independent statement of the reference
 * algorithm on machines where /root/reference is absent (the GPU box).
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here bit-for-bit
 * against the unmodified reference sources compiled from /root/reference into
 * oracle/_ref/libabsorption_ref.so (oracle/Makefile), and against the golden vectors
 * under tests/golden/ that were generated from that same library
 * (tests/golden/make_golden.py).
 *
 * Unlike the reference this file does not touch sqlite: the caller hands over the
 * line list as arrays in database row order, which is what the reference's
 * "select ... from transition where molecule_id == N" yields (absorption.c:67-73).
 *
 * Arithmetic is written so that, compiled without FMA contraction, every double
 * operation happens in the same order as in the reference; citations give the
 * reference file:line each block follows (paths relative to /root/reference/pyLBL/c_lib).
 */
#include <math.h>
#include <stddef.h>
#include <string.h>

/* ---- per-line Voigt state (voigt.c:7-15, 33-53, 87-144) ------------------------- */
typedef struct
{
    double centre;   /* shifted line centre nu' */
    double repwid;   /* sqrt(ln 2)/alpha                      voigt.c:13 */
    double y, yq;    /* repwid*gamma and its square           voigt.c:14-15 */
    double cof;      /* sw*rsqrpi*repwid is applied per point voigt.c:188 */
    double sw;
    double lim0, lim1, lim2, lim3, lim4; /* region limits     voigt.c:34-53 */
    double yrrtpi;
    double a0, d0, d2;                         /* region 1    voigt.c:91-93 */
    double h0, h2, h4, h6, e0, e2, e4;         /* region 2    voigt.c:105-111 */
    double z0, z2, z4, z6, z8, p0, p2, p4, p6, p8; /* region 3 voigt.c:123-143 */
} oracle_profile;

static const double k_c[6] = {1.0117281, -0.75197147, 0.012557727,
                              0.010022008, -0.00024206814, 0.00000050084806};
static const double k_s[6] = {1.393237, 0.23115241, -0.15535147,
                              0.0062183662, 0.000091908299, -0.00000062752596};
static const double k_t[6] = {0.31424038, 0.94778839, 1.5976826,
                              2.2795071, 3.0206370, 3.8897249};

static void profile_setup(oracle_profile *f, double nu, double alpha, double gamma, double sw)
{
    double const sqrln2 = sqrt(log(2.));
    double const rsqrpi = 1. / sqrt(M_PI);
    f->centre = nu;
    f->sw = sw;
    f->repwid = sqrln2 / alpha;
    f->y = f->repwid * gamma;
    f->yq = f->y * f->y;
    double const y = f->y, yq = f->yq;
    f->yrrtpi = y * rsqrpi;
    f->lim0 = sqrt(15100. + y * (40. - y * 3.6));
    f->lim1 = (y >= 8.425) ? 0. : sqrt(164. - y * (4.3 + y * 1.8));
    f->lim2 = 6.8 - y;
    f->lim3 = 2.4 * y;
    f->lim4 = 18.1 * y + 1.65;
    if (y <= 0.000001)
    {
        f->lim1 = f->lim0;
        f->lim2 = f->lim0;
    }
    /* The reference fills these lazily on the first point of each region; they are pure
     * functions of y so computing them up front gives the same bits. */
    f->a0 = yq + 0.5;
    f->d0 = f->a0 * f->a0;
    f->d2 = yq + yq - 1.;
    f->h0 = 0.5625 + yq * (4.5 + yq * (10.5 + yq * (6.0 + yq)));
    f->h2 = -4.5 + yq * (9.0 + yq * (6.0 + yq * 4.0));
    f->h4 = 10.5 - yq * (6.0 - yq * 6.0);
    f->h6 = -6.0 + yq * 4.0;
    f->e0 = 1.875 + yq * (8.25 + yq * (5.5 + yq));
    f->e2 = 5.25 + yq * (1.0 + yq * 3.0);
    f->e4 = 0.75 * f->h6;
    f->z0 = 272.1014 + y * (1280.829 + y * (2802.870 + y * (3764.966
            + y * (3447.629 + y * (2256.981 + y * (1074.409 + y * (369.1989
            + y * (88.26741 + y * (13.39880 + y)))))))));
    f->z2 = 211.678 + y * (902.3066 + y * (1758.336 + y * (2037.310
            + y * (1549.675 + y * (793.4273 + y * (266.2987
            + y * (53.59518 + y * 5.0)))))));
    f->z4 = 78.86585 + y * (308.1852 + y * (497.3014 + y * (479.2576
            + y * (269.2916 + y * (80.39278 + y * 10.0)))));
    f->z6 = 22.03523 + y * (55.02933 + y * (92.75679 + y * (53.59518
            + y * 10.0)));
    f->z8 = 1.496460 + y * (13.39880 + y * 5.0);
    f->p0 = 153.5168 + y * (549.3954 + y * (919.4955 + y * (946.8970
            + y * (662.8097 + y * (328.2151 + y * (115.3772 + y * (27.93941
            + y * (4.264678 + y * 0.3183291))))))));
    f->p2 = -34.16955 + y * (-1.322256 + y * (124.5975 + y * (189.7730
            + y * (139.4665 + y * (56.81652 + y * (12.79458
            + y * 1.2733163))))));
    f->p4 = 2.584042 + y * (10.46332 + y * (24.01655 + y * (29.81482
            + y * (12.79568 + y * 1.9099744))));
    f->p6 = -0.07272979 + y * (0.9377051 + y * (4.266322 + y * 1.273316));
    f->p8 = 0.0005480304 + y * 0.3183291;
}

/* Region code of one point: 0 = Lorentz (y>=70.55 branch), 1 = region 0 of W4,
 * 2..4 = W4 regions 1..3, 5 = CPF12 i, 6 = CPF12 ii.  Used for the region histogram. */
static int point_region(oracle_profile const *f, double abx)
{
    if (f->y >= 70.55) return 0;
    if (abx >= f->lim0) return 1;
    if (abx >= f->lim1) return 2;
    if (abx >= f->lim2) return 3;
    if (abx < f->lim3) return 4;
    return (abx <= f->lim4) ? 5 : 6;
}

/* Contribution of one line to one grid point (voigt.c:21-25 and voigt.c:74-188). */
static double profile_point(oracle_profile const *f, double v)
{
    double const rsqrpi = 1. / sqrt(M_PI);
    double const y = f->y, yq = f->yq, repwid = f->repwid;
    double xi = (v - f->centre) * repwid;
    if (y >= 70.55)
    {
        return f->sw * repwid * y / (M_PI * (xi * xi + yq));   /* voigt.c:24 */
    }
    double abx = fabs(xi);
    double xq = abx * abx;
    double buf;
    if (abx >= f->lim0)
    {
        buf = f->yrrtpi / (xq + yq);                             /* voigt.c:82 */
    }
    else if (abx >= f->lim1)
    {
        double d = rsqrpi / (f->d0 + xq * (f->d2 + xq));         /* voigt.c:95-96 */
        buf = d * y * (f->a0 + xq);
    }
    else if (abx >= f->lim2)
    {
        double d = rsqrpi / (f->h0 + xq * (f->h2 + xq * (f->h4 + xq * (f->h6 + xq))));
        buf = d * y * (f->e0 + xq * (f->e2 + xq * (f->e4 + xq))); /* voigt.c:113-114 */
    }
    else if (abx < f->lim3)
    {
        /* voigt.c:145-146; note the truncated literal for sqrt(pi). */
        double d = 1.7724538 / (f->z0 + xq * (f->z2 + xq * (f->z4 + xq * (f->z6 + xq * (f->z8 + xq)))));
        buf = d * (f->p0 + xq * (f->p2 + xq * (f->p4 + xq * (f->p6 + xq * f->p8))));
    }
    else
    {
        /* CPF12, voigt.c:151-186. */
        double const y0 = 1.5;
        double const y0py0 = y0 + y0;
        double const y0q = y0 * y0;
        double ypy0 = y + y0;
        double ypy0q = ypy0 * ypy0;
        double mq[6], pq[6], mf[6], pf[6], xm[6], xp[6], ym[6], yp[6];
        int j;
        for (j = 0; j < 6; ++j)
        {
            double d = xi - k_t[j];
            mq[j] = d * d;
            mf[j] = 1. / (mq[j] + ypy0q);
            xm[j] = mf[j] * d;
            ym[j] = mf[j] * ypy0;
            d = xi + k_t[j];
            pq[j] = d * d;
            pf[j] = 1. / (pq[j] + ypy0q);
            xp[j] = pf[j] * d;
            yp[j] = pf[j] * ypy0;
        }
        buf = 0.;
        if (abx <= f->lim4)
        {
            for (j = 0; j < 6; ++j)
            {
                buf += k_c[j] * (ym[j] + yp[j]) - k_s[j] * (xm[j] - xp[j]);
            }
        }
        else
        {
            double yf = y + y0py0;
            for (j = 0; j < 6; ++j)
            {
                buf += (k_c[j] * (mq[j] * mf[j] - y0 * ym[j]) + k_s[j] * yf * xm[j]) / (mq[j] + y0q)
                     + (k_c[j] * (pq[j] * pf[j] - y0 * yp[j]) - k_s[j] * yf * xp[j]) / (pq[j] + y0q);
            }
            buf = y * buf + exp(-xq);
        }
    }
    return f->sw * rsqrpi * repwid * buf;                        /* voigt.c:188 */
}

/* voigt(): accumulate one line into k[start..end] (voigt.c:4-191). */
void lbl_oracle_voigt(double const *v, int start, int end, double nu, double alpha,
                      double gamma, double sw, double *k)
{
    oracle_profile f;
    profile_setup(&f, nu, alpha, gamma, sw);
    int i;
    for (i = start; i <= end; ++i)
    {
        k[i] += profile_point(&f, v[i]);
    }
}

/* Region histogram of one line over v[start..end]; hist has 7 slots (point_region). */
void lbl_oracle_regions(double const *v, int start, int end, double nu, double alpha,
                        double gamma, long long *hist)
{
    oracle_profile f;
    profile_setup(&f, nu, alpha, gamma, 1.0);
    int i;
    for (i = start; i <= end; ++i)
    {
        double xi = (v[i] - f.centre) * f.repwid;
        hist[point_region(&f, fabs(xi))] += 1;
    }
}

/* total_partition_function(): spectral_database.c:97-104.  `t`/`q` are the flat
 * isotopologue-major tables, num_t entries per isotopologue. */
double lbl_oracle_tips(double const *t, double const *q, int num_t, double temperature, int iso)
{
    double const *tt = t + (ptrdiff_t)iso * num_t;
    double const *qq = q + (ptrdiff_t)iso * num_t;
    int i = (int)(floor(temperature)) - (int)(tt[0]);
    return qq[i] + (qq[i + 1] - qq[i]) * (temperature - tt[i]) / (tt[i + 1] - tt[i]);
}

/* The scaling part of spectra() (spectra.c:12-45): out = {nu', alpha, gamma, sw'}. */
void lbl_oracle_scale_line(double temperature, double pressure, double abundance,
                           double nu0, double sw0, double gamma_air, double gamma_self,
                           double n_air, double elower, double delta_air, double mass,
                           double q_ref, double q_t, double *out)
{
    double const vlight = 2.99792458e8;
    double const pa_to_atm = 9.86923e-6;
    double const r2 = 2 * log(2) * 8314.472;
    double const c2 = 1.4387752;

    double p = pressure * pa_to_atm;
    double partial_pressure = p * abundance;
    double tfact = 296. / temperature;
    double nu = nu0 + p * delta_air;                                       /* :22 */
    double gamma = (gamma_air * (p - partial_pressure) +
                    gamma_self * partial_pressure) * pow(tfact, n_air);    /* :25-26 */
    double alpha = (nu0 / vlight) * sqrt(r2 * temperature / mass);         /* :29 */
    double sb = exp(elower * c2 * (temperature - 296.) / (temperature * 296.)); /* :33 */
    double g = exp((-c2 * nu0) / temperature);                             /* :36 */
    double gref = exp((-c2 * nu0) / 296.);                                 /* :37 */
    double se = (1. - g) / (1. - gref);                                    /* :38 */
    double sq = q_ref / q_t;                                               /* :41-42 */
    double sw = sw0 * sb * se * sq * 0.01 * 0.01;                          /* :45 */
    out[0] = nu;
    out[1] = alpha;
    out[2] = gamma;
    out[3] = sw;
}

/* Window of one line on the grid (spectra.c:48-62).  Returns 0 if the line is skipped
 * (s >= n), else 1 with the clamped inclusive indices in *s and *e. */
int lbl_oracle_window(double nu_shifted, double v_first, int n, int n_per_v, int cut_off,
                      int *s_out, int *e_out)
{
    int s = (floor(nu_shifted) - cut_off - v_first) * n_per_v;
    if (s >= n)
    {
        return 0;
    }
    if (s < 0)
    {
        s = 0;
    }
    int e = (floor(nu_shifted) + cut_off + 1 - v_first) * n_per_v;
    if (e >= n)
    {
        e = n - 1;
    }
    *s_out = s;
    *e_out = e;
    return 1;
}

/* absorption(): absorption.c:19-99 with the sqlite reads replaced by arrays in database
 * row order.  `mass` has 32 slots indexed isoid-1 (spectral_database.c:108-133); the
 * local_iso_id==0 -> 10 mapping of spectral_database.c:173-177 is applied here.
 *
 * Extra outputs (may be NULL): n_evals = sum over processed lines of (e-s+1);
 * n_active = number of lines reached before the early break (absorption.c:80-83);
 * win = 2*n_lines ints (s,e per processed line, -1,-1 for skipped ones).
 * Returns 0; a line with e < s (possible only in the reference's UB corner, quirk Q9)
 * contributes nothing here.
 */
int lbl_oracle_absorption(double pressure, double temperature, double volume_mixing_ratio,
                          int v0, int vn, int n_per_v, double *k,
                          int n_lines, double const *nu, double const *sw,
                          double const *gamma_air, double const *gamma_self,
                          double const *n_air, double const *elower,
                          double const *delta_air, int const *local_iso_id,
                          double const *mass, int num_iso, int num_t,
                          double const *tips_t, double const *tips_q,
                          int cut_off, int remove_pedestal,
                          double *v_work, long long *n_evals, int *n_active, int *win)
{
    double dv = 1. / n_per_v;                                    /* absorption.c:33 */
    int n = (vn - v0) * n_per_v;
    int i;
    for (i = 0; i < n; ++i)
    {
        v_work[i] = v0 + i * dv;                                 /* absorption.c:39 */
    }
    memset(k, 0, sizeof(double) * n);
    long long evals = 0;
    int line;
    (void)num_iso;
    for (line = 0; line < n_lines; ++line)
    {
        if (nu[line] > vn + cut_off + 1 || nu[line] < v0 - (cut_off + 1))
        {
            break;                                               /* absorption.c:80-83 */
        }
        int iso = local_iso_id[line];
        if (iso == 0)
        {
            iso = 10;
        }
        double scaled[4];
        double q_ref = lbl_oracle_tips(tips_t, tips_q, num_t, 296., iso - 1);
        double q_t = lbl_oracle_tips(tips_t, tips_q, num_t, temperature, iso - 1);
        lbl_oracle_scale_line(temperature, pressure, volume_mixing_ratio, nu[line], sw[line],
                              gamma_air[line], gamma_self[line], n_air[line], elower[line],
                              delta_air[line], mass[iso - 1], q_ref, q_t, scaled);
        int s, e;
        if (win)
        {
            win[2 * line] = -1;
            win[2 * line + 1] = -1;
        }
        if (!lbl_oracle_window(scaled[0], v_work[0], n, n_per_v, cut_off, &s, &e))
        {
            continue;
        }
        if (win)
        {
            win[2 * line] = s;
            win[2 * line + 1] = e;
        }
        if (e < s)
        {
            continue;
        }
        evals += (long long)(e - s + 1);
        lbl_oracle_voigt(v_work, s, e, scaled[0], scaled[1], scaled[2], scaled[3], k);
        if (remove_pedestal != 0)
        {
            double pedestal = k[s];                              /* spectra.c:68-72 */
            if (k[e] < k[s])
            {
                pedestal = k[e];
            }
            for (i = s; i <= e; ++i)
            {
                k[i] -= pedestal;
            }
        }
    }
    if (n_evals)
    {
        *n_evals = evals;
    }
    if (n_active)
    {
        *n_active = line;
    }
    return 0;
}
