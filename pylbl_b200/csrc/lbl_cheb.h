// lbl_cheb.h -- interpolation tables of the cell-tiled summation kernel (K2c), host side.
//
// Chebyshev nodes of the first kind on the cell interval [0, (n_per_v-1)/n_per_v] (offsets
// from the cell's first grid point) and the Lagrange basis of those nodes evaluated at the
// grid offsets r/n_per_v, in barycentric form and long double arithmetic.  weights[k][r].
#pragma once

#include <cmath>
#include <vector>

namespace lbl
{

inline void build_cheb_tables(int n_nodes, int n_per_v, std::vector<double>& nodes,
                              std::vector<double>& weights)
{
    const long double pi = 3.14159265358979323846264338327950288L;
    const long double half = (long double)(n_per_v - 1) / (2.0L * n_per_v);
    std::vector<long double> x(n_nodes), w(n_nodes);
    nodes.assign(n_nodes, 0.);
    weights.assign((size_t)n_per_v * n_nodes, 0.);
    for (int k = 0; k < n_nodes; ++k)
    {
        const long double theta = (2 * k + 1) * pi / (2.0L * n_nodes);
        x[k] = half + half * cosl(theta);
        w[k] = ((k & 1) ? -1.0L : 1.0L) * sinl(theta);
        nodes[k] = (double)x[k];
    }
    for (int r = 0; r < n_per_v; ++r)
    {
        const long double t = (long double)r / n_per_v;
        int hit = -1;
        long double denom = 0.0L;
        for (int k = 0; k < n_nodes; ++k)
        {
            if (fabsl(t - x[k]) < 1e-15L)
            {
                hit = k;
            }
            else
            {
                denom += w[k] / (t - x[k]);
            }
        }
        for (int k = 0; k < n_nodes; ++k)
        {
            double val;
            if (hit >= 0)
            {
                val = (k == hit) ? 1.0 : 0.0;
            }
            else
            {
                val = (double)((w[k] / (t - x[k])) / denom);
            }
            weights[(size_t)k * n_per_v + r] = val;   // node-major: a warp reads consecutive r
        }
    }
}

// Values at the n first-kind Chebyshev nodes -> Chebyshev coefficients of the interpolant:
//   p(s) = sum_j c_j T_j(s),  c_j = sum_k M[k][j] F_k,  M[k][j] = (2 - [j == 0])/n * cos(j theta_k),
// theta_k = (2k+1) pi / (2n), node k at s_k = cos(theta_k) (the same nodes build_cheb_tables
// places on the cell interval).  Stored k-major so that lane j reads consecutive addresses.
inline void build_cheb_transform(int n_nodes, std::vector<double>& m)
{
    const long double pi = 3.14159265358979323846264338327950288L;
    m.assign((size_t)n_nodes * n_nodes, 0.);
    for (int k = 0; k < n_nodes; ++k)
    {
        const long double theta = (2 * k + 1) * pi / (2.0L * n_nodes);
        for (int j = 0; j < n_nodes; ++j)
        {
            const long double scale = (j == 0 ? 1.0L : 2.0L) / n_nodes;
            m[(size_t)k * n_nodes + j] = (double)(scale * cosl(j * theta));
        }
    }
}

}  // namespace lbl
