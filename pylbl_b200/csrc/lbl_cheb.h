// lbl_cheb.h -- interpolation tables of the cell-tiled summation kernel (K2c), host side.
//
// Chebyshev nodes of the first kind on the cell interval [0, (n_per_v-1)/n_per_v] (offsets
// from the cell's first grid point), and the transform from the values at those nodes to the
// Chebyshev coefficients of the interpolant; long double arithmetic.
#pragma once

#include <cmath>
#include <vector>

namespace lbl
{

// Node k sits at half + half*cos(theta_k), theta_k = (2k+1) pi / (2n), half = (n_per_v-1)/(2 n_per_v):
// in the interval's own coordinate s in [-1, 1] that is s_k = cos(theta_k).
inline void build_cheb_nodes(int n_nodes, int n_per_v, std::vector<double>& nodes)
{
    const long double pi = 3.14159265358979323846264338327950288L;
    const long double half = (long double)(n_per_v - 1) / (2.0L * n_per_v);
    nodes.assign(n_nodes, 0.);
    for (int k = 0; k < n_nodes; ++k)
    {
        const long double theta = (2 * k + 1) * pi / (2.0L * n_nodes);
        nodes[k] = (double)(half + half * cosl(theta));
    }
}

// Values at the n nodes -> Chebyshev coefficients of the interpolant:
//   p(s) = sum_j c_j T_j(s),  c_j = sum_k M[k][j] F_k,  M[k][j] = (2 - [j == 0])/n * cos(j theta_k).
// Stored k-major so that lane j reads consecutive addresses.
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
