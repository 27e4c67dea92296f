// Host-only: cutting a grid into contiguous spectral bands of about equal cost (the arithmetic
// behind lbl_gas_band_edges; also compiled into the CPU test library, tests/emu/emu.cpp).
#pragma once

#include <algorithm>
#include <vector>

namespace lbl
{

// cum[c] = cost of cells [0, c) (ncell + 1 entries, non-decreasing); prefix[c] = what a band that
// ENDS at cell c pays on top of its own cells (the pedestal recurrence over the rows before
// it; non-decreasing).  Writes n_bands + 1 edges, edges[0] = 0 and edges[n_bands] = ncell:
// the partition with the smallest largest band cost, found by bisection on that cost -- for a
// given bound the bands are taken greedily as wide as it allows, which is optimal because a
// band's cost grows with its end.  Bands are non-empty while there are cells to give them.
inline void partition_bands(const std::vector<double>& cum, const std::vector<double>& prefix, int n_bands,
                            int* edges)
{
    const int ncell = (int)cum.size() - 1;
    auto bands_for = [&](double T, int* out) {
        int lo = 0, used = 0;
        while (lo < ncell && used < n_bands)
        {
            int l = lo, h = ncell;
            while (l < h)
            {
                const int m = (l + h + 1) / 2;
                if (cum[m] - cum[lo] + prefix[m] <= T) l = m;
                else h = m - 1;
            }
            if (l == lo) return false;   // not even one cell fits
            lo = l;
            if (out) out[++used] = lo;
            else ++used;
        }
        return lo == ncell;
    };
    double t_lo = 0., t_hi = cum[ncell] + prefix[ncell];
    for (int it = 0; it < 60; ++it)
    {
        const double mid = 0.5 * (t_lo + t_hi);
        if (bands_for(mid, nullptr)) t_hi = mid;
        else t_lo = mid;
    }
    edges[0] = 0;
    for (int b = 1; b <= n_bands; ++b) edges[b] = ncell;
    bands_for(t_hi, edges);
    // fewer bands than asked for (a grid of few cells, or one cell that outweighs the rest):
    // split the widest until every band that can be non-empty is
    std::vector<int> e(edges, edges + n_bands + 1);
    e.erase(std::unique(e.begin(), e.end()), e.end());
    while ((int)e.size() < n_bands + 1 && (int)e.size() - 1 < ncell)
    {
        size_t widest = 0;
        for (size_t k = 0; k + 1 < e.size(); ++k)
        {
            if (e[k + 1] - e[k] > e[widest + 1] - e[widest]) widest = k;
        }
        if (e[widest + 1] - e[widest] < 2) break;
        e.insert(e.begin() + widest + 1, (e[widest] + e[widest + 1]) / 2);
    }
    for (int b = 0; b <= n_bands; ++b) edges[b] = b < (int)e.size() ? e[b] : ncell;
    edges[n_bands] = ncell;
}

}  // namespace lbl
