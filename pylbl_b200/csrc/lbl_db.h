// lbl_db.h -- reads one molecule of a pyLBL spectral database (sqlite) into host arrays.
//
// Replaces the per-call sqlite traffic of the reference (pyLBL/c_lib/absorption.c:45-79,
// pyLBL/c_lib/spectral_database.c:19-180): the same four queries are run ONCE per
// (database, formula) and the rows are kept in database row order.
#pragma once

#include <string>
#include <vector>

namespace lbl
{

struct MoleculeData
{
    int molecule_id = -1;
    bool has_tips = false;          // false -> the reference returns an all-zero spectrum
    int num_iso = 0, num_t = 0;     // TIPS table shape (spectral_database.c:59-90)
    std::vector<double> tips_t, tips_q;
    double iso_mass[32];            // indexed isoid-1 (spectral_database.c:108-133)
    // transition rows, database row order (absorption.c:67-73)
    std::vector<double> nu, sw, gamma_air, gamma_self, n_air, elower, delta_air, mass;
    std::vector<int> iso;           // local_iso_id with the 0 -> 10 rule applied (:173-177)
    bool sorted = true;             // nu non-decreasing in row order
    double max_abs_delta = 0.;      // max |delta_air|
    double min_mass = 0.;           // smallest positive mass among the lines
};

// Returns 0 on success, 1 on error (message in err), like the reference's helpers.
int read_molecule(const char* path, const char* formula, MoleculeData& out, std::string& err);

// ---- packed line-list cache (lbl_pack.cpp) -------------------------------------------------
struct PackInfo
{
    std::string formula;
    long long n_lines = 0;
    int num_iso = 0, num_t = 0;
    bool sorted = true, has_tips = false;
    long long source_size = 0, source_mtime = 0;
};

// Writes everything read_molecule() produced to one flat binary file (atomically).
int write_pack(const char* path, const char* formula, const MoleculeData& m, long long source_size,
               long long source_mtime, std::string& err);
// Reads it back (header_only: just `info`); verifies magic, version and checksum.
int read_pack(const char* path, MoleculeData& out, PackInfo& info, bool header_only, std::string& err);

}  // namespace lbl
