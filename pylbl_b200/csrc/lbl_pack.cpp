// lbl_pack.cpp -- the packed line-list cache.
//
// One file per (database, formula): everything lbl_gas_open() needs, as flat little-endian
// arrays, so that later runs skip sqlite altogether.  The reference pays the sqlite read on
// EVERY absorption() call (pyLBL/c_lib/absorption.c:45-79: open, prepare, step over all rows;
// 0.84 us per line); lbl_gas_open pays it once per handle, a pack once per database.
//
// Layout (all 8-byte aligned):
//   PackHeader
//   f64 tips_t[num_iso*num_t], tips_q[num_iso*num_t]
//   f64 nu[n], sw[n], gamma_air[n], gamma_self[n], n_air[n], elower[n], delta_air[n], mass[n]
//   i32 iso[n]  (+ 4 bytes of padding when n is odd)
//   u64 FNV-1a of everything before it, header included
// Rows keep DATABASE ROW ORDER: the reference's early break (absorption.c:80-83) and its
// accumulated pedestal (spectra.c:66-78) depend on it.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <unistd.h>
#include <vector>

#include "lbl_db.h"

namespace lbl
{
namespace
{

constexpr char kMagic[8] = {'L', 'B', 'L', 'P', 'A', 'C', 'K', '1'};
constexpr uint32_t kPackVersion = 2;   // 2: the checksum covers the header too

struct PackHeader
{
    char magic[8];
    uint32_t version;
    uint32_t flags;        // bit 0: has_tips, bit 1: rows nu-sorted
    int64_t n_lines;
    int32_t molecule_id, num_iso, num_t, reserved;
    double max_abs_delta, min_mass;
    double iso_mass[32];
    char formula[32];
    int64_t source_size, source_mtime;   // of the sqlite file the pack was made from (0 = unknown)
};
static_assert(sizeof(PackHeader) % 8 == 0, "payload must start 8-byte aligned");

uint64_t fnv1a(const void* data, size_t n, uint64_t h)
{
    const unsigned char* p = static_cast<const unsigned char*>(data);
    for (size_t i = 0; i < n; ++i)
    {
        h ^= p[i];
        h *= 1099511628211ull;
    }
    return h;
}

struct Writer
{
    FILE* f;
    uint64_t hash = 14695981039346656037ull;
    bool ok = true;
    void put(const void* p, size_t n)
    {
        if (n == 0) return;
        hash = fnv1a(p, n, hash);
        ok = ok && fwrite(p, 1, n, f) == n;
    }
};

struct Reader
{
    FILE* f;
    uint64_t hash = 14695981039346656037ull;
    bool ok = true;
    void get(void* p, size_t n)
    {
        if (n == 0) return;
        ok = ok && fread(p, 1, n, f) == n;
        if (ok) hash = fnv1a(p, n, hash);
    }
};

}  // namespace

int write_pack(const char* path, const char* formula, const MoleculeData& m, long long source_size,
               long long source_mtime, std::string& err)
{
    // A name of its own per writer: several ranks may build the same pack at the same time.
    const std::string tmp = std::string(path) + ".tmp." + std::to_string((long long)getpid()) + "." +
                            std::to_string((unsigned long long)(uintptr_t)&err);
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f)
    {
        err = std::string("Error: cannot write ") + tmp;
        return 1;
    }
    PackHeader h;
    memset(&h, 0, sizeof(h));
    memcpy(h.magic, kMagic, 8);
    h.version = kPackVersion;
    h.flags = (m.has_tips ? 1u : 0u) | (m.sorted ? 2u : 0u);
    h.n_lines = (int64_t)m.nu.size();
    h.molecule_id = m.molecule_id;
    h.num_iso = m.num_iso;
    h.num_t = m.num_t;
    h.max_abs_delta = m.max_abs_delta;
    h.min_mass = m.min_mass;
    memcpy(h.iso_mass, m.iso_mass, sizeof(h.iso_mass));
    strncpy(h.formula, formula, sizeof(h.formula) - 1);
    h.source_size = source_size;
    h.source_mtime = source_mtime;
    bool ok = true;
    Writer w{f};
    w.put(&h, sizeof(h));
    w.put(m.tips_t.data(), sizeof(double) * m.tips_t.size());
    w.put(m.tips_q.data(), sizeof(double) * m.tips_q.size());
    const std::vector<double>* cols[8] = {&m.nu, &m.sw, &m.gamma_air, &m.gamma_self,
                                          &m.n_air, &m.elower, &m.delta_air, &m.mass};
    for (const std::vector<double>* c : cols) w.put(c->data(), sizeof(double) * c->size());
    w.put(m.iso.data(), sizeof(int) * m.iso.size());
    if (m.iso.size() & 1)
    {
        const int pad = 0;
        w.put(&pad, sizeof(pad));
    }
    ok = ok && w.ok && fwrite(&w.hash, 1, sizeof(w.hash), f) == sizeof(w.hash);
    ok = (fclose(f) == 0) && ok;
    if (!ok)
    {
        remove(tmp.c_str());
        err = std::string("Error: writing ") + path + " failed.";
        return 1;
    }
    if (rename(tmp.c_str(), path) != 0)   // atomic: readers never see a partial pack
    {
        remove(tmp.c_str());
        // another writer's complete pack already being there is as good as ours
        FILE* there = fopen(path, "rb");
        if (!there)
        {
            err = std::string("Error: writing ") + path + " failed.";
            return 1;
        }
        fclose(there);
    }
    return 0;
}

int read_pack(const char* path, MoleculeData& out, PackInfo& info, bool header_only, std::string& err)
{
    FILE* f = fopen(path, "rb");
    if (!f)
    {
        err = std::string("Error: cannot open line-list pack ") + path;
        return 1;
    }
    PackHeader h;
    if (fread(&h, 1, sizeof(h), f) != sizeof(h) || memcmp(h.magic, kMagic, 8) != 0)
    {
        fclose(f);
        err = std::string("Error: ") + path + " is not a line-list pack.";
        return 1;
    }
    const bool tips_flag = (h.flags & 1u) != 0;
    if (h.version != kPackVersion || h.n_lines < 0 || h.n_lines > 0x7fffffff || h.num_iso < 0 ||
        h.num_iso > 32 || h.num_t < 0 || (int64_t)h.num_iso * h.num_t > (1 << 28) ||
        (tips_flag && (h.num_t < 2 || h.num_iso < 1)) || (!tips_flag && (h.num_t != 0 || h.num_iso != 0)))
    {
        fclose(f);
        err = std::string("Error: ") + path + ": unsupported pack version or corrupt header.";
        return 1;
    }
    h.formula[sizeof(h.formula) - 1] = 0;
    info.formula = h.formula;
    info.n_lines = h.n_lines;
    info.num_iso = h.num_iso;
    info.num_t = h.num_t;
    info.sorted = (h.flags & 2u) != 0;
    info.has_tips = (h.flags & 1u) != 0;
    info.source_size = h.source_size;
    info.source_mtime = h.source_mtime;
    const size_t n = (size_t)h.n_lines;
    const size_t nt = (size_t)h.num_iso * (size_t)h.num_t;
    out = MoleculeData();
    out.molecule_id = h.molecule_id;
    out.has_tips = info.has_tips;
    out.num_iso = h.num_iso;
    out.num_t = h.num_t;
    out.sorted = info.sorted;
    out.max_abs_delta = h.max_abs_delta;
    out.min_mass = h.min_mass;
    memcpy(out.iso_mass, h.iso_mass, sizeof(out.iso_mass));
    Reader r{f};
    r.hash = fnv1a(&h, sizeof(h), r.hash);
    out.tips_t.resize(nt);
    out.tips_q.resize(nt);
    r.get(out.tips_t.data(), sizeof(double) * nt);
    r.get(out.tips_q.data(), sizeof(double) * nt);
    std::vector<double>* cols[8] = {&out.nu, &out.sw, &out.gamma_air, &out.gamma_self,
                                    &out.n_air, &out.elower, &out.delta_air, &out.mass};
    for (std::vector<double>* c : cols)
    {
        c->resize(n);
        r.get(c->data(), sizeof(double) * n);
    }
    out.iso.resize(n);
    r.get(out.iso.data(), sizeof(int) * n);
    if (n & 1)
    {
        int pad = 0;
        r.get(&pad, sizeof(pad));
    }
    uint64_t stored = 0;
    const bool tail = fread(&stored, 1, sizeof(stored), f) == sizeof(stored);
    fclose(f);
    if (!r.ok || !tail || stored != r.hash)
    {
        err = std::string("Error: ") + path + ": truncated pack or checksum mismatch.";
        return 1;
    }
    for (size_t i = 0; i < n; ++i)
    {
        if (out.iso[i] < 1 || out.iso[i] > 32)
        {
            err = std::string("Error: ") + path + ": local_iso_id outside 1..32.";
            return 1;
        }
    }
    (void)header_only;   // the whole file is always read: the checksum is what vouches for it
    return 0;
}

}  // namespace lbl
