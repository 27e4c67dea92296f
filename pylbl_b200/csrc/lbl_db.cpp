// lbl_db.cpp -- sqlite reader for the packer (see lbl_db.h).
//
// The image carries libsqlite3.so.0 but no sqlite3.h, so the handful of entry points used
// are resolved with dlopen/dlsym against their stable public C signatures.
#include "lbl_db.h"

#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>

namespace lbl
{
namespace
{

struct sqlite3;
struct sqlite3_stmt;
constexpr int kSqliteOk = 0;
constexpr int kSqliteRow = 100;
constexpr int kSqliteDone = 101;
constexpr int kOpenReadOnly = 0x00000001;

struct SqliteApi
{
    void* lib = nullptr;
    int (*open_v2)(const char*, sqlite3**, int, const char*) = nullptr;
    int (*close)(sqlite3*) = nullptr;
    const char* (*errmsg)(sqlite3*) = nullptr;
    int (*prepare_v2)(sqlite3*, const char*, int, sqlite3_stmt**, const char**) = nullptr;
    int (*step)(sqlite3_stmt*) = nullptr;
    int (*finalize)(sqlite3_stmt*) = nullptr;
    int (*column_int)(sqlite3_stmt*, int) = nullptr;
    double (*column_double)(sqlite3_stmt*, int) = nullptr;
    int (*bind_text)(sqlite3_stmt*, int, const char*, int, void (*)(void*)) = nullptr;
    std::string error;
};

SqliteApi& api()
{
    static SqliteApi a;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libsqlite3.so.0", "libsqlite3.so",
                               "/usr/lib/x86_64-linux-gnu/libsqlite3.so.0"};
        for (const char* n : names)
        {
            a.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
            if (a.lib) break;
        }
        if (!a.lib)
        {
            a.error = "cannot load libsqlite3";
            return;
        }
#define LBL_SYM(field, name) \
        a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.lib, name)); \
        if (!a.field) a.error = std::string("libsqlite3 lacks ") + name;
        LBL_SYM(open_v2, "sqlite3_open_v2")
        LBL_SYM(close, "sqlite3_close")
        LBL_SYM(errmsg, "sqlite3_errmsg")
        LBL_SYM(prepare_v2, "sqlite3_prepare_v2")
        LBL_SYM(step, "sqlite3_step")
        LBL_SYM(finalize, "sqlite3_finalize")
        LBL_SYM(column_int, "sqlite3_column_int")
        LBL_SYM(column_double, "sqlite3_column_double")
        LBL_SYM(bind_text, "sqlite3_bind_text")
#undef LBL_SYM
    });
    return a;
}

struct Connection
{
    SqliteApi& s;
    sqlite3* db = nullptr;
    explicit Connection(SqliteApi& a) : s(a) {}
    ~Connection() { if (db) s.close(db); }
};

struct Statement
{
    SqliteApi& s;
    sqlite3_stmt* st = nullptr;
    explicit Statement(SqliteApi& a) : s(a) {}
    ~Statement() { if (st) s.finalize(st); }
    // true while a row is available
    bool next(bool& failed)
    {
        const int rc = s.step(st);
        if (rc == kSqliteRow) return true;
        if (rc != kSqliteDone) failed = true;
        return false;
    }
};

bool prepare(Connection& c, Statement& st, const char* sql, std::string& err)
{
    if (c.s.prepare_v2(c.db, sql, -1, &st.st, nullptr) != kSqliteOk)
    {
        err = std::string("Error: ") + c.s.errmsg(c.db);
        return false;
    }
    return true;
}

}  // namespace

int read_molecule(const char* path, const char* formula, MoleculeData& out, std::string& err)
{
    SqliteApi& s = api();
    if (!s.error.empty())
    {
        err = s.error;
        return 1;
    }
    Connection con(s);
    // Read-only: unlike sqlite3_open (spectral_database.c:22) this never creates a file.
    if (s.open_v2(path, &con.db, kOpenReadOnly, nullptr) != kSqliteOk || con.db == nullptr)
    {
        err = std::string("Error: failed to open ") + path +
              (con.db ? std::string(": ") + s.errmsg(con.db) : std::string("."));
        return 1;
    }
    char query[256];
    bool failed = false;

    // molecule_id(), spectral_database.c:137-159.
    {
        Statement st(s);
        if (!prepare(con, st, "select molecule from molecule_alias where alias == ?1", err)) return 1;
        s.bind_text(st.st, 1, formula, -1, nullptr);
        out.molecule_id = -1;
        if (st.next(failed)) out.molecule_id = s.column_int(st.st, 0);
        if (failed || out.molecule_id == -1)
        {
            err = std::string("Error: molecule ") + formula + " not found in database.";
            return 1;
        }
    }

    // tips_data(), spectral_database.c:49-93.
    {
        Statement st(s);
        snprintf(query, sizeof(query),
                 "select isotopologue_id, temperature, data from tips where molecule_id == %d",
                 out.molecule_id);
        if (!prepare(con, st, query, err)) return 1;
        int current = -1;
        out.num_iso = 0;
        out.tips_t.clear();
        out.tips_q.clear();
        while (st.next(failed))
        {
            const int iso = s.column_int(st.st, 0);
            if (iso != current)
            {
                out.num_iso++;
                current = iso;
            }
            out.tips_t.push_back(s.column_double(st.st, 1));
            out.tips_q.push_back(s.column_double(st.st, 2));
        }
        if (failed)
        {
            err = std::string("Error: ") + s.errmsg(con.db);
            return 1;
        }
        out.has_tips = !out.tips_t.empty();
        if (out.has_tips)
        {
            out.num_t = (int)(out.tips_t.size() / out.num_iso);
            if ((size_t)out.num_t * out.num_iso != out.tips_t.size())
            {
                err = "Error: tips data is not rectangular.";
                return 1;
            }
            if (out.num_t < 2)
            {
                err = "Error: tips table needs at least two temperatures.";
                return 1;
            }
        }
    }

    // mass_data(), spectral_database.c:108-133.
    {
        Statement st(s);
        std::memset(out.iso_mass, 0, sizeof(out.iso_mass));
        snprintf(query, sizeof(query),
                 "select isoid, mass from isotopologue where molecule_id == %d", out.molecule_id);
        if (!prepare(con, st, query, err)) return 1;
        while (st.next(failed))
        {
            int i = s.column_int(st.st, 0);
            if (i == 0) i = 10;  // "Weird HITRAN counting."
            if (i >= 32 || i < 1)
            {
                err = "Error: buffer is too small, increase num_mass.";
                return 1;
            }
            out.iso_mass[i - 1] = s.column_double(st.st, 1);
        }
        if (failed)
        {
            err = std::string("Error: ") + s.errmsg(con.db);
            return 1;
        }
    }

    // Line parameters, absorption.c:67-79 + spectral_database.c:163-180.
    {
        Statement st(s);
        snprintf(query, sizeof(query),
                 "select nu, sw, gamma_air, gamma_self, n_air, elower, delta_air, "
                 "local_iso_id from transition where molecule_id == %d", out.molecule_id);
        if (!prepare(con, st, query, err)) return 1;
        double prev = -INFINITY;
        out.sorted = true;
        out.max_abs_delta = 0.;
        out.min_mass = 0.;
        while (st.next(failed))
        {
            const double nu = s.column_double(st.st, 0);
            int iso = s.column_int(st.st, 7);
            if (iso == 0) iso = 10;
            if (iso < 1 || iso > 32)
            {
                err = "Error: local_iso_id outside 1..32.";
                return 1;
            }
            // (iso <= num_iso is checked per grid, for the rows the reference would touch:
            // make_plan in lbl_api.cu)
            const double delta = s.column_double(st.st, 6);
            const double m = out.iso_mass[iso - 1];
            out.nu.push_back(nu);
            out.sw.push_back(s.column_double(st.st, 1));
            out.gamma_air.push_back(s.column_double(st.st, 2));
            out.gamma_self.push_back(s.column_double(st.st, 3));
            out.n_air.push_back(s.column_double(st.st, 4));
            out.elower.push_back(s.column_double(st.st, 5));
            out.delta_air.push_back(delta);
            out.mass.push_back(m);
            out.iso.push_back(iso);
            if (!(nu >= prev)) out.sorted = false;
            prev = nu;
            if (std::fabs(delta) > out.max_abs_delta) out.max_abs_delta = std::fabs(delta);
            if (m > 0. && (out.min_mass == 0. || m < out.min_mass)) out.min_mass = m;
        }
        if (failed)
        {
            err = std::string("Error: ") + s.errmsg(con.db);
            return 1;
        }
    }
    return 0;
}

}  // namespace lbl
