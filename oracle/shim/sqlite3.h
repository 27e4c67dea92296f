/* Minimal declaration shim for the system libsqlite3.so.0 (test infrastructure).
 *
 * The image ships the sqlite runtime library but not its development header.
 * The reference C sources include "sqlite3.h" and use only the handful of
 * entry points declared below (open/close/errmsg/prepare_v2/step/finalize/
 * column_int/column_double), so these declarations are enough to compile them
 * unmodified from /root/reference and link against the system library.
 * Values of SQLITE_OK / SQLITE_DONE are part of sqlite's stable public ABI.
 */
#ifndef PYLBL_B200_ORACLE_SQLITE3_SHIM_H
#define PYLBL_B200_ORACLE_SQLITE3_SHIM_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sqlite3 sqlite3;
typedef struct sqlite3_stmt sqlite3_stmt;

#define SQLITE_OK 0
#define SQLITE_ROW 100
#define SQLITE_DONE 101

int sqlite3_open(const char *filename, sqlite3 **db);
int sqlite3_close(sqlite3 *db);
const char *sqlite3_errmsg(sqlite3 *db);
int sqlite3_prepare_v2(sqlite3 *db, const char *sql, int nbyte,
                       sqlite3_stmt **stmt, const char **tail);
int sqlite3_step(sqlite3_stmt *stmt);
int sqlite3_finalize(sqlite3_stmt *stmt);
int sqlite3_column_int(sqlite3_stmt *stmt, int col);
double sqlite3_column_double(sqlite3_stmt *stmt, int col);

#ifdef __cplusplus
}
#endif

#endif
