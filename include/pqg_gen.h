/*
 * pqg_gen.h -- synthetic workload / fixture generator (C-ABI).
 *
 * Turns columnar host arrays into the image of a Parquet file that is BYTE-IDENTICAL to
 * what the reference's ParquetWriter (src/writer/parquet_writer.cpp:376-581, with
 * include/writer/rle_bp_encoder.hpp and src/writer/thrift_writer.cpp) writes for the same
 * values: same per-row-group dictionary decision (distinct <= non_null / 5, first-seen
 * order, :255-283), same ~1 KB page splitting (:56-98), RLE-only definition levels
 * (:103-135), RLE/bit-packed index stream, same Thrift footer.  tests/test_gen_cpu.py
 * asserts the identity against the compiled reference (oracle/_ref).
 *
 * Why it exists: the reference writer runs at 0.2-4 Mvalues/s on one thread and keeps a
 * whole row group as 48-byte Values (SURVEY.md section 7 "Fixture cost"), so the
 * BASELINE.json configurations (1e8 .. 1e9 rows) cannot be produced with it inside a
 * benchmark run.  This generator is multi-threaded (one task per column chunk) and works
 * on flat arrays.  It is NOT on the decode path: nothing in pqg.h / pqg_reader.h calls it.
 *
 * Limits (checked, reported through pqgen_last_error): flat REQUIRED / OPTIONAL columns of
 * BOOLEAN, INT32, INT64, FLOAT, DOUBLE, BYTE_ARRAY; FLOAT/DOUBLE columns must not contain
 * NaN (the reference's std::map<variant> ordering is undefined for NaN).
 */
#ifndef PQG_GEN_H
#define PQG_GEN_H
#include <stddef.h>
#include <stdint.h>

#include "pqg.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pqgen_col {
    const char* name;
    int32_t type;       /* PQG_* physical type */
    int32_t repetition; /* 0 REQUIRED, 1 OPTIONAL */
    int32_t converted;  /* ConvertedType or -1 */
    /* values of ALL rows of the file, row groups concatenated; null rows are ignored:
     *   BOOLEAN: 1 byte per row; INT32/FLOAT: 4 bytes; INT64/DOUBLE: 8 bytes */
    const void* fixed;
    /* BYTE_ARRAY: rows + 1 offsets into chars */
    const uint64_t* str_off;
    const uint8_t* chars;
    const uint8_t* is_null; /* 1 byte per row, or NULL = no nulls */
} pqgen_col;

typedef struct pqgen_job pqgen_job;

PQG_API const char* pqgen_last_error(void);
/* Encode every column chunk (threads <= 0: one per hardware thread).  rg_rows[r] = rows of
 * row group r.  Returns NULL on error. */
PQG_API pqgen_job* pqgen_encode(const pqgen_col* cols, int32_t n_cols, const int64_t* rg_rows,
                                int32_t n_rgs, int32_t threads);
PQG_API uint64_t pqgen_size(const pqgen_job* job);       /* bytes of the file image */
PQG_API int pqgen_emit(const pqgen_job* job, uint8_t* dst, uint64_t cap); /* 0 = ok */
PQG_API int pqgen_write_file(const pqgen_job* job, const char* path);
PQG_API void pqgen_free(pqgen_job* job);

/* Synthetic BYTE_ARRAY columns of the BASELINE.json configurations, filled natively (multi-threaded; row i of the
 * file depends on (kind, seed, first_row + i) only, so shards generated apart agree with the whole):
 *   PQGEN_EMAILS   configs[3]: "user<9 digits>@mail<3 digits>.example.com" (33 bytes); rows in blocks of 2000, every
 *                  fifth block is noise -- alternately without '@' ('.' instead) and ending in ".co!" -- so that whole
 *                  pages fail the anchored e-mail pattern and both polarities of the predicate prune pages
 *   PQGEN_CITY64K  configs[2]: "city_<6 digits>_x" (13 bytes), 65 536 distinct values, uniform
 * All strings of a kind have one length (pqgen_string_len): chars receives rows * len bytes, str_off (may be NULL)
 * rows + 1 offsets starting at off_base; is_null (may be NULL) gets 1 with probability null_permille / 1000. */
enum { PQGEN_EMAILS = 0, PQGEN_CITY64K = 1 };
PQG_API uint32_t pqgen_string_len(int32_t kind);
PQG_API int pqgen_fill_strings(int32_t kind, uint64_t first_row, uint64_t rows, uint64_t seed, uint8_t* chars, uint64_t* str_off,
                               uint64_t off_base, uint8_t* is_null, uint32_t null_permille, int32_t threads);

#ifdef __cplusplus
}
#endif
#endif /* PQG_GEN_H */
