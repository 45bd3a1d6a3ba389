/*
 * pqg_reader.h -- C-ABI over the host-side reader (namespace pqg::ParquetReader /
 * ColumnReader / StringColumnIterator in duckdb-parquet-parser_b200/host/pq_reader.hpp),
 * i.e. the reference-facing API of the hot path, for callers that cannot link C++
 * (ctypes, cgo, JNI, N-API).  Every decode below runs on the GPU through pqg.h; the host
 * only parses Thrift metadata and page headers.
 *
 * Reference interface each entry point stands for (paths relative to the reference repo):
 *   pqr_open                      ParquetReader::open                src/reader/parquet_reader.cpp:14-61
 *   pqr_num_* / pqr_column_info   schema inspection                  src/reader/parquet_reader.cpp:65-121
 *   pqr_read_column*              ParquetReader::read_column*        src/reader/parquet_reader.cpp:125-165
 *   pqr_read_pages                ColumnReader::read_pages           src/reader/column_reader.cpp:73-126
 *   pqr_page_index / pqr_read_page_data / pqr_read_pages_chunk
 *                                 raw page API                       src/reader/parquet_reader.cpp:182-238
 *   pqr_string_iterator_dump      StringColumnIterator               src/reader/parquet_reader.cpp:282-465
 *   pqr_chunk_index               chunk-index prototype              src/main.cpp:21-32
 *   pqr_regex_prune               parser --regex-column mode         README.md:54-64
 *   pqr_page_chunk_index          index_test                         README.md:66-72
 * Errors: functions return 0 / a count on success and a negative value on failure; the
 * message (same text as the reference's std::runtime_error where one exists) is returned
 * by pqr_last_error() for the calling thread.
 */
#ifndef PQG_READER_H
#define PQG_READER_H
#include <stddef.h>
#include <stdint.h>

#include "pqg.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pqr_reader pqr_reader;

/* std::vector<Value> in flat form: is_null, variant alternative (0 bool, 1 int32, 2 int64,
 * 3 float, 4 double, 5 string), payload bits, string bytes.  Free with pqr_valdump_free. */
typedef struct pqr_valdump {
    int64_t n;
    uint8_t* is_null;
    uint8_t* vidx;
    uint64_t* fixed;
    uint64_t* str_off; /* n + 1 */
    uint8_t* chars;
    int64_t chars_len;
} pqr_valdump;

typedef struct pqr_pagedump {
    int64_t n_pages;
    int32_t* page_num;
    int32_t* page_type;
    int32_t* num_values;
    int64_t* first_value; /* n_pages + 1 */
    pqr_valdump values;
} pqr_pagedump;

typedef struct pqr_strdump {
    int64_t n;
    uint64_t* pos;
    uint64_t* off; /* n + 1 */
    uint8_t* chars;
} pqr_strdump;

typedef struct pqr_colinfo {
    char name[256];
    int32_t type;
    int32_t column_index;
    int32_t max_def_level;
    int32_t max_rep_level;
    int32_t repetition; /* -1 if absent */
    int32_t converted;  /* -1 if absent */
} pqr_colinfo;

typedef struct pqr_page_entry {
    uint64_t data_offset, data_size, row_group_idx, column_idx;
} pqr_page_entry;

/* columnar decode result (host memory owned by the library; free with pqr_columnar_free) */
typedef struct pqr_columnar {
    int32_t type;
    uint32_t width;
    uint64_t num_slots;
    int32_t has_validity;
    uint32_t n_chunks;
    const uint8_t* values;
    const uint32_t* validity;
    const uint32_t* offsets;      /* layout: see pqg.h */
    const uint64_t* char_bases;   /* n_chunks + 1 */
    const uint8_t* chars;
    uint64_t chars_size;
    const uint64_t* chunk_row_base; /* n_chunks */
    uint64_t bytes_in, bytes_out;
    float kernel_ms;
    void* owner;
} pqr_columnar;

/* descriptor tables of a column (file offsets), for callers that drive pqg.h directly */
typedef struct pqr_tables {
    uint32_t n_chunks, n_pages;
    pqg_chunk_desc* chunks;
    pqg_page_desc* pages;
    uint64_t total_slots;
} pqr_tables;

PQG_API const char* pqr_last_error(void);

PQG_API pqr_reader* pqr_open(const char* path, int device);
PQG_API pqr_reader* pqr_open_memory(const uint8_t* data, uint64_t size, int device);
PQG_API void pqr_close(pqr_reader* r);
/* Beyond the reference, off by default: SNAPPY-compressed chunks and DATA_PAGE_V2 pages of flat columns decode
 * through pqr_read_column* (pqg_plan_create_ext) instead of failing with "Only uncompressed parquet files are supported" /
 * PQG_ERR_UNSUPPORTED like the reference refuses or skips them (src/reader/column_reader.cpp:13-15,66-67).  The streaming
 * reads (pqr_read_columns_into, pqr_read_strings_into) take such chunks without per-chunk overlap; the table exports and the
 * dictionary-form reads keep refusing them. */
PQG_API void pqr_set_extensions(pqr_reader* r, int on);

PQG_API int64_t pqr_num_rows(const pqr_reader* r);
PQG_API int64_t pqr_num_row_groups(const pqr_reader* r);
PQG_API int64_t pqr_num_columns(const pqr_reader* r);
PQG_API int64_t pqr_num_pages(const pqr_reader* r);
PQG_API int64_t pqr_row_group_num_rows(const pqr_reader* r, int rg);
PQG_API int pqr_column_info(const pqr_reader* r, int col, pqr_colinfo* out);
PQG_API int pqr_find_column(const pqr_reader* r, const char* name);
PQG_API int pqr_schema_string(const pqr_reader* r, char* buf, int64_t cap);
PQG_API double pqr_page_scan_seconds(const pqr_reader* r);
PQG_API uint64_t pqr_file_size(const pqr_reader* r);

PQG_API int64_t pqr_page_index(const pqr_reader* r, pqr_page_entry* out, int64_t cap);
PQG_API int64_t pqr_read_page_data(const pqr_reader* r, int64_t id, uint8_t* buf, int64_t cap);
PQG_API int64_t pqr_read_pages_chunk(const pqr_reader* r, int64_t s, int64_t e, int64_t max_bytes,
                                     uint8_t* buf, int64_t cap);

/* the reference's Value-returning calls, dumped */
PQG_API int pqr_read_column_by_idx(pqr_reader* r, int rg, int col, pqr_valdump* out);
PQG_API int pqr_read_column(pqr_reader* r, const char* name, pqr_valdump* out);
PQG_API int pqr_read_column_rg(pqr_reader* r, const char* name, int64_t rg, pqr_valdump* out);
PQG_API int pqr_read_pages(pqr_reader* r, int rg, int col, pqr_pagedump* out);
PQG_API int pqr_string_iterator_dump(pqr_reader* r, const char* name, pqr_strdump* out);
PQG_API void pqr_valdump_free(pqr_valdump* d);
PQG_API void pqr_pagedump_free(pqr_pagedump* d);
PQG_API void pqr_strdump_free(pqr_strdump* d);

/* columnar decode: rg < 0 = all row groups */
PQG_API int pqr_read_columnar(pqr_reader* r, int col, int rg, pqr_columnar* out);
PQG_API void pqr_columnar_free(pqr_columnar* c);
PQG_API int pqr_column_tables(const pqr_reader* r, int col, int rg, pqr_tables* out);
/* Descriptor tables of SEVERAL fixed-width columns of the same value width (INT32/FLOAT or
 * INT64/DOUBLE) for ONE plan: column k of the call owns the output slots [k * S, (k + 1) * S) of
 * the plan's value / validity buffers (S = slots of one column in the selected row groups =
 * total_slots / n_cols).  Chunks are listed largest dictionary first, so that the long-running
 * CTAs of a launch start first.  What read_column does for one column (src/reader/
 * parquet_reader.cpp:133-165), several columns per launch. */
PQG_API int pqr_columns_tables(const pqr_reader* r, const int* cols, int n_cols, int rg, pqr_tables* out);
PQG_API void pqr_tables_free(pqr_tables* t);

/* Streaming read of fixed-width columns (the fast path of ParquetReader::read_column for callers
 * that take columnar buffers instead of std::vector<Value>): host file image in, caller-owned
 * host buffers out, H2D / decode / D2H pipelined per row group over all requested columns.
 * Plans (descriptor tables + device buffers) are cached per column until pqr_release_plans. */
typedef struct pqr_dst {
    void* values; uint64_t values_cap;        /* bytes */
    uint32_t* validity; uint64_t validity_cap; /* 32-bit words; may be NULL / 0 */
} pqr_dst;
typedef struct pqr_read_stats {
    uint64_t num_slots; uint32_t width; int32_t has_validity;
    uint64_t bytes_in, bytes_out, h2d_bytes, d2h_bytes;
} pqr_read_stats;
PQG_API int pqr_read_columns_into(pqr_reader* r, const int32_t* cols, int32_t n_cols, int32_t rg,
                                  const pqr_dst* dsts, pqr_read_stats* stats);
/* Dictionary-form read (late materialisation) of a BYTE_ARRAY column that is dictionary-encoded
 * throughout: uint32 dictionary index per slot (0 for nulls) + validity, pipelined like
 * pqr_read_columns_into; the string of slot i is entry indices[i] of the dictionary of i's row
 * group (pqr_chunk_dictionary, parsed on the host from the dictionary page).  Fails with
 * "... not dictionary-encoded throughout" for columns with PLAIN pages. */
PQG_API int pqr_read_dictionary_indices_into(pqr_reader* r, int32_t col, int64_t rg_begin, int64_t rg_end,
                                             const pqr_dst* dst, pqr_read_stats* stats);
/* Pipelined read of a BYTE_ARRAY column into the caller's (ideally pinned) buffers -- no counterpart in the reference, whose
 * read_column materialises one Value per slot (src/reader/parquet_reader.cpp:125-165).  One cached plan per row group,
 * alternating between two contexts of the device: the upload and size pass of row group k + 1 overlap the copy pass and the
 * D2H of row group k.  Layout as pqr_read_columnar: `offsets` holds total_slots + n_chunks entries (chunk c owns
 * [row_base(c) + c, row_base(c + 1) + c], relative to the chunk's chars), chunk c's bytes start at chars[char_bases[c]]
 * (n_chunks + 1 entries), `validity` covers the whole range (may be NULL).  A chars buffer that is too small fails with the
 * number of bytes needed so far in the message; the uncompressed size of the column's pages is always enough for PLAIN
 * columns, rows x longest dictionary entry for dictionary columns. */
typedef struct pqr_strings_dst {
    uint32_t* offsets; uint64_t offsets_cap;    /* entries */
    uint8_t* chars; uint64_t chars_cap;         /* bytes */
    uint32_t* validity; uint64_t validity_cap;  /* words */
    uint64_t* char_bases; uint64_t char_bases_cap;
} pqr_strings_dst;
typedef struct pqr_strings_stats {
    uint64_t num_slots, n_chunks, chars_size;
    int32_t has_validity, reserved;
    uint64_t bytes_in, bytes_out, h2d_bytes, d2h_bytes;
} pqr_strings_stats;
PQG_API int pqr_read_strings_into(pqr_reader* r, int32_t col, int64_t rg_begin, int64_t rg_end, const pqr_strings_dst* dst,
                                  pqr_strings_stats* stats);
/* offsets: n_entries + 1 uint32 into chars; pass NULL buffers to query the sizes first */
PQG_API int pqr_chunk_dictionary(const pqr_reader* r, int32_t col, int64_t rg, uint32_t* offsets, int64_t offsets_cap,
                                 uint8_t* chars, int64_t chars_cap, int64_t* n_entries, int64_t* n_bytes);
PQG_API void pqr_release_plans(pqr_reader* r);

/* chunk-index prototype (src/main.cpp:21-32): tuple_to_chunk has num_rows entries;
 * returns "Total chunks" */
PQG_API int64_t pqr_chunk_index(pqr_reader* r, const char* name, uint64_t chunk_size,
                                uint64_t* tuple_to_chunk, int64_t num_rows);
/* regex page pruning: bits[] one byte per data page of the column in global page order
 * (1 = some value satisfies the predicate, 0 = page can be pruned); returns page count */
PQG_API int64_t pqr_regex_prune(pqr_reader* r, int col, const char* pattern, int neg,
                                uint8_t* bits, int64_t cap, float* kernel_ms);
/* index_test: page-level 4 KB chunk index of a column; returns the number of chunks */
PQG_API int64_t pqr_page_chunk_index(pqr_reader* r, int col, uint64_t chunk_size, uint32_t* page_chunk,
                                     uint32_t* page_off, uint32_t* chunk_first_page, int64_t cap,
                                     int64_t* first_global_page, int64_t* n_col_pages);

/* ---- multi-GPU sharding (row groups are independent: no device-side collective) ----------
 * One reader per GPU / process; every shard handles a contiguous run of row groups; the host
 * concatenates page bitmaps in shard order and chains the chunk index through carry_in /
 * carry_out (SURVEY.md section 8 e). */
/* out_begin: n_shards + 1 row-group boundaries, balanced by the byte size of the column's chunks
 * (col < 0: all columns) */
PQG_API int pqr_shard_row_groups(const pqr_reader* r, int col, int n_shards, int32_t* out_begin);
PQG_API int64_t pqr_regex_prune_rgs(pqr_reader* r, int col, int64_t rg_begin, int64_t rg_end, const char* pattern, int neg,
                                    uint8_t* bits, int64_t cap, float* kernel_ms);
/* ids: id_base + shard-local chunk id for the shard's non-null rows (local 0 = the chunk the
 * previous shard left open), 0 for nulls; returns the shard's chunk count n (next shard:
 * id_base + n - 1, carry_in = carry_out), <0 on error */
PQG_API int64_t pqr_chunk_index_rgs(pqr_reader* r, const char* name, int64_t rg_begin, int64_t rg_end, uint64_t chunk_size,
                                    uint64_t carry_in, uint32_t id_base, uint32_t* ids, int64_t cap, uint64_t* carry_out);
/* chunk index of a shard in phases (pqg.h: pqg_chunk_index_prepare / _stitch / _emit): prepare on every shard at once,
 * stitch in shard order (microseconds, host), emit on every shard at once */
typedef struct pqr_chunk_job pqr_chunk_job;
PQG_API pqr_chunk_job* pqr_chunk_index_prepare_rgs(pqr_reader* r, const char* name, int64_t rg_begin, int64_t rg_end, uint64_t chunk_size,
                                                   uint64_t* num_slots, float* decode_ms, float* prepare_ms);
PQG_API int64_t pqr_chunk_index_stitch(pqr_chunk_job* job, uint64_t carry_in, uint64_t* carry_out); /* the shard's chunk count, <0 on error */
PQG_API int pqr_chunk_index_emit(pqr_chunk_job* job, uint32_t id_base, uint32_t* ids, int64_t cap, float* kernel_ms);
PQG_API void pqr_chunk_job_free(pqr_chunk_job* job);
/* descriptor tables of the row groups [rg_begin, rg_end) of a column (file offsets) */
PQG_API int pqr_column_tables_rgs(const pqr_reader* r, int col, int64_t rg_begin, int64_t rg_end, pqr_tables* out);
PQG_API int pqr_read_columns_into_rgs(pqr_reader* r, const int32_t* cols, int32_t n_cols, int64_t rg_begin, int64_t rg_end,
                                      const pqr_dst* dsts, pqr_read_stats* stats);

#ifdef __cplusplus
}
#endif
#endif
