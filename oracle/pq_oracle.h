/*
 * pq_oracle.h -- TEST INFRASTRUCTURE: CPU restatement (plain C) of the reference's
 * Parquet page-decode path, used only as a checker.
 *
 * PARITY PINNING: this restatement is pinned against the reference itself compiled here
 * (oracle/_ref/libpqref.so, built by `make -C oracle ref` from /root/reference sources) in
 * tests/test_oracle_vs_ref.py, and against the committed golden fixtures in tests/golden/
 * (written by the reference's ParquetWriter and dumped by the reference's ParquetReader;
 * generator: tests/golden/make_golden.py).  The regex page-pruning rows (SURVEY 8 a-19,
 * a-20) have no reference source: for those the oracle is "parity unpinned" and the frozen
 * spec in SURVEY.md section 8 is the contract (see regex_oracle.c).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may import, call, link or execute anything under oracle/.  The product (libpqg.so) never
 * does; it fails loudly when its CUDA path is unavailable.
 *
 * All file:line citations are relative to /root/reference/.
 */
#ifndef PQ_ORACLE_H
#define PQ_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#include "valdump.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_file orc_file;

typedef struct orc_colinfo {
    char name[256];
    int32_t type;
    int32_t column_index;
    int32_t max_def_level;
    int32_t max_rep_level;
    int32_t repetition; /* -1 if absent */
    int32_t converted;  /* -1 if absent */
} orc_colinfo;

typedef struct orc_page_entry {
    uint64_t data_offset, data_size, row_group_idx, column_idx;
} orc_page_entry;

typedef struct orc_strdump {
    int64_t n;
    uint64_t* pos;
    uint64_t* off; /* n + 1 */
    uint8_t* chars;
} orc_strdump;

const char* orc_last_error(void);

/* ParquetReader::open (src/reader/parquet_reader.cpp:14-61) */
orc_file* orc_open(const char* path);
orc_file* orc_open_mem(const uint8_t* data, size_t size); /* borrows data */
void orc_close(orc_file* f);

int64_t orc_num_rows(const orc_file* f);
int64_t orc_num_row_groups(const orc_file* f);
int64_t orc_num_columns(const orc_file* f);
int64_t orc_num_pages(const orc_file* f);
int64_t orc_row_group_num_rows(const orc_file* f, int rg);
int orc_column_info(const orc_file* f, int col, orc_colinfo* out);
int orc_find_column(const orc_file* f, const char* name);

/* read_column* (src/reader/parquet_reader.cpp:125-165) over ColumnReader::read_all
 * (src/reader/column_reader.cpp:18-71) */
int orc_read_column_by_idx(orc_file* f, int rg, int col, valdump* out);
int orc_read_column(orc_file* f, const char* name, valdump* out);
/* ColumnReader::read_pages (src/reader/column_reader.cpp:73-126) */
int orc_read_pages(orc_file* f, int rg, int col, pagedump* out);

/* page index + raw page API (src/reader/parquet_reader.cpp:182-238,559-605) */
int64_t orc_page_index(const orc_file* f, orc_page_entry* out, int64_t cap);
int64_t orc_read_page_data(orc_file* f, int64_t id, uint8_t* buf, int64_t cap);
int64_t orc_read_pages_chunk(orc_file* f, int64_t s, int64_t e, int64_t max_bytes,
                             uint8_t* buf, int64_t cap);

/* StringColumnIterator (src/reader/parquet_reader.cpp:282-465) drained */
int orc_string_iterator_dump(orc_file* f, const char* name, orc_strdump* out);
void orc_strdump_free(orc_strdump* d);

/* chunk-index prototype (src/main.cpp:21-32); returns "Total chunks" */
int64_t orc_chunk_index(orc_file* f, const char* name, uint64_t chunk_size,
                        uint64_t* tuple_to_chunk, int64_t num_rows);

/* page-level 4 KB chunk index, frozen spec SURVEY.md 8 a-20 (README.md:66-72): the
 * column's data pages in global-id order are packed greedily, a chunk closes when its
 * byte size >= chunk_size.  page_chunk / page_off have one entry per data page of the
 * column; chunk_first_page receives the first column-local page of each chunk (cap
 * entries available).  Returns the number of chunks. */
int64_t orc_page_chunk_index(const orc_file* f, int col, uint64_t chunk_size,
                             uint32_t* page_chunk, uint32_t* page_off,
                             uint32_t* chunk_first_page, int64_t cap,
                             int64_t* first_global_page, int64_t* n_col_pages);

/* RleDecoder::get_batch (include/reader/rle_decoder.hpp:17-95).  `size` bounds header
 * parsing like the reference's size_; `avail` (>= size) is how many bytes are readable
 * behind `data` (the reference reads literal bits without any bound, :59-62; past `avail`
 * the restatement reads zeros). */
void orc_rle_decode_i32(const uint8_t* data, uint32_t size, uint32_t avail, int bit_width,
                        int32_t* out, uint32_t count);
void orc_rle_decode_i16(const uint8_t* data, uint32_t size, uint32_t avail, int bit_width,
                        int16_t* out, uint32_t count);

void orc_valdump_free(valdump* d);
void orc_pagedump_free(pagedump* d);

/* ---- regex page pruning, frozen spec SURVEY.md 8 a-19 (regex_oracle.c) ---- */
/* Backtracking matcher over the supported RE2 subset; search (partial-match) semantics.
 * Returns 1 match, 0 no match, <0 unsupported/parse error (message in orc_last_error). */
int orc_regex_search(const char* pattern, const uint8_t* text, int64_t len);
const char* orc_regex_last_error(void);
/* per data page of (all row groups of) column `col`: bit = OR over non-null values of
 * (neg ? !match : match).  bits[] has one byte per data page of the column in global page
 * order; returns the number of pages, <0 on error. */
int64_t orc_regex_prune(orc_file* f, int col, const char* pattern, int neg,
                        uint8_t* bits, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif
