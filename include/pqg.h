/*
 * pqg.h -- C-ABI of the B200 (sm_100a) Parquet page decoder / page-pruning scanner.
 *
 * This is the drop-in boundary beneath the reference's reader classes.  The reference
 * (sputnik89/duckdb-parquet-parser) has no FFI layer of its own: its hot path is the C++
 * call chain
 *     ParquetReader::read_column / read_column_by_idx   src/reader/parquet_reader.cpp:125-165
 *       -> ColumnReader::read_all / read_pages          src/reader/column_reader.cpp:18-126
 *            -> read_dictionary_page / read_data_page   src/reader/column_reader.cpp:128-225
 *                 -> RleDecoder::get_batch              include/reader/rle_decoder.hpp:17-95
 *                 -> read_plain_value                   src/reader/column_reader.cpp:227-268
 *     StringColumnIterator::decode_next_page            src/reader/parquet_reader.cpp:347-465
 *     chunk-index prototype                             src/main.cpp:21-32
 *     --regex-column / --regex / --neg-regex, index_test  README.md:54-72 (source absent)
 * Each entry point below names the reference code it replaces.  Thrift footer / page-header
 * parsing and page-offset indexing stay in host C++ (include/pqg_reader.h); the host emits
 * the flat descriptor tables defined here and calls these functions.
 *
 * Rules of the boundary
 *   - plain C types only: pointers, sizes, POD structs; no C++/torch types;
 *   - every function returns a pqg_status (0 = ok) and never throws; the message of the
 *     last failure on a context is available from pqg_last_error();
 *   - there is NO CPU fallback: without a usable CUDA device every compute entry point
 *     fails with PQG_ERR_CUDA;
 *   - a context owns one CUDA stream (or borrows the caller's); calls on one context are
 *     serialised on it; use one context per GPU / host thread for multi-GPU;
 *   - `image` buffers are an exact copy of a byte range of the parquet file; the
 *     allocation behind a wrapped device pointer must be readable for 64 bytes past
 *     `size` (pqg_upload pads by itself).
 */
#ifndef PQG_H
#define PQG_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PQG_API __attribute__((visibility("default")))

typedef enum pqg_status {
    PQG_OK = 0,
    PQG_ERR_CUDA = 1,        /* no device / CUDA runtime failure */
    PQG_ERR_ARG = 2,         /* invalid argument */
    PQG_ERR_UNSUPPORTED = 3, /* compressed / DATA_PAGE_V2 / FLBA / DELTA_* / bit width > 32 ... */
    PQG_ERR_PAGE = 4,        /* a page failed to decode (truncated, bad run, ...) */
    PQG_ERR_REGEX = 5,       /* pattern rejected by the DFA compiler */
    PQG_ERR_NOMEM = 6
} pqg_status;

/* per-page decode error codes (pqg_page_error.code) */
enum {
    PQG_PAGE_OK = 0,
    PQG_PAGE_TRUNCATED = 1,     /* ByteBuffer::check would throw (include/common.hpp:162-168) */
    PQG_PAGE_BAD_BIT_WIDTH = 2, /* dictionary index bit width > 32 */
    PQG_PAGE_BAD_RUN = 3,       /* zero-length RLE run / empty literal run: reference UB */
    PQG_PAGE_DICT_TRUNCATED = 4,/* dictionary page shorter than its entries */
    PQG_PAGE_CHARS_OVERFLOW = 5,/* a column chunk decodes to >= 4 GiB of string bytes */
    PQG_PAGE_LAYOUT = 6,        /* internal: a string page broke the byte count the fast path assumed (the plan re-runs with the exact size pass) */
    PQG_PAGE_CHARS_CAP = 7,     /* internal: the chars buffer of the previous run is too small (the plan re-sizes it and re-runs) */
    PQG_PAGE_DECOMPRESS = 8     /* pqg_plan_create_ext plans: a page did not decompress to its uncompressed_page_size (corrupt SNAPPY data / wrong sizes) */
};

/* Physical types: reference enum ParquetType (include/common.hpp:16-25). */
enum {
    PQG_BOOLEAN = 0, PQG_INT32 = 1, PQG_INT64 = 2, PQG_INT96 = 3, PQG_FLOAT = 4,
    PQG_DOUBLE = 5, PQG_BYTE_ARRAY = 6, PQG_FIXED_LEN_BYTE_ARRAY = 7
};

/* One column chunk (one column of one row group).  Replaces the state ColumnReader keeps
 * per chunk (src/reader/column_reader.cpp:3-30): type, levels, dictionary page. */
typedef struct pqg_chunk_desc {
    uint64_t dict_off;        /* image offset of the dictionary page PAYLOAD (has_dict) */
    uint64_t out_row_base;    /* first output slot of this chunk in the column */
    uint64_t num_values;      /* level entries (= output slots) in this chunk */
    uint32_t dict_size;       /* dictionary payload bytes */
    uint32_t dict_num_values; /* DictionaryPageHeader.num_values */
    uint32_t first_page;      /* first entry of this chunk in the page table */
    uint32_t n_pages;         /* data pages of this chunk (contiguous in the page table) */
    uint32_t row_group;       /* informational */
    uint32_t column;          /* informational */
    int16_t max_def;          /* ColumnInfo.max_def_level */
    int16_t max_rep;          /* ColumnInfo.max_rep_level */
    uint8_t phys_type;        /* PQG_* physical type */
    uint8_t has_dict;         /* a DICTIONARY_PAGE precedes the data pages */
    uint8_t reserved[2];
} pqg_chunk_desc;

#define PQG_PAGE_FLAG_DICT 1u /* DataPageHeader.encoding is PLAIN_DICTIONARY / RLE_DICTIONARY */
#define PQG_PAGE_FLAG_V2 2u   /* the page is a DATA_PAGE_V2: pqg_plan_create rejects the plan (PQG_ERR_UNSUPPORTED) */
#define PQG_PAGE_FLAG_LEVELS_SEEN 8u /* the producer of the table looked at the page's level bytes: PQG_PAGE_FLAG_NO_NULLS is set where it holds */
#define PQG_PAGE_FLAG_NO_NULLS 4u /* routing hint (optional; the kernels verify it): the definition levels of this page of an OPTIONAL
                                   * flat column are ONE RLE run of level 1 covering all its values -- what nullable-by-default writers
                                   * emit for a page without nulls.  Such pages decode as REQUIRED pages in the tile kernel whatever
                                   * their size; OPTIONAL pages of more than 1024 values with LEVELS_SEEN and without NO_NULLS go straight to the block
                                   * decode of oversized pages instead of being looked at and handed over by the tile kernel. */
/* bits 8..15 of pqg_page_desc.flags: the Encoding enum value of DataPageHeader.encoding (0 PLAIN,
 * 2 PLAIN_DICTIONARY, 8 RLE_DICTIONARY are decoded; DELTA_BINARY_PACKED 5, DELTA_LENGTH_BYTE_ARRAY 6,
 * DELTA_BYTE_ARRAY 7, BYTE_STREAM_SPLIT 9 and anything else make pqg_plan_create fail with
 * PQG_ERR_UNSUPPORTED -- the reference decodes such pages as PLAIN and returns garbage,
 * src/reader/column_reader.cpp:173-222) */
#define PQG_PAGE_ENCODING(flags) (((flags) >> 8) & 0xffu)
#define PQG_PAGE_FLAGS(dict, encoding) (((dict) ? PQG_PAGE_FLAG_DICT : 0u) | (((uint32_t)(encoding) & 0xffu) << 8))

/* One DATA_PAGE.  Replaces PageHeader + the per-page cursor of read_all
 * (src/reader/column_reader.cpp:32-64) and PageIndexEntry
 * (include/reader/parquet_reader.hpp:12-17). */
typedef struct pqg_page_desc {
    uint64_t payload_off;  /* image offset of the first payload byte (after the header) */
    uint64_t out_row_base; /* first output slot of this page in the column */
    uint32_t payload_size; /* compressed_page_size */
    uint32_t num_values;   /* DataPageHeader.num_values */
    uint32_t chunk_idx;    /* index into the chunk table */
    uint32_t flags;        /* PQG_PAGE_FLAG_* | encoding << 8 */
} pqg_page_desc;

typedef struct pqg_page_error {
    uint32_t count; /* pages that failed */
    uint32_t page;  /* page-table index of one failing page (the lowest seen) */
    uint32_t code;  /* PQG_PAGE_* */
    uint32_t pos, need, size; /* ByteBuffer::check numbers where applicable */
} pqg_page_error;

typedef struct pqg_ctx pqg_ctx;   /* one GPU + one stream + caches */
typedef struct pqg_buf pqg_buf;   /* a device-resident byte range of the file */
typedef struct pqg_plan pqg_plan; /* descriptor tables + outputs of one column decode */
typedef struct pqg_dfa pqg_dfa;   /* host-compiled regex automaton */

/* per-kernel-family device times of the last pqg_plan_run, CUDA events on the context's
 * stream (milliseconds); only recorded after pqg_ctx_set_profiling(ctx, 1). */
typedef struct pqg_timings {
    float dict_ms;    /* dictionary preparation */
    float fixed_ms;   /* fixed-width plans: the TMA tile kernel (PLAIN copy + regular dictionary pages) */
    float str_size_ms;/* string pass 1: per-page byte totals + scan */
    float str_copy_ms;/* string pass 2: offsets + chars */
    float total_ms;
    uint32_t launches;/* kernels launched by the run */
    float general_ms; /* fixed-width plans: the general kernel (levels, RLE runs, big pages) */
    uint32_t tile_launches; /* fixed-width plans: launches of the tile kernel inside fixed_ms */
} pqg_timings;

/* ---- context ------------------------------------------------------------------------- */
/* stream: a cudaStream_t to borrow (e.g. torch's current stream), or NULL to create one. */
PQG_API int pqg_ctx_create(int device, void* stream, pqg_ctx** out);
PQG_API void pqg_ctx_destroy(pqg_ctx* ctx);
PQG_API const char* pqg_last_error(const pqg_ctx* ctx); /* ctx may be NULL: creation errors */
PQG_API int pqg_ctx_sync(pqg_ctx* ctx);
PQG_API int pqg_ctx_set_profiling(pqg_ctx* ctx, int on);
PQG_API int pqg_device_count(void); /* 0 when no CUDA device is usable */
PQG_API uint64_t pqg_kernel_launches(const pqg_ctx* ctx); /* total kernels launched so far */

/* ---- file image: replaces ParquetReader::read_range (parquet_reader.cpp:173-178) ------ */
PQG_API int pqg_upload(pqg_ctx* ctx, const void* host_bytes, uint64_t size, pqg_buf** out);
/* wrap caller-owned device memory (16-byte aligned).  `capacity` = bytes readable behind dev_ptr:
 * the kernels read whole 16-byte vectors and TMA tiles, so capacity must be >= size + 64
 * (PQG_ERR_ARG otherwise); the bytes past `size` are never interpreted. */
PQG_API int pqg_wrap_device(pqg_ctx* ctx, const void* dev_ptr, uint64_t size, uint64_t capacity, pqg_buf** out);
/* an empty device image to be filled range by range (a column's chunks packed together):
 * asynchronous H2D of `n` host bytes to image offset `dst_off` on the context's stream */
PQG_API int pqg_buf_alloc(pqg_ctx* ctx, uint64_t size, pqg_buf** out);
PQG_API int pqg_buf_write(pqg_ctx* ctx, pqg_buf* buf, uint64_t dst_off, const void* host_bytes, uint64_t n);
PQG_API uint64_t pqg_buf_size(const pqg_buf* buf);
PQG_API void pqg_buf_free(pqg_ctx* ctx, pqg_buf* buf);
PQG_API const void* pqg_buf_device_ptr(const pqg_buf* buf);
/* pinned host staging helpers (cudaHostAlloc / cudaFreeHost) for callers without torch */
PQG_API void* pqg_host_alloc(uint64_t size);
PQG_API void pqg_host_free(void* p);

/* ---- decode: replaces ColumnReader::read_all for a set of chunks of ONE column -------- */
/* All chunks must share phys_type.  chunks / pages are HOST arrays; they are copied.    */
PQG_API int pqg_plan_create(pqg_ctx* ctx, const pqg_buf* image, const pqg_chunk_desc* chunks,
                            uint32_t n_chunks, const pqg_page_desc* pages, uint32_t n_pages,
                            pqg_plan** out);
/* Dictionary-form output of a BYTE_ARRAY column that is dictionary-encoded throughout (late
 * materialisation, Arrow DictionaryArray style): the plan behaves like a 4-byte fixed-width plan
 * whose values are the uint32 DICTIONARY INDEX of every slot (0 for nulls, validity as usual;
 * an index beyond its chunk's dictionary is a null, like the reference treats it).  The
 * dictionaries themselves are the chunks' dictionary pages, which the caller already has.
 * All fixed-width entry points apply (run, run_pipelined, download).  PQG_ERR_UNSUPPORTED when
 * some chunk has no dictionary or falls back to PLAIN pages. */
PQG_API int pqg_plan_create_dict_indices(pqg_ctx* ctx, const pqg_buf* image, const pqg_chunk_desc* chunks,
                                         uint32_t n_chunks, const pqg_page_desc* pages, uint32_t n_pages,
                                         pqg_plan** out);
/* ---- beyond the reference: DATA_PAGE_V2 and SNAPPY pages ---------------- */
/* The reference refuses compressed chunks (src/reader/column_reader.cpp:13-15) and skips DATA_PAGE_V2
 * (:66-67); pqg_plan_create does the same (PQG_ERR_UNSUPPORTED).  This entry point takes, next to the
 * usual tables, what those pages need from their headers.  The plan keeps its OWN device image: at the
 * start of every run one kernel rewrites each dictionary / data page of `image` into the DATA_PAGE
 * layout the decode kernels know (length word in front of V2 definition levels, SNAPPY blocks decoded),
 * then the run proceeds as for pqg_plan_create.  In `pages`, payload_off / payload_size describe the
 * bytes as STORED (compressed size), flags carry no PQG_PAGE_FLAG_V2; in `chunks`, dict_size likewise.
 * Flat columns only (max_rep == 0: PQG_ERR_UNSUPPORTED otherwise); codecs other than SNAPPY:
 * PQG_ERR_UNSUPPORTED.  pqg_plan_run_pipelined and pqg_plan_set_image do not apply to such plans. */
#define PQG_CODEC_UNCOMPRESSED 0u
#define PQG_CODEC_SNAPPY 1u
#define PQG_PAGE_EXT_V2 1u
typedef struct pqg_page_ext {
    uint32_t uncompressed_size; /* PageHeader.uncompressed_page_size (levels included) */
    uint32_t def_len;           /* DATA_PAGE_V2: definition_levels_byte_length (DATA_PAGE: 0) */
    uint32_t rep_len;           /* DATA_PAGE_V2: repetition_levels_byte_length (DATA_PAGE: 0) */
    uint32_t kind;              /* PQG_PAGE_EXT_V2 | codec << 8 -- codec of the compressed part: the whole payload of a DATA_PAGE, the
                                 * value section of a DATA_PAGE_V2 (PQG_CODEC_UNCOMPRESSED when its is_compressed is false) */
} pqg_page_ext;
typedef struct pqg_chunk_ext {
    uint32_t dict_uncompressed_size; /* PageHeader.uncompressed_page_size of the dictionary page */
    uint32_t dict_codec;             /* PQG_CODEC_* of the dictionary page */
} pqg_chunk_ext;
PQG_API int pqg_plan_create_ext(pqg_ctx* ctx, const pqg_buf* image, const pqg_chunk_desc* chunks, uint32_t n_chunks,
                                const pqg_page_desc* pages, uint32_t n_pages, const pqg_page_ext* page_ext,
                                const pqg_chunk_ext* chunk_ext, pqg_plan** out);
PQG_API void pqg_plan_destroy(pqg_ctx* ctx, pqg_plan* plan);
/* tuning switches of a plan (A/B measurements; the defaults are the measured best).  PQG_OPT_PARTITIONED_DICT (default 0):
 * dictionaries of 32 KB .. 512 KB of values in REQUIRED-only 4/8-byte plans are split over the shared memories of 1..4
 * sibling 1024-thread CTAs that read the same page tiles and emit only the values of their part, instead of being gathered
 * from L2.  Measured on B200 (bench.py, cfg2 i64_d16 / f64_d16, 100 M values): 0.60 ms against 0.49 ms for the L2 gather
 * (256-thread CTAs: 1.21 ms), so it stays off; kept for A/B runs. */
enum { PQG_OPT_PARTITIONED_DICT = 1,
       PQG_OPT_REGEX_TILE_BARRIER = 2 /* regex scan: CTA-wide barrier per page tile (default 1) instead of the last-warp refill */ };
PQG_API int pqg_plan_set_option(pqg_plan* plan, int option, int value);
/* re-point a plan at another resident image with the same layout (pipelined ingest) */
PQG_API int pqg_plan_set_image(pqg_ctx* ctx, pqg_plan* plan, const pqg_buf* image);
/* enqueue the decode on the context's stream.  Asynchronous, except that the first run of
 * a BYTE_ARRAY plan waits for the size pass to learn the output size. */
PQG_API int pqg_plan_run(pqg_ctx* ctx, pqg_plan* plan);
/* Pipelined decode of a fixed-width plan straight from HOST memory into HOST memory.  For every
 * chunk of the plan, in table order: H2D of the chunk's byte ranges on the context's copy-in
 * stream -> decode of that chunk on the compute stream -> D2H of its rows (and validity words) on
 * the copy-out stream, chained by events, so both transfer directions overlap the kernels.
 * `image` must be the plan's image and come from pqg_buf_alloc; `ranges` (sorted by chunk)
 * name the host bytes of each chunk and where they go in the image.  host_values /
 * host_validity (may be NULL) receive the decoded column; pinned memory keeps the copies
 * asynchronous.  Returns after enqueueing: several plans can be in flight on one context;
 * pqg_plan_finish waits for this plan and reports page errors.  BYTE_ARRAY plans are refused
 * (PQG_ERR_UNSUPPORTED): their output size needs the size pass first. */
typedef struct pqg_h2d_range {
    const void* host;    /* source bytes */
    uint64_t image_off;  /* destination offset in the plan's image */
    uint64_t len;
    uint32_t chunk;      /* table chunk the bytes belong to */
    uint32_t reserved;
} pqg_h2d_range;
PQG_API int pqg_plan_run_pipelined(pqg_ctx* ctx, pqg_plan* plan, pqg_buf* image, const pqg_h2d_range* ranges,
                                   uint32_t n_ranges, void* host_values, uint32_t* host_validity);
/* wait for the run and report page errors; PQG_ERR_PAGE when err->count > 0.
 * An out-of-range dictionary index is a NULL in the reference whatever the column's repetition
 * (src/reader/column_reader.cpp:190-194).  A plan whose chunks are all REQUIRED carries no validity
 * bitmap; when such an index shows up, pqg_plan_finish adds the bitmap and decodes the plan again
 * (pqg_plan_validity is non-NULL afterwards).  After pqg_plan_run_pipelined that is not possible
 * (the host buffers are the caller's): PQG_ERR_PAGE "... out-of-range dictionary index ...". */
PQG_API int pqg_plan_finish(pqg_ctx* ctx, pqg_plan* plan, pqg_page_error* err);
PQG_API int pqg_plan_timings(const pqg_plan* plan, pqg_timings* out);
/* mean over the last `last_n` (0 = all kept, at most 8) profiled runs whose events have
 * completed; no synchronisation between runs is needed to collect them */
PQG_API int pqg_plan_timings_avg(const pqg_plan* plan, uint32_t last_n, pqg_timings* out, uint32_t* n_used);

/* Output layout (device memory owned by the plan), n = total slots of the column:
 *   values   : n * width bytes, width = 1 BOOLEAN, 4 INT32/FLOAT, 8 INT64/DOUBLE, 12 INT96;
 *              null slots hold 0.  Absent for BYTE_ARRAY.
 *   validity : ceil(n/32) uint32 words, bit (i & 31) of word i >> 5 set = slot i non-null.
 *              NULL pointer when every chunk has max_def == 0 (all slots valid).
 *   offsets  : BYTE_ARRAY only.  Chunk c owns entries [row_base_c + c, row_base_c + c + n_c]
 *              (n_c + 1 uint32, Arrow style), byte offsets relative to the chunk's chars.
 *   chars    : BYTE_ARRAY only.  Chunk c's bytes start at char_base[c]; char_base has
 *              n_chunks + 1 uint64 entries (host copy via pqg_plan_char_bases).
 */
PQG_API uint64_t pqg_plan_num_slots(const pqg_plan* plan);
PQG_API uint32_t pqg_plan_value_width(const pqg_plan* plan);
PQG_API const void* pqg_plan_values(const pqg_plan* plan);
PQG_API const uint32_t* pqg_plan_validity(const pqg_plan* plan);
PQG_API const uint32_t* pqg_plan_offsets(const pqg_plan* plan);
PQG_API const uint8_t* pqg_plan_chars(const pqg_plan* plan);
PQG_API uint64_t pqg_plan_chars_size(const pqg_plan* plan);
PQG_API int pqg_plan_char_bases(pqg_ctx* ctx, const pqg_plan* plan, uint64_t* out, uint32_t n);
PQG_API uint64_t pqg_plan_bytes_in(const pqg_plan* plan);  /* page + dictionary payload bytes */
PQG_API uint64_t pqg_plan_bytes_out(const pqg_plan* plan); /* algorithmic output bytes */
/* D2H of the outputs into caller (ideally pinned) memory; any pointer may be NULL.
 * Asynchronous on the context's stream; call pqg_ctx_sync before reading. */
PQG_API int pqg_plan_download(pqg_ctx* ctx, const pqg_plan* plan, void* values, uint32_t* validity,
                              uint32_t* offsets, uint8_t* chars);

/* ---- device-resident consumer: predicate on a decoded column ------------- */
/* Rows of a decoded (run + finished) fixed-width plan whose value compares true against a constant -- the decoded column
 * never leaves the device, one bit per slot comes back (row_bits: ceil(slots / 32) host words, may be NULL) together with
 * the number of matches.  value_type: PQG_INT32 / PQG_INT64 / PQG_FLOAT / PQG_DOUBLE, must match the plan's value width;
 * `constant` points at one value of that type; nulls never match; NaN compares like IEEE (only != is true).  Synchronous.
 * The reference has no counterpart (it materialises Values on the host). */
enum { PQG_CMP_EQ = 0, PQG_CMP_NE = 1, PQG_CMP_LT = 2, PQG_CMP_LE = 3, PQG_CMP_GT = 4, PQG_CMP_GE = 5 };
PQG_API int pqg_plan_filter(pqg_ctx* ctx, pqg_plan* plan, int value_type, int op, const void* constant,
                            uint32_t* row_bits, uint64_t* n_match, float* kernel_ms);

/* ---- regex page pruning: replaces the parser's --regex-column mode (README.md:54-64) -- */
/* RE2-syntax subset compiled on the host to a byte DFA; unsupported syntax is rejected
 * with an explicit message in err (PQG_ERR_REGEX).  No device needed. */
PQG_API int pqg_regex_compile(const char* pattern, pqg_dfa** out, char* err, size_t errlen);
PQG_API void pqg_dfa_free(pqg_dfa* dfa);
PQG_API uint32_t pqg_dfa_num_states(const pqg_dfa* dfa);
/* host-side run of the same tables (test hook, and the CLI's dictionary pre-check) */
PQG_API int pqg_dfa_match_host(const pqg_dfa* dfa, const uint8_t* text, uint64_t len);
/* page_bits: DEVICE or HOST? -> host array of ceil(n_pages/32) uint32 words written after
 * the scan: bit p set = page p (page-table order) holds a value v with (neg ? !m(v) : m(v)).
 * The plan must be a BYTE_ARRAY plan; nulls never match.  Synchronous. */
PQG_API int pqg_regex_scan(pqg_ctx* ctx, pqg_plan* plan, const pqg_dfa* dfa, int neg,
                           uint32_t* page_bits, float* kernel_ms);

/* ---- 4 KB chunk index: replaces src/main.cpp:21-32 and index_test (README.md:66-72) ---- */
/* Tuple-level map over a decoded BYTE_ARRAY plan (run + finish first): weight of a
 * non-null value = decimal digits of its length + its length; greedy chunks close at
 * >= chunk_size.  `carry_in` = bytes already in the open chunk when this shard starts
 * (0 on the first shard; multi-GPU shards are chained by the host through `carry_out` and
 * chunk counts).  tuple_to_chunk: HOST array of num_slots uint32: id_base + shard-local chunk
 * id for non-null values (local id 0 = the open chunk carried in), 0 for nulls.
 * n_chunks = local id of the last value + 1; the next shard's id_base = id_base + n_chunks - 1. */
PQG_API int pqg_chunk_index(pqg_ctx* ctx, pqg_plan* plan, uint64_t chunk_size, uint64_t carry_in,
                            uint32_t id_base, uint32_t* tuple_to_chunk, uint64_t* n_chunks,
                            uint64_t* carry_out, float* kernel_ms);
/* The same in three phases, for shards of ONE column decoded on several GPUs (SURVEY.md section 8 e): everything that
 * is heavy -- weights, prefix sums, the speculative walks (prepare) and the materialisation of cuts and ids (emit) --
 * needs no knowledge of the other shards and runs on all GPUs at once; only `stitch` (one 8-byte device lookup + one
 * table lookup per 256 KB of weight on the host) runs in shard order, handing carry_out / chunk counts to the next shard:
 *     prepare(shard r) on every r   ->   for r = 0..: stitch(r, carry_in = carry_out(r-1))   ->   emit(r, id_base_r) on every r
 * with id_base_r = sum over earlier shards of (n_chunks - 1).  pqg_chunk_index == prepare + stitch + emit. */
typedef struct pqg_chunk_job pqg_chunk_job;
PQG_API int pqg_chunk_index_prepare(pqg_ctx* ctx, pqg_plan* plan, uint64_t chunk_size, pqg_chunk_job** out, float* kernel_ms);
PQG_API int pqg_chunk_index_stitch(pqg_ctx* ctx, pqg_chunk_job* job, uint64_t carry_in, uint64_t* n_chunks, uint64_t* carry_out);
/* tuple_to_chunk: HOST array of num_slots uint32, or NULL to keep the ids on the device (pqg_chunk_job_ids) */
PQG_API int pqg_chunk_index_emit(pqg_ctx* ctx, pqg_chunk_job* job, uint32_t id_base, uint32_t* tuple_to_chunk, float* kernel_ms);
PQG_API const uint32_t* pqg_chunk_job_ids(const pqg_chunk_job* job);       /* device pointer, valid after emit */
PQG_API uint64_t pqg_chunk_job_total_weight(const pqg_chunk_job* job);     /* sum of the shard's weights */
PQG_API void pqg_chunk_job_free(pqg_ctx* ctx, pqg_chunk_job* job);

/* Page-level map (index_test): pages packed greedily by payload size in page-table order.
 * HOST outputs: page_chunk[n_pages], page_off[n_pages], chunk_first_page[cap]. */
PQG_API int pqg_page_chunk_index(pqg_ctx* ctx, const uint32_t* page_sizes, uint32_t n_pages,
                                 uint64_t chunk_size, uint32_t* page_chunk, uint32_t* page_off,
                                 uint32_t* chunk_first_page, uint32_t cap, uint32_t* n_chunks);

#ifdef __cplusplus
}
#endif
#endif /* PQG_H */
