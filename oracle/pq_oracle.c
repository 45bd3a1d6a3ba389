/*
 * pq_oracle.c -- TEST INFRASTRUCTURE: CPU restatement (plain C11) of the reference's
 * Parquet read path.  See pq_oracle.h for the pinning statement and the usage rule
 * (checker only; never on the product path).
 *
 * Each function cites the reference code it follows; paths are relative to
 * /root/reference/.  The code is written from the behaviour of those functions, in C
 * (the reference is C++17 with std::variant / std::vector), not copied from them.
 */
#define _GNU_SOURCE
#include "pq_oracle.h"

#include <fcntl.h>
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

static _Thread_local char g_err[512];

const char* orc_last_error(void) { return g_err; }

static void set_err(const char* msg) {
    if (g_err[0] == 0) snprintf(g_err, sizeof(g_err), "%s", msg);
}
static void clear_err(void) { g_err[0] = 0; }

/* ── byte cursor: ByteBuffer (include/common.hpp:110-173) ───────────────────────────── */

typedef struct {
    const uint8_t* data;
    size_t size;
    size_t pos;
    int failed;
} cursor;

/* ByteBuffer::check (include/common.hpp:162-168): throws with this exact text. */
static int cur_check(cursor* c, size_t n) {
    if (c->failed) return 0;
    if (c->pos + n > c->size) {
        char msg[200];
        snprintf(msg, sizeof(msg), "ByteBuffer: read beyond end (pos=%zu need=%zu size=%zu)",
                 c->pos, n, c->size);
        set_err(msg);
        c->failed = 1;
        return 0;
    }
    return 1;
}
static uint8_t cur_byte(cursor* c) {
    if (!cur_check(c, 1)) return 0;
    return c->data[c->pos++];
}
static const uint8_t* cur_bytes(cursor* c, size_t n) {
    if (!cur_check(c, n)) return NULL;
    const uint8_t* p = c->data + c->pos;
    c->pos += n;
    return p;
}
static uint32_t cur_u32(cursor* c) {
    const uint8_t* p = cur_bytes(c, 4);
    uint32_t v = 0;
    if (p) memcpy(&v, p, 4);
    return v;
}
/* ByteBuffer::read_varint (include/common.hpp:135-146) */
static uint64_t cur_varint(cursor* c) {
    uint64_t result = 0;
    int shift = 0;
    for (;;) {
        uint8_t b = cur_byte(c);
        if (c->failed) return 0;
        result |= (uint64_t)(b & 0x7F) << shift;
        if ((b & 0x80) == 0) break;
        shift += 7;
        if (shift > 63) { set_err("varint too long"); c->failed = 1; return 0; }
    }
    return result;
}
static int64_t cur_zigzag(cursor* c) {
    uint64_t v = cur_varint(c);
    return (int64_t)((v >> 1) ^ (~(v & 1) + 1));
}

/* ── Thrift compact protocol: ThriftReader (src/reader/thrift.cpp:6-119) ─────────────── */

enum { CT_STOP = 0, CT_TRUE = 1, CT_FALSE = 2, CT_I8 = 3, CT_I16 = 4, CT_I32 = 5, CT_I64 = 6,
       CT_DOUBLE = 7, CT_BINARY = 8, CT_LIST = 9, CT_SET = 10, CT_MAP = 11, CT_STRUCT = 12 };

typedef struct { int16_t id; uint8_t type; } field_hdr;

/* read_field_begin (src/reader/thrift.cpp:6-21); *last is the enclosing struct's
 * last_field_id_ (the reference keeps it on a std::stack, :57-65). */
static field_hdr th_field(cursor* c, int16_t* last) {
    field_hdr h = {0, CT_STOP};
    uint8_t b = cur_byte(c);
    if (c->failed || b == CT_STOP) return h;
    h.type = b & 0x0F;
    int16_t delta = (b >> 4) & 0x0F;
    if (delta != 0) h.id = (int16_t)(*last + delta);
    else h.id = (int16_t)cur_zigzag(c);
    *last = h.id;
    return h;
}
static int32_t th_i32(cursor* c) { return (int32_t)cur_zigzag(c); }
static int64_t th_i64(cursor* c) { return cur_zigzag(c); }
/* read_string (src/reader/thrift.cpp:36-40); returns malloc'd NUL-terminated copy */
static char* th_string(cursor* c) {
    uint32_t len = (uint32_t)cur_varint(c);
    const uint8_t* p = cur_bytes(c, len);
    char* s = (char*)malloc((size_t)len + 1);
    if (p) memcpy(s, p, len);
    s[p ? len : 0] = 0;
    return s;
}
typedef struct { uint8_t elem_type; int32_t count; } list_hdr;
static list_hdr th_list(cursor* c) {
    list_hdr l;
    uint8_t b = cur_byte(c);
    uint8_t nib = (b >> 4) & 0x0F;
    l.elem_type = b & 0x0F;
    l.count = (nib == 0x0F) ? (int32_t)cur_varint(c) : nib;
    return l;
}
/* skip (src/reader/thrift.cpp:67-119) */
static void th_skip(cursor* c, uint8_t type) {
    if (c->failed) return;
    switch (type) {
        case CT_TRUE: case CT_FALSE: break;
        case CT_I8: cur_byte(c); break;
        case CT_I16: case CT_I32: case CT_I64: cur_varint(c); break;
        case CT_DOUBLE: cur_bytes(c, 8); break;
        case CT_BINARY: { uint32_t len = (uint32_t)cur_varint(c); cur_bytes(c, len); break; }
        case CT_LIST: case CT_SET: {
            list_hdr l = th_list(c);
            for (int32_t i = 0; i < l.count && !c->failed; i++) th_skip(c, l.elem_type);
            break;
        }
        case CT_MAP: {
            int32_t count = (int32_t)cur_varint(c);
            if (count > 0) {
                uint8_t kv = cur_byte(c);
                for (int32_t i = 0; i < count && !c->failed; i++) {
                    th_skip(c, (kv >> 4) & 0x0F);
                    th_skip(c, kv & 0x0F);
                }
            }
            break;
        }
        case CT_STRUCT: {
            int16_t last = 0;
            for (;;) {
                field_hdr h = th_field(c, &last);
                if (c->failed || h.type == CT_STOP) break;
                th_skip(c, h.type);
            }
            break;
        }
        default: {
            char msg[64];
            snprintf(msg, sizeof(msg), "ThriftReader::skip: unknown type %d", (int)type);
            set_err(msg);
            c->failed = 1;
        }
    }
}

/* ── metadata (include/reader/metadata.hpp, src/reader/metadata.cpp) ─────────────────── */

typedef struct {
    int has_type; int32_t type;
    int has_rep; int32_t repetition;
    char* name;
    int has_children; int32_t num_children;
    int has_conv; int32_t converted;
} schema_elem;

typedef struct {
    int has_meta;
    int32_t type, codec;
    int64_t num_values, total_uncompressed, total_compressed, data_page_offset;
    int has_dict_off; int64_t dictionary_page_offset;
} chunk_meta;

typedef struct {
    chunk_meta* cols; int32_t ncols;
    int64_t total_byte_size, num_rows;
} row_group;

struct orc_file {
    const uint8_t* data; size_t size; int owned_map;
    int32_t version; int64_t num_rows;
    schema_elem* schema; int32_t nschema;
    row_group* rgs; int32_t nrgs;
    orc_colinfo* cols; int32_t ncols;
    orc_page_entry* pages; int64_t npages;
};

/* SchemaElement::deserialize (src/reader/metadata.cpp:5-22) */
static void parse_schema_elem(cursor* c, schema_elem* e) {
    memset(e, 0, sizeof(*e));
    int16_t last = 0;
    for (;;) {
        field_hdr h = th_field(c, &last);
        if (c->failed || h.type == CT_STOP) break;
        switch (h.id) {
            case 1: e->has_type = 1; e->type = th_i32(c); break;
            case 2: th_i32(c); break;
            case 3: e->has_rep = 1; e->repetition = th_i32(c); break;
            case 4: free(e->name); e->name = th_string(c); break;
            case 5: e->has_children = 1; e->num_children = th_i32(c); break;
            case 6: e->has_conv = 1; e->converted = th_i32(c); break;
            case 7: case 8: case 9: th_i32(c); break;
            default: th_skip(c, h.type); break;
        }
    }
    if (!e->name) e->name = strdup("");
}

/* ColumnMetaData::deserialize (src/reader/metadata.cpp:36-64) */
static void parse_column_meta(cursor* c, chunk_meta* m) {
    int16_t last = 0;
    m->has_meta = 1;
    m->type = 1; /* ParquetType::INT32 default, metadata.hpp:32 */
    for (;;) {
        field_hdr h = th_field(c, &last);
        if (c->failed || h.type == CT_STOP) break;
        switch (h.id) {
            case 1: m->type = th_i32(c); break;
            case 2: { list_hdr l = th_list(c); for (int32_t i = 0; i < l.count && !c->failed; i++) th_i32(c); break; }
            case 3: { list_hdr l = th_list(c); for (int32_t i = 0; i < l.count && !c->failed; i++) free(th_string(c)); break; }
            case 4: m->codec = th_i32(c); break;
            case 5: m->num_values = th_i64(c); break;
            case 6: m->total_uncompressed = th_i64(c); break;
            case 7: m->total_compressed = th_i64(c); break;
            case 9: m->data_page_offset = th_i64(c); break;
            case 10: th_i64(c); break;
            case 11: m->has_dict_off = 1; m->dictionary_page_offset = th_i64(c); break;
            default: th_skip(c, h.type); break;
        }
    }
}

/* ColumnChunk::deserialize (src/reader/metadata.cpp:68-86) */
static void parse_column_chunk(cursor* c, chunk_meta* m) {
    memset(m, 0, sizeof(*m));
    int16_t last = 0;
    for (;;) {
        field_hdr h = th_field(c, &last);
        if (c->failed || h.type == CT_STOP) break;
        switch (h.id) {
            case 1: free(th_string(c)); break;
            case 2: th_i64(c); break;
            case 3: parse_column_meta(c, m); break;
            default: th_skip(c, h.type); break;
        }
    }
}

/* RowGroup::deserialize (src/reader/metadata.cpp:159-180) */
static void parse_row_group(cursor* c, row_group* rg) {
    memset(rg, 0, sizeof(*rg));
    int16_t last = 0;
    for (;;) {
        field_hdr h = th_field(c, &last);
        if (c->failed || h.type == CT_STOP) break;
        switch (h.id) {
            case 1: {
                list_hdr l = th_list(c);
                for (int32_t i = 0; i < l.count && !c->failed; i++) {
                    rg->cols = (chunk_meta*)realloc(rg->cols, sizeof(chunk_meta) * (size_t)(rg->ncols + 1));
                    parse_column_chunk(c, &rg->cols[rg->ncols++]);
                }
                break;
            }
            case 2: rg->total_byte_size = th_i64(c); break;
            case 3: rg->num_rows = th_i64(c); break;
            default: th_skip(c, h.type); break;
        }
    }
}

/* FileMetaData::deserialize (src/reader/metadata.cpp:198-242) */
static void parse_file_meta(cursor* c, orc_file* f) {
    int16_t last = 0;
    for (;;) {
        field_hdr h = th_field(c, &last);
        if (c->failed || h.type == CT_STOP) break;
        switch (h.id) {
            case 1: f->version = th_i32(c); break;
            case 2: {
                list_hdr l = th_list(c);
                for (int32_t i = 0; i < l.count && !c->failed; i++) {
                    f->schema = (schema_elem*)realloc(f->schema, sizeof(schema_elem) * (size_t)(f->nschema + 1));
                    parse_schema_elem(c, &f->schema[f->nschema++]);
                }
                break;
            }
            case 3: f->num_rows = th_i64(c); break;
            case 4: {
                list_hdr l = th_list(c);
                for (int32_t i = 0; i < l.count && !c->failed; i++) {
                    f->rgs = (row_group*)realloc(f->rgs, sizeof(row_group) * (size_t)(f->nrgs + 1));
                    parse_row_group(c, &f->rgs[f->nrgs++]);
                }
                break;
            }
            case 5: {
                list_hdr l = th_list(c);
                for (int32_t i = 0; i < l.count && !c->failed; i++) th_skip(c, CT_STRUCT);
                break;
            }
            case 6: free(th_string(c)); break;
            default: th_skip(c, h.type); break;
        }
    }
}

/* PageHeader (include/reader/metadata.hpp:58-88; src/reader/metadata.cpp:90-155) */
typedef struct {
    int32_t type;                 /* default DATA_PAGE = 0 */
    int32_t uncompressed_page_size, compressed_page_size;
    int has_dph; int32_t dph_num_values, dph_encoding;
    int has_dict; int32_t dict_num_values;
    size_t header_size;
} page_header;

static void parse_data_page_header(cursor* c, page_header* ph) {
    int16_t last = 0;
    ph->has_dph = 1; ph->dph_num_values = 0; ph->dph_encoding = 0;
    for (;;) {
        field_hdr h = th_field(c, &last);
        if (c->failed || h.type == CT_STOP) break;
        switch (h.id) {
            case 1: ph->dph_num_values = th_i32(c); break;
            case 2: ph->dph_encoding = th_i32(c); break;
            case 3: case 4: th_i32(c); break;
            default: th_skip(c, h.type); break;
        }
    }
}
static void parse_dict_page_header(cursor* c, page_header* ph) {
    int16_t last = 0;
    ph->has_dict = 1; ph->dict_num_values = 0;
    for (;;) {
        field_hdr h = th_field(c, &last);
        if (c->failed || h.type == CT_STOP) break;
        switch (h.id) {
            case 1: ph->dict_num_values = th_i32(c); break;
            case 2: th_i32(c); break;
            case 3: break; /* bool lives in the field header */
            default: th_skip(c, h.type); break;
        }
    }
}

/* Reads the page header the way every reference loop does: a fixed 256-byte window at
 * `off` (HEADER_READ_SIZE, src/reader/column_reader.cpp:34-39), zero-filled past EOF
 * (read_range never checks the stream, src/reader/parquet_reader.cpp:173-178). */
static int read_page_header(const orc_file* f, size_t off, page_header* ph) {
    uint8_t win[256];
    memset(win, 0, sizeof(win));
    if (off < f->size) {
        size_t n = f->size - off < 256 ? f->size - off : 256;
        memcpy(win, f->data + off, n);
    }
    cursor c = {win, 256, 0, 0};
    memset(ph, 0, sizeof(*ph));
    int16_t last = 0;
    for (;;) {
        field_hdr h = th_field(&c, &last);
        if (c.failed || h.type == CT_STOP) break;
        switch (h.id) {
            case 1: ph->type = th_i32(&c); break;
            case 2: ph->uncompressed_page_size = th_i32(&c); break;
            case 3: ph->compressed_page_size = th_i32(&c); break;
            case 4: th_i32(&c); break;
            case 5: parse_data_page_header(&c, ph); break;
            case 7: parse_dict_page_header(&c, ph); break;
            default: th_skip(&c, h.type); break;
        }
    }
    ph->header_size = c.pos;
    return c.failed ? -1 : 0;
}

/* read_range (src/reader/parquet_reader.cpp:173-178): fresh zero-initialised buffer */
static uint8_t* read_range(const orc_file* f, size_t off, size_t len) {
    uint8_t* b = (uint8_t*)calloc(len + 16, 1); /* +16: slack for literal over-reads */
    if (off < f->size) {
        size_t n = f->size - off < len ? f->size - off : len;
        memcpy(b, f->data + off, n);
    }
    return b;
}

/* ── schema walk: build_column_info (src/reader/parquet_reader.cpp:484-557) ───────────── */

static int has_kids(const orc_file* f, int i) {
    return f->schema[i].has_children && f->schema[i].num_children > 0;
}
static int skip_subtree(const orc_file* f, int idx) {
    int children = f->schema[idx].has_children ? f->schema[idx].num_children : 0;
    idx++;
    for (int i = 0; i < children && idx < f->nschema; i++) {
        if (has_kids(f, idx)) idx = skip_subtree(f, idx);
        else idx++;
    }
    return idx;
}
static void build_columns(orc_file* f, int idx, int end, int16_t def, int16_t rep, int* col_index) {
    while (idx < end) {
        const schema_elem* e = &f->schema[idx];
        int16_t my_def = def, my_rep = rep;
        if (e->has_rep) {
            if (e->repetition == 1) my_def++;
            else if (e->repetition == 2) { my_def++; my_rep++; }
        }
        if (has_kids(f, idx)) {
            int remaining = e->num_children;
            idx++;
            int j = idx;
            while (remaining > 0 && j < end) {
                remaining--;
                if (has_kids(f, j)) j = skip_subtree(f, j);
                else j++;
            }
            build_columns(f, idx, j, my_def, my_rep, col_index);
            idx = j;
        } else {
            f->cols = (orc_colinfo*)realloc(f->cols, sizeof(orc_colinfo) * (size_t)(f->ncols + 1));
            orc_colinfo* ci = &f->cols[f->ncols++];
            memset(ci, 0, sizeof(*ci));
            snprintf(ci->name, sizeof(ci->name), "%s", e->name);
            ci->type = e->has_type ? e->type : 6; /* value_or(BYTE_ARRAY), :533 */
            ci->column_index = (*col_index)++;
            ci->max_def_level = my_def;
            ci->max_rep_level = my_rep;
            ci->repetition = e->has_rep ? e->repetition : -1;
            ci->converted = e->has_conv ? e->converted : -1;
            idx++;
        }
    }
}

static size_t chunk_start(const chunk_meta* m) {
    int64_t off = m->data_page_offset;
    if (m->has_dict_off && m->dictionary_page_offset < off) off = m->dictionary_page_offset;
    return (size_t)off;
}

/* build_page_index (src/reader/parquet_reader.cpp:559-605) */
static int build_page_index(orc_file* f) {
    int64_t cap = 0;
    for (int32_t rg = 0; rg < f->nrgs; rg++) {
        for (int32_t ci = 0; ci < f->rgs[rg].ncols; ci++) {
            const chunk_meta* m = &f->rgs[rg].cols[ci];
            if (!m->has_meta) continue;
            size_t cur = chunk_start(m);
            int64_t values_read = 0;
            while (values_read < m->num_values) {
                if (cur >= f->size) { set_err("page walk ran past end of file (reference would not terminate)"); return -1; }
                page_header ph;
                if (read_page_header(f, cur, &ph) != 0) return -1;
                cur += ph.header_size;
                if (ph.type == 0 || ph.type == 3) {
                    if (f->npages == cap) {
                        cap = cap ? cap * 2 : 1024;
                        f->pages = (orc_page_entry*)realloc(f->pages, sizeof(orc_page_entry) * (size_t)cap);
                    }
                    orc_page_entry* e = &f->pages[f->npages++];
                    e->data_offset = cur;
                    e->data_size = (uint64_t)(size_t)ph.compressed_page_size;
                    e->row_group_idx = (uint64_t)rg;
                    e->column_idx = (uint64_t)ci;
                    if (ph.type == 0 && ph.has_dph) values_read += ph.dph_num_values;
                }
                cur += (size_t)ph.compressed_page_size;
            }
        }
    }
    return 0;
}

/* ParquetReader::open (src/reader/parquet_reader.cpp:14-61) */
static orc_file* open_common(orc_file* f) {
    if (f->size < 12) { set_err("Error: file too small to be a Parquet file"); goto fail; }
    if (memcmp(f->data, "PAR1", 4) != 0) { set_err("Error: missing PAR1 magic at start"); goto fail; }
    if (memcmp(f->data + f->size - 4, "PAR1", 4) != 0) { set_err("Error: missing PAR1 magic at end"); goto fail; }
    uint32_t footer_len;
    memcpy(&footer_len, f->data + f->size - 8, 4);
    if ((size_t)footer_len + 8 > f->size) { set_err("Error: invalid footer length"); goto fail; }
    cursor c = {f->data + f->size - 8 - footer_len, footer_len, 0, 0};
    parse_file_meta(&c, f);
    if (c.failed) goto fail;
    if (f->nschema > 0) {
        int col_index = 0;
        build_columns(f, 1, f->nschema, 0, 0, &col_index);
    }
    if (build_page_index(f) != 0) goto fail;
    return f;
fail:
    orc_close(f);
    return NULL;
}

orc_file* orc_open(const char* path) {
    clear_err();
    int fd = open(path, O_RDONLY);
    if (fd < 0) { set_err("Error: cannot open file"); return NULL; }
    struct stat st;
    fstat(fd, &st);
    orc_file* f = (orc_file*)calloc(1, sizeof(orc_file));
    f->size = (size_t)st.st_size;
    if (f->size > 0) {
        void* p = mmap(NULL, f->size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (p == MAP_FAILED) { close(fd); free(f); set_err("mmap failed"); return NULL; }
        f->data = (const uint8_t*)p;
        f->owned_map = 1;
    }
    close(fd);
    return open_common(f);
}

orc_file* orc_open_mem(const uint8_t* data, size_t size) {
    clear_err();
    orc_file* f = (orc_file*)calloc(1, sizeof(orc_file));
    f->data = data;
    f->size = size;
    return open_common(f);
}

void orc_close(orc_file* f) {
    if (!f) return;
    for (int32_t i = 0; i < f->nschema; i++) free(f->schema[i].name);
    free(f->schema);
    for (int32_t i = 0; i < f->nrgs; i++) free(f->rgs[i].cols);
    free(f->rgs);
    free(f->cols);
    free(f->pages);
    if (f->owned_map && f->data) munmap((void*)f->data, f->size);
    free(f);
}

int64_t orc_num_rows(const orc_file* f) { return f->num_rows; }
int64_t orc_num_row_groups(const orc_file* f) { return f->nrgs; }
int64_t orc_num_columns(const orc_file* f) { return f->ncols; }
int64_t orc_num_pages(const orc_file* f) { return f->npages; }
int64_t orc_row_group_num_rows(const orc_file* f, int rg) { return f->rgs[rg].num_rows; }

int orc_column_info(const orc_file* f, int col, orc_colinfo* out) {
    clear_err();
    if (col < 0 || col >= f->ncols) {
        char msg[64];
        snprintf(msg, sizeof(msg), "Column index %d out of range", col);
        set_err(msg);
        return -1;
    }
    *out = f->cols[col];
    return 0;
}

/* find_column (src/reader/parquet_reader.cpp:93-97); the name map keeps the LAST column
 * of a given name (build_column_index overwrites, :477-482). */
int orc_find_column(const orc_file* f, const char* name) {
    for (int i = f->ncols - 1; i >= 0; i--)
        if (strcmp(f->cols[i].name, name) == 0) return i;
    return -1;
}

int64_t orc_page_index(const orc_file* f, orc_page_entry* out, int64_t cap) {
    for (int64_t i = 0; i < f->npages && i < cap; i++) out[i] = f->pages[i];
    return f->npages;
}

/* ── RleDecoder (include/reader/rle_decoder.hpp:6-108) ────────────────────────────────── */

typedef struct {
    const uint8_t* data;
    uint32_t size, avail, pos;
    uint8_t bw;
    uint32_t repeat_count, literal_count;
    uint64_t current_value;
    uint32_t literal_pos; /* byte offset of the literal run inside data */
    uint32_t literal_bit_offset;
    int has_literal_pos;
    int undefined; /* the reference would dereference a null/stale pointer here */
} rle_dec;

static void rle_init(rle_dec* d, const uint8_t* data, uint32_t size, uint32_t avail, uint8_t bw) {
    memset(d, 0, sizeof(*d));
    d->data = data; d->size = size; d->avail = avail < size ? size : avail; d->bw = bw;
}
/* read_varint32 (:76-86) */
static uint32_t rle_varint32(rle_dec* d) {
    uint32_t result = 0;
    int shift = 0;
    while (d->pos < d->size) {
        uint8_t b = d->data[d->pos++];
        if (shift < 32) result |= (uint32_t)(b & 0x7F) << shift;
        if ((b & 0x80) == 0) break;
        shift += 7;
    }
    return result;
}
/* next_counts (:37-53) */
static int rle_next_counts(rle_dec* d) {
    if (d->pos >= d->size) return 0;
    uint32_t indicator = rle_varint32(d);
    if (indicator & 1) {
        d->literal_count = (indicator >> 1) * 8;
        d->literal_pos = d->pos;
        d->has_literal_pos = 1;
        d->literal_bit_offset = 0;
    } else {
        d->repeat_count = indicator >> 1;
        /* read_fixed_width_value (:88-95): value bytes are NOT masked to bit_width */
        uint32_t need = ((uint32_t)d->bw + 7) / 8;
        uint64_t val = 0;
        for (uint32_t i = 0; i < need && d->pos < d->size; i++) {
            if (i < 8) val |= (uint64_t)d->data[d->pos] << (i * 8);
            d->pos++;
        }
        d->current_value = val;
    }
    return 1;
}
/* read_literal_value (:55-74): LSB-first, bit by bit, no bound (bounded here by avail) */
static uint64_t rle_literal(rle_dec* d) {
    if (d->bw == 0) return 0;
    if (!d->has_literal_pos) { d->undefined = 1; return 0; }
    uint64_t val = 0;
    for (uint8_t i = 0; i < d->bw; i++) {
        uint32_t byte_idx = d->literal_pos + d->literal_bit_offset / 8;
        uint32_t bit_idx = d->literal_bit_offset % 8;
        uint8_t byte = byte_idx < d->avail ? d->data[byte_idx] : 0;
        if ((byte & (1u << bit_idx)) && i < 64) val |= (uint64_t)1 << i;
        d->literal_bit_offset++;
    }
    if (d->literal_count == 1) d->pos = d->literal_pos + (d->literal_bit_offset + 7) / 8;
    return val;
}
/* get_batch (:17-34) */
static uint64_t rle_next(rle_dec* d, int* exhausted) {
    if (d->repeat_count == 0 && d->literal_count == 0) {
        if (!rle_next_counts(d)) { *exhausted = 1; return 0; }
    }
    if (d->repeat_count > 0) {
        d->repeat_count--;
        return d->current_value;
    }
    uint64_t v = rle_literal(d);
    d->literal_count--; /* wraps exactly like the reference's uint32_t */
    return v;
}

void orc_rle_decode_i32(const uint8_t* data, uint32_t size, uint32_t avail, int bit_width,
                        int32_t* out, uint32_t count) {
    rle_dec d;
    rle_init(&d, data, size, avail, (uint8_t)bit_width);
    int exhausted = 0;
    for (uint32_t i = 0; i < count; i++) {
        uint64_t v = exhausted ? 0 : rle_next(&d, &exhausted);
        out[i] = (int32_t)(uint32_t)v;
    }
}
void orc_rle_decode_i16(const uint8_t* data, uint32_t size, uint32_t avail, int bit_width,
                        int16_t* out, uint32_t count) {
    rle_dec d;
    rle_init(&d, data, size, avail, (uint8_t)bit_width);
    int exhausted = 0;
    for (uint32_t i = 0; i < count; i++) {
        uint64_t v = exhausted ? 0 : rle_next(&d, &exhausted);
        out[i] = (int16_t)(uint16_t)v;
    }
}

/* bit_width (src/reader/column_reader.cpp:270-276) */
static uint8_t level_bit_width(int16_t max_level) {
    if (max_level <= 0) return 0;
    uint8_t bw = 0;
    int16_t v = max_level;
    while (v > 0) { bw++; v >>= 1; }
    return bw;
}

/* ── growable vector<Value> in dump form ──────────────────────────────────────────────── */

typedef struct {
    int64_t n, cap;
    uint8_t* is_null; uint8_t* vidx; uint64_t* fixed; uint64_t* str_off;
    uint8_t* chars; int64_t chars_len, chars_cap;
} vvec;

static void vv_reserve(vvec* v, int64_t extra) {
    if (v->n + extra + 1 <= v->cap) return;
    int64_t cap = v->cap ? v->cap : 1024;
    while (cap < v->n + extra + 1) cap *= 2;
    v->is_null = (uint8_t*)realloc(v->is_null, (size_t)cap);
    v->vidx = (uint8_t*)realloc(v->vidx, (size_t)cap);
    v->fixed = (uint64_t*)realloc(v->fixed, (size_t)cap * 8);
    v->str_off = (uint64_t*)realloc(v->str_off, (size_t)cap * 8);
    v->cap = cap;
}
static void vv_push_fixed(vvec* v, int is_null, int vidx, uint64_t bits) {
    vv_reserve(v, 1);
    v->is_null[v->n] = (uint8_t)is_null;
    v->vidx[v->n] = (uint8_t)vidx;
    v->fixed[v->n] = bits;
    v->str_off[v->n] = (uint64_t)v->chars_len;
    v->n++;
}
static void vv_push_str(vvec* v, const uint8_t* p, size_t len) {
    vv_reserve(v, 1);
    if (v->chars_len + (int64_t)len + 1 > v->chars_cap) {
        int64_t cap = v->chars_cap ? v->chars_cap : 4096;
        while (cap < v->chars_len + (int64_t)len + 1) cap *= 2;
        v->chars = (uint8_t*)realloc(v->chars, (size_t)cap);
        v->chars_cap = cap;
    }
    v->is_null[v->n] = 0;
    v->vidx[v->n] = 5;
    v->fixed[v->n] = 0;
    v->str_off[v->n] = (uint64_t)v->chars_len;
    if (len) memcpy(v->chars + v->chars_len, p, len);
    v->chars_len += (int64_t)len;
    v->n++;
}
/* Value::null() (include/common.hpp:181): is_null with variant alternative 0 */
static void vv_push_null(vvec* v) { vv_push_fixed(v, 1, 0, 0); }
static void vv_push_copy(vvec* v, const vvec* src, int64_t i) {
    if (src->vidx[i] == 5 && !src->is_null[i])
        vv_push_str(v, src->chars + src->str_off[i], (size_t)(src->str_off[i + 1] - src->str_off[i]));
    else
        vv_push_fixed(v, src->is_null[i], src->vidx[i], src->fixed[i]);
}
static void vv_seal(vvec* v) {
    vv_reserve(v, 0);
    v->str_off[v->n] = (uint64_t)v->chars_len;
}
static void vv_free(vvec* v) {
    free(v->is_null); free(v->vidx); free(v->fixed); free(v->str_off); free(v->chars);
    memset(v, 0, sizeof(*v));
}
static void vv_to_dump(vvec* v, valdump* out) {
    vv_seal(v);
    if (!v->chars) v->chars = (uint8_t*)malloc(1);
    out->n = v->n; out->is_null = v->is_null; out->vidx = v->vidx; out->fixed = v->fixed;
    out->str_off = v->str_off; out->chars = v->chars; out->chars_len = v->chars_len;
    memset(v, 0, sizeof(*v));
}
void orc_valdump_free(valdump* d) {
    free(d->is_null); free(d->vidx); free(d->fixed); free(d->str_off); free(d->chars);
    memset(d, 0, sizeof(*d));
}
void orc_pagedump_free(pagedump* d) {
    free(d->page_num); free(d->page_type); free(d->num_values); free(d->first_value);
    orc_valdump_free(&d->values);
    memset(d, 0, sizeof(*d));
}

/* ── value decode: ColumnReader (src/reader/column_reader.cpp) ────────────────────────── */

/* read_plain_value (:227-268) */
static int read_plain_value(cursor* c, int32_t type, vvec* out) {
    switch (type) {
        case 0: { uint8_t b = cur_byte(c); if (c->failed) return -1; vv_push_fixed(out, 0, 0, b != 0); return 0; }
        case 1: { uint32_t v = cur_u32(c); if (c->failed) return -1; vv_push_fixed(out, 0, 1, v); return 0; }
        case 2: { const uint8_t* p = cur_bytes(c, 8); if (!p) return -1; uint64_t v; memcpy(&v, p, 8); vv_push_fixed(out, 0, 2, v); return 0; }
        case 4: { uint32_t v = cur_u32(c); if (c->failed) return -1; vv_push_fixed(out, 0, 3, v); return 0; }
        case 5: { const uint8_t* p = cur_bytes(c, 8); if (!p) return -1; uint64_t v; memcpy(&v, p, 8); vv_push_fixed(out, 0, 4, v); return 0; }
        case 6: {
            uint32_t len = cur_u32(c);
            if (c->failed) return -1;
            const uint8_t* p = cur_bytes(c, len);
            if (!p && len) return -1;
            if (c->failed) return -1;
            vv_push_str(out, p, len);
            return 0;
        }
        case 7: set_err("FIXED_LEN_BYTE_ARRAY not supported without type_length"); return -1;
        case 3: { /* INT96 becomes a *string* "INT96(high:low)" (:257-264) */
            const uint8_t* p = cur_bytes(c, 12);
            if (!p) return -1;
            int64_t low; int32_t high;
            memcpy(&low, p, 8); memcpy(&high, p + 8, 4);
            char s[64];
            int n = snprintf(s, sizeof(s), "INT96(%" PRId32 ":%" PRId64 ")", high, low);
            vv_push_str(out, (const uint8_t*)s, (size_t)n);
            return 0;
        }
        default: {
            char msg[64];
            snprintf(msg, sizeof(msg), "Unsupported type: %d", (int)type);
            set_err(msg);
            return -1;
        }
    }
}

/* read_dictionary_page (:128-138) */
static int read_dictionary_page(const uint8_t* data, int32_t size, int32_t num_values,
                                int32_t type, vvec* dict) {
    vv_free(dict);
    cursor c = {data, (size_t)size, 0, 0};
    for (int32_t i = 0; i < num_values; i++)
        if (read_plain_value(&c, type, dict) != 0) return -1;
    vv_seal(dict);
    return 0;
}

/* read_data_page (:140-225).  Appends header.num_values slots to out. */
static int read_data_page(const uint8_t* data, int32_t size, int32_t num_values, int32_t encoding,
                          int32_t type, int16_t max_def, int16_t max_rep,
                          const vvec* dictionary, vvec* out) {
    cursor c = {data, (size_t)size, 0, 0};
    if (num_values < 0) { set_err("negative num_values"); return -1; }
    int16_t* def = (int16_t*)malloc(sizeof(int16_t) * (size_t)(num_values + 1));
    for (int32_t i = 0; i < num_values; i++) def[i] = max_def;
    int rc = -1;
    if (max_def > 0) {
        uint32_t def_len = cur_u32(&c);
        if (c.failed) goto done;
        orc_rle_decode_i16(data + c.pos, def_len, (uint32_t)(c.size - c.pos), level_bit_width(max_def),
                           def, (uint32_t)num_values);
        if (!cur_bytes(&c, def_len) && def_len) goto done;
        if (c.failed) goto done;
    }
    if (max_rep > 0) { /* decoded by the reference, then unused (:157-164) */
        uint32_t rep_len = cur_u32(&c);
        if (c.failed) goto done;
        if (!cur_bytes(&c, rep_len) && rep_len) goto done;
        if (c.failed) goto done;
    }
    int32_t num_non_null = 0;
    for (int32_t i = 0; i < num_values; i++) if (def[i] == max_def) num_non_null++;

    int use_dict = (encoding == 2 || encoding == 8);
    if (use_dict && dictionary) {
        uint8_t bw = cur_byte(&c);
        if (c.failed) goto done;
        if (bw > 64) { set_err("dictionary index bit width > 64 (reference behaviour undefined)"); goto done; }
        int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * (size_t)(num_non_null + 1));
        uint32_t rem = (uint32_t)(c.size - c.pos);
        orc_rle_decode_i32(data + c.pos, rem, rem, bw, idx, (uint32_t)num_non_null);
        int32_t ip = 0;
        for (int32_t i = 0; i < num_values; i++) {
            if (def[i] < max_def) { vv_push_null(out); continue; }
            int32_t k = idx[ip++];
            if (k >= 0 && (int64_t)k < dictionary->n) vv_push_copy(out, dictionary, k);
            else vv_push_null(out); /* out-of-range index -> null (:190-194) */
        }
        free(idx);
    } else if (type == 0) { /* BOOLEAN PLAIN: LSB-first bits, consumed by non-null slots (:197-212) */
        int32_t bit_idx = 0;
        uint8_t cur = 0;
        for (int32_t i = 0; i < num_values; i++) {
            if (def[i] < max_def) { vv_push_null(out); continue; }
            if (bit_idx % 8 == 0) { cur = cur_byte(&c); if (c.failed) goto done; }
            vv_push_fixed(out, 0, 0, (cur >> (bit_idx % 8)) & 1);
            bit_idx++;
        }
    } else {
        for (int32_t i = 0; i < num_values; i++) {
            if (def[i] < max_def) { vv_push_null(out); continue; }
            if (read_plain_value(&c, type, out) != 0) goto done;
        }
    }
    rc = 0;
done:
    free(def);
    return rc;
}

typedef struct {
    pagedump* pages; /* NULL for read_all */
    int64_t pcap;
} page_sink;

/* ColumnReader ctor checks (:3-16) + read_all (:18-71) / read_pages (:73-126) */
static int read_chunk(orc_file* f, const chunk_meta* m, int32_t type, int16_t max_def,
                      int16_t max_rep, vvec* out, page_sink* sink) {
    if (!m->has_meta) { set_err("ColumnChunk has no metadata"); return -1; }
    if (m->codec != 0) { set_err("Only uncompressed parquet files are supported"); return -1; }
    size_t cur = chunk_start(m);
    int64_t values_read = 0;
    int has_dict = 0;
    vvec dict;
    memset(&dict, 0, sizeof(dict));
    int page_num = 0;
    int rc = -1;
    while (values_read < m->num_values) {
        if (cur >= f->size) { set_err("page walk ran past end of file (reference would not terminate)"); goto done; }
        page_header ph;
        if (read_page_header(f, cur, &ph) != 0) goto done;
        cur += ph.header_size;
        int32_t page_size = ph.compressed_page_size;
        uint8_t* page = read_range(f, cur, (size_t)page_size);
        int32_t rec_type = -1, rec_nv = 0;
        int64_t before = out->n;
        if (ph.type == 2) {
            if (!ph.has_dict) { free(page); set_err("bad_optional_access"); goto done; }
            if (read_dictionary_page(page, page_size, ph.dict_num_values, type, &dict) != 0) { free(page); goto done; }
            has_dict = 1;
            rec_type = 2; rec_nv = ph.dict_num_values;
        } else if (ph.type == 0) {
            if (!ph.has_dph) { free(page); set_err("bad_optional_access"); goto done; }
            if (read_data_page(page, page_size, ph.dph_num_values, ph.dph_encoding, type, max_def,
                               max_rep, has_dict ? &dict : NULL, out) != 0) { free(page); goto done; }
            values_read += ph.dph_num_values;
            rec_type = 0; rec_nv = ph.dph_num_values;
        }
        free(page);
        cur += (size_t)page_size;
        if (sink) {
            pagedump* pd = sink->pages;
            if (rec_type >= 0) {
                if (pd->n_pages + 2 >= sink->pcap) {
                    sink->pcap = sink->pcap ? sink->pcap * 2 : 256;
                    pd->page_num = (int32_t*)realloc(pd->page_num, 4 * (size_t)sink->pcap);
                    pd->page_type = (int32_t*)realloc(pd->page_type, 4 * (size_t)sink->pcap);
                    pd->num_values = (int32_t*)realloc(pd->num_values, 4 * (size_t)sink->pcap);
                    pd->first_value = (int64_t*)realloc(pd->first_value, 8 * (size_t)sink->pcap);
                }
                pd->page_num[pd->n_pages] = page_num;
                pd->page_type[pd->n_pages] = rec_type;
                pd->num_values[pd->n_pages] = rec_nv;
                pd->first_value[pd->n_pages] = before;
                pd->n_pages++;
                pd->first_value[pd->n_pages] = out->n;
            }
            page_num++; /* counts dictionary and unknown pages too (:104,115,122) */
        }
    }
    rc = 0;
done:
    vv_free(&dict);
    return rc;
}

static int check_rg_col(orc_file* f, int rg, int col) {
    if (rg < 0 || rg >= f->nrgs) { set_err("Invalid row group index"); return -1; }
    if (col < 0 || col >= f->ncols) { set_err("Invalid column index"); return -1; }
    if (f->cols[col].column_index >= f->rgs[rg].ncols) { set_err("column chunk missing"); return -1; }
    return 0;
}

int orc_read_column_by_idx(orc_file* f, int rg, int col, valdump* out) {
    clear_err();
    memset(out, 0, sizeof(*out));
    if (check_rg_col(f, rg, col) != 0) return -1;
    const orc_colinfo* ci = &f->cols[col];
    vvec v;
    memset(&v, 0, sizeof(v));
    int rc = read_chunk(f, &f->rgs[rg].cols[ci->column_index], ci->type,
                        (int16_t)ci->max_def_level, (int16_t)ci->max_rep_level, &v, NULL);
    if (rc != 0) { vv_free(&v); return -1; }
    vv_to_dump(&v, out);
    return 0;
}

int orc_read_column(orc_file* f, const char* name, valdump* out) {
    clear_err();
    memset(out, 0, sizeof(*out));
    int col = orc_find_column(f, name);
    if (col < 0) {
        char msg[300];
        snprintf(msg, sizeof(msg), "Column not found: %s", name);
        set_err(msg);
        return -1;
    }
    const orc_colinfo* ci = &f->cols[col];
    vvec v;
    memset(&v, 0, sizeof(v));
    for (int rg = 0; rg < f->nrgs; rg++) {
        if (check_rg_col(f, rg, col) != 0 ||
            read_chunk(f, &f->rgs[rg].cols[ci->column_index], ci->type, (int16_t)ci->max_def_level,
                       (int16_t)ci->max_rep_level, &v, NULL) != 0) {
            vv_free(&v);
            return -1;
        }
    }
    vv_to_dump(&v, out);
    return 0;
}

int orc_read_pages(orc_file* f, int rg, int col, pagedump* out) {
    clear_err();
    memset(out, 0, sizeof(*out));
    if (check_rg_col(f, rg, col) != 0) return -1;
    const orc_colinfo* ci = &f->cols[col];
    vvec v;
    memset(&v, 0, sizeof(v));
    page_sink sink = {out, 0};
    sink.pcap = 256;
    out->page_num = (int32_t*)calloc(256, 4);
    out->page_type = (int32_t*)calloc(256, 4);
    out->num_values = (int32_t*)calloc(256, 4);
    out->first_value = (int64_t*)calloc(256, 8);
    int rc = read_chunk(f, &f->rgs[rg].cols[ci->column_index], ci->type,
                        (int16_t)ci->max_def_level, (int16_t)ci->max_rep_level, &v, &sink);
    if (rc != 0) { vv_free(&v); orc_pagedump_free(out); return -1; }
    vv_to_dump(&v, &out->values);
    return 0;
}

/* ── raw page API (src/reader/parquet_reader.cpp:182-238) ─────────────────────────────── */

int64_t orc_read_page_data(orc_file* f, int64_t id, uint8_t* buf, int64_t cap) {
    clear_err();
    if (id < 0 || id >= f->npages) {
        char msg[96];
        snprintf(msg, sizeof(msg), "Global page ID %" PRId64 " out of range", id);
        set_err(msg);
        return -1;
    }
    const orc_page_entry* e = &f->pages[id];
    if ((int64_t)e->data_size > cap) { set_err("buffer too small"); return -2; }
    uint8_t* p = read_range(f, e->data_offset, e->data_size);
    memcpy(buf, p, e->data_size);
    free(p);
    return (int64_t)e->data_size;
}

int64_t orc_read_pages_chunk(orc_file* f, int64_t s, int64_t e, int64_t max_bytes,
                             uint8_t* buf, int64_t cap) {
    clear_err();
    char msg[96];
    if (s < 0 || s >= f->npages) { snprintf(msg, sizeof(msg), "Start page ID %" PRId64 " out of range", s); set_err(msg); return -1; }
    if (e < 0 || e >= f->npages) { snprintf(msg, sizeof(msg), "End page ID %" PRId64 " out of range", e); set_err(msg); return -1; }
    if (s > e) { set_err("Start page ID must be <= end page ID"); return -1; }
    int64_t n = 0;
    for (int64_t i = s; i <= e; i++) {
        int64_t remaining = max_bytes - n;
        if (remaining == 0) break;
        int64_t to_read = (int64_t)f->pages[i].data_size < remaining ? (int64_t)f->pages[i].data_size : remaining;
        if (n + to_read > cap) { set_err("buffer too small"); return -2; }
        uint8_t* p = read_range(f, f->pages[i].data_offset, (size_t)to_read);
        memcpy(buf + n, p, (size_t)to_read);
        free(p);
        n += to_read;
    }
    return n;
}

/* ── StringColumnIterator (src/reader/parquet_reader.cpp:282-465), drained ────────────── */

typedef struct {
    uint64_t* pos; uint64_t* off; uint8_t* chars;
    int64_t n, cap, clen, ccap;
} strsink;

static void ss_push(strsink* s, uint64_t pos, const uint8_t* p, size_t len) {
    if (s->n + 2 > s->cap) {
        s->cap = s->cap ? s->cap * 2 : 1024;
        s->pos = (uint64_t*)realloc(s->pos, 8 * (size_t)s->cap);
        s->off = (uint64_t*)realloc(s->off, 8 * (size_t)s->cap);
    }
    if (s->clen + (int64_t)len + 1 > s->ccap) {
        int64_t cap = s->ccap ? s->ccap : 4096;
        while (cap < s->clen + (int64_t)len + 1) cap *= 2;
        s->chars = (uint8_t*)realloc(s->chars, (size_t)cap);
        s->ccap = cap;
    }
    s->pos[s->n] = pos;
    s->off[s->n] = (uint64_t)s->clen;
    if (len) memcpy(s->chars + s->clen, p, len);
    s->clen += (int64_t)len;
    s->n++;
    s->off[s->n] = (uint64_t)s->clen;
}

static int string_iterate(orc_file* f, int col, strsink* sink) {
    const orc_colinfo* ci = &f->cols[col];
    int16_t max_def = (int16_t)ci->max_def_level, max_rep = (int16_t)ci->max_rep_level;
    size_t row_group_base = 0;
    for (int rg = 0; rg < f->nrgs; rg++) {
        if (ci->column_index >= f->rgs[rg].ncols || !f->rgs[rg].cols[ci->column_index].has_meta) {
            set_err("bad_optional_access");
            return -1;
        }
        const chunk_meta* m = &f->rgs[rg].cols[ci->column_index];
        size_t cur = chunk_start(m);
        int64_t values_read = 0;
        int has_dict = 0;
        vvec dict;
        memset(&dict, 0, sizeof(dict));
        while (values_read < m->num_values) {
            if (cur >= f->size) { vv_free(&dict); set_err("page walk ran past end of file (reference would not terminate)"); return -1; }
            page_header ph;
            if (read_page_header(f, cur, &ph) != 0) { vv_free(&dict); return -1; }
            cur += ph.header_size;
            int32_t page_size = ph.compressed_page_size;
            uint8_t* page = read_range(f, cur, (size_t)page_size);
            if (ph.type == 2) { /* (:380-393) always BYTE_ARRAY parse */
                if (!ph.has_dict || read_dictionary_page(page, page_size, ph.dict_num_values, 6, &dict) != 0) {
                    if (!ph.has_dict) set_err("bad_optional_access");
                    free(page); vv_free(&dict); return -1;
                }
                has_dict = 1;
            } else if (ph.type == 0) { /* (:395-458) */
                if (!ph.has_dph) { free(page); vv_free(&dict); set_err("bad_optional_access"); return -1; }
                int32_t nv = ph.dph_num_values;
                size_t base_pos = row_group_base + (size_t)values_read;
                cursor c = {page, (size_t)page_size, 0, 0};
                int16_t* def = (int16_t*)malloc(sizeof(int16_t) * (size_t)(nv + 1));
                for (int32_t i = 0; i < nv; i++) def[i] = max_def;
                int bad = 0;
                if (max_def > 0) {
                    uint32_t def_len = cur_u32(&c);
                    if (!c.failed) {
                        orc_rle_decode_i16(page + c.pos, def_len, (uint32_t)(c.size - c.pos),
                                           level_bit_width(max_def), def, (uint32_t)nv);
                        cur_bytes(&c, def_len);
                    }
                }
                if (!c.failed && max_rep > 0) { uint32_t rep_len = cur_u32(&c); if (!c.failed) cur_bytes(&c, rep_len); }
                int32_t nn = 0;
                for (int32_t i = 0; i < nv; i++) if (def[i] == max_def) nn++;
                int use_dict = (ph.dph_encoding == 2 || ph.dph_encoding == 8);
                if (!c.failed && use_dict && has_dict) {
                    uint8_t bw = cur_byte(&c);
                    if (!c.failed) {
                        int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nn + 1));
                        uint32_t rem = (uint32_t)(c.size - c.pos);
                        orc_rle_decode_i32(page + c.pos, rem, rem, bw, idx, (uint32_t)nn);
                        int32_t ip = 0;
                        for (int32_t i = 0; i < nv; i++) {
                            if (def[i] != max_def) continue;
                            int32_t k = idx[ip++];
                            /* out-of-range index is DROPPED here (:436-439) */
                            if (k >= 0 && (int64_t)k < dict.n)
                                ss_push(sink, base_pos + (size_t)i, dict.chars + dict.str_off[k],
                                        (size_t)(dict.str_off[k + 1] - dict.str_off[k]));
                        }
                        free(idx);
                    }
                } else if (!c.failed) {
                    for (int32_t i = 0; i < nv && !c.failed; i++) {
                        if (def[i] != max_def) continue;
                        uint32_t len = cur_u32(&c);
                        if (c.failed) break;
                        const uint8_t* p = cur_bytes(&c, len);
                        if (c.failed) break;
                        ss_push(sink, base_pos + (size_t)i, p, len);
                    }
                }
                bad = c.failed;
                free(def);
                if (bad) { free(page); vv_free(&dict); return -1; }
                values_read += nv;
            }
            free(page);
            cur += (size_t)page_size;
        }
        vv_free(&dict);
        row_group_base += (size_t)f->rgs[rg].num_rows;
    }
    return 0;
}

static int resolve_string_column(orc_file* f, const char* name) {
    int col = orc_find_column(f, name);
    char msg[400];
    if (col < 0) { snprintf(msg, sizeof(msg), "Column not found: %s", name); set_err(msg); return -1; }
    if (f->cols[col].type != 6) {
        static const char* names[] = {"BOOLEAN", "INT32", "INT64", "INT96", "FLOAT", "DOUBLE",
                                      "BYTE_ARRAY", "FIXED_LEN_BYTE_ARRAY"};
        int t = f->cols[col].type;
        snprintf(msg, sizeof(msg), "Column '%s' is not BYTE_ARRAY (type: %s)", name,
                 (t >= 0 && t < 8) ? names[t] : "UNKNOWN");
        set_err(msg);
        return -1;
    }
    return col;
}

int orc_string_iterator_dump(orc_file* f, const char* name, orc_strdump* out) {
    clear_err();
    memset(out, 0, sizeof(*out));
    int col = resolve_string_column(f, name);
    if (col < 0) return -1;
    strsink s;
    memset(&s, 0, sizeof(s));
    if (string_iterate(f, col, &s) != 0) { free(s.pos); free(s.off); free(s.chars); return -1; }
    if (!s.off) { s.off = (uint64_t*)calloc(1, 8); }
    if (!s.pos) { s.pos = (uint64_t*)calloc(1, 8); }
    if (!s.chars) { s.chars = (uint8_t*)malloc(1); }
    out->n = s.n; out->pos = s.pos; out->off = s.off; out->chars = s.chars;
    return 0;
}
void orc_strdump_free(orc_strdump* d) {
    free(d->pos); free(d->off); free(d->chars);
    memset(d, 0, sizeof(*d));
}

/* number of decimal digits of std::to_string(size_t) */
static size_t dec_digits(uint64_t v) {
    size_t d = 1;
    while (v >= 10) { v /= 10; d++; }
    return d;
}

/* chunk-index prototype (src/main.cpp:21-32) */
int64_t orc_chunk_index(orc_file* f, const char* name, uint64_t chunk_size,
                        uint64_t* tuple_to_chunk, int64_t num_rows) {
    clear_err();
    int col = resolve_string_column(f, name);
    if (col < 0) return -1;
    strsink s;
    memset(&s, 0, sizeof(s));
    if (string_iterate(f, col, &s) != 0) { free(s.pos); free(s.off); free(s.chars); return -1; }
    for (int64_t i = 0; i < num_rows; i++) tuple_to_chunk[i] = 0;
    uint64_t chunk_bytes = 0, chunk_id = 0;
    for (int64_t i = 0; i < s.n; i++) {
        uint64_t len = s.off[i + 1] - s.off[i];
        if (chunk_bytes >= chunk_size) { chunk_bytes = 0; chunk_id++; }
        chunk_bytes += dec_digits(len) + len; /* to_string(len) + bytes (:30) */
        if ((int64_t)s.pos[i] < num_rows) tuple_to_chunk[s.pos[i]] = chunk_id;
    }
    free(s.pos); free(s.off); free(s.chars);
    return (int64_t)(chunk_id + 1);
}

/* page-level chunk index, frozen spec SURVEY.md section 8 a-20 */
int64_t orc_page_chunk_index(const orc_file* f, int col, uint64_t chunk_size,
                             uint32_t* page_chunk, uint32_t* page_off,
                             uint32_t* chunk_first_page, int64_t cap,
                             int64_t* first_global_page, int64_t* n_col_pages) {
    clear_err();
    if (col < 0 || col >= f->ncols) { set_err("Invalid column index"); return -1; }
    uint64_t want = (uint64_t)f->cols[col].column_index;
    int64_t local = 0, chunk_id = 0, first = -1;
    uint64_t bytes = 0;
    for (int64_t g = 0; g < f->npages; g++) {
        if (f->pages[g].column_idx != want) continue;
        if (first < 0) first = g;
        if (local == 0) { if (cap > 0) chunk_first_page[0] = 0; }
        else if (bytes >= chunk_size) {
            bytes = 0;
            chunk_id++;
            if (chunk_id < cap) chunk_first_page[chunk_id] = (uint32_t)local;
        }
        page_chunk[local] = (uint32_t)chunk_id;
        page_off[local] = (uint32_t)bytes;
        bytes += f->pages[g].data_size;
        local++;
    }
    if (first_global_page) *first_global_page = first;
    if (n_col_pages) *n_col_pages = local;
    return local ? chunk_id + 1 : 0;
}
