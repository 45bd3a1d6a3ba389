// pqg_reader_c.cpp -- extern "C" surface of the host reader (include/pqg_reader.h).
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <string>

#include "pq_reader.hpp"
#include "pq_scan.hpp"
#include "pqg_reader.h"

using namespace pqg;

struct pqr_reader { ParquetReader r; };

static thread_local std::string g_err;

template <typename F>
static auto guarded(F&& f, decltype(f()) on_error) -> decltype(f()) {
    try { return f(); }
    catch (const std::exception& e) { g_err = e.what(); return on_error; }
    catch (...) { g_err = "unknown error"; return on_error; }
}

static void dump_values(const std::vector<Value>& v, pqr_valdump* out) {
    int64_t n = static_cast<int64_t>(v.size());
    out->n = n;
    out->is_null = static_cast<uint8_t*>(std::calloc(n + 1, 1));
    out->vidx = static_cast<uint8_t*>(std::calloc(n + 1, 1));
    out->fixed = static_cast<uint64_t*>(std::calloc(n + 1, 8));
    out->str_off = static_cast<uint64_t*>(std::calloc(n + 1, 8));
    uint64_t total = 0;
    for (const Value& x : v) if (x.data.index() == 5) total += std::get<std::string>(x.data).size();
    out->chars = static_cast<uint8_t*>(std::malloc(total + 1));
    out->chars_len = static_cast<int64_t>(total);
    uint64_t pos = 0;
    for (int64_t i = 0; i < n; i++) {
        const Value& x = v[i];
        out->is_null[i] = x.is_null;
        out->vidx[i] = static_cast<uint8_t>(x.data.index());
        out->str_off[i] = pos;
        uint64_t bits = 0;
        switch (x.data.index()) {
            case 0: bits = std::get<bool>(x.data); break;
            case 1: { uint32_t u; std::memcpy(&u, &std::get<int32_t>(x.data), 4); bits = u; break; }
            case 2: std::memcpy(&bits, &std::get<int64_t>(x.data), 8); break;
            case 3: { uint32_t u; std::memcpy(&u, &std::get<float>(x.data), 4); bits = u; break; }
            case 4: std::memcpy(&bits, &std::get<double>(x.data), 8); break;
            case 5: {
                const std::string& s = std::get<std::string>(x.data);
                std::memcpy(out->chars + pos, s.data(), s.size());
                pos += s.size();
                break;
            }
        }
        out->fixed[i] = bits;
    }
    out->str_off[n] = pos;
}

extern "C" {

const char* pqr_last_error(void) { return g_err.c_str(); }

pqr_reader* pqr_open(const char* path, int device) {
    return guarded([&]() -> pqr_reader* {
        auto* h = new pqr_reader();
        h->r.set_device(device);
        if (!h->r.open(path)) { g_err = h->r.open_error(); delete h; return nullptr; }
        return h;
    }, nullptr);
}

pqr_reader* pqr_open_memory(const uint8_t* data, uint64_t size, int device) {
    return guarded([&]() -> pqr_reader* {
        auto* h = new pqr_reader();
        h->r.set_device(device);
        if (!h->r.open_memory(data, size)) { g_err = h->r.open_error(); delete h; return nullptr; }
        return h;
    }, nullptr);
}

void pqr_close(pqr_reader* r) { delete r; }
void pqr_set_extensions(pqr_reader* r, int on) { if (r) r->r.set_extensions(on != 0); }

int64_t pqr_num_rows(const pqr_reader* r) { return r->r.num_rows(); }
int64_t pqr_num_row_groups(const pqr_reader* r) { return static_cast<int64_t>(r->r.num_row_groups()); }
int64_t pqr_num_columns(const pqr_reader* r) { return static_cast<int64_t>(r->r.num_columns()); }
int64_t pqr_num_pages(const pqr_reader* r) { return static_cast<int64_t>(r->r.num_pages()); }
int64_t pqr_row_group_num_rows(const pqr_reader* r, int rg) {
    return guarded([&]() -> int64_t { return r->r.metadata().row_groups.at(rg).num_rows; }, -1);
}
double pqr_page_scan_seconds(const pqr_reader* r) { return r->r.page_scan_seconds(); }
uint64_t pqr_file_size(const pqr_reader* r) { return r->r.file_size(); }

int pqr_column_info(const pqr_reader* r, int col, pqr_colinfo* out) {
    return guarded([&]() -> int {
        const ColumnInfo& ci = r->r.column(static_cast<size_t>(col));
        std::memset(out, 0, sizeof(*out));
        std::strncpy(out->name, ci.name.c_str(), sizeof(out->name) - 1);
        out->type = static_cast<int32_t>(ci.type);
        out->column_index = ci.column_index;
        out->max_def_level = ci.max_def_level;
        out->max_rep_level = ci.max_rep_level;
        out->repetition = ci.repetition ? static_cast<int32_t>(*ci.repetition) : -1;
        out->converted = ci.converted_type ? static_cast<int32_t>(*ci.converted_type) : -1;
        return 0;
    }, -1);
}

int pqr_find_column(const pqr_reader* r, const char* name) { return r->r.find_column(name); }

int pqr_schema_string(const pqr_reader* r, char* buf, int64_t cap) {
    std::string s = r->r.schema_string();
    if (static_cast<int64_t>(s.size()) + 1 > cap) return -1;
    std::memcpy(buf, s.c_str(), s.size() + 1);
    return static_cast<int>(s.size());
}

int64_t pqr_page_index(const pqr_reader* r, pqr_page_entry* out, int64_t cap) {
    int64_t n = static_cast<int64_t>(r->r.num_pages());
    for (int64_t i = 0; i < n && i < cap; i++) {
        const PageIndexEntry& e = r->r.page_index_entry(static_cast<size_t>(i));
        out[i] = {e.data_offset, e.data_size, e.row_group_idx, e.column_idx};
    }
    return n;
}

int64_t pqr_read_page_data(const pqr_reader* r, int64_t id, uint8_t* buf, int64_t cap) {
    return guarded([&]() -> int64_t {
        auto d = r->r.read_page_data(static_cast<size_t>(id));
        if (static_cast<int64_t>(d.size()) > cap) { g_err = "buffer too small"; return -2; }
        std::memcpy(buf, d.data(), d.size());
        return static_cast<int64_t>(d.size());
    }, -1);
}

int64_t pqr_read_pages_chunk(const pqr_reader* r, int64_t s, int64_t e, int64_t max_bytes, uint8_t* buf, int64_t cap) {
    return guarded([&]() -> int64_t {
        auto d = r->r.read_pages_chunk(static_cast<size_t>(s), static_cast<size_t>(e), static_cast<size_t>(max_bytes));
        if (static_cast<int64_t>(d.size()) > cap) { g_err = "buffer too small"; return -2; }
        std::memcpy(buf, d.data(), d.size());
        return static_cast<int64_t>(d.size());
    }, -1);
}

int pqr_read_column_by_idx(pqr_reader* r, int rg, int col, pqr_valdump* out) {
    return guarded([&]() -> int { dump_values(r->r.read_column_by_idx(rg, col), out); return 0; }, -1);
}
int pqr_read_column(pqr_reader* r, const char* name, pqr_valdump* out) {
    return guarded([&]() -> int { dump_values(r->r.read_column(name), out); return 0; }, -1);
}
int pqr_read_column_rg(pqr_reader* r, const char* name, int64_t rg, pqr_valdump* out) {
    return guarded([&]() -> int { dump_values(r->r.read_column(name, static_cast<size_t>(rg)), out); return 0; }, -1);
}

// Drives the mirror of ColumnReader exactly like the reference's README shows: a
// read_range callback over the reader plus the chunk's metadata.
int pqr_read_pages(pqr_reader* r, int rg, int col, pqr_pagedump* out) {
    return guarded([&]() -> int {
        ParquetReader& pr = r->r;
        const ColumnInfo& ci = pr.column(static_cast<size_t>(col));
        const ColumnChunk& chunk = pr.metadata().row_groups.at(rg).columns.at(ci.column_index);
        ColumnReader cr([&pr](size_t o, size_t l) { return pr.read_range(o, l); }, chunk, ci.type,
                        ci.max_def_level, ci.max_rep_level);
        std::vector<PageResult> pages = cr.read_pages();
        int64_t np = static_cast<int64_t>(pages.size());
        out->n_pages = np;
        out->page_num = static_cast<int32_t*>(std::calloc(np + 1, 4));
        out->page_type = static_cast<int32_t*>(std::calloc(np + 1, 4));
        out->num_values = static_cast<int32_t*>(std::calloc(np + 1, 4));
        out->first_value = static_cast<int64_t*>(std::calloc(np + 1, 8));
        std::vector<Value> all;
        for (int64_t p = 0; p < np; p++) {
            out->page_num[p] = pages[p].page_num;
            out->page_type[p] = static_cast<int32_t>(pages[p].type);
            out->num_values[p] = pages[p].num_values;
            out->first_value[p] = static_cast<int64_t>(all.size());
            for (auto& v : pages[p].values) all.push_back(std::move(v));
        }
        out->first_value[np] = static_cast<int64_t>(all.size());
        dump_values(all, &out->values);
        return 0;
    }, -1);
}

int pqr_string_iterator_dump(pqr_reader* r, const char* name, pqr_strdump* out) {
    return guarded([&]() -> int {
        StringColumnIterator it = r->r.column_iterator(name);
        std::vector<uint64_t> pos, off;
        std::string chars;
        off.push_back(0);
        while (it.has_next()) {
            auto [p, len, ptr] = it.next();
            pos.push_back(p);
            chars.append(ptr, len);
            off.push_back(chars.size());
        }
        out->n = static_cast<int64_t>(pos.size());
        out->pos = static_cast<uint64_t*>(std::malloc((pos.size() + 1) * 8));
        out->off = static_cast<uint64_t*>(std::malloc(off.size() * 8));
        out->chars = static_cast<uint8_t*>(std::malloc(chars.size() + 1));
        std::memcpy(out->pos, pos.data(), pos.size() * 8);
        std::memcpy(out->off, off.data(), off.size() * 8);
        std::memcpy(out->chars, chars.data(), chars.size());
        return 0;
    }, -1);
}

void pqr_valdump_free(pqr_valdump* d) {
    std::free(d->is_null); std::free(d->vidx); std::free(d->fixed); std::free(d->str_off); std::free(d->chars);
    std::memset(d, 0, sizeof(*d));
}
void pqr_pagedump_free(pqr_pagedump* d) {
    std::free(d->page_num); std::free(d->page_type); std::free(d->num_values); std::free(d->first_value);
    pqr_valdump_free(&d->values);
    std::memset(d, 0, sizeof(*d));
}
void pqr_strdump_free(pqr_strdump* d) {
    std::free(d->pos); std::free(d->off); std::free(d->chars);
    std::memset(d, 0, sizeof(*d));
}

struct ColumnarOwner {
    DecodedColumn d;
    std::vector<uint64_t> row_base;
};

int pqr_read_columnar(pqr_reader* r, int col, int rg, pqr_columnar* out) {
    return guarded([&]() -> int {
        auto* o = new ColumnarOwner();
        try { o->d = r->r.read_column_columnar(col, rg); } catch (...) { delete o; throw; }
        for (const auto& c : o->d.chunks) o->row_base.push_back(c.out_row_base);
        const DecodedColumn& d = o->d;
        std::memset(out, 0, sizeof(*out));
        out->type = static_cast<int32_t>(d.type);
        out->width = d.width;
        out->num_slots = d.num_slots;
        out->has_validity = d.has_validity;
        out->n_chunks = static_cast<uint32_t>(d.chunks.size());
        out->values = d.values.data();
        out->validity = d.validity.data();
        out->offsets = d.offsets.data();
        out->char_bases = d.char_bases.data();
        out->chars = d.chars.data();
        out->chars_size = d.chars.size();
        out->chunk_row_base = o->row_base.data();
        out->bytes_in = d.bytes_in;
        out->bytes_out = d.bytes_out;
        out->kernel_ms = d.kernel_ms;
        out->owner = o;
        return 0;
    }, -1);
}
int pqr_read_columns_into(pqr_reader* r, const int32_t* cols, int32_t n_cols, int32_t rg, const pqr_dst* dsts, pqr_read_stats* stats) {
    return guarded([&]() -> int {
        std::vector<int> ci(cols, cols + (n_cols > 0 ? n_cols : 0));
        std::vector<ColumnDst> d(ci.size());
        std::vector<ColumnReadStats> st(ci.size());
        for (size_t i = 0; i < ci.size(); i++) d[i] = ColumnDst{dsts[i].values, dsts[i].values_cap, dsts[i].validity, dsts[i].validity_cap};
        r->r.read_columns_into(ci.data(), static_cast<int>(ci.size()), rg, d.data(), st.data());
        if (stats) for (size_t i = 0; i < ci.size(); i++)
            stats[i] = pqr_read_stats{st[i].num_slots, st[i].width, st[i].has_validity, st[i].bytes_in, st[i].bytes_out, st[i].h2d_bytes, st[i].d2h_bytes};
        return 0;
    }, -1);
}
int pqr_read_columns_into_rgs(pqr_reader* r, const int32_t* cols, int32_t n_cols, int64_t rg_begin, int64_t rg_end, const pqr_dst* dsts,
                              pqr_read_stats* stats) {
    return guarded([&]() -> int {
        if (rg_begin < 0 || rg_end < rg_begin) throw std::runtime_error("Invalid row group index");
        std::vector<int> ci(cols, cols + (n_cols > 0 ? n_cols : 0));
        std::vector<ColumnDst> d(ci.size());
        std::vector<ColumnReadStats> st(ci.size());
        for (size_t i = 0; i < ci.size(); i++) d[i] = ColumnDst{dsts[i].values, dsts[i].values_cap, dsts[i].validity, dsts[i].validity_cap};
        r->r.read_columns_into_range(ci.data(), static_cast<int>(ci.size()), static_cast<size_t>(rg_begin), static_cast<size_t>(rg_end), d.data(), st.data());
        if (stats) for (size_t i = 0; i < ci.size(); i++)
            stats[i] = pqr_read_stats{st[i].num_slots, st[i].width, st[i].has_validity, st[i].bytes_in, st[i].bytes_out, st[i].h2d_bytes, st[i].d2h_bytes};
        return 0;
    }, -1);
}
int pqr_read_dictionary_indices_into(pqr_reader* r, int32_t col, int64_t rg_begin, int64_t rg_end, const pqr_dst* dst, pqr_read_stats* stats) {
    return guarded([&]() -> int {
        if (rg_begin < 0 || rg_end < rg_begin || !dst) throw std::runtime_error("Invalid row group index");
        ColumnReadStats st;
        r->r.read_dictionary_indices_into(col, static_cast<size_t>(rg_begin), static_cast<size_t>(rg_end),
                                          ColumnDst{dst->values, dst->values_cap, dst->validity, dst->validity_cap}, &st);
        if (stats) *stats = pqr_read_stats{st.num_slots, st.width, st.has_validity, st.bytes_in, st.bytes_out, st.h2d_bytes, st.d2h_bytes};
        return 0;
    }, -1);
}
int pqr_read_strings_into(pqr_reader* r, int32_t col, int64_t rg_begin, int64_t rg_end, const pqr_strings_dst* dst, pqr_strings_stats* stats) {
    return guarded([&]() -> int {
        if (rg_begin < 0 || rg_end < rg_begin || !dst) throw std::runtime_error("Invalid row group index");
        StringsReadStats st;
        r->r.read_strings_into_range(col, static_cast<size_t>(rg_begin), static_cast<size_t>(rg_end),
                                     StringsDst{dst->offsets, dst->offsets_cap, dst->chars, dst->chars_cap, dst->validity, dst->validity_cap,
                                                dst->char_bases, dst->char_bases_cap}, &st);
        if (stats) *stats = pqr_strings_stats{st.num_slots, st.n_chunks, st.chars_size, st.has_validity, 0, st.bytes_in, st.bytes_out, st.h2d_bytes, st.d2h_bytes};
        return 0;
    }, -1);
}
int pqr_chunk_dictionary(const pqr_reader* r, int32_t col, int64_t rg, uint32_t* offsets, int64_t offsets_cap, uint8_t* chars,
                         int64_t chars_cap, int64_t* n_entries, int64_t* n_bytes) {
    return guarded([&]() -> int {
        if (rg < 0) throw std::runtime_error("Invalid row group index");
        std::vector<uint32_t> off;
        std::vector<uint8_t> ch;
        r->r.chunk_dictionary(col, static_cast<size_t>(rg), off, ch);
        if (n_entries) *n_entries = static_cast<int64_t>(off.size()) - 1;
        if (n_bytes) *n_bytes = static_cast<int64_t>(ch.size());
        if (offsets) {
            if (offsets_cap < static_cast<int64_t>(off.size())) throw std::runtime_error("pqr_chunk_dictionary: offsets buffer too small");
            std::memcpy(offsets, off.data(), off.size() * 4);
        }
        if (chars && !ch.empty()) {
            if (chars_cap < static_cast<int64_t>(ch.size())) throw std::runtime_error("pqr_chunk_dictionary: chars buffer too small");
            std::memcpy(chars, ch.data(), ch.size());
        }
        return 0;
    }, -1);
}
void pqr_release_plans(pqr_reader* r) { if (r) r->r.release_plans(); }

int pqr_shard_row_groups(const pqr_reader* r, int col, int n_shards, int32_t* out_begin) {
    return guarded([&]() -> int {
        auto v = r->r.shard_row_groups(col, n_shards);
        for (size_t i = 0; i < v.size(); i++) out_begin[i] = v[i];
        return 0;
    }, -1);
}
int64_t pqr_regex_prune_rgs(pqr_reader* r, int col, int64_t rg_begin, int64_t rg_end, const char* pattern, int neg, uint8_t* bits,
                            int64_t cap, float* kernel_ms) {
    return guarded([&]() -> int64_t {
        if (rg_begin < 0 || rg_end < rg_begin) throw std::runtime_error("Invalid row group index");
        return regex_prune_rgs(r->r, col, static_cast<size_t>(rg_begin), static_cast<size_t>(rg_end), pattern, neg != 0, bits, cap, kernel_ms);
    }, -1);
}
int64_t pqr_chunk_index_rgs(pqr_reader* r, const char* name, int64_t rg_begin, int64_t rg_end, uint64_t chunk_size, uint64_t carry_in,
                            uint32_t id_base, uint32_t* ids, int64_t cap, uint64_t* carry_out) {
    return guarded([&]() -> int64_t {
        if (rg_begin < 0 || rg_end < rg_begin) throw std::runtime_error("Invalid row group index");
        return chunk_index_rgs(r->r, name, static_cast<size_t>(rg_begin), static_cast<size_t>(rg_end), chunk_size, carry_in, id_base, ids, cap, carry_out);
    }, -1);
}

pqr_chunk_job* pqr_chunk_index_prepare_rgs(pqr_reader* r, const char* name, int64_t rg_begin, int64_t rg_end, uint64_t chunk_size,
                                           uint64_t* num_slots, float* decode_ms, float* prepare_ms) {
    return guarded([&]() -> pqr_chunk_job* {
        if (rg_begin < 0 || rg_end < rg_begin) throw std::runtime_error("Invalid row group index");
        ChunkIndexJob* j = chunk_index_prepare_rgs(r->r, name, static_cast<size_t>(rg_begin), static_cast<size_t>(rg_end), chunk_size);
        if (num_slots) *num_slots = j->num_slots;
        if (decode_ms) *decode_ms = j->decode_ms;
        if (prepare_ms) *prepare_ms = j->prepare_ms;
        return reinterpret_cast<pqr_chunk_job*>(j);
    }, static_cast<pqr_chunk_job*>(nullptr));
}
int64_t pqr_chunk_index_stitch(pqr_chunk_job* job, uint64_t carry_in, uint64_t* carry_out) {
    return guarded([&]() -> int64_t {
        if (!job) throw std::runtime_error("pqr_chunk_index_stitch: no job");
        return chunk_index_stitch(*reinterpret_cast<ChunkIndexJob*>(job), carry_in, carry_out);
    }, static_cast<int64_t>(-1));
}
int pqr_chunk_index_emit(pqr_chunk_job* job, uint32_t id_base, uint32_t* ids, int64_t cap, float* kernel_ms) {
    return guarded([&]() -> int {
        if (!job) throw std::runtime_error("pqr_chunk_index_emit: no job");
        chunk_index_emit(*reinterpret_cast<ChunkIndexJob*>(job), id_base, ids, cap, kernel_ms);
        return 0;
    }, -1);
}
void pqr_chunk_job_free(pqr_chunk_job* job) { delete reinterpret_cast<ChunkIndexJob*>(job); }

void pqr_columnar_free(pqr_columnar* c) {
    if (c && c->owner) delete static_cast<ColumnarOwner*>(c->owner);
    if (c) std::memset(c, 0, sizeof(*c));
}

int pqr_column_tables(const pqr_reader* r, int col, int rg, pqr_tables* out) {
    return guarded([&]() -> int {
        if (col < 0 || col >= static_cast<int>(r->r.num_columns())) throw std::runtime_error("Invalid column index");
        if (rg >= static_cast<int>(r->r.num_row_groups())) throw std::runtime_error("Invalid row group index");
        ColumnTables t = r->r.column_tables(col, rg);
        if (t.ext) throw std::runtime_error("tables of compressed / DATA_PAGE_V2 chunks are not exported (decode them through pqr_read_column*)");
        out->n_chunks = static_cast<uint32_t>(t.chunks.size());
        out->n_pages = static_cast<uint32_t>(t.pages.size());
        out->total_slots = t.total_slots;
        out->chunks = static_cast<pqg_chunk_desc*>(std::malloc(sizeof(pqg_chunk_desc) * (t.chunks.size() + 1)));
        out->pages = static_cast<pqg_page_desc*>(std::malloc(sizeof(pqg_page_desc) * (t.pages.size() + 1)));
        std::memcpy(out->chunks, t.chunks.data(), sizeof(pqg_chunk_desc) * t.chunks.size());
        std::memcpy(out->pages, t.pages.data(), sizeof(pqg_page_desc) * t.pages.size());
        return 0;
    }, -1);
}
int pqr_column_tables_rgs(const pqr_reader* r, int col, int64_t rg_begin, int64_t rg_end, pqr_tables* out) {
    return guarded([&]() -> int {
        if (col < 0 || col >= static_cast<int>(r->r.num_columns())) throw std::runtime_error("Invalid column index");
        if (rg_begin < 0 || rg_end < rg_begin || rg_end > static_cast<int64_t>(r->r.num_row_groups())) throw std::runtime_error("Invalid row group index");
        ColumnTables t = r->r.column_tables_range(col, static_cast<size_t>(rg_begin), static_cast<size_t>(rg_end));
        if (t.ext) throw std::runtime_error("tables of compressed / DATA_PAGE_V2 chunks are not exported (decode them through pqr_read_column*)");
        out->n_chunks = static_cast<uint32_t>(t.chunks.size());
        out->n_pages = static_cast<uint32_t>(t.pages.size());
        out->total_slots = t.total_slots;
        out->chunks = static_cast<pqg_chunk_desc*>(std::malloc(sizeof(pqg_chunk_desc) * (t.chunks.size() + 1)));
        out->pages = static_cast<pqg_page_desc*>(std::malloc(sizeof(pqg_page_desc) * (t.pages.size() + 1)));
        std::memcpy(out->chunks, t.chunks.data(), sizeof(pqg_chunk_desc) * t.chunks.size());
        std::memcpy(out->pages, t.pages.data(), sizeof(pqg_page_desc) * t.pages.size());
        return 0;
    }, -1);
}
int pqr_columns_tables(const pqr_reader* r, const int* cols, int n_cols, int rg, pqr_tables* out) {
    return guarded([&]() -> int {
        if (!cols || n_cols <= 0) throw std::runtime_error("pqr_columns_tables: no columns");
        if (rg >= static_cast<int>(r->r.num_row_groups())) throw std::runtime_error("Invalid row group index");
        auto width_of = [](int phys) { return phys == PQG_INT32 || phys == PQG_FLOAT ? 4 : (phys == PQG_INT64 || phys == PQG_DOUBLE ? 8 : 0); };
        std::vector<ColumnTables> ts;
        int width = 0;
        uint64_t slots = 0;
        for (int k = 0; k < n_cols; k++) {
            if (cols[k] < 0 || cols[k] >= static_cast<int>(r->r.num_columns())) throw std::runtime_error("Invalid column index");
            ts.push_back(r->r.column_tables(cols[k], rg));
            const ColumnTables& t = ts.back();
            if (t.ext) throw std::runtime_error("tables of compressed / DATA_PAGE_V2 chunks are not exported (decode them through pqr_read_column*)");
            for (const pqg_chunk_desc& c : t.chunks) {
                const int w = width_of(c.phys_type);
                if (w == 0 || (width && w != width)) throw std::runtime_error("pqr_columns_tables: the columns of one plan must share a 4- or 8-byte value width");
                width = w;
            }
            if (k && t.total_slots != slots) throw std::runtime_error("pqr_columns_tables: the columns must have the same number of slots");
            slots = t.total_slots;
        }
        struct Item { int k; uint32_t c; uint64_t dict_bytes; };
        std::vector<Item> items;
        size_t n_pages = 0;
        for (int k = 0; k < n_cols; k++) {
            for (uint32_t c = 0; c < ts[k].chunks.size(); c++) {
                const pqg_chunk_desc& d = ts[k].chunks[c];
                items.push_back({k, c, d.has_dict ? static_cast<uint64_t>(d.dict_num_values) * width : 0});
            }
            n_pages += ts[k].pages.size();
        }
        std::stable_sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.dict_bytes > b.dict_bytes; });
        out->n_chunks = static_cast<uint32_t>(items.size());
        out->n_pages = static_cast<uint32_t>(n_pages);
        out->total_slots = slots * static_cast<uint64_t>(n_cols);
        out->chunks = static_cast<pqg_chunk_desc*>(std::malloc(sizeof(pqg_chunk_desc) * (items.size() + 1)));
        out->pages = static_cast<pqg_page_desc*>(std::malloc(sizeof(pqg_page_desc) * (n_pages + 1)));
        uint32_t pg = 0;
        for (uint32_t i = 0; i < items.size(); i++) {
            const ColumnTables& t = ts[items[i].k];
            const uint64_t base = slots * static_cast<uint64_t>(items[i].k);
            pqg_chunk_desc d = t.chunks[items[i].c];
            const uint32_t first = d.first_page;
            d.first_page = pg;
            d.out_row_base += base;
            out->chunks[i] = d;
            for (uint32_t q = 0; q < d.n_pages; q++) {
                pqg_page_desc pd = t.pages[first + q];
                pd.chunk_idx = i;
                pd.out_row_base += base;
                out->pages[pg++] = pd;
            }
        }
        return 0;
    }, -1);
}
void pqr_tables_free(pqr_tables* t) {
    if (!t) return;
    std::free(t->chunks); std::free(t->pages);
    std::memset(t, 0, sizeof(*t));
}

int64_t pqr_chunk_index(pqr_reader* r, const char* name, uint64_t chunk_size, uint64_t* tuple_to_chunk, int64_t num_rows) {
    return guarded([&]() -> int64_t { return chunk_index(r->r, name, chunk_size, tuple_to_chunk, num_rows); }, -1);
}
int64_t pqr_regex_prune(pqr_reader* r, int col, const char* pattern, int neg, uint8_t* bits, int64_t cap, float* kernel_ms) {
    return guarded([&]() -> int64_t { return regex_prune(r->r, col, pattern, neg != 0, bits, cap, kernel_ms); }, -1);
}
int64_t pqr_page_chunk_index(pqr_reader* r, int col, uint64_t chunk_size, uint32_t* page_chunk, uint32_t* page_off,
                             uint32_t* chunk_first_page, int64_t cap, int64_t* first_global_page, int64_t* n_col_pages) {
    return guarded([&]() -> int64_t {
        return page_chunk_index(r->r, col, chunk_size, page_chunk, page_off, chunk_first_page, cap, first_global_page, n_col_pages);
    }, -1);
}

} // extern "C"
