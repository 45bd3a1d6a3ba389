// ref_shim.cpp -- TEST INFRASTRUCTURE. A thin extern "C" driver around the UNMODIFIED
// reference sources (compiled from /root/reference where they lie, see oracle/Makefile;
// nothing is copied into this repo).  Output: oracle/_ref/libpqref.so (git-ignored).
//
// It exposes exactly the reference entry points the hot path replaces, so that the
// parity tests and bench.py's cpu_baseline / --impl reference legs can call the real
// thing:
//   ParquetWriter::write_row_group           src/writer/parquet_writer.cpp:376   (fixture source)
//   ParquetReader::open / read_column*       src/reader/parquet_reader.cpp:14,125-165
//   ColumnReader::read_all / read_pages      src/reader/column_reader.cpp:18,73
//   page index + raw page API                src/reader/parquet_reader.cpp:182-238,559-605
//   StringColumnIterator                     src/reader/parquet_reader.cpp:282-465
//   chunk-index prototype loop               src/main.cpp:21-32 (driven over the real iterator)
//   RleDecoder / RleBpEncoder                include/reader/rle_decoder.hpp, include/writer/rle_bp_encoder.hpp
//
// Only tests/, smoke() and bench.py's CPU-baseline legs may load this library.
#include "reader/parquet_reader.hpp"
#include "writer/parquet_writer.hpp"
#include "writer/rle_bp_encoder.hpp"
#include "valdump.h"

#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#define REF_API extern "C" __attribute__((visibility("default")))

static thread_local std::string g_err;

REF_API const char* ref_last_error() { return g_err.c_str(); }

// ── helpers ────────────────────────────────────────────────────────────────────────

static void dump_values(const std::vector<Value>& v, valdump* out) {
    int64_t n = static_cast<int64_t>(v.size());
    out->n = n;
    out->is_null = static_cast<uint8_t*>(std::calloc(n + 1, 1));
    out->vidx = static_cast<uint8_t*>(std::calloc(n + 1, 1));
    out->fixed = static_cast<uint64_t*>(std::calloc(n + 1, 8));
    out->str_off = static_cast<uint64_t*>(std::calloc(n + 1, 8));
    uint64_t total = 0;
    for (int64_t i = 0; i < n; i++) {
        if (v[i].data.index() == 5) total += std::get<std::string>(v[i].data).size();
    }
    out->chars = static_cast<uint8_t*>(std::malloc(total + 1));
    out->chars_len = static_cast<int64_t>(total);
    uint64_t pos = 0;
    for (int64_t i = 0; i < n; i++) {
        const Value& x = v[i];
        out->is_null[i] = x.is_null ? 1 : 0;
        out->vidx[i] = static_cast<uint8_t>(x.data.index());
        out->str_off[i] = pos;
        uint64_t bits = 0;
        switch (x.data.index()) {
            case 0: bits = std::get<bool>(x.data) ? 1 : 0; break;
            case 1: { int32_t t = std::get<int32_t>(x.data); uint32_t u; std::memcpy(&u, &t, 4); bits = u; break; }
            case 2: { int64_t t = std::get<int64_t>(x.data); std::memcpy(&bits, &t, 8); break; }
            case 3: { float t = std::get<float>(x.data); uint32_t u; std::memcpy(&u, &t, 4); bits = u; break; }
            case 4: { double t = std::get<double>(x.data); std::memcpy(&bits, &t, 8); break; }
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

REF_API void ref_valdump_free(valdump* d) {
    std::free(d->is_null); std::free(d->vidx); std::free(d->fixed);
    std::free(d->str_off); std::free(d->chars);
    std::memset(d, 0, sizeof(*d));
}

REF_API void ref_pagedump_free(pagedump* d) {
    std::free(d->page_num); std::free(d->page_type); std::free(d->num_values);
    std::free(d->first_value);
    ref_valdump_free(&d->values);
    std::memset(d, 0, sizeof(*d));
}

// ── writer (fixture generator: "files written by the repo's own writer") ───────────

struct ref_colspec {
    const char* name;
    int32_t type;        // ParquetType
    int32_t repetition;  // FieldRepetitionType
    int32_t converted;   // ConvertedType or -1
};

// One column of one row group in columnar form.  fixed: 8-byte slots holding the payload
// bits (bool/i32/float in the low bytes); strings: off[n+1] + chars.
struct ref_colin {
    const uint8_t* is_null;
    const uint64_t* fixed;
    const uint64_t* str_off;
    const uint8_t* chars;
};

struct RefWriter {
    ParquetWriter* w;
    std::vector<ColumnSpec> specs;
};

REF_API void* ref_writer_open(const char* path, int ncols, const ref_colspec* cols) {
    try {
        auto* rw = new RefWriter();
        for (int i = 0; i < ncols; i++) {
            ColumnSpec s;
            s.name = cols[i].name;
            s.type = static_cast<ParquetType>(cols[i].type);
            s.repetition = static_cast<FieldRepetitionType>(cols[i].repetition);
            if (cols[i].converted >= 0) s.converted_type = static_cast<ConvertedType>(cols[i].converted);
            rw->specs.push_back(s);
        }
        rw->w = new ParquetWriter(path, rw->specs);
        return rw;
    } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}

REF_API int ref_writer_write_row_group(void* h, int64_t nrows, const ref_colin* cols) {
    auto* rw = static_cast<RefWriter*>(h);
    try {
        std::vector<std::vector<Value>> data(rw->specs.size());
        for (size_t c = 0; c < rw->specs.size(); c++) {
            auto& col = data[c];
            col.reserve(static_cast<size_t>(nrows));
            const ref_colin& in = cols[c];
            for (int64_t i = 0; i < nrows; i++) {
                if (in.is_null && in.is_null[i]) { col.push_back(Value::null()); continue; }
                switch (rw->specs[c].type) {
                    case ParquetType::BOOLEAN: col.push_back(Value::from_bool(in.fixed[i] != 0)); break;
                    case ParquetType::INT32: { int32_t t; std::memcpy(&t, &in.fixed[i], 4); col.push_back(Value::from_i32(t)); break; }
                    case ParquetType::INT64: { int64_t t; std::memcpy(&t, &in.fixed[i], 8); col.push_back(Value::from_i64(t)); break; }
                    case ParquetType::FLOAT: { float t; std::memcpy(&t, &in.fixed[i], 4); col.push_back(Value::from_float(t)); break; }
                    case ParquetType::DOUBLE: { double t; std::memcpy(&t, &in.fixed[i], 8); col.push_back(Value::from_double(t)); break; }
                    case ParquetType::BYTE_ARRAY:
                        col.push_back(Value::from_string(std::string(
                            reinterpret_cast<const char*>(in.chars + in.str_off[i]),
                            static_cast<size_t>(in.str_off[i + 1] - in.str_off[i]))));
                        break;
                    default: throw std::runtime_error("ref_shim: unsupported writer type");
                }
            }
        }
        rw->w->write_row_group(data);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

REF_API int ref_writer_close(void* h) {
    auto* rw = static_cast<RefWriter*>(h);
    int rc = 0;
    try { rw->w->close(); } catch (const std::exception& e) { g_err = e.what(); rc = -1; }
    delete rw->w;
    delete rw;
    return rc;
}

// ── reader ─────────────────────────────────────────────────────────────────────────

REF_API void* ref_reader_open(const char* path) {
    auto* r = new ParquetReader();
    try {
        if (!r->open(path)) { g_err = "open failed"; delete r; return nullptr; }
    } catch (const std::exception& e) { g_err = e.what(); delete r; return nullptr; }
    return r;
}
REF_API void ref_reader_close(void* h) { delete static_cast<ParquetReader*>(h); }
REF_API int64_t ref_num_rows(void* h) { return static_cast<ParquetReader*>(h)->num_rows(); }
REF_API int64_t ref_num_row_groups(void* h) { return static_cast<int64_t>(static_cast<ParquetReader*>(h)->num_row_groups()); }
REF_API int64_t ref_num_columns(void* h) { return static_cast<int64_t>(static_cast<ParquetReader*>(h)->num_columns()); }
REF_API int64_t ref_num_pages(void* h) { return static_cast<int64_t>(static_cast<ParquetReader*>(h)->num_pages()); }
REF_API int64_t ref_row_group_num_rows(void* h, int rg) {
    return static_cast<ParquetReader*>(h)->metadata().row_groups[rg].num_rows;
}

struct ref_colinfo {
    char name[256];
    int32_t type;
    int32_t column_index;
    int32_t max_def_level;
    int32_t max_rep_level;
    int32_t repetition;  // -1 if absent
    int32_t converted;   // -1 if absent
};

REF_API int ref_column_info(void* h, int col, ref_colinfo* out) {
    try {
        const ColumnInfo& ci = static_cast<ParquetReader*>(h)->column(static_cast<size_t>(col));
        std::memset(out, 0, sizeof(*out));
        std::strncpy(out->name, ci.name.c_str(), sizeof(out->name) - 1);
        out->type = static_cast<int32_t>(ci.type);
        out->column_index = ci.column_index;
        out->max_def_level = ci.max_def_level;
        out->max_rep_level = ci.max_rep_level;
        out->repetition = ci.repetition.has_value() ? static_cast<int32_t>(*ci.repetition) : -1;
        out->converted = ci.converted_type.has_value() ? static_cast<int32_t>(*ci.converted_type) : -1;
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

REF_API int ref_find_column(void* h, const char* name) {
    return static_cast<ParquetReader*>(h)->find_column(name);
}

REF_API int ref_schema_string(void* h, char* buf, int64_t cap) {
    std::string s = static_cast<ParquetReader*>(h)->schema_string();
    if (static_cast<int64_t>(s.size()) + 1 > cap) return -1;
    std::memcpy(buf, s.c_str(), s.size() + 1);
    return static_cast<int>(s.size());
}

REF_API int ref_read_column_by_idx(void* h, int rg, int col, valdump* out) {
    try {
        auto v = static_cast<ParquetReader*>(h)->read_column_by_idx(rg, col);
        dump_values(v, out);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

REF_API int ref_read_column(void* h, const char* name, valdump* out) {
    try {
        auto v = static_cast<ParquetReader*>(h)->read_column(name);
        dump_values(v, out);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

REF_API int ref_read_column_rg(void* h, const char* name, int64_t rg, valdump* out) {
    try {
        auto v = static_cast<ParquetReader*>(h)->read_column(name, static_cast<size_t>(rg));
        dump_values(v, out);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

REF_API int ref_read_pages(void* h, int rg, int col, pagedump* out) {
    try {
        auto* r = static_cast<ParquetReader*>(h);
        const ColumnInfo& ci = r->column(static_cast<size_t>(col));
        const auto& chunk = r->metadata().row_groups.at(rg).columns.at(ci.column_index);
        ColumnReader cr([r](size_t o, size_t l) { return r->read_range(o, l); }, chunk,
                        ci.type, ci.max_def_level, ci.max_rep_level);
        auto pages = cr.read_pages();
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
            all.insert(all.end(), pages[p].values.begin(), pages[p].values.end());
        }
        out->first_value[np] = static_cast<int64_t>(all.size());
        dump_values(all, &out->values);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

struct ref_page_entry { uint64_t data_offset, data_size, row_group_idx, column_idx; };

REF_API int ref_page_index_entry(void* h, int64_t id, ref_page_entry* out) {
    try {
        const PageIndexEntry& e = static_cast<ParquetReader*>(h)->page_index_entry(static_cast<size_t>(id));
        out->data_offset = e.data_offset; out->data_size = e.data_size;
        out->row_group_idx = e.row_group_idx; out->column_idx = e.column_idx;
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

// whole page index in one call (n entries must be available in out)
REF_API int64_t ref_page_index(void* h, ref_page_entry* out, int64_t cap) {
    auto* r = static_cast<ParquetReader*>(h);
    int64_t n = static_cast<int64_t>(r->num_pages());
    for (int64_t i = 0; i < n && i < cap; i++) {
        const PageIndexEntry& e = r->page_index_entry(static_cast<size_t>(i));
        out[i] = {e.data_offset, e.data_size, e.row_group_idx, e.column_idx};
    }
    return n;
}

REF_API int64_t ref_read_page_data(void* h, int64_t id, uint8_t* buf, int64_t cap) {
    try {
        auto d = static_cast<ParquetReader*>(h)->read_page_data(static_cast<size_t>(id));
        if (static_cast<int64_t>(d.size()) > cap) { g_err = "buffer too small"; return -2; }
        std::memcpy(buf, d.data(), d.size());
        return static_cast<int64_t>(d.size());
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

REF_API int64_t ref_read_pages_chunk(void* h, int64_t s, int64_t e, int64_t max_bytes,
                                     uint8_t* buf, int64_t cap) {
    try {
        auto d = static_cast<ParquetReader*>(h)->read_pages_chunk(
            static_cast<size_t>(s), static_cast<size_t>(e), static_cast<size_t>(max_bytes));
        if (static_cast<int64_t>(d.size()) > cap) { g_err = "buffer too small"; return -2; }
        std::memcpy(buf, d.data(), d.size());
        return static_cast<int64_t>(d.size());
    } catch (const std::exception& ex) { g_err = ex.what(); return -1; }
}

// StringColumnIterator drained into (positions, offsets, chars).
struct ref_strdump {
    int64_t n;
    uint64_t* pos;
    uint64_t* off; /* n + 1 */
    uint8_t* chars;
};

REF_API int ref_string_iterator_dump(void* h, const char* name, ref_strdump* out) {
    try {
        auto it = static_cast<ParquetReader*>(h)->column_iterator(name);
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
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

REF_API void ref_strdump_free(ref_strdump* d) {
    std::free(d->pos); std::free(d->off); std::free(d->chars);
    std::memset(d, 0, sizeof(*d));
}

// The loop of src/main.cpp:21-32, verbatim in behaviour, over the real iterator, with the
// file, column and chunk size as arguments instead of the hard-coded ones (main.cpp:7-8).
// tuple_to_chunk must hold num_rows entries; returns chunk_id + 1 ("Total chunks").
REF_API int64_t ref_chunk_index(void* h, const char* name, uint64_t chunk_size,
                                uint64_t* tuple_to_chunk, int64_t num_rows) {
    try {
        auto* r = static_cast<ParquetReader*>(h);
        for (int64_t i = 0; i < num_rows; i++) tuple_to_chunk[i] = 0;
        StringColumnIterator it = r->column_iterator(name);
        std::string chunk;
        size_t chunk_id = 0;
        while (it.has_next()) {
            auto [pos, string_len, string] = it.next();
            if (chunk.size() >= chunk_size) {
                chunk.clear();
                chunk_id++;
            }
            chunk += std::to_string(string_len) + std::string(string, string_len);
            if (static_cast<int64_t>(pos) < num_rows) tuple_to_chunk[pos] = chunk_id;
        }
        return static_cast<int64_t>(chunk_id + 1);
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

// ── header-only codecs, exposed for unit tests of bit widths 1..32 ─────────────────

// RleBpEncoder (include/writer/rle_bp_encoder.hpp): returns encoded length, or -needed.
REF_API int64_t ref_rle_encode(const uint32_t* values, int64_t n, int bit_width,
                               uint8_t* out, int64_t cap) {
    RleBpEncoder enc(static_cast<uint8_t>(bit_width));
    for (int64_t i = 0; i < n; i++) enc.WriteValue(values[i]);
    std::vector<uint8_t> buf;
    enc.FinishWrite(buf);
    if (static_cast<int64_t>(buf.size()) > cap) return -static_cast<int64_t>(buf.size());
    std::memcpy(out, buf.data(), buf.size());
    return static_cast<int64_t>(buf.size());
}

// RleDecoder::get_batch<int32_t> (include/reader/rle_decoder.hpp:17-34).  The caller must
// pad `data` with >= 8 readable bytes past `size` (the reference reads literal bits without
// a bounds check, rle_decoder.hpp:59-62).
REF_API void ref_rle_decode_i32(const uint8_t* data, uint32_t size, int bit_width,
                                int32_t* out, uint32_t count) {
    RleDecoder d(data, size, static_cast<uint8_t>(bit_width));
    d.get_batch<int32_t>(out, count);
}
REF_API void ref_rle_decode_i16(const uint8_t* data, uint32_t size, int bit_width,
                                int16_t* out, uint32_t count) {
    RleDecoder d(data, size, static_cast<uint8_t>(bit_width));
    d.get_batch<int16_t>(out, count);
}

// ── CPU baseline timing: the reference's own read path, T threads, own reader each ─
// (BASELINE.md section 3).  Work items are (rg, col) pairs handed out by an atomic
// counter; every vector<Value> is freed before the next chunk is read.
REF_API double ref_time_read_chunks(const char* path, const int32_t* rgs, const int32_t* cols,
                                    int64_t n_items, int threads, int64_t* values_out) {
    // open() (footer + build_page_index) happens before the clock starts: the timed
    // region is read_column_by_idx only, like the GPU arm's decode call.
    std::vector<std::unique_ptr<ParquetReader>> readers;
    for (int t = 0; t < threads; t++) {
        readers.emplace_back(new ParquetReader());
        if (!readers.back()->open(path)) { g_err = "ref_time_read_chunks: open failed"; return -1.0; }
    }
    std::atomic<int64_t> next{0};
    std::atomic<int64_t> total{0};
    std::atomic<int> failed{0};
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) {
        ParquetReader* r = readers[t].get();
        pool.emplace_back([&, r]() {
            try {
                for (;;) {
                    int64_t i = next.fetch_add(1);
                    if (i >= n_items) break;
                    auto v = r->read_column_by_idx(rgs[i], cols[i]);
                    total += static_cast<int64_t>(v.size());
                }
            } catch (...) { failed = 1; }
        });
    }
    for (auto& th : pool) th.join();
    auto t1 = std::chrono::steady_clock::now();
    if (values_out) *values_out = total.load();
    if (failed) { g_err = "ref_time_read_chunks: a worker failed"; return -1.0; }
    return std::chrono::duration<double>(t1 - t0).count();
}
