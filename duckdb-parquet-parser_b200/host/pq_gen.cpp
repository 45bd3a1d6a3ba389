// pq_gen.cpp -- synthetic workload / fixture generator, see include/pqg_gen.h.
//
// Columnar arrays -> Parquet file image, byte-identical to the reference's ParquetWriter
// (src/writer/parquet_writer.cpp) for the same values, but organised for throughput: one
// task per column chunk on a thread pool, flat arrays instead of 48-byte Values, an
// open-addressing hash for the dictionary decision, chunk buffers laid out by a prefix sum.
// Not on the decode path.
#include "pqg_gen.h"

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

thread_local std::string g_err;

constexpr size_t kMaxPage = 1024; // MAX_UNCOMPRESSED_PAGE_SIZE, include/writer/parquet_writer.hpp:35

struct Bytes {
    std::vector<uint8_t> v;
    void u8(uint8_t b) { v.push_back(b); }
    void raw(const void* p, size_t n) { const uint8_t* q = static_cast<const uint8_t*>(p); v.insert(v.end(), q, q + n); }
    void le32(uint32_t x) { raw(&x, 4); }
    void varint(uint64_t x) { while (x >= 0x80) { v.push_back(static_cast<uint8_t>(x | 0x80)); x >>= 7; } v.push_back(static_cast<uint8_t>(x)); }
    void zigzag(int64_t x) { varint((static_cast<uint64_t>(x) << 1) ^ static_cast<uint64_t>(x >> 63)); }
};

// Thrift compact field header with the short (delta) form; `last` tracks the previous id.
struct Thrift {
    Bytes& b;
    int last = 0;
    void field(int id, int type) {
        int delta = id - last;
        if (delta > 0 && delta <= 15) b.u8(static_cast<uint8_t>((delta << 4) | type));
        else { b.u8(static_cast<uint8_t>(type)); b.zigzag(id); }
        last = id;
    }
    void i32(int id, int64_t x) { field(id, 5); b.zigzag(x); }
    void i64(int id, int64_t x) { field(id, 6); b.zigzag(x); }
    void str(int id, const std::string& s) { field(id, 8); b.varint(s.size()); b.raw(s.data(), s.size()); }
    void list(int id, int elem, int64_t n) {
        field(id, 9);
        if (n < 15) b.u8(static_cast<uint8_t>((n << 4) | elem));
        else { b.u8(static_cast<uint8_t>(0xF0 | elem)); b.varint(static_cast<uint64_t>(n)); }
    }
    void stop() { b.u8(0); }
};

int width_of(int type) {
    switch (type) {
        case PQG_BOOLEAN: return 1;
        case PQG_INT32: case PQG_FLOAT: return 4;
        case PQG_INT64: case PQG_DOUBLE: return 8;
        default: return 0;
    }
}

uint8_t index_bit_width(uint32_t dict_size) { // compute_bit_width(dict_size - 1), minimum 1
    uint32_t m = dict_size > 0 ? dict_size - 1 : 0;
    if (m == 0) return 1;
    uint8_t bw = 0;
    while (m) { bw++; m >>= 1; }
    return bw;
}

// PageHeader for a data page: {1 type=DATA_PAGE, 2 size, 3 size, 5 {1 n, 2 encoding, 3 RLE, 4 RLE}}
void data_page_header(Bytes& out, uint32_t size, uint32_t n, int encoding) {
    Thrift t{out};
    t.i32(1, 0); t.i32(2, size); t.i32(3, size);
    t.field(5, 12);
    Thrift h{out};
    h.i32(1, n); h.i32(2, encoding); h.i32(3, 3); h.i32(4, 3); h.stop();
    t.stop();
}

void dict_page_header(Bytes& out, uint32_t size, uint32_t n) {
    Thrift t{out};
    t.i32(1, 2); t.i32(2, size); t.i32(3, size);
    t.field(7, 12);
    Thrift h{out};
    h.i32(1, n); h.i32(2, 2); h.stop();
    t.stop();
}

struct Chunk {
    Bytes bytes;
    bool dict = false;
    uint64_t dict_page_bytes = 0; // header + payload of the dictionary page
    int64_t num_values = 0;
    std::string err;
};

struct ColView {
    const pqgen_col* c;
    int64_t a, b; // row range of the chunk
    bool null_at(int64_t i) const { return c->is_null && c->is_null[i]; }
};

// definition levels of rows [p0, p1): maximal RLE runs, 1 value byte (bit width 1)
void emit_def_levels(Bytes& pay, const ColView& cv, int64_t p0, int64_t p1) {
    size_t len_pos = pay.v.size();
    pay.le32(0);
    int64_t i = p0;
    while (i < p1) {
        bool nul = cv.null_at(i);
        int64_t j = i + 1;
        while (j < p1 && cv.null_at(j) == nul) j++;
        pay.varint(static_cast<uint64_t>(j - i) << 1);
        pay.u8(nul ? 0 : 1);
        i = j;
    }
    uint32_t len = static_cast<uint32_t>(pay.v.size() - len_pos - 4);
    std::memcpy(&pay.v[len_pos], &len, 4);
}

// RLE / bit-packed hybrid exactly as RleBpEncoder (include/writer/rle_bp_encoder.hpp) emits
// it, restated over the whole index sequence: at a fresh position a maximal run of >= 4
// equal values (or a run that reaches the end of the input) becomes an RLE run; anything
// else opens ONE bit-packed group of 8 values (zero padded at the end of the input).
void emit_indices(Bytes& pay, const uint32_t* idx, size_t n, uint8_t bw) {
    const uint32_t nbytes = (bw + 7u) / 8u;
    size_t i = 0;
    while (i < n) {
        size_t run = 1;
        while (i + run < n && idx[i + run] == idx[i]) run++;
        if (run >= 4 || i + run == n) {
            pay.varint(static_cast<uint64_t>(run) << 1);
            uint32_t v = idx[i];
            for (uint32_t k = 0; k < nbytes; k++) { pay.u8(static_cast<uint8_t>(v)); v >>= 8; }
            i += run;
        } else {
            pay.u8(0x03);
            uint64_t acc = 0;
            uint32_t bits = 0;
            for (size_t k = 0; k < 8; k++) {
                uint64_t v = (i + k < n) ? idx[i + k] : 0;
                acc |= v << bits;
                bits += bw;
                while (bits >= 8) { pay.u8(static_cast<uint8_t>(acc)); acc >>= 8; bits -= 8; }
            }
            // 8 * bw bits is a whole number of bytes: nothing left
            i += 8;
        }
    }
}

struct FixedDict {
    std::vector<uint64_t> keys; // slot -> canonical key
    std::vector<uint32_t> slot_idx;
    std::vector<uint64_t> values; // dictionary index -> first-seen payload bits
    uint64_t mask = 0;
    void init(size_t cap_entries) {
        size_t n = 64;
        while (n < cap_entries * 2 + 2) n <<= 1;
        keys.assign(n, 0);
        slot_idx.assign(n, UINT32_MAX);
        mask = n - 1;
    }
    static uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
    uint32_t find_or_add(uint64_t key, uint64_t bits) {
        uint64_t s = mix(key) & mask;
        for (;;) {
            if (slot_idx[s] == UINT32_MAX) {
                slot_idx[s] = static_cast<uint32_t>(values.size());
                keys[s] = key;
                values.push_back(bits);
                return slot_idx[s];
            }
            if (keys[s] == key) return slot_idx[s];
            s = (s + 1) & mask;
        }
    }
};

uint64_t load_bits(const pqgen_col* c, int w, int64_t i) {
    uint64_t x = 0;
    std::memcpy(&x, static_cast<const uint8_t*>(c->fixed) + static_cast<size_t>(i) * w, w);
    if (c->type == PQG_BOOLEAN) x = x != 0;
    return x;
}

void encode_chunk(const ColView& cv, Chunk& out) {
    const pqgen_col* c = cv.c;
    const int64_t rows = cv.b - cv.a;
    const bool optional = c->repetition == 1;
    const int w = width_of(c->type);
    const bool is_str = c->type == PQG_BYTE_ARRAY;
    out.num_values = rows;
    if (rows == 0) return;

    // ---- dictionary decision (analyze_column, parquet_writer.cpp:255-283) ----
    std::vector<uint32_t> idx; // per non-null row, in order
    uint32_t dict_size = 0;
    FixedDict fd;
    std::vector<std::string_view> sdict;
    int64_t non_null = 0;
    bool use_dict = true, nan_seen = false;
    const size_t bail = static_cast<size_t>(rows / 5); // distinct > rows/5 => distinct > non_null/5
    idx.reserve(static_cast<size_t>(rows));
    if (!is_str) {
        fd.init(bail + 1);
        for (int64_t i = cv.a; i < cv.b; i++) {
            if (cv.null_at(i)) continue;
            non_null++;
            uint64_t bits = load_bits(c, w, i), key = bits;
            if (c->type == PQG_FLOAT) {
                float f; std::memcpy(&f, &bits, 4);
                if (std::isnan(f)) { nan_seen = true; continue; }
                if (f == 0.0f) key = 0;
            } else if (c->type == PQG_DOUBLE) {
                double d; std::memcpy(&d, &bits, 8);
                if (std::isnan(d)) { nan_seen = true; continue; }
                if (d == 0.0) key = 0;
            }
            idx.push_back(fd.find_or_add(key, bits));
            if (fd.values.size() > bail) { use_dict = false; break; }
        }
        dict_size = static_cast<uint32_t>(fd.values.size());
    } else {
        std::unordered_map<std::string_view, uint32_t> m;
        m.reserve(bail + 1);
        for (int64_t i = cv.a; i < cv.b; i++) {
            if (cv.null_at(i)) continue;
            non_null++;
            std::string_view s(reinterpret_cast<const char*>(c->chars) + c->str_off[i], c->str_off[i + 1] - c->str_off[i]);
            auto it = m.find(s);
            if (it == m.end()) { it = m.emplace(s, static_cast<uint32_t>(sdict.size())).first; sdict.push_back(s); }
            idx.push_back(it->second);
            if (sdict.size() > bail) { use_dict = false; break; }
        }
        dict_size = static_cast<uint32_t>(sdict.size());
    }
    if (use_dict && (dict_size == 0 || dict_size > static_cast<uint64_t>(non_null) / 5)) use_dict = false;
    // NaNs are left out of the distinct count: they can only add keys, so a PLAIN verdict
    // stands; a dictionary verdict would depend on std::map's behaviour for unordered keys.
    if (use_dict && nan_seen) { out.err = "NaN in a FLOAT/DOUBLE column that would be dictionary-encoded"; return; }

    Bytes& o = out.bytes;
    Bytes pay;
    if (use_dict) {
        out.dict = true;
        // dictionary page: PLAIN values in first-seen order (:285-312)
        if (!is_str) {
            for (uint64_t v : fd.values) pay.raw(&v, static_cast<size_t>(w));
        } else {
            for (auto s : sdict) { pay.le32(static_cast<uint32_t>(s.size())); pay.raw(s.data(), s.size()); }
        }
        dict_page_header(o, static_cast<uint32_t>(pay.v.size()), dict_size);
        o.raw(pay.v.data(), pay.v.size());
        out.dict_page_bytes = o.v.size();
        const uint8_t bw = index_bit_width(dict_size);
        size_t per_page = kMaxPage / std::max<size_t>(1, (bw + 7) / 8);
        if (per_page == 0) per_page = 1;
        o.v.reserve(o.v.size() + static_cast<size_t>(non_null) * (bw + 1) / 8 + static_cast<size_t>(rows) / per_page * 64 + static_cast<size_t>(rows) / 4 + 1024);
        size_t k = 0; // position in idx
        for (int64_t p0 = cv.a; p0 < cv.b; p0 += static_cast<int64_t>(per_page)) {
            int64_t p1 = std::min<int64_t>(cv.b, p0 + static_cast<int64_t>(per_page));
            pay.v.clear();
            size_t nn = static_cast<size_t>(p1 - p0);
            if (optional) {
                emit_def_levels(pay, cv, p0, p1);
                nn = 0;
                for (int64_t i = p0; i < p1; i++) nn += !cv.null_at(i);
            }
            pay.u8(bw);
            emit_indices(pay, idx.data() + k, nn, bw);
            k += nn;
            data_page_header(o, static_cast<uint32_t>(pay.v.size()), static_cast<uint32_t>(p1 - p0), 8 /* RLE_DICTIONARY */);
            o.raw(pay.v.data(), pay.v.size());
        }
        return;
    }
    // ---- PLAIN pages: close when the estimated size reaches 1 KB (:56-80) ----
    idx.clear(); idx.shrink_to_fit();
    o.v.reserve(static_cast<size_t>(rows) * (is_str ? 8 : w) + static_cast<size_t>(rows) / 16 + 1024);
    int64_t p0 = cv.a;
    size_t est = 0;
    auto flush = [&](int64_t p1) { // rows [p0, p1)
        pay.v.clear();
        if (optional) emit_def_levels(pay, cv, p0, p1);
        for (int64_t i = p0; i < p1; i++) {
            if (cv.null_at(i)) continue;
            if (is_str) {
                uint64_t s0 = c->str_off[i], s1 = c->str_off[i + 1];
                pay.le32(static_cast<uint32_t>(s1 - s0));
                pay.raw(c->chars + s0, s1 - s0);
            } else {
                uint64_t bits = load_bits(c, w, i);
                pay.raw(&bits, static_cast<size_t>(w));
            }
        }
        data_page_header(o, static_cast<uint32_t>(pay.v.size()), static_cast<uint32_t>(p1 - p0), 0 /* PLAIN */);
        o.raw(pay.v.data(), pay.v.size());
    };
    for (int64_t i = cv.a; i < cv.b; i++) {
        if (!cv.null_at(i)) est += is_str ? 4 + (c->str_off[i + 1] - c->str_off[i]) : static_cast<size_t>(w);
        if (est >= kMaxPage) { flush(i + 1); p0 = i + 1; est = 0; }
    }
    if (p0 < cv.b) flush(cv.b);
}

} // namespace

struct pqgen_job {
    std::vector<Chunk> chunks; // [rg * n_cols + col]
    std::vector<uint64_t> chunk_off;
    Bytes footer;              // FileMetaData + u32 length + "PAR1"
    uint64_t size = 0;
};

extern "C" {

const char* pqgen_last_error(void) { return g_err.c_str(); }

pqgen_job* pqgen_encode(const pqgen_col* cols, int32_t n_cols, const int64_t* rg_rows, int32_t n_rgs, int32_t threads) {
    if (!cols || n_cols <= 0 || n_rgs < 0 || (n_rgs && !rg_rows)) { g_err = "pqgen_encode: bad argument"; return nullptr; }
    for (int c = 0; c < n_cols; c++) {
        const pqgen_col& k = cols[c];
        bool ok = k.name && (k.repetition == 0 || k.repetition == 1) &&
                  ((k.type == PQG_BYTE_ARRAY && k.str_off && (k.chars || true)) || (width_of(k.type) && k.fixed) );
        int64_t total = 0;
        for (int r = 0; r < n_rgs; r++) total += rg_rows[r];
        if (total == 0) ok = k.name && (k.type == PQG_BYTE_ARRAY || width_of(k.type));
        if (!ok) { g_err = std::string("pqgen_encode: unsupported column spec: ") + (k.name ? k.name : "?"); return nullptr; }
        if (k.repetition == 0 && k.is_null) {
            for (int64_t i = 0; i < total; i++) if (k.is_null[i]) { g_err = std::string("pqgen_encode: null in REQUIRED column ") + k.name; return nullptr; }
        }
    }
    pqgen_job* job = new pqgen_job();
    const size_t n_chunks = static_cast<size_t>(n_rgs) * n_cols;
    job->chunks.resize(n_chunks);
    std::vector<int64_t> base(n_rgs + 1, 0);
    for (int r = 0; r < n_rgs; r++) base[r + 1] = base[r] + rg_rows[r];
    int nt = threads > 0 ? threads : static_cast<int>(std::thread::hardware_concurrency());
    if (nt < 1) nt = 1;
    if (static_cast<size_t>(nt) > n_chunks) nt = static_cast<int>(std::max<size_t>(n_chunks, 1));
    std::atomic<size_t> next{0};
    auto work = [&]() {
        for (;;) {
            size_t t = next.fetch_add(1);
            if (t >= n_chunks) break;
            int r = static_cast<int>(t / n_cols), c = static_cast<int>(t % n_cols);
            ColView cv{&cols[c], base[r], base[r + 1]};
            try { encode_chunk(cv, job->chunks[t]); }
            catch (const std::exception& e) { job->chunks[t].err = e.what(); }
        }
    };
    std::vector<std::thread> pool;
    for (int i = 1; i < nt; i++) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    for (size_t t = 0; t < n_chunks; t++) {
        if (!job->chunks[t].err.empty()) {
            g_err = "pqgen_encode: column " + std::string(cols[t % n_cols].name) + ": " + job->chunks[t].err;
            delete job;
            return nullptr;
        }
    }
    // layout: "PAR1", chunks in (row group, column) order, footer
    job->chunk_off.resize(n_chunks);
    uint64_t pos = 4;
    for (size_t t = 0; t < n_chunks; t++) { job->chunk_off[t] = pos; pos += job->chunks[t].bytes.v.size(); }
    // footer (close(), parquet_writer.cpp:462-581)
    Bytes& f = job->footer;
    Thrift md{f};
    md.i32(1, 2);
    md.list(2, 12, 1 + n_cols);
    { Thrift s{f}; s.str(4, "schema"); s.i32(5, n_cols); s.stop(); }
    for (int c = 0; c < n_cols; c++) {
        Thrift s{f};
        s.i32(1, cols[c].type); s.i32(3, cols[c].repetition); s.str(4, cols[c].name);
        if (cols[c].converted >= 0) s.i32(6, cols[c].converted);
        s.stop();
    }
    md.i64(3, base[n_rgs]);
    md.list(4, 12, n_rgs);
    for (int r = 0; r < n_rgs; r++) {
        Thrift rg{f};
        rg.list(1, 12, n_cols);
        int64_t rg_total = 0;
        for (int c = 0; c < n_cols; c++) {
            const Chunk& ck = job->chunks[static_cast<size_t>(r) * n_cols + c];
            const int64_t start = static_cast<int64_t>(job->chunk_off[static_cast<size_t>(r) * n_cols + c]);
            const int64_t size = static_cast<int64_t>(ck.bytes.v.size());
            rg_total += size;
            Thrift cc{f};
            cc.i64(2, start);
            cc.field(3, 12);
            Thrift m{f};
            m.i32(1, cols[c].type);
            if (ck.dict) { m.list(2, 5, 2); f.zigzag(0); f.zigzag(8); } else { m.list(2, 5, 1); f.zigzag(0); }
            m.list(3, 8, 1);
            { std::string nm = cols[c].name; f.varint(nm.size()); f.raw(nm.data(), nm.size()); }
            m.i32(4, 0);
            m.i64(5, ck.num_values);
            m.i64(6, size);
            m.i64(7, size);
            m.i64(9, ck.dict ? start + static_cast<int64_t>(ck.dict_page_bytes) : start);
            if (ck.dict) m.i64(11, start);
            m.stop();
            cc.stop();
        }
        rg.i64(2, rg_total);
        rg.i64(3, rg_rows[r]);
        rg.stop();
    }
    md.stop();
    uint32_t flen = static_cast<uint32_t>(f.v.size());
    f.le32(flen);
    f.raw("PAR1", 4);
    job->size = pos + f.v.size();
    return job;
}

uint64_t pqgen_size(const pqgen_job* job) { return job ? job->size : 0; }

int pqgen_emit(const pqgen_job* job, uint8_t* dst, uint64_t cap) {
    if (!job || !dst || cap < job->size) { g_err = "pqgen_emit: bad argument / buffer too small"; return -1; }
    std::memcpy(dst, "PAR1", 4);
    const size_t n = job->chunks.size();
    int nt = static_cast<int>(std::min<size_t>(std::max<size_t>(n, 1), std::thread::hardware_concurrency() ? std::thread::hardware_concurrency() : 1));
    std::atomic<size_t> next{0};
    auto work = [&]() {
        for (;;) {
            size_t t = next.fetch_add(1);
            if (t >= n) break;
            const auto& v = job->chunks[t].bytes.v;
            if (!v.empty()) std::memcpy(dst + job->chunk_off[t], v.data(), v.size());
        }
    };
    std::vector<std::thread> pool;
    for (int i = 1; i < nt; i++) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    std::memcpy(dst + job->size - job->footer.v.size(), job->footer.v.data(), job->footer.v.size());
    return 0;
}

int pqgen_write_file(const pqgen_job* job, const char* path) {
    if (!job || !path) { g_err = "pqgen_write_file: bad argument"; return -1; }
    FILE* fp = std::fopen(path, "wb");
    if (!fp) { g_err = std::string("pqgen_write_file: cannot open ") + path; return -1; }
    bool ok = std::fwrite("PAR1", 1, 4, fp) == 4;
    for (const auto& c : job->chunks) if (ok && !c.bytes.v.empty()) ok = std::fwrite(c.bytes.v.data(), 1, c.bytes.v.size(), fp) == c.bytes.v.size();
    if (ok) ok = std::fwrite(job->footer.v.data(), 1, job->footer.v.size(), fp) == job->footer.v.size();
    if (std::fclose(fp) != 0) ok = false;
    if (!ok) { g_err = std::string("pqgen_write_file: short write to ") + path; return -1; }
    return 0;
}

void pqgen_free(pqgen_job* job) { delete job; }

static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ULL;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
    return x ^ (x >> 31);
}

uint32_t pqgen_string_len(int32_t kind) { return kind == PQGEN_EMAILS ? 33u : (kind == PQGEN_CITY64K ? 13u : 0u); }

int pqgen_fill_strings(int32_t kind, uint64_t first_row, uint64_t rows, uint64_t seed, uint8_t* chars, uint64_t* str_off,
                       uint64_t off_base, uint8_t* is_null, uint32_t null_permille, int32_t threads) {
    const uint32_t L = pqgen_string_len(kind);
    if (!L || (!chars && rows)) { g_err = "pqgen_fill_strings: bad argument"; return -1; }
    int nt = threads > 0 ? threads : static_cast<int>(std::thread::hardware_concurrency());
    if (nt < 1) nt = 1;
    const uint64_t per = (rows + static_cast<uint64_t>(nt) - 1) / static_cast<uint64_t>(nt);
    auto work = [&](uint64_t a, uint64_t b) {
        for (uint64_t i = a; i < b; i++) {
            const uint64_t row = first_row + i;
            const uint64_t r = splitmix64(seed * 0x100000001b3ULL + row);
            uint8_t* d = chars + i * L;
            if (kind == PQGEN_EMAILS) {
                std::memcpy(d, "user000000000@mail000.example.com", 33);
                uint64_t u = 100000000ULL + (r % 900000000ULL);
                for (int k = 12; k >= 4; k--) { d[k] = static_cast<uint8_t>('0' + u % 10); u /= 10; }
                uint32_t m = static_cast<uint32_t>((r >> 40) % 1000u);
                for (int k = 20; k >= 18; k--) { d[k] = static_cast<uint8_t>('0' + m % 10); m /= 10; }
                const uint64_t blk = row / 2000;
                if (blk % 5 == 3) { if ((blk / 5) % 2 == 0) d[13] = '.'; else d[32] = '!'; }
            } else {
                std::memcpy(d, "city_000000_x", 13);
                uint32_t c = static_cast<uint32_t>(r & 0xffffu);
                for (int k = 10; k >= 5; k--) { d[k] = static_cast<uint8_t>('0' + c % 10); c /= 10; }
            }
            if (str_off) str_off[i] = off_base + i * L;
            if (is_null) is_null[i] = (splitmix64(r ^ 0x5bd1e995ULL) % 1000u) < null_permille ? 1 : 0;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; t++) {
        const uint64_t a = std::min(rows, per * static_cast<uint64_t>(t)), b = std::min(rows, a + per);
        if (a < b) pool.emplace_back(work, a, b);
    }
    work(0, std::min(rows, per));
    for (auto& th : pool) th.join();
    if (str_off) str_off[rows] = off_base + rows * L;
    return 0;
}

} // extern "C"
