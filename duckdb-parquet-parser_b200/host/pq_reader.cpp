// pq_reader.cpp -- see pq_reader.hpp.  Host logic only; every value is decoded on the GPU.
#include "pq_reader.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <mutex>
#include <sstream>
#include <thread>
#include <tuple>

namespace pqg {

// ── Value ────────────────────────────────────────────────────────────────────────────────
std::string Value::to_string() const {
    if (is_null) return "NULL";
    return std::visit([](auto&& arg) -> std::string {
        using T = std::decay_t<decltype(arg)>;
        if constexpr (std::is_same_v<T, bool>) return arg ? "true" : "false";
        else if constexpr (std::is_same_v<T, std::string>) return arg;
        else return std::to_string(arg);
    }, data);
}

// ── Device ───────────────────────────────────────────────────────────────────────────────
Device::Device(int device) : device_(device) {
    if (pqg_ctx_create(device, nullptr, &ctx_) != PQG_OK)
        throw std::runtime_error(std::string("GPU decoder unavailable: ") + pqg_last_error(nullptr));
}
Device::~Device() { if (ctx2_) pqg_ctx_destroy(ctx2_); pqg_ctx_destroy(ctx_); }
pqg_ctx* Device::ctx2() {
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (!ctx2_ && pqg_ctx_create(device_, nullptr, &ctx2_) != PQG_OK)
        throw std::runtime_error(std::string("GPU decoder unavailable: ") + pqg_last_error(nullptr));
    return ctx2_;
}

Device& Device::get(int device) {
    static std::mutex mu;
    static std::map<int, std::unique_ptr<Device>> devices;
    if (device < 0) {
        const char* e = std::getenv("PQG_DEVICE");
        device = e ? std::atoi(e) : 0;
    }
    std::lock_guard<std::mutex> lock(mu);
    auto it = devices.find(device);
    if (it == devices.end()) it = devices.emplace(device, std::unique_ptr<Device>(new Device(device))).first;
    return *it->second;
}

// ── DecodedColumn ────────────────────────────────────────────────────────────────────────
size_t DecodedColumn::chunk_of_slot(uint64_t i) const {
    size_t lo = 0, hi = chunks.size();
    while (hi - lo > 1) {
        size_t mid = (lo + hi) / 2;
        if (chunks[mid].out_row_base <= i) lo = mid; else hi = mid;
    }
    return lo;
}

std::pair<const uint8_t*, uint32_t> DecodedColumn::string_at(size_t c, uint64_t i) const {
    const pqg_chunk_desc& ck = chunks[c];
    const uint32_t* off = offsets.data() + ck.out_row_base + c;
    uint64_t s = i - ck.out_row_base;
    return {chars.data() + char_bases[c] + off[s], off[s + 1] - off[s]};
}

void DecodedColumn::append_values(std::vector<Value>& out, uint64_t first, uint64_t last) const {
    if (first >= last) return;
    out.reserve(out.size() + (last - first));
    size_t c = type == ParquetType::BYTE_ARRAY ? chunk_of_slot(first) : 0;
    for (uint64_t i = first; i < last; i++) {
        if (!slot_valid(i)) { out.push_back(Value::null()); continue; }
        const uint8_t* p = values.data() + i * width;
        switch (type) {
            case ParquetType::BOOLEAN: out.push_back(Value::from_bool(*p != 0)); break;
            case ParquetType::INT32: { int32_t v; std::memcpy(&v, p, 4); out.push_back(Value::from_i32(v)); break; }
            case ParquetType::INT64: { int64_t v; std::memcpy(&v, p, 8); out.push_back(Value::from_i64(v)); break; }
            case ParquetType::FLOAT: { float v; std::memcpy(&v, p, 4); out.push_back(Value::from_float(v)); break; }
            case ParquetType::DOUBLE: { double v; std::memcpy(&v, p, 8); out.push_back(Value::from_double(v)); break; }
            case ParquetType::INT96: {
                // the reference renders INT96 as a string (column_reader.cpp:257-264)
                int64_t low; int32_t high;
                std::memcpy(&low, p, 8); std::memcpy(&high, p + 8, 4);
                out.push_back(Value::from_string("INT96(" + std::to_string(high) + ":" + std::to_string(low) + ")"));
                break;
            }
            case ParquetType::BYTE_ARRAY: {
                while (c + 1 < chunks.size() && chunks[c + 1].out_row_base <= i) c++;
                auto [sp, len] = string_at(c, i);
                out.push_back(Value::from_string(std::string(reinterpret_cast<const char*>(sp), len)));
                break;
            }
            default: throw std::runtime_error("Unsupported type: " + std::to_string(static_cast<int>(type)));
        }
    }
}

std::vector<Value> DecodedColumn::to_values() const {
    std::vector<Value> out;
    append_values(out, 0, num_slots);
    return out;
}

// ── GPU decode of a set of table chunks ──────────────────────────────────────────────────
namespace {

struct Range { uint64_t src_off; uint64_t len; uint64_t dst_off; };

// Packs the byte ranges a table needs; rewrites nothing (tables are built against the
// packed offsets by the callers).
struct PackedImage {
    std::vector<Range> ranges;
    uint64_t size = 0;
    uint64_t add(uint64_t src_off, uint64_t len) { // returns dst_off
        uint64_t dst = (size + 15) & ~uint64_t(15);
        dst += src_off & 15; // keep the 16-byte phase of the file: aligned pages stay aligned
        ranges.push_back({src_off, len, dst});
        size = dst + len;
        return dst;
    }
};

void throw_ctx(pqg_ctx* ctx, const char* what) {
    throw std::runtime_error(std::string(what) + ": " + pqg_last_error(ctx));
}

DecodedColumn decode_packed(Device& dev, const uint8_t* src_base, const PackedImage& img, const ColumnTables& t,
                            ParquetType type) {
    DecodedColumn out;
    out.type = type;
    out.chunks = t.chunks;
    out.pages = t.pages;
    out.page_row_group = t.page_row_group;
    out.num_slots = t.total_slots;
    static const uint32_t widths[] = {1, 4, 8, 12, 4, 8, 0, 0};
    out.width = widths[static_cast<int>(type) & 7];
    if (type == ParquetType::FIXED_LEN_BYTE_ARRAY)
        throw std::runtime_error("FIXED_LEN_BYTE_ARRAY not supported without type_length");
    if (static_cast<int>(type) < 0 || static_cast<int>(type) > 7)
        throw std::runtime_error("Unsupported type: " + std::to_string(static_cast<int>(type)));
    if (t.chunks.empty()) {
        if (type == ParquetType::BYTE_ARRAY) out.char_bases.assign(1, 0);
        return out;
    }
    pqg_ctx* ctx = dev.ctx();
    pqg_buf* buf = nullptr;
    if (pqg_buf_alloc(ctx, img.size, &buf) != PQG_OK) throw_ctx(ctx, "device image");
    struct Guard {
        pqg_ctx* c; pqg_buf* b; pqg_plan* p = nullptr;
        ~Guard() { if (p) pqg_plan_destroy(c, p); if (b) pqg_buf_free(c, b); }
    } g{ctx, buf};
    for (const Range& r : img.ranges)
        if (pqg_buf_write(ctx, buf, r.dst_off, src_base + r.src_off, r.len) != PQG_OK) throw_ctx(ctx, "upload");
    const int crc = t.ext ? pqg_plan_create_ext(ctx, buf, t.chunks.data(), static_cast<uint32_t>(t.chunks.size()), t.pages.data(),
                                                static_cast<uint32_t>(t.pages.size()), t.page_ext.data(), t.chunk_ext.data(), &g.p)
                          : pqg_plan_create(ctx, buf, t.chunks.data(), static_cast<uint32_t>(t.chunks.size()), t.pages.data(),
                                            static_cast<uint32_t>(t.pages.size()), &g.p);
    if (crc != PQG_OK) {
        // keep the reference's wording for unsupported types
        throw std::runtime_error(pqg_last_error(ctx));
    }
    pqg_ctx_set_profiling(ctx, 1);
    if (pqg_plan_run(ctx, g.p) != PQG_OK) throw_ctx(ctx, "decode");
    pqg_page_error pe;
    int rc = pqg_plan_finish(ctx, g.p, &pe);
    if (rc == PQG_ERR_PAGE) throw std::runtime_error(pqg_last_error(ctx));
    if (rc != PQG_OK) throw_ctx(ctx, "decode");
    pqg_timings tm;
    pqg_plan_timings(g.p, &tm);
    out.kernel_ms = tm.total_ms;
    out.bytes_in = pqg_plan_bytes_in(g.p);
    out.bytes_out = pqg_plan_bytes_out(g.p);
    out.has_validity = pqg_plan_validity(g.p) != nullptr;
    if (out.has_validity) out.validity.resize((out.num_slots + 31) / 32);
    if (type == ParquetType::BYTE_ARRAY) {
        out.offsets.resize(out.num_slots + t.chunks.size());
        out.chars.resize(pqg_plan_chars_size(g.p));
        out.char_bases.resize(t.chunks.size() + 1);
        pqg_plan_char_bases(ctx, g.p, out.char_bases.data(), static_cast<uint32_t>(out.char_bases.size()));
    } else {
        out.values.resize(out.num_slots * out.width);
    }
    if (pqg_plan_download(ctx, g.p, out.values.empty() ? nullptr : out.values.data(),
                          out.validity.empty() ? nullptr : out.validity.data(),
                          out.offsets.empty() ? nullptr : out.offsets.data(),
                          out.chars.empty() ? nullptr : out.chars.data()) != PQG_OK) throw_ctx(ctx, "download");
    if (pqg_ctx_sync(ctx) != PQG_OK) throw_ctx(ctx, "sync");
    return out;
}

} // namespace

DecodedColumn decode_column(Device& dev, const uint8_t* image, size_t image_size, const ColumnTables& t) {
    PackedImage img;
    img.ranges.push_back({0, image_size, 0});
    img.size = image_size;
    ParquetType type = t.chunks.empty() ? ParquetType::INT32 : static_cast<ParquetType>(t.chunks[0].phys_type);
    return decode_packed(dev, image, img, t, type);
}

// ── ColumnReader ─────────────────────────────────────────────────────────────────────────
ColumnReader::ColumnReader(ReadRangeFunc read_range, const ColumnChunk& chunk, ParquetType type,
                           int16_t max_def_level, int16_t max_rep_level)
    : read_range_(std::move(read_range)), type_(type), max_def_level_(max_def_level), max_rep_level_(max_rep_level) {
    if (!chunk.meta_data.has_value()) throw std::runtime_error("ColumnChunk has no metadata");
    meta_ = &chunk.meta_data.value();
    if (meta_->codec != CompressionCodec::UNCOMPRESSED)
        throw std::runtime_error("Only uncompressed parquet files are supported");
}

// Fetches the chunk's bytes through the caller's read_range callback (one large read
// instead of the reference's two reads per page) and walks its page headers.
ColumnReader::Loaded ColumnReader::load() {
    Loaded L;
    L.file_off = chunk_start_offset(*meta_);
    if (meta_->num_values <= 0) return L;
    size_t want = meta_->total_compressed_size > 0 ? static_cast<size_t>(meta_->total_compressed_size) : size_t(1) << 20;
    L.bytes = read_range_(static_cast<size_t>(L.file_off), want);
    for (int attempt = 0;; attempt++) {
        try {
            L.pages.clear();
            walk_chunk_pages(L.bytes.data(), L.file_off, L.bytes.size(), L.file_off, meta_->num_values, L.pages);
            if (!L.pages.empty()) {
                const PageRecord& last = L.pages.back();
                uint64_t end = last.payload_off + last.payload_size;
                if (end > L.file_off + L.bytes.size()) throw FormatError("short");
            }
            break;
        } catch (const FormatError&) {
            // metadata understated the chunk size: fetch more and retry
            if (attempt >= 6) throw;
            auto more = read_range_(static_cast<size_t>(L.file_off + L.bytes.size()), L.bytes.size() * 3 + 4096);
            if (more.empty()) throw;
            L.bytes.insert(L.bytes.end(), more.begin(), more.end());
        }
    }
    return L;
}

DecodedColumn ColumnReader::read_columnar() {
    Loaded L = load();
    ColumnTables t;
    append_chunk_tables(t, L.pages, L.file_off, type_, max_def_level_, max_rep_level_, 0, 0);
    DecodedColumn d = decode_column(Device::get(), L.bytes.data(), L.bytes.size(), t);
    d.type = type_;
    return d;
}

std::vector<Value> ColumnReader::read_all() { return read_columnar().to_values(); }

std::vector<PageResult> ColumnReader::read_pages() {
    Loaded L = load();
    ColumnTables t;
    append_chunk_tables(t, L.pages, L.file_off, type_, max_def_level_, max_rep_level_, 0, 0);
    DecodedColumn d = decode_column(Device::get(), L.bytes.data(), L.bytes.size(), t);
    d.type = type_;
    std::vector<PageResult> out;
    int page_num = 0;
    size_t data_page = 0;
    for (const PageRecord& r : L.pages) {
        if (r.type == PageType::DICTIONARY_PAGE) {
            out.push_back({page_num, PageType::DICTIONARY_PAGE, r.num_values, {}});
        } else if (r.type == PageType::DATA_PAGE) {
            const pqg_page_desc& pd = t.pages[data_page++];
            PageResult pr{page_num, PageType::DATA_PAGE, r.num_values, {}};
            d.append_values(pr.values, pd.out_row_base, pd.out_row_base + pd.num_values);
            out.push_back(std::move(pr));
        }
        page_num++; // dictionary and unknown pages count too (reference column_reader.cpp:104,115,122)
    }
    return out;
}

// ── ParquetReader ────────────────────────────────────────────────────────────────────────
ParquetReader::ParquetReader() = default;
ParquetReader::~ParquetReader() {
    close_file();
    for (int i = 0; i < 2; i++) if (vscratch_[i]) { pqg_host_free(vscratch_[i]); vscratch_[i] = nullptr; vscratch_words_[i] = 0; }
}

void ParquetReader::close_file() {
    if (mapped_ && data_) munmap(const_cast<uint8_t*>(data_), file_size_);
    data_ = nullptr;
    mapped_ = false;
}

bool ParquetReader::open(const std::string& filename) {
    close_file();
    int fd = ::open(filename.c_str(), O_RDONLY);
    if (fd < 0) {
        open_error_ = "Error: cannot open file " + filename;
        std::cerr << open_error_ << std::endl;
        return false;
    }
    struct stat st;
    fstat(fd, &st);
    file_size_ = static_cast<size_t>(st.st_size);
    if (file_size_ < 12) {
        ::close(fd);
        open_error_ = "Error: file too small to be a Parquet file";
        std::cerr << open_error_ << std::endl;
        return false;
    }
    void* p = mmap(nullptr, file_size_, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (p == MAP_FAILED) {
        open_error_ = "Error: cannot map file " + filename;
        std::cerr << open_error_ << std::endl;
        return false;
    }
    data_ = static_cast<const uint8_t*>(p);
    mapped_ = true;
    return finish_open();
}

bool ParquetReader::open_memory(const uint8_t* data, size_t size) {
    close_file();
    data_ = data;
    file_size_ = size;
    mapped_ = false;
    if (size < 12) {
        open_error_ = "Error: file too small to be a Parquet file";
        std::cerr << open_error_ << std::endl;
        return false;
    }
    return finish_open();
}

bool ParquetReader::finish_open() {
    auto err = [&](const std::string& m) { open_error_ = m; std::cerr << m << std::endl; return false; };
    if (std::memcmp(data_, "PAR1", 4) != 0) return err("Error: missing PAR1 magic at start");
    if (std::memcmp(data_ + file_size_ - 4, "PAR1", 4) != 0) return err("Error: missing PAR1 magic at end");
    uint32_t footer_length;
    std::memcpy(&footer_length, data_ + file_size_ - 8, 4);
    if (static_cast<size_t>(footer_length) + 8 > file_size_) return err("Error: invalid footer length");
    try {
        metadata_ = parse_file_metadata(data_ + file_size_ - 8 - footer_length, footer_length);
    } catch (const std::exception& e) { return err(std::string("Error: bad footer: ") + e.what()); }
    columns_ = build_column_info(metadata_);
    column_name_to_idx_.clear();
    for (size_t i = 0; i < columns_.size(); i++) column_name_to_idx_[columns_[i].name] = i;

    // page scan: one task per column chunk, all host threads
    auto t0 = std::chrono::steady_clock::now();
    const size_t nrg = metadata_.row_groups.size();
    chunk_pages_.assign(nrg, {});
    chunk_first_page_.assign(nrg, {});
    std::vector<std::pair<size_t, size_t>> tasks;
    for (size_t rg = 0; rg < nrg; rg++) {
        chunk_pages_[rg].resize(metadata_.row_groups[rg].columns.size());
        chunk_first_page_[rg].assign(metadata_.row_groups[rg].columns.size(), 0);
        for (size_t c = 0; c < metadata_.row_groups[rg].columns.size(); c++) tasks.push_back({rg, c});
    }
    std::atomic<size_t> next{0};
    std::mutex emu;
    std::string first_error;
    auto worker = [&]() {
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= tasks.size()) break;
            auto [rg, c] = tasks[i];
            const ColumnChunk& chunk = metadata_.row_groups[rg].columns[c];
            if (!chunk.meta_data) continue;
            try {
                walk_chunk_pages(data_, 0, file_size_, chunk_start_offset(*chunk.meta_data),
                                 chunk.meta_data->num_values, chunk_pages_[rg][c]);
            } catch (const std::exception& e) {
                std::lock_guard<std::mutex> lock(emu);
                if (first_error.empty())
                    first_error = "row group " + std::to_string(rg) + " column " + std::to_string(c) + ": " + e.what();
            }
        }
    };
    unsigned nt = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), 32));
    nt = static_cast<unsigned>(std::min<size_t>(nt, std::max<size_t>(tasks.size(), 1)));
    std::vector<std::thread> pool;
    for (unsigned i = 1; i < nt; i++) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
    if (!first_error.empty()) return err("Error: page scan failed: " + first_error);

    // global page ids in (row group, column, page) order; DATA_PAGE / DATA_PAGE_V2 only
    // (reference build_page_index, parquet_reader.cpp:588-599)
    page_index_.clear();
    size_t total = 0;
    for (auto& rg : chunk_pages_) for (auto& c : rg) for (auto& r : c) total += r.counted;
    page_index_.reserve(total);
    for (size_t rg = 0; rg < nrg; rg++)
        for (size_t c = 0; c < chunk_pages_[rg].size(); c++) {
            chunk_first_page_[rg][c] = page_index_.size();
            for (const PageRecord& r : chunk_pages_[rg][c])
                if (r.counted) page_index_.push_back({static_cast<size_t>(r.payload_off), r.payload_size, rg, c});
        }
    scan_seconds_ = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return true;
}

size_t ParquetReader::num_columns() const { return columns_.size(); }
int64_t ParquetReader::num_rows() const { return metadata_.num_rows; }
size_t ParquetReader::num_row_groups() const { return metadata_.row_groups.size(); }

std::vector<std::string> ParquetReader::column_names() const {
    std::vector<std::string> names;
    for (const auto& c : columns_) names.push_back(c.name);
    return names;
}

const ColumnInfo& ParquetReader::column(size_t col_idx) const {
    if (col_idx >= columns_.size()) throw std::runtime_error("Column index " + std::to_string(col_idx) + " out of range");
    return columns_[col_idx];
}
const ColumnInfo& ParquetReader::column(const std::string& name) const {
    int idx = find_column(name);
    if (idx < 0) throw std::runtime_error("Column not found: " + name);
    return columns_[static_cast<size_t>(idx)];
}
int ParquetReader::find_column(const std::string& name) const {
    auto it = column_name_to_idx_.find(name);
    return it == column_name_to_idx_.end() ? -1 : static_cast<int>(it->second);
}

std::string ParquetReader::schema_string() const {
    std::ostringstream ss;
    ss << "Schema:\n";
    for (size_t i = 0; i < columns_.size(); i++) {
        const auto& col = columns_[i];
        ss << "  " << i << ": " << col.name << " (" << col.type_name();
        if (col.converted_type && *col.converted_type != ConvertedType::NONE) ss << ", converted=" << col.converted_type_string();
        if (col.repetition) {
            static const char* reps[] = {", REQUIRED", ", OPTIONAL", ", REPEATED"};
            int r = static_cast<int>(*col.repetition);
            if (r >= 0 && r < 3) ss << reps[r];
        }
        ss << ")\n";
    }
    ss << "Rows: " << metadata_.num_rows << "\n";
    ss << "Row groups: " << metadata_.row_groups.size() << "\n";
    return ss.str();
}

ColumnTables ParquetReader::column_tables(int col_idx, int row_group_idx) const {
    size_t rg0 = row_group_idx < 0 ? 0 : static_cast<size_t>(row_group_idx);
    size_t rg1 = row_group_idx < 0 ? metadata_.row_groups.size() : rg0 + 1;
    return column_tables_range(col_idx, rg0, rg1);
}

ColumnTables ParquetReader::column_tables_range(int col_idx, size_t rg0, size_t rg1) const {
    // tables against FILE offsets (image byte 0 = file byte 0)
    ColumnTables t;
    const ColumnInfo& ci = columns_.at(static_cast<size_t>(col_idx));
    if (rg1 > metadata_.row_groups.size() || rg0 > rg1) throw std::runtime_error("Invalid row group index");
    for (size_t rg = rg0; rg < rg1; rg++) {
        const ColumnChunk& chunk = metadata_.row_groups[rg].columns.at(static_cast<size_t>(ci.column_index));
        if (!chunk.meta_data) throw std::runtime_error("ColumnChunk has no metadata");
        const CompressionCodec codec = chunk.meta_data->codec;
        if (codec != CompressionCodec::UNCOMPRESSED && !(extensions_ && codec == CompressionCodec::SNAPPY))
            throw std::runtime_error(extensions_ ? std::string("Only uncompressed and SNAPPY-compressed parquet files are supported (codec ") +
                                                       compression_name(codec) + ")"
                                                 : std::string("Only uncompressed parquet files are supported"));
        append_chunk_tables(t, chunk_pages_[rg][static_cast<size_t>(ci.column_index)], 0, ci.type,
                            ci.max_def_level, ci.max_rep_level, static_cast<uint32_t>(rg), static_cast<uint32_t>(col_idx), extensions_, codec);
    }
    return t;
}

// this column's byte ranges, packed (one range per table chunk); rebases the tables onto the packed image
static PackedImage pack_column(ColumnTables& t, uint64_t file_size) {
    PackedImage img;
    for (pqg_chunk_desc& c : t.chunks) {
        uint64_t lo = c.has_dict ? c.dict_off : UINT64_MAX, hi = c.has_dict ? c.dict_off + c.dict_size : 0;
        for (uint32_t q = c.first_page; q < c.first_page + c.n_pages; q++) {
            lo = std::min(lo, t.pages[q].payload_off);
            hi = std::max(hi, t.pages[q].payload_off + t.pages[q].payload_size);
        }
        if (lo == UINT64_MAX) { lo = 0; hi = 0; }
        if (hi > file_size) throw std::runtime_error("ByteBuffer: read beyond end (pos=" + std::to_string(lo) + " need=" +
                                                     std::to_string(hi - lo) + " size=" + std::to_string(file_size) + ")");
        uint64_t dst = img.add(lo, hi - lo);
        int64_t delta = static_cast<int64_t>(dst) - static_cast<int64_t>(lo);
        if (c.has_dict) c.dict_off = static_cast<uint64_t>(static_cast<int64_t>(c.dict_off) + delta);
        for (uint32_t q = c.first_page; q < c.first_page + c.n_pages; q++)
            t.pages[q].payload_off = static_cast<uint64_t>(static_cast<int64_t>(t.pages[q].payload_off) + delta);
    }
    return img;
}

DecodedColumn ParquetReader::read_column_columnar(int col_idx, int row_group_idx) {
    if (row_group_idx >= static_cast<int>(metadata_.row_groups.size())) throw std::runtime_error("Invalid row group index");
    if (col_idx < 0 || col_idx >= static_cast<int>(columns_.size())) throw std::runtime_error("Invalid column index");
    const ColumnInfo& ci = columns_[static_cast<size_t>(col_idx)];
    ColumnTables t = column_tables(col_idx, row_group_idx);
    PackedImage img = pack_column(t, file_size_);
    return decode_packed(Device::get(device_), data_, img, t, ci.type);
}

// ── pipelined reads with cached plans ───────────────────────────────────────────────────
struct CachedPlan {
    ColumnTables t;
    PackedImage img;
    std::vector<pqg_h2d_range> ranges;
    pqg_ctx* ctx = nullptr;
    pqg_buf* buf = nullptr;
    pqg_plan* plan = nullptr;
    ParquetType type = ParquetType::INT32;
    uint32_t width = 0;
    uint64_t h2d_bytes = 0;
    ~CachedPlan() {
        if (plan) pqg_plan_destroy(ctx, plan);
        if (buf) pqg_buf_free(ctx, buf);
    }
};

CachedPlan& ParquetReader::cached_plan(int col_idx, int row_group_idx) {
    if (row_group_idx >= static_cast<int>(metadata_.row_groups.size())) throw std::runtime_error("Invalid row group index");
    size_t rg0 = row_group_idx < 0 ? 0 : static_cast<size_t>(row_group_idx);
    size_t rg1 = row_group_idx < 0 ? metadata_.row_groups.size() : rg0 + 1;
    return cached_plan_range(col_idx, rg0, rg1);
}

CachedPlan& ParquetReader::cached_plan_range(int col_idx, size_t rg0, size_t rg1, bool dict_indices, int lane) {
    if (rg1 > metadata_.row_groups.size() || rg0 > rg1) throw std::runtime_error("Invalid row group index");
    if (col_idx < 0 || col_idx >= static_cast<int>(columns_.size())) throw std::runtime_error("Invalid column index");
    auto key = std::make_tuple(col_idx, rg0, rg1, (dict_indices ? 1 : 0) + 2 * lane);
    auto it = plans_.find(key);
    if (it != plans_.end()) return *it->second;
    auto cp = std::make_unique<CachedPlan>();
    cp->type = columns_[static_cast<size_t>(col_idx)].type;
    static const uint32_t widths[] = {1, 4, 8, 12, 4, 8, 0, 0};
    cp->width = widths[static_cast<int>(cp->type) & 7];
    cp->t = column_tables_range(col_idx, rg0, rg1);
    if (cp->t.ext && dict_indices) throw std::runtime_error("dictionary-form reads do not cover compressed / DATA_PAGE_V2 chunks (extensions)");
    cp->img = pack_column(cp->t, file_size_);
    cp->ctx = lane ? Device::get(device_).ctx2() : Device::get(device_).ctx();
    if (cp->t.chunks.empty()) { auto& ref = *cp; plans_[key] = std::move(cp); return ref; }
    if (pqg_buf_alloc(cp->ctx, cp->img.size, &cp->buf) != PQG_OK) throw_ctx(cp->ctx, "device image");
    auto create = dict_indices ? pqg_plan_create_dict_indices : pqg_plan_create;
    const int crc = cp->t.ext ? pqg_plan_create_ext(cp->ctx, cp->buf, cp->t.chunks.data(), static_cast<uint32_t>(cp->t.chunks.size()), cp->t.pages.data(),
                                                    static_cast<uint32_t>(cp->t.pages.size()), cp->t.page_ext.data(), cp->t.chunk_ext.data(), &cp->plan)
                              : create(cp->ctx, cp->buf, cp->t.chunks.data(), static_cast<uint32_t>(cp->t.chunks.size()), cp->t.pages.data(),
                                       static_cast<uint32_t>(cp->t.pages.size()), &cp->plan);
    if (crc != PQG_OK) throw std::runtime_error(pqg_last_error(cp->ctx));
    if (dict_indices) cp->width = 4;
    for (size_t c = 0; c < cp->img.ranges.size(); c++) {
        const Range& r = cp->img.ranges[c];
        cp->ranges.push_back(pqg_h2d_range{data_ + r.src_off, r.dst_off, r.len, static_cast<uint32_t>(c), 0});
        cp->h2d_bytes += r.len;
    }
    auto& ref = *cp;
    plans_[key] = std::move(cp);
    return ref;
}

void ParquetReader::release_plans() { plans_.clear(); }

ParquetReader::DevicePlan ParquetReader::device_plan(int col_idx, int row_group_idx, bool upload) {
    if (row_group_idx >= static_cast<int>(metadata_.row_groups.size())) throw std::runtime_error("Invalid row group index");
    size_t rg0 = row_group_idx < 0 ? 0 : static_cast<size_t>(row_group_idx);
    size_t rg1 = row_group_idx < 0 ? metadata_.row_groups.size() : rg0 + 1;
    return device_plan_range(col_idx, rg0, rg1, upload);
}

std::vector<int32_t> ParquetReader::shard_row_groups(int col_idx, int n_shards) const {
    if (n_shards < 1) throw std::runtime_error("shard_row_groups: n_shards must be >= 1");
    if (col_idx >= static_cast<int>(columns_.size())) throw std::runtime_error("Invalid column index");
    const size_t nrg = metadata_.row_groups.size();
    std::vector<uint64_t> bytes(nrg, 0);
    uint64_t total = 0;
    for (size_t rg = 0; rg < nrg; rg++) {
        const auto& cols = metadata_.row_groups[rg].columns;
        for (size_t c = 0; c < cols.size(); c++) {
            if (col_idx >= 0 && static_cast<int>(c) != columns_[static_cast<size_t>(col_idx)].column_index) continue;
            if (cols[c].meta_data) bytes[rg] += static_cast<uint64_t>(std::max<int64_t>(cols[c].meta_data->total_compressed_size, 0));
        }
        total += bytes[rg];
    }
    // boundary s = first row group whose byte prefix reaches s/n of the total (contiguous, ordered)
    std::vector<int32_t> out(static_cast<size_t>(n_shards) + 1, static_cast<int32_t>(nrg));
    out[0] = 0;
    uint64_t acc = 0;
    size_t rg = 0;
    for (int s = 1; s < n_shards; s++) {
        const double target = static_cast<double>(total) * s / n_shards;
        while (rg < nrg && static_cast<double>(acc) + bytes[rg] * 0.5 < target) acc += bytes[rg++];
        out[static_cast<size_t>(s)] = static_cast<int32_t>(rg);
    }
    return out;
}

ParquetReader::DevicePlan ParquetReader::device_plan_range(int col_idx, size_t rg0, size_t rg1, bool upload) {
    CachedPlan& cp = cached_plan_range(col_idx, rg0, rg1);
    if (upload && cp.buf) {
        for (const pqg_h2d_range& r : cp.ranges)
            if (pqg_buf_write(cp.ctx, cp.buf, r.image_off, r.host, r.len) != PQG_OK) throw_ctx(cp.ctx, "upload");
    }
    return DevicePlan{cp.ctx, cp.buf, cp.plan, static_cast<uint32_t>(cp.t.pages.size()), static_cast<uint32_t>(cp.t.chunks.size())};
}

void ParquetReader::read_columns_into(const int* col_idx, int n_cols, int row_group_idx, const ColumnDst* dsts, ColumnReadStats* stats) {
    if (row_group_idx >= static_cast<int>(metadata_.row_groups.size())) throw std::runtime_error("Invalid row group index");
    size_t rg0 = row_group_idx < 0 ? 0 : static_cast<size_t>(row_group_idx);
    size_t rg1 = row_group_idx < 0 ? metadata_.row_groups.size() : rg0 + 1;
    read_columns_into_range(col_idx, n_cols, rg0, rg1, dsts, stats);
}

void ParquetReader::run_pipelined(const std::vector<CachedPlan*>& cps, const ColumnDst* dsts, ColumnReadStats* stats) {
    // enqueue every column (asynchronous), then wait column by column
    std::vector<int> ext_no_validity; // extension columns enqueued without a validity bitmap (see below)
    for (int i = 0; i < static_cast<int>(cps.size()); i++) {
        CachedPlan& cp = *cps[static_cast<size_t>(i)];
        if (!cp.plan) continue;
        const bool has_validity = pqg_plan_validity(cp.plan) != nullptr;
        uint32_t* vdst = nullptr;
        if (has_validity && dsts[i].validity) {
            if (dsts[i].validity_cap < (cp.t.total_slots + 31) / 32) throw std::runtime_error("read_columns_into: validity buffer too small");
            vdst = dsts[i].validity;
        }
        if (cp.t.ext) {
            // compressed / DATA_PAGE_V2 chunks (extensions): the page rewrite needs the whole image, so this column runs as
            // upload -> rewrite + decode -> download on the context's stream (asynchronous; other columns proceed meanwhile)
            for (const pqg_h2d_range& r : cp.ranges)
                if (pqg_buf_write(cp.ctx, cp.buf, r.image_off, r.host, r.len) != PQG_OK) throw_ctx(cp.ctx, "upload");
            if (pqg_plan_run(cp.ctx, cp.plan) != PQG_OK) throw_ctx(cp.ctx, "decode");
            if (pqg_plan_download(cp.ctx, cp.plan, dsts[i].values, vdst, nullptr, nullptr) != PQG_OK) throw_ctx(cp.ctx, "download");
            ext_no_validity.push_back(has_validity ? -1 : i);
            continue;
        }
        if (pqg_plan_run_pipelined(cp.ctx, cp.plan, cp.buf, cp.ranges.data(), static_cast<uint32_t>(cp.ranges.size()),
                                   dsts[i].values, vdst) != PQG_OK) throw_ctx(cp.ctx, "decode");
    }
    std::string first_error;
    for (int i = 0; i < static_cast<int>(cps.size()); i++) {
        CachedPlan& cp = *cps[static_cast<size_t>(i)];
        ColumnReadStats st;
        st.num_slots = cp.t.total_slots;
        st.width = cp.width;
        if (cp.plan) {
            pqg_page_error pe;
            int rc = pqg_plan_finish(cp.ctx, cp.plan, &pe);
            if (rc != PQG_OK && first_error.empty()) first_error = pqg_last_error(cp.ctx);
            st.has_validity = pqg_plan_validity(cp.plan) != nullptr;
            // (pqg_plan_finish added a validity bitmap and decoded again: the values already copied out are stale)
            if (st.has_validity && std::find(ext_no_validity.begin(), ext_no_validity.end(), i) != ext_no_validity.end() && first_error.empty())
                first_error = "out-of-range dictionary index in a REQUIRED column chunk (a null in the reference): decode this column through read_column";
            st.bytes_in = pqg_plan_bytes_in(cp.plan);
            st.bytes_out = pqg_plan_bytes_out(cp.plan);
            st.h2d_bytes = cp.h2d_bytes;
            st.d2h_bytes = cp.t.total_slots * cp.width + (st.has_validity && dsts[i].validity ? ((cp.t.total_slots + 31) / 32) * 4 : 0);
        }
        if (stats) stats[i] = st;
    }
    if (!first_error.empty()) throw std::runtime_error(first_error);
}

void ParquetReader::read_dictionary_indices_into(int col_idx, size_t rg0, size_t rg1, const ColumnDst& dst, ColumnReadStats* stats) {
    if (col_idx < 0 || col_idx >= static_cast<int>(columns_.size())) throw std::runtime_error("Invalid column index");
    const ColumnInfo& ci = columns_[static_cast<size_t>(col_idx)];
    if (ci.type != ParquetType::BYTE_ARRAY)
        throw std::runtime_error("Column '" + ci.name + "' is not BYTE_ARRAY (type: " + parquet_type_name(ci.type) + ")");
    CachedPlan& cp = cached_plan_range(col_idx, rg0, rg1, true);
    const uint64_t need = cp.t.total_slots * 4;
    if (dst.values_cap < need || (need && !dst.values)) throw std::runtime_error("read_dictionary_indices_into: indices buffer too small");
    std::vector<CachedPlan*> cps{&cp};
    run_pipelined(cps, &dst, stats);
}

void ParquetReader::read_strings_into_range(int col_idx, size_t rg0, size_t rg1, const StringsDst& dst, StringsReadStats* stats) {
    if (col_idx < 0 || col_idx >= static_cast<int>(columns_.size())) throw std::runtime_error("Invalid column index");
    if (rg1 > metadata_.row_groups.size() || rg0 > rg1) throw std::runtime_error("Invalid row group index");
    const ColumnInfo& ci = columns_[static_cast<size_t>(col_idx)];
    if (ci.type != ParquetType::BYTE_ARRAY)
        throw std::runtime_error("Column '" + ci.name + "' is not BYTE_ARRAY (type: " + parquet_type_name(ci.type) + ")");
    const size_t K = rg1 - rg0;
    std::vector<CachedPlan*> cps(K);
    std::vector<uint64_t> row_base(K + 1, 0), chunk_base(K + 1, 0);
    for (size_t k = 0; k < K; k++) {
        cps[k] = &cached_plan_range(col_idx, rg0 + k, rg0 + k + 1, false, static_cast<int>(k & 1));
        row_base[k + 1] = row_base[k] + cps[k]->t.total_slots;
        chunk_base[k + 1] = chunk_base[k] + cps[k]->t.chunks.size();
    }
    const uint64_t slots = row_base[K], n_chunks = chunk_base[K], vwords = (slots + 31) / 32;
    if (dst.offsets_cap < slots + n_chunks || ((slots + n_chunks) && !dst.offsets)) throw std::runtime_error("read_strings_into: offsets buffer too small");
    if (dst.char_bases_cap < n_chunks + 1 || !dst.char_bases) throw std::runtime_error("read_strings_into: char_bases buffer too small");
    if (dst.validity && dst.validity_cap < vwords) throw std::runtime_error("read_strings_into: validity buffer too small");
    if (dst.validity) std::memset(dst.validity, 0, vwords * 4);
    StringsReadStats st;
    st.num_slots = slots; st.n_chunks = n_chunks;
    struct Flight { bool on = false; size_t k = 0; uint64_t char_base = 0, chars_size = 0; bool direct = false; } fl[2];
    std::string first_error;
    auto land = [&](int lane) { // wait for the lane's row group, merge what needs the host
        Flight& f = fl[lane];
        if (!f.on) return;
        f.on = false;
        CachedPlan& cp = *cps[f.k];
        pqg_page_error pe;
        if (pqg_plan_finish(cp.ctx, cp.plan, &pe) != PQG_OK) { if (first_error.empty()) first_error = pqg_last_error(cp.ctx); return; }
        if (pqg_plan_chars_size(cp.plan) != f.chars_size) { // (a plan that met other data than in its previous run: fetch again)
            if (first_error.empty()) first_error = "read_strings_into: the string bytes of a row group changed between two reads of the same file";
            return;
        }
        const uint64_t nc = cp.t.chunks.size(), n = cp.t.total_slots;
        std::vector<uint64_t> local(nc + 1);
        pqg_plan_char_bases(cp.ctx, cp.plan, local.data(), static_cast<uint32_t>(nc + 1));
        for (uint64_t c = 0; c < nc; c++) dst.char_bases[chunk_base[f.k] + c] = f.char_base + local[c];
        const bool has_v = pqg_plan_validity(cp.plan) != nullptr;
        if (has_v) st.has_validity = 1;
        if (dst.validity && n && f.direct) {
            // (the validity words went straight to their place: the row group starts and ends on word boundaries)
            if (!has_v) std::memset(dst.validity + (row_base[f.k] >> 5), 0xFF, ((n + 31) / 32) * 4);
            if (n & 31) dst.validity[(row_base[f.k] >> 5) + n / 32] &= (1u << (n & 31)) - 1u;
        } else if (dst.validity && n) { // the row group's bits at their place in the range's bitmap
            const uint64_t b0 = row_base[f.k];
            const uint32_t sh = static_cast<uint32_t>(b0 & 31);
            uint32_t* out = dst.validity + (b0 >> 5);
            const uint64_t nw = (n + 31) / 32;
            for (uint64_t w = 0; w < nw; w++) {
                uint32_t bits = has_v ? vscratch_[lane][w] : 0xffffffffu;
                if (w == nw - 1 && (n & 31)) bits &= (1u << (n & 31)) - 1u;
                out[w] |= bits << sh;
                if (sh && (bits >> (32 - sh))) out[w + 1] |= bits >> (32 - sh);
            }
        }
        st.bytes_in += pqg_plan_bytes_in(cp.plan);
        st.bytes_out += pqg_plan_bytes_out(cp.plan);
        st.h2d_bytes += cp.h2d_bytes;
        st.d2h_bytes += f.chars_size + 4 * (n + nc) + (has_v && dst.validity ? ((n + 31) / 32) * 4 : 0);
    };
    uint64_t char_base = 0;
    for (size_t k = 0; k < K && first_error.empty(); k++) {
        const int lane = static_cast<int>(k & 1);
        land(lane);
        if (!first_error.empty()) break;
        CachedPlan& cp = *cps[k];
        if (!cp.plan) continue; // a row group without pages of this column
        for (const pqg_h2d_range& r : cp.ranges)
            if (pqg_buf_write(cp.ctx, cp.buf, r.image_off, r.host, r.len) != PQG_OK) throw_ctx(cp.ctx, "upload");
        if (pqg_plan_run(cp.ctx, cp.plan) != PQG_OK) { first_error = pqg_last_error(cp.ctx); break; }
        const uint64_t csize = pqg_plan_chars_size(cp.plan);
        if (char_base + csize > dst.chars_cap || (csize && !dst.chars)) {
            first_error = "read_strings_into: chars buffer too small (" + std::to_string(char_base + csize) + " bytes needed up to row group " +
                          std::to_string(rg0 + k) + ", " + std::to_string(dst.chars_cap) + " given)";
            pqg_page_error pe;
            pqg_plan_finish(cp.ctx, cp.plan, &pe);
            break;
        }
        const uint64_t nw = (cp.t.total_slots + 31) / 32;
        uint32_t* vdst = nullptr;
        // row groups that start on a word boundary (and end on one, or end the range) download their validity words in place
        const bool direct = (row_base[k] & 31) == 0 && ((cp.t.total_slots & 31) == 0 || k + 1 == K);
        if (dst.validity && pqg_plan_validity(cp.plan) && direct) vdst = dst.validity + (row_base[k] >> 5);
        else if (dst.validity && pqg_plan_validity(cp.plan)) {
            if (vscratch_words_[lane] < nw + 1) {
                if (vscratch_[lane]) pqg_host_free(vscratch_[lane]);
                vscratch_[lane] = static_cast<uint32_t*>(pqg_host_alloc((nw + 1) * 4));
                vscratch_words_[lane] = vscratch_[lane] ? nw + 1 : 0;
                if (!vscratch_[lane]) throw std::runtime_error("read_strings_into: pinned staging allocation failed");
            }
            vdst = vscratch_[lane];
        }
        if (pqg_plan_download(cp.ctx, cp.plan, nullptr, vdst, dst.offsets + row_base[k] + chunk_base[k], dst.chars + char_base) != PQG_OK)
            throw_ctx(cp.ctx, "download");
        fl[lane] = Flight{true, k, char_base, csize, direct};
        char_base += csize;
    }
    land(0);
    land(1);
    if (!first_error.empty()) throw std::runtime_error(first_error);
    dst.char_bases[n_chunks] = char_base;
    st.chars_size = char_base;
    if (stats) *stats = st;
}

void ParquetReader::chunk_dictionary(int col_idx, size_t rg, std::vector<uint32_t>& offsets, std::vector<uint8_t>& chars) const {
    offsets.assign(1, 0);
    chars.clear();
    if (col_idx < 0 || col_idx >= static_cast<int>(columns_.size())) throw std::runtime_error("Invalid column index");
    if (rg >= metadata_.row_groups.size()) throw std::runtime_error("Invalid row group index");
    const ColumnInfo& ci = columns_[static_cast<size_t>(col_idx)];
    if (ci.type != ParquetType::BYTE_ARRAY)
        throw std::runtime_error("Column '" + ci.name + "' is not BYTE_ARRAY (type: " + parquet_type_name(ci.type) + ")");
    for (const PageRecord& r : chunk_pages_[rg][static_cast<size_t>(ci.column_index)]) {
        if (r.type != PageType::DICTIONARY_PAGE) continue;
        // a later dictionary page replaces the earlier one (reference column_reader.cpp:48-54); chunks
        // that switch dictionaries are split by the descriptor tables -- report the last one here
        offsets.assign(1, 0);
        chars.clear();
        uint64_t pos = r.payload_off, end = r.payload_off + r.payload_size;
        for (int32_t k = 0; k < r.num_values; k++) {
            if (pos + 4 > end) throw std::runtime_error("ByteBuffer: read beyond end (pos=" + std::to_string(pos - r.payload_off) + " need=4 size=" + std::to_string(r.payload_size) + ")");
            uint32_t len;
            std::memcpy(&len, data_ + pos, 4);
            if (pos + 4 + len > end) throw std::runtime_error("ByteBuffer: read beyond end (pos=" + std::to_string(pos + 4 - r.payload_off) + " need=" + std::to_string(len) + " size=" + std::to_string(r.payload_size) + ")");
            chars.insert(chars.end(), data_ + pos + 4, data_ + pos + 4 + len);
            offsets.push_back(static_cast<uint32_t>(chars.size()));
            pos += 4 + len;
        }
    }
}

void ParquetReader::read_columns_into_range(const int* col_idx, int n_cols, size_t rg0, size_t rg1, const ColumnDst* dsts, ColumnReadStats* stats) {
    if (n_cols < 0 || (n_cols && (!col_idx || !dsts))) throw std::runtime_error("read_columns_into: bad argument");
    std::vector<CachedPlan*> cps;
    for (int i = 0; i < n_cols; i++) {
        CachedPlan& cp = cached_plan_range(col_idx[i], rg0, rg1);
        if (cp.type == ParquetType::BYTE_ARRAY || cp.width == 0)
            throw std::runtime_error("read_columns_into: fixed-width columns only (use read_column_columnar for BYTE_ARRAY)");
        const uint64_t need = cp.t.total_slots * cp.width;
        if (dsts[i].values_cap < need || (need && !dsts[i].values)) throw std::runtime_error("read_columns_into: values buffer too small");
        cps.push_back(&cp);
    }
    run_pipelined(cps, dsts, stats);
}

std::vector<Value> ParquetReader::read_column(const std::string& col_name, size_t row_group_idx) {
    int col_idx = find_column(col_name);
    if (col_idx < 0) throw std::runtime_error("Column not found: " + col_name);
    return read_column_by_idx(static_cast<int>(row_group_idx), col_idx);
}

std::vector<Value> ParquetReader::read_column(const std::string& col_name) {
    int col_idx = find_column(col_name);
    if (col_idx < 0) throw std::runtime_error("Column not found: " + col_name);
    return read_column_columnar(col_idx, -1).to_values();
}

std::vector<Value> ParquetReader::read_column_by_idx(int row_group_idx, int col_idx) {
    if (row_group_idx < 0 || row_group_idx >= static_cast<int>(metadata_.row_groups.size()))
        throw std::runtime_error("Invalid row group index");
    if (col_idx < 0 || col_idx >= static_cast<int>(columns_.size())) throw std::runtime_error("Invalid column index");
    return read_column_columnar(col_idx, row_group_idx).to_values();
}

const FileMetaData& ParquetReader::metadata() const { return metadata_; }
const std::vector<ColumnInfo>& ParquetReader::columns() const { return columns_; }
size_t ParquetReader::file_size() const { return file_size_; }

std::vector<uint8_t> ParquetReader::read_range(size_t offset, size_t length) {
    // zero-filled past EOF like the reference's unchecked ifstream read (parquet_reader.cpp:173-178)
    std::vector<uint8_t> buf(length);
    if (offset < file_size_) std::memcpy(buf.data(), data_ + offset, std::min(length, file_size_ - offset));
    return buf;
}

size_t ParquetReader::num_pages() const { return page_index_.size(); }

size_t ParquetReader::first_page_id(size_t rg, size_t col_idx) const {
    return chunk_first_page_.at(rg).at(static_cast<size_t>(columns_.at(col_idx).column_index));
}

std::vector<uint8_t> ParquetReader::read_page_data(size_t global_page_id) const {
    if (global_page_id >= page_index_.size())
        throw std::runtime_error("Global page ID " + std::to_string(global_page_id) + " out of range");
    const auto& e = page_index_[global_page_id];
    return const_cast<ParquetReader*>(this)->read_range(e.data_offset, e.data_size);
}

std::vector<uint8_t> ParquetReader::read_pages_chunk(size_t start_page_id, size_t end_page_id, size_t max_bytes) const {
    if (start_page_id >= page_index_.size()) throw std::runtime_error("Start page ID " + std::to_string(start_page_id) + " out of range");
    if (end_page_id >= page_index_.size()) throw std::runtime_error("End page ID " + std::to_string(end_page_id) + " out of range");
    if (start_page_id > end_page_id) throw std::runtime_error("Start page ID must be <= end page ID");
    std::vector<uint8_t> result;
    for (size_t i = start_page_id; i <= end_page_id; i++) {
        const auto& e = page_index_[i];
        size_t remaining = max_bytes - result.size();
        if (remaining == 0) break;
        size_t n = std::min(e.data_size, remaining);
        auto part = const_cast<ParquetReader*>(this)->read_range(e.data_offset, n);
        result.insert(result.end(), part.begin(), part.end());
    }
    return result;
}

const PageIndexEntry& ParquetReader::page_index_entry(size_t global_page_id) const {
    if (global_page_id >= page_index_.size())
        throw std::runtime_error("Global page ID " + std::to_string(global_page_id) + " out of range");
    return page_index_[global_page_id];
}

PageIterator::PageIterator(ParquetReader& reader, size_t start, size_t end)
    : reader_(reader), start_(start), end_(end), current_(start) {}
bool PageIterator::has_next() const { return current_ < end_; }
RawPage PageIterator::next() {
    if (!has_next()) throw std::runtime_error("PageIterator: no more pages");
    const auto& e = reader_.page_index_entry(current_);
    RawPage page{current_, e.row_group_idx, e.column_idx, reader_.read_page_data(current_)};
    current_++;
    return page;
}
void PageIterator::reset() { current_ = start_; }

PageIterator ParquetReader::page_iterator() { return PageIterator(*this, 0, page_index_.size()); }
PageIterator ParquetReader::page_iterator(size_t start_page_id, size_t end_page_id) {
    if (start_page_id > page_index_.size()) throw std::runtime_error("start_page_id out of range");
    if (end_page_id > page_index_.size()) throw std::runtime_error("end_page_id out of range");
    if (start_page_id > end_page_id) throw std::runtime_error("start_page_id must be <= end_page_id");
    return PageIterator(*this, start_page_id, end_page_id);
}

// ── StringColumnIterator ─────────────────────────────────────────────────────────────────
StringColumnIterator ParquetReader::column_iterator(const std::string& col_name) {
    int col_idx = find_column(col_name);
    if (col_idx < 0) throw std::runtime_error("Column not found: " + col_name);
    const auto& ci = columns_[static_cast<size_t>(col_idx)];
    if (ci.type != ParquetType::BYTE_ARRAY)
        throw std::runtime_error("Column '" + col_name + "' is not BYTE_ARRAY (type: " + parquet_type_name(ci.type) + ")");
    return StringColumnIterator(*this, static_cast<size_t>(col_idx));
}

StringColumnIterator::StringColumnIterator(ParquetReader& reader, size_t col_idx) : reader_(reader), col_idx_(col_idx) {
    if (reader_.num_row_groups() == 0) { done_ = true; return; }
    rg_idx_ = 0;
    row_group_base_ = 0;
    if (!load_next_row_group()) done_ = true;
}

// Decodes row groups (on the GPU, one at a time) until one has a non-null string.
bool StringColumnIterator::load_next_row_group() {
    while (rg_idx_ < reader_.num_row_groups()) {
        prev_ = cur_;
        cur_ = std::make_shared<DecodedColumn>(reader_.read_column_columnar(static_cast<int>(col_idx_), static_cast<int>(rg_idx_)));
        slot_ = 0;
        chunk_ = 0;
        advance_to_valid();
        if (slot_ < cur_->num_slots) return true;
        row_group_base_ += static_cast<size_t>(reader_.metadata().row_groups[rg_idx_].num_rows);
        rg_idx_++;
    }
    return false;
}

void StringColumnIterator::advance_to_valid() {
    const DecodedColumn& d = *cur_;
    while (slot_ < d.num_slots && !d.slot_valid(slot_)) slot_++;
    while (chunk_ + 1 < d.chunks.size() && d.chunks[chunk_ + 1].out_row_base <= slot_) chunk_++;
}

bool StringColumnIterator::has_next() const { return !done_; }

std::tuple<size_t, size_t, const char*> StringColumnIterator::next() {
    if (done_) throw std::runtime_error("StringColumnIterator: no more strings");
    auto [p, len] = cur_->string_at(chunk_, slot_);
    std::tuple<size_t, size_t, const char*> result{row_group_base_ + static_cast<size_t>(slot_), len,
                                                   reinterpret_cast<const char*>(p)};
    slot_++;
    advance_to_valid();
    if (slot_ >= cur_->num_slots) {
        row_group_base_ += static_cast<size_t>(reader_.metadata().row_groups[rg_idx_].num_rows);
        rg_idx_++;
        if (!load_next_row_group()) done_ = true; // prev_ keeps the returned pointer alive
    }
    return result;
}

} // namespace pqg
