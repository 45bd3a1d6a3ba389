// pq_reader.hpp -- host-side mirror of the reference's reader interface for the hot path,
// with the decode bodies running on the GPU through the C-ABI of include/pqg.h.
//
// Same class / method names, argument meaning and error text as the reference so that a
// caller can switch by changing the namespace:
//   ParquetReader        include/reader/parquet_reader.hpp:79-138
//   ColumnReader         include/reader/column_reader.hpp:19-41
//   StringColumnIterator include/reader/parquet_reader.hpp:28-62
//   PageIterator         include/reader/parquet_reader.hpp:64-77
//   Value                include/common.hpp:177-201
// plus columnar accessors (DecodedColumn) that avoid the 48-byte Value altogether -- the
// reference spends ~78 % of its time copying Values.
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <unordered_map>
#include <variant>
#include <vector>

#include "pq_format.hpp"

namespace pqg {

struct Value {
    bool is_null = true;
    std::variant<bool, int32_t, int64_t, float, double, std::string> data;

    static Value null() { return Value{true, {}}; }
    static Value from_bool(bool v) { return Value{false, v}; }
    static Value from_i32(int32_t v) { return Value{false, v}; }
    static Value from_i64(int64_t v) { return Value{false, v}; }
    static Value from_float(float v) { return Value{false, v}; }
    static Value from_double(double v) { return Value{false, v}; }
    static Value from_string(std::string v) { return Value{false, std::move(v)}; }
    std::string to_string() const;
};

struct PageIndexEntry {
    size_t data_offset;   // file offset where the page data starts (after the header)
    size_t data_size;     // compressed_page_size
    size_t row_group_idx;
    size_t column_idx;
};

struct RawPage {
    size_t page_id;
    size_t row_group_idx;
    size_t column_idx;
    std::vector<uint8_t> data;
};

struct PageResult {
    int page_num;
    PageType type;
    int32_t num_values;
    std::vector<Value> values; // decoded values for data pages; empty for dictionary pages
};

using ReadRangeFunc = std::function<std::vector<uint8_t>(size_t, size_t)>;

// One GPU context per device, created on first use.  PQG_DEVICE selects the default device.
class Device {
public:
    static Device& get(int device = -1);
    pqg_ctx* ctx() const { return ctx_; }
    pqg_ctx* ctx2();              // a second context (own streams) on the same device: the pipelined string reads alternate between the two
    int index() const { return device_; }
    ~Device();
private:
    explicit Device(int device);
    pqg_ctx* ctx_ = nullptr;
    pqg_ctx* ctx2_ = nullptr;
    int device_ = 0;
};

// Columnar result of a column decode, on the host.
struct DecodedColumn {
    ParquetType type = ParquetType::INT32;
    uint32_t width = 0;                 // bytes per fixed-width value (0 for BYTE_ARRAY)
    uint64_t num_slots = 0;             // level entries = output slots
    bool has_validity = false;          // false: every slot is non-null
    std::vector<uint8_t> values;        // num_slots * width
    std::vector<uint32_t> validity;     // bit i set = slot i non-null
    std::vector<uint32_t> offsets;      // BYTE_ARRAY: per table chunk n_c + 1 entries (see pqg.h)
    std::vector<uint64_t> char_bases;   // BYTE_ARRAY: n_chunks + 1
    std::vector<uint8_t> chars;
    std::vector<pqg_chunk_desc> chunks; // table chunks (row bases) to interpret offsets
    std::vector<pqg_page_desc> pages;
    std::vector<uint32_t> page_row_group;
    uint64_t bytes_in = 0, bytes_out = 0;
    float kernel_ms = 0;

    bool slot_valid(uint64_t i) const { return !has_validity || ((validity[i >> 5] >> (i & 31)) & 1u); }
    // BYTE_ARRAY: (pointer, length) of slot i inside table chunk c (i is column-global)
    std::pair<const uint8_t*, uint32_t> string_at(size_t chunk, uint64_t i) const;
    size_t chunk_of_slot(uint64_t i) const;
    // the reference's output type, slot by slot (nulls, variant alternative, payload bits)
    std::vector<Value> to_values() const;
    void append_values(std::vector<Value>& out, uint64_t first, uint64_t last) const;
};

// Decode the pages described by `t` out of `image` (host bytes; image byte 0 = the file
// offset the tables were built against).  upload -> plan -> run -> download.
DecodedColumn decode_column(Device& dev, const uint8_t* image, size_t image_size, const ColumnTables& t);

class ColumnReader {
public:
    ColumnReader(ReadRangeFunc read_range, const ColumnChunk& chunk, ParquetType type,
                 int16_t max_def_level, int16_t max_rep_level);

    std::vector<Value> read_all();
    std::vector<PageResult> read_pages();
    DecodedColumn read_columnar(); // same decode, columnar result

private:
    struct Loaded { std::vector<uint8_t> bytes; uint64_t file_off; std::vector<PageRecord> pages; };
    Loaded load();
    ReadRangeFunc read_range_;
    const ColumnMetaData* meta_;
    ParquetType type_;
    int16_t max_def_level_;
    int16_t max_rep_level_;
};

class ParquetReader;
struct CachedPlan; // pq_reader.cpp: descriptor tables + device image + decode plan of one column

// Destination of a pipelined column read (caller-owned host memory; pinned memory keeps the
// device->host copies asynchronous).
struct ColumnDst {
    void* values = nullptr;          // num_slots * width bytes
    uint64_t values_cap = 0;         // bytes available behind `values`
    uint32_t* validity = nullptr;    // ceil(num_slots / 32) words, or nullptr
    uint64_t validity_cap = 0;       // words available behind `validity`
};
// Destination of a pipelined BYTE_ARRAY read (read_strings_into_range): the caller's (ideally pinned) buffers.
//   offsets    total_slots + n_chunks entries: chunk c owns [chunk_row_base[c] + c, chunk_row_base[c + 1] + c], offsets relative to the chunk's chars
//   chars      all string bytes, chunk after chunk; char_bases[c] = where chunk c's bytes start (n_chunks + 1 entries)
//   validity   ceil(total_slots / 32) words over the whole range (may be null)
struct StringsDst {
    uint32_t* offsets = nullptr; uint64_t offsets_cap = 0;   // entries
    uint8_t* chars = nullptr; uint64_t chars_cap = 0;        // bytes
    uint32_t* validity = nullptr; uint64_t validity_cap = 0; // words
    uint64_t* char_bases = nullptr; uint64_t char_bases_cap = 0;
};
struct StringsReadStats {
    uint64_t num_slots = 0, n_chunks = 0, chars_size = 0;
    int32_t has_validity = 0;
    uint64_t bytes_in = 0, bytes_out = 0, h2d_bytes = 0, d2h_bytes = 0;
};
struct ColumnReadStats {
    uint64_t num_slots = 0;
    uint32_t width = 0;
    int32_t has_validity = 0;
    uint64_t bytes_in = 0, bytes_out = 0;   // algorithmic payload / output bytes
    uint64_t h2d_bytes = 0, d2h_bytes = 0;  // bytes actually copied each way
};

class StringColumnIterator {
public:
    bool has_next() const;
    std::tuple<size_t, size_t, const char*> next(); // (global_pos, string_len, string_ptr)

private:
    friend class ParquetReader;
    StringColumnIterator(ParquetReader& reader, size_t col_idx);
    bool load_next_row_group();

    ParquetReader& reader_;
    size_t col_idx_;
    size_t rg_idx_ = 0;
    size_t row_group_base_ = 0;
    std::shared_ptr<DecodedColumn> cur_; // the row group being iterated (kept alive for the pointers)
    std::shared_ptr<DecodedColumn> prev_; // pointers returned by next() stay valid one row group longer
    uint64_t slot_ = 0;
    size_t chunk_ = 0;
    bool done_ = false;
    void advance_to_valid();
};

class PageIterator {
public:
    PageIterator(ParquetReader& reader, size_t start, size_t end);
    bool has_next() const;
    RawPage next();
    void reset();
private:
    ParquetReader& reader_;
    size_t start_, end_, current_;
};

class ParquetReader {
public:
    ParquetReader();
    ~ParquetReader();
    ParquetReader(const ParquetReader&) = delete;
    ParquetReader& operator=(const ParquetReader&) = delete;

    bool open(const std::string& filename);
    // same, over bytes already in memory (e.g. pinned host memory); borrows the pointer
    bool open_memory(const uint8_t* data, size_t size);
    const std::string& open_error() const { return open_error_; }

    // ── schema inspection ──
    size_t num_columns() const;
    int64_t num_rows() const;
    size_t num_row_groups() const;
    std::vector<std::string> column_names() const;
    const ColumnInfo& column(size_t col_idx) const;
    const ColumnInfo& column(const std::string& name) const;
    int find_column(const std::string& name) const;
    std::string schema_string() const;

    // ── column reading (GPU decode) ──
    std::vector<Value> read_column(const std::string& col_name, size_t row_group_idx);
    std::vector<Value> read_column(const std::string& col_name);
    std::vector<Value> read_column_by_idx(int row_group_idx, int col_idx);
    // columnar variants: row_group_idx < 0 = all row groups
    DecodedColumn read_column_columnar(int col_idx, int row_group_idx = -1);
    ColumnTables column_tables(int col_idx, int row_group_idx = -1) const;
    ColumnTables column_tables_range(int col_idx, size_t rg_begin, size_t rg_end) const; // row groups [begin, end)
    // Streaming read of fixed-width columns from the (host) file image into caller-owned host
    // buffers: per row group H2D -> decode -> D2H on three streams, all requested columns in
    // one pipeline.  Descriptor tables, the device image and the plan of a column are built on
    // first use and kept until release_plans() / close.  row_group_idx < 0 = all row groups.
    void read_columns_into(const int* col_idx, int n_cols, int row_group_idx, const ColumnDst* dsts, ColumnReadStats* stats);
    void read_columns_into_range(const int* col_idx, int n_cols, size_t rg_begin, size_t rg_end, const ColumnDst* dsts, ColumnReadStats* stats);
    // Dictionary-form read of a BYTE_ARRAY column that is dictionary-encoded throughout (late
    // materialisation): uint32 dictionary indices per slot (0 for nulls) + validity, pipelined like
    // read_columns_into.  Throws "... not dictionary-encoded throughout" otherwise.  The strings
    // of slot i are entry indices[i] of chunk_dictionary(row group of i).
    void read_dictionary_indices_into(int col_idx, size_t rg_begin, size_t rg_end, const ColumnDst& dst, ColumnReadStats* stats);
    // Pipelined read of a BYTE_ARRAY column (the reference has no counterpart: its read_column materialises Values one by
    // one, src/reader/parquet_reader.cpp:125-165): one cached plan per row group, alternating between two contexts of the
    // device, so that the upload and size pass of row group k + 1 overlap the copy pass and the D2H of row group k.
    // The output sizes are data dependent: `chars_cap` too small fails with the bytes needed so far in the message.
    void read_strings_into_range(int col_idx, size_t rg_begin, size_t rg_end, const StringsDst& dst, StringsReadStats* stats);
    // the dictionary page of one column chunk, parsed on the host: n + 1 offsets into chars
    void chunk_dictionary(int col_idx, size_t rg, std::vector<uint32_t>& offsets, std::vector<uint8_t>& chars) const;
    void release_plans();
    // The cached device-side state of a column: descriptor tables, device image, decode plan.
    // upload = true copies the column's bytes into the device image (synchronously ordered on
    // the context's stream).  plan is null for a column without pages.
    struct DevicePlan { pqg_ctx* ctx; pqg_buf* image; pqg_plan* plan; uint32_t n_pages; uint32_t n_chunks; };
    DevicePlan device_plan(int col_idx, int row_group_idx, bool upload);
    DevicePlan device_plan_range(int col_idx, size_t rg_begin, size_t rg_end, bool upload);
    // contiguous split of the row groups over n_shards, balanced by the byte size of the column's
    // chunks (col_idx < 0: of all columns); returns n_shards + 1 boundaries (multi-GPU sharding)
    std::vector<int32_t> shard_row_groups(int col_idx, int n_shards) const;

    StringColumnIterator column_iterator(const std::string& col_name);

    // ── raw page data API ──
    size_t num_pages() const;
    std::vector<uint8_t> read_page_data(size_t global_page_id) const;
    const PageIndexEntry& page_index_entry(size_t global_page_id) const;
    std::vector<uint8_t> read_pages_chunk(size_t start_page_id, size_t end_page_id, size_t max_bytes) const;
    PageIterator page_iterator();
    PageIterator page_iterator(size_t start_page_id, size_t end_page_id);

    // ── accessors ──
    const FileMetaData& metadata() const;
    const std::vector<ColumnInfo>& columns() const;
    size_t file_size() const;
    std::vector<uint8_t> read_range(size_t offset, size_t length);
    const uint8_t* file_data() const { return data_; }
    const std::vector<PageRecord>& chunk_pages(size_t rg, size_t chunk_col) const { return chunk_pages_[rg][chunk_col]; }
    double page_scan_seconds() const { return scan_seconds_; }
    // global page id of the first data page of (rg, leaf column)
    size_t first_page_id(size_t rg, size_t col_idx) const;

    void set_device(int device) { device_ = device; }
    // Beyond the reference: with extensions on, SNAPPY-compressed chunks and DATA_PAGE_V2 pages of flat columns
    // decode (read_column*, read_column_columnar; pqg_plan_create_ext) instead of being refused / skipped like the reference
    // does (src/reader/column_reader.cpp:13-15,66-67).  Off by default: the reader then behaves exactly like the reference.
    void set_extensions(bool on) { extensions_ = on; }
    bool extensions() const { return extensions_; }
    int device() const { return device_; }

private:
    bool finish_open();
    void close_file();

    const uint8_t* data_ = nullptr;
    size_t file_size_ = 0;
    bool mapped_ = false;
    int device_ = -1;
    bool extensions_ = false;
    std::string open_error_;
    double scan_seconds_ = 0;
    FileMetaData metadata_;
    std::vector<ColumnInfo> columns_;
    std::unordered_map<std::string, size_t> column_name_to_idx_;
    std::vector<PageIndexEntry> page_index_;
    std::vector<std::vector<std::vector<PageRecord>>> chunk_pages_; // [rg][chunk column]
    std::vector<std::vector<size_t>> chunk_first_page_;              // [rg][chunk column] -> global id
    std::map<std::tuple<int, size_t, size_t, int>, std::unique_ptr<CachedPlan>> plans_; // (column, row groups [begin, end), dictionary form + 2 * context lane)
    uint32_t* vscratch_[2] = {nullptr, nullptr}; // pinned validity staging of the pipelined string reads, one per lane
    uint64_t vscratch_words_[2] = {0, 0};
    CachedPlan& cached_plan(int col_idx, int row_group_idx);
    CachedPlan& cached_plan_range(int col_idx, size_t rg_begin, size_t rg_end, bool dict_indices = false, int lane = 0);
    void run_pipelined(const std::vector<CachedPlan*>& cps, const ColumnDst* dsts, ColumnReadStats* stats);
};

} // namespace pqg
