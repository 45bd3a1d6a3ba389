// pq_format.hpp -- host-side Parquet footer / page-header parsing (Thrift compact
// protocol) and the flat descriptor tables handed to the CUDA decoder.
//
// This is the part the north star keeps on the host: it replaces the reference's
// ThriftReader (src/reader/thrift.cpp), the *::deserialize functions
// (src/reader/metadata.cpp), build_column_info (src/reader/parquet_reader.cpp:484-557)
// and build_page_index (:559-605).  Written from the Parquet / Thrift-compact formats, not
// from those files: a single-pass, allocation-free cursor over an in-memory image, page
// headers of any size (the reference reads a fixed 256-byte window per header), and a
// multi-threaded page scan (one task per column chunk) because files written by the
// reference's writer have ~1 KB pages, i.e. 10^6..10^8 headers.
#pragma once
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "pqg.h"

namespace pqg {

// Parquet format enums (values fixed by the format; same names as the reference's
// include/common.hpp:16-88 so that call sites read the same).
enum class ParquetType : int32_t { BOOLEAN = 0, INT32 = 1, INT64 = 2, INT96 = 3, FLOAT = 4, DOUBLE = 5, BYTE_ARRAY = 6, FIXED_LEN_BYTE_ARRAY = 7 };
enum class Encoding : int32_t { PLAIN = 0, GROUP_VAR_INT = 1, PLAIN_DICTIONARY = 2, RLE = 3, BIT_PACKED = 4, DELTA_BINARY_PACKED = 5, DELTA_LENGTH_BYTE_ARRAY = 6, DELTA_BYTE_ARRAY = 7, RLE_DICTIONARY = 8, BYTE_STREAM_SPLIT = 9 };
enum class CompressionCodec : int32_t { UNCOMPRESSED = 0, SNAPPY = 1, GZIP = 2, LZO = 3, BROTLI = 4, LZ4 = 5, ZSTD = 6, LZ4_RAW = 7 };
enum class PageType : int32_t { DATA_PAGE = 0, INDEX_PAGE = 1, DICTIONARY_PAGE = 2, DATA_PAGE_V2 = 3 };
enum class FieldRepetitionType : int32_t { REQUIRED = 0, OPTIONAL = 1, REPEATED = 2 };
enum class ConvertedType : int32_t { NONE = -1, UTF8 = 0, MAP = 1, MAP_KEY_VALUE = 2, LIST = 3, ENUM = 4, DECIMAL = 5, DATE = 6, TIME_MILLIS = 7, TIME_MICROS = 8, TIMESTAMP_MILLIS = 9, TIMESTAMP_MICROS = 10, UINT_8 = 11, UINT_16 = 12, UINT_32 = 13, UINT_64 = 14, INT_8 = 15, INT_16 = 16, INT_32 = 17, INT_64 = 18, JSON = 19, BSON = 20, INTERVAL = 21 };

const char* parquet_type_name(ParquetType t);
const char* encoding_name(Encoding e);
const char* compression_name(CompressionCodec c);
const char* page_type_name(PageType t);
const char* converted_type_name(ConvertedType ct);

// ── metadata (same field names as the reference's include/reader/metadata.hpp) ──────────
struct SchemaElement {
    std::optional<ParquetType> type;
    std::optional<int32_t> type_length;
    std::optional<FieldRepetitionType> repetition_type;
    std::string name;
    std::optional<int32_t> num_children;
    std::optional<ConvertedType> converted_type;
    std::optional<int32_t> scale, precision, field_id;
};

struct ColumnMetaData {
    ParquetType type = ParquetType::INT32;
    std::vector<Encoding> encodings;
    std::vector<std::string> path_in_schema;
    CompressionCodec codec = CompressionCodec::UNCOMPRESSED;
    int64_t num_values = 0;
    int64_t total_uncompressed_size = 0;
    int64_t total_compressed_size = 0;
    int64_t data_page_offset = 0;
    std::optional<int64_t> index_page_offset;
    std::optional<int64_t> dictionary_page_offset;
};

struct ColumnChunk {
    std::optional<std::string> file_path;
    int64_t file_offset = 0;
    std::optional<ColumnMetaData> meta_data;
};

struct RowGroup {
    std::vector<ColumnChunk> columns;
    int64_t total_byte_size = 0;
    int64_t num_rows = 0;
};

struct KeyValue {
    std::string key;
    std::optional<std::string> value;
};

struct FileMetaData {
    int32_t version = 0;
    std::vector<SchemaElement> schema;
    int64_t num_rows = 0;
    std::vector<RowGroup> row_groups;
    std::vector<KeyValue> key_value_metadata;
    std::optional<std::string> created_by;
};

struct DataPageHeader {
    int32_t num_values = 0;
    Encoding encoding = Encoding::PLAIN;
    Encoding definition_level_encoding = Encoding::RLE;
    Encoding repetition_level_encoding = Encoding::RLE;
};
struct DictionaryPageHeader {
    int32_t num_values = 0;
    Encoding encoding = Encoding::PLAIN_DICTIONARY;
    bool is_sorted = false;
};
struct PageHeader {
    PageType type = PageType::DATA_PAGE;
    int32_t uncompressed_page_size = 0;
    int32_t compressed_page_size = 0;
    std::optional<int32_t> crc;
    std::optional<DataPageHeader> data_page_header;
    std::optional<DictionaryPageHeader> dictionary_page_header;
    bool is_v2 = false;
    std::optional<DataPageHeader> v2_header; // DATA_PAGE_V2: num_values + encoding (the reference steps over such pages)
    int32_t v2_def_len = 0, v2_rep_len = 0;  // DATA_PAGE_V2: definition_ / repetition_levels_byte_length
    bool v2_compressed = true;               // DATA_PAGE_V2: is_compressed (default true)
};

// Leaf column description (reference include/reader/column_info.hpp).
struct ColumnInfo {
    std::string name;
    ParquetType type = ParquetType::BYTE_ARRAY;
    int column_index = 0;
    int16_t max_def_level = 0;
    int16_t max_rep_level = 0;
    std::optional<FieldRepetitionType> repetition;
    std::optional<ConvertedType> converted_type;

    std::string type_name() const { return parquet_type_name(type); }
    std::string converted_type_string() const {
        return (converted_type && *converted_type != ConvertedType::NONE) ? converted_type_name(*converted_type) : "NONE";
    }
    bool is_required() const { return repetition && *repetition == FieldRepetitionType::REQUIRED; }
    bool is_optional() const { return repetition && *repetition == FieldRepetitionType::OPTIONAL; }
    bool is_repeated() const { return repetition && *repetition == FieldRepetitionType::REPEATED; }
};

// Thrown for malformed Thrift / out-of-bounds reads.  The text mirrors the reference's
// ByteBuffer::check message (include/common.hpp:162-168) so callers see the same error.
struct FormatError : std::runtime_error { using std::runtime_error::runtime_error; };

// Parses a FileMetaData struct from [data, data+size).
FileMetaData parse_file_metadata(const uint8_t* data, size_t size);
// Parses one PageHeader at data (at most `avail` bytes readable); returns the header size.
size_t parse_page_header(const uint8_t* data, size_t avail, PageHeader& out);
// Leaf columns with their max definition / repetition levels.
std::vector<ColumnInfo> build_column_info(const FileMetaData& md);

// ── page walk of one column chunk ────────────────────────────────────────────────────────
struct PageRecord {
    uint64_t payload_off;  // file offset of the first payload byte
    uint32_t payload_size; // compressed_page_size
    int32_t num_values;    // data pages: level entries; dictionary pages: entries
    PageType type;
    Encoding encoding;
    bool counted;          // contributes to the global page index (DATA_PAGE / DATA_PAGE_V2)
    uint32_t uncompressed_size; // PageHeader.uncompressed_page_size
    uint32_t v2_def_len, v2_rep_len; // DATA_PAGE_V2 level lengths
    bool v2_compressed;    // DATA_PAGE_V2: the value section is compressed with the chunk's codec
    bool one_level_run;    // data pages: the payload starts like <u32 length><RLE run of value 1 covering num_values> (PQG_PAGE_FLAG_NO_NULLS for max_def 1 columns)
};

// Walks the pages of a chunk exactly like every reference loop does
// (`while (values_read < num_values)`, src/reader/column_reader.cpp:32), but over an
// in-memory image [base_off, base_off + size) of the file.  `pages` receives every page in
// file order (dictionary, data, other).  Throws FormatError on malformed headers.
void walk_chunk_pages(const uint8_t* image, uint64_t image_file_off, uint64_t image_size,
                      uint64_t chunk_start, int64_t num_values, std::vector<PageRecord>& pages);

inline uint64_t chunk_start_offset(const ColumnMetaData& m) {
    int64_t off = m.data_page_offset;
    if (m.dictionary_page_offset && *m.dictionary_page_offset < off) off = *m.dictionary_page_offset;
    return static_cast<uint64_t>(off);
}

// Descriptor tables of one column over a list of chunks (one entry per row group, in
// order).  A chunk that switches dictionaries mid-way becomes several table chunks.
struct ColumnTables {
    std::vector<pqg_chunk_desc> chunks;
    std::vector<pqg_page_desc> pages;
    // extension mode (DATA_PAGE_V2 / SNAPPY, pqg_plan_create_ext): parallel to chunks / pages; `ext` = some entry needs it
    std::vector<pqg_chunk_ext> chunk_ext;
    std::vector<pqg_page_ext> page_ext;
    bool ext = false;
    std::vector<uint32_t> page_row_group; // per page: source row group (for views)
    uint64_t total_slots = 0;
};

// Appends the tables of one column chunk.  image_file_off = file offset of image byte 0.
// extensions: list DATA_PAGE_V2 pages as decodable pages and record `codec` (pqg_plan_create_ext); otherwise V2 pages are
// listed with PQG_PAGE_FLAG_V2 (refused at plan creation, like the reference never decodes them)
void append_chunk_tables(ColumnTables& t, const std::vector<PageRecord>& pages, uint64_t image_file_off,
                         ParquetType type, int16_t max_def, int16_t max_rep, uint32_t rg, uint32_t col,
                         bool extensions = false, CompressionCodec codec = CompressionCodec::UNCOMPRESSED);

} // namespace pqg
