// pq_format.cpp -- see pq_format.hpp.
#include "pq_format.hpp"

#include <cstring>

namespace pqg {

const char* parquet_type_name(ParquetType t) {
    static const char* n[] = {"BOOLEAN", "INT32", "INT64", "INT96", "FLOAT", "DOUBLE", "BYTE_ARRAY", "FIXED_LEN_BYTE_ARRAY"};
    int i = static_cast<int>(t);
    return (i >= 0 && i < 8) ? n[i] : "UNKNOWN";
}
const char* encoding_name(Encoding e) {
    static const char* n[] = {"PLAIN", "GROUP_VAR_INT", "PLAIN_DICTIONARY", "RLE", "BIT_PACKED", "DELTA_BINARY_PACKED",
                              "DELTA_LENGTH_BYTE_ARRAY", "DELTA_BYTE_ARRAY", "RLE_DICTIONARY", "BYTE_STREAM_SPLIT"};
    int i = static_cast<int>(e);
    return (i >= 0 && i < 10) ? n[i] : "UNKNOWN";
}
const char* compression_name(CompressionCodec c) {
    static const char* n[] = {"UNCOMPRESSED", "SNAPPY", "GZIP", "LZO", "BROTLI", "LZ4", "ZSTD", "LZ4_RAW"};
    int i = static_cast<int>(c);
    return (i >= 0 && i < 8) ? n[i] : "UNKNOWN";
}
const char* page_type_name(PageType t) {
    static const char* n[] = {"DATA_PAGE", "INDEX_PAGE", "DICTIONARY_PAGE", "DATA_PAGE_V2"};
    int i = static_cast<int>(t);
    return (i >= 0 && i < 4) ? n[i] : "UNKNOWN";
}
const char* converted_type_name(ConvertedType ct) {
    static const char* n[] = {"UTF8", "MAP", "MAP_KEY_VALUE", "LIST", "ENUM", "DECIMAL", "DATE", "TIME_MILLIS",
                              "TIME_MICROS", "TIMESTAMP_MILLIS", "TIMESTAMP_MICROS", "UINT_8", "UINT_16", "UINT_32",
                              "UINT_64", "INT_8", "INT_16", "INT_32", "INT_64", "JSON", "BSON", "INTERVAL"};
    int i = static_cast<int>(ct);
    if (i == -1) return "NONE";
    return (i >= 0 && i < 22) ? n[i] : "UNKNOWN";
}

namespace {

// Thrift compact-protocol wire types
enum : uint8_t { T_STOP = 0, T_TRUE = 1, T_FALSE = 2, T_I8 = 3, T_I16 = 4, T_I32 = 5, T_I64 = 6,
                 T_DOUBLE = 7, T_BINARY = 8, T_LIST = 9, T_SET = 10, T_MAP = 11, T_STRUCT = 12 };

// Pointer cursor; every struct parser keeps its own "previous field id" on the C++ stack.
class Cursor {
public:
    Cursor(const uint8_t* p, size_t n) : begin_(p), p_(p), end_(p + n) {}
    size_t offset() const { return static_cast<size_t>(p_ - begin_); }

    uint8_t byte() { need(1); return *p_++; }
    uint64_t varint() {
        uint64_t v = 0;
        for (int shift = 0;; shift += 7) {
            if (shift > 63) throw FormatError("varint too long");
            uint8_t b = byte();
            v |= static_cast<uint64_t>(b & 0x7F) << shift;
            if (!(b & 0x80)) return v;
        }
    }
    int64_t zigzag() { uint64_t v = varint(); return static_cast<int64_t>((v >> 1) ^ (~(v & 1) + 1)); }
    int32_t i32() { return static_cast<int32_t>(zigzag()); }
    int64_t i64() { return zigzag(); }
    std::string str() {
        uint32_t len = static_cast<uint32_t>(varint());
        need(len);
        std::string s(reinterpret_cast<const char*>(p_), len);
        p_ += len;
        return s;
    }
    void skip_bytes(size_t n) { need(n); p_ += n; }

    // field header: returns false at STOP
    bool field(int16_t& last_id, int16_t& id, uint8_t& type) {
        uint8_t b = byte();
        if (b == T_STOP) return false;
        type = b & 0x0F;
        uint8_t delta = b >> 4;
        id = delta ? static_cast<int16_t>(last_id + delta) : static_cast<int16_t>(zigzag());
        last_id = id;
        return true;
    }
    void list(uint8_t& elem_type, int32_t& count) {
        uint8_t b = byte();
        elem_type = b & 0x0F;
        const uint64_t n = (b >> 4) == 0x0F ? varint() : (b >> 4);
        // every element takes at least one byte on the wire: a larger count is a corrupt header
        if (n > static_cast<uint64_t>(end_ - p_)) throw FormatError("thrift list longer than its buffer");
        count = static_cast<int32_t>(n);
    }
    void skip(uint8_t type, int depth = 0) {
        if (depth > 64) throw FormatError("thrift nesting too deep");
        switch (type) {
            case T_TRUE: case T_FALSE: break;
            case T_I8: byte(); break;
            case T_I16: case T_I32: case T_I64: varint(); break;
            case T_DOUBLE: skip_bytes(8); break;
            case T_BINARY: skip_bytes(static_cast<uint32_t>(varint())); break;
            case T_LIST: case T_SET: {
                uint8_t et; int32_t n;
                list(et, n);
                if (et == T_TRUE || et == T_FALSE) skip_bytes(static_cast<size_t>(n)); // bool elements are one byte each inside a list
                else for (int32_t i = 0; i < n; i++) skip(et, depth + 1);
                break;
            }
            case T_MAP: {
                const uint64_t n = varint();
                if (n > static_cast<uint64_t>(end_ - p_)) throw FormatError("thrift map longer than its buffer");
                if (n > 0) {
                    uint8_t kv = byte();
                    auto elem = [&](uint8_t t) { if (t == T_TRUE || t == T_FALSE) skip_bytes(1); else skip(t, depth + 1); };
                    for (uint64_t i = 0; i < n; i++) { elem(kv >> 4); elem(kv & 0x0F); }
                }
                break;
            }
            case T_STRUCT: {
                int16_t last = 0, id; uint8_t t;
                while (field(last, id, t)) skip(t, depth + 1);
                break;
            }
            default: throw FormatError("ThriftReader::skip: unknown type " + std::to_string(type));
        }
    }

private:
    void need(size_t n) {
        if (static_cast<size_t>(end_ - p_) < n)
            throw FormatError("ByteBuffer: read beyond end (pos=" + std::to_string(offset()) + " need=" +
                              std::to_string(n) + " size=" + std::to_string(static_cast<size_t>(end_ - begin_)) + ")");
    }
    const uint8_t* begin_;
    const uint8_t* p_;
    const uint8_t* end_;
};

SchemaElement parse_schema_element(Cursor& c) {
    SchemaElement e;
    int16_t last = 0, id; uint8_t t;
    while (c.field(last, id, t)) {
        switch (id) {
            case 1: e.type = static_cast<ParquetType>(c.i32()); break;
            case 2: e.type_length = c.i32(); break;
            case 3: e.repetition_type = static_cast<FieldRepetitionType>(c.i32()); break;
            case 4: e.name = c.str(); break;
            case 5: e.num_children = c.i32(); break;
            case 6: e.converted_type = static_cast<ConvertedType>(c.i32()); break;
            case 7: e.scale = c.i32(); break;
            case 8: e.precision = c.i32(); break;
            case 9: e.field_id = c.i32(); break;
            default: c.skip(t);
        }
    }
    return e;
}

ColumnMetaData parse_column_meta(Cursor& c) {
    ColumnMetaData m;
    int16_t last = 0, id; uint8_t t;
    while (c.field(last, id, t)) {
        switch (id) {
            case 1: m.type = static_cast<ParquetType>(c.i32()); break;
            case 2: { uint8_t et; int32_t n; c.list(et, n); for (int32_t i = 0; i < n; i++) m.encodings.push_back(static_cast<Encoding>(c.i32())); break; }
            case 3: { uint8_t et; int32_t n; c.list(et, n); for (int32_t i = 0; i < n; i++) m.path_in_schema.push_back(c.str()); break; }
            case 4: m.codec = static_cast<CompressionCodec>(c.i32()); break;
            case 5: m.num_values = c.i64(); break;
            case 6: m.total_uncompressed_size = c.i64(); break;
            case 7: m.total_compressed_size = c.i64(); break;
            case 9: m.data_page_offset = c.i64(); break;
            case 10: m.index_page_offset = c.i64(); break;
            case 11: m.dictionary_page_offset = c.i64(); break;
            default: c.skip(t);
        }
    }
    return m;
}

ColumnChunk parse_column_chunk(Cursor& c) {
    ColumnChunk cc;
    int16_t last = 0, id; uint8_t t;
    while (c.field(last, id, t)) {
        switch (id) {
            case 1: cc.file_path = c.str(); break;
            case 2: cc.file_offset = c.i64(); break;
            case 3: cc.meta_data = parse_column_meta(c); break;
            default: c.skip(t);
        }
    }
    return cc;
}

RowGroup parse_row_group(Cursor& c) {
    RowGroup rg;
    int16_t last = 0, id; uint8_t t;
    while (c.field(last, id, t)) {
        switch (id) {
            case 1: { uint8_t et; int32_t n; c.list(et, n); rg.columns.reserve(n > 0 ? n : 0); for (int32_t i = 0; i < n; i++) rg.columns.push_back(parse_column_chunk(c)); break; }
            case 2: rg.total_byte_size = c.i64(); break;
            case 3: rg.num_rows = c.i64(); break;
            default: c.skip(t);
        }
    }
    return rg;
}

} // namespace

FileMetaData parse_file_metadata(const uint8_t* data, size_t size) {
    Cursor c(data, size);
    FileMetaData md;
    int16_t last = 0, id; uint8_t t;
    while (c.field(last, id, t)) {
        switch (id) {
            case 1: md.version = c.i32(); break;
            case 2: { uint8_t et; int32_t n; c.list(et, n); for (int32_t i = 0; i < n; i++) md.schema.push_back(parse_schema_element(c)); break; }
            case 3: md.num_rows = c.i64(); break;
            case 4: { uint8_t et; int32_t n; c.list(et, n); for (int32_t i = 0; i < n; i++) md.row_groups.push_back(parse_row_group(c)); break; }
            case 5: {
                uint8_t et; int32_t n;
                c.list(et, n);
                for (int32_t i = 0; i < n; i++) {
                    KeyValue kv;
                    int16_t l2 = 0, id2; uint8_t t2;
                    while (c.field(l2, id2, t2)) {
                        if (id2 == 1) kv.key = c.str();
                        else if (id2 == 2) kv.value = c.str();
                        else c.skip(t2);
                    }
                    md.key_value_metadata.push_back(std::move(kv));
                }
                break;
            }
            case 6: md.created_by = c.str(); break;
            default: c.skip(t);
        }
    }
    return md;
}

size_t parse_page_header(const uint8_t* data, size_t avail, PageHeader& ph) {
    Cursor c(data, avail);
    ph = PageHeader();
    int16_t last = 0, id; uint8_t t;
    while (c.field(last, id, t)) {
        switch (id) {
            case 1: ph.type = static_cast<PageType>(c.i32()); break;
            case 2: ph.uncompressed_page_size = c.i32(); break;
            case 3: ph.compressed_page_size = c.i32(); break;
            case 4: ph.crc = c.i32(); break;
            case 5: {
                DataPageHeader d;
                int16_t l2 = 0, id2; uint8_t t2;
                while (c.field(l2, id2, t2)) {
                    switch (id2) {
                        case 1: d.num_values = c.i32(); break;
                        case 2: d.encoding = static_cast<Encoding>(c.i32()); break;
                        case 3: d.definition_level_encoding = static_cast<Encoding>(c.i32()); break;
                        case 4: d.repetition_level_encoding = static_cast<Encoding>(c.i32()); break;
                        default: c.skip(t2);
                    }
                }
                ph.data_page_header = d;
                break;
            }
            case 7: {
                DictionaryPageHeader d;
                int16_t l2 = 0, id2; uint8_t t2;
                while (c.field(l2, id2, t2)) {
                    switch (id2) {
                        case 1: d.num_values = c.i32(); break;
                        case 2: d.encoding = static_cast<Encoding>(c.i32()); break;
                        case 3: d.is_sorted = (t2 == T_TRUE); break;
                        default: c.skip(t2);
                    }
                }
                ph.dictionary_page_header = d;
                break;
            }
            case 8: { // DataPageHeaderV2 {1 num_values, 2 num_nulls, 3 num_rows, 4 encoding, 5 / 6 level byte lengths, 7 is_compressed}
                ph.is_v2 = true;
                DataPageHeader d;
                int16_t l2 = 0, id2; uint8_t t2;
                while (c.field(l2, id2, t2)) {
                    switch (id2) {
                        case 1: d.num_values = c.i32(); break;
                        case 4: d.encoding = static_cast<Encoding>(c.i32()); break;
                        case 5: ph.v2_def_len = c.i32(); break;
                        case 6: ph.v2_rep_len = c.i32(); break;
                        case 7: ph.v2_compressed = (t2 == T_TRUE); break;
                        default: c.skip(t2);
                    }
                }
                ph.v2_header = d;
                break;
            }
            default: c.skip(t);
        }
    }
    return c.offset();
}

namespace {
int schema_subtree_end(const FileMetaData& md, int idx) {
    int kids = md.schema[idx].num_children.value_or(0);
    idx++;
    for (int i = 0; i < kids && idx < static_cast<int>(md.schema.size()); i++) {
        if (md.schema[idx].num_children.value_or(0) > 0) idx = schema_subtree_end(md, idx);
        else idx++;
    }
    return idx;
}
void walk_schema(const FileMetaData& md, int idx, int end, int16_t def, int16_t rep, int& leaf, std::vector<ColumnInfo>& out) {
    while (idx < end) {
        const SchemaElement& e = md.schema[idx];
        int16_t d = def, r = rep;
        if (e.repetition_type) {
            if (*e.repetition_type == FieldRepetitionType::OPTIONAL) d++;
            else if (*e.repetition_type == FieldRepetitionType::REPEATED) { d++; r++; }
        }
        if (e.num_children.value_or(0) > 0) {
            int sub_end = idx + 1, remaining = *e.num_children;
            while (remaining > 0 && sub_end < end) {
                remaining--;
                sub_end = md.schema[sub_end].num_children.value_or(0) > 0 ? schema_subtree_end(md, sub_end) : sub_end + 1;
            }
            walk_schema(md, idx + 1, sub_end, d, r, leaf, out);
            idx = sub_end;
        } else {
            ColumnInfo ci;
            ci.name = e.name;
            ci.type = e.type.value_or(ParquetType::BYTE_ARRAY);
            ci.column_index = leaf++;
            ci.max_def_level = d;
            ci.max_rep_level = r;
            ci.repetition = e.repetition_type;
            ci.converted_type = e.converted_type;
            out.push_back(std::move(ci));
            idx++;
        }
    }
}
} // namespace

std::vector<ColumnInfo> build_column_info(const FileMetaData& md) {
    std::vector<ColumnInfo> out;
    if (md.schema.empty()) return out;
    int leaf = 0;
    walk_schema(md, 1, static_cast<int>(md.schema.size()), 0, 0, leaf, out);
    return out;
}

void walk_chunk_pages(const uint8_t* image, uint64_t image_file_off, uint64_t image_size,
                      uint64_t chunk_start, int64_t num_values, std::vector<PageRecord>& pages) {
    uint64_t cur = chunk_start;
    int64_t values_read = 0;
    const uint64_t image_end = image_file_off + image_size;
    while (values_read < num_values) {
        if (cur < image_file_off || cur >= image_end)
            throw FormatError("page walk left the column chunk image at offset " + std::to_string(cur));
        PageHeader ph;
        size_t hsize = parse_page_header(image + (cur - image_file_off), static_cast<size_t>(image_end - cur), ph);
        PageRecord r;
        r.payload_off = cur + hsize;
        r.payload_size = static_cast<uint32_t>(ph.compressed_page_size);
        r.type = ph.type;
        r.encoding = Encoding::PLAIN;
        r.num_values = 0;
        r.counted = false;
        r.one_level_run = false;
        r.uncompressed_size = ph.uncompressed_page_size < 0 ? 0u : static_cast<uint32_t>(ph.uncompressed_page_size);
        r.v2_def_len = ph.v2_def_len < 0 ? 0u : static_cast<uint32_t>(ph.v2_def_len);
        r.v2_rep_len = ph.v2_rep_len < 0 ? 0u : static_cast<uint32_t>(ph.v2_rep_len);
        r.v2_compressed = ph.v2_compressed;
        if (ph.compressed_page_size < 0) throw FormatError("negative page size");
        if (ph.type == PageType::DICTIONARY_PAGE) {
            if (!ph.dictionary_page_header) throw FormatError("bad_optional_access: dictionary page without its header");
            r.num_values = ph.dictionary_page_header->num_values;
            r.encoding = ph.dictionary_page_header->encoding;
        } else if (ph.type == PageType::DATA_PAGE) {
            r.counted = true;
            if (ph.data_page_header) {
                r.num_values = ph.data_page_header->num_values;
                r.encoding = ph.data_page_header->encoding;
                if (r.num_values < 0) throw FormatError("negative num_values");
                values_read += r.num_values;
                // routing hint for OPTIONAL flat columns: levels = one RLE run of level 1 over the whole page (no nulls)
                const uint64_t po = r.payload_off - image_file_off;
                if (r.payload_size >= 6 && po + r.payload_size <= image_size) {
                    const uint8_t* pl = image + po;
                    const uint32_t def_len = static_cast<uint32_t>(pl[0]) | static_cast<uint32_t>(pl[1]) << 8 | static_cast<uint32_t>(pl[2]) << 16 | static_cast<uint32_t>(pl[3]) << 24;
                    if (def_len >= 2 && def_len <= r.payload_size - 4) {
                        uint64_t ind = 0;
                        uint32_t shift = 0, hp = 0;
                        bool complete = false;
                        while (hp < def_len && hp < 5) { const uint8_t b = pl[4 + hp++]; ind |= static_cast<uint64_t>(b & 0x7f) << shift; shift += 7; if (!(b & 0x80)) { complete = true; break; } }
                        r.one_level_run = complete && !(ind & 1) && (ind >> 1) >= static_cast<uint64_t>(r.num_values) && hp < def_len && pl[4 + hp] == 1;
                    }
                }
            } else {
                // the reference dereferences an empty optional here (bad_optional_access)
                throw FormatError("bad_optional_access: data page without its header");
            }
        } else if (ph.type == PageType::DATA_PAGE_V2) {
            // gets a global id in the reference's page index and is never decoded there (column_reader.cpp:66-67: the
            // reference's loop does not even advance past it).  Here the page's values are counted so that the walk ends,
            // and every decode of the chunk is refused with an explicit error (pqg_plan_create: PQG_PAGE_FLAG_V2).
            r.counted = true;
            if (ph.v2_header) {
                r.num_values = ph.v2_header->num_values;
                r.encoding = ph.v2_header->encoding;
                if (r.num_values < 0) throw FormatError("negative num_values");
                values_read += r.num_values;
            }
        }
        pages.push_back(r);
        cur = r.payload_off + r.payload_size;
    }
}

void append_chunk_tables(ColumnTables& t, const std::vector<PageRecord>& pages, uint64_t image_file_off,
                         ParquetType type, int16_t max_def, int16_t max_rep, uint32_t rg, uint32_t col,
                         bool extensions, CompressionCodec codec) {
    const uint32_t pcodec = static_cast<uint32_t>(codec); // PQG_CODEC_* share the format's numbering for 0 / 1
    if (extensions && codec != CompressionCodec::UNCOMPRESSED) t.ext = true;
    // (one entry per page: grow once -- a 100 M-row PLAIN column has 780 K pages; the extension records only where they are used)
    auto grow = [](auto& v, size_t extra) { // (geometric: the tables of a column are appended chunk by chunk)
        if (v.capacity() < v.size() + extra) v.reserve(std::max(v.size() + extra, v.capacity() * 2));
    };
    grow(t.pages, pages.size());
    grow(t.page_row_group, pages.size());
    if (extensions) grow(t.page_ext, pages.size());
    auto open_chunk = [&](const PageRecord* dict) {
        pqg_chunk_desc c;
        std::memset(&c, 0, sizeof(c));
        c.out_row_base = t.total_slots;
        c.first_page = static_cast<uint32_t>(t.pages.size());
        c.row_group = rg; c.column = col;
        c.max_def = max_def; c.max_rep = max_rep;
        c.phys_type = static_cast<uint8_t>(type);
        if (dict) {
            c.has_dict = 1;
            c.dict_off = dict->payload_off - image_file_off;
            c.dict_size = dict->payload_size;
            c.dict_num_values = dict->num_values < 0 ? 0u : static_cast<uint32_t>(dict->num_values);
        }
        t.chunks.push_back(c);
        if (extensions) t.chunk_ext.push_back(pqg_chunk_ext{dict ? dict->uncompressed_size : 0u, dict ? pcodec : 0u});
    };
    bool opened = false;
    for (const PageRecord& r : pages) {
        if (r.type == PageType::DICTIONARY_PAGE) {
            // a new dictionary replaces the previous one for the following data pages
            // (column_reader.cpp:48-54): start a new table chunk
            if (opened && t.chunks.back().n_pages == 0) { t.chunks.pop_back(); if (extensions) t.chunk_ext.pop_back(); }
            open_chunk(&r);
            opened = true;
        } else if (r.type == PageType::DATA_PAGE) {
            if (!opened) { open_chunk(nullptr); opened = true; }
            pqg_chunk_desc& c = t.chunks.back();
            pqg_page_desc p;
            std::memset(&p, 0, sizeof(p));
            p.payload_off = r.payload_off - image_file_off;
            p.out_row_base = t.total_slots;
            p.payload_size = r.payload_size;
            p.num_values = static_cast<uint32_t>(r.num_values);
            p.chunk_idx = static_cast<uint32_t>(t.chunks.size() - 1);
            p.flags = PQG_PAGE_FLAGS(r.encoding == Encoding::PLAIN_DICTIONARY || r.encoding == Encoding::RLE_DICTIONARY, static_cast<int32_t>(r.encoding));
            if (max_def == 1 && max_rep == 0) p.flags |= PQG_PAGE_FLAG_LEVELS_SEEN | (r.one_level_run ? PQG_PAGE_FLAG_NO_NULLS : 0u);
            t.pages.push_back(p);
            if (extensions) t.page_ext.push_back(pqg_page_ext{r.uncompressed_size, 0u, 0u, pcodec << 8});
            t.page_row_group.push_back(rg);
            c.n_pages++;
            c.num_values += p.num_values;
            t.total_slots += p.num_values;
        }
        else if (r.type == PageType::DATA_PAGE_V2 && extensions) {
            // decodable through pqg_plan_create_ext: a data page like any other + its level lengths
            if (!opened) { open_chunk(nullptr); opened = true; }
            pqg_chunk_desc& c = t.chunks.back();
            pqg_page_desc p;
            std::memset(&p, 0, sizeof(p));
            p.payload_off = r.payload_off - image_file_off;
            p.out_row_base = t.total_slots;
            p.payload_size = r.payload_size;
            p.num_values = static_cast<uint32_t>(r.num_values);
            p.chunk_idx = static_cast<uint32_t>(t.chunks.size() - 1);
            p.flags = PQG_PAGE_FLAGS(r.encoding == Encoding::PLAIN_DICTIONARY || r.encoding == Encoding::RLE_DICTIONARY, static_cast<int32_t>(r.encoding));
            t.pages.push_back(p);
            t.page_ext.push_back(pqg_page_ext{r.uncompressed_size, r.v2_def_len, r.v2_rep_len, PQG_PAGE_EXT_V2 | ((r.v2_compressed ? pcodec : 0u) << 8)});
            t.page_row_group.push_back(rg);
            t.ext = true;
            c.n_pages++;
            c.num_values += p.num_values;
            t.total_slots += p.num_values;
        }
        else if (r.type == PageType::DATA_PAGE_V2) {
            // listed so that pqg_plan_create refuses the chunk (PQG_ERR_UNSUPPORTED); it contributes no slots
            if (!opened) { open_chunk(nullptr); opened = true; }
            pqg_page_desc p;
            std::memset(&p, 0, sizeof(p));
            p.payload_off = r.payload_off - image_file_off;
            p.out_row_base = t.total_slots;
            p.payload_size = r.payload_size;
            p.chunk_idx = static_cast<uint32_t>(t.chunks.size() - 1);
            p.flags = PQG_PAGE_FLAG_V2 | PQG_PAGE_FLAGS(false, static_cast<int32_t>(r.encoding));
            t.pages.push_back(p);
            if (extensions) t.page_ext.push_back(pqg_page_ext{0u, 0u, 0u, 0u});
            t.page_row_group.push_back(rg);
            t.chunks.back().n_pages++;
        }
        // INDEX_PAGE / unknown: skipped like the reference (column_reader.cpp:66-67)
    }
    if (opened && t.chunks.back().n_pages == 0) { t.chunks.pop_back(); if (extensions) t.chunk_ext.pop_back(); }
}

} // namespace pqg
