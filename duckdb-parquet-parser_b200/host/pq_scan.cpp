// pq_scan.cpp -- see pq_scan.hpp: host drivers of the regex page-pruning scan and the 4 KB
// chunk indexes.  The host only prepares descriptor tables and moves bytes; matching and
// chunking run on the GPU (csrc/pqg_scan.cu).
#include "pq_scan.hpp"

#include <cstring>
#include <stdexcept>
#include <vector>

#include "pq_regex.hpp"

namespace pqg {

namespace {
struct DfaGuard {
    pqg_dfa* d = nullptr;
    ~DfaGuard() { pqg_dfa_free(d); }
};

int string_column(ParquetReader& r, const std::string& col_name) {
    int col = r.find_column(col_name);
    if (col < 0) throw std::runtime_error("Column not found: " + col_name);
    const ColumnInfo& ci = r.column(static_cast<size_t>(col));
    if (ci.type != ParquetType::BYTE_ARRAY)
        throw std::runtime_error("Column '" + col_name + "' is not BYTE_ARRAY (type: " + parquet_type_name(ci.type) + ")");
    return col;
}
} // namespace

int64_t regex_prune(ParquetReader& r, int col, const std::string& pattern, bool neg, uint8_t* bits, int64_t cap, float* kernel_ms) {
    if (col < 0 || col >= static_cast<int>(r.num_columns())) throw std::runtime_error("Invalid column index");
    const ColumnInfo& ci = r.column(static_cast<size_t>(col));
    if (ci.type != ParquetType::BYTE_ARRAY)
        throw std::runtime_error("Column '" + ci.name + "' is not BYTE_ARRAY (type: " + parquet_type_name(ci.type) + ")");
    DfaGuard g;
    char err[512] = {0};
    if (pqg_regex_compile(pattern.c_str(), &g.d, err, sizeof(err)) != PQG_OK) throw std::runtime_error(err);
    ParquetReader::DevicePlan dp = r.device_plan(col, -1, true);
    if (kernel_ms) *kernel_ms = 0;
    if (!dp.plan) return 0;
    std::vector<uint32_t> words((static_cast<size_t>(dp.n_pages) + 31) / 32 + 1, 0);
    if (pqg_regex_scan(dp.ctx, dp.plan, g.d, neg ? 1 : 0, words.data(), kernel_ms) != PQG_OK)
        throw std::runtime_error(pqg_last_error(dp.ctx));
    for (int64_t p = 0; p < static_cast<int64_t>(dp.n_pages) && p < cap; p++) bits[p] = (words[static_cast<size_t>(p) >> 5] >> (p & 31)) & 1u;
    return static_cast<int64_t>(dp.n_pages);
}

int64_t chunk_index(ParquetReader& r, const std::string& col_name, uint64_t chunk_size, uint64_t* tuple_to_chunk, int64_t num_rows) {
    int col = string_column(r, col_name);
    for (int64_t i = 0; i < num_rows; i++) tuple_to_chunk[i] = 0;
    ParquetReader::DevicePlan dp = r.device_plan(col, -1, true);
    if (!dp.plan) return 1; // no values: chunk 0 only
    if (pqg_plan_run(dp.ctx, dp.plan) != PQG_OK) throw std::runtime_error(pqg_last_error(dp.ctx));
    pqg_page_error pe;
    if (pqg_plan_finish(dp.ctx, dp.plan, &pe) != PQG_OK) throw std::runtime_error(pqg_last_error(dp.ctx));
    const uint64_t n = pqg_plan_num_slots(dp.plan);
    std::vector<uint32_t> ids(n + 1);
    uint64_t n_chunks = 0, carry = 0;
    if (pqg_chunk_index(dp.ctx, dp.plan, chunk_size, 0, ids.data(), &n_chunks, &carry, nullptr) != PQG_OK)
        throw std::runtime_error(pqg_last_error(dp.ctx));
    for (uint64_t i = 0; i < n && static_cast<int64_t>(i) < num_rows; i++) tuple_to_chunk[i] = ids[i];
    return static_cast<int64_t>(n_chunks);
}

int64_t page_chunk_index(ParquetReader& r, int col, uint64_t chunk_size, uint32_t* page_chunk, uint32_t* page_off,
                         uint32_t* chunk_first_page, int64_t cap, int64_t* first_global_page, int64_t* n_col_pages) {
    if (col < 0 || col >= static_cast<int>(r.num_columns())) throw std::runtime_error("Invalid column index");
    const size_t want = static_cast<size_t>(r.column(static_cast<size_t>(col)).column_index);
    std::vector<uint32_t> sizes;
    int64_t first = -1;
    for (size_t g = 0; g < r.num_pages(); g++) {
        const PageIndexEntry& e = r.page_index_entry(g);
        if (e.column_idx != want) continue;
        if (first < 0) first = static_cast<int64_t>(g);
        sizes.push_back(static_cast<uint32_t>(e.data_size));
    }
    if (first_global_page) *first_global_page = first;
    if (n_col_pages) *n_col_pages = static_cast<int64_t>(sizes.size());
    if (sizes.empty()) return 0;
    pqg_ctx* ctx = Device::get(r.device()).ctx();
    uint32_t n_chunks = 0;
    std::vector<uint32_t> firsts(sizes.size() + 1);
    if (pqg_page_chunk_index(ctx, sizes.data(), static_cast<uint32_t>(sizes.size()), chunk_size, page_chunk, page_off, firsts.data(),
                             static_cast<uint32_t>(firsts.size()), &n_chunks) != PQG_OK)
        throw std::runtime_error(pqg_last_error(ctx));
    for (int64_t c = 0; c < static_cast<int64_t>(n_chunks) && c < cap; c++) chunk_first_page[c] = firsts[static_cast<size_t>(c)];
    return static_cast<int64_t>(n_chunks);
}

} // namespace pqg
