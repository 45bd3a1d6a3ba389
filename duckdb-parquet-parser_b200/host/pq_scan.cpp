// pq_scan.cpp -- see pq_scan.hpp: host drivers of the regex page-pruning scan and the 4 KB
// chunk indexes.  The host only prepares descriptor tables and moves bytes; matching and
// chunking run on the GPU (csrc/pqg_scan.cu).
#include "pq_scan.hpp"

#include <cstring>
#include <memory>
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
    return regex_prune_rgs(r, col, 0, r.num_row_groups(), pattern, neg, bits, cap, kernel_ms);
}

int64_t regex_prune_rgs(ParquetReader& r, int col, size_t rg_begin, size_t rg_end, const std::string& pattern, bool neg,
                        uint8_t* bits, int64_t cap, float* kernel_ms) {
    if (col < 0 || col >= static_cast<int>(r.num_columns())) throw std::runtime_error("Invalid column index");
    const ColumnInfo& ci = r.column(static_cast<size_t>(col));
    if (ci.type != ParquetType::BYTE_ARRAY)
        throw std::runtime_error("Column '" + ci.name + "' is not BYTE_ARRAY (type: " + parquet_type_name(ci.type) + ")");
    DfaGuard g;
    char err[512] = {0};
    if (pqg_regex_compile(pattern.c_str(), &g.d, err, sizeof(err)) != PQG_OK) throw std::runtime_error(err);
    ParquetReader::DevicePlan dp = r.device_plan_range(col, rg_begin, rg_end, true);
    if (kernel_ms) *kernel_ms = 0;
    if (!dp.plan) return 0;
    std::vector<uint32_t> words((static_cast<size_t>(dp.n_pages) + 31) / 32 + 1, 0);
    if (pqg_regex_scan(dp.ctx, dp.plan, g.d, neg ? 1 : 0, words.data(), kernel_ms) != PQG_OK)
        throw std::runtime_error(pqg_last_error(dp.ctx));
    for (int64_t p = 0; p < static_cast<int64_t>(dp.n_pages) && p < cap; p++) bits[p] = (words[static_cast<size_t>(p) >> 5] >> (p & 31)) & 1u;
    return static_cast<int64_t>(dp.n_pages);
}

ChunkIndexJob::~ChunkIndexJob() { if (job) pqg_chunk_job_free(ctx, job); }

ChunkIndexJob* chunk_index_prepare_rgs(ParquetReader& r, const std::string& col_name, size_t rg_begin, size_t rg_end, uint64_t chunk_size) {
    int col = string_column(r, col_name);
    ParquetReader::DevicePlan dp = r.device_plan_range(col, rg_begin, rg_end, true);
    auto j = std::make_unique<ChunkIndexJob>();
    j->ctx = dp.ctx;
    if (!dp.plan) return j.release(); // no values: the open chunk only
    pqg_ctx_set_profiling(dp.ctx, 1);
    if (pqg_plan_run(dp.ctx, dp.plan) != PQG_OK) throw std::runtime_error(pqg_last_error(dp.ctx));
    pqg_page_error pe;
    if (pqg_plan_finish(dp.ctx, dp.plan, &pe) != PQG_OK) throw std::runtime_error(pqg_last_error(dp.ctx));
    pqg_timings tm;
    if (pqg_plan_timings(dp.plan, &tm) == PQG_OK) j->decode_ms = tm.total_ms;
    j->num_slots = pqg_plan_num_slots(dp.plan);
    if (pqg_chunk_index_prepare(dp.ctx, dp.plan, chunk_size, &j->job, &j->prepare_ms) != PQG_OK)
        throw std::runtime_error(pqg_last_error(dp.ctx));
    return j.release();
}

int64_t chunk_index_stitch(ChunkIndexJob& j, uint64_t carry_in, uint64_t* carry_out) {
    if (carry_out) *carry_out = carry_in;
    if (!j.job) return 1;
    uint64_t n_chunks = 0, carry = 0;
    if (pqg_chunk_index_stitch(j.ctx, j.job, carry_in, &n_chunks, &carry) != PQG_OK) throw std::runtime_error(pqg_last_error(j.ctx));
    if (carry_out) *carry_out = carry;
    return static_cast<int64_t>(n_chunks);
}

void chunk_index_emit(ChunkIndexJob& j, uint32_t id_base, uint32_t* ids, int64_t cap, float* kernel_ms) {
    if (kernel_ms) *kernel_ms = 0;
    if (!j.job) return;
    if (ids && static_cast<uint64_t>(cap) < j.num_slots) throw std::runtime_error("chunk_index: output buffer too small");
    if (pqg_chunk_index_emit(j.ctx, j.job, id_base, ids, kernel_ms) != PQG_OK) throw std::runtime_error(pqg_last_error(j.ctx));
}

int64_t chunk_index_rgs(ParquetReader& r, const std::string& col_name, size_t rg_begin, size_t rg_end, uint64_t chunk_size,
                        uint64_t carry_in, uint32_t id_base, uint32_t* ids, int64_t cap, uint64_t* carry_out) {
    std::unique_ptr<ChunkIndexJob> j(chunk_index_prepare_rgs(r, col_name, rg_begin, rg_end, chunk_size));
    if (j->job && static_cast<uint64_t>(cap) < j->num_slots) throw std::runtime_error("chunk_index: output buffer too small");
    const int64_t n_chunks = chunk_index_stitch(*j, carry_in, carry_out);
    chunk_index_emit(*j, id_base, ids, cap, nullptr);
    return n_chunks;
}

int64_t chunk_index(ParquetReader& r, const std::string& col_name, uint64_t chunk_size, uint64_t* tuple_to_chunk, int64_t num_rows) {
    string_column(r, col_name);
    for (int64_t i = 0; i < num_rows; i++) tuple_to_chunk[i] = 0;
    std::vector<uint32_t> ids(static_cast<size_t>(r.num_rows()) + 1);
    uint64_t carry = 0;
    int64_t n_chunks = chunk_index_rgs(r, col_name, 0, r.num_row_groups(), chunk_size, 0, 0, ids.data(), static_cast<int64_t>(ids.size()), &carry);
    for (int64_t i = 0; i < num_rows && i < r.num_rows(); i++) tuple_to_chunk[i] = ids[static_cast<size_t>(i)];
    return n_chunks;
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
