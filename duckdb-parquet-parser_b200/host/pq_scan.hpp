// pq_scan.hpp -- host drivers of the page-pruning scan and the 4 KB chunk indexes on top
// of the C-ABI (pqg_regex_scan / pqg_chunk_index / pqg_page_chunk_index in include/pqg.h).
//   regex_prune       parser --regex-column/--regex/--neg-regex   reference README.md:54-64
//   chunk_index       tuple-level chunk map                        reference src/main.cpp:21-32
//   page_chunk_index  index_test                                   reference README.md:66-72
#pragma once
#include <cstdint>
#include <string>

#include "pq_reader.hpp"

namespace pqg {

// bits[p] (one byte per data page of the column, global page order) = 1 when some non-null
// value v of page p satisfies (neg ? !match(v) : match(v)).  Returns the page count.
int64_t regex_prune(ParquetReader& r, int col, const std::string& pattern, bool neg, uint8_t* bits,
                    int64_t cap, float* kernel_ms);

// tuple_to_chunk[pos] for every non-null string in row order (nulls stay 0); returns the
// reference's "Total chunks".
int64_t chunk_index(ParquetReader& r, const std::string& col_name, uint64_t chunk_size,
                    uint64_t* tuple_to_chunk, int64_t num_rows);

// ── multi-GPU shards: the same operations over the row groups [rg_begin, rg_end) ──────────
// bits[] covers only the shard's pages (concatenate the shards in order to get the column's
// bitmap: pages are numbered in row-group order).
int64_t regex_prune_rgs(ParquetReader& r, int col, size_t rg_begin, size_t rg_end, const std::string& pattern, bool neg,
                        uint8_t* bits, int64_t cap, float* kernel_ms);
// tuple_to_chunk_local[i] for the shard's rows (shard-local chunk ids: 0 = the chunk left open by
// the previous shard); carry_in = bytes in that open chunk.  Returns the shard's chunk count
// (ids used = 0 .. count-1); the next shard's ids start at (sum of counts - 1 per shard).
int64_t chunk_index_rgs(ParquetReader& r, const std::string& col_name, size_t rg_begin, size_t rg_end, uint64_t chunk_size,
                        uint64_t carry_in, uint32_t id_base, uint32_t* tuple_to_chunk_local, int64_t cap, uint64_t* carry_out);

// The same in phases (include/pqg.h: pqg_chunk_index_prepare / _stitch / _emit): `prepare` uploads and decodes the
// shard and runs everything that does not need the previous shard's carry; only `stitch` is ordered between shards.
struct ChunkIndexJob {
    pqg_ctx* ctx = nullptr;
    pqg_chunk_job* job = nullptr;   // null: the shard has no values
    uint64_t num_slots = 0;
    float decode_ms = 0, prepare_ms = 0;
    ~ChunkIndexJob();
};
ChunkIndexJob* chunk_index_prepare_rgs(ParquetReader& r, const std::string& col_name, size_t rg_begin, size_t rg_end, uint64_t chunk_size);
// returns the shard's chunk count (local ids 0 .. count-1)
int64_t chunk_index_stitch(ChunkIndexJob& j, uint64_t carry_in, uint64_t* carry_out);
void chunk_index_emit(ChunkIndexJob& j, uint32_t id_base, uint32_t* ids, int64_t cap, float* kernel_ms);

int64_t page_chunk_index(ParquetReader& r, int col, uint64_t chunk_size, uint32_t* page_chunk,
                         uint32_t* page_off, uint32_t* chunk_first_page, int64_t cap,
                         int64_t* first_global_page, int64_t* n_col_pages);

} // namespace pqg
