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

int64_t page_chunk_index(ParquetReader& r, int col, uint64_t chunk_size, uint32_t* page_chunk,
                         uint32_t* page_off, uint32_t* chunk_first_page, int64_t cap,
                         int64_t* first_global_page, int64_t* n_col_pages);

} // namespace pqg
