// pq_scan.cpp -- see pq_scan.hpp.  (regex / chunk-index drivers: implemented next)
#include "pq_scan.hpp"

#include <stdexcept>

namespace pqg {
int64_t regex_prune(ParquetReader&, int, const std::string&, bool, uint8_t*, int64_t, float*) {
    throw std::runtime_error("regex_prune: not implemented yet");
}
int64_t chunk_index(ParquetReader&, const std::string&, uint64_t, uint64_t*, int64_t) {
    throw std::runtime_error("chunk_index: not implemented yet");
}
int64_t page_chunk_index(ParquetReader&, int, uint64_t, uint32_t*, uint32_t*, uint32_t*, int64_t, int64_t*, int64_t*) {
    throw std::runtime_error("page_chunk_index: not implemented yet");
}
} // namespace pqg
