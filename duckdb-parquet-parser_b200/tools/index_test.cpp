// index_test -- "Groups a column's data pages into 4 KB chunks and builds an inverted index
// that maps byte offsets back to source pages" (reference README.md:66-72; its source is
// absent from the checkout -- frozen spec: SURVEY.md section 8 a-20).  The greedy packing
// runs on the GPU (pqg_page_chunk_index); lookups are a binary search over the result.
//
//   index_test <parquet_file> <column_name> [chunk_size]
#include <algorithm>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "pq_reader.hpp"
#include "pq_scan.hpp"

using namespace pqg;

int main(int argc, char* argv[]) {
    if (argc < 3) { std::cerr << "usage: index_test <parquet_file> <column_name> [chunk_size]\n"; return 2; }
    const uint64_t chunk_size = argc > 3 ? std::strtoull(argv[3], nullptr, 10) : 4096;
    try {
        ParquetReader reader;
        if (!reader.open(argv[1])) return 1;
        int col = reader.find_column(argv[2]);
        if (col < 0) throw std::runtime_error(std::string("Column not found: ") + argv[2]);
        const size_t cap = reader.num_pages() + 1;
        std::vector<uint32_t> page_chunk(cap), page_off(cap), chunk_first(cap);
        int64_t first = 0, n_pages = 0;
        int64_t n_chunks = page_chunk_index(reader, col, chunk_size, page_chunk.data(), page_off.data(), chunk_first.data(),
                                            static_cast<int64_t>(cap), &first, &n_pages);
        std::cout << "Column: " << argv[2] << "\nData pages: " << n_pages << " (first global page id " << first << ")\n";
        std::cout << "Chunk size: " << chunk_size << "\nTotal chunks: " << n_chunks << "\n";
        // self-check of the inverted index: every page is found again from (chunk, byte offset)
        int64_t bad = 0;
        for (int64_t p = 0; p < n_pages; p++) {
            uint32_t c = page_chunk[static_cast<size_t>(p)];
            int64_t lo = chunk_first[c], hi = (c + 1 < n_chunks ? chunk_first[c + 1] : n_pages);
            // last page of the chunk whose offset is <= the byte offset
            auto it = std::upper_bound(page_off.begin() + lo, page_off.begin() + hi, page_off[static_cast<size_t>(p)]);
            int64_t found = (it - page_off.begin()) - 1;
            // zero-size pages share an offset with their successor: accept any page at that offset
            if (page_off[static_cast<size_t>(found)] != page_off[static_cast<size_t>(p)]) bad++;
        }
        std::cout << "Lookup self-check: " << (bad ? "FAILED" : "ok") << "\n";
        for (int64_t c = 0; c < std::min<int64_t>(n_chunks, 8); c++) {
            int64_t lo = chunk_first[static_cast<size_t>(c)], hi = (c + 1 < n_chunks ? chunk_first[static_cast<size_t>(c) + 1] : n_pages);
            std::cout << "  chunk " << c << ": pages [" << lo + first << ", " << hi + first << ")\n";
        }
        return bad ? 1 : 0;
    } catch (const std::exception& e) {
        std::cerr << "Error: " << e.what() << std::endl;
        return 1;
    }
}
