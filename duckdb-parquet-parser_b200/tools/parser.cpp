// parser -- the reference's CLI (README.md:46-64) on top of the GPU decoder.
//
//   parser <file>                                              schema, row groups, page sizes
//   parser <file> --regex-column <col> --regex <pat> [--neg-regex]
//                                                              pages with no value satisfying the predicate
//   parser <file> --chunk-index <col> [--chunk-size N]         the prototype of src/main.cpp:10-37
//
// The reference's own main() ignores argv (src/main.cpp:10-12); the flags are the ones its
// README documents.  Every value-level operation runs on the GPU (no CPU fallback).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "pq_reader.hpp"
#include "pq_scan.hpp"

using namespace pqg;

static int usage() {
    std::cerr << "usage: parser <parquet_file> [--regex-column <column> --regex <pattern> [--neg-regex]]\n"
                 "       parser <parquet_file> --chunk-index <column> [--chunk-size <bytes>]\n";
    return 2;
}

int main(int argc, char* argv[]) {
    if (argc < 2) return usage();
    std::string file = argv[1], regex_col, pattern, chunk_col;
    bool neg = false, have_pattern = false;
    uint64_t chunk_size = 4096;
    for (int i = 2; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char* { if (i + 1 >= argc) { usage(); std::exit(2); } return argv[++i]; };
        if (a == "--regex-column") regex_col = next();
        else if (a == "--regex") { pattern = next(); have_pattern = true; }
        else if (a == "--neg-regex") neg = true;
        else if (a == "--chunk-index") chunk_col = next();
        else if (a == "--chunk-size") chunk_size = std::strtoull(next(), nullptr, 10);
        else return usage();
    }
    try {
        ParquetReader reader;
        if (!reader.open(file)) return 1;
        if (!regex_col.empty() || have_pattern) {
            if (regex_col.empty() || !have_pattern) return usage();
            int col = reader.find_column(regex_col);
            if (col < 0) throw std::runtime_error("Column not found: " + regex_col);
            std::vector<uint8_t> bits(reader.num_pages() + 1);
            float ms = 0;
            int64_t n = regex_prune(reader, col, pattern, neg, bits.data(), static_cast<int64_t>(bits.size()), &ms);
            int64_t first = static_cast<int64_t>(reader.num_row_groups() ? reader.first_page_id(0, static_cast<size_t>(col)) : 0);
            int64_t pruned = 0;
            std::cout << "Column: " << regex_col << "  pattern: " << pattern << (neg ? "  (negated)" : "") << "\n";
            std::cout << "Pages with no matching value (column-local id / global id):\n";
            // global ids: data pages are numbered (row group, column, page); walk the index
            std::vector<size_t> global;
            size_t want = static_cast<size_t>(reader.column(static_cast<size_t>(col)).column_index);
            for (size_t g = 0; g < reader.num_pages(); g++) if (reader.page_index_entry(g).column_idx == want) global.push_back(g);
            for (int64_t p = 0; p < n; p++) {
                if (bits[static_cast<size_t>(p)]) continue;
                pruned++;
                std::cout << "  " << p << " / " << global[static_cast<size_t>(p)] << "\n";
            }
            (void)first;
            std::cout << "Pages scanned: " << n << "\nPages prunable: " << pruned << "\nGPU scan time: " << ms << " ms\n";
            return 0;
        }
        if (!chunk_col.empty()) {
            std::vector<uint64_t> t2c(static_cast<size_t>(reader.num_rows()) + 1);
            int64_t chunks = chunk_index(reader, chunk_col, chunk_size, t2c.data(), reader.num_rows());
            std::cout << "Total tuples: " << reader.num_rows() << std::endl;
            std::cout << "Total chunks: " << chunks << std::endl;
            return 0;
        }
        // schema and page layout
        std::cout << reader.schema_string();
        std::cout << "Rows: " << reader.num_rows() << "  Row groups: " << reader.num_row_groups() << "  Data pages: " << reader.num_pages() << "\n";
        for (size_t rg = 0; rg < reader.num_row_groups(); rg++) {
            const RowGroup& g = reader.metadata().row_groups[rg];
            std::cout << "Row group " << rg << ": " << g.num_rows << " rows, " << g.total_byte_size << " bytes\n";
            for (size_t c = 0; c < reader.num_columns(); c++) {
                const ColumnInfo& ci = reader.column(c);
                const auto& pages = reader.chunk_pages(rg, static_cast<size_t>(ci.column_index));
                uint64_t bytes = 0, data_pages = 0, dict_pages = 0;
                for (const PageRecord& p : pages) {
                    bytes += p.payload_size;
                    if (p.type == PageType::DICTIONARY_PAGE) dict_pages++; else if (p.counted) data_pages++;
                }
                std::cout << "  " << ci.name << " (" << ci.type_name() << "): " << data_pages << " data pages, " << dict_pages
                          << " dictionary pages, " << bytes << " payload bytes\n";
            }
        }
        return 0;
    } catch (const std::exception& e) {
        std::cerr << "Error: " << e.what() << std::endl;
        return 1;
    }
}
