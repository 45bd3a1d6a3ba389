// pq_regex.hpp -- regex -> DFA tables shared between the host compiler (pq_regex.cpp) and
// the scan driver that uploads them (csrc/pqg_scan.cu).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

struct pqg_dfa;

namespace pqg {

struct CompiledDfa {
    uint8_t cls[256];             // byte -> equivalence class
    uint32_t n_classes = 0;
    uint32_t n_states = 0;
    uint32_t start = 0;
    uint32_t dead = UINT32_MAX;   // absorbing non-accepting state, if any
    std::vector<uint16_t> trans;  // [n_states][n_classes]; state 0 is the absorbing ACCEPT
    std::vector<uint8_t> accept;  // [n_states]: accepting when the input ends here
};

// throws std::runtime_error with an explicit reason for unsupported / malformed patterns
CompiledDfa compile_regex(const std::string& pattern, uint32_t max_states);
bool dfa_match(const CompiledDfa& d, const uint8_t* text, uint64_t len);
const CompiledDfa& dfa_tables(const pqg_dfa* d);

} // namespace pqg
