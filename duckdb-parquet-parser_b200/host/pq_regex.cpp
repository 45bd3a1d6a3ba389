// pq_regex.cpp -- host regex -> DFA compiler (implemented next).
#include <cstdio>
#include <cstring>
#include "pqg.h"
extern "C" {
int pqg_regex_compile(const char*, pqg_dfa** out, char* err, size_t errlen) {
    if (out) *out = nullptr;
    if (err && errlen) std::snprintf(err, errlen, "regex compiler not implemented yet");
    return PQG_ERR_REGEX;
}
void pqg_dfa_free(pqg_dfa*) {}
uint32_t pqg_dfa_num_states(const pqg_dfa*) { return 0; }
int pqg_dfa_match_host(const pqg_dfa*, const uint8_t*, uint64_t) { return -1; }
}
