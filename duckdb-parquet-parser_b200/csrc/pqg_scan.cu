// pqg_scan.cu -- regex page scan and chunk-index kernels (implemented next).
#include <cstdio>
#include "pqg_internal.h"
extern "C" {
int pqg_regex_scan(pqg_ctx*, pqg_plan*, const pqg_dfa*, int, uint32_t*, float*) { return PQG_ERR_UNSUPPORTED; }
int pqg_chunk_index(pqg_ctx*, pqg_plan*, uint64_t, uint64_t, uint32_t*, uint64_t*, uint64_t*, float*) { return PQG_ERR_UNSUPPORTED; }
int pqg_page_chunk_index(pqg_ctx*, const uint32_t*, uint32_t, uint64_t, uint32_t*, uint32_t*, uint32_t*, uint32_t, uint32_t*) { return PQG_ERR_UNSUPPORTED; }
}
