// pqg_internal.h -- shared between the CUDA kernels (pqg_decode.cu, pqg_scan.cu) and the
// C-ABI implementation (pqg_api.cu).  Not installed; the public surface is include/pqg.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pqg.h"

namespace pqg {

constexpr int kWarpsPerCta = 8;
constexpr int kThreadsPerCta = kWarpsPerCta * 32;
constexpr int kSlotBytes = 2048;          // largest page payload staged in shared memory
constexpr int kSlotAlloc = kSlotBytes + 48; // + misalignment (<=15) + over-read padding
constexpr int kIdxWords = 512;            // per-warp index scratch: 1024 x u16 or 512 x u32
constexpr int kTileNarrow = 1024;         // slots per tile when indices fit 16 bits
constexpr int kTileWide = 512;
constexpr int kMaxSmemDictBytes = 32 * 1024; // dictionaries up to this size are staged per CTA
constexpr int kImagePad = 64;             // readable bytes required past the image end

// Device-side chunk record (built by the host API from pqg_chunk_desc).
struct DevChunk {
    uint64_t dict_off;       // image offset of the dictionary payload
    uint64_t out_row_base;
    uint64_t num_values;
    uint64_t dict_arena_off; // byte offset of this chunk's prepared dictionary in the arena
    uint64_t char_base;      // BYTE_ARRAY: first byte of this chunk in `chars` (written by the scan)
    uint32_t dict_size;
    uint32_t dict_n;
    uint32_t first_page;
    uint32_t n_pages;
    int16_t max_def;
    int16_t max_rep;
    uint8_t phys_type;
    uint8_t has_dict;
    uint8_t def_bw;
    uint8_t rep_bw;
    uint32_t dict_ok_n;      // dictionary entries that parsed (written by the prepare kernel)
};

// First failing page (lowest page-table index) of a run.
struct DevErr {
    unsigned long long key;  // (page << 32) | code, atomicMin; ~0 = none
    uint32_t count;
    uint32_t d_page, d_pos, d_need, d_size; // details, valid when d_page == key >> 32
    uint32_t pad;
};

struct DecodeParams {
    const uint8_t* image;
    uint64_t image_size;
    DevChunk* chunks;
    const pqg_page_desc* pages;
    uint32_t page_begin;     // page-table range handled by this launch
    uint32_t page_end;
    uint32_t pages_per_cta;
    uint32_t n_chunks;
    uint8_t* dict_arena;
    uint8_t* values;
    uint32_t* validity;
    uint32_t* offsets;
    uint8_t* chars;
    uint32_t* page_chars;    // BYTE_ARRAY pass 1 output: string bytes per page
    uint32_t* page_char_base;// exclusive prefix inside the chunk
    DevErr* err;
};

// launchers (pqg_decode.cu)
cudaError_t launch_dict_prepare(const DecodeParams& p, uint32_t n_chunks, int width, cudaStream_t s);
cudaError_t launch_decode_fixed(const DecodeParams& p, int width, bool boolean_plain, int sm_count, cudaStream_t s);
cudaError_t launch_str_sizes(const DecodeParams& p, int sm_count, cudaStream_t s);
cudaError_t launch_str_scan(const DecodeParams& p, uint64_t* total_chars, cudaStream_t s);
cudaError_t launch_str_copy(const DecodeParams& p, int sm_count, cudaStream_t s);
size_t decode_smem_bytes(bool with_dict);

// regex / chunk index (pqg_scan.cu)
struct DevDfa {
    const uint8_t* cls;   // 256 byte -> class
    const uint16_t* trans;// [n_states][n_classes]
    const uint8_t* accept;// [n_states] 1 = accepting at end of input
    uint32_t n_states, n_classes, start;
};
cudaError_t launch_regex_scan(const DecodeParams& p, const DevDfa& dfa, int neg, uint32_t* dict_match,
                              uint32_t* page_bits, int sm_count, cudaStream_t s, uint32_t* launches);
cudaError_t launch_chunk_index(const uint32_t* offsets_or_null, const DecodeParams& p, uint64_t n_slots,
                               uint64_t chunk_size, uint64_t carry_in, uint32_t* tuple_to_chunk,
                               uint64_t* scratch, uint64_t* result, cudaStream_t s, uint32_t* launches);

} // namespace pqg
