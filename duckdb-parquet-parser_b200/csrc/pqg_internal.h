// pqg_internal.h -- shared between the CUDA kernels (pqg_decode.cu, pqg_scan.cu) and the
// C-ABI implementation (pqg_api.cu).  Not installed; the public surface is include/pqg.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pqg.h"

// Debug build (make EXTRA=-DPQG_DEBUG_ASSERTS): device-side checks of the shared / global indices the kernels compute
// (compute-sanitizer is not available on the GPU pool; the GPU suite is run once under this build, see profiles/README.md)
#if defined(PQG_DEBUG_ASSERTS) && defined(__CUDACC__)
#include <cstdio>
#define PQG_ASSERT(c) do { if (!(c)) { printf("PQG_ASSERT failed: %s  (%s:%d, block %d thread %d)\n", #c, __FILE__, __LINE__, blockIdx.x, threadIdx.x); __trap(); } } while (0)
#else
#define PQG_ASSERT(c) ((void)0)
#endif

namespace pqg {

constexpr int kWarpsPerCta = 8;
constexpr int kThreadsPerCta = kWarpsPerCta * 32;
constexpr int kSlotBytes = 2048;          // largest page payload staged in shared memory
constexpr int kSlotAlloc = kSlotBytes + 48; // + misalignment (<=15) + over-read padding
constexpr int kIdxWords = 512;            // per-warp index scratch: 1024 x u16 or 512 x u32
constexpr int kTileNarrow = 1024;         // slots per tile when indices fit 16 bits
constexpr int kTileWide = 512;
constexpr int kMaxSmemDictBytes = 32 * 1024; // dictionaries up to this size are staged per CTA
// Dictionaries beyond that and up to kPartMaxParts x kPartDictBytes: PARTITIONED mode of the tile kernel -- 2^k CTAs read
// the same tiles, each keeps 1 / 2^k of the dictionary in its shared memory and emits only the values whose index falls
// into its part (local ld.shared instead of one L1TEX wavefront + one 32-byte L2 sector per gathered value)
constexpr int kPartDictBytes = 128 * 1024;
constexpr int kPartMaxParts = 4;
constexpr int kStageMaxLen = 48;          // strings up to this length go through the per-warp staging buffer
constexpr int kStageBytes32 = 32 * kStageMaxLen + 32; // 32 strings + alignment phase (multiple of 16)
constexpr int kImagePad = 64;             // readable bytes required past the image end
// TMA-staged tile pipeline of the fast fixed-width kernel (pqg_tiles.cu)
constexpr int kTileBytes = 8192;          // image bytes per tile (16-byte aligned range covering whole pages)
constexpr int kTileBytesLarge = 16384;     // tiles of OPTIONAL fixed-width plans (level bytes make their pages ~1.8 KB: 8 pages per tile)
constexpr int kTileBytesMid = 10240;       // string plans whose pages run just past 1 KB (eight of them per tile for the regex scan)
constexpr int kTilePages = 8;             // pages per tile (one per warp)
constexpr int kTileStages = 2;            // ring depth per CTA (measured: 2 x 8 KB beats 3-4 stages and 16 KB tiles: the shared-memory
                                          // carve-out eats L1, which the large-dictionary gathers and the PLAIN copy both feel)

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
    // BYTE_ARRAY dictionaries whose entries are all <= 15 bytes also get a 16-byte-per-entry table
    // (15 zero-padded bytes + the length) behind the {start, len} entries: one 16-byte load per value
    uint32_t dict_minlen;    // BYTE_ARRAY, written by the prepare kernels: shortest / longest dictionary entry
    uint32_t dict_maxlen;    //   (all <= 15 bytes: the padded table is valid; equal: one common length)
    uint64_t dict_pad_off;   // arena offset of the padded table
    uint32_t dict_seg_first; // BYTE_ARRAY: first record of this chunk in the plan's dictionary-segment scratch
    uint32_t pad_;
#ifdef __CUDACC__
    __device__ __forceinline__ bool dict_short() const { return dict_ok_n > 0 && dict_maxlen <= 15u; }
    // common length of all entries, ~0u if they differ
    __device__ __forceinline__ uint32_t dict_len() const { return (dict_ok_n > 0 && dict_minlen == dict_maxlen) ? dict_maxlen : 0xffffffffu; }
#endif
};
// one segment of a BYTE_ARRAY dictionary page (k_dict_seg / _link / _emit): up to kDictCand speculative entry positions,
// each with the end of its walk and its entry count; the link picks the one the chain from the left really enters at
constexpr int kDictCand = 4;
struct DictSeg { uint32_t start[kDictCand], end[kDictCand]; uint16_t cnt[kDictCand]; uint32_t base, pick; };
constexpr uint32_t kDictSeg = 128;                  // bytes per segment

struct FlatCtl { uint32_t scan_cursor, rank_cursor, emit_cursor, nblk, handed, pad; }; // (pqg_flat.cu) cursors of the flat kernels, blocks handed out, pages handed on

// First failing page (lowest page-table index) of a run, plus the per-run work counters
// (one memset resets everything).
struct DevErr {
    unsigned long long key;  // (page << 32) | code, atomicMin; ~0 = none
    uint32_t count;
    uint32_t d_page, d_pos, d_need, d_size; // details, valid when d_page == key >> 32
    uint32_t bad_index;      // out-of-range dictionary indices seen in REQUIRED chunks
    uint32_t pad;
    // work counters, reset before every (sub-)run: keep them last and together
    uint32_t slow_count;     // pages the tile kernel handed to the general kernel
    uint32_t slow_cursor;    // work-stealing cursor of the general kernel
    FlatCtl flat;
};
constexpr size_t kWorkCounterBytes = 8 + sizeof(FlatCtl); // slow_count .. the end

// Flat decode of oversized / OPTIONAL pages (pqg_flat.cu): per work-list entry, and the cursors of its three kernels
// (reset with the other work counters of a (sub-)run).
struct FlatPage { uint32_t status, nn, vpos, bw, blk0, page; }; // status 0: decode; vpos: values / first index-run header inside the payload
struct __align__(16) FlatBlk { // one 1024-slot block, everything its warp needs in one read (written by k_flat_ranks)
    uint64_t src;      // image offset: PLAIN -- the block's first value; dictionary page -- the index stream (behind the bit-width byte)
    uint64_t row;      // output slot of the block's first slot
    uint32_t nslots;   // 1 .. 1024; 0: nothing to do (the page went to the general kernel)
    uint32_t rank0;    // values of the page in front of the block
    uint32_t cp_pos, cp_first; // dictionary page: the checkpoint in front of value rank0 {run header position, first value of that run}
    uint32_t slen;     // dictionary page: length of the index stream
    uint32_t info;     // bits 0..5: index bit width; bit 8: dictionary page; bit 9: no validity to read (REQUIRED chunk)
    uint32_t chunk;
    uint32_t pad;
};

// A run of consecutive pages of one chunk whose bytes are staged with one bulk copy.
struct __align__(32) TileDesc {
    uint64_t byte_lo;        // image offset, multiple of 16
    uint32_t byte_len;       // multiple of 16, <= kTileBytes
    uint32_t first_page;
    uint32_t n_pages;        // <= kTilePages
    uint32_t chunk_idx;
    uint32_t pad[2];
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
    DictSeg* dict_segs;      // BYTE_ARRAY plans: scratch of the dictionary preparation
    uint8_t* values;
    uint32_t* validity;
    uint32_t* offsets;
    uint8_t* chars;
    uint32_t* page_chars;    // BYTE_ARRAY pass 1 output: string bytes per page
    uint32_t* page_char_base;// exclusive prefix inside the chunk
    DevErr* err;
    // fixed-width plans: fast tiles + the list of pages for the general kernel
    const TileDesc* tiles;
    uint32_t tile_lo, tile_hi; // tile range of this launch (a chunk range of the plan)
    uint32_t tiles_per_cta;
    uint32_t dict_smem;      // bytes of shared memory reserved for a staged dictionary
    uint32_t slow_lo, slow_hi; // host-listed slow pages of this launch: slow_pages[slow_lo, slow_hi)
    uint32_t* slow_pages;    // host-listed slow pages (n_slow_host entries)
    uint32_t* slow_append;   // pages handed over on the device: slow_append[0 .. err->slow_count); capacity: every page of the plan
    uint32_t slow_cap;       // ... that capacity (debug asserts)
    uint64_t n_slots;        // output slots of the plan (debug asserts)
    uint32_t chunk_lo;       // first chunk of this launch (dictionary preparation)
    uint32_t tile_sync;      // 1: CTA-wide barrier per tile instead of the last-warp refill (the regex scan: issue bound)
    uint32_t opt_idx;        // OPTIONAL plans with foreign-looking pages (> 1024 slots / > 4 KB): the tile kernel carries a per-warp index buffer
    uint32_t part_bits;      // partitioned-dictionary launch: log2 of the CTAs that share a tile span (0: off)
    uint32_t part_entries;   // ... dictionary entries per part (a power of two)
    uint32_t handover_hint;  // pages the tile kernel handed to the general kernel in the previous run of the plan (~0u: unknown)
    uint32_t skip_dict_pad;  // 1: the run does not materialise strings (regex scan): no padded short-string table
    uint32_t tile_bytes;     // tile size the plan's tiles were cut for (kTileBytes / kTileBytesLarge)
    uint32_t identity_dict;  // dictionary-form output: emit the dictionary INDEX of every slot instead of the entry
    uint32_t check_layout;   // BYTE_ARRAY copy pass: page byte counts came from the page headers -- verify them
    uint32_t exact_sizes;    // BYTE_ARRAY size pass: no optimistic per-page byte counts (a page broke its count in an earlier run)
    // flat decode of the slow list (plans with OPTIONAL chunks or host-listed pages, 4/8-byte values): null when off
    FlatCtl* flat;
    FlatPage* flat_pages;    // per work-list entry (capacity slow_cap)
    FlatBlk* flat_blk;       // per 1024-slot block of a listed page
    uint2* flat_ckpt;        // per 1024 VALUES of a listed dictionary page: {run header position, first value of that run}
    uint32_t flat_blk_cap;
    uint32_t* flat_append;   // pages the flat kernels hand to the general kernel: flat_append[0 .. flat->handed)
    const uint32_t* append_count; // general kernel: number of device-appended pages behind slow_append (null: err->slow_count)
    uint64_t chars_cap;      // BYTE_ARRAY copy pass launched without a host sync: capacity of `chars` ...
    const uint64_t* total_chars; // ... and the device-side grand total it must hold (null: the host already checked)
};

// Rewrite of DATA_PAGE_V2 / SNAPPY pages into the layout the decode kernels know (pqg_ext.cu): one record per dictionary
// page and per data page of a plan created by pqg_plan_create_ext
constexpr uint32_t kXformV2 = 1u, kXformPrefix = 2u; // XformRec::kind bits (bits 8..15: PQG_CODEC_* of the compressed part)
constexpr uint32_t kXformSynthBytes = 10u;           // length word + the one RLE run written for a V2 page of an OPTIONAL column without level bytes
struct XformRec {
    uint64_t src_off, dst_off; // in the caller's image / in the plan's image
    uint32_t src_size, dst_size;
    uint32_t kind, def_len, rep_len, num_values;
    uint32_t page;             // page-table index for error reports (dictionary pages: their chunk's first page)
    uint32_t pad;
};
struct DevErr;
cudaError_t launch_xform(const uint8_t* src, uint8_t* dst, const XformRec* recs, uint32_t n, DevErr* err, int sm_count, cudaStream_t s);

// launchers (pqg_decode.cu)
// width 0 = BYTE_ARRAY: three launches (segments, link, emit), max_dict_blocks = blocks of 256 threads per dictionary
cudaError_t launch_dict_prepare(const DecodeParams& p, uint32_t n_chunks, int width, uint32_t max_dict_blocks, cudaStream_t s);
inline uint32_t dict_prepare_launches(int width) { return width == 0 ? 3u : 1u; }
cudaError_t launch_decode_fixed(const DecodeParams& p, int width, bool boolean_plain, int sm_count, cudaStream_t s);
// fast path (pqg_tiles.cu): PLAIN / regular-dictionary pages of REQUIRED 4- and 8-byte chunks
cudaError_t launch_fixed_tiles(const DecodeParams& p, int width, int sm_count, cudaStream_t s);
cudaError_t launch_fixed_tiles_part(const DecodeParams& p, int width, int sm_count, cudaStream_t s); // partitioned dictionary: ONE chunk per launch
bool chunk_is_tileable(int phys_type, int max_def, int max_rep);
// the slow list of 4/8-byte plans as a flat list of 1024-slot blocks (pqg_flat.cu); what it cannot take goes to flat_append
// idx16: no dictionary of the plan has more than 65535 entries (the emission keeps 16-bit indices: 4 CTAs per SM instead of 3)
cudaError_t launch_flat_pages(const DecodeParams& p, int width, int sm_count, bool any_dict, bool idx16, cudaStream_t s);
uint32_t flat_launches();
cudaError_t launch_str_sizes(const DecodeParams& p, int sm_count, cudaStream_t s);
cudaError_t launch_str_scan(const DecodeParams& p, uint64_t* total_chars, cudaStream_t s);
cudaError_t launch_str_copy(const DecodeParams& p, bool any_dict, int sm_count, cudaStream_t s);
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
