// pqg_scan.cu -- regex page-pruning scan and the greedy chunk indexes.
//
//   k_regex_dict     one lane per dictionary entry: predicate of the entry
//   k_regex_tiles    the fast scan on the TMA tile pipeline, one warp per data page: OR over
//                    the page's non-null values of (neg ? !match : match); PLAIN pages: parallel
//                    length-prefix discovery, then the DFA (absolute-address table in shared
//                    memory) with one lane per string; dictionary pages look the predicate of
//                    their indices up.
//   k_regex_pages    the same for every other page shape (slow list, work stealing).
//                    Together they replace the parser's --regex-column/--regex/--neg-regex mode
//                    (reference README.md:54-64; source absent, frozen spec SURVEY.md 8 a-19).
//   chain engine     greedy ">= chunk_size closes the chunk" chunking over a weight sequence
//                    (src/main.cpp:21-32 tuple level; README.md:66-72 page level): device-wide
//                    exclusive scan of the weights, then the cut chain c' = first prefix value
//                    >= c + chunk_size.  The chain is sequential by nature; it is broken up by
//                    speculation: weight space is cut into tiles of 64 chunk sizes, every slot
//                    that can be the first cut of a tile (prefix in [tile, tile + chunk_size])
//                    walks the tile in parallel, the host stitches the tiles (one lookup per
//                    tile) and a last pass materialises cuts and chunk ids.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "pq_regex.hpp"
#include "pqg_tilepipe.cuh"

struct pqg_ctx;
struct pqg_plan;

namespace pqg {
// accessors implemented in pqg_api.cu
DecodeParams plan_params(const pqg_plan* p);
cudaStream_t ctx_stream(const pqg_ctx* c);
int ctx_sm_count(const pqg_ctx* c);
int ctx_device(const pqg_ctx* c);
void ctx_add_launches(pqg_ctx* c, uint32_t n);
int ctx_fail(pqg_ctx* c, int code, const std::string& m);
bool plan_is_str(const pqg_plan* p);
int plan_width(const pqg_plan* p);
bool plan_ran(const pqg_plan* p);
bool plan_run_pending(const pqg_plan* p);
bool plan_regex_tile_sync(const pqg_plan* p);
bool plan_any_dict(const pqg_plan* p);
uint64_t plan_slots(const pqg_plan* p);
const std::vector<pqg_chunk_desc>& plan_chunks(const pqg_plan* p);
const std::vector<pqg_page_desc>& plan_pages(const pqg_plan* p);
uint32_t plan_max_page_values(const pqg_plan* p);
size_t plan_dict_arena_bytes(const pqg_plan* p);
uint32_t plan_str_dict_blocks(const pqg_plan* p);

namespace {

// ---------------------------------------------------------------------------------------------
// regex
// ---------------------------------------------------------------------------------------------
struct DfaDev {
    const uint16_t* trans; // wide: [n_states][256]; else [n_states][n_classes]
    const uint8_t* cls;    // byte -> class (class-indexed tables only)
    const uint8_t* accept; // [n_states]
    uint32_t n_states, n_classes, start, dead;
    uint32_t wide;         // transition table indexed by the byte itself
    uint32_t table_bytes;  // bytes of trans
    uint32_t in_smem;      // tables are staged in shared memory
    uint32_t scaled;       // wide table staged in shared memory: the word-wise fast DFA loop applies
};

// state 0 is the absorbing accept, `dead` the absorbing reject (host compiler, pq_regex.cpp)
__device__ __forceinline__ bool dfa_run(const DfaDev& D, const uint16_t* trans, const uint8_t* cls, const uint8_t* accept,
                                        const uint8_t* text, uint32_t len) {
    uint32_t s = D.start;
    if (D.wide) { // byte-indexed table, entries pre-scaled by 256
        const uint32_t dead = D.dead << 8;
        s <<= 8;
        for (uint32_t i = 0; i < len && s != 0 && s != dead; i++) s = trans[s + text[i]];
        return accept[s >> 8] != 0;
    } else {
        for (uint32_t i = 0; i < len && s != 0 && s != D.dead; i++) s = trans[s * D.n_classes + cls[text[i]]];
    }
    return accept[s] != 0;
}

// Byte-indexed table staged in shared memory with ABSOLUTE entries: every entry is the
// shared-space byte address of the next state's row (row = 256 x u16), so a step is
// "address = state + 2 * byte; state = ld.shared.u16 [address]".  Text comes from aligned words
// (one funnel shift per word realigns it), eight bytes per loop trip while the string lasts, the
// absorbing states are tested once per trip.
__device__ __forceinline__ uint32_t dfa_step(uint32_t s, uint32_t w, uint32_t k) {
    uint32_t nxt;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(nxt) : "r"(s + 2u * __byte_perm(w, 0u, 0x4440u + k)));
    return nxt;
}
template <class LD>
__device__ __forceinline__ uint32_t dfa_run_abs(uint32_t s, uint32_t acc_s, uint32_t dead_s, const LD& ld, uint32_t apos, uint32_t len) {
    uint32_t wi = apos >> 2;
    const uint32_t sh = (apos & 3u) * 8u;
    uint32_t w0 = ld(wi);
    uint32_t left = len;
    while (left >= 8u && s != acc_s && s != dead_s) {
        const uint32_t w1 = ld(wi + 1), w2 = ld(wi + 2);
        const uint32_t a = __funnelshift_r(w0, w1, sh), b = __funnelshift_r(w1, w2, sh);
        s = dfa_step(s, a, 0); s = dfa_step(s, a, 1); s = dfa_step(s, a, 2); s = dfa_step(s, a, 3);
        s = dfa_step(s, b, 0); s = dfa_step(s, b, 1); s = dfa_step(s, b, 2); s = dfa_step(s, b, 3);
        w0 = w2; wi += 2; left -= 8u;
    }
    if (left >= 8u) return s; // absorbed
    if (left >= 4u) {
        const uint32_t w1 = ld(wi + 1);
        const uint32_t a = __funnelshift_r(w0, w1, sh);
        s = dfa_step(s, a, 0); s = dfa_step(s, a, 1); s = dfa_step(s, a, 2); s = dfa_step(s, a, 3);
        w0 = w1; wi++; left -= 4u;
    }
    if (left) {
        const uint32_t a = __funnelshift_r(w0, ld(wi + 1), sh);
        s = dfa_step(s, a, 0);
        if (left > 1u) s = dfa_step(s, a, 1);
        if (left > 2u) s = dfa_step(s, a, 2);
    }
    return s;
}

struct RegexParams {
    DecodeParams P;
    DfaDev D;
    uint8_t* dict_match;  // per dictionary entry (arena entry index): predicate (neg applied)
    uint32_t* page_bits;  // bit per page-table entry
    int neg;
    uint32_t cand_cap;    // tile scan: length-prefix candidates per page (u16 positions in shared memory)
};

__global__ void __launch_bounds__(256) k_regex_dict(RegexParams R) {
    const DevChunk& ck = R.P.chunks[blockIdx.y];
    if (!ck.has_dict) return;
    const uint2* ent = reinterpret_cast<const uint2*>(R.P.dict_arena + ck.dict_arena_off);
    uint8_t* out = R.dict_match + ck.dict_arena_off / 8;
    const uint8_t* chars = R.P.image + ck.dict_off;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < ck.dict_ok_n; i += gridDim.x * blockDim.x) {
        uint2 e = ent[i];
        bool m = dfa_run(R.D, R.D.trans, R.D.cls, R.D.accept, chars + e.x, e.y);
        out[i] = (R.neg ? !m : m) ? 1 : 0;
    }
}

__device__ __forceinline__ bool regex_page(const RegexParams& R, const uint16_t* trans, const uint8_t* cls, const uint8_t* accept,
                                           uint32_t q, const DevChunk& ck, WarpScratch& ws) {
    const DecodeParams& P = R.P;
    const uint32_t l = lane_id();
    const pqg_page_desc pd = P.pages[q];
    if (pd.num_values == 0) return false;
    PageCtx c;
    if (!page_begin(P, q, pd, ck, ws, c)) return false;
    const uint8_t* vals = c.pg + c.vals_pos;
    const uint32_t vavail = c.size - c.vals_pos;
    const bool pwide = !c.dict && c.size > 65535u;
    const bool wide = c.dict ? c.wide : pwide;
    const uint32_t T_ = wide ? kTileWide : kTileNarrow;
    const bool single = c.n <= T_;
    const uint32_t dict_n = ck.dict_ok_n;
    const uint8_t* dmatch = R.dict_match + ck.dict_arena_off / 8;
    bool regular = false, hit = false;
    RegStream rs{};
    uint32_t nn_before = 0, wpos = 0;
    for (uint32_t ts = 0; ts < c.n; ts += T_) {
        const uint32_t t = min(T_, c.n - ts);
        uint32_t bad = 0;
        const uint32_t nn = levels_tile(c.defw, ws, t, ck.max_def, single, &bad);
        if (bad) { if (l == 0) report_error(P.err, q, bad); return false; }
        if (c.dict) {
            if (ts == 0 && (single || !c.has_def)) regular = check_regular2(c.idxw.s, c.idxw.len, c.bw, single ? nn : c.n, &rs);
            if (!regular) {
                indices_tile(c.idxw, ws, nn, c.wide, &bad);
                if (bad) { if (l == 0) report_error(P.err, q, bad); return false; }
                __syncwarp();
            }
            for (uint32_t k = l; k < nn; k += 32) {
                uint32_t ix = regular ? regular_index2(rs, nn_before + k) : idx_load(ws.idx, k, c.wide);
                if (ix < dict_n && dmatch[ix]) hit = true; // out-of-range index: no value (null / dropped)
            }
        } else {
            uint32_t epos = 0, eneed = 0;
            if (!walk_strings(vals, vavail, &wpos, nn, ws, wide, true, &epos, &eneed)) {
                if (l == 0) report_error(P.err, q, PQG_PAGE_TRUNCATED, c.vals_pos + epos, eneed, c.size);
                return false;
            }
            for (uint32_t k = l; k < nn; k += 32) {
                uint32_t pp = idx_load(ws.idx, k, wide);
                uint32_t len = ld32u(vals + pp);
                bool m = dfa_run(R.D, trans, cls, accept, vals + pp + 4, len);
                if (R.neg ? !m : m) hit = true;
            }
        }
        nn_before += nn;
        __syncwarp();
        if (__any_sync(0xffffffffu, hit)) return true; // the page bit is an OR: stop early
    }
    return __any_sync(0xffffffffu, hit);
}

// stage [trans][cls 256][accept] into shared memory when they fit.  `absolute`: byte-indexed
// entries (next state * 256, u16 units) become shared-space byte addresses of the next row
// (entry * 2 + address of the table); two entries per 32-bit word, no carry between the halves
// because every address stays below 64 KiB (checked by the launcher).
__device__ __forceinline__ void stage_tables(const RegexParams& R, uint8_t* stab, const uint16_t*& trans, const uint8_t*& cls,
                                             const uint8_t*& accept, bool absolute = false) {
    trans = R.D.trans; cls = R.D.cls; accept = R.D.accept;
    if (R.D.in_smem) {
        const uint32_t total = R.D.table_bytes + 256u + R.D.n_states;
        const uint8_t* src = reinterpret_cast<const uint8_t*>(R.D.trans); // the three tables are contiguous on the device
        const uint32_t rebase = static_cast<uint32_t>(__cvta_generic_to_shared(stab)) * 0x10001u;
        const uint32_t n_trans16 = absolute ? R.D.table_bytes / 16u : 0u;
        // 16-byte vectors (both sides are 16-byte aligned; the device blob is padded)
        for (uint32_t i = threadIdx.x; i < (total + 15u) / 16u; i += blockDim.x) {
            uint4 v = reinterpret_cast<const uint4*>(src)[i];
            if (i < n_trans16) { v.x = (v.x << 1) + rebase; v.y = (v.y << 1) + rebase; v.z = (v.z << 1) + rebase; v.w = (v.w << 1) + rebase; }
            reinterpret_cast<uint4*>(stab)[i] = v;
        }
        __syncthreads();
        trans = reinterpret_cast<const uint16_t*>(stab);
        cls = stab + R.D.table_bytes;
        accept = cls + 256;
    }
}

// The general scan: every page shape, pages taken from the slow list (appended by the fast
// kernel) through a work-stealing cursor.
__global__ void __launch_bounds__(kThreadsPerCta) k_regex_pages(RegexParams R) {
    extern __shared__ __align__(128) uint8_t smem[];
    WarpScratch& ws = reinterpret_cast<WarpScratch*>(smem)[warp_id()];
    const uint16_t* trans; const uint8_t* cls; const uint8_t* accept;
    stage_tables(R, smem + sizeof(WarpScratch) * kWarpsPerCta, trans, cls, accept);
    const DecodeParams& P = R.P;
    const uint32_t n_host = P.slow_hi - P.slow_lo;
    const uint32_t total = n_host + P.err->slow_count;
    for (;;) {
        uint32_t i = 0;
        if (lane_id() == 0) i = atomicAdd(&P.err->slow_cursor, 1u);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= total) break;
        const uint32_t q = i < n_host ? P.slow_pages[P.slow_lo + i] : P.slow_append[i - n_host];
        const DevChunk& ck = P.chunks[P.pages[q].chunk_idx];
        bool hit = regex_page(R, trans, cls, accept, q, ck, ws);
        if (hit && lane_id() == 0) atomicOr(&R.page_bits[q >> 5], 1u << (q & 31u));
        __syncwarp();
    }
}

// The fast scan: one warp per page for what the reference's writer emits -- pages that fit the
// shared-memory slot, flat columns, RLE-run definition levels, PLAIN strings (parallel
// length-prefix discovery, then one lane per string through the DFA) or single-group
// bit-packed dictionary indices (predicate looked up per index).  Everything else goes to
// the slow list.
__device__ __forceinline__ void rx_to_slow(const DecodeParams& P, uint32_t q) {
    uint32_t k = atomicAdd(&P.err->slow_count, 1u);
    PQG_ASSERT(k < P.slow_cap);
    P.slow_append[k] = q;
}

constexpr uint32_t kRxCand = 1024; // most length-prefix candidates per page the tile scan keeps (u16 positions); 512 when no page needs more

// shared memory: [tables, padded to 128 bytes][tile pipeline][candidates per warp]
__host__ __device__ inline uint32_t rx_table_pad(uint32_t table_bytes, uint32_t n_states, bool in_smem) {
    return in_smem ? ((table_bytes + 256u + n_states + 127u) & ~127u) : 0u;
}

template <int TB>
__global__ void __launch_bounds__(kThreadsPerCta, 5) k_regex_tiles(RegexParams R) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* pipe = smem + rx_table_pad(R.D.table_bytes, R.D.n_states, R.D.in_smem != 0);
    uint16_t* cand = reinterpret_cast<uint16_t*>(pipe + tile_pipe_bytes(TB)) + warp_id() * R.cand_cap;
    const uint16_t* trans; const uint8_t* cls; const uint8_t* accept;
    stage_tables(R, smem, trans, cls, accept, R.D.scaled != 0);
    const DecodeParams& P = R.P;
    const uint32_t l = lane_id();
    const uint32_t trans_s = smem_u32(trans); // meaningful when the tables are staged (R.D.scaled implies it)
    const uint32_t start_s = trans_s + R.D.start * 512u, dead_s = trans_s + R.D.dead * 512u; // dead == ~0: never equal
    int max_def = 0;
    bool has_dict = false;
    uint32_t dict_n = 0;
    const uint8_t* dmatch = nullptr;
    tile_pipeline<TB>(P, pipe,
        [&](uint32_t chunk, uint64_t*, uint32_t&) {
            const DevChunk& ck = P.chunks[chunk];
            max_def = ck.max_def; has_dict = ck.has_dict; dict_n = ck.dict_ok_n;
            dmatch = R.dict_match + ck.dict_arena_off / 8;
        },
        [&](uint32_t q, const pqg_page_desc& pd, const uint8_t* pg) {
            const uint32_t n = pd.num_values, size = pd.payload_size;
            if (n == 0) return;
            if (max_def > 1) { if (l == 0) rx_to_slow(P, q); return; }
            uint32_t pos = 0, nn = n;
            bool slow = false;
            if (max_def == 1) {
                // definition levels: <varint < 128><level byte> RLE runs only, else the general scan
                uint32_t def_len = size >= 4 ? ld32u(pg) : 0xffffffffu;
                if (size < 4 || def_len > size - 4 || (def_len & 1u)) slow = true;
                else {
                    const uint8_t* s = pg + 4;
                    const uint32_t nr = def_len >> 1;
                    bool ok = true;
                    uint32_t carry = 0, present = 0;
                    for (uint32_t base = 0; base < nr && carry < n; base += 32) {
                        uint32_t r = base + l, cnt = 0, val = 0;
                        if (r < nr) { uint32_t b = s[2 * r]; ok = ok && ((b & 0x81u) == 0u) && b != 0u; cnt = b >> 1; val = s[2 * r + 1]; }
                        uint32_t incl = warp_incl_scan(cnt);
                        uint32_t start = carry + incl - cnt;
                        if (cnt && start < n && val >= 1u) present += min(cnt, n - start);
                        carry += __shfl_sync(0xffffffffu, incl, 31);
                    }
                    if (!__all_sync(0xffffffffu, ok)) slow = true;
                    nn = __reduce_add_sync(0xffffffffu, present);
                    pos = 4 + def_len;
                }
            }
            bool hit = false;
            if (!slow) {
                if ((pd.flags & PQG_PAGE_FLAG_DICT) && has_dict) {
                    const uint32_t bw = pos < size ? pg[pos] : 99u;
                    RegStream rs;
                    if (bw > 32u || !check_regular2(pg + pos + 1, size - pos - 1, bw, nn, &rs)) slow = true;
                    else {
                        for (uint32_t k0 = 0; k0 < nn && !hit; k0 += 32) {
                            uint32_t k = k0 + l;
                            bool h = false;
                            if (k < nn) { uint32_t ix = regular_index2(rs, k); h = ix < dict_n && dmatch[ix]; }
                            hit = __any_sync(0xffffffffu, h);
                        }
                    }
                } else {
                    const uint8_t* vals = pg + pos;
                    const uint32_t va = smem_u32(vals);
                    const SmemWords ld{va & ~3u};
                    const uint32_t o = va & 3u;
                    uint32_t endp = 0, c1 = 0, len1 = 0;
                    // Values of ONE length (header arithmetic): the section is nn x (4 + len) bytes, so string k sits at
                    // k * stride -- verified exactly (every prefix must read len; by induction those ARE the prefixes of
                    // the chain), and no candidate search is needed.
                    const uint32_t vsec = size - pos;
                    // (no division: the first prefix names the only possible common length)
                    const uint32_t stride = (nn && vsec >= 4u) ? __funnelshift_r(ld(o >> 2), ld((o >> 2) + 1u), (o & 3u) * 8u) + 4u : 0u;
                    bool uniform = nn > 0 && stride >= 4u && stride <= 65535u && stride * nn == vsec && vsec <= 65535u && nn <= R.cand_cap;
                    if (uniform) {
                        const uint32_t ulen = stride - 4u;
                        bool same = true;
                        for (uint32_t k = l; k < nn; k += 32) {
                            const uint32_t a = k * stride + o;
                            same = same && __funnelshift_r(ld(a >> 2), ld((a >> 2) + 1u), (a & 3u) * 8u) == ulen;
                        }
                        uniform = __all_sync(0xffffffffu, same);
                        if (uniform) {
                            for (uint32_t k0 = 0; k0 < nn && !hit; k0 += 32) {
                                const uint32_t k = k0 + l;
                                bool h = false;
                                if (k < nn) {
                                    bool m;
                                    if (R.D.scaled) {
                                        const uint32_t st = dfa_run_abs(start_s, trans_s, dead_s, ld, k * stride + o + 4u, ulen);
                                        m = accept[(st - trans_s) >> 9] != 0;
                                    } else m = dfa_run(R.D, trans, cls, accept, vals + k * stride + 4, ulen);
                                    h = R.neg ? !m : m;
                                }
                                hit = __any_sync(0xffffffffu, h);
                            }
                        }
                    }
                    if (uniform) { /* done */ }
                    else if (!find_headers_w(ld, o, size - pos, nn, cand, R.cand_cap, &endp, &c1, &len1)) slow = true;
                    else {
                        for (uint32_t k0 = 0; k0 < nn && !hit; k0 += 32) {
                            uint32_t k = k0 + l;
                            bool h = false;
                            if (k < nn) {
                                const uint32_t c = k0 ? cand[k] : c1;
                                const uint32_t len = k0 ? ld.u16at(c + o) : len1;
                                bool m;
                                if (R.D.scaled) {
                                    const uint32_t s = dfa_run_abs(start_s, trans_s, dead_s, ld, c + o + 4u, len);
                                    m = accept[(s - trans_s) >> 9] != 0;
                                } else m = dfa_run(R.D, trans, cls, accept, vals + c + 4, len);
                                h = R.neg ? !m : m;
                            }
                            hit = __any_sync(0xffffffffu, h);
                        }
                    }
                }
            }
            __syncwarp();
            if (slow) { if (l == 0) rx_to_slow(P, q); return; }
            if (hit && l == 0) atomicOr(&R.page_bits[q >> 5], 1u << (q & 31u));
        });
}

// ---------------------------------------------------------------------------------------------
// weights + device-wide exclusive scan
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t dec_digits(uint32_t v) { // length of std::to_string(v)
    uint32_t d = 1;
    while (v >= 10u) { v /= 10u; d++; }
    return d;
}

// weight of slot i: tuple level = decimal digits of the length + the length for non-null strings
struct StrWeights {
    const uint32_t* offsets;   // plan offsets (chunk c owns [row_base_c + c, ...])
    const uint32_t* validity;  // may be null
    const uint64_t* chunk_row_base; // n_chunks + 1 (device)
    uint32_t n_chunks;
    __device__ uint32_t chunk_of(uint64_t i) const {
        uint32_t lo = 0, hi = n_chunks;
        while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (chunk_row_base[mid] <= i) lo = mid; else hi = mid; }
        return lo;
    }
    // weight of slot i; c = chunk of some slot <= i (a cursor: threads walk consecutive slots)
    __device__ uint32_t at(uint64_t i, uint32_t& c) const {
        if (validity && !((validity[i >> 5] >> (i & 31)) & 1u)) return 0;
        while (c + 1 < n_chunks && chunk_row_base[c + 1] <= i) c++;
        const uint32_t* off = offsets + i + c;
        uint32_t len = off[1] - off[0];
        return dec_digits(len) + len;
    }
    // zero weight <=> null (a non-null string weighs at least the one digit of its length)
    __device__ bool is_zero(uint64_t i) const { return validity && !((validity[i >> 5] >> (i & 31)) & 1u); }
};
struct ArrWeights {
    const uint32_t* w;
    __device__ uint32_t chunk_of(uint64_t) const { return 0; }
    __device__ uint32_t at(uint64_t i, uint32_t&) const { return w[i]; }
    __device__ bool is_zero(uint64_t i) const { return w[i] == 0; }
};

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8; // per thread
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint64_t block_excl_scan(uint64_t v, uint64_t* total_out) {
    __shared__ uint64_t wsum[kScanThreads / 32];
    const uint32_t l = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint64_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint64_t t = __shfl_up_sync(0xffffffffu, incl, d); if (l >= static_cast<uint32_t>(d)) incl += t; }
    if (l == 31) wsum[w] = incl;
    __syncthreads();
    uint64_t base = 0, total = 0;
#pragma unroll
    for (int i = 0; i < kScanThreads / 32; i++) { uint64_t x = wsum[i]; if (i < static_cast<int>(w)) base += x; total += x; }
    __syncthreads();
    *total_out = total;
    return base + incl - v;
}

template <class Src>
__global__ void __launch_bounds__(kScanThreads) k_scan_reduce(Src src, uint64_t n, uint64_t* block_sums) {
    // every thread takes kScanItems consecutive slots (the same split as k_scan_write): one chunk lookup per thread
    const uint64_t base = static_cast<uint64_t>(blockIdx.x) * kScanTile + static_cast<uint64_t>(threadIdx.x) * kScanItems;
    uint64_t sum = 0;
    if (base < n) {
        uint32_t c = src.chunk_of(base);
#pragma unroll
        for (int j = 0; j < kScanItems; j++) { uint64_t i = base + j; if (i < n) sum += src.at(i, c); }
    }
    uint64_t total;
    block_excl_scan(sum, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block of 1024 threads: exclusive scan of the block sums in place; grand total -> sums[nb].
// Every thread owns a contiguous run of the sums (sequential inside the run, one block-wide scan of the run totals).
__global__ void __launch_bounds__(1024) k_scan_sums(uint64_t* sums, uint64_t nb) {
    __shared__ uint64_t wsum[32];
    const uint32_t tid = threadIdx.x, l = tid & 31u, w = tid >> 5;
    const uint64_t per = (nb + 1023) / 1024, a = min(nb, tid * per), b = min(nb, a + per);
    uint64_t run = 0;
    for (uint64_t i = a; i < b; i++) run += sums[i];
    uint64_t incl = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint64_t t = __shfl_up_sync(0xffffffffu, incl, d); if (l >= static_cast<uint32_t>(d)) incl += t; }
    if (l == 31) wsum[w] = incl;
    __syncthreads();
    if (w == 0) {
        uint64_t x = wsum[l], xi = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint64_t t = __shfl_up_sync(0xffffffffu, xi, d); if (l >= static_cast<uint32_t>(d)) xi += t; }
        wsum[l] = xi - x;
    }
    __syncthreads();
    uint64_t ex = wsum[w] + incl - run;
    for (uint64_t i = a; i < b; i++) { const uint64_t v = sums[i]; sums[i] = ex; ex += v; }
    if (tid == 1023) sums[nb] = ex;
}

// P[i] = exclusive prefix of the weights, P[n] = total
template <class Src>
__global__ void __launch_bounds__(kScanThreads) k_scan_write(Src src, uint64_t n, const uint64_t* block_sums, uint64_t* P) {
    const uint64_t base = static_cast<uint64_t>(blockIdx.x) * kScanTile + static_cast<uint64_t>(threadIdx.x) * kScanItems;
    uint32_t w[kScanItems];
    uint64_t sum = 0;
    uint32_t c = base < n ? src.chunk_of(base) : 0u;
#pragma unroll
    for (int j = 0; j < kScanItems; j++) { uint64_t i = base + j; w[j] = i < n ? src.at(i, c) : 0; sum += w[j]; }
    uint64_t total;
    uint64_t ex = block_sums[blockIdx.x] + block_excl_scan(sum, &total);
#pragma unroll
    for (int j = 0; j < kScanItems; j++) { uint64_t i = base + j; if (i < n) P[i] = ex; ex += w[j]; }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanThreads - 1) P[n] = block_sums[gridDim.x]; // grand total
}

// ---------------------------------------------------------------------------------------------
// the cut chain
// ---------------------------------------------------------------------------------------------
// first slot k in [0, n) with P[k] >= x, else n
__device__ __forceinline__ uint64_t lower_bound_P(const uint64_t* P, uint64_t n, uint64_t x) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (P[mid] < x) lo = mid + 1; else hi = mid; }
    return lo;
}
// same, knowing P[k] < x: gallop forward from k (cuts are a chunk size apart: a short hop)
__device__ __forceinline__ uint64_t lower_bound_from(const uint64_t* P, uint64_t n, uint64_t k, uint64_t x) {
    uint64_t lo = k, step = 1, hi = k + 1;
    while (hi < n && P[hi] < x) { lo = hi; step <<= 1; hi = min(n, hi + step); }
    while (lo + 1 < hi) { uint64_t mid = (lo + hi) >> 1; if (P[mid] < x) lo = mid; else hi = mid; }
    return hi;
}

// a weight-space tile spans this many chunk sizes: at least kChainTileChunks, more for long columns so that the
// tile count stays near kChainTilesTarget (every tile has ~(slots per chunk) candidates whose walk results go to the
// host for the stitch: fewer, longer tiles = fewer candidates to copy; the total walking work does not change)
constexpr uint32_t kChainTileChunks = 64, kChainTileChunksMax = 4096, kChainTilesTarget = 2048;

// per weight-space tile t: klo[t] = first slot with P >= t*L, ncand[t] = candidate first cuts
__global__ void k_chain_tiles(const uint64_t* P, uint64_t n, uint64_t S, uint64_t L, uint64_t T, uint64_t* klo, uint32_t* ncand) {
    uint64_t t = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t > T) return;
    if (t == T) { klo[T] = n; return; }
    // candidates: slots a .. b inclusive, a = first slot with P >= t*L, b = first slot with P >= t*L + S
    uint64_t a = lower_bound_P(P, n, t * L);
    klo[t] = a;
    uint64_t cnt = 0;
    if (a < n) {
        uint64_t b = P[a] >= t * L + S ? a : lower_bound_from(P, n, a, t * L + S);
        cnt = b < n ? b - a + 1 : n - a;
    }
    ncand[t] = static_cast<uint32_t>(cnt);
}

// next[k] = the cut that follows a cut at slot k: first slot with P >= P[k] + S.  A CTA stages the prefix values of its
// kNextBlock slots and of the ~1000 slots behind them in shared memory (a chunk of S bytes spans S / mean weight slots: ~110
// for the strings of the benchmarks) and every thread binary-searches there -- 8 .. 12 shared-memory probes per slot instead
// of a chain of dependent global loads (3.0 ms per 320 M slots before); a target beyond the window falls back to the
// galloping search in global memory.
constexpr uint32_t kNextBlock = 2048, kNextWin = 3072;
__global__ void __launch_bounds__(256) k_chain_next(const uint64_t* P, uint64_t n, uint64_t S, uint32_t* next) {
    __shared__ uint64_t sp[kNextWin];
    const uint64_t k0 = static_cast<uint64_t>(blockIdx.x) * kNextBlock;
    if (k0 >= n) return;
    const uint32_t win = static_cast<uint32_t>(min(static_cast<uint64_t>(kNextWin), n - k0));
    for (uint32_t i = threadIdx.x; i < win; i += 256) sp[i] = P[k0 + i];
    __syncthreads();
    const uint32_t cnt = min(kNextBlock, win);
    for (uint32_t i = threadIdx.x; i < cnt; i += 256) {
        const uint64_t x = sp[i] + S;
        uint64_t r;
        if (sp[win - 1] < x) r = k0 + win >= n ? n : lower_bound_from(P, n, k0 + win - 1, x);
        else {
            uint32_t lo = i, hi = win - 1; // sp[lo] < x <= sp[hi]
            if (i + 256u < hi && sp[i + 256u] >= x) hi = i + 256u;
            while (lo + 1u < hi) { const uint32_t mid = (lo + hi) >> 1; if (sp[mid] < x) lo = mid; else hi = mid; }
            r = k0 + hi;
        }
        next[k0 + i] = static_cast<uint32_t>(r);
    }
}

// every candidate of every tile walks its tile along next[]: number of cuts inside, first cut after it, prefix value of
// the last cut inside.  (Candidates x cuts per tile = slots per chunk x cuts in all: about one step per slot of the column.)
__global__ void __launch_bounds__(128) k_chain_walk(const uint64_t* P, const uint32_t* next, const uint64_t* klo,
                                                     const uint32_t* ncand, const uint64_t* cand_base, uint64_t* res_exit, uint32_t* res_cnt,
                                                     uint64_t* res_last) {
    const uint64_t t = blockIdx.x;
    const uint64_t end_slot = klo[t + 1]; // first slot of the next tile: P >= (t + 1) * L from there on
    for (uint32_t j = threadIdx.x; j < ncand[t]; j += blockDim.x) {
        uint64_t k = klo[t] + j, last = 0;
        uint32_t cnt = 0;
        while (k < end_slot) { cnt++; last = k; k = next[k]; }
        res_exit[cand_base[t] + j] = k;
        res_cnt[cand_base[t] + j] = cnt;
        res_last[cand_base[t] + j] = cnt ? P[last] : 0;
    }
}

// out[0] = total weight, out[1] = number of slots that can still start a chunk: with zero-weight (null) slots
// masked, the slots behind the last weighted one never do (the reference only cuts when a value follows)
__global__ void k_chain_totals(const uint64_t* P, uint64_t n, bool mask_zero, uint64_t* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const uint64_t total = P[n];
        out[0] = total;
        out[1] = mask_zero ? lower_bound_P(P, n, total) : n;
    }
}
// first cut of the sequence: first slot with carry_in + P >= S, i.e. P >= x0
__global__ void k_chain_entry(const uint64_t* P, uint64_t n_eff, uint64_t x0, uint64_t* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = lower_bound_P(P, n_eff, x0);
}

// materialise the cuts: tile t walks from its true entry
__global__ void k_chain_emit(const uint32_t* next, uint64_t n, uint64_t T, const uint64_t* klo, const uint64_t* tile_entry,
                             const uint64_t* tile_base, uint64_t* cuts) {
    uint64_t t = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= T) return;
    uint64_t k = tile_entry[t], o = tile_base[t];
    if (k >= n || k < klo[t]) return; // no cut starts in this tile
    const uint64_t end_slot = klo[t + 1];
    while (k < end_slot) { cuts[o++] = k; k = next[k]; }
}

// ids[i] = number of cuts at slots <= i (0 for zero-weight slots when mask_zero).  A block takes kIdsBlock consecutive slots;
// the cuts that fall into them (a handful: one per chunk size of weight) are staged in shared memory and every slot counts
// the ones at or before it there.
constexpr uint32_t kIdsBlock = 4096, kIdsCuts = 1024;
template <class Src>
__global__ void __launch_bounds__(256) k_chain_ids(Src src, uint64_t n, const uint64_t* cuts, uint64_t n_cuts, bool mask_zero, uint32_t id_base, uint32_t* ids) {
    __shared__ uint64_t first_s, last_s;
    __shared__ uint32_t cs[kIdsCuts]; // block-relative slots of the cuts inside the block
    const uint64_t b0 = static_cast<uint64_t>(blockIdx.x) * kIdsBlock, b1 = min(n, b0 + kIdsBlock);
    if (threadIdx.x < 2) { // cuts < b0 (thread 0), cuts < b1 (thread 1)
        const uint64_t x = threadIdx.x ? b1 : b0;
        uint64_t lo = 0, hi = n_cuts;
        while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (cuts[mid] < x) lo = mid + 1; else hi = mid; }
        if (threadIdx.x) last_s = lo; else first_s = lo;
    }
    __syncthreads();
    const uint64_t first = first_s;
    const uint32_t nc = static_cast<uint32_t>(last_s - first);
    const bool staged = nc <= kIdsCuts;
    if (staged) for (uint32_t j = threadIdx.x; j < nc; j += blockDim.x) cs[j] = static_cast<uint32_t>(cuts[first + j] - b0);
    __syncthreads();
    for (uint32_t j = threadIdx.x; b0 + j < b1; j += blockDim.x) {
        const uint64_t i = b0 + j;
        uint64_t c;
        if (staged) { // cuts at block-relative slots <= j
            uint32_t lo = 0, hi = nc;
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (cs[mid] <= j) lo = mid + 1; else hi = mid; }
            c = first + lo;
        } else {
            uint64_t lo = first, hi = first + nc;
            while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (cuts[mid] <= i) lo = mid + 1; else hi = mid; }
            c = lo;
        }
        ids[i] = (mask_zero && src.is_zero(i)) ? 0u : id_base + static_cast<uint32_t>(c);
    }
}

// page-level extras: byte offset of every page inside its chunk
__global__ void k_page_offsets(const uint64_t* P, uint64_t n, const uint32_t* ids, const uint64_t* cuts, uint32_t* off) {
    uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t id = ids[i];
    off[i] = static_cast<uint32_t>(P[i] - (id ? P[cuts[id - 1]] : 0));
}
__global__ void k_narrow(const uint64_t* in, uint64_t n, uint32_t* out) {
    uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = static_cast<uint32_t>(in[i]);
}
__global__ void k_chunk_row_bases(const DevChunk* chunks, uint32_t n_chunks, uint64_t n_slots, uint64_t* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_chunks) out[i] = chunks[i].out_row_base;
    if (i == n_chunks) out[i] = n_slots;
}

struct DevMem { // stream-ordered scratch (the device's memory pool keeps freed blocks); frees on scope exit
    cudaStream_t stream = nullptr;
    std::vector<void*> ptrs;
    explicit DevMem(cudaStream_t s) : stream(s) {}
    ~DevMem() { for (void* p : ptrs) cudaFreeAsync(p, stream); }
    template <class T> cudaError_t alloc(T** out, size_t count) {
        void* p = nullptr;
        cudaError_t e = cudaMallocAsync(&p, std::max<size_t>(count * sizeof(T), 16), stream);
        if (e == cudaSuccess) { ptrs.push_back(p); *out = static_cast<T*>(p); }
        return e;
    }
};

struct EventPair { // RAII: the timing events of one call
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t create() { cudaError_t e = cudaEventCreate(&e0); return e != cudaSuccess ? e : cudaEventCreate(&e1); }
    ~EventPair() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
};

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return e_; } while (0)

// Greedy chunking of a weight sequence src(0..n), in three phases so that shards of one column (one per GPU)
// overlap everything but a host loop:
//   prepare  weights -> exclusive prefix sums P -> per weight-space tile the candidate first cuts -> every candidate
//            walks its tile (cuts inside, exit, prefix value of its last cut).  Needs no carry: runs on all shards at once.
//   stitch   the carry of the previous shard fixes the first cut (one lower bound on the device, 8 bytes back), then one
//            table lookup per tile on the host -> number of cuts, carry for the next shard.  The only serial step.
//   emit     cuts and chunk ids materialised on the device (runs on all shards at once again).
struct ChainJob {
    DevMem mem;
    cudaStream_t s;
    uint64_t n_all = 0, n = 0, S = 0, L = 0, T = 0, total = 0;
    bool mask_zero = false;
    uint64_t* d_P = nullptr;
    uint64_t* d_entry = nullptr;
    uint64_t* d_klo = nullptr;   // T + 1 tile starts (slots)
    uint32_t* d_next = nullptr;  // n: the cut that follows a cut at each slot
    std::vector<uint64_t> klo, cand_base, r_exit, r_last, tile_entry, tile_base;
    std::vector<uint32_t> ncand, r_cnt;
    uint64_t n_cuts = 0, last_cut_P = 0, carry_in = 0;
    bool stitched = false;
    uint64_t* d_cuts = nullptr;
    uint32_t* d_ids = nullptr; // n_all entries (after emit)
    uint32_t launches = 0;
    explicit ChainJob(cudaStream_t st) : mem(st), s(st) {}
};

template <class Src>
cudaError_t chain_prepare(ChainJob& J, Src src, uint64_t n, uint64_t S, bool mask_zero) {
    cudaStream_t s = J.s;
    J.n_all = n; J.S = S; J.mask_zero = mask_zero;
    const uint64_t nb = (n + kScanTile - 1) / kScanTile;
    uint64_t* d_sums = nullptr;
    CK(J.mem.alloc(&d_sums, nb + 2));
    CK(J.mem.alloc(&J.d_P, n + 1));
    if (nb) {
        k_scan_reduce<Src><<<static_cast<unsigned>(nb), kScanThreads, 0, s>>>(src, n, d_sums);
        k_scan_sums<<<1, 1024, 0, s>>>(d_sums, nb);
        k_scan_write<Src><<<static_cast<unsigned>(nb), kScanThreads, 0, s>>>(src, n, d_sums, J.d_P);
        J.launches += 3;
    } else {
        CK(cudaMemsetAsync(J.d_P, 0, 8, s));
    }
    CK(J.mem.alloc(&J.d_entry, 2));
    k_chain_totals<<<1, 32, 0, s>>>(J.d_P, n, mask_zero, J.d_entry);
    J.launches++;
    uint64_t h_tot[2];
    CK(cudaMemcpyAsync(h_tot, J.d_entry, 16, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    J.total = h_tot[0];
    J.n = h_tot[1]; // the chain only sees slots that can start a chunk
    {
        const uint64_t chunks = J.total / S + 1;
        uint64_t m = (chunks + kChainTilesTarget - 1) / kChainTilesTarget;
        m = m < kChainTileChunks ? kChainTileChunks : (m > kChainTileChunksMax ? kChainTileChunksMax : m);
        J.L = S * m;
    }
    J.T = J.total / J.L + 1;
    const uint64_t T = J.T;
    if (J.n >= 0xffffffffull) return cudaErrorInvalidValue; // next[] holds 32-bit slots (a plan of >= 4 G slots is split by the host)
    uint64_t* d_klo = nullptr; uint32_t* d_ncand = nullptr;
    CK(J.mem.alloc(&d_klo, T + 1));
    CK(J.mem.alloc(&d_ncand, T + 1));
    CK(J.mem.alloc(&J.d_next, J.n + 1));
    J.d_klo = d_klo;
    k_chain_tiles<<<static_cast<unsigned>((T + 1 + 127) / 128), 128, 0, s>>>(J.d_P, J.n, S, J.L, T, d_klo, d_ncand);
    if (J.n) k_chain_next<<<static_cast<unsigned>((J.n + kNextBlock - 1) / kNextBlock), 256, 0, s>>>(J.d_P, J.n, S, J.d_next);
    J.launches += 2;
    J.klo.resize(T + 1);
    J.ncand.resize(T);
    CK(cudaMemcpyAsync(J.klo.data(), d_klo, (T + 1) * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(J.ncand.data(), d_ncand, T * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    J.cand_base.assign(T + 1, 0);
    for (uint64_t t = 0; t < T; t++) J.cand_base[t + 1] = J.cand_base[t] + J.ncand[t];
    const uint64_t n_cand = J.cand_base[T];
    uint64_t* d_cand_base = nullptr; uint64_t* d_exit = nullptr; uint32_t* d_cnt = nullptr; uint64_t* d_last = nullptr;
    CK(J.mem.alloc(&d_cand_base, T + 1));
    CK(J.mem.alloc(&d_exit, n_cand));
    CK(J.mem.alloc(&d_cnt, n_cand));
    CK(J.mem.alloc(&d_last, n_cand));
    CK(cudaMemcpyAsync(d_cand_base, J.cand_base.data(), (T + 1) * 8, cudaMemcpyHostToDevice, s));
    k_chain_walk<<<static_cast<unsigned>(T), 128, 0, s>>>(J.d_P, J.d_next, d_klo, d_ncand, d_cand_base, d_exit, d_cnt, d_last);
    J.launches++;
    J.r_exit.resize(n_cand); J.r_cnt.resize(n_cand); J.r_last.resize(n_cand);
    if (n_cand) {
        CK(cudaMemcpyAsync(J.r_exit.data(), d_exit, n_cand * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(J.r_cnt.data(), d_cnt, n_cand * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(J.r_last.data(), d_last, n_cand * 8, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    return cudaGetLastError();
}

inline cudaError_t chain_stitch(ChainJob& J, uint64_t carry_in) {
    cudaStream_t s = J.s;
    const uint64_t x0 = carry_in >= J.S ? 0 : J.S - carry_in;
    k_chain_entry<<<1, 32, 0, s>>>(J.d_P, J.n, x0, J.d_entry);
    J.launches++;
    uint64_t e = 0;
    CK(cudaMemcpyAsync(&e, J.d_entry, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const uint64_t T = J.T, n = J.n;
    J.tile_entry.assign(T, 0);
    J.tile_base.assign(T + 1, 0);
    J.last_cut_P = 0;
    for (uint64_t t = 0; t < T; t++) { // one lookup per tile
        J.tile_entry[t] = e;
        J.tile_base[t + 1] = J.tile_base[t];
        if (e < n && e >= J.klo[t] && e < J.klo[t + 1]) {
            const uint64_t j = e - J.klo[t];
            if (j >= J.ncand[t]) return cudaErrorAssert; // cannot happen: an entry is always a candidate
            const uint64_t ci = J.cand_base[t] + j;
            if (J.r_cnt[ci]) J.last_cut_P = J.r_last[ci];
            J.tile_base[t + 1] += J.r_cnt[ci];
            e = J.r_exit[ci];
        }
    }
    J.n_cuts = J.tile_base[T];
    J.carry_in = carry_in;
    J.stitched = true;
    return cudaSuccess;
}
inline uint64_t chain_carry_out(const ChainJob& J) { return J.n_cuts ? J.total - J.last_cut_P : J.carry_in + J.total; }

template <class Src>
cudaError_t chain_emit(ChainJob& J, Src src, uint32_t id_base) {
    cudaStream_t s = J.s;
    const uint64_t T = J.T;
    uint64_t* d_tile_entry = nullptr; uint64_t* d_tile_base = nullptr;
    CK(J.mem.alloc(&d_tile_entry, T));
    CK(J.mem.alloc(&d_tile_base, T + 1));
    CK(J.mem.alloc(&J.d_cuts, J.n_cuts));
    CK(J.mem.alloc(&J.d_ids, J.n_all + 1));
    CK(cudaMemcpyAsync(d_tile_entry, J.tile_entry.data(), T * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_tile_base, J.tile_base.data(), (T + 1) * 8, cudaMemcpyHostToDevice, s));
    k_chain_emit<<<static_cast<unsigned>((T + 127) / 128), 128, 0, s>>>(J.d_next, J.n, T, J.d_klo, d_tile_entry, d_tile_base, J.d_cuts);
    J.launches++;
    if (J.n_all) {
        k_chain_ids<Src><<<static_cast<unsigned>((J.n_all + kIdsBlock - 1) / kIdsBlock), 256, 0, s>>>(src, J.n_all, J.d_cuts, J.n_cuts, J.mask_zero, id_base, J.d_ids);
        J.launches++;
    }
    return cudaGetLastError();
}

// ---- device-resident consumer: a comparison predicate over a decoded fixed-width column ----------------
// bit per slot = value <op> constant, nulls never match; one value per lane, a word of the bitmap per warp step
template <typename T>
__global__ void __launch_bounds__(256) k_filter(const T* values, const uint32_t* validity, uint64_t n, int op, T c, uint32_t* bits, unsigned long long* count) {
    const uint64_t nwords = (n + 31) / 32;
    const uint32_t l = threadIdx.x & 31u;
    unsigned long long local = 0;
    for (uint64_t w = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; w < nwords; w += (static_cast<uint64_t>(gridDim.x) * blockDim.x) >> 5) {
        const uint64_t i = w * 32 + l;
        bool m = false;
        if (i < n) {
            const T v = values[i];
            m = op == PQG_CMP_EQ ? v == c : op == PQG_CMP_NE ? v != c : op == PQG_CMP_LT ? v < c : op == PQG_CMP_LE ? v <= c : op == PQG_CMP_GT ? v > c : v >= c;
        }
        uint32_t word = __ballot_sync(0xffffffffu, m);
        if (validity) word &= validity[w];
        if (l == 0) { bits[w] = word; local += __popc(word); }
    }
    if (l == 0 && local) atomicAdd(count, local);
}

template <typename T>
static cudaError_t run_filter(const void* values, const uint32_t* validity, uint64_t n, int op, const void* constant, uint32_t* bits,
                              unsigned long long* count, int sm_count, cudaStream_t s) {
    T c;
    std::memcpy(&c, constant, sizeof(T));
    const uint64_t warps = (n + 31) / 32;
    const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(std::max<uint64_t>((warps + 7) / 8, 1), static_cast<uint64_t>(sm_count) * 16u));
    k_filter<T><<<grid, 256, 0, s>>>(static_cast<const T*>(values), validity, n, op, c, bits, count);
    return cudaGetLastError();
}

} // namespace
} // namespace pqg

using namespace pqg;

#define CUF(ctx, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return ctx_fail(ctx, PQG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

extern "C" {

int pqg_plan_filter(pqg_ctx* ctx, pqg_plan* plan, int value_type, int op, const void* constant, uint32_t* row_bits, uint64_t* n_match, float* kernel_ms) {
    if (!ctx || !plan || !constant) return ctx_fail(ctx, PQG_ERR_ARG, "pqg_plan_filter: bad argument");
    if (plan_is_str(plan) || !plan_ran(plan) || plan_run_pending(plan))
        return ctx_fail(ctx, PQG_ERR_ARG, "pqg_plan_filter: the plan must be a decoded (run + finished) fixed-width plan");
    if (op < PQG_CMP_EQ || op > PQG_CMP_GE) return ctx_fail(ctx, PQG_ERR_ARG, "pqg_plan_filter: unknown comparison");
    const int width = plan_width(plan);
    const int want = (value_type == PQG_INT32 || value_type == PQG_FLOAT) ? 4 : (value_type == PQG_INT64 || value_type == PQG_DOUBLE) ? 8 : 0;
    if (!want || want != width) return ctx_fail(ctx, PQG_ERR_ARG, "pqg_plan_filter: value_type must be INT32 / INT64 / FLOAT / DOUBLE and match the plan's value width");
    CUF(ctx, cudaSetDevice(ctx_device(ctx)));
    cudaStream_t s = ctx_stream(ctx);
    const uint64_t n = plan_slots(plan), words = (n + 31) / 32;
    DevMem mem(s);
    uint32_t* d_bits = nullptr;
    unsigned long long* d_count = nullptr;
    CUF(ctx, mem.alloc(&d_bits, words + 1));
    CUF(ctx, mem.alloc(&d_count, 1));
    CUF(ctx, cudaMemsetAsync(d_count, 0, 8, s));
    EventPair ev;
    CUF(ctx, ev.create());
    CUF(ctx, cudaEventRecord(ev.e0, s));
    const void* vals = pqg_plan_values(plan);
    const uint32_t* validity = pqg_plan_validity(plan);
    cudaError_t e = cudaSuccess;
    if (n) {
        switch (value_type) {
            case PQG_INT32: e = run_filter<int32_t>(vals, validity, n, op, constant, d_bits, d_count, ctx_sm_count(ctx), s); break;
            case PQG_INT64: e = run_filter<int64_t>(vals, validity, n, op, constant, d_bits, d_count, ctx_sm_count(ctx), s); break;
            case PQG_FLOAT: e = run_filter<float>(vals, validity, n, op, constant, d_bits, d_count, ctx_sm_count(ctx), s); break;
            default: e = run_filter<double>(vals, validity, n, op, constant, d_bits, d_count, ctx_sm_count(ctx), s); break;
        }
    }
    CUF(ctx, e);
    ctx_add_launches(ctx, 1);
    CUF(ctx, cudaEventRecord(ev.e1, s));
    unsigned long long h_count = 0;
    CUF(ctx, cudaMemcpyAsync(&h_count, d_count, 8, cudaMemcpyDeviceToHost, s));
    if (row_bits && words) CUF(ctx, cudaMemcpyAsync(row_bits, d_bits, words * 4, cudaMemcpyDeviceToHost, s));
    CUF(ctx, cudaStreamSynchronize(s));
    if (n_match) *n_match = h_count;
    if (kernel_ms) { float ms = 0; cudaEventElapsedTime(&ms, ev.e0, ev.e1); *kernel_ms = ms; }
    return PQG_OK;
}

int pqg_regex_scan(pqg_ctx* ctx, pqg_plan* plan, const pqg_dfa* dfa, int neg, uint32_t* page_bits, float* kernel_ms) {
    if (!ctx || !plan || !dfa || !page_bits) return ctx_fail(ctx, PQG_ERR_ARG, "pqg_regex_scan: bad argument");
    if (!plan_is_str(plan)) return ctx_fail(ctx, PQG_ERR_ARG, "pqg_regex_scan: the plan must describe a BYTE_ARRAY column");
    // the scan prepares the chunks' dictionaries in its own way (no padded table): not between a run and its finish
    if (plan_run_pending(plan)) return ctx_fail(ctx, PQG_ERR_ARG, "pqg_regex_scan: the plan has an unfinished decode (call pqg_plan_finish first)");
    CUF(ctx, cudaSetDevice(ctx_device(ctx)));
    cudaStream_t s = ctx_stream(ctx);
    const CompiledDfa& d = dfa_tables(dfa);
    DecodeParams P = plan_params(plan);
    const uint32_t n_pages = P.page_end;
    const size_t words = (static_cast<size_t>(n_pages) + 31) / 32;
    std::memset(page_bits, 0, words * 4);
    if (n_pages == 0) { if (kernel_ms) *kernel_ms = 0; return PQG_OK; }
    // device tables: [trans][cls 256][accept]; byte-indexed transitions when they stay small
    const bool wide = static_cast<size_t>(d.n_states) * 256 * 2 <= 40 * 1024 && d.n_states < 255;
    std::vector<uint8_t> blob;
    const uint32_t table_bytes = wide ? d.n_states * 512u : static_cast<uint32_t>(d.trans.size() * 2);
    blob.resize(static_cast<size_t>(table_bytes) + 256 + d.n_states);
    uint16_t* tr = reinterpret_cast<uint16_t*>(blob.data());
    if (wide) {
        for (uint32_t st = 0; st < d.n_states; st++)
            for (uint32_t b = 0; b < 256; b++) // entries pre-scaled: next state * 256
                tr[st * 256u + b] = static_cast<uint16_t>(d.trans[static_cast<size_t>(st) * d.n_classes + d.cls[b]] << 8);
    } else {
        std::memcpy(tr, d.trans.data(), d.trans.size() * 2);
    }
    std::memcpy(blob.data() + table_bytes, d.cls, 256);
    std::memcpy(blob.data() + table_bytes + 256, d.accept.data(), d.n_states);
    DevMem mem(s);
    uint8_t* d_blob = nullptr; uint8_t* d_dmatch = nullptr; uint32_t* d_bits = nullptr;
    // the scan's own error record / work counters and hand-over list: the plan's decode state stays untouched
    DevErr* d_err = nullptr; uint32_t* d_append = nullptr;
    CUF(ctx, mem.alloc(&d_err, 1));
    CUF(ctx, mem.alloc(&d_append, static_cast<size_t>(n_pages) + 1));
    P.err = d_err; P.slow_append = d_append; P.slow_cap = n_pages + 1;
    CUF(ctx, mem.alloc(&d_blob, blob.size() + 16));
    CUF(ctx, mem.alloc(&d_dmatch, plan_dict_arena_bytes(plan) / 8 + 16));
    CUF(ctx, mem.alloc(&d_bits, words + 1));
    CUF(ctx, cudaMemcpyAsync(d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice, s));
    CUF(ctx, cudaMemsetAsync(d_bits, 0, (words + 1) * 4, s));
    CUF(ctx, cudaMemsetAsync(P.err, 0xFF, 8, s));
    CUF(ctx, cudaMemsetAsync(reinterpret_cast<uint8_t*>(P.err) + 8, 0, sizeof(DevErr) - 8, s));
    RegexParams R;
    R.P = P;
    // ~660 warp instructions per page: the scan is issue bound, and warps spinning on the next
    // stage's mbarrier cost more than idling at a barrier (measured 0.618 vs 0.650 ms per 20 M strings)
    R.P.tile_sync = plan_regex_tile_sync(plan) ? 1u : 0u;
    R.D.trans = reinterpret_cast<const uint16_t*>(d_blob);
    R.D.cls = d_blob + table_bytes;
    R.D.accept = d_blob + table_bytes + 256;
    R.D.n_states = d.n_states; R.D.n_classes = d.n_classes; R.D.start = d.start; R.D.dead = d.dead;
    R.D.wide = wide; R.D.table_bytes = table_bytes;
    const size_t tab_smem = static_cast<size_t>(table_bytes) + 256 + d.n_states;
    R.D.in_smem = tab_smem <= 48 * 1024;
    R.D.scaled = wide && R.D.in_smem;
    R.dict_match = d_dmatch; R.page_bits = d_bits; R.neg = neg ? 1 : 0;
    EventPair ev;
    CUF(ctx, ev.create());
    const cudaEvent_t e0 = ev.e0, e1 = ev.e1;
    uint32_t launches = 0;
    CUF(ctx, cudaEventRecord(e0, s));
    if (plan_any_dict(plan)) {
        P.skip_dict_pad = 1; // entries {start, len} only: the scan reads the dictionary chars in place
        CUF(ctx, launch_dict_prepare(P, P.n_chunks, 0, plan_str_dict_blocks(plan), s));
        k_regex_dict<<<dim3(64, P.n_chunks), 256, 0, s>>>(R);
        launches += 1 + dict_prepare_launches(0);
    }
    const size_t tab_pad = R.D.in_smem ? ((tab_smem + 15) & ~size_t(15)) : 0;
    R.cand_cap = plan_max_page_values(plan) <= kRxCand / 2 ? kRxCand / 2 : kRxCand;
    // the plan cut its tiles for 8 KB or (pages of ~1 KB and more: eight of them per tile, one per warp) 10 / 16 KB
    const int tb = static_cast<int>(R.P.tile_bytes);
    const size_t smem_fast = static_cast<size_t>(tile_pipe_bytes(tb)) + kWarpsPerCta * R.cand_cap * 2 + rx_table_pad(table_bytes, d.n_states, R.D.in_smem != 0);
    const size_t smem_slow = decode_smem_bytes(false) + tab_pad;
    auto* fast = tb == kTileBytesLarge ? k_regex_tiles<kTileBytesLarge> : tb == kTileBytesMid ? k_regex_tiles<kTileBytesMid> : k_regex_tiles<kTileBytes>;
    CUF(ctx, cudaFuncSetAttribute(fast, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_fast)));
    CUF(ctx, cudaFuncSetAttribute(fast, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CUF(ctx, cudaFuncSetAttribute(k_regex_pages, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_slow)));
    if (R.P.tile_hi > R.P.tile_lo) {
        int resident = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, fast, kThreadsPerCta, smem_fast);
        const uint32_t grid = tile_grid(R.P.tile_hi - R.P.tile_lo, ctx_sm_count(ctx), resident, &R.P.tiles_per_cta);
        fast<<<grid, kThreadsPerCta, smem_fast, s>>>(R);
        launches++;
    }
    // host-listed (oversized) pages + whatever the tile kernel handed over
    k_regex_pages<<<static_cast<unsigned>(ctx_sm_count(ctx)) * 2u, kThreadsPerCta, smem_slow, s>>>(R);
    launches++;
    CUF(ctx, cudaEventRecord(e1, s));
    CUF(ctx, cudaGetLastError());
    DevErr herr;
    CUF(ctx, cudaMemcpyAsync(page_bits, d_bits, words * 4, cudaMemcpyDeviceToHost, s));
    CUF(ctx, cudaMemcpyAsync(&herr, P.err, sizeof(DevErr), cudaMemcpyDeviceToHost, s));
    CUF(ctx, cudaStreamSynchronize(s));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (kernel_ms) *kernel_ms = ms;
    ctx_add_launches(ctx, launches);
    if (std::getenv("PQG_DEBUG")) std::fprintf(stderr, "[pqg] regex scan: %u pages, %u through the general kernel, %.3f ms\n", n_pages, herr.slow_count, ms);
    if (herr.count) {
        char msg[160];
        std::snprintf(msg, sizeof(msg), "regex scan: page %u failed to decode (code %u)", static_cast<uint32_t>(herr.key >> 32),
                      static_cast<uint32_t>(herr.key & 0xffffffffu));
        return ctx_fail(ctx, PQG_ERR_PAGE, msg);
    }
    return PQG_OK;
}

// ---- tuple-level chunk index in phases (multi-GPU: prepare on every shard at once, stitch in shard order, emit at once) ----
struct pqg_chunk_job {
    ChainJob J;
    StrWeights src{};
    float prepare_ms = 0;
    explicit pqg_chunk_job(cudaStream_t s) : J(s) {}
};

int pqg_chunk_index_prepare(pqg_ctx* ctx, pqg_plan* plan, uint64_t chunk_size, pqg_chunk_job** out, float* kernel_ms) {
    if (!ctx || !plan || !out) return ctx_fail(ctx, PQG_ERR_ARG, "pqg_chunk_index_prepare: bad argument");
    *out = nullptr;
    if (chunk_size == 0) return ctx_fail(ctx, PQG_ERR_ARG, "pqg_chunk_index: chunk_size must be > 0");
    if (!plan_is_str(plan) || !plan_ran(plan) || plan_run_pending(plan))
        return ctx_fail(ctx, PQG_ERR_ARG, "pqg_chunk_index: needs a decoded BYTE_ARRAY plan (pqg_plan_run + pqg_plan_finish first)");
    CUF(ctx, cudaSetDevice(ctx_device(ctx)));
    cudaStream_t s = ctx_stream(ctx);
    DecodeParams P = plan_params(plan);
    const uint64_t n = plan_slots(plan);
    pqg_chunk_job* job = new (std::nothrow) pqg_chunk_job(s);
    if (!job) return ctx_fail(ctx, PQG_ERR_NOMEM, "out of memory");
    struct Guard { pqg_chunk_job* j; ~Guard() { delete j; } } g{job};
    uint64_t* d_rb = nullptr;
    CUF(ctx, job->J.mem.alloc(&d_rb, P.n_chunks + 2));
    EventPair ev;
    CUF(ctx, ev.create());
    CUF(ctx, cudaEventRecord(ev.e0, s));
    k_chunk_row_bases<<<(P.n_chunks + 1 + 127) / 128, 128, 0, s>>>(P.chunks, P.n_chunks, n, d_rb);
    job->src = StrWeights{P.offsets, P.validity, d_rb, P.n_chunks};
    cudaError_t ce = chain_prepare(job->J, job->src, n, chunk_size, true);
    if (ce != cudaSuccess) return ctx_fail(ctx, PQG_ERR_CUDA, std::string("chunk index: ") + cudaGetErrorString(ce));
    CUF(ctx, cudaEventRecord(ev.e1, s));
    CUF(ctx, cudaStreamSynchronize(s));
    cudaEventElapsedTime(&job->prepare_ms, ev.e0, ev.e1);
    if (kernel_ms) *kernel_ms = job->prepare_ms;
    g.j = nullptr;
    *out = job;
    return PQG_OK;
}

int pqg_chunk_index_stitch(pqg_ctx* ctx, pqg_chunk_job* job, uint64_t carry_in, uint64_t* n_chunks, uint64_t* carry_out) {
    if (!ctx || !job) return ctx_fail(ctx, PQG_ERR_ARG, "pqg_chunk_index_stitch: bad argument");
    CUF(ctx, cudaSetDevice(ctx_device(ctx)));
    cudaError_t ce = chain_stitch(job->J, carry_in);
    if (ce != cudaSuccess) return ctx_fail(ctx, PQG_ERR_CUDA, std::string("chunk index stitch: ") + cudaGetErrorString(ce));
    if (n_chunks) *n_chunks = job->J.n_cuts + 1;
    if (carry_out) *carry_out = chain_carry_out(job->J);
    return PQG_OK;
}

int pqg_chunk_index_emit(pqg_ctx* ctx, pqg_chunk_job* job, uint32_t id_base, uint32_t* tuple_to_chunk, float* kernel_ms) {
    if (!ctx || !job) return ctx_fail(ctx, PQG_ERR_ARG, "pqg_chunk_index_emit: bad argument");
    if (!job->J.stitched) return ctx_fail(ctx, PQG_ERR_ARG, "pqg_chunk_index_emit: call pqg_chunk_index_stitch first");
    CUF(ctx, cudaSetDevice(ctx_device(ctx)));
    cudaStream_t s = job->J.s;
    EventPair ev;
    CUF(ctx, ev.create());
    CUF(ctx, cudaEventRecord(ev.e0, s));
    cudaError_t ce = chain_emit(job->J, job->src, id_base);
    if (ce != cudaSuccess) return ctx_fail(ctx, PQG_ERR_CUDA, std::string("chunk index emit: ") + cudaGetErrorString(ce));
    CUF(ctx, cudaEventRecord(ev.e1, s));
    if (tuple_to_chunk && job->J.n_all) CUF(ctx, cudaMemcpyAsync(tuple_to_chunk, job->J.d_ids, job->J.n_all * 4, cudaMemcpyDeviceToHost, s));
    CUF(ctx, cudaStreamSynchronize(s));
    float ms = 0;
    cudaEventElapsedTime(&ms, ev.e0, ev.e1);
    if (kernel_ms) *kernel_ms = ms;
    ctx_add_launches(ctx, job->J.launches + 1);
    job->J.launches = 0;
    return PQG_OK;
}

const uint32_t* pqg_chunk_job_ids(const pqg_chunk_job* job) { return job ? job->J.d_ids : nullptr; }
uint64_t pqg_chunk_job_total_weight(const pqg_chunk_job* job) { return job ? job->J.total : 0; }

void pqg_chunk_job_free(pqg_ctx* ctx, pqg_chunk_job* job) {
    if (!job) return;
    if (ctx) cudaSetDevice(ctx_device(ctx));
    delete job;
}

int pqg_chunk_index(pqg_ctx* ctx, pqg_plan* plan, uint64_t chunk_size, uint64_t carry_in, uint32_t id_base, uint32_t* tuple_to_chunk,
                    uint64_t* n_chunks, uint64_t* carry_out, float* kernel_ms) {
    pqg_chunk_job* job = nullptr;
    float ms0 = 0, ms1 = 0;
    int rc = pqg_chunk_index_prepare(ctx, plan, chunk_size, &job, &ms0);
    if (rc != PQG_OK) return rc;
    rc = pqg_chunk_index_stitch(ctx, job, carry_in, n_chunks, carry_out);
    if (rc == PQG_OK) rc = pqg_chunk_index_emit(ctx, job, id_base, tuple_to_chunk, &ms1);
    pqg_chunk_job_free(ctx, job);
    if (kernel_ms) *kernel_ms = ms0 + ms1;
    return rc;
}

int pqg_page_chunk_index(pqg_ctx* ctx, const uint32_t* page_sizes, uint32_t n_pages, uint64_t chunk_size, uint32_t* page_chunk,
                         uint32_t* page_off, uint32_t* chunk_first_page, uint32_t cap, uint32_t* n_chunks) {
    if (!ctx || (!page_sizes && n_pages)) return ctx_fail(ctx, PQG_ERR_ARG, "pqg_page_chunk_index: bad argument");
    if (chunk_size == 0) return ctx_fail(ctx, PQG_ERR_ARG, "pqg_page_chunk_index: chunk_size must be > 0");
    if (n_chunks) *n_chunks = 0;
    if (n_pages == 0) return PQG_OK;
    CUF(ctx, cudaSetDevice(ctx_device(ctx)));
    cudaStream_t s = ctx_stream(ctx);
    ChainJob J(s);
    uint32_t* d_w = nullptr; uint32_t* d_off = nullptr; uint32_t* d_first = nullptr;
    CUF(ctx, J.mem.alloc(&d_w, n_pages));
    CUF(ctx, J.mem.alloc(&d_off, n_pages));
    CUF(ctx, cudaMemcpyAsync(d_w, page_sizes, static_cast<size_t>(n_pages) * 4, cudaMemcpyHostToDevice, s));
    ArrWeights src{d_w};
    // the first page never closes a chunk (there is nothing to close): carry_in = 0
    cudaError_t ce = chain_prepare(J, src, n_pages, chunk_size, false);
    if (ce == cudaSuccess) ce = chain_stitch(J, 0);
    if (ce == cudaSuccess) ce = chain_emit(J, src, 0);
    if (ce != cudaSuccess) return ctx_fail(ctx, PQG_ERR_CUDA, std::string("page chunk index: ") + cudaGetErrorString(ce));
    k_page_offsets<<<(n_pages + 255) / 256, 256, 0, s>>>(J.d_P, n_pages, J.d_ids, J.d_cuts, d_off);
    if (page_chunk) CUF(ctx, cudaMemcpyAsync(page_chunk, J.d_ids, static_cast<size_t>(n_pages) * 4, cudaMemcpyDeviceToHost, s));
    if (page_off) CUF(ctx, cudaMemcpyAsync(page_off, d_off, static_cast<size_t>(n_pages) * 4, cudaMemcpyDeviceToHost, s));
    if (chunk_first_page && cap) {
        chunk_first_page[0] = 0;
        uint64_t take = std::min<uint64_t>(J.n_cuts, cap - 1);
        if (take) {
            CUF(ctx, J.mem.alloc(&d_first, take));
            k_narrow<<<static_cast<unsigned>((take + 255) / 256), 256, 0, s>>>(J.d_cuts, take, d_first);
            CUF(ctx, cudaMemcpyAsync(chunk_first_page + 1, d_first, take * 4, cudaMemcpyDeviceToHost, s));
        }
    }
    CUF(ctx, cudaStreamSynchronize(s));
    if (n_chunks) *n_chunks = static_cast<uint32_t>(J.n_cuts + 1);
    ctx_add_launches(ctx, J.launches + 2);
    return PQG_OK;
}

} // extern "C"
