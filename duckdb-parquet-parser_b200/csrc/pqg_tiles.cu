// pqg_tiles.cu -- the fast fixed-width decode kernel: TMA-staged page tiles.
//
// Covers flat INT32/INT64/FLOAT/DOUBLE chunks, REQUIRED or OPTIONAL (max_def <= 1), as the
// reference's writer emits them (src/writer/parquet_writer.cpp:376-460) and as foreign writers do
// while their pages fit a tile:
//   PLAIN pages          payload = the value array        (reader: column_reader.cpp:213-222,227-248)
//   dictionary pages     u8 bit width + the RLE / bit-packed hybrid stream: single bit-packed
//                        groups "03 <bw bytes>" with positional index extraction (the writer's
//                        shape), any other well-formed stream run by run (hybrid_runs_page)
//                        (reader: column_reader.cpp:174-196 over rle_decoder.hpp:17-95)
//   definition levels    RLE runs <varint < 128><level> (the writer's shape) -> validity image by
//                        toggle bits + prefix XOR; one run covering the page -> REQUIRED path
// Anything else found in a tile (bit-packed levels, out-of-range indices, truncated or malformed
// pages, bit width > 32) is NOT decoded here: the page is appended to the slow list and the
// general kernel (pqg_decode.cu) handles it right after, with the reference's error reporting.
// BOOLEAN / INT96 / nested chunks and pages larger than a tile never enter a tile (the host lists
// them for the big-page and general kernels).
//
// Data movement (HBM-bound, no tensor cores):
//   * the plan's host side cuts every chunk into tiles: <= 8 consecutive pages whose bytes
//     (page headers in between included -- pages of a chunk are contiguous in the file) fit
//     8 KB (16 KB for plans with OPTIONAL chunks).  Tile bytes + the tile's page descriptors are
//     staged into a kTileStages-deep shared-memory ring with cp.async.bulk (1-D TMA, UBLKCP)
//     completing on an mbarrier; 8 warps decode one page each out of shared memory; the last warp
//     to finish a stage refills it (pqg_tilepipe.cuh).  No register staging.
//   * a dictionary that fits (<= 32 KB of values) is staged once per chunk with the same
//     bulk copy; larger ones are gathered from L2 (the 126 MB L2 holds even the 8 MB
//     dictionary of a 2^20-key chunk; outputs are written with streaming stores so they do
//     not evict it).
//   * outputs: 8 or 4 bytes per lane, consecutive lanes -> consecutive slots (coalesced).
#include <algorithm>

#include "pqg_tilepipe.cuh"

namespace pqg {
namespace {

constexpr int kLevelScratchBytes = kWarpsPerCta * 64 * 4; // validity image + rank bases per warp
constexpr int kOptIdxBytes = kWarpsPerCta * 1024 * 2;   // u16 dictionary indices of a 1024-slot sub-tile per warp (opt_page_runs)

template <int W> struct FElem;
template <> struct FElem<4> { using T = uint32_t; };
template <> struct FElem<8> { using T = uint64_t; };

template <int W> __device__ __forceinline__ typename FElem<W>::T ld_elem(const uint8_t* p);
template <> __device__ __forceinline__ uint32_t ld_elem<4>(const uint8_t* p) { return ld32u(p); }
template <> __device__ __forceinline__ uint64_t ld_elem<8>(const uint8_t* p) { return ld64u(p); }

template <typename T> __device__ __forceinline__ T lds_elem(uint32_t saddr);
template <> __device__ __forceinline__ uint32_t lds_elem<uint32_t>(uint32_t saddr) {
    uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr)); return v;
}
template <> __device__ __forceinline__ uint64_t lds_elem<uint64_t>(uint32_t saddr) {
    uint64_t v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(saddr)); return v;
}

template <typename T> __device__ __forceinline__ void st_stream(T* p, T v);
template <> __device__ __forceinline__ void st_stream<uint32_t>(uint32_t* p, uint32_t v) { __stcs(p, v); }
template <> __device__ __forceinline__ void st_stream<uint64_t>(uint64_t* p, uint64_t v) {
    __stcs(reinterpret_cast<unsigned long long*>(p), static_cast<unsigned long long>(v));
}

__device__ __forceinline__ void to_slow(const DecodeParams& P, uint32_t q) {
    // one lane
    uint32_t k = atomicAdd(&P.err->slow_count, 1u);
    PQG_ASSERT(k < P.slow_cap);
    P.slow_append[k] = q;
}

// random dictionary gather from global memory: L2 only (ld.global.cg; measured 2-3 % ahead of
// ld.global.nc and ld.global.nc.L1::no_allocate, scripts/ubench_gather.cu)
template <typename T> __device__ __forceinline__ T ldg_gather(const T* p);
template <> __device__ __forceinline__ uint64_t ldg_gather<uint64_t>(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
template <> __device__ __forceinline__ uint32_t ldg_gather<uint32_t>(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// How a dictionary index turns into a value -- resolved ONCE per page and compiled into the
// inner loops (a run-time choice per value cost ~18 instructions and divergent branches):
enum { kDictIdent = 0, kDictSmem = 1, kDictGlobal = 2 };
template <int M> struct ModeTag { static constexpr int value = M; };
struct DictRef { const void* gptr; uint32_t saddr; };
template <typename T, int W, int MODE>
__device__ __forceinline__ T dict_get(const DictRef& d, uint32_t ix) {
    if constexpr (MODE == kDictIdent) return static_cast<T>(ix);
    else if constexpr (MODE == kDictSmem) { PQG_ASSERT(ix * W < static_cast<uint32_t>(kMaxSmemDictBytes)); return lds_elem<T>(d.saddr + ix * W); }
    else return ldg_gather<T>(static_cast<const T*>(d.gptr) + ix);
}
template <class F>
__device__ __forceinline__ void with_dict_mode(int mode, F&& f) {
    switch (mode) {
        case kDictIdent: f(ModeTag<kDictIdent>{}); break;
        case kDictSmem: f(ModeTag<kDictSmem>{}); break;
        default: f(ModeTag<kDictGlobal>{}); break;
    }
}

// RleDecoder::get_batch (include/reader/rle_decoder.hpp:17-95) for a whole page of dictionary
// indices, run by run: every lane decodes the (warp-uniform) run header; a literal run of g groups
// is 8 g consecutive bw-bit values -- lane k takes values k, k + 32, ...; an RLE run is one
// dictionary value stored count times.  Returns false on anything the reference handles by
// zero-filling, throwing or producing nulls (stream exhausted early, zero-length RLE run, value
// bytes cut off, index >= dictionary size): the general kernel redoes those pages.
template <int W, int MODE>
__device__ __noinline__ bool hybrid_runs_page(const DecodeParams& P, const uint8_t* s, uint32_t len, uint32_t bw, uint32_t n,
                                             const DictRef& dref, uint32_t dict_n, typename FElem<W>::T* out) {
    using T = typename FElem<W>::T;
    const uint32_t l = lane_id();
    const uint32_t sa = smem_u32(s);
    const SmemWords ldw{sa & ~3u};
    const uint32_t bit0 = (sa & 3u) * 8u;
    const uint32_t nb = (bw + 7u) >> 3, imask = bw >= 32u ? 0xffffffffu : ((1u << bw) - 1u);
    uint32_t pos = 0, v = 0;
    bool bad = false;
    while (v < n) {
        uint32_t hdr = 0, shift = 0;
        for (;;) { // varint32 (rle_decoder.hpp:76-86)
            if (pos >= len || shift > 28u) return false;
            const uint32_t b = s[pos++];
            hdr |= (b & 0x7fu) << shift;
            if (!(b & 0x80u)) break;
            shift += 7u;
        }
        if (hdr & 1u) {
            const uint32_t groups = hdr >> 1;
            const uint64_t bytes = static_cast<uint64_t>(groups) * bw;
            if (pos + bytes > len) return false;
            const uint32_t take = static_cast<uint32_t>(min(static_cast<uint64_t>(groups) * 8u, static_cast<uint64_t>(n - v)));
            const uint32_t base = bit0 + pos * 8u;
            for (uint32_t k = l; k < take; k += 32) {
                const uint32_t bit = base + k * bw;
                const uint32_t ix = __funnelshift_r(ldw(bit >> 5), ldw((bit >> 5) + 1u), bit & 31u) & imask;
                bad = bad || ix >= dict_n;
                st_stream<T>(out + v + k, dict_get<T, W, MODE>(dref, ix < dict_n ? ix : 0u));
            }
            pos += static_cast<uint32_t>(bytes);
            v += take;
        } else {
            const uint32_t cnt = hdr >> 1;
            if (cnt == 0 || pos + nb > len) return false;
            uint32_t ix = 0;
            for (uint32_t i = 0; i < nb; i++) ix |= static_cast<uint32_t>(s[pos + i]) << (8u * i); // not masked, like the reference (:88-95)
            if (ix >= dict_n) return false;
            const T x = dict_get<T, W, MODE>(dref, ix);
            const uint32_t take = min(cnt, n - v);
            for (uint32_t k = l; k < take; k += 32) st_stream<T>(out + v + k, x);
            pos += nb;
            v += take;
        }
    }
    return !__any_sync(0xffffffffu, bad);
}

// One warp decodes one page out of the staged tile.  `pg` = first payload byte (shared).
// PART: partitioned-dictionary launch -- this CTA holds entries [part_lo, part_lo + P.part_entries) of the dictionary in shared
// memory (dictp) and stores only the values whose index lies there; its sibling CTAs emit the rest of the same page.
template <int W, bool PART = false>
__device__ __forceinline__ bool fast_page(const DecodeParams& P, uint32_t q, const pqg_page_desc& pd, const uint8_t* pg,
                                          bool chunk_has_dict, const uint8_t* dictp, uint32_t dict_n, bool dict_in_smem, uint32_t part_lo = 0) {
    using T = typename FElem<W>::T;
    const uint32_t l = lane_id();
    const uint32_t n = pd.num_values, size = pd.payload_size;
    if (n == 0) return true;
    PQG_ASSERT(pd.out_row_base + n <= P.n_slots);
    T* out = reinterpret_cast<T*>(P.values) + pd.out_row_base;
    if (!((pd.flags & PQG_PAGE_FLAG_DICT) && chunk_has_dict)) {
        // PLAIN: read_plain_value per slot == a shifted copy
        if constexpr (PART) { if (part_lo != 0) return true; } // (a PLAIN fallback page inside a dictionary chunk: one sibling copies it)
        if (static_cast<uint64_t>(n) * W > size) { if (l == 0) to_slow(P, q); return false; }
        uint32_t e = l;
        for (; e + 96 < n; e += 128) {
            T v0 = ld_elem<W>(pg + static_cast<size_t>(e) * W);
            T v1 = ld_elem<W>(pg + static_cast<size_t>(e + 32) * W);
            T v2 = ld_elem<W>(pg + static_cast<size_t>(e + 64) * W);
            T v3 = ld_elem<W>(pg + static_cast<size_t>(e + 96) * W);
            st_stream<T>(out + e, v0); st_stream<T>(out + e + 32, v1);
            st_stream<T>(out + e + 64, v2); st_stream<T>(out + e + 96, v3);
        }
        for (; e < n; e += 32) st_stream<T>(out + e, ld_elem<W>(pg + static_cast<size_t>(e) * W));
        return true;
    }
    // dictionary indices: u8 bit width, then the RLE / bit-packed hybrid stream
    if (size < 1) { if (l == 0) to_slow(P, q); return false; }
    const uint32_t bw = pg[0];
    const uint8_t* s = pg + 1;
    RegStream rs;
    if (bw > 32) { if (l == 0) to_slow(P, q); return false; }
    const DictRef dref{dictp, dict_in_smem ? smem_u32(dictp) : 0u};
    const int mode = P.identity_dict ? kDictIdent : (dict_in_smem ? kDictSmem : kDictGlobal);
    if constexpr (PART) {
        // the writer's stream shape only; anything else goes to the general kernel (handed over by ONE sibling)
        if (bw > 32 || !check_regular2(s, size - 1, bw, n, &rs)) { if (l == 0 && part_lo == 0) to_slow(P, q); return false; }
        const uint32_t sa = smem_u32(s);
        const SmemWords ldw{sa & ~3u};
        const uint32_t gs = 1u + bw, imask = bw >= 32u ? 0xffffffffu : ((1u << bw) - 1u);
        const uint32_t bit_l = (sa & 3u) * 8u + (((l >> 3) * gs + 1u) << 3) + (l & 7u) * bw;
        const uint32_t step32 = 32u * gs, dsa = smem_u32(dictp), pn = P.part_entries;
        bool bad = false;
        auto index_at = [&](uint32_t bitpos, uint32_t k) -> uint32_t {
            const uint32_t ix = __funnelshift_r(ldw(bitpos >> 5), ldw((bitpos >> 5) + 1u), bitpos & 31u) & imask;
            return k >= rs.tail_start ? rs.tail_val : ix;
        };
        uint32_t bit = bit_l, v = l;
        for (; v + 96 < n; v += 128, bit += 4u * step32) {
            const uint32_t i0 = index_at(bit, v), i1 = index_at(bit + step32, v + 32);
            const uint32_t i2 = index_at(bit + 2u * step32, v + 64), i3 = index_at(bit + 3u * step32, v + 96);
            bad = bad || i0 >= dict_n || i1 >= dict_n || i2 >= dict_n || i3 >= dict_n;
            const uint32_t r0 = i0 - part_lo, r1 = i1 - part_lo, r2 = i2 - part_lo, r3 = i3 - part_lo; // wraps below the part
            if (r0 < pn && i0 < dict_n) st_stream<T>(out + v, lds_elem<T>(dsa + r0 * W));
            if (r1 < pn && i1 < dict_n) st_stream<T>(out + v + 32, lds_elem<T>(dsa + r1 * W));
            if (r2 < pn && i2 < dict_n) st_stream<T>(out + v + 64, lds_elem<T>(dsa + r2 * W));
            if (r3 < pn && i3 < dict_n) st_stream<T>(out + v + 96, lds_elem<T>(dsa + r3 * W));
        }
        for (; v < n; v += 32, bit += step32) {
            const uint32_t i0 = index_at(bit, v), r0 = i0 - part_lo;
            bad = bad || i0 >= dict_n;
            if (r0 < pn && i0 < dict_n) st_stream<T>(out + v, lds_elem<T>(dsa + r0 * W));
        }
        const bool any_bad = __any_sync(0xffffffffu, bad);
        if (any_bad && l == 0 && part_lo == 0) { to_slow(P, q); atomicAdd(&P.err->bad_index, 1u); }
        return !any_bad;
    }
    if (!check_regular2(s, size - 1, bw, n, &rs)) {
        // any other well-formed hybrid stream (RLE runs between the groups, literal runs of several
        // groups as foreign writers emit them): runs in sequence, the warp expands each one together
        bool ok = false;
        with_dict_mode(mode, [&](auto tag) { ok = hybrid_runs_page<W, decltype(tag)::value>(P, s, size - 1u, bw, n, dref, dict_n, out); });
        if (!ok && l == 0) to_slow(P, q); // malformed / truncated / out-of-range index: the general kernel reports like the reference
        return ok;
    }
    // index bits straight from aligned shared-memory words; value v sits in group v >> 3 at
    // bit ((v >> 3) * (1 + bw) + 1) * 8 + (v & 7) * bw of the stream: +32 values = +4 groups
    const uint32_t sa = smem_u32(s);
    const SmemWords ldw{sa & ~3u};
    const uint32_t gs = 1u + bw, imask = bw >= 32u ? 0xffffffffu : ((1u << bw) - 1u);
    const uint32_t bit_l = (sa & 3u) * 8u + (((l >> 3) * gs + 1u) << 3) + (l & 7u) * bw;
    const uint32_t step32 = 32u * gs;
    bool bad = false;
    with_dict_mode(mode, [&](auto tag) {
        constexpr int MODE = decltype(tag)::value;
        auto index_at = [&](uint32_t bitpos, uint32_t k) -> uint32_t {
            const uint32_t ix = __funnelshift_r(ldw(bitpos >> 5), ldw((bitpos >> 5) + 1u), bitpos & 31u) & imask;
            return k >= rs.tail_start ? rs.tail_val : ix;
        };
        uint32_t bit = bit_l, v = l;
        for (; v + 96 < n; v += 128, bit += 4u * step32) {
            const uint32_t i0 = index_at(bit, v), i1 = index_at(bit + step32, v + 32);
            const uint32_t i2 = index_at(bit + 2u * step32, v + 64), i3 = index_at(bit + 3u * step32, v + 96);
            bad = bad || i0 >= dict_n || i1 >= dict_n || i2 >= dict_n || i3 >= dict_n;
            // (an out-of-range index reads entry 0 instead of branching: the page is redone anyway)
            const T x0 = dict_get<T, W, MODE>(dref, i0 < dict_n ? i0 : 0u);
            const T x1 = dict_get<T, W, MODE>(dref, i1 < dict_n ? i1 : 0u);
            const T x2 = dict_get<T, W, MODE>(dref, i2 < dict_n ? i2 : 0u);
            const T x3 = dict_get<T, W, MODE>(dref, i3 < dict_n ? i3 : 0u);
            st_stream<T>(out + v, x0); st_stream<T>(out + v + 32, x1);
            st_stream<T>(out + v + 64, x2); st_stream<T>(out + v + 96, x3);
        }
        for (; v < n; v += 32, bit += step32) {
            const uint32_t i0 = index_at(bit, v);
            bad = bad || i0 >= dict_n;
            st_stream<T>(out + v, dict_get<T, W, MODE>(dref, i0 < dict_n ? i0 : 0u));
        }
    });
    const bool any_bad = __any_sync(0xffffffffu, bad);
    if (any_bad && l == 0) { to_slow(P, q); atomicAdd(&P.err->bad_index, 1u); }
    return !any_bad;
}

// ---- OPTIONAL pages, any well-formed streams (foreign writers) ---------------------------------------------------
// What fast_page_opt does not take -- definition levels in literal (bit-packed) runs or RLE runs longer than 63, index
// streams with RLE runs / literal runs of several groups, pages of more than 1024 slots -- decoded by the same warp out
// of the staged tile instead of being handed to the general kernel (which walks run headers with one lane and reads the
// page in place from global memory).  RleDecoder::get_batch (include/reader/rle_decoder.hpp:17-95) restated run by run:
// every lane parses the (warp-uniform) run header; the warp expands the run together -- a literal level run IS validity
// bits (copied word-wise into the image), an RLE run a bit range; index runs go to a per-warp u16 buffer addressed by
// rank.  The page is taken in sub-tiles of 1024 slots; the run state carries over.  Returns false for what the general
// kernel must do (bit width > 16, truncated / zero-length runs: the reference's error paths).
struct RunCursor {
    const uint8_t* s;   // stream bytes (shared)
    uint32_t len;       // bytes
    uint32_t pos;       // next header
    uint32_t rem;       // values left in the current run
    uint32_t lit;       // literal run
    uint32_t val;       // RLE value (unmasked)
    uint32_t bit;       // bit offset of the next literal value
};
// parse the next run header (warp-uniform); false: zero-length run or data beyond the stream
__device__ __forceinline__ bool run_next(RunCursor& c, uint32_t bw) {
    uint32_t ind = 0, shift = 0;
    for (;;) {
        if (c.pos >= c.len || shift > 28u) return false;
        const uint32_t b = c.s[c.pos++];
        ind |= (b & 0x7fu) << shift;
        if (!(b & 0x80u)) break;
        shift += 7u;
    }
    if (ind & 1u) {
        const uint64_t cnt = static_cast<uint64_t>(ind >> 1) * 8u, bytes = (cnt * bw + 7u) >> 3;
        if (cnt == 0 || cnt > 0x7fffffffull || c.pos + bytes > c.len) return false;
        c.rem = static_cast<uint32_t>(cnt); c.lit = 1; c.bit = c.pos * 8u;
        c.pos += static_cast<uint32_t>(bytes);
    } else {
        const uint32_t nb = (bw + 7u) >> 3;
        if ((ind >> 1) == 0 || c.pos + nb > c.len) return false;
        uint32_t v = 0;
        for (uint32_t i = 0; i < nb && i < 4u; i++) v |= static_cast<uint32_t>(c.s[c.pos + i]) << (8u * i);
        c.rem = ind >> 1; c.lit = 0; c.val = v;
        c.pos += nb;
    }
    return true;
}

constexpr uint32_t kOptMaxSlots = 8192; // pages beyond that (in a <= 16 KB tile: bit width <= 8 or mostly nulls) go to the general kernel

template <int W>
__device__ __noinline__ bool opt_page_runs(const DecodeParams& P, const pqg_page_desc& pd, const uint8_t* pg, bool dict_page,
                                           const uint8_t* dictp, uint32_t dict_n, bool dict_in_smem, uint32_t* vwords, uint32_t* rankbase,
                                           uint16_t* idx16) {
    using T = typename FElem<W>::T;
    const uint32_t l = lane_id();
    const uint32_t n = pd.num_values, size = pd.payload_size;
    const uint32_t def_len = ld32u(pg); // (validated by the caller)
    RunCursor lv{pg + 4, def_len, 0, 0, 0, 0, 0}, ix{nullptr, 0, 0, 0, 0, 0, 0};
    uint32_t vpos = 4u + def_len, bw = 0;
    if (dict_page) {
        if (vpos >= size) return false;
        bw = pg[vpos];
        if (bw > 16u) return false;
        ix.s = pg + vpos + 1u; ix.len = size - vpos - 1u;
    }
    const uint8_t* vals = pg + vpos;
    const DictRef dref{dictp, dict_in_smem ? smem_u32(dictp) : 0u};
    const int mode = !dict_page ? -1 : (P.identity_dict ? kDictIdent : (dict_in_smem ? kDictSmem : kDictGlobal));
    uint32_t nn_before = 0;
    bool lv_done = false; // level stream exhausted: the remaining slots are null (rle_decoder.hpp:21-24)
    for (uint32_t ts = 0; ts < n; ts += 1024u) {
        const uint32_t t = min(1024u, n - ts);
        // ---- validity image of the sub-tile
        vwords[l] = 0;
        __syncwarp();
        for (uint32_t produced = 0; produced < t && !lv_done;) {
            if (lv.rem == 0) {
                if (lv.pos >= lv.len) { lv_done = true; break; }
                if (!run_next(lv, 1u)) return false;
            }
            const uint32_t take = min(lv.rem, t - produced);
            if (lv.lit || (lv.val & 0xffu) >= 1u) { // (an RLE run of level 0 leaves zeros)
                const uint32_t w0 = produced >> 5, w1 = (produced + take - 1u) >> 5;
                for (uint32_t w = w0 + l; w <= w1; w += 32) {
                    const uint32_t lo = max(w * 32u, produced), hi = min(w * 32u + 32u, produced + take), cnt = hi - lo;
                    uint32_t bits = cnt >= 32u ? 0xffffffffu : ((1u << cnt) - 1u);
                    if (lv.lit) bits &= ldbits(lv.s, lv.bit + (lo - produced), cnt);
                    if (bits) atomicOr(&vwords[w], bits << (lo & 31u));
                }
            }
            if (lv.lit) lv.bit += take;
            lv.rem -= take;
            produced += take;
        }
        __syncwarp();
        const uint32_t c = __popc(vwords[l]);
        const uint32_t incl = warp_incl_scan(c);
        rankbase[l] = incl - c;
        const uint32_t nn = __shfl_sync(0xffffffffu, incl, 31);
        __syncwarp();
        // ---- the sub-tile's values: dictionary indices by rank, or the PLAIN bounds
        if (dict_page) {
            for (uint32_t produced = 0; produced < nn;) {
                if (ix.rem == 0) {
                    if (ix.pos >= ix.len) { // exhausted: the remaining indices read as 0
                        for (uint32_t k = produced + l; k < nn; k += 32) idx16[k] = 0;
                        break;
                    }
                    if (!run_next(ix, bw)) return false;
                    if (!ix.lit && ix.val > 0xffffu) return false; // (cannot be held in the u16 buffer: out of range anyway, the general kernel nulls it)
                }
                const uint32_t take = min(ix.rem, nn - produced);
                if (ix.lit) {
                    for (uint32_t k = l; k < take; k += 32) idx16[produced + k] = static_cast<uint16_t>(ldbits(ix.s, ix.bit + k * bw, bw));
                    ix.bit += take * bw;
                } else {
                    for (uint32_t k = l; k < take; k += 32) idx16[produced + k] = static_cast<uint16_t>(ix.val);
                }
                ix.rem -= take;
                produced += take;
            }
            __syncwarp();
        } else if (static_cast<uint64_t>(nn_before + nn) * W > size - vpos) return false;
        // ---- emission: one slot per lane, 32 per step (the validity word of a step is warp-uniform)
        T* outp = reinterpret_cast<T*>(P.values) + pd.out_row_base + ts;
        auto emit = [&](auto tag) {
            constexpr int MODE = decltype(tag)::value; // -1: PLAIN
            for (uint32_t j = 0; j < t; j += 32) {
                const uint32_t wv = vwords[j >> 5];
                const bool valid = (wv >> l) & 1u;
                const uint32_t k = rankbase[j >> 5] + __popc(wv & ((1u << l) - 1u));
                T x = T(0);
                bool bad = false;
                if constexpr (MODE < 0) {
                    if (valid) x = ld_elem<W>(vals + static_cast<size_t>(nn_before + k) * W);
                } else {
                    const uint32_t i = valid ? static_cast<uint32_t>(idx16[k]) : 0u;
                    const bool in_range = i < dict_n;
                    bad = valid && !in_range; // NULL in the reference (column_reader.cpp:190-194)
                    const T v = dict_get<T, W, MODE>(dref, in_range ? i : 0u);
                    x = (valid && in_range) ? v : T(0);
                }
                if (j + l < t) st_stream<T>(outp + j + l, x);
                if constexpr (MODE >= 0) {
                    const uint32_t b = __ballot_sync(0xffffffffu, bad);
                    if (b && l == 0) vwords[j >> 5] = wv & ~b;
                }
            }
        };
        if (mode < 0) emit(ModeTag<-1>{});
        else with_dict_mode(mode, emit);
        __syncwarp();
        { // validity: the image shifted to its position in the column's bitmap, one word per lane
            const uint64_t a0 = pd.out_row_base + ts;
            const uint32_t head = static_cast<uint32_t>(a0 & 31u), nwords = (t + 31u) >> 5;
            uint32_t* vp = P.validity + (a0 >> 5);
            const uint32_t cur = l < nwords ? vwords[l] : 0u;
            const uint32_t prev = (l > 0 && l <= nwords) ? vwords[l - 1] : 0u;
            const uint32_t gw = head ? ((cur << head) | (prev >> (32u - head))) : cur;
            const uint32_t total = head + t, gwords = (total + 31u) >> 5;
            if (l < gwords) {
                const bool full = (l > 0 || head == 0) && (l + 1u) * 32u <= total;
                if (full) vp[l] = gw; else if (gw) atomicOr(&vp[l], gw);
            }
            if (l == 0 && gwords > 32u) {
                const uint32_t last = vwords[31] >> (32u - head);
                if (last) atomicOr(&vp[32], last);
            }
        }
        nn_before += nn;
        __syncwarp();
    }
    return true;
}

// OPTIONAL (max_def == 1) pages: definition levels as the writer emits them -- RLE runs
// <varint < 128><level byte> only (src/writer/parquet_writer.cpp:103-135) -- are verified and
// expanded in parallel (one lane per run, warp prefix sum of the run lengths, bit ranges OR-ed
// into a per-warp validity image), ranks come from popcounts, and every lane then emits one
// slot per step: its value (PLAIN shifted load or dictionary lookup) or 0, plus the validity
// word by ballot.  Out-of-range dictionary indices become nulls right here
// (column_reader.cpp:190-194).  Anything else (bit-packed levels, > 1024 slots, ...) -> slow list.
template <int W>
__device__ __forceinline__ void fast_page_opt(const DecodeParams& P, uint32_t q, const pqg_page_desc& pd, const uint8_t* pg,
                                              bool chunk_has_dict, const uint8_t* dictp, uint32_t dict_n, bool dict_in_smem,
                                              uint32_t* vwords, uint32_t* rankbase, uint16_t* idx16) {
    using T = typename FElem<W>::T;
    const uint32_t l = lane_id();
    const uint32_t n = pd.num_values, size = pd.payload_size;
    if (n == 0) return;
    PQG_ASSERT(pd.out_row_base + n <= P.n_slots);
    if (size < 4u) { if (l == 0) to_slow(P, q); return; }
    const uint32_t def_len = ld32u(pg);
    if (def_len > size - 4u) { if (l == 0) to_slow(P, q); return; }
    // a page without nulls (nullable-by-default writers): ONE RLE run of level 1 covering all
    // n slots -- decode it as a REQUIRED page and set its validity range
    if (def_len >= 2u) {
        uint32_t ind = 0, shift = 0, hp = 0;
        while (hp < def_len) { const uint32_t b = pg[4u + hp++]; if (shift < 32) ind |= (b & 0x7Fu) << shift; if (!(b & 0x80u)) break; shift += 7; }
        if (!(ind & 1u) && (ind >> 1) >= n && hp < def_len && pg[4u + hp] == 1u) {
            pqg_page_desc pv = pd;
            pv.payload_size = size - 4u - def_len;
            if (!fast_page<W>(P, q, pv, pg + 4u + def_len, chunk_has_dict, dictp, dict_n, dict_in_smem)) return; // the general kernel redoes it
            const uint64_t a0 = pd.out_row_base, a1 = a0 + n;
            const uint64_t w0 = a0 >> 5, w1 = (a1 - 1) >> 5;
            for (uint64_t w = w0 + l; w <= w1; w += 32) {
                uint32_t m = 0xffffffffu;
                if (w == w0) m &= ~0u << (a0 & 31u);
                if (w == w1 && (a1 & 31u)) m &= (1u << (a1 & 31u)) - 1u;
                if (m == 0xffffffffu) P.validity[w] = m; else atomicOr(&P.validity[w], m);
            }
            return;
        }
    }
    // anything that is not the writer's shape: the run-by-run decode (plans with the index buffer), else the general kernel
    auto general = [&]() {
        const bool dp = (pd.flags & PQG_PAGE_FLAG_DICT) && chunk_has_dict;
        // (pages beyond 4 KB: one or two of them fill a 16 KB tile, i.e. one or two busy warps per CTA -- the general kernel, a
        //  warp per page over all resident warps, is the better shape: measured 0.13 vs 0.26 ms per 4884 8-KB pyarrow pages)
        if (!(idx16 && size <= 4096u && n <= kOptMaxSlots && opt_page_runs<W>(P, pd, pg, dp, dictp, dict_n, dict_in_smem, vwords, rankbase, idx16)) && l == 0) to_slow(P, q);
    };
    if (n > 1024u || (def_len & 1u)) { general(); return; }
    const uint8_t* s = pg + 4;
    const uint32_t nr = def_len >> 1;
    // Validity image without shared-memory atomics: every run boundary where the level flips
    // sets ONE toggle bit at the run's first slot (lanes that hit the same word are combined
    // with match.any + redux.or, one lane writes); the levels are then the prefix XOR of the
    // toggle bits (5 shift-xors inside a word, parity carried across words by a warp scan).
    vwords[l] = 0;
    __syncwarp();
    bool ok = true;
    uint32_t carry = 0, last_lv = 0;
    const uint32_t sdef = smem_u32(s);
    for (uint32_t base = 0; base < nr && carry < n; base += 32) {
        const uint32_t r = base + l;
        uint32_t cnt = 0, lv = 0;
        if (r < nr) { const uint32_t bl = SmemWords{sdef}.u16at(2u * r), b = bl & 0xffu; ok = ok && ((b & 0x81u) == 0u) && b != 0u; cnt = b >> 1; lv = (bl >> 8) >= 1u ? 1u : 0u; }
        const uint32_t incl = warp_incl_scan(cnt);
        const uint32_t start = carry + incl - cnt;
        uint32_t prev = __shfl_up_sync(0xffffffffu, lv, 1);
        if (l == 0) prev = last_lv;
        const bool tog = r < nr && start < n && lv != prev;
        if (tog) atomicXor(&vwords[start >> 5], 1u << (start & 31u));
        const uint32_t last_lane = min(31u, nr - 1u - base);
        last_lv = __shfl_sync(0xffffffffu, lv, last_lane);
        carry += __shfl_sync(0xffffffffu, incl, 31);
        __syncwarp();
    }
    // a stream that ends early leaves the remaining slots at level 0 (rle_decoder.hpp:21-24)
    if (l == 0 && last_lv == 1u && carry < n) vwords[carry >> 5] ^= 1u << (carry & 31u);
    __syncwarp();
    {
        uint32_t t = vwords[l];
        const uint32_t par = __popc(t) & 1u;
        t ^= t << 1; t ^= t << 2; t ^= t << 4; t ^= t << 8; t ^= t << 16;
        uint32_t px = par; // inclusive xor-scan of the word parities
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, px, d); if (l >= static_cast<uint32_t>(d)) px ^= o; }
        if ((px ^ par) & 1u) t = ~t; // odd number of toggles before this word
        const uint32_t lo = l * 32u; // slots beyond n are not part of the page
        t &= lo >= n ? 0u : (n - lo >= 32u ? 0xffffffffu : ((1u << (n - lo)) - 1u));
        vwords[l] = t;
    }
    if (!__all_sync(0xffffffffu, ok)) { general(); return; }
    __syncwarp();
    const uint32_t c = __popc(vwords[l]);
    const uint32_t incl = warp_incl_scan(c);
    rankbase[l] = incl - c;
    const uint32_t nn = __shfl_sync(0xffffffffu, incl, 31);
    __syncwarp();
    uint32_t pos = 4u + def_len;
    const bool dict_page = (pd.flags & PQG_PAGE_FLAG_DICT) && chunk_has_dict;
    uint32_t bw = 0;
    RegStream rs{};
    const uint8_t* vals = pg + pos;
    if (dict_page) {
        if (pos >= size) { if (l == 0) to_slow(P, q); return; }
        bw = pg[pos];
        vals = pg + pos + 1;
        if (bw > 32u || !check_regular2(vals, size - pos - 1u, bw, nn, &rs)) { general(); return; }
    } else if (static_cast<uint64_t>(nn) * W > size - pos) { if (l == 0) to_slow(P, q); return; }
    // emission: 64 page-relative slots per step, two adjacent slots per lane (one 16-byte store
    // for 8-byte values); the validity image is written to global memory afterwards in one step.
    // Branch-free per value: null lanes compute with a clamped rank and discard the result.
    T* outp = reinterpret_cast<T*>(P.values) + pd.out_row_base;
    const bool pair_aligned = (reinterpret_cast<uintptr_t>(outp) & (2 * W - 1)) == 0;
    const uint32_t va = smem_u32(vals);
    const SmemWords ldw{va & ~3u};
    const uint32_t bit0 = (va & 3u) * 8u;
    const uint32_t gs = 1u + bw, imask = bw >= 32u ? 0xffffffffu : ((1u << bw) - 1u);
    const uint32_t kmax = nn ? nn - 1u : 0u;
    const DictRef dref{dictp, dict_in_smem ? smem_u32(dictp) : 0u};
    const int mode = !dict_page ? -1 : (P.identity_dict ? kDictIdent : (dict_in_smem ? kDictSmem : kDictGlobal));
    auto emit = [&](auto tag) {
        constexpr int MODE = decltype(tag)::value; // -1: PLAIN
        auto value_of = [&](uint32_t k, bool valid, bool* bad) -> T {
            if constexpr (MODE < 0) {
                return valid ? ld_elem<W>(vals + k * W) : T(0);
            } else {
                const uint32_t bit = bit0 + (((k >> 3) * gs + 1u) << 3) + (k & 7u) * bw;
                uint32_t ix = __funnelshift_r(ldw(bit >> 5), ldw((bit >> 5) + 1u), bit & 31u) & imask;
                ix = k >= rs.tail_start ? rs.tail_val : ix;
                const bool in_range = ix < dict_n;
                *bad = valid && !in_range; // NULL in the reference (column_reader.cpp:190-194)
                const T x = dict_get<T, W, MODE>(dref, in_range ? ix : 0u);
                return (valid && in_range) ? x : T(0);
            }
        };
        for (uint32_t j = 0; j < n; j += 64) {
            const uint32_t s0 = j + 2u * l;
            if (s0 < n) {
                const uint32_t wv = vwords[s0 >> 5], b = s0 & 31u;
                const bool v0 = (wv >> b) & 1u, v1 = (s0 + 1u < n) && ((wv >> (b + 1u)) & 1u);
                const uint32_t k0 = rankbase[s0 >> 5] + __popc(wv & ((1u << b) - 1u));
                bool bad0 = false, bad1 = false;
                const T x0 = value_of(min(k0, kmax), v0, &bad0);
                const T x1 = value_of(min(k0 + (v0 ? 1u : 0u), kmax), v1, &bad1);
                if (bad0 | bad1) atomicAnd(&vwords[s0 >> 5], ~((bad0 ? 1u : 0u) << b | (bad1 ? 2u : 0u) << b));
                if (s0 + 1u < n && pair_aligned) {
                    if constexpr (W == 8) { __stcs(reinterpret_cast<ulonglong2*>(outp + s0), make_ulonglong2(x0, x1)); }
                    else { __stcs(reinterpret_cast<uint2*>(outp + s0), make_uint2(x0, x1)); }
                } else {
                    st_stream<T>(outp + s0, x0);
                    if (s0 + 1u < n) st_stream<T>(outp + s0 + 1u, x1);
                }
            }
        }
    };
    if (mode < 0) emit(ModeTag<-1>{});
    else with_dict_mode(mode, emit);
    __syncwarp();
    // validity: the page's image shifted to its position in the column's bitmap, one word per lane
    {
        const uint32_t head = static_cast<uint32_t>(pd.out_row_base & 31u), nwords = (n + 31u) >> 5;
        uint32_t* vp = P.validity + (pd.out_row_base >> 5);
        const uint32_t cur = l < nwords ? vwords[l] : 0u;
        const uint32_t prev = (l > 0 && l <= nwords) ? vwords[l - 1] : 0u;
        const uint32_t gw = head ? ((cur << head) | (prev >> (32u - head))) : cur;
        const uint32_t total = head + n, gwords = (total + 31u) >> 5;
        if (l < gwords) {
            const bool full = (l > 0 || head == 0) && (l + 1u) * 32u <= total;
            if (full) vp[l] = gw; else if (gw) atomicOr(&vp[l], gw);
        }
        if (l == 0 && gwords > 32u) { // head pushes the last bits into a 33rd word
            const uint32_t last = vwords[31] >> (32u - head);
            if (last) atomicOr(&vp[32], last);
        }
    }
    __syncwarp();
}

// OPT = the plan has OPTIONAL chunks.  The REQUIRED-only instantiation stays at <= 64 registers
// (4 CTAs/SM); carrying the level code costs 8 more and a whole CTA per SM (measured: PLAIN
// 0.262 -> 0.284 ms, dictionary bw 8 0.177 -> 0.203 ms per 100 M values).
template <int W, int TB, bool OPT>
__global__ void __launch_bounds__(kThreadsPerCta, OPT ? 3 : 4) k_fixed_tiles(const DecodeParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    // (a third stage for OPTIONAL plans changed nothing: 0.285 vs 0.282 ms per 40 M slots, profiles/README.md)
    constexpr int ST = kTileStages;
    uint32_t* vwords = reinterpret_cast<uint32_t*>(smem + tile_pipe_bytes(TB, ST)) + warp_id() * 64; // per warp: 32 validity words + 32 rank bases
    // plans with foreign-looking OPTIONAL pages: + 1024 u16 dictionary indices per warp (opt_page_runs)
    uint16_t* idx16 = (OPT && P.opt_idx) ? reinterpret_cast<uint16_t*>(smem + tile_pipe_bytes(TB, ST) + kLevelScratchBytes) + warp_id() * 1024 : nullptr;
    uint8_t* sdict = smem + tile_pipe_bytes(TB, ST) + kLevelScratchBytes + ((OPT && P.opt_idx) ? kOptIdxBytes : 0);
    uint32_t dict_n = 0;
    const uint8_t* dictp = nullptr;
    bool has_dict = false, dict_in_smem = false;
    int max_def = 0;
    tile_pipeline<TB, ST>(P, smem,
        [&](uint32_t chunk, uint64_t* bar, uint32_t& phase) {
            const DevChunk& ck = P.chunks[chunk];
            has_dict = ck.has_dict;
            max_def = ck.max_def;
            dict_n = ck.dict_ok_n;
            dictp = P.dict_arena + ck.dict_arena_off;
            dict_in_smem = false;
            const uint32_t dbytes = (dict_n * W + 15u) & ~15u;
            if (has_dict && dbytes && dbytes <= P.dict_smem) {
                __syncthreads(); // nobody reads the previous dictionary any more
                if (threadIdx.x == 0) { mbar_expect_tx(bar, dbytes); bulk_g2s(sdict, dictp, dbytes, bar); }
                mbar_wait(bar, phase);
                phase ^= 1;
                dictp = sdict;
                dict_in_smem = true;
            }
        },
        [&](uint32_t q, const pqg_page_desc& pd, const uint8_t* pg) {
            if constexpr (OPT) {
                if (max_def > 0) { fast_page_opt<W>(P, q, pd, pg, has_dict, dictp, dict_n, dict_in_smem, vwords, vwords + 32, idx16); return; }
            }
            fast_page<W>(P, q, pd, pg, has_dict, dictp, dict_n, dict_in_smem);
        });
}

// Partitioned dictionary: ONE chunk per launch (P.chunk_lo), 1024-thread CTAs = kPartGroups groups of 8 warps.  The CTA keeps
// entries [part_lo, part_lo + P.part_entries) of the chunk's dictionary in shared memory (staged once, with one bulk copy);
// its 2^part_bits - 1 siblings (neighbouring block ids) keep the other parts and read the same tiles.  Every group runs its
// own tile pipeline over its own span of tiles and stores only the values whose index falls into the CTA's part.
// Why 1024 threads: with the dictionary part taking 128 KB only one CTA fits an SM, and 8 warps cannot hide the latency of the
// index -> ld.shared -> store chain (measured with 256-thread CTAs: 1.21 ms per 100 M values against 0.49 ms through L2).
constexpr int kPartGroups = 4;
template <int W>
__global__ void __launch_bounds__(kThreadsPerCta * kPartGroups, 1) k_fixed_tiles_part(const DecodeParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* dbar = reinterpret_cast<uint64_t*>(smem);
    uint8_t* sdict = smem + 128;
    uint8_t* pipes = sdict + P.dict_smem;
    const DevChunk& ck = P.chunks[P.chunk_lo];
    const uint32_t dict_n = ck.dict_ok_n;
    const uint32_t part_lo = (blockIdx.x & ((1u << P.part_bits) - 1u)) * P.part_entries;
    const uint32_t cnt = (ck.has_dict && part_lo < dict_n) ? min(P.part_entries, dict_n - part_lo) : 0u;
    const uint32_t pbytes = (cnt * W + 15u) & ~15u;
    if (threadIdx.x == 0) {
        mbar_init(dbar, 1);
        fence_mbar_init();
        if (pbytes) { mbar_expect_tx(dbar, pbytes); bulk_g2s(sdict, P.dict_arena + ck.dict_arena_off + static_cast<size_t>(part_lo) * W, pbytes, dbar); }
    }
    __syncthreads();
    if (pbytes) mbar_wait(dbar, 0);
    const uint32_t group = threadIdx.x / kThreadsPerCta;
    const uint32_t sub = (blockIdx.x >> P.part_bits) * kPartGroups + group; // this group's span of the chunk's tiles
    PipeGroup g;
    g.bar_id = 1u + group;
    g.tid = threadIdx.x % kThreadsPerCta;
    g.warp = g.tid >> 5;
    g.t0 = min(P.tile_hi, P.tile_lo + sub * P.tiles_per_cta);
    g.t1 = min(P.tile_hi, g.t0 + P.tiles_per_cta);
    const bool has_dict = ck.has_dict;
    tile_pipeline<kTileBytes>(P, pipes + group * tile_pipe_bytes(kTileBytes),
        [&](uint32_t, uint64_t*, uint32_t&) {},
        [&](uint32_t q, const pqg_page_desc& pd, const uint8_t* pg) { fast_page<W, true>(P, q, pd, pg, has_dict, sdict, dict_n, true, part_lo); },
        &g);
}

} // namespace

bool chunk_is_tileable(int phys_type, int max_def, int max_rep) {
    const bool w48 = phys_type == PQG_INT32 || phys_type == PQG_FLOAT || phys_type == PQG_INT64 || phys_type == PQG_DOUBLE;
    return w48 && max_def <= 1 && max_rep <= 0;
}

template <int W, int TB, bool OPT>
static cudaError_t launch_tiles_t(DecodeParams p, int sm_count, cudaStream_t s) {
    const size_t smem = static_cast<size_t>(tile_pipe_bytes(TB, kTileStages)) + kLevelScratchBytes + ((OPT && p.opt_idx) ? kOptIdxBytes : 0) + p.dict_smem;
    // (attributes are per device: set on every launch, it is cheap)
    cudaError_t e = cudaFuncSetAttribute(k_fixed_tiles<W, TB, OPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    // contiguous tile spans per CTA (a CTA stages a chunk's dictionary once)
    int resident = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k_fixed_tiles<W, TB, OPT>, kThreadsPerCta, smem);
    const uint32_t grid = tile_grid(p.tile_hi - p.tile_lo, sm_count, resident, &p.tiles_per_cta);
    k_fixed_tiles<W, TB, OPT><<<grid, kThreadsPerCta, smem, s>>>(p);
    return cudaGetLastError();
}

// one chunk (p.chunk_lo, tiles [tile_lo, tile_hi)); p.part_bits / part_entries / dict_smem set by the caller
template <int W>
static cudaError_t launch_part_t(DecodeParams p, int sm_count, cudaStream_t s) {
    const size_t smem = 128 + p.dict_smem + static_cast<size_t>(kPartGroups) * tile_pipe_bytes(kTileBytes);
    cudaError_t e = cudaFuncSetAttribute(k_fixed_tiles_part<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    const uint32_t parts = 1u << p.part_bits, n_tiles = p.tile_hi - p.tile_lo;
    // one wave: sm_count CTAs = sm_count / parts spans x parts; every span is cut into kPartGroups sub-spans
    uint32_t spans = std::max<uint32_t>(1u, static_cast<uint32_t>(sm_count) / parts);
    uint32_t per = (n_tiles + spans * kPartGroups - 1) / (spans * kPartGroups);
    if (per < 2) per = 2;
    p.tiles_per_cta = per;
    spans = (n_tiles + per * kPartGroups - 1) / (per * kPartGroups);
    k_fixed_tiles_part<W><<<spans * parts, kThreadsPerCta * kPartGroups, smem, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_fixed_tiles_part(const DecodeParams& p, int width, int sm_count, cudaStream_t s) {
    if (p.tile_hi <= p.tile_lo) return cudaSuccess;
    if (width == 4) return launch_part_t<4>(p, sm_count, s);
    if (width == 8) return launch_part_t<8>(p, sm_count, s);
    return cudaErrorInvalidValue;
}

cudaError_t launch_fixed_tiles(const DecodeParams& p, int width, int sm_count, cudaStream_t s) {
    if (p.tile_hi <= p.tile_lo) return cudaSuccess;
    // plans with OPTIONAL chunks: 16 KB tiles + the level code; REQUIRED-only plans: the lean kernel
    const bool opt = p.tile_bytes == static_cast<uint32_t>(kTileBytesLarge);
    if (width == 4) return opt ? launch_tiles_t<4, kTileBytesLarge, true>(p, sm_count, s) : launch_tiles_t<4, kTileBytes, false>(p, sm_count, s);
    if (width == 8) return opt ? launch_tiles_t<8, kTileBytesLarge, true>(p, sm_count, s) : launch_tiles_t<8, kTileBytes, false>(p, sm_count, s);
    return cudaErrorInvalidValue;
}

} // namespace pqg
