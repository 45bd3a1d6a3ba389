// pqg_decode.cu -- sm_100a page-decode kernels (HBM-bound integer/byte work, no tensor cores).
//
//   k_dict_prepare   per column chunk: dictionary page -> aligned value table (fixed width:
//                    byte-shifted vector copy) or {start,len} entry table + 16-byte padded
//                    entries for short-string dictionaries (BYTE_ARRAY, one CTA per dictionary).
//                    Replaces ColumnReader::read_dictionary_page (column_reader.cpp:128-138).
//   k_decode_fixed   the GENERAL fixed-width kernel, one warp per data page from the slow list
//                    (what the tile kernel of pqg_tiles.cu does not take): levels + PLAIN /
//                    dictionary values -> values[] + validity bits.  Replaces read_data_page
//                    (column_reader.cpp:140-225) and read_plain_value (:227-268) for
//                    BOOLEAN / INT32 / INT64 / INT96 / FLOAT / DOUBLE.
//   k_str_pages<0>   BYTE_ARRAY pass 1: string bytes per page.
//   k_str_scan_*     exclusive scans: page bases inside a chunk, chunk bases in the column.
//   k_str_pages<1>   BYTE_ARRAY pass 2: Arrow-style offsets + chars.
//
// Grid sizing: CTAs take contiguous spans of the page table (so a CTA stages a chunk's
// dictionary once); the grid is a multiple of the SM count when there is enough work.
#include <algorithm>

#include "pqg_page.cuh"

namespace pqg {
namespace {

template <int W> struct Elem;
template <> struct Elem<1> { using T = uint8_t; };
template <> struct Elem<4> { using T = uint32_t; };
template <> struct Elem<8> { using T = uint64_t; };
struct U96 { uint32_t a, b, c; };
template <> struct Elem<12> { using T = U96; };

template <int W> __device__ __forceinline__ typename Elem<W>::T zero_val();
template <> __device__ __forceinline__ uint8_t zero_val<1>() { return 0; }
template <> __device__ __forceinline__ uint32_t zero_val<4>() { return 0; }
template <> __device__ __forceinline__ uint64_t zero_val<8>() { return 0; }
template <> __device__ __forceinline__ U96 zero_val<12>() { return U96{0, 0, 0}; }

template <int W> __device__ __forceinline__ typename Elem<W>::T load_plain(const uint8_t* p);
template <> __device__ __forceinline__ uint8_t load_plain<1>(const uint8_t* p) { return *p != 0; }
template <> __device__ __forceinline__ uint32_t load_plain<4>(const uint8_t* p) { return ld32u(p); }
template <> __device__ __forceinline__ uint64_t load_plain<8>(const uint8_t* p) { return ld64u(p); }
template <> __device__ __forceinline__ U96 load_plain<12>(const uint8_t* p) {
    return U96{ld32u(p), ld32u(p + 4), ld32u(p + 8)};
}

__device__ __forceinline__ uint4 shift_bytes(uint4 a, uint4 b, uint32_t sh) {
    // bytes sh..sh+15 of the 32-byte pair (a,b); sh is warp-uniform
    uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t bs = (sh & 3u) * 8u;
    uint4 r;
    switch (sh >> 2) {
        case 0: r.x = __funnelshift_r(w[0], w[1], bs); r.y = __funnelshift_r(w[1], w[2], bs); r.z = __funnelshift_r(w[2], w[3], bs); r.w = __funnelshift_r(w[3], w[4], bs); break;
        case 1: r.x = __funnelshift_r(w[1], w[2], bs); r.y = __funnelshift_r(w[2], w[3], bs); r.z = __funnelshift_r(w[3], w[4], bs); r.w = __funnelshift_r(w[4], w[5], bs); break;
        case 2: r.x = __funnelshift_r(w[2], w[3], bs); r.y = __funnelshift_r(w[3], w[4], bs); r.z = __funnelshift_r(w[4], w[5], bs); r.w = __funnelshift_r(w[5], w[6], bs); break;
        default: r.x = __funnelshift_r(w[3], w[4], bs); r.y = __funnelshift_r(w[4], w[5], bs); r.z = __funnelshift_r(w[5], w[6], bs); r.w = __funnelshift_r(w[6], w[7], bs); break;
    }
    return r;
}

// ---------------------------------------------------------------------------------------------
// dictionary preparation
// ---------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(256) k_dict_prepare(DecodeParams P) {
    // grid: (blocks per chunk, chunks)
    DevChunk& ck = P.chunks[P.chunk_lo + blockIdx.y];
    const bool first = blockIdx.x == 0 && threadIdx.x == 0;
    if (!ck.has_dict) { if (first) ck.dict_ok_n = 0; return; }
    const uint8_t* src = P.image + ck.dict_off;
    uint8_t* dst = P.dict_arena + ck.dict_arena_off;
    const uint32_t n = ck.dict_n, size = ck.dict_size;
    if constexpr (W != 0) {
        // PLAIN fixed width: read_plain_value per entry (column_reader.cpp:229-248,257-264)
        uint32_t ok = min(n, size / W);
        if (first) {
            ck.dict_ok_n = ok;
            // the reference throws from ByteBuffer::check while reading entry `ok`
            if (ok < n) report_error(P.err, ck.first_page, PQG_PAGE_DICT_TRUNCATED, ok * W, W, size);
        }
        typename Elem<W>::T* out = reinterpret_cast<typename Elem<W>::T*>(dst);
        const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
        if constexpr (W == 4 || W == 8) {
            // raw bits: a byte-shifted copy in 16-byte vectors (dst is 16-byte aligned in the arena)
            const uint32_t m = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(src) & 15u);
            const uint8_t* a = src - m;
            const uint32_t nvec = static_cast<uint32_t>((static_cast<uint64_t>(ok) * W) >> 4);
            for (uint32_t j = gtid; j < nvec; j += gstride) {
                const uint4 v0 = ldg_nc16(a + 16ull * j);
                reinterpret_cast<uint4*>(dst)[j] = m ? shift_bytes(v0, ldg_nc16(a + 16ull * j + 16), m) : v0;
            }
            for (uint32_t i = nvec * (16u / W) + gtid; i < ok; i += gstride) out[i] = load_plain<W>(src + static_cast<size_t>(i) * W);
        } else {
            for (uint32_t i = gtid; i < ok; i += gstride) out[i] = load_plain<W>(src + static_cast<size_t>(i) * W);
        }
    }
}

// ---- BYTE_ARRAY dictionaries: u32 length prefix + bytes (column_reader.cpp:249-253) -----------------------------
// Entry table {start (byte offset of the chars inside the dictionary payload), len} + a 16-byte padded copy of every
// entry (15 bytes + the length; meaningful when all entries are <= 15 bytes: one vector load per value).
// The prefix chain is sequential by format; it is cut into kDictSeg-byte segments, one THREAD each, over as many CTAs
// as the dictionary needs (a 1.1 MB dictionary: 8.6 K threads; one 1024-thread CTA took 0.31 ms of a 0.56 ms decode):
//   k_dict_seg   every thread collects the first kDictCand positions of its segment that LOOK like a prefix eight links
//                deep (string inside the payload, and so for its 8 successors) and walks the chain from each of them to
//                the end of its segment: (start, end, entries).  More than one candidate, because the bytes in front of
//                a prefix often pass the test too: "x" + the next prefix <0d 00 00> reads as a length whose string ends
//                exactly on a later prefix whenever the entries have one size;
//   k_dict_link  one warp per dictionary links the segments in order: a segment is entered exactly where the previous
//                walk ended -- the candidate that starts there is the true one, no candidate there means the speculation
//                failed (exactness check); 32 segments per step when the chain runs straight through them, one by one
//                where a long string jumps over segments; on a failed link or a truncated page one lane walks the whole
//                page like the reference does, with its error;
//   k_dict_emit  every thread walks its segment again from the picked candidate and writes its entries at its base index.
constexpr uint32_t kSegNone = 0xffffffffu;

__device__ __forceinline__ uint32_t dict_next_of(const uint8_t* src, uint32_t size, uint32_t p) {
    // position after the string whose prefix is at p, kSegNone if it does not fit
    if (static_cast<uint64_t>(p) + 4u > size) return kSegNone;
    const uint32_t len = ld32u(src + p);
    const uint64_t e = static_cast<uint64_t>(p) + 4u + len;
    return e <= size ? static_cast<uint32_t>(e) : kSegNone;
}

__global__ void __launch_bounds__(256) k_dict_seg(DecodeParams P) {
    DevChunk& ck = P.chunks[P.chunk_lo + blockIdx.y];
    if (!ck.has_dict || ck.dict_n == 0) return;
    const uint8_t* src = P.image + ck.dict_off;
    const uint32_t size = ck.dict_size;
    const uint32_t nseg = (size + kDictSeg - 1) / kDictSeg;
    DictSeg* segs = P.dict_segs + ck.dict_seg_first;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nseg; t += gridDim.x * blockDim.x) {
        const uint32_t lo = t * kDictSeg, hi = min(size, lo + kDictSeg);
        DictSeg sg;
#pragma unroll
        for (int c = 0; c < kDictCand; c++) { sg.start[c] = kSegNone; sg.end[c] = kSegNone; sg.cnt[c] = 0; }
        sg.base = kSegNone; sg.pick = 0;
        int nc = 0;
        if (t == 0) { sg.start[0] = 0; nc = 1; }
        else {
            // positions in groups of 8: their first links are independent loads (they overlap); the survivors get the
            // deep test one by one; the scan stops at kDictCand accepted candidates (usually within two entries)
            for (uint32_t g = lo; g < hi && nc < kDictCand; g += 8u) {
                uint32_t m = 0;
#pragma unroll
                for (uint32_t i = 0; i < 8u; i++) {
                    const uint32_t p = g + i;
                    if (p < hi && dict_next_of(src, size, p) != kSegNone) m |= 1u << i;
                }
                while (m && nc < kDictCand) {
                    const uint32_t p = g + static_cast<uint32_t>(__ffs(static_cast<int>(m)) - 1);
                    m &= m - 1;
                    uint32_t q = p;
                    int depth = 0;
                    for (; depth < 8; depth++) { q = dict_next_of(src, size, q); if (q == kSegNone) break; if (q == size) { depth = 8; break; } }
                    if (depth >= 8) {
#pragma unroll
                        for (int c = 0; c < kDictCand; c++) if (c == nc) sg.start[c] = p;
                        nc++;
                    }
                }
            }
        }
        { // the candidates' walks, in lockstep (independent chains: their loads overlap)
            uint32_t p[kDictCand], cnt[kDictCand];
            bool go[kDictCand];
#pragma unroll
            for (int c = 0; c < kDictCand; c++) { p[c] = sg.start[c]; cnt[c] = 0; go[c] = sg.start[c] != kSegNone && p[c] < hi; }
            while (go[0] || go[1] || go[2] || go[3]) {
#pragma unroll
                for (int c = 0; c < kDictCand; c++) {
                    if (!go[c]) continue;
                    const uint32_t q = dict_next_of(src, size, p[c]);
                    if (q == kSegNone) { go[c] = false; continue; }
                    cnt[c]++; p[c] = q;
                    go[c] = q < hi;
                }
            }
#pragma unroll
            for (int c = 0; c < kDictCand; c++) if (sg.start[c] != kSegNone) { sg.end[c] = p[c]; sg.cnt[c] = static_cast<uint16_t>(cnt[c]); }
        }
        segs[t] = sg;
    }
}

static_assert(kDictCand == 4, "k_dict_seg / k_dict_link unroll four candidates");
constexpr int kLinkThreads = 256;
__global__ void __launch_bounds__(kLinkThreads) k_dict_link(DecodeParams P) {
    DevChunk& ck = P.chunks[P.chunk_lo + blockIdx.x];
    const uint32_t tid = threadIdx.x, l = tid & 31u;
    if (!ck.has_dict) { if (tid == 0) { ck.dict_ok_n = 0; ck.dict_minlen = 0xffffffffu; ck.dict_maxlen = 0; } return; }
    const uint8_t* src = P.image + ck.dict_off;
    const uint32_t n = ck.dict_n, size = ck.dict_size;
    const uint32_t nseg = (size + kDictSeg - 1) / kDictSeg;
    DictSeg* segs = P.dict_segs + ck.dict_seg_first;
    // ---- the straight case, kLinkThreads segments per step: every walk ends inside the next segment, exactly one candidate
    //      of every segment starts where a candidate of its left neighbour ends, and those picks form an exact chain from 0
    {
        __shared__ uint32_t s_pe[kLinkThreads], s_w[kLinkThreads / 32], s_bad;
        uint32_t carry_base = 0, carry_cur = 0;
        bool straight = n > 0;
        if (tid == 0) s_bad = 0;
        __syncthreads();
        for (uint32_t c0 = 0; c0 < nseg && straight; c0 += kLinkThreads) {
            const uint32_t t = c0 + tid;
            const bool live = t < nseg;
            uint32_t pstart = kSegNone, pe = kSegNone, pcnt = 0;
            int pick = -1, npick = 0;
            if (live) {
                const DictSeg sg = segs[t];
                uint32_t lend[kDictCand] = {kSegNone, kSegNone, kSegNone, kSegNone};
                if (t > 0) {
#pragma unroll
                    for (int d = 0; d < kDictCand; d++) lend[d] = segs[t - 1].end[d];
                }
#pragma unroll
                for (int c = 0; c < kDictCand; c++) {
                    bool sup = false;
                    if (sg.start[c] != kSegNone) {
                        if (t == 0) sup = sg.start[c] == 0u;
                        else {
#pragma unroll
                            for (int d = 0; d < kDictCand; d++) sup = sup || lend[d] == sg.start[c];
                        }
                    }
                    if (sup) { if (pick < 0) { pick = c; pstart = sg.start[c]; pe = sg.end[c]; pcnt = sg.cnt[c]; } npick++; }
                }
            }
            s_pe[tid] = pe;
            __syncthreads();
            const uint32_t thi = min(size, (t + 1) * kDictSeg), next_hi = min(size, (t + 2) * kDictSeg);
            const uint32_t want = tid == 0 ? carry_cur : s_pe[tid - 1];
            const bool fine = !live || (npick == 1 && pstart == want && pe >= thi && (pe < next_hi || pe == size));
            // block-wide exclusive scan of the entry counts
            const uint32_t incl = warp_incl_scan(pcnt);
            if (l == 31) s_w[tid >> 5] = incl;
            if (!fine) s_bad = 1;
            __syncthreads();
            uint32_t wbase = 0, total = 0;
#pragma unroll
            for (int i = 0; i < kLinkThreads / 32; i++) { const uint32_t x = s_w[i]; if (i < static_cast<int>(tid >> 5)) wbase += x; total += x; }
            if (s_bad || carry_base + total > n) { straight = false; break; }
            if (live) { segs[t].base = carry_base + wbase + incl - pcnt; segs[t].pick = static_cast<uint32_t>(pick); }
            carry_base += total;
            carry_cur = s_pe[min(static_cast<uint32_t>(kLinkThreads) - 1u, nseg - 1u - c0)];
            __syncthreads();
        }
        if (straight && carry_base == n) {
            if (tid == 0) { ck.dict_ok_n = n; ck.dict_minlen = 0xffffffffu; ck.dict_maxlen = 0; }
            return;
        }
    }
    // ---- anything else (a string longer than a segment, trailing bytes behind the entries, ambiguous candidates): one warp,
    //      segment by segment where needed
    __syncthreads();
    for (uint32_t t = tid; t < nseg; t += kLinkThreads) segs[t].base = kSegNone;
    __syncthreads();
    if (tid >= 32) return;
    __shared__ DictSeg sb[32];
    uint32_t cur = 0, base = 0, ok = n > 0 ? 1u : 0u;
    for (uint32_t b0 = 0; b0 < nseg && ok && base < n; b0 += 32) {
        const uint32_t t = b0 + l;
        const bool live = t < nseg;
        DictSeg sg;
        if (live) sg = segs[t];
        else {
#pragma unroll
            for (int c = 0; c < kDictCand; c++) { sg.start[c] = kSegNone; sg.end[c] = kSegNone; sg.cnt[c] = 0; }
        }
        const uint32_t thi = min(size, (t + 1) * kDictSeg), next_hi = min(size, (t + 2) * kDictSeg);
        // straight run: the walk of every segment ends inside the next one (or at the page end).  Then the candidate a
        // segment is entered at is the one some candidate of its left neighbour ends on; picked that way in parallel,
        // then verified as an exact chain from `cur`.
        uint32_t pend[kDictCand];
#pragma unroll
        for (int c = 0; c < kDictCand; c++) pend[c] = __shfl_up_sync(0xffffffffu, sg.end[c], 1);
        int pick = -1, npick = 0;
#pragma unroll
        for (int c = 0; c < kDictCand; c++) {
            bool sup = false;
            if (sg.start[c] != kSegNone) {
                if (l == 0) sup = sg.start[c] == cur;
                else {
#pragma unroll
                    for (int d = 0; d < kDictCand; d++) sup = sup || pend[d] == sg.start[c];
                }
            }
            if (sup) { if (pick < 0) pick = c; npick++; }
        }
        uint32_t pstart = kSegNone, pe = kSegNone, pcnt = 0;
#pragma unroll
        for (int c = 0; c < kDictCand; c++) if (c == pick) { pstart = sg.start[c]; pe = sg.end[c]; pcnt = sg.cnt[c]; }
        const uint32_t left_end = __shfl_up_sync(0xffffffffu, pe, 1);
        const uint32_t want = l == 0 ? cur : left_end;
        const bool fine = !live || (npick == 1 && pstart == want && pe >= thi && (pe < next_hi || pe == size));
        const uint32_t incl = warp_incl_scan(live ? pcnt : 0u);
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (__all_sync(0xffffffffu, fine) && base + total <= n) {
            if (live) { segs[t].base = base + incl - pcnt; segs[t].pick = static_cast<uint32_t>(pick); }
            base += total;
            cur = __shfl_sync(0xffffffffu, pe, min(31u, nseg - 1u - b0));
            continue;
        }
        // one by one (a string longer than a segment, the end of the entries, ambiguous candidates, a failed speculation)
        sb[l] = sg;
        __syncwarp();
        if (l == 0) {
            for (uint32_t i = 0; i < 32u && b0 + i < nseg; i++) {
                const uint32_t tt = b0 + i, tlo = tt * kDictSeg, hh = min(size, tlo + kDictSeg);
                uint32_t bse = kSegNone, pk = 0;
                if (ok && base < n && cur < hh) {              // (cur >= hh: a string spans the whole segment)
                    int c = 0;
                    while (c < kDictCand && sb[i].start[c] != cur) c++;
                    if (c == kDictCand) ok = 0;                 // nobody starts where the chain enters: speculation failed
                    else {
                        bse = base; pk = static_cast<uint32_t>(c);
                        base += sb[i].cnt[c];
                        cur = sb[i].end[c];
                        if (cur < hh && base < n) ok = 0;      // the chain broke before n entries: truncated page
                    }
                }
                segs[tt].base = bse;
                segs[tt].pick = pk;
            }
        }
        cur = __shfl_sync(0xffffffffu, cur, 0);
        base = __shfl_sync(0xffffffffu, base, 0);
        ok = __shfl_sync(0xffffffffu, ok, 0);
        __syncwarp();
    }
    if (base < n) ok = 0;
    // (segments behind the last entry keep base == kSegNone from k_dict_seg; a segment that would run past n entries is
    //  cut by k_dict_emit's k < n)
    if (ok) {
        if (l == 0) { ck.dict_ok_n = n; ck.dict_minlen = 0xffffffffu; ck.dict_maxlen = 0; }
        return;
    }
    // fallback: nothing from the segments; one lane walks the page like the reference (and reports like it)
    for (uint32_t t = l; t < nseg; t += 32) segs[t].base = kSegNone;
    if (l == 0) {
        uint2* ent = reinterpret_cast<uint2*>(P.dict_arena + ck.dict_arena_off);
        uint32_t pos = 0, k = 0, mn = 0xffffffffu;
        for (; k < n; k++) {
            if (static_cast<uint64_t>(pos) + 4 > size) { report_error(P.err, ck.first_page, PQG_PAGE_DICT_TRUNCATED, pos, 4, size); break; }
            const uint32_t len = ld32u(src + pos);
            if (static_cast<uint64_t>(pos) + 4 + len > size) { report_error(P.err, ck.first_page, PQG_PAGE_DICT_TRUNCATED, pos + 4, len, size); break; }
            ent[k] = make_uint2(pos + 4, len);
            mn = min(mn, len);
            pos += 4 + len;
        }
        ck.dict_ok_n = k;
        ck.dict_minlen = mn;
        ck.dict_maxlen = 0xfffffffeu; // no padded table, no common length: the general paths only
    }
}

__global__ void __launch_bounds__(256) k_dict_emit(DecodeParams P) {
    DevChunk& ck = P.chunks[P.chunk_lo + blockIdx.y];
    if (!ck.has_dict || ck.dict_n == 0) return;
    const uint8_t* src = P.image + ck.dict_off;
    const uint32_t n = ck.dict_n, size = ck.dict_size;
    const uint32_t nseg = (size + kDictSeg - 1) / kDictSeg;
    const DictSeg* segs = P.dict_segs + ck.dict_seg_first;
    uint2* ent = reinterpret_cast<uint2*>(P.dict_arena + ck.dict_arena_off);
    uint4* pad = reinterpret_cast<uint4*>(P.dict_arena + ck.dict_pad_off);
    const bool want_pad = !P.skip_dict_pad;
    uint32_t mn = 0xffffffffu, mx = 0;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nseg; t += gridDim.x * blockDim.x) {
        const DictSeg sg = segs[t];
        if (sg.base == kSegNone) continue;
        const uint32_t hi = min(size, (t + 1) * kDictSeg);
        uint32_t p = sg.start[0], k = sg.base;
#pragma unroll
        for (int c = 1; c < kDictCand; c++) if (static_cast<uint32_t>(c) == sg.pick) p = sg.start[c];
        while (p < hi && k < n) {
            const uint32_t q = dict_next_of(src, size, p);
            if (q == kSegNone) break;
            const uint32_t len = q - p - 4u;
            ent[k] = make_uint2(p + 4u, len);
            mn = min(mn, len); mx = max(mx, len);
            if (want_pad) { // zero-padded first 15 bytes + the length in the top byte (used only if every entry is <= 15 bytes)
                const uint8_t* sp = src + p + 4u;
                uint32_t w[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    uint32_t v = 4u * i < len ? ld32u(sp + 4 * i) : 0u;
                    const uint32_t rem = len > 4u * i ? len - 4u * i : 0u;
                    if (rem < 4u) v &= (1u << (8u * rem)) - 1u;
                    w[i] = v;
                }
                pad[k] = make_uint4(w[0], w[1], w[2], (w[3] & 0x00ffffffu) | (min(len, 255u) << 24));
            }
            k++;
            p = q;
        }
    }
    mn = __reduce_min_sync(0xffffffffu, mn); mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31u) == 0 && mx >= mn) { atomicMin(&ck.dict_minlen, mn); atomicMax(&ck.dict_maxlen, mx); }
}

// ---------------------------------------------------------------------------------------------
// PLAIN, REQUIRED, 4/8-byte values: the page payload IS the value array (shifted by the
// payload's misalignment).  Straight global->global copy, 16-byte vectors.
// ---------------------------------------------------------------------------------------------
template <int W>
__device__ __forceinline__ void plain_copy(const DecodeParams& P, uint32_t q, const pqg_page_desc& pd) {
    const uint32_t l = lane_id();
    const uint32_t n = pd.num_values;
    const uint64_t bytes = static_cast<uint64_t>(n) * W;
    if (bytes > pd.payload_size) {
        if (l == 0) report_error(P.err, q, PQG_PAGE_TRUNCATED, (pd.payload_size / W) * W, W, pd.payload_size);
        return;
    }
    const uint8_t* src = P.image + pd.payload_off;
    uint8_t* dst = P.values + pd.out_row_base * W;
    // head: bring dst to a 16-byte boundary (a whole number of elements)
    uint32_t head = static_cast<uint32_t>((16u - (reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u);
    if (head > bytes) head = static_cast<uint32_t>(bytes);
    using T = typename Elem<W>::T;
    if (l < head / W) reinterpret_cast<T*>(dst)[l] = load_plain<W>(src + l * W);
    const uint8_t* s2 = src + head;
    uint8_t* d2 = dst + head;
    const uint32_t nvec = static_cast<uint32_t>((bytes - head) >> 4);
    const uint32_t sh = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(s2) & 15u);
    const uint8_t* a = s2 - sh;
    if (sh == 0) {
#pragma unroll 2
        for (uint32_t j = l; j < nvec; j += 32) reinterpret_cast<uint4*>(d2)[j] = ldg_nc16(a + 16u * j);
    } else {
#pragma unroll 2
        for (uint32_t j = l; j < nvec; j += 32) {
            uint4 v0 = *reinterpret_cast<const uint4*>(a + 16u * j);      // cached: the next lane's
            uint4 v1 = *reinterpret_cast<const uint4*>(a + 16u * j + 16); // v0 is this lane's v1
            reinterpret_cast<uint4*>(d2)[j] = shift_bytes(v0, v1, sh);
        }
    }
    const uint32_t done = head + (nvec << 4);
    const uint32_t tail = static_cast<uint32_t>(bytes - done) / W;
    if (l < tail) reinterpret_cast<T*>(dst + done)[l] = load_plain<W>(src + done + l * W);
}

// ---------------------------------------------------------------------------------------------
// fixed-width data pages
// ---------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T dict_value(const DecodeParams& P, const uint8_t* dictp, uint32_t ix) {
    return reinterpret_cast<const T*>(dictp)[ix];
}
template <> __device__ __forceinline__ uint32_t dict_value<uint32_t>(const DecodeParams& P, const uint8_t* dictp, uint32_t ix) {
    return P.identity_dict ? ix : reinterpret_cast<const uint32_t*>(dictp)[ix]; // dictionary-form output keeps the index
}

template <int W, bool BOOLP>
__device__ __forceinline__ void decode_fixed_page(const DecodeParams& P, uint32_t q, const DevChunk& ck,
                                                  const uint8_t* dictp, WarpScratch& ws) {
    using T = typename Elem<W>::T;
    const uint32_t l = lane_id();
    const pqg_page_desc pd = P.pages[q];
    if (pd.num_values == 0) return;
    const bool dict_page = (pd.flags & PQG_PAGE_FLAG_DICT) && ck.has_dict;
    if constexpr ((W == 4 || W == 8) && !BOOLP) {
        if (!dict_page && ck.max_def <= 0 && ck.max_rep <= 0) { plain_copy<W>(P, q, pd); return; }
    }
    PageCtx c;
    if (!page_begin(P, q, pd, ck, ws, c)) return;
    const uint8_t* vals = c.pg + c.vals_pos;
    const uint32_t vavail = c.size - c.vals_pos;
    const uint32_t T_ = c.wide ? kTileWide : kTileNarrow;
    const bool single = c.n <= T_;
    const uint32_t dict_n = ck.dict_ok_n;
    T* out = reinterpret_cast<T*>(P.values);
    bool regular = false;
    RegStream rs{};
    uint32_t nn_before = 0;
    for (uint32_t ts = 0; ts < c.n; ts += T_) {
        const uint32_t t = min(T_, c.n - ts);
        uint32_t bad = 0;
        const uint32_t nn = levels_tile(c.defw, ws, t, ck.max_def, single, &bad);
        if (bad) { if (l == 0) report_error(P.err, q, bad); return; }
        if (dict_page) {
            if (ts == 0 && (single || !c.has_def)) regular = check_regular2(c.idxw.s, c.idxw.len, c.bw, single ? nn : c.n, &rs);
            if (!regular) {
                indices_tile(c.idxw, ws, nn, c.wide, &bad);
                if (bad) { if (l == 0) report_error(P.err, q, bad); return; }
                __syncwarp();
            }
        } else if (BOOLP) {
            uint32_t need = (nn_before + nn + 7) >> 3;
            if (need > vavail) { if (l == 0) report_error(P.err, q, PQG_PAGE_TRUNCATED, c.size, 1, c.size); return; }
        } else {
            if (static_cast<uint64_t>(nn_before + nn) * W > vavail) {
                if (l == 0) report_error(P.err, q, PQG_PAGE_TRUNCATED, c.vals_pos + (vavail / W) * W, W, c.size);
                return;
            }
        }
        // ---- emit: 32 absolute slots per step, aligned to validity words ----
        const uint64_t abs0 = pd.out_row_base + ts;
        const uint64_t gend = abs0 + t;
        for (uint64_t g = abs0 & ~uint64_t(31); g < gend; g += 32) {
            const int64_t s = static_cast<int64_t>(g + l) - static_cast<int64_t>(abs0);
            const bool in = s >= 0 && s < static_cast<int64_t>(t);
            bool valid = false;
            uint32_t k = 0;
            if (in) {
                uint32_t wv = ws.valid[s >> 5];
                valid = (wv >> (s & 31)) & 1u;
                k = ws.rankbase[s >> 5] + __popc(wv & ((1u << (s & 31)) - 1u));
            }
            T v = zero_val<W>();
            if (valid) {
                if (dict_page) {
                    uint32_t ix = regular ? regular_index2(rs, nn_before + k) : idx_load(ws.idx, k, c.wide);
                    if (ix < dict_n) v = dict_value<T>(P, dictp, ix);
                    else { // out-of-range index -> null (column_reader.cpp:190-194)
                        valid = false;
                        if (!P.validity) atomicAdd(&P.err->bad_index, 1u); // REQUIRED-only plan: pqg_plan_finish adds a validity bitmap and re-runs
                    }
                } else if (BOOLP) {
                    uint32_t kk = nn_before + k;
                    if constexpr (W == 1) v = (vals[kk >> 3] >> (kk & 7u)) & 1u;
                } else {
                    v = load_plain<W>(vals + static_cast<size_t>(nn_before + k) * W);
                }
            }
            if (in) out[g + l] = v;
            if (P.validity) {
                const uint32_t m = __ballot_sync(0xffffffffu, valid), inm = __ballot_sync(0xffffffffu, in);
                if (l == 0) {
                    if (g >= abs0 && g + 32 <= gend) P.validity[g >> 5] = m;
                    else if (ck.max_def <= 0) { if (inm & ~m) atomicAnd(&P.validity[g >> 5], ~(inm & ~m)); } // REQUIRED chunks start all-valid (k_validity_ranges)
                    else if (m) atomicOr(&P.validity[g >> 5], m);
                }
            }
        }
        nn_before += nn;
        __syncwarp();
    }
}

// The general kernel: every page shape, one warp per page, pages taken from the slow list
// (listed by the host for OPTIONAL / BOOLEAN / INT96 / oversized pages, appended by the tile
// kernel for the rest) through a work-stealing cursor.
template <int W, bool BOOLP>
__global__ void __launch_bounds__(kThreadsPerCta) k_decode_fixed(DecodeParams P) {
    extern __shared__ __align__(16) uint8_t smem[];
    WarpScratch& ws = reinterpret_cast<WarpScratch*>(smem)[warp_id()];
    const uint32_t n_host = P.slow_hi - P.slow_lo;
    const uint32_t total = n_host + (P.append_count ? *P.append_count : P.err->slow_count);
    for (;;) {
        uint32_t i = 0;
        if (lane_id() == 0) i = atomicAdd(&P.err->slow_cursor, 1u);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= total) break;
        const uint32_t q = i < n_host ? P.slow_pages[P.slow_lo + i] : P.slow_append[i - n_host];
        PQG_ASSERT(i < n_host || i - n_host < P.slow_cap);
        const DevChunk& ck = P.chunks[P.pages[q].chunk_idx];
        decode_fixed_page<W, BOOLP>(P, q, ck, P.dict_arena + ck.dict_arena_off, ws);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// BYTE_ARRAY
// ---------------------------------------------------------------------------------------------
// copy one string of at most 4 * NW bytes: NW unaligned word loads, then byte stores
template <int NW>
__device__ __forceinline__ void stage_string(uint8_t* dst, const uint8_t* sp, uint32_t len) {
    uint32_t wbuf[NW];
#pragma unroll
    for (int i = 0; i < NW; i++) wbuf[i] = 4u * i < len ? ld32u(sp + 4 * i) : 0u;
#pragma unroll
    for (int i = 0; i < NW; i++) {
        const uint32_t w = wbuf[i], b = 4u * i;
        if (b < len) dst[b] = static_cast<uint8_t>(w);
        if (b + 1u < len) dst[b + 1u] = static_cast<uint8_t>(w >> 8);
        if (b + 2u < len) dst[b + 2u] = static_cast<uint8_t>(w >> 16);
        if (b + 3u < len) dst[b + 3u] = static_cast<uint8_t>(w >> 24);
    }
}

// the same from a padded dictionary entry held in registers (len <= 15)
__device__ __forceinline__ void stage_words(uint8_t* dst, const uint4& v, uint32_t len) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t b = 4u * i;
        if (b < len) dst[b] = static_cast<uint8_t>(w[i]);
        if (b + 1u < len) dst[b + 1u] = static_cast<uint8_t>(w[i] >> 8);
        if (b + 2u < len) dst[b + 2u] = static_cast<uint8_t>(w[i] >> 16);
        if (b + 3u < len) dst[b + 3u] = static_cast<uint8_t>(w[i] >> 24);
    }
}

// the complete 16-byte vectors of a warp's staging buffer -> global memory (aligned), the buffer left zeroed; explicit
// shared-space addressing (the generic form cost 36 warp instructions per 32 vectors, 15 % of the cfg3 copy pass)
__device__ __forceinline__ void flush_stage_vectors(uint32_t stage_s, uint8_t* gdst, uint32_t nfull, bool skip_first) {
    for (uint32_t j = lane_id(); j < nfull; j += 32) {
        uint4 v;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(stage_s + 16u * j));
        asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(stage_s + 16u * j), "r"(0u) : "memory");
        if (j || !skip_first) *reinterpret_cast<uint4*>(gdst + 16u * j) = v; // (first vector of a tile: the bytes in front belong to another page)
    }
}

// ---- copy pass, short-string dictionaries ----------------------------------------------------
// One tile (t <= 1024 slots) of a dictionary page whose index stream is regular and staged in shared
// memory, against a dictionary with the 16-byte padded table (all entries <= 15 bytes): the lean
// form of the loop in decode_str_page -- two groups of 32 slots per trip (both gathers in flight
// together), 32-bit shared-memory addressing, positional index extraction, ONE 16-byte load per
// value, chars packed into the zeroed staging buffer with word-wise red.shared.or (5 per string
// instead of 15 byte stores) and flushed as full aligned 16-byte vectors: the partial vector at
// the end of a trip stays in the buffer for the next one, only the first and last bytes of the
// tile are stored byte-wise.  Validity comes from the tile's image afterwards (words, atomics
// only where a word is shared with a neighbouring page).  Returns the string bytes of the tile.
__device__ __forceinline__ uint32_t copy_short_dict_tile(const DecodeParams& P, WarpScratch& ws, const RegStream& rs, uint32_t nn_before,
                                                         uint32_t t, const uint4* dpad, uint32_t dict_n, uint32_t* offs_tile, uint8_t* dst,
                                                         uint32_t off0, uint64_t abs_slot0, bool& stage_dirty) {
    const uint32_t l = lane_id();
    const uint32_t sa = static_cast<uint32_t>(__cvta_generic_to_shared(rs.s));
    const SmemWords ldw{sa & ~3u};
    const uint32_t bit0 = (sa & 3u) * 8u, bw = rs.bw, gs = 1u + bw;
    const uint32_t imask = bw >= 32u ? 0xffffffffu : ((1u << bw) - 1u);
    const uint32_t stw = static_cast<uint32_t>(__cvta_generic_to_shared(ws.stage)); // 16-byte aligned
    if (stage_dirty) {
        for (uint32_t i = l; i < static_cast<uint32_t>(kStageBytes32) / 16u; i += 32) reinterpret_cast<uint4*>(ws.stage)[i] = make_uint4(0, 0, 0, 0);
        stage_dirty = false;
    }
    __syncwarp();
    const uint32_t mis0 = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(dst) & 15u);
    uint8_t* const abase = dst - mis0;      // aligned-space origin: byte p of the tile's chars sits at abase[mis0 + p]
    uint32_t outp = 0;                      // chars produced so far
    uint32_t sbase = 0;                     // aligned-space position of stage byte 0 (multiple of 16)
    auto fetch = [&](uint32_t s, uint32_t* wv_out, bool* bad) -> uint4 {
        uint4 pv = make_uint4(0, 0, 0, 0);
        if (s < t) {
            const uint32_t wv = ws.valid[s >> 5];
            *wv_out = wv;
            if ((wv >> (s & 31u)) & 1u) {
                const uint32_t k = nn_before + ws.rankbase[s >> 5] + __popc(wv & ((1u << (s & 31u)) - 1u));
                const uint32_t bit = bit0 + (((k >> 3) * gs + 1u) << 3) + (k & 7u) * bw;
                uint32_t ix = __funnelshift_r(ldw(bit >> 5), ldw((bit >> 5) + 1u), bit & 31u) & imask;
                ix = k >= rs.tail_start ? rs.tail_val : ix;
                if (ix < dict_n) pv = ldg_nc16(reinterpret_cast<const uint8_t*>(dpad + ix));
                else *bad = true;   // no value: NULL in the reference (column_reader.cpp:190-194)
            }
        }
        return pv;
    };
    auto put = [&](const uint4& pv, uint32_t apos) { // OR the (zero-padded) string into the stage at aligned-space position apos
        const uint32_t o = apos - sbase, sh = (o & 3u) * 8u, a = stw + (o & ~3u);
        const uint32_t w3 = pv.w & 0x00ffffffu;
        const uint32_t x0 = pv.x << sh, x1 = __funnelshift_l(pv.x, pv.y, sh), x2 = __funnelshift_l(pv.y, pv.z, sh);
        const uint32_t x3 = __funnelshift_l(pv.z, w3, sh), x4 = __funnelshift_l(w3, 0u, sh);
        if (x0) asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(x0) : "memory");
        if (x1) asm volatile("red.shared.or.b32 [%0+4], %1;" ::"r"(a), "r"(x1) : "memory");
        if (x2) asm volatile("red.shared.or.b32 [%0+8], %1;" ::"r"(a), "r"(x2) : "memory");
        if (x3) asm volatile("red.shared.or.b32 [%0+12], %1;" ::"r"(a), "r"(x3) : "memory");
        if (x4) asm volatile("red.shared.or.b32 [%0+16], %1;" ::"r"(a), "r"(x4) : "memory");
    };
    for (uint32_t g = 0; g < t; g += 64) {
        const uint32_t s0 = g + l, s1 = g + 32u + l;
        bool bad0 = false, bad1 = false;
        uint32_t wv0 = 0, wv1 = 0;
        const uint4 p0 = fetch(s0, &wv0, &bad0);
        const uint4 p1 = fetch(s1, &wv1, &bad1);
        const uint32_t len0 = p0.w >> 24, len1 = p1.w >> 24;
        const uint32_t incl0 = warp_incl_scan(len0), incl1 = warp_incl_scan(len1);
        const uint32_t tot0 = __shfl_sync(0xffffffffu, incl0, 31), tot1 = __shfl_sync(0xffffffffu, incl1, 31);
        const uint32_t my0 = outp + incl0 - len0, my1 = outp + tot0 + incl1 - len1;
        if (s0 < t) offs_tile[s0] = off0 + my0;
        if (s1 < t) offs_tile[s1] = off0 + my1;
        // out-of-range indices: the slot turns NULL (after every lane has read its validity word)
        const uint32_t b0 = __ballot_sync(0xffffffffu, bad0), b1 = __ballot_sync(0xffffffffu, bad1);
        if (b0 | b1) {
            __syncwarp();
            if (l == 0) {
                if (b0) ws.valid[g >> 5] &= ~b0;
                if (b1) ws.valid[(g >> 5) + 1u] &= ~b1;
                if (!P.validity) atomicAdd(&P.err->bad_index, __popc(b0) + __popc(b1)); // REQUIRED column: finish re-runs with a validity bitmap
            }
        }
        const uint32_t total = tot0 + tot1;
        if (total) {
            if (len0) put(p0, mis0 + my0);
            if (len1) put(p1, mis0 + my1);
            __syncwarp();
            outp += total;
            const uint32_t end = mis0 + outp;                   // aligned-space end of the staged bytes
            const uint32_t nfull = (end >> 4) - (sbase >> 4);   // complete vectors in the stage
            if (nfull) {
                if (sbase == 0 && mis0) { // first vector: the bytes in front belong to another page
                    if (l >= mis0 && l < 16u) abase[l] = ws.stage[l];
                    __syncwarp();
                }
                flush_stage_vectors(stw, abase + sbase, nfull, sbase == 0 && mis0 != 0);
                __syncwarp();
                if (l == 0) { // the partial vector moves to the front
                    const uint4 v = *reinterpret_cast<const uint4*>(ws.stage + 16u * nfull);
                    *reinterpret_cast<uint4*>(ws.stage + 16u * nfull) = make_uint4(0, 0, 0, 0);
                    *reinterpret_cast<uint4*>(ws.stage) = v;
                }
                sbase += 16u * nfull;
                __syncwarp();
            }
        } else {
            __syncwarp();
        }
    }
    // the bytes behind the last complete vector (and a first vector that never filled up)
    {
        const uint32_t end = mis0 + outp;
        const uint32_t from = sbase == 0 ? mis0 : sbase;
        if (from + l < end) abase[from + l] = ws.stage[from - sbase + l]; // < 16 bytes
        __syncwarp();
        if (l == 0) *reinterpret_cast<uint4*>(ws.stage) = make_uint4(0, 0, 0, 0);
    }
    // validity: the tile's image shifted to its position in the column's bitmap, one word per lane
    if (P.validity) {
        __syncwarp();
        const uint32_t head = static_cast<uint32_t>(abs_slot0 & 31u), nwords = (t + 31u) >> 5;
        uint32_t* vp = P.validity + (abs_slot0 >> 5);
        const uint32_t cur = l < nwords ? ws.valid[l] : 0u;
        const uint32_t prev = (l > 0 && l <= nwords) ? ws.valid[l - 1] : 0u;
        const uint32_t gw = head ? ((cur << head) | (prev >> (32u - head))) : cur;
        const uint32_t totb = head + t, gwords = (totb + 31u) >> 5;
        if (l < gwords) {
            const bool full = (l > 0 || head == 0) && (l + 1u) * 32u <= totb;
            if (full) vp[l] = gw; else if (gw) atomicOr(&vp[l], gw);
        }
        if (l == 0 && gwords > 32u) { // head pushes the last bits into a 33rd word
            const uint32_t last = ws.valid[31] >> (32u - head);
            if (last) atomicOr(&vp[32], last);
        }
    }
    __syncwarp();
    return outp;
}

// ---- copy pass, short-string dictionaries, lanes by RANK -----------------------------------------
// The same job as copy_short_dict_tile for tiles in which EVERY index is known to be in range (2^bw <= dictionary
// entries, the unmasked value of a trailing RLE run checked by the caller), so that the k-th non-null slot of the tile
// is the k-th string of its chars: the chars loop runs over ranks -- no validity / rank lookups, and for dictionaries
// with one common entry length (UNIFORM) no prefix sums either: string k starts at k * ulen.  Two groups of 32
// strings per trip (two gathers in flight); the strings are OR-ed into the zeroed staging buffer word-wise and
// leave as full aligned 16-byte vectors.  Offsets are written by SLOT afterwards, four per lane: the offset of a
// slot is the start of the string of its rank (for a null: of the next string).  Returns the string bytes of the tile.
template <bool UNIFORM>
__device__ __forceinline__ uint32_t copy_ranked_dict_tile(const DecodeParams& P, WarpScratch& ws, const RegStream& rs, uint32_t nn_before,
                                                          uint32_t t, uint32_t nn, const uint4* dpad, uint32_t ulen, uint32_t* offs_tile,
                                                          uint8_t* dst, uint32_t off0, uint64_t abs_slot0, bool& stage_dirty) {
    const uint32_t l = lane_id();
    const uint32_t sa = static_cast<uint32_t>(__cvta_generic_to_shared(rs.s));
    const SmemWords ldw{sa & ~3u};
    const uint32_t bit0 = (sa & 3u) * 8u, bw = rs.bw, gs = 1u + bw;
    const uint32_t imask = bw >= 32u ? 0xffffffffu : ((1u << bw) - 1u);
    const uint32_t stw = static_cast<uint32_t>(__cvta_generic_to_shared(ws.stage)); // 16-byte aligned
    if (stage_dirty) { // (the OR-based copies leave the buffer zeroed: only a byte-staging pass dirties it)
        for (uint32_t i = l; i < static_cast<uint32_t>(kStageBytes32) / 16u; i += 32) reinterpret_cast<uint4*>(ws.stage)[i] = make_uint4(0, 0, 0, 0);
        stage_dirty = false;
    }
    uint16_t* pos16 = reinterpret_cast<uint16_t*>(ws.idx); // start of string k inside the tile's chars (variable lengths only)
    __syncwarp();
    const uint32_t mis0 = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(dst) & 15u);
    uint8_t* const abase = dst - mis0;
    uint32_t outp = 0, sbase = 0;
    auto fetch = [&](uint32_t k) -> uint4 {
        uint4 pv = make_uint4(0, 0, 0, 0);
        if (k < nn) {
            const uint32_t kk = nn_before + k;
            const uint32_t bit = bit0 + (((kk >> 3) * gs + 1u) << 3) + (kk & 7u) * bw;
            uint32_t ix = __funnelshift_r(ldw(bit >> 5), ldw((bit >> 5) + 1u), bit & 31u) & imask;
            ix = kk >= rs.tail_start ? rs.tail_val : ix;
            pv = ldg_nc16(reinterpret_cast<const uint8_t*>(dpad + ix));
        }
        return pv;
    };
    auto put = [&](const uint4& pv, uint32_t apos) {
        const uint32_t o = apos - sbase, sh = (o & 3u) * 8u, a = stw + (o & ~3u);
        PQG_ASSERT((o & ~3u) + 20u <= static_cast<uint32_t>(kStageBytes32));
        const uint32_t w3 = pv.w & 0x00ffffffu;
        const uint32_t x0 = pv.x << sh, x1 = __funnelshift_l(pv.x, pv.y, sh), x2 = __funnelshift_l(pv.y, pv.z, sh);
        const uint32_t x3 = __funnelshift_l(pv.z, w3, sh), x4 = __funnelshift_l(w3, 0u, sh);
        if (x0) asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(x0) : "memory");
        if (x1) asm volatile("red.shared.or.b32 [%0+4], %1;" ::"r"(a), "r"(x1) : "memory");
        if (x2) asm volatile("red.shared.or.b32 [%0+8], %1;" ::"r"(a), "r"(x2) : "memory");
        if (x3) asm volatile("red.shared.or.b32 [%0+12], %1;" ::"r"(a), "r"(x3) : "memory");
        if (x4) asm volatile("red.shared.or.b32 [%0+16], %1;" ::"r"(a), "r"(x4) : "memory");
    };
    // the dictionary entries of the NEXT 64 values are requested before this trip's are staged (four 16-byte gathers in
    // flight per lane: the L2 latency of a 1 MB dictionary was a third of a page's time at 16 warps per SM)
    uint4 q0 = fetch(l), q1 = fetch(32u + l);
    for (uint32_t g = 0; g < nn; g += 64) {
        const uint32_t k0 = g + l, k1 = g + 32u + l;
        const uint4 p0 = q0, p1 = q1;
        if (g + 64u < nn) { q0 = fetch(k0 + 64u); q1 = fetch(k1 + 64u); } // (two trips ahead: 0.407 vs 0.399 ms, no gain)
        uint32_t my0, my1, total;
        if constexpr (UNIFORM) {
            const uint32_t n0 = min(32u, nn - g), n1 = nn - g > 32u ? min(32u, nn - g - 32u) : 0u;
            my0 = outp + l * ulen;
            my1 = outp + (n0 + l) * ulen;
            total = (n0 + n1) * ulen;
        } else {
            const uint32_t len0 = p0.w >> 24, len1 = p1.w >> 24;
            const uint32_t incl0 = warp_incl_scan(len0), incl1 = warp_incl_scan(len1);
            const uint32_t tot0 = __shfl_sync(0xffffffffu, incl0, 31), tot1 = __shfl_sync(0xffffffffu, incl1, 31);
            my0 = outp + incl0 - len0;
            my1 = outp + tot0 + incl1 - len1;
            total = tot0 + tot1;
            if (k0 < nn) pos16[k0] = static_cast<uint16_t>(my0);
            if (k1 < nn) pos16[k1] = static_cast<uint16_t>(my1);
        }
        if (total) {
            if (k0 < nn) put(p0, mis0 + my0);
            if (k1 < nn) put(p1, mis0 + my1);
            __syncwarp();
            outp += total;
            const uint32_t end = mis0 + outp;
            const uint32_t nfull = (end >> 4) - (sbase >> 4);
            if (nfull) {
                if (sbase == 0 && mis0) { // first vector: the bytes in front belong to another page
                    if (l >= mis0 && l < 16u) abase[l] = ws.stage[l];
                    __syncwarp();
                }
                flush_stage_vectors(stw, abase + sbase, nfull, sbase == 0 && mis0 != 0);
                __syncwarp();
                if (l == 0) { // the partial vector moves to the front
                    const uint4 v = *reinterpret_cast<const uint4*>(ws.stage + 16u * nfull);
                    *reinterpret_cast<uint4*>(ws.stage + 16u * nfull) = make_uint4(0, 0, 0, 0);
                    *reinterpret_cast<uint4*>(ws.stage) = v;
                }
                sbase += 16u * nfull;
                __syncwarp();
            }
        } else {
            __syncwarp();
        }
    }
    { // the bytes behind the last complete vector (and a first vector that never filled up)
        const uint32_t end = mis0 + outp;
        const uint32_t from = sbase == 0 ? mis0 : sbase;
        if (from + l < end) abase[from + l] = ws.stage[from - sbase + l]; // < 16 bytes
        __syncwarp();
        if (l == 0) *reinterpret_cast<uint4*>(ws.stage) = make_uint4(0, 0, 0, 0);
    }
    // offsets by slot, four consecutive slots per lane (they share a validity word)
    for (uint32_t s0 = 4u * l; s0 < t; s0 += 128u) {
        const uint32_t wv = ws.valid[s0 >> 5], b = s0 & 31u;
        uint32_t r = ws.rankbase[s0 >> 5] + __popc(wv & ((1u << b) - 1u));
        uint32_t o[4];
#pragma unroll
        for (uint32_t j = 0; j < 4u; j++) {
            if constexpr (UNIFORM) o[j] = off0 + r * ulen;
            else o[j] = off0 + (r < nn ? static_cast<uint32_t>(pos16[r]) : outp);
            r += (wv >> (b + j)) & 1u;
        }
        uint32_t* op = offs_tile + s0;
        if (s0 + 3u < t && (reinterpret_cast<uintptr_t>(op) & 15u) == 0) *reinterpret_cast<uint4*>(op) = make_uint4(o[0], o[1], o[2], o[3]);
        else {
#pragma unroll
            for (uint32_t j = 0; j < 4u; j++) if (s0 + j < t) op[j] = o[j];
        }
    }
    // validity: the tile's image shifted to its position in the column's bitmap, one word per lane
    if (P.validity) {
        const uint32_t head = static_cast<uint32_t>(abs_slot0 & 31u), nwords = (t + 31u) >> 5;
        uint32_t* vp = P.validity + (abs_slot0 >> 5);
        const uint32_t cur = l < nwords ? ws.valid[l] : 0u;
        const uint32_t prev = (l > 0 && l <= nwords) ? ws.valid[l - 1] : 0u;
        const uint32_t gw = head ? ((cur << head) | (prev >> (32u - head))) : cur;
        const uint32_t totb = head + t, gwords = (totb + 31u) >> 5;
        if (l < gwords) {
            const bool full = (l > 0 || head == 0) && (l + 1u) * 32u <= totb;
            if (full) vp[l] = gw; else if (gw) atomicOr(&vp[l], gw);
        }
        if (l == 0 && gwords > 32u) { // head pushes the last bits into a 33rd word
            const uint32_t last = ws.valid[31] >> (32u - head);
            if (last) atomicOr(&vp[32], last);
        }
    }
    __syncwarp();
    return outp;
}

// ---- copy pass, PLAIN pages, lanes by RANK ------------------------------------------------------
// One tile (= the whole page, staged in the shared slot) whose length prefixes were found (ws.idx holds the u16 position
// of every prefix inside the value section, in order) and whose strings are all <= kStageMaxLen bytes.  The chars of
// the page are the value section minus the prefixes: string k starts at pos[k] - 4 k.  Every lane ORs its string into
// the zeroed staging buffer WORD-wise -- aligned words of the slot, one funnel shift for the phase difference between
// source and destination, first / last word masked -- instead of byte by byte (33 byte stores per e-mail address
// before), and the buffer leaves as full aligned 16-byte vectors.  Offsets by slot afterwards (nulls: next string).
template <int NW>
__device__ __forceinline__ void stage_or_words(uint32_t stage_s, uint32_t o, uint32_t sp, uint32_t len) {
    // `len` bytes at shared address `sp` -> stage byte offset `o`
    const uint32_t d = o & 3u, src0 = sp - d, sa = src0 & ~3u, sh = (src0 & 3u) * 8u;
    const uint32_t nw = (d + len + 3u) >> 2, da = stage_s + (o & ~3u), tail = (d + len) & 3u;
    PQG_ASSERT(nw <= static_cast<uint32_t>(NW) && (o & ~3u) + 4u * nw <= static_cast<uint32_t>(kStageBytes32));
    uint32_t wp;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wp) : "r"(sa));
#pragma unroll
    for (int j = 0; j < NW; j++) {
        if (static_cast<uint32_t>(j) < nw) {
            uint32_t wn;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wn) : "r"(sa + 4u * (j + 1)));
            uint32_t x = __funnelshift_r(wp, wn, sh);
            if (j == 0) x &= 0xffffffffu << (8u * d);
            if (static_cast<uint32_t>(j) == nw - 1u && tail) x &= (1u << (8u * tail)) - 1u;
            if (x) asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(da + 4u * j), "r"(x) : "memory");
            wp = wn;
        }
    }
}

__device__ __forceinline__ uint32_t copy_ranked_plain_tile(const DecodeParams& P, WarpScratch& ws, const uint8_t* vals, uint32_t t, uint32_t nn,
                                                           uint32_t maxlen, uint32_t* offs_tile, uint8_t* dst, uint32_t off0, uint64_t abs_slot0,
                                                           bool& stage_dirty) {
    const uint32_t l = lane_id();
    const uint16_t* pos16 = reinterpret_cast<const uint16_t*>(ws.idx);
    const uint32_t stw = static_cast<uint32_t>(__cvta_generic_to_shared(ws.stage));
    const uint32_t vs = static_cast<uint32_t>(__cvta_generic_to_shared(vals));
    if (stage_dirty) {
        for (uint32_t i = l; i < static_cast<uint32_t>(kStageBytes32) / 16u; i += 32) reinterpret_cast<uint4*>(ws.stage)[i] = make_uint4(0, 0, 0, 0);
        stage_dirty = false;
    }
    __syncwarp();
    const uint32_t mis0 = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(dst) & 15u);
    uint8_t* const abase = dst - mis0;
    uint32_t outp = 0, sbase = 0;
    for (uint32_t g = 0; g < nn; g += 32) {
        const uint32_t k = g + l;
        uint32_t len = 0, pp = 0;
        if (k < nn) { pp = pos16[k]; len = SmemWords{vs & ~3u}.u16at((vs & 3u) + pp) | (SmemWords{vs & ~3u}.u16at((vs & 3u) + pp + 2u) << 16); }
        const uint32_t my = pp - 4u * k;                       // start of string k in the page's chars
        const uint32_t last = min(nn, g + 32u) - 1u;           // the group ends where its last string ends
        const uint32_t lend = __shfl_sync(0xffffffffu, my + len, last - g);
        const uint32_t total = lend - outp;
        if (total) {
            if (len) {
                const uint32_t o = mis0 + my - sbase;
                if (maxlen <= 16u) stage_or_words<5>(stw, o, vs + pp + 4u, len);
                else if (maxlen <= 32u) stage_or_words<9>(stw, o, vs + pp + 4u, len);
                else stage_or_words<13>(stw, o, vs + pp + 4u, len);
            }
            __syncwarp();
            outp += total;
            const uint32_t end = mis0 + outp;
            const uint32_t nfull = (end >> 4) - (sbase >> 4);
            if (nfull) {
                if (sbase == 0 && mis0) { // first vector: the bytes in front belong to another page
                    if (l >= mis0 && l < 16u) abase[l] = ws.stage[l];
                    __syncwarp();
                }
                flush_stage_vectors(stw, abase + sbase, nfull, sbase == 0 && mis0 != 0);
                __syncwarp();
                if (l == 0) { // the partial vector moves to the front
                    const uint4 v = *reinterpret_cast<const uint4*>(ws.stage + 16u * nfull);
                    *reinterpret_cast<uint4*>(ws.stage + 16u * nfull) = make_uint4(0, 0, 0, 0);
                    *reinterpret_cast<uint4*>(ws.stage) = v;
                }
                sbase += 16u * nfull;
                __syncwarp();
            }
        } else {
            __syncwarp();
        }
    }
    { // the bytes behind the last complete vector (and a first vector that never filled up)
        const uint32_t end = mis0 + outp;
        const uint32_t from = sbase == 0 ? mis0 : sbase;
        if (from + l < end) abase[from + l] = ws.stage[from - sbase + l]; // < 16 bytes
        __syncwarp();
        if (l == 0) *reinterpret_cast<uint4*>(ws.stage) = make_uint4(0, 0, 0, 0);
    }
    // offsets by slot, four consecutive slots per lane (they share a validity word)
    for (uint32_t s0 = 4u * l; s0 < t; s0 += 128u) {
        const uint32_t wv = ws.valid[s0 >> 5], b = s0 & 31u;
        uint32_t r = ws.rankbase[s0 >> 5] + __popc(wv & ((1u << b) - 1u));
        uint32_t o[4];
#pragma unroll
        for (uint32_t j = 0; j < 4u; j++) {
            o[j] = off0 + (r < nn ? static_cast<uint32_t>(pos16[r]) - 4u * r : outp);
            r += (wv >> (b + j)) & 1u;
        }
        uint32_t* op = offs_tile + s0;
        if (s0 + 3u < t && (reinterpret_cast<uintptr_t>(op) & 15u) == 0) *reinterpret_cast<uint4*>(op) = make_uint4(o[0], o[1], o[2], o[3]);
        else {
#pragma unroll
            for (uint32_t j = 0; j < 4u; j++) if (s0 + j < t) op[j] = o[j];
        }
    }
    if (P.validity) {
        const uint32_t head = static_cast<uint32_t>(abs_slot0 & 31u), nwords = (t + 31u) >> 5;
        uint32_t* vp = P.validity + (abs_slot0 >> 5);
        const uint32_t cur = l < nwords ? ws.valid[l] : 0u;
        const uint32_t prev = (l > 0 && l <= nwords) ? ws.valid[l - 1] : 0u;
        const uint32_t gw = head ? ((cur << head) | (prev >> (32u - head))) : cur;
        const uint32_t totb = head + t, gwords = (totb + 31u) >> 5;
        if (l < gwords) {
            const bool full = (l > 0 || head == 0) && (l + 1u) * 32u <= totb;
            if (full) vp[l] = gw; else if (gw) atomicOr(&vp[l], gw);
        }
        if (l == 0 && gwords > 32u) {
            const uint32_t last = ws.valid[31] >> (32u - head);
            if (last) atomicOr(&vp[32], last);
        }
    }
    __syncwarp();
    return outp;
}

// LEAN: the plan has dictionary chunks -- compile the short-string-dictionary copy path in (kept out of the
// instantiation that PLAIN-only plans run: its registers cost the PLAIN path 7 %)
template <bool COPY, bool LEAN>
__device__ __forceinline__ void decode_str_page(const DecodeParams& P, uint32_t q, const pqg_page_desc& pd, const DevChunk& ck, WarpScratch& ws,
                                                const uint8_t* prestaged, bool& stage_dirty) {
    const uint32_t l = lane_id();
    if (pd.num_values == 0) { if (!COPY && l == 0) P.page_chars[q] = 0; return; }
    if constexpr (!COPY) {
        // Size pass, dictionaries whose entries all have one length (ck.dict_len) and whose size covers every bw-bit
        // index: the page's bytes are that length times its present slots -- counted from the definition-level runs
        // right where they lie in global memory (the writer's <varint < 128><level> layout), nothing staged, no index
        // decoded.  Optimistic in one point: an RLE run INSIDE the index stream carries an unmasked value that may be
        // out of range (a null in the reference, so fewer bytes); the copy pass compares every page's bytes with this
        // count and the plan falls back to the exact size pass (P.exact_sizes) on a mismatch.
        // PLAIN pages the same way: the value section is exactly its present values (4-byte prefix + bytes each), so the
        // page's bytes are the section minus 4 per present slot -- header arithmetic, verified by the copy pass as well.
        const bool dict_page = (pd.flags & PQG_PAGE_FLAG_DICT) && ck.has_dict;
        if (!P.exact_sizes && (!dict_page || ck.dict_len() != 0xffffffffu) && ck.max_rep <= 0 && ck.max_def <= 1) {
            const uint8_t* src = P.image + pd.payload_off;
            const uint32_t size = pd.payload_size, n = pd.num_values;
            uint32_t pos = 0, nn = n;
            bool ok = true;
            if (ck.max_def == 1) {
                ok = size >= 4u;
                const uint32_t def_len = ok ? ld32u(src) : 0u;
                ok = ok && def_len <= size - 4u && !(def_len & 1u);
                if (ok) {
                    const uint8_t* s = src + 4;
                    const uint32_t nr = def_len >> 1;
                    uint32_t total = 0, present = 0;
                    for (uint32_t r = l; r < nr; r += 32) {
                        const uint32_t b = s[2 * r], v = s[2 * r + 1];
                        ok = ok && ((b & 0x81u) == 0u) && b != 0u;
                        total += b >> 1;
                        if (v >= 1u) present += b >> 1;
                    }
                    total = __reduce_add_sync(0xffffffffu, total);
                    nn = __reduce_add_sync(0xffffffffu, present);
                    ok = __all_sync(0xffffffffu, ok) && total <= n; // (runs reaching past the page: exact path)
                    pos = 4u + def_len;
                }
            }
            if (ok && dict_page) {
                const uint32_t bw = pos < size ? src[pos] : 99u;
                if (bw < 32u && ck.dict_ok_n >= (1u << bw)) {
                    if (l == 0) P.page_chars[q] = ck.dict_len() * nn;
                    return;
                }
            } else if (ok && static_cast<uint64_t>(pos) + 4ull * nn <= size) {
                if (l == 0) P.page_chars[q] = size - pos - 4u * nn;
                return;
            }
        }
    }
    PageCtx c;
    if (!page_begin(P, q, pd, ck, ws, c, prestaged)) { if (!COPY && l == 0) P.page_chars[q] = 0; return; }
    const uint8_t* vals = c.pg + c.vals_pos;
    const uint32_t vavail = c.size - c.vals_pos;
    // plain pages keep prefix positions in ws.idx: u16 unless the page is large
    const bool pwide = !c.dict && c.size > 65535u;
    const bool wide = c.dict ? c.wide : pwide;
    const uint32_t T_ = wide ? kTileWide : kTileNarrow;
    const bool single = c.n <= T_;
    const uint2* dent = reinterpret_cast<const uint2*>(P.dict_arena + ck.dict_arena_off);
    const uint8_t* dchars = P.image + ck.dict_off;
    const uint32_t dict_n = ck.dict_ok_n;
    const bool dshort = c.dict && ck.dict_short();
    const uint4* dpad = reinterpret_cast<const uint4*>(P.dict_arena + ck.dict_pad_off);
    uint32_t* offs = COPY ? P.offsets + ck.out_row_base + (&ck - P.chunks) : nullptr; // chunk c owns [row_base + c, ...]
    uint8_t* chars = COPY ? P.chars + ck.char_base : nullptr;
    uint64_t page_bytes = 0;                       // running string bytes of this page
    const uint32_t page_base = COPY ? P.page_char_base[q] : 0;
    bool regular = false, plain_found = false;
    RegStream rs{};
    uint32_t nn_before = 0, wpos = 0, plain_maxlen = 0xffffffffu;
    for (uint32_t ts = 0; ts < c.n; ts += T_) {
        const uint32_t t = min(T_, c.n - ts);
        uint32_t bad = 0;
        const uint32_t nn = levels_tile(c.defw, ws, t, ck.max_def, single, &bad);
        if (bad) { if (l == 0) { report_error(P.err, q, bad); if (!COPY) P.page_chars[q] = 0; } return; }
        uint32_t tile_pos0 = wpos;
        if (c.dict) {
            if (ts == 0 && (single || !c.has_def)) regular = check_regular2(c.idxw.s, c.idxw.len, c.bw, single ? nn : c.n, &rs);
            if (!regular) {
                indices_tile(c.idxw, ws, nn, c.wide, &bad);
                if (bad) { if (l == 0) { report_error(P.err, q, bad); if (!COPY) P.page_chars[q] = 0; } return; }
                __syncwarp();
            }
        } else {
            // parallel length-prefix discovery first (text pages); sequential walk otherwise
            bool found = false;
            if (single && !wide) {
                // values of ONE length: the section is nn x (4 + len) bytes and string k sits at k * stride -- verified exactly
                // (every prefix reads len: by induction those are the prefixes of the chain); no candidate search
                const uint32_t stride = nn ? vavail / nn : 0u;
                if (nn && stride >= 4u && stride * nn == vavail && vavail <= 65535u && nn <= static_cast<uint32_t>(kIdxWords) * 2u) {
                    bool same = true;
                    for (uint32_t k = l; k < nn; k += 32) same = same && ld32u(vals + k * stride) == stride - 4u;
                    if (__all_sync(0xffffffffu, same)) {
                        for (uint32_t k = l; k < nn; k += 32) reinterpret_cast<uint16_t*>(ws.idx)[k] = static_cast<uint16_t>(k * stride);
                        __syncwarp();
                        found = true;
                        wpos = vavail;
                        plain_maxlen = stride - 4u;
                    }
                }
                if (!found) {
                    uint32_t endp = 0;
                    found = find_headers(vals, vavail, nn, reinterpret_cast<uint16_t*>(ws.idx), static_cast<uint32_t>(kIdxWords) * 2u, &endp);
                    if (found) {
                        wpos = endp;
                        if (COPY) { // longest string of the page (the word-wise staging takes <= kStageMaxLen bytes per string)
                            uint32_t mx = 0;
                            for (uint32_t k = l; k < nn; k += 32) mx = max(mx, ld32u(vals + reinterpret_cast<const uint16_t*>(ws.idx)[k]));
                            plain_maxlen = __reduce_max_sync(0xffffffffu, mx);
                        }
                    }
                }
                plain_found = found;
            }
            uint32_t epos = 0, eneed = 0;
            if (!found && !walk_strings(vals, vavail, &wpos, nn, ws, wide, COPY, &epos, &eneed)) {
                if (l == 0) { report_error(P.err, q, PQG_PAGE_TRUNCATED, c.vals_pos + epos, eneed, c.size); if (!COPY) P.page_chars[q] = 0; }
                return;
            }
        }
        if (!COPY) {
            if (c.dict) { // sum the lengths of the referenced dictionary entries
                uint32_t sum = 0;
                const uint32_t ulen = ck.dict_len(); // every entry has this length (~0u: lengths differ)
                // one common length and no index can be out of range (2^bw <= entries; the unmasked value
                // of a trailing RLE run checked apart): the page's bytes follow from the count alone
                const bool all_in_range = regular && c.bw < 32u && dict_n >= (1u << c.bw) &&
                                          (rs.tail_start >= nn_before + nn || rs.tail_val < dict_n);
                if (ulen != 0xffffffffu && all_in_range) {
                    if (l == 0) sum = ulen * nn;
                } else {
                    for (uint32_t k = l; k < nn; k += 32) {
                        uint32_t ix = regular ? regular_index2(rs, nn_before + k) : idx_load(ws.idx, k, c.wide);
                        if (ix < dict_n) sum += ulen != 0xffffffffu ? ulen : dent[ix].y;
                    }
                }
                for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
                page_bytes += sum;
            } else {
                page_bytes += (wpos - tile_pos0) - 4ull * nn;
            }
            nn_before += nn;
            __syncwarp();
            continue;
        }
        // ---- COPY: offsets, then chars, 32 slots per step ----
        const uint64_t slot0 = (pd.out_row_base - ck.out_row_base) + ts; // chunk-relative slot
        if constexpr (LEAN) if (dshort && regular && c.size <= static_cast<uint32_t>(kSlotBytes)) { // (the page is in the shared slot)
            // every index in range (the trailing RLE run's unmasked value included): the k-th present slot is the k-th string
            const bool all_in_range = c.bw < 32u && dict_n >= (1u << c.bw) && (rs.tail_start >= nn_before + nn || rs.tail_val < dict_n);
            if (all_in_range) {
                const uint32_t ulen = ck.dict_len();
                uint8_t* dstp = chars + page_base + page_bytes;
                const uint32_t o0 = static_cast<uint32_t>(page_base + page_bytes);
                page_bytes += ulen != 0xffffffffu
                    ? copy_ranked_dict_tile<true>(P, ws, rs, nn_before, t, nn, dpad, ulen, offs + slot0, dstp, o0, pd.out_row_base + ts, stage_dirty)
                    : copy_ranked_dict_tile<false>(P, ws, rs, nn_before, t, nn, dpad, 0u, offs + slot0, dstp, o0, pd.out_row_base + ts, stage_dirty);
                nn_before += nn;
                __syncwarp();
                continue;
            }
            page_bytes += copy_short_dict_tile(P, ws, rs, nn_before, t, dpad, dict_n, offs + slot0, chars + page_base + page_bytes,
                                               static_cast<uint32_t>(page_base + page_bytes), pd.out_row_base + ts, stage_dirty);
            nn_before += nn;
            __syncwarp();
            continue;
        }
        if (!c.dict && plain_found && plain_maxlen <= static_cast<uint32_t>(kStageMaxLen) && c.size <= static_cast<uint32_t>(kSlotBytes)) {
            // PLAIN page in the shared slot, prefixes known, short strings: lanes by rank, word-wise staging
            page_bytes += copy_ranked_plain_tile(P, ws, vals, t, nn, plain_maxlen, offs + slot0, chars + page_base + page_bytes,
                                                 static_cast<uint32_t>(page_base + page_bytes), pd.out_row_base + ts, stage_dirty);
            nn_before += nn;
            __syncwarp();
            continue;
        }
        stage_dirty = true; // the byte-staging path below writes the buffer without clearing it
        for (uint32_t g = 0; g < t; g += 32) {
            const uint32_t s = g + l;
            const bool in = s < t;
            bool valid = false;
            uint32_t k = 0;
            if (in) {
                uint32_t wv = ws.valid[s >> 5];
                valid = (wv >> (s & 31)) & 1u;
                k = ws.rankbase[s >> 5] + __popc(wv & ((1u << (s & 31)) - 1u));
            }
            uint32_t len = 0;
            const uint8_t* sp = nullptr; // source bytes of this lane's string
            uint4 pv = make_uint4(0, 0, 0, 0);   // short-string dictionaries: the padded entry itself
            if (valid) {
                if (c.dict) {
                    uint32_t ix = regular ? regular_index2(rs, nn_before + k) : idx_load(ws.idx, k, c.wide);
                    if (ix >= dict_n) { valid = false; if (!P.validity) atomicAdd(&P.err->bad_index, 1u); } // null in the reference
                    else if (dshort) { pv = ldg_nc16(reinterpret_cast<const uint8_t*>(dpad + ix)); len = pv.w >> 24; }
                    else { uint2 e = dent[ix]; sp = dchars + e.x; len = e.y; }
                } else {
                    uint32_t pp = idx_load(ws.idx, k, wide);
                    len = ld32u(vals + pp);
                    sp = vals + pp + 4;
                }
            }
            const uint32_t incl = warp_incl_scan(len);
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            const uint64_t gbase = page_base + page_bytes;          // chunk-relative byte offset of this group
            const uint32_t myoff = incl - len;                      // group-relative
            if (in) offs[slot0 + s] = static_cast<uint32_t>(gbase + myoff);
            if (P.validity) {
                // validity words are indexed by absolute slot; groups are not word aligned in general
                uint32_t m = __ballot_sync(0xffffffffu, valid);
                uint64_t a0 = pd.out_row_base + ts + g;
                uint32_t sh = static_cast<uint32_t>(a0 & 31u);
                if (l == 0 && m) {
                    atomicOr(&P.validity[a0 >> 5], m << sh);
                    if (sh && (m >> (32u - sh))) atomicOr(&P.validity[(a0 >> 5) + 1], m >> (32u - sh));
                }
            }
            // chars of the group: bytes [0,total) go to dstg.
            if (total) {
                uint8_t* dstg = chars + gbase;
                const uint32_t maxlen = __reduce_max_sync(0xffffffffu, len);
                if (maxlen <= static_cast<uint32_t>(kStageMaxLen)) {
                    // short strings: every lane copies its own string into the warp's staging
                    // buffer (same 16-byte phase as the destination), then the warp flushes the
                    // contiguous bytes with aligned 16-byte vectors
                    const uint32_t mis = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(dstg) & 15u);
                    uint8_t* st = ws.stage + mis;
                    // all loads first (the compiler cannot prove that the staging stores do not
                    // alias the source, and would otherwise serialise one L2 round trip per byte);
                    // unrolled for the warp's longest string: 4, 8 or 12 words
                    if (dshort) stage_words(st + myoff, pv, len);
                    else if (maxlen <= 16u) stage_string<4>(st + myoff, sp, len);
                    else if (maxlen <= 32u) stage_string<8>(st + myoff, sp, len);
                    else stage_string<kStageMaxLen / 4>(st + myoff, sp, len);
                    __syncwarp();
                    const uint32_t head = min(total, (16u - mis) & 15u);
                    if (l < head) dstg[l] = st[l];
                    const uint32_t body = (total - head) >> 4;
                    for (uint32_t j = l; j < body; j += 32)
                        *reinterpret_cast<uint4*>(dstg + head + 16u * j) = *reinterpret_cast<const uint4*>(st + head + 16u * j);
                    const uint32_t done = head + 16u * body;
                    if (l < total - done) dstg[done + l] = st[done + l];
                    __syncwarp();
                } else {
                    // long strings: the whole warp copies one string at a time, 32 consecutive bytes per step
                    const unsigned long long spl = reinterpret_cast<unsigned long long>(sp);
                    for (int k = 0; k < 32; k++) {
                        const uint32_t lk = __shfl_sync(0xffffffffu, len, k);
                        if (!lk) continue;
                        const uint8_t* sk = reinterpret_cast<const uint8_t*>(__shfl_sync(0xffffffffu, spl, k));
                        uint8_t* dk = dstg + __shfl_sync(0xffffffffu, myoff, k);
                        for (uint32_t b = l; b < lk; b += 32) dk[b] = sk[b];
                    }
                }
            }
            page_bytes += total;
        }
        nn_before += nn;
        __syncwarp();
    }
    if (!COPY) {
        if (l == 0) {
            if (page_bytes > 0xffffffffull) { report_error(P.err, q, PQG_PAGE_CHARS_OVERFLOW); page_bytes = 0; }
            P.page_chars[q] = static_cast<uint32_t>(page_bytes);
        }
    } else {
        // byte counts taken from the page headers (no size pass) hold only if the page is exactly its values
        if (P.check_layout && !c.dict && l == 0 && page_bytes != static_cast<uint64_t>(vavail) - 4ull * nn_before)
            report_error(P.err, q, PQG_PAGE_LAYOUT);
        // the size pass counted this page optimistically (see there): hold it to its word
        if (!P.check_layout && !P.exact_sizes && l == 0 && page_bytes != P.page_chars[q]) report_error(P.err, q, PQG_PAGE_LAYOUT);
        // the last page of the chunk closes the Arrow offsets array
        if (l == 0 && q + 1 == ck.first_page + ck.n_pages)
            offs[(pd.out_row_base - ck.out_row_base) + c.n] = static_cast<uint32_t>(page_base + page_bytes);
    }
}

// ---- copy pass, the dominant page shape of PLAIN string columns, without the generic page context -----------------
// REQUIRED PLAIN page staged in shared memory whose values all have ONE length <= kStageMaxLen (header arithmetic: the
// section is n x (4 + len) bytes; every prefix is checked, so the positions are exact).  No levels, no validity, no
// candidate search, no position table: string k sits at k * stride + 4, its chars go to k * len, its offset is base +
// k * len.  Same word-wise staging and vector flush as copy_ranked_plain_tile.  Returns false when the page is not of
// that shape (the caller takes the general path).
__device__ __forceinline__ bool lean_plain_page(const DecodeParams& P, uint32_t q, const pqg_page_desc& pd, const DevChunk& ck, WarpScratch& ws,
                                                const uint8_t* buf, bool& stage_dirty) {
    const uint32_t l = lane_id();
    const uint32_t n = pd.num_values, size = pd.payload_size;
    if (n == 0 || n > 1024u || size > static_cast<uint32_t>(kSlotBytes)) return false;
    const uint32_t stride = size / n;
    if (stride < 4u || stride * n != size) return false;
    const uint32_t ulen = stride - 4u;
    const uint32_t vs = static_cast<uint32_t>(__cvta_generic_to_shared(buf)) + static_cast<uint32_t>(pd.payload_off & 15u);
    const SmemWords ldw{vs & ~3u};
    bool same = true;
    for (uint32_t k = l; k < n; k += 32) {
        const uint32_t a = (vs & 3u) + k * stride;
        same = same && __funnelshift_r(ldw(a >> 2), ldw((a >> 2) + 1u), (a & 3u) * 8u) == ulen;
    }
    if (!__all_sync(0xffffffffu, same)) return false;
    const uint32_t page_base = P.page_char_base[q];
    uint32_t* offs = P.offsets + ck.out_row_base + (&ck - P.chunks) + (pd.out_row_base - ck.out_row_base);
    for (uint32_t k = l; k < n; k += 32) offs[k] = page_base + k * ulen;
    const uint32_t total = n * ulen;
    if (ulen >= 4u) {
        // Strings of four bytes and more: no staging at all.  Output byte j is source byte j + 4 * (j / ulen + 1), so a
        // lane builds one aligned 16-byte vector of the chars out of the page in shared memory -- per output word two
        // unaligned source words (the one at the byte, the one past the next prefix) and one byte permute that takes the
        // first `left in this string` bytes from the first and the rest from the second.  j / ulen: multiply-high by
        // ceil(2^32 / ulen), exact below 2^16.
        uint8_t* dst = P.chars + ck.char_base + page_base;
        const uint32_t vo = vs & 3u;
        const uint32_t magic = 0xffffffffu / ulen + 1u;
        const uint32_t head = min(total, (16u - static_cast<uint32_t>(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u);
        const uint32_t nvec = (total - head) >> 4;
        for (uint32_t v = l; v < nvec; v += 32) {
            uint32_t j = head + 16u * v;
            uint32_t k = __umulhi(j, magic), r = j - k * ulen;
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t a = vo + j + 4u * k + 4u, sh = (a & 3u) * 8u;
                const uint32_t x0 = ldw(a >> 2), x1 = ldw((a >> 2) + 1u), x2 = ldw((a >> 2) + 2u);
                const uint32_t c = min(ulen - r, 4u); // bytes of this word that string k still has
                w[i] = __byte_perm(__funnelshift_r(x0, x1, sh), __funnelshift_r(x1, x2, sh), 0x3210u + ((0x4444u << (4u * c)) & 0xffffu));
                j += 4u; r += 4u;
                if (r >= ulen) { r -= ulen; k++; }
            }
            *reinterpret_cast<uint4*>(dst + head + 16u * v) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        // the bytes in front of the first and behind the last aligned vector (< 16 each): lanes 0..15 / 16..31
        const uint32_t jb = l < 16u ? l : head + 16u * nvec + (l - 16u);
        if (l < 16u ? l < head : jb < total) {
            const uint32_t a = vs + jb + 4u * __umulhi(jb, magic) + 4u;
            uint32_t b;
            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(b) : "r"(a));
            dst[jb] = static_cast<uint8_t>(b);
        }
    } else if (ulen) {
        uint8_t* dst = P.chars + ck.char_base + page_base;
        const uint32_t stw = static_cast<uint32_t>(__cvta_generic_to_shared(ws.stage));
        if (stage_dirty) {
            for (uint32_t i = l; i < static_cast<uint32_t>(kStageBytes32) / 16u; i += 32) reinterpret_cast<uint4*>(ws.stage)[i] = make_uint4(0, 0, 0, 0);
            stage_dirty = false;
        }
        __syncwarp();
        const uint32_t mis0 = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(dst) & 15u);
        uint8_t* const abase = dst - mis0;
        uint32_t outp = 0, sbase = 0;
        for (uint32_t g = 0; g < n; g += 32) {
            const uint32_t k = g + l, cnt = min(32u, n - g);
            if (k < n) {
                const uint32_t o = mis0 + k * ulen - sbase, sp = vs + k * stride + 4u;
                if (ulen <= 16u) stage_or_words<5>(stw, o, sp, ulen);
                else if (ulen <= 32u) stage_or_words<9>(stw, o, sp, ulen);
                else stage_or_words<13>(stw, o, sp, ulen);
            }
            __syncwarp();
            outp += cnt * ulen;
            const uint32_t end = mis0 + outp;
            const uint32_t nfull = (end >> 4) - (sbase >> 4);
            if (nfull) {
                if (sbase == 0 && mis0) { // first vector: the bytes in front belong to another page
                    if (l >= mis0 && l < 16u) abase[l] = ws.stage[l];
                    __syncwarp();
                }
                flush_stage_vectors(stw, abase + sbase, nfull, sbase == 0 && mis0 != 0);
                __syncwarp();
                if (l == 0) { // the partial vector moves to the front
                    const uint4 v = *reinterpret_cast<const uint4*>(ws.stage + 16u * nfull);
                    *reinterpret_cast<uint4*>(ws.stage + 16u * nfull) = make_uint4(0, 0, 0, 0);
                    *reinterpret_cast<uint4*>(ws.stage) = v;
                }
                sbase += 16u * nfull;
                __syncwarp();
            }
        }
        const uint32_t end = mis0 + outp;
        const uint32_t from = sbase == 0 ? mis0 : sbase;
        if (from + l < end) abase[from + l] = ws.stage[from - sbase + l]; // < 16 bytes
        __syncwarp();
        if (l == 0) *reinterpret_cast<uint4*>(ws.stage) = make_uint4(0, 0, 0, 0);
        __syncwarp();
    }
    // the page's bytes against what the size pass / the page headers said; the last page closes the offsets array
    if (l == 0) {
        if (!P.exact_sizes && total != P.page_chars[q]) report_error(P.err, q, PQG_PAGE_LAYOUT);
        if (q + 1 == ck.first_page + ck.n_pages) offs[n] = page_base + total;
    }
    return true;
}

// Persistent CTAs; every warp takes the next batch of pages_per_cta consecutive pages from a device
// counter (the size pass counts in DevErr::slow_count, the copy pass in DevErr::slow_cursor; both
// are zeroed at the start of a run): all resident warps stay busy until the pages run out, whatever
// the spread of the per-page work (a fixed page per warp left 40 % of the warp slots idle).
// Measured on 40 M-row columns (scripts/gpu_variants.sh, profiles/README.md): the copy pass runs best with 2 CTAs per SM
// (128 registers: no spills; 136 KB of shared memory leaves the L1 its share for the dictionary gathers) and the next
// page staged ahead with cp.async -- cfg4 PLAIN 2.43 -> 2.17 ms, cfg3 dictionary 0.47 -> 0.46 ms against 3 CTAs without
// prefetch; 3 CTAs WITH the second buffer lose (2.81 / 0.63 ms: 204 KB of shared memory starve the L1).  The size pass
// stages nothing and keeps 3 CTAs.
#ifndef PQG_STR_PREFETCH
#define PQG_STR_PREFETCH 1
#endif
template <bool COPY, bool LEAN>
__global__ void __launch_bounds__(kThreadsPerCta, COPY ? 2 : 3) k_str_pages(DecodeParams P) {
    extern __shared__ __align__(16) uint8_t smem[];
    WarpScratch& ws = reinterpret_cast<WarpScratch*>(smem)[warp_id()];
    uint32_t* cursor = COPY ? &P.err->slow_cursor : &P.err->slow_count;
    const uint32_t n = P.page_end - P.page_begin, batch = P.pages_per_cta;
    if constexpr (COPY) {
        // launched without waiting for the size pass (the chars buffer of the previous run is reused): the grand total
        // must fit, otherwise nothing is written and pqg_plan_finish sizes the buffer and runs again
        if (P.total_chars && *P.total_chars > P.chars_cap) {
            if (blockIdx.x == 0 && threadIdx.x == 0) report_error(P.err, 0, PQG_PAGE_CHARS_CAP);
            return;
        }
    }
    // Copy pass: pages are staged one ahead -- while a page is decoded out of one shared buffer, the next page of the
    // batch streams into the other (cp.async), so its global latency (descriptor -> payload) hides behind the decode.
    uint8_t* alt = smem + sizeof(WarpScratch) * kWarpsPerCta + warp_id() * kSlotAlloc;
    bool stage_dirty = true;
    for (;;) {
        uint32_t i = 0;
        if (lane_id() == 0) i = atomicAdd(cursor, batch);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= n) break;
        const uint32_t i1 = min(n, i + batch);
        if constexpr (COPY && PQG_STR_PREFETCH) {
            uint32_t cur = 0;
            pqg_page_desc pd = P.pages[P.page_begin + i];
            page_stage_async(P, pd, ws.slot);
            page_stage_commit();
            for (; i < i1; i++) {
                const uint32_t q = P.page_begin + i;
                pqg_page_desc pn = pd;
                if (i + 1 < i1) {
                    pn = P.pages[q + 1];
                    page_stage_async(P, pn, cur ? ws.slot : alt);
                    page_stage_commit();
                    page_stage_wait<1>();
                } else page_stage_wait<0>();
                const DevChunk& ck = P.chunks[pd.chunk_idx];
                const uint8_t* buf = cur ? alt : ws.slot;
                const bool plain_req = !((pd.flags & PQG_PAGE_FLAG_DICT) && ck.has_dict) && ck.max_def <= 0 && ck.max_rep <= 0 && !P.validity;
                if (!(plain_req && lean_plain_page(P, q, pd, ck, ws, buf, stage_dirty)))
                    decode_str_page<COPY, LEAN>(P, q, pd, ck, ws, buf, stage_dirty);
                __syncwarp();
                pd = pn;
                cur ^= 1u;
            }
        } else {
            for (; i < i1; i++) {
                const uint32_t q = P.page_begin + i;
                const pqg_page_desc pd = P.pages[q];
                const DevChunk& ck = P.chunks[pd.chunk_idx];
                decode_str_page<COPY, LEAN>(P, q, pd, ck, ws, nullptr, stage_dirty);
                __syncwarp();
            }
        }
    }
}

// per chunk: exclusive scan of page_chars over the chunk's pages; chunk total -> char_base (temp)
__global__ void __launch_bounds__(1024) k_str_scan_pages(DecodeParams P) {
    __shared__ uint64_t wsum[32];
    __shared__ uint64_t carry_s;
    DevChunk& ck = P.chunks[blockIdx.x];
    const uint32_t l = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < ck.n_pages; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint64_t v = i < ck.n_pages ? P.page_chars[ck.first_page + i] : 0;
        uint64_t incl = v;
        for (int d = 1; d < 32; d <<= 1) { uint64_t t = __shfl_up_sync(0xffffffffu, incl, d); if (l >= (uint32_t)d) incl += t; }
        if (l == 31) wsum[w] = incl;
        __syncthreads();
        if (w == 0) {
            uint64_t x = wsum[l], xi = x;
            for (int d = 1; d < 32; d <<= 1) { uint64_t t = __shfl_up_sync(0xffffffffu, xi, d); if (l >= (uint32_t)d) xi += t; }
            wsum[l] = xi - x;
        }
        __syncthreads();
        uint64_t excl = carry_s + wsum[w] + incl - v;
        if (i < ck.n_pages) {
            if (excl + v > 0xffffffffull) { if (excl <= 0xffffffffull) report_error(P.err, ck.first_page + i, PQG_PAGE_CHARS_OVERFLOW); }
            P.page_char_base[ck.first_page + i] = static_cast<uint32_t>(excl);
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) ck.char_base = carry_s; // chunk total for now
}

// single block: chunk totals -> exclusive chunk bases; grand total -> *total
__global__ void k_str_scan_chunks(DecodeParams P, uint64_t* chunk_bases, uint64_t* total) {
    if (threadIdx.x == 0) {
        uint64_t acc = 0;
        for (uint32_t c = 0; c < P.n_chunks; c++) {
            uint64_t t = P.chunks[c].char_base;
            P.chunks[c].char_base = acc;
            chunk_bases[c] = acc;
            acc += t;
        }
        chunk_bases[P.n_chunks] = acc;
        *total = acc;
    }
}


} // namespace

size_t decode_smem_bytes(bool with_dict) {
    return sizeof(WarpScratch) * kWarpsPerCta + (with_dict ? kMaxSmemDictBytes : 0);
}

cudaError_t launch_dict_prepare(const DecodeParams& p, uint32_t n_chunks, int width, uint32_t max_dict_blocks, cudaStream_t s) {
    if (n_chunks == 0) return cudaSuccess;
    if (width == 0) { // strings: segments over many CTAs, one warp per dictionary to link them, entries over many CTAs
        dim3 grid(max_dict_blocks ? max_dict_blocks : 1u, n_chunks);
        k_dict_seg<<<grid, 256, 0, s>>>(p);
        k_dict_link<<<n_chunks, kLinkThreads, 0, s>>>(p);
        k_dict_emit<<<grid, 256, 0, s>>>(p);
        return cudaGetLastError();
    }
    dim3 grid(max_dict_blocks, n_chunks);
    switch (width) {
        case 1: k_dict_prepare<1><<<grid, 256, 0, s>>>(p); break;
        case 4: k_dict_prepare<4><<<grid, 256, 0, s>>>(p); break;
        case 8: k_dict_prepare<8><<<grid, 256, 0, s>>>(p); break;
        case 12: k_dict_prepare<12><<<grid, 256, 0, s>>>(p); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

template <int W, bool BOOLP>
static cudaError_t launch_fixed_t(DecodeParams p, int sm_count, uint32_t slow_hint, cudaStream_t s) {
    const size_t smem = decode_smem_bytes(false);
    cudaError_t e = cudaFuncSetAttribute(k_decode_fixed<W, BOOLP>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    // persistent CTAs over the work-stealing list; the list length is only known on the device
    // (the tile kernel appends to it): size the grid for the host-known part plus what the tile
    // kernel handed over in the previous run of the plan; nothing known yet -> every resident slot
    // (one CTA per SM left 3/4 of the warp slots of this latency-bound kernel empty)
    const uint32_t cap = static_cast<uint32_t>(sm_count) * 3u;
    uint32_t grid = cap;
    if (p.handover_hint != 0xffffffffu) {
        const uint32_t want = (slow_hint + p.handover_hint + kWarpsPerCta - 1) / kWarpsPerCta;
        grid = want < static_cast<uint32_t>(sm_count) ? static_cast<uint32_t>(sm_count) : (want > cap ? cap : want);
    }
    k_decode_fixed<W, BOOLP><<<grid, kThreadsPerCta, smem, s>>>(p);
    return cudaGetLastError();
}

// `boolean_plain`: BOOLEAN chunks (PLAIN pages are bit-packed; dictionary entries are bytes).
cudaError_t launch_decode_fixed(const DecodeParams& p, int width, bool boolean_plain, int sm_count, cudaStream_t s) {
    if (boolean_plain) return launch_fixed_t<1, true>(p, sm_count, p.slow_hi - p.slow_lo, s);
    switch (width) {
        case 4: return launch_fixed_t<4, false>(p, sm_count, p.slow_hi - p.slow_lo, s);
        case 8: return launch_fixed_t<8, false>(p, sm_count, p.slow_hi - p.slow_lo, s);
        case 12: return launch_fixed_t<12, false>(p, sm_count, p.slow_hi - p.slow_lo, s);
        default: return cudaErrorInvalidValue;
    }
}

template <bool COPY, bool LEAN>
static cudaError_t launch_str_t(DecodeParams p, int sm_count, cudaStream_t s) {
    const size_t smem = decode_smem_bytes(false) + (COPY && PQG_STR_PREFETCH ? static_cast<size_t>(kWarpsPerCta) * kSlotAlloc : 0); // + the second staging buffer per warp
    uint32_t n = p.page_end - p.page_begin;
    if (n == 0) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(k_str_pages<COPY, LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    int resident = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k_str_pages<COPY, LEAN>, kThreadsPerCta, smem);
    if (resident < 1) resident = 1;
    const uint32_t warps = static_cast<uint32_t>(sm_count) * static_cast<uint32_t>(resident) * kWarpsPerCta;
    // ~16 grabs per warp: few enough atomics on the one counter, fine enough for the tail
    p.pages_per_cta = std::min<uint32_t>(16u, std::max<uint32_t>(1u, n / (warps * 16u)));
    const uint32_t grid = std::min<uint32_t>((n + kWarpsPerCta - 1) / kWarpsPerCta, static_cast<uint32_t>(sm_count) * static_cast<uint32_t>(resident));
    k_str_pages<COPY, LEAN><<<grid, kThreadsPerCta, smem, s>>>(p);
    return cudaGetLastError();
}
cudaError_t launch_str_sizes(const DecodeParams& p, int sm_count, cudaStream_t s) { return launch_str_t<false, false>(p, sm_count, s); }
cudaError_t launch_str_copy(const DecodeParams& p, bool any_dict, int sm_count, cudaStream_t s) {
    return any_dict ? launch_str_t<true, true>(p, sm_count, s) : launch_str_t<true, false>(p, sm_count, s);
}

// total_chars: device pointer to [n_chunks + 1 chunk bases][grand total]
cudaError_t launch_str_scan(const DecodeParams& p, uint64_t* bases_and_total, cudaStream_t s) {
    if (p.n_chunks == 0) return cudaSuccess;
    k_str_scan_pages<<<p.n_chunks, 1024, 0, s>>>(p);
    k_str_scan_chunks<<<1, 32, 0, s>>>(p, bases_and_total, bases_and_total + p.n_chunks + 1);
    return cudaGetLastError();
}

} // namespace pqg
