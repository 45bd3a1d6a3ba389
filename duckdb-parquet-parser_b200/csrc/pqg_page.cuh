// pqg_page.cuh -- warp-level building blocks shared by every page kernel.
//
// One warp owns one data page.  The page payload (<= kSlotBytes) is staged into the warp's
// shared-memory slot as an exact byte image (16-byte aligned source rounded down, so the
// page starts `payload_off & 15` bytes into the slot); larger pages are read in place from
// global memory through the same code (generic pointers).
//
// Replaces, per page: read_data_page's level / index decoding
// (reference src/reader/column_reader.cpp:143-182) and RleDecoder::get_batch
// (reference include/reader/rle_decoder.hpp:17-95), restated as
//   run discovery  -> two regular layouts (the reference writer's) are recognised and expanded fully in parallel;
//                     any other stream is taken run by run with warp-uniform header parsing (every lane reads the
//                     same header bytes), so that
//   run expansion  -> is done by the whole warp on every run (lanes by validity word / by value),
//   slot emission  -> one lane per output slot, 32 slots per step, coalesced stores.
#pragma once
#include "pqg_internal.h"

namespace pqg {

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t warp_id() { return threadIdx.x >> 5; }

// ---- unaligned little-endian loads built from aligned 32-bit words ---------------------
// (valid for shared and global generic addresses; may touch up to 8 bytes past the field:
// the shared slot and the image are padded for that)
__device__ __forceinline__ uint32_t ld32u(const uint8_t* p) {
    uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* q = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
    return __funnelshift_r(q[0], q[1], static_cast<uint32_t>(a & 3) * 8);
}
__device__ __forceinline__ uint64_t ld64u(const uint8_t* p) {
    uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* q = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
    uint32_t sh = static_cast<uint32_t>(a & 3) * 8;
    uint32_t w0 = q[0], w1 = q[1], w2 = q[2];
    return (static_cast<uint64_t>(__funnelshift_r(w1, w2, sh)) << 32) | __funnelshift_r(w0, w1, sh);
}
// bw (0..32) bits at bit offset `bit` from base, LSB first (rle_decoder.hpp:55-65)
__device__ __forceinline__ uint32_t ldbits(const uint8_t* base, uint32_t bit, uint32_t bw) {
    uintptr_t a = reinterpret_cast<uintptr_t>(base + (bit >> 3));
    const uint32_t* q = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
    uint32_t sh = static_cast<uint32_t>(a & 3) * 8 + (bit & 7); // 0..31, sh + bw <= 63
    uint32_t v = __funnelshift_r(q[0], q[1], sh);
    return bw >= 32 ? v : (v & ((1u << bw) - 1u));
}
// same, but bytes at or past `avail` read as zero (slow; only for runs that overhang)
static __device__ __noinline__ uint32_t ldbits_bounded(const uint8_t* base, uint32_t bit, uint32_t bw, uint32_t avail) {
    uint64_t acc = 0;
    uint32_t b0 = bit >> 3;
    for (uint32_t i = 0; i < 5; i++) {
        uint32_t b = b0 + i;
        uint64_t byte = b < avail ? base[b] : 0;
        acc |= byte << (8 * i);
    }
    acc >>= (bit & 7);
    return bw >= 32 ? static_cast<uint32_t>(acc) : (static_cast<uint32_t>(acc) & ((1u << bw) - 1u));
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
    uint32_t l = lane_id();
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (l >= static_cast<uint32_t>(d)) v += t;
    }
    return v;
}

__device__ __forceinline__ uint4 ldg_nc16(const uint8_t* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// ---- errors ------------------------------------------------------------------------------
__device__ __forceinline__ void report_error(DevErr* e, uint32_t page, uint32_t code,
                                             uint32_t pos = 0, uint32_t need = 0, uint32_t size = 0) {
    // called by one lane
    atomicAdd(&e->count, 1u);
    unsigned long long key = (static_cast<unsigned long long>(page) << 32) | code;
    unsigned long long old = atomicMin(&e->key, key);
    if (key < old) { e->d_pos = pos; e->d_need = need; e->d_size = size; __threadfence(); e->d_page = page; }
}

// ---- per-warp scratch ----------------------------------------------------------------------
struct __align__(16) WarpScratch {
    uint8_t slot[kSlotAlloc];
    uint32_t idx[kIdxWords];
    uint32_t valid[32];
    uint32_t rankbase[32];
    uint8_t stage[kStageBytes32]; // BYTE_ARRAY copy pass: the chars of 32 short strings, flushed with aligned vectors
};

// ---- run discovery ---------------------------------------------------------------------------
struct Walker {
    const uint8_t* s;  // stream bytes
    uint32_t len;      // header parsing bound (RleDecoder::size_)
    uint32_t avail;    // bytes readable behind s (up to the page end)
    uint32_t pos, bw;
    uint32_t rem, lit, val, bit, next_pos;
    uint32_t bad;
};

__device__ __forceinline__ void walker_init(Walker& w, const uint8_t* s, uint32_t len, uint32_t avail, uint32_t bw) {
    w.s = s; w.len = len; w.avail = avail; w.pos = 0; w.bw = bw;
    w.rem = 0; w.lit = 0; w.val = 0; w.bit = 0; w.next_pos = 0; w.bad = 0;
}

// ---- run discovery, warp-uniform: EVERY lane parses the next run header (the same bytes: a broadcast load), so the warp
// can expand the run together -- no lane-0 walk, no 32-value pieces, no run table.  RleDecoder::next_counts
// (rle_decoder.hpp:37-53): returns false when the stream is exhausted (remaining outputs read as 0, :21-24); w.bad is set
// for a zero-length run (undefined behaviour in the reference).
__device__ __forceinline__ bool walker_next_run(Walker& w) {
    if (w.pos >= w.len) return false;
    const uint8_t* s = w.s;
    uint32_t ind = 0, shift = 0;
    while (w.pos < w.len) { // read_varint32 (rle_decoder.hpp:76-86)
        const uint32_t b = s[w.pos++];
        if (shift < 32) ind |= (b & 0x7Fu) << shift;
        if (!(b & 0x80u)) break;
        shift += 7;
    }
    if (ind & 1u) {
        w.rem = (ind >> 1) * 8u; w.lit = 1; w.bit = w.pos * 8u;
        w.next_pos = w.pos + static_cast<uint32_t>((static_cast<uint64_t>(w.rem) * w.bw + 7) >> 3);
    } else {
        w.rem = ind >> 1; w.lit = 0;
        const uint32_t nb = (w.bw + 7) >> 3;
        uint32_t v = 0;
        for (uint32_t i = 0; i < nb && w.pos < w.len; i++) { // value bytes are not masked (:88-95)
            if (i < 4) v |= static_cast<uint32_t>(s[w.pos]) << (8 * i);
            w.pos++;
        }
        w.val = v;
    }
    if (w.rem == 0) w.bad = PQG_PAGE_BAD_RUN;
    return true;
}
// OR `cnt` bits (all ones, or the stream bits from `srcbit` on when src != null) into words[] from bit `start` on; lanes by word
__device__ __forceinline__ void or_bits_warp(uint32_t* words, uint32_t start, uint32_t cnt, const uint8_t* src, uint32_t srcbit) {
    const uint32_t w0 = start >> 5, w1 = (start + cnt - 1u) >> 5;
    for (uint32_t w = w0 + lane_id(); w <= w1; w += 32) {
        const uint32_t lo = max(w * 32u, start), hi = min(w * 32u + 32u, start + cnt), c = hi - lo;
        uint32_t bits = c >= 32u ? 0xffffffffu : ((1u << c) - 1u);
        if (src) bits &= ldbits(src, srcbit + (lo - start), c);
        if (bits) atomicOr(&words[w], bits << (lo & 31u));
    }
}

__device__ __forceinline__ void set_bits_range(uint32_t* words, uint32_t start, uint32_t cnt) {
    uint32_t end = start + cnt; // cnt > 0
    uint32_t w0 = start >> 5, w1 = (end - 1) >> 5;
    for (uint32_t w = w0; w <= w1; w++) {
        uint32_t lo = (w == w0) ? (start & 31u) : 0u;
        uint32_t hi = (w == w1) ? (((end - 1) & 31u) + 1u) : 32u;
        uint32_t m = (hi == 32u ? 0xffffffffu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
        atomicOr(&words[w], m);
    }
}
__device__ __forceinline__ void set_bits_word(uint32_t* words, uint32_t start, uint32_t m) {
    if (!m) return;
    uint32_t sh = start & 31u;
    atomicOr(&words[start >> 5], m << sh);
    if (sh && (m >> (32u - sh))) atomicOr(&words[(start >> 5) + 1], m >> (32u - sh));
}

// level >= max_def means "value present" (reference tests def < max_def for null)
__device__ __forceinline__ bool level_present(uint32_t lv, int max_def) {
    return static_cast<int>(static_cast<int16_t>(static_cast<uint16_t>(lv))) >= max_def;
}

// Decode the validity of the next t (<= 1024) slots into ws.valid / ws.rankbase.
// Returns the number of present values in the tile; *bad receives walker errors.
__device__ __forceinline__ uint32_t levels_tile(Walker& w, WarpScratch& ws, uint32_t t, int max_def,
                                                bool first_single_tile, uint32_t* bad) {
    const uint32_t l = lane_id();
    if (max_def <= 0) {
        uint32_t lo = l * 32u;
        uint32_t m = (lo >= t) ? 0u : ((t - lo >= 32u) ? 0xffffffffu : ((1u << (t - lo)) - 1u));
        ws.valid[l] = m;
        ws.rankbase[l] = lo < t ? lo : t;
        __syncwarp();
        return t;
    }
    ws.valid[l] = 0;
    __syncwarp();
    bool done = false;
    // Regular layout: every run is <varint < 128><1 value byte>, i.e. RLE runs shorter than
    // 64 with a 1-byte level -- what the reference writer emits for scattered nulls
    // (src/writer/parquet_writer.cpp:103-135).  Verified in parallel (by induction over the
    // even bytes), then expanded with one lane per run and no sequential walk.
    if (first_single_tile && w.bw <= 8 && w.len >= 2 && !(w.len & 1u) && w.len <= w.avail) {
        const uint8_t* s = w.s;
        uint32_t nr = w.len >> 1;
        bool ok = true;
        {
            // no bit ranges, no loops over words: every run whose level class (present / null) differs from its
            // predecessor's sets ONE toggle bit at its first slot; the validity image is the prefix XOR of the toggles.
            // The layout is verified on the way (a header byte that is not <even, 2..126> voids the image: general walk).
            // two runs per lane and step (64 runs per warp prefix sum)
            uint32_t carry = 0, last_present = 0;
            for (uint32_t base = 0; base < nr && carry < t; base += 64) {
                const uint32_t r0 = base + 2u * l, r1 = r0 + 1u;
                uint32_t cnt0 = 0, cnt1 = 0, pres0 = 0, pres1 = 0;
                if (r0 < nr) {
                    const uint32_t b = s[2 * r0];
                    ok = ok && ((b & 0x81u) == 0u) && b != 0u;
                    cnt0 = b >> 1; pres0 = level_present(s[2 * r0 + 1], max_def) ? 1u : 0u;
                }
                if (r1 < nr) {
                    const uint32_t b = s[2 * r1];
                    ok = ok && ((b & 0x81u) == 0u) && b != 0u;
                    cnt1 = b >> 1; pres1 = level_present(s[2 * r1 + 1], max_def) ? 1u : 0u;
                }
                const uint32_t both = cnt0 + cnt1;
                const uint32_t incl = warp_incl_scan(both);
                const uint32_t start0 = carry + incl - both, start1 = start0 + cnt0;
                // class of the run in front of this lane's first run: the second run of the lane below (its last one in range)
                uint32_t prev = __shfl_up_sync(0xffffffffu, pres1, 1);
                if (l == 0) prev = last_present;
                if (r0 < nr && start0 < t && pres0 != prev) atomicXor(&ws.valid[start0 >> 5], 1u << (start0 & 31u));
                if (r1 < nr && start1 < t && pres1 != pres0) atomicXor(&ws.valid[start1 >> 5], 1u << (start1 & 31u));
                const uint32_t last_run = min(63u, nr - 1u - base);            // last run of this step, 0..63
                const uint32_t lp0 = __shfl_sync(0xffffffffu, pres0, last_run >> 1), lp1 = __shfl_sync(0xffffffffu, pres1, last_run >> 1);
                last_present = (last_run & 1u) ? lp1 : lp0;
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
            __syncwarp();
            // a short stream leaves the remaining slots null (rle_decoder.hpp:21-24)
            if (l == 0 && last_present && carry < t) ws.valid[carry >> 5] ^= 1u << (carry & 31u);
            __syncwarp();
            uint32_t x = ws.valid[l];
            const uint32_t par = __popc(x) & 1u;
            x ^= x << 1; x ^= x << 2; x ^= x << 4; x ^= x << 8; x ^= x << 16;
            uint32_t px = par; // inclusive xor-scan of the word parities
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, px, d); if (l >= static_cast<uint32_t>(d)) px ^= o; }
            if ((px ^ par) & 1u) x = ~x; // an odd number of toggles in front of this word
            const uint32_t lo = l * 32u;
            x &= lo >= t ? 0u : (t - lo >= 32u ? 0xffffffffu : ((1u << (t - lo)) - 1u));
            __syncwarp();
            if (__all_sync(0xffffffffu, ok)) {
                ws.valid[l] = x;
                w.pos = w.len; // consumed
                done = true;
            } else {
                ws.valid[l] = 0; // not the writer's layout after all
                __syncwarp();
            }
        }
    }
    // any other stream: run by run, the whole warp on each run (the Walker state is the same in every lane)
    {
        const uint32_t bw = w.bw, avail = w.avail;
        const uint8_t* s = w.s;
        uint32_t produced = 0;
        while (!done && produced < t) {
            if (w.rem == 0) {
                if (!walker_next_run(w)) break;            // exhausted: the remaining slots read level 0 = null (max_def > 0 here)
                if (w.bad) { *bad = w.bad; return 0; }
            }
            const uint32_t take = min(w.rem, t - produced);
            if (!w.lit) {
                if (level_present(w.val, max_def)) or_bits_warp(ws.valid, produced, take, nullptr, 0);
            } else {
                const bool inb = ((static_cast<uint64_t>(w.bit) + static_cast<uint64_t>(take) * bw + 7) >> 3) <= avail;
                if (inb && bw == 1 && max_def == 1) {
                    or_bits_warp(ws.valid, produced, take, s, w.bit); // the packed bits ARE the validity bits
                } else {
                    for (uint32_t j0 = 0; j0 < take; j0 += 32) {
                        const uint32_t j = j0 + l;
                        bool pres = false;
                        if (j < take) {
                            const uint32_t lv = inb ? ldbits(s, w.bit + j * bw, bw) : ldbits_bounded(s, w.bit + j * bw, bw, avail);
                            pres = level_present(lv, max_def);
                        }
                        const uint32_t m = __ballot_sync(0xffffffffu, pres);
                        if (l == 0) set_bits_word(ws.valid, produced + j0, m);
                    }
                }
                w.bit += take * bw;
            }
            w.rem -= take;
            if (w.lit && w.rem == 0) w.pos = w.next_pos;
            produced += take;
        }
    }
    __syncwarp();
    uint32_t c = __popc(ws.valid[l]);
    uint32_t incl = warp_incl_scan(c);
    ws.rankbase[l] = incl - c;
    __syncwarp();
    return __shfl_sync(0xffffffffu, incl, 31);
}

__device__ __forceinline__ void idx_store(uint32_t* buf, uint32_t k, uint32_t v, bool wide) {
    if (wide) buf[k] = v;
    else reinterpret_cast<uint16_t*>(buf)[k] = static_cast<uint16_t>(v);
}
__device__ __forceinline__ uint32_t idx_load(const uint32_t* buf, uint32_t k, bool wide) {
    return wide ? buf[k] : static_cast<uint32_t>(reinterpret_cast<const uint16_t*>(buf)[k]);
}

// Decode the next `cnt` dictionary indices of the stream into ws.idx[0..cnt).
// `wide` = indices may need more than 16 bits (bw > 16): cnt <= 512, else cnt <= 1024.
__device__ __forceinline__ void indices_tile(Walker& w, WarpScratch& ws, uint32_t cnt_total, bool wide, uint32_t* bad) {
    const uint32_t l = lane_id();
    const uint32_t bw = w.bw, avail = w.avail;
    const uint8_t* s = w.s;
    uint32_t produced = 0;
    while (produced < cnt_total) { // run by run, the whole warp on each run (uniform Walker state)
        if (w.rem == 0) {
            if (!walker_next_run(w)) { // exhausted: the remaining indices read as 0 (rle_decoder.hpp:21-24)
                for (uint32_t k = produced + l; k < cnt_total; k += 32) idx_store(ws.idx, k, 0u, wide);
                break;
            }
            if (w.bad) { *bad = w.bad; return; }
        }
        const uint32_t take = min(w.rem, cnt_total - produced);
        if (w.lit) {
            const bool inb = ((static_cast<uint64_t>(w.bit) + static_cast<uint64_t>(take) * bw + 7) >> 3) <= avail;
            for (uint32_t j = l; j < take; j += 32)
                idx_store(ws.idx, produced + j, inb ? ldbits(s, w.bit + j * bw, bw) : ldbits_bounded(s, w.bit + j * bw, bw, avail), wide);
            w.bit += take * bw;
        } else {
            for (uint32_t j = l; j < take; j += 32) idx_store(ws.idx, produced + j, w.val, wide);
        }
        w.rem -= take;
        if (w.lit && w.rem == 0) w.pos = w.next_pos;
        produced += take;
    }
    __syncwarp();
}

// All runs are single bit-packed groups "03 <bw bytes>" (what RleBpEncoder emits for data
// without 4-fold repeats, include/writer/rle_bp_encoder.hpp:93-98): group g starts at byte
// g * (1 + bw).  Verified in parallel; then every lane can address its own value directly.
__device__ __forceinline__ bool check_regular(const uint8_t* s, uint32_t len, uint32_t bw, uint32_t count) {
    if (count == 0) return true;
    uint32_t groups = (count + 7) >> 3;
    if (static_cast<uint64_t>(groups) * (1u + bw) > len) return false;
    bool ok = true;
    for (uint32_t g = lane_id(); g < groups; g += 32) ok = ok && (s[g * (1u + bw)] == 0x03u);
    return __all_sync(0xffffffffu, ok);
}
__device__ __forceinline__ uint32_t regular_index(const uint8_t* s, uint32_t bw, uint32_t k) {
    uint32_t g = k >> 3;
    return ldbits(s, ((g * (1u + bw) + 1u) << 3) + (k & 7u) * bw, bw);
}

// The writer's index stream in full generality of RleBpEncoder for data without 4-fold
// repeats (include/writer/rle_bp_encoder.hpp:49-61,93-106): floor(count / 8) single bit-packed
// groups "03 <bw bytes>", then for the last count % 8 values EITHER one more zero-padded group OR
// -- when the trailing values are all equal -- one RLE run "<r << 1> <ceil(bw/8) value bytes>"
// (FinishWrite flushes a pending run of 1..3 equal values as RLE).
struct RegStream {
    const uint8_t* s;
    uint32_t bw;
    uint32_t tail_start; // first value index served by the trailing RLE run (count: none)
    uint32_t tail_val;
};
__device__ __forceinline__ bool check_regular2(const uint8_t* s, uint32_t len, uint32_t bw, uint32_t count, RegStream* out) {
    out->s = s; out->bw = bw; out->tail_start = count; out->tail_val = 0;
    if (count == 0) return true;
    const uint32_t G = count >> 3, r = count & 7u, gs = 1u + bw;
    if (static_cast<uint64_t>(G) * gs > len) return false;
    bool ok = true;
    for (uint32_t g = lane_id(); g < G; g += 32) ok = ok && (s[g * gs] == 0x03u);
    if (r) {
        const uint32_t p = G * gs;
        if (p >= len) ok = false;
        else {
            const uint32_t h = s[p], nb = (bw + 7u) >> 3;
            if (h == 0x03u && p + gs <= len) { /* padded group */ }
            else if (h == (r << 1) && p + 1u + nb <= len) {
                uint32_t v = 0;
                for (uint32_t i = 0; i < nb && i < 4u; i++) v |= static_cast<uint32_t>(s[p + 1u + i]) << (8u * i);
                out->tail_start = G * 8u;
                out->tail_val = v; // not masked, like the reference (rle_decoder.hpp:88-95)
            } else ok = false;
        }
    }
    return __all_sync(0xffffffffu, ok);
}
__device__ __forceinline__ uint32_t regular_index2(const RegStream& rs, uint32_t k) {
    return k >= rs.tail_start ? rs.tail_val : regular_index(rs.s, rs.bw, k);
}

// ---- BYTE_ARRAY PLAIN: length-prefix walk -----------------------------------------------------
// Lane 0 walks `cnt` length-prefixed strings starting at byte `pos` of the value section and
// records each prefix position in ws.idx (u16 when !wide).  Returns the end position in *pos.
// TODO(perf): candidate-and-verify parallel walk for ASCII pages.
__device__ __forceinline__ bool walk_strings(const uint8_t* vals, uint32_t vavail, uint32_t* pos_io, uint32_t cnt,
                                             WarpScratch& ws, bool wide, bool record, uint32_t* epos, uint32_t* eneed) {
    uint32_t ok = 1, pos = *pos_io;
    if (lane_id() == 0) {
        for (uint32_t k = 0; k < cnt; k++) {
            if (static_cast<uint64_t>(pos) + 4 > vavail) { ok = 0; *epos = pos; *eneed = 4; break; }
            uint32_t len = ld32u(vals + pos);
            if (static_cast<uint64_t>(pos) + 4 + len > vavail) { ok = 0; *epos = pos + 4; *eneed = len; break; }
            if (record) idx_store(ws.idx, k, pos, wide);
            pos += 4 + len;
        }
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    *pos_io = __shfl_sync(0xffffffffu, pos, 0);
    *epos = __shfl_sync(0xffffffffu, *epos, 0);
    *eneed = __shfl_sync(0xffffffffu, *eneed, 0);
    __syncwarp();
    return ok != 0;
}

// ---- BYTE_ARRAY PLAIN: parallel length-prefix discovery --------------------------------------
// A length prefix of a string shorter than 64 KiB is <lo, hi, 0, 0>: candidates are the byte
// positions whose bytes +2 and +3 are zero and whose string stays inside the value section.
// Every lane scans its 1/32 of the section (zero-byte bit tricks, 4 positions per step), the
// candidates are compacted in order (warp prefix sum of the per-lane counts) into `out` (u16
// positions), and the chain is verified in parallel: cand[0] == 0, cand[i] + 4 + len ==
// cand[i+1], the nn-th string ends inside the section.  On text pages the candidates are exactly
// the prefixes; anything else (strings of 0-3 bytes producing false candidates, NUL-heavy
// binary data, strings >= 64 KiB) fails the check and the caller falls back to the sequential
// walk.  Returns the end position of the nn-th string in *end_pos.
// Word loaders: the scan below reads ALIGNED 32-bit words of the buffer behind the value
// section; shared memory gets explicit ld.shared (no generic-address overhead).
struct SmemWords {
    uint32_t base; // shared-state-space address of the aligned word 0
    __device__ __forceinline__ uint32_t operator()(uint32_t wi) const {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + 4u * wi));
        return v;
    }
    __device__ __forceinline__ uint32_t u16at(uint32_t apos) const { // unaligned u16 at byte apos
        uint32_t b0, b1;
        asm volatile("ld.shared.u8 %0, [%2];\n\tld.shared.u8 %1, [%2+1];" : "=r"(b0), "=r"(b1) : "r"(base + apos));
        return b0 | (b1 << 8);
    }
};
struct GenericWords {
    const uint32_t* base;
    __device__ __forceinline__ uint32_t operator()(uint32_t wi) const { return base[wi]; }
    __device__ __forceinline__ uint32_t u16at(uint32_t apos) const {
        const uint8_t* b = reinterpret_cast<const uint8_t*>(base) + apos;
        return static_cast<uint32_t>(b[0]) | (static_cast<uint32_t>(b[1]) << 8);
    }
};
__device__ __forceinline__ uint32_t zero_byte_mask32(uint32_t w) { // bit 8i+7 set iff byte i == 0 (exact)
    return ~(((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w | 0x7f7f7f7fu);
}
// u16 at byte position apos of the word stream
template <class LD>
__device__ __forceinline__ uint32_t ld16_at(const LD& ld, uint32_t apos) { return ld.u16at(apos); }

// `ld` addresses aligned words; the value section starts `o` (0..3) bytes into word 0 ("A-space"
// = byte positions counted from word 0).  Lane l owns the R <= 64 A-space positions from l * R.
//  1. candidate flags, 4 positions per step: zero-byte masks of two neighbouring words, funnel
//     shifted so that bit 8k+7 says "bytes k+2 and k+3 are zero", accumulated with one shift + one
//     LOP3 per step into two 32-bit words; flag of position 4i+k (i = step, k = byte) sits at bit
//     8k+i of its word.
//  2. a flag whose successor position is flagged too is dropped: the byte before a prefix
//     <len, 0, 0, 0> of a string shorter than 256 bytes is always such a shadow.  (Dropping is a
//     heuristic only: whatever survives is verified as an exact chain below.)
//  3. the 1-4 survivors per lane are range-checked against their own length and kept sorted in
//     registers; warp prefix sum; ordered write; parallel chain verification.
// *lane_pos / *lane_len (optional) return cand[lane] and its length for lane < nn.
template <class LD>
__device__ __forceinline__ bool find_headers_w(const LD& ld, uint32_t o, uint32_t vavail, uint32_t nn, uint16_t* out, uint32_t cap,
                                               uint32_t* end_pos, uint32_t* lane_pos = nullptr, uint32_t* lane_len = nullptr) {
    const uint32_t l = lane_id();
    if (nn == 0) { *end_pos = 0; return true; }
    if (vavail < 4u || vavail > 65535u || nn > cap) return false;
    const uint32_t last = vavail - 4u;                      // last vals-relative position of a prefix
    const uint32_t R = ((o + last) / 32u + 4u) & ~3u;       // A-space positions per lane, multiple of 4
    if (R > 64u) return false;
    const uint32_t a0 = l * R;                              // first A-space position of this lane
    uint32_t lo = 0, hi = 0;
    const uint32_t steps = R >> 2;                          // 1..16, warp-uniform
    if (a0 <= o + last) {
        const uint32_t w0i = a0 >> 2;
        uint32_t z0 = zero_byte_mask32(ld(w0i));
#pragma unroll
        for (uint32_t i = 0; i < 16; i++) {
            if (i >= steps) break;
            const uint32_t z1 = zero_byte_mask32(ld(w0i + i + 1));
            const uint32_t zz = __funnelshift_r(z0, z1, 16) & __funnelshift_r(z0, z1, 24);
            if (i < 8) lo = (lo >> 1) | zz; else hi = (hi >> 1) | zz;
            z0 = z1;
        }
        lo >>= 8u - min(steps, 8u);
        if (steps > 8u) hi >>= 16u - steps;
        // drop flags whose successor position is flagged (bit + 8 inside a step, bit - 23 across steps)
        const uint32_t slo = (lo >> 8) | ((lo & 0xfeu) << 23) | (hi << 31);
        const uint32_t shi = (hi >> 8) | ((hi & 0xfeu) << 23);
        lo &= ~slo;
        hi &= ~shi;
    }
    uint32_t c0 = 0xffffffffu, c1 = 0xffffffffu, c2 = 0xffffffffu, c3 = 0xffffffffu;
    uint32_t cnt = 0;
    bool over = false;
#pragma unroll
    for (uint32_t half = 0; half < 2; half++) {
        uint32_t m = half ? hi : lo;
        while (m) {
            const uint32_t b = __ffs(static_cast<int>(m)) - 1;
            m &= m - 1;
            const uint32_t apos = a0 + half * 32u + (((b & 7u) << 2) | (b >> 3));
            const uint32_t pos = apos - o;                  // wraps (and fails `<= last`) in front of the section
            const uint32_t len = ld.u16at(apos);
            if (pos <= last && pos + 4u + len <= vavail) {
                uint32_t v = pos, t;                        // sorted insert; what falls off the end is an overflow
                t = min(c0, v); v = max(c0, v); c0 = t;
                t = min(c1, v); v = max(c1, v); c1 = t;
                t = min(c2, v); v = max(c2, v); c2 = t;
                t = min(c3, v); v = max(c3, v); c3 = t;
                over = over || v != 0xffffffffu;
                cnt++;
            }
        }
    }
    if (__any_sync(0xffffffffu, over)) return false;
    const uint32_t incl = warp_incl_scan(cnt);
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if (total < nn || total > cap) return false;
    const uint32_t w = incl - cnt;
    if (cnt > 0) out[w] = static_cast<uint16_t>(c0);
    if (cnt > 1) out[w + 1] = static_cast<uint16_t>(c1);
    if (cnt > 2) out[w + 2] = static_cast<uint16_t>(c2);
    if (cnt > 3) out[w + 3] = static_cast<uint16_t>(c3);
    __syncwarp();
    bool ok = true;
    uint32_t endp = 0;
    for (uint32_t i = l; i < nn; i += 32) {
        const uint32_t ci = out[i];
        const uint32_t len = ld.u16at(ci + o);
        const uint32_t nx = ci + 4u + len;
        if (i == l) { if (lane_pos) *lane_pos = ci; if (lane_len) *lane_len = len; }
        if (i == 0 && ci != 0) ok = false;
        if (i + 1 < nn) { if (out[i + 1] != nx) ok = false; }
        else endp = nx;
    }
    ok = __all_sync(0xffffffffu, ok);
    endp = __reduce_max_sync(0xffffffffu, endp);
    *end_pos = endp;
    __syncwarp();
    return ok;
}

// generic-pointer front end (shared slot or global memory)
__device__ __forceinline__ bool find_headers(const uint8_t* vals, uint32_t vavail, uint32_t nn, uint16_t* out, uint32_t cap,
                                             uint32_t* end_pos) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(vals);
    GenericWords ld{reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3))};
    return find_headers_w(ld, static_cast<uint32_t>(a & 3u), vavail, nn, out, cap, end_pos);
}

// ---- page prolog -----------------------------------------------------------------------------
struct PageCtx {
    const uint8_t* pg;   // payload bytes (shared slot or global)
    uint32_t size;
    uint32_t n;          // num_values (level entries)
    uint32_t vals_pos;   // byte position of the value section inside the payload
    uint32_t bw;         // dictionary index bit width
    bool dict;           // indices into the chunk dictionary
    bool has_def;
    bool wide;
    Walker defw, idxw;
};

// Stage the payload and parse [def levels][rep levels][bit width] like read_data_page
// (column_reader.cpp:143-182).  Returns false (after reporting) when the page is unusable.
// asynchronous staging of a page into a slot-sized shared buffer (cp.async, 16 bytes per copy; the page then starts
// `payload_off & 15` bytes into the buffer, exactly as page_begin lays it out).  Pages beyond kSlotBytes are not staged.
__device__ __forceinline__ void page_stage_async(const DecodeParams& P, const pqg_page_desc& pd, uint8_t* buf) {
    if (pd.payload_size > static_cast<uint32_t>(kSlotBytes) || pd.num_values == 0) return;
    const uint32_t shift = static_cast<uint32_t>(pd.payload_off & 15u);
    const uint8_t* a = P.image + pd.payload_off - shift;
    const uint32_t nvec = (shift + pd.payload_size + 15u) >> 4;
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(buf));
    for (uint32_t j = lane_id(); j < nvec; j += 32)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 16u * j), "l"(a + 16u * j) : "memory");
}
__device__ __forceinline__ void page_stage_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING> __device__ __forceinline__ void page_stage_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory");
    __syncwarp();
}

// `prestaged`: the page already sits in that buffer (page_stage_async + wait); null: staged here into ws.slot
__device__ __forceinline__ bool page_begin(const DecodeParams& P, uint32_t q, const pqg_page_desc& pd,
                                           const DevChunk& ck, WarpScratch& ws, PageCtx& c, const uint8_t* prestaged = nullptr) {
    const uint32_t l = lane_id();
    c.size = pd.payload_size;
    c.n = pd.num_values;
    c.dict = (pd.flags & PQG_PAGE_FLAG_DICT) && ck.has_dict;
    c.has_def = ck.max_def > 0;
    const uint8_t* src = P.image + pd.payload_off;
    if (prestaged && c.size <= static_cast<uint32_t>(kSlotBytes)) {
        c.pg = prestaged + static_cast<uint32_t>(pd.payload_off & 15u);
    } else if (c.size <= static_cast<uint32_t>(kSlotBytes)) {
        uint32_t shift = static_cast<uint32_t>(pd.payload_off & 15u);
        const uint8_t* a = src - shift;
        uint32_t nvec = (shift + c.size + 15u) >> 4;
        uint4* dst = reinterpret_cast<uint4*>(ws.slot);
        for (uint32_t j = l; j < nvec; j += 32) dst[j] = ldg_nc16(a + 16u * j);
        __syncwarp();
        c.pg = ws.slot + shift;
    } else {
        // read in place: bring the first 64 KB of the page towards the SM first (every lane a line),
        // otherwise each dependent bit-field load of the decode below waits for DRAM on its own
        const uint32_t pre = min(c.size, 65536u);
        for (uint32_t off = l * 128u; off < pre; off += 32u * 128u) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + off));
        c.pg = src;
    }
    uint32_t pos = 0;
    if (c.has_def) {
        if (c.size < 4) { if (l == 0) report_error(P.err, q, PQG_PAGE_TRUNCATED, 0, 4, c.size); return false; }
        uint32_t def_len = ld32u(c.pg);
        if (def_len > c.size - 4) { if (l == 0) report_error(P.err, q, PQG_PAGE_TRUNCATED, 4, def_len, c.size); return false; }
        walker_init(c.defw, c.pg + 4, def_len, c.size - 4, ck.def_bw);
        pos = 4 + def_len;
    }
    if (ck.max_rep > 0) { // decoded by the reference, then unused (:157-164): skipped here
        if (c.size - pos < 4) { if (l == 0) report_error(P.err, q, PQG_PAGE_TRUNCATED, pos, 4, c.size); return false; }
        uint32_t rep_len = ld32u(c.pg + pos);
        pos += 4;
        if (rep_len > c.size - pos) { if (l == 0) report_error(P.err, q, PQG_PAGE_TRUNCATED, pos, rep_len, c.size); return false; }
        pos += rep_len;
    }
    c.bw = 0;
    if (c.dict) {
        if (pos >= c.size) { if (l == 0) report_error(P.err, q, PQG_PAGE_TRUNCATED, pos, 1, c.size); return false; }
        c.bw = c.pg[pos];
        pos++;
        if (c.bw > 32) { if (l == 0) report_error(P.err, q, PQG_PAGE_BAD_BIT_WIDTH, pos, c.bw, c.size); return false; }
        walker_init(c.idxw, c.pg + pos, c.size - pos, c.size - pos, c.bw);
    }
    c.vals_pos = pos;
    c.wide = c.dict && c.bw > 16;
    return true;
}

} // namespace pqg
