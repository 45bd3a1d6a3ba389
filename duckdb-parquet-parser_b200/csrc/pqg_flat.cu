// pqg_flat.cu -- host-listed pages of 4/8-byte plans (oversized pages of foreign writers, OPTIONAL pages beyond the
// tile shapes), decoded as a FLAT list of 1024-slot blocks instead of one warp / one CTA per page.
//
// A page of 20 K .. 120 K slots with scattered nulls is a few hundred to a few thousand runs: whoever owns the whole page
// (a warp in the general kernel, a CTA in the retired big-page kernel) walks them one step at a time, every step waiting
// on the loads of the step before -- 0.4 .. 0.9 TB/s on pyarrow files.  Here the serial part is cut down to what IS
// serial and everything else is spread over every warp of the device (reference semantics: read_data_page,
// src/reader/column_reader.cpp:140-225; RleDecoder, include/reader/rle_decoder.hpp:17-95):
//
// The work list = the host-listed pages of the launch + the pages the tile kernel handed over (OPTIONAL pages of more than
// 1024 slots, irregular index streams); all three kernels are persistent and take pages / blocks from device cursors.
//
//   k_flat_scan   one warp per page: the run-boundary walk.  Definition-level runs are expanded on the fly into the
//                 column's validity bitmap (a warp per run: word-wise OR of ones or of the literal bits); dictionary-index
//                 runs are only VISITED -- every 1024th value gets a checkpoint {stream position of its run header,
//                 first value of that run}.  Everything irregular (nested levels, bad runs, overhanging literal data,
//                 bit width > 32) hands the page to the general kernel, which also does all error reporting.
//   k_flat_ranks  one warp per page: popcount of the page's validity bits per 1024-slot block -> the rank (index of the
//                 first non-null value) of every block; PLAIN pages are checked against their payload size.
//   k_flat_emit   one warp per BLOCK, every block of every page in one grid: validity words -> ranks -> values.  PLAIN
//                 blocks read their contiguous value range, dictionary blocks restart the index stream at their
//                 checkpoint, extract their (<= 1024) indices into shared memory and gather.  Four 32-slot steps are in
//                 flight per warp.  Out-of-range indices become nulls here (column_reader.cpp:190-194).
#include "pqg_page.cuh"

namespace pqg {
namespace {

constexpr uint32_t kFlatUnroll = 4; // 32-slot steps per trip of the emission loops (loads in flight per lane)

template <int W> struct FlatElem;
template <> struct FlatElem<4> { using T = uint32_t; };
template <> struct FlatElem<8> { using T = uint64_t; };
template <int W> __device__ __forceinline__ typename FlatElem<W>::T flat_ld(const uint8_t* p);
template <> __device__ __forceinline__ uint32_t flat_ld<4>(const uint8_t* p) { return ld32u(p); }
template <> __device__ __forceinline__ uint64_t flat_ld<8>(const uint8_t* p) { return ld64u(p); }

// shared-state-space accesses by address (the generic forms cost an address conversion per access)
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
// the block's index buffer: 32-bit entries, or 16-bit ones when every dictionary of the plan has <= 65535 entries (an index
// beyond 16 bits is out of range there whatever its value: it is stored as 0xffff and stays out of range)
template <bool IDX16> __device__ __forceinline__ void sts_idx(uint32_t base, uint32_t i, uint32_t v) {
    if constexpr (IDX16) asm volatile("st.shared.u16 [%0], %1;" ::"r"(base + 2u * i), "h"(static_cast<uint16_t>(min(v, 0xffffu))) : "memory");
    else sts32(base + 4u * i, v);
}
template <bool IDX16> __device__ __forceinline__ uint32_t lds_idx(uint32_t base, uint32_t i) {
    if constexpr (IDX16) { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(base + 2u * i)); return v; }
    else return lds32(base + 4u * i);
}
template <int W> __device__ __forceinline__ typename FlatElem<W>::T lds_elem(uint32_t a);
template <> __device__ __forceinline__ uint32_t lds_elem<4>(uint32_t a) { return lds32(a); }
template <> __device__ __forceinline__ uint64_t lds_elem<8>(uint32_t a) { uint64_t v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a)); return v; }

__device__ __forceinline__ void prefetch_l1(const uint8_t* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
// lanes prefetch the 128-byte lines of [p + from, p + to)
__device__ __forceinline__ void prefetch_range(const uint8_t* p, uint32_t from, uint32_t to) {
    for (uint32_t o = (from & ~127u) + 128u * lane_id(); o < to; o += 32u * 128u) prefetch_l1(p + o);
}

__device__ __forceinline__ void flat_hand_over(const DecodeParams& P, FlatPage& fp, uint32_t q) {
    // called by one lane
    fp.status = 1;
    const uint32_t k = atomicAdd(&P.flat->handed, 1u);
    PQG_ASSERT(k < P.slow_cap);
    P.flat_append[k] = q;
}
// entry e of the work list: host-listed pages first, then what the tile kernel appended
__device__ __forceinline__ uint32_t flat_list_len(const DecodeParams& P) { return (P.slow_hi - P.slow_lo) + P.err->slow_count; }
__device__ __forceinline__ uint32_t flat_list_page(const DecodeParams& P, uint32_t e) {
    const uint32_t n_host = P.slow_hi - P.slow_lo;
    return e < n_host ? P.slow_pages[P.slow_lo + e] : P.slow_append[e - n_host];
}
__device__ __forceinline__ uint32_t warp_take(uint32_t* cursor) {
    uint32_t e = 0;
    if (lane_id() == 0) e = atomicAdd(cursor, 1u);
    return __shfl_sync(0xffffffffu, e, 0);
}

// ---- run-boundary scan of a 32-byte window of an RLE / bit-packed hybrid stream, by the warp ---------------------------
// RleDecoder::next_counts (rle_decoder.hpp:37-53) reads header after header; with scattered nulls a level stream is a run
// every two or three bytes and that walk is the whole cost of a page.  Here lane i takes stream byte pos + i for a run
// header and works out where that run would end; the true headers are the chain from lane 0, marked by pointer doubling
// inside the warp (5 rounds of shuffle + or-reduce cover the 32 positions); the lanes on the chain then hold one run each.
// Headers the window cannot settle (varints of more than two bytes, zero-length runs, data or value bytes that overhang the
// stream) end the chain: the caller takes that one run with the sequential walker, which owns the semantics of those cases.
// (A CTA per page with 256-byte windows -- per-warp exit tables chained through shared memory -- was built and measured:
//  1.75x the instructions per stream byte and one eighth of the pages in flight; slower on every file of bench_foreign.py.)
struct RunAt { uint32_t cnt, lit, data, end; }; // cnt 0: not settled here; data: literal -- bit offset of the run's data, RLE -- the value; end: stream position behind the run
__device__ __forceinline__ RunAt parse_run_at(const uint8_t* s, uint32_t len, uint32_t a, uint32_t bw) {
    RunAt r{0, 0, 0, 0};
    if (a >= len) return r;
    const uint32_t nb = (bw + 7u) >> 3;
    const uint32_t b0 = s[a], b1 = a + 1u < len ? s[a + 1u] : 0x80u;
    uint32_t ind = b0, hl = 1;
    if (b0 & 0x80u) {
        if (b1 & 0x80u) return r;
        ind = (b0 & 0x7fu) | (b1 << 7); hl = 2;
    }
    if (ind & 1u) {
        const uint32_t groups = ind >> 1, db = groups * bw;
        if (groups && a + hl + db <= len) { r.cnt = groups * 8u; r.lit = 1; r.data = (a + hl) * 8u; r.end = a + hl + db; }
    } else {
        const uint32_t cnt = ind >> 1;
        if (cnt && a + hl + nb <= len) {
            uint32_t v = 0;
            for (uint32_t k = 0; k < nb && k < 4u; k++) v |= static_cast<uint32_t>(s[a + hl + k]) << (8u * k); // not masked (rle_decoder.hpp:88-95)
            r.cnt = cnt; r.data = v; r.end = a + hl + nb;
        }
    }
    return r;
}
// returns the stream bytes the chain covers; r.cnt = 0 for lanes that are not run headers on the chain;
// *cut = the chain ended at a header for the sequential walker (at pos + return)
__device__ __forceinline__ uint32_t window_runs(const uint8_t* s, uint32_t len, uint32_t pos, uint32_t bw, RunAt& r, bool* cut) {
    const uint32_t l = lane_id();
    r = parse_run_at(s, len, pos + l, bw);
    const uint32_t adv = r.cnt ? r.end - pos : 0u;          // window-relative end of the run
    const uint32_t next = r.cnt ? min(adv, 32u) : 33u;      // 0..31: the next header inside the window, 32: beyond it, 33: not settled here
    uint32_t mask = 1u, j = next;
#pragma unroll
    for (int round = 0; round < 5; round++) { // round k marks chain distances 2^k .. 2^(k+1) - 1, then squares the jumps
        const uint32_t contrib = (((mask >> l) & 1u) && j < 32u) ? (1u << j) : 0u;
        mask |= __reduce_or_sync(0xffffffffu, contrib);
        const uint32_t jj = __shfl_sync(0xffffffffu, j, j & 31u);
        if (j < 32u) j = jj;
    }
    const uint32_t unsettled = __ballot_sync(0xffffffffu, ((mask >> l) & 1u) && next == 33u);
    uint32_t used;
    if (unsettled) {
        const uint32_t h = static_cast<uint32_t>(__ffs(static_cast<int>(unsettled))) - 1u; // the chain is one path: nothing behind h is on it
        mask &= (1u << h) - 1u;
        used = h;
        *cut = true;
    } else {
        used = __shfl_sync(0xffffffffu, adv, 31 - __clz(static_cast<int>(mask)));
        *cut = false;
    }
    if (!((mask >> l) & 1u)) r.cnt = 0;
    return used;
}

// clear bits [start, start + cnt) of words[] (the page's own slots only: neighbours share the edge words)
__device__ __forceinline__ void clear_bits_warp(uint32_t* words, uint32_t start, uint32_t cnt) {
    if (cnt == 0) return;
    const uint32_t w0 = start >> 5, w1 = (start + cnt - 1u) >> 5;
    for (uint32_t w = w0 + lane_id(); w <= w1; w += 32) {
        const uint32_t lo = max(w * 32u, start), hi = min(w * 32u + 32u, start + cnt), c = hi - lo;
        const uint32_t bits = (c >= 32u ? 0xffffffffu : ((1u << c) - 1u)) << (lo & 31u);
        atomicAnd(&words[w], ~bits);
    }
}

// ---- the run-boundary walk: one warp per page ---------------------------------------------------------------------------
template <int W>
__device__ __forceinline__ void flat_scan_page(const DecodeParams& P, uint32_t e) {
    const uint32_t l = lane_id();
    const uint32_t q = flat_list_page(P, e);
    const pqg_page_desc pd = P.pages[q];
    const DevChunk& ck = P.chunks[pd.chunk_idx];
    FlatPage& fp = P.flat_pages[e];
    const uint32_t n = pd.num_values, size = pd.payload_size;
    // the page's blocks (and checkpoints): a range of the block arrays
    uint32_t blk0 = 0;
    if (l == 0) {
        blk0 = atomicAdd(&P.flat->nblk, static_cast<uint32_t>((static_cast<uint64_t>(n) + 1023u) >> 10));
        PQG_ASSERT(blk0 + ((static_cast<uint64_t>(n) + 1023u) >> 10) <= P.flat_blk_cap);
        fp.status = n ? 0u : 2u; fp.vpos = 0; fp.bw = 0; fp.nn = n; fp.blk0 = blk0; fp.page = q;
    }
    blk0 = __shfl_sync(0xffffffffu, blk0, 0);
    if (n == 0) return; // nothing to decode, nothing to hand over
    PQG_ASSERT(pd.out_row_base + n <= P.n_slots);
    if (ck.max_rep > 0 || ck.max_def > 1 || n > 0x7fffffffu) { if (l == 0) flat_hand_over(P, fp, q); return; } // (32-bit slot arithmetic below)
    // pages the general kernel stages whole in shared memory (what the tile kernel hands over from writer-shaped files):
    // one mostly empty block each here, a single pass there
    if (n <= 1024u && size <= static_cast<uint32_t>(kSlotBytes)) { if (l == 0) flat_hand_over(P, fp, q); return; }
    const uint8_t* pg = P.image + pd.payload_off;
    uint32_t pos = 0, def_len = 0;
    if (ck.max_def == 1) {
        if (size < 4u) { if (l == 0) flat_hand_over(P, fp, q); return; }
        def_len = ld32u(pg);
        if (def_len > size - 4u) { if (l == 0) flat_hand_over(P, fp, q); return; }
        pos = 4u + def_len;
    }
    // the walks stay 2 KB .. 6 KB ahead of themselves in L1 (lanes by 128-byte line)
    auto keep_ahead = [&](uint32_t at, uint32_t& pf_end) { // `at`: payload byte the walk is at
        if (at + 2048u > pf_end && pf_end < size) { const uint32_t from = max(pf_end, at), to = min(size, from + 4096u); prefetch_range(pg, from, to); pf_end = to; }
    };
    const bool dict_page = (pd.flags & PQG_PAGE_FLAG_DICT) && ck.has_dict;
    // ---- dictionary indices first (nothing is written for the page until this walk is through) ----
    if (dict_page) {
        if (pos >= size) { if (l == 0) flat_hand_over(P, fp, q); return; }
        const uint32_t bw = pg[pos];
        if (bw > 32u) { if (l == 0) flat_hand_over(P, fp, q); return; }
        const uint8_t* stream = pg + pos + 1u;
        const uint32_t slen = size - pos - 1u;
        Walker w;
        walker_init(w, stream, slen, slen, bw);
        uint2* ckpt = P.flat_ckpt + blk0; // one entry per 1024 values
        uint32_t v = 0, pf = pos; // values visited (the page holds at most n)
        bool bad = false, seq = false;
        while (v < n) {
            keep_ahead(pos + 1u + w.pos, pf);
            if (!seq) {
                // a run that fills the window by itself (literal runs of foreign writers: 504 values) needs no chain
                const RunAt head = parse_run_at(stream, slen, w.pos, bw);
                if (head.cnt && head.end - w.pos >= 32u) {
                    const uint32_t cnt = min(head.cnt, n - v);
                    for (uint32_t j = ((v + 1023u) >> 10) + l; (j << 10) < v + cnt; j += 32) ckpt[j] = make_uint2(w.pos, v);
                    v += cnt;
                    w.pos = head.end;
                    continue;
                }
                RunAt r;
                bool cut;
                const uint32_t used = window_runs(stream, slen, w.pos, bw, r, &cut);
                const uint32_t incl = warp_incl_scan(r.cnt);
                const uint32_t v0 = v + incl - r.cnt;
                if (r.cnt && v0 < n) { // checkpoints of the multiples of 1024 this lane's run holds
                    const uint32_t vend = v0 + min(r.cnt, n - v0);
                    for (uint32_t j = (v0 + 1023u) >> 10; (j << 10) < vend; j++) ckpt[j] = make_uint2(w.pos + l, v0);
                }
                v = min(n, v + __shfl_sync(0xffffffffu, incl, 31));
                w.pos += used;
                seq = cut;
                continue;
            }
            // one run by the sequential walker (it owns long varints, zero-length runs, the stream's end)
            seq = false;
            const uint32_t hdr = w.pos;
            if (!walker_next_run(w)) break;
            if (w.bad) { bad = true; break; }
            const uint32_t cnt = min(w.rem, n - v);
            if (w.lit) {
                // literal data beyond the page: bounded reads in the general kernel
                if (static_cast<uint64_t>(w.rem) * bw > static_cast<uint64_t>(slen - w.pos) * 8u) { bad = true; break; }
                w.pos = w.next_pos;
            }
            for (uint32_t j = ((v + 1023u) >> 10) + l; (j << 10) < v + cnt; j += 32) ckpt[j] = make_uint2(hdr, v);
            v += cnt;
        }
        if (bad) { if (l == 0) flat_hand_over(P, fp, q); return; }
        // stream exhausted: the remaining values read as index 0 (rle_decoder.hpp:21-24) -- checkpoints at the stream's end
        for (uint32_t j = ((v + 1023u) >> 10) + l; (j << 10) < n; j += 32) ckpt[j] = make_uint2(slen, v);
        if (l == 0) { fp.vpos = pos + 1u; fp.bw = bw; }
    } else if (l == 0) fp.vpos = pos;
    // ---- definition levels -> the column's validity bitmap ----
    if (ck.max_def == 1) {
        PQG_ASSERT(P.validity != nullptr);
        uint32_t* gv = P.validity + (pd.out_row_base >> 5);
        const uint32_t b0 = static_cast<uint32_t>(pd.out_row_base & 31u);
        Walker w;
        walker_init(w, pg + 4, def_len, size - 4u, 1u);
        uint32_t slots = 0, pf = 0;
        bool bad = false, seq = false;
        while (slots < n) {
            keep_ahead(4u + w.pos, pf);
            if (!seq) {
                RunAt r;
                bool cut;
                const uint32_t used = window_runs(w.s, def_len, w.pos, 1u, r, &cut);
                const uint32_t incl = warp_incl_scan(r.cnt);
                const uint32_t s0 = slots + incl - r.cnt;
                const uint32_t cnt_l = (r.cnt && s0 < n) ? min(r.cnt, n - s0) : 0u;
                const bool present = r.lit || level_present(r.data, 1);
                if (cnt_l && cnt_l <= 64u && present) { // short runs: every lane expands its own (one or two words of bits)
                    for (uint32_t k = 0; k < cnt_l; k += 32) {
                        const uint32_t c = min(32u, cnt_l - k);
                        uint32_t m = c >= 32u ? 0xffffffffu : ((1u << c) - 1u);
                        if (r.lit) m &= ldbits(w.s, r.data + k, c);
                        set_bits_word(gv, b0 + s0 + k, m);
                    }
                }
                uint32_t longer = __ballot_sync(0xffffffffu, cnt_l > 64u && present); // long runs: the warp expands them together
                while (longer) {
                    const int src = __ffs(static_cast<int>(longer)) - 1;
                    longer &= longer - 1u;
                    const uint32_t rs0 = __shfl_sync(0xffffffffu, s0, src), rc = __shfl_sync(0xffffffffu, cnt_l, src);
                    const uint32_t rl = __shfl_sync(0xffffffffu, r.lit, src), rd = __shfl_sync(0xffffffffu, r.data, src);
                    or_bits_warp(gv, b0 + rs0, rc, rl ? w.s : nullptr, rd);
                }
                slots = min(n, slots + __shfl_sync(0xffffffffu, incl, 31));
                w.pos += used;
                seq = cut;
                continue;
            }
            // one run by the sequential walker
            seq = false;
            if (!walker_next_run(w)) break; // a short stream leaves the remaining slots null (rle_decoder.hpp:21-24)
            if (w.bad) { bad = true; break; }
            const uint32_t cnt = min(w.rem, n - slots);
            if (w.lit) {
                if (static_cast<uint64_t>(w.rem) > static_cast<uint64_t>(def_len - w.pos) * 8u) { bad = true; break; }
                or_bits_warp(gv, b0 + slots, cnt, w.s, w.bit);
                w.pos = w.next_pos;
            } else if (level_present(w.val, 1)) or_bits_warp(gv, b0 + slots, cnt, nullptr, 0);
            slots += cnt;
        }
        if (bad) { // the general kernel redoes the page from a clean slate
            __syncwarp();
            clear_bits_warp(gv, b0, slots);
            if (l == 0) flat_hand_over(P, fp, q);
            return;
        }
    }
}

template <int W>
__global__ void __launch_bounds__(kThreadsPerCta) k_flat_scan(const DecodeParams P) {
    const uint32_t total = flat_list_len(P);
    if (blockIdx.x * kWarpsPerCta >= total) return; // (more warps than pages: no traffic on the cursor)
    for (;;) {
        const uint32_t e = warp_take(&P.flat->scan_cursor);
        if (e >= total) break;
        flat_scan_page<W>(P, e);
        __syncwarp();
    }
}

// ---- ranks of the blocks ----------------------------------------------------------------------------------------------
// validity word of page-relative slots [s, s + 32) (s < n), bits at and beyond slot n cleared
__device__ __forceinline__ uint32_t page_valid_word(const uint32_t* validity, uint64_t row_base, uint32_t s, uint32_t n, bool all_valid) {
    uint32_t w = 0xffffffffu;
    if (!all_valid) {
        const uint64_t g = row_base + s;
        const uint32_t sh = static_cast<uint32_t>(g & 31u);
        const uint32_t lo = validity[g >> 5], hi = sh ? validity[(g >> 5) + 1u] : 0u;
        w = __funnelshift_r(lo, hi, sh);
    }
    const uint32_t left = n - s;
    return left >= 32u ? w : (w & ((1u << left) - 1u));
}

template <int W>
__device__ __forceinline__ void flat_rank_page(const DecodeParams& P, uint32_t e) {
    const uint32_t l = lane_id();
    FlatPage& fp = P.flat_pages[e];
    const uint32_t q = fp.page;
    const pqg_page_desc pd = P.pages[q];
    const uint32_t n = pd.num_values;
    const uint32_t nblk = static_cast<uint32_t>((static_cast<uint64_t>(n) + 1023u) >> 10);
    FlatBlk* blk = P.flat_blk + fp.blk0;
    if (fp.status) { for (uint32_t b = l; b < nblk; b += 32) blk[b].nslots = 0; return; }
    const DevChunk& ck = P.chunks[pd.chunk_idx];
    const bool all_valid = ck.max_def <= 0;
    const bool dict_page = (pd.flags & PQG_PAGE_FLAG_DICT) && ck.has_dict;
    const uint32_t vpos = fp.vpos, bw = fp.bw;
    uint32_t run = 0;
    for (uint32_t b0 = 0; b0 < nblk; b0 += 32) { // lane = block
        const uint32_t b = b0 + l;
        uint32_t c = 0;
        if (b < nblk) {
            if (all_valid) c = min(1024u, n - (b << 10));
            else for (uint32_t k = 0; k < 32u && (b << 10) + 32u * k < n; k++) c += __popc(page_valid_word(P.validity, pd.out_row_base, (b << 10) + 32u * k, n, false));
        }
        const uint32_t incl = warp_incl_scan(c);
        if (b < nblk) {
            const uint32_t rank0 = run + incl - c;
            FlatBlk d;
            d.row = pd.out_row_base + (b << 10);
            d.nslots = min(1024u, n - (b << 10));
            d.rank0 = rank0;
            d.cp_pos = 0; d.cp_first = 0; d.slen = 0;
            d.info = bw | (dict_page ? 0x100u : 0u) | (all_valid ? 0x200u : 0u);
            d.chunk = pd.chunk_idx; d.pad = 0;
            if (dict_page) {
                d.src = pd.payload_off + vpos;
                d.slen = pd.payload_size - vpos;
                if (c) { const uint2 cp = P.flat_ckpt[fp.blk0 + (rank0 >> 10)]; d.cp_pos = cp.x; d.cp_first = cp.y; }
            } else d.src = pd.payload_off + vpos + static_cast<uint64_t>(rank0) * W;
            blk[b] = d;
        }
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
    // PLAIN: the values must be there (the general kernel reports the truncated page)
    if (!dict_page && static_cast<uint64_t>(run) * W > pd.payload_size - vpos) {
        __syncwarp();
        for (uint32_t b = l; b < nblk; b += 32) blk[b].nslots = 0;
        if (l == 0) flat_hand_over(P, fp, q);
    }
}

template <int W>
__global__ void __launch_bounds__(kThreadsPerCta) k_flat_ranks(const DecodeParams P) {
    const uint32_t total = flat_list_len(P);
    if (blockIdx.x * kWarpsPerCta >= total) return;
    for (;;) {
        const uint32_t e = warp_take(&P.flat->rank_cursor);
        if (e >= total) break;
        flat_rank_page<W>(P, e);
        __syncwarp();
    }
}

// ---- emission -----------------------------------------------------------------------------------------------------------
template <int W, bool IDX16>
__device__ __forceinline__ void flat_emit_block(const DecodeParams& P, uint32_t g, uint32_t idx_s, const uint8_t* sdict, uint32_t sdict_chunk) {
    using T = typename FlatElem<W>::T;
    const uint32_t l = lane_id();
    const FlatBlk d = P.flat_blk[g];
    if (d.nslots == 0) return;
    const uint32_t nslots = d.nslots, rank0 = d.rank0;
    const bool all_valid = d.info & 0x200u, dict_page = d.info & 0x100u;
    // this lane's validity word: slots [32 l, + 32) of the block
    const uint32_t vw = 32u * l < nslots ? page_valid_word(P.validity, d.row, 32u * l, nslots, all_valid) : 0u;
    const uint32_t c = __popc(vw);
    const uint32_t incl = warp_incl_scan(c);
    const uint32_t excl = incl - c;                       // values of the block in front of this lane's word
    const uint32_t cntb = __shfl_sync(0xffffffffu, incl, 31);
    T* out = reinterpret_cast<T*>(P.values) + d.row;
    const uint32_t nsteps = (nslots + 31u) >> 5;
    if (!dict_page) {
        const uint8_t* src = P.image + d.src;
        const uint32_t kmax = cntb ? cntb - 1u : 0u;
        for (uint32_t st = 0; st < nsteps; st += kFlatUnroll) {
            T x[kFlatUnroll];
            bool v[kFlatUnroll];
#pragma unroll
            for (uint32_t u = 0; u < kFlatUnroll; u++) {
                const uint32_t sw = min(st + u, 31u);
                const uint32_t word = __shfl_sync(0xffffffffu, vw, sw), base = __shfl_sync(0xffffffffu, excl, sw);
                v[u] = (word >> l) & 1u;
                const uint32_t k = min(base + __popc(word & ((1u << l) - 1u)), kmax); // (clamped: null slots load and discard)
                x[u] = cntb ? flat_ld<W>(src + static_cast<size_t>(k) * W) : T(0);
            }
#pragma unroll
            for (uint32_t u = 0; u < kFlatUnroll; u++) {
                const uint32_t sl = 32u * (st + u) + l;
                if (st + u < nsteps && sl < nslots) __stcs(out + sl, v[u] ? x[u] : T(0));
            }
        }
        return;
    }
    // ---- dictionary block: indices of values [rank0, rank0 + cntb) -> shared memory ----
    const uint32_t bw = d.info & 63u;
    const uint8_t* stream = P.image + d.src;
    const uint32_t slen = d.slen;
    const DevChunk& ck = P.chunks[d.chunk];
    if (cntb) {
        const uint2 cp = make_uint2(d.cp_pos, d.cp_first);
        // bytes the walk will touch: the values skipped + taken, their run headers
        const uint32_t est = static_cast<uint32_t>((static_cast<uint64_t>(rank0 - cp.y + cntb) * bw + 7u) >> 3) + 256u;
        prefetch_range(stream, cp.x, min(slen, cp.x + min(est, 16384u)));
        Walker w;
        walker_init(w, stream, slen, slen, bw);
        w.pos = cp.x;
        uint32_t v = cp.y;
        const uint32_t vend = rank0 + cntb;
        const uint32_t imask = bw >= 32u ? 0xffffffffu : ((1u << bw) - 1u);
        while (v < vend) {
            if (!walker_next_run(w)) break;
            const uint32_t cnt = w.rem; // (validated by k_flat_scan: no bad runs, literal data inside the page)
            const uint32_t lo = max(v, rank0), hi = min(v + min(cnt, vend - v), vend);
            if (w.lit) {
                for (uint32_t t = lo + l; t < hi; t += 32u * kFlatUnroll) { // four loads in flight per lane, then the stores
                    uint32_t x[kFlatUnroll];
#pragma unroll
                    for (uint32_t u = 0; u < kFlatUnroll; u++) x[u] = ldbits(stream, w.bit + (min(t + 32u * u, hi - 1u) - v) * bw, bw) & imask;
#pragma unroll
                    for (uint32_t u = 0; u < kFlatUnroll; u++) if (t + 32u * u < hi) sts_idx<IDX16>(idx_s, t + 32u * u - rank0, x[u]);
                }
                w.pos = w.next_pos;
            } else {
                for (uint32_t t = lo + l; t < hi; t += 32) sts_idx<IDX16>(idx_s, t - rank0, w.val);
            }
            v += min(cnt, vend - v);
        }
        for (uint32_t t = max(v, rank0) + l; t < vend; t += 32) sts_idx<IDX16>(idx_s, t - rank0, 0u); // exhausted stream: zeros
    }
    __syncwarp();
    const T* dict = reinterpret_cast<const T*>(P.dict_arena + ck.dict_arena_off);
    const uint32_t dict_n = ck.dict_ok_n;
    const bool staged = d.chunk == sdict_chunk; // the CTA holds this chunk's dictionary in shared memory
    const uint32_t sdict_s = static_cast<uint32_t>(__cvta_generic_to_shared(sdict));
    uint32_t* gv = P.validity;
    uint32_t nbad = 0;
    for (uint32_t st = 0; st < nsteps; st += kFlatUnroll) {
        uint32_t ix[kFlatUnroll];
        bool in[kFlatUnroll];
#pragma unroll
        for (uint32_t u = 0; u < kFlatUnroll; u++) {
            const uint32_t sw = min(st + u, 31u);
            const uint32_t word = __shfl_sync(0xffffffffu, vw, sw), base = __shfl_sync(0xffffffffu, excl, sw);
            const bool valid = st + u < nsteps && ((word >> l) & 1u);
            ix[u] = lds_idx<IDX16>(idx_s, min(base + __popc(word & ((1u << l) - 1u)), 1023u)); // (null slots read and discard)
            in[u] = valid && ix[u] < dict_n;
            if (valid && !in[u]) { // out-of-range index: NULL (column_reader.cpp:190-194)
                const uint64_t gs = d.row + 32u * (st + u) + l;
                if (gv) atomicAnd(&gv[gs >> 5], ~(1u << (gs & 31u)));
                else nbad++;
            }
        }
        T x[kFlatUnroll];
#pragma unroll
        for (uint32_t u = 0; u < kFlatUnroll; u++) {
            x[u] = 0;
            if (in[u]) x[u] = P.identity_dict ? static_cast<T>(ix[u]) : (staged ? lds_elem<W>(sdict_s + static_cast<uint32_t>(W) * ix[u]) : __ldg(dict + ix[u]));
        }
#pragma unroll
        for (uint32_t u = 0; u < kFlatUnroll; u++) {
            const uint32_t sl = 32u * (st + u) + l;
            if (st + u < nsteps && sl < nslots) __stcs(out + sl, x[u]);
        }
    }
    if (nbad) atomicAdd(&P.err->bad_index, nbad); // REQUIRED-only plan: pqg_plan_finish adds a validity bitmap and re-runs
}

// Persistent CTAs take batches of kFlatBatch consecutive blocks (consecutive blocks = consecutive pages of a chunk) and keep
// the dictionary of the batch's chunk in shared memory when it fits (<= kMaxSmemDictBytes, as the tile kernel does): a
// gather from L1 costs one wavefront per lane, 40 M of them were the whole emission time of a 4096-entry dictionary column.
constexpr uint32_t kFlatBatch = 32;
template <int W, bool IDX16>
__global__ void __launch_bounds__(kThreadsPerCta) k_flat_emit(const DecodeParams P, uint32_t with_dict) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ uint32_t s_batch, s_chunk;
    constexpr uint32_t kIdxBytes = IDX16 ? 2048u : 4096u;
    const uint32_t idx_s = static_cast<uint32_t>(__cvta_generic_to_shared(smem)) + warp_id() * kIdxBytes; // the block's dictionary indices, by value
    uint8_t* sdict = smem + kWarpsPerCta * kIdxBytes;
    const uint32_t total = P.flat->nblk;
    if (blockIdx.x * kFlatBatch >= total) return;
    uint32_t cur_chunk = 0xffffffffu;
    for (;;) {
        __syncthreads(); // everyone is through with the previous batch
        if (threadIdx.x == 0) s_batch = atomicAdd(&P.flat->emit_cursor, kFlatBatch);
        __syncthreads();
        const uint32_t g0 = s_batch;
        if (g0 >= total) break;
        const uint32_t g1 = min(total, g0 + kFlatBatch);
        if (with_dict && !P.identity_dict) {
            if (warp_id() == 0) { // chunk of the batch's first dictionary block
                const uint32_t g = g0 + lane_id();
                bool want = false;
                uint32_t c = 0;
                if (g < g1) { const FlatBlk& b = P.flat_blk[g]; want = b.nslots && (b.info & 0x100u); c = b.chunk; }
                const uint32_t m = __ballot_sync(0xffffffffu, want);
                const uint32_t pick = m ? __shfl_sync(0xffffffffu, c, __ffs(static_cast<int>(m)) - 1) : 0xffffffffu;
                if (lane_id() == 0) s_chunk = pick;
            }
            __syncthreads();
            const uint32_t want_chunk = s_chunk;
            if (want_chunk != 0xffffffffu && want_chunk != cur_chunk) {
                const DevChunk& ck = P.chunks[want_chunk];
                const uint32_t bytes = ck.dict_ok_n * static_cast<uint32_t>(W);
                cur_chunk = 0xffffffffu;
                if (ck.dict_ok_n <= static_cast<uint32_t>(kMaxSmemDictBytes) / W) {
                    const uint4* src = reinterpret_cast<const uint4*>(P.dict_arena + ck.dict_arena_off);
                    for (uint32_t i = threadIdx.x; i < (bytes + 15u) >> 4; i += kThreadsPerCta) reinterpret_cast<uint4*>(sdict)[i] = src[i];
                    cur_chunk = want_chunk;
                }
                __syncthreads();
            }
        }
        for (uint32_t g = g0 + warp_id(); g < g1; g += kWarpsPerCta) {
            flat_emit_block<W, IDX16>(P, g, idx_s, sdict, cur_chunk);
            __syncwarp();
        }
    }
}

} // namespace

uint32_t flat_launches() { return 3; }

// p.flat (the cursors) must be zero: the caller resets it with the other work counters of the (sub-)run
template <int W, bool IDX16>
static cudaError_t launch_flat_t(const DecodeParams& p, int sm_count, bool any_dict, cudaStream_t s) {
    const unsigned grid = static_cast<unsigned>(sm_count) * 6u; // 48 warps per SM (<= 48 registers)
    // emission of plans with dictionaries: per-warp index buffers + the staged dictionary = 48 / 64 KB per CTA, 4 / 3 CTAs per SM
    const size_t smem = any_dict ? static_cast<size_t>(kWarpsPerCta) * (IDX16 ? 2048u : 4096u) + kMaxSmemDictBytes : 0;
    const unsigned egrid = static_cast<unsigned>(sm_count) * (any_dict ? (IDX16 ? 4u : 3u) : 6u);
    cudaError_t e = cudaFuncSetAttribute(k_flat_emit<W, IDX16>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    k_flat_scan<W><<<grid, kThreadsPerCta, 0, s>>>(p);
    k_flat_ranks<W><<<grid, kThreadsPerCta, 0, s>>>(p);
    k_flat_emit<W, IDX16><<<egrid, kThreadsPerCta, smem, s>>>(p, any_dict ? 1u : 0u);
    return cudaGetLastError();
}
cudaError_t launch_flat_pages(const DecodeParams& p, int width, int sm_count, bool any_dict, bool idx16, cudaStream_t s) {
    if (width == 4) return idx16 ? launch_flat_t<4, true>(p, sm_count, any_dict, s) : launch_flat_t<4, false>(p, sm_count, any_dict, s);
    if (width == 8) return idx16 ? launch_flat_t<8, true>(p, sm_count, any_dict, s) : launch_flat_t<8, false>(p, sm_count, any_dict, s);
    return cudaErrorInvalidValue;
}

} // namespace pqg
