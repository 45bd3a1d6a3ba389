// pqg_bigpage.cu -- oversized fixed-width pages: one CTA per page.
//
// Foreign writers (pyarrow, DuckDB, ...) emit 64 KB .. 1 MB data pages with tens of thousands
// of values; the reference's own writer never does (include/writer/parquet_writer.hpp:35).
// One warp per page, as the general kernel works, leaves such files with a few hundred warps
// of work.  Here the whole CTA takes the page (reference semantics: read_data_page,
// src/reader/column_reader.cpp:140-225):
//   * definition levels: handled when the page has none, or when the first RLE run covers
//     the whole page with the maximum level (a column without nulls in this page -- what
//     nullable-by-default writers produce most of the time);
//   * PLAIN values: CTA-wide shifted 16-byte vector copy;
//   * dictionary indices: the RLE / bit-packed hybrid stream is staged through shared memory
//     in 48 KB chunks; ONE thread walks the run headers of the chunk into a run table
//     (start value, kind, bit offset or value) -- the run-boundary scan -- and all threads
//     expand the runs in parallel (binary search of the value's run, bit extraction,
//     dictionary gather, coalesced stores).
// Pages this kernel does not take (nulls present, nested levels, bit width > 32, bad runs,
// out-of-range indices, truncation) are handed to the general kernel through the slow list,
// which also does all error reporting.
#include "pqg_page.cuh"

namespace pqg {
namespace {

constexpr int kBigThreads = 256;
constexpr int kBigChunk = 40 * 1024;   // staged stream bytes per step
constexpr int kScanWin = 4096;         // bytes of the staged slice the parallel run-boundary scan looks at per step (16 per thread)
constexpr int kSeqRunBudget = 48;      // runs the one-thread scan may walk before the stream counts as dense (-> parallel scan)
constexpr uint32_t kNxtInvalid = 0xffffu, kNxtBad = 0xfffeu; // header / data reach beyond the staged slice; zero-length run
constexpr int kBigRuns = 1024;         // run-table entries per step
constexpr uint32_t kEmitUnroll = 4;    // slots per thread and trip of the emission loops (loads in flight per lane)
constexpr int kBigSlots = 131072;      // slots per page the null-aware path can hold (validity image + ranks in shared memory)

template <int W> struct BElem;
template <> struct BElem<4> { using T = uint32_t; };
template <> struct BElem<8> { using T = uint64_t; };
template <int W> __device__ __forceinline__ typename BElem<W>::T ld_val(const uint8_t* p);
template <> __device__ __forceinline__ uint32_t ld_val<4>(const uint8_t* p) { return ld32u(p); }
template <> __device__ __forceinline__ uint64_t ld_val<8>(const uint8_t* p) { return ld64u(p); }

struct BigSmem {
    uint32_t run_first[kBigRuns + 1]; // first value index of run r (run_first[n_runs] = end)
    uint32_t run_data[kBigRuns];      // literal: bit offset inside the staged chunk; RLE: the value
    uint8_t run_lit[kBigRuns];
    uint32_t n_runs, next_pos, next_val, status; // status: 0 ok, 1 hand the page over
    uint32_t flag, all_valid, nn, lo_slot, hi_slot;
    uint16_t jmp[kScanWin];           // run-boundary scan: position after the run whose header would start at byte i, doubled per round
    uint16_t endp[kBigRuns + 2];      // ... position after run r
    uint32_t reach[kScanWin / 32];    // ... byte i IS a run header (on the chain from byte 0)
    uint32_t changed, n_true, inv_seen, dense;
    unsigned long long vtotal;
    uint32_t pv[kBigSlots / 32];      // page-relative validity words (pages with nulls)
    uint32_t rb[kBigSlots / 32 + 1];  // rank of the first slot of every word (exclusive popcount prefix)
    __align__(16) uint8_t chunk[kBigChunk + 32];
};

__device__ __forceinline__ void hand_over(const DecodeParams& P, uint32_t q) {
    uint32_t k = atomicAdd(&P.err->slow_count, 1u);
    PQG_ASSERT(k < P.slow_cap);
    P.slow_append[k] = q;
}

// stage `take` bytes of a stream slice into S.chunk keeping the source's 16-byte phase; returns the phase
__device__ __forceinline__ const uint8_t* stage_slice(BigSmem& S, const uint8_t* g, uint32_t take) {
    const uint32_t mis = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(g) & 15u);
    const uint32_t nvec = (mis + take + 15u) >> 4;
    for (uint32_t j = threadIdx.x; j < nvec; j += kBigThreads) reinterpret_cast<uint4*>(S.chunk)[j] = ldg_nc16(g - mis + 16u * j);
    __syncthreads();
    return S.chunk + mis;
}

// Run-boundary scan of a staged slice by ONE thread (RleDecoder::next_counts,
// rle_decoder.hpp:37-53): fills the run table for values [v0, ...) until the table or the slice
// is full.  spos = stream position of cb[0], slen = stream length.
__device__ __forceinline__ void scan_runs(BigSmem& S, const uint8_t* cb, uint32_t take, uint32_t bw, uint32_t n, uint32_t v0,
                                          uint32_t spos, uint32_t slen, uint32_t max_runs = kBigRuns) {
    const uint32_t nb = (bw + 7u) >> 3;
    uint32_t p = 0, v = v0, r = 0, st = 0;
    while (r < max_runs && v < n) {
        if (spos + p >= slen) { // stream exhausted: the remaining values read as 0 (:21-24)
            S.run_first[r] = v; S.run_data[r] = 0; S.run_lit[r] = 0; r++; v = n; break;
        }
        uint32_t ind = 0, shift = 0, hp = p;
        bool complete = false;
        while (hp < take) { uint32_t b = cb[hp++]; if (shift < 32) ind |= (b & 0x7Fu) << shift; if (!(b & 0x80u)) { complete = true; break; } shift += 7; }
        if (!complete) break; // the header continues in the next slice
        if (ind & 1u) {
            const uint64_t cnt = static_cast<uint64_t>(ind >> 1) * 8u;
            const uint64_t dbytes = (cnt * bw + 7u) >> 3;
            if (cnt == 0) { st = 1; break; }
            if (hp + dbytes > take) { if (p == 0) st = 1; break; } // does not fit: restage from this header
            S.run_first[r] = v; S.run_data[r] = hp * 8u; S.run_lit[r] = 1; r++;
            v += cnt < static_cast<uint64_t>(n - v) ? static_cast<uint32_t>(cnt) : (n - v);
            p = hp + static_cast<uint32_t>(dbytes);
        } else {
            const uint32_t cnt = ind >> 1;
            if (cnt == 0) { st = 1; break; }
            if (hp + nb > take) { if (p == 0) st = 1; break; }
            uint32_t val = 0;
            for (uint32_t i = 0; i < nb && i < 4u; i++) val |= static_cast<uint32_t>(cb[hp + i]) << (8u * i);
            S.run_first[r] = v; S.run_data[r] = val; S.run_lit[r] = 0; r++;
            v += min(cnt, n - v);
            p = hp + nb;
        }
    }
    if (r == 0 && st == 0) st = 1; // no progress: the general kernel decides
    S.run_first[r] = v;
    S.n_runs = r; S.next_pos = spos + p; S.next_val = v; S.status = st;
}

// ---- the run-boundary scan, in parallel (all kBigThreads threads) -----------------------------------------------
// RleDecoder::next_counts (rle_decoder.hpp:37-53) walks the run headers one after the other: header -> size of its
// payload -> next header.  Streams of many short runs (definition levels of columns with scattered nulls as foreign
// writers emit them: literal groups and RLE runs of 8+ in turns; run-heavy dictionary indices) make that walk the whole
// cost of a page.  Here EVERY byte of a window is taken for a header: jmp[i] = the position the run starting at i would end
// at.  The true headers are the chain 0 -> jmp[0] -> ...; it is marked by pointer doubling -- round r marks the successors
// of everything marked so far (reaching chain distance 2^(r+1)) and squares the jump table -- in log2(runs) rounds, then
// the marked headers are compacted in order (block scans of their number and of their value counts) into the same run
// table scan_runs fills.  Same contract as scan_runs: runs for values [v0, ...) until the table, the window or the values
// end; S.status = 1 hands the page over.
struct RunHdr { uint32_t next, cnt, lit, data; };
__device__ __forceinline__ RunHdr parse_run_header(const uint8_t* cb, uint32_t take, uint32_t i, uint32_t bw, uint32_t nb) {
    RunHdr h{kNxtInvalid, 0, 0, 0};
    uint32_t ind = 0, shift = 0, hp = i;
    bool complete = false;
    while (hp < take && hp < i + 10u) { const uint32_t b = cb[hp++]; if (shift < 32) ind |= (b & 0x7Fu) << shift; if (!(b & 0x80u)) { complete = true; break; } shift += 7; }
    if (!complete) return h; // the header continues in the next slice (or is no varint at all)
    if (ind & 1u) {
        const uint64_t cnt = static_cast<uint64_t>(ind >> 1) * 8u, dbytes = (cnt * bw + 7u) >> 3;
        if (cnt == 0) { h.next = kNxtBad; return h; }
        if (hp + dbytes > take) return h;
        h.next = hp + static_cast<uint32_t>(dbytes); h.cnt = static_cast<uint32_t>(cnt); h.lit = 1; h.data = hp * 8u;
    } else {
        const uint32_t cnt = ind >> 1;
        if (cnt == 0) { h.next = kNxtBad; return h; }
        if (hp + nb > take) return h;
        uint32_t val = 0;
        for (uint32_t k = 0; k < nb && k < 4u; k++) val |= static_cast<uint32_t>(cb[hp + k]) << (8u * k);
        h.next = hp + nb; h.cnt = cnt; h.lit = 0; h.data = val;
    }
    return h;
}

__device__ __forceinline__ unsigned long long block_excl_scan_u64(unsigned long long v, unsigned long long* total) {
    __shared__ unsigned long long wsum[kBigThreads / 32];
    const uint32_t l = threadIdx.x & 31u, w = threadIdx.x >> 5;
    unsigned long long incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d); if (l >= static_cast<uint32_t>(d)) incl += t; }
    if (l == 31) wsum[w] = incl;
    __syncthreads();
    unsigned long long base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < kBigThreads / 32; i++) { const unsigned long long x = wsum[i]; if (i < static_cast<int>(w)) base += x; tot += x; }
    __syncthreads();
    *total = tot;
    return base + incl - v;
}

__device__ __noinline__ void scan_runs_parallel(BigSmem& S, const uint8_t* cb, uint32_t take, uint32_t bw, uint32_t n, uint32_t v0,
                                               uint32_t spos, uint32_t slen) {
    const uint32_t tid = threadIdx.x;
    const uint32_t nb = (bw + 7u) >> 3;
    const uint32_t W = min(take, static_cast<uint32_t>(kScanWin));
    constexpr uint32_t kPer = kScanWin / kBigThreads; // positions per thread (contiguous)
    // A: the jump table
    for (uint32_t k = 0; k < kPer; k++) {
        const uint32_t i = tid * kPer + k;
        PQG_ASSERT(take <= static_cast<uint32_t>(kBigChunk) + 32u);
        if (i < W) S.jmp[i] = static_cast<uint16_t>(parse_run_header(cb, take, i, bw, nb).next);
    }
    if (tid < kScanWin / 32) S.reach[tid] = tid == 0 ? 1u : 0u;
    if (tid == 0) { S.status = 0; S.inv_seen = 0; }
    __syncthreads();
    // B: mark the chain by pointer doubling
    for (int round = 0; round < 13; round++) {
        if (tid == 0) S.changed = 0;
        __syncthreads();
        const uint32_t mine = (S.reach[(tid * kPer) >> 5] >> ((tid * kPer) & 31u)) & ((1u << kPer) - 1u);
        uint32_t m = mine;
        while (m) {
            const uint32_t k = static_cast<uint32_t>(__ffs(static_cast<int>(m)) - 1);
            m &= m - 1;
            const uint32_t j = S.jmp[tid * kPer + k];
            if (j < W) {
                PQG_ASSERT((j >> 5) < static_cast<uint32_t>(kScanWin) / 32u);
                const uint32_t bit = 1u << (j & 31u);
                if (!(atomicOr(&S.reach[j >> 5], bit) & bit)) S.changed = 1;
            }
        }
        __syncthreads();
        if (!S.changed) break;
        uint16_t nj[kPer];
#pragma unroll
        for (uint32_t k = 0; k < kPer; k++) {
            const uint32_t i = tid * kPer + k;
            uint32_t j = i < W ? S.jmp[i] : kNxtInvalid;
            if (j < W) j = S.jmp[j];
            nj[k] = static_cast<uint16_t>(j);
        }
        __syncthreads();
#pragma unroll
        for (uint32_t k = 0; k < kPer; k++) { const uint32_t i = tid * kPer + k; if (i < W) S.jmp[i] = nj[k]; }
        __syncthreads();
    }
    // C: the marked headers, in order: thread t takes the 16 positions it owns
    const uint32_t mine = (S.reach[(tid * kPer) >> 5] >> ((tid * kPer) & 31u)) & ((1u << kPer) - 1u);
    unsigned long long vsum = 0;
    uint32_t usable = 0; // headers with a complete run inside the slice
    {
        uint32_t m = mine;
        while (m) {
            const uint32_t k = static_cast<uint32_t>(__ffs(static_cast<int>(m)) - 1);
            m &= m - 1;
            const RunHdr h = parse_run_header(cb, take, tid * kPer + k, bw, nb);
            if (h.next == kNxtBad) S.status = 1;           // zero-length run: the general kernel reports it
            else if (h.next == kNxtInvalid) S.inv_seen = 1; // the chain ends here: restage from this header
            else { usable++; vsum += h.cnt; }
        }
    }
    unsigned long long vtot = 0, rtot = 0;
    const unsigned long long vbase = block_excl_scan_u64(vsum, &vtot);
    const uint32_t rbase = static_cast<uint32_t>(block_excl_scan_u64(usable, &rtot));
    {
        uint32_t m = mine, r = rbase;
        unsigned long long v = vbase;
        while (m) {
            const uint32_t k = static_cast<uint32_t>(__ffs(static_cast<int>(m)) - 1);
            m &= m - 1;
            const RunHdr h = parse_run_header(cb, take, tid * kPer + k, bw, nb);
            if (h.next >= kNxtBad) continue;
            if (r <= static_cast<uint32_t>(kBigRuns)) { // (entry kBigRuns: only the first value, the sentinel behind a full table)
                const unsigned long long vs = static_cast<unsigned long long>(v0) + v;
                S.run_first[r] = vs < n ? static_cast<uint32_t>(vs) : n;
            }
            if (r < static_cast<uint32_t>(kBigRuns)) {
                S.run_data[r] = h.data;
                S.run_lit[r] = static_cast<uint8_t>(h.lit);
                S.endp[r] = static_cast<uint16_t>(h.next);
            }
            r++;
            v += h.cnt;
        }
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t R = static_cast<uint32_t>(min(rtot, static_cast<unsigned long long>(kBigRuns)));
        const unsigned long long vend = static_cast<unsigned long long>(v0) + vtot;
        if (rtot <= static_cast<unsigned long long>(kBigRuns)) S.run_first[R] = vend < n ? static_cast<uint32_t>(vend) : n;
        uint32_t lo = 0, hi = R; // runs whose first value lies inside the page
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (S.run_first[mid] < n) lo = mid + 1; else hi = mid; }
        uint32_t nr = lo;
        uint32_t next_val = nr ? S.run_first[nr] : v0, next_pos = nr ? S.endp[nr - 1] : 0u;
        if (nr == 0 && !S.status && spos < slen) S.status = 1; // no progress inside the slice: the general kernel decides
        if (!S.status && spos + next_pos >= slen && next_val < n && nr < static_cast<uint32_t>(kBigRuns)) {
            // stream exhausted: the remaining values read as 0 (rle_decoder.hpp:21-24)
            S.run_first[nr] = next_val; S.run_data[nr] = 0; S.run_lit[nr] = 0; nr++;
            next_val = n;
            S.run_first[nr] = n;
        }
        S.n_runs = nr; S.next_pos = spos + next_pos; S.next_val = next_val;
    }
    __syncthreads();
}

__device__ __forceinline__ uint32_t find_run(const BigSmem& S, uint32_t nr, uint32_t v) {
    uint32_t lo = 0, hi = nr; // run with run_first[r] <= v < run_first[r + 1]
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (S.run_first[mid] <= v) lo = mid; else hi = mid; }
    return lo;
}

template <int W>
__global__ void __launch_bounds__(kBigThreads) k_big_pages(const DecodeParams P) {
    using T = typename BElem<W>::T;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    BigSmem& S = *reinterpret_cast<BigSmem*>(smem_raw);
    const uint32_t tid = threadIdx.x;
    const uint32_t n_host = P.slow_hi - P.slow_lo;
    for (uint32_t it = blockIdx.x; it < n_host; it += gridDim.x) {
        const uint32_t q = P.slow_pages[P.slow_lo + it];
        const pqg_page_desc pd = P.pages[q];
        const DevChunk& ck = P.chunks[pd.chunk_idx];
        const uint32_t n = pd.num_values, size = pd.payload_size;
        if (n == 0) continue;
        PQG_ASSERT(pd.out_row_base + n <= P.n_slots);
        const uint8_t* pg = P.image + pd.payload_off;
        __syncthreads();
        // ---- levels ----
        // none / one RLE run of the maximum level covering the page: every slot present;
        // otherwise (max_def == 1): validity image + ranks in shared memory
        if (tid == 0) {
            uint32_t st = 0, pos = 0, allv = 1;
            if (ck.max_rep > 0 || ck.max_def > 1) st = 1;
            else if (ck.max_def == 1) {
                if (size < 4) st = 1;
                else {
                    const uint32_t def_len = ld32u(pg);
                    if (def_len > size - 4) st = 1;
                    else {
                        uint32_t ind = 0, shift = 0, p = 4;
                        while (p < 4 + def_len) { uint32_t b = pg[p++]; if (shift < 32) ind |= (b & 0x7Fu) << shift; if (!(b & 0x80u)) break; shift += 7; }
                        if (def_len < 2 || (ind & 1u) || (ind >> 1) < n || p >= 4 + def_len || pg[p] != 1u) allv = 0;
                        if (!allv && (n > static_cast<uint32_t>(kBigSlots) || def_len > static_cast<uint32_t>(kBigChunk))) st = 1;
                        pos = 4 + def_len;
                    }
                }
            }
            S.status = st; S.next_pos = pos; S.all_valid = allv; S.nn = n;
        }
        __syncthreads();
        if (S.status) { if (tid == 0) hand_over(P, q); continue; }
        uint32_t pos = S.next_pos;
        const bool all_valid = S.all_valid != 0;
        const uint32_t nwords = (n + 31u) >> 5;
        bool ok = true;
        if (!all_valid) {
            // definition levels (bit width 1) -> page-relative validity words
            const uint32_t def_len = pos - 4u;
            for (uint32_t w = tid; w < nwords; w += kBigThreads) S.pv[w] = 0;
            const uint8_t* cb = stage_slice(S, pg + 4, def_len);
            uint32_t spos = 0, vdone = 0;
            bool dense = false; // many short runs: the one-thread header walk would be the whole cost -> parallel run-boundary scan
            while (vdone < n && ok) {
                if (!dense) {
                    if (tid == 0) scan_runs(S, cb + spos, def_len - spos, 1u, n, vdone, spos, def_len, kSeqRunBudget);
                    __syncthreads();
                    dense = !S.status && S.n_runs == static_cast<uint32_t>(kSeqRunBudget) && S.next_val < n;
                    __syncthreads();
                }
                if (dense) scan_runs_parallel(S, cb + spos, def_len - spos, 1u, n, vdone, spos, def_len);
                if (S.status) { ok = false; break; }
                const uint32_t nr = S.n_runs, vend = S.next_val;
                const uint8_t* rb_base = cb + spos; // literal bit offsets are relative to the scanned slice
                for (uint32_t w = (vdone >> 5) + tid; w <= ((vend - 1u) >> 5); w += kBigThreads) {
                    uint32_t sl = max(w * 32u, vdone);
                    const uint32_t hi = min(w * 32u + 32u, vend);
                    uint32_t bits = 0, r = find_run(S, nr, sl);
                    while (sl < hi) {
                        const uint32_t rend = S.run_first[r + 1];
                        const uint32_t cnt = min(hi, rend) - sl;
                        uint32_t m;
                        if (S.run_lit[r]) m = ldbits(rb_base, S.run_data[r] + (sl - S.run_first[r]), cnt);
                        else m = S.run_data[r] >= 1u ? (cnt >= 32u ? 0xffffffffu : ((1u << cnt) - 1u)) : 0u;
                        bits |= m << (sl & 31u);
                        sl += cnt;
                        r++;
                    }
                    PQG_ASSERT(w < static_cast<uint32_t>(kBigSlots) / 32u);
                    if (bits) atomicOr(&S.pv[w], bits);
                }
                __syncthreads();
                vdone = vend;
                spos = S.next_pos;
            }
            if (!ok) { __syncthreads(); if (tid == 0) hand_over(P, q); continue; }
            // ranks: exclusive prefix of the word popcounts (256 threads x up to 16 words)
            {
                const uint32_t per = (nwords + kBigThreads - 1) / kBigThreads;
                const uint32_t w0 = tid * per, w1 = min(nwords, w0 + per);
                uint32_t sum = 0;
                for (uint32_t w = w0; w < w1; w++) sum += __popc(S.pv[w]);
                __shared__ uint32_t wsum[kBigThreads / 32];
                const uint32_t l = tid & 31u, wp = tid >> 5;
                uint32_t incl = warp_incl_scan(sum);
                if (l == 31) wsum[wp] = incl;
                __syncthreads();
                uint32_t base = 0, total = 0;
                for (uint32_t i = 0; i < kBigThreads / 32; i++) { uint32_t x = wsum[i]; if (i < wp) base += x; total += x; }
                uint32_t run = base + incl - sum;
                for (uint32_t w = w0; w < w1; w++) { S.rb[w] = run; run += __popc(S.pv[w]); }
                if (tid == 0) { S.rb[nwords] = total; S.nn = total; }
                __syncthreads();
            }
        }
        const uint32_t nn = S.nn;
        const bool dict_page = (pd.flags & PQG_PAGE_FLAG_DICT) && ck.has_dict;
        T* out = reinterpret_cast<T*>(P.values) + pd.out_row_base;
        if (!dict_page) {
            if (static_cast<uint64_t>(nn) * W > size - pos) { if (tid == 0) hand_over(P, q); continue; }
            const uint8_t* src = pg + pos;
            if (all_valid) {
                // ---- PLAIN, no nulls: shifted copy, 16-byte vectors once the destination is aligned ----
                const uint64_t bytes = static_cast<uint64_t>(n) * W;
                uint8_t* dst = reinterpret_cast<uint8_t*>(out);
                uint32_t head = static_cast<uint32_t>((16u - (reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u);
                if (head > bytes) head = static_cast<uint32_t>(bytes);
                if (tid < head / W) out[tid] = ld_val<W>(src + tid * W);
                const uint8_t* s2 = src + head;
                uint8_t* d2 = dst + head;
                const uint32_t nvec = static_cast<uint32_t>((bytes - head) >> 4);
                const uint32_t sh = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(s2) & 15u);
                const uint8_t* a = s2 - sh;
                const uint32_t bs = (sh & 3u) * 8u, ws = sh >> 2;
                for (uint32_t j = tid; j < nvec; j += kBigThreads) {
                    const uint4 v0 = ldg_nc16(a + 16u * j);
                    uint4 r = v0;
                    if (sh) {
                        const uint4 v1 = ldg_nc16(a + 16u * j + 16);
                        const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                        r.x = __funnelshift_r(w[ws], w[ws + 1], bs); r.y = __funnelshift_r(w[ws + 1], w[ws + 2], bs);
                        r.z = __funnelshift_r(w[ws + 2], w[ws + 3], bs); r.w = __funnelshift_r(w[ws + 3], w[ws + 4], bs);
                    }
                    __stcs(reinterpret_cast<uint4*>(d2) + j, r);
                }
                const uint32_t done = head + (nvec << 4);
                const uint32_t tail = static_cast<uint32_t>(bytes - done) / W;
                if (tid < tail) reinterpret_cast<T*>(dst + done)[tid] = ld_val<W>(src + done + tid * W);
            } else {
                // ---- PLAIN with nulls: slot -> rank -> value; four slots per thread and trip so that four loads are in
                // flight per lane (one at a time left the kernel waiting on global latency: 28 % of its stall samples) ----
                const uint32_t kmax = nn ? nn - 1u : 0u;
                for (uint32_t s0 = tid; s0 < n; s0 += kEmitUnroll * kBigThreads) {
                    T x[kEmitUnroll];
                    bool v[kEmitUnroll];
#pragma unroll
                    for (uint32_t u = 0; u < kEmitUnroll; u++) {
                        const uint32_t sl = min(s0 + u * kBigThreads, n - 1u);
                        const uint32_t wv = S.pv[sl >> 5];
                        v[u] = nn && ((wv >> (sl & 31u)) & 1u);
                        const uint32_t k = min(S.rb[sl >> 5] + __popc(wv & ((1u << (sl & 31u)) - 1u)), kmax); // (clamped: null slots load and discard)
                        x[u] = nn ? ld_val<W>(src + static_cast<size_t>(k) * W) : T(0);
                    }
#pragma unroll
                    for (uint32_t u = 0; u < kEmitUnroll; u++) {
                        const uint32_t sl = s0 + u * kBigThreads;
                        if (sl < n) __stcs(out + sl, v[u] ? x[u] : T(0));
                    }
                }
            }
        } else {
            // ---- dictionary indices: staged slices, run table by one thread, parallel expansion ----
            if (pos >= size) { if (tid == 0) hand_over(P, q); continue; }
            const uint32_t bw = pg[pos];
            pos++;
            if (bw > 32) { if (tid == 0) hand_over(P, q); continue; }
            const uint8_t* stream = pg + pos;
            const uint32_t slen = size - pos;
            const T* dict = reinterpret_cast<const T*>(P.dict_arena + ck.dict_arena_off);
            const uint32_t dict_n = ck.dict_ok_n;
            const uint32_t imask = bw >= 32u ? 0xffffffffu : ((1u << bw) - 1u);
            uint32_t spos = 0, vdone = 0, slot_lo = 0; // next run header, values decoded, slots emitted
            if (nn == 0) { for (uint32_t sl = tid; sl < n; sl += kBigThreads) __stcs(out + sl, T(0)); }
            bool dense = false;
            while (vdone < nn && ok) {
                // (dense streams are consumed a scan window at a time: stage less)
                const uint32_t take = min(dense ? 2u * static_cast<uint32_t>(kScanWin) : static_cast<uint32_t>(kBigChunk), slen - min(slen, spos));
                const uint8_t* cb = stage_slice(S, stream + spos, take);
                if (!dense) {
                    if (tid == 0) scan_runs(S, cb, take, bw, nn, vdone, spos, slen, kSeqRunBudget);
                    __syncthreads();
                    dense = !S.status && S.n_runs == static_cast<uint32_t>(kSeqRunBudget) && S.next_val < nn && S.next_pos < spos + take;
                    __syncthreads();
                }
                if (dense) scan_runs_parallel(S, cb, take, bw, nn, vdone, spos, slen);
                if (tid == 0 && !S.status) {
                    // slots of this batch: up to (excluding) the slot of value next_val; the last batch takes the rest
                    uint32_t hi_slot = n;
                    if (!all_valid && S.next_val < nn) {
                        uint32_t lo = 0, hi = nwords; // last word with rb[w] <= next_val
                        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (S.rb[mid] <= S.next_val) lo = mid; else hi = mid; }
                        uint32_t need = S.next_val - S.rb[lo], wv = S.pv[lo], b = 0;
                        for (; b < 32; b++) { if ((wv >> b) & 1u) { if (need == 0) break; need--; } }
                        hi_slot = lo * 32u + b;
                    } else if (all_valid) hi_slot = S.next_val;
                    S.hi_slot = hi_slot;
                }
                __syncthreads();
                if (S.status) { ok = false; break; }
                const uint32_t nr = S.n_runs, vend = S.next_val, slot_hi = S.hi_slot;
                // four slots per thread and trip (index extraction first, then the four dictionary gathers in flight together);
                // the lanes of a warp hold consecutive slots, so their values are consecutive too: ONE binary search per warp
                // and step (for its first lane's value, warp-uniform shared loads), every lane walks on from there
                for (uint32_t s0 = slot_lo + tid; s0 - (tid & 31u) < slot_hi; s0 += kEmitUnroll * kBigThreads) {
                    uint32_t ix[kEmitUnroll];
                    bool val[kEmitUnroll];
#pragma unroll
                    for (uint32_t u = 0; u < kEmitUnroll; u++) {
                        const uint32_t sl = s0 + u * kBigThreads;
                        ix[u] = 0; val[u] = false;
                        if (sl - (tid & 31u) >= slot_hi) continue; // warp-uniform
                        const uint32_t slc = min(sl, slot_hi - 1u);
                        uint32_t v = slc;
                        bool valid = sl < slot_hi;
                        if (!all_valid) {
                            const uint32_t wv = S.pv[slc >> 5];
                            valid = valid && ((wv >> (slc & 31u)) & 1u);
                            v = S.rb[slc >> 5] + __popc(wv & ((1u << (slc & 31u)) - 1u));
                        }
                        const uint32_t v0 = __shfl_sync(0xffffffffu, v, 0);
                        uint32_t r = find_run(S, nr, min(v0, vend - 1u));
                        if (valid) {
                            while (v >= S.run_first[r + 1]) r++; // v < vend = run_first[nr]
                            ix[u] = S.run_lit[r] ? (ldbits(cb, S.run_data[r] + (v - S.run_first[r]) * bw, bw) & imask) : S.run_data[r];
                            val[u] = true;
                        }
                    }
                    T x[kEmitUnroll];
#pragma unroll
                    for (uint32_t u = 0; u < kEmitUnroll; u++) {
                        const bool in = val[u] && ix[u] < dict_n;
                        if (val[u] && !in) ok = false; // NULL in the reference: the general kernel redoes the page
                        x[u] = 0;
                        if (in) x[u] = P.identity_dict ? static_cast<T>(ix[u]) : __ldg(dict + ix[u]);
                    }
#pragma unroll
                    for (uint32_t u = 0; u < kEmitUnroll; u++) {
                        const uint32_t sl = s0 + u * kBigThreads;
                        if (sl < slot_hi) __stcs(out + sl, x[u]);
                    }
                }
                ok = __syncthreads_and(ok);
                vdone = vend;
                slot_lo = slot_hi;
                spos = S.next_pos;
            }
            if (!ok) { if (tid == 0) hand_over(P, q); continue; }
        }
        // ---- validity ----
        if (P.validity && ck.max_def > 0) {
            const uint64_t a0 = pd.out_row_base;
            if (all_valid) {
                const uint64_t a1 = a0 + n, w0 = a0 >> 5, w1 = (a1 - 1) >> 5;
                for (uint64_t w = w0 + tid; w <= w1; w += kBigThreads) {
                    uint32_t m = 0xffffffffu;
                    if (w == w0) m &= ~0u << (a0 & 31u);
                    if (w == w1 && (a1 & 31u)) m &= (1u << (a1 & 31u)) - 1u;
                    if (m == 0xffffffffu) P.validity[w] = m; else atomicOr(&P.validity[w], m);
                }
            } else {
                const uint32_t sh = static_cast<uint32_t>(a0 & 31u);
                uint32_t* gv = P.validity + (a0 >> 5);
                for (uint32_t w = tid; w < nwords; w += kBigThreads) {
                    const uint32_t m = S.pv[w];
                    if (!m) continue;
                    atomicOr(&gv[w], m << sh);
                    if (sh && (m >> (32u - sh))) atomicOr(&gv[w + 1], m >> (32u - sh));
                }
            }
        }
    }
}

} // namespace

cudaError_t launch_big_pages(const DecodeParams& p, int width, int sm_count, cudaStream_t s) {
    const uint32_t n = p.slow_hi - p.slow_lo;
    if (n == 0) return cudaSuccess;
    const size_t smem = sizeof(BigSmem);
    const uint32_t grid = n < static_cast<uint32_t>(sm_count) * 4u ? n : static_cast<uint32_t>(sm_count) * 4u;
    cudaError_t e;
    if (width == 4) {
        e = cudaFuncSetAttribute(k_big_pages<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        k_big_pages<4><<<grid, kBigThreads, smem, s>>>(p);
    } else if (width == 8) {
        e = cudaFuncSetAttribute(k_big_pages<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        k_big_pages<8><<<grid, kBigThreads, smem, s>>>(p);
    } else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

} // namespace pqg
