// pqg_ext.cu -- pages the reference does not read: DATA_PAGE_V2 framing and SNAPPY-compressed pages.
//
// The reference refuses compressed chunks (src/reader/column_reader.cpp:13-15) and never decodes a DATA_PAGE_V2
// (:66-67), so there is no parity target in it; pyarrow is the oracle of the tests.  Plans created through
// pqg_plan_create_ext keep every decode kernel as it is: ONE extra launch at the start of a run rewrites each dictionary
// page and data page of the caller's image into the plan's own image in the layout the kernels know --
//   DATA_PAGE    [u32 def_len][def levels][values]                       (max_def > 0; levels absent otherwise)
// -- a warp per page:
//   * uncompressed DATA_PAGE / dictionary page: a copy;
//   * DATA_PAGE_V2 (levels in front of the values WITHOUT a length word, never compressed; lengths in the page header):
//     length word + definition levels + values; a page of an OPTIONAL column that stores no level bytes gets one RLE run
//     "every slot present";
//   * SNAPPY (the whole DATA_PAGE payload, the value section of a DATA_PAGE_V2, a dictionary page): the raw-format
//     decoder below.  Elements are parsed by every lane (the same bytes: broadcast loads), literals and copies are moved
//     by the warp together; an overlapping copy (offset < length: a repeating pattern) reads source byte i mod offset,
//     so it needs no byte-serial loop either.
// Nested columns (repetition levels) stay out: the reference reads definition levels in front of repetition levels
// (DESIGN 8), there is nothing to be compatible with.
#include "pqg_page.cuh"

namespace pqg {
namespace {

// `n` bytes src -> dst by the warp (any alignment; the ranges do not overlap)
__device__ __forceinline__ void warp_copy(uint8_t* dst, const uint8_t* src, uint32_t n) {
    const uint32_t l = lane_id();
    const uint32_t head = min(n, static_cast<uint32_t>((4u - (reinterpret_cast<uintptr_t>(dst) & 3u)) & 3u));
    if (l < head) dst[l] = src[l];
    const uint32_t nw = (n - head) >> 2;
    for (uint32_t w = l; w < nw; w += 32) *reinterpret_cast<uint32_t*>(dst + head + 4u * w) = ld32u(src + head + 4u * w);
    const uint32_t done = head + 4u * nw;
    if (done + l < n) dst[done + l] = src[done + l];
}

// raw SNAPPY block `in[0, n)` -> exactly `expect` bytes at `out`; false: corrupt input (nothing may be assumed about out).
// `hist` = the warp's kSnappyHist bytes of shared memory: a ring of the most recent output (position p at p & (kSnappyHist - 1)).
// A copy element reads what the warp itself wrote a moment ago; from global memory that is an L2 round trip per element
// (the stores went past the L1), from the ring a shared-memory load.  Offsets beyond the ring still go to global memory.
constexpr uint32_t kSnappyHist = 4096;
__device__ __forceinline__ bool warp_snappy(const uint8_t* in, uint32_t n, uint8_t* out, uint32_t expect, uint8_t* hist) {
    const uint32_t l = lane_id();
    uint32_t ip = 0, ulen = 0, shift = 0;
    for (;;) { // preamble: uncompressed length, varint
        if (ip >= n || shift > 28u) return false;
        const uint32_t b = in[ip++];
        ulen |= (b & 0x7fu) << shift;
        if (!(b & 0x80u)) break;
        shift += 7u;
    }
    if (ulen != expect) return false;
    uint32_t op = 0;
    while (ip < n) {
        const uint32_t tag = in[ip++];
        const uint32_t type = tag & 3u;
        if (type == 0u) { // literal
            uint32_t len = (tag >> 2) + 1u;
            if (len > 60u) {
                const uint32_t nb = len - 60u; // 1 .. 4 length bytes
                if (ip + nb > n) return false;
                len = 0;
                for (uint32_t k = 0; k < nb; k++) len |= static_cast<uint32_t>(in[ip + k]) << (8u * k);
                if (len == 0xffffffffu) return false;
                len += 1u;
                ip += nb;
            }
            if (len > n - ip || len > expect - op) return false;
            if (len <= 64u) { // short: bytes by lane, to the output and to the ring
                for (uint32_t i = l; i < len; i += 32) { const uint8_t b = in[ip + i]; out[op + i] = b; hist[(op + i) & (kSnappyHist - 1u)] = b; }
            } else {
                warp_copy(out + op, in + ip, len);
                const uint32_t keep = min(len, kSnappyHist); // the ring only ever needs the last kSnappyHist bytes
                for (uint32_t i = len - keep + l; i < len; i += 32) hist[(op + i) & (kSnappyHist - 1u)] = in[ip + i];
            }
            ip += len;
            op += len;
        } else {
            uint32_t len, off;
            if (type == 1u) {
                if (ip >= n) return false;
                len = 4u + ((tag >> 2) & 7u);
                off = ((tag >> 5) << 8) | in[ip];
                ip += 1u;
            } else if (type == 2u) {
                if (ip + 2u > n) return false;
                len = (tag >> 2) + 1u;
                off = static_cast<uint32_t>(in[ip]) | static_cast<uint32_t>(in[ip + 1]) << 8;
                ip += 2u;
            } else {
                if (ip + 4u > n) return false;
                len = (tag >> 2) + 1u;
                off = static_cast<uint32_t>(in[ip]) | static_cast<uint32_t>(in[ip + 1]) << 8 | static_cast<uint32_t>(in[ip + 2]) << 16 |
                      static_cast<uint32_t>(in[ip + 3]) << 24;
                ip += 4u;
            }
            if (off == 0u || off > op || len > expect - op) return false;
            // (len <= 64: the bytes read all lie in [op - off, op), written before this element; an overlapping copy -- offset <
            //  length, a repeating pattern -- reads source byte i mod offset)
            if (off + 64u <= kSnappyHist) { // the source is still in the ring, and the bytes written now do not alias it
                for (uint32_t i = l; i < len; i += 32) {
                    const uint8_t b = hist[(op - off + (i < off ? i : i % off)) & (kSnappyHist - 1u)];
                    out[op + i] = b;
                    hist[(op + i) & (kSnappyHist - 1u)] = b;
                }
            } else {
                const uint8_t* from = out + op - off;
                for (uint32_t i = l; i < len; i += 32) {
                    const uint8_t b = __ldcg(from + (i < off ? i : i % off));
                    out[op + i] = b;
                    hist[(op + i) & (kSnappyHist - 1u)] = b;
                }
            }
            op += len;
        }
        __syncwarp(); // the element's bytes, written by any lane, are the next elements' source
    }
    return op == expect;
}

__global__ void __launch_bounds__(kThreadsPerCta) k_xform(const uint8_t* src, uint8_t* dst, const XformRec* recs, uint32_t n, DevErr* err) {
    __shared__ __align__(16) uint8_t s_hist[kWarpsPerCta][kSnappyHist];
    const uint32_t l = lane_id();
    const uint32_t nwarps = gridDim.x * kWarpsPerCta;
    for (uint32_t i = blockIdx.x * kWarpsPerCta + warp_id(); i < n; i += nwarps) {
        const XformRec r = recs[i];
        const uint8_t* in = src + r.src_off;
        uint8_t* out = dst + r.dst_off;
        uint32_t in_size = r.src_size, out_size = r.dst_size;
        bool ok = true;
        if (r.kind & kXformV2) {
            // levels: repetition levels (flat columns: none), then definition levels, both stored as they are
            if (static_cast<uint64_t>(r.rep_len) + r.def_len > in_size) ok = false;
            else {
                uint32_t lev = 0;
                if (r.kind & kXformPrefix) {
                    if (r.def_len) {
                        if (l == 0) { out[0] = r.def_len & 0xffu; out[1] = (r.def_len >> 8) & 0xffu; out[2] = (r.def_len >> 16) & 0xffu; out[3] = r.def_len >> 24; }
                        warp_copy(out + 4, in + r.rep_len, r.def_len);
                        lev = 4u + r.def_len;
                    } else { // no level bytes stored: one RLE run "every slot present" (its count as a 5-byte varint: fixed size)
                        if (l == 0) {
                            const uint32_t v = r.num_values << 1;
                            out[0] = static_cast<uint8_t>(kXformSynthBytes - 4u); out[1] = 0; out[2] = 0; out[3] = 0;
                            for (uint32_t j = 0; j < 4u; j++) out[4 + j] = static_cast<uint8_t>(((v >> (7u * j)) & 0x7fu) | 0x80u);
                            out[8] = static_cast<uint8_t>((v >> 28) & 0x7fu);
                            out[9] = 1;
                        }
                        lev = kXformSynthBytes;
                    }
                }
                in += r.rep_len + r.def_len;
                in_size -= r.rep_len + r.def_len;
                if (lev > out_size) ok = false;
                else { out += lev; out_size -= lev; }
            }
        }
        if (ok) {
            const uint32_t codec = (r.kind >> 8) & 0xffu;
            if (codec == PQG_CODEC_UNCOMPRESSED) {
                if (in_size != out_size) ok = false; else warp_copy(out, in, in_size);
            } else if (codec == PQG_CODEC_SNAPPY) ok = warp_snappy(in, in_size, out, out_size, s_hist[warp_id()]);
            else ok = false;
        }
        if (!ok && l == 0) report_error(err, r.page, PQG_PAGE_DECOMPRESS);
        __syncwarp();
    }
}

} // namespace

cudaError_t launch_xform(const uint8_t* src, uint8_t* dst, const XformRec* recs, uint32_t n, DevErr* err, int sm_count, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const uint32_t want = (n + kWarpsPerCta - 1) / kWarpsPerCta, cap = static_cast<uint32_t>(sm_count) * 8u;
    k_xform<<<want < cap ? want : cap, kThreadsPerCta, 0, s>>>(src, dst, recs, n, err);
    return cudaGetLastError();
}

} // namespace pqg
