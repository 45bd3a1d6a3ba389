// pqg_api.cu -- implementation of the C-ABI in include/pqg.h: contexts, device images,
// decode plans.  No CPU fallback: every compute entry point needs a CUDA device.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "pqg_internal.h"

using namespace pqg;

struct pqg_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t h2d = nullptr, d2h = nullptr; // copy-in / copy-out streams of the pipelined path (lazy)
    int sm_count = 148;
    bool profiling = false;
    uint64_t launches = 0;
    std::string err;
};

struct pqg_buf {
    uint8_t* d = nullptr;
    uint64_t size = 0;
    uint64_t capacity = 0; // readable bytes behind d (>= size + kImagePad)
    bool owned = false;
};

struct pqg_plan {
    const pqg_buf* image = nullptr;
    std::vector<pqg_chunk_desc> chunks;
    std::vector<pqg_page_desc> pages;
    std::vector<pqg_page_desc> virt_pages;   // device page table = pages + virt_pages (slices of oversized PLAIN pages)
    std::vector<uint32_t> virt_parent;       // page-table index each virtual page was cut from
    DevChunk* d_chunks = nullptr;
    pqg_page_desc* d_pages = nullptr;
    uint8_t* d_dict = nullptr;
    DictSeg* d_dict_segs = nullptr;  // BYTE_ARRAY plans: segment records of the dictionary preparation
    uint32_t str_dict_blocks = 1;    // ... and its blocks of 256 threads per dictionary
    size_t dict_bytes = 0;
    TileDesc* d_tiles = nullptr;      // fast path: TMA-staged page tiles (fixed-width plans)
    uint32_t n_tiles = 0;
    uint32_t* d_slow_pages = nullptr; // pages for the general kernel: host-listed, then device-appended
    uint32_t n_slow_host = 0;
    // flat decode of the slow list (pqg_flat.cu): fixed-width 4/8-byte plans with OPTIONAL chunks or host-listed pages
    bool flat_on = false;
    FlatPage* d_flat_pages = nullptr;
    FlatBlk* d_flat_blk = nullptr;
    uint2* d_flat_ckpt = nullptr;
    uint32_t* d_flat_append = nullptr;
    uint32_t flat_blk_cap = 0;
    // pqg_plan_create_ext: the plan's own image (`image` points at it) is rebuilt from the caller's at the start of every run
    const pqg_buf* ext_src = nullptr;
    pqg_buf* ext_image = nullptr;
    XformRec* d_xform = nullptr;
    uint32_t n_xform = 0;
    uint32_t max_dict_n = 0;                 // most entries of one dictionary of the plan
    bool flat_ran = false;                   // the current run launched the flat kernels
    uint32_t tile_handover_seen = 0xffffffffu; // pages the tile kernel handed over in the last finished run (~0u: none finished yet)
    uint32_t dict_smem = 0;
    uint32_t max_dict_blocks = 1;
    uint32_t tile_bytes = kTileBytes;
    // BYTE_ARRAY plans whose pages are all PLAIN REQUIRED: byte counts from the page headers
    // (payload - 4 * values), scans on the host, no size pass and no mid-run synchronisation;
    // the copy pass verifies every page and the plan falls back to the size pass on a mismatch
    bool host_sizes = false, force_exact = false;
    bool chars_sized = false; // a finished run sized d_chars: later runs launch the copy pass without waiting for the size pass
    bool identity = false; // dictionary-form output of a BYTE_ARRAY column: uint32 dictionary indices per slot
    std::vector<uint32_t> chunk_tile_begin;  // n_chunks + 1: tiles of chunk c = [begin[c], begin[c+1])
    std::vector<uint32_t> chunk_slow_begin;  // n_chunks + 1: host-listed slow pages of chunk c
    std::vector<cudaEvent_t> pipe_ev;        // pipelined path: 2 events per chunk (H2D done, decode done)
    cudaEvent_t ev_idle = nullptr;           // last decode of the previous run (guards the image buffer)
    bool pipelined_in_flight = false;
    bool any_dict = false, any_def = false, is_str = false, is_bool = false;
    bool forced_validity = false; // a REQUIRED chunk held an out-of-range dictionary index (null in the reference): validity added, plan re-run
    bool run_pending = false;     // pqg_plan_run / run_pipelined enqueued, pqg_plan_finish not called yet
    bool regex_tile_sync = true;  // regex scan: CTA barrier per tile (pqg_plan_set_option)
    bool opt_idx = false;         // OPTIONAL fixed-width plan with foreign-looking pages (see DecodeParams::opt_idx)
    bool no_part = true;          // partitioned-dictionary mode: measured slower than the L2 gather (0.60 vs 0.49 ms per 100 M values): opt-in (pqg_plan_set_option)
    int phys = 0, width = 0;
    uint64_t n_slots = 0;
    uint32_t tile_launches = 0;              // tile-kernel launches of the current run
    uint32_t handover_seen = 0xffffffffu;    // pages handed to the general kernel in the last finished run (~0u: none finished yet)
    mutable uint32_t max_page_values = 0xffffffffu; // most num_values of one page (computed on first use)
    uint8_t* d_values = nullptr;
    uint32_t* d_validity = nullptr;
    uint64_t* d_required_ranges = nullptr;   // plans with validity: slot ranges of the REQUIRED chunks
    uint32_t n_required_ranges = 0;
    uint32_t* d_offsets = nullptr;
    uint8_t* d_chars = nullptr;
    uint64_t chars_cap = 0, chars_size = 0;
    uint32_t* d_page_chars = nullptr;
    uint32_t* d_page_char_base = nullptr;
    uint64_t* d_bases = nullptr;  // n_chunks + 1 bases, then the grand total
    uint64_t* h_bases = nullptr;  // pinned mirror
    DevErr* d_err = nullptr;
    DevErr* h_err = nullptr;      // pinned
    static constexpr int kTimingSlots = 8;   // ring: the last 8 profiled runs keep their events
    cudaEvent_t evr[kTimingSlots][5] = {};
    cudaEvent_t* ev = evr[0];                // events of the current run
    uint64_t runs_timed = 0;
    bool timed = false;
    pqg_timings tm{};
    uint32_t last_launches = 0;
    uint64_t bytes_in = 0, bytes_out = 0;
    bool ran = false;
};

struct pqg_dfa; // pqg_regex_host.cpp

static std::string g_create_err;

static int fail(pqg_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg; else g_create_err = msg;
    return code;
}
static int cuda_fail(pqg_ctx* ctx, cudaError_t e, const char* what) {
    return fail(ctx, PQG_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(ctx, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(ctx, e_, #call); } while (0)

extern "C" {

int pqg_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int pqg_ctx_create(int device, void* stream, pqg_ctx** out) {
    if (!out) return fail(nullptr, PQG_ERR_ARG, "pqg_ctx_create: out is NULL");
    *out = nullptr;
    int n = pqg_device_count();
    if (n <= 0) return fail(nullptr, PQG_ERR_CUDA, "pqg_ctx_create: no CUDA device is available (this library has no CPU fallback)");
    if (device < 0 || device >= n) return fail(nullptr, PQG_ERR_ARG, "pqg_ctx_create: bad device index");
    pqg_ctx* c = new (std::nothrow) pqg_ctx();
    if (!c) return fail(nullptr, PQG_ERR_NOMEM, "out of memory");
    c->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { delete c; return cuda_fail(nullptr, e, "cudaSetDevice"); }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { delete c; return cuda_fail(nullptr, e, "cudaGetDeviceProperties"); }
    if (prop.major < 10) {
        delete c;
        return fail(nullptr, PQG_ERR_CUDA, "pqg_ctx_create: this build targets sm_100a (Blackwell B200) only");
    }
    c->sm_count = prop.multiProcessorCount;
    { // scratch of the scan / chunk-index calls comes from the stream-ordered pool: keep freed blocks cached
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    if (stream) { c->stream = static_cast<cudaStream_t>(stream); c->own_stream = false; }
    else {
        e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete c; return cuda_fail(nullptr, e, "cudaStreamCreate"); }
        c->own_stream = true;
    }
    *out = c;
    return PQG_OK;
}

void pqg_ctx_destroy(pqg_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->h2d) cudaStreamDestroy(ctx->h2d);
    if (ctx->d2h) cudaStreamDestroy(ctx->d2h);
    delete ctx;
}

const char* pqg_last_error(const pqg_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int pqg_ctx_sync(pqg_ctx* ctx) {
    if (!ctx) return PQG_ERR_ARG;
    CU(ctx, cudaSetDevice(ctx->device));
    if (ctx->h2d) CU(ctx, cudaStreamSynchronize(ctx->h2d));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->d2h) CU(ctx, cudaStreamSynchronize(ctx->d2h));
    return PQG_OK;
}

int pqg_ctx_set_profiling(pqg_ctx* ctx, int on) {
    if (!ctx) return PQG_ERR_ARG;
    ctx->profiling = on != 0;
    return PQG_OK;
}

uint64_t pqg_kernel_launches(const pqg_ctx* ctx) { return ctx ? ctx->launches : 0; }

void* pqg_host_alloc(uint64_t size) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, size ? size : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void pqg_host_free(void* p) { if (p) cudaFreeHost(p); }

int pqg_upload(pqg_ctx* ctx, const void* host_bytes, uint64_t size, pqg_buf** out) {
    if (!ctx || !out || (!host_bytes && size)) return fail(ctx, PQG_ERR_ARG, "pqg_upload: bad argument");
    CU(ctx, cudaSetDevice(ctx->device));
    pqg_buf* b = new (std::nothrow) pqg_buf();
    if (!b) return fail(ctx, PQG_ERR_NOMEM, "out of memory");
    cudaError_t e = cudaMalloc(&b->d, size + kImagePad);
    if (e != cudaSuccess) { delete b; return cuda_fail(ctx, e, "cudaMalloc(image)"); }
    b->size = size;
    b->capacity = size + kImagePad;
    b->owned = true;
    e = cudaMemsetAsync(b->d + size, 0, kImagePad, ctx->stream);
    if (e == cudaSuccess && size) e = cudaMemcpyAsync(b->d, host_bytes, size, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) { cudaFree(b->d); delete b; return cuda_fail(ctx, e, "cudaMemcpyAsync(image)"); }
    *out = b;
    return PQG_OK;
}

int pqg_buf_alloc(pqg_ctx* ctx, uint64_t size, pqg_buf** out) {
    if (!ctx || !out) return fail(ctx, PQG_ERR_ARG, "pqg_buf_alloc: bad argument");
    CU(ctx, cudaSetDevice(ctx->device));
    pqg_buf* b = new (std::nothrow) pqg_buf();
    if (!b) return fail(ctx, PQG_ERR_NOMEM, "out of memory");
    cudaError_t e = cudaMalloc(&b->d, size + kImagePad);
    if (e != cudaSuccess) { delete b; return cuda_fail(ctx, e, "cudaMalloc(image)"); }
    b->size = size;
    b->capacity = size + kImagePad;
    b->owned = true;
    e = cudaMemsetAsync(b->d + size, 0, kImagePad, ctx->stream);
    if (e != cudaSuccess) { cudaFree(b->d); delete b; return cuda_fail(ctx, e, "cudaMemsetAsync(image pad)"); }
    *out = b;
    return PQG_OK;
}

int pqg_buf_write(pqg_ctx* ctx, pqg_buf* buf, uint64_t dst_off, const void* host_bytes, uint64_t n) {
    if (!ctx || !buf || (!host_bytes && n)) return fail(ctx, PQG_ERR_ARG, "pqg_buf_write: bad argument");
    if (dst_off + n > buf->size) return fail(ctx, PQG_ERR_ARG, "pqg_buf_write: range outside the image");
    CU(ctx, cudaSetDevice(ctx->device));
    if (n) CU(ctx, cudaMemcpyAsync(buf->d + dst_off, host_bytes, n, cudaMemcpyHostToDevice, ctx->stream));
    return PQG_OK;
}

uint64_t pqg_buf_size(const pqg_buf* buf) { return buf ? buf->size : 0; }

int pqg_wrap_device(pqg_ctx* ctx, const void* dev_ptr, uint64_t size, uint64_t capacity, pqg_buf** out) {
    if (!ctx || !out || !dev_ptr) return fail(ctx, PQG_ERR_ARG, "pqg_wrap_device: bad argument");
    if (reinterpret_cast<uintptr_t>(dev_ptr) & 15u) return fail(ctx, PQG_ERR_ARG, "pqg_wrap_device: pointer must be 16-byte aligned");
    // the kernels read 16-byte vectors / whole tiles: up to kImagePad bytes behind the image must be readable
    if (capacity < size + kImagePad) return fail(ctx, PQG_ERR_ARG, "pqg_wrap_device: the allocation must be readable for 64 bytes past `size` (capacity >= size + 64)");
    pqg_buf* b = new (std::nothrow) pqg_buf();
    if (!b) return fail(ctx, PQG_ERR_NOMEM, "out of memory");
    b->d = const_cast<uint8_t*>(static_cast<const uint8_t*>(dev_ptr));
    b->size = size;
    b->capacity = capacity;
    b->owned = false;
    *out = b;
    return PQG_OK;
}

void pqg_buf_free(pqg_ctx* ctx, pqg_buf* buf) {
    if (!buf) return;
    if (ctx) cudaSetDevice(ctx->device);
    if (buf->owned && buf->d) cudaFree(buf->d);
    delete buf;
}

const void* pqg_buf_device_ptr(const pqg_buf* buf) { return buf ? buf->d : nullptr; }

// ---------------------------------------------------------------------------------------------

static int type_width(int phys) {
    switch (phys) {
        case PQG_BOOLEAN: return 1;
        case PQG_INT32: case PQG_FLOAT: return 4;
        case PQG_INT64: case PQG_DOUBLE: return 8;
        case PQG_INT96: return 12;
        default: return 0;
    }
}
static uint8_t level_bw(int16_t m) { // ColumnReader::bit_width (column_reader.cpp:270-276)
    uint8_t bw = 0;
    int v = m;
    while (v > 0) { bw++; v >>= 1; }
    return bw;
}

void pqg_plan_destroy(pqg_ctx* ctx, pqg_plan* p) {
    if (!p) return;
    if (ctx) cudaSetDevice(ctx->device);
    cudaFree(p->d_chunks); cudaFree(p->d_pages); cudaFree(p->d_dict); cudaFree(p->d_dict_segs); cudaFree(p->d_values);
    cudaFree(p->d_validity); cudaFree(p->d_required_ranges); cudaFree(p->d_offsets); cudaFree(p->d_chars); cudaFree(p->d_page_chars);
    cudaFree(p->d_page_char_base); cudaFree(p->d_bases); cudaFree(p->d_err); cudaFree(p->d_tiles); cudaFree(p->d_slow_pages);
    cudaFree(p->d_xform); if (p->ext_image) { cudaFree(p->ext_image->d); delete p->ext_image; }
    cudaFree(p->d_flat_pages); cudaFree(p->d_flat_blk); cudaFree(p->d_flat_ckpt); cudaFree(p->d_flat_append);
    if (p->h_bases) cudaFreeHost(p->h_bases);
    if (p->h_err) cudaFreeHost(p->h_err);
    for (auto& slot : p->evr) for (auto& e : slot) if (e) cudaEventDestroy(e);
    for (auto& e : p->pipe_ev) if (e) cudaEventDestroy(e);
    if (p->ev_idle) cudaEventDestroy(p->ev_idle);
    delete p;
}

// Plans that mix REQUIRED and OPTIONAL chunks (several columns in one plan): the slots of the
// REQUIRED chunks read "valid" in the plan's validity bitmap.  ranges: [2 * i] = first slot,
// [2 * i + 1] = one past the last slot; one CTA per range.
__global__ void k_validity_ranges(uint32_t* validity, const uint64_t* ranges) {
    const uint64_t a0 = ranges[2 * blockIdx.x], a1 = ranges[2 * blockIdx.x + 1];
    if (a1 <= a0) return;
    const uint64_t w0 = a0 >> 5, w1 = (a1 - 1) >> 5;
    for (uint64_t w = w0 + threadIdx.x; w <= w1; w += blockDim.x) {
        uint32_t m = 0xffffffffu;
        if (w == w0) m &= ~0u << (a0 & 31u);
        if (w == w1 && (a1 & 31u)) m &= (1u << (a1 & 31u)) - 1u;
        if (m == 0xffffffffu) validity[w] = m; else atomicOr(&validity[w], m);
    }
}
static cudaError_t reset_validity(pqg_plan* p, cudaStream_t s);

static int plan_create_impl(pqg_ctx* ctx, const pqg_buf* image, const pqg_chunk_desc* chunks, uint32_t n_chunks,
                            const pqg_page_desc* pages, uint32_t n_pages, bool dict_indices, pqg_plan** out);

int pqg_plan_create(pqg_ctx* ctx, const pqg_buf* image, const pqg_chunk_desc* chunks, uint32_t n_chunks,
                    const pqg_page_desc* pages, uint32_t n_pages, pqg_plan** out) {
    return plan_create_impl(ctx, image, chunks, n_chunks, pages, n_pages, false, out);
}

int pqg_plan_create_dict_indices(pqg_ctx* ctx, const pqg_buf* image, const pqg_chunk_desc* chunks, uint32_t n_chunks,
                                 const pqg_page_desc* pages, uint32_t n_pages, pqg_plan** out) {
    return plan_create_impl(ctx, image, chunks, n_chunks, pages, n_pages, true, out);
}

int pqg_plan_create_ext(pqg_ctx* ctx, const pqg_buf* image, const pqg_chunk_desc* chunks, uint32_t n_chunks,
                        const pqg_page_desc* pages, uint32_t n_pages, const pqg_page_ext* page_ext,
                        const pqg_chunk_ext* chunk_ext, pqg_plan** out) {
    if (!ctx || !image || !out || (!chunks && n_chunks) || (!pages && n_pages) || (!page_ext && n_pages) || (!chunk_ext && n_chunks))
        return fail(ctx, PQG_ERR_ARG, "pqg_plan_create_ext: bad argument");
    // the layout of the plan's own image: per chunk the dictionary page, then its data pages, each as a DATA_PAGE payload
    std::vector<pqg_chunk_desc> ck(chunks, chunks + n_chunks);
    std::vector<pqg_page_desc> pg(pages, pages + n_pages);
    std::vector<XformRec> recs;
    recs.reserve(static_cast<size_t>(n_pages) + n_chunks);
    uint64_t off = 0;
    auto place = [&](uint64_t bytes) { const uint64_t at = off; off = (off + bytes + 15u) & ~uint64_t(15); return at; };
    for (uint32_t c = 0; c < n_chunks; c++) {
        const pqg_chunk_desc& s = chunks[c];
        if (s.first_page > n_pages || s.n_pages > n_pages - s.first_page) return fail(ctx, PQG_ERR_ARG, "pqg_plan_create_ext: chunk page range outside the page table");
        if (s.has_dict) {
            const uint32_t codec = chunk_ext[c].dict_codec;
            if (codec > PQG_CODEC_SNAPPY) return fail(ctx, PQG_ERR_UNSUPPORTED, "pqg_plan_create_ext: dictionary page codec " + std::to_string(codec) + " is not supported (SNAPPY only)");
            if (s.dict_off + s.dict_size > image->size) return fail(ctx, PQG_ERR_ARG, "pqg_plan_create_ext: dictionary page outside the image");
            const uint32_t usz = codec ? chunk_ext[c].dict_uncompressed_size : static_cast<uint32_t>(s.dict_size);
            XformRec r{};
            r.src_off = s.dict_off; r.src_size = static_cast<uint32_t>(s.dict_size);
            r.dst_off = place(usz); r.dst_size = usz;
            r.kind = codec << 8; r.page = s.first_page;
            recs.push_back(r);
            ck[c].dict_off = r.dst_off; ck[c].dict_size = usz;
        }
        for (uint32_t q = s.first_page; q < s.first_page + s.n_pages; q++) {
            const pqg_page_desc& p = pages[q];
            const pqg_page_ext& e = page_ext[q];
            const uint32_t codec = (e.kind >> 8) & 0xffu;
            const bool v2 = e.kind & PQG_PAGE_EXT_V2;
            if (codec > PQG_CODEC_SNAPPY) return fail(ctx, PQG_ERR_UNSUPPORTED, "pqg_plan_create_ext: page codec " + std::to_string(codec) + " is not supported (SNAPPY only)");
            if (p.payload_off + p.payload_size > image->size) return fail(ctx, PQG_ERR_ARG, "pqg_plan_create_ext: page outside the image");
            if (p.flags & PQG_PAGE_FLAG_V2) return fail(ctx, PQG_ERR_ARG, "pqg_plan_create_ext: DATA_PAGE_V2 pages are marked in pqg_page_ext.kind, not in flags");
            XformRec r{};
            r.src_off = p.payload_off; r.src_size = p.payload_size;
            r.kind = codec << 8; r.page = q; r.num_values = p.num_values;
            uint32_t usz;
            if (v2) {
                if (s.max_rep > 0 || e.rep_len) return fail(ctx, PQG_ERR_UNSUPPORTED, "pqg_plan_create_ext: DATA_PAGE_V2 of a nested column (repetition levels) is not supported");
                if (static_cast<uint64_t>(e.def_len) + e.rep_len > e.uncompressed_size || static_cast<uint64_t>(e.def_len) + e.rep_len > p.payload_size)
                    return fail(ctx, PQG_ERR_ARG, "pqg_plan_create_ext: level lengths exceed the page");
                const uint32_t values = e.uncompressed_size - e.def_len - e.rep_len; // uncompressed size of the value section
                uint32_t lev = 0;
                r.kind |= kXformV2;
                if (s.max_def > 0) { r.kind |= kXformPrefix; lev = e.def_len ? 4u + e.def_len : kXformSynthBytes; }
                else if (e.def_len) return fail(ctx, PQG_ERR_ARG, "pqg_plan_create_ext: definition levels in a page of a REQUIRED column");
                r.def_len = e.def_len; r.rep_len = e.rep_len;
                usz = lev + values;
            } else usz = codec ? e.uncompressed_size : p.payload_size;
            r.dst_off = place(usz); r.dst_size = usz;
            recs.push_back(r);
            pg[q].payload_off = r.dst_off; pg[q].payload_size = usz;
        }
    }
    pqg_buf* own = new pqg_buf();
    own->size = off; own->capacity = off + kImagePad; own->owned = true;
    if (cudaSetDevice(ctx->device) != cudaSuccess || cudaMalloc(reinterpret_cast<void**>(&own->d), own->capacity) != cudaSuccess) {
        delete own;
        return cuda_fail(ctx, cudaGetLastError(), "cudaMalloc(plan image)");
    }
    cudaMemsetAsync(own->d + off, 0, kImagePad, ctx->stream);
    pqg_plan* p = nullptr;
    const int rc = plan_create_impl(ctx, own, ck.data(), n_chunks, pg.data(), n_pages, false, &p);
    if (rc != PQG_OK) { cudaFree(own->d); delete own; return rc; }
    p->ext_src = image; p->ext_image = own;
    p->n_xform = static_cast<uint32_t>(recs.size());
    if (!recs.empty()) {
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p->d_xform), recs.size() * sizeof(XformRec));
        if (e == cudaSuccess) e = cudaMemcpy(p->d_xform, recs.data(), recs.size() * sizeof(XformRec), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { pqg_plan_destroy(ctx, p); return cuda_fail(ctx, e, "cudaMalloc(plan transform table)"); }
    }
    *out = p;
    return PQG_OK;
}

static int plan_create_impl(pqg_ctx* ctx, const pqg_buf* image, const pqg_chunk_desc* chunks, uint32_t n_chunks,
                            const pqg_page_desc* pages, uint32_t n_pages, bool dict_indices, pqg_plan** out) {
    if (!ctx || !image || !out || (!chunks && n_chunks) || (!pages && n_pages))
        return fail(ctx, PQG_ERR_ARG, "pqg_plan_create: bad argument");
    *out = nullptr;
    if (n_chunks == 0) return fail(ctx, PQG_ERR_ARG, "pqg_plan_create: no chunks");
    CU(ctx, cudaSetDevice(ctx->device));
    pqg_plan* p = new (std::nothrow) pqg_plan();
    if (!p) return fail(ctx, PQG_ERR_NOMEM, "out of memory");
    p->image = image;
    p->chunks.assign(chunks, chunks + n_chunks);
    p->pages.assign(pages, pages + n_pages);
    p->phys = chunks[0].phys_type;
    auto bail = [&](int code, const std::string& m) { pqg_plan_destroy(ctx, p); return fail(ctx, code, m); };
    if (p->phys == PQG_FIXED_LEN_BYTE_ARRAY)
        return bail(PQG_ERR_UNSUPPORTED, "FIXED_LEN_BYTE_ARRAY not supported without type_length");
    if (p->phys < 0 || p->phys > PQG_FIXED_LEN_BYTE_ARRAY)
        return bail(PQG_ERR_UNSUPPORTED, "Unsupported type: " + std::to_string(p->phys));
    p->is_str = p->phys == PQG_BYTE_ARRAY;
    p->is_bool = p->phys == PQG_BOOLEAN;
    p->width = type_width(p->phys);
    if (dict_indices) {
        // dictionary-form output (Arrow DictionaryArray style): the column must be dictionary-encoded
        // throughout; the plan then is a 4-byte fixed-width plan whose "dictionary" is the identity
        if (!p->is_str) return bail(PQG_ERR_ARG, "pqg_plan_create_dict_indices: BYTE_ARRAY columns only");
        for (uint32_t c = 0; c < n_chunks; c++) {
            if (chunks[c].n_pages && !chunks[c].has_dict) return bail(PQG_ERR_UNSUPPORTED, "column chunk has no dictionary page: not dictionary-encoded throughout");
            for (uint32_t q = chunks[c].first_page; q < chunks[c].first_page + chunks[c].n_pages && q < n_pages; q++)
                if (!(pages[q].flags & PQG_PAGE_FLAG_DICT)) return bail(PQG_ERR_UNSUPPORTED, "column falls back to PLAIN pages: not dictionary-encoded throughout");
        }
        p->is_str = false;
        p->identity = true;
        p->width = 4;
    }

    std::vector<DevChunk> dc(n_chunks);
    size_t arena = 0, n_segs = 0;
    uint32_t max_segs = 0;
    uint64_t slots = 0;
    for (uint32_t c = 0; c < n_chunks; c++) {
        const pqg_chunk_desc& s = chunks[c];
        // values are moved as raw bits: chunks of several columns may share a plan when their value width
        // is the same (INT32/FLOAT, INT64/DOUBLE); BYTE_ARRAY, BOOLEAN and INT96 plans stay single-type
        if (s.phys_type != p->phys && (p->is_str || p->is_bool || p->width == 12 || type_width(s.phys_type) != p->width ||
                                       s.phys_type == PQG_BOOLEAN || s.phys_type == PQG_INT96))
            return bail(PQG_ERR_ARG, "pqg_plan_create: chunks of one plan must share a physical type (or a 4- / 8-byte value width)");
        if (static_cast<uint64_t>(s.first_page) + s.n_pages > n_pages) return bail(PQG_ERR_ARG, "pqg_plan_create: chunk page range out of bounds");
        DevChunk& d = dc[c];
        std::memset(&d, 0, sizeof(d));
        d.dict_off = s.dict_off; d.out_row_base = s.out_row_base; d.num_values = s.num_values;
        d.dict_size = s.dict_size; d.dict_n = s.dict_num_values; d.first_page = s.first_page; d.n_pages = s.n_pages;
        d.max_def = s.max_def; d.max_rep = s.max_rep; d.phys_type = s.phys_type; d.has_dict = s.has_dict;
        d.def_bw = level_bw(s.max_def); d.rep_bw = level_bw(s.max_rep);
        if (s.has_dict) {
            if (s.dict_off + s.dict_size > image->size) return bail(PQG_ERR_ARG, "pqg_plan_create: dictionary page outside the image");
            p->bytes_in += s.dict_size;
            if (p->identity) {
                d.dict_ok_n = s.dict_num_values; // the indices ARE the output: no prepared dictionary on the device
            } else {
                p->any_dict = true;
                d.dict_arena_off = arena;
                size_t ent = p->is_str ? 8 : static_cast<size_t>(p->width);
                arena += (static_cast<size_t>(s.dict_num_values) * ent + 31) & ~size_t(15);
                if (p->is_str) { // 16-byte padded entries for short-string dictionaries (filled by the prepare kernel)
                    d.dict_pad_off = arena;
                    arena += static_cast<size_t>(s.dict_num_values) * 16 + 16;
                    const uint32_t segs = (s.dict_size + kDictSeg - 1) / kDictSeg + 1;
                    if (n_segs + segs > 0xffffffffull) return bail(PQG_ERR_UNSUPPORTED, "pqg_plan_create: more than 512 GB of dictionary pages in one plan");
                    d.dict_seg_first = static_cast<uint32_t>(n_segs);
                    n_segs += segs;
                    max_segs = std::max(max_segs, segs);
                }
            }
        }
        if (s.max_def > 0) p->any_def = true;
        for (uint32_t q = s.first_page; q < s.first_page + s.n_pages; q++) {
            if (pages[q].chunk_idx != c) return bail(PQG_ERR_ARG, "pqg_plan_create: pages of a chunk must be contiguous");
            if (pages[q].flags & PQG_PAGE_FLAG_V2)
                return bail(PQG_ERR_UNSUPPORTED, "DATA_PAGE_V2 pages are not supported (row group " + std::to_string(s.row_group) + ", column " + std::to_string(s.column) + ")");
            const uint32_t enc = PQG_PAGE_ENCODING(pages[q].flags);
            if (enc != 0u && enc != 2u && enc != 8u) {
                static const char* names[] = {"PLAIN", "GROUP_VAR_INT", "PLAIN_DICTIONARY", "RLE", "BIT_PACKED", "DELTA_BINARY_PACKED",
                                              "DELTA_LENGTH_BYTE_ARRAY", "DELTA_BYTE_ARRAY", "RLE_DICTIONARY", "BYTE_STREAM_SPLIT"};
                return bail(PQG_ERR_UNSUPPORTED, std::string("data page encoding ") + (enc < 10u ? names[enc] : std::to_string(enc).c_str()) +
                                                     " is not supported (row group " + std::to_string(s.row_group) + ", column " + std::to_string(s.column) + ")");
            }
            if (pages[q].payload_off + pages[q].payload_size > image->size) return bail(PQG_ERR_ARG, "pqg_plan_create: page outside the image");
            p->bytes_in += pages[q].payload_size;
        }
        slots = std::max<uint64_t>(slots, s.out_row_base + s.num_values);
    }
    p->n_slots = slots;
    p->dict_bytes = arena;

    // fixed-width plans: cut tileable chunks into TMA tiles, list everything else for the
    // general kernel
    std::vector<TileDesc> tiles;
    std::vector<uint32_t> slow;
    {
        // OPTIONAL fixed-width pages carry ~0.9 level bytes per slot: larger tiles keep 8 pages (= 8 warps) per tile
        // (strings: the tiles feed the regex scan, one page per warp -- pages whose stride in the file is past 1 KB
        // would fill an 8 KB tile with seven or fewer, so their plans get the 10 KB or 16 KB tiles)
        uint64_t str_span = 0, str_pages = 0;
        if (p->is_str) for (uint32_t c = 0; c < n_chunks; c++) {
            const pqg_chunk_desc& s = chunks[c];
            if (s.n_pages == 0) continue;
            const pqg_page_desc& a = pages[s.first_page]; const pqg_page_desc& b = pages[s.first_page + s.n_pages - 1];
            if (b.payload_off + b.payload_size > a.payload_off) { str_span += b.payload_off + b.payload_size - a.payload_off; str_pages += s.n_pages; }
        }
        const uint64_t str_need = str_pages ? str_span * kTilePages / str_pages + 64 : 0; // eight average pages
        p->tile_bytes = p->is_str ? (str_need <= kTileBytes ? kTileBytes : str_need <= kTileBytesMid ? kTileBytesMid : kTileBytesLarge)
                                  : (p->any_def ? kTileBytesLarge : kTileBytes);
        const uint64_t tile_cap = p->tile_bytes;
        uint32_t max_dict_n = 0;
        const bool flat_ok = !p->is_str && !p->is_bool && (p->width == 4 || p->width == 8);
        for (uint32_t c = 0; c < n_chunks; c++) {
            const pqg_chunk_desc& s = chunks[c];
            p->chunk_tile_begin.push_back(static_cast<uint32_t>(tiles.size()));
            p->chunk_slow_begin.push_back(static_cast<uint32_t>(slow.size()));
            if (s.has_dict) max_dict_n = std::max(max_dict_n, s.dict_num_values);
            // strings: every flat chunk is tiled (the regex scan runs on tiles; the page kernels
            // classify pages themselves); fixed width: REQUIRED 4/8-byte chunks only
            const bool tileable = p->is_str ? s.max_rep <= 0
                                 : (p->identity ? (s.max_def <= 1 && s.max_rep <= 0) : chunk_is_tileable(s.phys_type, s.max_def, s.max_rep));
            if (tileable && s.has_dict && !p->is_str && !p->identity) {
                uint64_t db = (static_cast<uint64_t>(s.dict_num_values) * p->width + 15) & ~uint64_t(15);
                if (db <= static_cast<uint64_t>(kMaxSmemDictBytes)) p->dict_smem = std::max<uint32_t>(p->dict_smem, static_cast<uint32_t>(db));
            }
            TileDesc cur{};
            uint64_t cur_end = 0;
            auto flush = [&]() { if (cur.n_pages) { cur.byte_len = static_cast<uint32_t>(((cur_end + 15) & ~uint64_t(15)) - cur.byte_lo); tiles.push_back(cur); cur = TileDesc{}; } };
            auto place = [&](uint32_t q, const pqg_page_desc& pg) { // q: device page-table index
                const uint64_t lo = pg.payload_off & ~uint64_t(15), end = pg.payload_off + pg.payload_size;
                if (cur.n_pages) {
                    const bool fits = cur.n_pages < static_cast<uint32_t>(kTilePages) && pg.payload_off >= cur.byte_lo &&
                                      ((end + 15) & ~uint64_t(15)) - cur.byte_lo <= tile_cap && q == cur.first_page + cur.n_pages;
                    if (!fits) flush();
                }
                if (!cur.n_pages) { cur.byte_lo = lo; cur.first_page = q; cur.chunk_idx = c; cur_end = end; }
                cur.n_pages++;
                cur_end = std::max(cur_end, end);
            };
            for (uint32_t q = s.first_page; q < s.first_page + s.n_pages; q++) {
                const pqg_page_desc& pg = pages[q];
                if (pg.num_values == 0) continue;
                const uint64_t lo = pg.payload_off & ~uint64_t(15), end = pg.payload_off + pg.payload_size;
                const bool too_big = ((end + 15) & ~uint64_t(15)) - lo > tile_cap;
                if (tileable && too_big && !p->is_str && s.max_def <= 0 && !((pg.flags & PQG_PAGE_FLAG_DICT) && s.has_dict)) {
                    // an oversized PLAIN REQUIRED page (foreign writers: 64 KB .. 1 MB) is the
                    // concatenation of smaller ones: cut it into 1 KB virtual pages, 8 per tile
                    const uint32_t W = static_cast<uint32_t>(p->width), sub_vals = 1024u / W;
                    uint32_t remaining = pg.num_values, size_left = pg.payload_size;
                    uint64_t off = pg.payload_off, row = pg.out_row_base;
                    while (remaining) {
                        const uint32_t nv = std::min(remaining, sub_vals);
                        uint32_t bytes = std::min(size_left, nv * W); // a truncated page shows up in its last slices
                        pqg_page_desc v = pg;
                        v.payload_off = off; v.out_row_base = row; v.payload_size = bytes; v.num_values = nv;
                        const uint32_t vq = n_pages + static_cast<uint32_t>(p->virt_pages.size());
                        p->virt_pages.push_back(v);
                        p->virt_parent.push_back(q);
                        place(vq, v);
                        off += bytes; size_left -= bytes; row += nv; remaining -= nv;
                    }
                    continue;
                }
                if (!tileable || too_big) { flush(); slow.push_back(q); continue; }
                // OPTIONAL pages of more than 1024 slots: the tile kernel takes them only when they hold no nulls (one level run:
                // a REQUIRED page).  When the table's producer looked at the levels, the others skip the tile kernel (it would stage them and
                // hand them over) and go to the block decode of the slow list directly; tables without the hint keep the device-side hand-over.
                if (flat_ok && s.max_def == 1 && pg.num_values > 1024u && (pg.flags & (PQG_PAGE_FLAG_LEVELS_SEEN | PQG_PAGE_FLAG_NO_NULLS)) == PQG_PAGE_FLAG_LEVELS_SEEN) { flush(); slow.push_back(q); continue; }
                // OPTIONAL pages beyond the writer's shapes: the tile kernel decodes the small ones itself and needs its index buffer
                if (!p->is_str && s.max_def == 1 && (pg.num_values > 1024u || pg.payload_size > 2048u)) p->opt_idx = true;
                place(q, pg);
            }
            flush();
        }
        p->chunk_tile_begin.push_back(static_cast<uint32_t>(tiles.size()));
        p->chunk_slow_begin.push_back(static_cast<uint32_t>(slow.size()));
        p->n_tiles = static_cast<uint32_t>(tiles.size());
        p->n_slow_host = static_cast<uint32_t>(slow.size());
        p->max_dict_blocks = std::max<uint32_t>(1, std::min<uint32_t>(64, (max_dict_n + 2047) / 2048));
        p->max_dict_n = max_dict_n;
    }

    std::vector<uint32_t> h_page_chars, h_page_base;
    if (p->is_str && n_pages) {
        bool all = true;
        h_page_chars.assign(n_pages + 1, 0);
        h_page_base.assign(n_pages + 1, 0);
        uint64_t col_total = 0;
        for (uint32_t c = 0; c < n_chunks && all; c++) {
            const pqg_chunk_desc& s = chunks[c];
            if (s.max_def > 0 || s.max_rep > 0) { all = false; break; }
            uint64_t acc = 0;
            for (uint32_t q = s.first_page; q < s.first_page + s.n_pages; q++) {
                const pqg_page_desc& pg = pages[q];
                if (((pg.flags & PQG_PAGE_FLAG_DICT) && s.has_dict) || static_cast<uint64_t>(pg.num_values) * 4 > pg.payload_size) { all = false; break; }
                const uint32_t bytes = pg.num_values ? pg.payload_size - 4u * pg.num_values : 0u;
                h_page_chars[q] = bytes;
                h_page_base[q] = static_cast<uint32_t>(acc);
                acc += bytes;
                if (acc > 0xffffffffull) { all = false; break; }
            }
            dc[c].char_base = col_total;
            col_total += acc;
        }
        p->host_sizes = all;
        if (all) p->chars_size = col_total;
    }
    auto alloc = [&](void** ptr, size_t bytes) -> cudaError_t { return cudaMalloc(ptr, bytes ? bytes : 16); };
    cudaError_t e;
#define PA(ptr, bytes) if ((e = alloc(reinterpret_cast<void**>(&(ptr)), (bytes))) != cudaSuccess) { pqg_plan_destroy(ctx, p); return cuda_fail(ctx, e, "cudaMalloc(plan)"); }
    PA(p->d_chunks, sizeof(DevChunk) * n_chunks);
    PA(p->d_pages, sizeof(pqg_page_desc) * std::max<size_t>(static_cast<size_t>(n_pages) + p->virt_pages.size(), 1));
    PA(p->d_err, sizeof(DevErr));
    if (p->any_dict) PA(p->d_dict, arena + 64);
    if (p->any_dict && p->is_str) { PA(p->d_dict_segs, n_segs * sizeof(DictSeg)); p->str_dict_blocks = std::max<uint32_t>(1, std::min<uint32_t>(64, (max_segs + 255) / 256)); }
    // host-listed pages, then room for EVERY page to be handed over on the device (the big-page kernel re-lists host-listed pages)
    PA(p->d_slow_pages, sizeof(uint32_t) * (slow.size() + static_cast<size_t>(n_pages) + p->virt_pages.size() + 2));
    PA(p->d_tiles, sizeof(TileDesc) * std::max<size_t>(tiles.size(), 1));
    p->flat_on = !p->is_str && !p->is_bool && (p->width == 4 || p->width == 8) && (p->any_def || !slow.empty());
    if (p->flat_on) {
        // every page may end up on the work list (the tile kernel hands over on the device): one entry per page, one block per 1024 slots
        uint64_t blocks = 1;
        for (uint32_t q = 0; q < n_pages; q++) blocks += (static_cast<uint64_t>(pages[q].num_values) + 1023u) >> 10;
        for (const pqg_page_desc& v : p->virt_pages) blocks += (static_cast<uint64_t>(v.num_values) + 1023u) >> 10;
        if (blocks > 0x7fffffffull) p->flat_on = false;
        else {
            const size_t cap = static_cast<size_t>(n_pages) + p->virt_pages.size() + 2;
            p->flat_blk_cap = static_cast<uint32_t>(blocks);
            PA(p->d_flat_pages, sizeof(FlatPage) * cap);
            PA(p->d_flat_append, sizeof(uint32_t) * cap);
            PA(p->d_flat_blk, sizeof(FlatBlk) * blocks);
            PA(p->d_flat_ckpt, sizeof(uint2) * blocks);
        }
    }
    if (p->any_def) PA(p->d_validity, ((slots + 31) / 32 + 1) * 4);
    if (p->any_def && !p->is_str) {
        std::vector<uint64_t> rr;
        for (uint32_t c = 0; c < n_chunks; c++)
            if (chunks[c].max_def == 0 && chunks[c].num_values) { rr.push_back(chunks[c].out_row_base); rr.push_back(chunks[c].out_row_base + chunks[c].num_values); }
        if (!rr.empty()) {
            PA(p->d_required_ranges, rr.size() * 8);
            p->n_required_ranges = static_cast<uint32_t>(rr.size() / 2);
            if ((e = cudaMemcpy(p->d_required_ranges, rr.data(), rr.size() * 8, cudaMemcpyHostToDevice)) != cudaSuccess) { pqg_plan_destroy(ctx, p); return cuda_fail(ctx, e, "cudaMemcpy(ranges)"); }
        }
    }
    if (p->is_str) {
        PA(p->d_offsets, (slots + n_chunks + 1) * 4);
        PA(p->d_page_chars, static_cast<size_t>(n_pages + 1) * 4);
        PA(p->d_page_char_base, static_cast<size_t>(n_pages + 1) * 4);
        PA(p->d_bases, (n_chunks + 2) * 8);
        if ((e = cudaHostAlloc(reinterpret_cast<void**>(&p->h_bases), (n_chunks + 2) * 8, cudaHostAllocDefault)) != cudaSuccess) {
            pqg_plan_destroy(ctx, p); return cuda_fail(ctx, e, "cudaHostAlloc");
        }
        std::memset(p->h_bases, 0, (n_chunks + 2) * 8);
        if (p->host_sizes) {
            for (uint32_t c = 0; c < n_chunks; c++) p->h_bases[c] = dc[c].char_base;
            p->h_bases[n_chunks] = p->h_bases[n_chunks + 1] = p->chars_size;
            PA(p->d_chars, p->chars_size + 64);
            p->chars_cap = p->chars_size;
            if ((e = cudaMemcpyAsync(p->d_page_chars, h_page_chars.data(), static_cast<size_t>(n_pages) * 4, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess ||
                (e = cudaMemcpyAsync(p->d_page_char_base, h_page_base.data(), static_cast<size_t>(n_pages) * 4, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess ||
                (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) { pqg_plan_destroy(ctx, p); return cuda_fail(ctx, e, "page size upload"); }
        }
    } else {
        PA(p->d_values, slots * p->width + 16);
    }
#undef PA
    if ((e = cudaHostAlloc(reinterpret_cast<void**>(&p->h_err), sizeof(DevErr), cudaHostAllocDefault)) != cudaSuccess) {
        pqg_plan_destroy(ctx, p); return cuda_fail(ctx, e, "cudaHostAlloc");
    }
    std::memset(p->h_err, 0, sizeof(DevErr));
    e = cudaMemcpyAsync(p->d_chunks, dc.data(), sizeof(DevChunk) * n_chunks, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && n_pages)
        e = cudaMemcpyAsync(p->d_pages, p->pages.data(), sizeof(pqg_page_desc) * n_pages, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && !p->virt_pages.empty())
        e = cudaMemcpyAsync(p->d_pages + n_pages, p->virt_pages.data(), sizeof(pqg_page_desc) * p->virt_pages.size(), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && !tiles.empty())
        e = cudaMemcpyAsync(p->d_tiles, tiles.data(), sizeof(TileDesc) * tiles.size(), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && !slow.empty())
        e = cudaMemcpyAsync(p->d_slow_pages, slow.data(), sizeof(uint32_t) * slow.size(), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { pqg_plan_destroy(ctx, p); return cuda_fail(ctx, e, "descriptor upload"); }
    for (auto& slot : p->evr) for (auto& ev : slot) {
        if ((e = cudaEventCreate(&ev)) != cudaSuccess) { pqg_plan_destroy(ctx, p); return cuda_fail(ctx, e, "cudaEventCreate"); }
    }
    // algorithmic output bytes (SURVEY.md section 8 d)
    if (!p->is_str) p->bytes_out = slots * p->width + (p->any_def ? (slots + 7) / 8 : 0);
    *out = p;
    return PQG_OK;
}

int pqg_plan_set_option(pqg_plan* plan, int option, int value) {
    if (!plan) return PQG_ERR_ARG;
    if (option == PQG_OPT_PARTITIONED_DICT) { plan->no_part = value == 0; return PQG_OK; }
    if (option == PQG_OPT_REGEX_TILE_BARRIER) { plan->regex_tile_sync = value != 0; return PQG_OK; }
    return PQG_ERR_ARG;
}

int pqg_plan_set_image(pqg_ctx* ctx, pqg_plan* plan, const pqg_buf* image) {
    if (!ctx || !plan || !image) return fail(ctx, PQG_ERR_ARG, "pqg_plan_set_image: bad argument");
    if (plan->ext_image) return fail(ctx, PQG_ERR_UNSUPPORTED, "pqg_plan_set_image: not for plans of pqg_plan_create_ext");
    if (image->size < plan->image->size) return fail(ctx, PQG_ERR_ARG, "pqg_plan_set_image: image is smaller than the planned one");
    plan->image = image;
    return PQG_OK;
}

static DecodeParams make_params(const pqg_plan* p) {
    DecodeParams P;
    std::memset(&P, 0, sizeof(P));
    P.image = p->image->d; P.image_size = p->image->size;
    P.chunks = p->d_chunks; P.pages = p->d_pages;
    P.page_begin = 0; P.page_end = static_cast<uint32_t>(p->pages.size());
    P.n_chunks = static_cast<uint32_t>(p->chunks.size());
    P.dict_arena = p->d_dict; P.dict_segs = p->d_dict_segs; P.values = p->d_values; P.validity = p->d_validity; P.handover_hint = p->handover_seen;
    P.offsets = p->d_offsets; P.chars = p->d_chars;
    P.page_chars = p->d_page_chars; P.page_char_base = p->d_page_char_base; P.err = p->d_err;
    P.tiles = p->d_tiles; P.tile_lo = 0; P.tile_hi = p->n_tiles; P.dict_smem = p->dict_smem;
    P.slow_lo = 0; P.slow_hi = p->n_slow_host; P.slow_pages = p->d_slow_pages; P.slow_append = p->d_slow_pages + p->n_slow_host;
    P.slow_cap = static_cast<uint32_t>(p->pages.size() + p->virt_pages.size() + 1); P.n_slots = p->n_slots;
    P.chunk_lo = 0;
    P.tile_bytes = p->tile_bytes;
    P.identity_dict = p->identity ? 1u : 0u;
    P.opt_idx = p->opt_idx ? 1u : 0u;
    P.exact_sizes = p->force_exact ? 1u : 0u;
    if (p->flat_on) {
        P.flat = &p->d_err->flat; P.flat_pages = p->d_flat_pages; P.flat_blk = p->d_flat_blk;
        P.flat_ckpt = p->d_flat_ckpt; P.flat_blk_cap = p->flat_blk_cap; P.flat_append = p->d_flat_append;
    }
    return P;
}

// reset of the error record + work counters at the start of a run
static cudaError_t reset_err(pqg_plan* p, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(p->d_err, 0xFF, 8, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(reinterpret_cast<uint8_t*>(p->d_err) + 8, 0, sizeof(DevErr) - 8, s);
    return e;
}

// decode of chunks [c0, c1) of a fixed-width plan on stream s; returns kernels launched or -1
static cudaError_t reset_validity(pqg_plan* p, cudaStream_t s) {
    if (!p->d_validity) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(p->d_validity, 0, ((p->n_slots + 31) / 32 + 1) * 4, s);
    if (e == cudaSuccess && p->n_required_ranges) {
        k_validity_ranges<<<p->n_required_ranges, 256, 0, s>>>(p->d_validity, p->d_required_ranges);
        e = cudaGetLastError();
    }
    return e;
}

static int fixed_subrun(pqg_ctx* ctx, pqg_plan* p, uint32_t c0, uint32_t c1, bool reset_counters, cudaStream_t s, cudaError_t* err,
                        cudaEvent_t ev_tiles_begin = nullptr, cudaEvent_t ev_tiles_end = nullptr) {
    DecodeParams P = make_params(p);
    int launches = 0;
    cudaError_t e = cudaSuccess;
    if (reset_counters) e = cudaMemsetAsync(&p->d_err->slow_count, 0, kWorkCounterBytes, s);
    P.chunk_lo = c0;
    P.tile_lo = p->chunk_tile_begin[c0]; P.tile_hi = p->chunk_tile_begin[c1];
    P.slow_lo = p->chunk_slow_begin[c0]; P.slow_hi = p->chunk_slow_begin[c1];
    bool any_dict = false;
    for (uint32_t c = c0; c < c1 && !p->identity; c++) any_dict = any_dict || p->chunks[c].has_dict; // dictionary-form plans keep no device dictionary
    if (e == cudaSuccess && any_dict) { e = launch_dict_prepare(P, c1 - c0, p->width, p->max_dict_blocks, s); launches += static_cast<int>(dict_prepare_launches(p->width)); }
    if (e == cudaSuccess && ev_tiles_begin) e = cudaEventRecord(ev_tiles_begin, s);
    if (e == cudaSuccess && P.tile_hi > P.tile_lo) {
        {
            // Dictionaries too large for shared memory are gathered from L2: keep the dictionaries of
            // the chunks that are in flight together within ~48 MB (a 1 M-entry INT64 dictionary is
            // 8 MB per chunk; with every chunk of a 320 M-row column in one launch the 126 MB L2
            // thrashed and the gather fell from 2.2 to 1.1 TB/s) -- one launch per chunk group.
            // Plans over several columns: a launch also ends where the KIND of work changes (PLAIN copy /
            // shared-memory dictionary / L2 gather) -- streaming PLAIN pages next to a gather evicts the
            // dictionaries (measured: everything in one launch ran 2.4x slower than one launch per column).
            // kind 3: dictionaries of 32 KB .. 512 KB in REQUIRED-only plans -- partitioned over 2^k sibling CTAs' shared memories
            const bool part_ok = !p->any_def && !p->identity && (p->width == 4 || p->width == 8) && !p->no_part;
            auto kind_of = [&](const pqg_chunk_desc& ck) {
                const uint64_t db = ck.has_dict ? static_cast<uint64_t>(ck.dict_num_values) * p->width : 0;
                if (!ck.has_dict) return 0;
                if (db <= static_cast<uint64_t>(kMaxSmemDictBytes)) return 1;
                return (part_ok && db <= static_cast<uint64_t>(kPartDictBytes) * kPartMaxParts) ? 3 : 2;
            };
            uint32_t g0 = c0;
            while (g0 < c1 && e == cudaSuccess) {
                uint64_t dict_bytes = 0;
                uint32_t g1 = g0;
                const int kind = kind_of(p->chunks[g0]);
                while (g1 < c1) {
                    const pqg_chunk_desc& ck = p->chunks[g1];
                    const uint64_t db = ck.has_dict ? static_cast<uint64_t>(ck.dict_num_values) * p->width : 0;
                    const bool in_smem = db <= static_cast<uint64_t>(kMaxSmemDictBytes);
                    if (g1 > g0 && (kind_of(ck) != kind || (!in_smem && dict_bytes + db > (48ull << 20)))) break;
                    if (!in_smem) dict_bytes += db;
                    g1++;
                }
                DecodeParams Pg = P;
                Pg.tile_lo = p->chunk_tile_begin[g0]; Pg.tile_hi = p->chunk_tile_begin[g1];
                // shared memory for the largest staged dictionary of THIS launch only (the carve-out eats L1)
                Pg.dict_smem = 0;
                for (uint32_t c = g0; c < g1; c++) {
                    const pqg_chunk_desc& ck = p->chunks[c];
                    const uint64_t db = ck.has_dict ? (static_cast<uint64_t>(ck.dict_num_values) * p->width + 15u) & ~15ull : 0;
                    if (db && db <= p->dict_smem) Pg.dict_smem = std::max<uint32_t>(Pg.dict_smem, static_cast<uint32_t>(db));
                }
                if (kind == 3) { // partitioned dictionaries: one launch per chunk (a CTA keeps one dictionary part for the whole launch)
                    for (uint32_t c = g0; c < g1 && e == cudaSuccess; c++) {
                        const uint64_t db = static_cast<uint64_t>(p->chunks[c].dict_num_values) * p->width;
                        DecodeParams Pc = P;
                        Pc.chunk_lo = c;
                        Pc.tile_lo = p->chunk_tile_begin[c]; Pc.tile_hi = p->chunk_tile_begin[c + 1];
                        uint32_t bits = 0;
                        while ((static_cast<uint64_t>(kPartDictBytes) << bits) < db) bits++;
                        Pc.part_bits = bits;
                        Pc.part_entries = static_cast<uint32_t>(kPartDictBytes / p->width);
                        Pc.dict_smem = static_cast<uint32_t>(std::min<uint64_t>(kPartDictBytes, (db + 15u) & ~15ull));
                        if (Pc.tile_hi > Pc.tile_lo) { e = launch_fixed_tiles_part(Pc, p->width, ctx->sm_count, s); launches++; p->tile_launches++; }
                    }
                    g0 = g1;
                    continue;
                }
                if (Pg.tile_hi > Pg.tile_lo) { e = launch_fixed_tiles(Pg, p->width, ctx->sm_count, s); launches++; p->tile_launches++; }
                g0 = g1;
            }
        }
    }
    if (e == cudaSuccess && ev_tiles_end) e = cudaEventRecord(ev_tiles_end, s);
    // plans with OPTIONAL chunks or host-listed pages (oversized, or of chunks the tile kernel does not take), 4/8-byte values:
    // the slow list -- host-listed pages and what the tile kernel handed over -- is decoded as a flat list of 1024-slot blocks;
    // what those kernels cannot take is listed again for the general kernel
    DecodeParams Pg = P;
    // (no host-listed pages and the tile kernel handed nothing over in the previous run: the general kernel alone -- it takes
    //  whatever shows up -- instead of three launches over an empty list)
    if (e == cudaSuccess && p->flat_on && (P.slow_hi > P.slow_lo || (P.tile_hi > P.tile_lo && p->tile_handover_seen != 0))) {
        p->flat_ran = true;
        // (dictionary-form plans keep no device dictionary but their pages carry indices all the same: index buffers needed)
        e = launch_flat_pages(P, p->width, ctx->sm_count, any_dict || p->identity, p->max_dict_n <= 0xffffu, s);
        launches += static_cast<int>(flat_launches());
        Pg.slow_hi = Pg.slow_lo;
        Pg.slow_append = P.flat_append;
        Pg.append_count = &p->d_err->flat.handed;
    }
    // the general kernel: every page shape, one warp per page
    if (e == cudaSuccess && (P.slow_hi > P.slow_lo || P.tile_hi > P.tile_lo)) {
        e = launch_decode_fixed(Pg, p->width, p->is_bool, ctx->sm_count, s);
        launches++;
    }
    *err = e;
    return e == cudaSuccess ? launches : -1;
}

int pqg_plan_run(pqg_ctx* ctx, pqg_plan* p) {
    if (!ctx || !p) return fail(ctx, PQG_ERR_ARG, "pqg_plan_run: bad argument");
    CU(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const bool prof = ctx->profiling;
    uint32_t launches = 0;
    p->timed = prof;
    p->tile_launches = 0;
    if (prof) { p->ev = p->evr[p->runs_timed % pqg_plan::kTimingSlots]; p->runs_timed++; }
    if (prof) CU(ctx, cudaEventRecord(p->ev[0], s));
    CU(ctx, reset_err(p, s));
    CU(ctx, reset_validity(p, s));
    if (p->n_xform) { // pqg_plan_create_ext: the plan's image from the caller's (V2 framing, SNAPPY)
        CU(ctx, launch_xform(p->ext_src->d, p->ext_image->d, p->d_xform, p->n_xform, p->d_err, ctx->sm_count, s));
        launches++;
    }
    DecodeParams P = make_params(p);
    if (!p->is_str) {
        // ev0 .. ev1 dictionary preparation, ev1 .. ev2 the tile kernel, ev2 .. ev3 the general kernel
        cudaError_t ce = cudaSuccess;
        int n = fixed_subrun(ctx, p, 0, P.n_chunks, false, s, &ce, prof ? p->ev[1] : nullptr, prof ? p->ev[2] : nullptr);
        if (n < 0) return cuda_fail(ctx, ce, "decode launch");
        launches += static_cast<uint32_t>(n);
        if (prof) CU(ctx, cudaEventRecord(p->ev[3], s));
    } else if (p->host_sizes && !p->force_exact) {
        // byte counts, page bases and chunk bases are known from the page headers: copy pass only
        if (prof) { CU(ctx, cudaEventRecord(p->ev[1], s)); CU(ctx, cudaEventRecord(p->ev[2], s)); }
        P.check_layout = 1;
        if (P.page_end) { CU(ctx, launch_str_copy(P, p->any_dict, ctx->sm_count, s)); launches++; }
        if (prof) CU(ctx, cudaEventRecord(p->ev[3], s));
    } else {
        if (p->any_dict) {
            CU(ctx, launch_dict_prepare(P, P.n_chunks, 0, p->str_dict_blocks, s));
            launches += dict_prepare_launches(0);
        }
        if (prof) CU(ctx, cudaEventRecord(p->ev[1], s));
        if (P.page_end) { CU(ctx, launch_str_sizes(P, ctx->sm_count, s)); launches++; }
        CU(ctx, launch_str_scan(P, p->d_bases, s));
        launches += 2;
        CU(ctx, cudaMemcpyAsync(p->h_bases, p->d_bases, (P.n_chunks + 2) * 8, cudaMemcpyDeviceToHost, s));
        if (prof) CU(ctx, cudaEventRecord(p->ev[2], s));
        if (p->chars_sized && p->d_chars) {
            // the same plan ran before: reuse its chars buffer and do not wait for the size pass -- the copy kernel
            // checks the grand total against the capacity on the device (PQG_PAGE_CHARS_CAP -> finish re-sizes and re-runs)
            P.chars = p->d_chars;
            P.chars_cap = p->chars_cap;
            P.total_chars = p->d_bases + P.n_chunks + 1;
        } else {
            // the output size is data dependent: wait for the size pass, (re)allocate if needed
            CU(ctx, cudaStreamSynchronize(s));
            uint64_t total = p->h_bases[P.n_chunks + 1];
            if (total > p->chars_cap || !p->d_chars) {
                cudaFree(p->d_chars);
                p->d_chars = nullptr;
                CU(ctx, cudaMalloc(reinterpret_cast<void**>(&p->d_chars), total + 64));
                p->chars_cap = total;
            }
            p->chars_size = total;
            P.chars = p->d_chars;
        }
        if (P.page_end) { CU(ctx, launch_str_copy(P, p->any_dict, ctx->sm_count, s)); launches++; }
        if (prof) CU(ctx, cudaEventRecord(p->ev[3], s));
    }
    if (prof) CU(ctx, cudaEventRecord(p->ev[4], s));
    p->tm.launches = launches;
    p->tm.tile_launches = p->tile_launches;
    p->last_launches = launches;
    ctx->launches += launches;
    p->ran = true;
    p->run_pending = true;
    return PQG_OK;
}

static void elapsed_of(const pqg_plan* p, cudaEvent_t* ev, pqg_timings* tm) {
    cudaEventElapsedTime(&tm->dict_ms, ev[0], ev[1]);
    if (p->is_str) {
        cudaEventElapsedTime(&tm->str_size_ms, ev[1], ev[2]);
        cudaEventElapsedTime(&tm->str_copy_ms, ev[2], ev[3]);
        tm->fixed_ms = 0;
        tm->general_ms = 0;
    } else {
        cudaEventElapsedTime(&tm->fixed_ms, ev[1], ev[2]);
        cudaEventElapsedTime(&tm->general_ms, ev[2], ev[3]);
        tm->str_size_ms = tm->str_copy_ms = 0;
    }
    cudaEventElapsedTime(&tm->total_ms, ev[0], ev[4]);
}

int pqg_plan_finish(pqg_ctx* ctx, pqg_plan* p, pqg_page_error* err) {
    if (!ctx || !p) return fail(ctx, PQG_ERR_ARG, "pqg_plan_finish: bad argument");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemcpyAsync(p->h_err, p->d_err, sizeof(DevErr), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    const bool was_pipelined = p->pipelined_in_flight;
    if (p->pipelined_in_flight) { // the copy-out stream carries this plan's last D2H
        if (ctx->d2h) CU(ctx, cudaStreamSynchronize(ctx->d2h));
        p->pipelined_in_flight = false;
    }
    p->run_pending = false;
    if (p->timed) elapsed_of(p, p->ev, &p->tm);
    if (p->is_str && !(p->host_sizes && !p->force_exact) && p->ran) p->chars_size = p->h_bases[p->chunks.size() + 1]; // of the run that just finished
    if (p->is_str) {
        uint64_t slots = p->n_slots;
        p->bytes_out = p->chars_size + 4 * (slots + p->chunks.size()) + (p->any_def ? (slots + 7) / 8 : 0);
    }
    pqg_page_error pe{};
    const DevErr& d = *p->h_err;
    p->tile_handover_seen = d.slow_count;
    p->handover_seen = p->flat_ran ? d.flat.handed : d.slow_count; // what the general kernel had to take
    p->flat_ran = false;
    if (std::getenv("PQG_DEBUG"))
        std::fprintf(stderr, "[pqg] plan: %zu pages, %u tiles, %u host-listed slow pages, %u handed over by the tile kernel (last sub-run), bad_index %u\n",
                     p->pages.size(), p->n_tiles, p->n_slow_host, d.slow_count, d.bad_index);
    pe.count = d.count;
    if (d.count) {
        pe.page = static_cast<uint32_t>(d.key >> 32);
        const uint32_t raw_page = pe.page;
        if (pe.page >= p->pages.size() && pe.page - p->pages.size() < p->virt_parent.size())
            pe.page = p->virt_parent[pe.page - p->pages.size()]; // a slice of an oversized page: report the page
        pe.code = static_cast<uint32_t>(d.key & 0xffffffffu);
        if (d.d_page == raw_page) { pe.pos = d.d_pos; pe.need = d.d_need; pe.size = d.d_size; }
    }
    if (err) *err = pe;
    if (!pe.count && d.bad_index && !p->d_validity) {
        // an out-of-range dictionary index in a REQUIRED chunk: NULL in the reference whatever the repetition
        // (column_reader.cpp:190-194).  The plan has no validity bitmap: add one and decode again.
        if (was_pipelined)
            return fail(ctx, PQG_ERR_PAGE, "out-of-range dictionary index in a REQUIRED column chunk (a null in the reference): "
                                           "decode this column through pqg_plan_run, which adds a validity bitmap");
        const uint64_t slots = p->n_slots;
        CU(ctx, cudaMalloc(reinterpret_cast<void**>(&p->d_validity), ((slots + 31) / 32 + 1) * 4));
        if (!p->is_str) { // every chunk is REQUIRED: all slots start valid, the general kernel clears the bad ones
            std::vector<uint64_t> rr;
            for (const pqg_chunk_desc& ck : p->chunks) if (ck.num_values) { rr.push_back(ck.out_row_base); rr.push_back(ck.out_row_base + ck.num_values); }
            if (!rr.empty()) {
                CU(ctx, cudaMalloc(reinterpret_cast<void**>(&p->d_required_ranges), rr.size() * 8));
                CU(ctx, cudaMemcpy(p->d_required_ranges, rr.data(), rr.size() * 8, cudaMemcpyHostToDevice));
                p->n_required_ranges = static_cast<uint32_t>(rr.size() / 2);
            }
            p->bytes_out += (slots + 7) / 8;
        }
        p->any_def = true;
        p->forced_validity = true;
        int rc = pqg_plan_run(ctx, p);
        if (rc != PQG_OK) return rc;
        return pqg_plan_finish(ctx, p, err);
    }
    if (pe.count && pe.code == PQG_PAGE_CHARS_CAP && p->is_str && p->chars_sized) {
        // the chars buffer kept from the previous run is too small for this one: size it by the host again
        p->chars_sized = false;
        int rc = pqg_plan_run(ctx, p);
        if (rc != PQG_OK) return rc;
        return pqg_plan_finish(ctx, p, err);
    }
    if (pe.count && pe.code == PQG_PAGE_LAYOUT && p->is_str && !p->force_exact) {
        // a page broke the byte count a fast path assumed (a PLAIN page carrying bytes beyond its values / an out-of-range
        // index inside an RLE run of a dictionary page): redo the column with the exact size pass
        p->force_exact = true;
        p->chars_sized = false;
        int rc = pqg_plan_run(ctx, p);
        if (rc != PQG_OK) return rc;
        return pqg_plan_finish(ctx, p, err);
    }
    if (p->is_str && !pe.count) p->chars_sized = true;
    if (pe.count) {
        char msg[256];
        if (pe.code == PQG_PAGE_TRUNCATED || pe.code == PQG_PAGE_DICT_TRUNCATED)
            std::snprintf(msg, sizeof(msg), "ByteBuffer: read beyond end (pos=%u need=%u size=%u)", pe.pos, pe.need, pe.size);
        else if (pe.code == PQG_PAGE_BAD_BIT_WIDTH)
            std::snprintf(msg, sizeof(msg), "page %u: dictionary index bit width %u > 32 is not supported", pe.page, pe.need);
        else if (pe.code == PQG_PAGE_BAD_RUN)
            std::snprintf(msg, sizeof(msg), "page %u: zero-length RLE/bit-packed run (undefined in the reference decoder)", pe.page);
        else if (pe.code == PQG_PAGE_DECOMPRESS)
            std::snprintf(msg, sizeof(msg), "page %u: the page (or its chunk's dictionary page) does not decompress to its uncompressed_page_size", pe.page);
        else if (pe.code == PQG_PAGE_CHARS_OVERFLOW)
            std::snprintf(msg, sizeof(msg), "page %u: column chunk exceeds 4 GiB of string bytes", pe.page);
        else
            std::snprintf(msg, sizeof(msg), "page %u: decode error %u", pe.page, pe.code);
        return fail(ctx, PQG_ERR_PAGE, msg);
    }
    return PQG_OK;
}

int pqg_plan_run_pipelined(pqg_ctx* ctx, pqg_plan* p, pqg_buf* image, const pqg_h2d_range* ranges, uint32_t n_ranges,
                           void* host_values, uint32_t* host_validity) {
    if (!ctx || !p || !image || (!ranges && n_ranges)) return fail(ctx, PQG_ERR_ARG, "pqg_plan_run_pipelined: bad argument");
    if (p->is_str) return fail(ctx, PQG_ERR_UNSUPPORTED, "pqg_plan_run_pipelined: BYTE_ARRAY plans need the size pass first; use pqg_plan_run");
    if (p->ext_image) return fail(ctx, PQG_ERR_UNSUPPORTED, "pqg_plan_run_pipelined: not for plans of pqg_plan_create_ext; use pqg_plan_run");
    if (image != p->image || !image->owned) return fail(ctx, PQG_ERR_ARG, "pqg_plan_run_pipelined: image must be the plan's own pqg_buf_alloc buffer");
    CU(ctx, cudaSetDevice(ctx->device));
    if (!ctx->h2d) CU(ctx, cudaStreamCreateWithFlags(&ctx->h2d, cudaStreamNonBlocking));
    if (!ctx->d2h) CU(ctx, cudaStreamCreateWithFlags(&ctx->d2h, cudaStreamNonBlocking));
    const uint32_t nc = static_cast<uint32_t>(p->chunks.size());
    if (p->pipe_ev.empty()) {
        p->pipe_ev.assign(2 * static_cast<size_t>(nc), nullptr);
        for (auto& e : p->pipe_ev) CU(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CU(ctx, cudaEventCreateWithFlags(&p->ev_idle, cudaEventDisableTiming));
        CU(ctx, cudaEventRecord(p->ev_idle, ctx->stream));
    }
    for (uint32_t i = 0; i < n_ranges; i++) {
        if (ranges[i].chunk >= nc || (i && ranges[i].chunk < ranges[i - 1].chunk) || ranges[i].image_off + ranges[i].len > image->size)
            return fail(ctx, PQG_ERR_ARG, "pqg_plan_run_pipelined: ranges must be sorted by chunk and lie inside the image");
    }
    cudaStream_t s = ctx->stream;
    const bool prof = ctx->profiling;
    p->timed = prof;
    p->tile_launches = 0;
    if (prof) { p->ev = p->evr[p->runs_timed % pqg_plan::kTimingSlots]; p->runs_timed++; CU(ctx, cudaEventRecord(p->ev[0], s)); CU(ctx, cudaEventRecord(p->ev[1], s)); }
    CU(ctx, reset_err(p, s));
    CU(ctx, reset_validity(p, s));
    // the image buffer is overwritten: wait for the kernels of the previous run of this plan
    CU(ctx, cudaStreamWaitEvent(ctx->h2d, p->ev_idle, 0));
    uint32_t launches = 0, r = 0;
    for (uint32_t c = 0; c < nc; c++) {
        for (; r < n_ranges && ranges[r].chunk == c; r++)
            if (ranges[r].len) CU(ctx, cudaMemcpyAsync(image->d + ranges[r].image_off, ranges[r].host, ranges[r].len, cudaMemcpyHostToDevice, ctx->h2d));
        cudaEvent_t ev_in = p->pipe_ev[2 * c], ev_k = p->pipe_ev[2 * c + 1];
        CU(ctx, cudaEventRecord(ev_in, ctx->h2d));
        CU(ctx, cudaStreamWaitEvent(s, ev_in, 0));
        cudaError_t ce = cudaSuccess;
        int n = fixed_subrun(ctx, p, c, c + 1, true, s, &ce);
        if (n < 0) return cuda_fail(ctx, ce, "decode launch");
        launches += static_cast<uint32_t>(n);
        CU(ctx, cudaEventRecord(ev_k, s));
        CU(ctx, cudaStreamWaitEvent(ctx->d2h, ev_k, 0));
        const pqg_chunk_desc& ck = p->chunks[c];
        if (host_values && ck.num_values)
            CU(ctx, cudaMemcpyAsync(static_cast<uint8_t*>(host_values) + ck.out_row_base * p->width, p->d_values + ck.out_row_base * p->width,
                                    ck.num_values * p->width, cudaMemcpyDeviceToHost, ctx->d2h));
        if (host_validity && p->d_validity && ck.num_values) {
            uint64_t w0 = ck.out_row_base >> 5, w1 = (ck.out_row_base + ck.num_values + 31) >> 5;
            CU(ctx, cudaMemcpyAsync(host_validity + w0, p->d_validity + w0, (w1 - w0) * 4, cudaMemcpyDeviceToHost, ctx->d2h));
        }
    }
    CU(ctx, cudaEventRecord(p->ev_idle, s));
    if (prof) { CU(ctx, cudaEventRecord(p->ev[2], s)); CU(ctx, cudaEventRecord(p->ev[3], s)); CU(ctx, cudaEventRecord(p->ev[4], s)); }
    p->tm.launches = launches;
    p->tm.tile_launches = p->tile_launches;
    p->last_launches = launches;
    ctx->launches += launches;
    p->ran = true;
    p->run_pending = true;
    p->pipelined_in_flight = true;
    return PQG_OK;
}

int pqg_plan_timings(const pqg_plan* plan, pqg_timings* out) {
    if (!plan || !out) return PQG_ERR_ARG;
    *out = plan->tm;
    return PQG_OK;
}

int pqg_plan_timings_avg(const pqg_plan* plan, uint32_t last_n, pqg_timings* out, uint32_t* n_used) {
    if (!plan || !out) return PQG_ERR_ARG;
    pqg_timings acc{};
    uint32_t n = 0;
    uint64_t have = plan->runs_timed < (uint64_t)pqg_plan::kTimingSlots ? plan->runs_timed : (uint64_t)pqg_plan::kTimingSlots;
    if (last_n == 0 || last_n > have) last_n = static_cast<uint32_t>(have);
    for (uint32_t i = 0; i < last_n; i++) {
        uint64_t run = plan->runs_timed - 1 - i;
        cudaEvent_t* ev = const_cast<cudaEvent_t*>(plan->evr[run % pqg_plan::kTimingSlots]);
        if (cudaEventQuery(ev[4]) != cudaSuccess) { cudaGetLastError(); continue; }
        pqg_timings t{};
        elapsed_of(plan, ev, &t);
        acc.dict_ms += t.dict_ms; acc.fixed_ms += t.fixed_ms; acc.str_size_ms += t.str_size_ms;
        acc.str_copy_ms += t.str_copy_ms; acc.total_ms += t.total_ms; acc.general_ms += t.general_ms;
        n++;
    }
    if (n) { acc.dict_ms /= n; acc.fixed_ms /= n; acc.str_size_ms /= n; acc.str_copy_ms /= n; acc.total_ms /= n; acc.general_ms /= n; }
    acc.launches = plan->last_launches;
    acc.tile_launches = plan->tile_launches;
    *out = acc;
    if (n_used) *n_used = n;
    return PQG_OK;
}

uint64_t pqg_plan_num_slots(const pqg_plan* p) { return p ? p->n_slots : 0; }
uint32_t pqg_plan_value_width(const pqg_plan* p) { return p ? static_cast<uint32_t>(p->width) : 0; }
const void* pqg_plan_values(const pqg_plan* p) { return p ? p->d_values : nullptr; }
const uint32_t* pqg_plan_validity(const pqg_plan* p) { return p ? p->d_validity : nullptr; }
const uint32_t* pqg_plan_offsets(const pqg_plan* p) { return p ? p->d_offsets : nullptr; }
const uint8_t* pqg_plan_chars(const pqg_plan* p) { return p ? p->d_chars : nullptr; }
uint64_t pqg_plan_chars_size(const pqg_plan* p) { return p ? p->chars_size : 0; }
uint64_t pqg_plan_bytes_in(const pqg_plan* p) { return p ? p->bytes_in : 0; }
uint64_t pqg_plan_bytes_out(const pqg_plan* p) { return p ? p->bytes_out : 0; }

int pqg_plan_char_bases(pqg_ctx* ctx, const pqg_plan* p, uint64_t* out, uint32_t n) {
    if (!ctx || !p || !out) return fail(ctx, PQG_ERR_ARG, "pqg_plan_char_bases: bad argument");
    if (!p->is_str || !p->ran) return fail(ctx, PQG_ERR_ARG, "pqg_plan_char_bases: not a decoded BYTE_ARRAY plan");
    uint32_t have = static_cast<uint32_t>(p->chunks.size()) + 1;
    for (uint32_t i = 0; i < n && i < have; i++) out[i] = p->h_bases[i];
    return PQG_OK;
}

int pqg_plan_download(pqg_ctx* ctx, const pqg_plan* p, void* values, uint32_t* validity, uint32_t* offsets, uint8_t* chars) {
    if (!ctx || !p) return fail(ctx, PQG_ERR_ARG, "pqg_plan_download: bad argument");
    CU(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    if (values && p->d_values) CU(ctx, cudaMemcpyAsync(values, p->d_values, p->n_slots * p->width, cudaMemcpyDeviceToHost, s));
    if (validity && p->d_validity) CU(ctx, cudaMemcpyAsync(validity, p->d_validity, ((p->n_slots + 31) / 32) * 4, cudaMemcpyDeviceToHost, s));
    if (offsets && p->d_offsets) CU(ctx, cudaMemcpyAsync(offsets, p->d_offsets, (p->n_slots + p->chunks.size()) * 4, cudaMemcpyDeviceToHost, s));
    if (chars && p->d_chars && p->chars_size) CU(ctx, cudaMemcpyAsync(chars, p->d_chars, p->chars_size, cudaMemcpyDeviceToHost, s));
    return PQG_OK;
}

} // extern "C"

// accessors for the other translation units (regex scan, chunk index)
namespace pqg {
DecodeParams plan_params(const pqg_plan* p) { return make_params(p); }
cudaStream_t ctx_stream(const pqg_ctx* c) { return c->stream; }
int ctx_sm_count(const pqg_ctx* c) { return c->sm_count; }
int ctx_device(const pqg_ctx* c) { return c->device; }
void ctx_add_launches(pqg_ctx* c, uint32_t n) { c->launches += n; }
int ctx_fail(pqg_ctx* c, int code, const std::string& m) { return fail(c, code, m); }
bool plan_is_str(const pqg_plan* p) { return p->is_str; }
int plan_width(const pqg_plan* p) { return p->is_bool ? 1 : p->width; }
bool plan_ran(const pqg_plan* p) { return p->ran; }
bool plan_run_pending(const pqg_plan* p) { return p->run_pending; }
bool plan_regex_tile_sync(const pqg_plan* p) { return p->regex_tile_sync; }
bool plan_any_dict(const pqg_plan* p) { return p->any_dict; }
size_t plan_dict_arena_bytes(const pqg_plan* p) { return p->dict_bytes; }
uint32_t plan_str_dict_blocks(const pqg_plan* p) { return p->str_dict_blocks; }
uint64_t plan_slots(const pqg_plan* p) { return p->n_slots; }
const std::vector<pqg_chunk_desc>& plan_chunks(const pqg_plan* p) { return p->chunks; }
const std::vector<pqg_page_desc>& plan_pages(const pqg_plan* p) { return p->pages; }
uint32_t plan_max_page_values(const pqg_plan* p) {
    if (p->max_page_values == 0xffffffffu) {
        uint32_t m = 0;
        for (const pqg_page_desc& pd : p->pages) m = std::max(m, pd.num_values);
        p->max_page_values = m;
    }
    return p->max_page_values;
}
} // namespace pqg
