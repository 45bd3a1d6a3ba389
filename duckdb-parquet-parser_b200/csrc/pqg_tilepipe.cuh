// pqg_tilepipe.cuh -- the TMA-staged page-tile pipeline shared by the tile kernels.
//
// The plan's host side cuts every chunk into tiles (TileDesc): <= kTilePages consecutive pages
// whose bytes (page headers in between included) fit kTileBytes.  A CTA owns a contiguous run
// of tiles.  Tile bytes + the tile's page descriptors are staged into a kTileStages-deep
// shared-memory ring with cp.async.bulk (1-D TMA, SASS UBLKCP) completing on an mbarrier per
// stage; the CTA's warps take one page each out of shared memory.  There is no CTA-wide barrier
// per tile: a warp that is done with a stage bumps the stage's counter, and the LAST warp to do so
// refills the stage kTileStages tiles ahead -- so warps run up to one tile apart and a slow page
// does not idle the other seven warps.
// No register staging, no per-warp global latency chain (descriptor -> chunk -> payload).
#pragma once
#include "pqg_page.cuh"

namespace pqg {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared (TMA); dst/src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// the suspend-time hint parks the warp in hardware instead of spinning through issue slots
constexpr uint32_t kMbarSuspendNs = 20000;
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(kMbarSuspendNs) : "memory");
}

// TB = image bytes per tile (kTileBytes, or kTileBytesLarge for plans whose pages are fat)
__host__ __device__ constexpr int tile_stage_bytes(int TB) { return TB + 16 + kTilePages * static_cast<int>(sizeof(pqg_page_desc)); }
// ST + 1 mbarriers (<= 32 bytes), stage counters (<= 32 bytes), per-stage tile meta (32 B) and prefetched descriptors (32 B)
__host__ __device__ constexpr int tile_bar_bytes(int ST = kTileStages) { return (64 + ST * 64 + 127) & ~127; }
constexpr int kBarBytes = tile_bar_bytes();
__host__ __device__ constexpr int tile_pipe_bytes(int TB, int ST = kTileStages) { return tile_bar_bytes(ST) + ST * tile_stage_bytes(TB); } // shared memory of the pipeline itself
constexpr int kTilePipeBytes = tile_pipe_bytes(kTileBytes);

struct TileMeta { uint64_t byte_lo; uint32_t first_page; uint32_t n_pages; uint32_t chunk_idx; uint32_t pad; };

// Runs the pipeline over the CTA's tiles.  `smem` = tile_pipe_bytes(TB) of 128-byte aligned shared
// memory.  on_chunk(chunk_idx, extra_bar, phase&) is called by all threads (CTA-uniform) when
// the chunk changes -- it may __syncthreads and stage per-chunk data with a bulk copy on
// extra_bar.  on_page(q, pd, payload) is called by one warp per page.
// A CTA of several 256-thread GROUPS runs one pipeline per group (its own ring, barriers and tile span): `grp` names the
// group's barrier (0 = the whole CTA: __syncthreads), its thread / warp ids and its tile span.
struct PipeGroup { uint32_t bar_id, tid, warp, t0, t1; };
__device__ __forceinline__ void pipe_sync(uint32_t bar_id) {
    if (bar_id == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(kThreadsPerCta) : "memory");
}

template <int TB = kTileBytes, int ST = kTileStages, class OnChunk, class OnPage>
__device__ __forceinline__ void tile_pipeline(const DecodeParams& P, uint8_t* smem, OnChunk&& on_chunk, OnPage&& on_page, const PipeGroup* grp = nullptr) {
    static_assert(ST >= 2 && ST <= 3, "barrier block layout: 4 mbarriers and 4 counters at most");
    constexpr int kTileStages = ST; // (shadows the global default inside the pipeline)
    constexpr int kBarBytes = tile_bar_bytes(ST);
    constexpr int kStageBytes = tile_stage_bytes(TB);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);          // [kTileStages] tiles, then one for per-chunk staging
    uint32_t* done = reinterpret_cast<uint32_t*>(smem + 32);     // [kTileStages] warps finished with the stage
    TileMeta* meta = reinterpret_cast<TileMeta*>(smem + 64);     // [kTileStages]
    TileDesc* ahead = reinterpret_cast<TileDesc*>(smem + 64 + kTileStages * 32); // [kTileStages] descriptor of the tile that refills the stage
    uint8_t* ring = smem + kBarBytes;
    const uint32_t t0 = grp ? grp->t0 : P.tile_lo + blockIdx.x * P.tiles_per_cta;
    const uint32_t t1 = grp ? grp->t1 : min(P.tile_hi, t0 + P.tiles_per_cta);
    const uint32_t tid = grp ? grp->tid : threadIdx.x, wid = grp ? grp->warp : warp_id(), bar_id = grp ? grp->bar_id : 0u;
    if (tid == 0) {
        for (int i = 0; i <= kTileStages; i++) mbar_init(&full[i], 1);
        for (int i = 0; i < kTileStages; i++) done[i] = 0;
        fence_mbar_init();
    }
    pipe_sync(bar_id);
    if (t0 >= t1) return;
    auto issue = [&](uint32_t t, const TileDesc& td) { // one thread
        const uint32_t st = (t - t0) % kTileStages;
        uint8_t* dst = ring + st * kStageBytes;
        meta[st] = TileMeta{td.byte_lo, td.first_page, td.n_pages, td.chunk_idx, 0};
        const uint32_t pbytes = td.n_pages * static_cast<uint32_t>(sizeof(pqg_page_desc));
        mbar_expect_tx(&full[st], td.byte_len + pbytes);
        bulk_g2s(dst, P.image + td.byte_lo, td.byte_len, &full[st]);
        bulk_g2s(dst + TB + 16, P.pages + td.first_page, pbytes, &full[st]);
    };
    if (tid == 0) {
        for (uint32_t t = t0; t < min(t1, t0 + kTileStages); t++) issue(t, P.tiles[t]);
        for (uint32_t t = t0 + kTileStages; t < min(t1, t0 + 2 * kTileStages); t++) ahead[(t - t0) % kTileStages] = P.tiles[t];
    }
    uint32_t cur_chunk = 0xffffffffu, extra_phase = 0;
    for (uint32_t t = t0; t < t1; t++) {
        const uint32_t st = (t - t0) % kTileStages;
        mbar_wait(&full[st], ((t - t0) / kTileStages) & 1u);
        const uint8_t* tile = ring + st * kStageBytes;
        const TileMeta tm = meta[st];
        if (tm.chunk_idx != cur_chunk) { // uniform across the CTA
            cur_chunk = tm.chunk_idx;
            on_chunk(cur_chunk, &full[kTileStages], extra_phase);
        }
        const pqg_page_desc* pds = reinterpret_cast<const pqg_page_desc*>(tile + TB + 16);
        for (uint32_t j = wid; j < tm.n_pages; j += kWarpsPerCta) {
            const pqg_page_desc pd = pds[j];
            PQG_ASSERT(pd.payload_off >= tm.byte_lo && pd.payload_off - tm.byte_lo + pd.payload_size <= static_cast<uint64_t>(TB) + 16u);
            on_page(tm.first_page + j, pd, tile + (pd.payload_off - tm.byte_lo));
        }
        if (P.tile_sync) { // the classic barrier per tile, refill by thread 0
            pipe_sync(bar_id);
            if (tid == 0 && t + kTileStages < t1) issue(t + kTileStages, P.tiles[t + kTileStages]);
            continue;
        }
        __syncwarp();
        if (lane_id() == 0) {
            // release the stage: my warp's reads are done; the last warp re-arms and refills it
            uint32_t old;
            asm volatile("atom.acq_rel.cta.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(&done[st])) : "memory");
            if (old == kWarpsPerCta - 1u) {
                done[st] = 0;
                if (t + kTileStages < t1) {
                    // the descriptor was fetched one round ago (by the thread that refilled this stage last
                    // time; its store is ordered before this point by the counter's release/acquire chain)
                    const TileDesc td = ahead[st];
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic reads of the stage before the bulk write
                    issue(t + kTileStages, td);
                    if (t + 2 * kTileStages < t1) ahead[st] = P.tiles[t + 2 * kTileStages];
                }
            }
        }
    }
}

// grid sizing shared by the tile kernels: contiguous tile spans, two waves of resident CTAs
inline uint32_t tile_grid(uint32_t n_tiles, int sm_count, int resident, uint32_t* tiles_per_cta) {
    if (resident < 1) resident = 1;
    uint32_t target = static_cast<uint32_t>(sm_count) * static_cast<uint32_t>(resident) * 2u;
    uint32_t per = (n_tiles + target - 1) / target;
    if (per < 4) per = 4;
    *tiles_per_cta = per;
    return (n_tiles + per - 1) / per;
}

} // namespace pqg
