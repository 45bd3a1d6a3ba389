// Microbenchmark (not part of the product): random 8-byte dictionary gathers from an L2-resident table
//   mode 0: ld.global.cg.u64 per lane (what k_fixed_tiles does for dictionaries > 32 KB)
//   mode 1: one 16-byte cp.async.bulk (UBLKCP, the TMA unit) per value into shared memory, then ld.shared
//   mode 2: half the values through each path
// Prints gathers per clock and SM.  build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_gather ubench_gather.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

constexpr int kThreads = 256;
constexpr int kBatch = 4; // bulk copies in flight per thread and round

__global__ void __launch_bounds__(kThreads, 4) k_gather(const uint64_t* __restrict__ table, uint32_t mask, int rounds, int mode, uint64_t* out) {
    __shared__ __align__(16) uint8_t slots[kThreads * kBatch * 16];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t tid = threadIdx.x;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t x = (blockIdx.x * kThreads + tid) * 2654435761u + 12345u;
    uint64_t acc = 0;
    uint32_t phase = 0;
    for (int r = 0; r < rounds; r++) {
        uint32_t ix[kBatch];
#pragma unroll
        for (int b = 0; b < kBatch; b++) { x = x * 1664525u + 1013904223u; ix[b] = (x >> 8) & mask; }
        const int n_tma = mode == 0 ? 0 : (mode == 1 ? kBatch : kBatch / 2);
        if (n_tma) {
            if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(kThreads * n_tma * 16) : "memory");
            __syncthreads();
            for (int b = 0; b < n_tma; b++) {
                const uint64_t* src = table + (ix[b] & ~1u);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];"
                             ::"r"(smem_u32(slots + (tid * kBatch + b) * 16)), "l"(src), "r"(smem_u32(&bar)) : "memory");
            }
        }
        for (int b = n_tma; b < kBatch; b++) {
            uint64_t v;
            asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(table + ix[b]));
            acc += v;
        }
        if (n_tma) {
            asm volatile(
                "{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)), "r"(phase) : "memory");
            phase ^= 1;
            for (int b = 0; b < n_tma; b++) acc += *reinterpret_cast<const uint64_t*>(slots + (tid * kBatch + b) * 16 + (ix[b] & 1u) * 8);
            __syncthreads();
        }
    }
    if (acc == 0x1234567) out[0] = acc;
}

// pipelined split: per warp two mbarriers and two slot buffers; the bulk copies of round r + 1 are in flight
// while the lanes do the ld.global gathers of round r.  n_tma of the kBatch values per lane go through the TMA unit.
__global__ void __launch_bounds__(kThreads, 4) k_gather_split(const uint64_t* __restrict__ table, uint32_t mask, int rounds, int n_tma, uint64_t* out) {
    __shared__ __align__(16) uint8_t slots[2][kThreads * kBatch * 16];
    __shared__ __align__(8) uint64_t bars[kThreads / 32][2];
    const uint32_t tid = threadIdx.x, l = tid & 31u, w = tid >> 5;
    if (l == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[w][0])), "r"(1) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[w][1])), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t x = (blockIdx.x * kThreads + tid) * 2654435761u + 12345u;
    uint64_t acc = 0;
    uint32_t ixn[kBatch], ixc[kBatch];
    auto draw = [&](uint32_t* ix) {
#pragma unroll
        for (int b = 0; b < kBatch; b++) { x = x * 1664525u + 1013904223u; ix[b] = (x >> 8) & mask; }
    };
    auto issue = [&](const uint32_t* ix, int buf) {
        if (l == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[w][buf])), "r"(32 * n_tma * 16) : "memory");
        __syncwarp();
        for (int b = 0; b < n_tma; b++)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];"
                         ::"r"(smem_u32(slots[buf] + (tid * kBatch + b) * 16)), "l"(table + (ix[b] & ~1u)), "r"(smem_u32(&bars[w][buf])) : "memory");
    };
    draw(ixc);
    issue(ixc, 0);
    uint32_t ph0 = 0, ph1 = 0;
    for (int r = 0; r < rounds; r++) {
        const int buf = r & 1;
        draw(ixn);
        if (r + 1 < rounds) issue(ixn, buf ^ 1);
        for (int b = n_tma; b < kBatch; b++) {
            uint64_t v;
            asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(table + ixc[b]));
            acc += v;
        }
        const uint32_t ph = buf ? ph1 : ph0;
        asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bars[w][buf])), "r"(ph) : "memory");
        if (buf) ph1 ^= 1; else ph0 ^= 1;
        for (int b = 0; b < n_tma; b++) acc += *reinterpret_cast<const uint64_t*>(slots[buf] + (tid * kBatch + b) * 16 + (ixc[b] & 1u) * 8);
        __syncwarp();
#pragma unroll
        for (int b = 0; b < kBatch; b++) ixc[b] = ixn[b];
    }
    if (acc == 0x1234567) out[0] = acc;
}

// texture path: tex1Dfetch<int2> over the same table; n_tex of the kBatch values per lane through the TEX pipe
__global__ void __launch_bounds__(kThreads, 4) k_gather_tex(const uint64_t* __restrict__ table, cudaTextureObject_t tex, uint32_t mask, int rounds, int n_tex, uint64_t* out) {
    const uint32_t tid = threadIdx.x;
    uint32_t x = (blockIdx.x * kThreads + tid) * 2654435761u + 12345u;
    uint64_t acc = 0;
    for (int r = 0; r < rounds; r++) {
        uint32_t ix[kBatch];
#pragma unroll
        for (int b = 0; b < kBatch; b++) { x = x * 1664525u + 1013904223u; ix[b] = (x >> 8) & mask; }
#pragma unroll
        for (int b = 0; b < kBatch; b++) {
            if (b < n_tex) { const int2 v = tex1Dfetch<int2>(tex, static_cast<int>(ix[b])); acc += static_cast<uint32_t>(v.x) + (static_cast<uint64_t>(static_cast<uint32_t>(v.y)) << 32); }
            else { uint64_t v; asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(table + ix[b])); acc += v; }
        }
    }
    if (acc == 0x1234567) out[0] = acc;
}

int main(int argc, char** argv) {
    const uint32_t entries = argc > 1 ? static_cast<uint32_t>(std::atoi(argv[1])) : 65536u; // power of two
    const int rounds = 2000;
    uint64_t* d_table; uint64_t* d_out;
    cudaMalloc(&d_table, static_cast<size_t>(entries) * 8 + 64);
    cudaMalloc(&d_out, 8);
    std::vector<uint64_t> h(entries);
    for (uint32_t i = 0; i < entries; i++) h[i] = i * 2654435761ull + 1;
    cudaMemcpy(d_table, h.data(), static_cast<size_t>(entries) * 8, cudaMemcpyHostToDevice);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int grid = prop.multiProcessorCount * 4;
    for (int mode = 0; mode < 3; mode++) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k_gather<<<grid, kThreads>>>(d_table, entries - 1, 50, mode, d_out);
        cudaEventRecord(e0);
        k_gather<<<grid, kThreads>>>(d_table, entries - 1, rounds, mode, d_out);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        const double gathers = static_cast<double>(grid) * kThreads * kBatch * rounds;
        std::printf("mode %d entries %u: %.3f ms, %.1f G gathers/s, %.3f gathers/clk/SM (at %d MHz nominal)  [%s]\n", mode, entries, ms,
                    gathers / ms / 1e6, gathers / (ms * 1e-3) / (clk_khz * 1e3) / prop.multiProcessorCount, clk_khz / 1000, cudaGetErrorString(e));
    }
    {
        cudaResourceDesc rd{}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = d_table;
        rd.res.linear.desc = cudaCreateChannelDesc<int2>(); rd.res.linear.sizeInBytes = static_cast<size_t>(entries) * 8;
        cudaTextureDesc td{}; td.readMode = cudaReadModeElementType;
        cudaTextureObject_t tex = 0;
        cudaError_t ce = cudaCreateTextureObject(&tex, &rd, &td, nullptr);
        for (int n_tex = 1; n_tex <= 4 && ce == cudaSuccess; n_tex++) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            k_gather_tex<<<grid, kThreads>>>(d_table, tex, entries - 1, 50, n_tex, d_out);
            cudaEventRecord(e0);
            k_gather_tex<<<grid, kThreads>>>(d_table, tex, entries - 1, rounds, n_tex, d_out);
            cudaEventRecord(e1);
            cudaError_t e = cudaDeviceSynchronize();
            float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
            const double gathers = static_cast<double>(grid) * kThreads * kBatch * rounds;
            std::printf("%d of %d through tex1Dfetch<int2>, entries %u: %.3f ms, %.1f G gathers/s, %.3f gathers/clk/SM  [%s]\n", n_tex, kBatch, entries, ms,
                        gathers / ms / 1e6, gathers / (ms * 1e-3) / (clk_khz * 1e3) / prop.multiProcessorCount, cudaGetErrorString(e));
        }
        if (ce != cudaSuccess) std::printf("texture object: %s\n", cudaGetErrorString(ce));
    }
    for (int n_tma = 1; n_tma <= 2; n_tma++) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k_gather_split<<<grid, kThreads>>>(d_table, entries - 1, 50, n_tma, d_out);
        cudaEventRecord(e0);
        k_gather_split<<<grid, kThreads>>>(d_table, entries - 1, rounds, n_tma, d_out);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        const double gathers = static_cast<double>(grid) * kThreads * kBatch * rounds;
        std::printf("pipelined split %d of %d through the TMA unit, entries %u: %.3f ms, %.1f G gathers/s, %.3f gathers/clk/SM  [%s]\n", n_tma, kBatch, entries, ms,
                    gathers / ms / 1e6, gathers / (ms * 1e-3) / (clk_khz * 1e3) / prop.multiProcessorCount, cudaGetErrorString(e));
    }
    return 0;
}
