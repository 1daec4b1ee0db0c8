// tools/probes/gather_probe.cu -- memory-pipeline probe (development aid, not product code):
// how fast can one B200 gather the neighbour rows of an R-MAT CSR (512-B rows, d=128) with
//   (a) per-lane LDG.128 loads, U rows in flight per 16-lane group (the shape of the force kernel)
//   (b) TMA bulk copies (cp.async.bulk, one 512-B row per copy) into a per-warp shared-memory ring
// and minimal math (sum of the rows)?  Gives the ceiling the force kernel's gather can reach.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_probe gather_probe.cu
//   ./gather_probe rowptr.u64 colids.u32 n
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

struct Item { uint32_t v, len; uint64_t e0; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(b) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void* d, const void* s, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(d)), "l"(s), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

constexpr int D = 128, V4 = D / 4;

// (a) groups of LPR lanes, one item per group, U rows in flight
template <int LPR, int U>
__global__ void __launch_bounds__(256) k_ldg(const Item* items, uint32_t n_items, const uint32_t* colids, const float* X, float* out) {
    constexpr int G = 32 / LPR, VPL = V4 / LPR;
    const int lane = threadIdx.x & 31, g = lane / LPR, l = lane % LPR;
    const uint32_t t = (blockIdx.x * 8 + (threadIdx.x >> 5)) * G + g;
    Item it{0, 0, 0};
    if (t < n_items) it = items[t];
    float4 acc[VPL];
    for (int k = 0; k < VPL; k++) acc[k] = __ldcg(reinterpret_cast<const float4*>(X + (size_t)it.v * D) + k * LPR + l);
    uint32_t cmax = it.len;
    for (int o = 16; o >= LPR; o >>= 1) cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
    for (uint32_t base = 0; base < cmax; base += LPR) {
        const uint32_t nb = it.len > base ? min((uint32_t)LPR, it.len - base) : 0u;
        const uint32_t mine = (uint32_t)l < nb ? __ldg(colids + it.e0 + base + l) : it.v;
        const uint32_t nbm = min((uint32_t)LPR, cmax - base);
        for (uint32_t t0 = 0; t0 < nbm; t0 += U) {
            float4 r[U][VPL];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const uint32_t j = __shfl_sync(0xffffffffu, mine, t0 + u, LPR);
#pragma unroll
                for (int k = 0; k < VPL; k++) r[u][k] = __ldcg(reinterpret_cast<const float4*>(X + (size_t)j * D) + k * LPR + l);
            }
#pragma unroll
            for (int u = 0; u < U; u++)
#pragma unroll
                for (int k = 0; k < VPL; k++) { acc[k].x += r[u][k].x; acc[k].y += r[u][k].y; acc[k].z += r[u][k].z; acc[k].w += r[u][k].w; }
        }
    }
    if (t < n_items && !(it.len & 0x80000000u))
        for (int k = 0; k < VPL; k++) __stcg(reinterpret_cast<float4*>(out + (size_t)it.v * D) + k * LPR + l, acc[k]);
}

// (b) one item per warp; ring of R row slots per warp filled by TMA bulk copies, S stages of R/S rows
template <int R, int S, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_tma(const Item* items, uint32_t n_items, const uint32_t* colids, const float* X, float* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int RS = R / S;                       // rows per stage
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem) + w * S;
    float* ring = reinterpret_cast<float*>(smem + 1024) + (size_t)w * R * D;
    if (lane == 0) for (int s = 0; s < S; s++) mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const uint32_t t = blockIdx.x * WARPS + w;
    if (t >= n_items) return;
    const Item it = items[t];
    float4 acc = __ldcg(reinterpret_cast<const float4*>(X + (size_t)it.v * D) + lane);
    const uint32_t nst = (it.len + RS - 1) / RS;    // stages this item needs
    uint32_t parity = 0;                             // bit s = parity of stage s
    auto issue = [&](uint32_t st) {                  // fill stage st % S with rows [st*RS, ...)
        const uint32_t first = st * RS, cnt = min((uint32_t)RS, it.len - first);
        const int slot = st % S;
        if (lane == 0) mbar_expect_tx(bars + slot, cnt * D * 4);
        __syncwarp();
        if ((uint32_t)lane < cnt) {
            const uint32_t j = __ldg(colids + it.e0 + first + lane);
            bulk_g2s(ring + (size_t)(slot * RS + lane) * D, X + (size_t)j * D, D * 4, bars + slot);
        }
    };
    for (uint32_t st = 0; st < min(nst, (uint32_t)S); st++) issue(st);
    for (uint32_t st = 0; st < nst; st++) {
        const int slot = st % S;
        mbar_wait(bars + slot, (parity >> slot) & 1u);
        parity ^= 1u << slot;
        const uint32_t cnt = min((uint32_t)RS, it.len - st * RS);
        for (uint32_t r = 0; r < cnt; r++) {
            const float4 x = *(reinterpret_cast<const float4*>(ring + (size_t)(slot * RS + r) * D) + lane);
            acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
        }
        __syncwarp();
        if (st + S < nst) issue(st + S);
    }
    __stcg(reinterpret_cast<float4*>(out + (size_t)it.v * D) + lane, acc);
}

int main(int argc, char** argv) {
    if (argc < 4) { printf("usage: %s rowptr.u64 colids.u32 n\n", argv[0]); return 1; }
    const uint64_t n = strtoull(argv[3], nullptr, 10);
    std::vector<uint64_t> rp(n + 1);
    FILE* f = fopen(argv[1], "rb"); if (!f || fread(rp.data(), 8, n + 1, f) != n + 1) { printf("bad rowptr\n"); return 1; } fclose(f);
    const uint64_t nnz = rp[n];
    std::vector<uint32_t> ci(nnz);
    f = fopen(argv[2], "rb"); if (!f || fread(ci.data(), 4, nnz, f) != nnz) { printf("bad colids\n"); return 1; } fclose(f);
    const uint32_t chunk = 128;
    std::vector<Item> items;
    for (uint64_t v = 0; v < n; v++) {
        uint64_t deg = rp[v + 1] - rp[v], e0 = rp[v];
        if (deg <= chunk) { items.push_back({(uint32_t)v, (uint32_t)deg, e0}); continue; }
        uint64_t nc = (deg + chunk - 1) / chunk, base = deg / nc, extra = deg % nc;
        for (uint64_t c = 0; c < nc; c++) { uint32_t len = (uint32_t)(base + (c < extra)); items.push_back({(uint32_t)v, len, e0}); e0 += len; }
    }
    // longest first, like the engine's plan
    std::stable_sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.len > b.len; });
    const uint32_t n_items = (uint32_t)items.size();
    printf("n %llu nnz %llu items %u\n", (unsigned long long)n, (unsigned long long)nnz, n_items);
    Item* d_items; uint32_t* d_ci; float *d_X, *d_out;
    CK(cudaMalloc(&d_items, sizeof(Item) * n_items)); CK(cudaMalloc(&d_ci, 4 * nnz));
    CK(cudaMalloc(&d_X, n * D * 4)); CK(cudaMalloc(&d_out, n * D * 4));
    CK(cudaMemcpy(d_items, items.data(), sizeof(Item) * n_items, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ci, ci.data(), 4 * nnz, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_X, 0, n * D * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double bytes = (double)(nnz + n) * D * 4 + (double)n * D * 4;
    auto report = [&](const char* name, float ms) { printf("%-28s %8.3f ms  %8.1f GB/s algorithmic\n", name, ms, bytes / ms / 1e6); };
#define RUN_LDG(LPR, U) { auto k = k_ldg<LPR, U>; const unsigned per = 8 * (32 / LPR); float best = 1e9; \
        for (int r = 0; r < 4; r++) { cudaEventRecord(e0); k<<<(n_items + per - 1) / per, 256>>>(d_items, n_items, d_ci, d_X, d_out); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms); } \
        char nm[64]; snprintf(nm, 64, "ldg LPR=%d U=%d", LPR, U); report(nm, best); }
    RUN_LDG(16, 2) RUN_LDG(16, 4) RUN_LDG(16, 8) RUN_LDG(32, 4) RUN_LDG(32, 8) RUN_LDG(32, 16) RUN_LDG(8, 2) RUN_LDG(8, 4)
#define RUN_TMA(R, S, W) { auto k = k_tma<R, S, W>; size_t sm = 1024 + (size_t)W * R * D * 4; CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); float best = 1e9; \
        for (int r = 0; r < 4; r++) { cudaEventRecord(e0); k<<<(n_items + W - 1) / W, W * 32, sm>>>(d_items, n_items, d_ci, d_X, d_out); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms); } \
        char nm[64]; snprintf(nm, 64, "tma R=%d S=%d warps=%d", R, S, W); report(nm, best); }
    RUN_TMA(8, 2, 8) RUN_TMA(16, 2, 8) RUN_TMA(16, 4, 8) RUN_TMA(32, 2, 4) RUN_TMA(32, 4, 4) RUN_TMA(16, 2, 4) RUN_TMA(8, 2, 4) RUN_TMA(32, 4, 8) RUN_TMA(64, 2, 4)
    return 0;
}
