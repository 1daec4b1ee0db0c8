// tools/probes/random_gather_probe.cu -- DRAM ceiling for RANDOM 512-byte (or 256-byte) row reads
// (development aid, not product code).  The measured HBM peak in MEASURED_PEAKS.json is a sequential
// copy; the force kernel's DRAM traffic is row gathers at uniformly random addresses of a table far
// larger than L2 (negatives, non-hub neighbours).  This probe reads M uniformly random rows of an
// N-row table with 16-lane groups (8 for 256-byte rows), U rows in flight per group, no arithmetic
// but a sum, and prints the achieved GB/s: the ceiling the gather can reach when every row misses L2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o random_gather_probe random_gather_probe.cu
//   ./random_gather_probe [table_GiB=8] [row_bytes=512]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return (uint32_t)((z ^ (z >> 31)) >> 20);
}

template <int ROWB, int U, int MINB>
__global__ void __launch_bounds__(256, MINB) k_gather(const float4* __restrict__ X, uint32_t nrows, uint32_t per_group, float* out) {
    constexpr int LPR = ROWB / 32;           // lanes per row: each lane reads 2 x 16 bytes of the row
    constexpr int G = 32 / LPR;
    const int lane = threadIdx.x & 31, g = lane / LPR, l = lane % LPR;
    const uint64_t gid = ((uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * G + g;
    float4 acc0 = make_float4(0, 0, 0, 0), acc1 = acc0;
    for (uint32_t t = 0; t < per_group; t += U) {
        float4 a[U], b[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t j = mix(gid * 0x9E3779B97F4A7C15ULL + t + u) % nrows;
            const float4* row = X + (size_t)j * (ROWB / 16);
            a[u] = __ldcg(row + l);
            b[u] = __ldcg(row + LPR + l);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            acc0.x += a[u].x; acc0.y += a[u].y; acc0.z += a[u].z; acc0.w += a[u].w;
            acc1.x += b[u].x; acc1.y += b[u].y; acc1.z += b[u].z; acc1.w += b[u].w;
        }
    }
    if (acc0.x + acc0.y + acc0.z + acc0.w + acc1.x + acc1.y + acc1.z + acc1.w == 12345.678f) out[0] = 1.f;
}

template <int ROWB, int U, int MINB>
static void run(const float4* X, uint32_t nrows, float* out, const char* name) {
    constexpr int G = 32 / (ROWB / 32);
    const uint32_t per_group = 256;
    const unsigned grid = 148 * MINB * 8;
    const double bytes = (double)grid * 8 * G * per_group * ROWB;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_gather<ROWB, U, MINB><<<grid, 256>>>(X, nrows, per_group, out);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(e0));
        k_gather<ROWB, U, MINB><<<grid, 256>>>(X, nrows, per_group, out);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    printf("{\"probe\": \"random_row_gather\", \"row_bytes\": %d, \"rows_in_flight_per_group\": %d, \"ctas_per_sm\": %d, \"layout\": \"%s\", \"GBps\": %.1f, \"ms\": %.3f}\n",
           ROWB, U, MINB, name, bytes / best / 1e6, best);
}

int main(int argc, char** argv) {
    const double gib = argc > 1 ? atof(argv[1]) : 8.0;
    const int rowb = argc > 2 ? atoi(argv[2]) : 512;
    const size_t bytes = (size_t)(gib * (1ull << 30));
    float4* X; float* out;
    CK(cudaMalloc(&X, bytes)); CK(cudaMalloc(&out, 64));
    CK(cudaMemset(X, 0, bytes));
    const uint32_t nrows = (uint32_t)(bytes / rowb);
    if (rowb == 512) {
        run<512, 2, 4>(X, nrows, out, "2 rows in flight, 4 CTAs/SM");
        run<512, 2, 5>(X, nrows, out, "2 rows in flight, 5 CTAs/SM");
        run<512, 4, 4>(X, nrows, out, "4 rows in flight, 4 CTAs/SM");
        run<512, 8, 2>(X, nrows, out, "8 rows in flight, 2 CTAs/SM");
        run<512, 8, 4>(X, nrows, out, "8 rows in flight, 4 CTAs/SM");
        run<512, 16, 2>(X, nrows, out, "16 rows in flight, 2 CTAs/SM");
    } else {
        run<256, 2, 5>(X, nrows, out, "2 rows in flight, 5 CTAs/SM");
        run<256, 4, 4>(X, nrows, out, "4 rows in flight, 4 CTAs/SM");
        run<256, 8, 4>(X, nrows, out, "8 rows in flight, 4 CTAs/SM");
        run<256, 16, 2>(X, nrows, out, "16 rows in flight, 2 CTAs/SM");
    }
    return 0;
}
