// tools/probes/nvls_probe.cu -- does this box support NVLink multicast (NVLS) objects, and does a
// multimem.st from one GPU land in every GPU's memory?  Single process, all visible GPUs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o nvls_probe nvls_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

#define DRV(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s_; cuGetErrorString(r_, &s_); printf("%s -> %s\n", #x, s_); return 1; } } while (0)
#define RT(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__global__ void mc_store(float* mc, int n, float v) {
    int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i < n) asm volatile("multimem.st.global.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(mc + i), "f"(v) : "memory");
}
__global__ void mc_store_bw(float4* mc, size_t n4) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i < n4; i += (size_t)gridDim.x * blockDim.x)
        asm volatile("multimem.st.global.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(mc + i), "f"(1.0f) : "memory");
}
__global__ void uc_store_bw(float4* a, float4* b, float4* c, size_t n4) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 v = make_float4(1, 1, 1, 1);
        if (a) a[i] = v; if (b) b[i] = v; if (c) c[i] = v;
    }
}

int main() {
    DRV(cuInit(0));
    int ndev = 0;
    RT(cudaGetDeviceCount(&ndev));
    printf("devices %d\n", ndev);
    std::vector<CUdevice> dev(ndev);
    for (int d = 0; d < ndev; d++) {
        DRV(cuDeviceGet(&dev[d], d));
        int mc = 0, fab = 0, posix = 0;
        cuDeviceGetAttribute(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev[d]);
        cuDeviceGetAttribute(&fab, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED, dev[d]);
        cuDeviceGetAttribute(&posix, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev[d]);
        printf("dev %d multicast %d fabric %d posix_fd %d\n", d, mc, fab, posix);
        RT(cudaSetDevice(d));
        RT(cudaFree(0));
    }
    if (ndev < 2) { printf("need >= 2 devices for the store test\n"); return 0; }
    const size_t want = 64ull << 20;
    CUmulticastObjectProp mp;
    memset(&mp, 0, sizeof(mp));
    mp.numDevices = ndev;
    mp.size = want;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t gran = 0;
    DRV(cuMulticastGetGranularity(&gran, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED));
    printf("multicast granularity (recommended) %zu\n", gran);
    mp.size = (want + gran - 1) / gran * gran;
    CUmemGenericAllocationHandle mc;
    DRV(cuMulticastCreate(&mc, &mp));
    for (int d = 0; d < ndev; d++) DRV(cuMulticastAddDevice(mc, dev[d]));
    std::vector<CUmemGenericAllocationHandle> mem(ndev);
    std::vector<CUdeviceptr> uc(ndev), mcva(ndev);
    for (int d = 0; d < ndev; d++) {
        RT(cudaSetDevice(d));
        CUmemAllocationProp ap;
        memset(&ap, 0, sizeof(ap));
        ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        ap.location.id = d;
        ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
        DRV(cuMemCreate(&mem[d], mp.size, &ap, 0));
        DRV(cuMulticastBindMem(mc, 0, mem[d], 0, mp.size, 0));
        CUmemAccessDesc ad;
        ad.location.type = CU_MEM_LOCATION_TYPE_DEVICE; ad.location.id = d; ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        DRV(cuMemAddressReserve(&uc[d], mp.size, gran, 0, 0));
        DRV(cuMemMap(uc[d], mp.size, 0, mem[d], 0));
        DRV(cuMemSetAccess(uc[d], mp.size, &ad, 1));
        DRV(cuMemAddressReserve(&mcva[d], mp.size, gran, 0, 0));
        DRV(cuMemMap(mcva[d], mp.size, 0, mc, 0));
        DRV(cuMemSetAccess(mcva[d], mp.size, &ad, 1));
        RT(cudaMemset((void*)uc[d], 0, mp.size));
    }
    for (int d = 0; d < ndev; d++) { RT(cudaSetDevice(d)); RT(cudaDeviceSynchronize()); }
    RT(cudaSetDevice(0));
    mc_store<<<4, 256>>>((float*)mcva[0], 4096, 3.5f);
    RT(cudaDeviceSynchronize());
    for (int d = 0; d < ndev; d++) {
        RT(cudaSetDevice(d));
        float h[4];
        RT(cudaMemcpy(h, (void*)uc[d], sizeof(h), cudaMemcpyDeviceToHost));
        printf("dev %d sees %g %g %g %g\n", d, h[0], h[1], h[2], h[3]);
    }
    // bandwidth: multicast store of 64 MiB from device 0 vs unicast stores to every peer
    RT(cudaSetDevice(0));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const size_t n4 = want / 16;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0); mc_store_bw<<<148 * 8, 256>>>((float4*)mcva[0], n4); cudaEventRecord(e1);
        RT(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("multimem.st 64 MiB -> %d devices: %.3f ms  (%.1f GB/s payload)\n", ndev, ms, want / ms / 1e6);
    }
    // unicast to peers needs peer mappings of the VMM allocations: grant device 0 access
    for (int d = 1; d < ndev; d++) {
        CUmemAccessDesc ad;
        ad.location.type = CU_MEM_LOCATION_TYPE_DEVICE; ad.location.id = 0; ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        DRV(cuMemSetAccess(uc[d], mp.size, &ad, 1));
    }
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        uc_store_bw<<<148 * 8, 256>>>((float4*)uc[1], ndev > 2 ? (float4*)uc[2] : nullptr, ndev > 3 ? (float4*)uc[3] : nullptr, n4);
        cudaEventRecord(e1);
        RT(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("unicast st 64 MiB -> %d peers: %.3f ms  (%.1f GB/s egress)\n", ndev > 4 ? 3 : ndev - 1, ms, (double)want * (ndev > 4 ? 3 : ndev - 1) / ms / 1e6);
    }
    printf("NVLS_OK\n");
    return 0;
}
