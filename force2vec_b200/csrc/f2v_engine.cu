// force2vec_b200/csrc/f2v_engine.cu -- the C ABI of include/f2v.h: device state, the
// per-minibatch work plan, kernel dispatch.  Host-side only bookkeeping; all arithmetic
// of the force step is in f2v_kernels.cuh.  There is no CPU fallback in this file.
#include "../../include/f2v.h"
#include "f2v_kernels.cuh"
#include "f2v_plan.hpp"

#include <cuda.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <time.h>

#include <algorithm>
#include <climits>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <new>
#include <string>
#include <unistd.h>
#include <vector>

using namespace f2v;

// ------------------------------------------------------------------ errors -------------
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
// the host side of the library (f2v_host.cpp) reports through the same per-thread message
namespace f2v {
int host_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
}  // namespace f2v
#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t _e = (call);                                                              \
        if (_e != cudaSuccess)                                                                \
            return fail(F2V_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                                  \
    } while (0)

// ------------------------------------------------------------------ NCCL (dlopen) ------
// NCCL is loaded lazily so that the single-GPU path has no link-time dependency on it.
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static std::mutex g_load_mutex;           // engines of one process may be driven by several threads
static int nccl_load() {
    std::lock_guard<std::mutex> lock(g_load_mutex);
    if (g_nccl.handle) return F2V_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(F2V_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
    g_nccl.GetUniqueId = (int (*)(ncclUniqueId*))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(ncclComm_t))dlsym(h, "ncclCommDestroy");
    g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllGather");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllGather)
        return fail(F2V_ERR_NCCL, "libnccl is missing required symbols");
    g_nccl.handle = h;
    return F2V_OK;
}
#define NC(call)                                                                                 \
    do {                                                                                         \
        int _r = (call);                                                                         \
        if (_r != 0)                                                                             \
            return fail(F2V_ERR_NCCL, "%s failed: %s", #call,                                    \
                        g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "nccl error");       \
    } while (0)
constexpr int kNcclFloat32 = 7;   // ncclFloat32 in nccl.h's ncclDataType_t

// ------------------------------------------------------------------ driver API (VMM / multicast)
// Resolved through the runtime (cudaGetDriverEntryPoint): no link-time dependency on libcuda.
struct DrvApi {
    bool loaded = false;
    CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
    CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
    CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
    CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = nullptr;
    CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
    CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long) = nullptr;
    CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t) = nullptr;
    CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags) = nullptr;
    CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
    CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
};
static DrvApi g_drv;
static int drv_load() {
    std::lock_guard<std::mutex> lock(g_load_mutex);
    if (g_drv.loaded) return F2V_OK;
    struct { const char* name; void** slot; } tab[] = {
        {"cuGetErrorString", (void**)&g_drv.GetErrorString}, {"cuDeviceGet", (void**)&g_drv.DeviceGet},
        {"cuDeviceGetAttribute", (void**)&g_drv.DeviceGetAttribute}, {"cuMulticastCreate", (void**)&g_drv.MulticastCreate},
        {"cuMulticastAddDevice", (void**)&g_drv.MulticastAddDevice}, {"cuMulticastBindMem", (void**)&g_drv.MulticastBindMem},
        {"cuMulticastUnbind", (void**)&g_drv.MulticastUnbind}, {"cuMulticastGetGranularity", (void**)&g_drv.MulticastGetGranularity},
        {"cuMemCreate", (void**)&g_drv.MemCreate}, {"cuMemRelease", (void**)&g_drv.MemRelease},
        {"cuMemAddressReserve", (void**)&g_drv.MemAddressReserve}, {"cuMemAddressFree", (void**)&g_drv.MemAddressFree},
        {"cuMemMap", (void**)&g_drv.MemMap}, {"cuMemUnmap", (void**)&g_drv.MemUnmap}, {"cuMemSetAccess", (void**)&g_drv.MemSetAccess},
        {"cuMemExportToShareableHandle", (void**)&g_drv.MemExportToShareableHandle},
        {"cuMemImportFromShareableHandle", (void**)&g_drv.MemImportFromShareableHandle},
        {"cuMemGetAllocationGranularity", (void**)&g_drv.MemGetAllocationGranularity},
    };
    for (auto& t : tab) {
        cudaDriverEntryPointQueryResult q;
        cudaError_t r = cudaGetDriverEntryPoint(t.name, t.slot, cudaEnableDefault, &q);
        if (r != cudaSuccess || q != cudaDriverEntryPointSuccess || !*t.slot)
            return fail(F2V_ERR_CUDA, "driver entry point %s is not available", t.name);
    }
    g_drv.loaded = true;
    return F2V_OK;
}
#define DRV(call)                                                                                  \
    do {                                                                                           \
        CUresult _r = (call);                                                                      \
        if (_r != CUDA_SUCCESS) {                                                                  \
            const char* _s = nullptr;                                                              \
            if (g_drv.GetErrorString) g_drv.GetErrorString(_r, &_s);                               \
            return fail(F2V_ERR_CUDA, "%s failed: %s (%s:%d)", #call, _s ? _s : "?", __FILE__, __LINE__); \
        }                                                                                          \
    } while (0)

// ---- tiny unix-socket helpers: the multicast object's file descriptor has to travel from rank 0
// to the other processes (SCM_RIGHTS), and binding needs two all-rank handshakes
static int sock_send_fd(int sock, int fd) {
    char byte = 'F';
    struct iovec io = {&byte, 1};
    char ctl[CMSG_SPACE(sizeof(int))];
    memset(ctl, 0, sizeof(ctl));
    struct msghdr msg;
    memset(&msg, 0, sizeof(msg));
    msg.msg_iov = &io; msg.msg_iovlen = 1; msg.msg_control = ctl; msg.msg_controllen = sizeof(ctl);
    struct cmsghdr* c = CMSG_FIRSTHDR(&msg);
    c->cmsg_level = SOL_SOCKET; c->cmsg_type = SCM_RIGHTS; c->cmsg_len = CMSG_LEN(sizeof(int));
    memcpy(CMSG_DATA(c), &fd, sizeof(int));
    return sendmsg(sock, &msg, 0) == 1 ? 0 : -1;
}
static int sock_recv_fd(int sock) {
    char byte = 0;
    struct iovec io = {&byte, 1};
    char ctl[CMSG_SPACE(sizeof(int))];
    memset(ctl, 0, sizeof(ctl));
    struct msghdr msg;
    memset(&msg, 0, sizeof(msg));
    msg.msg_iov = &io; msg.msg_iovlen = 1; msg.msg_control = ctl; msg.msg_controllen = sizeof(ctl);
    if (recvmsg(sock, &msg, 0) != 1) return -1;
    struct cmsghdr* c = CMSG_FIRSTHDR(&msg);
    if (!c || c->cmsg_type != SCM_RIGHTS) return -1;
    int fd = -1;
    memcpy(&fd, CMSG_DATA(c), sizeof(int));
    return fd;
}
static int sock_byte(int sock, bool send_it, char b) {
    if (send_it) return send(sock, &b, 1, 0) == 1 ? 0 : -1;
    char got = 0;
    return (recv(sock, &got, 1, MSG_WAITALL) == 1 && got == b) ? 0 : -1;
}
// The hand-off must never block for ever: a rank that failed (or died) before its handshake would
// otherwise leave every peer in accept()/recv().  Receive and send time-outs turn that into an error.
static void sock_timeouts(int fd, int seconds) {
    struct timeval tv = {seconds, 0};
    setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));     // also bounds accept() on a listening socket
    setsockopt(fd, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof(tv));
}
constexpr int kHandoffTimeoutS = 120;
static void sock_addr(struct sockaddr_un* a, socklen_t* len, const char* name) {
    memset(a, 0, sizeof(*a));
    a->sun_family = AF_UNIX;
    const size_t n = strlen(name);
    memcpy(a->sun_path + 1, name, n);            // abstract namespace: no file system entry
    *len = (socklen_t)(offsetof(struct sockaddr_un, sun_path) + 1 + n);
}

// ------------------------------------------------------------------ engine -------------
struct Plan {
    uint32_t batch = 0, chunk = 0, par = 0, min_chunk = 0;
    bool walk = false;
    int rank = 0, world = 1, assign = 0;
    uint64_t first_row = 0, nrows = 0;      // row range covered (whole table for epochs)
    uint64_t nb = 0;
    std::vector<uint64_t> item_ptr;         // nb+1 offsets into items / hub
    std::vector<uint32_t> n_hub;            // hub-chunk items at the front of each minibatch
    Item* d_items = nullptr;
    HubInfo* d_hub = nullptr;
    uint64_t cap_items = 0;
    uint32_t max_slots = 0;
};

struct f2v_engine {
    int device = 0;
    uint64_t n = 0, nnz = 0;
    uint32_t dim = 0;
    std::vector<uint64_t> h_rowptr;          // host copy: the plan is built from degrees
    uint64_t* d_rowptr = nullptr;
    uint32_t* d_colids = nullptr;
    float* d_Xall = nullptr;                 // one allocation: table 0 then table 1 (rows_alloc rows each)
    float* d_X[2] = {nullptr, nullptr};      // ping-pong tables inside d_Xall; cur holds the live embedding
    int cur = 0;
    uint64_t rows_alloc = 0;                 // rows allocated per table (n padded for all-gather)
    float* d_lut = nullptr;
    bool lut_set = false;
    uint32_t* d_neg = nullptr;
    uint64_t neg_cap = 0, neg_count = 0, neg_off = 0;
    uint32_t* d_walks = nullptr;
    bool walks_set = false;
    float* d_stage = nullptr;
    uint64_t stage_cap = 0;
    float* d_partials = nullptr;
    uint32_t* d_counters = nullptr;
    uint64_t slots_cap = 0;
    Plan epoch_plan, step_plan;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t copy_stream = nullptr;      // device->host row copies that overlap the epoch (f2v_run_epoch_host)
    cudaEvent_t ev_rows = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool ev_valid = false;
    int epoch_mode = 0;
    int variant = -1;                        // d=128 lane layout: -1 auto, see launch_batch
    int neg_smem = 1;
    int auto_flow = 0;                       // epoch mode 0: batches up to this size run the dataflow epoch kernel (0 = never)
    bool flow_used = false;                  // a dataflow launch reports a wait time-out through d_done[1]
    uint32_t min_chunk = 0;                  // lower bound of the adaptive hub chunk (0 = default_min_chunk(batch))
    int par = 9472;                          // adaptive-chunk target: 148 SMs x 64 lane groups (0 = fixed chunk)
    uint64_t launches = 0;
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    int sm_count = 0;
    // peer-store exchange (f2v_comm_peer_*): the other ranks' tables and flag pages, mapped
    // through CUDA IPC (other processes) or used directly (engines of this process)
    bool peer_mode = false;
    float* peerX[kMaxWorld][2] = {};
    uint64_t* peer_flags[kMaxWorld] = {};    // base of rank r's flag page
    bool peer_ipc[kMaxWorld] = {};           // mapping opened with cudaIpcOpenMemHandle
    uint64_t* d_flags = nullptr;             // local flag page, kMaxWorld * kFlagStride u64
    uint32_t* d_done = nullptr;
    uint32_t* d_flow = nullptr;              // dataflow epoch: [ticket (128-byte line)] [cnt nb] [done nb]
    uint64_t flow_cap = 0;
    uint64_t step_id = 0;                    // minibatch steps published so far (same on every rank)
    // NVLink multicast (NVLS) exchange: tables + flag page live in one VMM allocation bound to a
    // multicast object shared by all ranks; a store through the multicast mapping lands everywhere
    int want_mc = 1;                         // option "multicast": use NVLS when every rank supports it
    int mc_inproc = 0;                       // option "multicast_in_process": the caller drives every engine of this process
                                             // from its own host thread, so the (blocking) hand-off may run between them
    bool mc_mode = false;
    CUmemGenericAllocationHandle mc_handle = 0, vmm_handle = 0;
    CUdeviceptr vmm_uc = 0, vmm_mc = 0;
    size_t vmm_size = 0;
    int listen_sock = -1;
    char sock_name[48] = "";
    // row-sharded tables: one flat virtual range, shard s = physical memory of rank s (VMM), mapped
    // on every rank; see shard_row() in f2v_kernels.cuh
    int want_shard = 0;                      // option "sharded"
    bool shard_mode = false;
    uint32_t shard_lg = 0, shard_rows = 0;
    CUmemGenericAllocationHandle shard_handles[2][kMaxWorld] = {};   // [table][rank]
    CUdeviceptr shard_va = 0;
    size_t shard_bytes = 0;                  // bytes of one shard of one table
    float* d_shard_stage = nullptr;          // staging chunk for host <-> sharded table copies
    int exchange_timeout_ms = 30000;         // option "exchange_timeout_ms": a peer that never publishes its step
                                             // is reported as an error instead of hanging the launch (0 = wait for ever)
    int trace = 0;                           // option "trace": one CUDA event per minibatch of the last epoch
    std::vector<cudaEvent_t> trace_ev;
    uint64_t trace_n = 0;
    int peer_debug = 0;                      // timing probes only: 1 = no peer row stores, 2 = no flag barrier
    int order = -1;                          // item order after the hub chunks: 0 descending degree, 1 light rows first,
                                             // 2 light rows interleaved; -1 = default (0 on one GPU, 1 on several)
    int pdl = 2;                             // programmatic dependent launch of consecutive minibatches (0 off, 1, 2)
    int peer_sig = 2;                        // who publishes a minibatch's exchange step: 2 (default) = CTA 0 of the NEXT
                                             // launch, after its dependency wait (launches stay PDL-chained); 1 = a 1-CTA
                                             // kernel after the force kernel; 0 = the force kernel's last CTA (a system
                                             // fence per CTA: measured slower)
};

// What a rank publishes for the peer-store exchange (f2v_comm_peer_export): fits F2V_PEER_BLOB.
struct PeerBlob {
    uint32_t magic;
    int32_t device;
    uint64_t pid;
    uint64_t n, dim, cur;
    uint64_t rows_alloc;
    uint64_t ptr[2];                         // tables (X[0] then X[1]), flags (valid inside process `pid`)
    cudaIpcMemHandle_t h[2];
    uint32_t shard;                          // this rank wants row-sharded tables (option "sharded")
    uint32_t mc_ok;                          // this rank can and wants to use NVLink multicast
    char sock_name[40];                      // abstract unix socket this rank listens on (fd hand-off)
};
static_assert(sizeof(PeerBlob) <= F2V_PEER_BLOB, "PeerBlob must fit the ABI's blob size");
constexpr uint32_t kPeerMagic = 0x46325650u;

static int ensure(void** p, uint64_t* cap, uint64_t need_bytes) {
    if (*cap >= need_bytes && *p) return F2V_OK;
    if (*p) CU(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    CU(cudaMalloc(p, need_bytes ? need_bytes : 16));
    *cap = need_bytes;
    return F2V_OK;
}

// Both tables in one allocation (combined-row addressing in the kernels).  Keeps the live table's
// first e->n rows when reallocating for a larger row count.
static int alloc_tables(f2v_engine* e, uint64_t rows) {
    if (2 * rows > 0xffffffffull) return fail(F2V_ERR_ARG, "table of %llu rows is too large for 32-bit combined row ids", (unsigned long long)rows);
    float* all = nullptr;
    const size_t tbl = sizeof(float) * rows * e->dim;
    CU(cudaMalloc((void**)&all, 2 * tbl));
    // on the engine's stream: it is a non-blocking stream, so a memset on the legacy default stream
    // would NOT be ordered before the uploads and kernels that follow
    CU(cudaMemsetAsync(all, 0, 2 * tbl, e->stream));
    if (e->d_Xall) {
        CU(cudaMemcpyAsync(all + (size_t)e->cur * rows * e->dim, e->d_X[e->cur], sizeof(float) * e->n * e->dim,
                           cudaMemcpyDeviceToDevice, e->stream));
        CU(cudaStreamSynchronize(e->stream));
        CU(cudaFree(e->d_Xall));
    }
    e->d_Xall = all;
    e->d_X[0] = all;
    e->d_X[1] = all + rows * e->dim;
    e->rows_alloc = rows;
    return F2V_OK;
}

static int ensure_tables(f2v_engine* e) {
    if (e->d_Xall) return F2V_OK;
    return alloc_tables(e, e->rows_alloc ? e->rows_alloc : e->n);
}

// Upload the host plan (f2v_plan.hpp) for rows [first_row, first_row+nrows) and size the
// hub-row partial buffers.  Cached on (batch, chunk, walk, rank, world, range).
static int build_plan(f2v_engine* e, Plan& pl, uint64_t first_row, uint64_t nrows, uint32_t batch,
                      uint32_t chunk, uint32_t par, bool walk, int rank, int world, int assign) {
    if (pl.d_items && pl.batch == batch && pl.chunk == chunk && pl.par == par && pl.min_chunk == e->min_chunk &&
        pl.walk == walk && pl.rank == rank &&
        pl.world == world && pl.assign == assign && pl.first_row == first_row && pl.nrows == nrows)
        return F2V_OK;
    HostPlan hp;
    build_host_plan(e->h_rowptr.data(), first_row, nrows, batch, chunk, par, walk, rank, world, assign, hp, e->min_chunk);
    const uint64_t nb = hp.nb, total = hp.items.size();
    std::vector<Item>& items = hp.items;
    std::vector<HubInfo>& hub = hp.hub;
    std::vector<uint64_t>& item_ptr = hp.item_ptr;
    std::vector<uint32_t>& n_hub = hp.n_hub;
    const uint32_t max_slots = hp.max_slots;
    if (pl.cap_items < total || !pl.d_items) {
        if (pl.d_items) CU(cudaFree(pl.d_items));
        if (pl.d_hub) CU(cudaFree(pl.d_hub));
        pl.d_items = nullptr; pl.d_hub = nullptr; pl.cap_items = 0;
        CU(cudaMalloc((void**)&pl.d_items, sizeof(Item) * (total ? total : 1)));
        CU(cudaMalloc((void**)&pl.d_hub, sizeof(HubInfo) * (total ? total : 1)));
        pl.cap_items = total;
    }
    // plans are (re)built rarely; a synchronous copy keeps the host vectors' lifetime simple
    CU(cudaStreamSynchronize(e->stream));
    if (total) {
        CU(cudaMemcpyAsync(pl.d_items, items.data(), sizeof(Item) * total, cudaMemcpyHostToDevice, e->stream));
        CU(cudaMemcpyAsync(pl.d_hub, hub.data(), sizeof(HubInfo) * total, cudaMemcpyHostToDevice, e->stream));
    }
    // on the engine's (non-blocking) stream, and finished before the host vectors go away: a copy on
    // the legacy stream would not be ordered before the launches that read the plan
    CU(cudaStreamSynchronize(e->stream));
    if (max_slots > e->slots_cap || !e->d_partials) {
        if (e->d_partials) CU(cudaFree(e->d_partials));
        if (e->d_counters) CU(cudaFree(e->d_counters));
        e->d_partials = nullptr; e->d_counters = nullptr;
        uint64_t slots = std::max<uint64_t>(max_slots, 1);
        CU(cudaMalloc((void**)&e->d_partials, sizeof(float) * slots * e->dim));
        CU(cudaMalloc((void**)&e->d_counters, sizeof(uint32_t) * slots));
        CU(cudaMemsetAsync(e->d_counters, 0, sizeof(uint32_t) * slots, e->stream));
        e->slots_cap = slots;
    }
    pl.batch = batch; pl.chunk = chunk; pl.par = par; pl.min_chunk = e->min_chunk; pl.walk = walk; pl.rank = rank; pl.world = world;
    pl.assign = assign;
    pl.first_row = first_row; pl.nrows = nrows; pl.nb = nb;
    pl.item_ptr.swap(item_ptr);
    pl.n_hub.swap(n_hub);
    pl.max_slots = max_slots;
    return F2V_OK;
}

// ------------------------------------------------------------------ dispatch -----------
template <class L, int MODEL>
static cudaError_t launch_batch_t(const BatchParams& p, cudaStream_t st, int sm_count) {
    (void)sm_count;
    if (p.n_items == 0) return cudaSuccess;
    auto kern = force_batch_kernel<L, MODEL>;
    const bool negs = L::kBulk && p.neg_in_smem;
    size_t smem = 0;
    if (negs || L::kStages > 0) smem = 128 + (negs ? (size_t)p.s * p.dim * sizeof(float) : 0);
    if constexpr (L::kStages > 0) {
        smem += L::kCtaBytes;                        // the lane groups' asynchronous-copy rings
        static bool carve = false;                   // (per instantiation) all of the SM's L1/shared array as shared
        if (!carve) {                                // memory: rows bypass L1, MINB CTAs of ring must fit
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            if (e != cudaSuccess) return e;
            carve = true;
        }
    }
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const unsigned per_cta = kWarpsPerCta * L::G;
    const unsigned grid = (p.n_items + per_cta - 1) / per_cta;
    if (p.pdl) {
        // programmatic dependent launch: may be scheduled while the previous minibatch drains
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kWarpsPerCta * 32); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, kern, p);
    }
    kern<<<grid, kWarpsPerCta * 32, smem, st>>>(p);
    return cudaGetLastError();
}

template <class L>
static cudaError_t launch_batch_m(int model, const BatchParams& p, cudaStream_t st, int sm_count) {
    switch (model) {
    case kTDist: return launch_batch_t<L, kTDist>(p, st, sm_count);
    case kSigmoid: return launch_batch_t<L, kSigmoid>(p, st, sm_count);
    default: return launch_batch_t<L, kWalk>(p, st, sm_count);
    }
}

static cudaError_t launch_batch(int model, const BatchParams& p, cudaStream_t st, int sm_count) {
    switch (p.dim) {
    case 32: return launch_batch_m<VecL<32, 8, 8>>(model, p, st, sm_count);
    case 64:
        // auto (-1): small launches run 4 rows in flight per 8-lane group (104 registers: latency-bound);
        // large ones 2 rows in flight at 64 registers / 4 CTAs per SM, the walk model (every item has
        // 5 + s pairs: occupancy-bound) at 48 registers / 5 CTAs per SM -- measured on R-MAT 20/22:
        // option 7 2.42 -> 1.46 ms, option 6 1.41 -> 1.22 ms per epoch at batch 65536
        // (batch 16384, option 7: 2.82 / 1.90 / 2.09 ms for the three layouts; batch 4096: 5.76 / 6.72 / 7.03)
        switch (p.variant >= 0 ? p.variant : (p.n_items >= 48000u && model == kWalk ? 4 : (p.n_items >= 8000u ? 1 : 0))) {
        case 21: return launch_batch_m<RingL<64, 8, 3, 4>>(model, p, st, sm_count);
        case 22: return launch_batch_m<RingL<64, 8, 4, 3>>(model, p, st, sm_count);
        case 1: return launch_batch_m<VecL<64, 8, 2, 4>>(model, p, st, sm_count);
        case 4: return launch_batch_m<VecL<64, 8, 2, 5>>(model, p, st, sm_count);
        default: return launch_batch_m<VecL<64, 8, 4>>(model, p, st, sm_count);
        }
    case 128:
        // auto (-1): 4 CTAs/SM at 64 registers; launches with many items run 5 CTAs/SM at 48 registers
        // (a few spilled values, 25 % more gathered rows in flight: measured 5-9 % faster from ~64 K items,
        // slower below); small launches are latency-bound, not occupancy-bound, and run 8 rows in
        // flight per group at 128 registers (5-11 % faster below ~12 K items)
        switch (p.variant >= 0 ? p.variant : (p.n_items >= 48000u ? 8 : (p.n_items < 12000u ? 11 : 3))) {
        case 21: return launch_batch_m<RingL<128, 16, 3, 4>>(model, p, st, sm_count);
        case 22: return launch_batch_m<RingL<128, 16, 4, 3>>(model, p, st, sm_count);
        case 0: return launch_batch_m<VecL<128, 16, 4, 3>>(model, p, st, sm_count);
        case 8: return launch_batch_m<VecL<128, 16, 2, 5>>(model, p, st, sm_count);
        case 11: return launch_batch_m<VecL<128, 16, 8, 2>>(model, p, st, sm_count);
        default: return launch_batch_m<VecL<128, 16, 2, 4>>(model, p, st, sm_count);   // 3
        }
    case 256: return launch_batch_m<VecL<256, 32, 4>>(model, p, st, sm_count);
    default: break;
    }
    if (p.dim <= 32) return launch_batch_m<GenL<1>>(model, p, st, sm_count);
    if (p.dim <= 64) return launch_batch_m<GenL<2>>(model, p, st, sm_count);
    if (p.dim <= 128) return launch_batch_m<GenL<4>>(model, p, st, sm_count);
    if (p.dim <= 256) return launch_batch_m<GenL<8>>(model, p, st, sm_count);
    if (p.dim <= 512) return launch_batch_m<GenL<16>>(model, p, st, sm_count);
    return launch_batch_m<GenL<32>>(model, p, st, sm_count);
}

// ---- dataflow epoch kernel (epoch mode 2): ordinary launch, grid = SMs x resident CTAs
template <class L, int MODEL>
static cudaError_t launch_flow_k(const FlowParams& fp, cudaStream_t st, int sm_count) {
    auto kern = force_flow_kernel<L, MODEL>;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kWarpsPerCta * 32, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    const unsigned per_cta = kWarpsPerCta * L::G;
    const unsigned want = (fp.total_items + per_cta - 1) / per_cta;
    const unsigned grid = std::max(1u, std::min(want, (unsigned)(sm_count * per_sm)));
    kern<<<grid, kWarpsPerCta * 32, 0, st>>>(fp);
    return cudaGetLastError();
}
template <class L>
static cudaError_t launch_flow_m(int model, const FlowParams& fp, cudaStream_t st, int sm_count) {
    switch (model) {
    case kTDist: return launch_flow_k<L, kTDist>(fp, st, sm_count);
    case kSigmoid: return launch_flow_k<L, kSigmoid>(fp, st, sm_count);
    default: return launch_flow_k<L, kWalk>(fp, st, sm_count);
    }
}
static cudaError_t launch_flow(int model, const FlowParams& fp, cudaStream_t st, int sm_count, int variant) {
    // (with many minibatches in flight a warp's latency matters more than occupancy: the layouts with the
    // most rows in flight per lane group measured fastest here -- d=128: 8 rows at 128 registers)
    (void)variant;
    const uint32_t dim = fp.p.dim;
    switch (dim) {
    case 32: return launch_flow_m<VecL<32, 8, 8>>(model, fp, st, sm_count);
    case 64: return launch_flow_m<VecL<64, 8, 4>>(model, fp, st, sm_count);
    case 128: return launch_flow_m<VecL<128, 16, 8, 2>>(model, fp, st, sm_count);
    case 256: return launch_flow_m<VecL<256, 32, 4>>(model, fp, st, sm_count);
    default: break;
    }
    if (dim <= 32) return launch_flow_m<GenL<1>>(model, fp, st, sm_count);
    if (dim <= 64) return launch_flow_m<GenL<2>>(model, fp, st, sm_count);
    if (dim <= 128) return launch_flow_m<GenL<4>>(model, fp, st, sm_count);
    if (dim <= 256) return launch_flow_m<GenL<8>>(model, fp, st, sm_count);
    if (dim <= 512) return launch_flow_m<GenL<16>>(model, fp, st, sm_count);
    return launch_flow_m<GenL<32>>(model, fp, st, sm_count);
}

static bool bulk_ok(const f2v_engine* e, uint32_t s, int bs_mode) {
    if (bs_mode != 0 || s == 0 || !e->neg_smem) return false;
    if (!(e->dim == 32 || e->dim == 64 || e->dim == 128 || e->dim == 256)) return false;
    return 128 + (size_t)s * e->dim * sizeof(float) <= 200 * 1024;
}

static int check_model(const f2v_engine* e, int model, uint32_t s, int bs_mode) {
    if (!e) return fail(F2V_ERR_ARG, "null engine");
    if (model != F2V_TDIST && model != F2V_SIGMOID && model != F2V_WALK)
        return fail(F2V_ERR_ARG, "model must be 5, 6 or 7 (got %d)", model);
    if (bs_mode != 0 && bs_mode != 1) return fail(F2V_ERR_ARG, "bs_mode must be 0 or 1");
    if (model != F2V_TDIST && !e->lut_set) return fail(F2V_ERR_STATE, "sigmoid table not set (f2v_set_lut)");
    if (model == F2V_WALK && !e->walks_set) return fail(F2V_ERR_STATE, "walks not set (f2v_set_walks / f2v_sample_walks)");
    (void)s;
    return F2V_OK;
}

static uint64_t neg_stride(int model, uint32_t batch, uint32_t s, int bs_mode) {
    return (bs_mode && model != F2V_WALK) ? (uint64_t)batch + s - 1 : (uint64_t)s;
}

// Row-sharded engine: rows [first, first+count) of the live table <-> a host buffer in identity
// layout, through a device staging chunk (the flat table is a permutation spread over all GPUs).
static int shard_copy(f2v_engine* e, bool to_table, uint64_t first, uint64_t count, float* host) {
    if (!to_table && e->step_id) {
        // reading other ranks' shards: wait until every rank has published the current exchange step
        BatchParams p{};
        p.n_peers = (uint32_t)(e->world - 1);
        p.rank = (uint32_t)e->rank; p.world = (uint32_t)e->world;
        p.flags = e->d_flags; p.done = e->d_done;
        p.timeout_ns = (uint64_t)e->exchange_timeout_ms * 1000000ull; p.timed_out = e->d_done + 1;
        p.wait_step = e->step_id;
        peer_sync_kernel<<<1, 32, 0, e->stream>>>(p);
        CU(cudaGetLastError());
    }
    const uint64_t chunk_rows = std::max<uint64_t>(1, (64ull << 20) / (sizeof(float) * e->dim));
    if (!e->d_shard_stage) CU(cudaMalloc((void**)&e->d_shard_stage, sizeof(float) * chunk_rows * e->dim));
    float* table = e->d_X[e->cur];
    for (uint64_t a = 0; a < count; a += chunk_rows) {
        const uint64_t c = std::min(chunk_rows, count - a);
        const size_t bytes = sizeof(float) * c * e->dim;
        const unsigned grid = (unsigned)std::min<uint64_t>((c * e->dim + 255) / 256, (uint64_t)e->sm_count * 16);
        if (to_table) {
            CU(cudaMemcpyAsync(e->d_shard_stage, host + a * e->dim, bytes, cudaMemcpyHostToDevice, e->stream));
            shard_copy_kernel<true><<<grid, 256, 0, e->stream>>>(table, e->d_shard_stage, first + a, c, e->dim, e->shard_lg,
                                                                  e->shard_rows, (uint32_t)e->rank);
        } else {
            shard_copy_kernel<false><<<grid, 256, 0, e->stream>>>(table, e->d_shard_stage, first + a, c, e->dim, e->shard_lg,
                                                                   e->shard_rows, (uint32_t)e->rank);
            CU(cudaMemcpyAsync(host + a * e->dim, e->d_shard_stage, bytes, cudaMemcpyDeviceToHost, e->stream));
        }
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(e->stream));        // one staging chunk: finish before it is reused
    }
    return F2V_OK;
}

// Publish one exchange step outside an epoch (after a sharded upload / a broadcast of uploaded
// rows): every rank must do it the same number of times.
static int publish_exchange_step(f2v_engine* e) {
    BatchParams p{};
    p.n_peers = (uint32_t)(e->world - 1);
    p.rank = (uint32_t)e->rank; p.world = (uint32_t)e->world;
    p.flags = e->d_flags; p.done = e->d_done;
    if (e->mc_mode) {
        p.mc_flag = (uint64_t*)((char*)e->vmm_mc + ((char*)e->d_flags - (char*)e->d_Xall)) + (size_t)e->rank * kFlagStride;
    } else {
        for (int q = 0, k = 0; q < e->world; q++) {
            if (q == e->rank) continue;
            p.peer_flag[k++] = e->peer_flags[q] + (size_t)e->rank * kFlagStride;
        }
    }
    p.wait_step = 0;
    p.signal_step = ++e->step_id;
    peer_sync_kernel<<<1, 32, 0, e->stream>>>(p);
    CU(cudaGetLastError());
    e->launches++;
    return F2V_OK;
}

// ------------------------------------------------------------------ C ABI --------------
extern "C" {

const char* f2v_last_error(void) { return g_err; }
int f2v_abi_version(void) { return F2V_ABI_VERSION; }

int f2v_device_count(void) {
    int c = 0;
    cudaError_t r = cudaGetDeviceCount(&c);
    if (r != cudaSuccess) return fail(F2V_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(r));
    return c;
}

int f2v_create(f2v_engine** out, int device_id, uint64_t n, uint64_t nnz, const uint64_t* rowptr,
               const uint32_t* colids, uint32_t dim) {
    if (!out) return fail(F2V_ERR_ARG, "out is null");
    *out = nullptr;
    if (n < 2 || n > 0xffffffffull) return fail(F2V_ERR_ARG, "n must be in [2, 2^32)");
    if (dim < 1 || dim > 1024) return fail(F2V_ERR_ARG, "dim must be in [1, 1024]");
    if (!rowptr || (nnz > 0 && !colids)) return fail(F2V_ERR_ARG, "null CSR arrays");
    if (rowptr[0] != 0 || rowptr[n] != nnz) return fail(F2V_ERR_ARG, "rowptr[0] must be 0 and rowptr[n] == nnz");
    // (all host threads: at R-MAT 26 these are 67 M and 2.1 G entries; the first offending index is reported)
    int64_t bad_row = INT64_MAX, bad_col = INT64_MAX;
#pragma omp parallel for schedule(static) reduction(min : bad_row)
    for (int64_t i = 0; i < (int64_t)n; i++)
        if (rowptr[i + 1] < rowptr[i] && i < bad_row) bad_row = i;
    if (bad_row != INT64_MAX) return fail(F2V_ERR_ARG, "rowptr not monotone at row %llu", (unsigned long long)bad_row);
#pragma omp parallel for schedule(static) reduction(min : bad_col)
    for (int64_t k = 0; k < (int64_t)nnz; k++)
        if (colids[k] >= n && k < bad_col) bad_col = k;
    if (bad_col != INT64_MAX)
        return fail(F2V_ERR_ARG, "colids[%llu] = %u out of range", (unsigned long long)bad_col, colids[bad_col]);
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device_id < 0 || device_id >= ndev) return fail(F2V_ERR_ARG, "device %d not in [0,%d)", device_id, ndev);
    CU(cudaSetDevice(device_id));
    f2v_engine* e = new (std::nothrow) f2v_engine();
    if (!e) return fail(F2V_ERR_NOMEM, "out of host memory");
    e->device = device_id; e->n = n; e->nnz = nnz; e->dim = dim;
    // everything below can fail (out of device memory on a large graph): one exit path releases
    // what was created so far (f2v_destroy accepts a partly built engine)
    int rc = [&]() -> int {
        try { e->h_rowptr.assign(rowptr, rowptr + n + 1); } catch (const std::bad_alloc&) { return fail(F2V_ERR_NOMEM, "out of host memory"); }
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, device_id));
        e->sm_count = prop.multiProcessorCount;
        if (prop.major < 10) return fail(F2V_ERR_CUDA, "device is sm_%d%d; this library is built for sm_100a only", prop.major, prop.minor);
        CU(cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking));
        e->stream = e->own_stream;
        CU(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&e->ev_rows, cudaEventDisableTiming));
        CU(cudaEventCreate(&e->ev0));
        CU(cudaEventCreate(&e->ev1));
        CU(cudaMalloc((void**)&e->d_rowptr, sizeof(uint64_t) * (n + 1)));
        CU(cudaMalloc((void**)&e->d_colids, sizeof(uint32_t) * (nnz ? nnz : 1)));
        CU(cudaMemcpyAsync(e->d_rowptr, rowptr, sizeof(uint64_t) * (n + 1), cudaMemcpyHostToDevice, e->stream));
        if (nnz) CU(cudaMemcpyAsync(e->d_colids, colids, sizeof(uint32_t) * nnz, cudaMemcpyHostToDevice, e->stream));
        CU(cudaStreamSynchronize(e->stream));
        e->rows_alloc = n;                       // tables are allocated on first use (ensure_tables): a row-sharded
                                                 // engine never holds a full-size table
        CU(cudaMalloc((void**)&e->d_lut, sizeof(float) * kLutAlloc));
        return F2V_OK;
    }();
    if (rc != F2V_OK) {
        char keep[sizeof(g_err)];
        memcpy(keep, g_err, sizeof(keep));       // f2v_destroy must not clobber the message
        f2v_destroy(e);
        cudaGetLastError();
        memcpy(g_err, keep, sizeof(keep));
        return rc;
    }
    *out = e;
    return F2V_OK;
}

int f2v_destroy(f2v_engine* e) {
    if (!e) return F2V_OK;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    if (e->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(e->comm);
    for (int r = 0; r < kMaxWorld; r++) {
        if (!e->peer_ipc[r]) continue;
        if (e->peerX[r][0]) cudaIpcCloseMemHandle(e->peerX[r][0]);
        cudaIpcCloseMemHandle(e->peer_flags[r]);
    }
    if (e->listen_sock >= 0) close(e->listen_sock);
    if (e->mc_mode) {
        // tables and flags live in the VMM allocation: unmap / unbind / release instead of cudaFree
        CUdevice dev;
        g_drv.DeviceGet(&dev, e->device);
        g_drv.MemUnmap(e->vmm_mc, e->vmm_size);
        g_drv.MemAddressFree(e->vmm_mc, e->vmm_size);
        g_drv.MulticastUnbind(e->mc_handle, dev, 0, e->vmm_size);
        g_drv.MemUnmap(e->vmm_uc, e->vmm_size);
        g_drv.MemAddressFree(e->vmm_uc, e->vmm_size);
        g_drv.MemRelease(e->vmm_handle);
        g_drv.MemRelease(e->mc_handle);
        e->d_Xall = nullptr; e->d_flags = nullptr;
    }
    if (e->shard_mode) {
        const size_t total = 2 * (size_t)e->world * e->shard_bytes;
        g_drv.MemUnmap(e->shard_va, total);
        g_drv.MemAddressFree(e->shard_va, total);
        for (int s = 0; s < e->world; s++) { g_drv.MemRelease(e->shard_handles[0][s]); g_drv.MemRelease(e->shard_handles[1][s]); }
        e->d_Xall = nullptr;
    }
    cudaFree(e->d_shard_stage);
    cudaFree(e->d_flags); cudaFree(e->d_done);
    cudaFree(e->d_rowptr); cudaFree(e->d_colids); cudaFree(e->d_Xall);
    cudaFree(e->d_lut); cudaFree(e->d_neg); cudaFree(e->d_walks); cudaFree(e->d_stage);
    cudaFree(e->d_partials); cudaFree(e->d_counters);
    cudaFree(e->epoch_plan.d_items); cudaFree(e->epoch_plan.d_hub);
    cudaFree(e->step_plan.d_items); cudaFree(e->step_plan.d_hub);
    cudaFree(e->d_flow);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    if (e->ev_rows) cudaEventDestroy(e->ev_rows);
    for (cudaEvent_t ev : e->trace_ev) cudaEventDestroy(ev);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->own_stream) cudaStreamDestroy(e->own_stream);
    delete e;
    return F2V_OK;
}

int f2v_host_alloc(void** p, uint64_t bytes) {
    if (!p) return fail(F2V_ERR_ARG, "null argument");
    *p = nullptr;
    CU(cudaMallocHost(p, bytes ? bytes : 16));
    return F2V_OK;
}

int f2v_host_free(void* p) {
    if (p) CU(cudaFreeHost(p));
    return F2V_OK;
}

int f2v_host_register(void* p, uint64_t bytes) {
    if (!p || !bytes) return fail(F2V_ERR_ARG, "null argument");
    CU(cudaHostRegister(p, bytes, cudaHostRegisterDefault));
    return F2V_OK;
}

int f2v_host_unregister(void* p) {
    if (p) CU(cudaHostUnregister(p));
    return F2V_OK;
}

int f2v_device_memory(const f2v_engine* e, uint64_t* free_bytes, uint64_t* total_bytes) {
    if (!e) return fail(F2V_ERR_ARG, "null engine");
    CU(cudaSetDevice(e->device));
    size_t f = 0, t = 0;
    CU(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return F2V_OK;
}

int f2v_set_stream(f2v_engine* e, void* cuda_stream) {
    if (!e) return fail(F2V_ERR_ARG, "null engine");
    CU(cudaSetDevice(e->device));
    CU(cudaStreamSynchronize(e->stream));
    e->stream = cuda_stream ? (cudaStream_t)cuda_stream : e->own_stream;
    return F2V_OK;
}

// Multi-GPU: did a launch give up waiting for a peer's exchange step?  (stream already synchronised)
static int check_exchange(f2v_engine* e) {
    if ((!e->peer_mode && !e->flow_used) || !e->d_done) return F2V_OK;
    uint32_t flag = 0;
    CU(cudaMemcpy(&flag, e->d_done + 1, sizeof(flag), cudaMemcpyDeviceToHost));
    if (flag == 2)
        return fail(F2V_ERR_STATE, "dataflow epoch: a warp waited more than 20 s for a minibatch to complete (internal error); "
                                   "the tables are not valid");
    if (flag)
        return fail(F2V_ERR_STATE, "multi-GPU exchange timed out after %d ms: a peer rank did not publish its step "
                                   "(did every rank issue the same calls?); the tables are not valid", e->exchange_timeout_ms);
    return F2V_OK;
}

int f2v_sync(f2v_engine* e) {
    if (!e) return fail(F2V_ERR_ARG, "null engine");
    CU(cudaSetDevice(e->device));
    CU(cudaStreamSynchronize(e->stream));
    return check_exchange(e);
}

int f2v_set_embeddings(f2v_engine* e, const float* X) {
    if (!e || !X) return fail(F2V_ERR_ARG, "null argument");
    CU(cudaSetDevice(e->device));
    if (e->shard_mode) {
        // this rank fills its own shard; the other shards are filled by their ranks, so the upload
        // ends with one exchange step: the next epoch's first launch waits for every rank's step
        int r = shard_copy(e, true, 0, e->n, const_cast<float*>(X));
        if (r) return r;
        return publish_exchange_step(e);
    }
    { int r = ensure_tables(e); if (r) return r; }
    CU(cudaMemcpyAsync(e->d_X[e->cur], X, sizeof(float) * e->n * e->dim, cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return F2V_OK;
}

int f2v_get_embeddings(f2v_engine* e, float* X) {
    if (!e || !X) return fail(F2V_ERR_ARG, "null argument");
    CU(cudaSetDevice(e->device));
    if (e->shard_mode) {                                             // remote shards are read over NVLink
        int r = shard_copy(e, false, 0, e->n, X);
        return r ? r : check_exchange(e);
    }
    { int r = ensure_tables(e); if (r) return r; }
    CU(cudaMemcpyAsync(X, e->d_X[e->cur], sizeof(float) * e->n * e->dim, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return check_exchange(e);
}

int f2v_get_rows(f2v_engine* e, uint64_t first_row, uint64_t nrows, float* rows) {
    if (!e || !rows) return fail(F2V_ERR_ARG, "null argument");
    if (first_row + nrows > e->n) return fail(F2V_ERR_ARG, "row range out of bounds");
    CU(cudaSetDevice(e->device));
    if (e->shard_mode) return shard_copy(e, false, first_row, nrows, rows);
    { int r = ensure_tables(e); if (r) return r; }
    CU(cudaMemcpyAsync(rows, e->d_X[e->cur] + first_row * e->dim, sizeof(float) * nrows * e->dim,
                       cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return F2V_OK;
}

int f2v_checksum(f2v_engine* e, uint64_t* out) {
    if (!e || !out) return fail(F2V_ERR_ARG, "null argument");
    CU(cudaSetDevice(e->device));
    { int r = e->shard_mode ? F2V_OK : ensure_tables(e); if (r) return r; }
    if (e->shard_mode && e->step_id) {
        // other ranks' shards are read: every rank must have published the current exchange step
        BatchParams p{};
        p.n_peers = (uint32_t)(e->world - 1);
        p.rank = (uint32_t)e->rank; p.world = (uint32_t)e->world;
        p.flags = e->d_flags; p.done = e->d_done;
        p.timeout_ns = (uint64_t)e->exchange_timeout_ms * 1000000ull; p.timed_out = e->d_done + 1;
        p.wait_step = e->step_id;
        peer_sync_kernel<<<1, 32, 0, e->stream>>>(p);
        CU(cudaGetLastError());
    }
    unsigned long long* d_sum = nullptr;
    CU(cudaMalloc((void**)&d_sum, sizeof(unsigned long long)));
    CU(cudaMemsetAsync(d_sum, 0, sizeof(unsigned long long), e->stream));
    checksum_kernel<<<(unsigned)e->sm_count * 8, 256, 0, e->stream>>>(e->d_X[e->cur], e->n, e->dim, e->shard_lg, e->shard_rows, d_sum);
    cudaError_t le = cudaGetLastError();
    unsigned long long h = 0;
    if (le == cudaSuccess) le = cudaMemcpyAsync(&h, d_sum, sizeof(h), cudaMemcpyDeviceToHost, e->stream);
    if (le == cudaSuccess) le = cudaStreamSynchronize(e->stream);
    cudaFree(d_sum);
    if (le != cudaSuccess) return fail(F2V_ERR_CUDA, "f2v_checksum: %s", cudaGetErrorString(le));
    *out = (uint64_t)h;
    return check_exchange(e);
}

int f2v_set_lut(f2v_engine* e, const float* t, uint32_t count) {
    if (!e || !t) return fail(F2V_ERR_ARG, "null argument");
    if (count != kLutSize && count != kLutSize + 1) return fail(F2V_ERR_ARG, "sigmoid table must have 2048 entries");
    float h[kLutAlloc];
    memcpy(h, t, sizeof(float) * kLutSize);
    for (int i = kLutSize; i < kLutAlloc; i++) h[i] = 1.0f;   // v == 6.0f exactly: defined as 1.0
    CU(cudaSetDevice(e->device));
    CU(cudaMemcpyAsync(e->d_lut, h, sizeof(h), cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    e->lut_set = true;
    return F2V_OK;
}

int f2v_set_negatives(f2v_engine* e, const uint32_t* idx, uint64_t count) {
    if (!e || (!idx && count)) return fail(F2V_ERR_ARG, "null argument");
    CU(cudaSetDevice(e->device));
    if (e->neg_cap < count || !e->d_neg) {
        CU(cudaStreamSynchronize(e->stream));
        if (e->d_neg) CU(cudaFree(e->d_neg));
        e->d_neg = nullptr;
        e->neg_cap = 0;
        CU(cudaMalloc((void**)&e->d_neg, sizeof(uint32_t) * (count ? count : 1)));
        e->neg_cap = count;
    }
    if (count) CU(cudaMemcpyAsync(e->d_neg, idx, sizeof(uint32_t) * count, cudaMemcpyHostToDevice, e->stream));
    e->neg_count = count;
    e->neg_off = 0;
    return F2V_OK;
}

int f2v_set_negative_offset(f2v_engine* e, uint64_t offset) {
    if (!e) return fail(F2V_ERR_ARG, "null engine");
    if (offset > e->neg_count) return fail(F2V_ERR_ARG, "offset beyond the resident stream");
    e->neg_off = offset;
    return F2V_OK;
}

static int ensure_walks(f2v_engine* e) {
    if (!e->d_walks) CU(cudaMalloc((void**)&e->d_walks, sizeof(uint32_t) * e->n * kWalkLen));
    return F2V_OK;
}

int f2v_set_walks(f2v_engine* e, const uint32_t* walks) {
    if (!e || !walks) return fail(F2V_ERR_ARG, "null argument");
    CU(cudaSetDevice(e->device));
    int r = ensure_walks(e);
    if (r) return r;
    CU(cudaMemcpyAsync(e->d_walks, walks, sizeof(uint32_t) * e->n * kWalkLen, cudaMemcpyHostToDevice, e->stream));
    e->walks_set = true;
    return F2V_OK;
}

int f2v_get_walks(f2v_engine* e, uint32_t* walks) {
    if (!e || !walks) return fail(F2V_ERR_ARG, "null argument");
    if (!e->walks_set) return fail(F2V_ERR_STATE, "no walks resident");
    CU(cudaSetDevice(e->device));
    CU(cudaMemcpyAsync(walks, e->d_walks, sizeof(uint32_t) * e->n * kWalkLen, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return F2V_OK;
}

int f2v_sample_walks(f2v_engine* e, uint64_t seed, uint64_t epoch) {
    if (!e) return fail(F2V_ERR_ARG, "null engine");
    CU(cudaSetDevice(e->device));
    int r = ensure_walks(e);
    if (r) return r;
    unsigned grid = (unsigned)((e->n + 255) / 256);
    walk_kernel<<<grid, 256, 0, e->stream>>>(e->n, e->nnz, e->d_rowptr, e->d_colids, e->d_walks, seed, epoch);
    CU(cudaGetLastError());
    e->launches++;
    e->walks_set = true;
    return F2V_OK;
}

int f2v_step(f2v_engine* e, int model, uint64_t first_row, uint32_t nrows, const uint32_t* neg_idx,
             uint32_t s, int bs_mode, float lr, const uint32_t* walks) {
    if (!e) return fail(F2V_ERR_ARG, "null engine");
    if (model == F2V_WALK) bs_mode = 0;     // -bs is ignored by option 7 (Test/Force2Vec.cpp:148-150)
    if (e->shard_mode) return fail(F2V_ERR_STATE, "f2v_step is a single-engine test entry point; not available on a row-sharded engine");
    CU(cudaSetDevice(e->device));
    { int r0 = ensure_tables(e); if (r0) return r0; }
    if (model == F2V_WALK && walks) { int r = f2v_set_walks(e, walks); if (r) return r; }
    int r = check_model(e, model, s, bs_mode);
    if (r) return r;
    if (nrows == 0) return F2V_OK;
    if (first_row + nrows > e->n) return fail(F2V_ERR_ARG, "row range out of bounds");
    if (s > 0 && !neg_idx) return fail(F2V_ERR_ARG, "neg_idx is null");
    r = f2v_set_negatives(e, neg_idx, neg_stride(model, nrows, s, bs_mode));
    if (r) return r;
    r = build_plan(e, e->step_plan, first_row, nrows, nrows, 128, (uint32_t)e->par, model == F2V_WALK, 0, 1, 0);
    if (r) return r;
    uint64_t cap_bytes = e->stage_cap;
    r = ensure((void**)&e->d_stage, &cap_bytes, sizeof(float) * (uint64_t)nrows * e->dim);
    if (r) return r;
    e->stage_cap = cap_bytes;
    BatchParams p{};
    p.items = e->step_plan.d_items; p.hub = e->step_plan.d_hub;
    p.n_items = (uint32_t)e->step_plan.item_ptr[1]; p.n_hub = e->step_plan.n_hub[0];
    p.lo = first_row; p.split = 0;
    p.Xb = e->d_Xall; p.off_lo = p.off_hi = (uint32_t)((uint64_t)e->cur * e->rows_alloc);
    p.out = e->d_stage; p.out_base = first_row;
    p.colids = e->d_colids; p.neg = e->d_neg; p.walks = e->d_walks; p.lut = e->d_lut;
    p.partials = e->d_partials; p.counters = e->d_counters;
    p.s = s; p.dim = e->dim; p.bs_mode = bs_mode; p.neg_in_smem = bulk_ok(e, s, bs_mode) ? 1 : 0;
    p.lr = lr; p.variant = e->variant;
    CU(launch_batch(model, p, e->stream, e->sm_count));
    e->launches++;
    // apply after the join (algorithms.cpp:629-639 / :913-921)
    CU(cudaMemcpyAsync(e->d_X[e->cur] + first_row * e->dim, e->d_stage, sizeof(float) * (uint64_t)nrows * e->dim,
                       cudaMemcpyDeviceToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return F2V_OK;
}

// One epoch.  X_out_host != nullptr (single GPU): finished rows are copied to the host buffer on a
// second stream while later minibatches still compute (a row is final once its minibatch is done).
static int run_epoch_impl(f2v_engine* e, int model, uint32_t batch, uint32_t s, int bs_mode, float lr, uint32_t chunk,
                          float* X_out_host) {
    if (!e) return fail(F2V_ERR_ARG, "null engine");
    if (model == F2V_WALK) bs_mode = 0;
    int r = check_model(e, model, s, bs_mode);
    if (r) return r;
    if (batch == 0) return fail(F2V_ERR_ARG, "batch must be > 0");
    if (e->world > 1 && !e->peer_mode && batch % e->world)
        return fail(F2V_ERR_ARG, "batch (%u) must be a multiple of the world size (%d)", batch, e->world);
    CU(cudaSetDevice(e->device));
    // default hub chunk (upper bound of the adaptive chunk length): 256 edges for rows of >= 512 bytes in batches
    // of >= 16 K rows, else 128; 64 on a multi-GPU engine whose share of a minibatch is small enough (< 16 K rows per rank)
    // for the longest item to be the critical path.  Measured (profiles/r2_tune.md section 8): R-MAT 20 d=128
    // B=65536: 1.76 / 1.65 / 1.74 / 1.80 ms at 128 / 256 / 512 / 1024; R-MAT 24 d=128: 40.8 / 39.2 / 38.9 / 37.9;
    // R-MAT 22 d=64: 4.73 ms at 128, 6.52 at 1024; R-MAT 24, N=2: 24.6 at 128, 26.9 at 64, 29.1 at 32.
    if (chunk == 0) {
        chunk = (e->dim >= 128 && batch >= 16384u) ? 256 : 128;     // (small batches: R-MAT 20, B=256: 33 ms at 128, 36 at 256)
        if (e->world > 1 && batch / (uint32_t)e->world < 16384u) chunk = 64;
    }
    const uint64_t nb = (e->n + batch - 1) / batch;
    const uint64_t W = neg_stride(model, batch, s, bs_mode);
    if (e->neg_count < e->neg_off + nb * W)
        return fail(F2V_ERR_STATE, "negative stream too short: have %llu (offset %llu), epoch needs %llu",
                    (unsigned long long)e->neg_count, (unsigned long long)e->neg_off, (unsigned long long)(nb * W));
    // tables: the all-gather works on whole minibatches, so pad the row count to nb*batch
    const uint64_t rows_needed = (e->world > 1 && !e->peer_mode) ? nb * batch : e->n;
    if (!e->shard_mode && e->rows_alloc < rows_needed) {
        CU(cudaStreamSynchronize(e->stream));
        r = alloc_tables(e, rows_needed);
        if (r) return r;
    }
    r = ensure_tables(e);
    if (r) return r;
    const int order = e->order >= 0 ? e->order : (e->peer_mode ? 1 : 0);
    // dataflow epoch (mode 2, or mode 0 = automatic at small batches): single-GPU engines, at least two minibatches
    const bool flow = e->world == 1 && nb >= 2 && nb * (uint64_t)batch < 0xffffffffull &&
                      (e->epoch_mode == 2 || (e->epoch_mode == 0 && e->auto_flow && batch <= (uint32_t)e->auto_flow));
    r = build_plan(e, e->epoch_plan, 0, e->n, batch, chunk, (uint32_t)e->par, model == F2V_WALK, e->rank, e->world,
                   flow ? kPlanFlow
                        : ((e->peer_mode ? kAssignBalanced : kAssignSlices) | (order == 1 ? kOrderLightFirst : 0) | (order == 2 ? kOrderInterleave : 0)));
    if (r) return r;
    const Plan& pl = e->epoch_plan;
    float* Xnew = e->d_X[1 - e->cur];
    CU(cudaEventRecord(e->ev0, e->stream));
    BatchParams p{};
    p.Xb = e->d_Xall; p.out = Xnew; p.out_base = 0;
    p.off_lo = (uint32_t)((uint64_t)(1 - e->cur) * e->rows_alloc);      // rows already updated this epoch: next table
    p.off_hi = (uint32_t)((uint64_t)e->cur * e->rows_alloc);            // the others: current table
    p.colids = e->d_colids; p.walks = e->d_walks; p.lut = e->d_lut;
    p.partials = e->d_partials; p.counters = e->d_counters;
    p.s = s; p.dim = e->dim; p.bs_mode = bs_mode; p.neg_in_smem = bulk_ok(e, s, bs_mode) ? 1 : 0;
    p.lr = lr; p.variant = e->variant;
    const uint64_t slice = batch / (uint64_t)e->world;
    if (e->peer_mode) {
        p.n_peers = (e->peer_debug & 2) ? 0u : (uint32_t)(e->world - 1);
        p.n_store = ((e->peer_debug & 1) || e->mc_mode || e->shard_mode) ? 0u : (uint32_t)(e->world - 1);
        if (e->shard_mode) { p.shard_lg = e->shard_lg; p.shard_rows = e->shard_rows; }
        if (e->mc_mode) {
            // the same offsets inside the multicast mapping: one store reaches every replica
            p.mc_out = (float*)((char*)e->vmm_mc + ((char*)Xnew - (char*)e->d_Xall));
            p.mc_flag = (uint64_t*)((char*)e->vmm_mc + ((char*)e->d_flags - (char*)e->d_Xall)) + (size_t)e->rank * kFlagStride;
        }
        p.rank = (uint32_t)e->rank; p.world = (uint32_t)e->world;
        p.flags = e->d_flags; p.done = e->d_done;
        p.timeout_ns = (uint64_t)e->exchange_timeout_ms * 1000000ull; p.timed_out = e->d_done + 1;
        for (int r = 0, k = 0; r < e->world && !e->mc_mode; r++) {
            if (r == e->rank) continue;
            p.peer_out[k] = e->shard_mode ? nullptr : e->peerX[r][1 - e->cur];
            p.peer_flag[k] = e->peer_flags[r] + (size_t)e->rank * kFlagStride;
            k++;
        }
    }
    if (flow) {
        // dataflow epoch: one ordinary launch, minibatches overlap, readers wait for their writers only
        const uint64_t words = 32 + 2 * nb;
        if (e->flow_cap < words || !e->d_flow) {
            if (e->d_flow) CU(cudaFree(e->d_flow));
            e->d_flow = nullptr; e->flow_cap = 0;
            CU(cudaMalloc((void**)&e->d_flow, sizeof(uint32_t) * words));
            e->flow_cap = words;
        }
        CU(cudaMemsetAsync(e->d_flow, 0, sizeof(uint32_t) * words, e->stream));
        FlowParams fp{};
        fp.p = p;
        fp.p.items = pl.d_items; fp.p.hub = pl.d_hub; fp.p.neg = e->d_neg + e->neg_off;
        fp.p.n_items = (uint32_t)pl.item_ptr[nb];
        fp.p.neg_in_smem = 0;                    // a CTA works on several minibatches at once: negatives come from L2
        fp.p.pdl = 0; fp.p.late_wait = 0; fp.p.wait_step = 0; fp.p.signal_step = 0;
        fp.p.flow_cnt = e->d_flow + 32;
        fp.p.flow_done = e->d_flow + 32 + nb;
        fp.p.flow_batch = batch; fp.p.flow_nb = (uint32_t)nb; fp.p.flow_n = (uint32_t)e->n;
        fp.p.timeout_ns = 20ull * 1000000000ull;     // a wait this long is a bug: report it, do not hang the GPU
        if (!e->d_done) { CU(cudaMalloc((void**)&e->d_done, sizeof(uint32_t) * 32)); CU(cudaMemsetAsync(e->d_done, 0, sizeof(uint32_t) * 32, e->stream)); }
        fp.p.timed_out = e->d_done + 1;
        fp.total_items = (uint32_t)pl.item_ptr[nb];
        fp.neg_stride = (uint32_t)W;
        fp.ticket = e->d_flow;
        if (fp.total_items) {
            CU(launch_flow(model, fp, e->stream, e->sm_count, e->variant));
            e->launches++;
        }
        if (X_out_host)
            CU(cudaMemcpyAsync(X_out_host, Xnew, sizeof(float) * e->n * e->dim, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaEventRecord(e->ev1, e->stream));
        e->ev_valid = true;
        e->flow_used = true;
        e->cur = 1 - e->cur;
        return F2V_OK;
    }
    uint64_t copy_lo = 0;
    if (e->trace) {
        while (e->trace_ev.size() < nb + 1) { cudaEvent_t ev; CU(cudaEventCreate(&ev)); e->trace_ev.push_back(ev); }
        CU(cudaEventRecord(e->trace_ev[0], e->stream));
        e->trace_n = nb;
    }
    bool first_launch = true;
    for (uint64_t b = 0; b < nb; b++) {
        p.items = pl.d_items + pl.item_ptr[b];
        p.hub = pl.d_hub + pl.item_ptr[b];
        p.n_items = (uint32_t)(pl.item_ptr[b + 1] - pl.item_ptr[b]);
        p.n_hub = pl.n_hub[b];
        p.lo = b * batch;
        p.split = (uint32_t)(b * batch);
        p.neg = e->d_neg + e->neg_off + b * W;
        // the epoch's first launch is an ordinary one: it follows host copies (negatives, walks, the
        // table) and must see a freshly invalidated L1; launches 1.. are programmatic dependents
        // (multi-GPU with the NCCL all-gather: a collective sits between two launches, no chaining)
        const bool chain = e->pdl && (e->world == 1 || (e->peer_mode && e->peer_sig == 2)) && !first_launch;
        p.pdl = chain ? (e->pdl >= 2 ? 2 : 1) : 0;
        if (e->peer_mode) {
            // minibatch b reads rows its peers stored during step_id (minibatch b-1); it publishes step_id+1
            p.wait_step = (e->peer_debug & 2) ? 0 : e->step_id;
            p.signal_step = ++e->step_id;
        }
        if (e->peer_mode && e->peer_sig == 2) {
            // "successor publishes": this launch's CTA 0 publishes the PREVIOUS minibatch's step once the
            // previous launch is complete (stream order or griddepcontrol.wait -- the kernel boundary has
            // performed its peer stores), every warp waits for the peers' flags after its own dependency
            // wait.  One launch per minibatch, no system-scope fence per CTA.  A minibatch with no rows
            // on this rank launches nothing: its step is implied by the next publication (flags are
            // monotone).
            p.publish_step = p.wait_step;
            p.signal_step = 0;
            p.late_wait = p.pdl != 0;        // the first (ordinary) launch waits at kernel entry: its own rows of the
                                             // current table may just have been uploaded / broadcast by a peer
            p.flow_batch = batch;            // (row -> exchange step mapping of the lazy wait)
            if (p.n_items) {
                CU(launch_batch(model, p, e->stream, e->sm_count));
                e->launches++;
                first_launch = false;
            }
        } else if (p.n_items) {
            const uint64_t sig = p.signal_step;
            if (e->peer_mode && e->peer_sig) p.signal_step = 0;
            CU(launch_batch(model, p, e->stream, e->sm_count));
            e->launches++;
            first_launch = false;
            if (e->peer_mode && e->peer_sig && p.n_peers) {
                // the force kernel has drained (all its peer stores are performed): publish the step
                BatchParams q = p;
                q.wait_step = 0; q.signal_step = sig;
                peer_sync_kernel<<<1, 32, 0, e->stream>>>(q);
                CU(cudaGetLastError());
                e->launches++;
            }
        } else if (e->peer_mode) {
            peer_sync_kernel<<<1, 32, 0, e->stream>>>(p);
            CU(cudaGetLastError());
            e->launches++;
        }
        if (e->trace) CU(cudaEventRecord(e->trace_ev[b + 1], e->stream));
        if (X_out_host) {
            // rows [copy_lo, hi) are final: hand them to the copy stream in pieces of >= 8 MiB
            const uint64_t hi = std::min<uint64_t>((b + 1) * (uint64_t)batch, e->n);
            if ((hi - copy_lo) * e->dim * sizeof(float) >= (8u << 20) || b + 1 == nb) {
                CU(cudaEventRecord(e->ev_rows, e->stream));
                CU(cudaStreamWaitEvent(e->copy_stream, e->ev_rows, 0));
                CU(cudaMemcpyAsync(X_out_host + copy_lo * e->dim, Xnew + copy_lo * e->dim,
                                   sizeof(float) * (hi - copy_lo) * e->dim, cudaMemcpyDeviceToHost, e->copy_stream));
                copy_lo = hi;
            }
        }
        if (e->world > 1 && !e->peer_mode) {
            // exchange the updated slices before the next minibatch reads them
            float* base = Xnew + b * batch * e->dim;
            NC(g_nccl.AllGather(base + (uint64_t)e->rank * slice * e->dim, base, slice * e->dim,
                                kNcclFloat32, e->comm, e->stream));
        }
    }
    if (e->peer_mode) {
        // the replica is complete once every peer has published the epoch's last step (this rank's own
        // last step is published here when the successor-publishes scheme is on: the last force launch
        // is complete in stream order)
        p.wait_step = (e->peer_debug & 2) ? 0 : e->step_id;
        p.signal_step = 0;
        p.publish_step = e->peer_sig == 2 ? e->step_id : 0;
        p.late_wait = 0; p.pdl = 0;
        peer_sync_kernel<<<1, 32, 0, e->stream>>>(p);
        CU(cudaGetLastError());
        e->launches++;
    }
    CU(cudaEventRecord(e->ev1, e->stream));
    e->ev_valid = true;
    e->cur = 1 - e->cur;
    return F2V_OK;
}

int f2v_run_epoch(f2v_engine* e, int model, uint32_t batch, uint32_t s, int bs_mode, float lr, uint32_t chunk) {
    return run_epoch_impl(e, model, batch, s, bs_mode, lr, chunk, nullptr);
}

int f2v_run_epoch_host(f2v_engine* e, int model, uint32_t batch, uint32_t s, int bs_mode, float lr,
                       uint32_t chunk, const float* X_in, const uint32_t* neg, uint64_t neg_count,
                       const uint32_t* walks, float* X_out) {
    if (!e) return fail(F2V_ERR_ARG, "null engine");
    CU(cudaSetDevice(e->device));
    int r;
    if (e->shard_mode) {
        if (X_in) { r = f2v_set_embeddings(e, X_in); if (r) return r; }
        if (neg) { r = f2v_set_negatives(e, neg, neg_count); if (r) return r; }
        if (walks) { r = f2v_set_walks(e, walks); if (r) return r; }
        r = run_epoch_impl(e, model, batch, s, bs_mode, lr, chunk, nullptr);
        if (r) return r;
        CU(cudaStreamSynchronize(e->stream));
        if (X_out) { r = f2v_get_embeddings(e, X_out); if (r) return r; }
        return check_exchange(e);
    }
    { r = ensure_tables(e); if (r) return r; }
    // multi-GPU with the peer exchange: rank r moves only rows [lo, hi) = its 1/world share of the
    // table over PCIe, in both directions; the other replicas get them over NVLink
    const bool sharded = e->peer_mode && e->world > 1;
    const uint64_t per = (e->n + (uint64_t)e->world - 1) / (uint64_t)e->world;
    const uint64_t lo = sharded ? std::min<uint64_t>(e->n, (uint64_t)e->rank * per) : 0;
    const uint64_t hi = sharded ? std::min<uint64_t>(e->n, lo + per) : e->n;
    if (X_in && hi > lo)
        CU(cudaMemcpyAsync(e->d_X[e->cur] + lo * e->dim, X_in + lo * e->dim, sizeof(float) * (hi - lo) * e->dim,
                           cudaMemcpyHostToDevice, e->stream));
    if (X_in && sharded) {
        BcastParams b{};
        const size_t off = (size_t)((e->d_X[e->cur] + lo * e->dim) - e->d_Xall);
        b.src = e->d_Xall + off;
        const uint64_t floats = (hi - lo) * e->dim;
        const bool vec = (off % 4 == 0) && (floats % 4 == 0);
        b.count = vec ? floats / 4 : floats;
        if (e->mc_mode) {
            b.mc = (float*)e->vmm_mc + off;      // also rewrites this rank's own rows with the same values
        } else {
            for (int q = 0, k = 0; q < e->world; q++) {
                if (q == e->rank) continue;
                b.peer[k++] = e->peerX[q][0] + off;
            }
            b.n_store = (uint32_t)(e->world - 1);
        }
        if (b.count) {
            const unsigned grid = (unsigned)std::min<uint64_t>((b.count + 255) / 256, (uint64_t)e->sm_count * 8);
            if (vec) bcast_rows_kernel<true><<<grid, 256, 0, e->stream>>>(b);
            else bcast_rows_kernel<false><<<grid, 256, 0, e->stream>>>(b);
            CU(cudaGetLastError());
            e->launches++;
        }
        // publish: one exchange step; the epoch's first minibatch waits for every rank's rows
        r = publish_exchange_step(e);
        if (r) return r;
    }
    if (neg) { r = f2v_set_negatives(e, neg, neg_count); if (r) return r; }
    if (walks) { r = f2v_set_walks(e, walks); if (r) return r; }
    // single GPU: the download is pipelined behind the minibatches; multi-GPU: a replica is complete
    // only after the epoch's last exchange, so it is copied afterwards
    const bool overlap = X_out && e->world == 1;
    r = run_epoch_impl(e, model, batch, s, bs_mode, lr, chunk, overlap ? X_out : nullptr);
    if (r) return r;
    if (X_out && !overlap && hi > lo)
        CU(cudaMemcpyAsync(X_out + lo * e->dim, e->d_X[e->cur] + lo * e->dim, sizeof(float) * (hi - lo) * e->dim,
                           cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    if (overlap) CU(cudaStreamSynchronize(e->copy_stream));
    return check_exchange(e);
}

int f2v_set_epoch_mode(f2v_engine* e, int mode) {
    if (!e) return fail(F2V_ERR_ARG, "null engine");
    if (mode == 1) return fail(F2V_ERR_ARG, "epoch mode 1 (persistent kernel with a grid barrier per minibatch) was removed: "
                                            "it lost to mode 0 at every batch size; mode 2 is the one-launch-per-epoch mode");
    if (mode != 0 && mode != 2) return fail(F2V_ERR_ARG, "epoch mode %d not available", mode);
    e->epoch_mode = mode;
    return F2V_OK;
}

int f2v_set_option(f2v_engine* e, const char* name, int64_t value) {
    if (!e || !name) return fail(F2V_ERR_ARG, "null argument");
    if (!strcmp(name, "variant")) e->variant = (int)value;
    else if (!strcmp(name, "neg_smem")) e->neg_smem = value != 0;
    else if (!strcmp(name, "par")) e->par = (int)value;
    else if (!strcmp(name, "peer_debug")) e->peer_debug = (int)value;
    else if (!strcmp(name, "peer_sig")) e->peer_sig = (int)value;
    else if (!strcmp(name, "pdl")) e->pdl = (int)value;
    else if (!strcmp(name, "order")) e->order = (int)value;
    else if (!strcmp(name, "multicast")) e->want_mc = (int)value;
    else if (!strcmp(name, "multicast_in_process")) e->mc_inproc = value != 0;
    else if (!strcmp(name, "trace")) e->trace = (int)value;
    else if (!strcmp(name, "exchange_timeout_ms")) e->exchange_timeout_ms = (int)std::max<int64_t>(0, value);
    else if (!strcmp(name, "sharded")) e->want_shard = value != 0;
    else if (!strcmp(name, "auto_flow")) e->auto_flow = (int)std::max<int64_t>(0, value);
    else if (!strcmp(name, "min_chunk")) e->min_chunk = (uint32_t)std::max<int64_t>(0, value);   // part of the plan cache key
    else return fail(F2V_ERR_ARG, "unknown option %s", name);
    return F2V_OK;
}

uint64_t f2v_launch_count(const f2v_engine* e) { return e ? e->launches : 0; }

int f2v_last_epoch_ms(f2v_engine* e, float* ms) {
    if (!e || !ms) return fail(F2V_ERR_ARG, "null argument");
    if (!e->ev_valid) return fail(F2V_ERR_STATE, "no epoch has run");
    CU(cudaSetDevice(e->device));
    CU(cudaEventSynchronize(e->ev1));
    CU(cudaEventElapsedTime(ms, e->ev0, e->ev1));
    return F2V_OK;
}

uint32_t f2v_shard_row(uint32_t vertex, uint32_t log2_world, uint32_t shard_rows) {
    return shard_row(vertex, log2_world, shard_rows);
}

int f2v_trace_ms(f2v_engine* e, float* ms, uint32_t cap, uint32_t* count) {
    if (!e || !ms || !count) return fail(F2V_ERR_ARG, "null argument");
    if (!e->trace || e->trace_n == 0) return fail(F2V_ERR_STATE, "no traced epoch (f2v_set_option trace 1, epoch mode 0)");
    CU(cudaSetDevice(e->device));
    CU(cudaStreamSynchronize(e->stream));
    const uint32_t n = (uint32_t)std::min<uint64_t>(cap, e->trace_n);
    for (uint32_t b = 0; b < n; b++) CU(cudaEventElapsedTime(ms + b, e->trace_ev[b], e->trace_ev[b + 1]));
    *count = n;
    return F2V_OK;
}

int f2v_comm_unique_id(void* id128) {
    if (!id128) return fail(F2V_ERR_ARG, "null argument");
    int r = nccl_load();
    if (r) return r;
    ncclUniqueId id;
    NC(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return F2V_OK;
}

int f2v_comm_init(f2v_engine* e, const void* id128, int rank, int world) {
    if (!e || !id128) return fail(F2V_ERR_ARG, "null argument");
    if (world < 1 || rank < 0 || rank >= world) return fail(F2V_ERR_ARG, "bad rank/world");
    if (e->comm || e->peer_mode) return fail(F2V_ERR_STATE, "communicator already initialised");
    int r = nccl_load();
    if (r) return r;
    CU(cudaSetDevice(e->device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    NC(g_nccl.CommInitRank(&e->comm, world, id, rank));
    e->rank = rank;
    e->world = world;
    return F2V_OK;
}

int f2v_comm_peer_export(f2v_engine* e, void* blob) {
    if (!e || !blob) return fail(F2V_ERR_ARG, "null argument");
    if (e->world > 1) return fail(F2V_ERR_STATE, "communicator already initialised");
    CU(cudaSetDevice(e->device));
    CU(cudaStreamSynchronize(e->stream));
    if (!e->d_flags) {
        CU(cudaMalloc((void**)&e->d_flags, sizeof(uint64_t) * kMaxWorld * kFlagStride));
        CU(cudaMemsetAsync(e->d_flags, 0, sizeof(uint64_t) * kMaxWorld * kFlagStride, e->stream));
        CU(cudaMalloc((void**)&e->d_done, sizeof(uint32_t) * 32));
        CU(cudaMemsetAsync(e->d_done, 0, sizeof(uint32_t) * 32, e->stream));
        CU(cudaStreamSynchronize(e->stream));        // peers may store into the flag page as soon as it is exported
    }
    PeerBlob b;
    memset(&b, 0, sizeof(b));
    b.magic = kPeerMagic; b.device = e->device; b.pid = (uint64_t)getpid();
    b.n = e->n; b.dim = e->dim; b.cur = (uint64_t)e->cur;
    b.rows_alloc = e->rows_alloc;
    b.shard = e->want_shard ? 1u : 0u;
    if (!e->want_shard) { int r = ensure_tables(e); if (r) return r; }
    void* ptrs[2] = {e->d_Xall, e->d_flags};
    for (int k = e->want_shard ? 1 : 0; k < 2; k++) {      // a sharded engine exports its flag page only
        b.ptr[k] = (uint64_t)(uintptr_t)ptrs[k];
        CU(cudaIpcGetMemHandle(&b.h[k], ptrs[k]));
    }
    // NVLink multicast: can this device do it, and is a hand-off socket up?
    b.mc_ok = 0;
    if ((e->want_mc || e->want_shard) && drv_load() == F2V_OK) {
        CUdevice dev;
        int mc = 0, posix = 0;
        if (g_drv.DeviceGet(&dev, e->device) == CUDA_SUCCESS &&
            g_drv.DeviceGetAttribute(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) == CUDA_SUCCESS &&
            g_drv.DeviceGetAttribute(&posix, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev) == CUDA_SUCCESS &&
            mc && posix) {
            if (e->listen_sock < 0) {
                snprintf(e->sock_name, sizeof(e->sock_name), "f2v-%d-%p", (int)getpid(), (void*)e);
                int s = socket(AF_UNIX, SOCK_STREAM, 0);
                struct sockaddr_un a;
                socklen_t alen;
                sock_addr(&a, &alen, e->sock_name);
                if (s >= 0) sock_timeouts(s, kHandoffTimeoutS);
                if (s >= 0 && bind(s, (struct sockaddr*)&a, alen) == 0 && listen(s, kMaxWorld) == 0) e->listen_sock = s;
                else if (s >= 0) close(s);
            }
            if (e->listen_sock >= 0) {
                b.mc_ok = 1u | (e->mc_inproc ? 2u : 0u) | (e->want_mc == 3 ? 4u : 0u);
                memcpy(b.sock_name, e->sock_name, std::min(sizeof(b.sock_name) - 1, strlen(e->sock_name)));
            }
        }
    }
    memset(blob, 0, F2V_PEER_BLOB);
    memcpy(blob, &b, sizeof(b));
    return F2V_OK;
}

// Move the tables and the flag page into one VMM allocation bound to a multicast object that all
// ranks share.  Rank 0 creates the object and hands its file descriptor to the others over a
// unix socket; two handshakes make sure every device is added before anyone binds and everything
// is bound before anyone stores.
static int mc_setup(f2v_engine* e, const PeerBlob* blobs, int rank, int world) {
    int r = drv_load();
    if (r) return r;
    CUdevice dev;
    DRV(g_drv.DeviceGet(&dev, e->device));
    const size_t tbl = sizeof(float) * e->rows_alloc * e->dim;
    const size_t flags_off = (2 * tbl + 255) / 256 * 256;
    const size_t need = flags_off + sizeof(uint64_t) * kMaxWorld * kFlagStride;
    CUmulticastObjectProp mp;
    memset(&mp, 0, sizeof(mp));
    mp.numDevices = (unsigned)world;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    mp.size = need;
    size_t gran = 0;
    DRV(g_drv.MulticastGetGranularity(&gran, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED));
    CUmemAllocationProp ap;
    memset(&ap, 0, sizeof(ap));
    ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    ap.location.id = e->device;
    ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t agran = 0;
    DRV(g_drv.MemGetAllocationGranularity(&agran, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    gran = std::max(gran, agran);
    const size_t size = (need + gran - 1) / gran * gran;
    mp.size = size;

    int conns[kMaxWorld];
    for (int k = 0; k < kMaxWorld; k++) conns[k] = -1;
    auto close_all = [&]() { for (int k = 0; k < kMaxWorld; k++) if (conns[k] >= 0) { close(conns[k]); conns[k] = -1; } };
    CUmemGenericAllocationHandle mc = 0;
    if (rank == 0) {
        // if the object cannot be created here (e.g. the fabric does not offer multicast to this
        // set of devices) every rank is told so and the caller falls back to unicast peer stores
        int fd = -1;
        bool have = e->want_mc != 2 &&            // option multicast = 2: exercise the fall-back (tests)
                    g_drv.MulticastCreate(&mc, &mp) == CUDA_SUCCESS;
        if (have && g_drv.MemExportToShareableHandle(&fd, mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) != CUDA_SUCCESS) {
            g_drv.MemRelease(mc);
            have = false;
        }
        for (int k = 1; k < world; k++) {
            int c = accept(e->listen_sock, nullptr, nullptr);
            if (c >= 0) sock_timeouts(c, kHandoffTimeoutS);
            char who = 0;
            if (c < 0 || recv(c, &who, 1, MSG_WAITALL) != 1 || who < 1 || who >= world || conns[(int)who] >= 0) {
                if (c >= 0) close(c);
                close_all();
                return fail(F2V_ERR_STATE, "multicast hand-off: bad connection");
            }
            conns[(int)who] = c;
        }
        for (int k = 1; k < world; k++)
            if (sock_byte(conns[k], true, have ? 'Y' : 'N')) { close_all(); return fail(F2V_ERR_STATE, "multicast hand-off: send failed"); }
        if (!have) { close_all(); return F2V_OK; }      // e->mc_mode stays false
        for (int k = 1; k < world; k++)
            if (sock_send_fd(conns[k], fd) != 0) { close_all(); return fail(F2V_ERR_STATE, "multicast hand-off: cannot send the handle"); }
        close(fd);
    } else {
        int c = socket(AF_UNIX, SOCK_STREAM, 0);
        if (c >= 0) sock_timeouts(c, kHandoffTimeoutS);
        struct sockaddr_un a;
        socklen_t alen;
        sock_addr(&a, &alen, blobs[0].sock_name);
        bool up = false;
        for (int tries = 0; tries < 600 && c >= 0; tries++) {
            if (connect(c, (struct sockaddr*)&a, alen) == 0) { up = true; break; }
            struct timespec ts = {0, 100 * 1000 * 1000};
            nanosleep(&ts, nullptr);
        }
        if (!up) { if (c >= 0) close(c); return fail(F2V_ERR_STATE, "multicast hand-off: rank 0 is not reachable"); }
        conns[0] = c;
        char who = (char)rank;
        if (send(c, &who, 1, 0) != 1) { close_all(); return fail(F2V_ERR_STATE, "multicast hand-off: send failed"); }
        char have = 0;
        if (recv(c, &have, 1, MSG_WAITALL) != 1) { close_all(); return fail(F2V_ERR_STATE, "multicast hand-off: no answer from rank 0"); }
        if (have != 'Y') { close_all(); return F2V_OK; }  // rank 0 could not create the object: unicast
        int fd = sock_recv_fd(c);
        if (fd < 0) { close_all(); return fail(F2V_ERR_STATE, "multicast hand-off: no handle received"); }
        CUresult ir = g_drv.MemImportFromShareableHandle(&mc, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
        close(fd);
        if (ir != CUDA_SUCCESS) { close_all(); return fail(F2V_ERR_CUDA, "cuMemImportFromShareableHandle failed (%d)", (int)ir); }
    }
    // all-rank handshake through rank 0: `tag` from everyone, then `tag+1` back to everyone
    auto handshake = [&](char tag) -> int {
        if (rank == 0) {
            for (int k = 1; k < world; k++) if (sock_byte(conns[k], false, tag)) return -1;
            for (int k = 1; k < world; k++) if (sock_byte(conns[k], true, (char)(tag + 1))) return -1;
            return 0;
        }
        if (sock_byte(conns[0], true, tag)) return -1;
        return sock_byte(conns[0], false, (char)(tag + 1));
    };
    // from here on a local failure must not leave the peers waiting in a handshake or leak the
    // handles: everything acquired is tracked and released on the error path, the sockets are closed
    // (the peers' next handshake byte then fails at once instead of after the time-out)
    CUmemGenericAllocationHandle mem = 0;
    CUdeviceptr uc = 0, mcva = 0;
    bool uc_mapped = false, mc_mapped = false, bound = false;
    int rr = [&]() -> int {
        DRV(g_drv.MulticastAddDevice(mc, dev));
        if (handshake('A')) return fail(F2V_ERR_STATE, "multicast hand-off: add-device handshake failed (a peer gave up?)");
        DRV(g_drv.MemCreate(&mem, size, &ap, 0));
        DRV(g_drv.MulticastBindMem(mc, 0, mem, 0, size, 0));
        bound = true;
        CUmemAccessDesc ad;
        memset(&ad, 0, sizeof(ad));
        ad.location.type = CU_MEM_LOCATION_TYPE_DEVICE; ad.location.id = e->device; ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        DRV(g_drv.MemAddressReserve(&uc, size, gran, 0, 0));
        DRV(g_drv.MemMap(uc, size, 0, mem, 0));
        uc_mapped = true;
        DRV(g_drv.MemSetAccess(uc, size, &ad, 1));
        DRV(g_drv.MemAddressReserve(&mcva, size, gran, 0, 0));
        DRV(g_drv.MemMap(mcva, size, 0, mc, 0));
        mc_mapped = true;
        DRV(g_drv.MemSetAccess(mcva, size, &ad, 1));
        // move the live state over (through the unicast mapping) and drop the cudaMalloc'ed tables
        CU(cudaMemsetAsync((void*)uc, 0, size, e->stream));
        CU(cudaStreamSynchronize(e->stream));
        CU(cudaMemcpyAsync((void*)uc, e->d_Xall, 2 * tbl, cudaMemcpyDeviceToDevice, e->stream));
        CU(cudaDeviceSynchronize());
        return F2V_OK;
    }();
    if (rr != F2V_OK) {
        close_all();
        if (mc_mapped) g_drv.MemUnmap(mcva, size);
        if (mcva) g_drv.MemAddressFree(mcva, size);
        if (bound) g_drv.MulticastUnbind(mc, dev, 0, size);
        if (uc_mapped) g_drv.MemUnmap(uc, size);
        if (uc) g_drv.MemAddressFree(uc, size);
        if (mem) g_drv.MemRelease(mem);
        g_drv.MemRelease(mc);
        return rr;
    }
    CU(cudaFree(e->d_Xall));
    CU(cudaFree(e->d_flags));
    e->d_Xall = (float*)uc;
    e->d_X[0] = e->d_Xall;
    e->d_X[1] = e->d_Xall + e->rows_alloc * e->dim;
    e->d_flags = (uint64_t*)((char*)uc + flags_off);
    e->mc_handle = mc; e->vmm_handle = mem; e->vmm_uc = uc; e->vmm_mc = mcva; e->vmm_size = size;
    e->mc_mode = true;
    if (handshake('C')) { close_all(); return fail(F2V_ERR_STATE, "multicast hand-off: bind handshake failed (a peer gave up?)"); }
    close_all();
    return F2V_OK;
}

// Row-sharded tables.  Every rank creates the physical memory of its shard (both tables) with the
// VMM API, the file descriptors are collected by rank 0 and handed to everybody over the unix
// socket, and every rank maps all shards into one flat virtual range: table t, shard s at
// ((t * world + s) * shard_bytes).  A gather of a remote shard's row is then an ordinary load that
// travels over NVLink; the kernels only permute the row index (shard_row()).
static int shard_setup(f2v_engine* e, const PeerBlob* blobs, int rank, int world) {
    int r = drv_load();
    if (r) return r;
    uint32_t lg = 0;
    while ((1 << lg) < world) lg++;
    if ((1 << lg) != world) return fail(F2V_ERR_ARG, "row-sharded tables need a power-of-two world size (got %d)", world);
    CUmemAllocationProp ap;
    memset(&ap, 0, sizeof(ap));
    ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    ap.location.id = e->device;
    ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t gran = 0;
    DRV(g_drv.MemGetAllocationGranularity(&gran, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    const size_t row_bytes = sizeof(float) * e->dim;
    // rows per shard: ceil(n / world) rounded up until the shard is a whole number of granules
    uint64_t rows = (e->n + (uint64_t)world - 1) / (uint64_t)world;
    size_t g = gran, rb = row_bytes;
    while (rb) { size_t t = g % rb; g = rb; rb = t; }            // g = gcd(gran, row_bytes)
    const uint64_t step = gran / g;                              // rows per smallest granule-aligned block
    rows = (rows + step - 1) / step * step;
    if (2ull * rows * world > 0xffffffffull) return fail(F2V_ERR_ARG, "sharded table too large for 32-bit row ids");
    const size_t shard_bytes = rows * row_bytes;

    // one physical allocation per table (cuMemMap maps whole allocations only)
    CUmemGenericAllocationHandle mine[2] = {0, 0};
    int fds[2][kMaxWorld];
    for (int t = 0; t < 2; t++)
        for (int k = 0; k < kMaxWorld; k++) fds[t][k] = -1;
    for (int t = 0; t < 2; t++) {
        DRV(g_drv.MemCreate(&mine[t], shard_bytes, &ap, 0));
        DRV(g_drv.MemExportToShareableHandle(&fds[t][rank], mine[t], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    }
    int conns[kMaxWorld];
    for (int k = 0; k < kMaxWorld; k++) conns[k] = -1;
    auto close_all = [&]() {
        for (int k = 0; k < kMaxWorld; k++) {
            if (conns[k] >= 0) close(conns[k]);
            conns[k] = -1;
            for (int t = 0; t < 2; t++) { if (fds[t][k] >= 0) close(fds[t][k]); fds[t][k] = -1; }
        }
    };
    if (rank == 0) {
        for (int k = 1; k < world; k++) {
            int c = accept(e->listen_sock, nullptr, nullptr);
            if (c >= 0) sock_timeouts(c, kHandoffTimeoutS);
            char who = 0;
            if (c < 0 || recv(c, &who, 1, MSG_WAITALL) != 1 || who < 1 || who >= world || conns[(int)who] >= 0) {
                if (c >= 0) close(c);
                close_all();
                return fail(F2V_ERR_STATE, "shard hand-off: bad connection");
            }
            conns[(int)who] = c;
            for (int t = 0; t < 2; t++) {
                fds[t][(int)who] = sock_recv_fd(c);
                if (fds[t][(int)who] < 0) { close_all(); return fail(F2V_ERR_STATE, "shard hand-off: no handle from rank %d", (int)who); }
            }
        }
        for (int k = 1; k < world; k++)
            for (int q = 0; q < world; q++)
                for (int t = 0; t < 2; t++)
                    if (q != k && sock_send_fd(conns[k], fds[t][q]) != 0) { close_all(); return fail(F2V_ERR_STATE, "shard hand-off: send failed"); }
    } else {
        int c = socket(AF_UNIX, SOCK_STREAM, 0);
        if (c >= 0) sock_timeouts(c, kHandoffTimeoutS);
        struct sockaddr_un a;
        socklen_t alen;
        sock_addr(&a, &alen, blobs[0].sock_name);
        bool up = false;
        for (int tries = 0; tries < 600 && c >= 0; tries++) {
            if (connect(c, (struct sockaddr*)&a, alen) == 0) { up = true; break; }
            struct timespec ts = {0, 100 * 1000 * 1000};
            nanosleep(&ts, nullptr);
        }
        if (!up) { if (c >= 0) close(c); close_all(); return fail(F2V_ERR_STATE, "shard hand-off: rank 0 is not reachable"); }
        conns[0] = c;
        char who = (char)rank;
        if (send(c, &who, 1, 0) != 1 || sock_send_fd(c, fds[0][rank]) != 0 || sock_send_fd(c, fds[1][rank]) != 0) {
            close_all();
            return fail(F2V_ERR_STATE, "shard hand-off: send failed");
        }
        for (int q = 0; q < world; q++) {
            if (q == rank) continue;
            for (int t = 0; t < 2; t++) {
                fds[t][q] = sock_recv_fd(c);
                if (fds[t][q] < 0) { close_all(); return fail(F2V_ERR_STATE, "shard hand-off: handle of rank %d missing", q); }
            }
        }
    }
    auto handshake = [&](char tag) -> int {
        if (rank == 0) {
            for (int k = 1; k < world; k++) if (sock_byte(conns[k], false, tag)) return -1;
            for (int k = 1; k < world; k++) if (sock_byte(conns[k], true, (char)(tag + 1))) return -1;
            return 0;
        }
        if (sock_byte(conns[0], true, tag)) return -1;
        return sock_byte(conns[0], false, (char)(tag + 1));
    };
    CUdeviceptr va = 0;
    const size_t total = 2 * (size_t)world * shard_bytes;
    DRV(g_drv.MemAddressReserve(&va, total, gran, 0, 0));
    for (int s = 0; s < world; s++) {
        for (int t = 0; t < 2; t++) {
            CUmemGenericAllocationHandle h = mine[t];
            if (s != rank) {
                CUresult ir = g_drv.MemImportFromShareableHandle(&h, (void*)(uintptr_t)fds[t][s], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
                if (ir != CUDA_SUCCESS) { close_all(); return fail(F2V_ERR_CUDA, "cuMemImportFromShareableHandle (shard %d) failed (%d)", s, (int)ir); }
            }
            e->shard_handles[t][s] = h;
            DRV(g_drv.MemMap(va + ((size_t)t * world + s) * shard_bytes, shard_bytes, 0, h, 0));
        }
    }
    CUmemAccessDesc ad;
    memset(&ad, 0, sizeof(ad));
    ad.location.type = CU_MEM_LOCATION_TYPE_DEVICE; ad.location.id = e->device; ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    DRV(g_drv.MemSetAccess(va, total, &ad, 1));
    // zero this rank's shard, then move any live state over (only the rows this rank stores)
    for (int t = 0; t < 2; t++)
        CU(cudaMemsetAsync((void*)(va + ((size_t)t * world + rank) * shard_bytes), 0, shard_bytes, e->stream));
    float* old_all = e->d_Xall;
    float* old_live = e->d_Xall ? e->d_X[e->cur] : nullptr;
    e->shard_va = va; e->shard_bytes = shard_bytes; e->shard_lg = lg; e->shard_rows = (uint32_t)rows;
    e->d_Xall = (float*)va;
    e->rows_alloc = rows * (uint64_t)world;
    e->d_X[0] = e->d_Xall;
    e->d_X[1] = e->d_Xall + e->rows_alloc * e->dim;
    e->rank = rank;
    if (old_live) {
        const uint64_t cells = e->n * e->dim;
        const unsigned grid = (unsigned)std::min<uint64_t>((cells + 255) / 256, (uint64_t)e->sm_count * 16);
        shard_copy_kernel<true><<<grid, 256, 0, e->stream>>>(e->d_X[e->cur], old_live, 0, e->n, e->dim, lg, (uint32_t)rows, (uint32_t)rank);
        CU(cudaGetLastError());
    }
    CU(cudaDeviceSynchronize());
    if (old_all) CU(cudaFree(old_all));
    e->shard_mode = true;
    // nobody reads a shard before its owner has zeroed / filled it
    if (handshake('S')) { close_all(); return fail(F2V_ERR_STATE, "shard hand-off: final handshake failed"); }
    close_all();
    return F2V_OK;
}

int f2v_comm_peer_init(f2v_engine* e, const void* blobs, int rank, int world) {
    if (!e || !blobs) return fail(F2V_ERR_ARG, "null argument");
    if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world)
        return fail(F2V_ERR_ARG, "bad rank/world (peer exchange supports up to %d ranks)", kMaxWorld);
    if (e->comm || e->peer_mode) return fail(F2V_ERR_STATE, "communicator already initialised");
    if (!e->d_flags) return fail(F2V_ERR_STATE, "call f2v_comm_peer_export first");
    CU(cudaSetDevice(e->device));
    const uint64_t me = (uint64_t)getpid();
    // NVLink multicast when every rank can and wants to, and the ranks are separate processes
    // (the hand-off below blocks, so engines of one process driven by one thread cannot use it)
    if (world > 1) {
        PeerBlob all[kMaxWorld];
        bool mc = true;
        int nshard = 0;
        for (int r = 0; r < world; r++) {
            memcpy(&all[r], (const char*)blobs + (size_t)r * F2V_PEER_BLOB, sizeof(PeerBlob));
            if (all[r].magic != kPeerMagic) return fail(F2V_ERR_ARG, "blob %d is not a peer blob", r);
            if (all[r].n != e->n || all[r].dim != e->dim || all[r].rows_alloc != e->rows_alloc)
                return fail(F2V_ERR_ARG, "rank %d holds a different table (n or dim)", r);
            if (all[r].cur != (uint64_t)e->cur) return fail(F2V_ERR_STATE, "rank %d is at a different table parity", r);
            mc = mc && all[r].mc_ok != 0;
            nshard += all[r].shard != 0;
            // engines of ONE process: only if the caller promised a host thread per engine (the hand-off blocks)
            for (int q = 0; q < r; q++) mc = mc && (all[q].pid != all[r].pid || ((all[q].mc_ok & 2u) && (all[r].mc_ok & 2u)));
        }
        if (nshard != 0 && nshard != world) return fail(F2V_ERR_ARG, "the option \"sharded\" must be set on every rank or on none");
        if (nshard) {
            // row-sharded tables: needs the VMM hand-off (separate processes, file-descriptor handles)
            if (!mc) return fail(F2V_ERR_STATE, "row-sharded tables need one process per GPU and VMM file-descriptor handles");
            int rr = shard_setup(e, all, rank, world);
            if (rr) return rr;
            mc = false;                        // rows are stored once, in their shard: nothing to multicast
        }
        // two ranks: one store per peer is the same NVLink egress as a multicast store, and the rank's own copy
        // does not loop through the switch (measured, R-MAT 24 option 5: 22.4 vs 24.5 ms per epoch); from
        // three ranks on multicast wins (egress 1x instead of (N-1)x).  "multicast" = 3 forces it at N=2.
        bool forced = true;                                        // (every rank must take the same decision: it is
        for (int r = 0; r < world; r++) forced = forced && (all[r].mc_ok & 4u);   // derived from the exchanged blobs only)
        if (mc && world == 2 && !forced) mc = false;
        if (mc) {
            int rr = mc_setup(e, all, rank, world);
            if (rr) return rr;
        }
        if (mc && e->mc_mode) {
            e->rank = rank;
            e->world = world;
            e->peer_mode = true;
            e->step_id = 0;
            if (e->listen_sock >= 0) { close(e->listen_sock); e->listen_sock = -1; }
            return F2V_OK;
        }
    }
    for (int r = 0; r < world; r++) {
        PeerBlob b;
        memcpy(&b, (const char*)blobs + (size_t)r * F2V_PEER_BLOB, sizeof(b));
        if (b.magic != kPeerMagic) return fail(F2V_ERR_ARG, "blob %d is not a peer blob", r);
        if (b.n != e->n || b.dim != e->dim || (!e->shard_mode && b.rows_alloc != e->rows_alloc))
            return fail(F2V_ERR_ARG, "rank %d holds a different table (n or dim)", r);
        if (b.cur != (uint64_t)e->cur) return fail(F2V_ERR_STATE, "rank %d is at a different table parity", r);
        if (r == rank) {
            if (!e->shard_mode && b.ptr[0] != (uint64_t)(uintptr_t)e->d_Xall) return fail(F2V_ERR_ARG, "blob %d is not this engine's", r);
            continue;
        }
        if (e->shard_mode) {
            // only the exchange flags are mapped through IPC; the tables are the flat VMM range
            void* m = nullptr;
            CU(cudaIpcOpenMemHandle(&m, b.h[1], cudaIpcMemLazyEnablePeerAccess));
            e->peer_flags[r] = (uint64_t*)m;
            e->peer_ipc[r] = true;
            continue;
        }
        if (b.pid == me) {
            // an engine of this process on another device: plain peer access
            int can = 0;
            CU(cudaDeviceCanAccessPeer(&can, e->device, b.device));
            if (!can) return fail(F2V_ERR_CUDA, "device %d cannot access device %d", e->device, b.device);
            cudaError_t pe = cudaDeviceEnablePeerAccess(b.device, 0);
            if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled)
                return fail(F2V_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(pe));
            cudaGetLastError();
            e->peerX[r][0] = (float*)(uintptr_t)b.ptr[0];
            e->peer_flags[r] = (uint64_t*)(uintptr_t)b.ptr[1];
        } else {
            void* m[2] = {nullptr, nullptr};
            for (int k = 0; k < 2; k++) CU(cudaIpcOpenMemHandle(&m[k], b.h[k], cudaIpcMemLazyEnablePeerAccess));
            e->peerX[r][0] = (float*)m[0];
            e->peer_flags[r] = (uint64_t*)m[1];
            e->peer_ipc[r] = true;
        }
        e->peerX[r][1] = e->peerX[r][0] + e->rows_alloc * e->dim;
    }
    e->rank = rank;
    e->world = world;
    e->peer_mode = world > 1;
    e->step_id = 0;
    return F2V_OK;
}

}  // extern "C"
