// force2vec_b200/csrc/f2v_kernels.cuh -- sm_100a device code of the Force2Vec force step.
//
// One warp owns one work item: a vertex of the minibatch (or, for hub rows, one chunk of a
// vertex's CSR row).  It walks the item's neighbour indices, gathers the neighbour embedding
// rows as coalesced float4 (a d=128 row is one 512-B warp transaction, a d=64 row half a
// warp so two neighbours ride in one instruction), reduces each pair's squared distance /
// dot product with warp shuffles, turns it into the t-distribution or LUT-sigmoid scalar and
// applies the learning-rate-scaled update in registers.  The minibatch's shared negative
// rows (bs=0) are staged once per CTA in shared memory with TMA bulk copies
// (cp.async.bulk + mbarrier -> SASS UBLKCP); per-vertex negatives (bs=1) are gathered like
// neighbours.  No tensor cores: the work is a sparse gather with ~3 flop/byte.
//
// Reference semantics restated here (file:line into /root/reference/sample/algorithms.cpp):
//   t-dist pair     :598-613 / :614-627     sigmoid pair  :854-868 / :898-911
//   walk pair       :1154-1170              fast_SM       :766-770      scale  :6-10
// Jacobi rule (reads see the pre-minibatch table, :588-639): rows < `split` are read from
// Xlo, rows >= split from Xhi, outputs go to `out`; the epoch driver ping-pongs two tables
// (Xlo = out = next table, Xhi = current table, split = first row of the minibatch), the
// single-step driver passes Xlo = Xhi = table and a staging buffer as `out`.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "f2v_plan.hpp"

// How gathered embedding rows (and an item's own row) are loaded: __ldcg (L2 only, default) or
// __ldca (L1 + L2: hub rows that most vertices gather then hit in the SM's own L1 instead of all
// converging on the few L2 slices that hold them).  L1 is legal for them: within an epoch a launch
// reads next-table rows only below its split (final since an earlier minibatch, never cached before
// they were final) and current-table rows (nobody writes them during the epoch), and the epoch's
// first launch is an ordinary launch that starts from an invalidated L1.  It is NOT legal for the
// hub partial sums other CTAs store during the same launch: fold_rows() always loads with __ldcg.
#ifndef F2V_ROW_LOAD
#define F2V_ROW_LOAD __ldcg
#endif

namespace f2v {

constexpr int kTDist = 5, kSigmoid = 6, kWalk = 7;
constexpr int kWalkLen = 5;
constexpr int kLutSize = 2048;
constexpr int kLutAlloc = 2052;          // padded to a multiple of 16 bytes
#ifndef F2V_WARPS
#define F2V_WARPS 8
#endif
constexpr int kWarpsPerCta = F2V_WARPS;   // warps per CTA (tuning: -DF2V_WARPS=9 caps the kernels at 56 registers)
constexpr unsigned kFull = 0xffffffffu;
constexpr int kMaxWorld = 8;             // ranks of one NVSwitch domain
constexpr int kFlagStride = 16;          // u64 per exchange flag: one 128-byte line each

struct BatchParams {
    const Item* items;
    const HubInfo* hub;
    uint32_t n_items;
    uint32_t n_hub;
    uint64_t lo;          // first row of the minibatch (bs=1 window origin)
    // Both embedding tables live in ONE allocation (Xb); vertex j is read from row
    // j + (j < split ? off_lo : off_hi) of it, so a gathered row costs one 32-bit select on the
    // index (done once per index block, lane-parallel) and one IMAD.WIDE instead of a 64-bit
    // compare + pointer select per row.  Epoch: off_lo = next table, off_hi = current table.
    uint32_t split;
    uint32_t off_lo, off_hi;
    const float* Xb;
    // Row-sharded tables (multi-GPU, tables too large for one GPU): the combined table is ONE flat
    // virtual range whose shard s (shard_rows rows of each table) is physical memory of GPU s,
    // mapped on every rank; vertex j lives in shard xorfold(j) & (W-1) at local row j >> shard_lg.
    // shard_lg = 0: not sharded (row = vertex id).
    uint32_t shard_lg;
    uint32_t shard_rows;
    float* out;
    uint64_t out_base;    // out row of vertex v = out + (v - out_base) * dim
    const uint32_t* colids;
    const uint32_t* neg;  // this minibatch's negative indices
    const uint32_t* walks;
    const float* lut;
    float* partials;
    uint32_t* counters;
    uint32_t s;
    uint32_t dim;
    int bs_mode;
    int neg_in_smem;
    int variant;          // layout variant of the d=128 kernels (tuning knob)
    float lr;
    // ---- peer-store exchange (multi-GPU): every finished row is also stored into the other
    // ranks' replicas of `out` over NVLink; one flag per (source rank) and minibatch step.
    uint32_t n_peers;                    // 0 = single GPU / NCCL exchange
    uint32_t rank, world;
    float* peer_out[kMaxWorld - 1];      // the peers' copies of `out` (same row indexing)
    uint64_t* peer_flag[kMaxWorld - 1];  // this rank's flag slot in each peer's flag page
    const uint64_t* flags;               // local flag page: flags[r * kFlagStride] written by rank r
    uint64_t wait_step;                  // before reading the next table: every peer's flag >= wait_step (0 = no wait)
    uint64_t signal_step;                // after the last CTA's stores: peers' flags := signal_step
    uint64_t publish_step;               // != 0: CTA 0 publishes this step once the PREDECESSOR launch is complete
                                         // (stream order / griddepcontrol.wait): the kernel boundary has flushed the
                                         // predecessor's peer stores, so no system-scope fence per CTA and no extra
                                         // launch is needed to hand minibatch b-1's rows to the peers
    uint32_t* done;                      // CTA arrival counter of the launch
    uint32_t n_store;                    // peers that receive unicast row stores (0 with multicast, or in timing probes)
    // NVLink multicast (NVLS): `mc_out` is the multicast mapping of the `out` table -- ONE store
    // lands in every rank's replica (this one included), replicated by the NVSwitch -- and
    // `mc_flag` the multicast mapping of this rank's exchange flag.  nullptr = unicast peer stores.
    float* mc_out;
    uint64_t* mc_flag;
    // A rank that never publishes its step (it failed, or was never launched) must not hang its
    // peers for ever: a wait longer than timeout_ns sets *timed_out and the launch carries on (its
    // results are then meaningless; the engine reports the failure at the next synchronisation).
    uint64_t timeout_ns;
    uint32_t* timed_out;
    // ---- programmatic dependent launch: this launch may start while its predecessor (the previous
    // minibatch) is still draining; everything the predecessor can have written is read only after
    // griddepcontrol.wait.  pdl = 2: the item's own row (from the current table, last written one
    // epoch ago) is also fetched before the wait.
    // ---- dataflow epoch (f2v_set_epoch_mode 2): one launch per epoch, minibatches overlap; a warp
    // waits only until the minibatches that wrote the rows it is about to read are complete.
    // flow_done[b] != 0: every row of minibatch b has been stored; flow_cnt[b]: rows finished so far.
    const uint32_t* flow_done;
    uint32_t* flow_cnt;
    uint32_t flow_batch;                 // rows per minibatch
    uint32_t flow_nb;                    // minibatches per epoch
    uint32_t flow_n;                     // rows of the table
    int pdl;
    int late_wait;        // multi-GPU: 1 = every warp waits for the peers' flags after its dependency wait
                          // (PDL-chained launches); 0 = one CTA-wide wait at kernel entry
};

// What changes from one minibatch to the next (the rest of BatchParams is constant over an epoch).
struct BatchVar {
    const Item* items;
    const HubInfo* hub;
    uint32_t n_items;
    uint32_t split;
    uint32_t lo;
    const uint32_t* neg;
};
__device__ __forceinline__ BatchVar batch_var(const BatchParams& p) {
    return BatchVar{p.items, p.hub, p.n_items, p.split, (uint32_t)p.lo, p.neg};
}

// Vertex id -> row of a table.  Sharded: shard(j) * shard_rows + (j >> lg) with
// shard(j) = (low digit of j) XOR hash(j >> lg): for any fixed high part the low digit maps onto
// the shards one-to-one (so (shard, j >> lg) is a dense unique row), and unlike j mod W it spreads
// R-MAT's hubs -- ids with few one-bits, nearly all congruent 0 mod W -- over the GPUs.
// Branch-free; lg = 0 (not sharded) gives row = j.
__host__ __device__ __forceinline__ uint32_t shard_row(uint32_t j, uint32_t lg, uint32_t shard_rows) {
    const uint32_t h = j >> lg;
#ifdef __CUDA_ARCH__
    const uint32_t mix = __umulhi(h, 0x9E3779B1u);
#else
    const uint32_t mix = (uint32_t)(((uint64_t)h * 0x9E3779B1ull) >> 32);
#endif
    return ((j ^ mix) & ((1u << lg) - 1u)) * shard_rows + h;
}
// Row of the combined (both tables) range that holds vertex j for a launch with the given split.
__device__ __forceinline__ uint32_t table_row(const BatchParams& p, uint32_t j, uint32_t split) {
    return shard_row(j, p.shard_lg, p.shard_rows) + (j < split ? p.off_lo : p.off_hi);
}

// ------------------------------------------------------------------ PTX helpers --------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// 16-byte asynchronous copy global -> shared (SASS: LDGSTS.E.BYPASS.128), bypassing L1; src_bytes = 0
// reads nothing and fills the destination with zeros.  Completion: cp.async.wait_group.
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src_gmem, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src_gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
// Programmatic dependent launch (sm_90+): wait for the prerequisite grid's completion and memory
// flush / allow the dependent grid to be scheduled.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// System-scope flag traffic of the peer-store exchange.
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Multicast stores (the address is a multicast mapping; the switch replicates the write).
__device__ __forceinline__ void mc_st_v4(float* p, float a, float b, float c, float d) {
    asm volatile("multimem.st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void mc_st_f32(float* p, float a) {
    asm volatile("multimem.st.global.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}
__device__ __forceinline__ void mc_st_release_sys(uint64_t* p, uint64_t v) {
    asm volatile("multimem.st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint64_t ld_relaxed_sys(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// Spin until *f >= want, or until the exchange time-out expires (checked every 256 polls).  The polls
// are relaxed loads (an acquire per poll would put a system-scope fence into the spin loop of every
// waiting warp); ONE acquire load after the value has been seen orders the row reads that follow.
__device__ __forceinline__ void wait_flag(const uint64_t* f, uint64_t want, uint64_t timeout_ns, uint32_t* timed_out) {
    uint64_t t0 = 0;
    for (uint32_t spins = 0; ld_relaxed_sys(f) < want; spins++) {
        if ((spins & 15u) == 15u) __nanosleep(32);
        if ((spins & 255u) == 255u && timeout_ns) {
            const uint64_t now = global_timer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > timeout_ns) { if (timed_out) atomicExch(timed_out, 1u); return; }
        }
    }
    (void)ld_acquire_sys(f);
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}

// ------------------------------------------------------------------ row layouts --------
// A warp is cut into G = 32/LPR groups of LPR lanes; every group owns one work item (its own
// vertex) and gathers that item's rows, U at a time.  VecL<D,LPR>: a row is D/4 float4, lane l
// of the group holds float4 l, l+LPR, ... (each load instruction covers LPR*16 contiguous
// bytes of the row: a d=128 row is one 512-B transaction with LPR=32, two 256-B ones with
// LPR=16).  Smaller LPR = more vertices in flight per warp (latency tolerance on short rows)
// and fewer redundant per-pair scalar instructions.
template <int D, int LPR_, int U_, int MINB_ = 1>
struct VecL {
    static constexpr int MINB = MINB_;
    static constexpr int V4 = D / 4;
    static constexpr int LPR = LPR_;
    static constexpr int G = 32 / LPR;
    static constexpr int VPL = V4 / LPR;
    static constexpr int NE = 4 * VPL;
    static constexpr int U = U_;
    static constexpr bool kBulk = true;
    static constexpr int kStages = 0;        // > 0: RingL (asynchronous shared-memory ring, gather_stream)
    static_assert(V4 % LPR == 0 && VPL >= 1 && U <= LPR && LPR % U == 0, "bad layout");
    __device__ static __forceinline__ size_t stride(uint32_t) { return D; }
    __device__ static __forceinline__ void load_g(float (&f)[NE], const float* row, int l, uint32_t) {
#pragma unroll
        for (int k = 0; k < VPL; k++) {
            float4 t = F2V_ROW_LOAD(reinterpret_cast<const float4*>(row) + k * LPR + l);
            f[4 * k + 0] = t.x; f[4 * k + 1] = t.y; f[4 * k + 2] = t.z; f[4 * k + 3] = t.w;
        }
    }
    // L2-only load: data other CTAs of the same launch may have written (hub partial sums)
    __device__ static __forceinline__ void load_cg(float (&f)[NE], const float* row, int l, uint32_t) {
#pragma unroll
        for (int k = 0; k < VPL; k++) {
            float4 t = __ldcg(reinterpret_cast<const float4*>(row) + k * LPR + l);
            f[4 * k + 0] = t.x; f[4 * k + 1] = t.y; f[4 * k + 2] = t.z; f[4 * k + 3] = t.w;
        }
    }
    __device__ static __forceinline__ void load_s(float (&f)[NE], const float* row, int l, uint32_t) {
#pragma unroll
        for (int k = 0; k < VPL; k++) {
            float4 t = *(reinterpret_cast<const float4*>(row) + k * LPR + l);
            f[4 * k + 0] = t.x; f[4 * k + 1] = t.y; f[4 * k + 2] = t.z; f[4 * k + 3] = t.w;
        }
    }
    __device__ static __forceinline__ void store_g(float* row, const float (&f)[NE], int l, uint32_t) {
#pragma unroll
        for (int k = 0; k < VPL; k++)
            __stcg(reinterpret_cast<float4*>(row) + k * LPR + l,
                   make_float4(f[4 * k + 0], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]));
    }
    __device__ static __forceinline__ void store_mc(float* row, const float (&f)[NE], int l, uint32_t) {
#pragma unroll
        for (int k = 0; k < VPL; k++)
            mc_st_v4(row + 4 * (k * LPR + l), f[4 * k + 0], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]);
    }
    // ---- arithmetic on a lane's fragment: packed fp32x2 (FFMA2 / FMUL2 / FADD2 on sm_100a;
    //      every component is rounded exactly like the scalar op)
    __device__ static __forceinline__ float dot(const float (&a)[NE], const float (&b)[NE]) {
        float2 s = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < NE; k += 2)
            s = __ffma2_rn(make_float2(a[k], a[k + 1]), make_float2(b[k], b[k + 1]), s);
        return s.x + s.y;
    }
    // acc += c * x   (single rounding per component)
    __device__ static __forceinline__ void axpy(float (&acc)[NE], float c, const float (&x)[NE]) {
        const float2 c2 = make_float2(c, c);
#pragma unroll
        for (int k = 0; k < NE; k += 2) {
            float2 r = __ffma2_rn(c2, make_float2(x[k], x[k + 1]), make_float2(acc[k], acc[k + 1]));
            acc[k] = r.x; acc[k + 1] = r.y;
        }
    }
    // acc = acc - round(c * x)   (two roundings per component, as the reference's float expression)
    __device__ static __forceinline__ void sub_mul(float (&acc)[NE], float c, const float (&x)[NE]) {
        const float2 cn = make_float2(-c, -c);
#pragma unroll
        for (int k = 0; k < NE; k += 2) {
            float2 r = __fadd2_rn(make_float2(acc[k], acc[k + 1]), __fmul2_rn(cn, make_float2(x[k], x[k + 1])));
            acc[k] = r.x; acc[k + 1] = r.y;
        }
    }
    // d = a - b, returns this lane's sum of d^2
    __device__ static __forceinline__ float diff_ss(float (&d)[NE], const float (&a)[NE], const float (&b)[NE]) {
        const float2 m1 = make_float2(-1.f, -1.f);
        float2 s = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < NE; k += 2) {
            float2 t = __ffma2_rn(make_float2(b[k], b[k + 1]), m1, make_float2(a[k], a[k + 1]));   // a - b, exact product
            d[k] = t.x; d[k + 1] = t.y;
            s = __ffma2_rn(t, t, s);
        }
        return s.x + s.y;
    }
    // acc = acc + round(w * clamp5(round(d * d1)))
    __device__ static __forceinline__ void clamp_acc(float (&acc)[NE], const float (&d)[NE], float d1, float w) {
        const float2 d2 = make_float2(d1, d1), w2 = make_float2(w, w);
#pragma unroll
        for (int k = 0; k < NE; k += 2) {
            float2 t = __fmul2_rn(make_float2(d[k], d[k + 1]), d2);
            t.x = fminf(fmaxf(t.x, -5.0f), 5.0f);
            t.y = fminf(fmaxf(t.y, -5.0f), 5.0f);
            float2 r = __fadd2_rn(make_float2(acc[k], acc[k + 1]), __fmul2_rn(w2, t));
            acc[k] = r.x; acc[k + 1] = r.y;
        }
    }
    // the same with the clamp known to be the identity (|d*d1| <= 5 proven by the caller): no FMNMX
    __device__ static __forceinline__ void scale_acc(float (&acc)[NE], const float (&d)[NE], float d1, float w) {
        const float2 d2 = make_float2(d1, d1), w2 = make_float2(w, w);
#pragma unroll
        for (int k = 0; k < NE; k += 2) {
            const float2 t = __fmul2_rn(make_float2(d[k], d[k + 1]), d2);
            float2 r = __fadd2_rn(make_float2(acc[k], acc[k + 1]), __fmul2_rn(w2, t));
            acc[k] = r.x; acc[k + 1] = r.y;
        }
    }
};

// RingL<D,LPR,S,MINB>: the VecL<D,LPR,2> fragment layout, but gathered rows do not land in registers:
// every lane group owns a ring of S stages x 2 rows in shared memory that it fills with 16-byte
// asynchronous copies (cp.async.cg -> SASS LDGSTS, L2 -> shared memory, no register, no L1
// allocation) S-1 stages ahead of the stage it computes on.  Bytes in flight per SM are then bounded
// by shared memory (up to ~190 KB) instead of by landing registers (~64 KB at full occupancy,
// and none while a warp computes), which is what the latency-bound gather needs.
// A lane reads back exactly the 16-byte pieces it copied itself, so cp.async.wait_group is the
// only synchronisation: no barrier, no cross-lane visibility.
template <int D, int LPR_, int S_, int MINB_>
struct RingL : VecL<D, LPR_, 2, MINB_> {
    static constexpr int kStages = S_;
    static constexpr uint32_t kRowBytes = D * 4;
    static constexpr uint32_t kGroupBytes = S_ * 2 * kRowBytes;                 // ring of one lane group
    static constexpr uint32_t kCtaBytes = kWarpsPerCta * (32 / LPR_) * kGroupBytes;
};

// GenL<NV>: any dim <= 32*NV; one group per warp, lane l holds elements l, l+32, ... (scalar
// loads, still coalesced).
template <int NV>
struct GenL {
    static constexpr int MINB = 1;
    static constexpr int LPR = 32;
    static constexpr int G = 1;
    static constexpr int NE = NV;
    static constexpr int U = NV <= 4 ? 4 : (NV <= 8 ? 2 : 1);
    static constexpr bool kBulk = false;
    static constexpr int kStages = 0;
    __device__ static __forceinline__ size_t stride(uint32_t dim) { return dim; }
    __device__ static __forceinline__ void load_g(float (&f)[NE], const float* row, int l, uint32_t dim) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            uint32_t e = k * 32 + l;
            f[k] = e < dim ? __ldcg(row + e) : 0.f;
        }
    }
    __device__ static __forceinline__ void load_cg(float (&f)[NE], const float* row, int l, uint32_t dim) { load_g(f, row, l, dim); }
    __device__ static __forceinline__ void load_s(float (&f)[NE], const float* row, int l, uint32_t dim) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            uint32_t e = k * 32 + l;
            f[k] = e < dim ? row[e] : 0.f;
        }
    }
    __device__ static __forceinline__ void store_g(float* row, const float (&f)[NE], int l, uint32_t dim) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            uint32_t e = k * 32 + l;
            if (e < dim) __stcg(row + e, f[k]);
        }
    }
    __device__ static __forceinline__ void store_mc(float* row, const float (&f)[NE], int l, uint32_t dim) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            uint32_t e = k * 32 + l;
            if (e < dim) mc_st_f32(row + e, f[k]);
        }
    }
    __device__ static __forceinline__ float dot(const float (&a)[NE], const float (&b)[NE]) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < NE; k++) s = fmaf(a[k], b[k], s);
        return s;
    }
    __device__ static __forceinline__ void axpy(float (&acc)[NE], float c, const float (&x)[NE]) {
#pragma unroll
        for (int k = 0; k < NE; k++) acc[k] = fmaf(c, x[k], acc[k]);
    }
    __device__ static __forceinline__ void sub_mul(float (&acc)[NE], float c, const float (&x)[NE]) {
#pragma unroll
        for (int k = 0; k < NE; k++) acc[k] = __fsub_rn(acc[k], __fmul_rn(c, x[k]));
    }
    __device__ static __forceinline__ float diff_ss(float (&d)[NE], const float (&a)[NE], const float (&b)[NE]) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < NE; k++) {
            d[k] = __fsub_rn(a[k], b[k]);
            s = fmaf(d[k], d[k], s);
        }
        return s;
    }
    __device__ static __forceinline__ void clamp_acc(float (&acc)[NE], const float (&d)[NE], float d1, float w) {
#pragma unroll
        for (int k = 0; k < NE; k++)
            acc[k] = __fadd_rn(acc[k], __fmul_rn(w, fminf(fmaxf(__fmul_rn(d[k], d1), -5.0f), 5.0f)));
    }
    __device__ static __forceinline__ void scale_acc(float (&acc)[NE], const float (&d)[NE], float d1, float w) {
#pragma unroll
        for (int k = 0; k < NE; k++) acc[k] = __fadd_rn(acc[k], __fmul_rn(w, __fmul_rn(d[k], d1)));
    }
};

// ------------------------------------------------------------------ scalar pieces ------
// scale(), algorithms.cpp:6-10 as compiled by the reference (-ffast-math: maxss/minss):
// NaN -> -MAXBOUND.  CUDA fmaxf/fminf return the non-NaN operand, which gives the same.
__device__ __forceinline__ float clamp5(float v) { return fminf(fmaxf(v, -5.0f), 5.0f); }

// fast_SM(), algorithms.cpp:766-770; sum and product in double, truncation, no interpolation.
// Branch-free: the index is taken from v clamped to [-6, 6] (entry 2048 exists and is 1.0f),
// then the reference's two range tests select 1 / 0.
__device__ __forceinline__ float fast_sm(const float* __restrict__ lut, float v) {
    const double res = (double)(float)(kLutSize / 12.0);   // SM_RESOLUTION, algorithms.h:49
    const float vc = fminf(fmaxf(v, -6.0f), 6.0f);
    const int i = (int)(((double)vc + 6.0) * res);
    float sg = __ldg(lut + i);
    sg = v > 6.0f ? 1.0f : sg;
    sg = v < -6.0f ? 0.0f : sg;
    return sg;
}

template <int LPR>
__device__ __forceinline__ float group_sum(float x) {
#pragma unroll
    for (int off = LPR / 2; off >= 1; off >>= 1) x += __shfl_xor_sync(kFull, x, off);
    return x;
}

// Per-pair scalar from the reduced squared distance / dot product `r`.
//   option 5: the factor d1 (algorithms.cpp:608 d1 = -2.0/(1.0+attrc); :622 d1 = 2.0/(repuls*(1.0+repuls)))
//   option 6/7: the coefficient of x_p: attractive STEP*degi*(1.0-sigma) formed in double and
//   rounded once (:866) == fma(-sd, sg, sd); repulsive STEP*sigma in float (:908)
template <int MODEL, bool ATTR>
__device__ __forceinline__ float pair_scalar(float r, float lr, float sd, const float* __restrict__ lut) {
    if (MODEL == kTDist)
        return ATTR ? __fdiv_rn(-2.0f, __fadd_rn(1.0f, r)) : __fdiv_rn(2.0f, __fmul_rn(r, __fadd_rn(1.0f, r)));
    const float sg = fast_sm(lut, r);
    return ATTR ? fmaf(-sd, sg, sd) : __fmul_rn(lr, sg);
}

// Option 5: can scale() (the clamp to +-5, algorithms.cpp:6-10) change any component of this pair?
// |diff_k| <= sqrt(r) with r = sum diff^2, so |diff_k * d1| <= 2 sqrt(r)/(1+r) <= 1 for an attractive
// pair and <= 2/(sqrt(r)(1+r)) <= 3.2 for a repulsive pair with r >= 0.25: the clamp is then the
// identity and its 2*NE min/max instructions (a quarter of the pair's instructions) are skipped.
// Anything else -- a close or identical negative (r < 0.25: the reference's NaN -> -5 quirk lives
// here), a non-finite r -- takes the clamped path.  The decision is taken per warp (one vote).
__device__ __forceinline__ bool may_clamp(float r, bool attr, bool valid) {
    return valid && !(attr ? r < __int_as_float(0x7f800000) : (r >= 0.25f && r < __int_as_float(0x7f800000)));
}

template <class L, int MODEL, bool ATTR>
__device__ __forceinline__ void pair_apply(float (&acc)[L::NE], const float (&xp)[L::NE], const float (&d)[L::NE],
                                           float sc, bool valid, float lr, bool clampless = false) {
    if (MODEL == kTDist) {                                                  // prev += STEP*scale(diff*d1)
        if (clampless) L::scale_acc(acc, d, valid ? sc : 0.f, valid ? lr : 0.f);
        else L::clamp_acc(acc, d, sc, valid ? lr : 0.f);
    }
    else if (ATTR) L::axpy(acc, valid ? sc : 0.f, xp);                     // prev += c*x_j
    else L::sub_mul(acc, valid ? sc : 0.f, xp);                            // prev -= (STEP*d1)*sample
}

// One (i, p) pair per group.  ATTR: attractive (neighbour / walk sample) or repulsive (negative).
// Executed by the whole warp (the reduction shuffles are warp-wide); groups with nothing to do
// pass valid = false and contribute exactly zero.
template <class L, int MODEL, bool ATTR>
__device__ __forceinline__ void pair_update(float (&acc)[L::NE], const float (&xi)[L::NE],
                                            const float (&xp)[L::NE], bool valid, float lr, float sd,
                                            const float* __restrict__ lut) {
    float d[L::NE];
    float r = MODEL == kTDist ? L::diff_ss(d, xi, xp) : L::dot(xi, xp);
    r = group_sum<L::LPR>(r);
    const float sc = pair_scalar<MODEL, ATTR>(r, lr, sd, lut);
    const bool clampless = MODEL == kTDist && !__any_sync(kFull, may_clamp(r, ATTR, valid));
    pair_apply<L, MODEL, ATTR>(acc, xp, d, sc, valid, lr, clampless);
}

// Two pairs per group at once: the two lane-partial sums are reduced with a halving butterfly
// (lower half of the group ends up with pair 0's total, upper half with pair 1's: log2(LPR)
// shuffles for both instead of 2*log2(LPR)), each half evaluates the scalar of ITS pair once,
// and one more shuffle swaps the results.  Updates are applied in pair order (0 then 1).
template <class L, int MODEL, bool ATTR>
__device__ __forceinline__ void pair2_update(float (&acc)[L::NE], const float (&xi)[L::NE],
                                             const float (&x0)[L::NE], const float (&x1)[L::NE],
                                             bool v0, bool v1, float lr, float sd,
                                             const float* __restrict__ lut, int l) {
    constexpr int H = L::LPR / 2;
    float d0[L::NE], d1[L::NE];
    const float p0 = MODEL == kTDist ? L::diff_ss(d0, xi, x0) : L::dot(xi, x0);
    const float p1 = MODEL == kTDist ? L::diff_ss(d1, xi, x1) : L::dot(xi, x1);
    const bool hi = (l & H) != 0;
    float keep = hi ? p1 : p0;
    keep += __shfl_xor_sync(kFull, hi ? p0 : p1, H);
#pragma unroll
    for (int off = H / 2; off >= 1; off >>= 1) keep += __shfl_xor_sync(kFull, keep, off);
    const float mine = pair_scalar<MODEL, ATTR>(keep, lr, sd, lut);
    const float other = __shfl_xor_sync(kFull, mine, H);
    const float s0 = hi ? other : mine, s1 = hi ? mine : other;
    // (each half of the group holds the reduced r of ITS pair: one vote covers both pairs of every group)
    const bool clampless = MODEL == kTDist && !__any_sync(kFull, may_clamp(keep, ATTR, hi ? v1 : v0));
    pair_apply<L, MODEL, ATTR>(acc, x0, d0, s0, v0, lr, clampless);
    pair_apply<L, MODEL, ATTR>(acc, x1, d1, s1, v1, lr, clampless);
}

__device__ __forceinline__ uint32_t warp_max(uint32_t v) { return __reduce_max_sync(kFull, v); }


// ------------------------------------------------------------------ dataflow epoch -----
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Called by the whole warp before it gathers rows of the NEXT table: `need` = 1 + the largest vertex
// id below the split among the ids it is about to read (0 = none); rows [0, rows_ok) are known to be
// complete (a per-warp register, monotone over the epoch).  Minibatches finish roughly in order and
// items are handed out in order, so the wait is "every minibatch up to the one holding row need-1
// is complete": the lanes poll 32 consecutive completion flags per load.  Everything a warp can wait
// for belongs to items handed out before its own, i.e. to warps that are running: no deadlock; the
// time-out only turns a programming error into a reported failure instead of a hung GPU.
__device__ __forceinline__ void flow_wait(const BatchParams& p, uint32_t need, uint32_t& rows_ok, int lane) {
    if (need <= rows_ok) return;
    uint32_t F = rows_ok / p.flow_batch;                 // minibatches known complete
    const uint32_t want = (need - 1) / p.flow_batch;     // must become < F
    uint64_t t0 = 0;
    for (uint32_t spins = 0; F <= want; spins++) {
        const uint32_t b = F + (uint32_t)lane;
        const uint32_t ok = b < p.flow_nb ? ld_acquire_gpu(p.flow_done + b) : 0u;
        const uint32_t mask = __ballot_sync(kFull, ok != 0u);
        const uint32_t lead = mask == kFull ? 32u : (uint32_t)__ffs((int)~mask) - 1u;
        F += lead;
        if (lead == 0u) {
            __nanosleep(64);
            if ((spins & 1023u) == 1023u && p.timeout_ns) {
                const uint64_t now = global_timer_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > p.timeout_ns) { if (p.timed_out && lane == 0) atomicExch(p.timed_out, 2u); break; }
            }
        }
    }
    rows_ok = max(rows_ok, min(F * p.flow_batch, p.flow_n));
}

// Multi-GPU, PDL-chained launches: the LAZY exchange wait.  Minibatch b's launch may only read a row
// of the next table written by a PEER once that peer has published the step of the row's minibatch --
// but most rows a warp gathers were written many minibatches ago, and only a fraction of the items
// touch the previous minibatch at all.  So instead of every warp waiting for every peer's previous
// step before it reads anything, a warp looks at the ids it is about to gather: `need` = 1 + the
// largest vertex id below the split (0 = none).  Rows [0, rows_ok) are known to be complete in every
// replica (per-warp register); if need exceeds it the lanes read the peers' step counters (relaxed
// loads, local memory) and derive how far back the slowest peer is: step wait_step covers rows below
// the split, each step before it one minibatch less.  Only a warp whose rows are really not there yet
// spins.  The exchange latency and the ranks' imbalance then hide behind the items that do not depend
// on the previous minibatch, instead of idling the whole GPU once per minibatch.
__device__ __forceinline__ void lazy_peer_wait(const BatchParams& p, uint32_t need, uint32_t split, uint32_t& rows_ok, int lane) {
    if (need <= rows_ok) return;
    const bool poll = (uint32_t)lane < p.world && ((uint32_t)lane != p.rank || p.mc_flag != nullptr);
    const uint64_t* f = p.flags + (size_t)lane * kFlagStride;
    uint64_t t0 = 0;
    uint32_t ok = rows_ok;
    for (uint32_t spins = 0;; spins++) {
        uint32_t lag = 0;                                  // minibatches this peer is behind wait_step
        if (poll) {
            const uint64_t v = ld_relaxed_sys(f);
            lag = v >= p.wait_step ? 0u : (uint32_t)min(p.wait_step - v, (uint64_t)0xffffffu);
        }
        lag = __reduce_max_sync(kFull, lag);
        const uint64_t behind = (uint64_t)lag * p.flow_batch;
        ok = behind >= split ? 0u : split - (uint32_t)behind;
        if (need <= ok) break;
        if ((spins & 7u) == 7u) __nanosleep(64);
        if ((spins & 255u) == 255u && p.timeout_ns) {
            const uint64_t now = global_timer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > p.timeout_ns) { if (p.timed_out && lane == 0) atomicExch(p.timed_out, 1u); break; }
        }
    }
    if (poll) (void)ld_acquire_sys(f);                     // order the row reads that follow
    __syncwarp();
    rows_ok = max(rows_ok, ok);
}

// Every group gathers its own `cnt` rows named by idx[0..cnt) (in order) and folds them into
// its acc.  Indices are fetched LPR at a time per group (coalesced) and broadcast inside the
// group by shuffle; U row loads per group (G*U per warp) are in flight.
template <class L, int MODEL, bool ATTR>
__device__ __forceinline__ void gather_pairs(float (&acc)[L::NE], const float (&xi)[L::NE],
                                             const uint32_t* __restrict__ idx, uint32_t cnt,
                                             uint32_t self, const BatchParams& p, uint32_t split, float sd, int l,
                                             const float* __restrict__ lut, uint32_t& rows_ok, bool have_first = false,
                                             uint32_t first = 0) {
    constexpr int LPR = L::LPR, U = L::U;
    const size_t rs = L::stride(p.dim);
    const float* const Xb = p.Xb;
    const uint32_t cnt_max = L::G > 1 ? warp_max(cnt) : cnt;
    for (uint32_t base = 0; base < cnt_max; base += LPR) {
        const uint32_t nb = cnt > base ? min((uint32_t)LPR, cnt - base) : 0u;
        const uint32_t nb_max = min((uint32_t)LPR, cnt_max - base);
        // (ld.global.cg, not the non-coherent path: negative and walk indices are rewritten between
        // epochs, and with dependent launches an SM's L1 can outlive a launch boundary)
        uint32_t mine = (have_first && base == 0) ? first : ((uint32_t)l < nb ? __ldcg(idx + base + l) : self);
        if (p.flow_done != nullptr)                      // dataflow epoch: the writers of the rows below the split must be done
            flow_wait(p, __reduce_max_sync(kFull, mine < split ? mine + 1u : 0u), rows_ok, (int)(threadIdx.x & 31));
        else if (p.late_wait)                            // multi-GPU: the peers that wrote them must have published their step
            lazy_peer_wait(p, __reduce_max_sync(kFull, mine < split ? mine + 1u : 0u), split, rows_ok, (int)(threadIdx.x & 31));
        mine = table_row(p, mine, split);                // row of the combined table
        for (uint32_t t0 = 0; t0 < nb_max; t0 += U) {
            float rows[U][L::NE];
            bool valid[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                // lanes past nb hold `self`, so an out-of-range slot re-reads the item's own row
                // (an L1/L2 hit) instead of branching; its pair is zeroed through `valid`
                const uint32_t slot = t0 + u;
                const uint32_t j = __shfl_sync(kFull, mine, slot, LPR);
                valid[u] = slot < nb;
                L::load_g(rows[u], Xb + (size_t)j * rs, l, p.dim);
            }
            if (U % 2 == 0) {
#pragma unroll
                for (int u = 0; u < U; u += 2)
                    pair2_update<L, MODEL, ATTR>(acc, xi, rows[u], rows[u + 1 < U ? u + 1 : u], valid[u],
                                                     valid[u + 1 < U ? u + 1 : u], p.lr, sd, lut, l);
            } else {
#pragma unroll
                for (int u = 0; u < U; u++)
                    pair_update<L, MODEL, ATTR>(acc, xi, rows[u], valid[u], p.lr, sd, lut);
            }
        }
    }
}


// Per-pair scalar / update with the pair's class (attractive or repulsive) known only at run time:
// the unified row stream of gather_stream() mixes an item's neighbours and its per-vertex negatives,
// and the lane groups of a warp sit at different positions of their streams.  Arithmetic identical
// to the templated versions above, operation for operation.
template <int MODEL>
__device__ __forceinline__ float pair_scalar_rt(float r, bool attr, float lr, float sd, const float* __restrict__ lut) {
    if (MODEL == kTDist) {
        const float one_r = __fadd_rn(1.0f, r);
        return __fdiv_rn(attr ? -2.0f : 2.0f, attr ? one_r : __fmul_rn(r, one_r));
    }
    const float sg = fast_sm(lut, r);
    return attr ? fmaf(-sd, sg, sd) : __fmul_rn(lr, sg);
}
template <class L, int MODEL>
__device__ __forceinline__ void pair_apply_rt(float (&acc)[L::NE], const float (&xp)[L::NE], const float (&d)[L::NE],
                                              float sc, bool valid, bool attr, float lr, bool clampless) {
    if (MODEL == kTDist) {
        if (clampless) L::scale_acc(acc, d, valid ? sc : 0.f, valid ? lr : 0.f);
        else L::clamp_acc(acc, d, sc, valid ? lr : 0.f);
    }
    else if (attr) L::axpy(acc, valid ? sc : 0.f, xp);
    else L::sub_mul(acc, valid ? sc : 0.f, xp);
}
template <class L, int MODEL>
__device__ __forceinline__ void pair2_update_rt(float (&acc)[L::NE], const float (&xi)[L::NE],
                                                const float (&x0)[L::NE], const float (&x1)[L::NE],
                                                bool v0, bool v1, bool a0, bool a1, float lr, float sd,
                                                const float* __restrict__ lut, int l) {
    constexpr int H = L::LPR / 2;
    float d0[L::NE], d1[L::NE];
    const float p0 = MODEL == kTDist ? L::diff_ss(d0, xi, x0) : L::dot(xi, x0);
    const float p1 = MODEL == kTDist ? L::diff_ss(d1, xi, x1) : L::dot(xi, x1);
    const bool hi = (l & H) != 0;
    float keep = hi ? p1 : p0;
    keep += __shfl_xor_sync(kFull, hi ? p0 : p1, H);
#pragma unroll
    for (int off = H / 2; off >= 1; off >>= 1) keep += __shfl_xor_sync(kFull, keep, off);
    const float mine = pair_scalar_rt<MODEL>(keep, hi ? a1 : a0, lr, sd, lut);
    const float other = __shfl_xor_sync(kFull, mine, H);
    const float s0 = hi ? other : mine, s1 = hi ? mine : other;
    const bool clampless = MODEL == kTDist && !__any_sync(kFull, may_clamp(keep, hi ? a1 : a0, hi ? v1 : v0));
    pair_apply_rt<L, MODEL>(acc, x0, d0, s0, v0, a0, lr, clampless);
    pair_apply_rt<L, MODEL>(acc, x1, d1, s1, v1, a1, lr, clampless);
}

// The asynchronous gather (RingL layouts).  An item's rows form ONE stream: its cntA attractive rows
// (CSR neighbours / walk samples, ids at idxA) followed by its cntB repulsive rows (per-vertex
// negatives, ids at idxB) -- the reference's order (algorithms.cpp:598-627).  The group copies the
// stream two rows (one stage) at a time into its shared-memory ring, S-1 stages ahead of the stage it
// computes on, so the copies of the next rows are in flight WHILE the current pair is reduced and
// applied, and a short row's negatives are already on their way while its neighbours are processed.
// Ids are fetched LPR at a time (one coalesced load per group), one block ahead of the copy pointer.
// All groups of the warp run the same number of stages (the reductions are warp-wide); slots past a
// group's own stream are zero-filled without touching memory and contribute exactly zero.
template <class L, int MODEL>
__device__ __forceinline__ void gather_stream(float (&acc)[L::NE], const float (&xi)[L::NE],
                                              const uint32_t* __restrict__ idxA, uint32_t cntA,
                                              const uint32_t* __restrict__ idxB, uint32_t cntB,
                                              uint32_t self, const BatchParams& p, uint32_t split, float sd, int l,
                                              const float* __restrict__ lut, uint32_t ring, uint32_t& rows_ok, bool have_first,
                                              uint32_t first) {
    constexpr int LPR = L::LPR, S = L::kStages, VPL = L::VPL;
    constexpr uint32_t RB = L::kRowBytes;
    static_assert(S >= 2 && LPR % 2 == 0, "ring layout");
    const uint32_t cnt = cntA + cntB;
    const uint32_t cnt_max = L::G > 1 ? warp_max(cnt) : cnt;
    if (cnt_max == 0) return;
    const uint32_t nst = (cnt_max + 1) >> 1;
    const float* const Xb = p.Xb;
    // this lane's id (as a row of the combined table) at stream position base + l
    auto load_ids = [&](uint32_t base) -> uint32_t {
        const uint32_t P = base + (uint32_t)l;
        uint32_t j = self;
        if (P < cntA) j = __ldcg(idxA + P);
        else if (P < cnt) j = __ldcg(idxB + (P - cntA));
        if (p.flow_done != nullptr)
            flow_wait(p, __reduce_max_sync(kFull, j < split ? j + 1u : 0u), rows_ok, (int)(threadIdx.x & 31));
        else if (p.late_wait)
            lazy_peer_wait(p, __reduce_max_sync(kFull, j < split ? j + 1u : 0u), split, rows_ok, (int)(threadIdx.x & 31));
        return table_row(p, j, split);
    };
    if (have_first && p.flow_done != nullptr)
        flow_wait(p, __reduce_max_sync(kFull, first < split ? first + 1u : 0u), rows_ok, (int)(threadIdx.x & 31));
    else if (have_first && p.late_wait)
        lazy_peer_wait(p, __reduce_max_sync(kFull, first < split ? first + 1u : 0u), split, rows_ok, (int)(threadIdx.x & 31));
    uint32_t ids_cur = have_first ? table_row(p, first, split) : load_ids(0);   // block of the copy pointer
    uint32_t ids_nxt = cnt_max > (uint32_t)LPR ? load_ids(LPR) : 0u;             // the block after it
    uint32_t blk = 0;                                                            // block index of ids_cur
    const uint32_t lane_off = (uint32_t)l * 16u;
    auto issue = [&](uint32_t st) {
        const uint32_t q0 = 2u * st;
        if (q0 / LPR != blk) {                       // the copy pointer enters the next id block (warp-uniform)
            blk++;
            ids_cur = ids_nxt;
            ids_nxt = cnt_max > (blk + 1) * LPR ? load_ids((blk + 1) * LPR) : 0u;
        }
        const uint32_t slot = ring + (st % S) * 2u * RB + lane_off;
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const uint32_t q = q0 + u;
            const uint32_t j = __shfl_sync(kFull, ids_cur, q % LPR, LPR);
            const uint32_t nbytes = q < cnt ? 16u : 0u;                        // past the stream: zero fill, no read
            const float* src = Xb + (size_t)j * (RB / 4) + (uint32_t)l * 4u;
#pragma unroll
            for (int k = 0; k < VPL; k++)
                cp_async16(slot + u * RB + k * LPR * 16u, src + k * LPR * 4, nbytes);
        }
    };
#pragma unroll
    for (int st = 0; st < S - 1; st++) {
        if ((uint32_t)st < nst) issue(st);
        cp_async_commit();
    }
    for (uint32_t k = 0; k < nst; k++) {
        if (k + S - 1 < nst) issue(k + S - 1);
        cp_async_commit();                           // one group per iteration (possibly empty): uniform accounting
        cp_async_wait<S - 1>();                      // stage k has landed (this lane's own pieces)
        const uint32_t slot = ring + (k % S) * 2u * RB + lane_off;
        float x0[L::NE], x1[L::NE];
#pragma unroll
        for (int c = 0; c < VPL; c++) {
            const float4 a = lds128(slot + c * LPR * 16u);
            const float4 b = lds128(slot + RB + c * LPR * 16u);
            x0[4 * c + 0] = a.x; x0[4 * c + 1] = a.y; x0[4 * c + 2] = a.z; x0[4 * c + 3] = a.w;
            x1[4 * c + 0] = b.x; x1[4 * c + 1] = b.y; x1[4 * c + 2] = b.z; x1[4 * c + 3] = b.w;
        }
        const uint32_t q0 = 2u * k;
        pair2_update_rt<L, MODEL>(acc, xi, x0, x1, q0 < cnt, q0 + 1 < cnt, q0 < cntA, q0 + 1 < cntA, p.lr, sd, lut, l);
    }
    cp_async_wait<0>();
}

// acc = rows[0] + rows[1] + ... + rows[cnt-1] (in that order), CU partial rows in flight.
template <class L>
__device__ __forceinline__ void fold_rows(float (&acc)[L::NE], const float* rows, uint32_t cnt, size_t rs,
                                          int l, uint32_t dim) {
    constexpr int NE = L::NE, CU = 4;
#pragma unroll
    for (int k = 0; k < NE; k++) acc[k] = 0.f;
    uint32_t c = 0;
    for (; c + CU <= cnt; c += CU) {
        float part[CU][NE];
#pragma unroll
        for (int u = 0; u < CU; u++) L::load_cg(part[u], rows + (size_t)(c + u) * rs, l, dim);
#pragma unroll
        for (int u = 0; u < CU; u++)
#pragma unroll
            for (int k = 0; k < NE; k++) acc[k] += part[u][k];
    }
    for (; c < cnt; c++) {
        float part[NE];
        L::load_cg(part, rows + (size_t)c * rs, l, dim);
#pragma unroll
        for (int k = 0; k < NE; k++) acc[k] += part[k];
    }
}

__device__ __forceinline__ void publish_step(const BatchParams& p, uint64_t step) {
    if (p.mc_flag != nullptr) { mc_st_release_sys(p.mc_flag, step); return; }
    for (uint32_t r = 0; r < p.n_peers; r++) st_release_sys(p.peer_flag[r], step);
}

// Predecessor complete (the caller has passed griddepcontrol.wait, or the launch is an ordinary
// stream-ordered one): its rows are performed in every replica, so its step can be published.
__device__ __forceinline__ void publish_predecessor(const BatchParams& p) {
    if (p.publish_step != 0 && p.n_peers != 0 && blockIdx.x == 0 && threadIdx.x == 0) publish_step(p, p.publish_step);
}

// G work items, one per group: items t_base .. t_base+G-1 (groups past n_items idle).
// s_neg: the staged negative rows (bs=0) or nullptr; neg_bar: their mbarrier.
template <class L, int MODEL>
__device__ __forceinline__ void process_items(const BatchParams& p, const BatchVar& bv, uint32_t t_base,
                                              const float* s_neg, uint64_t* neg_bar, int lane,
                                              const float* __restrict__ lut, uint32_t ring_base, uint32_t& rows_ok) {
    constexpr int NE = L::NE, LPR = L::LPR;
    constexpr bool kRing = L::kStages > 0;
    const int g = lane / LPR, l = lane % LPR;
    const size_t rs = L::stride(p.dim);
    const uint32_t t = t_base + g;
    bool active = t < bv.n_items;
    Item it{0, 0, 0};
    if (active) it = bv.items[t];
    if (it.v == kNoVertex) { active = false; it.v = (uint32_t)bv.lo; }   // padding item of a dataflow plan
    const bool is_chunk = (it.len & kChunkFlag) != 0;
    const uint32_t len = it.len & ~kChunkFlag;
    const uint32_t v = it.v;
    uint32_t deg = len;
    HubInfo h{0, 1, 0, 0};
    if (is_chunk) { h = bv.hub[t]; deg = h.deg; }
    // the item's first block of neighbour ids is static data: fetch it before the dependency wait
    const uint32_t* const nbr = MODEL == kWalk ? p.walks + (size_t)v * kWalkLen : p.colids + it.e0;
    const uint32_t nbr_cnt = MODEL == kWalk ? (active ? (uint32_t)kWalkLen : 0u) : len;
    const bool early_idx = p.pdl != 0 && MODEL != kWalk;     // walks may have been sampled by the predecessor kernel
    uint32_t first = v;
    if (early_idx && (uint32_t)l < min((uint32_t)LPR, nbr_cnt)) first = __ldg(nbr + l);
    // asynchronous-ring layouts: a row's per-vertex negatives (bs=1, or shared negatives that are not
    // staged in shared memory) ride in the same stream as its neighbours; hub chunks take theirs after
    // the fold.  The negative stream is uploaded before the epoch's first (ordinary) launch: static here.
    const uint32_t* const nidx = bv.neg + ((p.bs_mode && active) ? (size_t)(v - bv.lo) : 0);
    const uint32_t negB = (kRing && s_neg == nullptr && active && !is_chunk) ? p.s : 0u;
    if (kRing && early_idx && (uint32_t)l >= nbr_cnt && (uint32_t)l < nbr_cnt + negB) first = __ldcg(nidx + ((uint32_t)l - nbr_cnt));
    float xi[NE];
    if (p.pdl == 1) { pdl_wait(); pdl_launch_dependents(); }
    if (active) L::load_g(xi, p.Xb + (size_t)table_row(p, v, bv.split) * rs, l, p.dim);
    else {
#pragma unroll
        for (int k = 0; k < NE; k++) xi[k] = 0.f;
    }
    // the dependent launch is released only after this one's own wait: when minibatch b+1 starts,
    // minibatch b-1 is therefore complete
    if (p.pdl == 2) { pdl_wait(); pdl_launch_dependents(); }
    // multi-GPU, PDL-chained launch: rows of the next table written by the peers' previous minibatch.
    // (The item's own row comes from the current table, which nobody writes between the epoch's first
    // launch -- an ordinary launch that waits at kernel entry, because an upload / broadcast may just
    // have rewritten the current table -- and the epoch's end.)
    // -> the wait itself is lazy: lazy_peer_wait() in the gather, on the ids about to be read
    float sd = 0.f;
    if (MODEL != kTDist) {
        // degi = 1.0/(deg+1) stored to float (algorithms.cpp:852,1159); STEP*degi in float
        float degi = (float)(1.0 / (double)(deg + 1u));
        sd = __fmul_rn(p.lr, degi);
    }
    // opt 6/7 accumulate onto y = x_i (algorithms.cpp:824-831); opt 5 onto 0 (:559-567)
    const bool start_at_xi = (MODEL != kTDist) && !is_chunk;
    float acc[NE];
#pragma unroll
    for (int k = 0; k < NE; k++) acc[k] = start_at_xi ? xi[k] : 0.f;

    uint32_t ring = 0;
    if constexpr (kRing) {
        ring = ring_base + (uint32_t)(((threadIdx.x >> 5) * L::G + g) * L::kGroupBytes);
        gather_stream<L, MODEL>(acc, xi, nbr, nbr_cnt, nidx, negB, v, p, bv.split, sd, l, lut, ring, rows_ok, early_idx, first);
    } else {
        gather_pairs<L, MODEL, true>(acc, xi, nbr, nbr_cnt, v, p, bv.split, sd, l, lut, rows_ok, early_idx, first);
    }

    // split rows: publish this chunk's partial sum; the last chunk of a fold block (kFoldBlock
    // consecutive chunks) to arrive folds the block in chunk order and -- rows with more than one
    // block -- publishes the block sum; the last block to arrive folds the block sums in block
    // order and finishes the row.  Deterministic (no float atomics), and the fold's critical path
    // is ~kFoldBlock + nchunks/kFoldBlock partial rows instead of nchunks.
    bool finish = active;
    const bool any_chunk = __any_sync(kFull, is_chunk);
    if (any_chunk) {
        const uint32_t slot0 = h.slot - h.chunk;
        const uint32_t nblk = (h.nchunks + kFoldBlock - 1) / kFoldBlock;
        const uint32_t blk = h.chunk / kFoldBlock;
        const uint32_t blk_n = min((uint32_t)kFoldBlock, h.nchunks - blk * kFoldBlock);
        if (is_chunk) L::store_g(p.partials + (size_t)h.slot * rs, acc, l, p.dim);
        __threadfence();
        __syncwarp();
        uint32_t old = 0;
        if (is_chunk && l == 0) old = atomicAdd(p.counters + slot0 + blk, 1u);
        old = __shfl_sync(kFull, old, 0, LPR);
        bool last = is_chunk && old == blk_n - 1;
        if (last) {
            __threadfence();
            if (l == 0) p.counters[slot0 + blk] = 0;     // every chunk of the block has arrived: re-arm
            fold_rows<L>(acc, p.partials + (size_t)(slot0 + blk * kFoldBlock) * rs, blk_n, rs, l, p.dim);
        }
        const bool second = last && nblk > 1;
        if (__any_sync(kFull, second)) {
            if (second) L::store_g(p.partials + (size_t)(slot0 + h.nchunks + blk) * rs, acc, l, p.dim);
            __threadfence();
            __syncwarp();
            old = 0;
            if (second && l == 0) old = atomicAdd(p.counters + slot0 + nblk, 1u);
            old = __shfl_sync(kFull, old, 0, LPR);
            const bool last2 = second && old == nblk - 1;
            if (last2) {
                __threadfence();
                if (l == 0) p.counters[slot0 + nblk] = 0;
                fold_rows<L>(acc, p.partials + (size_t)(slot0 + h.nchunks) * rs, nblk, rs, l, p.dim);
            }
            if (second) last = last2;
        }
        if (is_chunk) finish = last;
    }

    // repulsive part: s negatives (algorithms.cpp:614-627, 898-911, 1172-1183)
    if (s_neg != nullptr) {
        mbar_wait(neg_bar, 0);
        uint32_t q = 0;
        for (; q + 2 <= p.s; q += 2) {
            float r0[NE], r1[NE];
            L::load_s(r0, s_neg + (size_t)q * rs, l, p.dim);
            L::load_s(r1, s_neg + (size_t)(q + 1) * rs, l, p.dim);
            pair2_update<L, MODEL, false>(acc, xi, r0, r1, finish, finish, p.lr, sd, lut, l);
        }
        if (q < p.s) {
            float row[NE];
            L::load_s(row, s_neg + (size_t)q * rs, l, p.dim);
            pair_update<L, MODEL, false>(acc, xi, row, finish, p.lr, sd, lut);
        }
    } else if constexpr (kRing) {
        // rows that were not split took their negatives in the stream above; a split row's negatives
        // follow the fold (the chunk that finished the row)
        if (any_chunk)
            gather_stream<L, MODEL>(acc, xi, nbr, 0u, nidx, (is_chunk && finish) ? p.s : 0u, v, p, bv.split, sd, l, lut,
                                        ring, rows_ok, false, 0u);
    } else {
        gather_pairs<L, MODEL, false>(acc, xi, nidx, finish ? p.s : 0u, v, p, bv.split, sd, l, lut, rows_ok);
    }
    if (finish) {
        if (MODEL == kTDist || is_chunk) {
#pragma unroll
            for (int k = 0; k < NE; k++) acc[k] = __fadd_rn(xi[k], acc[k]);   // X[i] += delta (:629-639)
        }
        // (out_base > 0 only for the teacher-forced single step, which is never sharded)
        const size_t off = (size_t)((uint64_t)shard_row(v, p.shard_lg, p.shard_rows) - p.out_base) * rs;
        // multi-GPU: the exchange is fused here -- the row goes straight into every replica, with one
        // multicast store (NVLS) or one store per peer
        if (p.mc_out != nullptr) {
            L::store_mc(p.mc_out + off, acc, l, p.dim);
        } else {
            L::store_g(p.out + off, acc, l, p.dim);
            for (uint32_t r = 0; r < p.n_store; r++) L::store_g(p.peer_out[r] + off, acc, l, p.dim);
        }
    }
    if (p.flow_done != nullptr) {
        // dataflow epoch: count this warp's finished rows into their minibatch (all lane groups of a warp
        // work on the same one); whoever completes it publishes the flag the readers of its rows poll
        __threadfence();
        const uint32_t k = (uint32_t)__popc(__ballot_sync(kFull, finish && l == 0));
        if (lane == 0 && k != 0u) {
            const uint32_t b = v / p.flow_batch;
            const uint32_t rows_b = min(p.flow_batch, p.flow_n - b * p.flow_batch);
            const uint32_t old = atomicAdd(p.flow_cnt + b, k);
            if (old + k == rows_b) { __threadfence(); st_release_gpu(const_cast<uint32_t*>(p.flow_done) + b, 1u); }
        }
    }
}

// Exchange barrier, entry side: before any row of the next table is read, every peer must have
// published the minibatch step that wrote it.  Threads 0..world-1 poll one flag each.
__device__ __forceinline__ void peer_wait(const BatchParams& p) {
    if (p.wait_step == 0) return;
    // with multicast this rank's own rows also come back through the switch: wait for its flag too
    if (threadIdx.x < p.world && (threadIdx.x != p.rank || p.mc_flag != nullptr)) {
        const uint64_t* f = p.flags + (size_t)threadIdx.x * kFlagStride;
        wait_flag(f, p.wait_step, p.timeout_ns, p.timed_out);
    }
    __syncthreads();
}

// Exchange barrier, exit side: the last CTA of the launch to finish its (local and peer) stores
// publishes signal_step in every peer's flag page.  Called by all threads of the CTA.
__device__ __forceinline__ void peer_signal(const BatchParams& p) {
    if (p.n_peers == 0 || p.signal_step == 0) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const uint32_t old = atomicAdd(p.done, 1u);
        if (old == gridDim.x - 1) {
            *p.done = 0;                         // re-arm for the next launch (stream-ordered)
            __threadfence_system();
            publish_step(p, p.signal_step);
        }
    }
}

// Stage the minibatch's s shared negative rows in shared memory (TMA bulk copies issued by
// the lanes of warp 0, completion on the CTA's mbarrier whose expected byte count the caller set).
template <class L>
__device__ __forceinline__ void stage_negatives(const BatchParams& p, const BatchVar& bv, float* s_neg, uint64_t* bar) {
    const size_t rs = L::stride(p.dim);
    const uint32_t row_bytes = (uint32_t)(rs * sizeof(float));
    for (uint32_t q = threadIdx.x; q < p.s; q += 32) {      // called by warp 0 after expect_tx
        const uint32_t j = __ldcg(bv.neg + q);
        const float* src = p.Xb + (size_t)table_row(p, j, bv.split) * rs;
        bulk_g2s(s_neg + (size_t)q * rs, src, row_bytes, bar);
    }
}

// ------------------------------------------------------------------ kernels ------------
// One launch per minibatch: one group of lanes per item and one pass per CTA; the hardware CTA scheduler
// balances the load (items are ordered hub chunks first, then rows by descending degree class, so it is
// longest-first).  The minibatch's shared negative rows (bs=0) are staged once per CTA by TMA bulk copies.
template <class L, int MODEL>
__global__ void __launch_bounds__(kWarpsPerCta * 32, L::MINB)
force_batch_kernel(const BatchParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    const bool negs = L::kBulk && p.neg_in_smem;
    float* s_neg = negs ? reinterpret_cast<float*>(smem_raw + 128) : nullptr;
    const uint32_t neg_bytes = negs ? (uint32_t)(p.s * L::stride(p.dim) * sizeof(float)) : 0u;
    const BatchVar bv = batch_var(p);
    if (!p.late_wait) { publish_predecessor(p); peer_wait(p); }
    if (negs) {
        if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
        __syncthreads();
        if (threadIdx.x < 32) {
            if (threadIdx.x == 0) mbar_expect_tx(bar, neg_bytes);
            if (p.pdl) pdl_wait();                  // negative rows may have been written by the previous minibatch
            if (p.late_wait) {
                publish_predecessor(p);
                uint32_t need = 0, ro = 0;                 // the staged negative rows that lie below the split
                for (uint32_t q = threadIdx.x; q < p.s; q += 32) {
                    const uint32_t jn = __ldcg(bv.neg + q);
                    if (jn < bv.split) need = max(need, jn + 1u);
                }
                lazy_peer_wait(p, __reduce_max_sync(kFull, need), bv.split, ro, (int)threadIdx.x);
            }
            if (p.wait_step || p.pdl) fence_proxy_async();   // rows written with generic stores (here or by peers) are read by TMA next
            __syncwarp();
            stage_negatives<L>(p, bv, s_neg, bar);
        }
    }
    const int lane = threadIdx.x & 31;
    const uint32_t gw = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    // asynchronous-ring layouts: the lane groups' rings follow the staged negatives
    const uint32_t ring_base = smem_u32(smem_raw) + 128u + neg_bytes;
    uint32_t rows_ok = 0;
    if (p.late_wait && !negs && blockIdx.x == 0 && threadIdx.x == 0) {
        // no staging warp in this launch: thread 0 of CTA 0 publishes the predecessor's step itself
        if (p.pdl) pdl_wait();
        publish_predecessor(p);
    }
    const uint32_t t_base = gw * L::G;
    if (t_base < bv.n_items) process_items<L, MODEL>(p, bv, t_base, s_neg, bar, lane, p.lut, ring_base, rows_ok);
    // the CTA's shared memory must stay allocated until the bulk copies have landed
    if (negs) mbar_wait(bar, 0);
    peer_signal(p);
}

// ------------------------------------------------------------------ dataflow epoch kernel
// The whole epoch in ONE ordinary launch with no barrier at all (f2v_set_epoch_mode 2).  The epoch's
// items (every minibatch's list, padded so that a warp's lane groups share a minibatch) are handed
// out in order through one ticket counter; a warp that is about to read rows of the next table waits
// -- flow_wait() -- only until the minibatches that write those rows are complete.  At the reference's
// batch sizes (256 / 384) almost no row read depends on the few minibatches in flight, so thousands
// of dependent minibatches overlap instead of paying a launch + drain (or a grid barrier) each, with
// exactly the reference's Jacobi semantics: a row below the minibatch's split is read from the next
// table after its writer finished, any other row from the current table, which nobody writes.
// Shared negatives (bs=0) are gathered from L2 here (a CTA works on several minibatches at once).
struct FlowParams {
    BatchParams p;               // items / hub: the epoch's plan; neg: the epoch's negative stream
    uint32_t total_items;
    uint32_t neg_stride;         // negative indices per minibatch
    uint32_t* ticket;            // next item (zeroed before the launch)
};
template <class L, int MODEL>
__global__ void __launch_bounds__(kWarpsPerCta * 32, L::MINB)
force_flow_kernel(const FlowParams fp) {
    const BatchParams& p = fp.p;
    const int lane = threadIdx.x & 31;
    uint32_t rows_ok = 0;
    constexpr uint32_t kTake = 4;            // lane-group rounds per ticket (one atomic per kTake * G items)
    for (;;) {
        uint32_t t0 = 0;
        if (lane == 0) t0 = atomicAdd(fp.ticket, kTake * (uint32_t)L::G);
        t0 = __shfl_sync(kFull, t0, 0);
        if (t0 >= fp.total_items) break;
        // a warp works through its rounds in order, so everything it can wait for still belongs to items
        // handed out earlier (to this warp or to others): the no-deadlock argument is unchanged
        for (uint32_t r = 0; r < kTake; r++) {
            const uint32_t t_base = t0 + r * (uint32_t)L::G;
            if (t_base >= fp.total_items) break;
            // minibatch of this round's items: that of the first one (padding keeps the lane groups of a
            // round inside one minibatch; a round that starts with a padding item is all padding)
            const uint32_t v0 = p.items[t_base].v;
            if (v0 == kNoVertex) continue;
            const uint32_t b = v0 / p.flow_batch;
            BatchVar bv;
            bv.items = p.items;
            bv.hub = p.hub;
            bv.n_items = fp.total_items;
            bv.split = b * p.flow_batch;
            bv.lo = bv.split;
            bv.neg = p.neg + (size_t)b * fp.neg_stride;
            process_items<L, MODEL>(p, bv, t_base, nullptr, nullptr, lane, p.lut, 0u, rows_ok);
        }
    }
}

// A launch with no rows on this rank still takes part in the exchange barrier; the epoch ends with
// a wait for every peer's last step (then this replica is complete).  One CTA.  publish_step is
// published BEFORE the wait (everything earlier in the stream is complete), signal_step after it.
__global__ void peer_sync_kernel(const BatchParams p) {
    publish_predecessor(p);
    peer_wait(p);
    peer_signal(p);
}

// Replicate a row range of this rank's table into every other replica (multi-GPU host-buffer
// epoch: each rank uploads 1/world of the table over its own PCIe link and the rest travels over
// NVLink).  dst of element i: the multicast mapping (one store reaches all) or each peer's table.
struct BcastParams {
    const float* src;
    float* mc;
    float* peer[kMaxWorld - 1];
    uint32_t n_store;
    uint64_t count;                      // float4 elements (VEC) or floats
};
template <bool VEC>
__global__ void bcast_rows_kernel(const BcastParams b) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < b.count; i += stride) {
        if (VEC) {
            const float4 v = __ldcg(reinterpret_cast<const float4*>(b.src) + i);
            if (b.mc) mc_st_v4(b.mc + 4 * i, v.x, v.y, v.z, v.w);
            for (uint32_t r = 0; r < b.n_store; r++) __stcg(reinterpret_cast<float4*>(b.peer[r]) + i, v);
        } else {
            const float v = __ldcg(b.src + i);
            if (b.mc) mc_st_f32(b.mc + i, v);
            for (uint32_t r = 0; r < b.n_store; r++) __stcg(b.peer[r] + i, v);
        }
    }
}

// Row-sharded tables: move rows [first, first+count) between an identity-layout staging buffer and
// the flat sharded table.  TO_TABLE: only the rows this rank stores are written (every rank runs
// the same upload, each fills its own shard); else every row is read (peer loads for remote shards).
template <bool TO_TABLE>
__global__ void shard_copy_kernel(float* table, float* staging, uint64_t first, uint64_t count, uint32_t dim,
                                  uint32_t lg, uint32_t shard_rows, uint32_t rank) {
    const uint64_t total = count * dim;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const uint64_t r = i / dim;
        const uint32_t k = (uint32_t)(i - r * dim);
        const uint32_t j = (uint32_t)(first + r);
        const uint32_t row = shard_row(j, lg, shard_rows);
        if (TO_TABLE) {
            if (row / shard_rows == rank) table[(size_t)row * dim + k] = staging[i];
        } else {
            staging[i] = __ldcg(table + (size_t)row * dim + k);
        }
    }
}

// Order-independent 64-bit checksum of the live table: sum over (vertex v, component k) of
// mix(v*dim + k, bits of X[v][k]) mod 2^64.  Any single differing bit changes the sum; used to
// compare a multi-GPU replica / sharded table with a single-GPU run without moving the tables.
__device__ __forceinline__ uint64_t checksum_mix(uint64_t pos, uint32_t bits) {
    uint64_t z = (pos + 1) * 0x9E3779B97F4A7C15ULL + bits;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__global__ void checksum_kernel(const float* table, uint64_t n, uint32_t dim, uint32_t lg, uint32_t shard_rows,
                                unsigned long long* out) {
    const uint64_t total = n * dim;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const uint64_t v = i / dim;
        const uint32_t k = (uint32_t)(i - v * dim);
        const uint32_t row = shard_row((uint32_t)v, lg, shard_rows);
        acc += checksum_mix(i, __float_as_uint(__ldcg(table + (size_t)row * dim + k)));
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(kFull, acc, off);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, (unsigned long long)acc);
}

// Counter-based draw for the device walk sampler (host mirror: oracle f2vo_counter_rand).
__host__ __device__ __forceinline__ uint32_t counter_rand(uint64_t seed, uint64_t epoch, uint64_t vertex, uint32_t step) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL * (vertex * 8u + step + 1u) + 0xD1B54A32D192ED03ULL * (epoch + 1u);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return (uint32_t)(z >> 33);
}

// Semi-random walks, one thread per start vertex (rule of algorithms.cpp:1097-1118).
__global__ void walk_kernel(uint64_t n, uint64_t nnz, const uint64_t* __restrict__ rowptr,
                            const uint32_t* __restrict__ colids, uint32_t* __restrict__ walks,
                            uint64_t seed, uint64_t epoch) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t w = i;
#pragma unroll
    for (int l = 0; l < kWalkLen; l++) {
        const uint64_t r0 = __ldg(rowptr + w), r1 = __ldg(rowptr + w + 1);
        const uint64_t dg = r1 - r0;
        uint64_t e = w;                       // vertex id used as an edge index (reference quirk)
        if (dg > 2) e = r0 + counter_rand(seed, epoch, i, (uint32_t)l) % (uint32_t)(dg - 1);
        else if (dg == 2) e = r0;
        const uint32_t nx = e < nnz ? __ldg(colids + e) : (uint32_t)w;
        walks[i * kWalkLen + l] = nx;
        w = nx;
    }
}

}  // namespace f2v
