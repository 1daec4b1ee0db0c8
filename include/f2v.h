/* include/f2v.h -- C ABI of the B200-native Force2Vec force-step engine (libf2v.so).
 *
 * The reference (HipGraph/Force2Vec) has no FFI/plugin interface; its only entry
 * points for this path are the C++ methods
 *     vector<float> algorithms::AlgoForce2VecNS | NSBS | NSRW | NSRWBS | NSRWEFF
 *         (INDEXTYPE ITERATIONS, INDEXTYPE NUMOFTHREADS, INDEXTYPE BATCHSIZE,
 *          INDEXTYPE ns, VALUETYPE lr)                  sample/algorithms.h:86-90
 * operating on `graph` (CSR) and `nCoordinates` (row-major n x DIM fp32,
 * sample/algorithms.h:53-54,68), and the unused per-vertex generator signature
 * sample/kgen/genDimFrc.base:36-57.  The functions below are what a C++ host that
 * replaces those method bodies binds; each one cites the
 * reference lines whose work it takes over.
 *
 * Conventions: plain pointers and sizes only; every function returns F2V_OK (0) or a
 * negative F2V_ERR_*; the message is available from f2v_last_error() (thread-local).
 * No exceptions cross the boundary.  The caller owns every host buffer; the library
 * owns all device memory.  One host thread drives one engine (one GPU).  There is no
 * CPU fallback: without a usable CUDA device every compute call fails with
 * F2V_ERR_CUDA.
 */
#ifndef F2V_H
#define F2V_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define F2V_ABI_VERSION 1

/* model ids = the reference CLI's -option values (Test/Force2Vec.cpp:139-150) */
#define F2V_TDIST   5   /* tForce2Vec: algorithms.cpp:544-652 (bs=0), 654-753 (bs=1)  */
#define F2V_SIGMOID 6   /* sForce2Vec: algorithms.cpp:778-932 (bs=0), 934-1060 (bs=1) */
#define F2V_WALK    7   /* rForce2Vec: algorithms.cpp:1063-1203                        */
#define F2V_WALKLEN 5   /* WALKLENGTH, algorithms.cpp:1074                             */
#define F2V_LUT_SIZE 2048 /* SM_TABLE_SIZE, algorithms.h:43                            */

#define F2V_OK          0
#define F2V_ERR_ARG    -1
#define F2V_ERR_CUDA   -2
#define F2V_ERR_NCCL   -3
#define F2V_ERR_STATE  -4
#define F2V_ERR_NOMEM  -5

typedef struct f2v_engine f2v_engine;

const char* f2v_last_error(void);
int  f2v_abi_version(void);
/* number of CUDA devices visible, or a negative error (no driver / no GPU) */
int  f2v_device_count(void);

/* ---- lifetime --------------------------------------------------------------------
 * f2v_create replaces the `algorithms` constructor (algorithms.h:60-70): it takes the
 * CSR (rowptr u64[n+1], colids u32[nnz], ascending within a row as CSR.h:154-186
 * produces) and uploads it to device `device_id`; the n x dim embedding table is
 * allocated on the device.  dim: any value >= 1 (fast paths for 32/64/128/256).      */
int f2v_create(f2v_engine** out, int device_id, uint64_t n, uint64_t nnz,
               const uint64_t* rowptr, const uint32_t* colids, uint32_t dim);
int f2v_destroy(f2v_engine* e);
/* Run on a caller-provided cudaStream_t (e.g. torch's current stream) instead of the
 * engine's own stream.  Pass NULL to go back to the engine's stream.                 */
int f2v_set_stream(f2v_engine* e, void* cuda_stream);
int f2v_sync(f2v_engine* e);

/* Page-locked host memory for the buffers handed to the calls below (copies from
 * pageable memory work too, at lower PCIe throughput).                               */
int f2v_host_alloc(void** p, uint64_t bytes);
int f2v_host_free(void* p);
/* Page-lock / release a range of caller-owned host memory in place (e.g. only the row range of a
 * full-size table that a multi-GPU rank moves over its own PCIe link in f2v_run_epoch_host).   */
int f2v_host_register(void* p, uint64_t bytes);
int f2v_host_unregister(void* p);
/* Free / total memory of the engine's device in bytes (reporting; either pointer may be NULL). */
int f2v_device_memory(const f2v_engine* e, uint64_t* free_bytes, uint64_t* total_bytes);

/* ---- state -----------------------------------------------------------------------
 * nCoordinates in / out (algorithms.h:54).  Host buffers of n*dim floats, row-major. */
int f2v_set_embeddings(f2v_engine* e, const float* X_host);
int f2v_get_embeddings(f2v_engine* e, float* X_host);
int f2v_get_rows(f2v_engine* e, uint64_t first_row, uint64_t nrows, float* rows_host);
/* Order-independent 64-bit checksum of the live table, computed on the device (sum over every
 * (vertex, component) of a hash of its position and bit pattern): equal tables give equal sums, any
 * differing bit changes it.  How a multi-GPU replica or a row-sharded table (remote shards are read
 * over NVLink) is compared with a single-GPU run at sizes where moving the tables is not an option
 * -- e.g. R-MAT 26 at d=128 (32 GiB), which the reference cannot run at all (algorithms.h:40,68:
 * 32-bit n*DIM).  Synchronises.                                                               */
int f2v_checksum(f2v_engine* e, uint64_t* checksum);
/* The sigmoid table built on the host with the reference expression
 * (init_SM_TABLE, algorithms.cpp:757-764): `count` = 2048 entries; entry 2048 (the
 * reference's out-of-bounds read for v == 6.0f exactly) is defined as 1.0f.          */
int f2v_set_lut(f2v_engine* e, const float* sm_table, uint32_t count);

/* ---- sample streams (made resident on the device) -----------------------------------
 * Negative indices for the coming step/epoch, in the order the reference draws them
 * (randIndex, algorithms.cpp:55-58 at :578,:687,:816,:967,:1126).  Layout per
 * minibatch b with stride W:  bs_mode 0 (and model 7): W = s;  bs_mode 1: W = batch+s-1
 * -- the first batch+s-1 of the reference's s*batch draws, the only ones it reads
 * (vertex k uses entries k..k+s-1, algorithms.cpp:719-720,1029-1030).                 */
int f2v_set_negatives(f2v_engine* e, const uint32_t* idx_host, uint64_t count);
/* Several epochs' streams may be uploaded at once; the next f2v_run_epoch reads the
 * resident stream from entry `offset` on (f2v_set_negatives resets it to 0).            */
int f2v_set_negative_offset(f2v_engine* e, uint64_t offset);
/* Walk samples, n*5 u32 (walksamples, algorithms.cpp:1075,1097-1118).                  */
int f2v_set_walks(f2v_engine* e, const uint32_t* walks_host);
int f2v_get_walks(f2v_engine* e, uint32_t* walks_host);
/* Device semi-random-walk sampler: same rule as algorithms.cpp:1097-1118, but the draw
 * for (epoch, vertex, step) comes from a counter-based generator (documented in
 * DESIGN.md; host mirror in oracle/f2v_oracle.c:f2vo_walks_counter) instead of the
 * serial libc stream, so it is parallel.  Fills the resident walk buffer.             */
int f2v_sample_walks(f2v_engine* e, uint64_t seed, uint64_t epoch);

/* ---- the hot path ------------------------------------------------------------------
 * f2v_step: ONE Jacobi minibatch over rows [first_row, first_row+nrows): every read of
 * X sees the pre-step table, then the rows are replaced (algorithms.cpp:588-639,
 * 833-921, 1142-1193).  neg_idx_host: s entries (bs_mode 0 / model 7) or nrows+s-1
 * (bs_mode 1).  walks_host: n*5 or NULL to use the resident walks.  Teacher-forced
 * entry point used by the parity tests; synchronous.  Hub rows are cut at 128 edges
 * (= f2v_run_epoch with chunk 128: bit-identical to the epoch's minibatch for that chunk). */
int f2v_step(f2v_engine* e, int model, uint64_t first_row, uint32_t nrows,
             const uint32_t* neg_idx_host, uint32_t s, int bs_mode, float lr,
             const uint32_t* walks_host);

/* f2v_run_epoch: one epoch = ceil(n/batch) dependent minibatches over contiguous row
 * ranges in natural order (algorithms.cpp:569-640), consuming the resident negative
 * stream (ceil(n/batch)*W entries) and, for model 7, the resident walks.  Same result
 * as a loop of f2v_step.  Asynchronous on the engine's stream.
 * chunk: hub rows longer than `chunk` edges are split across warps (0 = default: 256 for dim >= 128
 * and batches of >= 16 K rows, else 128; 64 on a multi-GPU engine whose share of a minibatch is below 16 K rows).  Results are bit-identical
 * across world sizes and epoch modes for equal `chunk`.                                */
int f2v_run_epoch(f2v_engine* e, int model, uint32_t batch, uint32_t s, int bs_mode,
                  float lr, uint32_t chunk);

/* Host-buffer epoch (the end-to-end call): upload X_in (nullable = keep resident),
 * the epoch's negatives and walks (nullable), run the epoch, download into X_out
 * (nullable).  Synchronous.  Single GPU: finished rows are copied back while later
 * minibatches still compute.  Multi-GPU engine with the peer exchange: every rank passes
 * the same full-size buffers but moves only rows [rank*ceil(n/world), ...) -- its share
 * -- over its PCIe link (the other replicas receive them over NVLink), and only that
 * row range of X_out is written on this rank.                                          */
int f2v_run_epoch_host(f2v_engine* e, int model, uint32_t batch, uint32_t s, int bs_mode,
                       float lr, uint32_t chunk, const float* X_in,
                       const uint32_t* neg_idx_host, uint64_t neg_count,
                       const uint32_t* walks_host, float* X_out);

/* Execution mode of f2v_run_epoch: 0 = one kernel launch per minibatch, chained by programmatic
 * dependent launch (default); 2 = the dataflow epoch: ONE ordinary launch per epoch, items handed out
 * in order, a warp waits only for the minibatches that wrote the rows it is about to read (single-GPU
 * engines; same bits as mode 0).  Mode 1 (a persistent kernel with a grid barrier per minibatch) lost
 * to mode 0 at every batch size and was removed.                                                    */
int f2v_set_epoch_mode(f2v_engine* e, int mode);
/* Tuning knobs (integers; defaults are the measured optimum, profiles/r1_tune_v5.md, r2_tune.md):
 *   "variant"   lane layout of the d=128 / d=64 kernels; -1 (default) = by launch size.  d=128:
 *               3 = 16 lanes per row, 2 rows in flight per group, 4 CTAs/SM at 64 registers;
 *               8 = the same at 5 CTAs/SM and 48 registers (launches with >= 48 K items);
 *               11 = 8 rows in flight at 128 registers (launches with < 12 K items); 0 = 4 rows in
 *               flight, 3 CTAs/SM; 21 / 22 = asynchronous shared-memory ring (cp.async stages, rows never
 *               land in registers; neighbours and per-vertex negatives in one stream), 3 stages at
 *               4 CTAs/SM / 4 stages at 3 CTAs/SM -- same bits, measured equal or slower (r2_tune.md).
 *               d=64: 0 = 4 rows in flight, 1 = 2 rows in flight at 4 CTAs/SM, 4 = at 5 CTAs/SM, 21 / 22.
 *   "neg_smem"  0 disables the TMA staging of shared negatives (gathered from L2 instead)
 *   "par"       lane groups a minibatch should fill (adaptive chunk length, 0 = fixed chunk)
 *   "min_chunk" lower bound of the adaptive chunk length (0 = default: 16 for batches <= 8192, else 8)
 *   "pdl"       programmatic dependent launch of consecutive minibatches: 0 off, 1, 2 (default)
 *   "auto_flow" epoch mode 0 switches to the dataflow epoch for batches up to this size (default 0 = never)
 *   "multicast" peer exchange through NVLink multicast stores: 1 (default) when supported, 0 never
 *   "multicast_in_process"  1 = the engines of this process are each driven by their own host thread, so
 *               they may set up the multicast exchange among themselves (f2v_train_gpus sets it when
 *               F2V_INPROC_MULTICAST=1; default off: that set-up has only run between processes so far)
 *   "sharded"   1 = row-sharded tables (set on every rank before f2v_comm_peer_export)
 *   "peer_sig"  who publishes a minibatch's exchange step: 2 (default) = CTA 0 of the next launch after
 *               its dependency wait; 1 = a 1-CTA kernel after the force kernel; 0 = the last CTA
 *   "trace"     1 = record per-minibatch times (f2v_trace_ms)
 *   "exchange_timeout_ms"  multi-GPU: how long a launch waits for a peer's exchange step before it
 *               gives up (default 30000; 0 = for ever); f2v_sync / f2v_get_embeddings then fail
 *   "order", "peer_debug"  development probes (tools/mgpu_probe.py)                                  */
int f2v_set_option(f2v_engine* e, const char* name, int64_t value);
/* Kernel launches issued by this engine since creation (force + sampler kernels).     */
uint64_t f2v_launch_count(const f2v_engine* e);
/* Device time of the last f2v_run_epoch in milliseconds (CUDA events on its stream);
 * synchronises.                                                                       */
int f2v_last_epoch_ms(f2v_engine* e, float* ms);

/* Per-minibatch device times of the last f2v_run_epoch (epoch mode 0) when the option "trace"
 * is 1: ms[b] = time from the end of minibatch b-1 to the end of minibatch b (its exchange
 * included).  Development / profiling aid.                                               */
int f2v_trace_ms(f2v_engine* e, float* ms, uint32_t cap, uint32_t* count);

/* ---- multi-GPU (one process per GPU) -----------------------------------------------
 * Every rank holds a full replica of X and the CSR.  Baseline exchange: each minibatch is
 * split into `world` contiguous slices; a rank updates its slice and the slices are exchanged
 * with an NCCL all-gather before the next minibatch.  id128: the 128-byte
 * ncclUniqueId produced by f2v_comm_unique_id on rank 0 and broadcast by the caller.
 * batch must be a multiple of world.                                                  */
int f2v_comm_unique_id(void* id128);
int f2v_comm_init(f2v_engine* e, const void* id128, int rank, int world);

/* Peer-store exchange (the default for N > 1): the exchange is fused into the force kernel.
 * Every rank maps the other ranks' tables (CUDA IPC between processes, direct peer access
 * between engines of one process); the kernel stores each finished row into its own replica
 * AND straight into every peer's replica over NVLink, so the transfer overlaps the compute row
 * by row.  Minibatches are separated by one system-scope step counter per rank: CTA 0 of launch
 * b+1 publishes step b once launch b is complete (the launches stay chained by programmatic
 * dependent launch), and a warp of launch b+1 waits for the peers' counters LAZILY -- only when
 * the ids it is about to gather lie in a minibatch a peer has not published yet -- so the
 * exchange latency hides behind the items that do not depend on the previous minibatch.  No host
 * round trip, no collective call.  Rows of a minibatch are dealt to the ranks by a degree-balanced greedy
 * partition (no contiguity needed).  Results equal the single-GPU run bit for bit.
 *   1. every rank: f2v_comm_peer_export(e, blob)         blob: F2V_PEER_BLOB bytes
 *   2. the caller all-gathers the blobs in rank order (torch.distributed, MPI, a file ...)
 *   3. every rank: f2v_comm_peer_init(e, blobs, rank, world)   world <= 8
 * All ranks must then issue the same sequence of f2v_run_epoch calls.  Where the devices
 * support it the stores go through an NVLink multicast mapping (one store reaches every
 * replica; option "multicast" = 0 forces one store per peer).
 *
 * Row-sharded tables (option "sharded" = 1 on every rank before the export; power-of-two
 * world, one process per GPU): instead of a replica per GPU each GPU stores 1/world of the
 * rows of both tables (VMM allocations mapped by every rank into one flat virtual range; vertex
 * j lives in shard (j xor hash(j / world)) mod world).  Gathers of remote rows cross NVLink, a finished row is
 * stored once, in its shard.  For tables that do not fit one GPU; slower than replicas when
 * they do.  f2v_set_embeddings (every rank passes the same table, each keeps its rows) must
 * then be called by all ranks the same number of times; f2v_get_embeddings returns the full
 * table on every rank; f2v_step is not available.  Results stay bit-identical.            */
/* Row of a sharded table that holds `vertex` (the placement function of the row-sharded mode:
 * shard = row / shard_rows).  Exposed for tests.                                          */
uint32_t f2v_shard_row(uint32_t vertex, uint32_t log2_world, uint32_t shard_rows);
#define F2V_PEER_BLOB 256
int f2v_comm_peer_export(f2v_engine* e, void* blob);
int f2v_comm_peer_init(f2v_engine* e, const void* blobs, int rank, int world);

#ifdef __cplusplus
}
#endif
#endif /* F2V_H */
