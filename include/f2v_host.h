/* include/f2v_host.h -- C ABI of the host side of the Force2Vec drop-in (same library,
 * libf2v.so): the pieces of the reference that sit either side of the force step and
 * that the engine needs bit-compatible -- the glibc rand() stream and the samplers that
 * consume it, the MatrixMarket -> CSR loader, the .embd writer -- plus a synthetic R-MAT
 * generator and the whole-run driver that replaces algorithms::AlgoForce2Vec*().
 * All arithmetic of the force step itself runs on the GPU through include/f2v.h.
 */
#ifndef F2V_HOST_H
#define F2V_HOST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- glibc srand()/rand() compatible stream (TYPE_3 additive feedback) ---------------
 * The reference seeds once with srand(1) (Test/Force2Vec.cpp:126) and consumes rand() in
 * serial sections only (algorithms.cpp:42,50,56), so a private generator with the same
 * recurrence reproduces its whole sample stream without touching libc state.            */
typedef struct f2v_rng f2v_rng;
f2v_rng* f2v_rng_create(uint32_t seed);
void     f2v_rng_destroy(f2v_rng* g);
int32_t  f2v_rng_next(f2v_rng* g);

/* randInitF (model 5, algorithms.cpp:47-53) / randInit (models 6,7, :38-45): n*dim draws. */
int f2v_init_embeddings(f2v_rng* g, int model, uint64_t n, uint32_t dim, float* X);
/* init_SM_TABLE (algorithms.cpp:757-764): 2048 floats.                                    */
int f2v_build_lut(float* table2048);

/* Entries of one epoch's negative stream in the ENGINE layout (include/f2v.h,
 * f2v_set_negatives): ceil(n/batch) * (bs_mode ? batch+s-1 : s).                          */
uint64_t f2v_neg_stream_len(int model, uint64_t n, uint32_t batch, uint32_t s, int bs_mode);
/* Draw one epoch's negatives exactly as the reference does -- per minibatch s draws, or
 * s*batch draws for bs=1 of which only the first batch+s-1 are kept (algorithms.cpp:
 * 577-578, 686-687, 812-816, 964-967; model 7: range min((b+1)*batch, n-1), :1125-1126). */
int f2v_draw_epoch_negatives(f2v_rng* g, int model, uint64_t n, uint32_t batch, uint32_t s,
                             int bs_mode, uint32_t* out);
/* Semi-random walks off the same stream, the reference's serial loop bit for bit (algorithms.cpp:1097-1118;
 * n*5 entries, and the stream is left where that loop leaves it).  Internally 24 walks are in flight at
 * speculative stream positions so that their cache misses overlap; they are committed in order.          */
int f2v_draw_walks(f2v_rng* g, uint64_t n, uint64_t nnz, const uint64_t* rowptr,
                   const uint32_t* colids, uint32_t* walks);

/* ---- graph IO ------------------------------------------------------------------------
 * MatrixMarket text -> CSR with the reference's semantics (IO.h:59-156, CSC.h:146-188,
 * CSR.h:154-186): 1-based ids; "symmetric" mirrors off-diagonal entries and drops
 * self-loops; general keeps entries as they are; duplicates are kept; column ids ascend
 * within a row.  rowptr/colids are malloc'ed; release with f2v_free.                      */
int  f2v_load_mtx(const char* path, uint64_t* n, uint64_t* nnz, uint64_t** rowptr, uint32_t** colids);
void f2v_free(void* p);
/* "N D" header, then "id v1 .. vD " per row with 6 significant digits and a trailing
 * space (algorithms.h:118-136).                                                           */
int  f2v_write_embd(const char* path, const float* X, uint64_t n, uint32_t dim);
/* Binary CSR cache (magic "F2VCSR01", n, nnz, rowptr u64, colids u32): the fast path for graphs
 * whose text form takes minutes to parse.  f2v_load_csr validates the arrays; release with f2v_free.
 * The CLI reads it when -input ends in ".f2vcsr".                                           */
int  f2v_write_csr(const char* path, uint64_t n, uint64_t nnz, const uint64_t* rowptr, const uint32_t* colids);
int  f2v_load_csr(const char* path, uint64_t* n, uint64_t* nnz, uint64_t** rowptr, uint32_t** colids);
/* "%.6g" of one value as the writer formats it (fast path + printf fall-back); out32: >= 32 bytes.
 * Exposed for tests.                                                                       */
int  f2v_format_g6(float v, char* out32);
/* Writes the lower triangle as "%%MatrixMarket matrix coordinate pattern symmetric".      */
int  f2v_write_mtx(const char* path, uint64_t n, const uint64_t* rowptr, const uint32_t* colids);

/* Graph500-style R-MAT (a,b,c,d = .57,.19,.19,.05), n = 2^scale, edge_factor*n edge draws,
 * ids not permuted, symmetrised, self-loops dropped, duplicates removed, rows sorted.     */
int  f2v_rmat_csr(int scale, int edge_factor, uint64_t seed, uint64_t* n, uint64_t* nnz,
                  uint64_t** rowptr, uint32_t** colids);

/* ---- work plan (exposed for tests of the host-side scheduling / multi-GPU slicing) ----
 * The plan the engine builds for rows [first_row, first_row+nrows) in minibatches of
 * `batch`: per minibatch, hub rows (degree > chunk; with par > 0 the chunk of a minibatch is
 * clamp(edges/par, batch <= 8192 ? 16 : 8, chunk)) cut into chunks, then the other rows by
 * descending degree class; with world > 1 only the rows of each minibatch that `rank` owns:
 * assign 0 = its contiguous slice of batch/world rows (NCCL all-gather exchange), assign 1 =
 * degree-balanced greedy partition (peer-store exchange); assign | 2 = the lightest rows
 * (degree 0..3) are scheduled right after the hub chunks instead of last.  items: 16-byte records {u32 v; u32 len (bit 31 = hub chunk); u64 e0};
 * hub: 16-byte records {u32 chunk; u32 nchunks; u32 slot; u32 deg}, parallel to items.
 * All four arrays are malloc'ed (f2v_free).                                                */
int f2v_plan_build(const uint64_t* rowptr, uint64_t first_row, uint64_t nrows, uint32_t batch,
                   uint32_t chunk, uint32_t par, int walk, int rank, int world, int assign, uint64_t* nb,
                   uint64_t** item_ptr, uint32_t** n_hub, void** items, void** hub);

/* ---- whole-run driver ------------------------------------------------------------------
 * What algorithms::AlgoForce2VecNS/NSBS/NSRW/NSRWBS/NSRWEFF do between omp_get_wtime()
 * start and end (algorithms.cpp:557-647 etc.): init from the stream, `iterations` epochs,
 * result in X_out (n*dim).  seconds = wall time of that span.  walk_sampler: 0 = host
 * walks off the libc-compatible stream (exact reference stream), 1 = device sampler.      */
typedef struct f2v_train_args {
    uint64_t n, nnz;
    const uint64_t* rowptr;
    const uint32_t* colids;
    uint32_t dim;
    int option;            /* 5, 6 or 7 */
    int bs;                /* 0 / 1 */
    uint32_t iterations;
    uint32_t batch;
    uint32_t nsamples;
    float lr;
    uint32_t seed;         /* srand seed; the reference uses 1 */
    int device;
    int walk_sampler;
    int epoch_mode;        /* f2v_set_epoch_mode */
    uint32_t chunk;        /* 0 = default */
} f2v_train_args;
int f2v_train(const f2v_train_args* a, float* X_out, double* seconds);
/* The same run on devices a->device .. a->device+gpus-1 of this node, driven from this process
 * (one host thread per GPU; replicated tables, the exchange fused into the force kernel).  The
 * result equals f2v_train's bit for bit for equal `chunk` (0 = the default, which differs when a
 * rank's share of a minibatch is below 16 K rows: 64 instead of 256 / 128).  gpus <= 1 is f2v_train.                                                 */
int f2v_train_gpus(const f2v_train_args* a, int gpus, float* X_out, double* seconds);

#ifdef __cplusplus
}
#endif
#endif
