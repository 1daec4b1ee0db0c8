/* oracle/f2v_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, CPU restatement of the reference's minibatch force path
 * (HipGraph/Force2Vec, sample/algorithms.cpp, options 5 / 6 / 7).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product (force2vec_b200/, bin/Force2Vec) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 *   (1) the reference's shipped golden embedding
 *       datasets/output/cora.mtxF2VNS384D128IT1200NS5.embd (copied as a data
 *       fixture to tests/golden/), and
 *   (2) outputs of the unmodified reference compiled here (oracle/_ref, see
 *       oracle/Makefile) -- committed under tests/golden/ by
 *       tests/golden/make_golden.py.
 */
#ifndef F2V_ORACLE_H
#define F2V_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { F2VO_TDIST = 5, F2VO_SIGMOID = 6, F2VO_WALK = 7 };
#define F2VO_WALKLEN 5
#define F2VO_LUT_SIZE 2048

/* glibc srand()/rand() (TYPE_3 additive feedback) re-implemented without libc state. */
typedef struct { int32_t r[31]; int f, b; } f2vo_rng;
void     f2vo_srand(f2vo_rng* g, uint32_t seed);
int32_t  f2vo_rand(f2vo_rng* g);

/* algorithms.cpp:38-53  randInit (opt 6/7) / randInitF (opt 5); n*dim draws, row-major. */
void f2vo_init_embeddings(f2vo_rng* g, int model, uint64_t n, uint32_t dim, float* X);

/* algorithms.cpp:757-764  sigmoid table, 2048 entries (+ entry 2048 := 1.0f, see .c). */
void f2vo_build_lut(float* lut2049);
/* algorithms.cpp:766-770 */
float f2vo_fast_sm(const float* lut, float v);

/* algorithms.cpp:1097-1118  semi-random walks for all n vertices, serial rand() stream. */
void f2vo_walks(f2vo_rng* g, uint64_t n, uint64_t nnz, const uint64_t* rowptr,
                const uint32_t* colids, uint32_t* walks /* n*5 */);

/* Number of rand() draws / index entries one minibatch consumes (Q2):
 *   bs=0: s;  bs=1: s*batch (always, even for the partial last batch).        */
uint64_t f2vo_draws_per_batch(int bs, uint32_t batch, uint32_t s);

/* Draw one minibatch's negative indices from the stream into idx[] (draws_per_batch
 * entries).  model 5/6: rand()%(n-1) (algorithms.cpp:55-58,578,815);
 * model 7: rand()%min((b+1)*batch, n-1) (algorithms.cpp:1125-1126).             */
void f2vo_draw_negatives(f2vo_rng* g, int model, int bs, uint64_t n, uint32_t batch,
                         uint32_t s, uint64_t b, uint32_t* idx);

/* One Jacobi minibatch [lo,hi) in place on X (algorithms.cpp:588-639 / 694-747 /
 * 833-921 / 985-1047 / 1142-1193).  idx = this batch's negative indices as drawn
 * (bs=1: vertex k uses idx[k .. k+s-1], the overlapping-window quirk at
 * algorithms.cpp:719-720,1029-1030).  walks only for model 7.  lut only for 6/7.
 * threads<=0 -> omp default.                                                      */
void f2vo_step(int model, int bs, uint64_t n, uint32_t dim, const uint64_t* rowptr,
               const uint32_t* colids, float* X, uint64_t lo, uint64_t hi,
               const uint32_t* idx, uint32_t s, float lr, const float* lut,
               const uint32_t* walks, int threads);

/* Whole run as the reference does it after srand(1) (Test/Force2Vec.cpp:126 ->
 * algorithms.cpp:544-652 etc.): init, then `iterations` epochs of ceil(n/batch)
 * minibatches.  If neg_log != NULL it receives every negative index drawn, in
 * order (iterations * nbatches * draws_per_batch entries); if walk_log != NULL it
 * receives each epoch's walks (iterations * n * 5).  X_init (nullable) receives
 * the initial embedding.  Returns 0, or -1 on bad arguments.                     */
int f2vo_run(int model, int bs, uint64_t n, uint64_t nnz, const uint64_t* rowptr,
             const uint32_t* colids, uint32_t dim, uint32_t iterations, uint32_t batch,
             uint32_t s, float lr, uint32_t seed, int threads,
             float* X_out, float* X_init, uint32_t* neg_log, uint32_t* walk_log);

/* Counter-based walk sampler (NOT in the reference: host mirror of the device
 * sampler kernel, force2vec_b200/csrc/f2v_kernels.cu f2v_walk_kernel).  Same
 * walk rule as algorithms.cpp:1097-1118 but the draw for (epoch, vertex, step) is
 * f2vo_counter_rand(seed, epoch, vertex, step) instead of the serial libc stream. */
uint32_t f2vo_counter_rand(uint64_t seed, uint64_t epoch, uint64_t vertex, uint32_t step);
void f2vo_walks_counter(uint64_t seed, uint64_t epoch, uint64_t n, uint64_t nnz,
                        const uint64_t* rowptr, const uint32_t* colids, uint32_t* walks);

#ifdef __cplusplus
}
#endif
#endif
