// force2vec_b200/csrc/f2v_host.cpp -- host side of the drop-in (include/f2v_host.h):
// glibc-compatible rand() stream and the samplers that consume it, MatrixMarket loader,
// .embd writer, R-MAT generator, and the whole-run driver that replaces the bodies of
// algorithms::AlgoForce2Vec*() with calls into the GPU engine (include/f2v.h).
#include "../../include/f2v_host.h"
#include "../../include/f2v.h"
#include "f2v_host.hpp"
#include "f2v_plan.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <cerrno>
#include <cmath>
#include <sys/stat.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

// ------------------------------------------------------------------ rand() stream ------
// glibc random_r.c TYPE_3: degree 31, separation 3.  srandom: r[0] = seed, r[i] =
// 16807*r[i-1] mod (2^31-1) (Schrage), then 310 outputs are discarded; random: r[f] += r[b],
// output r[f] >> 1.  Call sites replaced: Test/Force2Vec.cpp:126, algorithms.cpp:42,50,56.
struct f2v_rng {
    uint32_t r[31];
    int f, b;
    inline uint32_t next() {
        uint32_t v = r[f] + r[b];
        r[f] = v;
        if (++f == 31) f = 0;
        if (++b == 31) b = 0;
        return v >> 1;
    }
};

extern "C" f2v_rng* f2v_rng_create(uint32_t seed) {
    f2v_rng* g = (f2v_rng*)malloc(sizeof(f2v_rng));
    if (!g) return nullptr;
    if (seed == 0) seed = 1;
    int32_t w = (int32_t)seed;
    g->r[0] = (uint32_t)w;
    for (int i = 1; i < 31; i++) {
        long hi = w / 127773, lo = w % 127773;
        long t = 16807 * lo - 2836 * hi;
        if (t < 0) t += 2147483647;
        w = (int32_t)t;
        g->r[i] = (uint32_t)w;
    }
    g->f = 3;
    g->b = 0;
    for (int i = 0; i < 310; i++) g->next();
    return g;
}
extern "C" void f2v_rng_destroy(f2v_rng* g) { free(g); }
extern "C" int32_t f2v_rng_next(f2v_rng* g) { return (int32_t)g->next(); }

// ---- jump-ahead.  The raw sequence obeys s[i] = s[i-31] + s[i-3] (mod 2^32), a linear recurrence
// with characteristic polynomial x^31 - x^28 - 1 over Z/2^32: s[m] = sum_k P_m[k] * w[k] with
// P_m = x^m mod that polynomial and w the current window of 31 values.  So the state after N draws
// costs O(31^2 log N) instead of N steps, and a long run of draws (the n*dim initial embedding: 2.1 G
// at R-MAT 24, d=128) can be cut into chunks that independent threads generate -- the same numbers,
// in the same places, as the reference's serial loop.
namespace {
struct RngPoly { uint32_t c[31]; };
inline void poly_reduce(uint32_t (&t)[61], RngPoly& out) {
    for (int k = 60; k >= 31; k--) { t[k - 3] += t[k]; t[k - 31] += t[k]; }       // x^k = x^(k-3) + x^(k-31)
    for (int k = 0; k < 31; k++) out.c[k] = t[k];
}
inline RngPoly poly_mul(const RngPoly& a, const RngPoly& b) {
    uint32_t t[61] = {0};
    for (int i = 0; i < 31; i++)
        for (int j = 0; j < 31; j++) t[i + j] += a.c[i] * b.c[j];
    RngPoly r;
    poly_reduce(t, r);
    return r;
}
inline RngPoly poly_mul_x(const RngPoly& a) {                                     // a * x
    RngPoly r;
    for (int k = 30; k >= 1; k--) r.c[k] = a.c[k - 1];
    r.c[0] = a.c[30];                    // x^31 = x^28 + 1
    r.c[28] += a.c[30];
    return r;
}
inline RngPoly poly_pow_x(uint64_t m) {                                           // x^m mod the polynomial
    RngPoly result{}, base{};
    result.c[0] = 1;
    base.c[1] = 1;
    while (m) {
        if (m & 1) result = poly_mul(result, base);
        base = poly_mul(base, base);
        m >>= 1;
    }
    return result;
}
// window of the generator, oldest value first: w[k] = s[i-31+k]
inline void rng_window(const f2v_rng& g, uint32_t (&w)[31]) {
    for (int k = 0; k < 31; k++) w[k] = g.r[(g.f + k) % 31];
}
// the generator as it will be after `steps` more draws, given P = x^steps
inline f2v_rng rng_jump(const uint32_t (&w)[31], RngPoly P) {
    f2v_rng out;
    for (int j = 0; j < 31; j++) {
        uint32_t v = 0;
        for (int k = 0; k < 31; k++) v += P.c[k] * w[k];
        out.r[j] = v;
        P = poly_mul_x(P);
    }
    out.f = 0;
    out.b = 28;
    return out;
}
template <bool TDIST>
inline void init_span(f2v_rng& g, float* X, uint64_t count) {
    const double denom = 2147483647.0 + 1.0;   // RAND_MAX + 1.0
    if (TDIST) for (uint64_t k = 0; k < count; k++) X[k] = (float)(-1.0 + 2.0 * (double)g.next() / denom);
    else for (uint64_t k = 0; k < count; k++) X[k] = (float)((double)g.next() / denom);
}
}  // namespace

extern "C" int f2v_init_embeddings(f2v_rng* g, int model, uint64_t n, uint32_t dim, float* X) {
    if (!g || !X) return f2v::host_fail(F2V_ERR_ARG, "f2v_init_embeddings: bad argument or malformed input");
    const uint64_t total = n * (uint64_t)dim;
    constexpr uint64_t kChunk = 1ull << 20;
    const uint64_t nchunks = (total + kChunk - 1) / kChunk;
    if (nchunks <= 1) {
        if (model == F2V_TDIST) init_span<true>(*g, X, total); else init_span<false>(*g, X, total);
        return F2V_OK;
    }
    // generator states at the chunk starts: P_(c*kChunk) by repeated multiplication with P_kChunk
    uint32_t w[31];
    rng_window(*g, w);
    std::vector<f2v_rng> start(nchunks + 1);
    start[0] = *g;
    const RngPoly step = poly_pow_x(kChunk);
    RngPoly P = step;
    for (uint64_t c = 1; c < nchunks; c++) { start[c] = rng_jump(w, P); P = poly_mul(P, step); }
    start[nchunks] = rng_jump(w, poly_pow_x(total));     // the stream continues here (negatives, walks)
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t c = 0; c < (int64_t)nchunks; c++) {
        f2v_rng local = start[c];
        const uint64_t lo = (uint64_t)c * kChunk, cnt = std::min(kChunk, total - lo);
        if (model == F2V_TDIST) init_span<true>(local, X + lo, cnt); else init_span<false>(local, X + lo, cnt);
    }
    *g = start[nchunks];
    return F2V_OK;
}

extern "C" int f2v_build_lut(float* t) {
    if (!t) return f2v::host_fail(F2V_ERR_ARG, "f2v_build_lut: bad argument or malformed input");
    for (int i = 0; i < F2V_LUT_SIZE; i++) {
        // VALUETYPE x = 2.0*SM_BOUND*i/SM_TABLE_SIZE - SM_BOUND; 1.0/(1+exp(-x)) with the float exp
        float x = (float)(2.0 * 6.0 * i / F2V_LUT_SIZE - 6.0);
        float e = 1 + std::exp(-x);
        t[i] = (float)(1.0 / e);
    }
    return F2V_OK;
}

extern "C" uint64_t f2v_neg_stream_len(int model, uint64_t n, uint32_t batch, uint32_t s, int bs_mode) {
    if (batch == 0) return 0;
    const uint64_t nb = (n + batch - 1) / batch;
    const uint64_t W = (bs_mode && model != F2V_WALK) ? (uint64_t)batch + s - 1 : (uint64_t)s;
    return nb * W;
}

extern "C" int f2v_draw_epoch_negatives(f2v_rng* g, int model, uint64_t n, uint32_t batch, uint32_t s,
                                        int bs_mode, uint32_t* out) {
    if (!g || !out || batch == 0 || n < 2) return f2v::host_fail(F2V_ERR_ARG, "f2v_draw_epoch_negatives: bad argument or malformed input");
    const uint64_t nb = (n + batch - 1) / batch;
    const bool window = bs_mode && model != F2V_WALK;
    const uint64_t W = window ? (uint64_t)batch + s - 1 : (uint64_t)s;
    const uint64_t draws = window ? (uint64_t)s * batch : (uint64_t)s;
    // bs=1: the reference consumes s*batch draws per minibatch but only ever reads the first
    // batch+s-1 of them; the unread ones are skipped with one jump-ahead per minibatch
    // (the skip length is the same every time, so its polynomial is computed once)
    const uint64_t skip = draws > W ? draws - W : 0;
    const bool jump = skip >= 4096;
    if (jump && nb > 1) {
        // every minibatch consumes exactly `draws` values: the generator states at the minibatch
        // starts follow from one polynomial, and the minibatches are drawn by independent threads
        const uint32_t maxv = (uint32_t)(n - 1);
        uint32_t w[31];
        rng_window(*g, w);
        std::vector<f2v_rng> start(nb + 1);
        start[0] = *g;
        const RngPoly step = poly_pow_x(draws);
        RngPoly P = step;
        for (uint64_t b = 1; b <= nb; b++) { start[b] = rng_jump(w, P); if (b < nb) P = poly_mul(P, step); }
#pragma omp parallel for schedule(static)
        for (int64_t b = 0; b < (int64_t)nb; b++) {
            f2v_rng local = start[b];
            uint32_t* o = out + (uint64_t)b * W;
            for (uint64_t k = 0; k < W; k++) o[k] = local.next() % maxv;
        }
        *g = start[nb];
        return F2V_OK;
    }
    RngPoly Pskip{};
    if (jump) Pskip = poly_pow_x(skip);
    for (uint64_t b = 0; b < nb; b++) {
        uint32_t maxv = (uint32_t)(n - 1);
        if (model == F2V_WALK) {
            uint64_t pre = (b + 1) * (uint64_t)batch;
            if (pre < maxv) maxv = (uint32_t)pre;
        }
        uint32_t* o = out + b * W;
        const uint64_t kept = std::min(draws, W);
        for (uint64_t k = 0; k < kept; k++) o[k] = g->next() % maxv;
        if (jump) {
            uint32_t w[31];
            rng_window(*g, w);
            *g = rng_jump(w, Pskip);
        } else {
            for (uint64_t k = kept; k < draws; k++) g->next();
        }
    }
    return F2V_OK;
}

// Semi-random walks off the libc-compatible stream (algorithms.cpp:1097-1118).  The reference's loop is serial by
// construction -- walk i+1 starts at the stream position where walk i stopped, and how many draws a walk consumes
// (one per visited vertex of degree > 2) depends on its path -- and every step is two dependent cache misses
// (rowptr[w], then colids[e]): ~260 ns per start vertex, 1.1 s per epoch at R-MAT 22 against 1.5 ms on the GPU.
// The numbers cannot change, but the misses can overlap: kWalkers walks are in flight at once, round-robin, each
// visit doing one half-step and issuing the prefetch for its next one.  A younger walk starts at the stream
// position its elders are PREDICTED to leave it at (every step still to be decided is assumed to consume a draw);
// when an elder meets a vertex that consumes nothing the positions of all younger walks move back by one and those
// that have already used a draw start over (R-MAT 22: 13 % of all steps use no draw, most of them the first step of
// an isolated start vertex, which is decided before anything younger starts; one walk in five meets a later one,
// one in three is restarted once).  Walks are committed in order, so the result -- and the stream position
// afterwards -- is the serial loop's, bit for bit (tests/test_host.py, against the oracle's serial loop).
namespace {
struct DrawRing {                    // draws [.., gen) of the stream; a window of kSize behind gen stays readable
    static constexpr uint64_t kSize = 4096;
    f2v_rng g;
    uint64_t gen = 0;
    uint32_t buf[kSize];
    explicit DrawRing(const f2v_rng& start) : g(start) {}
    inline uint32_t at(uint64_t k) {
        while (gen <= k) { buf[gen & (kSize - 1)] = g.next(); gen++; }
        return buf[k & (kSize - 1)];
    }
};
struct Walker {
    uint64_t i, w, e, o;             // start vertex, current vertex, edge index of the current step, first draw
    uint32_t c, decided, l, phase;   // draws used, steps whose draw decision is made, steps finished, 0 = needs rowptr / 1 = needs colids
    uint32_t out[F2V_WALKLEN];
};
}  // namespace

extern "C" int f2v_draw_walks(f2v_rng* g, uint64_t n, uint64_t nnz, const uint64_t* rowptr,
                              const uint32_t* colids, uint32_t* walks) {
    if (!g || !rowptr || !walks) return f2v::host_fail(F2V_ERR_ARG, "f2v_draw_walks: bad argument or malformed input");
    constexpr int kWalkers = 24;
    static_assert((uint64_t)kWalkers * F2V_WALKLEN + 64 < DrawRing::kSize, "draw window too small");
    const f2v_rng start = *g;
    DrawRing ring(start);
    Walker ws[kWalkers];
    int head = 0, live = 0;                              // ws[(head + k) % kWalkers], k < live: oldest first
    uint64_t next_i = 0, committed = 0;                  // next start vertex; draws consumed by the committed walks
    auto slot = [&](int k) -> Walker& { return ws[(head + k) % kWalkers]; };
    // first half of a step: the draw decision and the edge index; returns whether a draw was used
    auto decide = [&](Walker& x) -> bool {
        const uint64_t r0 = rowptr[x.w], dg = rowptr[x.w + 1] - r0;
        x.e = x.w;                                       // the reference indexes colids[] with the vertex id here
        x.decided++;
        const bool draw = dg > 2;
        if (draw) { x.e = r0 + ring.at(x.o + x.c) % (uint32_t)(dg - 1); x.c++; }
        else if (dg == 2) x.e = r0;
        if (x.e < nnz) __builtin_prefetch(colids + x.e, 0, 0);
        x.phase = 1;
        return draw;
    };
    // (re)start a walk at stream position o.  A start vertex's rowptr entries are sequential reads, so its first
    // decision (40 % of R-MAT's vertices are isolated) is taken at once, before anything younger is positioned
    auto start_walk = [&](Walker& x, uint64_t o) {
        x.w = x.i; x.o = o; x.c = 0; x.decided = 0; x.l = 0; x.phase = 0;
        decide(x);
    };
    // positions of the walks younger than slot k follow from their elders' predictions; a walk whose position
    // changes after it has used a draw starts over (which changes ITS prediction: keep going down the line)
    auto reposition = [&](int k) {
        for (int q = k + 1; q < live; q++) {
            const Walker& prev = slot(q - 1);
            Walker& x = slot(q);
            const uint64_t o = prev.o + prev.c + (F2V_WALKLEN - prev.decided);
            if (o == x.o) break;                         // nothing further down depends on a changed value
            if (x.c != 0) start_walk(x, o);
            else x.o = o;
        }
    };
    while (live > 0 || next_i < n) {
        while (live < kWalkers && next_i < n) {
            Walker& x = slot(live);
            x.i = next_i++;
            start_walk(x, live ? slot(live - 1).o + slot(live - 1).c + (F2V_WALKLEN - slot(live - 1).decided) : committed);
            live++;
        }
        for (int k = 0; k < live; k++) {
            Walker& x = slot(k);
            if (x.l == F2V_WALKLEN) continue;            // finished, waits for its elders to commit
            if (x.phase == 0) {
                if (!decide(x)) reposition(k);           // no draw used: everything younger moves back by one
            } else {
                const uint32_t nx = x.e < nnz ? colids[x.e] : (uint32_t)x.w;
                x.out[x.l++] = nx;
                x.w = nx;
                x.phase = 0;
                if (x.l < F2V_WALKLEN) __builtin_prefetch(rowptr + nx, 0, 0);
            }
        }
        while (live > 0 && slot(0).l == F2V_WALKLEN) {   // commit in order: the head's position is always exact
            const Walker& x = slot(0);
            memcpy(walks + x.i * F2V_WALKLEN, x.out, sizeof(x.out));
            committed = x.o + x.c;
            head = (head + 1) % kWalkers;
            live--;
        }
    }
    // the stream continues after the draws the walks consumed (the ring has drawn a few more)
    *g = start;
    if (committed) {
        uint32_t w[31];
        rng_window(start, w);
        *g = rng_jump(w, poly_pow_x(committed));
    }
    return F2V_OK;
}

// ------------------------------------------------------------------ graph IO -----------
static int csr_from_pairs(uint64_t n, std::vector<uint32_t>& src, std::vector<uint32_t>& dst, bool dedupe,
                          uint64_t* nnz_out, uint64_t** rowptr_out, uint32_t** colids_out) {
    const uint64_t m = src.size();
    uint64_t* rowptr = (uint64_t*)calloc(n + 1, sizeof(uint64_t));
    if (!rowptr) return f2v::host_fail(F2V_ERR_NOMEM, "csr_from_pairs: out of host memory");
    // counting and scattering with all host threads: the order in which a row's entries arrive does not matter,
    // every row is sorted below (equal column ids are indistinguishable)
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < (int64_t)m; k++) __atomic_fetch_add(&rowptr[src[k] + 1], 1ull, __ATOMIC_RELAXED);
    for (uint64_t i = 0; i < n; i++) rowptr[i + 1] += rowptr[i];
    uint32_t* colids = (uint32_t*)malloc(sizeof(uint32_t) * (m ? m : 1));
    if (!colids) { free(rowptr); return f2v::host_fail(F2V_ERR_NOMEM, "csr_from_pairs: out of host memory"); }
    {
        std::vector<uint64_t> cur(rowptr, rowptr + n);
#pragma omp parallel for schedule(static)
        for (int64_t k = 0; k < (int64_t)m; k++) colids[__atomic_fetch_add(&cur[src[k]], 1ull, __ATOMIC_RELAXED)] = dst[k];
    }
    std::vector<uint32_t>().swap(src);
    std::vector<uint32_t>().swap(dst);
    std::vector<uint64_t> newdeg(dedupe ? n : 0);
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        uint32_t* b = colids + rowptr[i];
        uint32_t* e = colids + rowptr[i + 1];
        std::sort(b, e);
        if (dedupe) newdeg[i] = (uint64_t)(std::unique(b, e) - b);
    }
    if (dedupe) {
        uint64_t w = 0;
        for (uint64_t i = 0; i < n; i++) {
            const uint64_t r0 = rowptr[i];
            rowptr[i] = w;
            if (w != r0) memmove(colids + w, colids + r0, sizeof(uint32_t) * newdeg[i]);
            w += newdeg[i];
        }
        rowptr[n] = w;
    }
    *nnz_out = rowptr[n];
    *rowptr_out = rowptr;
    *colids_out = colids;
    return F2V_OK;
}

extern "C" int f2v_load_mtx(const char* path, uint64_t* n_out, uint64_t* nnz_out, uint64_t** rowptr_out,
                            uint32_t** colids_out) {
    if (!path || !n_out || !nnz_out || !rowptr_out || !colids_out) return f2v::host_fail(F2V_ERR_ARG, "f2v_load_mtx: bad argument or malformed input");
    struct stat st;
    if (stat(path, &st) != 0) return f2v::host_fail(F2V_ERR_ARG, "f2v_load_mtx: cannot stat %s: %s", path, strerror(errno));
    if (!S_ISREG(st.st_mode)) return f2v::host_fail(F2V_ERR_ARG, "f2v_load_mtx: %s is not a regular file", path);
    FILE* f = fopen(path, "rb");
    if (!f) return f2v::host_fail(F2V_ERR_ARG, "f2v_load_mtx: cannot open %s: %s", path, strerror(errno));
    const long sz = (long)st.st_size;
    std::vector<char> buf;
    try { buf.resize((size_t)sz + 1); } catch (const std::bad_alloc&) { fclose(f); return f2v::host_fail(F2V_ERR_NOMEM, "f2v_load_mtx: out of host memory (%ld-byte file)", sz); }
    if (sz > 0 && fread(buf.data(), 1, (size_t)sz, f) != (size_t)sz) { fclose(f); return f2v::host_fail(F2V_ERR_ARG, "f2v_load_mtx: short read on %s", path); }
    fclose(f);
    buf[sz] = 0;
    const char* p = buf.data();
    const char* end = p + sz;
    bool symmetric = false;
    // header / comment lines start with '%' (IO.h:66-75); "symmetric" anywhere in them mirrors
    while (p < end && *p == '%') {
        const char* eol = (const char*)memchr(p, '\n', end - p);
        if (!eol) eol = end;
        std::string line(p + 1, eol);
        if (line.find("symmetric") != std::string::npos) symmetric = true;
        p = eol < end ? eol + 1 : end;
    }
    auto parse_u = [&](uint64_t& v) -> bool {
        while (p < end && (*p == ' ' || *p == '\t' || *p == '\r')) p++;
        if (p >= end || *p < '0' || *p > '9') return false;
        v = 0;
        while (p < end && *p >= '0' && *p <= '9') v = v * 10 + (uint64_t)(*p++ - '0');
        return true;
    };
    auto skip_line = [&]() {
        const char* eol = (const char*)memchr(p, '\n', end - p);
        p = eol ? eol + 1 : end;
    };
    uint64_t m = 0, ncol = 0, cnt = 0;
    if (!parse_u(m) || !parse_u(ncol) || !parse_u(cnt)) return f2v::host_fail(F2V_ERR_ARG, "f2v_load_mtx: %s: no \"rows cols entries\" size line", path);
    skip_line();
    const uint64_t n = std::max(m, ncol);
    if (n < 1 || n > 0xffffffffull) return f2v::host_fail(F2V_ERR_ARG, "f2v_load_mtx: %s: matrix size %llu out of range", path, (unsigned long long)n);
    // The entry lines are parsed by all host threads: the data region is cut at line boundaries, every
    // piece yields its (row, col) pairs in file order, and the first `cnt` entries of the file are
    // kept (the size line decides how many there are, IO.h:95-106); which thread parsed an entry does
    // not matter, the CSR builder sorts every row.
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    const uint64_t region = (uint64_t)(end - p);
    if (region < (1u << 20)) nt = 1;
    std::vector<const char*> cut(nt + 1);
    cut[0] = p;
    cut[nt] = end;
    for (int t = 1; t < nt; t++) {
        const char* q = p + region * (uint64_t)t / (uint64_t)nt;
        const char* eol = (const char*)memchr(q, '\n', end - q);
        cut[t] = eol ? eol + 1 : end;
    }
    for (int t = 1; t <= nt; t++) if (cut[t] < cut[t - 1]) cut[t] = cut[t - 1];
    std::vector<std::vector<uint32_t>> rows(nt), cols(nt);
    std::vector<int> bad(nt, 0);
    // malformed or out-of-range entries are recorded as (0xffffffff, position) and judged afterwards:
    // only entries among the first `cnt` of the file are errors
#pragma omp parallel for schedule(static, 1) num_threads(nt)
    for (int t = 0; t < nt; t++) {
        const char* q = cut[t];
        const char* qe = cut[t + 1];
        auto num = [&](uint64_t& v) -> bool {
            while (q < qe && (*q == ' ' || *q == '\t' || *q == '\r')) q++;
            if (q >= qe || *q < '0' || *q > '9') return false;
            v = 0;
            while (q < qe && *q >= '0' && *q <= '9') v = v * 10 + (uint64_t)(*q++ - '0');
            return true;
        };
        std::vector<uint32_t>& R = rows[t];
        std::vector<uint32_t>& Cc = cols[t];
        R.reserve((size_t)((qe - q) / 8));
        Cc.reserve((size_t)((qe - q) / 8));
        while (q < qe) {
            while (q < qe && (*q == '\n' || *q == '\r')) q++;
            if (q >= qe) break;
            uint64_t r = 0, c = 0;
            const bool ok = num(r) && num(c) && r >= 1 && c >= 1 && r <= n && c <= n;
            const char* eol = (const char*)memchr(q, '\n', qe - q);   // the value column is never used
            q = eol ? eol + 1 : qe;
            R.push_back(ok ? (uint32_t)(r - 1) : 0xffffffffu);
            Cc.push_back(ok ? (uint32_t)(c - 1) : 0u);
        }
    }
    uint64_t have = 0;
    for (int t = 0; t < nt; t++) have += rows[t].size();
    if (have < cnt) return f2v::host_fail(F2V_ERR_ARG, "f2v_load_mtx: %s: %llu entries, the size line says %llu", path, (unsigned long long)have, (unsigned long long)cnt);
    std::vector<uint32_t> src, dst;
    src.reserve(symmetric ? 2 * cnt : cnt);
    dst.reserve(symmetric ? 2 * cnt : cnt);
    uint64_t taken = 0;
    for (int t = 0; t < nt && taken < cnt; t++) {
        const uint64_t m_t = std::min<uint64_t>(rows[t].size(), cnt - taken);
        for (uint64_t k = 0; k < m_t; k++) {
            const uint32_t r = rows[t][k], c = cols[t][k];
            if (r == 0xffffffffu) return f2v::host_fail(F2V_ERR_ARG, "f2v_load_mtx: %s: entry %llu is malformed or out of range", path, (unsigned long long)(taken + k + 1));
            if (symmetric) {
                if (r == c) continue;                          // self-loops dropped (IO.h:130-134)
                src.push_back(r); dst.push_back(c);
                src.push_back(c); dst.push_back(r);            // mirrored (IO.h:122-129)
            } else {
                src.push_back(r); dst.push_back(c);
            }
        }
        taken += m_t;
        std::vector<uint32_t>().swap(rows[t]);
        std::vector<uint32_t>().swap(cols[t]);
    }
    *n_out = n;
    return csr_from_pairs(n, src, dst, /*dedupe=*/false, nnz_out, rowptr_out, colids_out);
}

extern "C" void f2v_free(void* p) { free(p); }

// "%.6g" of a float without printf: what `ostream << float` emits (algorithms.h:126-134).  The six
// significant digits come from one double multiplication by a power of ten; when the scaled value
// is within 1e-6 of a rounding tie (where the product's error could flip the last digit) or the
// value is not finite, snprintf decides.  Returns the number of characters written (no terminator).
static int format_g6(char* out, float v) {
    static const double kPow10[] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11, 1e12, 1e13,
                                    1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    if (v == 0.0f) {
        int k = 0;
        if (std::signbit(v)) out[k++] = '-';
        out[k++] = '0';
        return k;
    }
    const double a = std::fabs((double)v);
    if (!(a >= 1e-17 && a < 1e17)) return snprintf(out, 32, "%.6g", (double)v);    // rare magnitudes, inf, nan
    int e10 = (int)std::floor(std::log10(a));
    double scaled = e10 <= 5 ? a * kPow10[5 - e10] : a / kPow10[e10 - 5];
    if (scaled < 1e5) { e10--; scaled = e10 <= 5 ? a * kPow10[5 - e10] : a / kPow10[e10 - 5]; }
    else if (scaled >= 1e6) { e10++; scaled = e10 <= 5 ? a * kPow10[5 - e10] : a / kPow10[e10 - 5]; }
    const double fl = std::floor(scaled), frac = scaled - fl;
    if (std::fabs(frac - 0.5) < 1e-6 || scaled < 1e5 || scaled >= 1e6) return snprintf(out, 32, "%.6g", (double)v);
    uint32_t digits = (uint32_t)fl + (frac > 0.5 ? 1u : 0u);
    if (digits >= 1000000u) { digits = 100000u; e10++; }
    char d[6];
    for (int k = 5; k >= 0; k--) { d[k] = (char)('0' + digits % 10); digits /= 10; }
    int nd = 6;
    while (nd > 1 && d[nd - 1] == '0') nd--;                   // %g drops trailing zeros
    int k = 0;
    if (v < 0) out[k++] = '-';
    if (e10 < -4 || e10 >= 6) {                                 // d.ddddde[+-]XX
        out[k++] = d[0];
        if (nd > 1) { out[k++] = '.'; for (int i = 1; i < nd; i++) out[k++] = d[i]; }
        out[k++] = 'e';
        int e = e10;
        if (e < 0) { out[k++] = '-'; e = -e; } else out[k++] = '+';
        if (e >= 100) { out[k++] = (char)('0' + e / 100); e %= 100; }
        out[k++] = (char)('0' + e / 10);
        out[k++] = (char)('0' + e % 10);
    } else if (e10 >= 0) {                                      // ddd.ddd
        const int ip = e10 + 1;                                 // digits before the point
        for (int i = 0; i < ip; i++) out[k++] = i < nd ? d[i] : '0';
        if (nd > ip) { out[k++] = '.'; for (int i = ip; i < nd; i++) out[k++] = d[i]; }
    } else {                                                    // 0.000ddd
        out[k++] = '0'; out[k++] = '.';
        for (int i = 0; i < -e10 - 1; i++) out[k++] = '0';
        for (int i = 0; i < nd; i++) out[k++] = d[i];
    }
    return k;
}

// exposed for the test that compares it with printf on millions of values
extern "C" int f2v_format_g6(float v, char* out32) { int k = format_g6(out32, v); out32[k] = 0; return k; }

extern "C" int f2v_write_embd(const char* path, const float* X, uint64_t n, uint32_t dim) {
    if (!path || !X) return f2v::host_fail(F2V_ERR_ARG, "f2v_write_embd: bad argument or malformed input");
    FILE* f = fopen(path, "wb");
    if (!f) return f2v::host_fail(F2V_ERR_ARG, "f2v_write_embd: bad argument or malformed input");
    fprintf(f, "%llu %u\n", (unsigned long long)n, dim);
    // rows are formatted in parallel blocks, written in order; "%.6g" is what ostream << float emits
    const uint64_t block = 4096;
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    // worst case per value: "-1.23457e-38 " = 13 characters; per row: a 20-digit id, a space, a newline
    const size_t row_cap = 24 + (size_t)dim * 16;
    std::vector<std::vector<char>> out(nt);
    std::vector<size_t> used(nt, 0);
    for (int t = 0; t < nt; t++) out[t].resize(row_cap * block);
    for (uint64_t base = 0; base < n; base += block * nt) {
#pragma omp parallel for schedule(static, 1) num_threads(nt)
        for (int t = 0; t < nt; t++) {
            char* s = out[t].data();
            size_t k = 0;
            uint64_t lo = base + (uint64_t)t * block, hi = std::min(n, lo + block);
            for (uint64_t i = lo; i < hi; i++) {
                k += (size_t)snprintf(s + k, 24, "%llu ", (unsigned long long)(i + 1));
                const float* row = X + i * dim;
                for (uint32_t d = 0; d < dim; d++) {
                    k += (size_t)format_g6(s + k, row[d]);
                    s[k++] = ' ';
                }
                s[k++] = '\n';
            }
            used[t] = lo < hi ? k : 0;
        }
        for (int t = 0; t < nt; t++)
            if (used[t]) fwrite(out[t].data(), 1, used[t], f);
    }
    int rc = ferror(f) ? F2V_ERR_ARG : F2V_OK;
    fclose(f);
    return rc;
}

// Binary CSR cache: "F2VCSR01", u64 n, u64 nnz, rowptr u64[n+1], colids u32[nnz].  The text loader
// parses ~10 M entries/s; a scale-24 graph (0.5 G entries) is minutes as text and seconds like this.
extern "C" int f2v_write_csr(const char* path, uint64_t n, uint64_t nnz, const uint64_t* rowptr, const uint32_t* colids) {
    if (!path || !rowptr || (nnz && !colids) || rowptr[n] != nnz) return f2v::host_fail(F2V_ERR_ARG, "f2v_write_csr: bad argument or malformed input");
    FILE* f = fopen(path, "wb");
    if (!f) return f2v::host_fail(F2V_ERR_ARG, "f2v_write_csr: bad argument or malformed input");
    const char magic[8] = {'F', '2', 'V', 'C', 'S', 'R', '0', '1'};
    bool ok = fwrite(magic, 1, 8, f) == 8 && fwrite(&n, 8, 1, f) == 1 && fwrite(&nnz, 8, 1, f) == 1 &&
              fwrite(rowptr, 8, n + 1, f) == n + 1 && (nnz == 0 || fwrite(colids, 4, nnz, f) == nnz);
    ok = fclose(f) == 0 && ok;
    return ok ? F2V_OK : F2V_ERR_ARG;
}

extern "C" int f2v_load_csr(const char* path, uint64_t* n_out, uint64_t* nnz_out, uint64_t** rowptr_out, uint32_t** colids_out) {
    if (!path || !n_out || !nnz_out || !rowptr_out || !colids_out) return f2v::host_fail(F2V_ERR_ARG, "f2v_load_csr: bad argument or malformed input");
    FILE* f = fopen(path, "rb");
    if (!f) return f2v::host_fail(F2V_ERR_ARG, "f2v_load_csr: bad argument or malformed input");
    char magic[8];
    uint64_t n = 0, nnz = 0;
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "F2VCSR01", 8) != 0 || fread(&n, 8, 1, f) != 1 ||
        fread(&nnz, 8, 1, f) != 1 || n < 1 || n > 0xffffffffull) { fclose(f); return f2v::host_fail(F2V_ERR_ARG, "f2v_load_csr: bad argument or malformed input"); }
    uint64_t* rp = (uint64_t*)malloc(sizeof(uint64_t) * (n + 1));
    uint32_t* ci = (uint32_t*)malloc(sizeof(uint32_t) * (nnz ? nnz : 1));
    bool ok = rp && ci && fread(rp, 8, n + 1, f) == n + 1 && (nnz == 0 || fread(ci, 4, nnz, f) == nnz);
    fclose(f);
    ok = ok && rp[0] == 0 && rp[n] == nnz;
    for (uint64_t i = 0; ok && i < n; i++) ok = rp[i + 1] >= rp[i];
    for (uint64_t k = 0; ok && k < nnz; k++) ok = ci[k] < n;
    if (!ok) { free(rp); free(ci); return ok ? F2V_OK : (rp && ci ? F2V_ERR_ARG : F2V_ERR_NOMEM); }
    *n_out = n; *nnz_out = nnz; *rowptr_out = rp; *colids_out = ci;
    return F2V_OK;
}

extern "C" int f2v_write_mtx(const char* path, uint64_t n, const uint64_t* rowptr, const uint32_t* colids) {
    if (!path || !rowptr) return f2v::host_fail(F2V_ERR_ARG, "f2v_write_mtx: bad argument or malformed input");
    FILE* f = fopen(path, "wb");
    if (!f) return f2v::host_fail(F2V_ERR_ARG, "f2v_write_mtx: bad argument or malformed input");
    uint64_t cnt = 0;
    for (uint64_t i = 0; i < n; i++)
        for (uint64_t e = rowptr[i]; e < rowptr[i + 1]; e++)
            if (colids[e] < i) cnt++;
    fprintf(f, "%%%%MatrixMarket matrix coordinate pattern symmetric\n%llu %llu %llu\n",
            (unsigned long long)n, (unsigned long long)n, (unsigned long long)cnt);
    std::vector<char> buf;
    buf.reserve(1 << 22);
    char tmp[48];
    for (uint64_t i = 0; i < n; i++) {
        for (uint64_t e = rowptr[i]; e < rowptr[i + 1]; e++) {
            if (colids[e] >= i) continue;
            int k = snprintf(tmp, sizeof(tmp), "%llu %u\n", (unsigned long long)(i + 1), colids[e] + 1);
            buf.insert(buf.end(), tmp, tmp + k);
        }
        if (buf.size() > (1u << 22) - 64) { fwrite(buf.data(), 1, buf.size(), f); buf.clear(); }
    }
    if (!buf.empty()) fwrite(buf.data(), 1, buf.size(), f);
    int rc = ferror(f) ? F2V_ERR_ARG : F2V_OK;
    fclose(f);
    return rc;
}

// ------------------------------------------------------------------ R-MAT --------------
static inline uint64_t splitmix(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// Edge k of the R-MAT stream: one quadrant choice per level from a generator keyed by
// (seed, k), so the edge list is independent of the thread count and is never stored.
static inline void rmat_edge(int scale, uint64_t seed, uint64_t k, uint32_t& u, uint32_t& v) {
    uint64_t st = seed * 0xD1B54A32D192ED03ULL + k * 0x9E3779B97F4A7C15ULL + 0x632BE59BD9B4E019ULL;
    uint32_t uu = 0, vv = 0;
    uint64_t bits = 0;
    int have = 0;
    for (int l = 0; l < scale; l++) {
        if (have == 0) { bits = splitmix(st); have = 2; }
        const uint32_t r = (uint32_t)bits;
        bits >>= 32;
        have--;
        // thresholds on 2^32: a = .57, a+b = .76, a+b+c = .95
        uint32_t q = r < 2448131358u ? 0 : (r < 3264175145u ? 1 : (r < 4080218931u ? 2 : 3));
        uu = (uu << 1) | (q >> 1);
        vv = (vv << 1) | (q & 1);
    }
    u = uu;
    v = vv;
}

extern "C" int f2v_rmat_csr(int scale, int edge_factor, uint64_t seed, uint64_t* n_out, uint64_t* nnz_out,
                            uint64_t** rowptr_out, uint32_t** colids_out) {
    if (scale < 2 || scale > 31 || edge_factor < 1 || !n_out || !nnz_out || !rowptr_out || !colids_out)
        return f2v::host_fail(F2V_ERR_ARG, "f2v_rmat_csr: bad argument or malformed input");
    const uint64_t n = 1ull << scale, m = (uint64_t)edge_factor * n;
    std::vector<uint32_t> degc(n, 0);
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < (int64_t)m; k++) {
        uint32_t u, v;
        rmat_edge(scale, seed, (uint64_t)k, u, v);
        if (u == v) continue;
        __atomic_fetch_add(&degc[u], 1u, __ATOMIC_RELAXED);
        __atomic_fetch_add(&degc[v], 1u, __ATOMIC_RELAXED);
    }
    uint64_t* rowptr = (uint64_t*)malloc(sizeof(uint64_t) * (n + 1));
    if (!rowptr) return f2v::host_fail(F2V_ERR_NOMEM, "f2v_rmat_csr: out of host memory");
    rowptr[0] = 0;
    for (uint64_t i = 0; i < n; i++) rowptr[i + 1] = rowptr[i] + degc[i];
    const uint64_t tot = rowptr[n];
    uint32_t* colids = (uint32_t*)malloc(sizeof(uint32_t) * (tot ? tot : 1));
    if (!colids) { free(rowptr); return f2v::host_fail(F2V_ERR_NOMEM, "f2v_rmat_csr: out of host memory"); }
    std::vector<uint64_t> cur(rowptr, rowptr + n);
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < (int64_t)m; k++) {
        uint32_t u, v;
        rmat_edge(scale, seed, (uint64_t)k, u, v);
        if (u == v) continue;
        colids[__atomic_fetch_add(&cur[u], 1ull, __ATOMIC_RELAXED)] = v;
        colids[__atomic_fetch_add(&cur[v], 1ull, __ATOMIC_RELAXED)] = u;
    }
    std::vector<uint64_t>().swap(cur);
    std::vector<uint64_t> newdeg(n);
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        uint32_t* b = colids + rowptr[i];
        uint32_t* e = colids + rowptr[i + 1];
        std::sort(b, e);
        newdeg[i] = (uint64_t)(std::unique(b, e) - b);
    }
    uint64_t w = 0;
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t r0 = rowptr[i];
        rowptr[i] = w;
        if (w != r0) memmove(colids + w, colids + r0, sizeof(uint32_t) * newdeg[i]);
        w += newdeg[i];
    }
    rowptr[n] = w;
    *n_out = n;
    *nnz_out = w;
    *rowptr_out = rowptr;
    *colids_out = colids;
    return F2V_OK;
}

// ------------------------------------------------------------------ plan ---------------
extern "C" int f2v_plan_build(const uint64_t* rowptr, uint64_t first_row, uint64_t nrows, uint32_t batch,
                              uint32_t chunk, uint32_t par, int walk, int rank, int world, int assign, uint64_t* nb,
                              uint64_t** item_ptr, uint32_t** n_hub, void** items, void** hub) {
    if (!rowptr || !nb || !item_ptr || !n_hub || !items || !hub || batch == 0 || world < 1 || rank < 0 ||
        rank >= world || chunk == 0 || assign < 0 || assign > 15)
        return f2v::host_fail(F2V_ERR_ARG, "f2v_plan_build: bad argument or malformed input");
    f2v::HostPlan hp;
    f2v::build_host_plan(rowptr, first_row, nrows, batch, chunk, par, walk != 0, rank, world, assign, hp);
    const size_t total = hp.item_ptr[hp.nb];
    *nb = hp.nb;
    *item_ptr = (uint64_t*)malloc(sizeof(uint64_t) * (hp.nb + 1));
    *n_hub = (uint32_t*)malloc(sizeof(uint32_t) * (hp.nb ? hp.nb : 1));
    *items = malloc(sizeof(f2v::Item) * (total ? total : 1));
    *hub = malloc(sizeof(f2v::HubInfo) * (total ? total : 1));
    if (!*item_ptr || !*n_hub || !*items || !*hub) return f2v::host_fail(F2V_ERR_NOMEM, "f2v_plan_build: out of host memory");
    memcpy(*item_ptr, hp.item_ptr.data(), sizeof(uint64_t) * (hp.nb + 1));
    if (hp.nb) memcpy(*n_hub, hp.n_hub.data(), sizeof(uint32_t) * hp.nb);
    if (total) {
        memcpy(*items, hp.items.data(), sizeof(f2v::Item) * total);
        memcpy(*hub, hp.hub.data(), sizeof(f2v::HubInfo) * total);
    }
    return F2V_OK;
}

// ------------------------------------------------------------------ driver -------------
extern "C" int f2v_train(const f2v_train_args* a, float* X_out, double* seconds) {
    if (!a || !X_out || !a->rowptr) return f2v::host_fail(F2V_ERR_ARG, "f2v_train: bad argument or malformed input");
    if (a->option != F2V_TDIST && a->option != F2V_SIGMOID && a->option != F2V_WALK) return f2v::host_fail(F2V_ERR_ARG, "f2v_train: bad argument or malformed input");
    const int model = a->option;
    const int bs = model == F2V_WALK ? 0 : (a->bs ? 1 : 0);   // -bs ignored by option 7
    f2v_engine* e = nullptr;
    int rc = f2v_create(&e, a->device, a->n, a->nnz, a->rowptr, a->colids, a->dim);
    if (rc) return rc;
    struct Guard { f2v_engine* e; f2v_rng* g; void* p0; void* p1; void* w;
                   ~Guard() { f2v_host_free(p0); f2v_host_free(p1); f2v_host_free(w); f2v_rng_destroy(g); f2v_destroy(e); } } gd{e, nullptr, nullptr, nullptr, nullptr};
    if (a->epoch_mode) { rc = f2v_set_epoch_mode(e, a->epoch_mode); if (rc) return rc; }
    auto t0 = std::chrono::steady_clock::now();      // algorithms.cpp:557 -- the timer starts before init
    f2v_rng* g = gd.g = f2v_rng_create(a->seed);
    if (!g) return f2v::host_fail(F2V_ERR_NOMEM, "f2v_train: out of host memory");
    f2v_init_embeddings(g, model, a->n, a->dim, X_out);
    rc = f2v_set_embeddings(e, X_out);
    if (rc) return rc;
    if (model != F2V_TDIST) {
        float lut[F2V_LUT_SIZE];
        f2v_build_lut(lut);
        rc = f2v_set_lut(e, lut, F2V_LUT_SIZE);
        if (rc) return rc;
    }
    const uint64_t slen = f2v_neg_stream_len(model, a->n, a->batch, a->nsamples, bs);
    uint32_t* negbuf[2] = {nullptr, nullptr};
    rc = f2v_host_alloc(&gd.p0, sizeof(uint32_t) * (slen ? slen : 1)); if (rc) return rc;
    rc = f2v_host_alloc(&gd.p1, sizeof(uint32_t) * (slen ? slen : 1)); if (rc) return rc;
    negbuf[0] = (uint32_t*)gd.p0; negbuf[1] = (uint32_t*)gd.p1;
    uint32_t* walks = nullptr;
    const bool host_walks = model == F2V_WALK && a->walk_sampler == 0;
    if (host_walks) {
        rc = f2v_host_alloc(&gd.w, sizeof(uint32_t) * a->n * F2V_WALKLEN); if (rc) return rc;
        walks = (uint32_t*)gd.w;
    }
    // Per epoch the reference draws (walks, then) negatives from the serial stream; the draws
    // for epoch it+1 are produced on the host while the GPU runs epoch it.
    auto draw = [&](uint32_t it) -> int {
        (void)it;
        if (host_walks) { int r = f2v_draw_walks(g, a->n, a->nnz, a->rowptr, a->colids, walks); if (r) return r; }
        return f2v_draw_epoch_negatives(g, model, a->n, a->batch, a->nsamples, bs, negbuf[it & 1]);
    };
    if (a->iterations > 0) { rc = draw(0); if (rc) return rc; }
    for (uint32_t it = 0; it < a->iterations; it++) {
        if (model == F2V_WALK) {
            if (host_walks) rc = f2v_set_walks(e, walks);
            else rc = f2v_sample_walks(e, a->seed, it);
            if (rc) return rc;
        }
        rc = f2v_set_negatives(e, negbuf[it & 1], slen);
        if (rc) return rc;
        if (host_walks) { rc = f2v_sync(e); if (rc) return rc; }   // walks buffer is reused by the next draw
        rc = f2v_run_epoch(e, model, a->batch, a->nsamples, bs, a->lr, a->chunk);
        if (rc) return rc;
        if (it + 1 < a->iterations) { rc = draw(it + 1); if (rc) return rc; }
        rc = f2v_sync(e);
        if (rc) return rc;
    }
    rc = f2v_get_embeddings(e, X_out);
    if (rc) return rc;
    auto t1 = std::chrono::steady_clock::now();      // algorithms.cpp:647 -- before the file write
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    return F2V_OK;
}

// ------------------------------------------------------------------ multi-GPU driver ----
// The same run on `gpus` devices of this node from ONE process: one host thread and one engine per
// device, replicated tables, the fused peer-store exchange (engines of one process reach each
// other's tables by plain peer access).  Thread 0 owns the libc-compatible stream and draws the
// walks / negatives of every epoch into host buffers all threads upload from, so the result is the
// single-GPU result bit for bit (for equal `chunk`).
namespace {
// Barrier whose wait() also tells every thread, identically, whether any thread had failed when the
// LAST one arrived.  Threads branch only on that value, so they all run the same number of barriers
// (a thread that read a shared flag on its own could leave the loop one barrier earlier than a peer
// and pair its post-loop wait with the peer's in-loop wait: a hang instead of an error).
struct FailBarrier {
    std::mutex m;
    std::condition_variable cv;
    int count = 0, gen = 0, parties;
    bool stop[2] = {false, false};       // decision of generation g lives in stop[g & 1]
    std::atomic<int>& failed;
    FailBarrier(int n, std::atomic<int>& f) : parties(n), failed(f) {}
    bool wait() {                        // true = somebody failed: stop together
        std::unique_lock<std::mutex> lk(m);
        const int g = gen;
        if (++count == parties) { stop[g & 1] = failed.load() != 0; count = 0; gen++; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
        return stop[g & 1];
    }
};
}  // namespace

// test hook: make the given rank fail at the given epoch's upload (F2V_TEST_FAIL_RANK / F2V_TEST_FAIL_EPOCH)
static int injected_failure(int rank, uint32_t epoch) {
    const char* fr = getenv("F2V_TEST_FAIL_RANK");
    if (!fr || atoi(fr) != rank) return F2V_OK;
    const char* fe = getenv("F2V_TEST_FAIL_EPOCH");
    return (fe ? (uint32_t)atoi(fe) : 0u) == epoch ? F2V_ERR_STATE : F2V_OK;
}

extern "C" int f2v_train_gpus(const f2v_train_args* a, int gpus, float* X_out, double* seconds) {
    if (gpus <= 1) return f2v_train(a, X_out, seconds);
    if (!a || !X_out || !a->rowptr) return f2v::host_fail(F2V_ERR_ARG, "f2v_train_gpus: bad argument or malformed input");
    if (a->option != F2V_TDIST && a->option != F2V_SIGMOID && a->option != F2V_WALK) return f2v::host_fail(F2V_ERR_ARG, "f2v_train_gpus: bad argument or malformed input");
    if (gpus > 8 || gpus > f2v_device_count()) return f2v::host_fail(F2V_ERR_ARG, "f2v_train_gpus: bad argument or malformed input");
    const int model = a->option;
    const int bs = model == F2V_WALK ? 0 : (a->bs ? 1 : 0);
    const int G = gpus;
    // the timed span is f2v_train's (and the reference's, algorithms.cpp:557-647): it starts after the engines
    // exist -- the reference's constructor has copied the graph before its timer starts -- and before the init draws
    std::chrono::steady_clock::time_point t0;
    f2v_rng* g = f2v_rng_create(a->seed);
    if (!g) return f2v::host_fail(F2V_ERR_NOMEM, "f2v_train_gpus: out of host memory");
    float lut[F2V_LUT_SIZE];
    f2v_build_lut(lut);
    const uint64_t slen = f2v_neg_stream_len(model, a->n, a->batch, a->nsamples, bs);
    std::vector<uint32_t> neg(slen ? slen : 1);
    const bool host_walks = model == F2V_WALK && a->walk_sampler == 0;
    std::vector<uint32_t> walks(host_walks ? a->n * (uint64_t)F2V_WALKLEN : 1);
    std::vector<char> blobs((size_t)G * F2V_PEER_BLOB);
    std::vector<int> status(G, F2V_OK);
    std::vector<std::string> errs(G);
    std::atomic<int> failed{0};
    FailBarrier bar(G, failed);
    auto worker = [&](int r) {
        f2v_engine* e = nullptr;
        int rc = F2V_OK;
        auto step = [&](int code) {          // record the first error of this thread; every thread learns of it
            if (code && !rc) { rc = code; errs[r] = f2v_last_error(); failed.store(1); }   // at the next barrier
        };
        step(f2v_create(&e, a->device + r, a->n, a->nnz, a->rowptr, a->colids, a->dim));
        if (!rc && a->epoch_mode) step(f2v_set_epoch_mode(e, a->epoch_mode));
        // one host thread per engine, so the engines of this process COULD run the (blocking) NVLink multicast
        // hand-off among themselves -- but that set-up has only ever run between processes (two in-process engines
        // exchange with one store per peer by the N = 2 rule), so it is opt-in here (F2V_INPROC_MULTICAST=1) and the
        // default for engines of one process is plain peer access with one store per peer
        if (!rc) {
            const char* mc = getenv("F2V_INPROC_MULTICAST");
            if (mc && atoi(mc) != 0) step(f2v_set_option(e, "multicast_in_process", 1));
        }
        if (!rc) step(f2v_comm_peer_export(e, blobs.data() + (size_t)r * F2V_PEER_BLOB));
        bool stop = bar.wait();
        if (!stop) step(f2v_comm_peer_init(e, blobs.data(), r, G));
        stop = bar.wait();
        if (r == 0) {                        // engines and exchange are up: the timed span starts with the init draws
            t0 = std::chrono::steady_clock::now();
            if (!stop) step(f2v_init_embeddings(g, model, a->n, a->dim, X_out));
        }
        stop = bar.wait();
        if (!stop) {
            if (model != F2V_TDIST) step(f2v_set_lut(e, lut, F2V_LUT_SIZE));
            step(f2v_set_embeddings(e, X_out));
        }
        stop = bar.wait();
        // every thread takes the same decisions (the barrier's return value), so every thread runs
        // the same number of barriers whatever fails where
        for (uint32_t it = 0; it < a->iterations && !stop; it++) {
            if (r == 0) {                    // the serial draws of this epoch (reference order: walks, then negatives)
                if (host_walks) step(f2v_draw_walks(g, a->n, a->nnz, a->rowptr, a->colids, walks.data()));
                step(f2v_draw_epoch_negatives(g, model, a->n, a->batch, a->nsamples, bs, neg.data()));
            }
            stop = bar.wait();
            if (stop) break;
            if (injected_failure(r, it)) step(f2v::host_fail(F2V_ERR_STATE, "injected failure on rank %d, epoch %u (test hook)", r, it));
            if (!rc && model == F2V_WALK) step(host_walks ? f2v_set_walks(e, walks.data()) : f2v_sample_walks(e, a->seed, it));
            if (!rc) step(f2v_set_negatives(e, neg.data(), slen));
            if (!rc) step(f2v_sync(e));      // the host buffers are redrawn by thread 0 after the next barrier
            stop = bar.wait();
            if (stop) break;                 // nobody has launched this epoch yet: safe to stop together
            step(f2v_run_epoch(e, model, a->batch, a->nsamples, bs, a->lr, a->chunk));
        }
        if (e && !rc) step(f2v_sync(e));
        stop = bar.wait();
        if (r == 0 && !stop) step(f2v_get_embeddings(e, X_out));
        bar.wait();                          // nobody unmaps a table a peer may still be storing into
        if (e) f2v_destroy(e);
        status[r] = rc;
    };
    std::vector<std::thread> th;
    for (int r = 1; r < G; r++) th.emplace_back(worker, r);
    worker(0);
    for (auto& t : th) t.join();
    f2v_rng_destroy(g);
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    for (int r = 0; r < G; r++)
        if (status[r]) { fprintf(stderr, "f2v_train_gpus: device %d: %s\n", a->device + r, errs[r].c_str()); return status[r]; }
    return F2V_OK;
}

// ------------------------------------------------------------------ C++ mirror ---------
namespace f2v {

bool SetInputMatricesAsCSR(Csr& A, const std::string& path, std::string* err) {
    // ".f2vcsr": the binary CSR cache (no reference counterpart); anything else is MatrixMarket text
    const bool binary = path.size() > 7 && path.compare(path.size() - 7, 7, ".f2vcsr") == 0;
    std::cout << (binary ? "Reading input matrices in binary (CSR cache)... " : "Reading input matrices in text (ascii)... ") << std::endl;
    std::cout << "Input File Directory:" << path << std::endl;
    uint64_t n = 0, nnz = 0;
    uint64_t* rp = nullptr;
    uint32_t* ci = nullptr;
    int rc = binary ? f2v_load_csr(path.c_str(), &n, &nnz, &rp, &ci) : f2v_load_mtx(path.c_str(), &n, &nnz, &rp, &ci);
    if (rc) {
        if (err) *err = std::string(binary ? "cannot read CSR cache file " : "cannot read MatrixMarket file ") + path;
        return false;
    }
    A.rows = n;
    A.nnz = nnz;
    A.rowptr.assign(rp, rp + n + 1);
    A.colids.assign(ci, ci + nnz);
    f2v_free(rp);
    f2v_free(ci);
    printf("Input Matrix: Rows = %llu, Columns= %llu, nnz = %llu\n", (unsigned long long)n,
           (unsigned long long)n, (unsigned long long)nnz);
    return true;
}

algorithms::algorithms(const Csr& A_csr, std::string input, std::string outputd, uint32_t dim, float gm, uint32_t)
    : graph(A_csr), GAMMA(gm), DIM(dim), filename(std::move(input)), outputdir(std::move(outputd)) {
    nCoordinates.assign((size_t)graph.rows * dim, 0.f);
}

std::vector<float> algorithms::run(int option, int bs, uint32_t iters, uint32_t batch, uint32_t ns, float lr,
                                   const char* banner, const std::string& tag) {
    f2v_train_args a{};
    a.n = graph.rows; a.nnz = graph.nnz; a.rowptr = graph.rowptr.data(); a.colids = graph.colids.data();
    a.dim = DIM; a.option = option; a.bs = bs; a.iterations = iters; a.batch = batch; a.nsamples = ns;
    a.lr = lr; a.seed = seed; a.device = device; a.walk_sampler = walk_sampler; a.epoch_mode = epoch_mode;
    a.chunk = chunk;
    double sec = 0;
    int rc = gpus > 1 ? f2v_train_gpus(&a, gpus, nCoordinates.data(), &sec) : f2v_train(&a, nCoordinates.data(), &sec);
    if (rc != F2V_OK) {
        // the reference's error convention: message + exit(1) (Test/Force2Vec.cpp:119,186)
        fprintf(stderr, "Force2Vec GPU engine error %d: %s\n", rc, f2v_last_error());
        exit(1);
    }
    std::cout << banner << sec << " seconds" << std::endl;
    writeToFile(tag + std::to_string(batch) + "D" + std::to_string(DIM) + "IT" + std::to_string(iters) + "NS" + std::to_string(ns));
    return std::vector<float>{(float)sec};
}

std::vector<float> algorithms::AlgoForce2VecNS(uint32_t IT, uint32_t, uint32_t B, uint32_t ns, float lr) {
    return run(F2V_TDIST, 0, IT, B, ns, lr, "Force2Vec Parallel Wall time required:", "F2VNS");
}
std::vector<float> algorithms::AlgoForce2VecNSBS(uint32_t IT, uint32_t, uint32_t B, uint32_t ns, float lr) {
    return run(F2V_TDIST, 1, IT, B, ns, lr, "Force2Vec Parallel Wall time required (with BS negative samples):", "F2VNS");
}
std::vector<float> algorithms::AlgoForce2VecNSRW(uint32_t IT, uint32_t, uint32_t B, uint32_t ns, float lr) {
    return run(F2V_SIGMOID, 0, IT, B, ns, lr, "Force2Vec Parallel Wall time required:", "F2VWNS");
}
std::vector<float> algorithms::AlgoForce2VecNSRWBS(uint32_t IT, uint32_t, uint32_t B, uint32_t ns, float lr) {
    return run(F2V_SIGMOID, 1, IT, B, ns, lr, "Force2Vec Parallel Wall time required (with BS negative samples):", "F2VWNS");
}
std::vector<float> algorithms::AlgoForce2VecNSRWEFF(uint32_t IT, uint32_t, uint32_t B, uint32_t ns, float lr) {
    return run(F2V_WALK, 0, IT, B, ns, lr, "Force2VecWNSEFF Parallel Wall time required:", "F2VWNSF");
}

void algorithms::writeToFile(std::string f) {
    size_t pos = filename.find_last_of('/');
    std::string lasttok = pos == std::string::npos ? filename : filename.substr(pos + 1);
    filename = outputdir + lasttok + f + ".embd";
    std::cout << "Creating output file in following directory:" << filename << std::endl;
    f2v_write_embd(filename.c_str(), nCoordinates.data(), graph.rows, DIM);
}

}  // namespace f2v
