// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A thin extern "C" shim, written for this repo, around the UNMODIFIED reference
// class `algorithms` (/root/reference/sample/algorithms.h:51-137).  It is compiled
// together with the reference's own sample/algorithms.cpp *where it lies* under
// /root/reference (see oracle/Makefile) into oracle/_ref/libf2vref*.so, so that
// tests can drive the reference's option 5/6/7 code in memory (full fp32
// precision, no .mtx/.embd round trip).  What the reference CLI does around the
// call (Test/Force2Vec.cpp:121-150) is mirrored here: build the CSR, construct
// `algorithms`, srand(1), dispatch on option/bs.
//
// No reference source is copied: this file only *includes* the reference header.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "algorithms.h"   // -I/root/reference/sample  (note: it #defines `t`)

extern VALUETYPE* sm_table;   // sample/algorithms.cpp:755
void init_SM_TABLE();          // sample/algorithms.cpp:757

extern "C" {

// Returns the reference's own wall-seconds (algorithms.cpp:557,647 span: init + all
// epochs) or a negative value on bad arguments.  X_out receives nCoordinates
// (n*dim floats) after the run.  outdir must end in '/' (algorithms.h:119-122);
// pass a non-existent directory to suppress the .embd write.
double f2vref_run(uint32_t n, uint32_t nnz, const uint32_t* rowptr, const uint32_t* colids,
                  uint32_t dim, int option, int bs, uint32_t iterations, uint32_t threads,
                  uint32_t batch, uint32_t ns, float lr, const char* outdir, float* X_out)
{
    if (n == 0 || rowptr == nullptr || (nnz > 0 && colids == nullptr)) return -1.0;
    CSR<INDEXTYPE, VALUETYPE> A(nnz, n, n);
    memcpy(A.rowptr, rowptr, sizeof(uint32_t) * (size_t)(n + 1));
    if (nnz > 0) memcpy(A.colids, colids, sizeof(uint32_t) * (size_t)nnz);
    std::vector<VALUETYPE> res;
    {
        algorithms algo(A, std::string("graph"), std::string(outdir ? outdir : "/nonexistent_f2vref/"),
                        dim, 1.0f, batch);
        srand(1);                              // Test/Force2Vec.cpp:126
        A.make_empty();                        // Test/Force2Vec.cpp:127
        switch (option) {
        case 5:
            res = bs ? algo.AlgoForce2VecNSBS(iterations, threads, batch, ns, lr)
                     : algo.AlgoForce2VecNS(iterations, threads, batch, ns, lr);
            break;
        case 6:
            res = bs ? algo.AlgoForce2VecNSRWBS(iterations, threads, batch, ns, lr)
                     : algo.AlgoForce2VecNSRW(iterations, threads, batch, ns, lr);
            break;
        case 7:
            res = algo.AlgoForce2VecNSRWEFF(iterations, threads, batch, ns, lr);
            break;
#ifdef AVX512
        case 8:
            res = algo.AlgoForce2VecNS_SREAL_D128_AVXZ(iterations, threads, batch, ns, lr);
            break;
        case 9:
            if (dim == 128) res = algo.AlgoForce2VecNSRW_SREAL_D128_AVXZ(iterations, threads, batch, ns, lr);
            else if (dim == 64) res = algo.AlgoForce2VecNSRWLB_SREAL_D64_AVXZ(iterations, threads, batch, ns, lr);
            else return -2.0;
            break;
        case 10:
            if (dim == 128) res = algo.AlgoForce2VecNSRWEFF_SREAL_D128_AVXZ(iterations, threads, batch, ns, lr);
            else if (dim == 64) res = algo.AlgoForce2VecNSRWEFF_SREAL_D64_AVXZ(iterations, threads, batch, ns, lr);
            else return -2.0;
            break;
        case 11:
            if (dim == 128) res = algo.AlgoForce2VecNSLB_SREAL_D128_AVXZ(iterations, threads, batch, ns, lr);
            else if (dim == 64) res = algo.AlgoForce2VecNSLB_SREAL_D64_AVXZ(iterations, threads, batch, ns, lr);
            else return -2.0;
            break;
#endif
        default:
            return -2.0;
        }
        if (X_out) memcpy(X_out, algo.nCoordinates, sizeof(float) * (size_t)n * dim);
    }
    return res.empty() ? -3.0 : (double)res[0];
}

// sample/algorithms.cpp:755-764: the reference's own sigmoid table (global sm_table).
void f2vref_lut(float* out2048)
{
    init_SM_TABLE();
    memcpy(out2048, sm_table, sizeof(float) * SM_TABLE_SIZE);
}

int f2vref_has_avx512(void)
{
#ifdef AVX512
    return 1;
#else
    return 0;
#endif
}

}  // extern "C"
