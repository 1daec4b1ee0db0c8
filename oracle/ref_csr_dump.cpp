// oracle/ref_csr_dump.cpp -- checker for the LOADER half of the drop-in boundary (test infrastructure).
// Like ref_binding.cpp this translation unit replaces sample/algorithms.cpp in a build of the
// reference's own CLI (Test/Force2Vec.cpp with its IO.h / CSC.h / CSR.h loaders, compiled from where
// they lie), but the hot-path methods only dump the CSR the reference driver hands them
// (Test/Force2Vec.cpp:121-127: SetInputMatricesAsCSR, Sorted(), the class's copy) to ./csr_dump.bin:
//   u64 rows, u64 nnz, u64 rowptr[rows+1], u32 colids[nnz].
// tests/test_host.py compares it with what force2vec_b200's own loader (f2v_load_mtx) builds from the
// same file: the byte-equality of the two CLIs' .embd files rests on equal neighbour order.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "algorithms.h"

static vector<VALUETYPE> dump(algorithms* self) {
    FILE* f = fopen("csr_dump.bin", "wb");
    if (!f) { perror("csr_dump.bin"); exit(1); }
    const uint64_t rows = self->graph.rows, nnz = self->graph.nnz;
    std::vector<uint64_t> rowptr(self->graph.rowptr, self->graph.rowptr + rows + 1);
    std::vector<uint32_t> colids(self->graph.colids, self->graph.colids + nnz);
    bool ok = fwrite(&rows, 8, 1, f) == 1 && fwrite(&nnz, 8, 1, f) == 1 &&
              fwrite(rowptr.data(), 8, rows + 1, f) == rows + 1 && (nnz == 0 || fwrite(colids.data(), 4, nnz, f) == nnz);
    ok = fclose(f) == 0 && ok;
    if (!ok) { fprintf(stderr, "csr_dump.bin: short write\n"); exit(1); }
    return vector<VALUETYPE>{(VALUETYPE)0};
}

vector<VALUETYPE> algorithms::AlgoForce2VecNS(INDEXTYPE, INDEXTYPE, INDEXTYPE, INDEXTYPE, VALUETYPE) { return dump(this); }
vector<VALUETYPE> algorithms::AlgoForce2VecNSBS(INDEXTYPE, INDEXTYPE, INDEXTYPE, INDEXTYPE, VALUETYPE) { return dump(this); }
vector<VALUETYPE> algorithms::AlgoForce2VecNSRW(INDEXTYPE, INDEXTYPE, INDEXTYPE, INDEXTYPE, VALUETYPE) { return dump(this); }
vector<VALUETYPE> algorithms::AlgoForce2VecNSRWBS(INDEXTYPE, INDEXTYPE, INDEXTYPE, INDEXTYPE, VALUETYPE) { return dump(this); }
vector<VALUETYPE> algorithms::AlgoForce2VecNSRWEFF(INDEXTYPE, INDEXTYPE, INDEXTYPE, INDEXTYPE, VALUETYPE) { return dump(this); }
vector<VALUETYPE> algorithms::AlgoForce2Vec(INDEXTYPE, INDEXTYPE, INDEXTYPE) { return dump(this); }
vector<VALUETYPE> algorithms::AlgoForce2VecFR(INDEXTYPE, INDEXTYPE, INDEXTYPE, INDEXTYPE, VALUETYPE) { return dump(this); }
vector<VALUETYPE> algorithms::AlgoForce2VecLL(INDEXTYPE, INDEXTYPE, INDEXTYPE, INDEXTYPE, VALUETYPE) { return dump(this); }
vector<VALUETYPE> algorithms::AlgoForce2VecFA(INDEXTYPE, INDEXTYPE, INDEXTYPE, INDEXTYPE, VALUETYPE) { return dump(this); }
