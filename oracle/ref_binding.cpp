// oracle/ref_binding.cpp -- INTEGRATION.md section B, compiled: the binding a reference maintainer
// adds.  This translation unit REPLACES sample/algorithms.cpp in a build of the reference's own CLI
// (Test/Force2Vec.cpp: argv parsing, the IO.h/CSC.h/CSR.h loaders, srand(1), Results.txt) and its own
// class (sample/algorithms.h: constructor, nCoordinates, writeToFile): the five hot-path method bodies
// (sample/algorithms.cpp:544-652, 654-753, 778-932, 934-1060, 1063-1203) become calls into libf2v.so
// through the C ABI (include/f2v.h, include/f2v_host.h).  Test infrastructure: oracle/Makefile builds
// oracle/_ref/Force2Vec_f2v from it where /root/reference exists; tests/test_gpu_x_boundary_and_sampler.py checks that
// its .embd is byte-identical to bin/Force2Vec's.  Nothing of the reference is copied: its headers and
// driver are compiled from where they lie.
#include "f2v.h"
#include "f2v_host.h"

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "algorithms.h"          // the reference's class (note: it #defines `t`)

static vector<VALUETYPE> f2v_run(algorithms* self, int option, int bs, INDEXTYPE ITERATIONS,
                                 INDEXTYPE BATCHSIZE, INDEXTYPE ns, VALUETYPE lr) {
    // CSR<unsigned,float> -> the ABI's CSR: rowptr widened to 64 bit, colids as they are
    std::vector<uint64_t> rowptr(self->graph.rowptr, self->graph.rowptr + self->graph.rows + 1);
    f2v_train_args a = {};
    a.n = self->graph.rows;
    a.nnz = self->graph.nnz;
    a.rowptr = rowptr.data();
    a.colids = self->graph.colids;
    a.dim = self->DIM;
    a.option = option;
    a.bs = bs;
    a.iterations = ITERATIONS;
    a.batch = BATCHSIZE;
    a.nsamples = ns;
    a.lr = lr;
    a.seed = 1;              // Test/Force2Vec.cpp:126 srand(1): the engine draws the same stream itself
    a.device = 0;
    double sec = 0;
    if (f2v_train(&a, self->nCoordinates, &sec) != F2V_OK) {       // error convention of the
        fprintf(stderr, "Force2Vec: %s\n", f2v_last_error());      // reference: message + exit(1)
        exit(1);
    }
    return vector<VALUETYPE>{(VALUETYPE)sec};
}

static string tag(const char* prefix, INDEXTYPE B, INDEXTYPE D, INDEXTYPE IT, INDEXTYPE ns) {
    return string(prefix) + to_string(B) + "D" + to_string(D) + "IT" + to_string(IT) + "NS" + to_string(ns);
}

vector<VALUETYPE> algorithms::AlgoForce2VecNS(INDEXTYPE ITERATIONS, INDEXTYPE NUMOFTHREADS, INDEXTYPE BATCHSIZE, INDEXTYPE ns, VALUETYPE lr) {
    vector<VALUETYPE> result = f2v_run(this, F2V_TDIST, 0, ITERATIONS, BATCHSIZE, ns, lr);
    cout << "Force2Vec Parallel Wall time required:" << result[0] << " seconds" << endl;
    writeToFile(tag("F2VNS", BATCHSIZE, this->DIM, ITERATIONS, ns));           // unchanged, algorithms.h:118
    return result;
}
vector<VALUETYPE> algorithms::AlgoForce2VecNSBS(INDEXTYPE ITERATIONS, INDEXTYPE NUMOFTHREADS, INDEXTYPE BATCHSIZE, INDEXTYPE ns, VALUETYPE lr) {
    vector<VALUETYPE> result = f2v_run(this, F2V_TDIST, 1, ITERATIONS, BATCHSIZE, ns, lr);
    cout << "Force2Vec Parallel Wall time required (with BS negative samples):" << result[0] << " seconds" << endl;
    writeToFile(tag("F2VNS", BATCHSIZE, this->DIM, ITERATIONS, ns));
    return result;
}
vector<VALUETYPE> algorithms::AlgoForce2VecNSRW(INDEXTYPE ITERATIONS, INDEXTYPE NUMOFTHREADS, INDEXTYPE BATCHSIZE, INDEXTYPE ns, VALUETYPE lr) {
    vector<VALUETYPE> result = f2v_run(this, F2V_SIGMOID, 0, ITERATIONS, BATCHSIZE, ns, lr);
    cout << "Force2Vec Parallel Wall time required:" << result[0] << " seconds" << endl;
    writeToFile(tag("F2VWNS", BATCHSIZE, this->DIM, ITERATIONS, ns));
    return result;
}
vector<VALUETYPE> algorithms::AlgoForce2VecNSRWBS(INDEXTYPE ITERATIONS, INDEXTYPE NUMOFTHREADS, INDEXTYPE BATCHSIZE, INDEXTYPE ns, VALUETYPE lr) {
    vector<VALUETYPE> result = f2v_run(this, F2V_SIGMOID, 1, ITERATIONS, BATCHSIZE, ns, lr);
    cout << "Force2Vec Parallel Wall time required (with BS negative samples):" << result[0] << " seconds" << endl;
    writeToFile(tag("F2VWNS", BATCHSIZE, this->DIM, ITERATIONS, ns));
    return result;
}
vector<VALUETYPE> algorithms::AlgoForce2VecNSRWEFF(INDEXTYPE ITERATIONS, INDEXTYPE NUMOFTHREADS, INDEXTYPE BATCHSIZE, INDEXTYPE ns, VALUETYPE lr) {
    vector<VALUETYPE> result = f2v_run(this, F2V_WALK, 0, ITERATIONS, BATCHSIZE, ns, lr);
    cout << "Force2VecWNSEFF Parallel Wall time required:" << result[0] << " seconds" << endl;
    writeToFile(tag("F2VWNSF", BATCHSIZE, this->DIM, ITERATIONS, ns));
    return result;
}

// options 1-4 are outside the accelerated path (out of scope): the driver still links
static vector<VALUETYPE> not_bound(const char* what) {
    fprintf(stderr, "%s is not bound to libf2v.so (options 5, 6, 7 are)\n", what);
    exit(1);
    return vector<VALUETYPE>();
}
vector<VALUETYPE> algorithms::AlgoForce2Vec(INDEXTYPE, INDEXTYPE, INDEXTYPE) { return not_bound("option 1"); }
vector<VALUETYPE> algorithms::AlgoForce2VecFR(INDEXTYPE, INDEXTYPE, INDEXTYPE, INDEXTYPE, VALUETYPE) { return not_bound("option 2"); }
vector<VALUETYPE> algorithms::AlgoForce2VecLL(INDEXTYPE, INDEXTYPE, INDEXTYPE, INDEXTYPE, VALUETYPE) { return not_bound("option 3"); }
vector<VALUETYPE> algorithms::AlgoForce2VecFA(INDEXTYPE, INDEXTYPE, INDEXTYPE, INDEXTYPE, VALUETYPE) { return not_bound("option 4"); }
