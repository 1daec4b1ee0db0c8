// force2vec_b200/csrc/f2v_host.hpp -- C++ host mirror of the reference's `class algorithms`
// (/root/reference/sample/algorithms.h:51-137) for options 5/6/7: same constructor
// arguments, same method names / argument meaning / return value ({wall seconds}), same
// output-file naming.  The method bodies run on the GPU through the C ABI (include/f2v.h);
// nothing here computes forces on the CPU.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace f2v {

// Sets the calling thread's f2v_last_error() message and returns `code` (defined in f2v_engine.cu).
int host_fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));

// CSR as the reference's CSR<INDEXTYPE,VALUETYPE> (sample/CSR.h:89-96) holds it, minus the
// unused `values`; rowptr widened to 64 bit (the reference wraps at 2^32).
struct Csr {
    uint64_t rows = 0, nnz = 0;
    std::vector<uint64_t> rowptr;
    std::vector<uint32_t> colids;
};

// SetInputMatricesAsCSR (sample/commonutility.h:44-54).  Returns false + message on error.
bool SetInputMatricesAsCSR(Csr& A, const std::string& path, std::string* err = nullptr);

class algorithms {
public:
    Csr graph;
    std::vector<float> nCoordinates;   // row-major rows x DIM (algorithms.h:54,68)
    float GAMMA = 1.0f;
    uint32_t DIM;
    std::string filename;
    std::string outputdir;
    int device = 0;                    // (first) CUDA device (no reference counterpart)
    int gpus = 1;                      // > 1: devices device..device+gpus-1, minibatches split across them
    int epoch_mode = 0;                // engine execution mode (include/f2v.h f2v_set_epoch_mode)
    int walk_sampler = 0;              // 0 = libc-stream host walks, 1 = device sampler
    uint32_t chunk = 0;                // hub-row chunk length (0 = default; equal values give equal bits on any GPU count)
    uint32_t seed = 1;                 // Test/Force2Vec.cpp:126 srand(1)

    algorithms(const Csr& A_csr, std::string input, std::string outputd, uint32_t dim, float gm, uint32_t bsize);

    // NUMOFTHREADS is accepted for signature compatibility and ignored (the work runs on the GPU).
    std::vector<float> AlgoForce2VecNS(uint32_t ITERATIONS, uint32_t NUMOFTHREADS, uint32_t BATCHSIZE, uint32_t ns, float lr);
    std::vector<float> AlgoForce2VecNSBS(uint32_t ITERATIONS, uint32_t NUMOFTHREADS, uint32_t BATCHSIZE, uint32_t ns, float lr);
    std::vector<float> AlgoForce2VecNSRW(uint32_t ITERATIONS, uint32_t NUMOFTHREADS, uint32_t BATCHSIZE, uint32_t ns, float lr);
    std::vector<float> AlgoForce2VecNSRWBS(uint32_t ITERATIONS, uint32_t NUMOFTHREADS, uint32_t BATCHSIZE, uint32_t ns, float lr);
    std::vector<float> AlgoForce2VecNSRWEFF(uint32_t ITERATIONS, uint32_t NUMOFTHREADS, uint32_t BATCHSIZE, uint32_t ns, float lr);

    // algorithms.h:118-136: <outputdir><basename(input)><f>.embd
    void writeToFile(std::string f);

private:
    std::vector<float> run(int option, int bs, uint32_t iters, uint32_t batch, uint32_t ns, float lr,
                           const char* banner, const std::string& tag);
};

}  // namespace f2v
