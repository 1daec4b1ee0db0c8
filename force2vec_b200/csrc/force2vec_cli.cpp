// force2vec_b200/csrc/force2vec_cli.cpp -- bin/Force2Vec, the drop-in command line.
// Same flags, defaults, stdout lines, output-file naming and Results.txt row as the
// reference driver (/root/reference/Test/Force2Vec.cpp:22-47 help, :54-116 argv loop,
// :121-150 dispatch, :191-198 Results.txt), for options 5, 6 and 7; the force step runs
// on the GPU.  Extra flags (ignored by the reference): -device <int>, -gpus <int>, -mode <0|2>, -chunk <int>
// (engine epoch mode), -walk <0|1> (0 = libc-stream host walks, 1 = device sampler).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>
#include <omp.h>

#include "f2v_host.hpp"

using namespace std;

static void helpmessage() {
    printf("\n");
    printf("Usage of Force2Vec tool:\n");
    printf("-input <string>, full path of input file (required).\n");
    printf("-output <string>, directory where output file will be stored. (default: current directory)\n");
    printf("-batch <int>, size of minibatch. (default:384)\n");
    printf("-iter <int>, number of iteration. (default:1200)\n");
    printf("-threads <int>, accepted for compatibility; the force step runs on the GPU.\n");
    printf("-dim <int>, size of embedding dimension. (default:128) \n");
    printf("-nsamples <int>, number of negative samples. (default:5) \n");
    printf("-lr <float>, learning rate of SGD. (default:0.02)\n");
    printf("-bs <int>, 1 = per-vertex negative samples (options 5, 6). (default:0)\n");
    printf("-option <int>, 5, 6 or 7. (default:5)\n");
    printf("        -option 5 - for t-distribution + negative sampling (tForce2Vec).\n");
    printf("        -option 6 - for sigmoid + negative sampling (sForce2Vec).\n");
    printf("        -option 7 - for sigmoid + semi-random walk (rForce2Vec).\n");
    printf("-device <int>, (first) CUDA device. (default:0)\n");
    printf("-gpus <int>, number of GPUs; minibatches are split across them. (default:1)\n");
    printf("-mode <int>, 0 = one launch per minibatch, 2 = one dataflow launch per epoch. (default:0)\n");
    printf("-chunk <int>, hub rows longer than this are split across warps; equal values give equal bits on any GPU count. (default:0 = auto)\n");
    printf("-walk <int>, option 7 walks: 0 = host (reference stream), 1 = device sampler. (default:0)\n");
    printf("-h, show help message.\n");
}

int main(int argc, char* argv[]) {
    float gamma = 1.0f, lr = 0.02f;
    uint32_t batchsize = 384, iterations = 1200, numberOfThreads = (uint32_t)omp_get_max_threads(), dim = 128,
             option = 5, nsamples = 5, bs = 0;
    int device = 0, mode = 0, walk = 0, gpus = 1, chunk = 0;
    string inputfile = "", outputfile = "", algoname = "Force2Vec:t-distribution with negative sampling",
           initname = "RAND";
    for (int p = 0; p < argc; p++) {
        const bool has_arg = p + 1 < argc;
        if (strcmp(argv[p], "-h") == 0) { helpmessage(); exit(1); }
        if (!has_arg) continue;
        if (strcmp(argv[p], "-input") == 0) inputfile = argv[p + 1];
        else if (strcmp(argv[p], "-output") == 0) outputfile = argv[p + 1];
        else if (strcmp(argv[p], "-batch") == 0) batchsize = atoi(argv[p + 1]);
        else if (strcmp(argv[p], "-iter") == 0) iterations = atoi(argv[p + 1]);
        else if (strcmp(argv[p], "-threads") == 0) numberOfThreads = atoi(argv[p + 1]);
        else if (strcmp(argv[p], "-dim") == 0) dim = atoi(argv[p + 1]);
        else if (strcmp(argv[p], "-gamma") == 0) gamma = (float)atof(argv[p + 1]);
        else if (strcmp(argv[p], "-bs") == 0) bs = atoi(argv[p + 1]);
        else if (strcmp(argv[p], "-lr") == 0) lr = (float)atof(argv[p + 1]);
        else if (strcmp(argv[p], "-nsamples") == 0) nsamples = atoi(argv[p + 1]);
        else if (strcmp(argv[p], "-device") == 0) device = atoi(argv[p + 1]);
        else if (strcmp(argv[p], "-gpus") == 0) gpus = atoi(argv[p + 1]);
        else if (strcmp(argv[p], "-mode") == 0) mode = atoi(argv[p + 1]);
        else if (strcmp(argv[p], "-walk") == 0) walk = atoi(argv[p + 1]);
        else if (strcmp(argv[p], "-chunk") == 0) chunk = atoi(argv[p + 1]);
        else if (strcmp(argv[p], "-option") == 0) {
            option = atoi(argv[p + 1]);
            if (option == 5) algoname = "Force2Vec:t-distribution with negative sampling";
            else if (option == 6) algoname = "Force2Vec:sigmoid with negative sampling";
            else if (option == 7) algoname = "Force2Vec:sigmoid based random-walk";
        }
    }
    if (inputfile.size() == 0) {
        printf("Valid input file needed!...\n");
        exit(1);
    }
    if (option < 5 || option > 7) {
        printf("This build implements -option 5, 6 and 7 (the GPU force-step path) only.\n");
        exit(1);
    }
    if (batchsize == 0 || dim == 0) {
        printf("-batch and -dim must be positive.\n");
        exit(1);
    }
    f2v::Csr A_csr;
    string err;
    if (!f2v::SetInputMatricesAsCSR(A_csr, inputfile, &err)) {
        printf("%s\n", err.c_str());
        exit(1);
    }
    f2v::algorithms algo(A_csr, inputfile, outputfile, dim, gamma, batchsize);
    algo.device = device;
    algo.gpus = gpus;
    algo.epoch_mode = mode;
    algo.walk_sampler = walk;
    algo.chunk = (uint32_t)(chunk > 0 ? chunk : 0);
    cout << "Running: " << algoname << endl;
    vector<float> outputvec;
    if (option == 5)
        outputvec = bs == 0 ? algo.AlgoForce2VecNS(iterations, numberOfThreads, batchsize, nsamples, lr)
                            : algo.AlgoForce2VecNSBS(iterations, numberOfThreads, batchsize, nsamples, lr);
    else if (option == 6)
        outputvec = bs == 0 ? algo.AlgoForce2VecNSRW(iterations, numberOfThreads, batchsize, nsamples, lr)
                            : algo.AlgoForce2VecNSRWBS(iterations, numberOfThreads, batchsize, nsamples, lr);
    else
        outputvec = algo.AlgoForce2VecNSRWEFF(iterations, numberOfThreads, batchsize, nsamples, lr);

    ofstream output;
    output.open("Results.txt", ofstream::app);
    output << "Algo:" << algoname << "\tInit:" << initname << "\tIteration:";
    output << iterations << "\tNumofthreads:" << numberOfThreads << "\tBatchSize:" << batchsize << "\tDimension:" << dim
           << "\tTime(sec.):";
    output << outputvec[0] << "\t";
    output << endl;
    output.close();
    return 0;
}
