# Top-level build: the sm_100a engine library (libf2v.so), the drop-in CLI (bin/Force2Vec)
# and the test checkers under oracle/.   `make` = everything; `make lib` = library only.
NVCC     ?= /usr/local/cuda/bin/nvcc
HOSTCXX   = /usr/bin/g++
ARCH      = -gencode arch=compute_100a,code=sm_100a
NVFLAGS   = $(ARCH) -O3 -std=c++17 -lineinfo -ccbin $(HOSTCXX) -Xcompiler -fPIC,-fopenmp,-Wall -Xptxas -v
PKG       = force2vec_b200
CSRC      = $(PKG)/csrc
LIB       = $(PKG)/lib/libf2v.so

all: lib cli oracle

lib: $(LIB)

$(LIB): $(CSRC)/f2v_engine.cu $(CSRC)/f2v_kernels.cuh $(CSRC)/f2v_host.cpp $(CSRC)/f2v_host.hpp include/f2v.h include/f2v_host.h
	mkdir -p $(PKG)/lib
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CSRC)/f2v_engine.cu $(CSRC)/f2v_host.cpp -lgomp -ldl 2> $(PKG)/lib/ptxas.log || (cat $(PKG)/lib/ptxas.log; false)

cli: bin/Force2Vec

bin/Force2Vec: $(CSRC)/force2vec_cli.cpp $(LIB) include/f2v.h include/f2v_host.h
	mkdir -p bin
	$(HOSTCXX) -O2 -std=c++17 -fopenmp -Wall -Iinclude -o $@ $(CSRC)/force2vec_cli.cpp -L$(PKG)/lib -lf2v -Wl,-rpath,'$$ORIGIN/../$(PKG)/lib'

oracle:
	$(MAKE) -C oracle

sass: $(LIB)
	mkdir -p profiles
	/usr/local/cuda/bin/cuobjdump -sass $(LIB) > profiles/libf2v.sass

clean:
	rm -rf $(PKG)/lib bin
	$(MAKE) -C oracle clean
.PHONY: all lib cli oracle sass clean
