# Build everything in-tree.  The .so files are git-ignored but travel to the GPU box.
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       ?= g++
CC        ?= gcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v $(if $(DEBUG),-DKA_DEBUG,)
PKG       := kmers.anno_b200
CSRC      := $(PKG)/csrc
HOST      := $(PKG)/host

LIB       := $(PKG)/libkmeranno.so
SYNTH     := $(PKG)/libkasynth.so
ORACLE    := oracle/libkaoracle.so
CLI       := $(PKG)/bin/kmers-anno
SELFTEST  := $(PKG)/bin/kmers-anno-selftest

all: $(LIB) $(SYNTH) $(ORACLE) $(CLI) $(SELFTEST)

CUSRC     := $(CSRC)/ka_kernels.cu $(CSRC)/ka_distance.cu $(CSRC)/ka_engine.cu $(CSRC)/ka_table.cu $(CSRC)/ka_route.cu $(CSRC)/ka_build_api.cu $(CSRC)/ka_line.cu

$(LIB): $(CUSRC) $(CSRC)/ka_common.cuh $(CSRC)/ka_kernels.cuh $(CSRC)/ka_engine_internal.cuh $(CSRC)/ka_line.cuh include/kmeranno.h
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CUSRC) -ldl 2> $(PKG)/ptxas.log || (cat $(PKG)/ptxas.log; exit 1)
	@grep -E "error|warning|spill|registers" $(PKG)/ptxas.log | grep -v "0 bytes spill" | head -40 || true

$(SYNTH): $(HOST)/ka_synth.cpp
	$(CXX) -O3 -std=c++17 -fPIC -shared -pthread -Wall -o $@ $<

$(ORACLE): oracle/ka_oracle.c oracle/ka_oracle_fast.c
	$(CC) -O2 -fPIC -shared -pthread -Wall -o $@ oracle/ka_oracle.c oracle/ka_oracle_fast.c

$(CLI): $(wildcard $(HOST)/*.cpp) $(wildcard $(HOST)/*.hpp) $(LIB)
	mkdir -p $(PKG)/bin
	$(CXX) -O2 -std=c++17 -Wall -Iinclude -o $@ $(HOST)/App.cpp $(HOST)/ApplyKmerProcessor.cpp $(HOST)/BuildKmerProcessor.cpp $(HOST)/GeneCopyProcessor.cpp $(HOST)/Genome.cpp $(HOST)/PackedBatch.cpp -L$(PKG) -lkmeranno -Wl,-rpath,'$$ORIGIN/..' -pthread

$(SELFTEST): $(HOST)/selftest.cpp $(HOST)/Genome.cpp $(wildcard $(HOST)/*.hpp)
	mkdir -p $(PKG)/bin
	$(CXX) -O2 -std=c++17 -Wall -Iinclude -o $@ $(HOST)/selftest.cpp $(HOST)/Genome.cpp

# bounds-checking build of the kernels (KA_CHECK): KMERANNO_LIB=kmers.anno_b200/libkmeranno_dbg.so python -m pytest ...
debuglib: $(CUSRC)
	$(NVCC) $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DKA_DEBUG -shared -o $(PKG)/libkmeranno_dbg.so $(CUSRC) -ldl

clean:
	rm -f $(LIB) $(SYNTH) $(ORACLE) $(CLI) $(SELFTEST) $(PKG)/ptxas.log

.PHONY: all clean debuglib
