# Build of the B200-native fixedL path.  nvcc cross-compiles sm_100a without a GPU.
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       ?= g++
NVCCFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC
CSRC      := $(sort $(wildcard tnml_b200/csrc/*.cu))
LIB       := tnml_b200/libtnml_b200.so
HOSTBIN   := tnml_b200/host/fixedL

all: lib host oracle

lib: $(LIB)
$(LIB): $(CSRC) $(wildcard tnml_b200/csrc/*.cuh) include/tnml_b200.h
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(CSRC) -ldl

# drop-in `fixedL <inputfile>` program (host C++ over the C-ABI)
host: $(HOSTBIN) tnml_b200/host/fulltest tnml_b200/host/hosttest
tnml_b200/host/hosttest: tnml_b200/host/hosttest.cc tnml_b200/host/itensor_lite.h tnml_b200/host/initial_w.h
	$(CXX) -O2 -std=c++17 -Wall -o $@ tnml_b200/host/hosttest.cc
tnml_b200/host/fulltest: tnml_b200/host/fulltest.cc tnml_b200/host/itensor_lite.h tnml_b200/host/mnist.h include/tnml_b200.h $(LIB)
	$(CXX) -O2 -std=c++17 -Wall -o $@ tnml_b200/host/fulltest.cc -Ltnml_b200 -ltnml_b200 -Wl,-rpath,'$$ORIGIN/..' -pthread
$(HOSTBIN): tnml_b200/host/fixedL.cc tnml_b200/host/initial_w.h tnml_b200/host/itensor_lite.h tnml_b200/host/mnist.h include/tnml_b200.h $(LIB)
	$(CXX) -O2 -std=c++17 -Wall -o $@ tnml_b200/host/fixedL.cc -Ltnml_b200 -ltnml_b200 -Wl,-rpath,'$$ORIGIN/..' -pthread

# CPU restatement used as checker / baseline only (never by the product)
oracle: oracle/_build/fixedl_ref_cpu
oracle/_build/fixedl_ref_cpu: oracle/fixedl_ref_cpu.cpp
	mkdir -p oracle/_build
	$(CXX) -O3 -march=native -std=c++17 -pthread -o $@ $<

clean:
	rm -f $(LIB) $(HOSTBIN) tnml_b200/host/fulltest tnml_b200/host/hosttest oracle/_build/fixedl_ref_cpu

.PHONY: all lib host oracle clean

# micro-benchmarks behind the design decisions (run on the GPU box)
TOOLS := tools/dmma_bench tools/dmma_lds_bench tools/imma_bench tools/lat_bench tools/krgemm_bench_base tools/umma_bench tools/oz_test
tools: $(TOOLS)
tools/krgemm_bench_base: tools/krgemm_bench.cu $(CSRC)
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o $@ $<
tools/oz_test: tools/oz_test.cu tnml_b200/csrc/tnml_ozaki.cu tnml_b200/csrc/tnml_kernels.cu tnml_b200/csrc/tnml_kernels.cuh
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o $@ tools/oz_test.cu tnml_b200/csrc/tnml_ozaki.cu tnml_b200/csrc/tnml_kernels.cu
tools/%: tools/%.cu
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o $@ $<
