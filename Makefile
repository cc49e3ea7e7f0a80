# Build of the product: libkcgpu.so (CUDA, sm_100a) and the host CLI `kmercamel` (C++ over the C ABI).
NVCC ?= nvcc
CXX ?= g++
CSRC := kmercamel_b200/csrc
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --extended-lambda \
           -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v
HDRS := $(wildcard $(CSRC)/*.cuh) include/kcgpu.h

all: $(CSRC)/libkcgpu.so host/kmercamel

$(CSRC)/kcgpu.o: $(CSRC)/kcgpu.cu $(HDRS)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(CSRC)/ptxas.log || (tail -50 $(CSRC)/ptxas.log; false)

$(CSRC)/framing.o: $(CSRC)/framing.cpp include/kcgpu.h
	$(CXX) -std=c++17 -O2 -fPIC -Wall -c $< -o $@

$(CSRC)/libkcgpu.so: $(CSRC)/kcgpu.o $(CSRC)/framing.o
	$(NVCC) -shared -o $@ $^

host/kmercamel: host/main.cpp include/kcgpu.h $(CSRC)/libkcgpu.so
	$(CXX) -std=c++17 -O2 -Wall -pthread -Iinclude host/main.cpp -o $@ -L$(CSRC) -lkcgpu -lz -Wl,-rpath,'$$ORIGIN/../$(CSRC)'

# tests-only: serial host emulation of the engine/emission control flow (no GPU needed, not part of the product)
tests/host_emul: tests/host_emul.cu $(HDRS)
	$(NVCC) -O2 -std=c++17 --extended-lambda -DKC_HOST_EMUL -Xcompiler -Wall,-Wno-unused-function -o $@ $<

clean:
	rm -f $(CSRC)/*.o $(CSRC)/libkcgpu.so host/kmercamel tests/host_emul $(CSRC)/ptxas.log

.PHONY: all clean
