# Build of the B200-native BreakDancerMax hot path.
#   make            -> breakdancer_b200/libbdk.so (C ABI: CUDA kernels + host decode/format),
#                      breakdancer_b200/bin/breakdancer_max (drop-in CLI), oracle/_build/libbdoracle.so
#   make ref        -> oracle/_ref/ (the unmodified reference, only where /root/reference exists)
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       := $(shell test -x /usr/bin/g++ && echo /usr/bin/g++ || echo g++)
CUDA_HOME ?= /usr/local/cuda
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := -ccbin $(CXX) $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-O3,-Wall,-Wno-unused-function --fmad=false -Xptxas -v
CXXFLAGS  := -O3 -std=c++17 -fPIC -Wall -I$(CUDA_HOME)/include
LDLIBS    := -L$(CUDA_HOME)/lib64 -lcudart_static -lz -lpthread -ldl -lrt

PKG   := breakdancer_b200
SRC   := $(PKG)/csrc
B     := build
LIB   := $(PKG)/libbdk.so
CLI   := $(PKG)/bin/breakdancer_max
B2C   := $(PKG)/bin/bam2cfg
ORA   := oracle/_build/libbdoracle.so

HOST_SRCS := $(SRC)/host/config.cpp $(SRC)/host/bam_io.cpp $(SRC)/host/format.cpp $(SRC)/host/options.cpp $(SRC)/host/support.cpp
HOST_OBJS := $(patsubst $(SRC)/host/%.cpp,$(B)/host_%.o,$(HOST_SRCS))
CU_HDRS   := $(wildcard $(SRC)/*.cuh) $(wildcard $(SRC)/*.h) $(wildcard $(SRC)/*.inl) include/bdk.h

all: $(LIB) $(CLI) $(B2C) $(ORA)

$(B)/host_%.o: $(SRC)/host/%.cpp $(SRC)/host/host.hpp $(SRC)/host/cli.hpp $(SRC)/host/bam2cfg_impl.hpp include/bdk.h include/bdk_host.h
	@mkdir -p $(B)
	$(CXX) $(CXXFLAGS) -c $< -o $@

$(B)/bdk_core.o: $(SRC)/bdk_core.cu $(CU_HDRS)
	@mkdir -p $(B)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(B)/ptxas_bdk_core.log || (cat $(B)/ptxas_bdk_core.log; false)

$(LIB): $(B)/bdk_core.o $(HOST_OBJS)
	$(CXX) -shared -o $@ $^ $(LDLIBS)

$(CLI): $(SRC)/host/main.cpp $(LIB) $(SRC)/host/host.hpp
	@mkdir -p $(PKG)/bin
	$(CXX) $(CXXFLAGS) $< -o $@ -L$(PKG) -lbdk -Wl,-rpath,'$$ORIGIN/..' $(LDLIBS)

$(B2C): $(SRC)/host/bam2cfg_main.cpp $(LIB) include/bdk_host.h
	@mkdir -p $(PKG)/bin
	$(CXX) $(CXXFLAGS) $< -o $@ -L$(PKG) -lbdk -Wl,-rpath,'$$ORIGIN/..' $(LDLIBS)

$(ORA): oracle/bd_oracle.cpp include/bdk.h
	@mkdir -p oracle/_build
	$(CXX) -O2 -std=c++17 -fPIC -Wall -shared $< -o $@

ref:
	bash oracle/build_ref.sh

clean:
	rm -rf $(B) $(LIB) $(CLI) $(B2C) oracle/_build

.PHONY: all ref clean
