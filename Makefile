# Builds the C-ABI CUDA library in-tree (sm_100a only) and the CPU-side test binaries.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall
SRC := sm3det_b200/csrc
OBJ := build/obj
LIB := sm3det_b200/lib/libsm3det_b200.so
SRCS := common.cu gemm_tc.cu ffn_fused.cu norm.cu stencil.cu moe.cu reduce.cu act.cu lsk.cu neck.cu capi.cu
OBJS := $(SRCS:%.cu=$(OBJ)/%.o)

all: $(LIB)

$(OBJ)/%.o: $(SRC)/%.cu $(wildcard $(SRC)/*.cuh) $(SRC)/kernels.h include/sm3det_b200.h
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	@mkdir -p sm3det_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart

build/gemm_test: tests/cuda/gemm_test.cu $(SRC)/gemm_tc.cu $(SRC)/common.cu $(SRC)/gemm_tc.cuh
	@mkdir -p build
	$(NVCC) $(ARCH) -O3 -std=c++17 -lineinfo -I $(SRC) tests/cuda/gemm_test.cu $(SRC)/gemm_tc.cu $(SRC)/common.cu -o $@

build/ffn_test: tests/cuda/ffn_test.cu $(SRC)/ffn_fused.cu $(SRC)/gemm_tc.cu $(SRC)/common.cu $(SRC)/gemm_tc.cuh $(SRC)/ffn_fused.cuh
	@mkdir -p build
	$(NVCC) $(ARCH) -O3 -std=c++17 -lineinfo -I $(SRC) tests/cuda/ffn_test.cu $(SRC)/ffn_fused.cu $(SRC)/gemm_tc.cu $(SRC)/common.cu -o $@

build/mma_bench: tests/cuda/mma_bench.cu $(SRC)/gemm_tc.cuh $(SRC)/common.cu
	@mkdir -p build
	$(NVCC) $(ARCH) -O3 -std=c++17 -lineinfo -I $(SRC) tests/cuda/mma_bench.cu $(SRC)/common.cu -o $@

clean:
	rm -rf build sm3det_b200/lib/*.so

.PHONY: all clean
