# Build the llark_b200 C-ABI library (sm_100a only) and the oracle's native pieces.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -Xptxas -v
SRC := $(wildcard llark_b200/csrc/*.cu)
OBJ := $(patsubst llark_b200/csrc/%.cu,build/%.o,$(SRC))
LIB := llark_b200/libllark_b200.so

all: $(LIB)

build/%.o: llark_b200/csrc/%.cu llark_b200/csrc/common.cuh include/llark_b200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lnccl

tools: build/gemm_check
build/gemm_check: tools/gemm_check.cu $(LIB)
	$(NVCC) $(ARCH) -std=c++17 -O2 -o $@ tools/gemm_check.cu -Lllark_b200 -lllark_b200 -Xlinker -rpath -Xlinker '$$ORIGIN/../llark_b200'

clean:
	rm -rf build $(LIB)
.PHONY: all clean tools
