# Build the llark_b200 C-ABI library (sm_100a only).
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

# the ALiBi instantiation #includes the attention source: rebuild it when that changes
build/attention_tc_alibi.o: llark_b200/csrc/attention_tc.cu

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lnccl

clean:
	rm -rf build $(LIB)
.PHONY: all clean
