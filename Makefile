# Builds the C-ABI CUDA library (sm_100a only) and the oracle's C pieces.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v --expt-relaxed-constexpr
SRC := $(wildcard mmtg_b200/csrc/*.cu)
OBJ := $(patsubst mmtg_b200/csrc/%.cu,build/%.o,$(SRC))
LIB := mmtg_b200/lib/libmmtg_b200.so

all: $(LIB)

build/%.o: mmtg_b200/csrc/%.cu $(wildcard mmtg_b200/csrc/*.cuh) $(wildcard mmtg_b200/csrc/*.h) include/mmtg_b200.h
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; false)

$(LIB): $(OBJ)
	@mkdir -p mmtg_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart_static -ldl -lrt -lpthread

clean:
	rm -rf build $(LIB)
.PHONY: all clean
