# Builds libtinyedm_b200.so (sm_100a only) in-tree; the oracle needs no compilation (torch CPU).
NVCC ?= nvcc
NVCCFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --use_fast_math -Iinclude
SRC := $(wildcard tinyedm_b200/csrc/*.cu)
OBJ := $(patsubst tinyedm_b200/csrc/%.cu,build/%.o,$(SRC))
LIB := tinyedm_b200/libtinyedm_b200.so

all: $(LIB)

build/%.o: tinyedm_b200/csrc/%.cu tinyedm_b200/csrc/common.cuh tinyedm_b200/csrc/kernels.h include/tinyedm_b200.h
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -shared -o $@ $(OBJ) -cudart static

clean:
	rm -rf build $(LIB)
.PHONY: all clean
