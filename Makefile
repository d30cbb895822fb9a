# Builds libpcnn.so (hand-written sm_100a CUDA kernels + C ABI) in-tree.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr
SRC := $(wildcard poisson_cnn_b200/csrc/*.cu)
OBJ := $(patsubst poisson_cnn_b200/csrc/%.cu,build/%.o,$(SRC))
LIB := poisson_cnn_b200/libpcnn.so

CHOST := build/pcnn_host

all: $(LIB) $(CHOST)

build/%.o: poisson_cnn_b200/csrc/%.cu poisson_cnn_b200/csrc/pcnn_common.cuh include/pcnn.h
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -Xptxas -v -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

build/engine.o: poisson_cnn_b200/csrc/engine_json.h poisson_cnn_b200/csrc/smallmap_stack.h
build/smallmap_stack.o: poisson_cnn_b200/csrc/smallmap_stack.h

$(LIB): $(OBJ)
	$(NVCC) -shared $(ARCH) -o $@ $(OBJ) -lcudart

# a host in plain C for the model-level ABI (gcc + libcudart only): examples/c_host/pcnn_host.c
$(CHOST): examples/c_host/pcnn_host.c include/pcnn.h $(LIB)
	gcc -O2 -Wall -Iinclude -I/usr/local/cuda/include $< -o $@ -Lpoisson_cnn_b200 -lpcnn -L/usr/local/cuda/lib64 -lcudart \
	    -Wl,-rpath,'$$ORIGIN/../poisson_cnn_b200' -Wl,-rpath,/usr/local/cuda/lib64

clean:
	rm -rf build $(LIB)
.PHONY: all clean
