# Builds the C-ABI shared library (sm_100a only) and the oracle's compiled pieces.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr
SRC       := $(wildcard nvp_b200/csrc/*.cu)
HDR       := $(wildcard nvp_b200/csrc/*.cuh) include/nvp_b200.h
LIB       := nvp_b200/libnvp_b200.so

all: $(LIB)

$(LIB): $(SRC) $(HDR)
	$(NVCC) $(NVCCFLAGS) -shared $(SRC) -o $@ -lcuda

ptxas-info: $(SRC) $(HDR)
	$(NVCC) $(NVCCFLAGS) -Xptxas -v -shared $(SRC) -o /tmp/nvp_ptxas_check.so -lcuda

clean:
	rm -f $(LIB)
.PHONY: all clean ptxas-info
