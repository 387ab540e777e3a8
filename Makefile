# Builds the C-ABI shared library (sm_100a only).  Objects are compiled separately so `make -j` parallelises.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr
# make TIMELINE=1: compile the in-kernel clock64 timeline of the fused forward kernel in (development only)
ifdef TIMELINE
NVCCFLAGS += -DNVP_TIMELINE
endif
SRC       := $(wildcard nvp_b200/csrc/*.cu)
HDR       := $(wildcard nvp_b200/csrc/*.cuh) include/nvp_b200.h
OBJDIR    := build/obj
OBJ       := $(patsubst nvp_b200/csrc/%.cu,$(OBJDIR)/%.o,$(SRC))
LIB       := nvp_b200/libnvp_b200.so

all: $(LIB)

$(OBJDIR)/%.o: nvp_b200/csrc/%.cu $(HDR)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared $(OBJ) -o $@ -lcuda

ptxas-info: $(SRC) $(HDR)
	$(NVCC) $(NVCCFLAGS) -Xptxas -v -shared $(SRC) -o /tmp/nvp_ptxas_check.so -lcuda

clean:
	rm -rf $(LIB) $(OBJDIR)
.PHONY: all clean ptxas-info
