# Build of libkatgpu (sm_100a CUDA + C-ABI) and the `kat` front-end.
#   make            -> kat_b200/lib/libkatgpu.so, kat_b200/bin/kat
#   make oracle     -> oracle/libkatoracle.so (+ oracle/_ref when /root/reference exists)
NVCC      ?= nvcc
CXX       ?= g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v
CSRC      := kat_b200/csrc
OBJDIR    := build
LIB       := kat_b200/lib/libkatgpu.so
KAT       := kat_b200/bin/kat

CU_SRC  := $(CSRC)/count_kernels.cu $(CSRC)/partition_kernels.cu $(CSRC)/tile_kernels.cu $(CSRC)/dist_kernels.cu $(CSRC)/parse_kernels.cu $(CSRC)/reduce_kernels.cu $(CSRC)/katgpu_api.cu $(CSRC)/katgpu_fastq.cu $(CSRC)/katgpu_walks.cu $(CSRC)/katgpu_dist.cu $(CSRC)/wide_kernels.cu $(CSRC)/katgpu_wide.cu
CU_OBJ  := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/%.o,$(CU_SRC))
HOST_SRC := $(wildcard $(CSRC)/host/*.cc)
HOST_OBJ := $(patsubst $(CSRC)/host/%.cc,$(OBJDIR)/host_%.o,$(HOST_SRC))

ifneq ($(strip $(HOST_SRC)),)
all: $(LIB) $(KAT)
else
all: $(LIB)
endif

$(OBJDIR)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) include/katgpu.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; false)

$(LIB): $(CU_OBJ)
	@mkdir -p $(dir $@)
	$(NVCC) $(ARCH) -shared -o $@ $^ -ldl

$(OBJDIR)/host_%.o: $(CSRC)/host/%.cc $(wildcard $(CSRC)/host/*.hpp) include/katgpu.h
	@mkdir -p $(OBJDIR)
	$(CXX) -O2 -std=c++17 -Wall -Wextra -Iinclude -c $< -o $@

$(KAT): $(HOST_OBJ) $(LIB)
	@mkdir -p $(dir $@)
	$(CXX) -o $@ $(HOST_OBJ) -Lkat_b200/lib -lkatgpu -lz -lpthread -Wl,-rpath,'$$ORIGIN/../lib'

oracle:
	$(MAKE) -C oracle all

clean:
	rm -rf $(OBJDIR) $(LIB) $(KAT)
.PHONY: all oracle clean
