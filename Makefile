# Builds libomc_b200.so (sm_100a only) in-tree so it travels to the GPU box with the snapshot.
NVCC ?= nvcc
PKG := optimalmatrixcompletion.jl_b200
SRC := $(PKG)/csrc/omc_api.cu
HDR := $(wildcard $(PKG)/csrc/*.cuh) include/omc_b200.h
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC $(EXTRA)

$(PKG)/libomc_b200.so: $(SRC) $(HDR)
	$(NVCC) $(NVFLAGS) -o $@ $(SRC)

clean:
	rm -f $(PKG)/libomc_b200.so
.PHONY: clean
