# Builds libomc_b200.so (sm_100a only) in-tree so it travels to the GPU box with the snapshot.
NVCC ?= nvcc
PKG := optimalmatrixcompletion.jl_b200
HDR := $(wildcard $(PKG)/csrc/*.cuh) $(wildcard $(PKG)/csrc/*.h) include/omc_b200.h
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $(EXTRA)
OBJ := $(PKG)/csrc/omc_api.o $(PKG)/csrc/omc_big.o $(PKG)/csrc/omc_comm.o

$(PKG)/libomc_b200.so: $(OBJ)
	$(NVCC) $(NVFLAGS) --shared -o $@ $(OBJ) -ldl

# omc_api.cu holds the C ABI and the round-1 kernels; omc_big.cu the batched large-block engine (separate translation unit)
$(PKG)/csrc/omc_api.o: $(PKG)/csrc/omc_api.cu $(filter-out %omc_big.cuh,$(HDR))
	$(NVCC) $(NVFLAGS) -c -o $@ $<

$(PKG)/csrc/omc_comm.o: $(PKG)/csrc/omc_comm.cu include/omc_b200.h
	$(NVCC) $(NVFLAGS) -c -o $@ $<

$(PKG)/csrc/omc_big.o: $(PKG)/csrc/omc_big.cu $(PKG)/csrc/omc_big.cuh $(PKG)/csrc/omc_big_host.h include/omc_b200.h
	$(NVCC) $(NVFLAGS) -c -o $@ $<

clean:
	rm -f $(PKG)/libomc_b200.so $(OBJ)
.PHONY: clean
