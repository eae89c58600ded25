# Builds libpnpinv.so (sm_100a only) in-tree.  `python -c "import __graft_entry__ as g; g.build()"` calls this.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr -Iinclude
SRC := pnpinversion_b200/csrc
OBJS := $(SRC)/gemm_sm100.o $(SRC)/norm.o $(SRC)/attention.o $(SRC)/attention_tc.o $(SRC)/epilogue.o $(SRC)/engine.o $(SRC)/vae.o $(SRC)/clip.o $(SRC)/probe.o
LIB := pnpinversion_b200/libpnpinv.so

all: $(LIB)

$(SRC)/%.o: $(SRC)/%.cu $(SRC)/pnp_internal.h $(SRC)/pnp_ptx.cuh $(SRC)/pnp_attn.h include/pnpinv.h
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -cudart static

clean:
	rm -f $(OBJS) $(LIB)
