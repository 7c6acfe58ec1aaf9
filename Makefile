# Top-level build: the product library (C ABI), the CPU oracle, and the bring-up tool.
NVCC  ?= nvcc
ARCH  := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo
PKG   := flash_attention_impls_b200
CSRC  := $(PKG)/csrc/fa_api.cu $(PKG)/csrc/fa_merge.cu $(PKG)/csrc/fa_host.cu $(PKG)/csrc/fa_ring.cu $(PKG)/csrc/fa_ref_shims.cu
HDRS  := $(wildcard $(PKG)/csrc/*.cuh) include/fa_b200.h

all: lib oracle tools

lib: $(PKG)/lib/libfa_b200.so
$(PKG)/lib/libfa_b200.so: $(CSRC) $(HDRS)
	mkdir -p $(PKG)/lib
	$(NVCC) $(NVFLAGS) -shared -Xcompiler -fPIC -o $@ $(CSRC)

oracle:
	$(MAKE) -C oracle liboracle.so ref

tools: tools/fa_selftest
tools/fa_selftest: tools/fa_selftest.cu $(PKG)/lib/libfa_b200.so oracle
	$(NVCC) $(NVFLAGS) -o $@ tools/fa_selftest.cu -L$(PKG)/lib -lfa_b200 -Loracle -loracle -ldl \
	  -Xlinker -rpath -Xlinker '$$ORIGIN/../$(PKG)/lib' -Xlinker -rpath -Xlinker '$$ORIGIN/../oracle'

sass: $(PKG)/lib/libfa_b200.so
	cuobjdump -sass $< > /tmp/libfa_b200.sass

clean:
	rm -f $(PKG)/lib/*.so tools/fa_selftest; $(MAKE) -C oracle clean

.PHONY: all lib oracle tools sass clean

# library variants for interleaved A/B timing (tools/gpu_ab.sh): make variant NAME=p3 DEFS=-DFA_P_FIRST_Q=3
variant:
	mkdir -p build/$(NAME)
	$(NVCC) $(NVFLAGS) $(DEFS) -shared -Xcompiler -fPIC -o build/$(NAME)/libfa_b200.so $(CSRC)
.PHONY: variant
