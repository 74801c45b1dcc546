# Top-level build: the CUDA library (product) and the oracle (checker).
PKG := block2-preview_b200
NVCC ?= /usr/local/cuda/bin/nvcc
NVCCFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v
SRCS := $(wildcard $(PKG)/csrc/*.cu)
HDRS := $(wildcard $(PKG)/csrc/*.h $(PKG)/csrc/*.cuh) include/b2g.h

all: lib oracle

lib: $(PKG)/libb2g.so tools/fp64_probe

$(PKG)/libb2g.so: $(SRCS) $(HDRS)
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(SRCS) -ldl 2> $(PKG)/csrc/ptxas.log || (cat $(PKG)/csrc/ptxas.log; false)

tools/fp64_probe: tools/fp64_probe.cu
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o $@ $<

oracle:
	$(MAKE) -C oracle _ref/liboracle.so
	if [ -d /root/reference/src ]; then $(MAKE) -C oracle -j2 ref; fi

clean:
	rm -f $(PKG)/libb2g.so tools/fp64_probe
.PHONY: all lib oracle clean
