# Builds the product library (CUDA, sm_100a) and the test-only SIMT-emulation library.
NVCC      ?= nvcc
CXX       ?= g++
CSRC      := ndrustfft_b200/csrc
SRCS      := $(CSRC)/ndfft_b200.cu
HDRS      := $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh) include/ndfft_b200.h
LIBDIR    := ndrustfft_b200/lib
LIB       := $(LIBDIR)/libndfft_b200.so
EMULIB    := tests/emu/libndfft_b200_emu.so
NVFLAGS   := -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-fvisibility=hidden \
             --expt-relaxed-constexpr -shared -cudart shared
CXXFLAGS  := -O2 -g -std=c++17 -fPIC -shared -DNDFB_EMU -Itests/emu -x c++

all: $(LIB) $(EMULIB)

lib: $(LIB)
emu: $(EMULIB)

$(LIB): $(SRCS) $(HDRS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -o $@ $(SRCS)

$(EMULIB): $(SRCS) $(HDRS) tests/emu/simt_emu.h
	$(CXX) $(CXXFLAGS) -o $@ $(SRCS)

clean:
	rm -f $(LIB) $(EMULIB)

.PHONY: all lib emu clean
