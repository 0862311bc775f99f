# Builds the product library (CUDA, sm_100a) and the test-only SIMT-emulation library.
NVCC      ?= nvcc
CXX       ?= g++
CSRC      := ndrustfft_b200/csrc
SRCS      := $(CSRC)/ndfft_b200.cu $(sort $(wildcard $(CSRC)/sfft_inst_*.cu) $(wildcard $(CSRC)/rsfft_inst_*.cu) $(wildcard $(CSRC)/bsfft_inst_*.cu) $(wildcard $(CSRC)/fs2_inst*.cu) $(wildcard $(CSRC)/sfft_inst_bulk*.cu))
HDRS      := $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh $(CSRC)/*.inc) include/ndfft_b200.h
# the generated instantiation files only see the kernel headers: host-side edits do not rebuild ~900 kernel instances
KHDRS     := $(CSRC)/common.h $(CSRC)/devapi.h $(CSRC)/butterflies.cuh $(CSRC)/sfft_kernel.cuh $(CSRC)/sfft_inst.h $(CSRC)/rsfft_tables.inc include/ndfft_b200.h
LIBDIR    := ndrustfft_b200/lib
LIB       := $(LIBDIR)/libndfft_b200.so
EMULIB    := tests/emu/libndfft_b200_emu.so
NVFLAGS   := -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-fvisibility=hidden \
             --expt-relaxed-constexpr -cudart shared $(EXTRA)
CXXFLAGS  := -O2 -g -std=c++17 -fPIC -DNDFB_EMU -Itests/emu
OBJDIR    := build
NVOBJS    := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/cuda/%.o,$(SRCS))
EMUOBJS   := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/emu/%.o,$(SRCS))

all: $(LIB) $(EMULIB) probes

# kernel headers embedded as strings for run-time compilation of further schedules (csrc/jit.h)
$(CSRC)/jit_sources.inc: $(CSRC)/common.h $(CSRC)/butterflies.cuh $(CSRC)/sfft_kernel.cuh tools/embed_src.py
	python3 tools/embed_src.py

lib: $(LIB)
emu: $(EMULIB)

$(OBJDIR)/cuda/ndfft_b200.o: $(CSRC)/ndfft_b200.cu $(HDRS) $(CSRC)/jit_sources.inc
	@mkdir -p $(OBJDIR)/cuda
	$(NVCC) $(NVFLAGS) -c -o $@ $<

# the software-pipelined column kernels have headers of their own (not in KHDRS: editing them rebuilds one file)
$(OBJDIR)/cuda/sfft_inst_pipe.o $(OBJDIR)/emu/sfft_inst_pipe.o: $(CSRC)/pipe_kernel.cuh $(CSRC)/pipe_inst.h

$(OBJDIR)/cuda/%.o: $(CSRC)/%.cu $(KHDRS)
	@mkdir -p $(OBJDIR)/cuda
	$(NVCC) $(NVFLAGS) -c -o $@ $<

$(OBJDIR)/emu/ndfft_b200.o: $(CSRC)/ndfft_b200.cu $(HDRS) tests/emu/simt_emu.h
	@mkdir -p $(OBJDIR)/emu
	$(CXX) $(CXXFLAGS) -x c++ -c -o $@ $<

$(OBJDIR)/emu/%.o: $(CSRC)/%.cu $(KHDRS) tests/emu/simt_emu.h
	@mkdir -p $(OBJDIR)/emu
	$(CXX) $(CXXFLAGS) -x c++ -c -o $@ $<

$(LIB): $(NVOBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) -shared -cudart shared -gencode arch=compute_100a,code=sm_100a -o $@ $(NVOBJS) -ldl

$(EMULIB): $(EMUOBJS)
	$(CXX) -shared -o $@ $(EMUOBJS)

# stand-alone measurement probes (tools/probes/*.cu -> binaries next to their sources)
PROBES    := $(patsubst %.cu,%,$(wildcard tools/probes/*.cu))
probes: $(PROBES)
tools/probes/%: tools/probes/%.cu
	$(NVCC) -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -o $@ $<

clean:
	rm -rf $(LIB) $(EMULIB) $(OBJDIR) $(PROBES)

.PHONY: all lib emu probes clean
