# Build recipe for libsmgpu.so (product), the smoothMesh CLI and the CPU oracle.
# __graft_entry__.build() runs `make all`.
NVCC      ?= nvcc
CXX       := $(shell test -x /usr/bin/g++ && echo /usr/bin/g++ || echo g++)
ARCH      := -gencode arch=compute_100a,code=sm_100a
# --fmad=false: FP64 results must match a scalar IEEE evaluation bit for bit
NVFLAGS   := -ccbin $(CXX) $(ARCH) -O3 -lineinfo --fmad=false -std=c++17 -Xcompiler -fPIC,-fopenmp,-ffp-contract=off,-Wall
CXXFLAGS  := -O3 -std=c++17 -fPIC -fopenmp -ffp-contract=off -Wall
CSRC      := smoothmesh_b200/csrc
LIBDIR    := smoothmesh_b200/lib
BINDIR    := smoothmesh_b200/bin
OBJDIR    := build

LIB       := $(LIBDIR)/libsmgpu.so
CLI       := $(BINDIR)/smoothMesh
ORACLE    := oracle/_build/liboracle.so
ORACLE_LM := oracle/_build/liboracle_libm.so

# oracle/_ref: the reference's own translation unit compiled against the OpenFOAM facade (oracle/Makefile.ref);
# only where the reference sources exist (the build container), the GPU box receives the built binary
REFSRC = /root/reference/src/smoothMesh.C
ifneq ($(wildcard $(REFSRC)),)
REFBIN = oracle/_ref/smoothMesh_ref
endif

all: $(LIB) $(CLI) $(ORACLE) $(ORACLE_LM) $(REFBIN)

oracle/_ref/smoothMesh_ref: oracle/of_facade/ref_main.cpp oracle/of_facade/of_support.cpp $(wildcard oracle/of_facade/*.H) oracle/oracle.cpp $(OBJDIR)/polymesh.o
	$(MAKE) -f oracle/Makefile.ref

$(OBJDIR)/%.o: $(CSRC)/%.cpp $(wildcard $(CSRC)/*.hpp) $(wildcard $(CSRC)/*.h) $(wildcard include/*.h)
	@mkdir -p $(OBJDIR)
	$(CXX) $(CXXFLAGS) -c $< -o $@

$(OBJDIR)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.hpp) $(wildcard $(CSRC)/*.h) $(wildcard $(CSRC)/*.cuh) $(wildcard include/*.h)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; false)

LIBOBJS := $(OBJDIR)/smgpu.o $(OBJDIR)/exchange.o $(OBJDIR)/polymesh.o $(OBJDIR)/topology.o $(OBJDIR)/boundary.o $(OBJDIR)/smmesh_api.o

$(LIB): $(LIBOBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) -ccbin $(CXX) $(ARCH) -shared -o $@ $(LIBOBJS) -Xcompiler -fopenmp -ldl

$(CLI): $(CSRC)/smoothmesh_cli.cpp $(LIB)
	@mkdir -p $(BINDIR)
	$(CXX) $(CXXFLAGS) -o $@ $< -L$(LIBDIR) -lsmgpu -Wl,-rpath,'$$ORIGIN/../lib'

$(ORACLE): oracle/oracle.cpp $(CSRC)/sm_math.h
	@mkdir -p oracle/_build
	$(CXX) $(CXXFLAGS) -shared -o $@ oracle/oracle.cpp

$(ORACLE_LM): oracle/oracle.cpp $(CSRC)/sm_math.h
	@mkdir -p oracle/_build
	$(CXX) $(CXXFLAGS) -DORACLE_LIBM_ACOS -shared -o $@ oracle/oracle.cpp

clean:
	rm -rf $(OBJDIR) $(LIBDIR) $(BINDIR) oracle/_build

.PHONY: all clean

# host set-up timing / fingerprint tool (not part of `all`)
build/setup_timing: tools/setup_timing.cpp $(OBJDIR)/topology.o $(OBJDIR)/polymesh.o
	$(CXX) $(CXXFLAGS) -I$(CSRC) -o $@ $^
