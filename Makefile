# Build recipes.  Everything is built in-tree (the .so / binaries travel to the GPU box with the
# snapshot; they are git-ignored).
#   make lib      -> polymutt_b200/lib/libpolymutt_b200.so   (CUDA C-ABI, sm_100a)
#   make cli      -> polymutt_b200/bin/polymutt-b200         (drop-in executable, links the C-ABI)
#   make oracle   -> oracle/_build/libpm_oracle.so + oracle/_build/polymutt_oracle_cli (test infra)
#   make ref      -> oracle/_ref/polymutt (the unmodified reference; only where /root/reference exists)
NVCC      ?= /usr/local/cuda/bin/nvcc
HOSTCXX   ?= g++
HOSTCC    ?= gcc
CUDA_ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(PM_DEFS) -O3 -std=c++17 -lineinfo $(CUDA_ARCH) -Xcompiler -fPIC -Xcompiler -fno-strict-aliasing -Iinclude -Ipolymutt_b200/csrc
CXXFLAGS  := -O2 -std=c++17 -fPIC -pthread -Wall -Wextra -Wno-unused-parameter -Iinclude -Ipolymutt_b200/csrc/host
# the oracle keeps the reference's arithmetic: no FMA contraction, no fast-math
OCFLAGS   := -O2 -std=gnu99 -fPIC -ffp-contract=off -Wall -Iinclude

HOST_SRC  := $(wildcard polymutt_b200/csrc/host/*.cpp)
HOST_LIB_SRC := $(filter-out polymutt_b200/csrc/host/main.cpp polymutt_b200/csrc/host/driver.cpp polymutt_b200/csrc/host/params.cpp polymutt_b200/csrc/host/vcf_writer.cpp polymutt_b200/csrc/host/vcf_mode.cpp polymutt_b200/csrc/host/glf.cpp polymutt_b200/csrc/host/glf_ingest.cpp polymutt_b200/csrc/host/glf_ingest.cpp polymutt_b200/csrc/host/pedigree.cpp,$(HOST_SRC))
FRONT_SRC := polymutt_b200/csrc/host/driver.cpp polymutt_b200/csrc/host/params.cpp polymutt_b200/csrc/host/vcf_writer.cpp polymutt_b200/csrc/host/vcf_mode.cpp polymutt_b200/csrc/host/glf.cpp polymutt_b200/csrc/host/glf_ingest.cpp polymutt_b200/csrc/host/pedigree.cpp
CU_SRC    := $(wildcard polymutt_b200/csrc/*.cu)
CU_HDR    := $(wildcard polymutt_b200/csrc/*.cuh) $(wildcard polymutt_b200/csrc/*.h) include/polymutt_b200.h

LIB ?= polymutt_b200/lib/libpolymutt_b200.so
CLI := polymutt_b200/bin/polymutt-b200
ORACLE_LIB := oracle/_build/libpm_oracle.so
ORACLE_CLI := oracle/_build/polymutt_oracle_cli
TOOLS := polymutt_b200/bin/pm-tools

.PHONY: all lib cli tools oracle ref clean timing
all: lib cli tools oracle

lib: $(LIB)
# One object per translation unit (so that `make -j` compiles them side by side); pm_post.cu (genotype posteriors:
# exact ties must break as in the reference) is compiled without FMA contraction.  The ptxas summaries
# (registers, spills) of every kernel land next to the objects.
OBJDIR ?= polymutt_b200/lib/obj
CU_OBJ := $(patsubst polymutt_b200/csrc/%.cu,$(OBJDIR)/%.o,$(CU_SRC))
HOSTLIB_OBJ := $(patsubst polymutt_b200/csrc/host/%.cpp,$(OBJDIR)/host_%.o,$(HOST_LIB_SRC))
$(OBJDIR)/pm_post.o: NVEXTRA := -fmad=false
$(OBJDIR)/%.o: polymutt_b200/csrc/%.cu $(CU_HDR) polymutt_b200/csrc/host/host_error.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) $(NVEXTRA) -Ipolymutt_b200/csrc/host -Xptxas -v -c -o $@ $< 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; exit 1)
$(OBJDIR)/host_%.o: polymutt_b200/csrc/host/%.cpp $(wildcard polymutt_b200/csrc/host/*.h) include/polymutt_b200.h
	@mkdir -p $(OBJDIR)
	$(HOSTCXX) $(CXXFLAGS) -c -o $@ $<
$(LIB): $(CU_OBJ) $(HOSTLIB_OBJ)
	$(NVCC) $(CUDA_ARCH) -shared -o $@ $(CU_OBJ) $(HOSTLIB_OBJ) -lcudart
	@cat $(OBJDIR)/*.ptxas.log > $(OBJDIR)/ptxas_all.log

# the same library with per-phase cycle counters in the wide kernel (scripts/gpu_phase_timing.py; never the product)
timing:
	$(MAKE) lib PM_DEFS=-DPM_PHASE_TIMING OBJDIR=polymutt_b200/lib/obj_timing LIB=polymutt_b200/lib/libpolymutt_b200_timing.so

tools: $(TOOLS)
$(TOOLS): polymutt_b200/csrc/tools/pm_tools.cpp $(FRONT_SRC) $(HOST_LIB_SRC) $(wildcard polymutt_b200/csrc/host/*.h)
	@mkdir -p polymutt_b200/bin
	$(HOSTCXX) $(CXXFLAGS) -o $@ polymutt_b200/csrc/tools/pm_tools.cpp polymutt_b200/csrc/host/glf.cpp polymutt_b200/csrc/host/glf_ingest.cpp polymutt_b200/csrc/host/pedigree.cpp polymutt_b200/csrc/host/vcf_writer.cpp polymutt_b200/csrc/host/params.cpp $(HOST_LIB_SRC) -lz

cli: $(CLI)
$(CLI): $(LIB) $(FRONT_SRC) polymutt_b200/csrc/host/main.cpp $(wildcard polymutt_b200/csrc/host/*.h)
	@mkdir -p polymutt_b200/bin
	$(HOSTCXX) $(CXXFLAGS) -o $@ polymutt_b200/csrc/host/main.cpp $(FRONT_SRC) -Lpolymutt_b200/lib -lpolymutt_b200 -lz -Wl,-rpath,'$$ORIGIN/../lib'

oracle: $(ORACLE_LIB) $(ORACLE_CLI)
$(ORACLE_LIB): oracle/pm_oracle.c oracle/pm_oracle.h include/polymutt_b200.h
	@mkdir -p oracle/_build
	$(HOSTCC) $(OCFLAGS) -shared -o $@ oracle/pm_oracle.c -lm
$(ORACLE_CLI): oracle/oracle_cli.cpp oracle/pm_oracle.c $(FRONT_SRC) $(HOST_LIB_SRC) $(wildcard polymutt_b200/csrc/host/*.h)
	@mkdir -p oracle/_build
	$(HOSTCC) $(OCFLAGS) -c -o oracle/_build/pm_oracle.o oracle/pm_oracle.c
	$(HOSTCXX) $(CXXFLAGS) -o $@ oracle/oracle_cli.cpp oracle/_build/pm_oracle.o $(FRONT_SRC) $(HOST_LIB_SRC) -lz -lm

ref:
	oracle/build_ref.sh

clean:
	rm -rf polymutt_b200/lib polymutt_b200/bin oracle/_build
