# Builds the in-tree shared libraries (sm_100a only; nvcc cross-compiles without a GPU).
#   laghos_b200/lib/liblaghos_b200.so   C ABI (include/laghos_b200.h) + C++ shim (lagb_laghos_run)
#   oracle/_build/liboracle.so          CPU oracle (test infrastructure)
NVCC ?= nvcc
CXX ?= g++
ARCH = -gencode arch=compute_100a,code=sm_100a
NVFLAGS = $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall -diag-suppress 128
CXXFLAGS = -O2 -std=c++17 -fPIC -Wall
CS = laghos_b200/csrc
OBJ = build/capi.o build/kernels_generic.o build/kernels_tuned.o build/kernels_l2.o build/problem_capi.o build/laghos_shim.o
DEV_HDRS = $(wildcard $(CS)/device/*.cuh) $(CS)/ctx.hpp include/laghos_b200.h
HOST_HDRS = $(wildcard $(CS)/host/*.hpp) include/laghos_b200.h

all: laghos_b200/lib/liblaghos_b200.so oracle/_build/liboracle.so

build/%.o: $(CS)/%.cu $(DEV_HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@
build/problem_capi.o: $(CS)/problem_capi.cpp $(HOST_HDRS)
	@mkdir -p build
	$(CXX) $(CXXFLAGS) -c $< -o $@
build/laghos_shim.o: laghos_b200/shim/laghos_shim.cpp laghos_b200/shim/laghos_shim.hpp $(HOST_HDRS)
	@mkdir -p build
	$(CXX) $(CXXFLAGS) -c $< -o $@
laghos_b200/lib/liblaghos_b200.so: $(OBJ)
	@mkdir -p laghos_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart -ldl
oracle/_build/liboracle.so: oracle/oracle_capi.cpp oracle/laghos_oracle.hpp oracle/own_tables.hpp oracle/smallmat.hpp $(HOST_HDRS)
	@mkdir -p oracle/_build
	$(CXX) -O3 -march=x86-64-v3 -std=c++17 -fPIC -shared -o $@ oracle/oracle_capi.cpp -lpthread
oracle/_build/oracle_cli: oracle/oracle_main.cpp oracle/laghos_oracle.hpp oracle/own_tables.hpp oracle/smallmat.hpp $(HOST_HDRS)
	@mkdir -p oracle/_build
	$(CXX) -O3 -march=x86-64-v3 -std=c++17 -o $@ oracle/oracle_main.cpp -lpthread

# the reference's own Sedov exact solution (self-contained: sedov_sol.cpp + two headers, no MFEM), compiled from the
# sources where they lie; only where /root/reference exists (this container), output into oracle/_ref/ (git-ignored)
REF_SEDOV = /root/reference/sedov
oracle/_ref/libsedov_ref.so: oracle/sedov_ref_capi.cpp
	@mkdir -p oracle/_ref
	$(CXX) -O2 -std=c++17 -fPIC -shared -I$(REF_SEDOV) -o $@ oracle/sedov_ref_capi.cpp $(REF_SEDOV)/sedov_sol.cpp
ref: $(if $(wildcard $(REF_SEDOV)/sedov_sol.cpp),oracle/_ref/libsedov_ref.so)

clean:
	rm -rf build laghos_b200/lib oracle/_build
.PHONY: all clean ref
