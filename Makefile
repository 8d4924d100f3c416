# cuppens-b200: build the CUDA library (sm_100a), the `cuppens` CLI and the test oracle.
# Target names and variables of the run recipes follow the reference Makefile
# (/root/reference/Makefile:27-54: all cuppen clean run runo runc runoc rune runec;
#  NUMTASKS DIM OUT SCHEME).
NVCC     ?= /usr/local/cuda/bin/nvcc
CC       ?= gcc
CSRC     := symmetric_eigenvalue_b200/csrc
LIBDIR   := symmetric_eigenvalue_b200/lib
LIB      := $(LIBDIR)/libcuppen_b200.so
SELFTEST := $(LIBDIR)/libcuppen_selftest.so
OBJ_NAME := cuppens
NVFLAGS  := --extended-lambda -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC \
            -Xcompiler -Wall -Xcompiler -Wno-unused-function --expt-relaxed-constexpr
HDRS     := $(wildcard $(CSRC)/*.h) include/cuppen_b200.h

all: cuppen
lib: $(LIB) $(SELFTEST)

$(LIBDIR)/hostio.o: $(CSRC)/hostio.c include/cuppen_b200.h
	@mkdir -p $(LIBDIR)
	$(CC) -O2 -fPIC -Wall -c $< -o $@

$(LIBDIR)/solver.o: $(CSRC)/solver.cu $(HDRS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> $(LIBDIR)/ptxas_solver.log || (cat $(LIBDIR)/ptxas_solver.log; false)

$(LIB): $(LIBDIR)/solver.o $(LIBDIR)/hostio.o
	$(NVCC) -shared -Xlinker -Bsymbolic -o $@ $^ -lcudart -ldl

# test / bench only: kernel self-tests and FP64 yardsticks (not part of the product ABI)
$(SELFTEST): $(CSRC)/selftest.cu $(HDRS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -shared -o $@ $< -lcudart

cuppen: $(LIB) $(CSRC)/cuppens_main.c
	$(CC) -O2 -Wall -Iinclude -o $(OBJ_NAME) $(CSRC)/cuppens_main.c -L$(LIBDIR) -lcuppen_b200 \
	    -Wl,-rpath,'$$ORIGIN/$(LIBDIR)' -lm

oracle:
	$(MAKE) -C oracle all

clean:
	rm -f $(OBJ_NAME) $(LIBDIR)/*.o $(LIBDIR)/*.so $(LIBDIR)/*.log

# This standard parameters will be overriden, if you call the Makefile and assign them as parameters.
# Example call: make run NUMTASKS=8 DIM=100   (NUMTASKS = reference leaves P, GPUS = number of B200s)
NUMTASKS=4
DIM=16
OUT=out.txt
SCHEME=1
GPUS=1

run:
	./$(OBJ_NAME) -p $(NUMTASKS) -g $(GPUS) -s $(SCHEME) -n $(DIM) $(OUT)
runo: run
	cat $(OUT)
runc: cuppen run
runoc: cuppen runo
rune:
	./$(OBJ_NAME) -p $(NUMTASKS) -g $(GPUS) -s $(SCHEME) -n $(DIM) -e $(OUT)
runec: cuppen rune

.PHONY: all lib cuppen oracle clean run runo runc runoc rune runec
