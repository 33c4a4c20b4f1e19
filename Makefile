# Builds the CUDA engine + host layer into blama_b200/lib/libblama_b200.so (sm_100a only) and the oracle.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v
CSRC      := blama_b200/csrc
HOST      := blama_b200/host
LIBDIR    := blama_b200/lib
OBJDIR    := build

CU_SRCS   := $(wildcard $(CSRC)/*.cu)
CU_OBJS   := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/%.o,$(CU_SRCS))
HOST_SRCS := $(wildcard $(HOST)/llama/*.cpp) $(wildcard $(HOST)/server/*.cpp) $(wildcard $(HOST)/*.cpp)
HOST_OBJS := $(patsubst $(HOST)/%.cpp,$(OBJDIR)/host/%.o,$(HOST_SRCS))
HDRS      := $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.hpp) $(wildcard $(HOST)/llama/*.hpp) $(wildcard $(HOST)/server/*.hpp) $(wildcard $(HOST)/*.hpp) include/blama_b200.h

BINDIR    := blama_b200/bin

all: $(LIBDIR)/libblama_b200.so $(BINDIR)/blama-server oracle

# the server executable (reference server/code/http/HttpServerMain.cpp main): links the library, configured by BLAMA_* variables
# (the C++ host classes are not exported from the library -- its boundary is the C ABI -- so the executable links their objects)
$(BINDIR)/blama-server: $(HOST)/server/main/HttpServerMain.cpp $(LIBDIR)/libblama_b200.so $(HDRS)
	@mkdir -p $(BINDIR)
	g++ -O2 -std=c++20 -Wall -I$(HOST) -Iinclude $< $(filter-out $(OBJDIR)/host/host_capi.o,$(HOST_OBJS)) -o $@ -L$(LIBDIR) -lblama_b200 -Wl,-rpath,'$$ORIGIN/../lib' -lpthread

$(OBJDIR)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(dir $@)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; exit 1)

$(OBJDIR)/host/%.o: $(HOST)/%.cpp $(HDRS)
	@mkdir -p $(dir $@)
	g++ -O2 -std=c++20 -fPIC -fvisibility=hidden -Wall -I$(HOST) -Iinclude -c $< -o $@

$(LIBDIR)/libblama_b200.so: $(CU_OBJS) $(HOST_OBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared -o $@ $^ -lpthread

oracle:
	$(MAKE) -s -C oracle

clean:
	rm -rf $(OBJDIR) $(LIBDIR)/libblama_b200.so $(BINDIR)/blama-server
.PHONY: all oracle clean
