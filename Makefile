# Builds the product: libvarscot_scan.so (CUDA kernels + C ABI) and the two drop-in executables at the
# path the reference driver calls them by (VARSCOT_pipeline/VARSCOT:296,307).  sm_100a only.
NVCC    ?= /usr/local/cuda/bin/nvcc
HOSTCXX ?= /usr/bin/g++
ARCH    := -gencode arch=compute_100a,code=sm_100a
EXTRA   ?=
NVFLAGS := $(ARCH) $(EXTRA) -O3 -std=c++17 -lineinfo -ccbin $(HOSTCXX) -Xcompiler -fPIC,-Wall,-Wno-unused-function --cudart static
CSRC    := varscot_b200/csrc
LIB     := varscot_b200/libvarscot_scan.so
BINDIR  := build/read_mapping_build
SRCS    := $(CSRC)/vs_device.cu $(CSRC)/vs_host.cpp $(CSRC)/vs_cli.cpp $(CSRC)/vs_vcf.cpp $(CSRC)/vs_merge.cpp
VPDIR   := build/variant_processing_build
HDRS    := $(CSRC)/vs_kernels.cuh $(CSRC)/vs_extract_block.inc $(CSRC)/vs_bucket.cuh $(CSRC)/vs_internal.h $(CSRC)/vs_genome.h include/varscot_scan.h

all: $(LIB) $(BINDIR)/bidir_mapping $(BINDIR)/bidir_index $(VPDIR)/vcf_loader $(VPDIR)/fasta_writer $(VPDIR)/bam_merger $(VPDIR)/bam_merger_ref_only

$(LIB): $(SRCS) $(HDRS)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(SRCS) -lpthread

$(BINDIR)/bidir_mapping: $(CSRC)/bidir_mapping_main.cpp $(LIB)
	@mkdir -p $(BINDIR)
	$(HOSTCXX) -O2 -o $@ $< -Lvarscot_b200 -lvarscot_scan -Wl,-rpath,'$$ORIGIN/../../varscot_b200' -lpthread -ldl -lrt

$(BINDIR)/bidir_index: $(CSRC)/bidir_index_main.cpp $(LIB)
	@mkdir -p $(BINDIR)
	$(HOSTCXX) -O2 -o $@ $< -Lvarscot_b200 -lvarscot_scan -Wl,-rpath,'$$ORIGIN/../../varscot_b200' -lpthread -ldl -lrt

$(VPDIR)/vcf_loader: $(CSRC)/vcf_loader_main.cpp $(LIB)
	@mkdir -p $(VPDIR)
	$(HOSTCXX) -O2 -o $@ $< -Lvarscot_b200 -lvarscot_scan -Wl,-rpath,'$$ORIGIN/../../varscot_b200' -lpthread -ldl -lrt

$(VPDIR)/fasta_writer: $(CSRC)/fasta_writer_main.cpp $(LIB)
	@mkdir -p $(VPDIR)
	$(HOSTCXX) -O2 -o $@ $< -Lvarscot_b200 -lvarscot_scan -Wl,-rpath,'$$ORIGIN/../../varscot_b200' -lpthread -ldl -lrt

$(VPDIR)/bam_merger: $(CSRC)/bam_merger_main.cpp $(LIB)
	@mkdir -p $(VPDIR)
	$(HOSTCXX) -O2 -o $@ $< -Lvarscot_b200 -lvarscot_scan -Wl,-rpath,'$$ORIGIN/../../varscot_b200' -lpthread -ldl -lrt

$(VPDIR)/bam_merger_ref_only: $(CSRC)/bam_merger_ref_only_main.cpp $(LIB)
	@mkdir -p $(VPDIR)
	$(HOSTCXX) -O2 -o $@ $< -Lvarscot_b200 -lvarscot_scan -Wl,-rpath,'$$ORIGIN/../../varscot_b200' -lpthread -ldl -lrt

oracle:
	$(MAKE) -C oracle all

clean:
	rm -f $(LIB) $(BINDIR)/bidir_mapping $(BINDIR)/bidir_index $(VPDIR)/vcf_loader $(VPDIR)/fasta_writer $(VPDIR)/bam_merger $(VPDIR)/bam_merger_ref_only
	$(MAKE) -C oracle clean

.PHONY: all oracle clean
