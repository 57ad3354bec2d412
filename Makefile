# Builds libscone_b200.so (sm_100a) in-tree.  `python __graft_entry__.py build` does the same.
NVCC ?= nvcc
SRC := scone_gcn_b200/csrc/scone_complex.cu scone_gcn_b200/csrc/scone_kernels.cu scone_gcn_b200/csrc/scone_model.cu scone_gcn_b200/csrc/scone_bunch.cu scone_gcn_b200/csrc/scone_slab.cu scone_gcn_b200/csrc/scone_rows.cu scone_gcn_b200/csrc/scone_fused.cu scone_gcn_b200/csrc/scone_plan_table.cu scone_gcn_b200/csrc/scone_umma.cu scone_gcn_b200/csrc/scone_dp.cu
OUT := scone_gcn_b200/libscone_b200.so
FLAGS := -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Iinclude -shared -Xptxas -v

$(OUT): $(SRC) include/scone_b200.h scone_gcn_b200/csrc/common.cuh scone_gcn_b200/csrc/slab_common.cuh scone_gcn_b200/csrc/fused.cuh
	$(NVCC) $(FLAGS) -o $@ $(SRC)

clean:
	rm -f $(OUT)
