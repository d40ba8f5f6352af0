#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
H=$(python -c "import __graft_entry__ as g; print(g.source_hash())")
(cd deer_vla_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DDEER_GEMM_PROF -DDEER_SRC_HASH=\"$H\" -Xcompiler -fPIC -shared engine.cu -o ../../scripts/libdeer_engine_prof.so) 2>&1 | tail -3
DEER_LIB=scripts/libdeer_engine_prof.so timeout 300 python scripts/gemm_kblock_prof.py 2>&1 | tee gpurun_out/r3d_kblock_prof.txt
