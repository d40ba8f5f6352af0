#!/bin/bash
# round 2, pass m: attention v6 (two-pass softmax, split PV): op tests, micro-benchmark + phase cycles, parity suite, step bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "attention or attn" > gpurun_out/r2m_attn_tests.log 2>&1; echo "attn op tests exit $?"; tail -3 gpurun_out/r2m_attn_tests.log
timeout 120 python scripts/attn_bench.py > gpurun_out/r2m_attn_bench.log 2>&1; echo "attn bench exit $?"; cat gpurun_out/r2m_attn_bench.log | tail -3
(cd deer_vla_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -DDEER_ATTN_PROF engine.cu -o /tmp/libdeer_prof.so) && DEER_LIB=/tmp/libdeer_prof.so timeout 300 python scripts/attn_bench.py > gpurun_out/r2m_attn_phases.log 2>&1; echo "attn phases exit $?"; tail -2 gpurun_out/r2m_attn_phases.log
timeout 1500 python -m pytest tests -m gpu -q --tb=short -s > gpurun_out/r2m_tests_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -2 gpurun_out/r2m_tests_gpu.log
grep -h "^\[full3b" gpurun_out/r2m_tests_gpu.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_bench_224.json 2> gpurun_out/r2m_bench_224.err; echo "bench224 exit $?"; cut -c1-200 gpurun_out/r2m_bench_224.json; tail -2 gpurun_out/r2m_bench_224.err
