mkdir -p gpurun_out
rm -f gpurun_out/summary15.txt
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -k "attention or pv_probe" > gpurun_out/tests15.log 2>&1; echo "attn tests exit $?" >> gpurun_out/summary15.txt
timeout 120 python scripts/attn_bench.py > gpurun_out/attn_bench15.log 2>&1
cat gpurun_out/summary15.txt; tail -6 gpurun_out/tests15.log; cat gpurun_out/attn_bench15.log
