mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -x -k "simt" > gpurun_out/ops_simt.log 2>&1; echo "ops_simt exit $?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -k "tcgen05 or flash" > gpurun_out/ops_tc.log 2>&1; echo "ops_tc exit $?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "tiny or errors" > gpurun_out/parity_tiny.log 2>&1; echo "parity_tiny exit $?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -s -k "full3b" > gpurun_out/parity_full.log 2>&1; echo "parity_full exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -5 gpurun_out/ops_simt.log gpurun_out/ops_tc.log gpurun_out/parity_tiny.log gpurun_out/parity_full.log
