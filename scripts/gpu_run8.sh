mkdir -p gpurun_out
rm -f gpurun_out/summary8.txt
timeout 900 python -m pytest tests/test_gpu_facade.py tests/test_gpu_parity.py -m gpu -q --tb=short -x > gpurun_out/tests8.log 2>&1; echo "tests exit $?" >> gpurun_out/summary8.txt
for ch in 0 8 16 32 64; do
  DEER_VIT_CHUNK=$ch timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --calib-steps 4 > gpurun_out/bench8_c$ch.json 2> gpurun_out/bench8_c$ch.err
  python -c "
import json; d=json.load(open('gpurun_out/bench8_c$ch.json')); print('chunk $ch', round(d['value'],1), 'steps/s', round(d['ms_per_step'],2), 'ms', 'gemm frac', round(d['roofline']['frac'],3), {k:round(v['tflops']) for k,v in d['roofline']['stages'].items()})" >> gpurun_out/summary8.txt
done
cat gpurun_out/summary8.txt; tail -3 gpurun_out/tests8.log
