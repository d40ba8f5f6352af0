mkdir -p gpurun_out
rm -f gpurun_out/summary18.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke18.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary18.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -x -k "graph or 9b" > gpurun_out/tests18.log 2>&1; echo "graph/9b tests exit $?" >> gpurun_out/summary18.txt
timeout 600 python bench.py --steps 20 --warmup 5 --batch 1 --no-cpu-baseline > gpurun_out/bench18_b1.json 2> gpurun_out/bench18_b1.err; echo "b1 exit $?" >> gpurun_out/summary18.txt
timeout 600 python bench.py --steps 20 --warmup 5 --batch 1 --no-cpu-baseline --graph > gpurun_out/bench18_b1g.json 2> gpurun_out/bench18_b1g.err; echo "b1 graph exit $?" >> gpurun_out/summary18.txt
cat gpurun_out/summary18.txt; tail -3 gpurun_out/smoke18.log; tail -12 gpurun_out/tests18.log
for f in b1 b1g; do python -c "
import json; d=json.load(open('gpurun_out/bench18_$f.json')); print('$f', round(d['value'],1), 'steps/s', round(d['ms_per_step'],3), 'ms/step e2e', round(d['e2e']['value'],1), d['exit_histogram'])"; done
tail -3 gpurun_out/bench18_b1g.err
