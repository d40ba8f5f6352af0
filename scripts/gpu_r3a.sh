#!/bin/bash
# snake row order (LayerNorm rows / ViT attention items from the end): bit-identity test, A/B bench, DRAM bytes with caches left alone
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ops.py -x -q -m gpu -k "snake or attention or layernorm or batch_invariance" 2>&1 | tail -3
for sn in 1 0 1 0; do
  DEER_SNAKE=$sn timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r3a_bench_snake${sn}.json 2> gpurun_out/r3a_bench_snake${sn}.err; echo "bench snake=$sn exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/r3a_bench_snake${sn}.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'], d['roofline']['frac'])"
done
for sn in 1 0; do
  DEER_SNAKE=$sn timeout 600 python bench.py --batch 28 --no-cpu-baseline > gpurun_out/r3a_bench28_snake${sn}.json 2> gpurun_out/r3a_bench28_snake${sn}.err; echo "bench28 snake=$sn exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/r3a_bench28_snake${sn}.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
done
for sn in 1 0; do
  DEER_SNAKE=$sn DEER_NCU_RANGE=1 timeout 900 ncu --profile-from-start off --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file gpurun_out/r3a_dram_snake${sn}.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --calib-windows 1 > gpurun_out/r3a_ncu_snake${sn}.log 2>&1; echo "ncu snake=$sn exit $?"
done
python - <<'PY'
import csv, collections
for sn in (1, 0):
    rows = list(csv.reader(open(f'gpurun_out/r3a_dram_snake{sn}.csv')))
    hdr = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    h = rows[hdr]; ki, mi, vi = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value')
    ui = h.index('Metric Unit')
    agg = collections.defaultdict(lambda: collections.defaultdict(float)); cnt = collections.Counter()
    for r in rows[hdr + 1:]:
        if len(r) <= vi: continue
        name = r[ki].split('(')[0].replace('void ', '')[:44]
        v = float(r[vi].replace(',', ''))
        u = r[ui]
        if 'byte' in u.lower():
            v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
        if 'second' in u.lower() or u in ('ns', 'us', 'ms'):
            v *= {'ns': 1e-3, 'us': 1, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1, 'msecond': 1e3}.get(u, 1)
        agg[name][r[mi]] += v
        if r[mi] == 'gpu__time_duration.sum': cnt[name] += 1
    print(f'== DEER_SNAKE={sn}: per step, caches not flushed between kernels')
    tot = collections.defaultdict(float)
    for name, m in sorted(agg.items(), key=lambda kv: -kv[1]['gpu__time_duration.sum'])[:8]:
        print(f"{name:46s} n={cnt[name]:4d} time {m['gpu__time_duration.sum']/1e3:8.3f} ms  dram read {m['dram__bytes_read.sum']/1e9:7.3f} GB  write {m['dram__bytes_write.sum']/1e9:7.3f} GB")
    for name, m in agg.items():
        for k, v in m.items(): tot[k] += v
    print(f"{'TOTAL':46s}        time {tot['gpu__time_duration.sum']/1e3:8.3f} ms  dram read {tot['dram__bytes_read.sum']/1e9:7.3f} GB  write {tot['dram__bytes_write.sum']/1e9:7.3f} GB")
PY
