"""profiles/r02_gemm_traffic.json from the per-launch ncu metrics of the TIMED steps of the bench command itself
(DEER_NCU_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed ... python bench.py --steps 1 --warmup 3, scripts/gpu_r2f.sh):
DRAM bytes per GEMM launch (bench.py reports it as roofline.traffic, labelled static) and the time-weighted tensor-pipe activity.
  python scripts/gemm_traffic.py gpurun_out/r2f_step_metrics.csv deer_b 224 "<when / commit>" > profiles/r02_gemm_traffic.json"""
import collections, csv, json, re, sys
rows = collections.OrderedDict()
for r in csv.DictReader(l for l in open(sys.argv[1]) if not l.startswith("==")):
    d = rows.setdefault(r["ID"], {"name": re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")})
    v = float(r["Metric Value"].replace(",", "")) if r["Metric Value"] not in ("", "n/a") else 0.0
    m, u = r["Metric Name"], r["Metric Unit"]
    if m == "gpu__time_duration.sum":
        d["us"] = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    elif m.startswith("dram__bytes"):
        d[m] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    elif m.startswith("sm__pipe_tensor"):
        d["tensor"] = v
g = [d for d in rows.values() if d["name"].startswith("gemm_tc")]
t = sum(d["us"] for d in g)
big = [d for d in g if d["us"] > 150]
print(json.dumps({
    "source": "DEER_NCU_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
              "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none python bench.py --steps 1 --warmup 3 "
              "(the timed step of the bench command; every gemm_tc2_kernel / gemm_tc_kernel launch)",
    "workload": sys.argv[2] if len(sys.argv) > 2 else "deer_b", "episodes_per_gpu": int(sys.argv[3]) if len(sys.argv) > 3 else 224,
    "captured_at": sys.argv[4] if len(sys.argv) > 4 else "",
    "launches": len(g),
    "dram_read_bytes_per_launch": sum(d.get("dram__bytes_read.sum", 0) for d in g) / len(g),
    "dram_write_bytes_per_launch": sum(d.get("dram__bytes_write.sum", 0) for d in g) / len(g),
    "traffic_bytes_per_launch": sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in g) / len(g),
    "avg_launch_us_under_ncu": t / len(g),
    "tensor_pipe_active_pct_time_weighted": sum(d["tensor"] * d["us"] for d in g) / t,
    "tensor_pipe_active_pct_large_launches": sum(d["tensor"] * d["us"] for d in big) / max(1e-9, sum(d["us"] for d in big)),
    "n_large_launches": len(big)}, indent=1))
