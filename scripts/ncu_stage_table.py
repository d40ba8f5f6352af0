"""Per-kernel and per-stage table from an ncu per-launch metrics CSV of ONE step (scripts/one_step.py):
time share, DRAM GB/s (dram bytes / duration) against the measured HBM peak, tensor-pipe active %.
GEMM launches are assigned to stages by launch order: ViT, Perceiver, then LLM+head (counts from one_step.json)."""
import collections, csv, json, re, sys
csv_path, meta_path = sys.argv[1], sys.argv[2]
meta = json.load(open(meta_path))
peaks = {"hbm_gbs": 6555.5}
try:
    peaks.update(json.load(open("MEASURED_PEAKS.json")))
except Exception:
    pass
rows = collections.OrderedDict()
for r in csv.DictReader(l for l in open(csv_path) if not l.startswith("==")):
    d = rows.setdefault(r["ID"], {"name": re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", ""), "grid": r.get("Grid Size", "")})
    v = float(r["Metric Value"].replace(",", "")) if r["Metric Value"] not in ("", "n/a") else 0.0
    m, u = r["Metric Name"], r["Metric Unit"]
    if m == "gpu__time_duration.sum":
        d["us"] = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    elif m.startswith("dram__bytes"):
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        d["dram"] = d.get("dram", 0.0) + v * scale
    elif m.startswith("sm__pipe_tensor"):
        d["tensor"] = v
L = list(rows.values())
tot = sum(x.get("us", 0) for x in L)
gs = meta["gemm_launches_by_stage"]
order, bounds, acc = ["vit", "perceiver", "llm", "head"], {}, 0
gi = 0
for x in L:
    if "gemm_tc" in x["name"] or "gemm_simt" in x["name"]:
        st = "vit" if gi < gs["vit"] else ("perceiver" if gi < gs["vit"] + gs["perceiver"] else "llm+head")
        x["stage"] = st
        gi += 1
print(f"# {csv_path}: {len(L)} launches, {tot / 1e3:.2f} ms (ncu-serialised, cold caches: compare shares); HBM peak {peaks['hbm_gbs']} GB/s")
print("## per kernel")
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for x in L:
    a = agg[x["name"]]; a[0] += 1; a[1] += x.get("us", 0); a[2] += x.get("dram", 0); a[3] += x.get("tensor", 0) * x.get("us", 0)
for k, (n, t, b, tw) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t / 1e3:9.3f} ms {100 * t / tot:6.2f}%  n={n:4d}  dram {b / (t * 1e-6) / 1e9:7.1f} GB/s ({100 * b / (t * 1e-6) / 1e9 / peaks['hbm_gbs']:5.1f}% of peak)  tensor-pipe {tw / t if t else 0:5.1f}%  {k}")
print("## GEMM launches per stage (tcgen05 kernels; tensor-pipe % is time-weighted sm__pipe_tensor_cycles_active pct_of_peak_sustained_elapsed)")
sagg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for x in L:
    if "stage" in x:
        a = sagg[x["stage"]]; a[0] += 1; a[1] += x.get("us", 0); a[2] += x.get("dram", 0); a[3] += x.get("tensor", 0) * x.get("us", 0)
for k, (n, t, b, tw) in sagg.items():
    print(f"{k:10s} n={n:4d} {t / 1e3:9.3f} ms  tensor-pipe {tw / t:5.1f}%  dram {b / (t * 1e-6) / 1e9:7.1f} GB/s")
big = [x for x in L if x.get("stage") == "llm+head" and x.get("us", 0) > 50]
if big:
    t = sum(x["us"] for x in big)
    print(f"llm GEMM launches > 50 us (the MPT-block / x-attn projections and MLPs): n={len(big)} {t / 1e3:.3f} ms tensor-pipe {sum(x.get('tensor', 0) * x['us'] for x in big) / t:.1f}%")
