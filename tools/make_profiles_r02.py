#!/usr/bin/env python
"""gpurun_out/ (scratch) -> profiles/ (tracked), round 2: launch list with per-kernel shares, ncu --set full summaries of
every hot kernel / shape, and profiles/ncu_traffic.json (DRAM bytes per score_kernel launch, read by bench.py)."""
import csv, io, json, os, subprocess, sys
from collections import OrderedDict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return (rows[0], rows[1], rows[2:]) if len(rows) > 2 else ([], [], [])

# ---- ncu summaries ----
reps = [f"{G}/{n}.ncu-rep" for n in ("r02_bench_c2", "r02_score_c3", "r02_score_c4", "r02_hypgen_c4", "r02_tri3", "r02_small2") if os.path.exists(f"{G}/{n}.ncu-rep")]
md = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), *reps], capture_output=True, text=True).stdout
head = ("# ncu `--set full --clock-control none` summaries, round 2 (B200)\n\n"
        "Captured by `tools/ncu_configs.sh`: the config-2 step of `bench.py` (r02_bench_c2: every kernel of the step), the scoring kernel at the "
        "config-3 shape (1M x 1M) and at the config-4 shape (1,024 of the 4,096 pairs x 4k x 4k), hypothesis generation at the config-4 shape, "
        "triangulation of 1M points, the fused small-problem kernel on the dino pair.  Times under ncu are cold-cache and serialised.\n\n")
open(f"{P}/r02_ncu.md", "w").write(head + md)
# ---- DRAM traffic of the scoring kernel at config 2 ----
if os.path.exists(f"{G}/r02_bench_c2.ncu-rep"):
    hdr, units, rows = raw(f"{G}/r02_bench_c2.ncu-rep")
    vals = []
    for r in rows:
        d = dict(zip(hdr, r))
        if "score_kernel" in d.get("Kernel Name", ""):
            def b(key):
                v, u = float(d[key].replace(",", "")), units[hdr.index(key)]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            vals.append(b("dram__bytes_read.sum") + b("dram__bytes_write.sum"))
    if vals:
        json.dump({"score_kernel_dram_bytes_per_launch": int(sum(vals) / len(vals)), "launches_averaged": len(vals),
                   "source": "profiles/r02_ncu.md: ncu --set full capture of `python bench.py --steps 2 --warmup 3 --no-extras` (gpurun_out/r02_bench_c2.ncu-rep), "
                             "dram__bytes_read.sum + dram__bytes_write.sum of score_kernel<8,1,256,1,0>"}, open(f"{P}/ncu_traffic.json", "w"), indent=1)
# ---- launch list ----
if os.path.exists(f"{G}/r02_launches.csv"):
    lines = [l for l in open(f"{G}/r02_launches.csv") if not l.startswith("==")]
    open(f"{P}/r02_launches.csv", "w").writelines(lines)
    agg = OrderedDict()
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].split("(")[0][:80]
        v = float(r["Metric Value"].replace(",", "")) * (1e3 if r.get("Metric Unit", "ns").startswith("us") else 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(t for _, t in agg.values())
    out = ["# ncu launch list, round 2 (`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 2 --warmup 3 --no-extras`)", "",
           "First 400 launches of the command: the headline config-2 region (and its warm-up / staged / solver-comparison regions).  Per-launch times "
           "under ncu are cold-cache and serialised: compare SHARES, not absolutes.", "", "| kernel | launches | mean ns | total ns | share |", "|---|---|---|---|---|"]
    for k, (c, t) in agg.items():
        out.append(f"| `{k}` | {c} | {t / c:.1f} | {t:.1f} | {100 * t / tot:.1f} % |")
    bj = f"{P}/r02_bench.json"
    if os.path.exists(bj):
        st = json.load(open(bj))["stage_ms"]
        out += ["", "bench.py (not under ncu) CUDA-event stage times of the same step, ms: " + ", ".join(f"{k} {v:.4f}" for k, v in st.items())]
    open(f"{P}/r02_launches.md", "w").write("\n".join(out) + "\n")
print("profiles written:", [f for f in sorted(os.listdir(P)) if f.startswith("r02") or f == "ncu_traffic.json"])
