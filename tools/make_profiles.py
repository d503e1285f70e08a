#!/usr/bin/env python
"""gpurun_out/ (scratch) -> profiles/ (tracked): bench lines, ncu launch list with per-kernel shares,
raw ncu metrics of the `--set full` capture and its summary table.  Run after tools/gpu_round4.sh."""
import csv
import io
import json
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"


def last_json(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


ours, ref = last_json(f"{G}/bench.json"), last_json(f"{G}/bench_ref.json")
json.dump(ours, open(f"{P}/{tag}_bench.json", "w"), indent=1)
json.dump(ref, open(f"{P}/{tag}_bench_reference.json", "w"), indent=1)

# ---- launch list ----
lines = [l for l in open(f"{G}/launches.csv") if not l.startswith("==")]
open(f"{P}/{tag}_launches.csv", "w").writelines(lines)
rows = list(csv.DictReader(io.StringIO("".join(lines))))
agg = OrderedDict()
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"].split("(")[0][:70]
    v = float(r["Metric Value"].replace(",", ""))
    if r.get("Metric Unit", "ns").startswith("us"):
        v *= 1e3
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
out = [f"# ncu launch list, round 1 (`ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 3 --warmup 3`)", "",
       "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.", "",
       "| kernel | launches | mean ns | total ns |", "|---|---|---|---|"]
for k, (c, t) in agg.items():
    out.append(f"| `{k}` | {c} | {t / c:.1f} | {t:.1f} |")
hot = [("ingest_xy_kernel", "ingest"), ("hypgen_kernel<128, 4, 0, 1>", "hypgen"), ("score_kernel<8, 1, 256, 1, 0>", "score"),
       ("select_pose_choose_kernel", "select"), ("triangulate_kernel", "triangulate")]
st = ours["stage_ms"]
bench_ms = {"ingest": st["ingest"], "hypgen": st["hypgen"], "score": st["score"],
            "select": st["select"] + st["pose_candidates"] + st["choose_pose"], "triangulate": st["triangulate"]}
means = {}
for pat, key in hot:
    for k, (c, t) in agg.items():
        if pat in k:
            means[key] = (k, t / c)
tot_ncu, tot_b = sum(v for _, v in means.values()), sum(bench_ms.values())
out += ["", "## Share of one hot-path step (default configuration: projector hypgen, packed scoring)", "",
        "| kernel | ncu mean ns | ncu share | bench.py CUDA-event ms | bench share |", "|---|---|---|---|---|"]
for key, (k, m) in means.items():
    out.append(f"| `{k}` | {m:.1f} | {100 * m / tot_ncu:.1f} % | {bench_ms[key]:.4f} | {100 * bench_ms[key] / tot_b:.1f} % |")
out += ["", f"bench.py (not under ncu): {ours['ms_per_step']:.4f} ms/step without per-stage events "
        f"({ours['stage_timing']['ms_per_step_with_stage_events']:.4f} ms with them: the per-kernel column above comes from that second region), "
        f"value {ours['value']:.4e} evals/s, e2e {ours['e2e']['value']:.4e} evals/s; score kernel share of the step "
        f"{100 * ours['roofline']['kernel_share_of_step']:.1f} % by CUDA events vs {100 * means['score'][1] / tot_ncu:.1f} % under ncu.",
        f"Reference arm on the same box: {ref['ms_per_step']:.1f} ms/step, {ref['value']:.4e} evals/s."]
open(f"{P}/{tag}_launches.md", "w").write("\n".join(out) + "\n")

# ---- full capture ----
rep = f"{G}/prof_path.ncu-rep"
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    raw = "".join(l + "\n" for l in raw.splitlines() if not l.startswith("=="))
    open(f"{G}/prof_path_raw.csv", "w").write(raw)
    open(f"{P}/{tag}_ncu_raw.csv", "w").write(raw)
    rr = list(csv.reader(io.StringIO(raw)))
    hdr, units = rr[0], rr[1]
    keys = [('gpu__time_duration.sum', 'duration'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'), ('launch__registers_per_thread', 'regs/thread'),
            ('sm__cycles_elapsed.avg.per_second', 'SM clock'), ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
            ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'FMA pipe cycles active %'), ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'FMA pipe inst %'),
            ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'ALU pipe inst %'), ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'XU pipe inst %'),
            ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots active %'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active % (occupancy)'),
            ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'), ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput %'),
            ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'stall: math pipe throttle'),
            ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'stall: wait (fixed latency)'),
            ('smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'stall: not selected'),
            ('smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'stall: no instruction (I-cache)'),
            ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'stall: short scoreboard (smem/MUFU)'),
            ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall: long scoreboard (global)'),
            ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'stall: barrier'),
            ('smsp__inst_executed.sum', 'warp instructions'), ('sass__inst_executed_local_loads', 'local (spill) loads')]
    seen = OrderedDict()
    for r in rr[2:]:
        if len(r) != len(hdr):
            continue
        seen.setdefault(r[hdr.index('Kernel Name')].split('(')[0], r)
    names = list(seen)
    o = ['# ncu `--set full` summary, round 1 (default configuration)', '',
         'Command: `ncu --set full --clock-control none --import-source on -k regex:score_kernel|hypgen_kernel|triangulate_kernel|select_pose -s 12 -c 8 python bench.py --steps 3 --warmup 3`',
         '(B200, BASELINE config 2: 10,000 correspondences x 65,536 hypotheses).  Numbers under ncu are NOT bench values.', '',
         '| metric | ' + ' | '.join('`' + n.replace('void ', '').replace('sfmb200::', '') + '`' for n in names) + ' |', '|---|' + '---|' * len(names)]
    for k, label in keys:
        if k not in hdr:
            continue
        i = hdr.index(k)
        vals = []
        for n in names:
            v = seen[n][i]
            try:
                v = '%.4g' % float(v.replace(',', ''))
            except ValueError:
                pass
            vals.append(v + ' ' + units[i])
        o.append('| ' + label + ' | ' + ' | '.join(vals) + ' |')
    sc = next((n for n in names if 'score_kernel' in n), None)
    traffic = None
    if sc:
        i1, i2 = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')

        def tobytes(v, u):
            f = float(v.replace(',', ''))
            return f * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
        traffic = tobytes(seen[sc][i1], units[i1]) + tobytes(seen[sc][i2], units[i2])
    o += ['', '## Reading', '',
          '* `score_kernel<8,1,256,1,0>` (packed FFMA2, 8 hypotheses/thread, one 256-thread CTA per SM, persistent stream-K grid of 148 CTAs, Sampson model, threshold folded into the coordinates: 17 FMA-pipe instructions per evaluation): the FMA pipe is the busy unit and the dominant stall is `math pipe throttle`: bound by the FP32 pipe as designed.  DRAM traffic per launch = E candidates (2.36 MB) + the duplicated scaled correspondence array (320 KB) + counts'
          + (f' = {traffic / 1e6:.2f} MB measured' if traffic else '') + ', ~0.07 % of HBM bandwidth: correspondence tiles are re-read from L2, never from HBM.  Algorithmic work: 6.5536e8 evaluations x 34 FLOP = 22.3 GFLOP per launch.  Scalar-vs-packed comparison: r01_ncu_score_variants.md.',
          '* `hypgen_kernel<128, 4, 0, 1>` = 8x8 Cholesky projector (default): 128 registers, 16 warps/SM.  `hypgen_kernel<128, 2, 0, 0>` (when present) = 9x9 Jacobi eigensolve: 255 registers, 8 warps/SM, top stall `no instruction` (a sweep is ~3,000 unrolled instructions).',
          '* `select_pose_choose_kernel`: 4 active lanes of dependent latency (3x3 SVD + 4x4 null vector + 4x4 inverse); 53 us as three kernels -> 16 us fused -> ~7-10 us with MUFU angles and the inverse-iteration null vector.',
          '* `triangulate_kernel` at 10k points is launch/latency bound (40 CTAs); at 1M points it runs at ~1,580 GB/s of 32 B/point traffic (r01_configs_n1.jsonl / DESIGN.md 3.4).']
    open(f"{P}/{tag}_ncu_summary.md", "w").write("\n".join(o) + "\n")
    print("score kernel DRAM traffic per launch (bytes):", traffic)
print("ours", ours["ms_per_step"], ours["value"], "e2e", ours["e2e"]["value"], "frac", ours["roofline"]["frac"])
print("ref", ref["ms_per_step"], ref["value"], "ratio value", ours["value"] / ref["value"], "e2e", ours["e2e"]["value"] / ref["value"])
