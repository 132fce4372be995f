"""Turns the ncu artefacts brought back in gpurun_out/ into small tracked summaries under profiles/."""
import csv, json, os, subprocess, sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r2"
GO, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(PR, exist_ok=True)
KEYS = ["gpu__time_duration.sum", "smsp__issue_active.avg", "sm__inst_executed.avg.per_cycle_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "smsp__average_warps_issue_stalled", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active", "gpu__dram_throughput",
        "sm__pipe_tensor_cycles_active", "sm__warps_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "lts__t_sector_hit_rate", "sm__inst_executed_pipe_tensor", "launch__shared_mem_per_block"]

# 1. launch list -> per-kernel mean/share
lp = os.path.join(GO, "launches_%s.csv" % TAG)
rows = [r for r in csv.DictReader(l for l in open(lp) if not l.startswith("=="))]
agg = defaultdict(list)
for r in rows:
    agg[r["Kernel Name"].split("(")[0]].append(float(r["Metric Value"]) / 1000.0)
nsteps = max(1, max((len(v) for k, v in agg.items() if "fusion" in k), default=1))      # one fused-kernel launch per step
per_step = {k: sum(v) / nsteps for k, v in agg.items()}                                  # time per step (a kernel may launch twice)
tot = sum(per_step.values())
with open(os.path.join(PR, "launches_%s.md" % TAG), "w") as f:
    f.write("# ncu launch list, `python bench.py --steps 3 --warmup 3` (cfg2), %s\n\n" % TAG)
    f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` — cold-cache, serialised: compare shares.\n\n")
    f.write("| kernel | launches | us per step | share of step |\n|---|---:|---:|---:|\n")
    for k, v in sorted(per_step.items(), key=lambda x: -x[1]):
        f.write("| `%s` | %d | %.1f | %.1f %% |\n" % (k, len(agg[k]), v, 100 * v / tot))
    f.write("\nsum per step: %.1f us (%d steps captured)\n" % (tot, nsteps))
os.replace(lp, os.path.join(PR, "launches_%s.csv" % TAG)) if False else None
import shutil
shutil.copy(lp, os.path.join(PR, "launches_%s.csv" % TAG))

# 2. full captures -> selected raw metrics
traffic = {}
for name, kern in (("pipe", "epi_fusion_pipe"), ("zgemm", "epi_zgemm"), ("stage", "epi_stage")):
    rep = os.path.join(GO, "prof_%s_%s.ncu-rep" % (kern, TAG))
    if not os.path.exists(rep):
        continue
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(out.splitlines()))
    hdr, units, vals = rr[0], rr[1], rr[2]
    sel = {h: (v, u) for h, u, v in zip(hdr, units, vals) if any(h.startswith(k) for k in KEYS) or h in ("Kernel Name",)}
    with open(os.path.join(PR, "ncu_%s_%s.md" % (name, TAG)), "w") as f:
        f.write("# ncu --set full, kernel `%s`, %s (one launch, cfg2)\n\n| metric | value | unit |\n|---|---:|---|\n" % (sel.get("Kernel Name", ("?",))[0], TAG))
        for h in sorted(sel):
            f.write("| %s | %s | %s |\n" % (h, sel[h][0], sel[h][1]))
    def num(k):
        v, u = sel.get(k, ("0", ""))
        x = float(v.replace(",", ""))
        mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
        return x * mult
    traffic[name] = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
if traffic:
    tp = os.path.join(PR, "traffic.json")
    cur = json.load(open(tp)) if os.path.exists(tp) else {}
    cur["cfg2"] = {"dram_bytes_per_launch": traffic.get("pipe"), "by_kernel": traffic, "step_total": sum(traffic.values()), "tag": TAG}
    json.dump(cur, open(tp, "w"), indent=1)
print(open(os.path.join(PR, "launches_%s.md" % TAG)).read()); print(traffic)
