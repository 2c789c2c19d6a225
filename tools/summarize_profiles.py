"""Turns the raw ncu outputs in gpurun_out/ into the tracked summaries under profiles/ (round 1)."""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")

def launch_table(fn):
    rows = list(csv.reader(open(fn)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]; kn, gs, bs, mv = h.index("Kernel Name"), h.index("Grid Size"), h.index("Block Size"), h.index("Metric Value")
    data = rows[hi + 1:]
    return [(r[kn], r[gs], r[bs], float(r[mv].replace(",", ""))) for r in data]

def short(name):
    n = name.split("(")[0]
    return n.replace("void ", "").replace("cb200::", "").replace("<unnamed>::", "").replace("unnamed>::", "").strip()

tab = launch_table(os.path.join(G, "launches_r1_cfg2.csv"))
# one training step = the launches between two sgd_update kernels; take the last complete step
idx = [i for i, t in enumerate(tab) if "sgd_update_kernel" in t[0]]
step = tab[idx[-2] + 1: idx[-1] + 1]
# the step starts after the previous step's transpose/decode tail: rotate so it starts at the first gemm of forward
names = [short(t[0]) for t in step]
total = sum(t[3] for t in step)
agg = {}
for t in step:
    k = short(t[0]); a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += t[3]
with open(os.path.join(P, "launches_r1_cfg2_step.csv"), "w") as f:
    f.write("kernel,grid,block,duration_ns\n")
    for t in step:
        f.write('"%s","%s","%s",%d\n' % (short(t[0]), t[1], t[2], t[3]))
summary = {"workload": "cfg2: nhidden=100 H=48 T=500 B=32 nc=83, one training step, 1xB200",
           "how": "ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)",
           "launches_per_step": len(step), "sum_kernel_us": total / 1e3,
           "share": {k: {"launches": v[0], "us": v[1] / 1e3, "share": v[1] / total} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])}}
# full-capture metrics of the top kernels
rep = os.path.join(G, "prof_r1_top.ncu-rep")
if os.path.exists(rep):
    out = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
            "sm__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_fp32.sum"]
    units = rows[1]
    caps = []
    for r in rows[2:]:
        d = {}
        for w in want:
            if w in hdr:
                i = hdr.index(w); d[w] = r[i] + (" " + units[i] if units[i] else "")
        d["Kernel Name"] = short(d["Kernel Name"])
        caps.append(d)
    summary["full_capture"] = caps
json.dump(summary, open(os.path.join(P, "r1_summary.json"), "w"), indent=1)
print(json.dumps(summary["share"], indent=1)[:2500])
for c in summary.get("full_capture", []):
    print(c)
