"""Turns the raw ncu outputs of round 2 in gpurun_out/ into the tracked summaries under profiles/.

    launches_r2_<tag>.csv   ncu --metrics gpu__time_duration.sum launch list of `tools/prof_step.py ...`   (one per config)
    prof_r2_<tag>.ncu-rep   ncu --set full capture of the recurrent kernels                                 (optional)
->  profiles/launches_r2_<tag>_step.csv, profiles/r2_summary.json
usage: python tools/summarize_profiles_r2.py tag:"description" [tag:"description" ...]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
PHASE = {"lstm_tcx_fwd": "lstm_fwd", "lstm_tcx_bwd": "lstm_bwd", "lstm_tc_fwd": "lstm_fwd", "lstm_tc_bwd": "lstm_bwd", "lstm_fwd": "lstm_fwd",
         "lstm_bwd": "lstm_bwd", "gemm_x_kernel": "wgrad_gemm_x", "split_transpose_kernel": "wgrad_split",
         "gemm_tn_kernel": "wgrad_gemm", "ctc_": "ctc_align", "sgd_update": "sgd_update", "peer_allreduce": "allreduce"}


def short(name):
    n = name.split("(")[0]
    return n.replace("void ", "").replace("cb200::", "").replace("<unnamed>::", "").replace("unnamed>::", "").strip()


def launch_table(fn):
    rows = list(csv.reader(open(fn)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]
    kn, gs, bs, mv = h.index("Kernel Name"), h.index("Grid Size"), h.index("Block Size"), h.index("Metric Value")
    return [(r[kn], r[gs], r[bs], float(r[mv].replace(",", ""))) for r in rows[hi + 1:] if len(r) > mv]


def to_bytes(text):
    v, u = text.split()[:2]
    return float(v.replace(",", "")) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "Tbyte": 1e12}[u]


def main():
    summary = {"how": "ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES); "
                      "full captures: ncu --set full --clock-control none --import-source on", "configs": {}, "full_capture": []}
    for arg in sys.argv[1:]:
        tag, _, desc = arg.partition(":")
        fn = os.path.join(G, "launches_r2_%s.csv" % tag)
        if os.path.exists(fn):
            tab = launch_table(fn)
            idx = [i for i, t in enumerate(tab) if "sgd_update_kernel" in t[0] or "peer_allreduce_update" in t[0]]
            step = tab[idx[-2] + 1: idx[-1] + 1] if len(idx) >= 2 else tab
            total = sum(t[3] for t in step)
            agg = {}
            for t in step:
                a = agg.setdefault(short(t[0]), [0, 0.0])
                a[0] += 1
                a[1] += t[3]
            with open(os.path.join(P, "launches_r2_%s_step.csv" % tag), "w") as f:
                f.write("kernel,grid,block,duration_ns\n")
                for t in step:
                    f.write('"%s","%s","%s",%d\n' % (short(t[0]), t[1], t[2], t[3]))
            summary["configs"][tag] = {"workload": desc, "launches_per_step": len(step), "sum_kernel_us": total / 1e3,
                                       "share": {k: {"launches": v[0], "us": v[1] / 1e3, "share": v[1] / total}
                                                 for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])}}
        rep = os.path.join(G, "prof_r2_%s.ncu-rep" % tag)
        if os.path.exists(rep):
            out = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
            rows = list(csv.reader(out.splitlines()))
            hdr, units = rows[0], rows[1]
            want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                    "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
                    "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
                    "sm__cycles_active.avg", "sm__cycles_elapsed.max", "l1tex__m_xbar2l1tex_read_bytes.sum"]
            for r in rows[2:]:
                d = {"config": tag, "workload": desc}
                for w in want:
                    if w in hdr:
                        i = hdr.index(w)
                        d[w] = r[i] + (" " + units[i] if units[i] else "")
                d["Kernel Name"] = short(d["Kernel Name"])
                for key, ph in PHASE.items():
                    if d["Kernel Name"].startswith(key):
                        d["phase"] = ph
                        break
                try:
                    d["dram_bytes_per_launch"] = to_bytes(d["dram__bytes_read.sum"]) + to_bytes(d["dram__bytes_write.sum"])
                except Exception:
                    pass
                summary["full_capture"].append(d)
    json.dump(summary, open(os.path.join(P, "r2_summary.json"), "w"), indent=1)
    for tag, c in summary["configs"].items():
        print(tag, c["workload"], "sum %.1f us" % c["sum_kernel_us"])
        for k, v in list(c["share"].items())[:8]:
            print("   %-34s x%-3d %10.1f us  %5.1f%%" % (k, v["launches"], v["us"], 100 * v["share"]))
    for c in summary["full_capture"]:
        print({k: c[k] for k in c if k not in ("workload",)})


if __name__ == "__main__":
    main()
