"""Summarise an `ncu --page source --csv` dump: hottest SASS instructions with their stall reasons.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K | python tools/ncu_src.py [topN] [loop_only]"""
import csv, sys
rows = list(csv.reader(sys.stdin))
import os
inst = int(os.environ.get("NCU_INSTANCE", "0"))        # which kernel instance of the dump to summarise
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = his[inst]
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:
    if r and r[0] in ("Kernel Name", "Address"):
        break                      # first kernel instance only
    if len(r) == len(hdr):
        data.append(r)
top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[col["# Samples"]] or 0) for r in data)
tot_inst = sum(int(r[col["Instructions Executed"]] or 0) for r in data)
print("total samples", tot, "total warp-instr", tot_inst, "n_sass", len(data))
agg = {s: sum(int(r[col[s]] or 0) for r in data) for s in stalls}
print("stall totals:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:10])
print("--- hottest by samples")
for r in sorted(data, key=lambda r: -int(r[col["# Samples"]] or 0))[:top]:
    st = sorted(((int(r[col[s]] or 0), s[6:]) for s in stalls if int(r[col[s]] or 0)), reverse=True)[:3]
    print("%6s %8s  %-70s %s" % (r[col["# Samples"]], r[col["Instructions Executed"]], r[col["Source"]][:70], st))
