"""A/B of the batched tensor-core recurrence (lstm_tc.cu) against the fp32 SIMT kernels on the device, for a list of
(nhidden, lines, tmin, tmax) cases.  usage: python tools/tc_selftest.py [small|sizes|timing ...] ; one JSON line per case."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clstm_b200  # noqa: E402

CASES = {
    "small": [(64, 8, 5, 12), (64, 130, 3, 9), (200, 16, 10, 30), (104, 40, 1, 20)],
    "sizes": [(200, 128, 20, 60), (400, 256, 20, 50), (800, 128, 10, 30), (400, 32, 30, 40), (256, 100, 17, 33)],
    "timing": [(200, 128, 200, 2000), (400, 256, 200, 2000), (400, 32, 200, 2000)],
    "timing800": [(800, 128, 512, 512)],
    "t3": [(200, 128, 200, 2000)],
    "xsmall": [(200, 16, 3, 8), (200, 40, 1, 20), (100, 20, 5, 12), (256, 33, 4, 9), (128, 128, 10, 30)],
    "t4": [(400, 256, 200, 2000)],
    "t4s": [(400, 32, 200, 2000)],
    "xwide": [(400, 20, 3, 9), (400, 40, 1, 20), (320, 33, 4, 9), (480, 17, 2, 7), (288, 16, 5, 9)],
    "t2": [(100, 32, 500, 500)],
    "t2b": [(100, 128, 500, 500)],
    "t3s": [(200, 32, 512, 512)],
    "xrest": [(100, 20, 5, 12), (256, 33, 4, 9), (128, 128, 10, 30), (96, 7, 2, 15), (160, 50, 1, 25), (200, 700, 2, 6)],
}

if __name__ == "__main__":
    xmode = "--x" in sys.argv                # the cluster-resident variant (lstm_tcx.cu)
    groups = [a for a in sys.argv[1:] if a != "--x"] or ["small"]
    for g in groups:
        for (no, B, t0, t1) in CASES[g]:
            try:
                r = clstm_b200.selftest_lstm(no, B, t0, t1, seed=7, cluster_resident=xmode)
                r.update(case=[no, B, t0, t1], ok=bool(max(r["d_gates"], r["d_cell"], r["d_h"], r["d_hprev"]) < 2e-5 and r["d_delta_rel"] < 1e-4))
            except Exception as e:  # noqa: BLE001
                r = dict(case=[no, B, t0, t1], ok=False, error=str(e))
            print(json.dumps(r), flush=True)
            if "error" in r:
                sys.exit(1)   # a trapped kernel poisons the context: stop here
