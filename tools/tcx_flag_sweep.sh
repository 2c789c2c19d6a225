#!/bin/bash
# timing of the cluster-resident recurrence under its build-time tuning switches (lstm_tcx.cu: CB200_TCX_FLAGS); rebuilds the
# library per value (nvcc must be on the box) and restores flags 0 at the end; logs under gpurun_out/
mkdir -p gpurun_out
for F in ${@:-0 1 2 8 16 32}; do   # (4 is no longer buildable, see lstm_tcx.cu)
  touch clstm_b200/csrc/lstm_tcx.cu; make -s -C clstm_b200/csrc TCX_FLAGS=$F > /dev/null || exit 1
  CLSTM_B200_TC_DBG=1 timeout 200 python tools/tc_selftest.py --x t4 t3 > gpurun_out/tcx_flags_$F.log 2>&1
  echo "== flags $F"; grep -h "ms_tc_fwd" gpurun_out/tcx_flags_$F.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['case'], 'ok' if d['ok'] else 'BAD', 'fwd %.2f bwd %.2f' % (d['ms_tc_fwd'], d['ms_tc_bwd']))"
done
touch clstm_b200/csrc/lstm_tcx.cu; make -s -C clstm_b200/csrc TCX_FLAGS=0 > /dev/null
