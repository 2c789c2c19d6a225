#!/bin/bash
# timing of the cluster-resident recurrence under its tuning switches (TcxArgs::flags); logs under gpurun_out/
mkdir -p gpurun_out
for F in ${@:-0 1 2 4 8 16}; do
  CLSTM_B200_TCX_FLAGS=$F CLSTM_B200_TC_DBG=1 timeout 200 python tools/tc_selftest.py --x t4 t3 > gpurun_out/tcx_flags_$F.log 2>&1
  echo "== flags $F"; grep -h "ms_tc_fwd" gpurun_out/tcx_flags_$F.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['case'], 'ok' if d['ok'] else 'BAD', 'fwd %.2f bwd %.2f' % (d['ms_tc_fwd'], d['ms_tc_bwd']))"
done
