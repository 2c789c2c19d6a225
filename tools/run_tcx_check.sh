#!/bin/bash
# A/B of the cluster-resident tensor-core recurrence (lstm_tcx.cu) on the device, one pass per process so that a trap in
# one pass does not hide the other.  usage: tools/run_tcx_check.sh [groups...]   (logs under gpurun_out/)
mkdir -p gpurun_out
G="${@:-xsmall}"
for only in fwd bwd; do
  CLSTM_B200_SELFTEST_ONLY=$only timeout 300 python tools/tc_selftest.py --x $G > gpurun_out/tcx_$only.log 2>&1
  echo "== $only exit $?"; tail -12 gpurun_out/tcx_$only.log
done
