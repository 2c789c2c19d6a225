#!/bin/bash
# ncu launch lists + full captures of round 2 (one GPU; see /opt/skills/guides/B200_PROFILING.md).  Outputs under gpurun_out/,
# summarised into profiles/ by tools/summarize_profiles_r2.py.
mkdir -p gpurun_out
M="--metrics gpu__time_duration.sum --clock-control none --csv"
ncu $M --log-file gpurun_out/launches_r2_cfg2.csv python tools/prof_step.py 100 32 500 3 > /dev/null 2>&1
ncu $M --log-file gpurun_out/launches_r2_cfg3.csv python tools/prof_step.py 200 128 200 3 2000 > /dev/null 2>&1
ncu $M --log-file gpurun_out/launches_r2_cfg4.csv python tools/prof_step.py 400 256 200 3 2000 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"lstm_tcx_fwd|lstm_tcx_bwd|gemm_x_kernel|split_transpose" -c 10 -o gpurun_out/prof_r2_cfg3 -f \
    python tools/prof_step.py 200 128 200 2 2000 > gpurun_out/prof_r2_cfg3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"lstm_tcx_fwd|lstm_tcx_bwd|gemm_x_kernel" -c 5 -o gpurun_out/prof_r2_cfg4 -f \
    python tools/prof_step.py 400 256 200 2 2000 > gpurun_out/prof_r2_cfg4.log 2>&1
ls -la gpurun_out/*.ncu-rep
