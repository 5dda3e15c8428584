#!/bin/bash
# ncu evidence of the round (run under gpurun, 1 GPU): launch list of a short bench run + full captures of the three
# dominant kernels (filter scan, K3 step GEMM of the CTA-pair kernel, refine).  Each capture only after the same command
# exited 0 without ncu.
set -o pipefail
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-library-baseline --no-parity"
$CMD > gpurun_out/r2_ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2_ncu_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:sl_filter_kernel -s 3 -c 1 -f -o gpurun_out/r2_sl_filter $CMD > gpurun_out/r2_ncu_filter.log 2>&1
echo "filter rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 47 -c 1 -f -o gpurun_out/r2_gemm_tc2 $CMD > gpurun_out/r2_ncu_gemm.log 2>&1
echo "gemm rc=$?"
ncu --set full --clock-control none --import-source on -k regex:sl_refine_kernel -s 1 -c 1 -f -o gpurun_out/r2_sl_refine $CMD > gpurun_out/r2_ncu_refine.log 2>&1
echo "refine rc=$?"
ls -la gpurun_out/ | grep r2_
