#!/bin/bash
# ncu capture of the CTA-pair K3 step GEMM (run under gpurun, 1 GPU), after the same command exited 0 without ncu.
set -o pipefail
export B=${B:-37888}
CMD="python tools/k3_ab.py"
$CMD > gpurun_out/r2b_k3_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2b_k3_plain.log; exit 1; }
cat gpurun_out/r2b_k3_plain.log
ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 25 -c 1 -f -o gpurun_out/r2b_gemm_tc2 $CMD > gpurun_out/r2b_ncu_gemm2.log 2>&1
echo "gemm2 rc=$?"
ls -la gpurun_out/ | grep r2b_
