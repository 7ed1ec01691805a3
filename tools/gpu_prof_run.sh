#!/bin/bash
# ncu evidence for one eager batch of the bench workload: launch list + full-set captures of selected kernels
# usage: gpu_prof_run.sh <round-tag> "<kernel:count:skip> ..."
mkdir -p gpurun_out
R=${1:-r1}
SPECS=${2:-"stem_kernel:1:0 gsf_gate_kernel:2:0 conv3x3g_kernel:2:0 gemm_tc_kernel:3:0"}
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$R.csv python tools/profile_step.py 13 bf16 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?" > gpurun_out/summary.txt
for spec in $SPECS; do
  IFS=: read -r name cnt skip <<< "$spec"
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$name -s $skip -c $cnt -f -o gpurun_out/prof_${R}_${name}_$skip python tools/profile_step.py 13 bf16 > gpurun_out/ncu_$name.log 2>&1
  echo "ncu $name rc=$?" >> gpurun_out/summary.txt
done
ls -la gpurun_out; cat gpurun_out/summary.txt
