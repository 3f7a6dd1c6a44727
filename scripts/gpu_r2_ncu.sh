#!/bin/bash
# round 2 ncu captures (the launch-list pass runs with RD_PDL=0: ncu reports LaunchFailed on programmatic-dependent graph edges;
# kernel durations are what is wanted there, not the overlap): launch list of two replayed cfg-5 steps + `--set full` of the dominant kernels (B=2, fp16)
mkdir -p gpurun_out
RD_PDL=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r02_train_step_launches.csv python scripts/ncu_targets.py step > gpurun_out/ncu_step.log 2>&1; echo "ncu launch list rc=$?"
tail -2 gpurun_out/ncu_step.log
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o /tmp/r02_kernels -f \
  --kernel-name 'regex:conv|wgrad|reduce_kernel|s_fwd|s_bwd|finalize|cand_kernel|resolve_kernel|greedy_kernel|merge_kernel' python scripts/ncu_targets.py kernels > gpurun_out/ncu_kernels.log 2>&1; echo "ncu full rc=$?"
python scripts/ncu_summary.py /tmp/r02_kernels.ncu-rep gpurun_out/r02_kernels_ncu_full.csv gpurun_out/r02_kernels_traffic.json
python scripts/launch_summary.py gpurun_out/r02_train_step_launches.csv > gpurun_out/r02_train_step_launch_summary.txt 2>&1; head -40 gpurun_out/r02_train_step_launch_summary.txt
ls -la /tmp/*.ncu-rep   # the report itself stays on the box (> 64 MiB with the weighted-NMS launches); the CSV summaries travel
