#!/bin/bash
# round 2, GPU call G: same-box A/B of the step changes + ncu launch list of two eager steps
mkdir -p gpurun_out
python scripts/ab_step.py > gpurun_out/r02_ab_step.jsonl 2> gpurun_out/ab_step.err; echo "ab rc=$?"; tail -c 300 gpurun_out/ab_step.err; cat gpurun_out/r02_ab_step.jsonl | cut -c1-120
RD_PDL=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r02_train_step_launches.csv python scripts/ncu_targets.py step > gpurun_out/ncu_step.log 2>&1; echo "ncu launch list rc=$?"
tail -2 gpurun_out/ncu_step.log
python scripts/launch_summary.py gpurun_out/r02_train_step_launches.csv > gpurun_out/r02_train_step_launch_summary.txt 2>&1; head -30 gpurun_out/r02_train_step_launch_summary.txt
python -m pytest tests/test_gpu_train.py -m gpu -q -k "train_layer" 2>&1 | tail -3
