#!/bin/bash
# ncu --set full captures of the training-step kernels (eager launches so that kernels appear in tape order):
#   bwd: four head layers of level 0 (128 ch @ 4x64x2656): bwd_reduce, bwd_apply, wgrad, dgrad conv
#   fwd: the first backbone layers (64 ch @ 4x64x2656): conv, stats, fwd_apply
#   loss: the three rpn_loss kernels of the first graph replay with real targets
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
F='regex:conv_kernel|wgrad_kernel|bwd_apply|bwd_reduce|stats_kernel|fwd_apply'
timeout 400 ncu --set full --clock-control none -k "$F" -s 991 -c 16 -o gpurun_out/prof_train_bwd -f \
  python scripts/train_bench.py --batch 4 --steps 1 --warmup 1 --eager > gpurun_out/ncu_train_bwd.log 2>&1; echo "ncu-bwd exit $?"
timeout 400 ncu --set full --clock-control none -k "$F" -s 645 -c 9 -o gpurun_out/prof_train_fwd -f \
  python scripts/train_bench.py --batch 4 --steps 1 --warmup 1 --eager > gpurun_out/ncu_train_fwd.log 2>&1; echo "ncu-fwd exit $?"
timeout 400 ncu --set full --clock-control none -k regex:rpn_loss_kernel -s 12 -c 3 -o gpurun_out/prof_train_loss -f \
  python scripts/train_bench.py --batch 4 --steps 1 --warmup 1 > gpurun_out/ncu_train_loss.log 2>&1; echo "ncu-loss exit $?"
for t in bwd fwd loss; do python scripts/ncu_summary.py gpurun_out/prof_train_$t.ncu-rep gpurun_out/r01_train_${t}_ncu_full.csv; done
ls -la gpurun_out/*.ncu-rep
