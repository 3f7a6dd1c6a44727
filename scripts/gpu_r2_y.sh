#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_train.py tests/test_gpu_train_step.py tests/test_gpu_parity_full.py -m gpu -q -x 2>&1 | tail -4
timeout 900 python scripts/ab_env.py "AB_BWD_SUMS=0" "AB_BWD_SUMS=1" "AB_BWD_SUMS=auto" | tee gpurun_out/ab_bwdsums3.jsonl
