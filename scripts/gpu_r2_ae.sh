#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_rpn_loss.py -m gpu -q -x 2>&1 | tail -4
timeout 1200 python -m pytest tests/test_gpu_train.py tests/test_gpu_train_step.py tests/test_gpu_parity_full.py tests/test_symbol.py -m gpu -q -x 2>&1 | tail -4
timeout 600 python scripts/ab_env.py "" | tee gpurun_out/ab_nhwc_loss.jsonl
