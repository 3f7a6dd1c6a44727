#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "meta_kernel_backward_from_nhwc" 2>&1 | tail -8
timeout 1200 python -m pytest tests/test_gpu_train.py tests/test_gpu_train_step.py tests/test_gpu_parity_full.py -m gpu -q -x 2>&1 | tail -4
timeout 600 python scripts/ab_env.py "" | tee gpurun_out/ab_meta_nhwc.jsonl
