#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -x -k "bn or layer" 2>&1 | tail -3
timeout 900 python scripts/ab_env.py RD_BN_REV=0 RD_BN_REV=6 RD_BN_REV=2 RD_BN_REV=4 RD_BN_REV=15 RD_BN_REV=10 | tee gpurun_out/ab_bn_rev.jsonl
