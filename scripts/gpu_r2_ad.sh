#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_parity.py -m gpu -q -x -k "conv or deconv or backbone or head or layer or stats or slice or dla" 2>&1 | tail -4
RD_CONV_PROF=1 timeout 120 python scripts/convtc_prof.py 2>&1 | grep -E "==|prof" | cut -c1-330
timeout 600 python scripts/ab_env.py "" | tee gpurun_out/ab_convtc16.jsonl
