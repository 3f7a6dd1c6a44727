#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q -x -k "transposed" 2>&1 | tail -5
timeout 300 python scripts/conv_t_ab.py 2>&1 | tail -6 | cut -c1-330
for t in 1 0; do RD_CONV_T64=$t timeout 200 python scripts/ab_overlap.py 2>&1 | tail -1; done
RD_CONV_T64=1 timeout 200 python scripts/ab_overlap.py 2>&1 | tail -1
