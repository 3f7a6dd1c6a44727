#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -x -k "transposed_orientation" 2>&1 | tail -4
timeout 300 python scripts/conv_t_tiles.py | tee gpurun_out/conv_t_tiles.jsonl
timeout 900 python scripts/ab_env.py "" "RD_CONVT_TN=256" | tee gpurun_out/ab_convt_tn.jsonl
