#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q -x -k "batchnorm_backward_sums or transposed_orientation or conv_epilogue_batch" 2>&1 | tail -5
RD_CONVT_PROF=1 timeout 120 python scripts/convt_prof.py 2>&1 | grep -E "==|prof"
timeout 300 python scripts/bwdsums_ab.py | tee gpurun_out/bwdsums_ab3.jsonl
timeout 300 python scripts/conv_t_tiles.py | tee gpurun_out/conv_t_tiles2.jsonl
