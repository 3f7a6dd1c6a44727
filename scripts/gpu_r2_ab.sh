#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "meta" 2>&1 | tail -4
timeout 300 python - <<'PY'
import sys, json, torch
sys.path.insert(0, ".")
import bench
r = bench.meta_kernel_leg(torch.device("cuda", 0), bench.peaks(), 20)
print(json.dumps(r)[:900])
PY
timeout 600 python scripts/ab_env.py "" | tee gpurun_out/ab_meta_epi.jsonl
