#!/bin/bash
# round 2, GPU call D: tests (stage-level parity new) + ncu captures
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
python - <<'PY'
import json
for f in ("parity_stages_f16","parity_stages_bf16"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        for k,e in d["report"].items(): print(f, k, {kk:(round(v,5) if isinstance(v,float) else v) for kk,v in e.items()})
    except Exception as e: print(f, e)
PY
bash scripts/gpu_r2_ncu.sh
