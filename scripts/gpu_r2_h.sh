#!/bin/bash
# round 2, GPU call H: full GPU suite + quick bench after the wgrad-reduce / one-slot-per-CTA statistics changes
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8
python bench.py --steps 20 --warmup 5 --quick > gpurun_out/bench_quick_h.json 2> gpurun_out/bench_quick_h.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_quick_h.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_quick_h.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"].get("value"))
r=d["roofline"]; print("roofline", r["achieved"], r["frac"], "sum", r["step_ms_sum_of_kernels"])
for k,v in r["families"].items(): print("   ",k,v)
PY
