#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_full.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_full.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"].get("value"), "launches", d["gpu_launches"], "clocks", d["clocks"])
r=d["roofline"]; print("roofline", r.get("achieved"), r.get("frac"), r.get("traffic"), "sum", r.get("step_ms_sum_of_kernels"), r.get("error"))
for k,v in (r.get("families") or {}).items(): print("   ",k,v)
for c in r.get("top_calls", []): print("      ", c)
for k in ("meta_kernel","postprocess","forward_b8","train_step_b4","cpu_baseline"):
    print(k, json.dumps(d.get(k))[:400])
PY
