#!/bin/bash
# round 2: the driver's end-of-round sequence on one GPU -- GPU test suite, smoke, reference arm, full bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_ref.json
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_full.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_full.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"].get("value"), "launches", d["gpu_launches"], "clocks", d["clocks"])
r=d["roofline"]; print("roofline", r["achieved"], r["frac"], r["traffic"], "sum", r["step_ms_sum_of_kernels"])
for k,v in r["families"].items(): print("   ",k,v)
for k in ("meta_kernel","postprocess","forward_b8","train_step_b4","cpu_baseline"):
    print(k, json.dumps(d.get(k))[:600])
PY
