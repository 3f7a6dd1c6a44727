#!/bin/bash
# round 2, GPU call A: full GPU test suite, smoke, bench (both storage types)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_f16.json 2> gpurun_out/bench_f16.err; echo "bench f16 rc=$?"; tail -c 600 gpurun_out/bench_f16.err
python bench.py --steps 10 --warmup 3 --dtype bf16 --quick > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench bf16 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_f16.json","gpurun_out/bench_bf16.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"].get("value"), d["e2e"].get("error"))
        r=d.get("roofline") or {}
        print("  roofline", r.get("achieved"), r.get("frac"), r.get("error"))
        for k,v in (r.get("families") or {}).items(): print("   ",k,v)
        for k in ("meta_kernel","postprocess","forward_b8","train_step_b4","cpu_baseline"):
            if k in d and d[k] is not None: print("  ",k, json.dumps(d[k])[:600])
    except Exception as e:
        print(f, "unreadable", e)
PY
