#!/bin/bash
# round 2, GPU call C: PDL A/B + tests
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
for pdl in 1 0; do
  RD_PDL=$pdl python bench.py --steps 10 --warmup 3 --quick > gpurun_out/bench_pdl$pdl.json 2> gpurun_out/bench_pdl$pdl.err; echo "bench pdl=$pdl rc=$?"; tail -c 300 gpurun_out/bench_pdl$pdl.err
done
python - <<'PY'
import json
for f in ("gpurun_out/bench_pdl1.json","gpurun_out/bench_pdl0.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"].get("value"), d["e2e"].get("error"))
        r=d.get("roofline") or {}
        print("  roofline", r.get("achieved"), r.get("frac"), r.get("error"), "sum", r.get("step_ms_sum_of_kernels"))
        for k,v in (r.get("families") or {}).items(): print("   ",k,v)
    except Exception as e:
        print(f, "unreadable", e)
for f in ("parity_derivative_f16",):
    try: print(f, json.load(open("gpurun_out/%s.json"%f)))
    except Exception as e: print(f, e)
PY
