#!/bin/bash
# round 2, GPU call I: transposed-orientation conv kernel -- correctness first (short timeout), then A/B, then the suite + bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q -x -k "transposed" 2>&1 | tail -15
rc=${PIPESTATUS[0]}
echo "transposed tests rc=$rc"
if [ "$rc" != "0" ]; then exit 0; fi
timeout 300 python scripts/conv_t_ab.py 2>&1 | tail -12 | tee gpurun_out/r02_conv_t_ab.jsonl
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
for t in 1 0; do
  RD_CONV_T=$t timeout 600 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/bench_convt$t.json 2> gpurun_out/bench_convt$t.err; echo "bench conv_t=$t rc=$?"; tail -c 200 gpurun_out/bench_convt$t.err
done
python - <<'PY'
import json
for f in ("gpurun_out/bench_convt1.json","gpurun_out/bench_convt0.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"].get("value"))
        r=d["roofline"]; print("  roofline", r["achieved"], r["frac"], "sum", r["step_ms_sum_of_kernels"])
        for k,v in r["families"].items(): print("   ",k,v)
        for t in r["top_calls"][:8]: print("     ", t)
    except Exception as e: print(f, "unreadable", e)
PY
