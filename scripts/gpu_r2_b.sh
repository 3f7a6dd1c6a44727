#!/bin/bash
# round 2, GPU call B: GPU test suite (incl. full-size parity), quick bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --quick > gpurun_out/bench_f16_quick.json 2> gpurun_out/bench_f16_quick.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_f16_quick.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_f16_quick.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"].get("value"), d["e2e"].get("error"))
        r=d.get("roofline") or {}
        print("  roofline", r.get("achieved"), r.get("frac"), r.get("error"), "sum", r.get("step_ms_sum_of_kernels"))
        for k,v in (r.get("families") or {}).items(): print("   ",k,v)
        for t in (r.get("top_calls") or []): print("     ", t)
    except Exception as e:
        print(f, "unreadable", e)
for f in ("parity_full_f16","parity_full_bf16"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, d["layers"], d["worst"])
    except Exception as e: print(f, e)
for f in ("parity_e2e_f16","parity_e2e_bf16","parity_cfg4_b8"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        if "grads" in d:
            cs=sorted(v["cos"] for v in d["grads"].values()); rs=sorted(v["rms"] for v in d["grads"].values())
            print(f, "head", d["head"], "loss", d["loss_ref"], d["loss_gpu"], "cos min/med", cs[0], cs[len(cs)//2], "rms med/max", rs[len(rs)//2], rs[-1])
        else: print(f, d)
    except Exception as e: print(f, e)
PY
