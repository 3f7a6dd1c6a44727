#!/bin/bash
# round 2, GPU call F: full GPU suite, wNMS timing, ncu launch list, full bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
python - <<'PY'
import time, torch, numpy as np, sys
sys.path.insert(0, ".")
from rangedet_b200 import ops, synth
for n in (20000, 50000, 100000):
    for clustered in (True, False):
        d = torch.from_numpy(synth.wnms_dets(n, seed=0, clustered=clustered)).cuda()
        ops.wnms_4c_device(d, 0.1, 0.5, False, 100); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3): o, k = ops.wnms_4c_device(d, 0.1, 0.5, False, 100)
        torch.cuda.synchronize()
        print("wnms n=%d clustered=%s: %.2f ms, kept %d" % (n, clustered, (time.perf_counter() - t0) / 3 * 1e3, k.numel()))
PY
RD_PDL=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r02_train_step_launches.csv python scripts/ncu_targets.py step > gpurun_out/ncu_step.log 2>&1; echo "ncu launch list rc=$?"
tail -2 gpurun_out/ncu_step.log
python scripts/launch_summary.py gpurun_out/r02_train_step_launches.csv > gpurun_out/r02_train_step_launch_summary.txt 2>&1; head -30 gpurun_out/r02_train_step_launch_summary.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_full.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_full.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"].get("value"), "launches", d["gpu_launches"])
print("roofline", d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["traffic"])
for k in ("meta_kernel","postprocess","forward_b8","train_step_b4","cpu_baseline"):
    print(k, json.dumps(d.get(k))[:700])
PY
