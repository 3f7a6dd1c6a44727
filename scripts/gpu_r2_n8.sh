#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench n4 rc=$?"; tail -c 300 gpurun_out/bench_n8.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_n8.json").read().strip().splitlines()[-1])
print("N=8 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"].get("value"), d["train_step"].get("allreduce"), d.get("clocks"))
PY
