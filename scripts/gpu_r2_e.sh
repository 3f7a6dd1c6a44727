#!/bin/bash
# round 2, GPU call E (2 GPUs): NCCL exchange parity + 2-rank bench
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -15
cat gpurun_out/dist_parity_n2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"; tail -c 500 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
print("N=2 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"].get("value"), d["train_step"].get("allreduce"))
PY
