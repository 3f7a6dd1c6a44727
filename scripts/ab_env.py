"""Same-box A/B of environment knobs (read once per process, hence one subprocess per configuration):
    python scripts/ab_env.py RD_BN_REV=0 RD_BN_REV=6 "RD_BN_REV=6 RD_CONV_T=0" ...
Each configuration: captured cfg-5 train step (B=2, fp16 storage), 5 warm-up + 30 timed replays, twice, interleaved.
One JSON line per run."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import json, os, sys, torch
sys.path.insert(0, %r)
from rangedet_b200 import synth, train
from rangedet_b200.model_params import make_params
B, H, W = int(os.environ.get("AB_BATCH", "2")), 64, 2656
dev = torch.device("cuda", 0)
T = synth.rpn_targets(B, seed=500)
g = torch.Generator(device=dev).manual_seed(600)
data = torch.randn((B, 8, H, W), device=dev, generator=g)
coord = torch.from_numpy(synth.range_image_coords(B, seed=700)).to(dev)
kw = {}
if os.environ.get("AB_BWD_SUMS"):
    kw["fuse_bwd_sums"] = {"0": False, "1": True, "auto": "auto"}[os.environ["AB_BWD_SUMS"]]
step = train.GraphedTrainStep(make_params(seed=0, device=dev), B, H, W, lr=0.0125, device=dev, act_dtype=torch.float16, **kw)
step.set_targets(T)
for _ in range(5):
    step.train_step(data, coord)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        step.train_step(data, coord)
    b.record()
    torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b) / 10)
print(json.dumps({"ms_per_step": best}))
''' % ROOT

if __name__ == "__main__":
    configs = sys.argv[1:] or [""]
    for rep in range(2):
        for c in configs:
            env = dict(os.environ)
            for kv in c.split():
                k, v = kv.split("=", 1)
                env[k] = v
            out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=600)
            line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else json.dumps({"error": out.stderr[-300:]})
            r = json.loads(line)
            r.update(config=c, rep=rep)
            print(json.dumps(r), flush=True)
