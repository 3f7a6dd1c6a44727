"""Does the weight-gradient stream overlap the BatchNorm passes when both leave room for each other on the SM?
Run once per environment setting (RD_WGRAD_SMEM_KB, RD_BN_UB are read once per process)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rangedet_b200 import synth, train
from rangedet_b200.model_params import make_params
H, W, B = 64, 2656, 2
dev = torch.device("cuda", 0)
T = synth.rpn_targets(B, seed=500)
g = torch.Generator(device=dev).manual_seed(600)
data = torch.randn((B, 8, H, W), device=dev, generator=g)
coord = torch.from_numpy(synth.range_image_coords(B, seed=700)).to(dev)
step = train.GraphedTrainStep(make_params(seed=0, device=dev), B, H, W, lr=0.0125, device=dev, act_dtype=torch.float16)
step.set_targets(T)
for _ in range(5):
    step.train_step(data, coord)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(30):
    step.train_step(data, coord)
b.record()
torch.cuda.synchronize()
print(json.dumps({"RD_WGRAD_SMEM_KB": os.environ.get("RD_WGRAD_SMEM_KB"), "RD_BN_UB": os.environ.get("RD_BN_UB"),
                  "ms_per_step": a.elapsed_time(b) / 30}), flush=True)
