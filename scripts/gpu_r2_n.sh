#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "wnms or postprocess or nms" 2>&1 | tail -6
timeout 300 python - <<'PY'
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
timeout 300 python -m pytest tests/test_postprocess.py tests/test_symbol.py -m gpu -q 2>&1 | tail -3
