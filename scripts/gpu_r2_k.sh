#!/bin/bash
for r in 25; do
  echo "== rings $r"
  RD_CONVT_RINGS=$r timeout 300 python scripts/conv_t_ab.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['Cin'], d['W'], 'T', d['T_plain']['us'], 'P', d['P_plain']['us'])
"
done
RD_CONVT_PROF=1 RD_CONVT_RINGS=33 timeout 120 python - <<'PY' 2>&1 | tail -3
import sys, torch
sys.path.insert(0, ".")
from rangedet_b200 import ops
DT = torch.float16
g = torch.Generator(device="cuda").manual_seed(0)
x = ops.to_nhwc_padded(torch.randn((2, 128, 64, 2656), device="cuda", generator=g), dtype=DT)
wt = ops.pack_conv_weight(torch.randn((128, 128, 3, 3), device="cuda", generator=g) * 0.03, dtype=DT)
y = torch.zeros((2, 66, 2658, 128), device="cuda", dtype=DT)
for _ in range(3): ops.conv2d_nhwc(x, wt, relu=False, out=y)
torch.cuda.synchronize()
PY
