"""Multi-GPU parity of the data-parallel exchange on the real kernels: torchrun, one rank per GPU, NCCL.
Needs >= 2 visible GPUs (run with `gpurun --gpus 2`); the host-side logic is covered on CPU by tests/test_dist_cpu.py."""
import json
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2])
def test_torchrun_nccl_exchange_bitexact(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs (this box has %d)" % (world, torch.cuda.device_count()))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "dist_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith("DIST_PARITY ")]
    assert r.returncode == 0 and lines, (r.stdout[-1500:], r.stderr[-3000:])
    res = json.loads(lines[-1][len("DIST_PARITY "):])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "dist_parity_n%d.json" % world), "w") as f:
        json.dump(res, f, indent=1)
    for k in ("init_differs", "bcast_args_equal", "bcast_aux_equal", "single_rank_recompute_bitexact", "allreduce_bitexact_sum",
              "grad_nonzero", "split_backward", "params_equal_after_update", "params_moved", "aux_drifted", "aux_equal_after_average",
              "lockstep_second_step"):
        assert res[k] is True, (k, res)
    assert res["aux_average_err"] < 1e-6, res
