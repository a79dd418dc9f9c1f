"""Multi-GPU parity on real peer GPUs: spawns torchrun on scripts/fused_allgather_check.py with min(2, device_count) ranks --
the fused NVLink peer-store path and the NCCL all-gather path against a single-GPU run, every rank comparing every row.
Skipped on boxes with fewer than two GPUs (the driver's 8-GPU bench covers the same comparison inside bench.py)."""
import json
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_fused_and_allgather_paths_on_two_gpus(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    out = os.path.join(ROOT, "gpurun_out", "fused_allgather_check_pytest.json")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(ROOT, "scripts", "fused_allgather_check.py"), "--rows-per-rank", "8192",
           "--out", out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    res = json.load(open(out))
    assert res["all_ok"] and res["n_gpus"] == 2 and all(res["checks"].values()), res
