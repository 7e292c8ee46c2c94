"""Multi-GPU correctness (`-m gpu`, skipped with fewer than 2 GPUs): every exchange mode of
handheld_super_resolution.distributed.main_sharded against the single-GPU path, in real NCCL / NVLink processes
(tools/check_multi_gpu.py under torch.distributed.run).  The row-sharded merge must be bit-identical."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.gpu
def test_sharded_modes_match_single_gpu():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (run with gpurun --gpus 2)")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "check_multi_gpu.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    lines = [json.loads(x) for x in p.stdout.splitlines() if x.startswith("{")]
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    json.dump(lines, open(os.path.join(out, "multi_gpu_check_n%d.json" % world), "w"), indent=1)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert len(lines) == 12 and all(x["ok"] for x in lines)
    assert all(x["max_abs_diff"] == 0.0 for x in lines if x["mode"] == "rows")
