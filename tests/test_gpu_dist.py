"""Multi-rank frame assembly (the reference's `concat` + scatter, main.hs:83,95,98-107) collected by
`pytest -m gpu`: runs tests/dist_check.py under torchrun with two ranks.  On a box with >= 2 GPUs that is one rank per
GPU over NCCL; on a one-GPU box both ranks share cuda:0 (same CUDA-IPC data plane, gloo for the fence).  Rank 0's frame
from every exchange mode and the shared host frame must be bit-identical to a single-GPU render."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_frame_assembly_bit_identical():
    import torch
    n_dev = torch.cuda.device_count()
    env = dict(os.environ)
    if n_dev < 2:
        env["YAHR_DIST_CHECK_SAME_DEVICE"] = "1"
    env.pop("YAHR_B200_HOST_STREAM", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "dist_check_pytest_n2.log"), "w") as f:
            f.write(r.stdout + "\n--- stderr ---\n" + r.stderr[-4000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "dist_check OK" in r.stdout
    assert "bit-equal=False" not in r.stdout and "equal=False" not in r.stdout
