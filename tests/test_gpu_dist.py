"""Multi-GPU (NCCL) correctness inside the `-m gpu` suite: launches tests/dist_check_gpu.py under torchrun
with one rank per GPU when the box has at least two GPUs (skipped on a 1-GPU box; the gloo world-2/3
tests in tests/test_parallel_gloo.py cover the same host logic on CPU)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 8])
def test_partitioned_path_over_nccl_matches_single_gpu_and_f64(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs, box has %d" % (world, torch.cuda.device_count()))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "tests", "dist_check_gpu.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert lines, p.stderr[-3000:]
    rep = json.loads(lines[-1])
    assert p.returncode == 0 and rep["ok"], rep
    for mode in ("allgather", "alltoall", "allgather+agg"):
        assert rep["modes"][mode]["exchange"] == mode.split("+")[0] and rep["modes"][mode]["ok"]
