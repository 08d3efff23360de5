"""The north-star multi-GPU workload on real GPUs (NCCL): one image ray-sharded over the ranks and all-gathered, the
SDF grid x-slab sharded and all-gathered, both bit-identical to the single-GPU result.  Needs >= 2 GPUs (skipped
otherwise); launched exactly like the driver launches bench.py (torch.distributed.run, 127.0.0.1 rendezvous)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_image_and_grid_equal_single_gpu(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-4000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("MGPU_RESULT")]
    assert line and "ok=1" in line[0], r.stdout[-4000:]
