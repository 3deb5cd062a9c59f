"""Multi-rank paths: host-side exchange plan on CPU (gloo, world_size 2 and 4) and, on a box with
>= 2 GPUs, bit-exact parity of the NCCL path with the oracle's rank emulation."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "multi_worker.py")


def _run(world, mode, kind, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER, mode, kind]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


@pytest.mark.parametrize("world,kind", [(2, "hex"), (4, "hex"), (2, "kelvin")])
def test_exchange_plan_gloo(world, kind):
    out = _run(world, "plan", kind, 29611 + world)
    assert out.count("plan ok") == world


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["hex", "kelvin", "hexlayers", "prismlayers"])
def test_multi_gpu_parity(kind):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 4 if n >= 4 else 2
    out = _run(world, "gpu", kind, 29631)
    assert "multi-GPU parity ok" in out
