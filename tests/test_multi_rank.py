"""Multi-rank paths: host-side exchange plan on CPU (gloo, world_size 2 and 4); on any GPU box the exchange
kernels (k_shared_pack / k_shared_merge / k_frozen_pack / k_frozen_or / k_finish_iter) through an in-process
group of N processor meshes on one device, bit-exact against the oracle's rank emulation; and, on a box with
>= 2 GPUs, the same parity over NCCL."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "multi_worker.py")


def _run(world, mode, kind, port, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER, mode, kind]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600, env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


@pytest.mark.parametrize("world,kind", [(2, "hex"), (4, "hex"), (2, "kelvin")])
def test_exchange_plan_gloo(world, kind):
    out = _run(world, "plan", kind, 29611 + world)
    assert out.count("plan ok") == world


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["peer-memory", "nccl"])
@pytest.mark.parametrize("kind", ["hex", "kelvin", "hexlayers", "prismlayers", "boundary"])
def test_multi_gpu_parity(kind, exchange):
    """One process per GPU: the peer-memory exchange (kernels store into the neighbours' mapped buffers over
    NVLink, flags with release / acquire) and the NCCL exchange, both bit-exact against the oracle's rank emulation."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 4 if n >= 4 else 2
    out = _run(world, "gpu", kind, 29631, env={"SMGPU_NO_P2P": "1"} if exchange == "nccl" else None)
    assert "multi-GPU parity ok" in out and f"exchange={exchange}" in out


@pytest.mark.gpu
@pytest.mark.parametrize("world,kind", [(2, "hex"), (4, "hex"), (8, "hex"), (3, "kelvin"), (4, "hexlayers"),
                                        (3, "prismlayers")])
def test_group_parity_one_gpu(world, kind):
    """N ranks as an in-process group on ONE device (smgpu_group_*): the interface records, the three-stage
    closest-point merge, the freeze-flag OR and the reduced statistics equal the oracle's rank emulation bit for
    bit -- iteration count, nFrozenPoints and residual per iteration, every rank's mask and points."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import multi_worker
    import smoothmesh_b200 as sm
    from oracle import Oracle
    parts = multi_worker.make_parts(world, kind)
    kw, iters = multi_worker.options(kind)
    members = [sm.Smoother(p, device=0, **kw) for p in parts]
    grp = sm.Group(members)
    log = grp.iterate(iters)
    o = Oracle([p.desc_arrays() for p in parts], **kw)
    n, nf, rs = o.iterate(iters)
    assert log.iterations == n
    assert np.array_equal(log.n_frozen, nf), (log.n_frozen, nf)
    assert np.array_equal(log.residual, rs)
    shared = 0
    for r, g in enumerate(members):
        assert np.array_equal(g.frozen(), o.get("frozen", r)), f"rank {r}: freeze mask differs"
        assert np.array_equal(g.points(), o.get("points", r)), f"rank {r}: points differ"
        shared += len(g.comm_local_shared())
    assert shared > 0 and log.launches > 0
    # a second run on the same group continues from the committed mesh, like the oracle
    log2 = grp.iterate(3)
    n2, nf2, rs2 = o.iterate(3)
    assert np.array_equal(log2.n_frozen, nf2) and np.array_equal(log2.residual, rs2)
    for r, g in enumerate(members):
        assert np.array_equal(g.points(), o.get("points", r))
    grp.close()
    for g in members:
        g.close()


@pytest.mark.gpu
def test_group_refuses_misuse():
    import smoothmesh_b200 as sm
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import multi_worker
    parts = multi_worker.make_parts(2, "hex")
    members = [sm.Smoother(p, device=0, rel_tol=0.0) for p in parts]
    grp = sm.Group(members)
    with pytest.raises(sm.SmoothMeshError, match="smgpu_group_iterate"):
        members[0].iterate(1)
    with pytest.raises(sm.SmoothMeshError, match="fresh handles"):
        sm.Group(members)
    grp.close()
