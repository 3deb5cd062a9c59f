"""Multi-GPU plumbing: one process per GPU, torch.distributed for the start-up handshake,
NCCL inside libsmgpu.so for the per-iteration interface exchanges.

The reference decomposes with decomposePar and runs `mpirun -np N smoothMesh -parallel`
(testcase/run_parallel); here every rank creates its Smoother from its processor mesh and
`init_comm` performs the three start-up steps documented in include/smgpu.h.
"""
from __future__ import annotations

import os

import numpy as np


def brick_dims(world):
    """px,py,pz with px*py*pz == world, as cubic as possible (power-of-two friendly)."""
    dims = [1, 1, 1]
    n, f = world, 2
    factors = []
    while n > 1:
        while n % f == 0:
            factors.append(f)
            n //= f
        f += 1
    for k in sorted(factors, reverse=True):
        dims[dims.index(min(dims))] *= k
    return tuple(sorted(dims, reverse=True))


def weak_scaling_part(n, world, rank, jitter_frac, seed):
    """Rank's brick (n^3 cells) of the weak-scaling hex block; jitter keyed on global point labels."""
    import smoothmesh_b200 as sm
    px, py, pz = brick_dims(world)
    mesh = sm.Mesh.hex_block_part(n, n, n, px, py, pz, rank, hi=(float(px), float(py), float(pz)))
    return mesh.jitter(jitter_frac * (1.0 / n), seed)


def gather_shared(my_gids: np.ndarray, dist):
    """All-gather of the variable-length processor-point lists -> (counts, concatenation)."""
    import torch
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    world = dist.get_world_size()
    cnt = torch.tensor([len(my_gids)], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, cnt)
    counts = np.array([int(c.item()) for c in counts], dtype=np.int64)
    m = int(counts.max()) if world else 0
    buf = torch.zeros(max(m, 1), dtype=torch.int64, device=dev)
    buf[: len(my_gids)] = torch.from_numpy(np.ascontiguousarray(my_gids)).to(dev)
    bufs = [torch.zeros(max(m, 1), dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(bufs, buf)
    allg = np.concatenate([b.cpu().numpy()[: counts[r]] for r, b in enumerate(bufs)]) if world else np.zeros(0, np.int64)
    return counts, allg


def init_comm(smoother, rank, world, dist, p2p_exchange=True, before_p2p=None):
    """The start-up steps of include/smgpu.h on every rank; before_p2p: a collective step that must sit between the
    communicator and the peer mapping (smgpu_enable_boundary_smoothing).  Returns True when the iteration loop uses
    the peer-memory exchange, False for the NCCL exchanges."""
    import torch
    import smoothmesh_b200 as sm
    counts, allg = gather_shared(smoother.comm_local_shared(), dist)
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    # local half first; the ranks agree on its outcome before anything collective inside the library runs, so
    # a rank that fails here (mesh without point_global_id, a point shared by too many ranks) cannot leave the
    # others waiting in ncclCommInitRank
    err = None
    try:
        smoother.comm_prepare(rank, world, counts, allg)
    except sm.SmoothMeshError as e:
        err = e
    bad = torch.tensor([1 if err else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(bad, op=dist.ReduceOp.MAX)
    if int(bad.item()):
        raise err if err else sm.SmoothMeshError("smgpu_comm_prepare failed on another rank")
    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid = torch.tensor(list(sm.Smoother.comm_unique_id()), dtype=torch.uint8, device=dev)
    dist.broadcast(uid, src=0)
    smoother.comm_init(rank, world, bytes(uid.cpu().numpy().tolist()), counts, allg)
    if before_p2p is not None:
        before_p2p()
    p2p = False
    if p2p_exchange and backend == "nccl" and not os.environ.get("SMGPU_NO_P2P"):
        # peer-memory exchange for the iteration loop (include/smgpu.h: smgpu_comm_p2p_*); every rank must end up in
        # the same mode, so the outcome is agreed on before anybody iterates
        ok = 1
        try:
            mine = torch.tensor(list(smoother.comm_p2p_export()), dtype=torch.uint8, device=dev)
        except sm.SmoothMeshError:
            mine = torch.zeros(64, dtype=torch.uint8, device=dev)
            ok = 0
        allh = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(world)]
        dist.all_gather(allh, mine)
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()):
            try:
                smoother.comm_p2p_connect(b"".join(bytes(t.cpu().numpy().tolist()) for t in allh))
            except sm.SmoothMeshError:
                ok = 0
            flag = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()):
                p2p = True
            else:
                smoother.comm_p2p_disable()
    return p2p


def shutdown_comm(smoother, dist):
    """Collective tear-down: every rank unmaps its peers' exchange blocks, the ranks synchronise, and only then
    the handle (and with it the rank's own block) goes away."""
    try:
        smoother.comm_p2p_disable()
    except Exception:
        pass
    dist.barrier()
    smoother.close()
