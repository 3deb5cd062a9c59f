"""Worker for the multi-rank tests; launched under torch.distributed.run (tests/ only).

mode=gpu : every rank owns one GPU and one brick of a jittered mesh; results are gathered on
           rank 0 and compared with the CPU oracle's rank emulation on the same bricks.
mode=plan: host-only (gloo, no GPU): the exchange plans of all ranks must be mutually consistent.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def make_parts(world, kind):
    import smoothmesh_b200 as sm
    from smoothmesh_b200 import multi
    px, py, pz = multi.brick_dims(world)
    if kind == "hex":
        mesh = sm.Mesh.hex_block(4 * px, 4 * py, 3 * pz, hi=(px, py, 0.75 * pz)).jitter(0.3 / 4, 4242)
        return mesh.decompose(px, py, pz)
    if kind == "hexlayers":
        # layer treatment on all six walls, bricks cut through the layers (BASELINE config 3 style options)
        mesh = sm.Mesh.hex_block(8 * px, 7 * py, 6 * pz, hi=(px, py, pz)).jitter(0.2 / 8, 99)
        return mesh.decompose(px, py, pz)
    if kind == "prismlayers":
        # the reference's own test mesh (testcase/: extruded tri/quad surface, layers on the hole walls),
        # from the committed golden fixture
        d = np.load(os.path.join(ROOT, "tests", "golden", "testcase_layers.npz"))
        mesh = sm.Mesh.from_arrays(d["points"], d["face_offsets"], d["face_verts"], d["owner"], d["neighbour"],
                                   int(d["n_cells"]), d["patch_start"], d["patch_size"], d["patch_kind"])
        return mesh.decompose(world, method="rcb")
    if kind == "kelvin":
        mesh = sm.Mesh.kelvin(3, 1.0).jitter(0.15 * 2 ** 0.5 / 4, 77)
        return mesh.decompose(world, method="rcb")
    if kind == "boundary":
        # boundary point smoothing of a box onto a slightly larger box (corners, feature edges, surfaces)
        mesh = sm.Mesh.hex_block(4 * px, 4 * py, 3 * pz, hi=(1.2, 1.0, 0.9)).jitter(0.03, 5)
        return mesh.decompose(px, py, pz)
    raise ValueError(kind)


def boundary_geometry():
    from test_boundary_smoothing_oracle import box_geometry
    hi = (1.2, 1.0, 0.9)
    ip, ie, _, _ = box_geometry((0, 0, 0), hi, 3)
    c = np.array(hi) / 2
    tp, te, tc, tt = box_geometry(c - 1.08 * c, c + 1.08 * c, 3)
    return dict(init_edges=(ip, ie), target_edges=(tp, te), surface=(tc, tt))


def options(kind):
    """(Smoother / Oracle keyword options, iterations) of a multi-rank parity case."""
    if kind == "hexlayers":
        return dict(rel_tol=0.0, layer_patches=[1, 1, 1, 0, 1, 1], max_layers=3, layer_expansion_ratio=1.2), 12
    if kind == "boundary":
        return dict(rel_tol=0.0, layer_patches=[1, 0, 1, 0, 0, 0], max_layers=2), 10
    if kind == "prismlayers":  # testcase/run_parallel:22
        return dict(rel_tol=0.0, min_edge_length=0.01, max_step_length=0.002, min_angle_deg=15.0, max_angle_deg=160.0,
                    layer_patches=[1, 0, 0, 0, 0, 0, 0]), 12
    return dict(rel_tol=1e-3, min_angle_deg=70.0, max_angle_deg=110.0, total_min_freeze=0), 25


def proc_points(part):
    s, z, k = part.patches
    off, verts = part.face_offsets, part.face_verts
    pts = set()
    for a, n, kind in zip(s, z, k):
        if kind == 1:
            pts.update(verts[off[a]:off[a + n]].tolist())
    return np.array(sorted(pts), dtype=np.int32)


def main():
    import faulthandler
    faulthandler.dump_traceback_later(240, exit=True)  # a hung collective must not hang the test run
    import torch
    import torch.distributed as dist
    mode, kind = sys.argv[1], sys.argv[2]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    import smoothmesh_b200 as sm
    from smoothmesh_b200 import multi
    parts = make_parts(world, kind)
    mine = parts[rank]
    if mode == "plan":
        dist.init_process_group("gloo")
        local = proc_points(mine)
        gids = mine.point_global_id[local]
        counts, allg = multi.gather_shared(gids, dist)
        sp, sr = sm.exchange_plan(rank, world, local, gids, counts, allg)
        # brute force: points whose global label also appears on another rank
        expect = {}
        for r, p in enumerate(parts):
            if r != rank:
                common = np.intersect1d(p.point_global_id[proc_points(p)], gids)
                if len(common):
                    expect[r] = common
        got = {}
        for point, r in zip(sp, sr):
            got.setdefault(int(r), []).append(int(mine.point_global_id[point]))
        assert sorted(got) == sorted(expect), (got.keys(), expect.keys())
        for r in expect:
            assert got[r] == expect[r].tolist(), f"slot order towards rank {r} is not ascending global label"
        # both sides of every pair must hold the same sequence
        objs = [None] * world
        dist.all_gather_object(objs, got)
        for r in got:
            assert objs[r][rank] == got[r]
        print(f"rank {rank}: plan ok, {len(sp)} slots to {sorted(got)}", flush=True)
        dist.barrier()
        dist.destroy_process_group()
        return

    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    kw, iters = options(kind)
    g = sm.Smoother(mine, device=local_rank, **kw)
    okw = dict(kw)
    if kind == "boundary":
        # the collective set-up of boundary point smoothing sits between the communicator and the peer mapping
        geo = boundary_geometry()
        p2p = multi.init_comm(g, rank, world, dist,
                              before_p2p=lambda: g.enable_boundary_smoothing(geo, [1, 1, 1, 1, 0, 1], 0.15))
        okw.update(smoothing_patches=[1, 1, 1, 1, 0, 1], geometry=geo, internal_smoothing_blending_fraction=0.15)
    else:
        p2p = multi.init_comm(g, rank, world, dist)  # SMGPU_NO_P2P=1 keeps the NCCL exchanges
    if mode == "debug":
        # step-by-step comparison against the oracle's rank emulation (every rank runs the oracle)
        from oracle import Oracle
        o = Oracle([p.desc_arrays() for p in parts], **kw)
        shared = set(proc_points(mine).tolist())
        for it in range(8):
            log = g.iterate(1)
            n, nf, rs = o.iterate(1)
            gp, op = g.points(), o.get("points", rank)
            gf, of = g.frozen(), o.get("frozen", rank)
            badp = np.nonzero((gp != op).any(axis=1))[0]
            badf = np.nonzero(gf != of)[0]
            print(f"[r{rank}] it{it + 1} nf gpu={log.n_frozen} orc={nf} res gpu={log.residual} orc={rs} "
                  f"point diffs={[(int(i), i in shared, float(np.abs(gp[i] - op[i]).max())) for i in badp[:6]]} "
                  f"mask diffs={[(int(i), i in shared, int(gf[i]), int(of[i]), int(o.get('isInternal', rank)[i])) for i in badf[:6]]}",
                  flush=True)
        dist.barrier()
        os._exit(0)
    log = g.iterate(iters)
    res = dict(n=log.iterations, nf=log.n_frozen, res=log.residual, pts=g.points(), fz=g.frozen())
    multi.shutdown_comm(g, dist)
    allres = [None] * world
    dist.all_gather_object(allres, res)
    if rank == 0:
        from oracle import Oracle
        o = Oracle([p.desc_arrays() for p in parts], **okw)
        n, nf, rs = o.iterate(iters)
        for r in range(world):
            a = allres[r]
            assert a["n"] == n, (a["n"], n)
            assert np.array_equal(a["nf"], nf), (r, a["nf"], nf)
            assert np.array_equal(a["res"], rs), (r, a["res"], rs)
            assert np.array_equal(a["fz"], o.get("frozen", r)), f"rank {r}: freeze mask differs"
            assert np.array_equal(a["pts"], o.get("points", r)), f"rank {r}: points differ"
        print(f"multi-GPU parity ok: world={world} kind={kind} exchange={'peer-memory' if p2p else 'nccl'} iterations={n} nFrozen[-1]={nf[-1]} "
              f"frozen internal={sum(int(o.get('frozen', r).sum()) for r in range(world))}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except BaseException:
        import traceback
        traceback.print_exc()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(1)  # never wait in collectives / communicator teardown after a failure
