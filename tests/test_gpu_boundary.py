"""Boundary point smoothing on the GPU against the oracle, which is itself pinned to the
reference's own translation unit (tests/test_boundary_smoothing_oracle.py): bit-exact iteration logs, freeze
masks and points on synthetic box cases (corners, feature edge strings, surfaces, partial patch selections,
with and without layer treatment) and on testcase4 exactly as shipped."""
import os

import numpy as np
import pytest

import smoothmesh_b200 as sm
from oracle import Oracle

from test_boundary_smoothing_oracle import box_geometry

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def synthetic(seed):
    rng = np.random.default_rng(seed)
    nx, ny, nz = rng.integers(3, 7, size=3)
    hi = tuple(rng.uniform(0.8, 1.6, size=3))
    h = min(hi[0] / nx, hi[1] / ny, hi[2] / nz)
    mesh = sm.Mesh.hex_block(int(nx), int(ny), int(nz), hi=hi).jitter(float(rng.uniform(0.05, 0.3)) * h, int(rng.integers(1, 10 ** 6)))
    seg = int(rng.integers(2, 6))
    ip, ie, _, _ = box_geometry((0, 0, 0), hi, seg)
    c, sc = np.array(hi) / 2, rng.uniform(0.9, 1.15, size=3)
    tp, te, tc, tt = box_geometry(c - sc * c, c + sc * c, seg)
    geo = dict(init_edges=(ip, ie), target_edges=(tp, te), surface=(tc, tt))
    flags = [int(rng.random() < 0.8) for _ in range(6)]
    flags[int(rng.integers(0, 6))] = 1
    kw = dict(rel_tol=0.0)
    layer = None
    if rng.random() < 0.5:
        layer = [int(rng.random() < 0.6) for _ in range(6)]
        kw["max_layers"] = int(rng.integers(1, 4))
    if rng.random() < 0.5:
        kw["min_angle_deg"], kw["max_angle_deg"] = float(rng.uniform(10, 60)), float(rng.uniform(120, 175))
    frac = float(rng.uniform(0.0, 0.5)) if rng.random() < 0.4 else 0.0
    return mesh, geo, flags, layer, kw, frac, int(rng.integers(4, 10))


def run_pair(mesh, geo, flags, layer, kw, frac, iters):
    g = sm.Smoother(mesh, layer_patches=layer, **kw)
    g.enable_boundary_smoothing(geo, flags, frac)
    o = Oracle(mesh.desc_arrays(), layer_patches=layer, smoothing_patches=flags, geometry=geo,
               internal_smoothing_blending_fraction=frac, **kw)
    n, nf, res = o.iterate(iters)
    log = g.iterate(iters)
    assert log.iterations == n
    assert np.array_equal(log.n_frozen, nf), (log.n_frozen, nf)
    assert np.array_equal(log.residual, res)
    assert np.array_equal(g.frozen(), o.get("frozen"))
    assert np.array_equal(g.points(), o.get("points"))
    return g, o


@pytest.mark.parametrize("seed", range(8))
def test_boundary_smoothing_matches_oracle_on_synthetic_boxes(seed):
    mesh, geo, flags, layer, kw, frac, iters = synthetic(seed)
    g, o = run_pair(mesh, geo, flags, layer, kw, frac, iters)
    moved = np.abs(g.points() - np.asarray(mesh.points)).max(axis=1) > 0
    assert moved[o.get("isInternal") == 0].any()            # boundary points do move


def test_testcase4_as_shipped():
    """testcase4/run_serial: layer treatment on `walls` + boundary point smoothing of every patch onto
    constant/geometry/*.obj, 200 iterations; the fixture was produced by the reference's own translation unit."""
    d = np.load(os.path.join(ROOT, "tests", "golden", "testcase4_boundary.npz"))
    mesh = sm.Mesh.from_arrays(d["points"], d["face_offsets"], d["face_verts"], d["owner"], d["neighbour"], int(d["n_cells"]),
                               d["patch_start"], d["patch_size"], d["patch_kind"])
    geo = dict(init_edges=(d["init_edges_points"], d["init_edges_edges"]),
               target_edges=(d["target_edges_points"], d["target_edges_edges"]),
               surface=(d["target_surfaces_points"], d["target_surfaces_tris"]))
    g = sm.Smoother(mesh, layer_patches=[1], layer_expansion_ratio=1.2, layer_edge_length=0.05, max_layers=3)
    g.enable_boundary_smoothing(geo, [1])
    log = g.iterate(200)
    assert log.iterations == int(d["iterations"]) and np.array_equal(log.n_frozen, d["n_frozen"])
    assert np.array_equal(g.points(), d["final_points"])


def test_boundary_smoothing_refusals():
    mesh = sm.Mesh.hex_block(4, 4, 4).jitter(0.02, 1)
    ip, ie, _, _ = box_geometry((0, 0, 0), (1, 1, 1), 2)
    tp, te, tc, tt = box_geometry((-3, -3, -3), (4, 4, 4), 2)       # perimeter far off: the reference's sanity check
    g = sm.Smoother(mesh)
    with pytest.raises(sm.SmoothMeshError, match="Perimeter"):
        g.enable_boundary_smoothing(dict(init_edges=(ip, ie), target_edges=(tp, te), surface=(tc, tt)), [1] * 6)
    parts = mesh.decompose(2, 1, 1)
    gp = sm.Smoother(parts[0])
    with pytest.raises(sm.SmoothMeshError, match="collective"):   # a processor mesh needs its communicator / group first
        gp.enable_boundary_smoothing(dict(init_edges=(ip, ie), target_edges=(ip, ie), surface=(tc, tt)), [1] * 6)


@pytest.mark.parametrize("case", ["testcase5", "testcase7", "testcase8"])
def test_shipped_boundary_cases(case):
    """testcase5/run_serial (500 iterations: layer treatment on `top`, boundary point smoothing with eight corner
    points), testcase7/run_serial (31 361 points, 21 edge strings, layer treatment on `walls`) and
    testcase8/run_serial exactly as shipped; fixtures produced by the reference's own translation unit."""
    d = np.load(os.path.join(ROOT, "tests", "golden", f"{case}_boundary.npz"))
    mesh = sm.Mesh.from_arrays(d["points"], d["face_offsets"], d["face_verts"], d["owner"], d["neighbour"], int(d["n_cells"]),
                               d["patch_start"], d["patch_size"], d["patch_kind"])
    target = "target_edges" if "target_edges_points" in d.files else "init_edges"
    geo = dict(init_edges=(d["init_edges_points"], d["init_edges_edges"]),
               target_edges=(d[target + "_points"], d[target + "_edges"]),
               surface=(d["target_surfaces_points"], d["target_surfaces_tris"]))
    okw = {str(k): float(v) for k, v in zip(d["opt_keys"], d["opt_vals"])}
    for k in ("max_layers", "min_layers"):
        if k in okw:
            okw[k] = int(okw[k])
    g = sm.Smoother(mesh, layer_patches=d["layer_patches"].tolist() or None, **okw)
    g.enable_boundary_smoothing(geo, [1] * len(d["patch_start"]))
    log = g.iterate(int(d["cli"][list(d["cli"]).index("-centroidalIters") + 1]))
    assert log.iterations == int(d["iterations"]) and np.array_equal(log.n_frozen, d["n_frozen"])
    assert np.array_equal(g.points(), d["final_points"])


def fine_box_surface(lo, hi, n):
    """The six faces of a box as 12 n^2 triangles (shared vertices along the box edges are duplicated per face)."""
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    pts, tris = [], []
    for axis in range(3):
        a1, a2 = (axis + 1) % 3, (axis + 2) % 3
        for side in (0, 1):
            base = len(pts)
            for j in range(n + 1):
                for i in range(n + 1):
                    p = np.zeros(3)
                    p[axis] = hi[axis] if side else lo[axis]
                    p[a1] = lo[a1] + (hi[a1] - lo[a1]) * i / n
                    p[a2] = lo[a2] + (hi[a2] - lo[a2]) * j / n
                    pts.append(p)
            for j in range(n):
                for i in range(n):
                    v = base + j * (n + 1) + i
                    tris += [[v, v + 1, v + n + 2], [v, v + n + 2, v + n + 1]]
    return np.array(pts), np.array(tris, dtype=np.int32)


def test_surface_ray_casts_through_the_bvh(monkeypatch):
    """The surface ray casts (findLine, src/boundaryPointSmoothing.C:702-736) walk a bounding volume hierarchy on
    the device; it only decides which triangles are tested, so the result must equal the visit-every-triangle
    search (SMGPU_NO_BVH=1) and the oracle bit for bit -- here on a target surface of 6 912 triangles, where rays
    through shared triangle edges and vertices (equal parameters on several triangles) are the common case."""
    hi = (1.2, 1.0, 0.9)
    mesh = sm.Mesh.hex_block(7, 6, 5, hi=hi).jitter(0.03, 5)
    ip, ie, _, _ = box_geometry((0, 0, 0), hi, 3)
    c = np.array(hi) / 2
    tp, te, _, _ = box_geometry(c - 1.08 * c, c + 1.08 * c, 3)
    tc, tt = fine_box_surface(c - 1.08 * c, c + 1.08 * c, 24)
    assert len(tt) == 6912
    geo = dict(init_edges=(ip, ie), target_edges=(tp, te), surface=(tc, tt))
    g, o = run_pair(mesh, geo, [1] * 6, None, dict(rel_tol=0.0), 0.0, 6)
    monkeypatch.setenv("SMGPU_NO_BVH", "1")
    b = sm.Smoother(mesh, rel_tol=0.0)
    b.enable_boundary_smoothing(geo, [1] * 6)
    b.iterate(6)
    assert np.array_equal(b.points(), g.points()) and np.array_equal(b.frozen(), g.frozen())


@pytest.mark.parametrize("seed,dims", [(0, (2, 1, 1)), (1, (2, 2, 1)), (2, (1, 2, 2)), (3, (3, 1, 1)), (5, (2, 2, 2)), (6, (2, 1, 2))])
def test_boundary_smoothing_on_processor_meshes(seed, dims):
    """Boundary point smoothing under the reference's -parallel semantics: the decomposed synthetic boxes as an
    in-process group on one GPU against the oracle's rank emulation (itself pinned to the reference's rank
    processes): the collective set-up (global mesh figures, synchronised hop counts and point normals) and the
    per-iteration synchronisations of src/boundaryPointSmoothing.C:660,668 and orthogonalBoundaryBlending.C:491
    carried by the interface records -- bit-exact logs, masks and points on every rank."""
    mesh, geo, flags, layer, kw, frac, iters = synthetic(seed)
    parts = mesh.decompose(*dims)
    members = [sm.Smoother(p, layer_patches=layer, device=0, **kw) for p in parts]
    grp = sm.Group(members)
    grp.enable_boundary_smoothing(geo, flags, frac)
    o = Oracle([p.desc_arrays() for p in parts], layer_patches=layer, smoothing_patches=flags, geometry=geo,
               internal_smoothing_blending_fraction=frac, **kw)
    n, nf, res = o.iterate(iters)
    log = grp.iterate(iters)
    assert log.iterations == n
    assert np.array_equal(log.n_frozen, nf), (log.n_frozen, nf)
    assert np.array_equal(log.residual, res)
    for r, g in enumerate(members):
        assert np.array_equal(g.frozen(), o.get("frozen", r)), f"rank {r}: freeze mask differs"
        assert np.array_equal(g.points(), o.get("points", r)), f"rank {r}: points differ"
    grp.close()
