"""Pins for the CPU oracle.  The reference ships no golden vectors and
cannot be built without OpenFOAM, so the oracle is pinned by known answers derivable from the
cited reference code, by an independent NumPy restatement, and by its libm-acos build."""
import math

import numpy as np
import pytest

import smoothmesh_b200 as sm
from oracle import Oracle, acos, edge_edge_angle
from oracle.oracle_np import NumpyOracle

from meshes import CASES, hex_jittered


def test_k1_uniform_block_is_a_fixed_point():
    n = 6
    o = Oracle(sm.Mesh.hex_block(n, n, n).desc_arrays())
    it, nf, res = o.iterate(10)
    assert it == 1 and nf[0] == (n + 1) ** 3 - (n - 1) ** 3
    assert res[0] < 1e-9


def test_k2_edge_edge_angle_range_and_nan_clamp():
    c = [0.0, 0.0, 0.0]
    assert edge_edge_angle(c, [1, 0, 0], [2, 0, 0]) == acos(0.99999)      # collinear: clamp at +MAX (:781-783)
    assert edge_edge_angle(c, [1, 0, 0], [-3, 0, 0]) == acos(-0.99999)    # opposite: clamp at -MAX
    assert abs(edge_edge_angle(c, [1, 0, 0], [0, 5, 0]) - math.pi / 2) < 1e-15
    assert abs(acos(0.99999) - 0.004472139) < 1e-8 and abs(acos(-0.99999) - 3.137120515) < 1e-8
    # zero-length edge -> NaN cosine -> std::min/std::max turn it into +MAX
    assert edge_edge_angle(c, [0, 0, 0], [1, 0, 0]) == acos(0.99999)


def test_sm_acos_within_one_ulp_of_libm():
    rng = np.random.default_rng(1)
    xs = np.concatenate([np.linspace(-0.99999, 0.99999, 4001), rng.uniform(-1, 1, 4000), [0.0, 0.5, -0.5, 1.0, -1.0]])
    for x in xs:
        a, b = acos(float(x)), math.acos(float(x))
        assert abs(a - b) <= np.spacing(b) if b > 0 else a == b
        assert acos(float(x), libm=True) == b


def test_k3_uniform_hex_face_angles_are_right_angles():
    o = Oracle(sm.Mesh.hex_block(3, 3, 3).desc_arrays())
    off, _ = o.csr("edgeCells")
    seen = set()
    for e in range(o.size("edges")):
        mn, mx = o.edge_face_angles(e)
        assert abs(mn - math.pi / 2) < 1e-12 and abs(mx - math.pi / 2) < 1e-12
        seen.add(int(off[e + 1] - off[e]))
    assert seen == {1, 2, 4}  # cube edge, flat boundary edge, interior edge


def test_k4_k5_k6_step_clamp_boundary_and_defaults():
    mesh = hex_jittered(6, 6, 6, 0.3, seed=5)
    o = Oracle(mesh.desc_arrays(), rel_tol=0.0)
    h = 1.0 / 6
    # K6: defaults from the minimum edge length (src/smoothMesh.C:1861-1865)
    assert o.prm.minEdgeLength == 0.5 * o.min_edge and o.prm.maxStepLength == 0.3 * o.prm.minEdgeLength
    assert o.min_edge < h
    x0 = np.array(mesh.points)
    o.iterate(1)
    blend, clamped = o.get("snapBlend"), o.get("snapClamped")
    d = np.linalg.norm(blend - x0, axis=1)
    step = np.linalg.norm(clamped - x0, axis=1)
    ms, rf = o.prm.maxStepLength, 0.5
    small = d <= ms
    assert np.allclose(step[small], rf * d[small], rtol=1e-12, atol=1e-18)      # K4: unclamped = relStepFrac * |d|
    assert np.allclose(step[~small], ms, rtol=1e-12) and (~small).any()        # K4: clamped = exactly maxStepLength
    # K5: boundary points never move and are always counted
    bnd = o.get("isInternal") == 0
    assert np.array_equal(o.get("points")[bnd], x0[bnd])
    _, nf, _ = o.iterate(3)
    assert (nf >= bnd.sum()).all()


def test_k7_scale_invariance_of_the_tiny_graded_block():
    # testcase8: one graded 3x3x3 block at 1e-8 scale must behave like the same block at unit scale
    def graded(scale):
        m = sm.Mesh.hex_block(3, 3, 3)
        p = m.points
        p[:] = (p ** 1.7) * scale
        return m
    a, b = Oracle(graded(1.0).desc_arrays()), Oracle(graded(1e-8).desc_arrays())
    na, fa, ra = a.iterate(50)
    nb, fb, rb = b.iterate(50)
    assert na == nb and np.array_equal(fa, fb)
    assert np.allclose(ra, rb, rtol=1e-6)
    assert np.allclose(a.get("points"), b.get("points") / 1e-8, rtol=0, atol=1e-9)


@pytest.mark.parametrize("case,opts", [("hex6_j25", dict()), ("layers_ar", dict()),
                                       ("hex_8x6x5_j45", dict(min_angle_deg=75.0, max_angle_deg=105.0)),
                                       ("kelvin3_j20", dict(min_angle_deg=60.0, max_angle_deg=120.0, total_min_freeze=1)),
                                       # boundary layer treatment (8f-2): all walls, and a subset with other options
                                       ("hex_8x6x5_j45", dict(layer_patches=[1, 1, 1, 1, 1, 1], max_layers=2)),
                                       ("hex6_j25", dict(layer_patches=[0, 1, 0, 0, 1, 1], max_layers=3, min_layers=2,
                                                         layer_expansion_ratio=1.2, layer_edge_length=0.05,
                                                         layer_max_blending_fraction=0.8, min_angle_deg=20.0)),
                                       ("layers_ar", dict(layer_patches=[1, 0, 0, 0, 0, 0], max_layers=4))])
def test_k8_independent_numpy_restatement_agrees(case, opts):
    mesh = CASES[case]()
    iters = 4
    o = Oracle(mesh.desc_arrays(), rel_tol=0.0, **opts)
    n1, f1, r1 = o.iterate(iters)
    p = NumpyOracle(mesh.desc_arrays(), rel_tol=0.0, **opts)
    n2, f2, r2 = p.iterate(iters)
    assert n1 == n2 and np.array_equal(f1, f2)
    assert np.allclose(r1, r2, rtol=1e-12)
    assert np.abs(o.get("points") - p.x).max() <= 1e-12
    assert np.array_equal(o.get("frozen").astype(bool), p.frozen)


@pytest.mark.parametrize("case", list(CASES))
def test_libm_acos_build_gives_the_same_masks(case):
    mesh = CASES[case]()
    kw = dict(rel_tol=0.0, min_angle_deg=70.0, max_angle_deg=110.0)
    a, b = Oracle(mesh.desc_arrays(), **kw), Oracle(mesh.desc_arrays(), libm=True, **kw)
    _, fa, ra = a.iterate(8)
    _, fb, rb = b.iterate(8)
    assert np.array_equal(fa, fb) and np.array_equal(a.get("frozen"), b.get("frozen"))
    assert np.abs(a.get("points") - b.get("points")).max() <= 1e-13


def test_rank_emulation_matches_serial_when_constraints_are_off():
    mesh = hex_jittered(8, 6, 4, 0.2, seed=11)
    kw = dict(rel_tol=0.0, edge_angle_constraint=0, face_angle_constraint=0)
    s = Oracle(mesh.desc_arrays(), **kw)
    parts = mesh.decompose(2, 2, 1)
    g = Oracle([p.desc_arrays() for p in parts], threads=4, **kw)
    s.iterate(6)
    g.iterate(6)
    ref = s.get("points")
    for r, p in enumerate(parts):
        assert np.abs(g.get("points", r) - ref[p.point_global_id]).max() <= 1e-12


def test_layer_setup_under_rank_emulation():
    """Boundary layer treatment on a decomposed mesh (src/orthogonalBoundaryBlending.C under -parallel).
    Derived from the cited code: the hop counts are max-synchronised after every Jacobi sweep (:124-130), so
    they equal the serial ones; interface copies of a point end every iteration with identical normals (sum
    :185, maxMagSqr :363) and positions; away from the interfaces the set-up normals are the serial ones."""
    mesh = hex_jittered(12, 10, 8, 0.2, seed=3)
    kw = dict(rel_tol=0.0, layer_patches=[1, 1, 1, 1, 1, 1], max_layers=3)
    s = Oracle(mesh.desc_arrays(), **kw)
    parts = mesh.decompose(2, 2, 1)
    g = Oracle([p.desc_arrays() for p in parts], **kw)
    hops = s.get("hopsToLayer")
    assert hops.max() == 4 and (hops == 0).sum() > 0        # maxLayers + 1 sweeps
    for r, p in enumerate(parts):
        assert np.array_equal(g.get("hopsToLayer", r), hops[p.point_global_id])
    s.iterate(3)
    g.iterate(3)
    owner = {}
    for r, p in enumerate(parts):
        nrm, pts = g.get("snapNormals", r), g.get("points", r)
        for i, gid in enumerate(p.point_global_id):
            if gid in owner:
                n0, x0 = owner[gid]
                assert np.array_equal(nrm[i], n0) and np.array_equal(pts[i], x0)
            else:
                owner[gid] = (nrm[i].copy(), pts[i].copy())
    # a layer point well inside rank 0 keeps the serial set-up normal (unit, pointing into the domain)
    serial_n = s.get("snapNormals")
    p0 = parts[0]
    x = np.asarray(mesh.points)[p0.point_global_id]
    inside = (x[:, 0] < 0.3) & (x[:, 1] < 0.3) & (hops[p0.point_global_id] >= 1)
    n0 = g.get("snapNormals", 0)
    assert inside.sum() > 10 and np.abs(n0[inside] - serial_n[p0.point_global_id][inside]).max() <= 1e-12
    mag = np.linalg.norm(n0[inside], axis=1)
    assert (np.abs(mag[mag > 0] - 1.0) < 1e-12).all()


def test_oracle_reports_reference_fatal_errors():
    arr = sm.Mesh.hex_block(2, 2, 2).desc_arrays()
    arr["patch_kind"] = arr["patch_kind"].copy()
    arr["patch_kind"][0] = 2  # empty patch -> src/smoothMesh.C:61-66
    with pytest.raises(RuntimeError, match="non-3D"):
        Oracle(arr)
