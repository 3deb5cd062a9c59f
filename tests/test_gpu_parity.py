"""GPU parity: libsmgpu.so (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): freeze masks, per-iteration nFrozen and iteration count bit-exact;
positions within 1e-9 x bounding-box diagonal.  The library is built --fmad=false with the oracle's
acos, so positions and residuals are in fact required to be bitwise equal here.
"""
import numpy as np
import pytest

import smoothmesh_b200 as sm
from oracle import Oracle

from meshes import CASES

pytestmark = pytest.mark.gpu

OPTION_SETS = {
    "default": dict(),
    # testcase/run_serial:18 minus -layerPatches
    "testcase_opts": dict(min_angle_deg=15.0, max_angle_deg=160.0),
    # BASELINE config 2: -totalMinFreeze true, tighter angle window so the freeze paths are busy
    "total_min_freeze": dict(total_min_freeze=1, min_angle_deg=60.0, max_angle_deg=120.0),
    "no_constraints": dict(edge_angle_constraint=0, face_angle_constraint=0),
    "org_geometry": dict(geometry_variant=1),
    # narrow angle window: many active points, self/neighbour freezes and stack re-visits
    "tight_angles": dict(min_angle_deg=80.0, max_angle_deg=100.0),
}


def _pair(mesh, **kw):
    g = sm.Smoother(mesh, **kw)
    o = Oracle(mesh.desc_arrays(), **kw)
    p = g.params
    assert p.min_edge_length == o.prm.minEdgeLength and p.max_step_length == o.prm.maxStepLength
    return g, o


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("opts", list(OPTION_SETS))
def test_stage_parity_first_iteration(case, opts):
    mesh = CASES[case]()
    kw = dict(OPTION_SETS[opts], rel_tol=0.0)
    g, o = _pair(mesh, **kw)
    n, nf, res = o.iterate(1)
    assert n == 1
    cc = g.op_cell_centres()
    assert np.array_equal(cc, o.get("snapCellCtr")), "cell centres differ"
    npred = g.op_predict()
    assert np.array_equal(npred, o.get("snapClamped")), "predictor output differs"
    fz = g.op_edge_constraints()
    ref = o.get("snapFrozenEdgeAngle") if kw.get("edge_angle_constraint", 1) else o.get("snapFrozenEdgeLen")
    assert np.array_equal(fz, ref), "freeze mask after edge constraints differs"
    if kw.get("face_angle_constraint", 1):
        fz = g.op_face_angle_constraint()
        assert np.array_equal(fz, o.get("snapFrozenFaceAngle")), "freeze mask after face-angle constraint differs"
    gnf, gres = g.op_commit()
    assert gnf == nf[0]
    assert gres == res[0]
    assert np.array_equal(g.points(), o.get("points"))


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("opts", ["default", "total_min_freeze", "tight_angles"])
def test_loop_parity(case, opts):
    mesh = CASES[case]()
    kw = dict(OPTION_SETS[opts], rel_tol=1e-3)
    g, o = _pair(mesh, **kw)
    iters = 40
    n, nf, res = o.iterate(iters)
    log = g.iterate(iters)
    assert log.iterations == n, "iteration count at convergence differs"
    assert np.array_equal(log.n_frozen, nf)
    assert np.array_equal(log.residual, res)
    assert np.array_equal(g.frozen(), o.get("frozen"))
    pts, ref = g.points(), o.get("points")
    diag = np.linalg.norm(ref.max(0) - ref.min(0))
    assert np.abs(pts - ref).max() <= 1e-9 * diag  # the stated tolerance
    assert np.array_equal(pts, ref)  # and in fact bitwise


def test_edge_face_angles_match_oracle():
    mesh = CASES["hex_8x6x5_j45"]()
    g, o = _pair(mesh)
    mn, mx = g.op_edge_face_angles()
    assert np.array_equal(g.edges(), o.get("edges"))
    for e in range(0, len(mn), 7):
        omn, omx = o.edge_face_angles(e)
        assert mn[e] == omn and mx[e] == omx


def test_connectivity_matches_oracle():
    mesh = CASES["kelvin3_j20"]()
    g, o = _pair(mesh)
    for name in ("pointCells", "pointPoints", "pointEdges", "edgeFaces"):
        goff, gval = g.csr(name)
        ooff, oval = o.csr(name)
        assert np.array_equal(goff, ooff) and np.array_equal(gval, oval), name
    # edgeCells: same sets (row order is not observable)
    goff, gval = g.csr("edgeCells")
    ooff, oval = o.csr("edgeCells")
    assert np.array_equal(goff, ooff)
    for e in range(len(goff) - 1):
        assert sorted(gval[goff[e]:goff[e + 1]]) == sorted(oval[ooff[e]:ooff[e + 1]])


def test_uniform_block_known_answer():
    # K1: uniform hex block -> nothing moves, stop at iteration 1, nFrozen = boundary points
    n = 7
    g = sm.Smoother(sm.Mesh.hex_block(n, n, n))
    log = g.iterate(10)
    assert log.iterations == 1
    assert log.n_frozen[0] == (n + 1) ** 3 - (n - 1) ** 3
    assert log.residual[0] < 1e-9


def test_determinism_and_boundary_fixed():
    mesh = CASES["hex6_j25"]()
    runs = []
    for _ in range(2):
        g = sm.Smoother(mesh, rel_tol=0.0)
        g.iterate(15)
        runs.append((g.points(), g.frozen()))
    assert np.array_equal(runs[0][0], runs[1][0]) and np.array_equal(runs[0][1], runs[1][1])
    o = Oracle(mesh.desc_arrays())
    bnd = o.get("isInternal") == 0
    assert np.array_equal(runs[0][0][bnd], np.asarray(mesh.points)[bnd])  # K5


def test_no_gpu_fallback_message():
    # a bad device ordinal must fail loudly, never fall back
    with pytest.raises(sm.SmoothMeshError):
        sm.Smoother(CASES["hex6_j25"](), device=99)


@pytest.mark.parametrize("case", ["hex_8x6x5_j45", "kelvin3_j20"])
@pytest.mark.parametrize("env", ["SMGPU_NO_FILTERS", "SMGPU_NO_F32", "SMGPU_NO_TILES", "SMGPU_FORCE_TILES",
                                 "SMGPU_NO_FUSED_FILTER", "SMGPU_OLD_TILES", "SMGPU_POINT_TILES", "SMGPU_NO_SHARE_MASK"])
def test_literal_path_without_filters(case, env, monkeypatch):
    # SMGPU_NO_FILTERS=1 disables the guard-banded cosine-space filters so that every point /
    # edge takes the literal evaluation; SMGPU_NO_F32=1 disables only the single-precision first
    # level; SMGPU_NO_TILES=1 replaces the fused geometry kernel by the per-face + per-cell pair and
    # SMGPU_FORCE_TILES=1 uses it on polyhedral meshes too (default: all-quad / all-hex meshes only).
    # SMGPU_NO_FUSED_FILTER=1 keeps the per-edge face-angle filter (k_face_current) instead of the per-cell one
    # fused into the geometry tiles, SMGPU_OLD_TILES=1 the first-generation tile kernel.
    # SMGPU_POINT_TILES=1 runs the predictor / edge constraints on point tiles (shared-memory staging) instead of
    # the per-point gather kernels.
    # Every combination must reproduce the oracle bit for bit.
    monkeypatch.setenv(env, "1")
    mesh = CASES[case]()
    kw = dict(OPTION_SETS["tight_angles"], rel_tol=0.0)
    g, o = _pair(mesh, **kw)
    n, nf, res = o.iterate(10)
    log = g.iterate(10)
    assert np.array_equal(log.n_frozen, nf) and np.array_equal(log.residual, res)
    assert np.array_equal(g.frozen(), o.get("frozen")) and np.array_equal(g.points(), o.get("points"))


@pytest.mark.parametrize("case", ["hex_8x6x5_j45", "kelvin3_j20"])
def test_renumbered_storage_matches_oracle_on_renumbered_mesh(case):
    # params.renumber = 1 is "renumberMesh, then smoothMesh": the checker is the oracle on the
    # Morton-renumbered mesh; the library returns points / masks in the caller's numbering
    mesh = CASES[case]()
    renum, point_old_of_new, _ = mesh.renumber()
    kw = dict(OPTION_SETS["tight_angles"], rel_tol=0.0)
    g = sm.Smoother(mesh, renumber=1, **kw)
    o = Oracle(renum.desc_arrays(), **kw)
    n, nf, res = o.iterate(10)
    log = g.iterate(10)
    assert np.array_equal(log.n_frozen, nf) and np.array_equal(log.residual, res)
    assert np.array_equal(g.points()[point_old_of_new], o.get("points"))
    assert np.array_equal(g.frozen()[point_old_of_new], o.get("frozen"))


def test_mesh_quality_acceptance_after_smoothing():
    # BASELINE acceptance: checkMesh quality (max non-orthogonality, skewness, min angle) of the GPU
    # result within 1e-6 relative of the reference result; and smoothing must improve the mesh
    mesh = CASES["hex6_j25"]()
    before = mesh.quality()
    g, o = _pair(mesh)
    g.iterate(60)
    o.iterate(60)
    a, b = CASES["hex6_j25"](), CASES["hex6_j25"]()
    a.points[:] = g.points()
    b.points[:] = o.get("points")
    qa, qb = a.quality(), b.quality()
    for k in ("max_non_ortho", "max_skewness", "min_edge_angle"):
        assert abs(qa[k] - qb[k]) <= 1e-6 * max(abs(qb[k]), 1e-30), k
    assert qa["max_non_ortho"] < 0.2 * before["max_non_ortho"] and qa["min_edge_angle"] > before["min_edge_angle"]


@pytest.mark.parametrize("case,opts", [("slab", "default"), ("slab", "tight_angles"), ("far", "default"), ("far", "tight_angles")])
def test_fused_face_filter_is_size_independent(case, opts):
    """The fused face-angle filter works in single precision relative to a tile-local origin: its error budget
    must not depend on how large the mesh is relative to its cells, or on where the mesh lies.  A long thin slab
    (bounding box 250 cells long) and a block far from the coordinate origin: bit-exact against the oracle, and
    at the default angle limits the filter certifies (almost) everything."""
    if case == "slab":
        mesh = sm.Mesh.hex_block(250, 4, 4, hi=(250.0, 4.0, 4.0)).jitter(0.25, 7)
    else:
        mesh = sm.Mesh.hex_block(12, 10, 9, lo=(4000.0, -7000.0, 9000.0), hi=(4001.2, -6999.0, 9000.9)).jitter(0.025, 8)
    kw = dict(OPTION_SETS[opts], rel_tol=0.0)
    g, o = _pair(mesh, **kw)
    n, nf, res = o.iterate(8)
    log = g.iterate(8)
    assert np.array_equal(log.n_frozen, nf) and np.array_equal(log.residual, res)
    assert np.array_equal(g.frozen(), o.get("frozen")) and np.array_equal(g.points(), o.get("points"))
    st = g.filter_stats()
    assert st["fused"]
    if opts == "default":
        assert st["suspect_points"] <= 0.02 * mesh.n_points, st


def test_shared_reciprocal_division_is_ieee_division():
    """The geometry kernels divide the three components of a vector by one scalar with a single reciprocal
    (kernels.cuh divShared); every quotient must be the IEEE quotient bit for bit: 3 x 10^9 components."""
    import ctypes as C
    bad = C.c_int64(-1)
    L = sm.lib()
    L.smgpu_selftest_division.argtypes = [C.c_int32, C.c_uint64, C.c_int64, C.c_void_p]
    for seed in (1, 20261017):
        assert L.smgpu_selftest_division(0, seed, 500_000_000, C.byref(bad)) == 0
        assert bad.value == 0


@pytest.mark.parametrize("case", list(CASES))
def test_point_tile_kernels_loop_parity(case, monkeypatch):
    """The tiled forms of the predictor and the edge constraints (k_predict_tiles / k_edge_tiles, opt-in through
    SMGPU_POINT_TILES=1): same device functions behind shared-memory staging, bit-exact on every mesh kind."""
    monkeypatch.setenv("SMGPU_POINT_TILES", "1")
    mesh = CASES[case]()
    kw = dict(OPTION_SETS["tight_angles"], rel_tol=1e-3)
    g, o = _pair(mesh, **kw)
    assert g.tile_stats()["point_tiles"] > 0
    n, nf, res = o.iterate(25)
    log = g.iterate(25)
    assert log.iterations == n and np.array_equal(log.n_frozen, nf) and np.array_equal(log.residual, res)
    assert np.array_equal(g.frozen(), o.get("frozen")) and np.array_equal(g.points(), o.get("points"))


@pytest.mark.parametrize("n,jit,iters", [(24, 0.35, 10), (40, 0.3, 6)])
def test_dense_face_angle_worklist_matches_the_sequential_walk(n, jit, iters):
    """-minAngle 80 -maxAngle 100 on a jittered block: almost every point is active and the freeze decisions
    cascade through the whole mesh, the regime where the reference's stack walk (src/smoothMesh.C:1347-1434) is
    most order-dependent.  The parallel schedule of k_face_resolve (fixed point over freeze times) must reproduce
    the oracle's sequential walk bit for bit."""
    from meshes import hex_jittered
    mesh = hex_jittered(n, n, n, jit, seed=99)
    kw = dict(OPTION_SETS["tight_angles"], rel_tol=0.0)
    g, o = _pair(mesh, **kw)
    on, nf, res = o.iterate(iters)
    log = g.iterate(iters)
    assert g.filter_stats()["active_points"] > 0.5 * mesh.n_points
    assert np.array_equal(log.n_frozen, nf) and np.array_equal(log.residual, res)
    assert np.array_equal(g.frozen(), o.get("frozen")) and np.array_equal(g.points(), o.get("points"))


@pytest.mark.parametrize("n,frac", [(24, 0.2), (18, 0.4)])
def test_hex_dominant_mesh_uses_both_tile_paths_and_matches_oracle(n, frac):
    """Hexahedra and prisms in one mesh: the fused geometry kernel takes its fixed-stride path on the all-hex
    tiles and the offset tables elsewhere, inside one launch; points, flags and log bit-exact against the oracle."""
    from meshes import mixed_hex_prism_block
    mesh = mixed_hex_prism_block(n, frac=frac)
    tiles = mesh.geom_tiles()
    assert 0 < tiles["uniform_tile_cells"] < mesh.n_cells
    g, o = sm.Smoother(mesh, rel_tol=0.0), Oracle(mesh.desc_arrays(), rel_tol=0.0)
    assert g.filter_stats()["fused"]
    n_it, nf, res = o.iterate(4)
    log = g.iterate(4)
    assert np.array_equal(log.n_frozen, nf) and np.array_equal(log.residual, res)
    assert np.array_equal(g.points(), o.get("points")) and np.array_equal(g.frozen(), o.get("frozen"))


def test_share_a_cell_bits_of_the_point_records_match_the_literal_test(monkeypatch):
    """High-aspect-ratio layers: the midpoint-of-two-closest-points blend is active at most points, so the
    hasCommonCell short cut (src/smoothMesh.C:383) decides positions.  The predictor reads it from the bits
    k_share_mask put into the point records; SMGPU_NO_SHARE_MASK=1 keeps the literal row intersection."""
    from meshes import prism_layers
    mesh = prism_layers(n=10, layers=8, thickness=0.01, frac=0.3, seed=17)
    g = sm.Smoother(mesh, rel_tol=0.0)
    log = g.iterate(5)
    o = Oracle(mesh.desc_arrays(), rel_tol=0.0)
    n, nf, res = o.iterate(5)
    assert np.array_equal(log.n_frozen, nf) and np.array_equal(log.residual, res)
    assert np.array_equal(g.points(), o.get("points"))
    monkeypatch.setenv("SMGPU_NO_SHARE_MASK", "1")
    lit = sm.Smoother(mesh, rel_tol=0.0)
    lit.iterate(5)
    assert np.array_equal(g.points(), lit.points()) and np.array_equal(g.frozen(), lit.frozen())
