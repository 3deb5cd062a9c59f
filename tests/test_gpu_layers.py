"""Prismatic boundary layer treatment (src/orthogonalBoundaryBlending.C, enabled by -layerPatches):
GPU path against the CPU oracle, stage by stage and over the loop."""
import re
import subprocess

import numpy as np
import pytest

import smoothmesh_b200 as sm
from oracle import Oracle

from meshes import hex_jittered

pytestmark = pytest.mark.gpu

# patch order of Mesh.hex_block: xMin, xMax, yMin, yMax, zMin, zMax
LAYER_CASES = {
    "flat_layers_zmin": (lambda: sm.Mesh.hex_block(6, 6, 8, hi=(1.0, 1.0, 0.4)).jitter(0.01, 3), [0, 0, 0, 0, 1, 0],
                         dict(layer_edge_length=0.02)),
    "two_patches_corner": (lambda: hex_jittered(7, 6, 6, 0.2, seed=8), [1, 0, 0, 1, 0, 0],
                           dict(max_layers=2, layer_expansion_ratio=1.2, layer_max_blending_fraction=0.5)),
    "all_walls_tight_angles": (lambda: hex_jittered(6, 6, 6, 0.3, seed=2), [1, 1, 1, 1, 1, 1],
                               dict(min_angle_deg=70.0, max_angle_deg=110.0, min_layers=2, max_layers=3)),
}


@pytest.mark.parametrize("case", list(LAYER_CASES))
def test_layer_stage_parity(case):
    build, flags, kw = LAYER_CASES[case]
    mesh = build()
    kw = dict(kw, rel_tol=0.0)
    g = sm.Smoother(mesh, layer_patches=flags, **kw)
    o = Oracle(mesh.desc_arrays(), layer_patches=flags, **kw)
    assert (o.get("hopsToLayer") >= 1).sum() > 0 and (o.get("pointToOuter") >= 0).sum() > 0
    n, nf, res = o.iterate(1)
    assert np.array_equal(g.op_layer_normals(), o.get("snapNormals")), "boundary point normals differ"
    g.op_cell_centres()
    g.op_predict()
    assert np.array_equal(g.op_layer_blend(), o.get("snapClamped")), "layer blend + second clamp differ"
    g.op_edge_constraints()
    fz = g.op_face_angle_constraint()
    assert np.array_equal(fz, o.get("snapFrozenFaceAngle"))
    gnf, gres = g.op_commit()
    assert gnf == nf[0] and gres == res[0]
    assert np.array_equal(g.points(), o.get("points"))


@pytest.mark.parametrize("case", list(LAYER_CASES))
def test_layer_loop_parity(case):
    build, flags, kw = LAYER_CASES[case]
    mesh = build()
    kw = dict(kw, rel_tol=1e-4)
    g = sm.Smoother(mesh, layer_patches=flags, **kw)
    o = Oracle(mesh.desc_arrays(), layer_patches=flags, **kw)
    n, nf, res = o.iterate(40)
    log = g.iterate(40)
    assert log.iterations == n and np.array_equal(log.n_frozen, nf) and np.array_equal(log.residual, res)
    assert np.array_equal(g.frozen(), o.get("frozen")) and np.array_equal(g.points(), o.get("points"))
    # the treatment must actually do something: compare with a run without layer patches
    plain = Oracle(mesh.desc_arrays(), **kw)
    plain.iterate(40)
    assert np.abs(plain.get("points") - o.get("points")).max() > 1e-4
    # a restart through set_points starts from fresh set-up normals, like a new invocation of the tool
    g.set_points(np.array(mesh.points))
    log2 = g.iterate(40)
    assert np.array_equal(log2.n_frozen, nf) and np.array_equal(g.points(), o.get("points"))


def test_layers_on_a_processor_mesh_need_the_collective_setup():
    # hop counts and set-up normals of a decomposed case are synchronised between the ranks
    # (src/orthogonalBoundaryBlending.C:124,185,363): iterating before smgpu_comm_init must fail, not guess
    parts = hex_jittered(6, 4, 4, 0.1).decompose(2, 1, 1)
    g = sm.Smoother(parts[0], layer_patches=[1, 1, 1, 1, 1, 1])
    with pytest.raises(sm.SmoothMeshError, match="smgpu_comm_init"):
        g.iterate(1)


def test_cli_layer_patches_wordre(tmp_path):
    mesh = sm.Mesh.hex_block(6, 6, 8, hi=(1.0, 1.0, 0.4)).jitter(0.01, 3)
    case = tmp_path / "case"
    mesh.write(case / "constant" / "polyMesh")
    (case / "system").mkdir()
    (case / "system" / "controlDict").write_text("startFrom latestTime;\ndeltaT 1;\nwriteFormat binary;\n")
    r = subprocess.run([sm.CLI_PATH, "-case", str(case), "-centroidalIters", "15", "-relTol", "0", "-layerPatches",
                        '("zM.n" notAPatch)', "-layerEdgeLength", "0.02", "-smoothingPatches", "()"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Enabled boundary layer treatment" in r.stdout
    o = Oracle(mesh.desc_arrays(), layer_patches=[0, 0, 0, 0, 1, 0], rel_tol=0.0, layer_edge_length=0.02)
    n, nf, res = o.iterate(15)
    lines = re.findall(r"Smoothing iteration=(\d+) nFrozenPoints=(\d+) residual=(\S+)", r.stdout)
    assert [int(b) for _, b, _ in lines] == nf.tolist()
    out = sm.Mesh.read(case / "constant" / "polyMesh")
    out.read_points(case / "15" / "polyMesh" / "points")
    assert np.array_equal(out.points, o.get("points"))
