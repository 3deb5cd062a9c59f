"""The stand-alone `smoothMesh` executable (reference command-line surface + polyMesh I/O) on a GPU."""
import os
import re
import subprocess

import numpy as np
import pytest

import smoothmesh_b200 as sm
from oracle import Oracle

from meshes import hex_jittered

pytestmark = pytest.mark.gpu


def make_case(tmp_path, mesh, write_format="binary"):
    case = tmp_path / "case"
    mesh.write(case / "constant" / "polyMesh")
    (case / "system").mkdir()
    (case / "system" / "controlDict").write_text(
        "FoamFile { version 2.0; format ascii; class dictionary; object controlDict; }\n"
        f"startFrom latestTime;\nstartTime 0;\ndeltaT 1;\nwriteFormat {write_format};\nwritePrecision 12;\n"
        "timeFormat general;\ntimePrecision 6;\n")
    return case


def test_cli_matches_oracle_and_writes_time_directories(tmp_path):
    mesh = hex_jittered(7, 6, 5, 0.35, seed=21)
    case = make_case(tmp_path, mesh)
    opts = ["-centroidalIters", "12", "-relTol", "0", "-minAngle", "60", "-maxAngle", "120", "-totalMinFreeze", "true",
            "-writeInterval", "5", "-smoothingPatches", "()"]
    r = subprocess.run([sm.CLI_PATH, "-case", str(case)] + opts, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    o = Oracle(mesh.desc_arrays(), rel_tol=0.0, min_angle_deg=60.0, max_angle_deg=120.0, total_min_freeze=1)
    n, nf, res = o.iterate(12)
    lines = re.findall(r"Smoothing iteration=(\d+) nFrozenPoints=(\d+) residual=(\S+)", r.stdout)
    assert [int(a) for a, _, _ in lines] == list(range(1, 13))
    assert [int(b) for _, b, _ in lines] == nf.tolist()                       # the reference's log line, :2396
    assert np.allclose([float(c) for _, _, c in lines], res, rtol=1e-5)
    assert "Maximum centroidalIters reached, stopping." in r.stdout
    # writes at (i+1) % writeInterval == 0 && i > 0, and at the end (:2416): times 5, 10, 12
    assert sorted(p.name for p in case.iterdir() if p.name[0].isdigit()) == ["10", "12", "5"]
    final = sm.Mesh.read(case / "constant" / "polyMesh")
    final.read_points(case / "12" / "polyMesh" / "points")
    assert np.array_equal(final.points, o.get("points"))                      # binary writeFormat: bit exact
    # restart from latestTime (testcase8/run_serial:16-18 runs the tool twice)
    r2 = subprocess.run([sm.CLI_PATH, "-case", str(case), "-centroidalIters", "3", "-relTol", "0", "-minAngle", "60",
                         "-maxAngle", "120", "-totalMinFreeze", "true"], capture_output=True, text=True)
    assert r2.returncode == 0 and "Create mesh for time = 12" in r2.stdout
    assert (case / "15" / "polyMesh" / "points").exists()


REF_BIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "smoothMesh_ref")


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/smoothMesh_ref not built")
@pytest.mark.parametrize("layers", [False, True])
def test_cli_log_equals_the_reference_translation_unit(tmp_path, layers):
    """Drop-in check at the outermost boundary: the GPU executable and the reference's own main() (compiled
    against the OpenFOAM facade, oracle/_ref) on copies of one case -- same command line, same standard output
    line by line (set-up messages, parameter echo, every iteration line, write messages; only the banner, the
    GPU timing line and ClockTime are the tools' own) and bit-identical points files."""
    mesh = hex_jittered(7, 6, 5, 0.35, seed=21)
    opts = ["-centroidalIters", "12", "-relTol", "0", "-minAngle", "50", "-maxAngle", "130", "-writeInterval", "5",
            "-smoothingPatches", "()"]
    if layers:
        opts += ["-layerPatches", '("x.*" zMin)', "-maxLayers", "3", "-layerExpansionRatio", "1.25"]
    outs = {}
    for tool, binary in (("gpu", sm.CLI_PATH), ("ref", REF_BIN)):
        case = make_case(tmp_path / tool, mesh)
        r = subprocess.run([binary, "-case", str(case)] + opts, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        lines = [ln.rstrip() for ln in r.stdout.splitlines()]
        lines = [ln for ln in lines if not ln.startswith(("smoothMesh (smoothmesh_b200", "GPU iteration time", "ClockTime"))]
        while lines and lines[0] == "":
            lines.pop(0)
        outs[tool] = (lines, case)
    assert outs["gpu"][0] == outs["ref"][0]
    for t in ("5", "10", "12"):
        a = (outs["gpu"][1] / t / "polyMesh" / "points").read_bytes()
        b = (outs["ref"][1] / t / "polyMesh" / "points").read_bytes()
        assert a == b, f"points file of time {t} differs"


def test_cli_ascii_precision_and_reltol_stop(tmp_path):
    mesh = hex_jittered(5, 5, 5, 0.2, seed=4)
    case = make_case(tmp_path, mesh, write_format="ascii")
    r = subprocess.run([sm.CLI_PATH, "-case", str(case), "-centroidalIters", "500"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Residual reached relTol, stopping." in r.stdout
    o = Oracle(mesh.desc_arrays())
    n, nf, res = o.iterate(500)
    assert len(re.findall(r"Smoothing iteration=", r.stdout)) == n             # iteration count at convergence
    out = sm.Mesh.read(case / "constant" / "polyMesh")
    out.read_points(case / str(n) / "polyMesh" / "points")
    assert np.allclose(out.points, o.get("points"), rtol=0, atol=1e-11)       # precision max(10, writePrecision)


def test_cli_parallel_on_processor_directories(tmp_path):
    """`smoothMesh -parallel` on a decomposed case (testcase/run_parallel:19-22: decomposePar, then
    mpirun -np N smoothMesh -parallel): one GPU per processor directory (NCCL between host threads), or -- on a
    box with fewer GPUs than directories -- all processor meshes as an in-process group on one GPU; points written
    per processor."""
    mesh = hex_jittered(8, 6, 5, 0.35, seed=33)
    case = make_case(tmp_path, mesh)
    r = subprocess.run([sm.CLI_PATH, "-case", str(case), "-decompose", "(2 1 1)"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    opts = ["-centroidalIters", "10", "-relTol", "0", "-minAngle", "60", "-maxAngle", "120", "-totalMinFreeze", "true",
            "-writeInterval", "4", "-smoothingPatches", "()"]
    r = subprocess.run([sm.CLI_PATH, "-case", str(case), "-parallel"] + opts, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    parts = [sm.Mesh.read_processor(case, k) for k in range(2)]
    o = Oracle([p.desc_arrays() for p in parts], rel_tol=0.0, min_angle_deg=60.0, max_angle_deg=120.0, total_min_freeze=1)
    n, nf, res = o.iterate(10)
    lines = re.findall(r"Smoothing iteration=(\d+) nFrozenPoints=(\d+) residual=(\S+)", r.stdout)
    assert [int(b) for _, b, _ in lines] == nf.tolist()
    assert np.allclose([float(c) for _, _, c in lines], res, rtol=1e-5)
    for k in range(2):
        d = case / f"processor{k}"
        assert sorted(p.name for p in d.iterdir() if p.name[0].isdigit()) == ["10", "4", "8"]
        parts[k].read_points(d / "10" / "polyMesh" / "points")
        assert np.array_equal(parts[k].points, o.get("points", rank=k))
    # the same case with boundary layer treatment, as most of the reference's run_parallel scripts do
    # (testcase/run_parallel:22); '(".*")' also matches the processor patches, as the reference's patchSet would
    parts = [sm.Mesh.read_processor(case, k) for k in range(2)]
    r = subprocess.run([sm.CLI_PATH, "-case", str(case), "-parallel", "-time", "constant", "-centroidalIters", "6",
                        "-relTol", "0", "-layerPatches", '(".*")', "-maxLayers", "2", "-smoothingPatches", "()"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    o = Oracle([p.desc_arrays() for p in parts], rel_tol=0.0, layer_patches=[1] * 16, max_layers=2)
    n, nf, res = o.iterate(6)
    lines = re.findall(r"Smoothing iteration=(\d+) nFrozenPoints=(\d+) residual=(\S+)", r.stdout)
    assert [int(b) for _, b, _ in lines] == nf.tolist()
    for k in range(2):
        parts[k].read_points(case / f"processor{k}" / "6" / "polyMesh" / "points")
        assert np.array_equal(parts[k].points, o.get("points", rank=k))


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/smoothMesh_ref not built")
def test_cli_boundary_point_smoothing_testcase4_as_shipped(tmp_path):
    """testcase4/run_serial exactly as shipped -- layer treatment on the wall patch and boundary point smoothing
    of every patch onto constant/geometry/*.obj -- through both executables: same log, bit-identical points."""
    import sys
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, golden)
    from make_golden import write_obj
    d = np.load(os.path.join(golden, "testcase4_boundary.npz"))
    mesh = sm.Mesh.from_arrays(d["points"], d["face_offsets"], d["face_verts"], d["owner"], d["neighbour"], int(d["n_cells"]),
                               d["patch_start"], d["patch_size"], d["patch_kind"])
    cli = [str(x) for x in d["cli"]]
    cli[cli.index("-layerPatches") + 1] = "(patch0)"
    cli[cli.index("-centroidalIters") + 1] = "60"
    outs = {}
    for tool, binary in (("gpu", sm.CLI_PATH), ("ref", REF_BIN)):
        case = make_case(tmp_path / tool, mesh)
        (case / "constant" / "geometry").mkdir()
        for f, key in (("initEdges.obj", "init_edges"), ("targetEdges.obj", "target_edges"), ("targetSurfaces.obj", "target_surfaces")):
            write_obj(case / "constant" / "geometry" / f, d[key + "_points"], d[key + "_edges"], d[key + "_tris"], key)
        r = subprocess.run([binary, "-case", str(case)] + cli, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
        lines = [ln.rstrip() for ln in r.stdout.splitlines()]
        lines = [ln for ln in lines if not ln.startswith(("smoothMesh (smoothmesh_b200", "GPU iteration time", "ClockTime"))]
        while lines and lines[0] == "":
            lines.pop(0)
        outs[tool] = (lines, case)
    assert "Enabled boundary point smoothing" in outs["gpu"][0]
    assert "- Detected number of feature edge points: 80" in outs["gpu"][0]
    assert outs["gpu"][0] == outs["ref"][0]
    a = (outs["gpu"][1] / "60" / "polyMesh" / "points").read_bytes()
    b = (outs["ref"][1] / "60" / "polyMesh" / "points").read_bytes()
    assert a == b


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/smoothMesh_ref not built")
def test_cli_restart_with_classification_lists(tmp_path):
    """Two invocations in a row (testcase8/run_serial:17-19) with boundary point smoothing: the second one starts
    from latestTime and takes the corner / feature-edge classes from the isCornerPoint / isFeatureEdgePoint lists the
    first one wrote.  Same logs and bit-identical points as two invocations of the reference's own main()."""
    from test_boundary_smoothing_oracle import _restart_case
    res = {}
    for tool, binary in (("gpu", sm.CLI_PATH), ("ref", REF_BIN)):
        d = tmp_path / tool
        d.mkdir()
        _restart_case(str(d))
        logs = []
        for run in (1, 2):
            r = subprocess.run([binary, "-case", str(d), "-centroidalIters", "6", "-relTol", "0"], capture_output=True, text=True, timeout=300)
            assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
            lines = [ln.rstrip() for ln in r.stdout.splitlines()]
            lines = [ln for ln in lines if not ln.startswith(("smoothMesh (smoothmesh_b200", "GPU iteration time", "ClockTime"))]
            while lines and lines[0] == "":
                lines.pop(0)
            logs.append(lines)
        res[tool] = (logs, d)
    assert "Found corners and feature edges in isCornerPoint and isFeatureEdgePoint files" in res["gpu"][0][1]
    assert res["gpu"][0] == res["ref"][0]
    for t in ("6", "12"):
        assert (res["gpu"][1] / t / "polyMesh" / "points").read_bytes() == (res["ref"][1] / t / "polyMesh" / "points").read_bytes()
        for f in ("isCornerPoint", "isFeatureEdgePoint"):
            assert (res["gpu"][1] / t / f).read_bytes() == (res["ref"][1] / t / f).read_bytes()


def _write_obj(path, points, edges=None, tris=None):
    with open(path, "w") as f:
        for p in np.asarray(points, float):
            f.write("v %.17g %.17g %.17g\n" % tuple(p))
        for e in (edges if edges is not None else []):
            f.write("l %d %d\n" % (e[0] + 1, e[1] + 1))
        for t in (tris if tris is not None else []):
            f.write("f %d %d %d\n" % (t[0] + 1, t[1] + 1, t[2] + 1))


def test_cli_parallel_with_boundary_point_smoothing(tmp_path):
    """`smoothMesh -parallel` with constant/geometry/*.obj present: boundary point smoothing on the processor
    meshes (collective set-up, per-iteration synchronisations inside the interface exchange), points per processor
    bit-identical to the oracle's rank emulation."""
    from test_boundary_smoothing_oracle import box_geometry
    hi = (1.2, 1.0, 0.8)
    mesh = sm.Mesh.hex_block(6, 5, 4, hi=hi).jitter(0.03, 3)
    case = make_case(tmp_path, mesh)
    ip, ie, _, _ = box_geometry((0, 0, 0), hi, 3)
    c = np.array(hi) / 2
    tp, te, tc, tt = box_geometry(c - 1.1 * c, c + 1.1 * c, 3)
    (case / "constant" / "geometry").mkdir()
    _write_obj(case / "constant" / "geometry" / "initEdges.obj", ip, edges=ie)
    _write_obj(case / "constant" / "geometry" / "targetEdges.obj", tp, edges=te)
    _write_obj(case / "constant" / "geometry" / "targetSurfaces.obj", tc, tris=tt)
    r = subprocess.run([sm.CLI_PATH, "-case", str(case), "-decompose", "(2 1 1)"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([sm.CLI_PATH, "-case", str(case), "-parallel", "-centroidalIters", "8", "-relTol", "0"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Enabled boundary point smoothing" in r.stdout
    parts = [sm.Mesh.read_processor(case, k) for k in range(2)]
    # what the files hold (OBJ text round trip of the geometry)
    geo = {}
    for key, name in (("init_edges", "initEdges"), ("target_edges", "targetEdges"), ("surface", "targetSurfaces")):
        pts, edges, tris = [], [], []
        for line in (case / "constant" / "geometry" / f"{name}.obj").read_text().splitlines():
            w = line.split()
            if w[0] == "v":
                pts.append([float(x) for x in w[1:4]])
            elif w[0] == "l":
                edges.append([int(w[1]) - 1, int(w[2]) - 1])
            elif w[0] == "f":
                tris.append([int(x) - 1 for x in w[1:4]])
        geo[key] = (np.array(pts), np.array(tris if key == "surface" else edges, dtype=np.int32))
    o = Oracle([p.desc_arrays() for p in parts], rel_tol=0.0, smoothing_patches=[1] * 16, geometry=geo)
    n, nf, res = o.iterate(8)
    lines = re.findall(r"Smoothing iteration=(\d+) nFrozenPoints=(\d+) residual=(\S+)", r.stdout)
    assert [int(b) for _, b, _ in lines] == nf.tolist()
    for k in range(2):
        parts[k].read_points(case / f"processor{k}" / "8" / "polyMesh" / "points")
        assert np.array_equal(parts[k].points, o.get("points", rank=k))
