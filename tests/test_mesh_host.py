"""Host-side polyMesh library: generators, file I/O round trips, decomposition, CLI surface."""
import os
import subprocess

import numpy as np
import pytest

import smoothmesh_b200 as sm

from meshes import hex_jittered
from oracle import Oracle


def test_hex_block_counts_and_numbering():
    m = sm.Mesh.hex_block(4, 3, 2)
    assert (m.n_points, m.n_cells) == (5 * 4 * 3, 24)
    assert m.n_internal_faces == 3 * 3 * 2 + 4 * 2 * 2 + 4 * 3 * 1
    assert m.n_faces == m.n_internal_faces + 2 * (3 * 2 + 4 * 2 + 4 * 3)
    # blockMesh numbering: point i + j(nx+1) + k(nx+1)(ny+1)
    assert np.allclose(m.points[1 + 5 * 2 + 20 * 1], [1 / 4, 2 / 3, 1 / 2])
    own, nei = m.owner, m.neighbour
    assert (nei > own[: len(nei)]).all()
    key = own[: len(nei)].astype(np.int64) * m.n_cells + nei
    assert (np.diff(key) > 0).all()  # upper-triangular order


def test_generic_builder_matches_structured_generator():
    nx, ny, nz = 3, 2, 2
    ref = sm.Mesh.hex_block(nx, ny, nz)
    pid = lambda i, j, k: i + j * (nx + 1) + k * (nx + 1) * (ny + 1)
    cells = []
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                q = lambda a, b, c: pid(i + a, j + b, k + c)
                cells.append([[q(0, 0, 0), q(0, 0, 1), q(0, 1, 1), q(0, 1, 0)], [q(1, 0, 0), q(1, 1, 0), q(1, 1, 1), q(1, 0, 1)],
                              [q(0, 0, 0), q(1, 0, 0), q(1, 0, 1), q(0, 0, 1)], [q(0, 1, 0), q(0, 1, 1), q(1, 1, 1), q(1, 1, 0)],
                              [q(0, 0, 0), q(0, 1, 0), q(1, 1, 0), q(1, 0, 0)], [q(0, 0, 1), q(1, 0, 1), q(1, 1, 1), q(0, 1, 1)]])
    m = sm.Mesh.from_cells(np.array(ref.points), cells)
    assert m.n_faces == ref.n_faces and m.n_internal_faces == ref.n_internal_faces
    assert np.array_equal(m.owner[: m.n_internal_faces], ref.owner[: ref.n_internal_faces])
    assert np.array_equal(m.neighbour, ref.neighbour)
    same = lambda a, b: sorted(a) == sorted(b)
    assert all(same(a, b) for a, b in zip(m.faces()[: m.n_internal_faces], ref.faces()[: ref.n_internal_faces]))


def test_kelvin_mesh_is_a_valid_polyhedral_mesh():
    m = sm.Mesh.kelvin(3)
    assert m.n_cells == 2 * 27
    sizes = np.diff(m.face_offsets)
    assert set(sizes.tolist()) == {4, 6}
    # every cell has 14 faces
    cnt = np.bincount(m.owner, minlength=m.n_cells) + np.bincount(m.neighbour, minlength=m.n_cells)
    assert (cnt == 14).all()


@pytest.mark.parametrize("binary", [False, True])
def test_polymesh_write_read_round_trip(tmp_path, binary):
    m = hex_jittered(4, 3, 3, 0.3)
    d = tmp_path / "constant" / "polyMesh"
    m.write(d, binary=binary, precision=17)
    r = sm.Mesh.read(d)
    assert np.array_equal(r.points, m.points)
    for a in ("face_offsets", "face_verts", "owner", "neighbour"):
        assert np.array_equal(getattr(r, a), getattr(m, a)), a
    assert r.patch_names == m.patch_names
    for x, y in zip(r.patches, m.patches):
        assert np.array_equal(x, y)


def test_reader_accepts_openfoam_grammar_variants(tmp_path):
    d = tmp_path / "polyMesh"
    sm.Mesh.hex_block(1, 1, 1).write(d)
    # comments anywhere, single-line lists, uniform list N{v} for owner (all faces owned by cell 0)
    owner = (d / "owner").read_text()
    head = owner[: owner.index("6\n(")]
    (d / "owner").write_text("// leading comment\n" + head + "/* block\n comment */ 6{0}\n")
    (d / "neighbour").write_text((d / "neighbour").read_text().replace("0\n(", "0()").replace("\n)\n", "\n"))
    m = sm.Mesh.read(d)
    assert m.n_cells == 1 and m.n_faces == 6 and m.n_internal_faces == 0


def test_reader_rejects_invalid_mesh(tmp_path):
    d = tmp_path / "polyMesh"
    sm.Mesh.hex_block(2, 1, 1).write(d)
    txt = (d / "neighbour").read_text()
    (d / "neighbour").write_text(txt.replace("1\n(\n1\n)", "1\n(\n0\n)"))
    with pytest.raises(sm.SmoothMeshError, match="neighbour"):
        sm.Mesh.read(d)


@pytest.mark.parametrize("dims,method", [((2, 2, 1), "bricks"), ((3, 1, 1), "bricks"), ((5,), "rcb")])
def test_decompose_produces_consistent_processor_meshes(dims, method):
    m = hex_jittered(6, 4, 3, 0.2)
    parts = m.decompose(*dims, method=method) if method == "bricks" else m.decompose(dims[0], method="rcb")
    assert sum(p.n_cells for p in parts) == m.n_cells
    all_cells = np.sort(np.concatenate([p.cell_global_id for p in parts]))
    assert np.array_equal(all_cells, np.arange(m.n_cells))
    for p in parts:
        assert np.array_equal(p.points, np.asarray(m.points)[p.point_global_id])
        s, z, k = p.patches
        assert (k[:6] == 0).all() and (k[6:] == 1).all()
    # every processor face appears once on each side, reversed
    nproc = sum(int(z[k == 1].sum()) for _, z, k in (p.patches for p in parts))
    cut = m.n_internal_faces - sum(p.n_internal_faces for p in parts)
    assert nproc == 2 * cut


def test_counter_rng_jitter_is_partition_independent():
    n = 4
    whole = sm.Mesh.hex_block(2 * n, n, n, hi=(2.0, 1.0, 1.0)).jitter(0.1 / n, 99)
    for r in range(2):
        part = sm.Mesh.hex_block_part(n, n, n, 2, 1, 1, r, hi=(2.0, 1.0, 1.0)).jitter(0.1 / n, 99)
        assert np.array_equal(part.points, np.asarray(whole.points)[part.point_global_id])


def test_cli_rejects_out_of_scope_features_loudly(tmp_path):
    case = tmp_path / "case"
    sm.Mesh.hex_block(2, 2, 2).write(case / "constant" / "polyMesh")
    (case / "system").mkdir()
    (case / "system" / "controlDict").write_text("deltaT 1;\nwriteFormat ascii;\n")
    # -parallel without processor directories fails loudly (boundary point smoothing itself is supported there)
    (case / "constant" / "geometry").mkdir()
    for f in ("targetSurfaces.obj", "initEdges.obj"):
        (case / "constant" / "geometry" / f).write_text("# empty\n")
    r = subprocess.run([sm.CLI_PATH, "-case", str(case), "-parallel"], capture_output=True, text=True)
    assert r.returncode != 0 and "no processor directories" in r.stderr
    # without a GPU the tool fails loudly (there is no CPU path)
    r = subprocess.run([sm.CLI_PATH, "-case", str(case)], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
    for f in ("targetSurfaces.obj", "initEdges.obj"):
        (case / "constant" / "geometry" / f).unlink()
    r = subprocess.run([sm.CLI_PATH, "-case", str(case), "-bogusOption", "1"], capture_output=True, text=True)
    assert r.returncode != 0 and "Invalid option" in r.stderr
    (case / "system" / "controlDict").write_text("deltaT 0;\n")
    r = subprocess.run([sm.CLI_PATH, "-case", str(case)], capture_output=True, text=True)
    assert r.returncode != 0 and "too small" in r.stderr


def test_quality_metrics_known_answers():
    q = sm.Mesh.hex_block(4, 3, 2, hi=(4.0, 3.0, 2.0)).quality()
    assert q["max_non_ortho"] < 1e-6 and q["max_skewness"] < 1e-12          # orthogonal uniform block
    assert abs(q["min_edge_angle"] - 90.0) < 1e-9 and abs(q["min_edge_length"] - 1.0) < 1e-12
    assert abs(q["min_volume"] - 1.0) < 1e-12
    # shear the block by 45 degrees in x-y: every y-normal face becomes 45 deg non-orthogonal
    m = sm.Mesh.hex_block(3, 3, 3)
    m.points[:, 0] += m.points[:, 1]
    q = m.quality()
    assert abs(q["max_non_ortho"] - 45.0) < 1e-9 and abs(q["min_edge_angle"] - 45.0) < 1e-9
    jq = hex_jittered(6, 6, 6, 0.3).quality()
    assert jq["max_non_ortho"] > 5 and jq["max_skewness"] > 0.05 and jq["min_volume"] > 0


@pytest.mark.parametrize("binary", [False, True])
def test_decomposed_case_round_trip(tmp_path, binary):
    """processor<k>/constant/polyMesh with pointProcAddressing / cellProcAddressing (decomposePar layout,
    what testcase/run_parallel:19-22 leaves on disk)."""
    m = hex_jittered(5, 4, 3, 0.2, seed=3)
    parts = m.decompose(2, 2, 1)
    sm.Mesh.write_decomposed(parts, tmp_path, binary=binary)
    for k, p in enumerate(parts):
        d = tmp_path / f"processor{k}" / "constant" / "polyMesh"
        for f in ("points", "faces", "owner", "neighbour", "boundary", "pointProcAddressing", "cellProcAddressing"):
            assert (d / f).exists(), f
        q = sm.Mesh.read_processor(tmp_path, k)
        assert np.array_equal(q.point_global_id, p.point_global_id)
        assert np.array_equal(q.cell_global_id, p.cell_global_id)
        assert np.array_equal(q.face_verts, p.face_verts) and np.array_equal(q.owner, p.owner)
        assert np.array_equal(q.neighbour, p.neighbour)
        for a, b in zip(q.patches, p.patches):
            assert np.array_equal(a, b)                          # processor patches keep their kind
        if binary:
            assert np.array_equal(q.points, p.points)
        else:
            assert np.allclose(q.points, p.points, rtol=0, atol=1e-15)
    # the boundary file names the neighbour rank of every processor patch
    txt = (tmp_path / "processor0" / "constant" / "polyMesh" / "boundary").read_text()
    assert "procBoundary0to1" in txt and "neighbProcNo" in txt and "myProcNo" in txt


def test_cli_decompose_utility_and_parallel_refusals(tmp_path):
    case = tmp_path / "case"
    m = hex_jittered(4, 4, 2, 0.2, seed=5)
    m.write(case / "constant" / "polyMesh")
    (case / "system").mkdir()
    (case / "system" / "controlDict").write_text("deltaT 1;\nwriteFormat binary;\n")
    r = subprocess.run([sm.CLI_PATH, "-case", str(case), "-decompose", "(2 1 1)"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    expect = m.decompose(2, 1, 1)
    for k in range(2):
        q = sm.Mesh.read_processor(case, k)
        assert np.array_equal(q.point_global_id, expect[k].point_global_id)
        assert np.array_equal(q.points, expect[k].points)
    r = subprocess.run([sm.CLI_PATH, "-case", str(case), "-decompose", "2"], capture_output=True, text=True)
    assert r.returncode != 0 and "(nx ny nz)" in r.stderr
    # -parallel without a GPU fails loudly (no CPU fallback); with fewer GPUs than processor directories the
    # processor meshes run as an in-process group on one device (tests/test_cli_gpu.py)
    r = subprocess.run([sm.CLI_PATH, "-case", str(case), "-parallel"], capture_output=True, text=True,
                       env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert r.returncode != 0 and "no CUDA device" in r.stderr


def test_geometry_tiles_cover_the_mesh():
    """Host tiling of the fused geometry kernel: smmesh_geom_tiles checks the invariants the kernel relies on
    (every cell in one tile, every face stored by one tile, slot references resolve to the cell's faces in
    OpenFOAM's accumulation order) and fails otherwise."""
    g = hex_jittered(16, 16, 16, 0.25).geom_tiles()
    assert g["tiles"] == 16 and g["max_faces"] == 896          # 8 x 8 x 4 bricks
    assert g["listed_faces"] < 1.2 * g["faces"]                # border faces are listed twice
    g = hex_jittered(7, 5, 3, 0.2).geom_tiles(max_cells=4, max_faces=20)   # ragged: budgets close tiles early
    assert g["tiles"] >= 27 and g["max_faces"] <= 20
    k = sm.Mesh.kelvin(3, 1.0).geom_tiles()                    # 14-faced cells: fewer cells per tile
    assert k["tiles"] >= 1 and k["max_faces"] <= 1024
    assert sm.Mesh.kelvin(2, 1.0).geom_tiles(max_cells=8, max_faces=10)["tiles"] == 0   # a cell does not fit


@pytest.mark.parametrize("binary", [False, True])
def test_label_list_and_obj_files(tmp_path, binary):
    """labelIOList files (isCornerPoint / isFeatureEdgePoint, src/smoothMesh.C:2039-2065) and the OBJ reader used for
    constant/geometry."""
    import ctypes as C
    L = sm.lib()
    L.smmesh_read_label_list.restype = C.c_int64
    L.smmesh_read_label_list.argtypes = [C.c_char_p, C.c_void_p, C.c_int64]
    L.smmesh_write_label_list.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_int32]
    for name, data in (("mixed", np.array([0, 1, 0, 0, 1, 1, 0], dtype=np.int32)), ("uniform", np.zeros(9, dtype=np.int32))):
        f = str(tmp_path / name).encode()
        assert L.smmesh_write_label_list(f, name.encode(), b"7", data.ctypes.data_as(C.c_void_p), data.size, int(binary)) == 0
        n = L.smmesh_read_label_list(f, None, 0)
        assert n == data.size
        back = np.full(n, -1, dtype=np.int32)
        L.smmesh_read_label_list(f, back.ctypes.data_as(C.c_void_p), n)
        assert np.array_equal(back, data)
        txt = open(f, "rb").read()
        assert b"object" in txt and name.encode() in txt
        if name == "uniform" and not binary:
            assert b"9{0}" in txt                      # OpenFOAM's compact form for uniform lists
    assert L.smmesh_read_label_list(str(tmp_path / "absent").encode(), None, 0) == -1
    obj = tmp_path / "g.obj"
    obj.write_text("# comment\no thing\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nl 1 2 3\nl 4 1\nf 1/1/1 2/2/1 3/3/1 4/4/1\n")
    L.smmesh_read_obj.argtypes = [C.c_char_p] + [C.c_void_p] * 6
    np_, ne, nt = C.c_int64(), C.c_int64(), C.c_int64()
    assert L.smmesh_read_obj(str(obj).encode(), C.byref(np_), None, C.byref(ne), None, C.byref(nt), None) == 0
    assert (np_.value, ne.value, nt.value) == (4, 3, 2)           # polyline -> 2 edges, +1; quad -> 2 triangles
    e, t = np.zeros((3, 2), np.int32), np.zeros((2, 3), np.int32)
    L.smmesh_read_obj(str(obj).encode(), None, None, None, e.ctypes.data_as(C.c_void_p), None, t.ctypes.data_as(C.c_void_p))
    assert e.tolist() == [[0, 1], [1, 2], [3, 0]] and t.tolist() == [[0, 1, 2], [0, 2, 3]]


@pytest.mark.parametrize("seed", range(8))
def test_layer_setup_matches_the_oracle(seed):
    """topology.cpp: buildLayerSetup (the product's one-time set-up of the boundary layer treatment) against the
    oracle's calculatePointHopsToBoundary / propagateOuterNeighInfo on random meshes and patch selections: hop
    counts, point-to-outer map, and which internal points end up with a propagated normal."""
    from oracle import Oracle
    rng = np.random.default_rng(500 + seed)
    if seed % 4 == 3:
        mesh = sm.Mesh.kelvin(3, 1.0).jitter(0.02, seed)
    else:
        nx, ny, nz = rng.integers(3, 9, size=3)
        mesh = hex_jittered(int(nx), int(ny), int(nz), float(rng.uniform(0.05, 0.4)), seed=int(seed))
    flags = [int(rng.random() < 0.6) for _ in range(mesh.n_patches)]
    flags[int(rng.integers(0, mesh.n_patches))] = 1
    max_layers = int(rng.integers(1, 6))
    L = mesh.layer_setup(flags, max_layers)
    o = Oracle(mesh.desc_arrays(), layer_patches=flags, max_layers=max_layers)
    assert np.array_equal(L["hops"], o.get("hopsToLayer"))
    assert np.array_equal(L["point_to_outer"], o.get("pointToOuter"))
    # the oracle's set-up normals are non-zero exactly where the product copies a boundary point's normal
    internal = o.get("isInternal").astype(bool)
    o.iterate(1)
    has_normal = np.abs(o.get("snapNormals")).sum(axis=1) > 0
    assert np.array_equal(has_normal[internal], (L["normal_source"] >= 0)[internal])


def test_bvh_ray_casts_equal_the_visit_every_triangle_search():
    """The triangle BVH of boundary.cpp (acceleration structure for the surface ray casts of boundary point
    smoothing) must return exactly what the reference definition returns -- same triangle, same hit point, bit for
    bit -- on random segments against a random triangle soup, a closed sphere-like surface (rays through shared
    edges and vertices included) and the target surface of testcase4."""
    import ctypes as C
    L = sm.lib()
    L.smmesh_ray_cast.argtypes = [C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32,
                                  C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(7)

    def cast(pts, tris, a, b, use_bvh):
        pts, tris = np.ascontiguousarray(pts, np.float64), np.ascontiguousarray(tris, np.int32)
        a, b = np.ascontiguousarray(a, np.float64), np.ascontiguousarray(b, np.float64)
        ht, hp = np.zeros(len(a), np.int32), np.zeros((len(a), 3))
        p = lambda x: x.ctypes.data_as(C.c_void_p)
        assert L.smmesh_ray_cast(len(pts), p(pts), len(tris), p(tris), len(a), p(a), p(b), use_bvh, p(ht), p(hp)) == 0
        return ht, hp

    surfaces = []
    soup = rng.uniform(-1, 1, size=(900, 3))
    surfaces.append((soup, np.arange(900).reshape(-1, 3)))
    # lat-long sphere: many triangles share vertices and edges
    nu, nv = 24, 12
    sp = np.array([[np.cos(2 * np.pi * i / nu) * np.sin(np.pi * j / nv), np.sin(2 * np.pi * i / nu) * np.sin(np.pi * j / nv),
                    np.cos(np.pi * j / nv)] for j in range(nv + 1) for i in range(nu)])
    st = []
    for j in range(nv):
        for i in range(nu):
            a, b = j * nu + i, j * nu + (i + 1) % nu
            st += [[a, b, a + nu], [b, b + nu, a + nu]]
    surfaces.append((sp, np.array(st)))
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "testcase4_boundary.npz"))
    surfaces.append((d["target_surfaces_points"], d["target_surfaces_tris"]))
    for pts, tris in surfaces:
        lo, hi = pts.min(axis=0) - 0.2, pts.max(axis=0) + 0.2
        a = rng.uniform(lo, hi, size=(4000, 3))
        b = rng.uniform(lo, hi, size=(4000, 3))
        # segments that start on mesh vertices / pass through vertices and edge midpoints: ties between triangles
        v = pts[rng.integers(0, len(pts), size=500)]
        a = np.concatenate([a, v - 0.3 * (v - pts.mean(axis=0)), v])
        b = np.concatenate([b, v + 0.3 * (v - pts.mean(axis=0)), v + rng.normal(size=(500, 3))])
        t0, p0 = cast(pts, tris, a, b, 0)
        t1, p1 = cast(pts, tris, a, b, 1)
        assert (t0 >= 0).sum() > 100
        assert np.array_equal(t0, t1) and np.array_equal(p0, p1)


@pytest.mark.parametrize("n,dims", [(3, (2, 1, 1)), (4, (2, 2, 1)), (4, (2, 2, 2)), (5, (3, 1, 2))])
def test_kelvin_bricks_generated_per_rank_form_the_global_mesh(n, dims):
    """BASELINE config 4 at its stated size is generated per rank (smmesh_gen_kelvin_part): the bricks must tile
    the global Kelvin mesh -- same cells, same points (matched through point_global_id, identical coordinates
    on every copy), every inter-brick face listed once by each side, and the same jitter on every copy."""
    px, py, pz = dims
    whole = sm.Mesh.kelvin(n, 1.0)
    parts = [sm.Mesh.kelvin_part(n, 1.0, px, py, pz, r) for r in range(px * py * pz)]
    assert sum(p.n_cells for p in parts) == whole.n_cells == 2 * n ** 3
    coords = {}
    for p in parts:
        for g, x in zip(p.point_global_id.tolist(), np.asarray(p.points).tolist()):
            assert coords.setdefault(g, x) == x
    assert len(coords) == whole.n_points
    assert sorted(map(tuple, coords.values())) == sorted(map(tuple, np.asarray(whole.points).tolist()))
    # faces: internal faces of the bricks + half the processor faces + wall faces = faces of the whole mesh
    n_int = sum(p.n_internal_faces for p in parts)
    n_proc = n_wall = 0
    pair = {}
    for r, p in enumerate(parts):
        s, z, k = p.patches
        for name, size, kind in zip(p.patch_names, z, k):
            if kind == sm.PATCH_PROCESSOR:
                n_proc += size
                other = int(name.split("to")[1])
                pair[(r, other)] = size
            else:
                n_wall += size
    assert all(pair[(b, a)] == v for (a, b), v in pair.items())
    assert n_int + n_proc // 2 + n_wall == whole.n_faces and n_wall == whole.n_faces - whole.n_internal_faces
    for p in parts:
        p.jitter(0.05, 11)
    moved = {}
    for p in parts:
        for g, x in zip(p.point_global_id.tolist(), np.asarray(p.points).tolist()):
            assert moved.setdefault(g, x) == x
    # the rank-emulating oracle accepts the bricks (interface points found through point_global_id)
    o = Oracle([p.desc_arrays() for p in parts], rel_tol=0.0)
    assert o.iterate(2)[0] == 2


def test_geometry_tiles_of_a_hex_dominant_mesh_keep_the_fast_path_per_tile():
    """A mesh of hexahedra and prisms: every tile whose faces are all quadrilaterals and whose cells are all
    hexahedra is on the kernel's fast path (fixed-stride reference copies, canonical hexahedron records), the other
    tiles go through the offset tables; smmesh_geom_tiles checks both representations against each other."""
    from meshes import mixed_hex_prism_block
    g = mixed_hex_prism_block(12).geom_tiles(max_cells=32, max_faces=200, max_points=200)
    assert 0 < g["uniform_tile_cells"] < 12 ** 3 + 4 * 12 * 12 and g["uniform_cell_edges"] == 0
    h = hex_jittered(8, 8, 8, 0.2).geom_tiles()
    assert h["uniform_tile_cells"] == 512 and h["uniform_cell_edges"] == 12


def test_setup_tables_do_not_depend_on_the_thread_count(tmp_path):
    """The set-up builds every connectivity table and the geometry tiles with per-thread row buffers, atomic
    cursors and parallel prefix sums (topology.cpp); the result must be the same for any number of threads.
    tools/setup_timing.cpp prints a fingerprint of every table."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-s", "build/setup_timing"], cwd=root, check=True, capture_output=True)
    from meshes import mixed_hex_prism_block
    mixed_hex_prism_block(7).write(str(tmp_path / "mixed" / "constant" / "polyMesh"))
    for args in (["hex", "19"], ["kelvin", "5"], ["dir", str(tmp_path / "mixed" / "constant" / "polyMesh")]):
        outs = []
        # the last entry: the runtime grants fewer threads than asked for (the row ranges are work items, not threads)
        for threads, extra in (("1", {}), ("3", {}), ("8", {}), ("8", {"OMP_THREAD_LIMIT": "3", "OMP_DYNAMIC": "true"})):
            env = dict(os.environ, OMP_NUM_THREADS=threads, **extra)
            r = subprocess.run([os.path.join(root, "build", "setup_timing")] + args, env=env, check=True, capture_output=True, text=True)
            assert "t.ecPair" in r.stdout and "T.hexRec" in r.stdout
            outs.append(r.stdout)
        assert outs[0] == outs[1] == outs[2] == outs[3], args
