"""SURVEY 8(f)-4 groundwork: the oracle's restatement of boundary point smoothing (src/boundaryPointSmoothing.C,
CPU only -- the CUDA path does not have this feature yet) against the reference's own translation unit
(oracle/_ref/smoothMesh_ref, see tests/test_reference_build.py).

Synthetic cases: a jittered hex block whose boundary is smoothed onto OBJ geometry generated here -- the
twelve box edges as polylines sharing the eight corner vertices (corner points, twelve feature edge strings),
a target box scaled about the centre (so corners, feature edge points and surface points all move) or the
initial box itself, the box surface as twelve triangles -- with random patch selections, optional boundary
layer treatment, -internalSmoothingBlendingFraction, angle limits, serial and as rank processes.  nFrozenPoints
per iteration and the final points must agree bit for bit; where the reference aborts, the oracle must too."""
import os
import re
import shutil
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import smoothmesh_b200 as sm
from oracle import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden import write_obj  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "smoothMesh_ref")


def box_geometry(lo, hi, seg):
    """Edges (12 box edges as polylines of `seg` segments sharing the 8 corner vertices) and surface (12 triangles)."""
    lo, hi = np.array(lo, float), np.array(hi, float)
    corners = np.array([[lo[0] if i==0 else hi[0], lo[1] if j==0 else hi[1], lo[2] if k==0 else hi[2]] for k in (0,1) for j in (0,1) for i in (0,1)])
    cid = lambda i,j,k: i + 2*j + 4*k
    pts = [c for c in corners]; edges=[]
    pairs=[]
    for k in (0,1):
        for j in (0,1): pairs.append((cid(0,j,k),cid(1,j,k)))
    for k in (0,1):
        for i in (0,1): pairs.append((cid(i,0,k),cid(i,1,k)))
    for j in (0,1):
        for i in (0,1): pairs.append((cid(i,j,0),cid(i,j,1)))
    for a,b in pairs:
        prev=a
        for s in range(1,seg):
            pts.append(corners[a] + (corners[b]-corners[a])*s/seg); edges.append([prev,len(pts)-1]); prev=len(pts)-1
        edges.append([prev,b])
    quads=[(0,2,3,1),(4,5,7,6),(0,1,5,4),(2,6,7,3),(0,4,6,2),(1,3,7,5)]
    tris=[]
    for q in quads: tris += [[q[0],q[1],q[2]],[q[0],q[2],q[3]]]
    return np.array(pts), np.array(edges,dtype=np.int32), corners, np.array(tris,dtype=np.int32)

def one(seed):
    rng=np.random.default_rng(seed)
    nx,ny,nz = rng.integers(3,7,size=3)
    hi = tuple(rng.uniform(0.8,1.6,size=3))
    m = sm.Mesh.hex_block(int(nx),int(ny),int(nz),hi=hi)
    h = min(hi[0]/nx,hi[1]/ny,hi[2]/nz)
    m = m.jitter(float(rng.uniform(0.05,0.3))*h, int(rng.integers(1,10**6)))
    seg=int(rng.integers(2,6))
    ip, ie, _, _ = box_geometry((0,0,0), hi, seg)
    # target: the box scaled about its centre (morph), same topology
    sc = rng.uniform(0.9,1.15,size=3); c=np.array(hi)/2
    tlo = c - sc*c; thi = c + sc*c
    tp, te, tc, tt = box_geometry(tlo, thi, seg)
    okw=dict(rel_tol=float(rng.choice([0.0,0.02]))); cli=["-relTol", repr(okw["rel_tol"])]
    names=[sm.lib().smmesh_patch_name(m._h,i).decode() for i in range(m.n_patches)]
    sflags=[int(rng.random()<0.8) for _ in names]
    if not any(sflags): sflags[0]=1
    okw["smoothing_patches"]=sflags; cli += ["-smoothingPatches","("+" ".join(n for n,f in zip(names,sflags) if f)+")"]
    if rng.random()<0.5:
        lflags=[int(rng.random()<0.6) for _ in names]
        if any(lflags):
            okw["layer_patches"]=lflags; cli += ["-layerPatches","("+" ".join(n for n,f in zip(names,lflags) if f)+")"]
            mlay=int(rng.integers(1,4)); okw["max_layers"]=mlay; cli += ["-maxLayers",str(mlay)]
    if rng.random()<0.4:
        f=float(rng.uniform(0.0,0.5)); okw["internal_smoothing_blending_fraction"]=f; cli += ["-internalSmoothingBlendingFraction", repr(f)]
    if rng.random()<0.5:
        mn=float(rng.uniform(10,60)); mx=float(rng.uniform(120,175)); okw["min_angle_deg"]=mn; okw["max_angle_deg"]=mx; cli += ["-minAngle",repr(mn),"-maxAngle",repr(mx)]
    use_target = rng.random()<0.7
    geo=dict(init_edges=(ip,ie), target_edges=(tp,te) if use_target else (ip,ie), surface=(tc if use_target else box_geometry((0,0,0),hi,seg)[2], tt))
    if not use_target:
        geo["surface"]=(box_geometry((0,0,0),hi,seg)[2], tt)
    okw["geometry"]=geo
    iters=int(rng.integers(3,10))
    par=None
    if rng.random()<0.4:
        par=(int(rng.integers(1,3)),int(rng.integers(1,3)),int(rng.integers(1,3)))
        if par==(1,1,1): par=(2,1,1)
    tmp=tempfile.mkdtemp(prefix="fzb_")
    try:
        m.write(tmp+"/constant/polyMesh"); os.makedirs(tmp+"/system"); os.makedirs(tmp+"/constant/geometry")
        open(tmp+"/system/controlDict","w").write("startFrom startTime;\nstartTime 0;\ndeltaT 1;\nwriteFormat binary;\n")
        write_obj(tmp+"/constant/geometry/initEdges.obj", ip, ie, None, "initEdges")
        if use_target: write_obj(tmp+"/constant/geometry/targetEdges.obj", tp, te, None, "targetEdges")
        write_obj(tmp+"/constant/geometry/targetSurfaces.obj", geo["surface"][0], None, tt, "targetSurfaces")
        args=[REF,"-case",tmp,"-centroidalIters",str(iters)]+cli
        if par:
            parts=m.decompose(*par); sm.Mesh.write_decomposed(parts,tmp,binary=True); args.insert(3,"-parallel")
        r=subprocess.run(args,capture_output=True,text=True,timeout=300)
        log=re.findall(r"Smoothing iteration=(\d+) nFrozenPoints=(\d+) residual=(\S+)", r.stdout)
        oerr=None
        try:
            if par:
                parts=[sm.Mesh.read_processor(tmp,k) for k in range(len(parts))]
                o=Oracle([p.desc_arrays() for p in parts], libm=True, **okw)
            else:
                o=Oracle(m.desc_arrays(), libm=True, **okw)
            n,nf,res=o.iterate(iters)
        except RuntimeError as e:
            oerr=str(e)
        if r.returncode!=0 or oerr:
            ok=(r.returncode!=0) and bool(oerr)
            return ok, f"seed {seed} par={par} both-fail={ok} ref_rc={r.returncode} oracle_err={oerr} ref_tail={r.stdout[-300:]} {r.stderr[-200:]}"
        ok=[int(b) for _,b,_ in log]==nf.tolist()
        if par:
            for k,p in enumerate(parts):
                p.read_points(f"{tmp}/processor{k}/{n}/polyMesh/points"); ok = ok and np.array_equal(p.points,o.get("points",rank=k))
        else:
            out=sm.Mesh.read(tmp+"/constant/polyMesh"); out.read_points(f"{tmp}/{n}/polyMesh/points"); ok = ok and np.array_equal(out.points,o.get("points"))
        moved = "enabled" if "Enabled boundary point smoothing" in r.stdout else "DISABLED"
        return ok, f"seed {seed} par={par} iters={n} nf={nf.tolist()} bsmooth={moved} ok={ok} opts={ {k:v for k,v in okw.items() if k!='geometry'} }"
    finally:
        shutil.rmtree(tmp,ignore_errors=True)


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/smoothMesh_ref not built")
@pytest.mark.parametrize("block", range(6))
def test_boundary_smoothing_restatement_agrees_with_the_reference_translation_unit(block):
    for seed in range(5 * block, 5 * block + 5):
        ok, msg = one(seed)
        assert ok, msg


def test_oracle_reproduces_testcase4_as_shipped():
    """tests/golden/testcase4_boundary.npz: testcase4/run_serial exactly as shipped (layer treatment on `walls`,
    boundary point smoothing of every patch onto constant/geometry/*.obj, 200 iterations), produced by the
    reference's own translation unit.  The oracle must give the same log and the same points, bit for bit."""
    d = np.load(os.path.join(ROOT, "tests", "golden", "testcase4_boundary.npz"))
    m = sm.Mesh.from_arrays(d["points"], d["face_offsets"], d["face_verts"], d["owner"], d["neighbour"], int(d["n_cells"]),
                            d["patch_start"], d["patch_size"], d["patch_kind"])
    geo = dict(init_edges=(d["init_edges_points"], d["init_edges_edges"]),
               target_edges=(d["target_edges_points"], d["target_edges_edges"]),
               surface=(d["target_surfaces_points"], d["target_surfaces_tris"]))
    for libm in (True, False):
        o = Oracle(m.desc_arrays(), libm=libm, layer_patches=[1], smoothing_patches=[1], geometry=geo,
                   layer_expansion_ratio=1.2, layer_edge_length=0.05, max_layers=3)
        n, nf, res = o.iterate(200)
        assert n == int(d["iterations"]) and np.array_equal(nf, d["n_frozen"])
        assert np.allclose(res, d["residual"], rtol=1e-5, atol=0)
        assert np.array_equal(o.get("points"), d["final_points"])


@pytest.mark.parametrize("seed", range(6))
def test_product_host_setup_matches_the_oracle(seed):
    """smoothmesh_b200/csrc/boundary.cpp (the product's one-time set-up of boundary point smoothing: edge strings,
    corner / feature-edge / smoothing-surface classes, corner targets, hop counts, inner-neighbour map) against the
    oracle's restatement, on the same synthetic box cases and on testcase4 as shipped."""
    rng = np.random.default_rng(1000 + seed)
    if seed == 0:
        d = np.load(os.path.join(ROOT, "tests", "golden", "testcase4_boundary.npz"))
        mesh = sm.Mesh.from_arrays(d["points"], d["face_offsets"], d["face_verts"], d["owner"], d["neighbour"], int(d["n_cells"]),
                                   d["patch_start"], d["patch_size"], d["patch_kind"])
        init, target = (d["init_edges_points"], d["init_edges_edges"]), (d["target_edges_points"], d["target_edges_edges"])
        surface, flags, lel = (d["target_surfaces_points"], d["target_surfaces_tris"]), [1], 0.05
    else:
        nx, ny, nz = rng.integers(3, 7, size=3)
        hi = tuple(rng.uniform(0.8, 1.6, size=3))
        mesh = sm.Mesh.hex_block(int(nx), int(ny), int(nz), hi=hi).jitter(0.1 * min(hi[0] / nx, hi[1] / ny, hi[2] / nz), int(seed))
        seg = int(rng.integers(2, 6))
        ip, ie, _, _ = box_geometry((0, 0, 0), hi, seg)
        c, sc = np.array(hi) / 2, rng.uniform(0.9, 1.15, size=3)
        tp, te, tc, tt = box_geometry(c - sc * c, c + sc * c, seg)
        init, target, surface = (ip, ie), (tp, te), (tc, tt)
        flags = [int(rng.random() < 0.8) for _ in range(6)]
        flags[int(rng.integers(0, 6))] = 1
        lel = -1.0
    o = Oracle(mesh.desc_arrays(), smoothing_patches=flags, layer_edge_length=lel,
               geometry=dict(init_edges=init, target_edges=target, surface=surface))
    b = mesh.boundary_setup(init, target, flags, layer_edge_length=lel)
    for mine, theirs in (("is_corner", "isCorner"), ("is_feature_edge", "isFeatureEdge"), ("is_smoothing_surface", "isSmoothingSurface"),
                         ("point_strings", "pointStrings"), ("hops_to_smoothing", "hopsToSmoothing"), ("point_to_inner", "pointToInner")):
        assert np.array_equal(b[mine], o.get(theirs)), mine
    corners = b["is_corner"] == 1
    assert np.array_equal(b["corner_points"][corners], o.get("cornerPoints")[corners])
    if seed > 0:
        assert corners.sum() == 8 and b["target_edge_strings"].max() == 11     # 8 box corners, 12 edge strings


def _restart_case(tmp, binary=True):
    hi = (1.0, 1.2, 0.9)
    mesh = sm.Mesh.hex_block(5, 4, 4, hi=hi).jitter(0.02, 3)
    ip, ie, _, _ = box_geometry((0, 0, 0), hi, 3)
    tp, te, tc, tt = box_geometry((-0.06, -0.05, -0.04), (1.07, 1.26, 0.95), 3)
    mesh.write(os.path.join(tmp, "constant", "polyMesh"))
    os.makedirs(os.path.join(tmp, "system"))
    os.makedirs(os.path.join(tmp, "constant", "geometry"))
    open(os.path.join(tmp, "system", "controlDict"), "w").write(
        "startFrom latestTime;\nstartTime 0;\ndeltaT 1;\nwriteFormat %s;\nwritePrecision 16;\n" % ("binary" if binary else "ascii"))
    write_obj(os.path.join(tmp, "constant", "geometry", "initEdges.obj"), ip, ie, None)
    write_obj(os.path.join(tmp, "constant", "geometry", "targetEdges.obj"), tp, te, None)
    write_obj(os.path.join(tmp, "constant", "geometry", "targetSurfaces.obj"), tc, None, tt)
    return mesh, dict(init_edges=(ip, ie), target_edges=(tp, te), surface=(tc, tt))


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/smoothMesh_ref not built")
def test_restart_uses_the_classification_lists_of_the_first_run(tmp_path):
    """The reference keeps isCornerPoint / isFeatureEdgePoint label lists with the mesh (src/smoothMesh.C:2039-2078)
    so that a second invocation -- testcase8/run_serial runs the tool twice -- classifies the moved boundary
    points as the first one did.  Two runs of the reference translation unit against two oracle runs."""
    tmp = str(tmp_path)
    mesh, geo = _restart_case(tmp)
    for run in (1, 2):
        r = subprocess.run([REF, "-case", tmp, "-centroidalIters", "6", "-relTol", "0"], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        assert ("Found corners and feature edges" in r.stdout) == (run == 2)
        assert "- Detected number of corner points: 8" in r.stdout
    assert sorted(os.listdir(os.path.join(tmp, "6"))) == ["isCornerPoint", "isFeatureEdgePoint", "polyMesh"]
    o1 = Oracle(mesh.desc_arrays(), libm=True, rel_tol=0.0, smoothing_patches=[1] * 6, geometry=geo)
    o1.iterate(6)
    out = sm.Mesh.read(os.path.join(tmp, "constant", "polyMesh"))
    out.read_points(os.path.join(tmp, "6", "polyMesh", "points"))
    assert np.array_equal(out.points, o1.get("points"))
    moved = sm.Mesh.read(os.path.join(tmp, "constant", "polyMesh"))
    moved.read_points(os.path.join(tmp, "6", "polyMesh", "points"))
    lists = dict(geo, is_corner_point=o1.get("isCorner").astype(np.int32), is_feature_edge_point=o1.get("isFeatureEdge").astype(np.int32))
    o2 = Oracle(moved.desc_arrays(), libm=True, rel_tol=0.0, smoothing_patches=[1] * 6, geometry=lists)
    o2.iterate(6)
    out.read_points(os.path.join(tmp, "12", "polyMesh", "points"))
    assert np.array_equal(out.points, o2.get("points"))
    # without the lists the moved corners would no longer be recognised
    o3 = Oracle(moved.desc_arrays(), libm=True, rel_tol=0.0, smoothing_patches=[1] * 6, geometry=geo)
    assert int(o3.get("isCorner").sum()) == 0 and int(o2.get("isCorner").sum()) == 8


@pytest.mark.parametrize("case", ["testcase5", "testcase7", "testcase8"])
def test_oracle_reproduces_shipped_boundary_cases(case):
    """testcase5/run_serial (layer treatment on `top`, boundary point smoothing of every patch, 500 iterations, eight
    corner points) and testcase8/run_serial exactly as shipped: fixtures produced by the reference's own translation
    unit (tests/golden/make_golden.py: shipped_boundary_cases); testcase7/run_serial likewise (31 361 points, 21 edge strings)."""
    d = np.load(os.path.join(ROOT, "tests", "golden", f"{case}_boundary.npz"))
    m = sm.Mesh.from_arrays(d["points"], d["face_offsets"], d["face_verts"], d["owner"], d["neighbour"], int(d["n_cells"]),
                            d["patch_start"], d["patch_size"], d["patch_kind"])
    target = "target_edges" if "target_edges_points" in d.files else "init_edges"
    geo = dict(init_edges=(d["init_edges_points"], d["init_edges_edges"]),
               target_edges=(d[target + "_points"], d[target + "_edges"]),
               surface=(d["target_surfaces_points"], d["target_surfaces_tris"]))
    okw = {str(k): float(v) for k, v in zip(d["opt_keys"], d["opt_vals"])}
    for k in ("max_layers", "min_layers"):
        if k in okw:
            okw[k] = int(okw[k])
    layer = d["layer_patches"].tolist() or None
    o = Oracle(m.desc_arrays(), layer_patches=layer, smoothing_patches=[1] * len(d["patch_start"]), geometry=geo, **okw)
    n, nf, res = o.iterate(int(d["cli"][list(d["cli"]).index("-centroidalIters") + 1]))
    assert n == int(d["iterations"]) and np.array_equal(nf, d["n_frozen"])
    assert np.array_equal(o.get("points"), d["final_points"])
