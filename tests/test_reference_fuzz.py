"""Randomised comparison of oracle.cpp with the reference's own translation unit (oracle/_ref/smoothMesh_ref,
see tests/test_reference_build.py): seeded random meshes (hex blocks of random shape and jitter, flat
high-aspect-ratio layers, Kelvin cells), random option sets (angle limits, totalMinFreeze, constraints on/off,
relStepFrac, relTol, boundary layer treatment on random patch subsets with random layer options), serial and
as 2-8 rank processes on random decompositions.  nFrozenPoints per iteration and the final points of every
(processor) mesh must agree bit for bit; where the reference aborts (FatalError), the oracle must report an
error too."""
import os
import re
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

import smoothmesh_b200 as sm
from oracle import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "smoothMesh_ref")

pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/smoothMesh_ref not built")


def one(seed, verbose=False, force_kind=None):
    rng=np.random.default_rng(seed)
    kind = rng.choice(["hex","hex","hex","kelvin","flat"])
    if force_kind:
        kind = force_kind
    if kind=="hex":
        nx,ny,nz = rng.integers(3,8,size=3)
        hi = tuple(rng.uniform(0.5,2.0,size=3))
        m = sm.Mesh.hex_block(int(nx),int(ny),int(nz),hi=hi)
        h = min(hi[0]/nx,hi[1]/ny,hi[2]/nz)
        m = m.jitter(float(rng.uniform(0.05,0.49))*h, int(rng.integers(1,10**6)))
    elif kind=="tets":
        import sys
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from meshes import tet_block
        m = tet_block(int(rng.integers(2,5)), float(rng.uniform(0.05,0.25)), int(rng.integers(1,10**6)))
    elif kind=="kelvin":
        m = sm.Mesh.kelvin(int(rng.integers(2,4)),1.0).jitter(float(rng.uniform(0.05,0.3))*(2**0.5)/4, int(rng.integers(1,10**6)))
    else:
        n=int(rng.integers(4,7)); L=int(rng.integers(3,6)); th=float(rng.uniform(0.01,0.05))
        m = sm.Mesh.hex_block(n,n,L,hi=(1.0,1.0,th*L)).jitter(float(rng.uniform(0.05,0.4))*th, int(rng.integers(1,10**6)))
    okw={}; cli=[]
    def opt(name_o, name_c, val, fmt=str):
        okw[name_o]=val; cli.extend([name_c, fmt(val)])
    opt("rel_tol","-relTol", float(rng.choice([0.0,0.0,0.02,0.2])), repr)
    if rng.random()<0.7:
        mn=float(rng.uniform(10,85)); mx=float(rng.uniform(95,175))
        opt("min_angle_deg","-minAngle",mn,repr); opt("max_angle_deg","-maxAngle",mx,repr)
    if rng.random()<0.4: okw["total_min_freeze"]=1; cli+=["-totalMinFreeze","true"]
    if rng.random()<0.2: okw["edge_angle_constraint"]=0; cli+=["-edgeAngleConstraint","false"]
    if rng.random()<0.2: okw["face_angle_constraint"]=0; cli+=["-faceAngleConstraint","false"]
    if rng.random()<0.3: opt("rel_step_frac","-relStepFrac", float(rng.uniform(0.2,1.0)), repr)
    npatch = m.n_patches
    names=[sm.lib().smmesh_patch_name(m._h,i).decode() for i in range(npatch)]
    if rng.random()<0.5:
        flags=[int(rng.random()<0.6) for _ in range(npatch)]
        if any(flags):
            okw["layer_patches"]=flags
            cli+=["-layerPatches","("+" ".join(n for n,f in zip(names,flags) if f)+")"]
            if rng.random()<0.5: opt("max_layers","-maxLayers", int(rng.integers(1,6)))
            if rng.random()<0.5: opt("layer_expansion_ratio","-layerExpansionRatio", float(rng.uniform(1.0,1.6)), repr)
            if rng.random()<0.5: opt("layer_max_blending_fraction","-layerMaxBlendingFraction", float(rng.uniform(0.1,1.0)), repr)
            if rng.random()<0.3: opt("layer_edge_length","-layerEdgeLength", float(rng.uniform(0.01,0.2)), repr)
    iters=int(rng.integers(3,12))
    par = None
    if rng.random()<0.5:
        par = (int(rng.integers(1,3)), int(rng.integers(1,3)), int(rng.integers(1,3))) if kind not in ("kelvin","tets") else (int(rng.integers(2,5)),)
        if kind not in ("kelvin","tets") and par==(1,1,1): par=(2,1,1)
    tmp=tempfile.mkdtemp(prefix="fz_")
    try:
        m.write(tmp+"/constant/polyMesh"); os.makedirs(tmp+"/system")
        open(tmp+"/system/controlDict","w").write("startFrom startTime;\nstartTime 0;\ndeltaT 1;\nwriteFormat binary;\n")
        args=[REF,"-case",tmp,"-centroidalIters",str(iters),"-smoothingPatches","()"]+cli
        if par:
            parts = m.decompose(*par) if kind not in ("kelvin","tets") else m.decompose(par[0],method="rcb")
            sm.Mesh.write_decomposed(parts,tmp,binary=True); args.insert(3,"-parallel")
        r=subprocess.run(args,capture_output=True,text=True,timeout=120)
        log=re.findall(r"Smoothing iteration=(\d+) nFrozenPoints=(\d+) residual=(\S+)", r.stdout)
        if par:
            parts=[sm.Mesh.read_processor(tmp,k) for k in range(len(parts))]
            try:
                o=Oracle([p.desc_arrays() for p in parts], libm=True, **okw); n,nf,res=o.iterate(iters); oerr=None
            except RuntimeError as e:
                oerr=str(e)
        else:
            try:
                o=Oracle(m.desc_arrays(), libm=True, **okw); n,nf,res=o.iterate(iters); oerr=None
            except RuntimeError as e:
                oerr=str(e)
        if r.returncode!=0 or oerr:
            ok = (r.returncode!=0) and bool(oerr)
            return ok, f"seed {seed} {kind} par={par} both-fail={ok} ref_rc={r.returncode} oracle_err={oerr} ref_err={r.stderr[-200:]}"
        ok = [int(b) for _,b,_ in log]==nf.tolist()
        if par:
            for k,p in enumerate(parts):
                p.read_points(f"{tmp}/processor{k}/{n}/polyMesh/points"); ok = ok and np.array_equal(p.points,o.get("points",rank=k))
        else:
            out=sm.Mesh.read(tmp+"/constant/polyMesh"); out.read_points(f"{tmp}/{n}/polyMesh/points"); ok = ok and np.array_equal(out.points,o.get("points"))
        return ok, f"seed {seed} {kind} par={par} iters={n} nf={nf[-1]} opts={okw} ok={ok}"
    finally:
        shutil.rmtree(tmp,ignore_errors=True)


@pytest.mark.parametrize("block", range(20))
def test_random_configurations_agree_with_the_reference_translation_unit(block):
    for seed in range(10 * block, 10 * block + 10):
        ok, msg = one(seed)
        assert ok, msg


@pytest.mark.parametrize("block", range(3))
def test_random_tetrahedral_configurations_agree_with_the_reference_translation_unit(block):
    for seed in range(5000 + 10 * block, 5000 + 10 * block + 10):
        ok, msg = one(seed, force_kind="tets")
        assert ok, msg
