import sys; sys.path.insert(0,'tests')
import numpy as np, smoothmesh_b200 as sm
from oracle import Oracle
import test_gpu_boundary as tb
for seed,dims in [(0,(2,1,1)),(1,(2,2,1)),(3,(3,1,1))]:
    mesh, geo, flags, layer, kw, frac, iters = tb.synthetic(seed)
    parts = mesh.decompose(*dims)
    members = [sm.Smoother(p, layer_patches=layer, device=0, **kw) for p in parts]
    grp = sm.Group(members)
    grp.enable_boundary_smoothing(geo, flags, frac)
    o = Oracle([p.desc_arrays() for p in parts], layer_patches=layer, smoothing_patches=flags, geometry=geo, internal_smoothing_blending_fraction=frac, **kw)
    print("case", seed, dims, "layer", layer, "frac", frac, "flags", flags)
    for it in range(iters):
        log = grp.iterate(1); n, nf, res = o.iterate(1)
        bad = False
        for r,g in enumerate(members):
            gp, op = g.points(), o.get("points", r)
            d = np.abs(gp-op).max(axis=1)
            idx = np.nonzero(d>0)[0]
            if len(idx):
                bad = True
                isint = o.get("isInternal", r)
                shared = set(g.comm_local_shared().tolist()) if False else None
                gid = np.asarray(parts[r].point_global_id)
                print(f"  it{it+1} rank{r}: {len(idx)} points differ, max {d.max():.3e}; first: ", [(int(i), int(isint[i]), float(d[i])) for i in idx[:6]])
        print(f"  it{it+1} nf gpu {log.n_frozen} orc {nf} res {log.residual} {res}")
        if bad: break
    grp.close()
