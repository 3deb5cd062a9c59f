import time, sys
sys.path.insert(0, "tests")
import smoothmesh_b200 as sm
from meshes import hex_jittered
small = sm.Smoother(hex_jittered(8, 8, 8, 0.2)); small.iterate(1)
m = hex_jittered(200, 200, 200, 0.25)
for rep in range(2):
    t = time.perf_counter(); g = sm.Smoother(m, rel_tol=0.0); print("create", time.perf_counter() - t, flush=True)
    log = g.iterate(3); print(log.n_frozen, log.ms / 3)
    g.close()
