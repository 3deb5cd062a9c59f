"""Properties at BASELINE.json's full single-GPU size (200^3 hex block, 8.1 M points), where the CPU
oracle is too slow to be the checker: size-independent invariants of the path, and the literal
evaluation path as an independent checker of the filtered production path."""
import numpy as np
import pytest

import smoothmesh_b200 as sm
from oracle import Oracle

from meshes import hex_jittered

pytestmark = pytest.mark.gpu
N = 200


@pytest.fixture(scope="module")
def mesh():
    return hex_jittered(N, N, N, 0.25, seed=12345)


def test_full_size_filtered_path_equals_literal_path(mesh, monkeypatch):
    iters = 3
    g = sm.Smoother(mesh, rel_tol=0.0)
    log = g.iterate(iters)
    pts, fz = g.points(), g.frozen()
    st = g.filter_stats()
    assert st["fused"] and st["suspect_points"] <= 0.001 * mesh.n_points, st   # the fast path is the one measured
    g.close()
    monkeypatch.setenv("SMGPU_NO_FILTERS", "1")
    monkeypatch.setenv("SMGPU_NO_TILES", "1")  # and the two-kernel geometry instead of the fused tile kernel
    lit = sm.Smoother(mesh, rel_tol=0.0)
    log2 = lit.iterate(iters)
    assert np.array_equal(log.n_frozen, log2.n_frozen) and np.array_equal(log.residual, log2.residual)
    assert np.array_equal(fz, lit.frozen())
    assert np.array_equal(pts, lit.points())


def test_full_size_invariants(mesh):
    x0 = np.array(mesh.points)
    g = sm.Smoother(mesh, rel_tol=0.0)
    p = g.params
    stats = g.mesh_stats()
    assert p.min_edge_length == 0.5 * stats["min_edge"] and p.max_step_length == 0.3 * p.min_edge_length   # K6
    log = g.iterate(6)
    pts = g.points()
    lattice = np.arange((N + 1) ** 3)
    i, j, k = lattice % (N + 1), (lattice // (N + 1)) % (N + 1), lattice // (N + 1) ** 2
    bnd = (i == 0) | (i == N) | (j == 0) | (j == N) | (k == 0) | (k == N)
    assert np.array_equal(pts[bnd], x0[bnd])                                   # K5: boundary points never move
    assert (log.n_frozen >= bnd.sum()).all() and stats["n_internal_points"] == (~bnd).sum()
    assert (log.residual <= 1.0 + 1e-12).all()                                 # K4: |step| <= maxStepLength
    # one more iteration by hand: step lengths obey the clamp, frozen points do not move
    before = g.points()
    log1 = g.iterate(1)
    after, fz = g.points(), g.frozen().astype(bool)
    step = np.linalg.norm(after - before, axis=1)
    assert step.max() <= p.max_step_length * (1 + 1e-12)
    assert np.array_equal(after[fz], before[fz])
    assert log1.n_frozen[0] == (fz | bnd).sum()                                # the count printed at :2396
    assert abs(log1.residual[0] - step.max() / p.max_step_length) <= 1e-12
    # smoothing a jittered uniform block drives it back towards the lattice
    h = 1.0 / N
    ideal = np.stack([i * h, j * h, k * h], axis=1)
    assert np.abs(after - ideal).max() < np.abs(x0 - ideal).max()


def test_full_size_uniform_block_is_a_fixed_point():
    g = sm.Smoother(sm.Mesh.hex_block(N, N, N))
    log = g.iterate(5)
    assert log.iterations == 1 and log.n_frozen[0] == (N + 1) ** 3 - (N - 1) ** 3 and log.residual[0] < 1e-9


def test_mid_size_oracle_parity():
    # the largest size the serial oracle finishes in seconds
    mesh = hex_jittered(48, 48, 48, 0.25, seed=12345)
    g, o = sm.Smoother(mesh, rel_tol=0.0), Oracle(mesh.desc_arrays(), rel_tol=0.0)
    n, nf, res = o.iterate(3)
    log = g.iterate(3)
    assert np.array_equal(log.n_frozen, nf) and np.array_equal(log.residual, res)
    assert np.array_equal(g.points(), o.get("points")) and np.array_equal(g.frozen(), o.get("frozen"))


def test_edge_cases_tiny_meshes_and_zero_iterations():
    for dims in [(1, 1, 1), (2, 1, 1), (2, 2, 2)]:
        m = sm.Mesh.hex_block(*dims)
        g, o = sm.Smoother(m), Oracle(m.desc_arrays())
        assert g.iterate(0).iterations == 0
        n, nf, res = o.iterate(4)
        log = g.iterate(4)
        assert log.iterations == n and np.array_equal(log.n_frozen, nf) and np.array_equal(log.residual, res)
        assert np.array_equal(g.points(), o.get("points"))
