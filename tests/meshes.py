"""Seeded synthetic meshes shared by the CPU and GPU tests (tests/ only)."""
import numpy as np

import smoothmesh_b200 as sm


def hex_jittered(nx, ny, nz, frac, seed=12345, hi=(1.0, 1.0, 1.0)):
    """Hex block with interior jitter U(-frac*h, frac*h), h = shortest cell side."""
    m = sm.Mesh.hex_block(nx, ny, nz, hi=hi)
    h = min(hi[0] / nx, hi[1] / ny, hi[2] / nz)
    return m.jitter(frac * h, seed)


def kelvin_jittered(n, frac, seed=7):
    m = sm.Mesh.kelvin(n, 1.0)
    # shortest Kelvin edge = sqrt(2)/4 * h
    return m.jitter(frac * (2 ** 0.5) / 4.0, seed)


def prism_layers(n=6, layers=5, thickness=0.02, frac=0.2, seed=3):
    """Flat, high-aspect-ratio hex layers: exercises the midpoint-of-two-closest-points rule."""
    m = sm.Mesh.hex_block(n, n, layers, hi=(1.0, 1.0, thickness * layers))
    return m.jitter(frac * thickness, seed)


CASES = {
    "hex6_j25": lambda: hex_jittered(6, 6, 6, 0.25),
    "hex_8x6x5_j45": lambda: hex_jittered(8, 6, 5, 0.45, seed=99),
    "kelvin3_j20": lambda: kelvin_jittered(3, 0.20),
    "layers_ar": lambda: prism_layers(),
}


def tet_block(n=3, frac=0.15, seed=5):
    """Conforming tetrahedral mesh: every cube of an n^3 lattice split into six tetrahedra around its main
    diagonal (Kuhn triangulation), interior points jittered.  Triangular faces, four-faced cells, points of
    valence up to 14: the generic (non-record) paths of the kernels and of the oracle."""
    import itertools
    idx = lambda i, j, k: i + (n + 1) * (j + (n + 1) * k)
    pts = np.array([[i / n, j / n, k / n] for k in range(n + 1) for j in range(n + 1) for i in range(n + 1)], dtype=float)
    cells = []
    for k in range(n):
        for j in range(n):
            for i in range(n):
                for perm in itertools.permutations(range(3)):
                    v = [np.array([i, j, k])]
                    for ax in perm:
                        step = np.zeros(3, dtype=int)
                        step[ax] = 1
                        v.append(v[-1] + step)
                    a, b, c, d = [idx(*map(int, q)) for q in v]
                    faces = [[a, b, c], [a, b, d], [a, c, d], [b, c, d]]
                    cc = pts[[a, b, c, d]].mean(axis=0)
                    out = []
                    for f in faces:
                        p = pts[f]
                        nrm = np.cross(p[1] - p[0], p[2] - p[0])
                        out.append(f if np.dot(nrm, p.mean(axis=0) - cc) > 0 else f[::-1])
                    cells.append(out)
    m = sm.Mesh.from_cells(pts, cells)
    return m.jitter(frac / n, seed)


CASES["tets3_j15"] = lambda: tet_block()
EXTRA_CASES = {"tets3_j15": CASES["tets3_j15"]}


def mixed_hex_prism_block(n=9, frac=0.2, seed=11):
    """Hex-dominant mesh: an n^3 block whose columns with i < n // 3 are split into two prisms each (the same
    diagonal in every layer, so the triangles of stacked prisms match), the rest stay hexahedra.  Triangular and
    quadrilateral faces, five- and six-faced cells: the per-tile fast path of the fused geometry kernel applies
    to the all-hex tiles only."""
    idx = lambda i, j, k: i + (n + 1) * (j + (n + 1) * k)
    pts = np.array([[i / n, j / n, k / n] for k in range(n + 1) for j in range(n + 1) for i in range(n + 1)], dtype=float)
    cells = []

    def oriented(faces, verts):
        cc = pts[list(verts)].mean(axis=0)
        out = []
        for f in faces:
            p = pts[f]
            nrm = np.cross(p[1] - p[0], p[2] - p[0])
            out.append(f if np.dot(nrm, p.mean(axis=0) - cc) > 0 else f[::-1])
        return out

    for k in range(n):
        for j in range(n):
            for i in range(n):
                a, b, c, d = idx(i, j, k), idx(i + 1, j, k), idx(i + 1, j + 1, k), idx(i, j + 1, k)
                e, f, g, h = idx(i, j, k + 1), idx(i + 1, j, k + 1), idx(i + 1, j + 1, k + 1), idx(i, j + 1, k + 1)
                if i < n // 3:
                    # two prisms sharing the vertical quad a-c-g-e
                    cells.append(oriented([[a, b, c], [e, f, g], [a, b, f, e], [b, c, g, f], [a, c, g, e]], (a, b, c, e, f, g)))
                    cells.append(oriented([[a, c, d], [e, g, h], [c, d, h, g], [d, a, e, h], [a, c, g, e]], (a, c, d, e, g, h)))
                else:
                    cells.append(oriented([[a, b, c, d], [e, f, g, h], [a, b, f, e], [b, c, g, f], [c, d, h, g], [d, a, e, h]],
                                          (a, b, c, d, e, f, g, h)))
    m = sm.Mesh.from_cells(pts, cells)
    return m.jitter(frac / n, seed)


CASES["mixed9_j20"] = lambda: mixed_hex_prism_block()
