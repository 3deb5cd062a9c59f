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
