"""oracle -- TEST INFRASTRUCTURE ONLY (parity checker + CPU baseline).

ctypes binding over oracle/_build/liboracle.so, the CPU restatement of the
reference's smoothing iteration (oracle/oracle.cpp).  Parity: bit-exact against the
reference's own translation unit compiled against an OpenFOAM facade (oracle/_ref,
tests/test_reference_build.py); the OpenFOAM semantics inside that facade are recalled,
not verified -- see the header of oracle.cpp.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_libs = {}


class OParams(C.Structure):
    _fields_ = [
        ("minEdgeLength", C.c_double),
        ("maxStepLength", C.c_double),
        ("relStepFrac", C.c_double),
        ("minAngle", C.c_double),
        ("maxAngle", C.c_double),
        ("relTol", C.c_double),
        ("totalMinFreeze", C.c_int32),
        ("edgeAngleConstraint", C.c_int32),
        ("faceAngleConstraint", C.c_int32),
        ("geometryVariant", C.c_int32),
        ("layerMaxBlendingFraction", C.c_double),
        ("layerEdgeLength", C.c_double),
        ("layerExpansionRatio", C.c_double),
        ("minLayers", C.c_int32),
        ("maxLayers", C.c_int32),
    ]


class _OMesh(C.Structure):
    _fields_ = [
        ("P", C.c_int64), ("C", C.c_int64), ("F", C.c_int64), ("Fi", C.c_int64),
        ("pts", C.c_void_p), ("fOff", C.c_void_p), ("fV", C.c_void_p), ("own", C.c_void_p), ("nei", C.c_void_p),
        ("nPatches", C.c_int32),
        ("pStart", C.c_void_p), ("pSize", C.c_void_p), ("pKind", C.c_void_p),
        ("pointGlobalId", C.c_void_p),
        ("pLayer", C.c_void_p),
        ("pSmooth", C.c_void_p),
    ]


def build():
    subprocess.check_call(["make", "-s", "oracle/_build/liboracle.so", "oracle/_build/liboracle_libm.so"], cwd=_ROOT)


def _lib(libm=False):
    key = bool(libm)
    if key not in _libs:
        path = os.path.join(_HERE, "_build", "liboracle_libm.so" if libm else "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_threads.argtypes = [C.c_void_p, C.c_int]
        L.orc_last_error.restype = C.c_char_p
        L.orc_iterate.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_mesh_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_set_params.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_set_params.restype = C.c_int
        L.orc_set_geometry.argtypes = [C.c_void_p] + [C.c_int64, C.c_void_p] * 6 + [C.c_double]
        L.orc_set_label_lists.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_sizes.restype = C.c_int64
        L.orc_sizes.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_get.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p, C.c_int64]
        L.orc_get_csr.restype = C.c_int64
        L.orc_get_csr.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_acos.restype = C.c_double
        L.orc_acos.argtypes = [C.c_double]
        L.orc_edgeEdgeAngle.restype = C.c_double
        L.orc_edgeEdgeAngle.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_edge_face_angles.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _libs[key] = L
    return _libs[key]


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def acos(x, libm=False):
    return _lib(libm).orc_acos(float(x))


def edge_edge_angle(c, p1, p2, libm=False):
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (c, p1, p2)]
    return _lib(libm).orc_edgeEdgeAngle(_p(a[0]), _p(a[1]), _p(a[2]))


class Oracle:
    """CPU reference run.  `meshes` is one dict (serial) or a list of dicts (rank emulation),
    each with the keys of smoothmesh_b200.Mesh.desc_arrays().  Negative minEdgeLength /
    maxStepLength select the reference defaults (src/smoothMesh.C:1861-1865)."""

    def __init__(self, meshes, min_edge_length=-1.0, max_step_length=-1.0, rel_step_frac=0.5, min_angle_deg=35.0,
                 max_angle_deg=160.0, rel_tol=0.02, total_min_freeze=0, edge_angle_constraint=1,
                 face_angle_constraint=1, geometry_variant=0, libm=False, threads=1, layer_patches=None,
                 layer_max_blending_fraction=0.3, layer_edge_length=-1.0, layer_expansion_ratio=1.3, min_layers=1,
                 max_layers=4, smoothing_patches=None, geometry=None, internal_smoothing_blending_fraction=0.0):
        """layer_patches / smoothing_patches: 0/1 flag per physical patch (-layerPatches / -smoothingPatches).
        geometry: dict(init_edges=(points, edges), target_edges=(points, edges), surface=(points, tris)) --
        the arrays of constant/geometry/*.obj; with it and a selected smoothing patch the run does boundary
        point smoothing (oracle only, the CUDA path does not have it)."""
        self.L = _lib(libm)
        if isinstance(meshes, dict):
            meshes = [meshes]
        self.n_ranks = len(meshes)
        self._keep = []
        arr = (_OMesh * self.n_ranks)()
        for r, m in enumerate(meshes):
            a = dict(
                pts=np.ascontiguousarray(m["points"], dtype=np.float64),
                fOff=np.ascontiguousarray(m["face_offsets"], dtype=np.int32),
                fV=np.ascontiguousarray(m["face_verts"], dtype=np.int32),
                own=np.ascontiguousarray(m["owner"], dtype=np.int32),
                nei=np.ascontiguousarray(m["neighbour"], dtype=np.int32),
                pStart=np.ascontiguousarray(m["patch_start"], dtype=np.int32),
                pSize=np.ascontiguousarray(m["patch_size"], dtype=np.int32),
                pKind=np.ascontiguousarray(m["patch_kind"], dtype=np.int32),
                gid=None if m.get("point_global_id") is None else np.ascontiguousarray(m["point_global_id"], dtype=np.int64),
                lay=None,
            )
            if layer_patches is not None:
                # flags of the physical patches; a processor mesh appends its processor patches (never selected)
                lay = np.zeros(a["pStart"].size, dtype=np.int32)
                k = min(lay.size, len(layer_patches))
                lay[:k] = np.asarray(layer_patches, dtype=np.int32)[:k]
                a["lay"] = lay
            self._keep.append(a)
            o = arr[r]
            o.P, o.C, o.F, o.Fi = a["pts"].size // 3, int(m["n_cells"]), a["own"].size, a["nei"].size
            o.pts, o.fOff, o.fV, o.own, o.nei = _p(a["pts"]), _p(a["fOff"]), _p(a["fV"]), _p(a["own"]), _p(a["nei"])
            o.nPatches = a["pStart"].size
            o.pStart, o.pSize, o.pKind = _p(a["pStart"]), _p(a["pSize"]), _p(a["pKind"])
            o.pointGlobalId = _p(a["gid"])
            o.pLayer = _p(a["lay"])
            a["smo"] = None
            if smoothing_patches is not None:
                smo = np.zeros(a["pStart"].size, dtype=np.int32)
                k = min(smo.size, len(smoothing_patches))
                smo[:k] = np.asarray(smoothing_patches, dtype=np.int32)[:k]
                a["smo"] = smo
            o.pSmooth = _p(a["smo"])
        self.prm = OParams(min_edge_length, max_step_length, rel_step_frac, min_angle_deg, max_angle_deg, rel_tol,
                           int(total_min_freeze), int(edge_angle_constraint), int(face_angle_constraint),
                           int(geometry_variant), layer_max_blending_fraction, layer_edge_length,
                           layer_expansion_ratio, int(min_layers), int(max_layers))
        self.h = self.L.orc_create(self.n_ranks, C.byref(arr), C.byref(self.prm))
        if not self.h:
            raise RuntimeError("oracle: " + self.L.orc_last_error().decode())
        self._keep = None
        if geometry is not None:
            g = []
            for key, width in (("init_edges", 2), ("target_edges", 2), ("surface", 3)):
                pts, idx = geometry[key]
                pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
                idx = np.ascontiguousarray(idx, dtype=np.int32).reshape(-1, width)
                g += [pts.shape[0], _p(pts), idx.shape[0], _p(idx)]
                self.__dict__.setdefault("_geo_keep", []).extend([pts, idx])
            self.L.orc_set_geometry(self.h, *g, float(internal_smoothing_blending_fraction))
            if geometry.get("is_corner_point") is not None:   # label lists of an earlier run (serial restart)
                a = np.ascontiguousarray(geometry["is_corner_point"], dtype=np.int32)
                b = np.ascontiguousarray(geometry["is_feature_edge_point"], dtype=np.int32)
                self.L.orc_set_label_lists(self.h, 0, a.size, _p(a), _p(b))
        mn, mx = C.c_double(), C.c_double()
        self.L.orc_mesh_stats(self.h, C.byref(mn), C.byref(mx))
        self.min_edge, self.max_edge = mn.value, mx.value
        if self.prm.minEdgeLength < 0:
            self.prm.minEdgeLength = 0.5 * self.min_edge
        if self.prm.maxStepLength < 0:
            self.prm.maxStepLength = 0.3 * self.prm.minEdgeLength
        if self.L.orc_set_params(self.h, C.byref(self.prm)) != 0:
            raise RuntimeError("oracle: " + self.L.orc_last_error().decode())
        self.L.orc_set_threads(self.h, int(threads))

    def __del__(self):
        try:
            if self.h:
                self.L.orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def size(self, what, rank=0):
        return int(self.L.orc_sizes(self.h, rank, {"points": 0, "cells": 1, "faces": 2, "edges": 3}[what]))

    def iterate(self, max_iters):
        nf = np.zeros(max(max_iters, 1), dtype=np.int64)
        res = np.zeros(max(max_iters, 1), dtype=np.float64)
        n = self.L.orc_iterate(self.h, int(max_iters), _p(nf), _p(res))
        if n < 0:
            raise RuntimeError("oracle: " + self.L.orc_last_error().decode())
        return n, nf[:n].copy(), res[:n].copy()

    _SHAPES = {
        "points": ("points", np.float64, 3), "cellCentres": ("cells", np.float64, 3),
        "faceCentres": ("faces", np.float64, 3), "faceAreas": ("faces", np.float64, 3),
        "snapCellCtr": ("cells", np.float64, 3), "snapCentroidal": ("points", np.float64, 3),
        "snapBlend": ("points", np.float64, 3), "snapClamped": ("points", np.float64, 3),
        "frozen": ("points", np.uint8, 1), "isInternal": ("points", np.uint8, 1),
        "snapFrozenEdgeLen": ("points", np.uint8, 1), "snapFrozenEdgeAngle": ("points", np.uint8, 1),
        "snapFrozenFaceAngle": ("points", np.uint8, 1), "snapCurMin": ("points", np.float64, 1),
        "snapCurMax": ("points", np.float64, 1), "edges": ("edges", np.int32, 2),
        "snapNormals": ("points", np.float64, 3), "snapLayerBlend": ("points", np.float64, 3),
        "hopsToLayer": ("points", np.int32, 1), "pointToOuter": ("points", np.int32, 1),
        "isCorner": ("points", np.uint8, 1), "isFeatureEdge": ("points", np.uint8, 1),
        "isSmoothingSurface": ("points", np.uint8, 1), "cornerPoints": ("points", np.float64, 3),
        "pointStrings": ("points", np.int32, 1), "hopsToSmoothing": ("points", np.int32, 1),
        "pointToInner": ("points", np.int32, 1),
    }

    def get(self, name, rank=0):
        what, dt, w = self._SHAPES[name]
        n = self.size(what, rank)
        out = np.zeros((n, w) if w > 1 else n, dtype=dt)
        got = self.L.orc_get(self.h, rank, name.encode(), _p(out), out.nbytes)
        if got < 0:
            raise RuntimeError(f"oracle: no array {name} (rc={got})")
        return out

    def csr(self, name, rank=0):
        rows = {"pointFaces": "points", "pointCells": "points", "pointPoints": "points", "pointEdges": "points",
                "edgeFaces": "edges", "edgeCells": "edges", "cellFaces": "cells", "cellPoints": "cells"}[name]
        n = self.size(rows, rank)
        tot = self.L.orc_get_csr(self.h, rank, name.encode(), None, None, 0)
        off = np.zeros(n + 1, dtype=np.int32)
        val = np.zeros(max(tot, 1), dtype=np.int32)
        self.L.orc_get_csr(self.h, rank, name.encode(), _p(off), _p(val), tot)
        return off, val[:tot]

    def edge_face_angles(self, edge, rank=0):
        mn, mx = C.c_double(), C.c_double()
        if self.L.orc_edge_face_angles(self.h, rank, int(edge), C.byref(mn), C.byref(mx)) != 0:
            raise RuntimeError("oracle: " + self.L.orc_last_error().decode())
        return mn.value, mx.value
