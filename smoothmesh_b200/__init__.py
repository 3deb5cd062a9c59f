"""smoothmesh_b200 -- thin ctypes binding over libsmgpu.so (include/smgpu.h, include/smmesh.h).

The product is the CUDA library and the C++ `smoothMesh` CLI; this module only
exists so that tests and bench.py can drive the C ABI from Python.  Names mirror
the reference's vocabulary (src/smoothMesh.C): `Mesh` is the polyMesh, `Smoother`
owns the device-resident state and runs the iteration loop of :2257-2437.

There is no CPU fallback: `Smoother` raises if the library or a CUDA device is
missing.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsmgpu.so")
CLI_PATH = os.path.join(_HERE, "bin", "smoothMesh")

_lib = None


class SmoothMeshError(RuntimeError):
    pass


class Params(C.Structure):
    """smgpu_params (include/smgpu.h); defaults = src/smoothMesh.C:1861-1914."""

    _fields_ = [
        ("min_edge_length", C.c_double),
        ("max_step_length", C.c_double),
        ("rel_step_frac", C.c_double),
        ("min_angle_deg", C.c_double),
        ("max_angle_deg", C.c_double),
        ("rel_tol", C.c_double),
        ("total_min_freeze", C.c_int32),
        ("edge_angle_constraint", C.c_int32),
        ("face_angle_constraint", C.c_int32),
        ("geometry_variant", C.c_int32),
        ("device", C.c_int32),
        ("renumber", C.c_int32),
        ("layer_max_blending_fraction", C.c_double),
        ("layer_edge_length", C.c_double),
        ("layer_expansion_ratio", C.c_double),
        ("min_layers", C.c_int32),
        ("max_layers", C.c_int32),
    ]


class _MeshDesc(C.Structure):
    _fields_ = [
        ("n_points", C.c_int64),
        ("n_cells", C.c_int64),
        ("n_faces", C.c_int64),
        ("n_internal_faces", C.c_int64),
        ("points", C.c_void_p),
        ("face_offsets", C.c_void_p),
        ("face_verts", C.c_void_p),
        ("owner", C.c_void_p),
        ("neighbour", C.c_void_p),
        ("n_patches", C.c_int32),
        ("patch_start", C.c_void_p),
        ("patch_size", C.c_void_p),
        ("patch_kind", C.c_void_p),
        ("point_global_id", C.c_void_p),
        ("patch_layer", C.c_void_p),
    ]


def lib():
    """Load libsmgpu.so (fails loudly if it has not been built)."""
    global _lib
    if _lib is None:
        path = os.environ.get("SMGPU_LIB", LIB_PATH)  # override: kernel-variant experiments only
        if not os.path.exists(path):
            raise SmoothMeshError(f"{path} not found: run `make` (or __graft_entry__.build()) first")
        L = C.CDLL(path, mode=C.RTLD_GLOBAL)
        L.smgpu_last_error.restype = C.c_char_p
        L.smgpu_version.restype = C.c_char_p
        L.smmesh_last_error.restype = C.c_char_p
        L.smmesh_patch_name.restype = C.c_char_p
        for f in ("smmesh_gen_hex_block", "smmesh_gen_kelvin", "smmesh_gen_kelvin_part", "smmesh_from_cells",
                  "smmesh_from_arrays", "smmesh_read"):
            getattr(L, f).restype = C.c_void_p
        for f in ("smmesh_points", "smmesh_points_mut", "smmesh_face_offsets", "smmesh_face_verts", "smmesh_owner",
                  "smmesh_neighbour", "smmesh_point_global_id", "smmesh_cell_global_id"):
            getattr(L, f).restype = C.c_void_p
            getattr(L, f).argtypes = [C.c_void_p]
        L.smmesh_size.restype = C.c_int64
        L.smmesh_size.argtypes = [C.c_void_p, C.c_int32]
        L.smmesh_free.argtypes = [C.c_void_p]
        L.smmesh_jitter.argtypes = [C.c_void_p, C.c_double, C.c_uint64]
        L.smmesh_patches.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.smmesh_patch_name.argtypes = [C.c_void_p, C.c_int32]
        L.smmesh_write.argtypes = [C.c_void_p, C.c_char_p, C.c_int32, C.c_int32]
        L.smmesh_read.argtypes = [C.c_char_p]
        L.smmesh_read_points.argtypes = [C.c_void_p, C.c_char_p]
        L.smmesh_write_points.argtypes = [C.c_void_p, C.c_int64, C.c_char_p, C.c_int32, C.c_int32, C.c_char_p]
        L.smmesh_gen_hex_block.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        L.smmesh_gen_kelvin.argtypes = [C.c_int32, C.c_double]
        L.smmesh_gen_kelvin_part.argtypes = [C.c_int32, C.c_double] + [C.c_int32] * 4
        L.smmesh_decompose.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
        L.smmesh_from_cells.argtypes = [C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int32, C.c_void_p, C.c_void_p]
        L.smmesh_from_arrays.argtypes = [C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                         C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.smgpu_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        for f in ("smgpu_destroy", "smgpu_get_params", "smgpu_set_params", "smgpu_get_points", "smgpu_set_points",
                  "smgpu_get_frozen", "smgpu_op_cell_centres", "smgpu_op_predict", "smgpu_op_edge_constraints",
                  "smgpu_op_face_angle_constraint", "smgpu_get_edges", "smgpu_op_layer_normals", "smgpu_op_layer_blend"):
            getattr(L, f).argtypes = [C.c_void_p] + ([C.c_void_p] if f != "smgpu_destroy" else [])
        L.smgpu_mesh_stats.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.smgpu_iterate.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.smgpu_last_timing.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.smgpu_op_commit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.smgpu_op_edge_face_angles.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.smgpu_get_csr.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.smgpu_profile.argtypes = [C.c_void_p, C.c_int32]
        L.smgpu_profile_get.argtypes = [C.c_void_p] * 5
        L.smgpu_comm_unique_id.argtypes = [C.c_void_p]
        L.smgpu_comm_init.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.smgpu_comm_local_shared.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.smgpu_comm_prepare.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        L.smgpu_comm_abort.argtypes = [C.c_void_p]
        L.smgpu_comm_p2p_export.argtypes = [C.c_void_p, C.c_void_p]
        L.smgpu_comm_p2p_connect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.smgpu_comm_p2p_disable.argtypes = [C.c_void_p]
        L.smgpu_group_create.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.smgpu_group_iterate.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.smgpu_group_destroy.argtypes = [C.c_void_p]
        L.smgpu_exchange_plan.restype = C.c_int64
        L.smgpu_exchange_plan.argtypes = [C.c_int32, C.c_int32, C.c_int64] + [C.c_void_p] * 6
        L.smmesh_quality.argtypes = [C.c_void_p, C.c_void_p]
        L.smmesh_geom_tiles.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
        L.smmesh_boundary_setup.argtypes = ([C.c_void_p] + [C.c_int64, C.c_void_p] * 4 + [C.c_void_p, C.c_double]
                                            + [C.c_void_p] * 8)
        L.smmesh_read_obj.argtypes = [C.c_char_p] + [C.c_void_p] * 6
        L.smmesh_write_decomposed.argtypes = [C.c_void_p, C.c_int32, C.c_char_p, C.c_int32]
        L.smmesh_read_processor.restype = C.c_void_p
        L.smmesh_read_processor.argtypes = [C.c_char_p, C.c_int32]
        L.smmesh_renumber.restype = C.c_void_p
        L.smmesh_renumber.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.smmesh_gen_hex_block_part.restype = C.c_void_p
        L.smmesh_gen_hex_block_part.argtypes = [C.c_int32] * 7 + [C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


class _BoundaryGeometry(C.Structure):
    _fields_ = [(n, t) for k in ("init", "target") for n, t in
                ((f"n_{k}_points", C.c_int64), (f"{k}_points", C.c_void_p), (f"n_{k}_edges", C.c_int64),
                 (f"{k}_edges", C.c_void_p))] + [("n_surface_points", C.c_int64), ("surface_points", C.c_void_p),
                                                 ("n_surface_tris", C.c_int64), ("surface_tris", C.c_void_p),
                                                 ("is_corner_point", C.c_void_p), ("is_feature_edge_point", C.c_void_p)]


def _boundary_geometry(geometry):
    """smgpu_boundary_geometry from dict(init_edges=(points, pairs), target_edges=..., surface=(points, triangles)[,
    is_corner_point=..., is_feature_edge_point=...]); returns (struct, arrays to keep alive)."""
    keep = []

    def arr(a, dt, w):
        a = np.ascontiguousarray(a, dtype=dt).reshape(-1, w)
        keep.append(a)
        return a
    g = _BoundaryGeometry()
    ip, ie = arr(geometry["init_edges"][0], np.float64, 3), arr(geometry["init_edges"][1], np.int32, 2)
    tp, te = arr(geometry["target_edges"][0], np.float64, 3), arr(geometry["target_edges"][1], np.int32, 2)
    sp, st = arr(geometry["surface"][0], np.float64, 3), arr(geometry["surface"][1], np.int32, 3)
    g.n_init_points, g.init_points, g.n_init_edges, g.init_edges = len(ip), _ptr(ip), len(ie), _ptr(ie)
    g.n_target_points, g.target_points, g.n_target_edges, g.target_edges = len(tp), _ptr(tp), len(te), _ptr(te)
    g.n_surface_points, g.surface_points, g.n_surface_tris, g.surface_tris = len(sp), _ptr(sp), len(st), _ptr(st)
    for key in ("is_corner_point", "is_feature_edge_point"):   # label lists of an earlier run (restart)
        if geometry.get(key) is not None:
            a = np.ascontiguousarray(geometry[key], dtype=np.int32)
            keep.append(a)
            setattr(g, key, _ptr(a))
    return g, keep


def default_params(**kw) -> Params:
    p = Params()
    lib().smgpu_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _view(ptr, n, dtype):
    if not ptr or n == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n)


PATCH_BOUNDARY, PATCH_PROCESSOR, PATCH_EMPTY = 0, 1, 2


class Mesh:
    """A polyMesh held by the host library (points, faces, owner, neighbour, boundary)."""

    def __init__(self, handle):
        if not handle:
            raise SmoothMeshError(lib().smmesh_last_error().decode())
        self._h = C.c_void_p(handle)

    def __del__(self):
        try:
            if self._h:
                lib().smmesh_free(self._h)
                self._h = None
        except Exception:
            pass

    # ---- constructors ----
    @staticmethod
    def hex_block(nx, ny, nz, lo=(0.0, 0.0, 0.0), hi=(1.0, 1.0, 1.0)) -> "Mesh":
        lo = np.asarray(lo, dtype=np.float64)
        hi = np.asarray(hi, dtype=np.float64)
        return Mesh(lib().smmesh_gen_hex_block(nx, ny, nz, _ptr(lo), _ptr(hi)))

    @staticmethod
    def hex_block_part(nx, ny, nz, px, py, pz, rank, lo=(0.0, 0.0, 0.0), hi=(1.0, 1.0, 1.0)) -> "Mesh":
        """Brick `rank` of a (nx*px, ny*py, nz*pz) block as a processor mesh (weak-scaling runs)."""
        lo = np.asarray(lo, dtype=np.float64)
        hi = np.asarray(hi, dtype=np.float64)
        return Mesh(lib().smmesh_gen_hex_block_part(nx, ny, nz, px, py, pz, rank, _ptr(lo), _ptr(hi)))

    @staticmethod
    def kelvin(n, h=1.0) -> "Mesh":
        return Mesh(lib().smmesh_gen_kelvin(n, h))

    @staticmethod
    def kelvin_part(n, h, px, py, pz, rank) -> "Mesh":
        """Brick `rank` of the Kelvin mesh as a processor mesh, generated locally (point_global_id = lattice slot)."""
        return Mesh(lib().smmesh_gen_kelvin_part(n, float(h), px, py, pz, rank))

    @staticmethod
    def read_processor(case_dir, k) -> "Mesh":
        """processor<k> mesh of a decomposed case, with its pointProcAddressing as point_global_id."""
        return Mesh(lib().smmesh_read_processor(str(case_dir).encode(), int(k)))

    @staticmethod
    def write_decomposed(parts, case_dir, binary=False):
        arr = (C.c_void_p * len(parts))(*[p._h for p in parts])
        if lib().smmesh_write_decomposed(arr, len(parts), str(case_dir).encode(), int(binary)) != 0:
            raise SmoothMeshError(lib().smmesh_last_error().decode())

    @staticmethod
    def read(polymesh_dir) -> "Mesh":
        return Mesh(lib().smmesh_read(str(polymesh_dir).encode()))

    @staticmethod
    def from_cells(points, cells, patch_of_face=None, patch_names=("walls",), patch_types=("wall",)) -> "Mesh":
        """cells: list of cells, each a list of outward-oriented faces (vertex-label lists)."""
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        cfo = [0]
        cvo = [0]
        cv = []
        cp = []
        for ci, cell in enumerate(cells):
            for fi, f in enumerate(cell):
                cv.extend(int(v) for v in f)
                cvo.append(len(cv))
                cp.append(patch_of_face(ci, fi, f) if patch_of_face else 0)
            cfo.append(len(cvo) - 1)
        cfo = np.asarray(cfo, dtype=np.int32)
        cvo = np.asarray(cvo, dtype=np.int32)
        cv = np.asarray(cv, dtype=np.int32)
        cp = np.asarray(cp, dtype=np.int32)
        names = (C.c_char_p * len(patch_names))(*[s.encode() for s in patch_names])
        types = (C.c_char_p * len(patch_types))(*[s.encode() for s in patch_types])
        return Mesh(lib().smmesh_from_cells(len(pts), _ptr(pts), len(cells), _ptr(cfo), _ptr(cvo), _ptr(cv), _ptr(cp),
                                            len(patch_names), names, types))

    @staticmethod
    def from_arrays(points, face_offsets, face_verts, owner, neighbour, n_cells, patch_start, patch_size,
                    patch_kind) -> "Mesh":
        a = [np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)]
        a += [np.ascontiguousarray(x, dtype=np.int32) for x in
              (face_offsets, face_verts, owner, neighbour, patch_start, patch_size, patch_kind)]
        return Mesh(lib().smmesh_from_arrays(len(a[0]), _ptr(a[0]), len(a[3]), _ptr(a[1]), _ptr(a[2]), _ptr(a[3]),
                                             len(a[4]), _ptr(a[4]), int(n_cells), len(a[5]), _ptr(a[5]), _ptr(a[6]),
                                             _ptr(a[7])))

    # ---- accessors (views into library memory; copy if you keep them) ----
    def _size(self, what):
        return int(lib().smmesh_size(self._h, what))

    n_points = property(lambda s: s._size(0))
    n_cells = property(lambda s: s._size(1))
    n_faces = property(lambda s: s._size(2))
    n_internal_faces = property(lambda s: s._size(3))
    n_patches = property(lambda s: s._size(5))

    @property
    def points(self):
        return _view(lib().smmesh_points_mut(self._h), 3 * self.n_points, np.float64).reshape(-1, 3)

    @property
    def face_offsets(self):
        return _view(lib().smmesh_face_offsets(self._h), self.n_faces + 1, np.int32)

    @property
    def face_verts(self):
        return _view(lib().smmesh_face_verts(self._h), self._size(4), np.int32)

    @property
    def owner(self):
        return _view(lib().smmesh_owner(self._h), self.n_faces, np.int32)

    @property
    def neighbour(self):
        return _view(lib().smmesh_neighbour(self._h), self.n_internal_faces, np.int32)

    @property
    def patches(self):
        n = self.n_patches
        s, z, k = (np.zeros(n, dtype=np.int32) for _ in range(3))
        lib().smmesh_patches(self._h, _ptr(s), _ptr(z), _ptr(k))
        return s, z, k

    @property
    def patch_names(self):
        return [lib().smmesh_patch_name(self._h, i).decode() for i in range(self.n_patches)]

    @property
    def point_global_id(self):
        p = lib().smmesh_point_global_id(self._h)
        return _view(p, self.n_points, np.int64) if p else None

    @property
    def cell_global_id(self):
        p = lib().smmesh_cell_global_id(self._h)
        return _view(p, self.n_cells, np.int64) if p else None

    def faces(self):
        off, v = self.face_offsets, self.face_verts
        return [v[off[i]:off[i + 1]].tolist() for i in range(self.n_faces)]

    # ---- operations ----
    def jitter(self, amp, seed=12345):
        lib().smmesh_jitter(self._h, float(amp), int(seed))
        return self

    def write(self, polymesh_dir, binary=False, precision=17):  # 17 significant digits round-trip a double
        if lib().smmesh_write(self._h, str(polymesh_dir).encode(), int(binary), int(precision)) != 0:
            raise SmoothMeshError(lib().smmesh_last_error().decode())

    def read_points(self, points_file):
        if lib().smmesh_read_points(self._h, str(points_file).encode()) != 0:
            raise SmoothMeshError(lib().smmesh_last_error().decode())

    def quality(self):
        """checkMesh-style figures: max/avg non-orthogonality [deg], max skewness, min edge-edge angle [deg], ..."""
        out = np.zeros(7)
        if lib().smmesh_quality(self._h, _ptr(out)) != 0:
            raise SmoothMeshError(lib().smmesh_last_error().decode())
        keys = ("max_non_ortho", "avg_non_ortho", "max_skewness", "min_edge_angle", "min_edge_length",
                "max_edge_length", "min_volume")
        return dict(zip(keys, out.tolist()))

    def renumber(self):
        """Morton renumbering (renumberMesh stand-in) -> (new Mesh, point_old_of_new, cell_old_of_new)."""
        pm = np.zeros(self.n_points, dtype=np.int32)
        cm = np.zeros(self.n_cells, dtype=np.int32)
        return Mesh(lib().smmesh_renumber(self._h, _ptr(pm), _ptr(cm))), pm, cm

    def layer_setup(self, patch_layer, max_layers=4):
        """Host set-up of the boundary layer treatment (include/smmesh.h: smmesh_layer_setup)."""
        flags = np.zeros(self.n_patches, dtype=np.int32)
        k = min(len(patch_layer), flags.size)
        flags[:k] = np.asarray(patch_layer, dtype=np.int32)[:k]
        out = [np.zeros(self.n_points, dtype=np.int32) for _ in range(3)]
        lib().smmesh_layer_setup.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        if lib().smmesh_layer_setup(self._h, _ptr(flags), int(max_layers), *[_ptr(a) for a in out]) != 0:
            raise SmoothMeshError(lib().smmesh_last_error().decode())
        return dict(hops=out[0], point_to_outer=out[1], normal_source=out[2])

    def boundary_setup(self, init_edges, target_edges, patch_smoothing, layer_edge_length=-1.0):
        """Host set-up of boundary point smoothing (include/smmesh.h: smmesh_boundary_setup); edges = (points, pairs)."""
        P = self.n_points
        ip, ie = (np.ascontiguousarray(init_edges[0], dtype=np.float64).reshape(-1, 3),
                  np.ascontiguousarray(init_edges[1], dtype=np.int32).reshape(-1, 2))
        tp, te = (np.ascontiguousarray(target_edges[0], dtype=np.float64).reshape(-1, 3),
                  np.ascontiguousarray(target_edges[1], dtype=np.int32).reshape(-1, 2))
        ps = np.zeros(self.n_patches, dtype=np.int32)
        k = min(len(patch_smoothing), ps.size)
        ps[:k] = np.asarray(patch_smoothing, dtype=np.int32)[:k]
        out = dict(is_corner=np.zeros(P, np.uint8), is_feature_edge=np.zeros(P, np.uint8),
                   is_smoothing_surface=np.zeros(P, np.uint8), corner_points=np.zeros((P, 3)),
                   point_strings=np.zeros(P, np.int32), hops_to_smoothing=np.zeros(P, np.int32),
                   point_to_inner=np.zeros(P, np.int32), target_edge_strings=np.zeros(len(te), np.int32))
        rc = lib().smmesh_boundary_setup(self._h, len(ip), _ptr(ip), len(ie), _ptr(ie), len(tp), _ptr(tp), len(te), _ptr(te),
                                         _ptr(ps), C.c_double(layer_edge_length), *[_ptr(out[k]) for k in
                                         ("is_corner", "is_feature_edge", "is_smoothing_surface", "corner_points",
                                          "point_strings", "hops_to_smoothing", "point_to_inner", "target_edge_strings")])
        if rc != 0:
            raise SmoothMeshError(lib().smmesh_last_error().decode())
        return out

    def geom_tiles(self, max_cells=256, max_faces=1024, max_points=1024):
        """Host-side tiling of the fused geometry kernel, checked:
        dict(tiles, listed_faces, max_faces, faces, max_points)."""
        out = (C.c_int64 * 8)()
        if lib().smmesh_geom_tiles(self._h, int(max_cells), int(max_faces), int(max_points), out) != 0:
            raise SmoothMeshError(lib().smmesh_last_error().decode())
        return dict(tiles=out[0], listed_faces=out[1], max_faces=out[2], faces=out[3], max_points=out[4],
                    edge_cell_pairs=out[5], uniform_cell_edges=out[6], uniform_tile_cells=out[7])

    def decompose(self, px, py=1, pz=1, method="bricks"):
        n = px * py * pz if method == "bricks" else px
        out = (C.c_void_p * n)()
        rc = lib().smmesh_decompose(self._h, 0 if method == "bricks" else 1, px, py, pz, out)
        if rc != 0:
            raise SmoothMeshError(lib().smmesh_last_error().decode())
        return [Mesh(out[i]) for i in range(n)]

    def desc_arrays(self, copy=True):
        """The arrays of smgpu_mesh_desc as contiguous numpy arrays (also what the oracle takes): copies, or with
        copy=False views of this mesh's own storage (valid while the mesh lives; smgpu_create copies what it keeps)."""
        s, z, k = self.patches
        gid = self.point_global_id
        arr = (lambda a: np.array(a)) if copy else (lambda a: a)
        return dict(points=arr(self.points), face_offsets=arr(self.face_offsets),
                    face_verts=arr(self.face_verts), owner=arr(self.owner), neighbour=arr(self.neighbour),
                    n_cells=self.n_cells, patch_start=s, patch_size=z, patch_kind=k,
                    point_global_id=None if gid is None else arr(gid))


@dataclass
class IterationLog:
    iterations: int
    n_frozen: np.ndarray  # int64 per iteration (what the reference prints at src/smoothMesh.C:2396)
    residual: np.ndarray  # float64 per iteration
    ms: float             # device time of the loop (CUDA events)
    launches: int


class Smoother:
    """Device-resident smoothing state; the C-ABI counterpart of the reference's main() loop."""

    def __init__(self, mesh: Mesh, params: Params | None = None, layer_patches=None, **kw):
        """layer_patches: 0/1 flag per patch (the patches -layerPatches selects); shorter lists are padded
        with 0, so the flags of the physical patches also fit a processor mesh (its processor patches
        come last).  On a processor mesh the layer set-up completes in comm_init (it is collective)."""
        L = lib()
        self.params_in = params if params is not None else default_params(**kw)
        self._arrays = mesh.desc_arrays(copy=False)  # views of the mesh's storage: alive during create
        a = self._arrays
        d = _MeshDesc()
        d.n_points, d.n_cells = len(a["points"]), a["n_cells"]
        d.n_faces, d.n_internal_faces = len(a["owner"]), len(a["neighbour"])
        d.points, d.face_offsets, d.face_verts = _ptr(a["points"]), _ptr(a["face_offsets"]), _ptr(a["face_verts"])
        d.owner, d.neighbour = _ptr(a["owner"]), _ptr(a["neighbour"])
        d.n_patches = len(a["patch_start"])
        d.patch_start, d.patch_size, d.patch_kind = _ptr(a["patch_start"]), _ptr(a["patch_size"]), _ptr(a["patch_kind"])
        d.point_global_id = _ptr(a["point_global_id"])
        lay = None
        if layer_patches is not None:
            lay = np.zeros(d.n_patches, dtype=np.int32)
            k = min(d.n_patches, len(layer_patches))
            lay[:k] = np.asarray(layer_patches, dtype=np.int32)[:k]
        d.patch_layer = _ptr(lay)
        h = C.c_void_p()
        rc = L.smgpu_create(C.byref(d), C.byref(self.params_in), C.byref(h))
        if rc != 0:
            raise SmoothMeshError(f"smgpu_create failed ({rc}): {L.smgpu_last_error().decode()}")
        self._h = h
        self.n_points, self.n_cells, self.n_patches = int(d.n_points), int(d.n_cells), int(d.n_patches)
        self._arrays = None

    def _ck(self, rc):
        if rc != 0:
            raise SmoothMeshError(f"libsmgpu error {rc}: {lib().smgpu_last_error().decode()}")

    def close(self):
        if getattr(self, "_h", None):
            lib().smgpu_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def params(self) -> Params:
        p = Params()
        self._ck(lib().smgpu_get_params(self._h, C.byref(p)))
        return p

    def set_params(self, p: Params):
        self._ck(lib().smgpu_set_params(self._h, C.byref(p)))

    def mesh_stats(self):
        mn, mx = C.c_double(), C.c_double()
        ni, ne = C.c_int64(), C.c_int64()
        self._ck(lib().smgpu_mesh_stats(self._h, C.byref(mn), C.byref(mx), C.byref(ni), C.byref(ne)))
        return dict(min_edge=mn.value, max_edge=mx.value, n_internal_points=ni.value, n_edges=ne.value)

    def iterate(self, max_iters) -> IterationLog:
        nf = np.zeros(max(max_iters, 1), dtype=np.int64)
        res = np.zeros(max(max_iters, 1), dtype=np.float64)
        done = C.c_int32()
        self._ck(lib().smgpu_iterate(self._h, int(max_iters), _ptr(nf), _ptr(res), C.byref(done)))
        ms, ln = C.c_double(), C.c_int64()
        self._ck(lib().smgpu_last_timing(self._h, C.byref(ms), C.byref(ln)))
        n = done.value
        return IterationLog(n, nf[:n].copy(), res[:n].copy(), ms.value, ln.value)

    def points(self, out=None):
        """mesh.points() of the last executed iteration; `out`: a (n_points, 3) float64 array to fill (reused buffers
        avoid the page faults of a fresh allocation)."""
        if out is None:
            out = np.empty((self.n_points, 3), dtype=np.float64)
        assert out.dtype == np.float64 and out.size == 3 * self.n_points and out.flags.c_contiguous
        self._ck(lib().smgpu_get_points(self._h, _ptr(out)))
        return out

    def set_points(self, pts):
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        assert pts.size == 3 * self.n_points
        self._ck(lib().smgpu_set_points(self._h, _ptr(pts)))

    def frozen(self):
        out = np.zeros(self.n_points, dtype=np.uint8)
        self._ck(lib().smgpu_get_frozen(self._h, _ptr(out)))
        return out

    # ---- operator-level entry points (one per reference L3 function) ----
    def op_cell_centres(self):
        out = np.zeros((self.n_cells, 3), dtype=np.float64)
        self._ck(lib().smgpu_op_cell_centres(self._h, _ptr(out)))
        return out

    def op_predict(self):
        out = np.zeros((self.n_points, 3), dtype=np.float64)
        self._ck(lib().smgpu_op_predict(self._h, _ptr(out)))
        return out

    def op_layer_normals(self):
        out = np.zeros((self.n_points, 3), dtype=np.float64)
        self._ck(lib().smgpu_op_layer_normals(self._h, _ptr(out)))
        return out

    def op_layer_blend(self):
        out = np.zeros((self.n_points, 3), dtype=np.float64)
        self._ck(lib().smgpu_op_layer_blend(self._h, _ptr(out)))
        return out

    def op_edge_constraints(self):
        out = np.zeros(self.n_points, dtype=np.uint8)
        self._ck(lib().smgpu_op_edge_constraints(self._h, _ptr(out)))
        return out

    def op_face_angle_constraint(self):
        out = np.zeros(self.n_points, dtype=np.uint8)
        self._ck(lib().smgpu_op_face_angle_constraint(self._h, _ptr(out)))
        return out

    def op_commit(self):
        nf, r = C.c_int64(), C.c_double()
        self._ck(lib().smgpu_op_commit(self._h, C.byref(nf), C.byref(r)))
        return nf.value, r.value

    def op_edge_face_angles(self):
        ne = self.mesh_stats()["n_edges"]
        mn, mx = np.zeros(ne), np.zeros(ne)
        self._ck(lib().smgpu_op_edge_face_angles(self._h, _ptr(mn), _ptr(mx)))
        return mn, mx

    def edges(self):
        ne = self.mesh_stats()["n_edges"]
        out = np.zeros((ne, 2), dtype=np.int32)
        self._ck(lib().smgpu_get_edges(self._h, _ptr(out)))
        return out

    def csr(self, name):
        n = C.c_int64()
        self._ck(lib().smgpu_get_csr(self._h, name.encode(), None, None, C.byref(n)))
        rows = self.mesh_stats()["n_edges"] if name.startswith("edge") else self.n_points
        off = np.zeros(rows + 1, dtype=np.int32)
        val = np.zeros(n.value, dtype=np.int32)
        self._ck(lib().smgpu_get_csr(self._h, name.encode(), _ptr(off), _ptr(val), C.byref(n)))
        return off, val

    def enable_boundary_smoothing(self, geometry, smoothing_patches, internal_smoothing_blending_fraction=0.0):
        """Boundary point smoothing (include/smgpu.h: smgpu_enable_boundary_smoothing).  geometry:
        dict(init_edges=(points, pairs), target_edges=(points, pairs), surface=(points, triangles));
        smoothing_patches: 0/1 per patch (shorter lists are padded with 0).  Collective on a processor mesh."""
        g, keep = _boundary_geometry(geometry)
        flags = np.zeros(self.n_patches, dtype=np.int32)
        k = min(len(smoothing_patches), flags.size)
        flags[:k] = np.asarray(smoothing_patches, dtype=np.int32)[:k]
        lib().smgpu_enable_boundary_smoothing.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
        self._ck(lib().smgpu_enable_boundary_smoothing(self._h, C.byref(g), _ptr(flags),
                                                      float(internal_smoothing_blending_fraction)))

    def boundary_classes(self):
        """(isCornerPoint, isFeatureEdgePoint) label lists, as the reference writes them with the mesh."""
        a, b = np.zeros(self.n_points, dtype=np.int32), np.zeros(self.n_points, dtype=np.int32)
        lib().smgpu_get_boundary_classes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self._ck(lib().smgpu_get_boundary_classes(self._h, _ptr(a), _ptr(b)))
        return a, b

    def tile_stats(self):
        out = (C.c_int64 * 8)()
        lib().smgpu_tile_stats.argtypes = [C.c_void_p, C.c_void_p]
        self._ck(lib().smgpu_tile_stats(self._h, out))
        return dict(tiles=int(out[0]), listed_faces=int(out[1]), listed_points=int(out[2]), smem_bytes=int(out[3]),
                    point_tiles=int(out[4]), pt_listed_points=int(out[5]), pt_listed_cells=int(out[6]),
                    edge_tile_smem_bytes=int(out[7]))

    def filter_stats(self):
        out = (C.c_int64 * 4)()
        lib().smgpu_filter_stats.argtypes = [C.c_void_p, C.c_void_p]
        self._ck(lib().smgpu_filter_stats(self._h, out))
        return dict(fused=bool(out[0]), suspect_points=int(out[1]), active_points=int(out[2]), paths=int(out[3]))

    def profile(self, enable=True):
        self._ck(lib().smgpu_profile(self._h, int(enable)))

    def profile_get(self):
        n = C.c_int32()
        names = (C.c_char_p * 16)()
        ms = (C.c_double * 16)()
        ln = (C.c_int64 * 16)()
        self._ck(lib().smgpu_profile_get(self._h, C.byref(n), names, ms, ln))
        return {names[i].decode(): dict(ms=ms[i], launches=ln[i]) for i in range(n.value)}

    # ---- multi-GPU ----
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        rc = lib().smgpu_comm_unique_id(buf)
        if rc != 0:
            raise SmoothMeshError(lib().smgpu_last_error().decode())
        return bytes(buf)

    def comm_local_shared(self) -> np.ndarray:
        """Global labels of this rank's processor-patch points (step 1 of the multi-GPU start-up)."""
        n = C.c_int64()
        self._ck(lib().smgpu_comm_local_shared(self._h, C.byref(n), None))
        g = np.zeros(max(n.value, 1), dtype=np.int64)
        self._ck(lib().smgpu_comm_local_shared(self._h, C.byref(n), _ptr(g)))
        return g[:n.value]

    def comm_prepare(self, rank, n_ranks, counts, all_gids):
        """Local half of the start-up (exchange plan, buffers); raises without touching any other rank."""
        counts = np.ascontiguousarray(counts, dtype=np.int64)
        all_gids = np.ascontiguousarray(all_gids, dtype=np.int64)
        if all_gids.size == 0:
            all_gids = np.zeros(1, dtype=np.int64)
        self._ck(lib().smgpu_comm_prepare(self._h, rank, n_ranks, _ptr(counts), _ptr(all_gids)))

    def comm_init(self, rank, n_ranks, unique_id: bytes, counts, all_gids):
        buf = (C.c_uint8 * 128)(*unique_id)
        counts = np.ascontiguousarray(counts, dtype=np.int64)
        all_gids = np.ascontiguousarray(all_gids, dtype=np.int64)
        if all_gids.size == 0:
            all_gids = np.zeros(1, dtype=np.int64)
        self._ck(lib().smgpu_comm_init(self._h, rank, n_ranks, buf, _ptr(counts), _ptr(all_gids)))

    def comm_abort(self):
        self._ck(lib().smgpu_comm_abort(self._h))

    def comm_p2p_export(self) -> bytes:
        buf = (C.c_uint8 * 64)()
        self._ck(lib().smgpu_comm_p2p_export(self._h, buf))
        return bytes(buf)

    def comm_p2p_connect(self, all_handles: bytes):
        """all_handles: the 64-byte IPC handles of all ranks, concatenated in rank order."""
        buf = (C.c_uint8 * len(all_handles))(*all_handles)
        self._ck(lib().smgpu_comm_p2p_connect(self._h, buf, None))

    def comm_p2p_disable(self):
        self._ck(lib().smgpu_comm_p2p_disable(self._h))


class Group:
    """In-process group (include/smgpu.h: smgpu_group_*): the processor meshes of one decomposed case as
    Smoothers on ONE device, rank = position in the list; the reference's `mpirun -np N` semantics without NCCL."""

    def __init__(self, smoothers):
        self.members = list(smoothers)
        arr = (C.c_void_p * len(self.members))(*[s._h for s in self.members])
        h = C.c_void_p()
        rc = lib().smgpu_group_create(arr, len(self.members), C.byref(h))
        if rc != 0:
            raise SmoothMeshError(f"smgpu_group_create failed ({rc}): {lib().smgpu_last_error().decode()}")
        self._h = h

    def iterate(self, max_iters) -> IterationLog:
        nf = np.zeros(max(max_iters, 1), dtype=np.int64)
        res = np.zeros(max(max_iters, 1), dtype=np.float64)
        done = C.c_int32()
        rc = lib().smgpu_group_iterate(self._h, int(max_iters), _ptr(nf), _ptr(res), C.byref(done))
        if rc != 0:
            raise SmoothMeshError(f"libsmgpu error {rc}: {lib().smgpu_last_error().decode()}")
        ms, ln = C.c_double(), C.c_int64()
        lib().smgpu_last_timing(self.members[0]._h, C.byref(ms), C.byref(ln))
        n = done.value
        return IterationLog(n, nf[:n].copy(), res[:n].copy(), ms.value, ln.value)

    def enable_boundary_smoothing(self, geometry, smoothing_patches, internal_smoothing_blending_fraction=0.0):
        """Boundary point smoothing on every member (collective set-up); smoothing_patches: 0/1 per physical patch."""
        g, keep = _boundary_geometry(geometry)
        flags = []
        for m in self.members:
            f = np.zeros(m.n_patches, dtype=np.int32)
            k = min(len(smoothing_patches), f.size)
            f[:k] = np.asarray(smoothing_patches, dtype=np.int32)[:k]
            flags.append(f)
        ptrs = (C.c_void_p * len(flags))(*[f.ctypes.data for f in flags])
        lib().smgpu_group_enable_boundary_smoothing.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
        rc = lib().smgpu_group_enable_boundary_smoothing(self._h, C.byref(g), ptrs, float(internal_smoothing_blending_fraction))
        if rc != 0:
            raise SmoothMeshError(f"libsmgpu error {rc}: {lib().smgpu_last_error().decode()}")

    def close(self):
        if getattr(self, "_h", None):
            lib().smgpu_group_destroy(self._h)
            self._h = None

    __del__ = close


def exchange_plan(rank, n_ranks, local, gids, counts, all_gids):
    """Host-only: (slot_point, slot_rank) of the exchange plan smgpu_comm_init builds."""
    local = np.ascontiguousarray(local, dtype=np.int32)
    gids = np.ascontiguousarray(gids, dtype=np.int64)
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    all_gids = np.ascontiguousarray(all_gids, dtype=np.int64)
    n = lib().smgpu_exchange_plan(rank, n_ranks, len(local), _ptr(local), _ptr(gids), _ptr(counts), _ptr(all_gids), None, None)
    if n < 0:
        raise SmoothMeshError(lib().smmesh_last_error().decode())
    sp, sr = np.zeros(max(n, 1), dtype=np.int32), np.zeros(max(n, 1), dtype=np.int32)
    lib().smgpu_exchange_plan(rank, n_ranks, len(local), _ptr(local), _ptr(gids), _ptr(counts), _ptr(all_gids), _ptr(sp), _ptr(sr))
    return sp[:n], sr[:n]
