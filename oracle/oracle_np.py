"""oracle_np -- TEST INFRASTRUCTURE ONLY.

A second, independent restatement of the smoothMesh iteration in plain Python/NumPy, written from
the prose contract in SURVEY.md section 8 (a0-a15) rather than from oracle/oracle.cpp, so that the
two can be checked against each other (pin K8).  Serial only, small meshes only (pure
Python loops).  It uses math.acos (libm), i.e. the reference's literal std::acos.

The prismatic boundary layer treatment (src/orthogonalBoundaryBlending.C) is restated
here as graph logic -- breadth-first hop levels, "unique parent" maps, per-iteration normals -- not as
the reference's sweeps, so that agreement with oracle.cpp checks the reading of the algorithm and not
just a transcription.
"""
from __future__ import annotations

import math

import numpy as np

GREAT, VSMALL, ROOTVSMALL = 1e15, 1e-300, 1e-150
COS_CLAMP = 0.99999


def _mag(v):
    return math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])


def _dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def _cross(a, b):
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])


def _clamped_acos(c):
    # std::max(-MAX, std::min(MAX, c)): NaN -> +MAX
    t = c if c < COS_CLAMP else COS_CLAMP
    t = t if -COS_CLAMP < t else -COS_CLAMP
    return math.acos(t)


def _same(a, b):
    return all(abs(a[i] - b[i]) <= VSMALL for i in range(3))


class NumpyOracle:
    def __init__(self, m, min_edge_length=-1.0, max_step_length=-1.0, rel_step_frac=0.5, min_angle_deg=35.0,
                 max_angle_deg=160.0, rel_tol=0.02, total_min_freeze=0, edge_angle_constraint=1,
                 face_angle_constraint=1, layer_patches=None, layer_max_blending_fraction=0.3, layer_edge_length=-1.0,
                 layer_expansion_ratio=1.3, min_layers=1, max_layers=4):
        self.x = np.array(m["points"], dtype=np.float64).reshape(-1, 3).copy()
        off, fv = m["face_offsets"], m["face_verts"]
        self.faces = [list(map(int, fv[off[i]:off[i + 1]])) for i in range(len(off) - 1)]
        self.own = list(map(int, m["owner"]))
        self.nei = list(map(int, m["neighbour"]))
        self.C = int(m["n_cells"])
        P = len(self.x)
        # a16: internal = not on a non-processor patch
        self.internal = np.ones(P, dtype=bool)
        for s, n, k in zip(m["patch_start"], m["patch_size"], m["patch_kind"]):
            if k == 1:
                continue
            for f in range(s, s + n):
                self.internal[self.faces[f]] = False
        # connectivity (appendix A.2)
        self.point_faces = [[] for _ in range(P)]
        for fi, f in enumerate(self.faces):
            for v in f:
                self.point_faces[v].append(fi)
        cells_of_face = [[self.own[f]] + ([self.nei[f]] if f < len(self.nei) else []) for f in range(len(self.faces))]
        self.point_cells = [sorted({c for f in pf for c in cells_of_face[f]}) for pf in self.point_faces]
        nbrs = [set() for _ in range(P)]
        for f in self.faces:
            for i, v in enumerate(f):
                w = f[(i + 1) % len(f)]
                nbrs[v].add(w)
                nbrs[w].add(v)
        self.point_points = [sorted(s) for s in nbrs]  # ascending neighbour label
        self.edge_faces = {}
        for fi, f in enumerate(self.faces):
            for i, v in enumerate(f):
                w = f[(i + 1) % len(f)]
                self.edge_faces.setdefault((min(v, w), max(v, w)), []).append(fi)
        self.cells_of_face = cells_of_face
        self.cell_points = [set() for _ in range(self.C)]
        for fi, f in enumerate(self.faces):
            for c in cells_of_face[fi]:
                self.cell_points[c].update(f)
        lens = [_mag(self.x[b] - self.x[a]) for (a, b) in self.edge_faces]
        self.min_edge = min(lens)
        self.min_edge_length = 0.5 * self.min_edge if min_edge_length < 0 else min_edge_length
        self.max_step_length = 0.3 * self.min_edge_length if max_step_length < 0 else max_step_length
        self.rel_step_frac, self.rel_tol = rel_step_frac, rel_tol
        self.small = math.pi * min_angle_deg / 180.0
        self.large = math.pi * max_angle_deg / 180.0
        self.total_min_freeze = bool(total_min_freeze)
        self.edge_angle, self.face_angle = bool(edge_angle_constraint), bool(face_angle_constraint)
        self.frozen = np.zeros(P, dtype=bool)
        self.patches = [(int(a), int(n), int(k)) for a, n, k in zip(m["patch_start"], m["patch_size"], m["patch_kind"])]
        self.layers = (layer_patches is not None and any(layer_patches) and layer_max_blending_fraction > 1e-15)
        if self.layers:
            self.layer_flags = [bool(layer_patches[i]) if i < len(layer_patches) else False for i in range(len(self.patches))]
            self.layer_frac, self.layer_ratio = layer_max_blending_fraction, layer_expansion_ratio
            self.layer_len = self.min_edge_length if layer_edge_length < 0 else layer_edge_length
            self.min_layers, self.max_layers = min_layers, max_layers
            self.setup_layers()

    # ---- prismatic boundary layer treatment ----
    def boundary_normals(self):
        """Per call: every boundary point adds the negated unit normals of its boundary faces (ascending face
        label) to whatever its normal was before -- the reference never zeroes the field -- then normals shorter
        than 0.1 at boundary points are dropped and every non-zero normal is rescaled to unit length."""
        touched = set()
        for a, n, k in self.patches:
            if k != 0:
                continue
            for f in range(a, a + n):
                unit = self.face_area[f] / _mag(self.face_area[f])
                for v in self.faces[f]:
                    self.normal[v] = self.normal[v] - unit
                    touched.add(v)
        for v in touched:
            if _mag(self.normal[v]) < 0.1:
                self.normal[v] = np.zeros(3)
        for v in range(len(self.x)):
            if not _same(self.normal[v], np.zeros(3)):
                self.normal[v] = self.normal[v] / _mag(self.normal[v])

    def setup_layers(self):
        P = len(self.x)
        # a boundary point belongs to the first patch (in patch order) that contains it
        first_patch = {}
        for pi, (a, n, k) in enumerate(self.patches):
            for f in range(a, a + n):
                for v in self.faces[f]:
                    first_patch.setdefault(v, pi)
        touches_internal = [any(self.internal[q] for q in self.point_points[v]) for v in range(P)]
        layer_surface = [(not self.internal[v]) and self.layer_flags[first_patch[v]] for v in range(P)]
        # hop levels: breadth-first from the layer-patch points that touch the interior, through internal
        # points only, max_layers + 1 levels deep
        hop = [-1] * P
        level = set()
        for pi, (a, n, k) in enumerate(self.patches):
            if self.layer_flags[pi]:
                for f in range(a, a + n):
                    level.update(v for v in self.faces[f] if (not self.internal[v]) and touches_internal[v])
        for v in level:
            hop[v] = 0
        for depth in range(1, self.max_layers + 2):
            nxt = {q for v in level for q in self.point_points[v] if hop[q] < 0 and self.internal[q]}
            for q in nxt:
                hop[q] = depth
            level = nxt
        self.hop = hop
        # set-up normals of the boundary points
        self.normal = np.zeros((P, 3))
        self.geometry()
        self.boundary_normals()
        # an internal point is tied to an outer point if exactly one neighbour lies one level closer to the
        # wall and that neighbour is internal or on a layer patch; an outer point claimed by two points
        # unties both, and an untied point unties everything that hangs below it
        self.outer = [-1] * P
        untied = set()
        claimed = {}
        for depth in range(1, self.max_layers + 2):
            for v in range(P):
                if hop[v] != depth:
                    continue
                parents = [q for q in self.point_points[v] if hop[q] == depth - 1]
                if len(parents) != 1:
                    continue
                q = parents[0]
                if not self.internal[q] and not layer_surface[q]:
                    continue
                if q in claimed:
                    untied.update((v, claimed[q]))
                    continue
                claimed[q] = v
                self.outer[v] = q
                if q in untied:
                    untied.add(v)
                else:
                    self.normal[v] = self.normal[q]
        for v in untied:
            self.normal[v] = np.zeros(3)
            self.outer[v] = -1

    def layer_blend(self):
        top = self.max_layers + 1
        for v in range(len(self.x)):
            if _same(self.normal[v], np.zeros(3)) or not self.internal[v] or self.hop[v] < 1:
                continue
            h = self.hop[v]
            length = self.layer_len * self.layer_ratio ** min(h - 1, top)
            slope = -self.layer_frac / (top - self.min_layers)
            frac = max(0.0, min(-slope * top + slope * h, self.layer_frac))
            ortho = self.x[self.outer[v]] + length * self.normal[v]
            self.new[v] = frac * ortho + (1.0 - frac) * self.new[v]
        for v in range(len(self.x)):  # second constrainMaxStepLength, every point
            d = self.new[v] - self.x[v]
            scale = self.max_step_length / (_mag(d) * self.rel_step_frac) if _mag(d) > self.max_step_length else 1.0
            self.new[v] = self.x[v] + (self.rel_step_frac * scale) * d

    # a0: OpenFOAM (openfoam.com) face centres / areas and cell centres
    def geometry(self):
        x = self.x
        fc, fa = [], []
        for f in self.faces:
            p = x[f]
            n = len(f)
            if n == 3:
                fc.append((1.0 / 3.0) * (p[0] + p[1] + p[2]))
                fa.append(0.5 * _cross(p[1] - p[0], p[2] - p[0]))
                continue
            est = p[0].copy()
            for i in range(1, n):
                est = est + p[i]
            est = est / float(n)
            sN, sA, sAc = np.zeros(3), 0.0, np.zeros(3)
            for i in range(n):
                a, b = p[i], p[(i + 1) % n]
                c = a + b + est
                nn = _cross(b - a, est - a)
                w = _mag(nn)
                sN, sA, sAc = sN + nn, sA + w, sAc + w * c
            if sA < ROOTVSMALL:
                fc.append(est)
                fa.append(np.zeros(3))
            else:
                fc.append(((1.0 / 3.0) * sAc) / sA)
                fa.append(0.5 * sN)
        F, Fi = len(self.faces), len(self.nei)
        est = np.zeros((self.C, 3))
        cnt = np.zeros(self.C)
        for f in range(F):
            est[self.own[f]] = est[self.own[f]] + fc[f]
            cnt[self.own[f]] += 1
        for f in range(Fi):
            est[self.nei[f]] = est[self.nei[f]] + fc[f]
            cnt[self.nei[f]] += 1
        est = np.array([est[c] / cnt[c] for c in range(self.C)])
        ctr = np.zeros((self.C, 3))
        vol = np.zeros(self.C)
        for f in range(F):
            c = self.own[f]
            pv = _dot(fa[f], fc[f] - est[c])
            ctr[c] = ctr[c] + pv * (0.75 * fc[f] + 0.25 * est[c])
            vol[c] += pv
        for f in range(Fi):
            c = self.nei[f]
            pv = _dot(fa[f], est[c] - fc[f])
            ctr[c] = ctr[c] + pv * (0.75 * fc[f] + 0.25 * est[c])
            vol[c] += pv
        self.cell_ctr = np.array([ctr[c] / vol[c] if abs(vol[c]) > VSMALL else est[c] for c in range(self.C)])
        self.face_area = fa

    # a1 + a3-a6: predictor
    def predict(self):
        x = self.x
        new = x.copy()
        for p in range(len(x)):
            cen = x[p]
            if self.internal[p] and self.point_cells[p]:
                s = np.zeros(3)
                for c in self.point_cells[p]:
                    s = s + self.cell_ctr[c]
                cen = s / float(len(self.point_cells[p]))
            # closest three eligible neighbours, stable order
            cand = [q for q in self.point_points[p] if self.internal[p] or not self.internal[q]]
            order = sorted(range(len(cand)), key=lambda k: _mag(x[p] - x[cand[k]]))  # Python's sort is stable
            q1, q2 = cand[order[0]], cand[order[1]]
            c1, c2 = x[q1] - x[p], x[q2] - x[p]
            c3 = x[cand[order[2]]] - x[p] if len(order) > 2 else np.array([GREAT] * 3)
            share = any(q2 in self.cell_points[c] for c in self.point_cells[q1])  # the two closest share a cell
            b = 0.0
            if not share and not _same(c1, np.zeros(3)) and not _same(c2, np.zeros(3)):
                r1, r2 = _mag(c2) / _mag(c1), _mag(c3) / _mag(c2)
                if self.internal[p]:
                    if r1 < 1.5 and r2 > 1.5:
                        b = min(1.0, max(0.0, (r2 - 1.5) / (3.0 - 1.5)))
                else:
                    b = min(1.0, max(0.0, (r1 - 1.0) / (2.0 - 1.0)))
            tgt = cen
            if b > 0.0:
                tgt = (1.0 - b) * cen + b * (x[p] + (c1 + c2) / 2.0)
            d = tgt - x[p]
            scale = self.max_step_length / (_mag(d) * self.rel_step_frac) if _mag(d) > self.max_step_length else 1.0
            new[p] = x[p] + (self.rel_step_frac * scale) * d
        self.new = new

    # a7
    def edge_shortening(self):
        x, new = self.x, self.new
        for p in range(len(x)):
            if self.frozen[p]:
                continue
            lc = min([GREAT] + [_mag(x[p] - x[q]) for q in self.point_points[p]])
            ln = min([GREAT] + [_mag(new[p] - x[q]) for q in self.point_points[p]])
            if self.total_min_freeze and min(ln, lc) < self.min_edge_length:
                self.frozen[p] = True
            elif ln < self.min_edge_length and ln < lc:
                self.frozen[p] = True

    @staticmethod
    def _corner_angle(c, a, b):
        u, v = a - c, b - c
        u, v = u / _mag(u), v / _mag(v)
        return _clamped_acos(_dot(u, v))

    # a8
    def edge_angles(self):
        x, new = self.x, self.new
        for p in range(len(x)):
            if self.frozen[p]:
                continue
            min_c = min_n = 1.7976931348623157e308
            for f in self.point_faces[p]:
                loop = self.faces[f]
                i = loop.index(p)
                a, b = loop[i - 1], loop[(i + 1) % len(loop)]
                cur = self._corner_angle(x[p], x[a], x[b])
                hyp = min(self._corner_angle(new[p], x[a], x[b]), self._corner_angle(new[p], new[a], new[b]),
                          self._corner_angle(new[p], x[a], new[b]), self._corner_angle(new[p], new[a], x[b]))
                min_c, min_n = min(min_c, cur), min(min_n, hyp)
            if min_n < self.small and min_n < min_c:
                self.frozen[p] = True

    # a10
    def edge_face_angles(self, edge, subst):
        a, b = edge
        pos = lambda v: subst.get(v, self.x[v])
        e0, e1 = pos(a), pos(b)
        mid = 0.5 * (e0 + e1)
        axis = (e1 - e0) / _mag(e1 - e0)

        def projected(pt):
            d = _dot(mid - pt, axis)
            q = pt + d * axis
            return (q - mid) / _mag(q - mid)

        fvec = {}
        for f in self.edge_faces[edge]:
            c = np.zeros(3)
            for v in self.faces[f]:
                c = c + pos(v)
            fvec[f] = projected(c / float(len(self.faces[f])))
        cells = []
        for f in self.edge_faces[edge]:
            for c in self.cells_of_face[f]:
                if c not in cells:
                    cells.append(c)
        lo, hi = 2.0 * math.pi, 0.0
        for c in cells:
            f0, f1 = [f for f in self.edge_faces[edge] if c in self.cells_of_face[f]]
            cv = projected(self.cell_ctr[c])
            ang = _clamped_acos(_dot(fvec[f0], cv)) + _clamped_acos(_dot(cv, fvec[f1]))
            lo, hi = min(lo, ang), max(hi, ang)
        return lo, hi

    def point_face_angles(self, p, subst):
        lo, hi = 2.0 * math.pi, 0.0
        for q in self.point_points[p]:
            a, b = self.edge_face_angles((min(p, q), max(p, q)), subst)
            lo, hi = min(lo, a), max(hi, b)
        return lo, hi

    # a11 + a12
    def face_angles(self):
        x, new = self.x, self.new
        P = len(x)
        cur_min, cur_max = np.full(P, 2.0 * math.pi), np.zeros(P)
        for e in self.edge_faces:
            a, b = self.edge_face_angles(e, {})
            for p in e:
                cur_min[p], cur_max[p] = min(cur_min[p], a), max(cur_max[p], b)
        self.cur_min, self.cur_max = cur_min, cur_max
        stack = list(range(P))
        while stack:
            p = stack.pop()
            if cur_min[p] > self.small and cur_max[p] < self.large:
                continue
            here = x[p] if self.frozen[p] else new[p]
            bad = lambda lo, hi: (lo < self.small and lo < cur_min[p]) or (hi > self.large and hi > cur_max[p])
            if not _same(here, x[p]):
                if bad(*self.point_face_angles(p, {p: here})):
                    here = x[p]
                    self.frozen[p] = True
            for q in self.point_points[p]:
                if self.frozen[q] or _same(new[q], x[q]):
                    continue
                if bad(*self.point_face_angles(p, {p: here, q: new[q]})):
                    self.frozen[q] = True
                    stack.append(q)

    def iterate(self, max_iters):
        n_frozen, residual = [], []
        for _ in range(max_iters):
            self.frozen[:] = False
            self.geometry()
            if self.layers:
                self.boundary_normals()
            self.predict()
            if self.layers:
                self.layer_blend()
            self.edge_shortening()
            if self.edge_angle:
                self.edge_angles()
            if self.face_angle:
                self.face_angles()
            keep = self.frozen | ~self.internal
            self.new[keep] = self.x[keep]
            n_frozen.append(int(keep.sum()))
            res = max(_mag(self.new[p] - self.x[p]) / self.max_step_length for p in range(len(self.x)))
            residual.append(res)
            self.x = self.new.copy()
            if res < self.rel_tol:
                break
        return len(n_frozen), np.array(n_frozen), np.array(residual)
