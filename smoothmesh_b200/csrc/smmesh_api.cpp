// smmesh_api.cpp -- C ABI (include/smmesh.h) over polymesh.cpp.
#include "../../include/smgpu.h"
#include "../../include/smmesh.h"
#include "polymesh.hpp"
#include "topology.hpp"
#include "boundary.hpp"

#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>

struct smmesh
{
    sm::PolyMesh m;
};

static thread_local std::string g_merr;

template <class Fn> static smmesh *guarded(Fn fn)
{
    try
    {
        smmesh *r = new smmesh;
        try
        {
            r->m = fn();
            r->m.check();
        }
        catch (...)
        {
            delete r;
            throw;
        }
        return r;
    }
    catch (const std::exception &e)
    {
        g_merr = e.what();
        return nullptr;
    }
}

extern "C"
{
    const char *smmesh_last_error(void) { return g_merr.c_str(); }
    void smmesh_free(smmesh *m) { delete m; }

    smmesh *smmesh_gen_hex_block(int32_t nx, int32_t ny, int32_t nz, const double lo[3], const double hi[3])
    {
        return guarded([&] { return sm::genHexBlock(nx, ny, nz, lo, hi); });
    }
    smmesh *smmesh_gen_hex_block_part(int32_t nx, int32_t ny, int32_t nz, int32_t px, int32_t py, int32_t pz, int32_t rank,
                                      const double lo[3], const double hi[3])
    {
        return guarded([&] { return sm::genHexBlockPart(nx, ny, nz, px, py, pz, rank, lo, hi); });
    }
    smmesh *smmesh_gen_kelvin(int32_t n, double h)
    {
        return guarded([&] { return sm::genKelvin(n, h); });
    }
    smmesh *smmesh_gen_kelvin_part(int32_t n, double h, int32_t px, int32_t py, int32_t pz, int32_t rank)
    {
        return guarded([&] { return sm::genKelvinPart(n, h, px, py, pz, rank); });
    }
    smmesh *smmesh_from_cells(int64_t n_points, const double *points, int32_t n_cells, const int32_t *cfo,
                              const int32_t *cvo, const int32_t *cv, const int32_t *cp, int32_t n_patches,
                              const char *const *names, const char *const *types)
    {
        return guarded([&] {
            std::vector<double> pts(points, points + 3 * n_points);
            std::vector<int32_t> a(cfo, cfo + n_cells + 1);
            const int32_t nf = a[n_cells];
            std::vector<int32_t> b(cvo, cvo + nf + 1), c(cv, cv + b[nf]);
            std::vector<int32_t> d;
            if (cp)
                d.assign(cp, cp + nf);
            std::vector<std::string> nm, ty;
            for (int i = 0; i < n_patches; ++i)
            {
                nm.push_back(names[i]);
                ty.push_back(types[i]);
            }
            return sm::buildFromCells(pts, a, b, c, d, nm, ty);
        });
    }
    smmesh *smmesh_from_arrays(int64_t n_points, const double *points, int64_t n_faces, const int32_t *face_offsets,
                               const int32_t *face_verts, const int32_t *owner, int64_t n_internal_faces,
                               const int32_t *neighbour, int64_t n_cells, int32_t n_patches, const int32_t *patch_start,
                               const int32_t *patch_size, const int32_t *patch_kind)
    {
        return guarded([&] {
            sm::PolyMesh m;
            m.points.assign(points, points + 3 * n_points);
            m.faceOffsets.assign(face_offsets, face_offsets + n_faces + 1);
            m.faceVerts.assign(face_verts, face_verts + face_offsets[n_faces]);
            m.owner.assign(owner, owner + n_faces);
            m.neighbour.assign(neighbour, neighbour + n_internal_faces);
            m.nCells = n_cells;
            for (int i = 0; i < n_patches; ++i)
            {
                sm::Patch p;
                p.name = "patch" + std::to_string(i);
                p.type = patch_kind[i] == SMGPU_PATCH_PROCESSOR ? "processor"
                         : patch_kind[i] == SMGPU_PATCH_EMPTY   ? "empty"
                                                                : "patch";
                p.start = patch_start[i];
                p.size = patch_size[i];
                m.patches.push_back(p);
            }
            return m;
        });
    }
    smmesh *smmesh_read(const char *dir)
    {
        return guarded([&] { return sm::readPolyMesh(dir); });
    }
    int smmesh_read_points(smmesh *m, const char *file)
    {
        try
        {
            std::vector<double> pts = sm::readPoints(file);
            if (pts.size() != m->m.points.size())
                throw std::runtime_error("point count differs from the mesh topology");
            m->m.points.assign(pts.begin(), pts.end());
        }
        catch (const std::exception &e)
        {
            g_merr = e.what();
            return SMGPU_ERR_IO;
        }
        return SMGPU_OK;
    }
    int smmesh_write(const smmesh *m, const char *dir, int32_t binary, int32_t precision)
    {
        try
        {
            sm::writePolyMesh(m->m, dir, binary != 0, precision);
        }
        catch (const std::exception &e)
        {
            g_merr = e.what();
            return SMGPU_ERR_IO;
        }
        return SMGPU_OK;
    }
    int smmesh_write_points(const double *points, int64_t n_points, const char *dir, int32_t binary, int32_t precision,
                            const char *location)
    {
        try
        {
            sm::writePoints(points, n_points, dir, binary != 0, precision, location);
        }
        catch (const std::exception &e)
        {
            g_merr = e.what();
            return SMGPU_ERR_IO;
        }
        return SMGPU_OK;
    }
    int smmesh_jitter(smmesh *m, double amp, uint64_t seed)
    {
        sm::jitterInterior(m->m, amp, seed);
        return SMGPU_OK;
    }
    int64_t smmesh_size(const smmesh *m, int32_t what)
    {
        switch (what)
        {
        case 0:
            return m->m.nPoints();
        case 1:
            return m->m.nCells;
        case 2:
            return m->m.nFaces();
        case 3:
            return m->m.nInternalFaces();
        case 4:
            return (int64_t)m->m.faceVerts.size();
        case 5:
            return (int64_t)m->m.patches.size();
        }
        return -1;
    }
    const double *smmesh_points(const smmesh *m) { return m->m.points.data(); }
    double *smmesh_points_mut(smmesh *m) { return m->m.points.data(); }
    const int32_t *smmesh_face_offsets(const smmesh *m) { return m->m.faceOffsets.data(); }
    const int32_t *smmesh_face_verts(const smmesh *m) { return m->m.faceVerts.data(); }
    const int32_t *smmesh_owner(const smmesh *m) { return m->m.owner.data(); }
    const int32_t *smmesh_neighbour(const smmesh *m) { return m->m.neighbour.data(); }
    int smmesh_patches(const smmesh *m, int32_t *start, int32_t *size, int32_t *kind)
    {
        for (size_t i = 0; i < m->m.patches.size(); ++i)
        {
            start[i] = m->m.patches[i].start;
            size[i] = m->m.patches[i].size;
            kind[i] = m->m.patches[i].kind();
        }
        return SMGPU_OK;
    }
    const char *smmesh_patch_name(const smmesh *m, int32_t i) { return m->m.patches[i].name.c_str(); }
    const int64_t *smmesh_point_global_id(const smmesh *m)
    {
        return m->m.pointGlobalId.empty() ? nullptr : m->m.pointGlobalId.data();
    }
    const int64_t *smmesh_cell_global_id(const smmesh *m)
    {
        return m->m.cellGlobalId.empty() ? nullptr : m->m.cellGlobalId.data();
    }
    int smmesh_write_decomposed(smmesh *const *parts, int32_t n_parts, const char *case_dir, int32_t binary)
    {
        try
        {
            std::vector<sm::PolyMesh> v;
            for (int i = 0; i < n_parts; ++i)
                v.push_back(parts[i]->m);
            sm::writeDecomposedCase(v, case_dir, binary != 0);
        }
        catch (const std::exception &e)
        {
            g_merr = e.what();
            return SMGPU_ERR_IO;
        }
        return SMGPU_OK;
    }
    smmesh *smmesh_read_processor(const char *case_dir, int32_t k)
    {
        return guarded([&] { return sm::readProcessorMesh(case_dir, k); });
    }
    int smmesh_quality(const smmesh *m, double out[7])
    {
        try
        {
            const sm::MeshQuality q = sm::computeQuality(m->m);
            out[0] = q.maxNonOrtho;
            out[1] = q.avgNonOrtho;
            out[2] = q.maxSkewness;
            out[3] = q.minEdgeAngle;
            out[4] = q.minEdgeLength;
            out[5] = q.maxEdgeLength;
            out[6] = q.minVolume;
        }
        catch (const std::exception &e)
        {
            g_merr = e.what();
            return SMGPU_ERR_MESH;
        }
        return SMGPU_OK;
    }
    int smmesh_geom_tiles(const smmesh *m, int32_t max_cells, int32_t max_faces, int32_t max_points, int64_t out[8])
    {
        try
        {
            const sm::Topology t = sm::buildTopology(m->m);
            const sm::GeomTiles G = sm::buildGeomTiles(m->m, t, max_cells, max_faces, max_points, true);
            // invariants the fused geometry kernel relies on
            std::vector<int32_t> cellSeen(t.C, 0), faceStored(t.F, 0);
            int64_t maxFaces = 0, maxPoints = 0;
            for (int32_t k = 0; k < G.nTiles; ++k)
            {
                const int32_t fb = G.tileFaceOff[k], nf = G.tileFaceOff[k + 1] - fb;
                const int32_t pb = G.tilePointOff[k], np = G.tilePointOff[k + 1] - pb;
                maxPoints = std::max<int64_t>(maxPoints, np);
                if (np > max_points)
                    throw std::runtime_error("tile over its point budget");
                for (int32_t i = 0; i < nf; ++i)
                {
                    const int32_t f = G.tileFaces[fb + i] & 0x7fffffff;
                    const int32_t rb = G.faceRefOff[fb + i], nv = G.faceRefOff[fb + i + 1] - rb;
                    if (nv != m->m.faceOffsets[f + 1] - m->m.faceOffsets[f])
                        throw std::runtime_error("face vertex count differs");
                    for (int32_t q = 0; q < nv; ++q)
                        if (G.faceRef[rb + q] >= np || G.tilePoints[pb + G.faceRef[rb + q]] != m->m.faceVerts[m->m.faceOffsets[f] + q])
                            throw std::runtime_error("face vertex reference does not resolve to the face's vertex");
                }
                const int32_t nc = G.tileCellOff[k + 1] - G.tileCellOff[k];
                maxFaces = std::max<int64_t>(maxFaces, nf);
                if (nf > max_faces || nc > max_cells || nc < 1)
                    throw std::runtime_error("tile over budget");
                for (int32_t i = 0; i < nf; ++i)
                {
                    const int32_t w = G.tileFaces[fb + i];
                    if (i > 0 && (w & 0x7fffffff) <= (G.tileFaces[fb + i - 1] & 0x7fffffff))
                        throw std::runtime_error("tile faces not ascending");
                    if (w < 0)
                        ++faceStored[w & 0x7fffffff];
                }
                for (int32_t s = G.tileCellOff[k]; s < G.tileCellOff[k + 1]; ++s)
                {
                    const int32_t c = G.tileCells[s];
                    ++cellSeen[c];
                    if (G.slotOff[s + 1] - G.slotOff[s] != t.cfOff[c + 1] - t.cfOff[c])
                        throw std::runtime_error("slot face count differs from the cell's");
                    for (int32_t j = 0; j < t.cfOff[c + 1] - t.cfOff[c]; ++j)
                    {
                        const uint16_t ref = G.slotRef[G.slotOff[s] + j];
                        const int32_t w = t.cf[t.cfOff[c] + j];
                        if ((ref & 0x7fff) >= nf || (G.tileFaces[fb + (ref & 0x7fff)] & 0x7fffffff) != (w & 0x7fffffff) ||
                            ((ref & 0x8000) != 0) != (w < 0))
                            throw std::runtime_error("slot reference does not resolve to the cell's face");
                    }
                }
            }
            // the (edge, cell) pairs of the fused face-angle filter: every pair names an edge of the mesh, one of
            // its cells, and the two faces findCellFacePair (src/smoothMesh.C:1042-1097) returns for them; all
            // pairs of the mesh are listed exactly once
            if (G.nTiles > 0 && !G.cellEdgeOff.empty())
            {
                int64_t listed = 0;
                for (int32_t k = 0; k < G.nTiles; ++k)
                {
                    const int32_t fb = G.tileFaceOff[k], pb = G.tilePointOff[k], cb = G.tileCellOff[k], nc = G.tileCellOff[k + 1] - cb;
                    for (int32_t i = 0; i < nc; ++i)
                    {
                        const int32_t slot = cb + i, c = G.tileCells[slot];
                        const int32_t n = G.cellEdgeOff[slot + 1] - G.cellEdgeOff[slot];
                        for (int32_t j = 0; j < n; ++j)
                        {
                            const size_t at = (size_t)G.cellEdgeOff[slot] + j;
                            const uint16_t *r = &G.cellEdgeRef[4 * at];
                            const int32_t p0 = G.tilePoints[pb + r[0]], p1 = G.tilePoints[pb + r[1]];
                            const int32_t f0 = G.tileFaces[fb + r[2]] & 0x7fffffff, f1 = G.tileFaces[fb + r[3]] & 0x7fffffff;
                            int32_t e = -1;
                            for (int32_t q = t.ppOff[p0]; q < t.ppOff[p0 + 1]; ++q)
                                if (t.pp[q] == p1)
                                    e = t.pe[q];
                            if (e < 0 || p0 >= p1)
                                throw std::runtime_error("(edge, cell) pair does not name an edge");
                            bool found = false;
                            for (int32_t q = t.ecOff[e]; q < t.ecOff[e + 1]; ++q)
                                if (t.ecCell[q] == c)
                                {
                                    const int32_t g0 = t.ef[t.efOff[e] + (t.ecPair[q] & 0xffff)], g1 = t.ef[t.efOff[e] + ((t.ecPair[q] >> 16) & 0xffff)];
                                    found = (g0 == f0 && g1 == f1) || (g0 == f1 && g1 == f0);
                                }
                            if (!found)
                                throw std::runtime_error("(edge, cell) pair does not match the cell's face pair at that edge");
                            ++listed;
                        }
                    }
                }
                if (listed != (int64_t)t.ecCell.size())
                    throw std::runtime_error("(edge, cell) pairs do not cover the mesh exactly once");
            }
            // uniform tiles: the fixed-stride copies equal the offset-table entries, and are laid out back to back
            if (G.nTiles > 0 && !G.tileUCellOff.empty())
            {
                int64_t uf = 0, uc = 0;
                for (int32_t k = 0; k < G.nTiles; ++k)
                {
                    if (G.tileUCellOff[k] < 0)
                        continue;
                    if (G.tileUCellOff[k] != uc || G.tileUFaceOff[k] != uf)
                        throw std::runtime_error("uniform tile offsets are not consecutive");
                    const int32_t fb = G.tileFaceOff[k], nf = G.tileFaceOff[k + 1] - fb, cb = G.tileCellOff[k], nc = G.tileCellOff[k + 1] - cb;
                    for (int32_t i = 0; i < nf; ++i)
                    {
                        if (G.faceRefOff[fb + i + 1] - G.faceRefOff[fb + i] != 4)
                            throw std::runtime_error("uniform tile with a face that is not a quadrilateral");
                        for (int q = 0; q < 4; ++q)
                            if (G.uFaceRef[4 * (uf + i) + q] != G.faceRef[G.faceRefOff[fb + i] + q])
                                throw std::runtime_error("uniform tile: face reference copy differs");
                    }
                    for (int32_t i = 0; i < nc; ++i)
                    {
                        if (G.slotOff[cb + i + 1] - G.slotOff[cb + i] != 6)
                            throw std::runtime_error("uniform tile with a cell that has not six faces");
                        for (int q = 0; q < 6; ++q)
                            if (G.uSlotRef[6 * (uc + i) + q] != G.slotRef[G.slotOff[cb + i] + q])
                                throw std::runtime_error("uniform tile: cell reference copy differs");
                    }
                    uf += nf;
                    uc += nc;
                }
                if (uc != G.nUniformCells || (int64_t)G.hexRec.size() != 16 * uc)
                    throw std::runtime_error("uniform cell count differs");
            }
            if (G.nTiles > 0)
            {
                for (int32_t c = 0; c < t.C; ++c)
                    if (cellSeen[c] != 1)
                        throw std::runtime_error("cell not in exactly one tile");
                for (int32_t f = 0; f < t.F; ++f)
                    if (faceStored[f] != 1)
                        throw std::runtime_error("face outputs not stored by exactly one tile");
            }
            out[0] = G.nTiles;
            out[1] = (int64_t)G.tileFaces.size();
            out[2] = maxFaces;
            out[3] = t.F;
            out[4] = maxPoints;
            out[5] = (int64_t)G.cellEdgeRef.size() / 4;
            out[6] = G.uniformCellEdges;
            out[7] = G.nUniformCells;
        }
        catch (const std::exception &e)
        {
            g_merr = e.what();
            return SMGPU_ERR_MESH;
        }
        return SMGPU_OK;
    }
    int smmesh_boundary_setup(const smmesh *m, int64_t n_init_points, const double *init_points, int64_t n_init_edges,
                              const int32_t *init_edges, int64_t n_target_points, const double *target_points,
                              int64_t n_target_edges, const int32_t *target_edges, const int32_t *patch_smoothing,
                              double layer_edge_length, uint8_t *is_corner, uint8_t *is_feature_edge,
                              uint8_t *is_smoothing_surface, double *corner_points, int32_t *point_strings,
                              int32_t *hops_to_smoothing, int32_t *point_to_inner, int32_t *target_edge_strings)
    {
        try
        {
            const sm::Topology t = sm::buildTopology(m->m);
            sm::EdgeMesh ie, te;
            ie.points.assign(init_points, init_points + 3 * n_init_points);
            ie.edges.assign(init_edges, init_edges + 2 * n_init_edges);
            ie.finish();
            te.points.assign(target_points, target_points + 3 * n_target_points);
            te.edges.assign(target_edges, target_edges + 2 * n_target_edges);
            te.finish();
            std::vector<int32_t> ps(patch_smoothing, patch_smoothing + m->m.patches.size());
            const double lel = layer_edge_length < 0 ? 0.5 * t.minEdgeLength : layer_edge_length;
            const sm::BoundarySetup B = sm::buildBoundarySetup(m->m, t, std::vector<double>(m->m.points.begin(), m->m.points.end()), ie, te, sm::TriSurface(), ps, lel);
            auto put = [](auto *dst, const auto &src) {
                if (dst)
                    std::copy(src.begin(), src.end(), dst);
            };
            put(is_corner, B.isCorner);
            put(is_feature_edge, B.isFeatureEdge);
            put(is_smoothing_surface, B.isSmoothingSurface);
            put(corner_points, B.cornerPoints);
            put(point_strings, B.pointStrings);
            put(hops_to_smoothing, B.hopsToSmoothing);
            put(point_to_inner, B.pointToInner);
            put(target_edge_strings, B.targetEdgeStrings);
        }
        catch (const std::exception &e)
        {
            g_merr = e.what();
            return SMGPU_ERR_MESH;
        }
        return SMGPU_OK;
    }
    int smmesh_layer_setup(const smmesh *m, const int32_t *patch_layer, int32_t max_layers, int32_t *hops,
                           int32_t *point_to_outer, int32_t *normal_source)
    {
        try
        {
            const sm::Topology t = sm::buildTopology(m->m);
            const std::vector<int32_t> flags(patch_layer, patch_layer + m->m.patches.size());
            const sm::LayerSetup L = sm::buildLayerSetup(m->m, t, flags, max_layers);
            if (hops)
                std::copy(L.hops.begin(), L.hops.end(), hops);
            if (point_to_outer)
                std::copy(L.pointToOuter.begin(), L.pointToOuter.end(), point_to_outer);
            if (normal_source)
                std::copy(L.normalSrc.begin(), L.normalSrc.end(), normal_source);
        }
        catch (const std::exception &e)
        {
            g_merr = e.what();
            return SMGPU_ERR_MESH;
        }
        return SMGPU_OK;
    }
    int smmesh_ray_cast(int64_t n_points, const double *points, int64_t n_tris, const int32_t *tris, int64_t n_rays,
                        const double *start, const double *end, int32_t use_bvh, int32_t *hit_tri, double *hit_point)
    {
        try
        {
            sm::TriSurface s;
            s.points.assign(points, points + 3 * n_points);
            s.tris.assign(tris, tris + 3 * n_tris);
            const sm::TriangleBvh bvh = use_bvh ? sm::buildTriangleBvh(s) : sm::TriangleBvh();
#pragma omp parallel for schedule(dynamic, 64)
            for (int64_t i = 0; i < n_rays; ++i)
            {
                double h[3] = {0, 0, 0};
                hit_tri[i] = sm::segmentSurfaceHit(s, use_bvh ? &bvh : nullptr, start + 3 * i, end + 3 * i, h);
                hit_point[3 * i] = h[0], hit_point[3 * i + 1] = h[1], hit_point[3 * i + 2] = h[2];
            }
        }
        catch (const std::exception &e)
        {
            g_merr = e.what();
            return SMGPU_ERR_MESH;
        }
        return SMGPU_OK;
    }
    int64_t smmesh_read_label_list(const char *file, int32_t *data, int64_t capacity)
    {
        try
        {
            const std::vector<int32_t> v = sm::readLabelIOList(file);
            if (data)
                std::copy(v.begin(), v.begin() + std::min<int64_t>(capacity, (int64_t)v.size()), data);
            return (int64_t)v.size();
        }
        catch (const std::exception &e)
        {
            g_merr = e.what();
            return -1;
        }
    }
    int smmesh_write_label_list(const char *file, const char *object, const char *location, const int32_t *data, int64_t n,
                                int32_t binary)
    {
        try
        {
            sm::writeLabelIOList(file, object, location, std::vector<int32_t>(data, data + n), binary != 0);
        }
        catch (const std::exception &e)
        {
            g_merr = e.what();
            return SMGPU_ERR_MESH;
        }
        return SMGPU_OK;
    }
    int smmesh_read_obj(const char *file, int64_t *n_points, double *points, int64_t *n_edges, int32_t *edges,
                        int64_t *n_tris, int32_t *tris)
    {
        try
        {
            std::vector<double> p;
            std::vector<int32_t> e, t;
            sm::readObj(file, p, e, t);
            if (n_points)
                *n_points = (int64_t)p.size() / 3;
            if (n_edges)
                *n_edges = (int64_t)e.size() / 2;
            if (n_tris)
                *n_tris = (int64_t)t.size() / 3;
            if (points)
                std::copy(p.begin(), p.end(), points);
            if (edges)
                std::copy(e.begin(), e.end(), edges);
            if (tris)
                std::copy(t.begin(), t.end(), tris);
        }
        catch (const std::exception &e)
        {
            g_merr = e.what();
            return SMGPU_ERR_MESH;
        }
        return SMGPU_OK;
    }
    smmesh *smmesh_renumber(const smmesh *m, int32_t *point_old_of_new, int32_t *cell_old_of_new)
    {
        return guarded([&] {
            std::vector<int32_t> pm, cm;
            sm::PolyMesh o = sm::renumberMorton(m->m, pm, cm);
            if (point_old_of_new)
                std::copy(pm.begin(), pm.end(), point_old_of_new);
            if (cell_old_of_new)
                std::copy(cm.begin(), cm.end(), cell_old_of_new);
            return o;
        });
    }
    int smmesh_decompose(const smmesh *m, int32_t method, int32_t px, int32_t py, int32_t pz, smmesh **parts_out)
    {
        try
        {
            const int n = method == 0 ? px * py * pz : px;
            std::vector<int32_t> part = method == 0 ? sm::partitionBricks(m->m, px, py, pz) : sm::partitionRCB(m->m, n);
            std::vector<sm::PolyMesh> parts = sm::decompose(m->m, part, n);
            for (int i = 0; i < n; ++i)
            {
                parts[i].check();
                parts_out[i] = new smmesh{std::move(parts[i])};
            }
        }
        catch (const std::exception &e)
        {
            g_merr = e.what();
            return SMGPU_ERR_MESH;
        }
        return SMGPU_OK;
    }
}

// ---- host-only view of the multi-GPU exchange plan (declared in smgpu.h) ----
#include "exchange.hpp"
extern "C" int64_t smgpu_exchange_plan(int32_t rank, int32_t n_ranks, int64_t n_local, const int32_t *local,
                                       const int64_t *gids, const int64_t *counts, const int64_t *all_gids,
                                       int32_t *slot_point, int32_t *slot_rank)
{
    try
    {
        std::vector<int32_t> loc(local, local + n_local);
        std::vector<int64_t> g(gids, gids + n_local), cnt(counts, counts + n_ranks);
        int64_t total = 0;
        for (int64_t c : cnt)
            total += c;
        std::vector<int64_t> all(all_gids, all_gids + total);
        const sm::ExchangePlan pl = sm::buildExchangePlan(rank, n_ranks, loc, g, cnt, all);
        for (size_t j = 0; j < pl.nbrRank.size(); ++j)
            for (int32_t s = pl.nbrOff[j]; s < pl.nbrOff[j + 1]; ++s)
            {
                if (slot_point)
                    slot_point[s] = pl.sendPoint[s];
                if (slot_rank)
                    slot_rank[s] = pl.nbrRank[j];
            }
        return (int64_t)pl.sendPoint.size();
    }
    catch (const std::exception &e)
    {
        g_merr = e.what();
        return -1;
    }
}
