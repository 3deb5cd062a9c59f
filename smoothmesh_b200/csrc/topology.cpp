// topology.cpp -- see topology.hpp.
#include "topology.hpp"
#include "sm_math.h"

#include <algorithm>
#include <array>
#include <parallel/algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

namespace sm
{

static void fail(const std::string &s) { throw std::runtime_error(s); }

Topology buildTopology(const PolyMesh &m)
{
    // SMGPU_TIMING=1: wall time of every set-up phase on stderr
    const bool timing = getenv("SMGPU_TIMING") && atoi(getenv("SMGPU_TIMING")) != 0;
    auto clock = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double tPrev = clock();
    auto tick = [&](const char *what) {
        if (!timing)
            return;
        const double tNow = clock();
        fprintf(stderr, "[smgpu set-up] %-28s %.3f s\n", what, tNow - tPrev);
        tPrev = tNow;
    };
    m.check();
    tick("mesh check");
    Topology t;
    const int64_t P = m.nPoints(), C = m.nCells, F = m.nFaces(), Fi = m.nInternalFaces();
    t.P = P;
    t.C = C;
    t.F = F;
    t.Fi = Fi;
    t.faceOff = m.faceOffsets;
    t.faceVerts = m.faceVerts;
    const int64_t FV = (int64_t)m.faceVerts.size();
    if (2 * FV >= (int64_t)INT32_MAX)
        fail("mesh too large for 32-bit offsets");

    // src/smoothMesh.C:52-80: internal = not on any non-processor patch; empty patches abort
    t.isInternal.assign(P, 1);
    for (const Patch &p : m.patches)
    {
        if (p.kind() == PATCH_PROCESSOR)
            continue;
        if (p.kind() == PATCH_EMPTY)
            fail("Smoothing of non-3D meshes (meshes with type empty patches) is not supported");
        for (int32_t f = p.start; f < p.start + p.size; ++f)
            for (int32_t k = m.faceOffsets[f]; k < m.faceOffsets[f + 1]; ++k)
                t.isInternal[m.faceVerts[k]] = 0;
    }

    // points on processor patches: the candidates for inter-rank sharing
    {
        std::vector<uint8_t> onProc(P, 0);
        for (const Patch &p : m.patches)
            if (p.kind() == PATCH_PROCESSOR)
                for (int32_t f = p.start; f < p.start + p.size; ++f)
                    for (int32_t k = m.faceOffsets[f]; k < m.faceOffsets[f + 1]; ++k)
                        onProc[m.faceVerts[k]] = 1;
        for (int64_t p = 0; p < P; ++p)
            if (onProc[p])
                t.procPoints.push_back((int32_t)p);
    }

    // ---- point -> face corners (pointFaces ascending) ----
    // counted and filled in parallel (atomic cursors), then every row is put in ascending face order
    t.cornerOff.assign(P + 1, 0);
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < FV; ++k)
    {
#pragma omp atomic update
        ++t.cornerOff[m.faceVerts[k] + 1];
    }
    for (int64_t p = 0; p < P; ++p)
        t.cornerOff[p + 1] += t.cornerOff[p];
    t.corner.resize(2 * FV);
    std::vector<int32_t> cornerFace(FV);
    {
        std::vector<int32_t> cur(t.cornerOff.begin(), t.cornerOff.end() - 1);
        int32_t maxFaceSize = 0;
#pragma omp parallel for schedule(static) reduction(max : maxFaceSize)
        for (int64_t f = 0; f < F; ++f)
        {
            const int32_t b = m.faceOffsets[f], n = m.faceOffsets[f + 1] - b;
            maxFaceSize = std::max(maxFaceSize, n);
            for (int32_t i = 0; i < n; ++i)
            {
                const int32_t v = m.faceVerts[b + i];
                int32_t slot;
#pragma omp atomic capture
                slot = cur[v]++;
                t.corner[2 * (int64_t)slot] = m.faceVerts[b + (i == 0 ? n - 1 : i - 1)];
                t.corner[2 * (int64_t)slot + 1] = m.faceVerts[b + (i == n - 1 ? 0 : i + 1)];
                cornerFace[slot] = (int32_t)f;
            }
        }
        t.maxFaceSize = maxFaceSize;
#pragma omp parallel for schedule(static)
        for (int64_t p = 0; p < P; ++p)
        { // insertion sort of the (short) row by (face, previous vertex, next vertex)
            const int32_t b = t.cornerOff[p], e = t.cornerOff[p + 1];
            for (int32_t i = b + 1; i < e; ++i)
            {
                const int32_t f = cornerFace[i], c0 = t.corner[2 * (int64_t)i], c1 = t.corner[2 * (int64_t)i + 1];
                int32_t j = i - 1;
                while (j >= b && (cornerFace[j] > f || (cornerFace[j] == f && (t.corner[2 * (int64_t)j] > c0 ||
                                                                              (t.corner[2 * (int64_t)j] == c0 && t.corner[2 * (int64_t)j + 1] > c1)))))
                {
                    cornerFace[j + 1] = cornerFace[j];
                    t.corner[2 * (int64_t)(j + 1)] = t.corner[2 * (int64_t)j];
                    t.corner[2 * (int64_t)(j + 1) + 1] = t.corner[2 * (int64_t)j + 1];
                    --j;
                }
                cornerFace[j + 1] = f;
                t.corner[2 * (int64_t)(j + 1)] = c0;
                t.corner[2 * (int64_t)(j + 1) + 1] = c1;
            }
        }
    }

    tick("point corners");
    // ---- point -> cells (ascending) and point -> points (ascending) ----
    std::vector<int32_t> pcCount(P), ppCount(P);
#pragma omp parallel
    {
        std::vector<int32_t> tmp;
#pragma omp for schedule(static)
        for (int64_t p = 0; p < P; ++p)
        {
            tmp.clear();
            for (int32_t s = t.cornerOff[p]; s < t.cornerOff[p + 1]; ++s)
            {
                const int32_t f = cornerFace[s];
                tmp.push_back(m.owner[f]);
                if (f < Fi)
                    tmp.push_back(m.neighbour[f]);
            }
            std::sort(tmp.begin(), tmp.end());
            pcCount[p] = (int32_t)(std::unique(tmp.begin(), tmp.end()) - tmp.begin());
            tmp.clear();
            for (int32_t s = 2 * t.cornerOff[p]; s < 2 * t.cornerOff[p + 1]; ++s)
                tmp.push_back(t.corner[s]);
            std::sort(tmp.begin(), tmp.end());
            ppCount[p] = (int32_t)(std::unique(tmp.begin(), tmp.end()) - tmp.begin());
        }
    }
    t.pcOff.assign(P + 1, 0);
    t.ppOff.assign(P + 1, 0);
    for (int64_t p = 0; p < P; ++p)
    {
        t.pcOff[p + 1] = t.pcOff[p] + pcCount[p];
        t.ppOff[p + 1] = t.ppOff[p] + ppCount[p];
        t.maxPointDegree = std::max(t.maxPointDegree, ppCount[p]);
    }
    t.pc.resize(t.pcOff[P]);
    t.pp.resize(t.ppOff[P]);
    std::vector<int32_t> upStart(P + 1, 0); // first edge label of point p = #edges (a,b) with a<p
#pragma omp parallel
    {
        std::vector<int32_t> tmp;
#pragma omp for schedule(static)
        for (int64_t p = 0; p < P; ++p)
        {
            tmp.clear();
            for (int32_t s = t.cornerOff[p]; s < t.cornerOff[p + 1]; ++s)
            {
                const int32_t f = cornerFace[s];
                tmp.push_back(m.owner[f]);
                if (f < Fi)
                    tmp.push_back(m.neighbour[f]);
            }
            std::sort(tmp.begin(), tmp.end());
            tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
            std::copy(tmp.begin(), tmp.end(), t.pc.begin() + t.pcOff[p]);
            tmp.clear();
            for (int32_t s = 2 * t.cornerOff[p]; s < 2 * t.cornerOff[p + 1]; ++s)
                tmp.push_back(t.corner[s]);
            std::sort(tmp.begin(), tmp.end());
            tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
            std::copy(tmp.begin(), tmp.end(), t.pp.begin() + t.ppOff[p]);
            int32_t up = 0;
            for (int32_t q : tmp)
                up += (q > (int32_t)p);
            upStart[p + 1] = up;
        }
    }
    for (int64_t p = 0; p < P; ++p)
        upStart[p + 1] += upStart[p];
    const int64_t E = upStart[P];
    t.E = E;
    if ((int64_t)t.pp.size() != 2 * E)
        fail("inconsistent edge connectivity (pointPoints is not symmetric)");

    tick("pointCells / pointPoints");
    // ---- edges, pointEdges ----
    t.edge.resize(2 * E);
    t.pe.resize(t.pp.size());
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < P; ++p)
    {
        int32_t e = upStart[p];
        for (int32_t s = t.ppOff[p]; s < t.ppOff[p + 1]; ++s)
        {
            const int32_t q = t.pp[s];
            if (q > p)
            {
                t.edge[2 * (int64_t)e] = (int32_t)p;
                t.edge[2 * (int64_t)e + 1] = q;
                t.pe[s] = e++;
            }
        }
    }
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < P; ++p)
        for (int32_t s = t.ppOff[p]; s < t.ppOff[p + 1]; ++s)
        {
            const int32_t q = t.pp[s];
            if (q < p)
            { // label lives in q's row: upStart[q] + rank of p among q's higher neighbours
                const int32_t *b = &t.pp[t.ppOff[q]], *e = &t.pp[t.ppOff[q + 1]];
                const int32_t *firstUp = std::upper_bound(b, e, q);
                const int32_t *pos = std::lower_bound(firstUp, e, (int32_t)p);
                t.pe[s] = upStart[q] + (int32_t)(pos - firstUp);
            }
        }

    tick("edges");
    // ---- edge -> faces (ascending), edge -> cells with face pairs ----
    std::vector<int32_t> efCount(E), ecCount(E);
    auto edgeFacesOf = [&](int64_t e, int32_t *out) {
        const int32_t p = t.edge[2 * e], q = t.edge[2 * e + 1];
        int32_t n = 0;
        for (int32_t s = t.cornerOff[p]; s < t.cornerOff[p + 1]; ++s)
            if (t.corner[2 * (int64_t)s] == q || t.corner[2 * (int64_t)s + 1] == q)
                out[n++] = cornerFace[s];
        return n;
    };
    int32_t maxRow = 0;
    for (int64_t p = 0; p < P; ++p)
        maxRow = std::max(maxRow, t.cornerOff[p + 1] - t.cornerOff[p]);
#pragma omp parallel
    {
        std::vector<int32_t> fs(maxRow), cs;
#pragma omp for schedule(static)
        for (int64_t e = 0; e < E; ++e)
        {
            const int32_t n = edgeFacesOf(e, fs.data());
            efCount[e] = n;
            cs.clear();
            for (int32_t i = 0; i < n; ++i)
            {
                const int32_t f = fs[i];
                if (std::find(cs.begin(), cs.end(), m.owner[f]) == cs.end())
                    cs.push_back(m.owner[f]);
                if (f < Fi && std::find(cs.begin(), cs.end(), m.neighbour[f]) == cs.end())
                    cs.push_back(m.neighbour[f]);
            }
            ecCount[e] = (int32_t)cs.size();
        }
    }
    t.efOff.assign(E + 1, 0);
    t.ecOff.assign(E + 1, 0);
    for (int64_t e = 0; e < E; ++e)
    {
        t.efOff[e + 1] = t.efOff[e] + efCount[e];
        t.ecOff[e + 1] = t.ecOff[e] + ecCount[e];
        t.maxEdgeFaces = std::max(t.maxEdgeFaces, efCount[e]);
    }
    t.ef.resize(t.efOff[E]);
    t.ecCell.resize(t.ecOff[E]);
    t.ecPair.resize(t.ecOff[E]);
    int bad = 0;
#pragma omp parallel
    {
        std::vector<int32_t> fs(maxRow), cs;
#pragma omp for schedule(static)
        for (int64_t e = 0; e < E; ++e)
        {
            const int32_t n = edgeFacesOf(e, fs.data());
            std::copy(fs.begin(), fs.begin() + n, t.ef.begin() + t.efOff[e]);
            cs.clear();
            for (int32_t i = 0; i < n; ++i)
            {
                const int32_t f = fs[i];
                if (std::find(cs.begin(), cs.end(), m.owner[f]) == cs.end())
                    cs.push_back(m.owner[f]);
                if (f < Fi && std::find(cs.begin(), cs.end(), m.neighbour[f]) == cs.end())
                    cs.push_back(m.neighbour[f]);
            }
            for (size_t ci = 0; ci < cs.size(); ++ci)
            {
                const int32_t c = cs[ci];
                int32_t f0 = -1, f1 = -1, cnt = 0;
                for (int32_t i = 0; i < n; ++i)
                {
                    const int32_t f = fs[i];
                    if (m.owner[f] == c || (f < Fi && m.neighbour[f] == c))
                    {
                        if (cnt == 0)
                            f0 = i;
                        else if (cnt == 1)
                            f1 = i;
                        ++cnt;
                    }
                }
                if (cnt != 2 || n >= 65536)
                {
#pragma omp atomic write
                    bad = (cnt > 2) ? 1 : 2;
                }
                t.ecCell[t.ecOff[e] + ci] = c;
                t.ecPair[t.ecOff[e] + ci] = (f0 & 0xffff) | (f1 << 16);
            }
        }
    }
    if (bad == 1)
        fail("Sanity broken, more than two edge faces belong to same cell");
    if (bad == 2)
        fail("Sanity broken, didn't find face pairs for cell");

    tick("edgeFaces / edgeCells");
    // ---- cell -> faces in OpenFOAM's accumulation order (owned faces ascending, then
    //      neighbour-side faces ascending; bit 31 marks the neighbour side) ----
    t.cfOff.assign(C + 1, 0);
#pragma omp parallel for schedule(static)
    for (int64_t f = 0; f < F; ++f)
    {
#pragma omp atomic update
        ++t.cfOff[m.owner[f] + 1];
        if (f < Fi)
        {
#pragma omp atomic update
            ++t.cfOff[m.neighbour[f] + 1];
        }
    }
    for (int64_t c = 0; c < C; ++c)
        t.cfOff[c + 1] += t.cfOff[c];
    t.cf.resize(t.cfOff[C]);
    {
        // filled in parallel, then every (short) row sorted: owned faces ascending first (bit 31 clear), then the
        // neighbour-side faces ascending -- as unsigned numbers that is exactly ascending order
        std::vector<int32_t> cur(t.cfOff.begin(), t.cfOff.end() - 1);
#pragma omp parallel for schedule(static)
        for (int64_t f = 0; f < F; ++f)
        {
            int32_t slot;
#pragma omp atomic capture
            slot = cur[m.owner[f]]++;
            t.cf[slot] = (int32_t)f;
            if (f < Fi)
            {
#pragma omp atomic capture
                slot = cur[m.neighbour[f]]++;
                t.cf[slot] = (int32_t)((uint32_t)f | 0x80000000u);
            }
        }
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < C; ++c)
        {
            uint32_t *b = reinterpret_cast<uint32_t *>(&t.cf[t.cfOff[c]]), *e = reinterpret_cast<uint32_t *>(&t.cf[t.cfOff[c + 1]]);
            std::sort(b, e);
        }
    }

    tick("cellFaces");
    // ---- findClosestPoints prerequisite (:354-362): two eligible neighbours per point ----
    for (int64_t p = 0; p < P; ++p)
    {
        int32_t eligible = 0;
        for (int32_t s = t.ppOff[p]; s < t.ppOff[p + 1]; ++s)
            eligible += (t.isInternal[p] || !t.isInternal[t.pp[s]]);
        if (eligible < 2)
            fail("Failed to find cLabel" + std::to_string(eligible + 1) + " for pointI " + std::to_string(p));
    }

    tick("eligibility check");
    // ---- getMeshStats :1495-1510 ----
    double mn = 1e300, mx = 0.0;
#pragma omp parallel for reduction(min : mn) reduction(max : mx) schedule(static)
    for (int64_t e = 0; e < E; ++e)
    {
        const double *a = &m.points[3 * (int64_t)t.edge[2 * e]], *b = &m.points[3 * (int64_t)t.edge[2 * e + 1]];
        const double dx = b[0] - a[0], dy = b[1] - a[1], dz = b[2] - a[2];
        const double len = std::sqrt(dx * dx + dy * dy + dz * dz);
        mn = std::min(mn, len);
        mx = std::max(mx, len);
    }
    t.minEdgeLength = mn;
    t.maxEdgeLength = mx;

    tick("mesh stats");
    // ---- fixed-size records for the common low-valence case (see topology.hpp) ----
    t.pointRec.assign(16 * P, 0);
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < P; ++p)
    {
        int32_t *r = &t.pointRec[16 * p];
        const int32_t npc = t.pcOff[p + 1] - t.pcOff[p], npp = t.ppOff[p + 1] - t.ppOff[p];
        const bool generic = npc > 8 || npp > 6;
        if (!generic)
        {
            for (int32_t k = 0; k < npc; ++k)
                r[k] = t.pc[t.pcOff[p] + k];
            for (int32_t k = 0; k < npp; ++k)
                r[8 + k] = t.pp[t.ppOff[p] + k];
        }
        r[14] = npc | (npp << 8) | (generic ? (int32_t)0x80000000u : 0);
        // corners (prev,next) as a mask over the 15 unordered pairs of the (<= 6) row positions
        if (!generic)
        {
            int32_t mask = 0;
            const int32_t *b = &t.pp[t.ppOff[p]];
            for (int32_t k = t.cornerOff[p]; k < t.cornerOff[p + 1]; ++k)
            {
                int32_t sa = (int32_t)(std::lower_bound(b, b + npp, t.corner[2 * (int64_t)k]) - b);
                int32_t sb = (int32_t)(std::lower_bound(b, b + npp, t.corner[2 * (int64_t)k + 1]) - b);
                if (sa == sb)
                { // degenerate face (prev == next): not expressible as a pair -> generic (literal) path
                    r[14] |= (int32_t)0x80000000u;
                    continue;
                }
                if (sa > sb)
                    std::swap(sa, sb);
                // index of pair (sa,sb), sa<sb, in the order (0,1),(0,2),..,(0,5),(1,2),..,(4,5)
                mask |= 1 << (sa * 6 - sa * (sa + 1) / 2 + (sb - sa - 1));
            }
            r[15] = mask;
        }
    }
    tick("point records");
    return t;
}

} // namespace sm

namespace sm
{
void buildEdgeRecords(Topology &t)
{
    const int64_t E = t.E;
    if (!t.edgeRec.empty())
        return;
    t.edgeRec.assign(12 * E, 0);
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < E; ++e)
    {
        int32_t *r = &t.edgeRec[12 * e];
        const int32_t nf = t.efOff[e + 1] - t.efOff[e], nc = t.ecOff[e + 1] - t.ecOff[e];
        const bool generic = nf > 4 || nc > 4;
        r[0] = t.edge[2 * e];
        r[1] = t.edge[2 * e + 1];
        // Faces in fan order around the edge with cell k between face k and face k+1 (mod nf), so the
        // kernel needs no per-cell face indices: walk face -> cell -> other face, starting from a face
        // that has only one cell at this edge if there is one (boundary fan), else from face 0.
        int32_t fOrd[4], cOrd[4];
        bool fan = !generic;
        if (fan)
        {
            const int32_t fb = t.efOff[e], cb = t.ecOff[e];
            int32_t cellsOfFace[4] = {0, 0, 0, 0};
            for (int32_t k = 0; k < nc; ++k)
            {
                ++cellsOfFace[t.ecPair[cb + k] & 0xffff];
                ++cellsOfFace[(t.ecPair[cb + k] >> 16) & 0xffff];
            }
            int32_t start = 0;
            for (int32_t i = 0; i < nf; ++i)
                if (cellsOfFace[i] == 1)
                {
                    start = i;
                    break;
                }
            bool usedCell[4] = {false, false, false, false};
            int32_t cur = start, nfo = 0, nco = 0;
            fOrd[nfo++] = cur;
            for (;;)
            {
                int32_t found = -1, other = -1;
                for (int32_t k = 0; k < nc; ++k)
                {
                    if (usedCell[k])
                        continue;
                    const int32_t f0 = t.ecPair[cb + k] & 0xffff, f1 = (t.ecPair[cb + k] >> 16) & 0xffff;
                    if (f0 == cur || f1 == cur)
                    {
                        found = k;
                        other = (f0 == cur) ? f1 : f0;
                        break;
                    }
                }
                if (found < 0)
                    break;
                usedCell[found] = true;
                cOrd[nco++] = found;
                if (other == start)
                    break; // closed fan
                if (nfo >= nf)
                {
                    fan = false;
                    break;
                }
                fOrd[nfo++] = other;
                cur = other;
            }
            // valid fans: open (nf == nc + 1) or closed (nf == nc), everything visited exactly once
            fan = fan && nfo == nf && nco == nc && (nf == nc || nf == nc + 1);
            for (int32_t i = 0; fan && i < nf; ++i)
                for (int32_t j = i + 1; j < nf; ++j)
                    if (fOrd[i] == fOrd[j])
                        fan = false;
            if (fan)
            {
                for (int32_t k = 0; k < nf; ++k)
                    r[2 + k] = t.ef[fb + fOrd[k]];
                for (int32_t k = 0; k < nc; ++k)
                    r[6 + k] = t.ecCell[cb + cOrd[k]];
            }
        }
        const int32_t meta = nf | (nc << 4) | (fan ? 0 : (int32_t)0x80000000u);
        r[10] = meta;
    }
}
} // namespace sm

// ------------------------------------------------- boundary layer treatment ----
namespace sm
{
// One-time set-up of the prismatic boundary layer treatment (src/smoothMesh.C:2190-2221 with
// src/boundaryPointSmoothing.C:301-423 and src/orthogonalBoundaryBlending.C:52-134, 244-391),
// serial form.  Everything here depends on topology and patch selection only; the normal
// *values* are computed on the device and copied along `normalSrc`.
LayerSetup buildLayerSetup(const PolyMesh &m, const Topology &t, const std::vector<int32_t> &patchLayer, int maxLayers)
{
    LayerSetup L;
    const int64_t P = t.P;
    // classifyBoundaryPoints: a boundary point is classified once, by the first patch containing it
    std::vector<uint8_t> visited(P, 0), connected(P, 0), layerSurface(P, 0);
    for (size_t pi = 0; pi < m.patches.size(); ++pi)
    {
        const Patch &pt = m.patches[pi];
        for (int32_t f = pt.start; f < pt.start + pt.size; ++f)
            for (int32_t k = m.faceOffsets[f]; k < m.faceOffsets[f + 1]; ++k)
            {
                const int32_t p = m.faceVerts[k];
                if (visited[p])
                    continue;
                visited[p] = 1;
                if (t.isInternal[p])
                    continue;
                for (int32_t s = t.ppOff[p]; s < t.ppOff[p + 1]; ++s)
                    if (t.isInternal[t.pp[s]])
                        connected[p] = 1;
                if (patchLayer[pi])
                    layerSurface[p] = 1;
            }
    }
    // calculatePointHopsToBoundary(layerPatchIds, maxIter = maxLayers + 1)
    L.hops.assign(P, -1);
    for (size_t pi = 0; pi < m.patches.size(); ++pi)
    {
        if (!patchLayer[pi])
            continue;
        const Patch &pt = m.patches[pi];
        for (int32_t f = pt.start; f < pt.start + pt.size; ++f)
            for (int32_t k = m.faceOffsets[f]; k < m.faceOffsets[f + 1]; ++k)
                if (connected[m.faceVerts[k]])
                    L.hops[m.faceVerts[k]] = 0;
    }
    std::vector<int32_t> newHops(P, -1);
    for (int iter = 0; iter < maxLayers + 1; ++iter)
    {
        for (int64_t p = 0; p < P; ++p)
        {
            if (L.hops[p] >= 0 || !t.isInternal[p])
                continue;
            int32_t mx = -1;
            for (int32_t s = t.ppOff[p]; s < t.ppOff[p + 1]; ++s)
                mx = std::max(mx, L.hops[t.pp[s]]);
            if (mx >= 0)
                newHops[p] = mx + 1;
        }
        for (int64_t p = 0; p < P; ++p)
            if (newHops[p] > L.hops[p])
                L.hops[p] = newHops[p];
    }
    // propagateOuterNeighInfo: maps towards the boundary and the source of the propagated normal.
    // normalSrc: >= 0 boundary point whose normal is copied, -1 none (zero normal), -2 UNDEF_VECTOR
    // (multiply connected, undone at :372-382)
    L.pointToOuter.assign(P, -1);
    L.normalSrc.assign(P, -1);
    for (int64_t p = 0; p < P; ++p)
        if (!t.isInternal[p])
            L.normalSrc[p] = (int32_t)p; // boundary points carry their own (device-computed) normal
    std::vector<int32_t> boundaryPointLabels(P, -1);
    std::vector<int32_t> firstWithLabel(P, -1); // findIndex(boundaryPointLabels, q): lowest point mapped to q
    for (int iter = 1; iter < maxLayers + 2; ++iter)
        for (int64_t p = 0; p < P; ++p)
        {
            if (L.hops[p] != iter)
                continue;
            int32_t n = 0, q = -1;
            for (int32_t s = t.ppOff[p]; s < t.ppOff[p + 1]; ++s)
                if (L.hops[t.pp[s]] == iter - 1)
                {
                    ++n;
                    q = t.pp[s];
                }
            if (n != 1)
                continue;
            if (!t.isInternal[q] && !layerSurface[q])
                continue;
            if (firstWithLabel[q] >= 0)
            {
                L.normalSrc[p] = -2;
                L.normalSrc[firstWithLabel[q]] = -2;
                continue;
            }
            L.pointToOuter[p] = q;
            L.normalSrc[p] = L.normalSrc[q];
            boundaryPointLabels[p] = q;
            firstWithLabel[q] = (int32_t)p; // points are visited in ascending order: the first is the lowest
        }
    for (int64_t p = 0; p < P; ++p)
        if (L.normalSrc[p] == -2)
        {
            L.normalSrc[p] = -1;
            L.pointToOuter[p] = -1;
        }
    // point -> boundary faces on non-processor, non-empty patches (ascending face label): the
    // accumulation order of calculateBoundaryPointNormals (:151-182)
    L.bfOff.assign(P + 1, 0);
    for (const Patch &pt : m.patches)
    {
        if (pt.kind() != PATCH_BOUNDARY)
            continue;
        for (int32_t f = pt.start; f < pt.start + pt.size; ++f)
            for (int32_t k = m.faceOffsets[f]; k < m.faceOffsets[f + 1]; ++k)
                ++L.bfOff[m.faceVerts[k] + 1];
    }
    for (int64_t p = 0; p < P; ++p)
        L.bfOff[p + 1] += L.bfOff[p];
    L.bf.resize(L.bfOff[P]);
    {
        std::vector<int32_t> cur(L.bfOff.begin(), L.bfOff.end() - 1);
        for (const Patch &pt : m.patches)
        {
            if (pt.kind() != PATCH_BOUNDARY)
                continue;
            for (int32_t f = pt.start; f < pt.start + pt.size; ++f)
                for (int32_t k = m.faceOffsets[f]; k < m.faceOffsets[f + 1]; ++k)
                    L.bf[cur[m.faceVerts[k]]++] = f;
        }
    }
    for (int32_t h : L.hops)
        L.maxHop = std::max(L.maxHop, h);
    return L;
}

LayerSetup buildLayerSetupParallel(const PolyMesh &m, const Topology &t, const std::vector<int32_t> &patchLayer,
                                   int maxLayers, std::vector<double> &normals, const LayerSync &sync)
{
    // tables that need no communication (classification flags are local in the reference too,
    // src/boundaryPointSmoothing.C:301-423)
    LayerSetup L = buildLayerSetup(m, t, patchLayer, maxLayers);
    const int64_t P = t.P;
    std::vector<uint8_t> visited(P, 0), connected(P, 0), layerSurface(P, 0);
    for (size_t pi = 0; pi < m.patches.size(); ++pi)
    {
        const Patch &pt = m.patches[pi];
        for (int32_t f = pt.start; f < pt.start + pt.size; ++f)
            for (int32_t k = m.faceOffsets[f]; k < m.faceOffsets[f + 1]; ++k)
            {
                const int32_t p = m.faceVerts[k];
                if (visited[p])
                    continue;
                visited[p] = 1;
                if (t.isInternal[p])
                    continue;
                for (int32_t s = t.ppOff[p]; s < t.ppOff[p + 1]; ++s)
                    if (t.isInternal[t.pp[s]])
                        connected[p] = 1;
                if (patchLayer[pi])
                    layerSurface[p] = 1;
            }
    }
    // calculatePointHopsToBoundary with the max-synchronisation after every sweep (:124-130)
    L.hops.assign(P, -1);
    for (size_t pi = 0; pi < m.patches.size(); ++pi)
    {
        if (!patchLayer[pi])
            continue;
        const Patch &pt = m.patches[pi];
        for (int32_t f = pt.start; f < pt.start + pt.size; ++f)
            for (int32_t k = m.faceOffsets[f]; k < m.faceOffsets[f + 1]; ++k)
                if (connected[m.faceVerts[k]])
                    L.hops[m.faceVerts[k]] = 0;
    }
    std::vector<int32_t> newHops(P, -1);
    for (int iter = 0; iter < maxLayers + 1; ++iter)
    {
        for (int64_t p = 0; p < P; ++p)
        {
            if (L.hops[p] >= 0 || !t.isInternal[p])
                continue;
            int32_t mx = -1;
            for (int32_t s = t.ppOff[p]; s < t.ppOff[p + 1]; ++s)
                mx = std::max(mx, L.hops[t.pp[s]]);
            if (mx >= 0)
                newHops[p] = mx + 1;
        }
        for (int64_t p = 0; p < P; ++p)
            if (newHops[p] > L.hops[p])
                L.hops[p] = newHops[p];
        sync.maxInt(L.hops);
    }
    // set-up call of calculateBoundaryPointNormals: the two sum-synchronisations (:185-198), the
    // sharp-edge zeroing (:200-218) and the normalisation (:222-230)
    auto magOf = [&](int64_t p) {
        const double x = normals[3 * p], y = normals[3 * p + 1], z = normals[3 * p + 2];
        return std::sqrt(x * x + y * y + z * z);
    };
    auto equals = [&](int64_t p, double v) {
        return sm_equal(normals[3 * p], v) && sm_equal(normals[3 * p + 1], v) && sm_equal(normals[3 * p + 2], v);
    };
    auto setAll = [&](int64_t p, double v) { normals[3 * p] = normals[3 * p + 1] = normals[3 * p + 2] = v; };
    std::vector<int32_t> nFaces(P);
    for (int64_t p = 0; p < P; ++p)
        nFaces[p] = L.bfOff[p + 1] - L.bfOff[p];
    sync.sumVec(normals);
    sync.sumInt(nFaces);
    for (int64_t p = 0; p < P; ++p)
        if (nFaces[p] >= 1 && magOf(p) < 0.1)
            setAll(p, 0.0);
    for (int64_t p = 0; p < P; ++p)
        if (!equals(p, 0.0))
        {
            const double mg = magOf(p);
            normals[3 * p] /= mg;
            normals[3 * p + 1] /= mg;
            normals[3 * p + 2] /= mg;
        }
    // propagateOuterNeighInfo on values, maxMagSqr-synchronised after every sweep (:363-369);
    // UNDEF_VECTOR (GREAT, GREAT, GREAT) marks multiply connected points and wins every combination
    L.pointToOuter.assign(P, -1);
    L.normalSrc.assign(P, -1);
    std::vector<int32_t> firstWithLabel(P, -1);
    for (int iter = 1; iter < maxLayers + 2; ++iter)
    {
        for (int64_t p = 0; p < P; ++p)
        {
            if (L.hops[p] != iter)
                continue;
            int32_t n = 0, q = -1;
            for (int32_t s = t.ppOff[p]; s < t.ppOff[p + 1]; ++s)
                if (L.hops[t.pp[s]] == iter - 1)
                {
                    ++n;
                    q = t.pp[s];
                }
            if (n != 1)
                continue;
            if (!t.isInternal[q] && !layerSurface[q])
                continue;
            if (firstWithLabel[q] >= 0)
            {
                setAll(p, SM_GREAT);
                setAll(firstWithLabel[q], SM_GREAT);
                continue;
            }
            L.pointToOuter[p] = q;
            for (int k = 0; k < 3; ++k)
                normals[3 * p + k] = normals[3 * q + k];
            firstWithLabel[q] = (int32_t)p;
        }
        sync.maxMagSqrVec(normals);
    }
    for (int64_t p = 0; p < P; ++p)
        if (equals(p, SM_GREAT))
        {
            setAll(p, 0.0);
            L.pointToOuter[p] = -1;
        }
    L.maxHop = 0;
    for (int32_t h : L.hops)
        L.maxHop = std::max(L.maxHop, h);
    return L;
}

namespace
{
inline uint64_t spreadBits21(uint64_t v)
{ // 21 bits -> every third bit
    v &= 0x1fffff;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}
} // namespace

GeomTiles buildGeomTiles(const PolyMesh &m, const Topology &t, int maxCells, int maxFaces, int maxPoints, bool keepPairs)
{
    GeomTiles G;
    const int64_t C = t.C, F = t.F, P = t.P;
    if (C == 0)
        return G;
    const bool timing = getenv("SMGPU_TIMING") && atoi(getenv("SMGPU_TIMING")) != 0;
    auto clock = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double tPrev = clock();
    auto tick = [&](const char *what) {
        if (!timing)
            return;
        const double tNow = clock();
        fprintf(stderr, "[smgpu set-up] %-28s %.3f s\n", what, tNow - tPrev);
        tPrev = tNow;
    };
    // cell order: Morton curve over the cells' vertex averages, quantised by the mean cell size so that
    // on block-structured meshes 2^k consecutive cells form a brick
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int64_t p = 0; p < P; ++p)
        for (int d = 0; d < 3; ++d)
        {
            lo[d] = std::min(lo[d], m.points[3 * p + d]);
            hi[d] = std::max(hi[d], m.points[3 * p + d]);
        }
    double vol = 1.0;
    int dims = 0;
    for (int d = 0; d < 3; ++d)
        if (hi[d] - lo[d] > 0)
        {
            vol *= hi[d] - lo[d];
            ++dims;
        }
    const double h = dims ? std::pow(vol / double(C), 1.0 / dims) : 1.0;
    const double invH = h > 0 ? 1.0 / h : 0.0;
    std::vector<std::pair<uint64_t, int32_t>> keys(C);
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < C; ++c)
    {
        double x[3] = {0, 0, 0};
        int n = 0;
        for (int32_t k = t.cfOff[c]; k < t.cfOff[c + 1]; ++k)
        {
            const int32_t f = t.cf[k] & 0x7fffffff;
            for (int32_t q = m.faceOffsets[f]; q < m.faceOffsets[f + 1]; ++q)
            {
                for (int d = 0; d < 3; ++d)
                    x[d] += m.points[3 * (int64_t)m.faceVerts[q] + d];
                ++n;
            }
        }
        uint64_t key = 0;
        for (int d = 0; d < 3; ++d)
        {
            double q = n ? (x[d] / n - lo[d]) * invH : 0.0;
            q = q < 0 ? 0 : (q > 2097151.0 ? 2097151.0 : q);
            key |= spreadBits21((uint64_t)q) << d;
        }
        keys[c] = {key, (int32_t)c};
    }
    tick("tiles: cell keys");
    __gnu_parallel::sort(keys.begin(), keys.end());
    tick("tiles: sort");
    if (maxFaces > 0x7fff || maxPoints > 0xffff)
        return GeomTiles();

    // Tiles = runs of consecutive cells of that order.  Start from runs of maxCells cells and halve a run
    // until its face and point lists fit the budgets (deterministic, and independent per initial run, so
    // the construction is parallel; halving only happens on polyhedral or very irregular meshes).
    struct Tile
    {
        int32_t a, b; // cells keys[a .. b)
        std::vector<int32_t> faces, points;
    };
    auto collect = [&](Tile &T) {
        T.faces.clear();
        T.points.clear();
        for (int32_t i = T.a; i < T.b; ++i)
        {
            const int32_t c = keys[i].second;
            for (int32_t k = t.cfOff[c]; k < t.cfOff[c + 1]; ++k)
                T.faces.push_back(t.cf[k] & 0x7fffffff);
        }
        std::sort(T.faces.begin(), T.faces.end());
        T.faces.erase(std::unique(T.faces.begin(), T.faces.end()), T.faces.end());
        for (int32_t f : T.faces)
            for (int32_t q = m.faceOffsets[f]; q < m.faceOffsets[f + 1]; ++q)
                T.points.push_back(m.faceVerts[q]);
        std::sort(T.points.begin(), T.points.end());
        T.points.erase(std::unique(T.points.begin(), T.points.end()), T.points.end());
    };
    // runs are closed at maxCells cells and at every aligned block of 2^8 keys (an 8 x 8 x 4 brick of a
    // block-structured mesh), so that a brick cut short by the mesh boundary does not shift all later tiles
    // off the brick grid (271^3: 33 full bricks and one of 7 cells per row)
    std::vector<std::pair<int32_t, int32_t>> runs;
    for (int64_t i = 0; i < C;)
    {
        int64_t j = i;
        const uint64_t block = keys[i].first >> 8;
        while (j < C && j - i < maxCells && (maxCells < 256 || (keys[j].first >> 8) == block))
            ++j;
        runs.push_back({(int32_t)i, (int32_t)j});
        i = j;
    }
    const int64_t nRuns = (int64_t)runs.size();
    std::vector<std::vector<Tile>> perRun(nRuns);
    bool cellTooLarge = false;
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t r = 0; r < nRuns; ++r)
    {
        std::vector<Tile> todo, done;
        todo.push_back({runs[r].first, runs[r].second, {}, {}});
        while (!todo.empty())
        {
            Tile T = std::move(todo.back());
            todo.pop_back();
            collect(T);
            if ((int)T.faces.size() <= maxFaces && (int)T.points.size() <= maxPoints)
                done.push_back(std::move(T));
            else if (T.b - T.a == 1)
            {
#pragma omp atomic write
                cellTooLarge = true;
            }
            else
            { // second half first on the stack, so the first half is finished first (keeps the order)
                const int32_t mid = T.a + (T.b - T.a) / 2;
                todo.push_back({mid, T.b, {}, {}});
                todo.push_back({T.a, mid, {}, {}});
            }
        }
        perRun[r] = std::move(done);
    }
    tick("tiles: face / point lists");
    if (cellTooLarge)
        return GeomTiles(); // a cell that does not fit a tile: the caller keeps the two-kernel path
    std::vector<Tile *> tiles;
    for (auto &v : perRun)
        for (Tile &T : v)
            tiles.push_back(&T);
    G.nTiles = (int32_t)tiles.size();
    // offsets
    G.tileCellOff.assign(G.nTiles + 1, 0);
    G.tileFaceOff.assign(G.nTiles + 1, 0);
    G.tilePointOff.assign(G.nTiles + 1, 0);
    std::vector<int64_t> slotRefOff(G.nTiles + 1, 0), faceRefBase(G.nTiles + 1, 0);
    for (int32_t k = 0; k < G.nTiles; ++k)
    {
        const Tile &T = *tiles[k];
        int64_t nSlotRefs = 0, nFaceRefs = 0;
        for (int32_t i = T.a; i < T.b; ++i)
            nSlotRefs += t.cfOff[keys[i].second + 1] - t.cfOff[keys[i].second];
        for (int32_t f : T.faces)
            nFaceRefs += m.faceOffsets[f + 1] - m.faceOffsets[f];
        G.tileCellOff[k + 1] = G.tileCellOff[k] + (T.b - T.a);
        G.tileFaceOff[k + 1] = G.tileFaceOff[k] + (int32_t)T.faces.size();
        G.tilePointOff[k + 1] = G.tilePointOff[k] + (int32_t)T.points.size();
        slotRefOff[k + 1] = slotRefOff[k] + nSlotRefs;
        faceRefBase[k + 1] = faceRefBase[k] + nFaceRefs;
    }
    if (faceRefBase[G.nTiles] >= (int64_t)INT32_MAX)
        return GeomTiles();
    // the first tile (in tile order) that lists a face stores the face's global outputs
    std::vector<int32_t> firstTile(F, -1);
    for (int32_t k = 0; k < G.nTiles; ++k)
        for (int32_t f : tiles[k]->faces)
            if (firstTile[f] < 0)
                firstTile[f] = k;
    tick("tiles: offsets, first tile");
    G.tileCells.resize(C);
    G.tileFaces.resize(G.tileFaceOff[G.nTiles]);
    G.tilePoints.resize(G.tilePointOff[G.nTiles]);
    G.slotOff.resize(C + 1);
    G.slotRef.resize(slotRefOff[G.nTiles]);
    G.faceRefOff.resize(G.tileFaceOff[G.nTiles] + 1);
    G.faceRef.resize(faceRefBase[G.nTiles]);
    G.slotOff[C] = (int32_t)slotRefOff[G.nTiles];
    G.faceRefOff[G.tileFaceOff[G.nTiles]] = (int32_t)faceRefBase[G.nTiles];
#pragma omp parallel for schedule(dynamic, 16)
    for (int32_t k = 0; k < G.nTiles; ++k)
    {
        const Tile &T = *tiles[k];
        std::vector<int32_t> cells;
        for (int32_t i = T.a; i < T.b; ++i)
            cells.push_back(keys[i].second);
        std::sort(cells.begin(), cells.end());
        auto localPoint = [&](int32_t p) { return (int32_t)(std::lower_bound(T.points.begin(), T.points.end(), p) - T.points.begin()); };
        auto localFace = [&](int32_t f) { return (int32_t)(std::lower_bound(T.faces.begin(), T.faces.end(), f) - T.faces.begin()); };
        std::copy(T.points.begin(), T.points.end(), G.tilePoints.begin() + G.tilePointOff[k]);
        int64_t fr = faceRefBase[k];
        for (size_t i = 0; i < T.faces.size(); ++i)
        {
            const int32_t f = T.faces[i];
            G.tileFaces[G.tileFaceOff[k] + i] = (firstTile[f] == k) ? (int32_t)(f | 0x80000000u) : f;
            G.faceRefOff[G.tileFaceOff[k] + i] = (int32_t)fr;
            for (int32_t q = m.faceOffsets[f]; q < m.faceOffsets[f + 1]; ++q)
                G.faceRef[fr++] = (uint16_t)localPoint(m.faceVerts[q]);
        }
        int64_t sr = slotRefOff[k];
        for (size_t i = 0; i < cells.size(); ++i)
        {
            const int32_t c = cells[i], slot = G.tileCellOff[k] + (int32_t)i;
            G.tileCells[slot] = c;
            G.slotOff[slot] = (int32_t)sr;
            for (int32_t q = t.cfOff[c]; q < t.cfOff[c + 1]; ++q)
            {
                const int32_t w = t.cf[q];
                G.slotRef[sr++] = (uint16_t)(localFace(w & 0x7fffffff) | (w < 0 ? 0x8000 : 0));
            }
        }
    }
    tick("tiles: references");
    // ---- (edge, cell) pairs of the fused face-angle filter: the edges of a cell are the vertex pairs that two
    //      of its faces share; both faces and both end points are already tile-local references ----
    {
        std::vector<int32_t> nPairs(C + 1, 0);
        bool closed = true;
        struct Half
        {
            uint16_t a, b, f;
        };
        auto cellPairs = [&](int32_t k, int32_t slot, std::vector<Half> &hs, uint16_t *out) -> int32_t {
            // half edges of the cell's faces, sorted by end points; a closed cell has every edge exactly twice
            hs.clear();
            const int32_t fb = G.tileFaceOff[k];
            for (int32_t q = G.slotOff[slot]; q < G.slotOff[slot + 1]; ++q)
            {
                const uint16_t li = G.slotRef[q] & 0x7fff;
                const int32_t rb = G.faceRefOff[fb + li], nv = G.faceRefOff[fb + li + 1] - rb;
                for (int32_t v = 0; v < nv; ++v)
                {
                    const uint16_t a = G.faceRef[rb + v], b = G.faceRef[rb + (v + 1 == nv ? 0 : v + 1)];
                    hs.push_back({std::min(a, b), std::max(a, b), li});
                }
            }
            std::sort(hs.begin(), hs.end(), [](const Half &x, const Half &y) {
                return x.a != y.a ? x.a < y.a : (x.b != y.b ? x.b < y.b : x.f < y.f);
            });
            if (hs.size() % 2)
                return -1;
            int32_t n = 0;
            for (size_t i = 0; i < hs.size(); i += 2)
            {
                if (hs[i].a != hs[i + 1].a || hs[i].b != hs[i + 1].b || hs[i].a == hs[i].b ||
                    (i + 2 < hs.size() && hs[i + 2].a == hs[i].a && hs[i + 2].b == hs[i].b))
                    return -1;
                if (out)
                {
                    out[4 * n] = hs[i].a, out[4 * n + 1] = hs[i].b;
                    out[4 * n + 2] = hs[i].f, out[4 * n + 3] = hs[i + 1].f;
                }
                ++n;
            }
            return n;
        };
        // canonical hexahedron record of a cell (see topology.hpp), checked against its pair list
        auto hexRecord = [&](int32_t k, int32_t slot, const uint16_t *pairs, uint16_t rec[16]) -> bool {
            const int32_t fb = G.tileFaceOff[k];
            if (G.slotOff[slot + 1] - G.slotOff[slot] != 6)
                return false;
            uint16_t fl[6], fv[6][4];
            for (int q = 0; q < 6; ++q)
            {
                fl[q] = G.slotRef[G.slotOff[slot] + q] & 0x7fff;
                const int32_t rb = G.faceRefOff[fb + fl[q]];
                if (G.faceRefOff[fb + fl[q] + 1] - rb != 4)
                    return false;
                for (int v = 0; v < 4; ++v)
                    fv[q][v] = G.faceRef[rb + v];
            }
            auto has = [&](int q, uint16_t p) { return fv[q][0] == p || fv[q][1] == p || fv[q][2] == p || fv[q][3] == p; };
            for (int q = 0; q < 16; ++q)
                rec[q] = 0;
            int B = -1;
            for (int q = 1; q < 6; ++q)
                if (!has(q, fv[0][0]) && !has(q, fv[0][1]) && !has(q, fv[0][2]) && !has(q, fv[0][3]))
                    B = (B < 0) ? q : 99;
            if (B < 1 || B > 5)
                return false;
            for (int e = 0; e < 4; ++e)
            {
                const uint16_t a = fv[0][e], b = fv[0][(e + 1) & 3];
                int S = -1;
                for (int q = 1; q < 6; ++q)
                    if (q != B && has(q, a) && has(q, b))
                        S = (S < 0) ? q : 99;
                if (S < 1 || S > 5)
                    return false;
                int at = 0;
                while (fv[S][at] != a)
                    ++at;
                // w_e: the neighbour of a in S's loop that is not b
                const uint16_t n1 = fv[S][(at + 1) & 3], n2 = fv[S][(at + 3) & 3];
                const uint16_t w = (n1 == b) ? n2 : n1;
                if (!((n1 == b || n2 == b) && has(B, w)))
                    return false;
                rec[e] = a;
                rec[4 + e] = w;
                rec[10 + e] = fl[S];
            }
            rec[8] = fl[0];
            rec[9] = fl[B];
            // the fixed pattern must reproduce the cell's pair list
            std::array<std::array<uint16_t, 4>, 12> want, got;
            for (int j = 0; j < 12; ++j)
                want[j] = {pairs[4 * j], pairs[4 * j + 1], std::min(pairs[4 * j + 2], pairs[4 * j + 3]), std::max(pairs[4 * j + 2], pairs[4 * j + 3])};
            int ng = 0;
            auto add = [&](uint16_t p0, uint16_t p1, uint16_t f0, uint16_t f1) {
                got[ng++] = {std::min(p0, p1), std::max(p0, p1), std::min(f0, f1), std::max(f0, f1)};
            };
            for (int e = 0; e < 4; ++e)
            {
                add(rec[e], rec[(e + 1) & 3], rec[8], rec[10 + e]);
                add(rec[4 + e], rec[4 + ((e + 1) & 3)], rec[9], rec[10 + e]);
                add(rec[e], rec[4 + e], rec[10 + ((e + 3) & 3)], rec[10 + e]);
            }
            std::sort(want.begin(), want.end());
            std::sort(got.begin(), got.end());
            return want == got;
        };
        for (int32_t k = 0; k < G.nTiles; ++k)
        {
            G.maxTileCells = std::max(G.maxTileCells, G.tileCellOff[k + 1] - G.tileCellOff[k]);
            G.maxTileFaces = std::max(G.maxTileFaces, G.tileFaceOff[k + 1] - G.tileFaceOff[k]);
            G.maxTilePoints = std::max(G.maxTilePoints, G.tilePointOff[k + 1] - G.tilePointOff[k]);
        }
        // one pass: pair counts, closedness, and the canonical record of every cell that is a topological
        // hexahedron; then the uniform tiles (all faces quadrilaterals, all cells such hexahedra) get their
        // fixed-stride reference copies, and the pair lists are materialised for the cells of the other tiles
        // (for all cells when the caller wants them for checking)
        std::vector<uint8_t> cellIsHex(C, 0);
        std::vector<uint16_t> recAll(16 * (size_t)C, 0); // cell-major, by slot
#pragma omp parallel
        {
            std::vector<Half> hs;
            std::vector<uint16_t> tmp;
#pragma omp for schedule(dynamic, 16)
            for (int32_t k = 0; k < G.nTiles; ++k)
            {
                const int32_t cb = G.tileCellOff[k], nc = G.tileCellOff[k + 1] - cb;
                for (int32_t i = 0; i < nc; ++i)
                {
                    const int32_t slot = cb + i;
                    int32_t nHalf = 0;
                    for (int32_t q = G.slotOff[slot]; q < G.slotOff[slot + 1]; ++q)
                    {
                        const int32_t at = G.tileFaceOff[k] + (G.slotRef[q] & 0x7fff);
                        nHalf += G.faceRefOff[at + 1] - G.faceRefOff[at];
                    }
                    tmp.resize(2 * (size_t)nHalf + 4);
                    const int32_t n = cellPairs(k, slot, hs, tmp.data());
                    if (n < 0)
                    {
#pragma omp atomic write
                        closed = false;
                    }
                    nPairs[slot + 1] = std::max(n, 0);
                    if (n == 12 && hexRecord(k, slot, tmp.data(), &recAll[16 * (size_t)slot]))
                        cellIsHex[slot] = 1;
                }
            }
        }
        G.tileUFaceOff.assign(G.nTiles, -1);
        G.tileUCellOff.assign(G.nTiles, -1);
        if (closed)
        {
            int64_t uf = 0, uc = 0;
            for (int32_t k = 0; k < G.nTiles; ++k)
            {
                bool uni = true;
                for (int32_t slot = G.tileCellOff[k]; slot < G.tileCellOff[k + 1] && uni; ++slot)
                    uni = cellIsHex[slot] != 0;
                for (int32_t i = G.tileFaceOff[k]; i < G.tileFaceOff[k + 1] && uni; ++i)
                    uni = G.faceRefOff[i + 1] - G.faceRefOff[i] == 4;
                if (!uni)
                    continue;
                G.tileUFaceOff[k] = (int32_t)uf;
                G.tileUCellOff[k] = (int32_t)uc;
                uf += G.tileFaceOff[k + 1] - G.tileFaceOff[k];
                uc += G.tileCellOff[k + 1] - G.tileCellOff[k];
            }
            G.nUniformCells = uc;
            G.uFaceRef.resize(4 * (size_t)uf);
            G.uSlotRef.resize(6 * (size_t)uc);
            G.hexRec.resize(16 * (size_t)uc);
#pragma omp parallel for schedule(dynamic, 16)
            for (int32_t k = 0; k < G.nTiles; ++k)
            {
                if (G.tileUCellOff[k] < 0)
                    continue;
                const int32_t fb = G.tileFaceOff[k], nf = G.tileFaceOff[k + 1] - fb, cb = G.tileCellOff[k], nc = G.tileCellOff[k + 1] - cb;
                const size_t ufb = (size_t)G.tileUFaceOff[k], ucb = (size_t)G.tileUCellOff[k];
                for (int32_t i = 0; i < nf; ++i)
                    for (int q = 0; q < 4; ++q)
                        G.uFaceRef[4 * (ufb + i) + q] = G.faceRef[G.faceRefOff[fb + i] + q];
                for (int32_t i = 0; i < nc; ++i)
                {
                    for (int q = 0; q < 6; ++q)
                        G.uSlotRef[6 * (ucb + i) + q] = G.slotRef[G.slotOff[cb + i] + q];
                    const uint16_t *rec = &recAll[16 * (size_t)(cb + i)];
                    uint16_t *o1 = &G.hexRec[16 * ucb + 8 * (size_t)i], *o2 = &G.hexRec[16 * ucb + 8 * (size_t)nc + 8 * (size_t)i];
                    for (int q = 0; q < 8; ++q)
                        o1[q] = rec[q], o2[q] = rec[8 + q];
                }
            }
            G.uniformCellEdges = C ? nPairs[1] : 0;
            for (int64_t c = 0; c < C && G.uniformCellEdges; ++c)
                if (nPairs[c + 1] != G.uniformCellEdges)
                    G.uniformCellEdges = 0;
            // pair lists: for the cells of the non-uniform tiles (all cells if asked for)
            int64_t total = 0;
            std::vector<uint8_t> wantPairs(C, 0);
            for (int32_t k = 0; k < G.nTiles; ++k)
                if (keepPairs || G.tileUCellOff[k] < 0)
                    for (int32_t slot = G.tileCellOff[k]; slot < G.tileCellOff[k + 1]; ++slot)
                    {
                        wantPairs[slot] = 1;
                        total += nPairs[slot + 1];
                    }
            if (total > 0 && total < (int64_t)INT32_MAX / 2)
            {
                G.cellEdgeOff.assign(C + 1, 0);
                for (int64_t c = 0; c < C; ++c)
                    G.cellEdgeOff[c + 1] = G.cellEdgeOff[c] + (wantPairs[c] ? nPairs[c + 1] : 0);
                G.cellEdgeRef.resize(4 * (size_t)total);
                for (int32_t k = 0; k < G.nTiles; ++k)
                    G.maxTileEdgePairs = std::max(G.maxTileEdgePairs, G.cellEdgeOff[G.tileCellOff[k + 1]] - G.cellEdgeOff[G.tileCellOff[k]]);
#pragma omp parallel
                {
                    std::vector<Half> hs;
#pragma omp for schedule(dynamic, 16)
                    for (int32_t k = 0; k < G.nTiles; ++k)
                        for (int32_t slot = G.tileCellOff[k]; slot < G.tileCellOff[k + 1]; ++slot)
                            if (wantPairs[slot])
                                cellPairs(k, slot, hs, G.cellEdgeRef.data() + 4 * (size_t)G.cellEdgeOff[slot]);
                }
            }
        }
        tick("tiles: (edge, cell) pairs");
    }
    return G;
}

PointTiles buildPointTiles(const PolyMesh &m, const Topology &t, int maxOwn, int maxHalo, int maxCells)
{
    PointTiles T;
    const int64_t P = t.P;
    if (P == 0 || maxHalo > 0xffff || maxCells > 0xffff)
        return T;
    const bool timing = getenv("SMGPU_TIMING") && atoi(getenv("SMGPU_TIMING")) != 0;
    auto clock = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double tPrev = clock();
    auto tick = [&](const char *what) {
        if (!timing)
            return;
        const double tNow = clock();
        fprintf(stderr, "[smgpu set-up] %-28s %.3f s\n", what, tNow - tPrev);
        tPrev = tNow;
    };
    // Morton order over the points, quantised (round to nearest) by the mean point spacing: on block-structured
    // meshes 2^8 consecutive keys are an 8 x 8 x 4 brick of lattice positions, jitter below half a spacing
    // does not move a point to another position
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int64_t p = 0; p < P; ++p)
        for (int d = 0; d < 3; ++d)
        {
            lo[d] = std::min(lo[d], m.points[3 * p + d]);
            hi[d] = std::max(hi[d], m.points[3 * p + d]);
        }
    double vol = 1.0;
    int dims = 0;
    for (int d = 0; d < 3; ++d)
        if (hi[d] - lo[d] > 0)
        {
            vol *= hi[d] - lo[d];
            ++dims;
        }
    const double h = dims ? std::pow(vol / double(std::max<int64_t>(t.C, 1)), 1.0 / dims) : 1.0;
    const double invH = h > 0 ? 1.0 / h : 0.0;
    std::vector<std::pair<uint64_t, int32_t>> keys(P);
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < P; ++p)
    {
        uint64_t key = 0;
        for (int d = 0; d < 3; ++d)
        {
            double q = (m.points[3 * p + d] - lo[d]) * invH + 0.5;
            q = q < 0 ? 0 : (q > 2097151.0 ? 2097151.0 : q);
            key |= spreadBits21((uint64_t)q) << d;
        }
        keys[p] = {key, (int32_t)p};
    }
    __gnu_parallel::sort(keys.begin(), keys.end());
    tick("point tiles: keys, sort");
    // runs: closed at every aligned 2^8 block of keys and at maxOwn points
    std::vector<std::pair<int32_t, int32_t>> runs;
    for (int64_t i = 0; i < P;)
    {
        int64_t j = i;
        const uint64_t block = keys[i].first >> 8;
        while (j < P && j - i < maxOwn && (keys[j].first >> 8) == block)
            ++j;
        runs.push_back({(int32_t)i, (int32_t)j});
        i = j;
    }
    struct Tile
    {
        int32_t a, b;
        std::vector<int32_t> own, halo, cells;
    };
    auto collect = [&](Tile &X) {
        X.own.clear();
        X.halo.clear();
        X.cells.clear();
        for (int32_t i = X.a; i < X.b; ++i)
            X.own.push_back(keys[i].second);
        std::sort(X.own.begin(), X.own.end());
        for (int32_t p : X.own)
        {
            for (int32_t k = t.ppOff[p]; k < t.ppOff[p + 1]; ++k)
                if (!std::binary_search(X.own.begin(), X.own.end(), t.pp[k]))
                    X.halo.push_back(t.pp[k]);
            for (int32_t k = t.pcOff[p]; k < t.pcOff[p + 1]; ++k)
                X.cells.push_back(t.pc[k]);
        }
        std::sort(X.halo.begin(), X.halo.end());
        X.halo.erase(std::unique(X.halo.begin(), X.halo.end()), X.halo.end());
        std::sort(X.cells.begin(), X.cells.end());
        X.cells.erase(std::unique(X.cells.begin(), X.cells.end()), X.cells.end());
    };
    std::vector<std::vector<Tile>> perRun(runs.size());
    bool tooLarge = false;
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t r = 0; r < (int64_t)runs.size(); ++r)
    {
        std::vector<Tile> todo, done;
        todo.push_back({runs[r].first, runs[r].second, {}, {}, {}});
        while (!todo.empty())
        {
            Tile X = std::move(todo.back());
            todo.pop_back();
            collect(X);
            if ((int)(X.own.size() + X.halo.size()) <= maxHalo && (int)X.cells.size() <= maxCells)
                done.push_back(std::move(X));
            else if (X.b - X.a == 1)
            {
#pragma omp atomic write
                tooLarge = true;
            }
            else
            {
                const int32_t mid = X.a + (X.b - X.a) / 2;
                todo.push_back({mid, X.b, {}, {}, {}});
                todo.push_back({X.a, mid, {}, {}, {}});
            }
        }
        perRun[r] = std::move(done);
    }
    tick("point tiles: lists");
    if (tooLarge)
        return PointTiles(); // a point whose stencil does not fit a tile: the caller keeps the per-point kernels
    std::vector<Tile *> tiles;
    for (auto &v : perRun)
        for (Tile &X : v)
            tiles.push_back(&X);
    T.nTiles = (int32_t)tiles.size();
    T.ownOff.assign(T.nTiles + 1, 0);
    T.haloOff.assign(T.nTiles + 1, 0);
    T.cellOff.assign(T.nTiles + 1, 0);
    for (int32_t k = 0; k < T.nTiles; ++k)
    {
        const Tile &X = *tiles[k];
        T.ownOff[k + 1] = T.ownOff[k] + (int32_t)X.own.size();
        T.haloOff[k + 1] = T.haloOff[k] + (int32_t)(X.own.size() + X.halo.size());
        T.cellOff[k + 1] = T.cellOff[k] + (int32_t)X.cells.size();
        T.maxOwn = std::max(T.maxOwn, (int32_t)X.own.size());
        T.maxHalo = std::max(T.maxHalo, (int32_t)(X.own.size() + X.halo.size()));
        T.maxCells = std::max(T.maxCells, (int32_t)X.cells.size());
    }
    T.halo.resize(T.haloOff[T.nTiles]);
    T.cell.resize(T.cellOff[T.nTiles]);
    T.rec.assign(16 * (size_t)P, 0);
#pragma omp parallel for schedule(dynamic, 16)
    for (int32_t k = 0; k < T.nTiles; ++k)
    {
        const Tile &X = *tiles[k];
        const int32_t no = (int32_t)X.own.size();
        std::copy(X.own.begin(), X.own.end(), T.halo.begin() + T.haloOff[k]);
        std::copy(X.halo.begin(), X.halo.end(), T.halo.begin() + T.haloOff[k] + no);
        std::copy(X.cells.begin(), X.cells.end(), T.cell.begin() + T.cellOff[k]);
        auto pointRef = [&](int32_t q) -> uint16_t {
            auto it = std::lower_bound(X.own.begin(), X.own.end(), q);
            if (it != X.own.end() && *it == q)
                return (uint16_t)(it - X.own.begin());
            return (uint16_t)(no + (std::lower_bound(X.halo.begin(), X.halo.end(), q) - X.halo.begin()));
        };
        for (int32_t i = 0; i < no; ++i)
        {
            const int32_t p = X.own[i];
            uint16_t *r = &T.rec[16 * (size_t)(T.ownOff[k] + i)];
            const int32_t *src = &t.pointRec[16 * (size_t)p];
            const int32_t meta = src[14];
            const int32_t npc = meta & 0xff, npp = (meta >> 8) & 0xff;
            if (meta < 0)
            {
                r[14] = 0x8000;
                continue;
            }
            for (int32_t j = 0; j < npc; ++j)
                r[j] = (uint16_t)(std::lower_bound(X.cells.begin(), X.cells.end(), src[j]) - X.cells.begin());
            for (int32_t j = 0; j < npp; ++j)
                r[8 + j] = pointRef(src[8 + j]);
            r[14] = (uint16_t)(npc | (npp << 4));
            r[15] = (uint16_t)src[15];
        }
    }
    tick("point tiles: records");
    return T;
}
} // namespace sm
