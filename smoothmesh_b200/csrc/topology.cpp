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
#include <omp.h>
#include <stdexcept>
#include <string>

namespace sm
{

static void fail(const std::string &s) { throw std::runtime_error(s); }

namespace
{
// Rows of variable length are computed once: every thread takes a contiguous range of rows (rangeOf), appends them
// to a private buffer (plain stores: vector::insert per row costs several times the row) and leaves the row lengths
// in the offset array; after the prefix sums (parScan) the buffers are copied to their places.
struct RowBuf
{
    Vec<int32_t> v;
    int64_t n = 0;
    explicit RowBuf(int64_t cap = 0) { v.resize(cap); }
    int32_t *need(int64_t k)
    {
        if (n + k > (int64_t)v.size())
            v.resize(std::max<int64_t>(2 * (int64_t)v.size(), n + k + 4096));
        return v.data() + n;
    }
};
inline void rangeOf(int64_t n, int tid, int nThreads, int64_t &lo, int64_t &hi)
{
    lo = n * tid / nThreads;
    hi = n * (tid + 1) / nThreads;
}
// off[0] = 0 and off[i + 1] = length of row i on entry, the running sums on exit; returns the longest row.  The
// rows are cut into the same nParts ranges as where they were filled (rangeOf), so the lengths are mostly read by
// the thread that wrote them.  The ranges are work items of an `omp for`, not thread numbers: the result does not
// depend on how many threads the runtime actually grants.
int32_t parScan(Vec<int32_t> &off, int64_t n, int nParts)
{
    std::vector<int64_t> sum(nParts + 1, 0);
    std::vector<int32_t> longest(nParts, 0);
    off[0] = 0;
#pragma omp parallel for schedule(static, 1)
    for (int part = 0; part < nParts; ++part)
    {
        int64_t lo, hi;
        rangeOf(n, part, nParts, lo, hi);
        int64_t s = 0;
        int32_t mx = 0;
        for (int64_t i = lo; i < hi; ++i)
        {
            s += off[i + 1];
            mx = std::max(mx, off[i + 1]);
        }
        sum[part + 1] = s;
        longest[part] = mx;
    }
    for (int k = 0; k < nParts; ++k)
        sum[k + 1] += sum[k];
    if (sum[nParts] >= (int64_t)INT32_MAX)
        fail("mesh too large for 32-bit offsets");
#pragma omp parallel for schedule(static, 1)
    for (int part = 0; part < nParts; ++part)
    {
        int64_t lo, hi;
        rangeOf(n, part, nParts, lo, hi);
        int64_t run = sum[part];
        for (int64_t i = lo; i < hi; ++i)
        {
            run += off[i + 1];
            off[i + 1] = (int32_t)run;
        }
    }
    return *std::max_element(longest.begin(), longest.end());
}
} // namespace

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
    parCopy(t.faceOff, m.faceOffsets.data(), m.faceOffsets.size());
    parCopy(t.faceVerts, m.faceVerts.data(), m.faceVerts.size());
    const int64_t FV = (int64_t)m.faceVerts.size();
    if (2 * FV >= (int64_t)INT32_MAX)
        fail("mesh too large for 32-bit offsets");

    // src/smoothMesh.C:52-80: internal = not on any non-processor patch; empty patches abort
    parFill(t.isInternal, (size_t)P, (uint8_t)1);
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
        for (const Patch &p : m.patches)
            if (p.kind() == PATCH_PROCESSOR)
                for (int32_t f = p.start; f < p.start + p.size; ++f)
                    for (int32_t k = m.faceOffsets[f]; k < m.faceOffsets[f + 1]; ++k)
                        t.procPoints.push_back(m.faceVerts[k]);
        std::sort(t.procPoints.begin(), t.procPoints.end());
        t.procPoints.erase(std::unique(t.procPoints.begin(), t.procPoints.end()), t.procPoints.end());
    }

    // ---- point -> face corners (pointFaces ascending) ----
    // counted and filled in parallel (atomic cursors), then every row is put in ascending face order
    parFill(t.cornerOff, (size_t)P + 1, (int32_t)0);
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < FV; ++k)
    {
#pragma omp atomic update
        ++t.cornerOff[m.faceVerts[k] + 1];
    }
    const int nThreads = omp_get_max_threads();
    tick("  corner count");
    const int32_t maxRow = parScan(t.cornerOff, P, nThreads);
    tick("  corner scan");
    t.corner.resize(2 * FV);
    Vec<int32_t> cornerFace(FV);
    {
        Vec<int32_t> cur;
        parCopy(cur, t.cornerOff.data(), (size_t)P);
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
        tick("  corner fill");
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
    // sorted insertion without duplicates into a short row
    auto insertUnique = [](int32_t *row, int32_t &n, int32_t v) {
        int32_t i = n;
        while (i > 0 && row[i - 1] > v)
            --i;
        if (i > 0 && row[i - 1] == v)
            return;
        for (int32_t j = n; j > i; --j)
            row[j] = row[j - 1];
        row[i] = v;
        ++n;
    };
    // ---- point -> cells (ascending) and point -> points (ascending) ----
    t.pcOff.resize(P + 1);
    t.ppOff.resize(P + 1);
    Vec<int32_t> upStart(P + 1); // first edge label of point p = #edges (a,b) with a<p
    {
        std::vector<RowBuf> pcPart(nThreads), ppPart(nThreads);
#pragma omp parallel for schedule(static, 1)
        for (int tid = 0; tid < nThreads; ++tid)
        { // tid numbers a range of rows (a work item), not a thread
            int64_t lo, hi;
            rangeOf(P, tid, nThreads, lo, hi);
            const int64_t rows = t.cornerOff[hi] - t.cornerOff[lo];
            RowBuf pcl(rows + rows / 8 + 64), ppl(rows + 64);
            for (int64_t p = lo; p < hi; ++p)
            {
                const int32_t nCorners = t.cornerOff[p + 1] - t.cornerOff[p];
                int32_t *rowC = pcl.need(2 * (int64_t)nCorners), *rowP = ppl.need(2 * (int64_t)nCorners);
                int32_t nc = 0, np = 0;
                for (int32_t s = t.cornerOff[p]; s < t.cornerOff[p + 1]; ++s)
                {
                    const int32_t f = cornerFace[s];
                    insertUnique(rowC, nc, m.owner[f]);
                    if (f < Fi)
                        insertUnique(rowC, nc, m.neighbour[f]);
                    insertUnique(rowP, np, t.corner[2 * (int64_t)s]);
                    insertUnique(rowP, np, t.corner[2 * (int64_t)s + 1]);
                }
                pcl.n += nc;
                ppl.n += np;
                int32_t up = 0;
                for (int32_t i = 0; i < np; ++i)
                    up += (rowP[i] > (int32_t)p);
                t.pcOff[p + 1] = nc;
                t.ppOff[p + 1] = np;
                upStart[p + 1] = up;
            }
            pcPart[tid] = std::move(pcl);
            ppPart[tid] = std::move(ppl);
        }
        parScan(t.pcOff, P, nThreads);
        t.maxPointDegree = parScan(t.ppOff, P, nThreads);
        parScan(upStart, P, nThreads);
        t.pc.resize(t.pcOff[P]);
        t.pp.resize(t.ppOff[P]);
#pragma omp parallel for schedule(static, 1)
        for (int tid = 0; tid < nThreads; ++tid)
        { // tid numbers a range of rows (a work item), not a thread
            int64_t lo, hi;
            rangeOf(P, tid, nThreads, lo, hi);
            std::copy(pcPart[tid].v.begin(), pcPart[tid].v.begin() + pcPart[tid].n, t.pc.begin() + t.pcOff[lo]);
            std::copy(ppPart[tid].v.begin(), ppPart[tid].v.begin() + ppPart[tid].n, t.pp.begin() + t.ppOff[lo]);
        }
    }
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
    t.efOff.resize(E + 1);
    t.ecOff.resize(E + 1);
    int bad = 0;
    {
        std::vector<RowBuf> efPart(nThreads), ccPart(nThreads), cpPart(nThreads);
#pragma omp parallel for schedule(static, 1)
        for (int tid = 0; tid < nThreads; ++tid)
        { // tid numbers a range of rows (a work item), not a thread
            int64_t lo, hi;
            rangeOf(E, tid, nThreads, lo, hi);
            // sum over the edges of their faces = sum over the faces of their vertices
            const int64_t guess = FV / nThreads + FV / (8 * nThreads) + 64;
            RowBuf efl(guess), ccl(guess), cpl(guess);
            std::vector<int32_t> cnt(2 * (size_t)maxRow + 2);
            for (int64_t e = lo; e < hi; ++e)
            {
                const int32_t p = t.edge[2 * e], q = t.edge[2 * e + 1];
                const int32_t nCorners = t.cornerOff[p + 1] - t.cornerOff[p];
                int32_t *fs = efl.need(nCorners), *cs = ccl.need(2 * (int64_t)nCorners), *pr = cpl.need(2 * (int64_t)nCorners);
                int32_t n = 0;
                for (int32_t s = t.cornerOff[p]; s < t.cornerOff[p + 1]; ++s)
                    if (t.corner[2 * (int64_t)s] == q || t.corner[2 * (int64_t)s + 1] == q)
                        fs[n++] = cornerFace[s];
                // the cells of those faces, in order of first appearance (owner before neighbour), each with the
                // positions of its first two faces in the row
                int32_t ncs = 0;
                bool over = false;
                for (int32_t i = 0; i < n; ++i)
                    for (int side = 0; side < 2; ++side)
                    {
                        if (side == 1 && fs[i] >= Fi)
                            break;
                        const int32_t c = side ? m.neighbour[fs[i]] : m.owner[fs[i]];
                        int32_t k = 0;
                        while (k < ncs && cs[k] != c)
                            ++k;
                        if (k == ncs)
                        {
                            cs[ncs] = c;
                            pr[ncs] = i & 0xffff;
                            cnt[ncs++] = 1;
                        }
                        else
                        {
                            if (cnt[k] == 1)
                                pr[k] |= i << 16;
                            else
                                over = true;
                            ++cnt[k];
                        }
                    }
                bool under = n >= 65536;
                for (int32_t k = 0; k < ncs; ++k)
                    if (cnt[k] == 1)
                    {
                        under = true;
                        pr[k] |= (int32_t)0xffff0000u; // no second face: position -1, as a search would leave it
                    }
                if (over || under)
                {
#pragma omp atomic write
                    bad = over ? 1 : 2;
                }
                efl.n += n;
                ccl.n += ncs;
                cpl.n += ncs;
                t.efOff[e + 1] = n;
                t.ecOff[e + 1] = ncs;
            }
            efPart[tid] = std::move(efl);
            ccPart[tid] = std::move(ccl);
            cpPart[tid] = std::move(cpl);
        }
        t.maxEdgeFaces = parScan(t.efOff, E, nThreads);
        parScan(t.ecOff, E, nThreads);
        t.ef.resize(t.efOff[E]);
        t.ecCell.resize(t.ecOff[E]);
        t.ecPair.resize(t.ecOff[E]);
#pragma omp parallel for schedule(static, 1)
        for (int tid = 0; tid < nThreads; ++tid)
        { // tid numbers a range of rows (a work item), not a thread
            int64_t lo, hi;
            rangeOf(E, tid, nThreads, lo, hi);
            std::copy(efPart[tid].v.begin(), efPart[tid].v.begin() + efPart[tid].n, t.ef.begin() + t.efOff[lo]);
            std::copy(ccPart[tid].v.begin(), ccPart[tid].v.begin() + ccPart[tid].n, t.ecCell.begin() + t.ecOff[lo]);
            std::copy(cpPart[tid].v.begin(), cpPart[tid].v.begin() + cpPart[tid].n, t.ecPair.begin() + t.ecOff[lo]);
        }
    }
    if (bad == 1)
        fail("Sanity broken, more than two edge faces belong to same cell");
    if (bad == 2)
        fail("Sanity broken, didn't find face pairs for cell");

    tick("edgeFaces / edgeCells");
    // ---- cell -> faces in OpenFOAM's accumulation order (owned faces ascending, then
    //      neighbour-side faces ascending; bit 31 marks the neighbour side) ----
    parFill(t.cfOff, (size_t)C + 1, (int32_t)0);
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
    parScan(t.cfOff, C, nThreads);
    t.cf.resize(t.cfOff[C]);
    {
        // filled in parallel, then every (short) row sorted: owned faces ascending first (bit 31 clear), then the
        // neighbour-side faces ascending -- as unsigned numbers that is exactly ascending order
        Vec<int32_t> cur;
        parCopy(cur, t.cfOff.data(), (size_t)C);
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
    t.pointRec.resize(16 * P);
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < P; ++p)
    {
        int32_t *r = &t.pointRec[16 * p];
        for (int k = 0; k < 16; ++k)
            r[k] = 0;
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

namespace
{
// label -> position in a tile's (short) list: open addressing, sized per tile, reset per tile
struct LocalMap
{
    std::vector<int32_t> key, val;
    uint32_t mask = 0;
    int shift = 0;
    void reset(size_t expected)
    {
        size_t cap = 1024;
        int bits = 10;
        while (cap < 2 * expected)
            cap *= 2, ++bits;
        if (key.size() < cap)
            key.resize(cap), val.resize(cap);
        std::fill(key.begin(), key.begin() + cap, -1);
        mask = (uint32_t)cap - 1;
        shift = 32 - bits;
    }
    uint32_t home(int32_t k) const { return ((uint32_t)k * 0x9E3779B1u) >> shift; }
    bool insert(int32_t k)
    { // true if k was not there
        uint32_t h = home(k);
        while (key[h] != -1)
        {
            if (key[h] == k)
                return false;
            h = (h + 1) & mask;
        }
        key[h] = k;
        return true;
    }
    int32_t &at(int32_t k)
    {
        uint32_t h = home(k);
        while (key[h] != k)
            h = (h + 1) & mask;
        return val[h];
    }
};

// Everything one tile contributes, built by the thread that accepted the tile; the global arrays are filled
// from these after the prefix sums over the tiles.
struct TileOut
{
    int32_t a = 0, b = 0;                 // cells keys[a .. b)
    std::vector<int32_t> cells;           // ascending
    std::vector<int32_t> faces, points;   // ascending
    std::vector<int32_t> faceRefOff;      // per listed face + 1, tile-local
    std::vector<uint16_t> faceRef;        // vertices as positions in points
    std::vector<int32_t> slotOff;         // per cell + 1, tile-local
    std::vector<uint16_t> slotRef;        // faces as positions in faces, bit 15 = neighbour side
    std::vector<int32_t> nPairs;          // per cell; -1 = not closed
    std::vector<uint16_t> pairs;          // 4 per pair, cell-major
    std::vector<uint16_t> rec;            // canonical hexahedron records, 16 per cell (valid where isHex)
    bool allHex = false, allQuads = false; // every cell a topological hexahedron / every listed face a quadrilateral
};

// The (edge, cell) pairs of one cell: half edges of its faces sorted by end points; a closed cell has every edge
// exactly twice.  keys: scratch; out: 4 x uint16 per pair (p0 < p1, then the two faces, ascending).  -1 if not closed.
inline int32_t cellPairsOf(const TileOut &T, int32_t i, std::vector<uint64_t> &keys, uint16_t *out)
{
    keys.clear();
    for (int32_t q = T.slotOff[i]; q < T.slotOff[i + 1]; ++q)
    {
        const uint16_t li = T.slotRef[q] & 0x7fff;
        const int32_t rb = T.faceRefOff[li], nv = T.faceRefOff[li + 1] - rb;
        for (int32_t v = 0; v < nv; ++v)
        {
            const uint16_t a = T.faceRef[rb + v], b = T.faceRef[rb + (v + 1 == nv ? 0 : v + 1)];
            keys.push_back((uint64_t)std::min(a, b) << 32 | (uint64_t)std::max(a, b) << 16 | li);
        }
    }
    std::sort(keys.begin(), keys.end());
    if (keys.size() % 2)
        return -1;
    int32_t n = 0;
    for (size_t i2 = 0; i2 < keys.size(); i2 += 2)
    {
        const uint64_t e0 = keys[i2] >> 16, e1 = keys[i2 + 1] >> 16;
        if (e0 != e1 || (e0 >> 16) == (e0 & 0xffff) || (i2 + 2 < keys.size() && (keys[i2 + 2] >> 16) == e0))
            return -1;
        out[4 * n] = (uint16_t)(e0 >> 16), out[4 * n + 1] = (uint16_t)(e0 & 0xffff);
        out[4 * n + 2] = (uint16_t)(keys[i2] & 0xffff), out[4 * n + 3] = (uint16_t)(keys[i2 + 1] & 0xffff);
        ++n;
    }
    return n;
}

// Canonical hexahedron record of cell i (see topology.hpp).  True only if the record's fixed pattern of twelve
// (edge, cell) pairs is exactly the cell's pair list: the six faces are distinct quadrilaterals, every pattern edge
// is a side of both of its faces, and the twelve edges are distinct -- then each face meets four distinct pattern
// edges, which are its four sides, so all 24 half edges are matched in pairs (the cell is closed) and nothing else is.
inline bool hexRecordOf(const TileOut &T, int32_t i, uint16_t rec[16])
{
    if (T.slotOff[i + 1] - T.slotOff[i] != 6)
        return false;
    uint16_t fl[6], fv[6][4];
    for (int q = 0; q < 6; ++q)
    {
        fl[q] = T.slotRef[T.slotOff[i] + q] & 0x7fff;
        const int32_t rb = T.faceRefOff[fl[q]];
        if (T.faceRefOff[fl[q] + 1] - rb != 4)
            return false;
        for (int v = 0; v < 4; ++v)
            fv[q][v] = T.faceRef[rb + v];
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
    int side[4];
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
        side[e] = S;
    }
    rec[8] = fl[0];
    rec[9] = fl[B];
    // six distinct faces (as slots of the cell and as faces of the tile)
    const int slots[6] = {0, B, side[0], side[1], side[2], side[3]};
    for (int x = 0; x < 6; ++x)
        for (int y = x + 1; y < 6; ++y)
            if (slots[x] == slots[y] || fl[slots[x]] == fl[slots[y]])
                return false;
    // every pattern edge is a side of both of its faces; the twelve edges are distinct
    auto isSide = [&](int q, uint16_t p0, uint16_t p1) {
        for (int v = 0; v < 4; ++v)
            if (fv[q][v] == p0)
                return fv[q][(v + 1) & 3] == p1 || fv[q][(v + 3) & 3] == p1;
        return false;
    };
    uint32_t edges[12];
    int ne = 0;
    auto add = [&](uint16_t p0, uint16_t p1, int q0, int q1) {
        edges[ne++] = (uint32_t)std::min(p0, p1) << 16 | std::max(p0, p1);
        return p0 != p1 && isSide(q0, p0, p1) && isSide(q1, p0, p1);
    };
    for (int e = 0; e < 4; ++e)
        if (!add(rec[e], rec[(e + 1) & 3], 0, side[e]) || !add(rec[4 + e], rec[4 + ((e + 1) & 3)], B, side[e]) ||
            !add(rec[e], rec[4 + e], side[(e + 3) & 3], side[e]))
            return false;
    for (int x = 0; x < 12; ++x)
        for (int y = x + 1; y < 12; ++y)
            if (edges[x] == edges[y])
                return false;
    return true;
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
    {
        double l0 = 1e300, l1 = 1e300, l2 = 1e300, h0 = -1e300, h1 = -1e300, h2 = -1e300;
#pragma omp parallel for schedule(static) reduction(min : l0, l1, l2) reduction(max : h0, h1, h2)
        for (int64_t p = 0; p < P; ++p)
        {
            l0 = std::min(l0, m.points[3 * p]), h0 = std::max(h0, m.points[3 * p]);
            l1 = std::min(l1, m.points[3 * p + 1]), h1 = std::max(h1, m.points[3 * p + 1]);
            l2 = std::min(l2, m.points[3 * p + 2]), h2 = std::max(h2, m.points[3 * p + 2]);
        }
        lo[0] = l0, lo[1] = l1, lo[2] = l2, hi[0] = h0, hi[1] = h1, hi[2] = h2;
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
    Vec<std::pair<uint64_t, int32_t>> keys(C);
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
    // runs are closed at maxCells cells and at every aligned block of 2^8 keys (an 8 x 8 x 4 brick of a
    // block-structured mesh), so that a brick cut short by the mesh boundary does not shift all later tiles
    // off the brick grid (271^3: 33 full bricks and one of 7 cells per row)
    const int alignBits = maxCells >= 256 ? 8 : (maxCells == 128 ? 7 : 0);
    std::vector<std::pair<int32_t, int32_t>> runs;
    for (int64_t i = 0; i < C;)
    {
        int64_t j = i;
        const uint64_t block = keys[i].first >> alignBits;
        while (j < C && j - i < maxCells && (alignBits == 0 || (keys[j].first >> alignBits) == block))
            ++j;
        runs.push_back({(int32_t)i, (int32_t)j});
        i = j;
    }
    const int64_t nRuns = (int64_t)runs.size();
    std::vector<std::vector<TileOut>> perRun(nRuns);
    bool cellTooLarge = false, closed = true;
    double tCollect = 0, tFinish = 0, tPairs = 0;
#pragma omp parallel reduction(+ : tCollect, tFinish, tPairs)
    {
        LocalMap faceMap, pointMap;
        std::vector<uint64_t> scratch;
        // lists of a run of cells; false if they do not fit the budgets
        auto collect = [&](TileOut &T) -> bool {
            size_t nSlots = 0;
            for (int32_t i = T.a; i < T.b; ++i)
                nSlots += t.cfOff[keys[i].second + 1] - t.cfOff[keys[i].second];
            faceMap.reset(nSlots);
            T.faces.clear();
            for (int32_t i = T.a; i < T.b; ++i)
            {
                const int32_t c = keys[i].second;
                for (int32_t k = t.cfOff[c]; k < t.cfOff[c + 1]; ++k)
                    if (faceMap.insert(t.cf[k] & 0x7fffffff))
                        T.faces.push_back(t.cf[k] & 0x7fffffff);
            }
            if ((int)T.faces.size() > maxFaces)
                return false;
            std::sort(T.faces.begin(), T.faces.end());
            size_t nVerts = 0;
            for (int32_t f : T.faces)
                nVerts += m.faceOffsets[f + 1] - m.faceOffsets[f];
            pointMap.reset(nVerts);
            T.points.clear();
            for (int32_t f : T.faces)
                for (int32_t q = m.faceOffsets[f]; q < m.faceOffsets[f + 1]; ++q)
                    if (pointMap.insert(m.faceVerts[q]))
                        T.points.push_back(m.faceVerts[q]);
            if ((int)T.points.size() > maxPoints)
                return false;
            std::sort(T.points.begin(), T.points.end());
            return true;
        };
        // references, pairs and records of an accepted tile (the maps still hold its labels)
        auto finish = [&](TileOut &T) {
            for (size_t i = 0; i < T.faces.size(); ++i)
                faceMap.at(T.faces[i]) = (int32_t)i;
            for (size_t i = 0; i < T.points.size(); ++i)
                pointMap.at(T.points[i]) = (int32_t)i;
            T.cells.clear();
            for (int32_t i = T.a; i < T.b; ++i)
                T.cells.push_back(keys[i].second);
            std::sort(T.cells.begin(), T.cells.end());
            T.faceRefOff.assign(1, 0);
            T.allQuads = true;
            for (int32_t f : T.faces)
            {
                for (int32_t q = m.faceOffsets[f]; q < m.faceOffsets[f + 1]; ++q)
                    T.faceRef.push_back((uint16_t)pointMap.at(m.faceVerts[q]));
                T.faceRefOff.push_back((int32_t)T.faceRef.size());
                T.allQuads = T.allQuads && m.faceOffsets[f + 1] - m.faceOffsets[f] == 4;
            }
            T.slotOff.assign(1, 0);
            for (int32_t c : T.cells)
            {
                for (int32_t q = t.cfOff[c]; q < t.cfOff[c + 1]; ++q)
                {
                    const int32_t w = t.cf[q];
                    T.slotRef.push_back((uint16_t)(faceMap.at(w & 0x7fffffff) | (w < 0 ? 0x8000 : 0)));
                }
                T.slotOff.push_back((int32_t)T.slotRef.size());
            }
            // (edge, cell) pairs of the fused face-angle filter: the edges of a cell are the vertex pairs that two
            // of its faces share; both faces and both end points are already tile-local references
            const double p0 = timing ? clock() : 0;
            const int32_t nc = (int32_t)T.cells.size();
            T.nPairs.assign(nc, 0);
            T.rec.assign(16 * (size_t)nc, 0);
            T.allHex = true;
            // cells that are hexahedra by the record's own check have twelve pairs and are closed; the pair list is
            // only materialised where it is read (tiles off the fast path, or on request)
            std::vector<uint8_t> isHex(nc, 0);
            for (int32_t i = 0; i < nc; ++i)
            {
                isHex[i] = hexRecordOf(T, i, &T.rec[16 * (size_t)i]) ? 1 : 0;
                T.allHex = T.allHex && isHex[i];
            }
            const bool listAll = keepPairs || !(T.allHex && T.allQuads);
            for (int32_t i = 0; i < nc; ++i)
            {
                if (isHex[i] && !listAll)
                {
                    T.nPairs[i] = 12;
                    continue;
                }
                int32_t nHalf = 0;
                for (int32_t q = T.slotOff[i]; q < T.slotOff[i + 1]; ++q)
                {
                    const int32_t li = T.slotRef[q] & 0x7fff;
                    nHalf += T.faceRefOff[li + 1] - T.faceRefOff[li];
                }
                const size_t at = T.pairs.size();
                T.pairs.resize(at + 2 * (size_t)nHalf + 4);
                const int32_t n = cellPairsOf(T, i, scratch, T.pairs.data() + at);
                T.nPairs[i] = n;
                T.pairs.resize(at + 4 * (size_t)std::max(n, 0));
            }
            tPairs += (timing ? clock() : 0) - p0;
        };
#pragma omp for schedule(dynamic, 16)
        for (int64_t r = 0; r < nRuns; ++r)
        {
            std::vector<std::pair<int32_t, int32_t>> todo;
            todo.push_back(runs[r]);
            while (!todo.empty())
            {
                TileOut T;
                T.a = todo.back().first, T.b = todo.back().second;
                todo.pop_back();
                const double c0 = timing ? clock() : 0;
                const bool fits = collect(T);
                const double c1 = timing ? clock() : 0;
                tCollect += c1 - c0;
                if (fits)
                {
                    finish(T);
                    tFinish += (timing ? clock() : 0) - c1;
                    for (int32_t n : T.nPairs)
                        if (n < 0)
                        {
#pragma omp atomic write
                            closed = false;
                        }
                    perRun[r].push_back(std::move(T));
                }
                else if (T.b - T.a == 1)
                {
#pragma omp atomic write
                    cellTooLarge = true;
                }
                else
                { // second half first on the stack, so the first half is finished first (keeps the order)
                    const int32_t mid = T.a + (T.b - T.a) / 2;
                    todo.push_back({mid, T.b});
                    todo.push_back({T.a, mid});
                }
            }
        }
    }
    if (timing)
        fprintf(stderr, "[smgpu set-up]   thread-seconds: lists %.3f, references + pairs %.3f (pairs %.3f)\n", tCollect, tFinish, tPairs);
    tick("tiles: lists, references, pairs");
    if (cellTooLarge)
        return GeomTiles(); // a cell that does not fit a tile: the caller keeps the two-kernel path
    std::vector<TileOut *> tiles;
    for (auto &v : perRun)
        for (TileOut &T : v)
            tiles.push_back(&T);
    G.nTiles = (int32_t)tiles.size();
    const int32_t nT = G.nTiles;
    // offsets; uniform tiles (all faces quadrilaterals, all cells topological hexahedra) are counted separately
    G.tileCellOff.assign(nT + 1, 0);
    G.tileFaceOff.assign(nT + 1, 0);
    G.tilePointOff.assign(nT + 1, 0);
    G.tileUFaceOff.assign(nT, -1);
    G.tileUCellOff.assign(nT, -1);
    std::vector<int64_t> slotRefOff(nT + 1, 0), faceRefBase(nT + 1, 0), pairBase(nT + 1, 0);
    int64_t uf = 0, uc = 0;
    for (int32_t k = 0; k < nT; ++k)
    {
        const TileOut &T = *tiles[k];
        G.tileCellOff[k + 1] = G.tileCellOff[k] + (int32_t)T.cells.size();
        G.tileFaceOff[k + 1] = G.tileFaceOff[k] + (int32_t)T.faces.size();
        G.tilePointOff[k + 1] = G.tilePointOff[k] + (int32_t)T.points.size();
        slotRefOff[k + 1] = slotRefOff[k] + (int64_t)T.slotRef.size();
        faceRefBase[k + 1] = faceRefBase[k] + (int64_t)T.faceRef.size();
        G.maxTileCells = std::max(G.maxTileCells, (int32_t)T.cells.size());
        G.maxTileFaces = std::max(G.maxTileFaces, (int32_t)T.faces.size());
        G.maxTilePoints = std::max(G.maxTilePoints, (int32_t)T.points.size());
        const bool uniform = closed && T.allHex && T.allQuads;
        if (uniform)
        {
            G.tileUFaceOff[k] = (int32_t)uf;
            G.tileUCellOff[k] = (int32_t)uc;
            uf += (int64_t)T.faces.size();
            uc += (int64_t)T.cells.size();
        }
        int64_t np = 0;
        if (closed && (keepPairs || !uniform))
            for (int32_t n : T.nPairs)
                np += n;
        pairBase[k + 1] = pairBase[k] + np;
        G.maxTileEdgePairs = std::max<int64_t>(G.maxTileEdgePairs, np);
    }
    if (faceRefBase[nT] >= (int64_t)INT32_MAX)
        return GeomTiles();
    // the first tile (in tile order) that lists a face stores the face's global outputs
    Vec<int32_t> firstTile;
    parFill(firstTile, (size_t)F, INT32_MAX);
#pragma omp parallel for schedule(dynamic, 64)
    for (int32_t k = 0; k < nT; ++k)
        for (int32_t f : tiles[k]->faces)
        {
            int32_t seen;
#pragma omp atomic read
            seen = firstTile[f];
            while (k < seen)
            { // atomic minimum
                if (__atomic_compare_exchange_n(&firstTile[f], &seen, k, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED))
                    break;
            }
        }
    tick("tiles: offsets, first tile");
    G.tileCells.resize(C);
    G.tileFaces.resize(G.tileFaceOff[nT]);
    G.tilePoints.resize(G.tilePointOff[nT]);
    G.slotOff.resize(C + 1);
    G.slotRef.resize(slotRefOff[nT]);
    G.faceRefOff.resize(G.tileFaceOff[nT] + 1);
    G.faceRef.resize(faceRefBase[nT]);
    G.slotOff[C] = (int32_t)slotRefOff[nT];
    G.faceRefOff[G.tileFaceOff[nT]] = (int32_t)faceRefBase[nT];
    G.nUniformCells = uc;
    G.uFaceRef.resize(4 * (size_t)uf);
    G.uSlotRef.resize(6 * (size_t)uc);
    G.hexRec.resize(16 * (size_t)uc);
    const int64_t totalPairs = closed ? pairBase[nT] : 0;
    const bool storePairs = totalPairs > 0 && totalPairs < (int64_t)INT32_MAX / 2;
    if (storePairs)
    {
        G.cellEdgeOff.resize(C + 1);
        G.cellEdgeOff[C] = (int32_t)totalPairs;
        G.cellEdgeRef.resize(4 * (size_t)totalPairs);
    }
    else
        G.maxTileEdgePairs = 0;
    const int32_t firstCellEdges = closed ? tiles[0]->nPairs[0] : 0;
    bool sameEdges = closed;
#pragma omp parallel for schedule(dynamic, 16) reduction(&& : sameEdges)
    for (int32_t k = 0; k < nT; ++k)
    {
        const TileOut &T = *tiles[k];
        const int32_t fb = G.tileFaceOff[k], cb = G.tileCellOff[k], nf = (int32_t)T.faces.size(), nc = (int32_t)T.cells.size();
        std::copy(T.points.begin(), T.points.end(), G.tilePoints.begin() + G.tilePointOff[k]);
        for (int32_t i = 0; i < nf; ++i)
        {
            const int32_t f = T.faces[i];
            G.tileFaces[fb + i] = (firstTile[f] == k) ? (int32_t)(f | 0x80000000u) : f;
            G.faceRefOff[fb + i] = (int32_t)(faceRefBase[k] + T.faceRefOff[i]);
        }
        std::copy(T.faceRef.begin(), T.faceRef.end(), G.faceRef.begin() + faceRefBase[k]);
        for (int32_t i = 0; i < nc; ++i)
        {
            G.tileCells[cb + i] = T.cells[i];
            G.slotOff[cb + i] = (int32_t)(slotRefOff[k] + T.slotOff[i]);
        }
        std::copy(T.slotRef.begin(), T.slotRef.end(), G.slotRef.begin() + slotRefOff[k]);
        if (G.tileUCellOff[k] >= 0)
        { // fixed-stride copies of the references, and the records: first halves of all cells, then second halves
            const size_t ufb = (size_t)G.tileUFaceOff[k], ucb = (size_t)G.tileUCellOff[k];
            std::copy(T.faceRef.begin(), T.faceRef.end(), G.uFaceRef.begin() + 4 * ufb);
            std::copy(T.slotRef.begin(), T.slotRef.end(), G.uSlotRef.begin() + 6 * ucb);
            for (int32_t i = 0; i < nc; ++i)
            {
                const uint16_t *rec = &T.rec[16 * (size_t)i];
                uint16_t *o1 = &G.hexRec[16 * ucb + 8 * (size_t)i], *o2 = &G.hexRec[16 * ucb + 8 * (size_t)nc + 8 * (size_t)i];
                for (int q = 0; q < 8; ++q)
                    o1[q] = rec[q], o2[q] = rec[8 + q];
            }
        }
        for (int32_t n : T.nPairs)
            sameEdges = sameEdges && n == firstCellEdges;
        if (storePairs)
        {
            const bool mine = pairBase[k + 1] > pairBase[k];
            int64_t at = pairBase[k];
            for (int32_t i = 0; i < nc; ++i)
            {
                G.cellEdgeOff[cb + i] = (int32_t)at;
                if (mine)
                    at += T.nPairs[i];
            }
            if (mine)
                std::copy(T.pairs.begin(), T.pairs.end(), G.cellEdgeRef.begin() + 4 * pairBase[k]);
        }
    }
    G.uniformCellEdges = sameEdges ? firstCellEdges : 0;
    tick("tiles: global arrays");
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
