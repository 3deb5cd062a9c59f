// boundary.cpp -- see boundary.hpp.  Citations are to /root/reference/src/boundaryPointSmoothing.C unless
// a file is named.
#include "boundary.hpp"
#include "sm_math.h"

#include <algorithm>
#include <cmath>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace sm
{
namespace
{
struct V
{
    double x, y, z;
};
inline V operator+(V a, V b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V operator-(V a, V b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V operator*(double s, V a) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(V a, V b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline double mag(V a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
inline V at(const std::vector<double> &p, int64_t i) { return {p[3 * i], p[3 * i + 1], p[3 * i + 2]}; }
[[noreturn]] void fail(const std::string &s) { throw std::runtime_error(s); }

// projectPointToEdge, :89-145
void projectPointToEdge(V pt, const EdgeMesh &em, int32_t edgeI, double distanceTolerance, V &projPoint, int32_t &edgePointI)
{
    edgePointI = -1;
    const int32_t startPointI = em.edges[2 * edgeI], endPointI = em.edges[2 * edgeI + 1];
    const V startPoint = at(em.points, startPointI), endPoint = at(em.points, endPointI);
    const double edgeLength = mag(endPoint - startPoint);
    const V c2pt = pt - startPoint, edgeVec = endPoint - startPoint;
    const double normalizedDotProd = dot(c2pt, edgeVec) / (edgeLength * edgeLength);
    const V testProjPoint = startPoint + normalizedDotProd * edgeVec;
    if (normalizedDotProd <= 1e-6) // ABS_TOL
    {
        projPoint = startPoint;
        if (mag(testProjPoint - startPoint) <= distanceTolerance)
            edgePointI = startPointI;
    }
    else if (normalizedDotProd >= (1.0 - 1e-6))
    {
        projPoint = endPoint;
        if (mag(testProjPoint - endPoint) <= distanceTolerance)
            edgePointI = endPointI;
    }
    else
        projPoint = testProjPoint;
}

// findClosestEdgeInfo, :206-263
void findClosestEdgeInfo(V pt, const EdgeMesh &em, int32_t requiredStringI, const std::vector<int32_t> &strings,
                         double distanceTolerance, V &projPoint, int32_t &closestEdgeStringI, int32_t &closestEdgePointI)
{
    double distance = SM_GREAT;
    projPoint = {SM_GREAT, SM_GREAT, SM_GREAT};
    closestEdgeStringI = closestEdgePointI = -1;
    for (int64_t edgeI = 0; edgeI < em.nEdges(); ++edgeI)
    {
        if (requiredStringI >= 0 && strings[edgeI] != requiredStringI)
            continue;
        V testProjPoint;
        int32_t edgePointI;
        projectPointToEdge(pt, em, (int32_t)edgeI, distanceTolerance, testProjPoint, edgePointI);
        const double testDistance = mag(testProjPoint - pt);
        if (testDistance < distance)
        {
            distance = testDistance;
            projPoint = testProjPoint;
            closestEdgePointI = edgePointI;
            if (em.nEdges() == (int64_t)strings.size())
                closestEdgeStringI = strings[edgeI];
        }
    }
    if (requiredStringI >= 0 && closestEdgeStringI == -1)
        fail("Internal sanity check failed: Did not find any edges with string index " + std::to_string(requiredStringI));
}

// checkEdgeMeshSanity, :20-82.  The perimeters use the reference's formula (max z PLUS min z, :76 and
// src/smoothMesh.C:1538); the test at :77 calls unqualified abs() on a double, which in a plain GCC build
// is ::abs(int): the ratio is truncated towards zero first.
void checkEdgeMeshSanity(const EdgeMesh &em, double meshMinEdgeLength, double meshPerimeter)
{
    double minEdgeLength = SM_VGREAT;
    double lo[3] = {SM_VGREAT, SM_VGREAT, SM_VGREAT}, hi[3] = {-SM_VGREAT, -SM_VGREAT, -SM_VGREAT};
    for (int64_t e = 0; e < em.nEdges(); ++e)
    {
        const V a = at(em.points, em.edges[2 * e]), b = at(em.points, em.edges[2 * e + 1]);
        const double edgeLength = mag(b - a);
        if (edgeLength < minEdgeLength)
            minEdgeLength = edgeLength;
        for (const V &q : {a, b})
        {
            const double c[3] = {q.x, q.y, q.z};
            for (int d = 0; d < 3; ++d)
            {
                if (c[d] < lo[d])
                    lo[d] = c[d];
                if (c[d] > hi[d])
                    hi[d] = c[d];
            }
        }
    }
    if (minEdgeLength < 1e-4 * meshMinEdgeLength)
        fail("Minimum edge length in edge mesh " + std::to_string(minEdgeLength) +
             " is too small in comparison to minimum edge length in polyMesh " + std::to_string(meshMinEdgeLength));
    const double emPerimeter = hi[0] - lo[0] + hi[1] - lo[1] + hi[2] + lo[2];
    const double ratio = (emPerimeter / meshPerimeter) - 1.0;
    if (std::abs((int)ratio) > 0.5)
        fail("Perimeter (sum of bounding box side lengths) of edge mesh " + std::to_string(emPerimeter) +
             " is too different in comparison to perimeter of polyMesh " + std::to_string(meshPerimeter));
}

// findContinuousEdgeMeshEdges :446-486, stringifyEdgeMeshEdges :492-551 (the "not a corner" test at :521 / :529
// indexes pointEdges with an EDGE label; kept, an edge label beyond the point list reads as "not 2")
void continuousEdges(const EdgeMesh &em, int32_t edgeI, int32_t &n1, int32_t &n2)
{
    n1 = n2 = -1;
    const int32_t p1 = em.edges[2 * edgeI], p2 = em.edges[2 * edgeI + 1];
    if (em.pointEdges[p1].size() == 2)
        n1 = em.pointEdges[p1][0] == edgeI ? em.pointEdges[p1][1] : em.pointEdges[p1][0];
    if (em.pointEdges[p2].size() == 2)
        n2 = em.pointEdges[p2][0] == edgeI ? em.pointEdges[p2][1] : em.pointEdges[p2][0];
}
void stringify(const EdgeMesh &em, std::vector<int32_t> &strings, int32_t edgeI, int32_t n1, int32_t n2, int32_t &nStrings)
{
    const int32_t s0 = strings[edgeI], s1 = n1 >= 0 ? strings[n1] : -1, s2 = n2 >= 0 ? strings[n2] : -1;
    const int32_t mx = std::max(std::max(s0, s1), s2);
    if (mx == -1)
        strings[edgeI] = ++nStrings;
    else if (s0 == -1)
        strings[edgeI] = mx;
    auto twoEdgesAt = [&](int32_t label) { return label < (int32_t)em.pointEdges.size() && em.pointEdges[label].size() == 2; };
    if (n1 >= 0 && s1 == -1 && twoEdgesAt(n1))
    {
        int32_t a, b;
        continuousEdges(em, n1, a, b);
        stringify(em, strings, n1, a, b, nStrings);
    }
    if (n2 >= 0 && s2 == -1 && twoEdgesAt(n2))
    {
        int32_t a, b;
        continuousEdges(em, n2, a, b);
        stringify(em, strings, n2, a, b, nStrings);
    }
}
} // namespace

void EdgeMesh::finish()
{
    pointEdges.assign(nPoints(), {});
    for (int64_t e = 0; e < nEdges(); ++e)
    {
        pointEdges[edges[2 * e]].push_back((int32_t)e);
        pointEdges[edges[2 * e + 1]].push_back((int32_t)e);
    }
}

void readObj(const std::string &file, std::vector<double> &points, std::vector<int32_t> &edges, std::vector<int32_t> &tris)
{
    std::ifstream in(file);
    if (!in.good())
        fail("cannot open " + file);
    std::string line;
    while (std::getline(in, line))
    {
        std::istringstream is(line);
        std::string tag;
        if (!(is >> tag))
            continue;
        if (tag == "v")
        {
            double x, y, z;
            if (is >> x >> y >> z)
            {
                points.push_back(x);
                points.push_back(y);
                points.push_back(z);
            }
        }
        else if (tag == "l" || tag == "f")
        {
            std::vector<int32_t> ids;
            std::string w;
            while (is >> w)
                ids.push_back(atoi(w.c_str()) - 1); // "7/1/3" -> 7
            if (tag == "l")
                for (size_t i = 0; i + 1 < ids.size(); ++i)
                {
                    edges.push_back(ids[i]);
                    edges.push_back(ids[i + 1]);
                }
            else
                for (size_t i = 1; i + 1 < ids.size(); ++i)
                {
                    tris.push_back(ids[0]);
                    tris.push_back(ids[i]);
                    tris.push_back(ids[i + 1]);
                }
        }
    }
}

void meshBoundingBox(const Topology &t, const std::vector<double> &points, double lo[3], double hi[3])
{
    for (int d = 0; d < 3; ++d)
        lo[d] = SM_VGREAT, hi[d] = -SM_VGREAT;
    for (int64_t e = 0; e < t.E; ++e)
        for (int s = 0; s < 2; ++s)
            for (int d = 0; d < 3; ++d)
            {
                const double c = points[3 * (int64_t)t.edge[2 * e + s] + d];
                if (c < lo[d])
                    lo[d] = c;
                if (c > hi[d])
                    hi[d] = c;
            }
}

BoundarySetup buildBoundarySetup(const PolyMesh &m, const Topology &t, const std::vector<double> &points, const EdgeMesh &initEdges,
                                 const EdgeMesh &targetEdgesIn, const TriSurface &surface, const std::vector<int32_t> &patchSmoothing,
                                 double layerEdgeLength, const std::vector<int32_t> &cornerIO, const std::vector<int32_t> &featureIO,
                                 const BoundaryParallel *par)
{
    BoundarySetup B;
    // src/smoothMesh.C:2067-2078: do the label lists of an earlier run hold classification data
    bool labelIOListsHaveData = false;
    for (int32_t v : cornerIO)
        labelIOListsHaveData = labelIOListsHaveData || v == 1;
    for (int32_t v : featureIO)
        labelIOListsHaveData = labelIOListsHaveData || v == 1;
    const int64_t P = t.P;
    B.targetEdges = targetEdgesIn;
    B.surface = surface;
    const double meshMinEdgeLength = par ? par->meshMinEdgeLength : t.minEdgeLength;
    B.distanceTolerance = 1e-4 * ((meshMinEdgeLength < layerEdgeLength) ? meshMinEdgeLength : layerEdgeLength); // :1921
    // getMeshStats' perimeter over the edge end points, src/smoothMesh.C:1495-1538
    double meshPerimeter;
    if (par)
        meshPerimeter = par->meshPerimeter;
    else
    {
        double lo[3], hi[3];
        meshBoundingBox(t, points, lo, hi);
        meshPerimeter = hi[0] - lo[0] + hi[1] - lo[1] + hi[2] + lo[2];
    }
    checkEdgeMeshSanity(initEdges, meshMinEdgeLength, meshPerimeter);
    checkEdgeMeshSanity(B.targetEdges, meshMinEdgeLength, meshPerimeter);
    // findEdgeMeshStrings, :557-590
    B.targetEdgeStrings.assign(B.targetEdges.nEdges(), -1);
    B.nStrings = -1;
    for (int64_t e = 0; e < B.targetEdges.nEdges(); ++e)
    {
        if (B.targetEdgeStrings[e] >= 0)
            continue;
        int32_t n1, n2;
        continuousEdges(B.targetEdges, (int32_t)e, n1, n2);
        stringify(B.targetEdges, B.targetEdgeStrings, (int32_t)e, n1, n2, B.nStrings);
    }
    // classifyBoundaryPoints, :269-440: once per boundary point, by the first patch that contains it
    B.isCorner.assign(P, 0);
    B.isFeatureEdge.assign(P, 0);
    B.isSmoothingSurface.assign(P, 0);
    B.isConnectedToInternal.assign(P, 0);
    B.cornerPoints.assign(3 * P, SM_GREAT);
    std::vector<uint8_t> visited(P, 0);
    for (size_t pi = 0; pi < m.patches.size(); ++pi)
        for (int32_t f = m.patches[pi].start; f < m.patches[pi].start + m.patches[pi].size; ++f)
            for (int32_t k = t.faceOff[f]; k < t.faceOff[f + 1]; ++k)
            {
                const int32_t p = t.faceVerts[k];
                if (visited[p])
                    continue;
                visited[p] = 1;
                if (t.isInternal[p])
                    continue;
                for (int32_t s = t.ppOff[p]; s < t.ppOff[p + 1]; ++s)
                    if (t.isInternal[t.pp[s]])
                        B.isConnectedToInternal[p] = 1;
                if (initEdges.nPoints() > 0 && B.targetEdges.nPoints() > 0)
                {
                    const V pt = at(points, p);
                    if (labelIOListsHaveData)
                    { // :336-340
                        B.isCorner[p] = (p < (int32_t)cornerIO.size() && cornerIO[p] == 1) ? 1 : 0;
                        B.isFeatureEdge[p] = (p < (int32_t)featureIO.size() && featureIO[p] == 1) ? 1 : 0;
                    }
                    else
                    {
                        V projPoint;
                        int32_t stringI, closestEdgePointI;
                        findClosestEdgeInfo(pt, initEdges, -1, B.targetEdgeStrings, B.distanceTolerance, projPoint, stringI,
                                            closestEdgePointI);
                        if (closestEdgePointI >= 0 && initEdges.pointEdges[closestEdgePointI].size() != 2)
                            B.isCorner[p] = 1;
                        else if (mag(pt - projPoint) < B.distanceTolerance)
                            B.isFeatureEdge[p] = 1;
                    }
                    if (B.isCorner[p])
                    { // findClosestEdgeMeshCornerPointIndex, :151-186
                        double distance = SM_GREAT;
                        int64_t closest = -1;
                        for (int64_t q = 0; q < B.targetEdges.nPoints(); ++q)
                        {
                            if (B.targetEdges.pointEdges[q].size() == 2)
                                continue;
                            const double d = mag(pt - at(B.targetEdges.points, q));
                            if (d < distance)
                            {
                                distance = d;
                                closest = q;
                            }
                        }
                        if (closest < 0)
                            fail("Did not find any eligible corner points in edge mesh");
                        for (int d = 0; d < 3; ++d)
                            B.cornerPoints[3 * (int64_t)p + d] = B.targetEdges.points[3 * closest + d];
                        ++B.nCorners;
                    }
                    if (B.isFeatureEdge[p])
                        ++B.nFeatureEdgePoints;
                }
                if (patchSmoothing[pi])
                {
                    B.isSmoothingSurface[p] = 1;
                    ++B.nSmoothingSurfacePoints;
                }
            }
    // calculatePointHopsToBoundary(smoothingPatchIds, maxIter = 2), src/smoothMesh.C:2218
    B.hopsToSmoothing.assign(P, -1);
    for (size_t pi = 0; pi < m.patches.size(); ++pi)
    {
        if (!patchSmoothing[pi])
            continue;
        for (int32_t f = m.patches[pi].start; f < m.patches[pi].start + m.patches[pi].size; ++f)
            for (int32_t k = t.faceOff[f]; k < t.faceOff[f + 1]; ++k)
                if (B.isConnectedToInternal[t.faceVerts[k]])
                    B.hopsToSmoothing[t.faceVerts[k]] = 0;
    }
    {
        std::vector<int32_t> newHops(P, -1);
        for (int iter = 0; iter < 2; ++iter)
        {
            for (int64_t p = 0; p < P; ++p)
            {
                if (B.hopsToSmoothing[p] >= 0 || !t.isInternal[p])
                    continue;
                int32_t mx = -1;
                for (int32_t s = t.ppOff[p]; s < t.ppOff[p + 1]; ++s)
                    mx = std::max(mx, B.hopsToSmoothing[t.pp[s]]);
                if (mx >= 0)
                    newHops[p] = mx + 1;
            }
            for (int64_t p = 0; p < P; ++p)
                if (newHops[p] > B.hopsToSmoothing[p])
                    B.hopsToSmoothing[p] = newHops[p];
            if (par && par->maxInt)
                par->maxInt(B.hopsToSmoothing);
        }
    }
    // propagateInnerNeighInfo, src/orthogonalBoundaryBlending.C:397-458
    B.pointToInner.assign(P, -1);
    for (int64_t p = 0; p < P; ++p)
    {
        if (!B.isSmoothingSurface[p] || !B.isConnectedToInternal[p])
            continue;
        if (B.hopsToSmoothing[p] != 0)
            fail(std::to_string(p) + " is not boundary point");
        int32_t n = 0, q = -1;
        for (int32_t s = t.ppOff[p]; s < t.ppOff[p + 1]; ++s)
            if (B.hopsToSmoothing[t.pp[s]] == 1)
            {
                ++n;
                q = t.pp[s];
            }
        if (n == 1)
            B.pointToInner[p] = q;
    }
    // src/smoothMesh.C:2234-2250: the target edge string of every feature edge point
    B.pointStrings.assign(P, -1);
    for (int64_t p = 0; p < P; ++p)
    {
        if (!B.isFeatureEdge[p])
            continue;
        V dummy;
        int32_t stringI, edgePointI;
        findClosestEdgeInfo(at(points, p), B.targetEdges, -1, B.targetEdgeStrings, B.distanceTolerance, dummy, stringI, edgePointI);
        B.pointStrings[p] = stringI;
    }
    for (int64_t p = 0; p < P; ++p)
        if (!t.isInternal[p])
            B.boundaryPoints.push_back((int32_t)p);
    return B;
}


// ------------------------------------------------------------------ surface ray casts ----
namespace
{
inline V crossV(V a, V b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
// Moeller-Trumbore segment / triangle test, the operation order shared with the device kernel and the oracle
inline bool triangleHit(const TriSurface &s, int32_t i, V start, V dir, double &t)
{
    const V p0 = at(s.points, s.tris[3 * i]), p1 = at(s.points, s.tris[3 * i + 1]), p2 = at(s.points, s.tris[3 * i + 2]);
    const V e1 = p1 - p0, e2 = p2 - p0;
    const V h = crossV(dir, e2);
    const double det = dot(e1, h);
    if (std::fabs(det) < SM_VSMALL)
        return false;
    const double inv = 1.0 / det;
    const V sv = start - p0;
    const double u = inv * dot(sv, h);
    if (u < 0.0 || u > 1.0)
        return false;
    const V q = crossV(sv, e1);
    const double v = inv * dot(dir, q);
    if (v < 0.0 || u + v > 1.0)
        return false;
    t = inv * dot(e2, q);
    return !(t < 0.0 || t > 1.0);
}
} // namespace

TriangleBvh buildTriangleBvh(const TriSurface &s, int leafSize)
{
    TriangleBvh B;
    const int32_t n = (int32_t)s.nTris();
    B.order.resize(n);
    for (int32_t i = 0; i < n; ++i)
        B.order[i] = i;
    std::vector<double> lo(3 * (size_t)n), hi(3 * (size_t)n), ctr(3 * (size_t)n);
    for (int32_t i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d)
        {
            const double a = s.points[3 * (size_t)s.tris[3 * i] + d], b = s.points[3 * (size_t)s.tris[3 * i + 1] + d],
                         c = s.points[3 * (size_t)s.tris[3 * i + 2] + d];
            lo[3 * (size_t)i + d] = std::min(a, std::min(b, c));
            hi[3 * (size_t)i + d] = std::max(a, std::max(b, c));
            ctr[3 * (size_t)i + d] = (a + b + c) / 3.0;
        }
    struct Job
    {
        int32_t node, first, count;
    };
    std::vector<Job> stack;
    auto newNode = [&]() {
        B.box.resize(B.box.size() + 6);
        B.left.push_back(-1);
        B.right.push_back(-1);
        B.first.push_back(0);
        B.count.push_back(0);
        return (int32_t)B.right.size() - 1;
    };
    if (n == 0)
        return B;
    stack.push_back({newNode(), 0, n});
    while (!stack.empty())
    {
        const Job j = stack.back();
        stack.pop_back();
        double blo[3] = {SM_VGREAT, SM_VGREAT, SM_VGREAT}, bhi[3] = {-SM_VGREAT, -SM_VGREAT, -SM_VGREAT};
        double clo[3] = {SM_VGREAT, SM_VGREAT, SM_VGREAT}, chi[3] = {-SM_VGREAT, -SM_VGREAT, -SM_VGREAT};
        for (int32_t k = j.first; k < j.first + j.count; ++k)
            for (int d = 0; d < 3; ++d)
            {
                const size_t q = 3 * (size_t)B.order[k] + d;
                blo[d] = std::min(blo[d], lo[q]);
                bhi[d] = std::max(bhi[d], hi[q]);
                clo[d] = std::min(clo[d], ctr[q]);
                chi[d] = std::max(chi[d], ctr[q]);
            }
        for (int d = 0; d < 3; ++d)
        { // inflate: the slab test must never reject a segment that a triangle test would accept
            const double pad = 1e-9 * (std::fabs(blo[d]) + std::fabs(bhi[d]) + (bhi[d] - blo[d])) + 1e-300;
            B.box[6 * (size_t)j.node + d] = blo[d] - pad;
            B.box[6 * (size_t)j.node + 3 + d] = bhi[d] + pad;
        }
        int axis = 0;
        for (int d = 1; d < 3; ++d)
            if (chi[d] - clo[d] > chi[axis] - clo[axis])
                axis = d;
        if (j.count <= leafSize || !(chi[axis] > clo[axis]))
        {
            B.first[j.node] = j.first;
            B.count[j.node] = j.count;
            continue;
        }
        const int32_t mid = j.first + j.count / 2;
        std::nth_element(B.order.begin() + j.first, B.order.begin() + mid, B.order.begin() + j.first + j.count,
                         [&](int32_t a, int32_t b) {
                             const double ca = ctr[3 * (size_t)a + axis], cb = ctr[3 * (size_t)b + axis];
                             return ca < cb || (ca == cb && a < b);
                         });
        const int32_t left = newNode();
        const int32_t right = newNode();
        B.left[j.node] = left;
        B.right[j.node] = right;
        stack.push_back({right, mid, j.first + j.count - mid});
        stack.push_back({left, j.first, mid - j.first});
    }
    return B;
}

int32_t segmentSurfaceHit(const TriSurface &s, const TriangleBvh *bvh, const double startA[3], const double endA[3], double hit[3])
{
    const V start = {startA[0], startA[1], startA[2]}, end = {endA[0], endA[1], endA[2]};
    const V dir = end - start;
    double best = 2.0;
    int32_t bestI = -1;
    auto consider = [&](int32_t i) {
        double t;
        if (triangleHit(s, i, start, dir, t) && (t < best || (t == best && i < bestI)))
        {
            best = t;
            bestI = i;
        }
    };
    if (!bvh)
        for (int32_t i = 0; i < (int32_t)s.nTris(); ++i)
            consider(i);
    else if (!bvh->right.empty())
    {
        const double o[3] = {start.x, start.y, start.z}, dv[3] = {dir.x, dir.y, dir.z};
        std::vector<int32_t> stack(1, 0);
        while (!stack.empty())
        {
            const int32_t node = stack.back();
            stack.pop_back();
            // slab test of the segment parameter range [0, min(1, best)] against the inflated box
            double t0 = 0.0, t1 = best < 1.0 ? best : 1.0;
            bool miss = false;
            for (int d = 0; d < 3 && !miss; ++d)
            {
                const double lo = bvh->box[6 * (size_t)node + d], hi = bvh->box[6 * (size_t)node + 3 + d];
                if (dv[d] == 0.0)
                    miss = o[d] < lo || o[d] > hi;
                else
                {
                    double a = (lo - o[d]) / dv[d], b = (hi - o[d]) / dv[d];
                    if (a > b)
                        std::swap(a, b);
                    // one ulp-scale slack on the parametric bounds keeps the test conservative
                    a -= 1e-12 * (1.0 + std::fabs(a));
                    b += 1e-12 * (1.0 + std::fabs(b));
                    t0 = a > t0 ? a : t0;
                    t1 = b < t1 ? b : t1;
                    miss = t0 > t1;
                }
            }
            if (miss)
                continue;
            if (bvh->right[node] < 0)
                for (int32_t k = bvh->first[node]; k < bvh->first[node] + bvh->count[node]; ++k)
                    consider(bvh->order[k]);
            else
            {
                stack.push_back(bvh->right[node]);
                stack.push_back(bvh->left[node]);
            }
        }
    }
    if (bestI < 0)
        return -1;
    const V p = start + best * dir;
    hit[0] = p.x, hit[1] = p.y, hit[2] = p.z;
    return bestI;
}

} // namespace sm
