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

BoundarySetup buildBoundarySetup(const PolyMesh &m, const Topology &t, const std::vector<double> &points, const EdgeMesh &initEdges,
                                 const EdgeMesh &targetEdgesIn, const TriSurface &surface, const std::vector<int32_t> &patchSmoothing,
                                 double layerEdgeLength, const std::vector<int32_t> &cornerIO, const std::vector<int32_t> &featureIO)
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
    B.distanceTolerance = 1e-4 * ((t.minEdgeLength < layerEdgeLength) ? t.minEdgeLength : layerEdgeLength); // :1921
    // getMeshStats' perimeter over the edge end points, src/smoothMesh.C:1495-1538
    double lo[3] = {SM_VGREAT, SM_VGREAT, SM_VGREAT}, hi[3] = {-SM_VGREAT, -SM_VGREAT, -SM_VGREAT};
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
    const double meshPerimeter = hi[0] - lo[0] + hi[1] - lo[1] + hi[2] + lo[2];
    checkEdgeMeshSanity(initEdges, t.minEdgeLength, meshPerimeter);
    checkEdgeMeshSanity(B.targetEdges, t.minEdgeLength, meshPerimeter);
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

} // namespace sm
