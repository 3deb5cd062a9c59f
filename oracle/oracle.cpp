// oracle.cpp -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the smoothMesh
// centroidal-smoothing iteration, used as the parity checker for the CUDA path
// and as the timed CPU baseline.  Nothing under smoothmesh_b200/ may call, link
// or import this file; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs do.
//
// PARITY: pinned against the reference's own code, with one caveat.  The reference (tkeskita/smoothMesh)
// needs OpenFOAM + wmake for a real build and ships no golden vectors, but its translation unit
// (src/smoothMesh.C and the two files it #includes) compiles UNMODIFIED, where it lies, against the
// minimal OpenFOAM facade in oracle/of_facade (oracle/_ref/smoothMesh_ref, recipe oracle/Makefile.ref).
// tests/test_reference_build.py runs that binary and this restatement on the same cases (hex, polyhedral,
// high aspect ratio, constraints on/off, boundary layer treatment, BASELINE configs 1 and 2): iteration
// counts, nFrozenPoints and final points agree bit for bit.  The caveat: what OpenFOAM itself computes
// (geometry formulas, connectivity row orders, VSMALL-tolerant vector equality, Foam::min/max, syncTools)
// is written from memory of OpenFOAM v2312/v12, marked [OF-recalled], and SHARED between this file and
// the facade -- the comparison verifies the reading of the reference's code, not those recalled semantics.
// The same binary run with -parallel (one forked process per processor directory, syncPointList /
// returnReduce over shared memory) is reproduced bit for bit by the rank emulation below.
// Further pins: the known-answer tests in tests/test_oracle_known_answers.py and an independent NumPy
// restatement (oracle/oracle_np.py).  Every function cites the reference lines it follows (paths relative
// to /root/reference).
//
// Build: see oracle/Makefile (g++ -O3 -ffp-contract=off -fopenmp).
//
// Layout: one `Rank` object per (emulated) MPI rank holding an OpenFOAM-style
// local mesh; a `Group` runs the iteration in lock-step over its ranks and
// performs the syncTools::syncPointList / returnReduce combinations between
// phases.  A serial run is a group of one rank with no shared points.

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stack>
#include <string>
#include <vector>

#ifdef ORACLE_LIBM_ACOS
// Literal reference behaviour (std::acos).  The default build uses the same
// sm_acos as the device code so that CPU and GPU agree bit for bit.
#define ORC_ACOS(x) std::acos(x)
#include "../smoothmesh_b200/csrc/sm_math.h"
#else
#include "../smoothmesh_b200/csrc/sm_math.h"
#define ORC_ACOS(x) sm_acos(x)
#endif

namespace
{

// ---------------------------------------------------------------- vector ----
// OpenFOAM Vector<double> semantics [OF-recalled]: component-wise + - ; s*v
// multiplies each component; v/s divides each component; a&b is the dot product
// summed x,y,z left to right; a^b the cross product; mag = sqrt(magSqr);
// operator== compares each component with mag(a-b) <= VSMALL.
struct V3
{
    double x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator/(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline V3 &operator+=(V3 &a, V3 b)
{
    a = a + b;
    return a;
}
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double magSqr(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
inline double mag(V3 a) { return std::sqrt(magSqr(a)); }
inline bool veq(V3 a, V3 b) { return sm_equal(a.x, b.x) && sm_equal(a.y, b.y) && sm_equal(a.z, b.z); }
inline double fmin_(double a, double b) { return (a < b) ? a : b; } // Foam::min [OF-recalled]
inline double fmax_(double a, double b) { return (a > b) ? a : b; } // Foam::max [OF-recalled]

const double GREAT = SM_GREAT, VSMALL = SM_VSMALL, VGREAT = SM_VGREAT;
const V3 UNDEF_VECTOR = {SM_GREAT, SM_GREAT, SM_GREAT}; // src/smoothMeshCommon.H:15
const V3 ZERO_VECTOR = {0, 0, 0};                       // src/smoothMeshCommon.H:16
const int UNDEF_LABEL = -1;                             // src/smoothMeshCommon.H:14

struct Params
{
    double minEdgeLength, maxStepLength, relStepFrac, minAngle, maxAngle, relTol;
    int32_t totalMinFreeze, edgeAngleConstraint, faceAngleConstraint, geometryVariant;
    // boundary layer treatment (src/smoothMesh.C:1892-1905); layerEdgeLength < 0 -> minEdgeLength
    double layerMaxBlendingFraction, layerEdgeLength, layerExpansionRatio;
    int32_t minLayers, maxLayers;
};

struct MeshIn
{
    int64_t P, C, F, Fi;
    const double *pts;
    const int32_t *fOff, *fV, *own, *nei;
    int32_t nPatches;
    const int32_t *pStart, *pSize, *pKind; // kind: 0 boundary, 1 processor, 2 empty
    const int64_t *pointGlobalId;          // may be null (serial)
    const int32_t *pLayer;                 // may be null: 1 = patch selected by -layerPatches
    const int32_t *pSmooth;                // may be null: 1 = patch selected by -smoothingPatches
};

thread_local std::string g_err;
std::string g_last_error;

struct Rank
{
    int P = 0, C = 0, F = 0, Fi = 0;
    std::vector<V3> pts;
    std::vector<int> fOff, fV, own, nei;
    std::vector<int> pStart, pSize, pKind, pLayer;
    std::vector<int64_t> gid;
    // boundary layer treatment state (src/orthogonalBoundaryBlending.C)
    bool doLayerTreatment = false;
    std::vector<uint8_t> isConnectedToInternal, isLayerSurface, isOuterNeighInProc, isSharpEdge;
    std::vector<int> hopsToLayer, pointToOuter, newHopCounts, nBoundaryFaces, boundaryPointLabels;
    std::vector<V3> pointNormals, outerNeighCoords;
    std::vector<double> layerLength, layerBlend; // per hop count
    std::vector<V3> snapNormals, snapLayerBlend;
    // boundary point smoothing state (src/boundaryPointSmoothing.C), SURVEY 8(f)-4
    struct EdgeMesh
    { // [OF-recalled] edgeMesh read from an OBJ file: points, edges, pointEdges in ascending edge order
        std::vector<V3> points;
        std::vector<std::array<int, 2>> edges;
        std::vector<std::vector<int>> pointEdges;
        void finish()
        {
            pointEdges.assign(points.size(), {});
            for (size_t e = 0; e < edges.size(); ++e)
            {
                pointEdges[edges[e][0]].push_back((int)e);
                pointEdges[edges[e][1]].push_back((int)e);
            }
        }
    };
    struct TriSurface
    {
        std::vector<V3> points;
        std::vector<std::array<int, 3>> tris;
    };
    bool haveGeometry = false, doBoundarySmoothing = false;
    EdgeMesh initEdges, targetEdges;
    TriSurface surf;
    std::vector<int> pSmooth;                   // patch selected by -smoothingPatches
    std::vector<int> cornerIO, featureIO;       // isCornerPoint / isFeatureEdgePoint lists of an earlier run (may be empty)
    double internalSmoothingBlendingFraction = 0.0;
    double distanceTolerance = 0.0, meshPerimeter = 0.0;
    V3 bbMin = {0, 0, 0}, bbMax = {0, 0, 0};
    std::vector<uint8_t> isFeatureEdge, isCorner, isSmoothingSurface, isInnerNeighInProc;
    std::vector<V3> cornerPoints, innerNeighCoords, featureEdgeProjections;
    std::vector<int> targetEdgeStrings, pointStrings, hopsToSmoothing, pointToInner, nFeatureEdgeProjections,
        nFaceCentroids;
    Params prm;

    // derived connectivity [OF-recalled row orders, SURVEY A.2]
    std::vector<std::vector<int>> pointFaces, pointCells, pointEdges, pointPoints, edgeFaces, edgeCells, cellFaces,
        cellPoints, pointNeighPoints;
    std::vector<std::array<int, 2>> edges;
    std::vector<uint8_t> isInternal;
    double meshMinEdge = 0, meshMaxEdge = 0;

    // geometry of the current mesh (OpenFOAM demand-driven data, cleared by movePoints)
    std::vector<V3> faceCtr, faceArea, cellCtr;
    bool geomValid = false;

    // per-iteration state
    std::vector<V3> sumC;          // centroidalSmoothing: cellPoints (sum of cell centres)
    std::vector<int64_t> nC;       // centroidalSmoothing: nPoints
    std::vector<V3> centroidal, newPts;
    std::vector<V3> cp1, cp2, cp3;
    std::vector<uint8_t> hasCommon, frozen;
    int64_t nFrozen = 0;
    double residual = 0;
    std::string err;

    // snapshots of the last iteration (for stage-level parity tests)
    std::vector<V3> snapCellCtr, snapCentroidal, snapBlend, snapClamped;
    std::vector<uint8_t> snapFrozenEdgeLen, snapFrozenEdgeAngle, snapFrozenFaceAngle;
    std::vector<double> snapCurMin, snapCurMax;

    int faceSize(int f) const { return fOff[f + 1] - fOff[f]; }
    const int *faceBegin(int f) const { return &fV[fOff[f]]; }

    bool init(const MeshIn &m, const Params &p);
    void buildConnectivity();
    bool buildCellFaces();
    void findInternalMeshPoints();
    void getMeshStats();
    void calcGeometry();
    void centroidalPartial();
    void centroidalFinish();
    bool findClosestLocal();
    void aspectRatioBlend();
    void constrainMaxStepLength();
    void restrictEdgeShortening();
    void restrictMinEdgeAngleDecrease();
    bool restrictFaceAngleDeterioration();
    void classifyBoundaryPoints();
    void layersBegin();
    bool boundaryBegin();
    void boundaryPointStrings();
    void innerNeighInfo();
    void featureEdgeProjectionsLocal();
    bool surfaceCentroidsLocal();
    bool projectBoundaryPoints();
    void updateInnerNeighCoordsLocal();
    bool projectPrismaticInternalPoints();
    void hopsInit();
    void hopsSweep();
    void hopsInitFor(std::vector<int> &hops, const std::vector<int> &patchFlags);
    void hopsSweepFor(std::vector<int> &hops);
    void boundaryNormalsLocal();
    void boundaryNormalsFinish();
    void outerInit();
    void outerSweep(int iter);
    void outerUndo();
    void layerTables();
    void updateNeighCoordsLocal();
    bool blendWithOrthogonalPoints();
    bool calcMinMaxFaceAngleForEdge(int edgeI, double &mn, double &mx, int pI1, V3 c1, int pI2, V3 c2);
    bool calcMinMaxFaceAngleForPoint(int pI1, V3 c1, int pI2, V3 c2, double &mn, double &mx);
    void restoreAndResidual();
    void movePoints();
};

// ------------------------------------------------------------------ setup ----
bool Rank::init(const MeshIn &m, const Params &p)
{
    P = (int)m.P;
    C = (int)m.C;
    F = (int)m.F;
    Fi = (int)m.Fi;
    prm = p;
    pts.resize(P);
    for (int i = 0; i < P; ++i)
        pts[i] = {m.pts[3 * i], m.pts[3 * i + 1], m.pts[3 * i + 2]};
    fOff.assign(m.fOff, m.fOff + F + 1);
    fV.assign(m.fV, m.fV + fOff[F]);
    own.assign(m.own, m.own + F);
    nei.assign(m.nei, m.nei + Fi);
    pStart.assign(m.pStart, m.pStart + m.nPatches);
    pSize.assign(m.pSize, m.pSize + m.nPatches);
    pKind.assign(m.pKind, m.pKind + m.nPatches);
    pLayer.assign(m.nPatches, 0);
    if (m.pLayer)
        pLayer.assign(m.pLayer, m.pLayer + m.nPatches);
    pSmooth.assign(m.nPatches, 0);
    if (m.pSmooth)
        pSmooth.assign(m.pSmooth, m.pSmooth + m.nPatches);
    if (m.pointGlobalId)
        gid.assign(m.pointGlobalId, m.pointGlobalId + P);
    // src/smoothMesh.C:61-66: empty patches are fatal
    for (int k : pKind)
        if (k == 2)
        {
            err = "Smoothing of non-3D meshes (meshes with type empty patches) is not supported";
            return false;
        }
    buildConnectivity();
    if (!buildCellFaces())
        return false;
    findInternalMeshPoints();
    getMeshStats();
    frozen.assign(P, 0);
    return true;
}

// ------------------------------------------------- boundary layer treatment ----
// ------------------------------------------------ boundary point smoothing ----
// SURVEY 8(f)-4, src/boundaryPointSmoothing.C.  Restated for the oracle only (the CUDA path does not
// have it yet); checked against the reference's own translation unit (oracle/_ref) on testcase4 as shipped
// and on synthetic cases (tests/test_reference_build.py).

// indexedOctree::findLine stand-in [OF-recalled: the intersection of the segment with the surface that is
// nearest to `start`]: every triangle is tested with the Moeller-Trumbore segment test, the smallest
// parameter wins, the lower triangle on ties.  The same definition (same operation order) is what the
// OpenFOAM facade of oracle/_ref uses; OpenFOAM's own triangle::intersection differs in rounding and in
// its edge tolerances.
static bool segmentSurfaceHit(const Rank::TriSurface &s, V3 start, V3 end, V3 &hitPoint)
{
    const V3 dir = end - start;
    double best = 2.0;
    int bestI = -1;
    for (size_t i = 0; i < s.tris.size(); ++i)
    {
        const V3 p0 = s.points[s.tris[i][0]], p1 = s.points[s.tris[i][1]], p2 = s.points[s.tris[i][2]];
        const V3 e1 = p1 - p0, e2 = p2 - p0;
        const V3 h = cross(dir, e2);
        const double det = dot(e1, h);
        if (std::fabs(det) < VSMALL)
            continue;
        const double inv = 1.0 / det;
        const V3 sv = start - p0;
        const double u = inv * dot(sv, h);
        if (u < 0.0 || u > 1.0)
            continue;
        const V3 q = cross(sv, e1);
        const double v = inv * dot(dir, q);
        if (v < 0.0 || u + v > 1.0)
            continue;
        const double t = inv * dot(e2, q);
        if (t < 0.0 || t > 1.0)
            continue;
        if (t < best)
        {
            best = t;
            bestI = (int)i;
        }
    }
    if (bestI < 0)
        return false;
    hitPoint = start + best * dir;
    return true;
}

// checkEdgeMeshSanity, src/boundaryPointSmoothing.C:20-82.  Both perimeters use the reference's formula
// (max z PLUS min z, :76 and src/smoothMesh.C:1538).  The test at :77 calls unqualified abs() on a double:
// in a plain GCC build (<cmath> / <cstdlib> only) that is ::abs(int), i.e. the ratio is truncated towards
// zero first, so the check fires only for |ratio - 1| >= 1; the shipped testcase4 (ratio - 1 = 0.744)
// relies on it.  Restated with that truncation.
static bool checkEdgeMeshSanity(const Rank::EdgeMesh &em, double meshMinEdgeLength, double meshPerimeter, std::string &err)
{
    double minEdgeLength = VGREAT;
    double bbMinX = VGREAT, bbMaxX = -VGREAT, bbMinY = VGREAT, bbMaxY = -VGREAT, bbMinZ = VGREAT, bbMaxZ = -VGREAT;
    for (auto &e : em.edges)
    {
        const V3 startPoint = em.points[e[0]], endPoint = em.points[e[1]];
        const double edgeLength = mag(endPoint - startPoint);
        if (edgeLength < minEdgeLength)
            minEdgeLength = edgeLength;
        for (const V3 &q : {startPoint, endPoint})
        {
            if (q.x < bbMinX)
                bbMinX = q.x;
            if (q.y < bbMinY)
                bbMinY = q.y;
            if (q.z < bbMinZ)
                bbMinZ = q.z;
            if (q.x > bbMaxX)
                bbMaxX = q.x;
            if (q.y > bbMaxY)
                bbMaxY = q.y;
            if (q.z > bbMaxZ)
                bbMaxZ = q.z;
        }
    }
    if (minEdgeLength < 1e-4 * meshMinEdgeLength) // REL_TOL
    {
        err = "Minimum edge length in edge mesh is too small in comparison to minimum edge length in polyMesh";
        return false;
    }
    const double emPerimeter = bbMaxX - bbMinX + bbMaxY - bbMinY + bbMaxZ + bbMinZ;
    const double ratio = (emPerimeter / meshPerimeter) - 1.0;
    if (std::abs((int)ratio) > 0.5) // ::abs(int), see above
    {
        err = "Perimeter (sum of bounding box side lengths) of edge mesh is too different in comparison to perimeter of polyMesh";
        return false;
    }
    return true;
}

// projectPointToEdge, :89-145
static void projectPointToEdge(V3 pt, const Rank::EdgeMesh &em, int edgeI, double distanceTolerance, V3 &projPoint, int &edgePointI)
{
    edgePointI = UNDEF_LABEL;
    const int startPointI = em.edges[edgeI][0], endPointI = em.edges[edgeI][1];
    const V3 startPoint = em.points[startPointI], endPoint = em.points[endPointI];
    const double edgeLength = mag(endPoint - startPoint);
    const V3 c2pt = pt - startPoint;
    const V3 edgeVec = endPoint - startPoint;
    const double normalizedDotProd = dot(c2pt, edgeVec) / (edgeLength * edgeLength);
    const V3 testProjPoint = startPoint + normalizedDotProd * edgeVec;
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

// findClosestEdgeMeshCornerPointIndex, :151-186; -1 where the reference aborts
static int findClosestEdgeMeshCornerPointIndex(V3 pt, const Rank::EdgeMesh &em)
{
    double distance = GREAT;
    int closestPointI = UNDEF_LABEL;
    for (size_t pointI = 0; pointI < em.points.size(); ++pointI)
    {
        if (em.pointEdges[pointI].size() == 2)
            continue;
        const double testDistance = mag(pt - em.points[pointI]);
        if (testDistance < distance)
        {
            distance = testDistance;
            closestPointI = (int)pointI;
        }
    }
    return closestPointI;
}

// findClosestEdgeInfo, :206-263; false where the reference aborts
static bool findClosestEdgeInfo(V3 pt, const Rank::EdgeMesh &em, int requiredStringI, const std::vector<int> &targetEdgeStrings,
                                double distanceTolerance, V3 &projPoint, int &closestEdgeI, int &closestEdgeStringI,
                                int &closestEdgePointI)
{
    double distance = GREAT;
    projPoint = UNDEF_VECTOR;
    closestEdgeI = closestEdgeStringI = closestEdgePointI = UNDEF_LABEL;
    for (size_t edgeI = 0; edgeI < em.edges.size(); ++edgeI)
    {
        if (requiredStringI >= 0 && targetEdgeStrings[edgeI] != requiredStringI)
            continue;
        V3 testProjPoint;
        int edgePointI;
        projectPointToEdge(pt, em, (int)edgeI, distanceTolerance, testProjPoint, edgePointI);
        const double testDistance = mag(testProjPoint - pt);
        if (testDistance < distance)
        {
            distance = testDistance;
            projPoint = testProjPoint;
            closestEdgeI = (int)edgeI;
            closestEdgePointI = edgePointI;
            if (em.edges.size() == targetEdgeStrings.size())
                closestEdgeStringI = targetEdgeStrings[edgeI];
        }
    }
    return !(requiredStringI >= 0 && closestEdgeStringI == UNDEF_LABEL);
}

// findContinuousEdgeMeshEdges (:446-486), stringifyEdgeMeshEdges (:492-551), findEdgeMeshStrings (:557-590).
// Note :521 / :529: the "neighbour is not a corner" test indexes pointEdges with an EDGE label; restated
// literally (an edge label beyond the point list reads as "not 2", i.e. no recursion).
static void findContinuousEdgeMeshEdges(const Rank::EdgeMesh &em, int edgeI, int &neighEdgeI1, int &neighEdgeI2)
{
    neighEdgeI1 = neighEdgeI2 = UNDEF_LABEL;
    const int pointI1 = em.edges[edgeI][0];
    if (em.pointEdges[pointI1].size() == 2)
    {
        int edgeI1 = em.pointEdges[pointI1][0];
        if (edgeI1 == edgeI)
            edgeI1 = em.pointEdges[pointI1][1];
        neighEdgeI1 = edgeI1;
    }
    const int pointI2 = em.edges[edgeI][1];
    if (em.pointEdges[pointI2].size() == 2)
    {
        int edgeI2 = em.pointEdges[pointI2][0];
        if (edgeI2 == edgeI)
            edgeI2 = em.pointEdges[pointI2][1];
        neighEdgeI2 = edgeI2;
    }
}
static void stringifyEdgeMeshEdges(const Rank::EdgeMesh &em, std::vector<int> &targetEdgeStrings, int edgeI, int neighEdgeI1,
                                   int neighEdgeI2, int &nStrings)
{
    const int stringI0 = targetEdgeStrings[edgeI];
    const int stringI1 = (neighEdgeI1 != UNDEF_LABEL) ? targetEdgeStrings[neighEdgeI1] : UNDEF_LABEL;
    const int stringI2 = (neighEdgeI2 != UNDEF_LABEL) ? targetEdgeStrings[neighEdgeI2] : UNDEF_LABEL;
    const int maxStringI = std::max(std::max(stringI0, stringI1), stringI2);
    if (maxStringI == UNDEF_LABEL)
    {
        ++nStrings;
        targetEdgeStrings[edgeI] = nStrings;
    }
    else if (stringI0 == UNDEF_LABEL)
        targetEdgeStrings[edgeI] = maxStringI;
    auto twoEdgesAt = [&](int label) { return label < (int)em.pointEdges.size() && em.pointEdges[label].size() == 2; };
    if (neighEdgeI1 != UNDEF_LABEL && stringI1 == UNDEF_LABEL && twoEdgesAt(neighEdgeI1))
    {
        int nn1, nn2;
        findContinuousEdgeMeshEdges(em, neighEdgeI1, nn1, nn2);
        stringifyEdgeMeshEdges(em, targetEdgeStrings, neighEdgeI1, nn1, nn2, nStrings);
    }
    if (neighEdgeI2 != UNDEF_LABEL && stringI2 == UNDEF_LABEL && twoEdgesAt(neighEdgeI2))
    {
        int nn1, nn2;
        findContinuousEdgeMeshEdges(em, neighEdgeI2, nn1, nn2);
        stringifyEdgeMeshEdges(em, targetEdgeStrings, neighEdgeI2, nn1, nn2, nStrings);
    }
}
static int findEdgeMeshStrings(std::vector<int> &targetEdgeStrings, const Rank::EdgeMesh &em)
{
    int nStrings = UNDEF_LABEL;
    targetEdgeStrings.assign(em.edges.size(), UNDEF_LABEL);
    for (size_t edgeI = 0; edgeI < em.edges.size(); ++edgeI)
    {
        if (targetEdgeStrings[edgeI] >= 0)
            continue;
        int n1, n2;
        findContinuousEdgeMeshEdges(em, (int)edgeI, n1, n2);
        stringifyEdgeMeshEdges(em, targetEdgeStrings, (int)edgeI, n1, n2, nStrings);
    }
    return nStrings;
}

// src/smoothMesh.C:2080-2098, :2131-2171: is boundary point smoothing on, and its one-time inputs.
// (The isCornerPoint / isFeatureEdgePoint label lists of :2039-2065 are not restated: classification
// always comes from the edge meshes.)
bool Rank::boundaryBegin()
{
    bool anySmoothingPatch = false;
    for (int f : pSmooth)
        anySmoothingPatch = anySmoothingPatch || f;
    doBoundarySmoothing = haveGeometry && anySmoothingPatch;
    const double layerEdgeLength = prm.layerEdgeLength < 0 ? prm.minEdgeLength : prm.layerEdgeLength;
    distanceTolerance = 1e-4 * fmin_(meshMinEdge, layerEdgeLength); // REL_TOL, :1921
    if (!doBoundarySmoothing)
        return true;
    if (!checkEdgeMeshSanity(initEdges, meshMinEdge, meshPerimeter, err) ||
        !checkEdgeMeshSanity(targetEdges, meshMinEdge, meshPerimeter, err))
        return false;
    findEdgeMeshStrings(targetEdgeStrings, targetEdges);
    return true;
}

// classifyBoundaryPoints, src/boundaryPointSmoothing.C:269-440: every boundary point is classified once,
// by the first patch (in patch order) that contains it.  false where the reference aborts.
void Rank::classifyBoundaryPoints()
{
    isConnectedToInternal.assign(P, 0);
    isLayerSurface.assign(P, 0);
    isFeatureEdge.assign(P, 0);
    isCorner.assign(P, 0);
    isSmoothingSurface.assign(P, 0);
    cornerPoints.assign(P, UNDEF_VECTOR);
    std::vector<uint8_t> visited(P, 0);
    for (size_t patchI = 0; patchI < pKind.size(); ++patchI)
        for (int f = pStart[patchI]; f < pStart[patchI] + pSize[patchI]; ++f)
            for (int k = fOff[f]; k < fOff[f + 1]; ++k)
            {
                const int pointI = fV[k];
                if (visited[pointI])
                    continue;
                visited[pointI] = 1;
                if (isInternal[pointI])
                    continue;
                for (int i : pointPoints[pointI])
                    if (isInternal[i])
                        isConnectedToInternal[pointI] = 1;
                if (!initEdges.points.empty() && !targetEdges.points.empty())
                {
                    const V3 pt = pts[pointI];
                    // src/smoothMesh.C:2067-2078 + :336-340: classes from the label lists of an earlier run
                    bool labelIOListsHaveData = false;
                    for (int v : cornerIO)
                        labelIOListsHaveData = labelIOListsHaveData || v == 1;
                    for (int v : featureIO)
                        labelIOListsHaveData = labelIOListsHaveData || v == 1;
                    if (labelIOListsHaveData)
                    {
                        isCorner[pointI] = (pointI < (int)cornerIO.size() && cornerIO[pointI] == 1) ? 1 : 0;
                        isFeatureEdge[pointI] = (pointI < (int)featureIO.size() && featureIO[pointI] == 1) ? 1 : 0;
                    }
                    else
                    {
                        V3 projPoint;
                        int dummy, dummy2, closestEdgePointI = UNDEF_LABEL;
                        findClosestEdgeInfo(pt, initEdges, -1, targetEdgeStrings, distanceTolerance, projPoint, dummy, dummy2,
                                            closestEdgePointI);
                        if (closestEdgePointI >= 0 && initEdges.pointEdges[closestEdgePointI].size() != 2)
                            isCorner[pointI] = 1;
                        else if (mag(pt - projPoint) < distanceTolerance)
                            isFeatureEdge[pointI] = 1;
                    }
                    if (isCorner[pointI])
                    {
                        const int c = findClosestEdgeMeshCornerPointIndex(pt, targetEdges);
                        if (c < 0)
                            err = "Did not find any eligible corner points in edge mesh";
                        else
                            cornerPoints[pointI] = targetEdges.points[c];
                    }
                }
                if (pLayer[patchI])
                    isLayerSurface[pointI] = 1;
                if (doBoundarySmoothing && pSmooth[patchI])
                    isSmoothingSurface[pointI] = 1;
            }
}

// src/smoothMesh.C:2234-2250: the target edge string every feature edge point snaps to
void Rank::boundaryPointStrings()
{
    pointStrings.assign(P, UNDEF_LABEL);
    for (int pointI = 0; pointI < P; ++pointI)
    {
        if (!isFeatureEdge[pointI])
            continue;
        V3 dummyPoint;
        int dummy, dummy2, pointStringI = UNDEF_LABEL;
        findClosestEdgeInfo(pts[pointI], targetEdges, -1, targetEdgeStrings, distanceTolerance, dummyPoint, dummy, pointStringI,
                            dummy2);
        pointStrings[pointI] = pointStringI;
    }
}

// propagateInnerNeighInfo, src/orthogonalBoundaryBlending.C:397-458
void Rank::innerNeighInfo()
{
    isInnerNeighInProc.assign(P, 0);
    pointToInner.assign(P, UNDEF_LABEL);
    for (int pointI = 0; pointI < P; ++pointI)
    {
        if (!isSmoothingSurface[pointI] || !isConnectedToInternal[pointI])
            continue;
        const int nHops = hopsToSmoothing[pointI];
        if (nHops != 0)
        {
            err = std::to_string(pointI) + " is not boundary point";
            continue;
        }
        int nNeighHops = 0, neighPointI = UNDEF_LABEL;
        for (int neighI : pointPoints[pointI])
            if (hopsToSmoothing[neighI] == nHops + 1)
            {
                ++nNeighHops;
                neighPointI = neighI;
            }
        if (nNeighHops == 1)
        {
            isInnerNeighInProc[pointI] = 1;
            pointToInner[pointI] = neighPointI;
        }
    }
}

// calculateFeatureEdgeProjections before its synchronisations, src/boundaryPointSmoothing.C:623-656
void Rank::featureEdgeProjectionsLocal()
{
    featureEdgeProjections.assign(P, ZERO_VECTOR);
    nFeatureEdgeProjections.assign(P, 0);
    for (int pointI = 0; pointI < P; ++pointI)
    {
        if (!isFeatureEdge[pointI])
            continue;
        for (int neighI : pointPoints[pointI])
        { // findNeighborSurfacePoints, :592-616
            if (isInternal[neighI] || isFeatureEdge[neighI] || isCorner[neighI])
                continue;
            V3 projPoint;
            int dummy, dummy2, dummy3;
            if (!findClosestEdgeInfo(pts[neighI], targetEdges, pointStrings[pointI], targetEdgeStrings, distanceTolerance, projPoint,
                                     dummy, dummy2, dummy3))
                err = "Internal sanity check failed: Did not find any edges with string index " +
                      std::to_string(pointStrings[pointI]);
            featureEdgeProjections[pointI] += projPoint;
            ++nFeatureEdgeProjections[pointI];
        }
    }
}

// calculateSurfaceCentroids, :781-838: its sums only feed a term multiplied by
// faceCentroidBlendingFraction = 0.0 (:869), so only its FatalError is observable
bool Rank::surfaceCentroidsLocal()
{
    nFaceCentroids.assign(P, 0);
    const int firstBoundaryFaceI = Fi;
    for (int pointI = 0; pointI < P; ++pointI)
    {
        if (isInternal[pointI])
            continue;
        for (int faceI : pointFaces[pointI])
            if (faceI >= firstBoundaryFaceI)
                ++nFaceCentroids[pointI];
        if (nFaceCentroids[pointI] == 0)
        {
            err = "did not find faceNeighbour for point " + std::to_string(pointI);
            return false;
        }
    }
    return true;
}

// projectBoundaryPointsToEdgesAndSurfaces after the synchronisations, :871-944, with findIntersection (:682-744)
bool Rank::projectBoundaryPoints()
{
    for (int pointI = 0; pointI < P; ++pointI)
    {
        if (isInternal[pointI])
            continue;
        if (isCorner[pointI])
        {
            newPts[pointI] = cornerPoints[pointI];
            continue;
        }
        if (isFeatureEdge[pointI])
        {
            newPts[pointI] = featureEdgeProjections[pointI] / double(nFeatureEdgeProjections[pointI]);
            continue;
        }
        if (isSharpEdge[pointI])
            frozen[pointI] = 1;
        else if (isSmoothingSurface[pointI])
        {
            const V3 pointNormal = pointNormals[pointI];
            double searchDistance = distanceTolerance;
            // faceCentroidBlendingFraction = 0.0: 0 * centroid + 1 * newPoint
            const V3 newPoint = 0.0 * ZERO_VECTOR + (1 - 0.0) * newPts[pointI];
            V3 surfPoint = UNDEF_VECTOR;
            for (int i = 0; i < 4; ++i)
            {
                searchDistance *= (1.0 / 1e-4); // 1 / REL_TOL
                if (veq(pointNormal, ZERO_VECTOR))
                {
                    err = "pointNormal is zero for pointI " + std::to_string(pointI);
                    return false;
                }
                V3 hitPoint1 = UNDEF_VECTOR, hitPoint2 = UNDEF_VECTOR, h;
                if (segmentSurfaceHit(surf, newPoint, newPoint + searchDistance * pointNormal, h))
                    hitPoint1 = h;
                if (segmentSurfaceHit(surf, newPoint, newPoint - searchDistance * pointNormal, h))
                    hitPoint2 = h;
                const double distance1 = mag(newPoint - hitPoint1), distance2 = mag(newPoint - hitPoint2);
                if (distance1 < distance2)
                    surfPoint = hitPoint1;
                else if (distance2 < distance1)
                    surfPoint = hitPoint2;
                else if (segmentSurfaceHit(surf, newPoint + searchDistance * pointNormal, newPoint - searchDistance * pointNormal, h))
                    surfPoint = h;
                else
                    surfPoint = UNDEF_VECTOR;
                if (!veq(surfPoint, UNDEF_VECTOR))
                {
                    newPts[pointI] = surfPoint;
                    break;
                }
            }
            if (veq(surfPoint, UNDEF_VECTOR))
            {
                err = "Did not find surface intersection for pointI " + std::to_string(pointI);
                return false;
            }
        }
    }
    return true;
}

// updateNeighCoords for the inner map, local part (src/orthogonalBoundaryBlending.C:472-487)
void Rank::updateInnerNeighCoordsLocal()
{
    innerNeighCoords.assign(P, UNDEF_VECTOR);
    for (int pointI = 0; pointI < P; ++pointI)
        if (isInnerNeighInProc[pointI])
            innerNeighCoords[pointI] = pts[pointToInner[pointI]];
}

// projectPrismaticInternalPointsToSurfaces, src/orthogonalBoundaryBlending.C:573-632
bool Rank::projectPrismaticInternalPoints()
{
    for (int pointI = 0; pointI < P; ++pointI)
    {
        if (!isSmoothingSurface[pointI] || !isConnectedToInternal[pointI] || pointToInner[pointI] < 0 || isFeatureEdge[pointI] ||
            isCorner[pointI] || isSharpEdge[pointI])
            continue;
        const V3 pointNormal = pointNormals[pointI];
        const V3 innerNeighCoord = innerNeighCoords[pointI];
        if (hopsToSmoothing[pointI] != 0 || veq(pointNormal, ZERO_VECTOR) || veq(innerNeighCoord, UNDEF_VECTOR))
        {
            err = "Point " + std::to_string(pointI) + " fails the sanity checks of projectPrismaticInternalPointsToSurfaces";
            return false;
        }
        const V3 cCoords = newPts[pointI];
        const V3 neighVec = cCoords - innerNeighCoord;
        const double dotProd = dot(neighVec, pointNormal);
        const V3 pVec = neighVec - dotProd * pointNormal;
        const V3 newCoords = cCoords - pVec;
        newPts[pointI] = internalSmoothingBlendingFraction * newCoords + (1 - internalSmoothingBlendingFraction) * newPts[pointI];
    }
    return true;
}

// calculatePointHopsToBoundary, src/orthogonalBoundaryBlending.C:52-134, in two parts: the seeding
// (:64-77) and one propagation sweep (:86-120); the caller synchronises after every sweep (:124-130).
void Rank::hopsInitFor(std::vector<int> &hops, const std::vector<int> &patchFlags)
{
    hops.assign(P, UNDEF_LABEL);
    for (size_t patchI = 0; patchI < pKind.size(); ++patchI)
    {
        if (!patchFlags[patchI])
            continue;
        for (int f = pStart[patchI]; f < pStart[patchI] + pSize[patchI]; ++f)
            for (int k = fOff[f]; k < fOff[f + 1]; ++k)
                if (isConnectedToInternal[fV[k]])
                    hops[fV[k]] = 0;
    }
    newHopCounts.assign(P, -1);
}
void Rank::hopsSweepFor(std::vector<int> &hops)
{
    for (int pointI = 0; pointI < P; ++pointI)
    {
        if (hops[pointI] >= 0 || !isInternal[pointI])
            continue;
        int maxHops = -1;
        for (int neighI : pointPoints[pointI])
            if (hops[neighI] > maxHops)
                maxHops = hops[neighI];
        if (maxHops >= 0)
            newHopCounts[pointI] = maxHops + 1;
    }
    for (int pointI = 0; pointI < P; ++pointI)
        if (newHopCounts[pointI] > hops[pointI])
            hops[pointI] = newHopCounts[pointI];
}
void Rank::hopsInit() { hopsInitFor(hopsToLayer, pLayer); }
void Rank::hopsSweep() { hopsSweepFor(hopsToLayer); }

// calculateBoundaryPointNormals, src/orthogonalBoundaryBlending.C:141-233, in two parts around the
// two sum-synchronisations at :185-198.  Note that it accumulates onto the normals of the previous
// call (no zeroing at :178; a shared point therefore also sums the previous normal of every copy)
// and re-normalises every non-zero normal, internal points included (:224-230).
void Rank::boundaryNormalsLocal()
{
    calcGeometry(); // patch.Sf() / patch.magSf() of the current mesh, :171-172
    nBoundaryFaces.assign(P, 0);
    for (size_t patchI = 0; patchI < pKind.size(); ++patchI)
    {
        if (pKind[patchI] != 0)
            continue;
        for (int f = pStart[patchI]; f < pStart[patchI] + pSize[patchI]; ++f)
        {
            const V3 Sf = faceArea[f] / mag(faceArea[f]);
            for (int k = fOff[f]; k < fOff[f + 1]; ++k)
            {
                pointNormals[fV[k]] = pointNormals[fV[k]] - Sf;
                ++nBoundaryFaces[fV[k]];
            }
        }
    }
}
void Rank::boundaryNormalsFinish()
{
    for (int pointI = 0; pointI < P; ++pointI)
    {
        if (nBoundaryFaces[pointI] < 1)
            continue;
        if (mag(pointNormals[pointI]) < 0.1)
        {
            pointNormals[pointI] = ZERO_VECTOR;
            isSharpEdge[pointI] = 1;
        }
        else
            isSharpEdge[pointI] = 0;
    }
    for (int pointI = 0; pointI < P; ++pointI)
        if (!veq(pointNormals[pointI], ZERO_VECTOR))
            pointNormals[pointI] = pointNormals[pointI] / mag(pointNormals[pointI]);
}

// propagateOuterNeighInfo, src/orthogonalBoundaryBlending.C:244-391: initialisation, one sweep per
// hop count (the caller synchronises the normals with maxMagSqr after every sweep, :363-369), undo.
void Rank::outerInit()
{
    isOuterNeighInProc.assign(P, 0);
    pointToOuter.assign(P, UNDEF_LABEL);
    boundaryPointLabels.assign(P, UNDEF_LABEL);
}
void Rank::outerSweep(int iter)
{
    for (int pointI = 0; pointI < P; ++pointI)
    {
        const int nHops = hopsToLayer[pointI];
        if (nHops != iter)
            continue;
        int nNeighHops = 0, neighPointI = UNDEF_LABEL;
        for (int neighI : pointPoints[pointI])
            if (hopsToLayer[neighI] == nHops - 1)
            {
                ++nNeighHops;
                neighPointI = neighI;
            }
        if (nNeighHops != 1)
            continue;
        if (!isInternal[neighPointI] && !isLayerSurface[neighPointI])
            continue;
        const auto it = std::find(boundaryPointLabels.begin(), boundaryPointLabels.end(), neighPointI);
        if (it != boundaryPointLabels.end())
        {
            pointNormals[pointI] = UNDEF_VECTOR;
            pointNormals[it - boundaryPointLabels.begin()] = UNDEF_VECTOR;
            continue;
        }
        isOuterNeighInProc[pointI] = 1;
        pointToOuter[pointI] = neighPointI;
        pointNormals[pointI] = pointNormals[neighPointI];
        boundaryPointLabels[pointI] = neighPointI;
    }
}
void Rank::outerUndo()
{
    for (int pointI = 0; pointI < P; ++pointI)
        if (veq(pointNormals[pointI], UNDEF_VECTOR))
        {
            pointNormals[pointI] = ZERO_VECTOR;
            isOuterNeighInProc[pointI] = 0;
            pointToOuter[pointI] = UNDEF_LABEL;
        }
}

// src/smoothMesh.C:2024-2033 and the local, communication-free pieces of :2190-2221; Group::setupLayers
// drives the phases and the synchronisations between them.
void Rank::layersBegin()
{
    bool anyLayerPatch = false;
    for (int f : pLayer)
        anyLayerPatch = anyLayerPatch || f;
    doLayerTreatment = anyLayerPatch && prm.layerMaxBlendingFraction > SM_SMALL;
    pointNormals.assign(P, ZERO_VECTOR);
    isSharpEdge.assign(P, 0);
}
// per-hop constants of blendWithOrthogonalPoints (:547-555); maxLayers there is maxLayers + 1 (:2300)
void Rank::layerTables()
{
    const double maxLayers = prm.maxLayers + 1, minLayers = prm.minLayers;
    const double layerEdgeLength = prm.layerEdgeLength < 0 ? prm.minEdgeLength : prm.layerEdgeLength;
    int maxHopSeen = 0;
    for (int h : hopsToLayer)
        maxHopSeen = std::max(maxHopSeen, h);
    layerLength.assign(maxHopSeen + 2, 0.0);
    layerBlend.assign(maxHopSeen + 2, 0.0);
    for (int nHops = 1; nHops <= maxHopSeen + 1; ++nHops)
    {
        const int maxHops = (int)fmin_(double(nHops - 1), maxLayers);
        layerLength[nHops] = layerEdgeLength * std::pow(prm.layerExpansionRatio, (double)maxHops);
        const double slope = -prm.layerMaxBlendingFraction / (maxLayers - minLayers);
        const double y0 = -slope * maxLayers;
        const double y = y0 + slope * nHops;
        layerBlend[nHops] = fmax_(0.0, fmin_(y, prm.layerMaxBlendingFraction));
    }
}

// updateNeighCoords, local part (:472-487); the caller synchronises with minMagSqr (:491-497)
void Rank::updateNeighCoordsLocal()
{
    outerNeighCoords.assign(P, UNDEF_VECTOR);
    for (int pointI = 0; pointI < P; ++pointI)
        if (isOuterNeighInProc[pointI])
            outerNeighCoords[pointI] = pts[pointToOuter[pointI]];
}

// blendWithOrthogonalPoints (:507-567)
bool Rank::blendWithOrthogonalPoints()
{
    for (int pointI = 0; pointI < P; ++pointI)
    {
        if (veq(pointNormals[pointI], ZERO_VECTOR) || !isInternal[pointI])
            continue;
        const int nHops = hopsToLayer[pointI];
        if (nHops < 1)
            continue;
        if (veq(outerNeighCoords[pointI], UNDEF_VECTOR))
        {
            err = "Sanity broken, outerNeighCoord is zero for pointI " + std::to_string(pointI);
            return false;
        }
        const double length = layerLength[nHops], blendFrac = layerBlend[nHops];
        const V3 orthoPoint = outerNeighCoords[pointI] + length * pointNormals[pointI];
        newPts[pointI] = blendFrac * orthoPoint + (1.0 - blendFrac) * newPts[pointI];
    }
    snapLayerBlend = newPts;
    return true;
}

// [OF-recalled] primitiveMesh addressing, SURVEY appendix A.2.
void Rank::buildConnectivity()
{
    // pointFaces: invert faces -> ascending face label
    pointFaces.assign(P, {});
    for (int f = 0; f < F; ++f)
        for (int k = fOff[f]; k < fOff[f + 1]; ++k)
            pointFaces[fV[k]].push_back(f);

    // cellPoints: union of the cell's face vertices, first seen
    // cell -> faces in plain "owner faces then neighbour faces" visiting order
    std::vector<std::vector<int>> cf(C);
    for (int f = 0; f < F; ++f)
        cf[own[f]].push_back(f);
    for (int f = 0; f < Fi; ++f)
        cf[nei[f]].push_back(f);
    cellPoints.assign(C, {});
    for (int c = 0; c < C; ++c)
        for (int f : cf[c])
            for (int k = fOff[f]; k < fOff[f + 1]; ++k)
                if (std::find(cellPoints[c].begin(), cellPoints[c].end(), fV[k]) == cellPoints[c].end())
                    cellPoints[c].push_back(fV[k]);

    // pointCells: ascending cell label
    pointCells.assign(P, {});
    for (int c = 0; c < C; ++c)
        for (int p : cellPoints[c])
            pointCells[p].push_back(c);

    // edges: unique (low,high) vertex pairs of consecutive face vertices,
    // numbered upper-triangular (meshes whose points are not sorted internal-first)
    std::vector<std::array<int, 2>> all;
    all.reserve(fV.size());
    for (int f = 0; f < F; ++f)
    {
        const int n = faceSize(f);
        const int *v = faceBegin(f);
        for (int i = 0; i < n; ++i)
        {
            int a = v[i], b = v[(i + 1) % n];
            if (a > b)
                std::swap(a, b);
            all.push_back({a, b});
        }
    }
    std::sort(all.begin(), all.end());
    all.erase(std::unique(all.begin(), all.end()), all.end());
    edges = all;
    const int E = (int)edges.size();

    // pointEdges sorted by edge label; pointPoints follows pointEdges
    pointEdges.assign(P, {});
    for (int e = 0; e < E; ++e)
    {
        pointEdges[edges[e][0]].push_back(e);
        pointEdges[edges[e][1]].push_back(e);
    }
    pointPoints.assign(P, {});
    for (int p = 0; p < P; ++p)
    {
        std::sort(pointEdges[p].begin(), pointEdges[p].end());
        for (int e : pointEdges[p])
            pointPoints[p].push_back(edges[e][0] == p ? edges[e][1] : edges[e][0]);
    }

    // edgeFaces: faces using the edge, ascending; edgeCells: first-seen unique
    auto edgeLabel = [&](int a, int b) {
        if (a > b)
            std::swap(a, b);
        std::array<int, 2> key = {a, b};
        return (int)(std::lower_bound(edges.begin(), edges.end(), key) - edges.begin());
    };
    edgeFaces.assign(E, {});
    for (int f = 0; f < F; ++f)
    {
        const int n = faceSize(f);
        const int *v = faceBegin(f);
        for (int i = 0; i < n; ++i)
            edgeFaces[edgeLabel(v[i], v[(i + 1) % n])].push_back(f);
    }
    edgeCells.assign(E, {});
    for (int e = 0; e < E; ++e)
        for (int f : edgeFaces[e])
        {
            auto &ec = edgeCells[e];
            if (std::find(ec.begin(), ec.end(), own[f]) == ec.end())
                ec.push_back(own[f]);
            if (f < Fi && std::find(ec.begin(), ec.end(), nei[f]) == ec.end())
                ec.push_back(nei[f]);
        }

    // generatePointNeighPoints, src/smoothMesh.C:190-217
    pointNeighPoints.assign(P, {});
    for (int pointI = 0; pointI < P; ++pointI)
        for (int cellI : pointCells[pointI])
            for (int pp : cellPoints[cellI])
            {
                if (pp == pointI)
                    continue;
                auto &l = pointNeighPoints[pointI];
                if (std::find(l.begin(), l.end(), pp) == l.end())
                    l.push_back(pp);
            }
}

// generateCellFaces, src/smoothMesh.C:1575-1620
bool Rank::buildCellFaces()
{
    cellFaces.assign(C, {});
    // fvMesh::owner()/neighbour() are the lduAddressing lower/upper lists, i.e.
    // internal faces only [OF-recalled; :1054 names its size nInternalFaces].
    for (int f = 0; f < Fi; ++f)
        cellFaces[own[f]].push_back(f);
    for (int f = 0; f < Fi; ++f)
        cellFaces[nei[f]].push_back(f);
    // boundary faces patch by patch (:1601-1617), processor patches included
    for (size_t patchI = 0; patchI < pKind.size(); ++patchI)
        for (int f = pStart[patchI]; f < pStart[patchI] + pSize[patchI]; ++f)
            cellFaces[own[f]].push_back(f);
    return true;
}

// findInternalMeshPoints, src/smoothMesh.C:40-91
void Rank::findInternalMeshPoints()
{
    isInternal.assign(P, 1);
    for (size_t patchI = 0; patchI < pKind.size(); ++patchI)
    {
        if (pKind[patchI] == 1) // processor patch
            continue;
        for (int f = pStart[patchI]; f < pStart[patchI] + pSize[patchI]; ++f)
            for (int k = fOff[f]; k < fOff[f + 1]; ++k)
                isInternal[fV[k]] = 0;
    }
}

// getMeshStats, src/smoothMesh.C:1478-1541 (edge length extrema only)
void Rank::getMeshStats()
{
    double minLength = VGREAT, maxLength = 0.0;
    bbMin = {VGREAT, VGREAT, VGREAT};
    bbMax = {-VGREAT, -VGREAT, -VGREAT};
    for (auto &e : edges)
    {
        const double length = mag(pts[e[1]] - pts[e[0]]);
        if (length < minLength)
            minLength = length;
        if (length > maxLength)
            maxLength = length;
        for (const V3 &q : {pts[e[0]], pts[e[1]]})
        {
            bbMin = {fmin_(q.x, bbMin.x), fmin_(q.y, bbMin.y), fmin_(q.z, bbMin.z)};
            bbMax = {fmax_(q.x, bbMax.x), fmax_(q.y, bbMax.y), fmax_(q.z, bbMax.z)};
        }
    }
    meshMinEdge = minLength;
    meshMaxEdge = maxLength;
    // :1538 as written: max z PLUS min z
    meshPerimeter = bbMax.x - bbMin.x + bbMax.y - bbMin.y + bbMax.z + bbMin.z;
}

// ---------------------------------------------------------------- geometry ----
// [OF-recalled] primitiveMesh::makeFaceCentresAndAreas / makeCellCentresAndVols.
// geometryVariant 0 = openfoam.com (v2312..v2506), 1 = openfoam.org (v12).
void Rank::calcGeometry()
{
    if (geomValid)
        return;
    faceCtr.resize(F);
    faceArea.resize(F);
    for (int facei = 0; facei < F; ++facei)
    {
        const int nPoints = faceSize(facei);
        const int *f = faceBegin(facei);
        if (nPoints == 3)
        {
            faceCtr[facei] = (1.0 / 3.0) * (pts[f[0]] + pts[f[1]] + pts[f[2]]);
            faceArea[facei] = 0.5 * cross(pts[f[1]] - pts[f[0]], pts[f[2]] - pts[f[0]]);
            continue;
        }
        V3 fCentre = pts[f[0]];
        for (int pi = 1; pi < nPoints; ++pi)
            fCentre += pts[f[pi]];
        fCentre = fCentre / double(nPoints);
        if (prm.geometryVariant == 0)
        {
            V3 sumN = {0, 0, 0}, sumAc = {0, 0, 0};
            double sumA = 0.0;
            for (int pi = 0; pi < nPoints; ++pi)
            {
                const V3 nextPoint = pts[f[(pi == nPoints - 1) ? 0 : pi + 1]];
                const V3 thisPoint = pts[f[pi]];
                const V3 c = thisPoint + nextPoint + fCentre;
                const V3 n = cross(nextPoint - thisPoint, fCentre - thisPoint);
                const double a = mag(n);
                sumN += n;
                sumA += a;
                sumAc += a * c;
            }
            if (sumA < SM_ROOTVSMALL)
            {
                faceCtr[facei] = fCentre;
                faceArea[facei] = {0, 0, 0};
            }
            else
            {
                faceCtr[facei] = ((1.0 / 3.0) * sumAc) / sumA;
                faceArea[facei] = 0.5 * sumN;
            }
        }
        else
        {
            V3 sumA = {0, 0, 0};
            for (int pi = 0; pi < nPoints; ++pi)
            {
                const V3 p0 = pts[f[pi]];
                const V3 p1 = pts[f[(pi == nPoints - 1) ? 0 : pi + 1]];
                sumA += cross(p1 - p0, fCentre - p0);
            }
            const double magSumA = mag(sumA);
            const V3 sumAHat = magSumA > 0 ? sumA / magSumA : V3{0, 0, 0};
            double sumAn = 0;
            V3 sumAnc = {0, 0, 0};
            for (int pi = 0; pi < nPoints; ++pi)
            {
                const V3 p0 = pts[f[pi]];
                const V3 p1 = pts[f[(pi == nPoints - 1) ? 0 : pi + 1]];
                const V3 a = cross(p1 - p0, fCentre - p0);
                const V3 c = p0 + p1 + fCentre;
                const double an = dot(a, sumAHat);
                sumAn += an;
                sumAnc += an * c;
            }
            faceCtr[facei] = (sumAn > VSMALL) ? ((1.0 / 3.0) * sumAnc) / sumAn : fCentre;
            faceArea[facei] = 0.5 * sumA;
        }
    }

    std::vector<V3> cEst(C, V3{0, 0, 0});
    std::vector<int> nCellFaces(C, 0);
    for (int facei = 0; facei < F; ++facei)
    {
        cEst[own[facei]] += faceCtr[facei];
        ++nCellFaces[own[facei]];
    }
    for (int facei = 0; facei < Fi; ++facei)
    {
        cEst[nei[facei]] += faceCtr[facei];
        ++nCellFaces[nei[facei]];
    }
    for (int c = 0; c < C; ++c)
        cEst[c] = cEst[c] / double(nCellFaces[c]);

    cellCtr.assign(C, V3{0, 0, 0});
    std::vector<double> cellVol(C, 0.0);
    for (int facei = 0; facei < F; ++facei)
    {
        const int c = own[facei];
        const double pyr3Vol = dot(faceArea[facei], faceCtr[facei] - cEst[c]);
        const V3 pc = (3.0 / 4.0) * faceCtr[facei] + (1.0 / 4.0) * cEst[c];
        cellCtr[c] += pyr3Vol * pc;
        cellVol[c] += pyr3Vol;
    }
    for (int facei = 0; facei < Fi; ++facei)
    {
        const int c = nei[facei];
        const double pyr3Vol = dot(faceArea[facei], cEst[c] - faceCtr[facei]);
        const V3 pc = (3.0 / 4.0) * faceCtr[facei] + (1.0 / 4.0) * cEst[c];
        cellCtr[c] += pyr3Vol * pc;
        cellVol[c] += pyr3Vol;
    }
    for (int c = 0; c < C; ++c)
    {
        if (std::fabs(cellVol[c]) > VSMALL)
            cellCtr[c] = cellCtr[c] / cellVol[c];
        else
            cellCtr[c] = cEst[c];
    }
    geomValid = true;
}

// -------------------------------------------------------------- predictor ----
// centroidalSmoothing, src/smoothMesh.C:96-131 (local partial sums; the hot
// path always runs with doBoundarySmoothing == false)
void Rank::centroidalPartial()
{
    calcGeometry();
    sumC.assign(P, V3{0, 0, 0});
    nC.assign(P, 0);
    for (int pointI = 0; pointI < P; ++pointI)
    {
        if (!doBoundarySmoothing && !isInternal[pointI]) // :116
            continue;
        const auto &pCells = pointCells[pointI];
        nC[pointI] = (int64_t)pCells.size();
        for (int celli : pCells)
            sumC[pointI] += cellCtr[celli];
    }
}

// src/smoothMesh.C:150-165
void Rank::centroidalFinish()
{
    centroidal = pts;
    for (int pointI = 0; pointI < P; ++pointI)
        if (nC[pointI])
            centroidal[pointI] = sumC[pointI] / double(nC[pointI]);
    snapCentroidal = centroidal;
}

// findAppropriateClosestPointLabel, src/smoothMesh.C:277-308
static int findAppropriateClosestPointLabel(const std::vector<int> &pointPoints, const std::vector<int> &sLabels,
                                            int pointI, const std::vector<uint8_t> &isInternalPoint, int stride)
{
    const bool isThisInternalPoint = isInternalPoint[pointI];
    int counter = 0;
    for (size_t i = 0; i < sLabels.size(); ++i)
    {
        const int labelI = sLabels[i];
        if ((!isThisInternalPoint) && (isInternalPoint[pointPoints[labelI]]))
            continue;
        if (counter == stride)
            return labelI;
        ++counter;
    }
    return UNDEF_LABEL;
}

// findClosestPoints (local initialisation part), src/smoothMesh.C:325-387
bool Rank::findClosestLocal()
{
    cp1.assign(P, ZERO_VECTOR);
    cp2.assign(P, ZERO_VECTOR);
    cp3.assign(P, ZERO_VECTOR);
    hasCommon.assign(P, 0);
    std::vector<double> edgeLengths;
    std::vector<int> sLabels;
    for (int pointI = 0; pointI < P; ++pointI)
    {
        const V3 cCoords = pts[pointI];
        const auto &pp = pointPoints[pointI];
        const int n = (int)pp.size();
        edgeLengths.assign(n, 0.0);
        for (int k = 0; k < n; ++k)
            edgeLengths[k] = mag(cCoords - pts[pp[k]]); // getPointDistance(points[n], cCoords): mag(cCoords - points[n])
        sLabels.resize(n);
        for (int k = 0; k < n; ++k)
            sLabels[k] = k;
        std::stable_sort(sLabels.begin(), sLabels.end(),
                         [&](int a, int b) { return edgeLengths[a] < edgeLengths[b]; }); // Foam::sortedOrder [OF-recalled]
        const int cLabel1 = findAppropriateClosestPointLabel(pp, sLabels, pointI, isInternal, 0);
        const int cLabel2 = findAppropriateClosestPointLabel(pp, sLabels, pointI, isInternal, 1);
        const int cLabel3 = findAppropriateClosestPointLabel(pp, sLabels, pointI, isInternal, 2);
        if (cLabel1 == UNDEF_LABEL || cLabel2 == UNDEF_LABEL)
        {
            char buf[160];
            snprintf(buf, sizeof buf, "Failed to find cLabel%d for pointI %d", cLabel1 == UNDEF_LABEL ? 1 : 2, pointI);
            err = buf;
            return false;
        }
        cp1[pointI] = pts[pp[cLabel1]] - cCoords;
        cp2[pointI] = pts[pp[cLabel2]] - cCoords;
        cp3[pointI] = (cLabel3 == UNDEF_LABEL) ? UNDEF_VECTOR : pts[pp[cLabel3]] - cCoords;
        const auto &nn = pointNeighPoints[pp[cLabel1]];
        hasCommon[pointI] = std::find(nn.begin(), nn.end(), pp[cLabel2]) != nn.end();
    }
    return true;
}

// isSmallerByVectorElements / isCloserPoint, src/smoothMesh.C:222-272
static bool isSmallerByVectorElements(V3 a, V3 b)
{
    const double va[3] = {a.x, a.y, a.z}, vb[3] = {b.x, b.y, b.z};
    for (int i = 0; i < 3; ++i)
    {
        if (va[i] < vb[i])
            return true;
        else if (va[i] > vb[i])
            return false;
    }
    return false;
}
static bool isCloserPoint(V3 point1, V3 point2)
{
    if (veq(point1, point2))
        return false;
    const double deltaDistance = mag(point1) - mag(point2);
    if (deltaDistance < VSMALL)
        return true;
    else if ((std::fabs(deltaDistance) < VSMALL) && isSmallerByVectorElements(point1, point2))
        return true;
    return false;
}

// calcARSmoothingRatio, src/smoothMesh.C:489-543
static double calcARSmoothingRatio(V3 c1, V3 c2, V3 c3, bool hasCommonCell, bool isInternalPoint)
{
    if (hasCommonCell)
        return 0.0;
    if (veq(c1, ZERO_VECTOR) || veq(c2, ZERO_VECTOR))
        return 0.0;
    const double lengthRatio1 = mag(c2) / mag(c1);
    const double lengthRatio2 = mag(c3) / mag(c2);
    if (isInternalPoint)
    {
        const double minRatio = 1.5, maxRatio = 3.0;
        if ((lengthRatio1 < minRatio) && (lengthRatio2 > minRatio))
        {
            const double frac = (lengthRatio2 - minRatio) / (maxRatio - minRatio);
            return fmin_(1.0, fmax_(0.0, frac));
        }
    }
    else
    {
        const double minRatio = 1.0, maxRatio = 2.0;
        const double frac = (lengthRatio1 - minRatio) / (maxRatio - minRatio);
        return fmin_(1.0, fmax_(0.0, frac));
    }
    return 0.0;
}

// aspectRatioSmoothing (blend part), src/smoothMesh.C:566-592
void Rank::aspectRatioBlend()
{
    newPts = centroidal;
    for (int pointI = 0; pointI < P; ++pointI)
    {
        const double blendFrac =
            calcARSmoothingRatio(cp1[pointI], cp2[pointI], cp3[pointI], hasCommon[pointI], isInternal[pointI]);
        if (blendFrac > 0.0)
        {
            const V3 aCoords = pts[pointI] + (cp1[pointI] + cp2[pointI]) / 2.0;
            newPts[pointI] = (1.0 - blendFrac) * centroidal[pointI] + blendFrac * aCoords;
        }
    }
    snapBlend = newPts;
}

// constrainMaxStepLength(..., doGlobalScaling=false), src/smoothMesh.C:684-754
void Rank::constrainMaxStepLength()
{
    const double maxStepLength = prm.maxStepLength, relStepFrac = prm.relStepFrac;
    for (int pointI = 0; pointI < P; ++pointI)
    {
        const V3 cCoords = pts[pointI];
        const V3 stepDir = newPts[pointI] - cCoords;
        double globalScale;
        if (mag(stepDir) > maxStepLength)
            globalScale = maxStepLength / (mag(stepDir) * relStepFrac);
        else
            globalScale = 1.0;
        newPts[pointI] = cCoords + (relStepFrac * globalScale) * stepDir;
    }
    snapClamped = newPts;
}

// ------------------------------------------------------------ constraints ----
// restrictEdgeShortening, src/smoothMesh.C:602-652
void Rank::restrictEdgeShortening()
{
    const double minEdgeLength = prm.minEdgeLength;
    for (int pointI = 0; pointI < P; ++pointI)
    {
        if (frozen[pointI])
            continue;
        const V3 cCoords = pts[pointI];
        const V3 nCoords = newPts[pointI];
        double shortestCurrentEdgeLength = GREAT, shortestNewEdgeLength = GREAT;
        for (int neighI : pointPoints[pointI])
        {
            const double testCurrentLength = mag(cCoords - pts[neighI]);
            if (testCurrentLength < shortestCurrentEdgeLength)
                shortestCurrentEdgeLength = testCurrentLength;
            const double testNewLength = mag(nCoords - pts[neighI]);
            if (testNewLength < shortestNewEdgeLength)
                shortestNewEdgeLength = testNewLength;
        }
        const double shortestLength = fmin_(shortestNewEdgeLength, shortestCurrentEdgeLength);
        if (prm.totalMinFreeze && (shortestLength < minEdgeLength))
            frozen[pointI] = 1;
        else if ((shortestNewEdgeLength < minEdgeLength) && (shortestNewEdgeLength < shortestCurrentEdgeLength))
            frozen[pointI] = 1;
    }
    snapFrozenEdgeLen = frozen;
}

// edgeEdgeAngle, src/smoothMesh.C:766-786
static double edgeEdgeAngle(V3 cCoords, V3 p1Coords, V3 p2Coords)
{
    V3 vec1 = p1Coords - cCoords;
    V3 vec2 = p2Coords - cCoords;
    vec1 = vec1 / mag(vec1);
    vec2 = vec2 / mag(vec2);
    const double cosA = dot(vec1, vec2);
    return ORC_ACOS(sm_clamp_cos(cosA));
}

// restrictMinEdgeAngleDecrease + calc_min_edge_angles + getNeighbourPoints,
// src/smoothMesh.C:793-930
void Rank::restrictMinEdgeAngleDecrease()
{
    const double smallAngle = M_PI * prm.minAngle / 180.0;
    for (int pointI = 0; pointI < P; ++pointI)
    {
        if (frozen[pointI])
            continue;
        double minCAngle = 1.7976931348623157e308, minNAngle = 1.7976931348623157e308; // DBL_MAX
        for (int faceI : pointFaces[pointI])
        {
            const int n = faceSize(faceI);
            const int *fp = faceBegin(faceI);
            int neighPI1 = 0, neighPI2 = 0;
            for (int k = 0; k < n; ++k)
                if (fp[k] == pointI)
                {
                    neighPI1 = fp[(k == 0) ? n - 1 : k - 1];
                    neighPI2 = fp[(k == n - 1) ? 0 : k + 1];
                    break;
                }
            const V3 cp0 = pts[pointI], c1 = pts[neighPI1], c2 = pts[neighPI2];
            const double cAngle = edgeEdgeAngle(cp0, c1, c2);
            const V3 np0 = newPts[pointI];
            const double nAngle0 = edgeEdgeAngle(np0, c1, c2);
            const V3 np1 = newPts[neighPI1], np2 = newPts[neighPI2];
            const double nAngle1 = edgeEdgeAngle(np0, np1, np2);
            const double nAngle2 = edgeEdgeAngle(np0, c1, np2);
            const double nAngle3 = edgeEdgeAngle(np0, np1, c2);
            const double nAngle = fmin_(fmin_(fmin_(nAngle0, nAngle1), nAngle2), nAngle3);
            if (cAngle < minCAngle)
                minCAngle = cAngle;
            if (nAngle < minNAngle)
                minNAngle = nAngle;
        }
        if ((minNAngle < smallAngle) && (minNAngle < minCAngle))
            frozen[pointI] = 1;
    }
    snapFrozenEdgeAngle = frozen;
}

// calcEdgeCenterEdgeAngle, src/smoothMesh.C:980-998
static double calcEdgeCenterEdgeAngle(V3 p0, V3 cC, V3 p1)
{
    const double cosA0 = dot(p0, cC);
    const double cosA1 = dot(cC, p1);
    const double angle0 = ORC_ACOS(sm_clamp_cos(cosA0));
    const double angle1 = ORC_ACOS(sm_clamp_cos(cosA1));
    return angle0 + angle1;
}

// calcMinMaxFaceAngleForEdge + calcFaceCenter + findCellFacePair +
// calcMinMaxFinalProjectedAngle, src/smoothMesh.C:1003-1231
bool Rank::calcMinMaxFaceAngleForEdge(int edgeI, double &minFaceAngle, double &maxFaceAngle, int pointI1, V3 coords1,
                                      int pointI2, V3 coords2)
{
    const auto &eFaces = edgeFaces[edgeI];
    const int nFaces = (int)eFaces.size();
    const int e0I = edges[edgeI][0], e1I = edges[edgeI][1];
    V3 e0 = pts[e0I];
    if (pointI1 >= 0 && e0I == pointI1)
        e0 = coords1;
    else if (pointI2 >= 0 && e0I == pointI2)
        e0 = coords2;
    V3 e1 = pts[e1I];
    if (pointI1 >= 0 && e1I == pointI1)
        e1 = coords1;
    else if (pointI2 >= 0 && e1I == pointI2)
        e1 = coords2;
    const V3 cCoords = 0.5 * (e0 + e1);
    const V3 eVec = (e1 - e0) / mag(e1 - e0);

    std::vector<V3> pVecs(nFaces, UNDEF_VECTOR);
    for (int i = 0; i < nFaces; ++i)
    {
        const int faceI = eFaces[i];
        // calcFaceCenter :1103-1130
        V3 center = {0, 0, 0};
        for (int k = fOff[faceI]; k < fOff[faceI + 1]; ++k)
        {
            const int pointI = fV[k];
            if (pointI1 >= 0 && pointI == pointI1)
                center += coords1;
            else if (pointI2 >= 0 && pointI == pointI2)
                center += coords2;
            else
                center += pts[pointI];
        }
        const V3 fCoords = center / double(faceSize(faceI));
        const V3 cf = cCoords - fCoords;
        const double dotProd = dot(cf, eVec);
        const V3 pCoords = fCoords + dotProd * eVec;
        pVecs[i] = (pCoords - cCoords) / mag(pCoords - cCoords);
    }

    const auto &eCells = edgeCells[edgeI];
    const int nCells = (int)eCells.size();
    double minAngle = 2.0 * M_PI, maxAngle = 0.0;
    for (int i = 0; i < nCells; ++i)
    {
        const int cellI = eCells[i];
        // findCellFacePair :1042-1097
        int face0I = UNDEF_LABEL, face1I = UNDEF_LABEL;
        for (int faceI : cellFaces[cellI])
        {
            const int faceIsI = (int)(std::find(eFaces.begin(), eFaces.end(), faceI) - eFaces.begin());
            if (faceIsI < nFaces)
            {
                if (face0I == UNDEF_LABEL)
                    face0I = faceIsI;
                else if (face1I == UNDEF_LABEL)
                    face1I = faceIsI;
                else
                {
                    err = "Sanity broken, more than two edge faces belong to same cell";
                    return false;
                }
            }
        }
        if (face0I == UNDEF_LABEL || face1I == UNDEF_LABEL || face0I == face1I)
        {
            err = "Sanity broken, didn't find face pairs for cell " + std::to_string(cellI);
            return false;
        }
        const V3 cellCenter = cellCtr[cellI];
        const V3 cf = cCoords - cellCenter;
        const double dotProd = dot(cf, eVec);
        const V3 pCoords = cellCenter + dotProd * eVec;
        const V3 cp = (pCoords - cCoords) / mag(pCoords - cCoords);
        const double angle = calcEdgeCenterEdgeAngle(pVecs[face0I], cp, pVecs[face1I]);
        if (angle < minAngle)
            minAngle = angle;
        if (angle > maxAngle)
            maxAngle = angle;
    }
    minFaceAngle = minAngle;
    maxFaceAngle = maxAngle;
    return true;
}

// calcMinMaxFaceAngleForPoint, src/smoothMesh.C:1276-1308
bool Rank::calcMinMaxFaceAngleForPoint(int pointI1, V3 coords1, int pointI2, V3 coords2, double &minFaceAngle,
                                       double &maxFaceAngle)
{
    minFaceAngle = 2.0 * M_PI;
    maxFaceAngle = 0.0;
    for (int edgeI : pointEdges[pointI1])
    {
        double minAngle, maxAngle;
        if (!calcMinMaxFaceAngleForEdge(edgeI, minAngle, maxAngle, pointI1, coords1, pointI2, coords2))
            return false;
        if (minFaceAngle > minAngle)
            minFaceAngle = minAngle;
        if (maxFaceAngle < maxAngle)
            maxFaceAngle = maxAngle;
    }
    return true;
}

// restrictFaceAngleDeterioration, src/smoothMesh.C:1320-1437
bool Rank::restrictFaceAngleDeterioration()
{
    calcGeometry(); // mesh.C() of the current mesh, :1218
    const int E = (int)edges.size();
    std::vector<double> curMinE(E, GREAT), curMaxE(E, GREAT);
    for (int e = 0; e < E; ++e) // calcCurrentMinMaxFaceAnglesForEdges :1252-1270
        if (!calcMinMaxFaceAngleForEdge(e, curMinE[e], curMaxE[e], -1, ZERO_VECTOR, -1, ZERO_VECTOR))
            return false;
    // mapCurrentMinMaxFaceAnglesToPoints :938-975
    std::vector<double> curMin(P, 2.0 * M_PI), curMax(P, 0.0);
    for (int e = 0; e < E; ++e)
        for (int s = 0; s < 2; ++s)
        {
            const int pointI = edges[e][s];
            if (curMin[pointI] > curMinE[e])
                curMin[pointI] = curMinE[e];
            if (curMax[pointI] < curMaxE[e])
                curMax[pointI] = curMaxE[e];
        }
    snapCurMin = curMin;
    snapCurMax = curMax;

    std::stack<int, std::vector<int>> pointStack;
    for (int pointI = 0; pointI < P; ++pointI)
        pointStack.push(pointI);
    const double smallAngle = M_PI * prm.minAngle / 180.0;
    const double largeAngle = M_PI * prm.maxAngle / 180.0;
    while (!pointStack.empty())
    {
        const int pointI = pointStack.top();
        pointStack.pop();
        if ((curMin[pointI] > smallAngle) && (curMax[pointI] < largeAngle))
            continue;
        const V3 cCoords = pts[pointI];
        V3 nCoords = newPts[pointI];
        if (frozen[pointI])
            nCoords = cCoords;
        if (!veq(nCoords, cCoords))
        {
            double newMin, newMax;
            if (!calcMinMaxFaceAngleForPoint(pointI, nCoords, -1, nCoords, newMin, newMax))
                return false;
            if (((newMin < smallAngle) && (newMin < curMin[pointI])) ||
                ((newMax > largeAngle) && (newMax > curMax[pointI])))
            {
                nCoords = cCoords;
                frozen[pointI] = 1;
            }
        }
        for (int neighPointI : pointPoints[pointI])
        {
            const V3 neighCoords = newPts[neighPointI];
            if (frozen[neighPointI])
                continue;
            if (veq(neighCoords, pts[neighPointI]))
                continue;
            double newMin, newMax;
            if (!calcMinMaxFaceAngleForPoint(pointI, nCoords, neighPointI, neighCoords, newMin, newMax))
                return false;
            if (((newMin < smallAngle) && (newMin < curMin[pointI])) ||
                ((newMax > largeAngle) && (newMax > curMax[pointI])))
            {
                frozen[neighPointI] = 1;
                pointStack.push(neighPointI);
            }
        }
    }
    snapFrozenFaceAngle = frozen;
    return true;
}

// restore + count + calculateResidual, src/smoothMesh.C:2384-2395, 1546-1570
void Rank::restoreAndResidual()
{
    nFrozen = 0;
    for (int pointI = 0; pointI < P; ++pointI)
        if (frozen[pointI] || (!isInternal[pointI] && !(doBoundarySmoothing && isSmoothingSurface[pointI]))) // :2387
        {
            newPts[pointI] = pts[pointI];
            ++nFrozen;
        }
    double maxStep = 0.0;
    for (int pointI = 0; pointI < P; ++pointI)
    {
        const double distance = mag(newPts[pointI] - pts[pointI]) / prm.maxStepLength;
        if (distance > maxStep)
            maxStep = distance;
    }
    residual = maxStep;
}

// mesh.movePoints, src/smoothMesh.C:2399 [OF-recalled: only points + cleared geometry are observable]
void Rank::movePoints()
{
    pts = newPts;
    geomValid = false;
}

// ------------------------------------------------------------------ group ----
struct Shared
{ // one globally shared point: its copies in ascending rank order
    std::vector<std::pair<int, int>> copies; // (rank, local point)
};

struct Group
{
    std::vector<Rank> ranks;
    std::vector<Shared> shared;
    int threads = 1;
    std::string err;

    void buildShared()
    {
        if (ranks.size() < 2)
            return;
        std::map<int64_t, std::vector<std::pair<int, int>>> m;
        for (size_t r = 0; r < ranks.size(); ++r)
            for (int p = 0; p < ranks[r].P; ++p)
                m[ranks[r].gid[p]].push_back({(int)r, p});
        for (auto &kv : m)
            if (kv.second.size() > 1)
                shared.push_back({kv.second});
    }

    // syncTools::syncPointList [OF-recalled, SURVEY A.3]: combine all copies
    // (here: ascending rank order) and write the result back to every copy.
    template <class T, class Op> void sync(std::vector<T> Rank::*field, Op op)
    {
        for (auto &s : shared)
        {
            T v = (ranks[s.copies[0].first].*field)[s.copies[0].second];
            for (size_t k = 1; k < s.copies.size(); ++k)
                op(v, (ranks[s.copies[k].first].*field)[s.copies[k].second]);
            for (auto &c : s.copies)
                (ranks[c.first].*field)[c.second] = v;
        }
    }

    template <class Fn> bool forRanks(Fn fn)
    {
        bool ok = true;
        const int n = (int)ranks.size();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) if (n > 1 && threads > 1)
        for (int r = 0; r < n; ++r)
            if (!fn(ranks[r]))
            {
#pragma omp critical
                {
                    ok = false;
                    err = ranks[r].err;
                }
            }
        return ok;
    }

    // one closest-point merge step across ranks, src/smoothMesh.C:391-469
    void mergeClosest(int position)
    {
        if (shared.empty())
            return;
        std::vector<std::vector<V3>> syncPts(ranks.size());
        for (size_t r = 0; r < ranks.size(); ++r)
            syncPts[r] = position == 1 ? ranks[r].cp1 : position == 2 ? ranks[r].cp2 : ranks[r].cp3;
        for (auto &s : shared)
        {
            V3 v = syncPts[s.copies[0].first][s.copies[0].second];
            for (size_t k = 1; k < s.copies.size(); ++k)
            { // minMagSqrEqOp: x = magSqr(x) <= magSqr(y) ? x : y
                const V3 y = syncPts[s.copies[k].first][s.copies[k].second];
                v = (magSqr(v) <= magSqr(y)) ? v : y;
            }
            for (auto &c : s.copies)
            {
                Rank &R = ranks[c.first];
                const int p = c.second;
                if (position == 1)
                {
                    if (isCloserPoint(v, R.cp1[p]))
                    {
                        R.cp3[p] = R.cp2[p];
                        R.cp2[p] = R.cp1[p];
                        R.cp1[p] = v;
                        R.hasCommon[p] = 0;
                    }
                }
                else if (position == 2)
                {
                    if (isCloserPoint(v, R.cp2[p]))
                    {
                        R.cp3[p] = R.cp2[p];
                        R.cp2[p] = v;
                        R.hasCommon[p] = 0;
                    }
                }
                else
                {
                    if (isCloserPoint(v, R.cp3[p]))
                        R.cp3[p] = v;
                }
            }
        }
    }

    // calculateBoundaryPointNormals with its two synchronisations (orthogonalBoundaryBlending.C:185-198)
    void calculateBoundaryPointNormals()
    {
        forRanks([](Rank &R) {
            if (R.doLayerTreatment || R.doBoundarySmoothing)
                R.boundaryNormalsLocal();
            return true;
        });
        if (ranks[0].doLayerTreatment || ranks[0].doBoundarySmoothing)
        {
            sync(&Rank::pointNormals, [](V3 &x, const V3 &y) { x = x + y; });    // plusEqOp<vector>
            sync(&Rank::nBoundaryFaces, [](int &x, const int &y) { x = x + y; }); // plusEqOp<label>
        }
        forRanks([](Rank &R) {
            if (R.doLayerTreatment || R.doBoundarySmoothing)
                R.boundaryNormalsFinish();
            return true;
        });
    }

    // One-time set-up of the boundary layer treatment and of boundary point smoothing,
    // src/smoothMesh.C:2024-2250, with the synchronisations of orthogonalBoundaryBlending.C:124-130 (max),
    // :185-198 (sum), :363-369 (maxMagSqr).  false on a FatalError-equivalent.
    bool setupLayers()
    {
        // getMeshStats' reductions (:1527-1534): extrema over all ranks
        if (ranks.size() > 1)
        {
            V3 lo = ranks[0].bbMin, hi = ranks[0].bbMax;
            for (auto &R : ranks)
            {
                lo = {fmin_(R.bbMin.x, lo.x), fmin_(R.bbMin.y, lo.y), fmin_(R.bbMin.z, lo.z)};
                hi = {fmax_(R.bbMax.x, hi.x), fmax_(R.bbMax.y, hi.y), fmax_(R.bbMax.z, hi.z)};
            }
            double mn = VGREAT, mx = 0.0;
            for (auto &R : ranks)
            {
                mn = fmin_(R.meshMinEdge, mn);
                mx = fmax_(R.meshMaxEdge, mx);
            }
            for (auto &R : ranks)
            {
                R.meshPerimeter = hi.x - lo.x + hi.y - lo.y + hi.z + lo.z;
                R.meshMinEdge = mn;
                R.meshMaxEdge = mx;
            }
        }
        for (auto &R : ranks)
        {
            R.layersBegin();
            if (!R.boundaryBegin())
            {
                err = R.err;
                return false;
            }
            R.classifyBoundaryPoints();
            if (!R.err.empty())
            {
                err = R.err;
                return false;
            }
        }
        const bool doLayers = ranks[0].doLayerTreatment, doBoundary = ranks[0].doBoundarySmoothing;
        if (!doLayers && !doBoundary)
            return true;
        const int maxIter = ranks[0].prm.maxLayers + 1;
        for (auto &R : ranks)
            R.hopsInit();
        for (int iter = 0; iter < maxIter; ++iter)
        {
            for (auto &R : ranks)
                R.hopsSweep();
            sync(&Rank::hopsToLayer, [](int &x, const int &y) { x = (x > y) ? x : y; }); // maxEqOp<label>
        }
        for (auto &R : ranks)
            R.hopsInitFor(R.hopsToSmoothing, R.pSmooth);
        for (int iter = 0; iter < 2; ++iter) // :2218, maxIter = 2
        {
            for (auto &R : ranks)
                R.hopsSweepFor(R.hopsToSmoothing);
            sync(&Rank::hopsToSmoothing, [](int &x, const int &y) { x = (x > y) ? x : y; });
        }
        calculateBoundaryPointNormals();
        for (auto &R : ranks)
            R.outerInit();
        for (int iter = 1; iter < maxIter + 1; ++iter)
        {
            for (auto &R : ranks)
                R.outerSweep(iter);
            // maxMagSqrEqOp<vector> [OF-recalled, ops.H]: x = (magSqr(x) >= magSqr(y)) ? x : y
            sync(&Rank::pointNormals, [](V3 &x, const V3 &y) { x = (magSqr(x) >= magSqr(y)) ? x : y; });
        }
        for (auto &R : ranks)
        {
            R.outerUndo();
            R.layerTables();
            R.innerNeighInfo();
            if (doBoundary)
                R.boundaryPointStrings();
            if (!R.err.empty())
            {
                err = R.err;
                return false;
            }
        }
        return true;
    }

    // One smoothing iteration, src/smoothMesh.C:2257-2399.  Returns false on a
    // FatalError-equivalent.
    bool iterate(int64_t &nFrozenSum, double &res)
    {
        // :2266 (the call is unconditional in the reference; its result is only used with layer treatment)
        calculateBoundaryPointNormals();
        forRanks([](Rank &R) {
            R.frozen.assign(R.P, 0); // :2262-2263
            if (R.doLayerTreatment)
                R.snapNormals = R.pointNormals;
            R.centroidalPartial();   // :2269
            R.snapCellCtr = R.cellCtr;
            return true;
        });
        sync(&Rank::sumC, [](V3 &x, const V3 &y) { x = x + y; });           // :134
        sync(&Rank::nC, [](int64_t &x, const int64_t &y) { x = x + y; });   // :142
        if (!forRanks([](Rank &R) {
                R.centroidalFinish();
                return R.findClosestLocal(); // :2276 -> :577
            }))
            return false;
        mergeClosest(1);
        mergeClosest(2);
        mergeClosest(3);
        sync(&Rank::hasCommon, [](uint8_t &x, const uint8_t &y) { x = x || y; }); // :472
        forRanks([](Rank &R) {
            R.aspectRatioBlend();
            R.constrainMaxStepLength(); // :2280
            if (R.doLayerTreatment)
                R.updateNeighCoordsLocal(); // :2286
            return true;
        });
        if (ranks[0].doLayerTreatment) // minMagSqrEqOp<vector>, orthogonalBoundaryBlending.C:491-497
            sync(&Rank::outerNeighCoords, [](V3 &x, const V3 &y) { x = (magSqr(x) <= magSqr(y)) ? x : y; });
        if (!forRanks([](Rank &R) {
                if (R.doLayerTreatment)
                { // :2288-2304
                    if (!R.blendWithOrthogonalPoints())
                        return false;
                    R.constrainMaxStepLength();
                }
                if (R.doBoundarySmoothing)
                { // :2310, and the local halves of calculateFeatureEdgeProjections / calculateSurfaceCentroids
                    R.updateInnerNeighCoordsLocal();
                    R.featureEdgeProjectionsLocal();
                    if (!R.err.empty() || !R.surfaceCentroidsLocal())
                        return false;
                }
                return true;
            }))
            return false;
        if (ranks[0].doBoundarySmoothing)
        {
            sync(&Rank::innerNeighCoords, [](V3 &x, const V3 &y) { x = (magSqr(x) <= magSqr(y)) ? x : y; }); // :491
            sync(&Rank::featureEdgeProjections, [](V3 &x, const V3 &y) { x = x + y; });                       // :660
            sync(&Rank::nFeatureEdgeProjections, [](int &x, const int &y) { x = x + y; });                    // :668
            sync(&Rank::nFaceCentroids, [](int &x, const int &y) { x = x + y; });                             // :830
        }
        if (!forRanks([](Rank &R) {
                if (R.doBoundarySmoothing)
                { // :2312-2355
                    if (!R.projectBoundaryPoints() || !R.projectPrismaticInternalPoints())
                        return false;
                    R.constrainMaxStepLength();
                }
                R.restrictEdgeShortening(); // :2359
                if (R.prm.edgeAngleConstraint)
                    R.restrictMinEdgeAngleDecrease(); // :2364
                if (R.prm.faceAngleConstraint)
                    return R.restrictFaceAngleDeterioration(); // :2370
                return true;
            }))
            return false;
        sync(&Rank::frozen, [](uint8_t &x, const uint8_t &y) { x = x || y; }); // :2374
        forRanks([](Rank &R) {
            R.restoreAndResidual();
            return true;
        });
        nFrozenSum = 0;
        res = 0.0;
        for (auto &R : ranks)
        { // returnReduce sum / max, :1567, :2396
            nFrozenSum += R.nFrozen;
            if (R.residual > res)
                res = R.residual;
        }
        forRanks([](Rank &R) {
            R.movePoints(); // :2399
            return true;
        });
        return true;
    }
};

template <class T> int copyOut(const std::vector<T> &v, void *out, int64_t capBytes)
{
    const int64_t n = (int64_t)(v.size() * sizeof(T));
    if (n > capBytes)
        return -2;
    memcpy(out, v.data(), (size_t)n);
    return (int)v.size();
}

} // namespace

// ================================================================== C API ====
extern "C"
{
    struct orc_mesh
    {
        int64_t P, C, F, Fi;
        const double *pts;
        const int32_t *fOff, *fV, *own, *nei;
        int32_t nPatches;
        const int32_t *pStart, *pSize, *pKind;
        const int64_t *pointGlobalId;
        const int32_t *pLayer;
        const int32_t *pSmooth;
    };
    struct orc_params
    {
        double minEdgeLength, maxStepLength, relStepFrac, minAngle, maxAngle, relTol;
        int32_t totalMinFreeze, edgeAngleConstraint, faceAngleConstraint, geometryVariant;
        double layerMaxBlendingFraction, layerEdgeLength, layerExpansionRatio;
        int32_t minLayers, maxLayers;
    };

    const char *orc_last_error() { return g_last_error.c_str(); }

    void *orc_create(int nRanks, const orc_mesh *meshes, const orc_params *p)
    {
        Group *g = new Group;
        g->ranks.resize(nRanks);
        Params prm;
        memcpy(&prm, p, sizeof(prm));
        for (int r = 0; r < nRanks; ++r)
        {
            MeshIn m;
            memcpy(&m, &meshes[r], sizeof(m));
            if (nRanks > 1 && !m.pointGlobalId)
            {
                g_last_error = "pointGlobalId required for nRanks > 1";
                delete g;
                return nullptr;
            }
            if (!g->ranks[r].init(m, prm))
            {
                g_last_error = g->ranks[r].err;
                delete g;
                return nullptr;
            }
        }
        g->buildShared();
        return g;
    }

    void orc_destroy(void *h) { delete (Group *)h; }
    // isCornerPoint / isFeatureEdgePoint label lists of an earlier run for one rank (restart); before orc_set_params
    void orc_set_label_lists(void *h, int rank, int64_t n, const int32_t *isCornerPoint, const int32_t *isFeatureEdgePoint)
    {
        Rank &R = ((Group *)h)->ranks[rank];
        R.cornerIO.assign(isCornerPoint, isCornerPoint + n);
        R.featureIO.assign(isFeatureEdgePoint, isFeatureEdgePoint + n);
    }

    // Inputs of boundary point smoothing (constant/geometry/initEdges.obj, targetEdges.obj -- the initial
    // edges again when that file is absent, src/smoothMesh.C:2148-2160 --, targetSurfaces.obj) as arrays,
    // and -internalSmoothingBlendingFraction.  Every rank reads the same files.  Call before orc_set_params.
    void orc_set_geometry(void *h, int64_t nInitPts, const double *initPts, int64_t nInitEdges, const int32_t *initEdges,
                          int64_t nTargetPts, const double *targetPts, int64_t nTargetEdges, const int32_t *targetEdges,
                          int64_t nSurfPts, const double *surfPts, int64_t nSurfTris, const int32_t *surfTris,
                          double internalSmoothingBlendingFraction)
    {
        Group *g = (Group *)h;
        for (auto &R : g->ranks)
        {
            auto fill = [](Rank::EdgeMesh &em, int64_t np, const double *p, int64_t ne, const int32_t *e) {
                em.points.clear();
                em.edges.clear();
                for (int64_t i = 0; i < np; ++i)
                    em.points.push_back({p[3 * i], p[3 * i + 1], p[3 * i + 2]});
                for (int64_t i = 0; i < ne; ++i)
                    em.edges.push_back({e[2 * i], e[2 * i + 1]});
                em.finish();
            };
            fill(R.initEdges, nInitPts, initPts, nInitEdges, initEdges);
            fill(R.targetEdges, nTargetPts, targetPts, nTargetEdges, targetEdges);
            R.surf.points.clear();
            R.surf.tris.clear();
            for (int64_t i = 0; i < nSurfPts; ++i)
                R.surf.points.push_back({surfPts[3 * i], surfPts[3 * i + 1], surfPts[3 * i + 2]});
            for (int64_t i = 0; i < nSurfTris; ++i)
                R.surf.tris.push_back({surfTris[3 * i], surfTris[3 * i + 1], surfTris[3 * i + 2]});
            R.haveGeometry = true;
            R.internalSmoothingBlendingFraction = internalSmoothingBlendingFraction;
        }
    }
    void orc_set_threads(void *h, int n) { ((Group *)h)->threads = n < 1 ? 1 : n; }

    // Mesh statistics used for option defaults, src/smoothMesh.C:1857-1865
    // (returnReduce min/max over ranks).
    void orc_mesh_stats(void *h, double *minEdge, double *maxEdge)
    {
        Group *g = (Group *)h;
        double mn = VGREAT, mx = 0;
        for (auto &R : g->ranks)
        {
            mn = std::min(mn, R.meshMinEdge);
            mx = std::max(mx, R.meshMaxEdge);
        }
        *minEdge = mn;
        *maxEdge = mx;
    }
    // Sets the resolved options and runs the one-time layer set-up (:2215-2221); 0 on success.
    int orc_set_params(void *h, const orc_params *p)
    {
        Group *g = (Group *)h;
        for (auto &R : g->ranks)
            memcpy(&R.prm, p, sizeof(Params));
        if (!g->setupLayers())
        {
            g_last_error = g->err;
            return -1;
        }
        return 0;
    }

    // The iteration loop with the stop rule of src/smoothMesh.C:2257, 2401-2411.
    // Returns the number of iterations executed, or -1 on error.
    int orc_iterate(void *h, int maxIters, int64_t *nFrozen, double *residual)
    {
        Group *g = (Group *)h;
        int done = 0;
        for (int i = 0; i < maxIters; ++i)
        {
            int64_t nf;
            double res;
            if (!g->iterate(nf, res))
            {
                g_last_error = g->err;
                return -1;
            }
            if (nFrozen)
                nFrozen[i] = nf;
            if (residual)
                residual[i] = res;
            ++done;
            if (res < g->ranks[0].prm.relTol)
                break;
        }
        return done;
    }

    int64_t orc_sizes(void *h, int rank, int what)
    {
        Rank &R = ((Group *)h)->ranks[rank];
        switch (what)
        {
        case 0:
            return R.P;
        case 1:
            return R.C;
        case 2:
            return R.F;
        case 3:
            return (int64_t)R.edges.size();
        }
        return -1;
    }

    // name -> array of the last executed iteration (or static data)
    int orc_get(void *h, int rank, const char *name, void *out, int64_t capBytes)
    {
        Rank &R = ((Group *)h)->ranks[rank];
        const std::string n = name;
        if (n == "points")
            return copyOut(R.pts, out, capBytes);
        if (n == "frozen")
            return copyOut(R.frozen, out, capBytes);
        if (n == "isInternal")
            return copyOut(R.isInternal, out, capBytes);
        if (n == "cellCentres")
        {
            R.calcGeometry();
            return copyOut(R.cellCtr, out, capBytes);
        }
        if (n == "faceCentres")
        {
            R.calcGeometry();
            return copyOut(R.faceCtr, out, capBytes);
        }
        if (n == "faceAreas")
        {
            R.calcGeometry();
            return copyOut(R.faceArea, out, capBytes);
        }
        if (n == "snapCellCtr")
            return copyOut(R.snapCellCtr, out, capBytes);
        if (n == "snapCentroidal")
            return copyOut(R.snapCentroidal, out, capBytes);
        if (n == "snapBlend")
            return copyOut(R.snapBlend, out, capBytes);
        if (n == "snapClamped")
            return copyOut(R.snapClamped, out, capBytes);
        if (n == "snapFrozenEdgeLen")
            return copyOut(R.snapFrozenEdgeLen, out, capBytes);
        if (n == "snapFrozenEdgeAngle")
            return copyOut(R.snapFrozenEdgeAngle, out, capBytes);
        if (n == "snapFrozenFaceAngle")
            return copyOut(R.snapFrozenFaceAngle, out, capBytes);
        if (n == "snapCurMin")
            return copyOut(R.snapCurMin, out, capBytes);
        if (n == "snapCurMax")
            return copyOut(R.snapCurMax, out, capBytes);
        if (n == "snapNormals")
            return copyOut(R.snapNormals, out, capBytes);
        if (n == "snapLayerBlend")
            return copyOut(R.snapLayerBlend, out, capBytes);
        if (n == "hopsToLayer")
            return copyOut(R.hopsToLayer, out, capBytes);
        if (n == "pointToOuter")
            return copyOut(R.pointToOuter, out, capBytes);
        if (n == "isCorner")
            return copyOut(R.isCorner, out, capBytes);
        if (n == "isFeatureEdge")
            return copyOut(R.isFeatureEdge, out, capBytes);
        if (n == "isSmoothingSurface")
            return copyOut(R.isSmoothingSurface, out, capBytes);
        if (n == "cornerPoints")
            return copyOut(R.cornerPoints, out, capBytes);
        if (n == "pointStrings")
            return copyOut(R.pointStrings, out, capBytes);
        if (n == "hopsToSmoothing")
            return copyOut(R.hopsToSmoothing, out, capBytes);
        if (n == "pointToInner")
            return copyOut(R.pointToInner, out, capBytes);
        if (n == "edges")
            return copyOut(R.edges, out, capBytes);
        return -1;
    }

    // ragged connectivity tables, flattened: returns total entries; offsets has n+1 entries
    int64_t orc_get_csr(void *h, int rank, const char *name, int32_t *offsets, int32_t *values, int64_t capValues)
    {
        Rank &R = ((Group *)h)->ranks[rank];
        const std::string n = name;
        const std::vector<std::vector<int>> *t = nullptr;
        if (n == "pointFaces")
            t = &R.pointFaces;
        else if (n == "pointCells")
            t = &R.pointCells;
        else if (n == "pointPoints")
            t = &R.pointPoints;
        else if (n == "pointEdges")
            t = &R.pointEdges;
        else if (n == "edgeFaces")
            t = &R.edgeFaces;
        else if (n == "edgeCells")
            t = &R.edgeCells;
        else if (n == "cellFaces")
            t = &R.cellFaces;
        else if (n == "cellPoints")
            t = &R.cellPoints;
        if (!t)
            return -1;
        int64_t tot = 0;
        for (auto &row : *t)
            tot += (int64_t)row.size();
        if (!offsets)
            return tot;
        if (tot > capValues)
            return -2;
        int64_t k = 0;
        for (size_t i = 0; i < t->size(); ++i)
        {
            offsets[i] = (int32_t)k;
            for (int v : (*t)[i])
                values[k++] = v;
        }
        offsets[t->size()] = (int32_t)k;
        return tot;
    }

    // scalar helpers exposed for known-answer tests
    double orc_acos(double x) { return ORC_ACOS(x); }
    double orc_edgeEdgeAngle(const double *c, const double *p1, const double *p2)
    {
        return edgeEdgeAngle({c[0], c[1], c[2]}, {p1[0], p1[1], p1[2]}, {p2[0], p2[1], p2[2]});
    }
    int orc_edge_face_angles(void *h, int rank, int edgeI, double *mn, double *mx)
    {
        Rank &R = ((Group *)h)->ranks[rank];
        R.calcGeometry();
        return R.calcMinMaxFaceAngleForEdge(edgeI, *mn, *mx, -1, ZERO_VECTOR, -1, ZERO_VECTOR) ? 0 : -1;
    }
}
