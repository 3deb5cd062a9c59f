// boundary.hpp -- one-time host set-up of boundary point smoothing (src/boundaryPointSmoothing.C and the
// calls around it in src/smoothMesh.C:2080-2250): inputs from constant/geometry/*.obj, sanity checks, edge
// strings, classification of the boundary points, hop counts to the smoothing patches, inner-neighbour map.
// Pure CPU code; the per-iteration projections run on the device (kernels.cuh: k_boundary_*).
#pragma once
#include "polymesh.hpp"
#include "topology.hpp"

#include <array>
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

namespace sm
{

// edgeMesh as OpenFOAM reads it from an OBJ file: points, edges, edges of every point in ascending edge order
struct EdgeMesh
{
    std::vector<double> points; // xyz
    std::vector<int32_t> edges; // pairs
    std::vector<std::vector<int32_t>> pointEdges;
    int64_t nPoints() const { return (int64_t)points.size() / 3; }
    int64_t nEdges() const { return (int64_t)edges.size() / 2; }
    void finish();
};
struct TriSurface
{
    std::vector<double> points; // xyz
    std::vector<int32_t> tris;  // triples
    int64_t nTris() const { return (int64_t)tris.size() / 3; }
};
// Wavefront OBJ: "v", "l" (polylines) and "f" (fan-triangulated) records
void readObj(const std::string &file, std::vector<double> &points, std::vector<int32_t> &edges, std::vector<int32_t> &tris);

// Bounding volume hierarchy over the target triangles: the acceleration structure for the surface ray casts
// (findLine).  It changes which triangles are tested, never the arithmetic of a test nor the choice among hits
// (smallest parameter, lower triangle label on ties), so a traversal returns exactly what the visit-every-triangle
// search returns.  Host builder and host traversal (the device kernel still visits every triangle; porting the
// traversal is the next step for large surfaces).
struct TriangleBvh
{
    // node i: box lo/hi (6 doubles, slightly inflated), then either two children or a leaf
    std::vector<double> box;           // 6 per node
    std::vector<int32_t> left, right;  // internal node: child nodes; leaf: -1
    std::vector<int32_t> first, count; // leaf: range in `order`
    std::vector<int32_t> order;  // triangle labels, leaf ranges are contiguous
};
TriangleBvh buildTriangleBvh(const TriSurface &s, int leafSize = 4);
// the intersection of the segment start -> end with the surface that is nearest to start: returns the triangle
// label (-1 = none) and the hit point; `useBvh = false` visits every triangle (the reference definition)
int32_t segmentSurfaceHit(const TriSurface &s, const TriangleBvh *bvh, const double start[3], const double end[3], double hit[3]);

struct BoundarySetup
{
    EdgeMesh targetEdges;
    TriSurface surface;
    std::vector<int32_t> targetEdgeStrings;                            // per target edge
    std::vector<uint8_t> isCorner, isFeatureEdge, isSmoothingSurface;  // per point
    std::vector<uint8_t> isConnectedToInternal;
    std::vector<double> cornerPoints;                                  // xyz per point (corner points only)
    std::vector<int32_t> pointStrings, hopsToSmoothing, pointToInner;  // per point, -1 = none
    std::vector<int32_t> boundaryPoints;                               // non-internal points, ascending
    double distanceTolerance = 0;
    int64_t nCorners = 0, nFeatureEdgePoints = 0, nSmoothingSurfacePoints = 0;
    int32_t nStrings = 0;
};
// Throws std::runtime_error with the reference's FatalError texts.  layerEdgeLength / minEdgeLength are
// the resolved options (distanceTolerance = REL_TOL min(meshMinEdgeLength, layerEdgeLength), :1921).
// cornerIO / featureIO: the isCornerPoint / isFeatureEdgePoint label lists of an earlier run (empty = none).
// When either holds a 1 the classes are taken from them instead of from the initial edges (:336-340).
// Processor mesh of a decomposed run (`par` non-null): the mesh-wide figures come from all ranks (getMeshStats'
// reductions, src/smoothMesh.C:1527-1534) and the hop counts to the smoothing patches are synchronised (max) after
// each of their two sweeps (src/orthogonalBoundaryBlending.C:124-130); everything else is rank-local in the
// reference too.
struct BoundaryParallel
{
    double meshMinEdgeLength = 0, meshPerimeter = 0;      // global
    std::function<void(std::vector<int32_t> &)> maxInt;   // syncPointList(maxEqOp<label>)
};
BoundarySetup buildBoundarySetup(const PolyMesh &patchesAndFaces, const Topology &t, const std::vector<double> &points,
                                 const EdgeMesh &initEdges, const EdgeMesh &targetEdges, const TriSurface &surface,
                                 const std::vector<int32_t> &patchSmoothing, double layerEdgeLength,
                                 const std::vector<int32_t> &cornerIO = {}, const std::vector<int32_t> &featureIO = {},
                                 const BoundaryParallel *par = nullptr);
// getMeshStats' bounding box over the points (src/smoothMesh.C:1495-1512): lo[3], hi[3]
void meshBoundingBox(const Topology &t, const std::vector<double> &points, double lo[3], double hi[3]);

} // namespace sm
