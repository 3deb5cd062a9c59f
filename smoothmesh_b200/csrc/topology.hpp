// topology.hpp -- derived connectivity for the smoothing kernels, built once on
// the host and uploaded to HBM.
//
// Row orders follow what OpenFOAM's primitiveMesh hands the reference (SURVEY.md
// appendix A.2): pointCells ascending cell label (sum order at
// src/smoothMesh.C:127-130), pointPoints / pointEdges ascending neighbour label
// (stable-sort tie-break :345-352, push order :1406-1431), pointFaces and
// edgeFaces ascending face label.  The reference's per-call searches
// (getNeighbourPoints :793-831, findCellFacePair :1042-1097, the share-a-cell
// lookup :383) depend only on topology, so they are resolved here once.
#pragma once
#include "polymesh.hpp"
#include <cstdint>
#include <functional>
#include <vector>

namespace sm
{

struct Topology
{
    int64_t P = 0, C = 0, F = 0, Fi = 0, E = 0;

    // point -> cells (ascending)
    Vec<int32_t> pcOff, pc;
    // point -> edge-connected points (ascending) and the matching edge labels
    Vec<int32_t> ppOff, pp, pe;
    // point -> face corners: for each face of pointFaces(p) (ascending) the previous and
    // next vertex of p in that face: corner[2*k], corner[2*k+1]
    Vec<int32_t> cornerOff, corner;
    // edges (lo,hi), numbered upper-triangular
    Vec<int32_t> edge;
    // edge -> faces (ascending)
    Vec<int32_t> efOff, ef;
    // edge -> cells, each with the positions (within the edge's face row) of the two
    // faces of that cell that meet at the edge: ecPair = f0 | f1<<16
    Vec<int32_t> ecOff, ecCell, ecPair;
    // faces (copied from the mesh; vertex loops)
    Vec<int32_t> faceOff, faceVerts;
    // cell -> faces in OpenFOAM's cell-centre accumulation order: faces the cell owns
    // (ascending), then faces it neighbours (ascending, bit 31 set)
    Vec<int32_t> cfOff, cf;
    // Fixed-size, vector-loadable records for the common low-valence case; the CSR tables above
    // stay authoritative and serve the generic path (flag bit 31 of the meta word).
    //   pointRec  16 words/point: pc[0..7], pp[8..13], [14] = nCells | nNbrs<<8 | generic<<31,
    //             [15] = mask over the 15 unordered pairs of pp positions that are face corners of the point
    //   edgeRec   12 words/edge : e0,e1, f[4], c[4], [10] = nf | nc<<4 | generic<<31; faces in fan order around
    //             the edge, cell k lies between face k and face (k+1) mod nf
    Vec<int32_t> pointRec, edgeRec;
    Vec<uint8_t> isInternal; // src/smoothMesh.C:40-91
    std::vector<int32_t> procPoints; // points on processor patches (ascending)
    double minEdgeLength = 0, maxEdgeLength = 0; // src/smoothMesh.C:1478-1541
    int32_t maxPointDegree = 0, maxFaceSize = 0, maxEdgeFaces = 0;
};

// One-time data of the prismatic boundary layer treatment (src/orthogonalBoundaryBlending.C).
struct LayerSetup
{
    std::vector<int32_t> hops;         // pointHopsToLayerBoundary, -1 = undefined
    std::vector<int32_t> pointToOuter; // pointToOuterPointMap, -1 = none
    std::vector<int32_t> normalSrc;    // boundary point whose set-up normal an internal point carries, -1 = zero
    std::vector<int32_t> bfOff, bf;    // point -> boundary faces (non-processor patches), ascending
    int32_t maxHop = 0;
};
LayerSetup buildLayerSetup(const PolyMesh &m, const Topology &t, const std::vector<int32_t> &patchLayer, int maxLayers);

// The same set-up for a processor mesh of a decomposed case.  The reference synchronises the hop counts
// (max), the boundary normals and face counts (sum) and the propagated normals (maxMagSqr) over the copies
// of every interface point between its sweeps (src/orthogonalBoundaryBlending.C:124-130, :185-198,
// :363-369); which copy wins depends on the normals' values, so the propagation runs on values here.
// `normals` (xyz per point) comes in as this rank's accumulated boundary normals of the set-up call
// (zero minus the unit normals of the point's boundary faces) and leaves as the set-up normals.
struct LayerSync
{
    std::function<void(std::vector<int32_t> &)> maxInt, sumInt;
    std::function<void(std::vector<double> &)> sumVec, maxMagSqrVec;
};
LayerSetup buildLayerSetupParallel(const PolyMesh &m, const Topology &t, const std::vector<int32_t> &patchLayer,
                                   int maxLayers, std::vector<double> &normals, const LayerSync &sync);

// Tiles of the fused geometry kernel (k_geom_tiles): compact groups of cells (consecutive cells of a
// space-filling-curve order, closed when a cell or face budget is reached) with the list of the faces
// their cells touch.  A thread block computes every listed face once into shared memory and then the
// centres of its cells from there, so face centres/areas never travel through HBM.  Faces on a tile
// border are listed (and computed, bit-identically) by both tiles.
struct GeomTiles
{
    int32_t nTiles = 0;
    Vec<int32_t> tileCellOff, tileCells; // cells of tile t: tileCells[tileCellOff[t] .. tileCellOff[t+1])
    Vec<int32_t> tileFaceOff, tileFaces; // its faces, ascending; bit 31: this tile stores the face's global outputs
    Vec<int32_t> slotOff;                // per cell slot (position in tileCells): offsets into slotRef
    Vec<uint16_t> slotRef;               // index of the face in the tile's list, in the order of Topology::cf; bit 15 = neighbour side
    // the points the tile's faces use (staged in shared memory before the face pass) and, per listed face,
    // its vertices as indices into that list
    Vec<int32_t> tilePointOff, tilePoints; // ascending point labels
    Vec<int32_t> faceRefOff;               // per listed face (position in tileFaces) + 1: offsets into faceRef
    Vec<uint16_t> faceRef;
    // The (edge, cell) pairs of the fused face-angle filter: for every cell slot, each edge of the cell with its
    // end points (indices into the tile's point list) and the two faces of the cell that meet at it (indices
    // into the tile's face list) -- what calcMinMaxFaceAngleForEdge (src/smoothMesh.C:1135-1231) visits for this
    // cell of the edge.  4 x uint16 per pair: p0, p1, f0, f1, cell-major through cellEdgeOff (per slot + 1);
    // uniformCellEdges > 0 when every cell has that many.  Empty when some cell is not closed (an edge not shared
    // by exactly two of its faces: the caller then keeps the per-edge kernel) and, unless asked for, on
    // all-hexahedra meshes, where hexRec replaces them.
    Vec<int32_t> cellEdgeOff;
    Vec<uint16_t> cellEdgeRef;
    // All-hexahedra meshes: the same pairs as a canonical record per cell slot, 16 x uint16:
    //   v0..v3 (vertex loop of one face A), w0..w3 (w_i = the vertex joined to v_i by an edge, on the opposite
    //   face B), A, B, S0..S3 (S_i = the side face through v_i, v_i+1, w_i+1, w_i), 2 x padding.
    // The twelve (edge, cell) pairs are then a fixed pattern: (v_i, v_i+1 | A, S_i), (w_i, w_i+1 | B, S_i),
    // (v_i, w_i | S_i-1, S_i).  Stored for the uniform tiles only (see below), per tile: the first 8 entries of all
    // its cells, then the second 8 (two conflict-free 16-byte reads per thread), at 16 * tileUCellOff[t].
    Vec<uint16_t> hexRec;
    // Per-tile fast path: a tile all of whose listed faces are quadrilaterals and all of whose cells are
    // topological hexahedra is "uniform" and reads fixed-stride copies of its references -- uFaceRef: 4 x uint16 per
    // listed face, uSlotRef: 6 x uint16 per cell, hexRec as above -- at tileUFaceOff[t] / tileUCellOff[t] (counted
    // over the uniform tiles only; -1 for the other tiles, which go through the offset tables).  A mesh of
    // hexahedra only has tileUFaceOff == tileFaceOff and tileUCellOff == tileCellOff.
    Vec<int32_t> tileUFaceOff, tileUCellOff;
    Vec<uint16_t> uFaceRef, uSlotRef;
    int64_t nUniformCells = 0;
    int32_t uniformCellEdges = 0;
    int32_t maxTileCells = 0, maxTileFaces = 0, maxTilePoints = 0, maxTileEdgePairs = 0;
};
// keepPairs: also materialise the pair lists of an all-hexahedra mesh (the kernel only reads hexRec there)
GeomTiles buildGeomTiles(const PolyMesh &m, const Topology &t, int maxCells, int maxFaces, int maxPoints, bool keepPairs = false);

// Tiles of the per-point kernels (k_predict_tiles, k_edge_tiles): compact groups of points (aligned bricks of a
// space-filling-curve order) with the list of the points their stencils read -- the tile's own points first,
// then their edge neighbours outside the tile -- and the list of the cells around them.  A thread block
// stages the positions of the listed points and the centres of the listed cells in shared memory once; the
// per-point rows become 16-bit references into those lists.
struct PointTiles
{
    int32_t nTiles = 0;
    std::vector<int32_t> ownOff;           // per tile + 1: its own points are slots ownOff[t] .. ownOff[t+1]
    std::vector<int32_t> haloOff, halo;    // per tile + 1; point labels: the tile's own points (ascending), then the halo (ascending)
    std::vector<int32_t> cellOff, cell;    // per tile + 1; cell labels, ascending
    // per own-point slot, 16 x uint16: cells[8] (pointCells row order), nbrs[6] (pointPoints row order), both as
    // indices into the tile's lists; [14] = nCells | nNbrs << 4 | generic << 15; [15] = corner-pair mask (as
    // Topology::pointRec[15]).  generic: the point takes the CSR path (valence above the record's capacity).
    std::vector<uint16_t> rec;
    int32_t maxOwn = 0, maxHalo = 0, maxCells = 0;
};
PointTiles buildPointTiles(const PolyMesh &m, const Topology &t, int maxOwn, int maxHalo, int maxCells);

// Throws std::runtime_error with the reference's FatalError texts where the
// reference would abort (empty patches :61-66, <2 eligible closest points
// :354-362, edge/cell face-pair sanity :1073,:1087).
Topology buildTopology(const PolyMesh &m);
// The 48-byte edge records (Topology::edgeRec) are only read by the per-edge face-angle filter, which the fused
// per-cell filter replaces on tiled meshes: built on first use.
void buildEdgeRecords(Topology &t);

} // namespace sm
