// polymesh.hpp -- host-side polyMesh container, file I/O, synthetic mesh
// generators and cell decomposition for the smoothMesh hot path.
//
// This is the stand-in for the part of OpenFOAM's polyMesh the reference uses
// (reading <case>/<time|constant>/polyMesh/{points,faces,owner,neighbour,boundary},
// src/smoothMesh.C:1814-1818, and writing points, :2416-2431); grammar as in
// SURVEY.md appendix A.1.
#pragma once
#include "bigvec.hpp"
#include <cstdint>
#include <string>
#include <vector>

namespace sm
{

enum PatchKind : int32_t
{
    PATCH_BOUNDARY = 0,
    PATCH_PROCESSOR = 1,
    PATCH_EMPTY = 2
};

struct Patch
{
    std::string name;
    std::string type; // "wall", "patch", "processor", "empty", ...
    int32_t start = 0, size = 0;
    int32_t myProc = -1, nbrProc = -1;
    int32_t kind() const { return type == "processor" ? PATCH_PROCESSOR : type == "empty" ? PATCH_EMPTY : PATCH_BOUNDARY; }
};

struct PolyMesh
{
    Vec<double> points;       // 3*nPoints, AoS
    Vec<int32_t> faceOffsets; // nFaces+1
    Vec<int32_t> faceVerts;
    Vec<int32_t> owner;     // nFaces
    Vec<int32_t> neighbour; // nInternalFaces
    std::vector<Patch> patches;
    int64_t nCells = 0;
    // decomposed meshes only: local -> global addressing (decomposePar's *ProcAddressing)
    std::vector<int64_t> pointGlobalId, cellGlobalId;

    int64_t nPoints() const { return (int64_t)points.size() / 3; }
    int64_t nFaces() const { return (int64_t)owner.size(); }
    int64_t nInternalFaces() const { return (int64_t)neighbour.size(); }
    // throws std::runtime_error describing the first violated polyMesh invariant
    void check() const;
};

// ---- generators -------------------------------------------------------------
// Structured hex block on [lo,hi]^3 with blockMesh's single-block numbering
// (point i + j(nx+1) + k(nx+1)(ny+1), cell i + j nx + k nx ny); six patches
// xMin,xMax,yMin,yMax,zMin,zMax of type `patchType`.
PolyMesh genHexBlock(int nx, int ny, int nz, const double lo[3], const double hi[3], const std::string &patchType = "wall");

// One brick (rank = ix + px*(iy + py*iz)) of the (nx*px) x (ny*py) x (nz*pz) block, in processor-mesh
// form (processor patches, pointGlobalId / cellGlobalId), generated locally for weak-scaling runs.
PolyMesh genHexBlockPart(int nx, int ny, int nz, int px, int py, int pz, int rank, const double lo[3], const double hi[3]);

// Kelvin-cell (truncated octahedron, BCC Voronoi) polyhedral mesh clipped to the
// box [0,n]^3*h: stand-in for polyDualMesh output.
PolyMesh genKelvin(int n, double h);
// One brick of that mesh (rank = ix + px*(iy + py*iz) of a px x py x pz split of the lattice) in processor-mesh
// form, generated locally (config 4 at its stated size never exists on one host).
PolyMesh genKelvinPart(int n, double h, int px, int py, int pz, int rank);

// Generic builder: cells given as lists of outward-oriented faces (vertex
// loops).  cellFaceOffsets[C+1] indexes faces; cfVertOffsets[NF+1] / cfVerts
// give each cell-face's loop; cfPatch[NF] is the patch id a face gets if it
// turns out to be a boundary face.  Produces a valid polyMesh (upper-triangular
// internal face order, owner<neighbour, boundary faces patch by patch).
PolyMesh buildFromCells(const std::vector<double> &points, const std::vector<int32_t> &cellFaceOffsets,
                        const std::vector<int32_t> &cfVertOffsets, const std::vector<int32_t> &cfVerts,
                        const std::vector<int32_t> &cfPatch, const std::vector<std::string> &patchNames,
                        const std::vector<std::string> &patchTypes);

// Displace every point that is not on a non-processor patch by i.i.d.
// U(-amp,amp) per component; counter-based RNG keyed on (seed, global point
// label, component) so every partition generates identical values.
void jitterInterior(PolyMesh &m, double amp, uint64_t seed);
double counterUniform(uint64_t seed, uint64_t label, uint32_t comp); // in [0,1)

// Morton (space-filling-curve) renumbering of points and cells (renumberMesh stand-in); the maps give the
// old label of every new label.
PolyMesh renumberMorton(const PolyMesh &m, std::vector<int32_t> &pointOldOfNew, std::vector<int32_t> &cellOldOfNew);

// checkMesh-style quality figures (acceptance criteria of BASELINE.json: max non-orthogonality, skewness,
// min angle); host-side, evaluated on request only.
struct MeshQuality
{
    double maxNonOrtho = 0, avgNonOrtho = 0; // degrees, internal faces
    double maxSkewness = 0;                  // internal and boundary faces
    double minEdgeAngle = 0;                 // degrees, smallest angle between consecutive face edges
    double minEdgeLength = 0, maxEdgeLength = 0, minVolume = 0;
};
MeshQuality computeQuality(const PolyMesh &m);

// ---- decomposition (decomposePar stand-in) ------------------------------------
// cellPart[c] in [0,nParts).  Produces OpenFOAM-style processor meshes: local
// points/cells/faces in ascending global order, inter-part faces become
// processor patches (one per neighbour part, ascending neighbour, faces in
// ascending global face order, reversed on the neighbour side), original
// patches kept (possibly empty) in front.
std::vector<PolyMesh> decompose(const PolyMesh &m, const std::vector<int32_t> &cellPart, int nParts);
// simple geometric partitioners
std::vector<int32_t> partitionBricks(const PolyMesh &m, int px, int py, int pz);
std::vector<int32_t> partitionRCB(const PolyMesh &m, int nParts);

// ---- file I/O -----------------------------------------------------------------
// dir = ".../polyMesh".  Reads ascii or binary (label=32, scalar=64) files.
PolyMesh readPolyMesh(const std::string &dir);
void writePolyMesh(const PolyMesh &m, const std::string &dir, bool binary = false, int precision = 17);
// Writes only the points file (what mesh.write() does after movePoints).
void writePoints(const double *pts, int64_t nPoints, const std::string &dir, bool binary, int precision,
                 const std::string &location);
std::vector<double> readPoints(const std::string &file);
// labelIOList files written next to the mesh (isCornerPoint / isFeatureEdgePoint)
void writeLabelIOList(const std::string &file, const std::string &object, const std::string &location,
                      const std::vector<int32_t> &v, bool binary);
std::vector<int32_t> readLabelIOList(const std::string &file);
// decomposed cases: processor<k>/constant/polyMesh with pointProcAddressing / cellProcAddressing
void writeDecomposedCase(const std::vector<PolyMesh> &parts, const std::string &caseDir, bool binary);
PolyMesh readProcessorMesh(const std::string &caseDir, int k);

} // namespace sm
