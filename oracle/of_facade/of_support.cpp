// of_support.cpp -- TEST INFRASTRUCTURE ONLY.  Second translation unit of oracle/_ref/smoothMesh_ref: mesh
// files (product reader / writer) and oracle.cpp's Rank for the addressing and geometry OpenFOAM would
// provide [OF-recalled].  See FacadeSupport.H for why this is kept apart from the reference's translation unit.
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

#include "../../smoothmesh_b200/csrc/sm_math.h"
#define ORACLE_LIBM_ACOS 1
namespace orc
{
#include "../oracle.cpp"
}
#include "FacadeSupport.H"
#include "polymesh.hpp"

static orc::Rank &rankOf(FacadeAddressing &A) { return *static_cast<orc::Rank *>(A.impl); }

void facadeUpdateGeometry(FacadeAddressing &A)
{
    orc::Rank &R = rankOf(A);
    R.calcGeometry();
    A.cellCentres.resize(3 * (size_t)R.C);
    for (int c = 0; c < R.C; ++c)
        A.cellCentres[3 * c] = R.cellCtr[c].x, A.cellCentres[3 * c + 1] = R.cellCtr[c].y, A.cellCentres[3 * c + 2] = R.cellCtr[c].z;
    A.faceCentres.resize(3 * (size_t)R.F);
    A.faceAreas.resize(3 * (size_t)R.F);
    for (int f = 0; f < R.F; ++f)
    {
        A.faceCentres[3 * f] = R.faceCtr[f].x, A.faceCentres[3 * f + 1] = R.faceCtr[f].y, A.faceCentres[3 * f + 2] = R.faceCtr[f].z;
        A.faceAreas[3 * f] = R.faceArea[f].x, A.faceAreas[3 * f + 1] = R.faceArea[f].y, A.faceAreas[3 * f + 2] = R.faceArea[f].z;
    }
}

void facadeMovePoints(FacadeAddressing &A, const std::vector<double> &points)
{
    orc::Rank &R = rankOf(A);
    A.points = points;
    for (int i = 0; i < R.P; ++i)
        R.pts[i] = {points[3 * i], points[3 * i + 1], points[3 * i + 2]};
    R.geomValid = false;
}

std::string facadeLoadMesh(const std::string &topoDir, const std::string &pointsFile, FacadeAddressing &A)
{
    try
    {
        sm::PolyMesh pm = sm::readPolyMesh(topoDir);
        if (!pointsFile.empty())
        {
            const std::vector<double> pts = sm::readPoints(pointsFile);
            pm.points.assign(pts.begin(), pts.end());
        }
        for (const sm::Patch &p : pm.patches)
        {
            A.patchStart.push_back(p.start);
            A.patchSize.push_back(p.size);
            A.patchKind.push_back(p.kind());
            A.patchName.push_back(p.name);
            A.patchType.push_back(p.type);
        }
        orc::MeshIn in = {};
        in.P = pm.nPoints();
        in.C = pm.nCells;
        in.F = pm.nFaces();
        in.Fi = pm.nInternalFaces();
        in.pts = pm.points.data();
        in.fOff = pm.faceOffsets.data();
        in.fV = pm.faceVerts.data();
        in.own = pm.owner.data();
        in.nei = pm.neighbour.data();
        in.nPatches = (int32_t)A.patchStart.size();
        in.pStart = A.patchStart.data();
        in.pSize = A.patchSize.data();
        in.pKind = A.patchKind.data();
        in.pointGlobalId = nullptr;
        in.pLayer = nullptr;
        orc::Params prm;
        memset(&prm, 0, sizeof prm);
        orc::Rank *R = new orc::Rank;
        A.impl = R;
        if (!R->init(in, prm))
            return R->err;
        A.P = R->P, A.C = R->C, A.F = R->F, A.Fi = R->Fi;
        A.points.assign(pm.points.begin(), pm.points.end());
        A.faceOff = R->fOff;
        A.faceVerts = R->fV;
        A.owner = R->own;
        A.neighbour = R->nei;
        for (auto &e : R->edges)
        {
            A.edgeA.push_back(e[0]);
            A.edgeB.push_back(e[1]);
        }
        A.pointPoints = R->pointPoints;
        A.pointCells = R->pointCells;
        A.pointFaces = R->pointFaces;
        A.pointEdges = R->pointEdges;
        A.edgeFaces = R->edgeFaces;
        A.edgeCells = R->edgeCells;
        A.cellPoints = R->cellPoints;
    }
    catch (const std::exception &e)
    {
        return e.what();
    }
    return "";
}

std::string facadeLoadGlobalIds(const std::string &rootCaseDir, int rank, FacadeAddressing &A)
{
    try
    {
        const sm::PolyMesh pm = sm::readProcessorMesh(rootCaseDir, rank);
        A.pointGlobalId = pm.pointGlobalId;
    }
    catch (const std::exception &e)
    {
        return e.what();
    }
    return "";
}

void facadeWritePoints(const FacadeAddressing &A, const std::string &dir, bool binary, int precision, const std::string &location)
{
    sm::writePoints(A.points.data(), A.P, dir, binary, precision, location);
}

namespace Foam
{
bool facadeReadLabels(const std::string &file, std::vector<int> &out)
{
    FILE *f = fopen(file.c_str(), "rb");
    if (!f)
        return false;
    fclose(f);
    try
    {
        const std::vector<int32_t> v = sm::readLabelIOList(file);
        out.assign(v.begin(), v.end());
    }
    catch (const std::exception &)
    {
        return false;
    }
    return true;
}
void facadeWriteLabels(const std::string &file, const std::string &object, const std::string &location, const std::vector<int> &v,
                       bool binary)
{
    sm::writeLabelIOList(file, object, location, std::vector<int32_t>(v.begin(), v.end()), binary);
}
} // namespace Foam
