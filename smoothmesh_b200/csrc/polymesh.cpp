// polymesh.cpp -- see polymesh.hpp.
#include "polymesh.hpp"

#include <algorithm>
#include <array>
#include <parallel/algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <numeric>
#include <sstream>
#include <stdexcept>
#include <unordered_map>

namespace sm
{

template <class T, class V> static void assignVec(Vec<T> &dst, const V &src) { dst.assign(src.begin(), src.end()); }


static void fail(const std::string &s) { throw std::runtime_error(s); }

// ----------------------------------------------------------------- check ----
void PolyMesh::check() const
{
    const int64_t P = nPoints(), F = nFaces(), Fi = nInternalFaces();
    if ((int64_t)faceOffsets.size() != F + 1)
        fail("faceOffsets size != nFaces+1");
    if (Fi > F)
        fail("more neighbours than faces");
    if (faceOffsets[0] != 0 || (int64_t)faceVerts.size() != faceOffsets[F])
        fail("faceVerts size mismatch");
    // the first violation in label order, found by all threads (meshes of 10^8 faces are checked on every create)
    const int64_t none = INT64_MAX;
    int64_t badFace = none, badVert = none, badNbr = none;
#pragma omp parallel for schedule(static) reduction(min : badFace)
    for (int64_t f = 0; f < F; ++f)
        if (faceOffsets[f + 1] - faceOffsets[f] < 3 || owner[f] < 0 || owner[f] >= nCells)
            badFace = std::min(badFace, f);
    if (badFace != none)
    {
        if (faceOffsets[badFace + 1] - faceOffsets[badFace] < 3)
            fail("face " + std::to_string(badFace) + " has fewer than 3 vertices");
        fail("owner out of range at face " + std::to_string(badFace));
    }
    const int64_t FV = (int64_t)faceVerts.size();
#pragma omp parallel for schedule(static) reduction(min : badVert)
    for (int64_t k = 0; k < FV; ++k)
        if (faceVerts[k] < 0 || faceVerts[k] >= P)
            badVert = std::min(badVert, k);
    if (badVert != none)
        fail("face vertex label out of range");
    auto nbrRange = [&](int64_t f) { return neighbour[f] <= owner[f] || neighbour[f] >= nCells; };
    auto nbrOrder = [&](int64_t f) {
        return f > 0 && (owner[f] < owner[f - 1] || (owner[f] == owner[f - 1] && neighbour[f] < neighbour[f - 1]));
    };
#pragma omp parallel for schedule(static) reduction(min : badNbr)
    for (int64_t f = 0; f < Fi; ++f)
        if (nbrRange(f) || nbrOrder(f))
            badNbr = std::min(badNbr, f);
    if (badNbr != none)
    {
        if (nbrRange(badNbr))
            fail("neighbour <= owner or out of range at face " + std::to_string(badNbr));
        fail("internal faces not in upper-triangular order at face " + std::to_string(badNbr));
    }
    int64_t next = Fi;
    for (const Patch &p : patches)
    {
        if (p.start != next)
            fail("patch " + p.name + " does not start where the previous one ends");
        next += p.size;
    }
    if (next != F)
        fail("patches do not cover the boundary faces");
}

// ------------------------------------------------------------- hex block ----
PolyMesh genHexBlock(int nx, int ny, int nz, const double lo[3], const double hi[3], const std::string &patchType)
{
    if (nx < 1 || ny < 1 || nz < 1)
        fail("genHexBlock: need at least one cell per direction");
    PolyMesh m;
    const int64_t px = nx + 1, py = ny + 1, pz = nz + 1;
    const int64_t P = px * py * pz, C = (int64_t)nx * ny * nz;
    const int64_t Fi = (int64_t)(nx - 1) * ny * nz + (int64_t)nx * (ny - 1) * nz + (int64_t)nx * ny * (nz - 1);
    const int64_t Fb = 2 * ((int64_t)ny * nz + (int64_t)nx * nz + (int64_t)nx * ny);
    if (4 * (Fi + Fb) >= (int64_t)INT32_MAX)
        fail("genHexBlock: mesh too large for 32-bit labels");
    m.nCells = C;
    m.points.resize(3 * P);
    for (int64_t k = 0; k < pz; ++k)
        for (int64_t j = 0; j < py; ++j)
            for (int64_t i = 0; i < px; ++i)
            {
                const int64_t p = i + j * px + k * px * py;
                m.points[3 * p + 0] = lo[0] + (hi[0] - lo[0]) * (double(i) / double(nx));
                m.points[3 * p + 1] = lo[1] + (hi[1] - lo[1]) * (double(j) / double(ny));
                m.points[3 * p + 2] = lo[2] + (hi[2] - lo[2]) * (double(k) / double(nz));
            }
    auto pid = [&](int64_t i, int64_t j, int64_t k) { return (int32_t)(i + j * px + k * px * py); };
    auto cid = [&](int64_t i, int64_t j, int64_t k) { return (int32_t)(i + j * nx + k * (int64_t)nx * ny); };
    m.faceOffsets.reserve(Fi + Fb + 1);
    m.faceVerts.reserve(4 * (Fi + Fb));
    m.owner.reserve(Fi + Fb);
    m.neighbour.reserve(Fi);
    m.faceOffsets.push_back(0);
    auto addFace = [&](int32_t a, int32_t b, int32_t c, int32_t d, int32_t own) {
        m.faceVerts.push_back(a);
        m.faceVerts.push_back(b);
        m.faceVerts.push_back(c);
        m.faceVerts.push_back(d);
        m.faceOffsets.push_back((int32_t)m.faceVerts.size());
        m.owner.push_back(own);
    };
    // faces normal to x / y / z at the upper side of cell (i,j,k), normal pointing +axis
    auto xFace = [&](int64_t i, int64_t j, int64_t k, int32_t own, bool flip) {
        const int32_t a = pid(i, j, k), b = pid(i, j + 1, k), c = pid(i, j + 1, k + 1), d = pid(i, j, k + 1);
        flip ? addFace(a, d, c, b, own) : addFace(a, b, c, d, own);
    };
    auto yFace = [&](int64_t i, int64_t j, int64_t k, int32_t own, bool flip) {
        const int32_t a = pid(i, j, k), b = pid(i, j, k + 1), c = pid(i + 1, j, k + 1), d = pid(i + 1, j, k);
        flip ? addFace(a, d, c, b, own) : addFace(a, b, c, d, own);
    };
    auto zFace = [&](int64_t i, int64_t j, int64_t k, int32_t own, bool flip) {
        const int32_t a = pid(i, j, k), b = pid(i + 1, j, k), c = pid(i + 1, j + 1, k), d = pid(i, j + 1, k);
        flip ? addFace(a, d, c, b, own) : addFace(a, b, c, d, own);
    };
    // internal faces, upper-triangular: per cell its +x, +y, +z neighbours (ascending labels)
    for (int64_t k = 0; k < nz; ++k)
        for (int64_t j = 0; j < ny; ++j)
            for (int64_t i = 0; i < nx; ++i)
            {
                const int32_t c = cid(i, j, k);
                if (i + 1 < nx)
                {
                    xFace(i + 1, j, k, c, false);
                    m.neighbour.push_back(cid(i + 1, j, k));
                }
                if (j + 1 < ny)
                {
                    yFace(i, j + 1, k, c, false);
                    m.neighbour.push_back(cid(i, j + 1, k));
                }
                if (k + 1 < nz)
                {
                    zFace(i, j, k + 1, c, false);
                    m.neighbour.push_back(cid(i, j, k + 1));
                }
            }
    auto beginPatch = [&](const char *name) {
        Patch p;
        p.name = name;
        p.type = patchType;
        p.start = (int32_t)m.owner.size();
        m.patches.push_back(p);
    };
    auto endPatch = [&]() { m.patches.back().size = (int32_t)m.owner.size() - m.patches.back().start; };
    beginPatch("xMin");
    for (int64_t k = 0; k < nz; ++k)
        for (int64_t j = 0; j < ny; ++j)
            xFace(0, j, k, cid(0, j, k), true);
    endPatch();
    beginPatch("xMax");
    for (int64_t k = 0; k < nz; ++k)
        for (int64_t j = 0; j < ny; ++j)
            xFace(nx, j, k, cid(nx - 1, j, k), false);
    endPatch();
    beginPatch("yMin");
    for (int64_t k = 0; k < nz; ++k)
        for (int64_t i = 0; i < nx; ++i)
            yFace(i, 0, k, cid(i, 0, k), true);
    endPatch();
    beginPatch("yMax");
    for (int64_t k = 0; k < nz; ++k)
        for (int64_t i = 0; i < nx; ++i)
            yFace(i, ny, k, cid(i, ny - 1, k), false);
    endPatch();
    beginPatch("zMin");
    for (int64_t j = 0; j < ny; ++j)
        for (int64_t i = 0; i < nx; ++i)
            zFace(i, j, 0, cid(i, j, 0), true);
    endPatch();
    beginPatch("zMax");
    for (int64_t j = 0; j < ny; ++j)
        for (int64_t i = 0; i < nx; ++i)
            zFace(i, j, nz, cid(i, j, nz - 1), false);
    endPatch();
    return m;
}

// --------------------------------------------------------- generic build ----
namespace
{
struct FaceRec
{
    uint64_t key;   // hash of the sorted vertex list
    int32_t cell;   // owning cell of this copy
    int32_t cf;     // index into the cell-face arrays
};
inline uint64_t mix64(uint64_t x)
{
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
} // namespace

PolyMesh buildFromCells(const std::vector<double> &points, const std::vector<int32_t> &cellFaceOffsets,
                        const std::vector<int32_t> &cfVertOffsets, const std::vector<int32_t> &cfVerts,
                        const std::vector<int32_t> &cfPatch, const std::vector<std::string> &patchNames,
                        const std::vector<std::string> &patchTypes)
{
    const int32_t C = (int32_t)cellFaceOffsets.size() - 1;
    const int32_t NF = (int32_t)cfVertOffsets.size() - 1;
    std::vector<FaceRec> recs(NF);
    std::vector<int32_t> tmp;
    auto sortedVerts = [&](int32_t cf, std::vector<int32_t> &out) {
        out.assign(cfVerts.begin() + cfVertOffsets[cf], cfVerts.begin() + cfVertOffsets[cf + 1]);
        std::sort(out.begin(), out.end());
    };
#pragma omp parallel for schedule(static)
    for (int32_t c = 0; c < C; ++c)
    {
        std::vector<int32_t> sv;
        for (int32_t cf = cellFaceOffsets[c]; cf < cellFaceOffsets[c + 1]; ++cf)
        {
            sv.assign(cfVerts.begin() + cfVertOffsets[cf], cfVerts.begin() + cfVertOffsets[cf + 1]);
            std::sort(sv.begin(), sv.end());
            uint64_t h = 0x1234567ull + sv.size();
            for (int32_t v : sv)
                h = mix64(h ^ (uint64_t)(uint32_t)v);
            recs[cf] = {h, c, cf};
        }
    }
    __gnu_parallel::sort(recs.begin(), recs.end(), [](const FaceRec &a, const FaceRec &b) {
        return a.key != b.key ? a.key < b.key : a.cell < b.cell;
    });
    struct Out
    {
        int32_t own, nei, cf, patch;
    };
    std::vector<Out> internal, boundary;
    std::vector<int32_t> t2;
    for (int32_t i = 0; i < NF;)
    {
        int32_t j = i + 1;
        while (j < NF && recs[j].key == recs[i].key)
            ++j;
        if (j - i == 1)
            boundary.push_back({recs[i].cell, -1, recs[i].cf, cfPatch.empty() ? 0 : cfPatch[recs[i].cf]});
        else if (j - i == 2)
        {
            sortedVerts(recs[i].cf, tmp);
            sortedVerts(recs[i + 1].cf, t2);
            if (tmp != t2)
                fail("buildFromCells: face hash collision");
            if (recs[i].cell == recs[i + 1].cell)
                fail("buildFromCells: a cell uses the same face twice");
            internal.push_back({recs[i].cell, recs[i + 1].cell, recs[i].cf, -1});
        }
        else
            fail("buildFromCells: face shared by more than two cells (or hash collision)");
        i = j;
    }
    __gnu_parallel::sort(internal.begin(), internal.end(),
                         [](const Out &a, const Out &b) { return a.own != b.own ? a.own < b.own : (a.nei != b.nei ? a.nei < b.nei : a.cf < b.cf); });
    std::stable_sort(boundary.begin(), boundary.end(), [](const Out &a, const Out &b) {
        return a.patch != b.patch ? a.patch < b.patch : (a.own != b.own ? a.own < b.own : a.cf < b.cf);
    });
    PolyMesh m;
    m.points.assign(points.begin(), points.end());
    m.nCells = C;
    m.faceOffsets.push_back(0);
    auto emit = [&](const Out &o) {
        for (int32_t k = cfVertOffsets[o.cf]; k < cfVertOffsets[o.cf + 1]; ++k)
            m.faceVerts.push_back(cfVerts[k]);
        m.faceOffsets.push_back((int32_t)m.faceVerts.size());
        m.owner.push_back(o.own);
    };
    for (const Out &o : internal)
    {
        emit(o);
        m.neighbour.push_back(o.nei);
    }
    const int32_t nPatches = (int32_t)std::max<size_t>(patchNames.size(), 1);
    size_t b = 0;
    for (int32_t p = 0; p < nPatches; ++p)
    {
        Patch pt;
        pt.name = patchNames.empty() ? "walls" : patchNames[p];
        pt.type = patchTypes.empty() ? "wall" : patchTypes[p];
        pt.start = (int32_t)m.owner.size();
        while (b < boundary.size() && boundary[b].patch == p)
            emit(boundary[b++]);
        pt.size = (int32_t)m.owner.size() - pt.start;
        m.patches.push_back(pt);
    }
    if (b != boundary.size())
        fail("buildFromCells: boundary face with patch id out of range");
    return m;
}

// ------------------------------------------------------------- Kelvin mesh ----
// face templates of a truncated octahedron relative to its centre (integer lattice units): 6 squares + 8
// hexagons, outward oriented
static std::vector<std::vector<std::array<int, 3>>> kelvinFaceTemplates()
{
    std::vector<std::vector<std::array<int, 3>>> tmpl;
    auto orient = [&](std::vector<std::array<int, 3>> loop) {
        // order by angle around the centroid, then make the normal point away from the origin
        double cx = 0, cy = 0, cz = 0;
        for (auto &v : loop)
        {
            cx += v[0];
            cy += v[1];
            cz += v[2];
        }
        cx /= loop.size();
        cy /= loop.size();
        cz /= loop.size();
        const double nl = std::sqrt(cx * cx + cy * cy + cz * cz);
        const double nx = cx / nl, ny = cy / nl, nz = cz / nl;
        double ux = loop[0][0] - cx, uy = loop[0][1] - cy, uz = loop[0][2] - cz;
        const double ul = std::sqrt(ux * ux + uy * uy + uz * uz);
        ux /= ul;
        uy /= ul;
        uz /= ul;
        const double wx = ny * uz - nz * uy, wy = nz * ux - nx * uz, wz = nx * uy - ny * ux;
        std::sort(loop.begin(), loop.end(), [&](const std::array<int, 3> &a, const std::array<int, 3> &b) {
            const double aa = std::atan2((a[0] - cx) * wx + (a[1] - cy) * wy + (a[2] - cz) * wz,
                                         (a[0] - cx) * ux + (a[1] - cy) * uy + (a[2] - cz) * uz);
            const double bb = std::atan2((b[0] - cx) * wx + (b[1] - cy) * wy + (b[2] - cz) * wz,
                                         (b[0] - cx) * ux + (b[1] - cy) * uy + (b[2] - cz) * uz);
            return aa < bb;
        });
        return loop; // counter-clockwise seen from outside (right-hand rule about the outward normal)
    };
    for (int axis = 0; axis < 3; ++axis)
        for (int s = -1; s <= 1; s += 2)
        {
            std::vector<std::array<int, 3>> loop;
            const int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
            for (int t = 0; t < 4; ++t)
            {
                std::array<int, 3> v = {0, 0, 0};
                v[axis] = 2 * s;
                v[(t % 2 == 0) ? a1 : a2] = (t < 2) ? 1 : -1;
                loop.push_back(v);
            }
            tmpl.push_back(orient(loop));
        }
    for (int sx = -1; sx <= 1; sx += 2)
        for (int sy = -1; sy <= 1; sy += 2)
            for (int sz = -1; sz <= 1; sz += 2)
            {
                std::vector<std::array<int, 3>> loop;
                int perm[3] = {0, 1, 2};
                do
                    loop.push_back({sx * perm[0], sy * perm[1], sz * perm[2]});
                while (std::next_permutation(perm, perm + 3));
                tmpl.push_back(orient(loop));
            }
    return tmpl;
}

PolyMesh genKelvin(int n, double h)
{
    // BCC lattice with lattice constant 4 in integer units (vertex coordinates are
    // integers); cell centres at (4i,4j,4k) and (4i+2,4j+2,4k+2).  A truncated
    // octahedron around c has the 24 vertices c + perm(0,+-1,+-2).
    if (n < 1)
        fail("genKelvin: n < 1");
    std::vector<std::array<int, 3>> centres;
    for (int k = 0; k < n; ++k)
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i)
            {
                centres.push_back({4 * i, 4 * j, 4 * k});
                centres.push_back({4 * i + 2, 4 * j + 2, 4 * k + 2});
            }
    std::sort(centres.begin(), centres.end(), [](const std::array<int, 3> &a, const std::array<int, 3> &b) {
        return a[2] != b[2] ? a[2] < b[2] : (a[1] != b[1] ? a[1] < b[1] : a[0] < b[0]);
    });
    const std::vector<std::vector<std::array<int, 3>>> tmpl = kelvinFaceTemplates();
    // vertices: integer triples -> ids in lexicographic (z,y,x) order, through a dense lattice index
    const int64_t G = 4 * (int64_t)n + 6; // coordinates lie in [-2, 4n+3]
    auto slot = [&](int x, int y, int z) { return ((int64_t)(z + 2) * G + (y + 2)) * G + (x + 2); };
    std::vector<int32_t> vid((size_t)(G * G * G), -1);
    for (auto &c : centres)
        for (auto &f : tmpl)
            for (auto &v : f)
                vid[slot(c[0] + v[0], c[1] + v[1], c[2] + v[2])] = 0;
    int32_t nv = 0;
    std::vector<double> pts;
    for (int64_t z = 0; z < G; ++z)
        for (int64_t y = 0; y < G; ++y)
            for (int64_t x = 0; x < G; ++x)
            {
                int32_t &id = vid[(size_t)((z * G + y) * G + x)];
                if (id == 0)
                {
                    id = nv++;
                    pts.push_back((x - 2) * h / 4.0);
                    pts.push_back((y - 2) * h / 4.0);
                    pts.push_back((z - 2) * h / 4.0);
                }
            }
    std::vector<int32_t> cfo{0}, cvo{0}, cv, cp;
    cv.reserve(centres.size() * 72);
    for (auto &c : centres)
    {
        for (auto &f : tmpl)
        {
            for (auto &v : f)
                cv.push_back(vid[slot(c[0] + v[0], c[1] + v[1], c[2] + v[2])]);
            cvo.push_back((int32_t)cv.size());
            cp.push_back(0);
        }
        cfo.push_back((int32_t)cvo.size() - 1);
    }
    return buildFromCells(pts, cfo, cvo, cv, cp, {"walls"}, {"wall"});
}

// One brick of the Kelvin mesh in processor-mesh form, generated locally: the lattice cells
// [i0,i1) x [j0,j1) x [k0,k1) of an n^3 lattice split px x py x pz ways (both truncated octahedra of a
// lattice cell belong to it).  Faces towards cells of other bricks become processor patches (one per
// neighbour rank, ascending); pointGlobalId is the vertex's slot in the global integer lattice, which every
// rank computes alike, so shared points match and the jitter (keyed on it) is identical on all copies.
PolyMesh genKelvinPart(int n, double h, int px, int py, int pz, int rank)
{
    if (n < 1 || px < 1 || py < 1 || pz < 1 || rank < 0 || rank >= px * py * pz)
        fail("genKelvinPart: bad arguments");
    const int parts[3] = {px, py, pz};
    const int mine[3] = {rank % px, (rank / px) % py, rank / (px * py)};
    auto lower = [&](int d, int q) { return (int)((int64_t)n * q / parts[d]); };
    int lo[3], hi[3];
    for (int d = 0; d < 3; ++d)
    {
        lo[d] = lower(d, mine[d]);
        hi[d] = lower(d, mine[d] + 1);
        if (hi[d] <= lo[d])
            fail("genKelvinPart: more bricks than lattice cells along an axis");
    }
    auto brickOf = [&](int d, int i) { // brick index along axis d of lattice cell i
        int q = (int)(((int64_t)i * parts[d]) / n);
        while (q + 1 < parts[d] && lower(d, q + 1) <= i)
            ++q;
        while (q > 0 && lower(d, q) > i)
            --q;
        return q;
    };
    // rank of the cell with centre c, -1 if there is no such cell
    auto rankOfCentre = [&](const int c[3]) {
        const int r0 = ((c[0] % 4) + 4) % 4;
        if ((r0 != 0 && r0 != 2) || ((c[1] % 4) + 4) % 4 != r0 || ((c[2] % 4) + 4) % 4 != r0)
            return -1;
        int q[3];
        for (int d = 0; d < 3; ++d)
        {
            if (c[d] - r0 < 0)
                return -1;
            const int i = (c[d] - r0) / 4;
            if (i >= n)
                return -1;
            q[d] = brickOf(d, i);
        }
        return q[0] + px * (q[1] + py * q[2]);
    };
    std::vector<std::array<int, 3>> centres;
    for (int k = lo[2]; k < hi[2]; ++k)
        for (int j = lo[1]; j < hi[1]; ++j)
            for (int i = lo[0]; i < hi[0]; ++i)
            {
                centres.push_back({4 * i, 4 * j, 4 * k});
                centres.push_back({4 * i + 2, 4 * j + 2, 4 * k + 2});
            }
    std::sort(centres.begin(), centres.end(), [](const std::array<int, 3> &a, const std::array<int, 3> &b) {
        return a[2] != b[2] ? a[2] < b[2] : (a[1] != b[1] ? a[1] < b[1] : a[0] < b[0]);
    });
    const std::vector<std::vector<std::array<int, 3>>> tmpl = kelvinFaceTemplates();
    // local dense lattice over the brick's bounding box
    int64_t base[3], ext[3];
    for (int d = 0; d < 3; ++d)
    {
        base[d] = 4 * (int64_t)lo[d] - 2;
        ext[d] = 4 * (int64_t)(hi[d] - lo[d]) + 6;
    }
    auto slot = [&](int x, int y, int z) { return ((z - base[2]) * ext[1] + (y - base[1])) * ext[0] + (x - base[0]); };
    std::vector<int32_t> vid((size_t)(ext[0] * ext[1] * ext[2]), -1);
    for (auto &c : centres)
        for (auto &f : tmpl)
            for (auto &v : f)
                vid[slot(c[0] + v[0], c[1] + v[1], c[2] + v[2])] = 0;
    const int64_t G = 4 * (int64_t)n + 6;
    int32_t nv = 0;
    std::vector<double> pts;
    std::vector<int64_t> gid;
    for (int64_t z = 0; z < ext[2]; ++z)
        for (int64_t y = 0; y < ext[1]; ++y)
            for (int64_t x = 0; x < ext[0]; ++x)
            {
                int32_t &id = vid[(size_t)((z * ext[1] + y) * ext[0] + x)];
                if (id == 0)
                {
                    id = nv++;
                    const int64_t gx = x + base[0], gy = y + base[1], gz = z + base[2];
                    pts.push_back(gx * h / 4.0);
                    pts.push_back(gy * h / 4.0);
                    pts.push_back(gz * h / 4.0);
                    gid.push_back(((gz + 2) * G + (gy + 2)) * G + (gx + 2));
                }
            }
    // neighbour ranks -> patch ids (0 = walls, then ascending neighbour rank)
    std::map<int, int> patchOfRank;
    std::vector<int32_t> cfo{0}, cvo{0}, cv, cp;
    std::vector<int> cpRank; // per cell face: -1 wall, own rank = internal candidate, else neighbour rank
    cv.reserve(centres.size() * 72);
    for (auto &c : centres)
    {
        for (auto &f : tmpl)
        {
            int sum[3] = {0, 0, 0};
            for (auto &v : f)
            {
                cv.push_back(vid[slot(c[0] + v[0], c[1] + v[1], c[2] + v[2])]);
                for (int d = 0; d < 3; ++d)
                    sum[d] += v[d];
            }
            cvo.push_back((int32_t)cv.size());
            // the face centroid is half way to the neighbour's centre
            const int nb[3] = {c[0] + 2 * sum[0] / (int)f.size(), c[1] + 2 * sum[1] / (int)f.size(), c[2] + 2 * sum[2] / (int)f.size()};
            const int r = rankOfCentre(nb);
            cpRank.push_back(r);
            if (r >= 0 && r != rank)
                patchOfRank[r] = 0;
        }
        cfo.push_back((int32_t)cvo.size() - 1);
    }
    std::vector<std::string> names{"walls"}, types{"wall"};
    for (auto &kv : patchOfRank)
    {
        kv.second = (int)names.size();
        names.push_back("procBoundary" + std::to_string(rank) + "to" + std::to_string(kv.first));
        types.push_back("processor");
    }
    cp.resize(cpRank.size());
    for (size_t i = 0; i < cpRank.size(); ++i)
        cp[i] = (cpRank[i] < 0 || cpRank[i] == rank) ? 0 : patchOfRank[cpRank[i]];
    PolyMesh m = buildFromCells(pts, cfo, cvo, cv, cp, names, types);
    // a face towards a cell of this brick must have been paired: a wall patch face whose neighbour exists here
    // would mean the templates and the neighbour rule disagree
    for (auto &kv : patchOfRank)
    {
        m.patches[kv.second].myProc = rank;
        m.patches[kv.second].nbrProc = kv.first;
    }
    m.pointGlobalId = gid;
    m.cellGlobalId.resize(centres.size());
    for (size_t i = 0; i < centres.size(); ++i)
    {
        const int b = centres[i][0] % 4 == 0 ? 0 : 1;
        const int64_t ci = (centres[i][0] - 2 * b) / 4, cj = (centres[i][1] - 2 * b) / 4, ck = (centres[i][2] - 2 * b) / 4;
        m.cellGlobalId[i] = 2 * ((ck * n + cj) * n + ci) + b;
    }
    return m;
}

// ------------------------------------------------------------------ jitter ----
double counterUniform(uint64_t seed, uint64_t label, uint32_t comp)
{
    const uint64_t r = mix64(mix64(seed ^ 0x5851f42d4c957f2dull) ^ mix64(label * 3ull + comp));
    return (double)(r >> 11) * (1.0 / 9007199254740992.0); // 53 bits -> [0,1)
}

void jitterInterior(PolyMesh &m, double amp, uint64_t seed)
{
    const int64_t P = m.nPoints();
    std::vector<uint8_t> onBoundary(P, 0);
    for (const Patch &p : m.patches)
    {
        if (p.kind() == PATCH_PROCESSOR)
            continue;
        for (int32_t f = p.start; f < p.start + p.size; ++f)
            for (int32_t k = m.faceOffsets[f]; k < m.faceOffsets[f + 1]; ++k)
                onBoundary[m.faceVerts[k]] = 1;
    }
    for (int64_t p = 0; p < P; ++p)
    {
        if (onBoundary[p])
            continue;
        const uint64_t g = m.pointGlobalId.empty() ? (uint64_t)p : (uint64_t)m.pointGlobalId[p];
        for (uint32_t c = 0; c < 3; ++c)
            m.points[3 * p + c] += amp * (2.0 * counterUniform(seed, g, c) - 1.0);
    }
}

// ------------------------------------------------------------- decompose ----
std::vector<int32_t> partitionBricks(const PolyMesh &m, int px, int py, int pz)
{
    // geometric bricks on cell centroids (vertex average of the cell's face vertices)
    const int64_t C = m.nCells;
    std::vector<double> cx(3 * C, 0.0);
    std::vector<int32_t> cnt(C, 0);
    auto acc = [&](int32_t c, int32_t f) {
        for (int32_t k = m.faceOffsets[f]; k < m.faceOffsets[f + 1]; ++k)
        {
            for (int d = 0; d < 3; ++d)
                cx[3 * c + d] += m.points[3 * (int64_t)m.faceVerts[k] + d];
            ++cnt[c];
        }
    };
    for (int64_t f = 0; f < m.nFaces(); ++f)
        acc(m.owner[f], (int32_t)f);
    for (int64_t f = 0; f < m.nInternalFaces(); ++f)
        acc(m.neighbour[f], (int32_t)f);
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int64_t c = 0; c < C; ++c)
        for (int d = 0; d < 3; ++d)
        {
            cx[3 * c + d] /= cnt[c];
            lo[d] = std::min(lo[d], cx[3 * c + d]);
            hi[d] = std::max(hi[d], cx[3 * c + d]);
        }
    const int np[3] = {px, py, pz};
    std::vector<int32_t> part(C);
    for (int64_t c = 0; c < C; ++c)
    {
        int idx[3];
        for (int d = 0; d < 3; ++d)
        {
            const double t = (hi[d] > lo[d]) ? (cx[3 * c + d] - lo[d]) / (hi[d] - lo[d]) : 0.0;
            idx[d] = std::min(np[d] - 1, (int)(t * np[d]));
        }
        part[c] = idx[0] + px * (idx[1] + py * idx[2]);
    }
    return part;
}

std::vector<int32_t> partitionRCB(const PolyMesh &m, int nParts)
{
    const int64_t C = m.nCells;
    std::vector<double> cx(3 * C, 0.0);
    std::vector<int32_t> cnt(C, 0);
    auto acc = [&](int32_t c, int32_t f) {
        for (int32_t k = m.faceOffsets[f]; k < m.faceOffsets[f + 1]; ++k)
        {
            for (int d = 0; d < 3; ++d)
                cx[3 * c + d] += m.points[3 * (int64_t)m.faceVerts[k] + d];
            ++cnt[c];
        }
    };
    for (int64_t f = 0; f < m.nFaces(); ++f)
        acc(m.owner[f], (int32_t)f);
    for (int64_t f = 0; f < m.nInternalFaces(); ++f)
        acc(m.neighbour[f], (int32_t)f);
    for (int64_t c = 0; c < C; ++c)
        for (int d = 0; d < 3; ++d)
            cx[3 * c + d] /= cnt[c];
    std::vector<int32_t> part(C, 0), idx(C);
    std::iota(idx.begin(), idx.end(), 0);
    // recursive bisection: split [b,e) of idx into parts [p0,p0+np)
    struct Job
    {
        int64_t b, e;
        int p0, np;
    };
    std::vector<Job> jobs{{0, C, 0, nParts}};
    while (!jobs.empty())
    {
        Job j = jobs.back();
        jobs.pop_back();
        if (j.np == 1)
        {
            for (int64_t i = j.b; i < j.e; ++i)
                part[idx[i]] = j.p0;
            continue;
        }
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (int64_t i = j.b; i < j.e; ++i)
            for (int d = 0; d < 3; ++d)
            {
                lo[d] = std::min(lo[d], cx[3 * (int64_t)idx[i] + d]);
                hi[d] = std::max(hi[d], cx[3 * (int64_t)idx[i] + d]);
            }
        int d = 0;
        for (int q = 1; q < 3; ++q)
            if ((hi[q] - lo[q]) > (hi[d] - lo[d]) * (1.0 + 1e-12))
                d = q;
        const int nl = j.np / 2;
        const int64_t mid = j.b + (j.e - j.b) * nl / j.np;
        std::nth_element(idx.begin() + j.b, idx.begin() + mid, idx.begin() + j.e, [&](int32_t a, int32_t b) {
            const double va = cx[3 * (int64_t)a + d], vb = cx[3 * (int64_t)b + d];
            return va != vb ? va < vb : a < b;
        });
        jobs.push_back({j.b, mid, j.p0, nl});
        jobs.push_back({mid, j.e, j.p0 + nl, j.np - nl});
    }
    return part;
}

std::vector<PolyMesh> decompose(const PolyMesh &m, const std::vector<int32_t> &cellPart, int nParts)
{
    const int64_t F = m.nFaces(), Fi = m.nInternalFaces(), P = m.nPoints(), C = m.nCells;
    std::vector<PolyMesh> out(nParts);
    // local cell numbering: ascending global label
    std::vector<int32_t> cellLocal(C);
    {
        std::vector<int32_t> n(nParts, 0);
        for (int64_t c = 0; c < C; ++c)
        {
            cellLocal[c] = n[cellPart[c]]++;
            out[cellPart[c]].cellGlobalId.push_back(c);
        }
        for (int r = 0; r < nParts; ++r)
            out[r].nCells = n[r];
    }
    for (int r = 0; r < nParts; ++r)
    {
        PolyMesh &o = out[r];
        // faces of this part, in the order: internal (both cells here, ascending global face),
        // original patches, then processor patches by ascending neighbour part
        std::vector<int32_t> internal;
        std::vector<std::vector<int32_t>> patchFaces(m.patches.size());
        std::map<int32_t, std::vector<int32_t>> procFaces; // neighbour part -> global faces
        for (int64_t f = 0; f < Fi; ++f)
        {
            const int32_t po = cellPart[m.owner[f]], pn = cellPart[m.neighbour[f]];
            if (po == r && pn == r)
                internal.push_back((int32_t)f);
            else if (po == r)
                procFaces[pn].push_back((int32_t)f);
            else if (pn == r)
                procFaces[po].push_back((int32_t)f);
        }
        for (size_t p = 0; p < m.patches.size(); ++p)
            for (int32_t f = m.patches[p].start; f < m.patches[p].start + m.patches[p].size; ++f)
                if (cellPart[m.owner[f]] == r)
                    patchFaces[p].push_back(f);
        // local internal faces must be upper-triangular in LOCAL labels; since local labels are
        // monotone in global labels and global faces are upper-triangular, ascending global face
        // order already is.
        std::vector<int32_t> pointLocal; // built lazily via map global->local in ascending global order
        std::vector<int32_t> used;
        auto collect = [&](int32_t f) {
            for (int32_t k = m.faceOffsets[f]; k < m.faceOffsets[f + 1]; ++k)
                used.push_back(m.faceVerts[k]);
        };
        for (int32_t f : internal)
            collect(f);
        for (auto &pf : patchFaces)
            for (int32_t f : pf)
                collect(f);
        for (auto &kv : procFaces)
            for (int32_t f : kv.second)
                collect(f);
        std::sort(used.begin(), used.end());
        used.erase(std::unique(used.begin(), used.end()), used.end());
        std::unordered_map<int32_t, int32_t> g2l;
        g2l.reserve(used.size() * 2);
        o.points.resize(3 * used.size());
        o.pointGlobalId.resize(used.size());
        for (size_t i = 0; i < used.size(); ++i)
        {
            g2l[used[i]] = (int32_t)i;
            o.pointGlobalId[i] = m.pointGlobalId.empty() ? used[i] : m.pointGlobalId[used[i]];
            for (int d = 0; d < 3; ++d)
                o.points[3 * i + d] = m.points[3 * (int64_t)used[i] + d];
        }
        (void)P;
        o.faceOffsets.push_back(0);
        auto emit = [&](int32_t f, bool reverse, int32_t ownGlobal) {
            const int32_t b = m.faceOffsets[f], e = m.faceOffsets[f + 1];
            if (!reverse)
                for (int32_t k = b; k < e; ++k)
                    o.faceVerts.push_back(g2l[m.faceVerts[k]]);
            else
            { // OpenFOAM face::reverseFace keeps vertex 0 first: (v0, vn-1, ..., v1)
                o.faceVerts.push_back(g2l[m.faceVerts[b]]);
                for (int32_t k = e - 1; k > b; --k)
                    o.faceVerts.push_back(g2l[m.faceVerts[k]]);
            }
            o.faceOffsets.push_back((int32_t)o.faceVerts.size());
            o.owner.push_back(cellLocal[ownGlobal]);
        };
        for (int32_t f : internal)
        {
            emit(f, false, m.owner[f]);
            o.neighbour.push_back(cellLocal[m.neighbour[f]]);
        }
        for (size_t p = 0; p < m.patches.size(); ++p)
        {
            Patch pt = m.patches[p];
            pt.start = (int32_t)o.owner.size();
            for (int32_t f : patchFaces[p])
                emit(f, false, m.owner[f]);
            pt.size = (int32_t)o.owner.size() - pt.start;
            o.patches.push_back(pt);
        }
        for (auto &kv : procFaces)
        {
            Patch pt;
            pt.name = "procBoundary" + std::to_string(r) + "to" + std::to_string(kv.first);
            pt.type = "processor";
            pt.myProc = r;
            pt.nbrProc = kv.first;
            pt.start = (int32_t)o.owner.size();
            for (int32_t f : kv.second)
            {
                const bool ownerHere = cellPart[m.owner[f]] == r;
                emit(f, !ownerHere, ownerHere ? m.owner[f] : m.neighbour[f]);
            }
            pt.size = (int32_t)o.owner.size() - pt.start;
            o.patches.push_back(pt);
        }
    }
    (void)F;
    return out;
}

// ---------------------------------------------------------------- file I/O ----
namespace
{
struct Lexer
{
    std::string s;
    size_t i = 0;
    void skipWs()
    {
        for (;;)
        {
            while (i < s.size() && isspace((unsigned char)s[i]))
                ++i;
            if (i + 1 < s.size() && s[i] == '/' && s[i + 1] == '/')
            {
                while (i < s.size() && s[i] != '\n')
                    ++i;
            }
            else if (i + 1 < s.size() && s[i] == '/' && s[i + 1] == '*')
            {
                i += 2;
                while (i + 1 < s.size() && !(s[i] == '*' && s[i + 1] == '/'))
                    ++i;
                i += 2;
            }
            else
                break;
        }
    }
    bool eof()
    {
        skipWs();
        return i >= s.size();
    }
    char peek()
    {
        skipWs();
        return i < s.size() ? s[i] : '\0';
    }
    void expect(char c)
    {
        skipWs();
        if (i >= s.size() || s[i] != c)
            fail(std::string("polyMesh parse error: expected '") + c + "' at offset " + std::to_string(i));
        ++i;
    }
    std::string word()
    {
        skipWs();
        size_t b = i;
        if (i < s.size() && s[i] == '"')
        {
            ++i;
            while (i < s.size() && s[i] != '"')
                ++i;
            ++i;
            return s.substr(b + 1, i - b - 2);
        }
        while (i < s.size() && !isspace((unsigned char)s[i]) && s[i] != ';' && s[i] != '(' && s[i] != ')' &&
               s[i] != '{' && s[i] != '}')
            ++i;
        return s.substr(b, i - b);
    }
    long long integer()
    {
        skipWs();
        char *end;
        long long v = strtoll(s.c_str() + i, &end, 10);
        if (end == s.c_str() + i)
            fail("polyMesh parse error: expected integer at offset " + std::to_string(i));
        i = end - s.c_str();
        return v;
    }
    double real()
    {
        skipWs();
        char *end;
        double v = strtod(s.c_str() + i, &end);
        if (end == s.c_str() + i)
            fail("polyMesh parse error: expected number at offset " + std::to_string(i));
        i = end - s.c_str();
        return v;
    }
};

struct Header
{
    bool binary = false;
    std::string cls;
};

std::string slurp(const std::string &file)
{
    std::ifstream in(file, std::ios::binary);
    if (!in)
        fail("cannot open " + file);
    std::stringstream ss;
    ss << in.rdbuf();
    return ss.str();
}

Header readHeader(Lexer &lx)
{
    Header h;
    if (lx.word() != "FoamFile")
        fail("polyMesh parse error: missing FoamFile header");
    lx.expect('{');
    while (lx.peek() != '}')
    {
        const std::string key = lx.word();
        std::string val;
        // value runs to ';' (may contain quoted strings)
        lx.skipWs();
        size_t b = lx.i;
        bool inq = false;
        while (lx.i < lx.s.size() && (inq || lx.s[lx.i] != ';'))
        {
            if (lx.s[lx.i] == '"')
                inq = !inq;
            ++lx.i;
        }
        val = lx.s.substr(b, lx.i - b);
        lx.expect(';');
        if (key == "format")
            h.binary = val.find("binary") != std::string::npos;
        if (key == "class")
            h.cls = val;
        if (key == "arch" && (val.find("label=64") != std::string::npos || val.find("scalar=32") != std::string::npos))
            fail("only label=32 / scalar=64 polyMesh files are supported");
    }
    lx.expect('}');
    return h;
}

template <class T> std::vector<T> readBinaryList(Lexer &lx, long long n)
{
    lx.expect('(');
    std::vector<T> v(n);
    if (lx.i + n * sizeof(T) > lx.s.size())
        fail("polyMesh parse error: truncated binary list");
    memcpy(v.data(), lx.s.data() + lx.i, n * sizeof(T));
    lx.i += n * sizeof(T);
    lx.expect(')');
    return v;
}

std::vector<int32_t> readLabelList(Lexer &lx, bool binary)
{
    const long long n = lx.integer();
    if (lx.peek() == '{')
    { // uniform list N{v}
        lx.expect('{');
        const int32_t v = (int32_t)lx.integer();
        lx.expect('}');
        return std::vector<int32_t>(n, v);
    }
    if (binary)
        return n ? readBinaryList<int32_t>(lx, n) : (lx.peek() == '(' ? (lx.expect('('), lx.expect(')'), std::vector<int32_t>()) : std::vector<int32_t>());
    std::vector<int32_t> v(n);
    lx.expect('(');
    for (long long k = 0; k < n; ++k)
        v[k] = (int32_t)lx.integer();
    lx.expect(')');
    return v;
}

std::string foamHeader(const std::string &cls, const std::string &location, const std::string &object, bool binary,
                       const std::string &note = "")
{
    std::ostringstream o;
    o << "/*--------------------------------*- C++ -*----------------------------------*\\\n"
      << "| smoothmesh_b200 polyMesh writer                                             |\n"
      << "\\*---------------------------------------------------------------------------*/\n"
      << "FoamFile\n{\n    version     2.0;\n    format      " << (binary ? "binary" : "ascii") << ";\n"
      << "    arch        \"LSB;label=32;scalar=64\";\n";
    if (!note.empty())
        o << "    note        \"" << note << "\";\n";
    o << "    class       " << cls << ";\n    location    \"" << location << "\";\n    object      " << object
      << ";\n}\n// * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * //\n\n";
    return o.str();
}
} // namespace

std::vector<double> readPoints(const std::string &file)
{
    Lexer lx;
    lx.s = slurp(file);
    const Header h = readHeader(lx);
    const long long n = lx.integer();
    std::vector<double> pts;
    if (h.binary)
        pts = readBinaryList<double>(lx, 3 * n);
    else
    {
        pts.resize(3 * n);
        lx.expect('(');
        for (long long k = 0; k < n; ++k)
        {
            lx.expect('(');
            pts[3 * k] = lx.real();
            pts[3 * k + 1] = lx.real();
            pts[3 * k + 2] = lx.real();
            lx.expect(')');
        }
        lx.expect(')');
    }
    return pts;
}

PolyMesh readPolyMesh(const std::string &dir)
{
    PolyMesh m;
    {
        const std::vector<double> pts = readPoints(dir + "/points");
        m.points.assign(pts.begin(), pts.end());
    }
    {
        Lexer lx;
        lx.s = slurp(dir + "/faces");
        const Header h = readHeader(lx);
        if (h.cls.find("faceCompactList") != std::string::npos)
        {
            assignVec(m.faceOffsets, readLabelList(lx, h.binary));
            assignVec(m.faceVerts, readLabelList(lx, h.binary));
        }
        else
        {
            if (h.binary)
                fail("binary faceList (non-compact) is not supported");
            const long long n = lx.integer();
            lx.expect('(');
            m.faceOffsets.push_back(0);
            for (long long k = 0; k < n; ++k)
            {
                const long long nv = lx.integer();
                lx.expect('(');
                for (long long q = 0; q < nv; ++q)
                    m.faceVerts.push_back((int32_t)lx.integer());
                lx.expect(')');
                m.faceOffsets.push_back((int32_t)m.faceVerts.size());
            }
            lx.expect(')');
        }
    }
    {
        Lexer lx;
        lx.s = slurp(dir + "/owner");
        const Header h = readHeader(lx);
        assignVec(m.owner, readLabelList(lx, h.binary));
    }
    {
        Lexer lx;
        lx.s = slurp(dir + "/neighbour");
        const Header h = readHeader(lx);
        assignVec(m.neighbour, readLabelList(lx, h.binary));
    }
    {
        Lexer lx;
        lx.s = slurp(dir + "/boundary");
        readHeader(lx);
        const long long n = lx.integer();
        lx.expect('(');
        for (long long k = 0; k < n; ++k)
        {
            Patch p;
            p.name = lx.word();
            lx.expect('{');
            while (lx.peek() != '}')
            {
                const std::string key = lx.word();
                if (lx.peek() == '{')
                { // nested dictionary: skip
                    int depth = 0;
                    do
                    {
                        if (lx.s[lx.i] == '{')
                            ++depth;
                        if (lx.s[lx.i] == '}')
                            --depth;
                        ++lx.i;
                    } while (depth > 0 && lx.i < lx.s.size());
                    continue;
                }
                lx.skipWs();
                size_t b = lx.i;
                int par = 0;
                while (lx.i < lx.s.size() && (par > 0 || lx.s[lx.i] != ';'))
                {
                    if (lx.s[lx.i] == '(')
                        ++par;
                    if (lx.s[lx.i] == ')')
                        --par;
                    ++lx.i;
                }
                std::string val = lx.s.substr(b, lx.i - b);
                lx.expect(';');
                while (!val.empty() && isspace((unsigned char)val.back()))
                    val.pop_back();
                if (key == "type")
                    p.type = val;
                else if (key == "nFaces")
                    p.size = atoi(val.c_str());
                else if (key == "startFace")
                    p.start = atoi(val.c_str());
                else if (key == "myProcNo")
                    p.myProc = atoi(val.c_str());
                else if (key == "neighbProcNo")
                    p.nbrProc = atoi(val.c_str());
            }
            lx.expect('}');
            m.patches.push_back(p);
        }
        lx.expect(')');
    }
    int32_t maxCell = -1;
    for (int32_t c : m.owner)
        maxCell = std::max(maxCell, c);
    for (int32_t c : m.neighbour)
        maxCell = std::max(maxCell, c);
    m.nCells = (int64_t)maxCell + 1;
    m.check();
    return m;
}

static void mkdirs(const std::string &dir)
{
    std::string cur;
    for (size_t i = 0; i <= dir.size(); ++i)
        if (i == dir.size() || dir[i] == '/')
        {
            if (!cur.empty())
            {
                const std::string cmd = cur;
                (void)cmd;
                if (::system(("mkdir -p '" + cur + "'").c_str()) != 0)
                    fail("cannot create directory " + cur);
            }
            if (i < dir.size())
                cur += dir[i];
        }
        else
            cur += dir[i];
}

void writePoints(const double *pts, int64_t nPoints, const std::string &dir, bool binary, int precision,
                 const std::string &location)
{
    mkdirs(dir);
    std::ofstream o(dir + "/points", std::ios::binary);
    if (!o)
        fail("cannot write " + dir + "/points");
    o << foamHeader("vectorField", location, "points", binary);
    o << nPoints << "\n(";
    if (binary)
        o.write((const char *)pts, sizeof(double) * 3 * nPoints);
    else
    {
        o << "\n";
        char buf[128];
        for (int64_t k = 0; k < nPoints; ++k)
        {
            snprintf(buf, sizeof buf, "(%.*g %.*g %.*g)\n", precision, pts[3 * k], precision, pts[3 * k + 1], precision,
                     pts[3 * k + 2]);
            o << buf;
        }
    }
    o << ")\n";
}

void writePolyMesh(const PolyMesh &m, const std::string &dir, bool binary, int precision)
{
    writePoints(m.points.data(), m.nPoints(), dir, binary, precision, "constant/polyMesh");
    const std::string note = "nPoints:" + std::to_string(m.nPoints()) + "  nCells:" + std::to_string(m.nCells) +
                             "  nFaces:" + std::to_string(m.nFaces()) +
                             "  nInternalFaces:" + std::to_string(m.nInternalFaces());
    auto writeLabels = [&](std::ofstream &o, const auto &v) {
        o << v.size() << "\n(";
        if (binary)
            o.write((const char *)v.data(), sizeof(int32_t) * v.size());
        else
        {
            o << "\n";
            for (int32_t x : v)
                o << x << "\n";
        }
        o << ")\n";
    };
    {
        std::ofstream o(dir + "/faces", std::ios::binary);
        if (binary)
        {
            o << foamHeader("faceCompactList", "constant/polyMesh", "faces", true);
            writeLabels(o, m.faceOffsets);
            o << "\n";
            writeLabels(o, m.faceVerts);
        }
        else
        {
            o << foamHeader("faceList", "constant/polyMesh", "faces", false);
            o << m.nFaces() << "\n(\n";
            for (int64_t f = 0; f < m.nFaces(); ++f)
            {
                o << (m.faceOffsets[f + 1] - m.faceOffsets[f]) << "(";
                for (int32_t k = m.faceOffsets[f]; k < m.faceOffsets[f + 1]; ++k)
                    o << m.faceVerts[k] << (k + 1 < m.faceOffsets[f + 1] ? " " : "");
                o << ")\n";
            }
            o << ")\n";
        }
    }
    {
        std::ofstream o(dir + "/owner", std::ios::binary);
        o << foamHeader("labelList", "constant/polyMesh", "owner", binary, note);
        writeLabels(o, m.owner);
    }
    {
        std::ofstream o(dir + "/neighbour", std::ios::binary);
        o << foamHeader("labelList", "constant/polyMesh", "neighbour", binary, note);
        writeLabels(o, m.neighbour);
    }
    {
        std::ofstream o(dir + "/boundary", std::ios::binary);
        o << foamHeader("polyBoundaryMesh", "constant/polyMesh", "boundary", false);
        o << m.patches.size() << "\n(\n";
        for (const Patch &p : m.patches)
        {
            o << "    " << p.name << "\n    {\n        type            " << p.type << ";\n";
            if (p.type == "wall")
                o << "        inGroups        1(wall);\n";
            o << "        nFaces          " << p.size << ";\n        startFace       " << p.start << ";\n";
            if (p.type == "processor")
                o << "        matchTolerance  0.0001;\n        myProcNo        " << p.myProc
                  << ";\n        neighbProcNo    " << p.nbrProc << ";\n";
            o << "    }\n";
        }
        o << ")\n";
    }
}

} // namespace sm

// ---------------------------------------------------- distributed hex block ----
namespace sm
{
// One brick of a (nx*px) x (ny*py) x (nz*pz) hex block, generated without ever building the
// global mesh: local numbering as genHexBlock, coordinates and pointGlobalId from the global
// lattice, sides that touch another brick turned into processor patches (listed after the
// six physical patches, ascending neighbour rank).
PolyMesh genHexBlockPart(int nx, int ny, int nz, int px, int py, int pz, int rank, const double lo[3], const double hi[3])
{
    const int ix = rank % px, iy = (rank / px) % py, iz = rank / (px * py);
    if (rank < 0 || iz >= pz)
        throw std::runtime_error("genHexBlockPart: rank out of range");
    const int64_t NX = (int64_t)nx * px, NY = (int64_t)ny * py, NZ = (int64_t)nz * pz;
    const double l[3] = {0, 0, 0}, h[3] = {1, 1, 1};
    PolyMesh m = genHexBlock(nx, ny, nz, l, h, "wall");
    const int64_t lx = nx + 1, ly = ny + 1;
    const int64_t P = m.nPoints();
    m.pointGlobalId.resize(P);
    for (int64_t p = 0; p < P; ++p)
    {
        const int64_t i = p % lx + (int64_t)ix * nx, j = (p / lx) % ly + (int64_t)iy * ny, k = p / (lx * ly) + (int64_t)iz * nz;
        m.points[3 * p + 0] = lo[0] + (hi[0] - lo[0]) * (double(i) / double(NX));
        m.points[3 * p + 1] = lo[1] + (hi[1] - lo[1]) * (double(j) / double(NY));
        m.points[3 * p + 2] = lo[2] + (hi[2] - lo[2]) * (double(k) / double(NZ));
        m.pointGlobalId[p] = i + j * (NX + 1) + k * (NX + 1) * (NY + 1);
    }
    m.cellGlobalId.resize(m.nCells);
    for (int64_t c = 0; c < m.nCells; ++c)
    {
        const int64_t i = c % nx + (int64_t)ix * nx, j = (c / nx) % ny + (int64_t)iy * ny, k = c / ((int64_t)nx * ny) + (int64_t)iz * nz;
        m.cellGlobalId[c] = i + j * NX + k * NX * NY;
    }
    // neighbour rank behind each of the six sides (xMin,xMax,yMin,yMax,zMin,zMax), -1 = physical
    const int nbr[6] = {ix > 0 ? rank - 1 : -1,           ix + 1 < px ? rank + 1 : -1,
                        iy > 0 ? rank - px : -1,          iy + 1 < py ? rank + px : -1,
                        iz > 0 ? rank - px * py : -1,     iz + 1 < pz ? rank + px * py : -1};
    // new boundary order: six physical patches (emptied where the side is internal), then
    // processor patches by ascending neighbour rank
    std::vector<int> procSides;
    for (int s = 0; s < 6; ++s)
        if (nbr[s] >= 0)
            procSides.push_back(s);
    std::sort(procSides.begin(), procSides.end(), [&](int a, int b) { return nbr[a] < nbr[b]; });
    const int64_t Fi = m.nInternalFaces();
    std::vector<int32_t> newOff(m.faceOffsets.begin(), m.faceOffsets.begin() + Fi + 1), newVerts(m.faceVerts.begin(), m.faceVerts.begin() + m.faceOffsets[Fi]),
        newOwner(m.owner.begin(), m.owner.begin() + Fi);
    std::vector<Patch> newPatches;
    auto appendSide = [&](int s) {
        const Patch &src = m.patches[s];
        for (int32_t f = src.start; f < src.start + src.size; ++f)
        {
            for (int32_t k = m.faceOffsets[f]; k < m.faceOffsets[f + 1]; ++k)
                newVerts.push_back(m.faceVerts[k]);
            newOff.push_back((int32_t)newVerts.size());
            newOwner.push_back(m.owner[f]);
        }
    };
    for (int s = 0; s < 6; ++s)
    {
        Patch p = m.patches[s];
        p.start = (int32_t)newOwner.size();
        if (nbr[s] < 0)
            appendSide(s);
        p.size = (int32_t)newOwner.size() - p.start;
        newPatches.push_back(p);
    }
    for (int s : procSides)
    {
        Patch p;
        p.name = "procBoundary" + std::to_string(rank) + "to" + std::to_string(nbr[s]);
        p.type = "processor";
        p.myProc = rank;
        p.nbrProc = nbr[s];
        p.start = (int32_t)newOwner.size();
        appendSide(s);
        p.size = (int32_t)newOwner.size() - p.start;
        newPatches.push_back(p);
    }
    assignVec(m.faceOffsets, newOff);
    assignVec(m.faceVerts, newVerts);
    assignVec(m.owner, newOwner);
    m.patches.swap(newPatches);
    return m;
}
} // namespace sm

// ------------------------------------------------------------- renumbering ----
namespace sm
{
namespace
{
inline uint64_t spread21(uint64_t v)
{ // interleave helper: 21 bits -> every third bit
    v &= 0x1fffff;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}
inline uint64_t mortonKey(const double *x, const double *lo, const double *inv)
{
    uint64_t k = 0;
    for (int d = 0; d < 3; ++d)
    {
        double t = (x[d] - lo[d]) * inv[d];
        t = t < 0 ? 0 : (t > 1 ? 1 : t);
        k |= spread21((uint64_t)(t * 2097151.0)) << d;
    }
    return k;
}
} // namespace

// Space-filling-curve (Morton) renumbering of points and cells, the stand-in for OpenFOAM's
// renumberMesh utility: produces a valid polyMesh (faces re-sorted upper-triangular, flipped
// where owner and neighbour swap) whose storage order makes the smoothing kernels' gathers
// local.  Labels change, so label-order-dependent results (stable-sort ties, worklist order,
// summation order) are those of the renumbered mesh -- exactly as if renumberMesh had been run
// before the reference.
PolyMesh renumberMorton(const PolyMesh &m, std::vector<int32_t> &pointOldOfNew, std::vector<int32_t> &cellOldOfNew)
{
    const int64_t P = m.nPoints(), C = m.nCells, F = m.nFaces(), Fi = m.nInternalFaces();
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, inv[3];
    for (int64_t p = 0; p < P; ++p)
        for (int d = 0; d < 3; ++d)
        {
            lo[d] = std::min(lo[d], m.points[3 * p + d]);
            hi[d] = std::max(hi[d], m.points[3 * p + d]);
        }
    for (int d = 0; d < 3; ++d)
        inv[d] = hi[d] > lo[d] ? 1.0 / (hi[d] - lo[d]) : 0.0;
    std::vector<std::pair<uint64_t, int32_t>> keys(P);
    for (int64_t p = 0; p < P; ++p)
        keys[p] = {mortonKey(&m.points[3 * p], lo, inv), (int32_t)p};
    std::sort(keys.begin(), keys.end());
    pointOldOfNew.resize(P);
    std::vector<int32_t> pointNew(P);
    for (int64_t i = 0; i < P; ++i)
    {
        pointOldOfNew[i] = keys[i].second;
        pointNew[keys[i].second] = (int32_t)i;
    }
    std::vector<double> cx(3 * C, 0.0);
    std::vector<int32_t> cnt(C, 0);
    auto acc = [&](int32_t c, int64_t f) {
        for (int32_t k = m.faceOffsets[f]; k < m.faceOffsets[f + 1]; ++k)
        {
            for (int d = 0; d < 3; ++d)
                cx[3 * (int64_t)c + d] += m.points[3 * (int64_t)m.faceVerts[k] + d];
            ++cnt[c];
        }
    };
    for (int64_t f = 0; f < F; ++f)
        acc(m.owner[f], f);
    for (int64_t f = 0; f < Fi; ++f)
        acc(m.neighbour[f], f);
    keys.resize(C);
    for (int64_t c = 0; c < C; ++c)
    {
        for (int d = 0; d < 3; ++d)
            cx[3 * c + d] /= cnt[c];
        keys[c] = {mortonKey(&cx[3 * c], lo, inv), (int32_t)c};
    }
    std::sort(keys.begin(), keys.end());
    cellOldOfNew.resize(C);
    std::vector<int32_t> cellNew(C);
    for (int64_t i = 0; i < C; ++i)
    {
        cellOldOfNew[i] = keys[i].second;
        cellNew[keys[i].second] = (int32_t)i;
    }
    struct IF
    {
        int32_t own, nei, f;
        bool flip;
    };
    std::vector<IF> ifs(Fi);
    for (int64_t f = 0; f < Fi; ++f)
    {
        const int32_t a = cellNew[m.owner[f]], b = cellNew[m.neighbour[f]];
        ifs[f] = {std::min(a, b), std::max(a, b), (int32_t)f, a > b};
    }
    std::sort(ifs.begin(), ifs.end(), [](const IF &x, const IF &y) { return x.own != y.own ? x.own < y.own : (x.nei != y.nei ? x.nei < y.nei : x.f < y.f); });
    PolyMesh o;
    o.nCells = C;
    o.points.resize(3 * P);
    for (int64_t i = 0; i < P; ++i)
        for (int d = 0; d < 3; ++d)
            o.points[3 * i + d] = m.points[3 * (int64_t)pointOldOfNew[i] + d];
    o.faceOffsets.push_back(0);
    auto emit = [&](int64_t f, bool flip) {
        const int32_t b = m.faceOffsets[f], e = m.faceOffsets[f + 1];
        if (!flip)
            for (int32_t k = b; k < e; ++k)
                o.faceVerts.push_back(pointNew[m.faceVerts[k]]);
        else
        {
            o.faceVerts.push_back(pointNew[m.faceVerts[b]]);
            for (int32_t k = e - 1; k > b; --k)
                o.faceVerts.push_back(pointNew[m.faceVerts[k]]);
        }
        o.faceOffsets.push_back((int32_t)o.faceVerts.size());
    };
    for (const IF &x : ifs)
    {
        emit(x.f, x.flip);
        o.owner.push_back(x.own);
        o.neighbour.push_back(x.nei);
    }
    o.patches = m.patches;
    for (int64_t f = Fi; f < F; ++f)
    {
        emit(f, false);
        o.owner.push_back(cellNew[m.owner[f]]);
    }
    if (!m.pointGlobalId.empty())
    {
        o.pointGlobalId.resize(P);
        for (int64_t i = 0; i < P; ++i)
            o.pointGlobalId[i] = m.pointGlobalId[pointOldOfNew[i]];
    }
    if (!m.cellGlobalId.empty())
    {
        o.cellGlobalId.resize(C);
        for (int64_t i = 0; i < C; ++i)
            o.cellGlobalId[i] = m.cellGlobalId[cellOldOfNew[i]];
    }
    return o;
}
} // namespace sm

// ------------------------------------------------------------ mesh quality ----
namespace sm
{
namespace
{
struct Q3
{
    double x, y, z;
};
inline Q3 operator+(Q3 a, Q3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Q3 operator-(Q3 a, Q3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Q3 operator*(double s, Q3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline double qdot(Q3 a, Q3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Q3 qcross(Q3 a, Q3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double qmag(Q3 a) { return std::sqrt(qdot(a, a)); }
} // namespace

// checkMesh-style quality figures ([OF-recalled] primitiveMeshTools::
// faceOrthogonality / faceSkewness) plus smoothMesh's own minimum edge-edge angle
// (src/smoothMesh.C:766-786, without the 0.99999 clamp).  Host-side, not on the iteration path.
MeshQuality computeQuality(const PolyMesh &m)
{
    const int64_t P = m.nPoints(), C = m.nCells, F = m.nFaces(), Fi = m.nInternalFaces();
    auto pt = [&](int32_t i) { return Q3{m.points[3 * (int64_t)i], m.points[3 * (int64_t)i + 1], m.points[3 * (int64_t)i + 2]}; };
    std::vector<Q3> fc(F), fa(F);
    for (int64_t f = 0; f < F; ++f)
    {
        const int32_t b = m.faceOffsets[f], n = m.faceOffsets[f + 1] - b;
        if (n == 3)
        {
            const Q3 p0 = pt(m.faceVerts[b]), p1 = pt(m.faceVerts[b + 1]), p2 = pt(m.faceVerts[b + 2]);
            fc[f] = (1.0 / 3.0) * (p0 + p1 + p2);
            fa[f] = 0.5 * qcross(p1 - p0, p2 - p0);
            continue;
        }
        Q3 est = {0, 0, 0};
        for (int32_t k = 0; k < n; ++k)
            est = est + pt(m.faceVerts[b + k]);
        est = (1.0 / n) * est;
        Q3 sN = {0, 0, 0}, sAc = {0, 0, 0};
        double sA = 0;
        for (int32_t k = 0; k < n; ++k)
        {
            const Q3 a = pt(m.faceVerts[b + k]), c = pt(m.faceVerts[b + (k + 1) % n]);
            const Q3 nn = qcross(c - a, est - a);
            const double w = qmag(nn);
            sN = sN + nn;
            sA += w;
            sAc = sAc + w * (a + c + est);
        }
        fc[f] = sA > 1e-150 ? (1.0 / (3.0 * sA)) * sAc : est;
        fa[f] = 0.5 * sN;
    }
    std::vector<Q3> est(C, Q3{0, 0, 0}), cc(C, Q3{0, 0, 0});
    std::vector<double> vol(C, 0.0);
    std::vector<int32_t> nf(C, 0);
    for (int64_t f = 0; f < F; ++f)
    {
        est[m.owner[f]] = est[m.owner[f]] + fc[f];
        ++nf[m.owner[f]];
    }
    for (int64_t f = 0; f < Fi; ++f)
    {
        est[m.neighbour[f]] = est[m.neighbour[f]] + fc[f];
        ++nf[m.neighbour[f]];
    }
    for (int64_t c = 0; c < C; ++c)
        est[c] = (1.0 / nf[c]) * est[c];
    auto pyr = [&](int32_t c, int64_t f, double sign) {
        const double v = sign * qdot(fa[f], fc[f] - est[c]);
        cc[c] = cc[c] + v * (0.75 * fc[f] + 0.25 * est[c]);
        vol[c] += v;
    };
    for (int64_t f = 0; f < F; ++f)
        pyr(m.owner[f], f, 1.0);
    for (int64_t f = 0; f < Fi; ++f)
        pyr(m.neighbour[f], f, -1.0);
    MeshQuality q;
    q.minVolume = 1e300;
    for (int64_t c = 0; c < C; ++c)
    {
        cc[c] = std::fabs(vol[c]) > 1e-300 ? (1.0 / vol[c]) * cc[c] : est[c];
        q.minVolume = std::min(q.minVolume, vol[c] / 3.0);
    }
    const double VS = 1e-150;
    double sumNonOrtho = 0;
    for (int64_t f = 0; f < Fi; ++f)
    {
        const Q3 d = cc[m.neighbour[f]] - cc[m.owner[f]];
        const double cosT = qdot(d, fa[f]) / (qmag(d) * qmag(fa[f]) + VS);
        const double ang = std::acos(std::max(-1.0, std::min(1.0, cosT))) * 180.0 / M_PI;
        q.maxNonOrtho = std::max(q.maxNonOrtho, ang);
        sumNonOrtho += ang;
        const Q3 cpf = fc[f] - cc[m.owner[f]];
        const Q3 sv = cpf - (qdot(fa[f], cpf) / (qdot(fa[f], d) + VS)) * d;
        const Q3 svHat = (1.0 / (qmag(sv) + VS)) * sv;
        double fd = 0.2 * qmag(d) + VS;
        for (int32_t k = m.faceOffsets[f]; k < m.faceOffsets[f + 1]; ++k)
            fd = std::max(fd, std::fabs(qdot(svHat, pt(m.faceVerts[k]) - fc[f])));
        q.maxSkewness = std::max(q.maxSkewness, qmag(sv) / fd);
    }
    q.avgNonOrtho = Fi ? sumNonOrtho / Fi : 0.0;
    for (int64_t f = Fi; f < F; ++f)
    {
        const Q3 cpf = fc[f] - cc[m.owner[f]];
        const Q3 nHat = (1.0 / (qmag(fa[f]) + VS)) * fa[f];
        const Q3 d = qdot(nHat, cpf) * nHat;
        const Q3 sv = cpf - (qdot(fa[f], cpf) / (qdot(fa[f], d) + VS)) * d;
        const Q3 svHat = (1.0 / (qmag(sv) + VS)) * sv;
        double fd = 0.4 * qmag(d) + VS;
        for (int32_t k = m.faceOffsets[f]; k < m.faceOffsets[f + 1]; ++k)
            fd = std::max(fd, std::fabs(qdot(svHat, pt(m.faceVerts[k]) - fc[f])));
        q.maxSkewness = std::max(q.maxSkewness, qmag(sv) / fd);
    }
    q.minEdgeLength = 1e300;
    q.minEdgeAngle = 180.0;
    for (int64_t f = 0; f < F; ++f)
    {
        const int32_t b = m.faceOffsets[f], n = m.faceOffsets[f + 1] - b;
        for (int32_t k = 0; k < n; ++k)
        {
            const Q3 c = pt(m.faceVerts[b + k]), a = pt(m.faceVerts[b + (k + n - 1) % n]), e = pt(m.faceVerts[b + (k + 1) % n]);
            const Q3 u = a - c, v = e - c;
            q.minEdgeLength = std::min(q.minEdgeLength, qmag(v));
            q.maxEdgeLength = std::max(q.maxEdgeLength, qmag(v));
            const double cosA = qdot(u, v) / (qmag(u) * qmag(v) + VS);
            q.minEdgeAngle = std::min(q.minEdgeAngle, std::acos(std::max(-1.0, std::min(1.0, cosA))) * 180.0 / M_PI);
        }
    }
    (void)P;
    return q;
}
} // namespace sm

// ------------------------------------------------ decomposed (processorN) cases ----
namespace sm
{
namespace
{
void writeLabelListFile(const std::string &file, const std::string &object, const std::string &location,
                        const std::vector<int64_t> &v, bool binary)
{
    std::ofstream o(file, std::ios::binary);
    if (!o)
        fail("cannot write " + file);
    o << foamHeader("labelIOList", location, object, binary);
    bool uniform = v.size() > 1;
    for (size_t i = 1; i < v.size() && uniform; ++i)
        uniform = v[i] == v[0];
    if (uniform && !binary)
    { // [OF-recalled] List<T>::writeList: uniform contiguous lists are written as N{value}
        o << v.size() << "{" << v[0] << "}\n";
        return;
    }
    o << v.size() << "\n(";
    if (binary)
    {
        std::vector<int32_t> t(v.begin(), v.end());
        o.write((const char *)t.data(), sizeof(int32_t) * t.size());
    }
    else
    {
        o << "\n";
        for (int64_t x : v)
            o << x << "\n";
    }
    o << ")\n";
}
} // namespace

// processor<k>/constant/polyMesh/* plus pointProcAddressing / cellProcAddressing, the layout
// decomposePar produces (testcase/run_parallel) and the reference reads under -parallel
void writeDecomposedCase(const std::vector<PolyMesh> &parts, const std::string &caseDir, bool binary)
{
    for (size_t k = 0; k < parts.size(); ++k)
    {
        const std::string dir = caseDir + "/processor" + std::to_string(k) + "/constant/polyMesh";
        writePolyMesh(parts[k], dir, binary, 17);
        const std::string loc = "constant/polyMesh";
        writeLabelListFile(dir + "/pointProcAddressing", "pointProcAddressing", loc, parts[k].pointGlobalId, binary);
        writeLabelListFile(dir + "/cellProcAddressing", "cellProcAddressing", loc, parts[k].cellGlobalId, binary);
    }
}

// one processor mesh of a decomposed case, with its addressing (topology from constant/)
PolyMesh readProcessorMesh(const std::string &caseDir, int k)
{
    const std::string dir = caseDir + "/processor" + std::to_string(k) + "/constant/polyMesh";
    PolyMesh m = readPolyMesh(dir);
    auto readIds = [&](const std::string &file) {
        Lexer lx;
        lx.s = slurp(file);
        const Header h = readHeader(lx);
        const std::vector<int32_t> v = readLabelList(lx, h.binary);
        return std::vector<int64_t>(v.begin(), v.end());
    };
    m.pointGlobalId = readIds(dir + "/pointProcAddressing");
    if ((int64_t)m.pointGlobalId.size() != m.nPoints())
        fail("pointProcAddressing size does not match the processor mesh");
    std::ifstream probe(dir + "/cellProcAddressing");
    if (probe.good())
        m.cellGlobalId = readIds(dir + "/cellProcAddressing");
    return m;
}

// labelIOList files next to the mesh (isCornerPoint / isFeatureEdgePoint, src/smoothMesh.C:2039-2065)
void writeLabelIOList(const std::string &file, const std::string &object, const std::string &location,
                      const std::vector<int32_t> &v, bool binary)
{
    writeLabelListFile(file, object, location, std::vector<int64_t>(v.begin(), v.end()), binary);
}
std::vector<int32_t> readLabelIOList(const std::string &file)
{
    Lexer lx;
    lx.s = slurp(file);
    const Header h = readHeader(lx);
    return readLabelList(lx, h.binary);
}
} // namespace sm
