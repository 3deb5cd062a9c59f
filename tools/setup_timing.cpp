// setup_timing.cpp -- wall time of the host-side set-up phases (connectivity build and geometry tiles) without a
// GPU -- the part of smgpu_create that runs before the first upload -- and a fingerprint of every table they
// produce, to compare two builds of topology.cpp on the same mesh.
//   make build/setup_timing
//   SMGPU_TIMING=1 build/setup_timing hex 200            (n^3 hex block)
//   build/setup_timing kelvin 12 | build/setup_timing dir <polyMesh directory>
#include "polymesh.hpp"
#include "topology.hpp"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

template <class V> static unsigned long long fingerprint(const V &v)
{
    unsigned long long h = 1469598103934665603ull;
    const unsigned char *p = reinterpret_cast<const unsigned char *>(v.data());
    const size_t n = v.size() * sizeof(v[0]);
    for (size_t i = 0; i < n; ++i)
        h = (h ^ p[i]) * 1099511628211ull;
    return h ^ v.size();
}

int main(int argc, char **argv)
{
    const std::string kind = argc > 1 ? argv[1] : "hex";
    const int reps = argc > 3 ? atoi(argv[3]) : 1;
    auto wall = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
    double t0 = wall();
    sm::PolyMesh m;
    if (kind == "hex")
    {
        const int n = argc > 2 ? atoi(argv[2]) : 200;
        m = sm::genHexBlock(n, n + (n < 100 ? 3 : 0), n - (n < 100 ? 2 : 0), lo, hi);
    }
    else if (kind == "kelvin")
        m = sm::genKelvin(argc > 2 ? atoi(argv[2]) : 10, 1.0);
    else
        m = sm::readPolyMesh(argv[2]);
    fprintf(stderr, "mesh %.3f s\n", wall() - t0);
    const bool prints = m.nCells < 3000000;
    for (int r = 0; r < reps; ++r)
    {
        t0 = wall();
        sm::Topology t = sm::buildTopology(m);
        const double t1 = wall();
        sm::GeomTiles G = sm::buildGeomTiles(m, t, 256, 1024, 1024);
        const double t2 = wall();
        fprintf(stderr, "== topology %.3f s, tiles %.3f s (%d tiles, %lld uniform cells)\n", t1 - t0, t2 - t1, G.nTiles,
                (long long)G.nUniformCells);
        if (prints && r == 0)
        {
            sm::GeomTiles K = sm::buildGeomTiles(m, t, 64, 300, 200, true);
            sm::buildEdgeRecords(t);
#define FP(x) printf("%-22s %016llx\n", #x, fingerprint(x))
            FP(t.pcOff), FP(t.pc), FP(t.ppOff), FP(t.pp), FP(t.pe), FP(t.cornerOff), FP(t.corner), FP(t.edge), FP(t.efOff), FP(t.ef);
            FP(t.ecOff), FP(t.ecCell), FP(t.ecPair), FP(t.faceOff), FP(t.faceVerts), FP(t.cfOff), FP(t.cf), FP(t.pointRec), FP(t.edgeRec);
            FP(t.isInternal), FP(t.procPoints);
            printf("scalars %lld %lld %.17g %.17g %d %d %d\n", (long long)t.E, (long long)t.P, t.minEdgeLength, t.maxEdgeLength,
                   t.maxPointDegree, t.maxFaceSize, t.maxEdgeFaces);
            for (const sm::GeomTiles *g : {&G, &K})
            {
                const sm::GeomTiles &T = *g;
                FP(T.tileCellOff), FP(T.tileCells), FP(T.tileFaceOff), FP(T.tileFaces), FP(T.slotOff), FP(T.slotRef), FP(T.tilePointOff);
                FP(T.tilePoints), FP(T.faceRefOff), FP(T.faceRef), FP(T.cellEdgeOff), FP(T.cellEdgeRef), FP(T.hexRec), FP(T.tileUFaceOff);
                FP(T.tileUCellOff), FP(T.uFaceRef), FP(T.uSlotRef);
                printf("tiles %d %lld %d %d %d %d %d\n", T.nTiles, (long long)T.nUniformCells, T.uniformCellEdges, T.maxTileCells,
                       T.maxTileFaces, T.maxTilePoints, T.maxTileEdgePairs);
            }
        }
    }
    return 0;
}
