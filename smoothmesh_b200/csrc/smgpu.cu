// smgpu.cu -- C ABI (include/smgpu.h) over the kernels in kernels.cuh.
// Host side: flatten mesh -> derived connectivity (topology.cpp) -> upload once
// -> launch the per-iteration kernel sequence with no host round trip inside a
// chunk of iterations.  There is no CPU fallback anywhere in this file.
#include "../../include/smgpu.h"
#include "kernels.cuh"
#include "polymesh.hpp"
#include "topology.hpp"
#include "boundary.hpp"

#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#include <omp.h>

using namespace smk;

namespace sm
{
struct Comm;
static void commSetupLayers(Comm *cm, struct ::smgpu_handle *h);
static bool commIsGroupMember(const Comm *cm);
// global getMeshStats figures and the hop-count synchronisation for the boundary point smoothing set-up of a rank
static void commBoundaryParallel(Comm *cm, struct ::smgpu_handle *h, const std::vector<double> &points, struct BoundaryParallel &par);
}

static thread_local std::string g_err;
static int setErr(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}
#define CK(call)                                                                                                       \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t e_ = (call);                                                                                       \
        if (e_ != cudaSuccess)                                                                                         \
            throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #call);             \
    } while (0)

struct smgpu_handle
{
    sm::Topology topo;
    smgpu_params prm;
    Dev d;
    std::vector<void *> allocs;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double lastMs = 0;
    int64_t lastLaunches = 0, launches = 0;
    int64_t nInternal = 0;
    int statCap = 0;
    sm::Comm *comm = nullptr;
    std::vector<int64_t> gid;
    // boundary layer treatment: one-time set-up data and per-hop tables
    sm::LayerSetup layer;
    int resolveBlocks = 1;
    bool doBoundary = false;            // boundary point smoothing enabled (smgpu_enable_boundary_smoothing)
    int64_t boundaryCounts[4] = {0, 0, 0, 0};
    std::vector<uint8_t> boundaryClass; // per point, bits as Dev::bClass
    std::vector<sm::Patch> patches;     // patch table of the mesh (boundary set-up needs it after create)
    bool useTiles = false; // fused geometry kernel over sm::GeomTiles
    bool tilesF = false;   // its second generation (k_geom_tiles_f: run-time strides, fused face-angle filter)
    bool tilesUniform = false, tilesHavePairs = false;
    int64_t tileListedFaces = 0, tileListedPoints = 0;
    bool usePointTiles = false; // k_predict_tiles / k_edge_tiles over sm::PointTiles
    size_t predictSmem = 0, edgeSmem = 0;
    int64_t ptListedPoints = 0, ptListedCells = 0;
    size_t tileSmem = 0;
    int tileMinBlocks = 2; // resident blocks per SM the kernel variant is compiled for (register budget)
    bool doLayers = false;
    bool anyLayerPatch = false;
    bool layersParallel = false, layersReady = false; // processor mesh: set-up runs in smgpu_comm_init
    sm::PolyMesh layerMesh;                           // patches + faces for that set-up
    std::vector<int32_t> patchLayerFlags;
    int *dHops = nullptr, *dPointToOuter = nullptr;
    double *dLayerLength = nullptr, *dLayerBlend = nullptr;
    P4 *normalsTmp = nullptr;
    // params.renumber: storage order = Morton order; old label of every stored point / cell
    std::vector<int32_t> pointOldOfNew, cellOldOfNew;
    smgpu_params prmRequested;              // as passed by the caller (negative = reference default)
    double meshMinEdge = 0, meshMaxEdge = 0; // getMeshStats; global (all-reduced) in multi-rank runs
    // option defaults of src/smoothMesh.C:1861-1865 from the (global) minimum edge length
    void resolveParams()
    {
        const int dev = prm.device;
        prm = prmRequested;
        prm.device = dev;
        if (prm.min_edge_length < 0)
            prm.min_edge_length = 0.5 * meshMinEdge;
        if (prm.max_step_length < 0)
            prm.max_step_length = 0.3 * prm.min_edge_length;
        applyParams();
    }

    // optional per-kernel timing (CUDA events on the launch stream)
    enum { K_FACE_GEOM, K_CELL, K_PREDICT, K_EDGE, K_FACE_CUR, K_COMPACT, K_FACE_TESTS, K_FACE_RESOLVE, K_COMMIT, K_EXCHANGE, K_LAYER, K_GEOM_TILES, K_X_PACK, K_X_MERGE, K_X_FROZEN, K_X_FINISH, K_NUM };
    bool profiling = false;
    std::vector<cudaEvent_t> evPool;
    std::vector<std::pair<int, size_t>> evUse; // (kernel id, index of start event)
    size_t evNext = 0;
    double profMs[K_NUM] = {0};
    int64_t profLaunches[K_NUM] = {0};
    void profBegin(int k)
    {
        if (!profiling)
            return;
        if (evNext + 2 > evPool.size())
        {
            evPool.resize(evPool.size() + 256);
            for (size_t i = evPool.size() - 256; i < evPool.size(); ++i)
                CK(cudaEventCreate(&evPool[i]));
        }
        evUse.push_back({k, evNext});
        CK(cudaEventRecord(evPool[evNext], stream));
    }
    void profEnd(int nLaunches)
    {
        if (!profiling)
            return;
        CK(cudaEventRecord(evPool[evNext + 1], stream));
        profLaunches[evUse.back().first] += nLaunches;
        evNext += 2;
    }
    void profCollect()
    {
        for (auto &u : evUse)
        {
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, evPool[u.second], evPool[u.second + 1]));
            profMs[u.first] += ms;
        }
        evUse.clear();
        evNext = 0;
    }

    template <class T> T *dalloc(size_t n)
    {
        void *p = nullptr;
        CK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T) + 16));
        allocs.push_back(p);
        return (T *)p;
    }
    template <class T, class A> T *upload(const std::vector<T, A> &v)
    {
        T *p = dalloc<T>(v.size());
        if (v.empty())
            return p;
        if (deferUploads)
            uploadJobs.push_back({p, v.data(), v.size() * sizeof(T)});
        else if (pipeReady && !uploadThread.joinable() && v.size() * sizeof(T) >= ((size_t)1 << 20))
        { // during the set-up, after the helper thread has finished: same pipe, from this thread
            CK(pipe.copy(p, v.data(), v.size() * sizeof(T), 8));
            CK(pipe.finish()); // the source may be a temporary
        }
        else
            CK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
        return p;
    }
    // One-time uploads go through a small page-locked double buffer: the host copy of chunk k + 1 into it (a few
    // threads) runs while the DMA of chunk k is in flight, several times the rate of cudaMemcpy from pageable memory.
    struct PinnedPipe
    {
        static constexpr size_t CHUNK = (size_t)32 << 20;
        char *buf[2] = {nullptr, nullptr};
        cudaEvent_t ev[2] = {nullptr, nullptr};
        bool busy[2] = {false, false};
        cudaStream_t s = nullptr;
        int next = 0;
        cudaError_t init()
        {
            const bool timing = getenv("SMGPU_TIMING") && atoi(getenv("SMGPU_TIMING")) != 0;
            auto wall = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
            const double t0 = wall();
            struct Report
            {
                bool on;
                double t0;
                decltype(wall) &w;
                ~Report()
                {
                    if (on)
                        fprintf(stderr, "[smgpu create]   page-locked double buffer  %.3f s\n", w() - t0);
                }
            } report{timing, t0, wall};
            cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
            for (int i = 0; i < 2 && e == cudaSuccess; ++i)
            {
                e = cudaMallocHost((void **)&buf[i], CHUNK);
                if (e == cudaSuccess)
                    e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
            }
            return e;
        }
        cudaError_t copy(void *dst, const void *src, size_t bytes, int threads)
        {
            for (size_t off = 0; off < bytes; off += CHUNK)
            {
                const size_t n = std::min(CHUNK, bytes - off);
                const int b = next;
                next ^= 1;
                if (busy[b])
                {
                    const cudaError_t e = cudaEventSynchronize(ev[b]);
                    if (e != cudaSuccess)
                        return e;
                }
                const char *from = static_cast<const char *>(src) + off;
                char *to = buf[b];
#pragma omp parallel for schedule(static) num_threads(threads)
                for (int64_t i = 0; i < (int64_t)n; i += 1 << 20)
                    memcpy(to + i, from + i, std::min<size_t>((size_t)1 << 20, n - i));
                cudaError_t e = cudaMemcpyAsync(static_cast<char *>(dst) + off, to, n, cudaMemcpyHostToDevice, s);
                if (e == cudaSuccess)
                    e = cudaEventRecord(ev[b], s);
                if (e != cudaSuccess)
                    return e;
                busy[b] = true;
            }
            return cudaSuccess;
        }
        // device -> host: chunk k + 2 is in flight while chunk k is copied out of the buffer
        cudaError_t copyOut(void *dstHost, const void *srcDev, size_t bytes, int threads)
        {
            cudaError_t e = finish();
            const size_t nChunks = (bytes + CHUNK - 1) / CHUNK;
            auto issue = [&](size_t k) {
                const size_t off = k * CHUNK, n = std::min(CHUNK, bytes - off);
                cudaError_t x = cudaMemcpyAsync(buf[k & 1], static_cast<const char *>(srcDev) + off, n, cudaMemcpyDeviceToHost, s);
                return x == cudaSuccess ? cudaEventRecord(ev[k & 1], s) : x;
            };
            for (size_t k = 0; k < std::min<size_t>(2, nChunks) && e == cudaSuccess; ++k)
                e = issue(k);
            for (size_t k = 0; k < nChunks && e == cudaSuccess; ++k)
            {
                e = cudaEventSynchronize(ev[k & 1]);
                if (e != cudaSuccess)
                    break;
                const size_t off = k * CHUNK, n = std::min(CHUNK, bytes - off);
                const char *from = buf[k & 1];
                char *to = static_cast<char *>(dstHost) + off;
#pragma omp parallel for schedule(static) num_threads(threads)
                for (int64_t i = 0; i < (int64_t)n; i += 1 << 20)
                    memcpy(to + i, from + i, std::min<size_t>((size_t)1 << 20, n - i));
                if (k + 2 < nChunks)
                    e = issue(k + 2);
            }
            busy[0] = busy[1] = false;
            next = 0;
            return e;
        }
        cudaError_t finish() { return s ? cudaStreamSynchronize(s) : cudaSuccess; }
        void destroy()
        {
            for (int i = 0; i < 2; ++i)
            {
                if (buf[i])
                    cudaFreeHost(buf[i]);
                if (ev[i])
                    cudaEventDestroy(ev[i]);
                buf[i] = nullptr, ev[i] = nullptr, busy[i] = false;
            }
            if (s)
                cudaStreamDestroy(s);
            s = nullptr;
        }
    } pipe;
    bool pipeReady = false;
    // Uploads of the connectivity tables run on a helper thread while the host builds the geometry tiles
    // (smgpu_create): upload() only allocates and queues while deferUploads is set, startUploads() hands the queue to
    // the thread, joinUploads() waits for it and reports its error.  The sources must stay untouched until then.
    struct UploadJob
    {
        void *dst;
        const void *src;
        size_t bytes;
    };
    std::vector<UploadJob> uploadJobs;
    bool deferUploads = false;
    std::thread uploadThread;
    cudaError_t uploadErr = cudaSuccess;
    void ensurePipe()
    {
        if (pipeReady)
            return;
        CK(pipe.init());
        pipeReady = true;
    }
    void startUploads(int device)
    {
        deferUploads = false;
        ensurePipe();
        std::vector<UploadJob> jobs;
        jobs.swap(uploadJobs);
        uploadThread = std::thread([this, device, jobs]() {
            cudaError_t e = cudaSetDevice(device);
            for (size_t i = 0; i < jobs.size() && e == cudaSuccess; ++i)
                e = pipe.copy(jobs[i].dst, jobs[i].src, jobs[i].bytes, 4);
            if (e == cudaSuccess)
                e = pipe.finish();
            uploadErr = e;
        });
    }
    void joinUploads()
    {
        if (uploadThread.joinable())
            uploadThread.join();
        CK(uploadErr);
    }
    void applyParams()
    {
        d.minEdgeLength = prm.min_edge_length;
        d.maxStepLength = prm.max_step_length;
        d.relStepFrac = prm.rel_step_frac;
        d.relTol = prm.rel_tol;
        d.smallAngle = M_PI * prm.min_angle_deg / 180.0; // src/smoothMesh.C:921, :1364
        d.largeAngle = M_PI * prm.max_angle_deg / 180.0; // :1365
        d.totalMinFreeze = prm.total_min_freeze;
        d.edgeAngleConstraint = prm.edge_angle_constraint;
        d.faceAngleConstraint = prm.face_angle_constraint;
        d.geometryVariant = prm.geometry_variant;
        // guard-banded cosine-space filters; disabled near the ends of the
        // angle range, where the literal path is always taken
        const double guard = 1e-9;
        d.edgeCosT = std::cos(d.smallAngle) - guard;
        d.edgeFilter = (d.edgeCosT > -0.9999 && d.edgeCosT < 0.9999) ? 1 : 0;
        d.faceCosHi = std::cos(d.smallAngle) - guard;
        d.faceCosLo = std::cos(d.largeAngle) + guard;
        d.faceFilter = (d.smallAngle > 1e-3 && d.largeAngle < M_PI - 1e-3 && d.smallAngle < d.largeAngle &&
                        d.faceCosLo < d.faceCosHi)
                           ? 1
                           : 0;
        // boundary layer treatment: per-hop target edge length and blending fraction
        // (blendWithOrthogonalPoints, src/orthogonalBoundaryBlending.C:547-555; its maxLayers is maxLayers + 1, :2300)
        doLayers = anyLayerPatch && prm.layer_max_blending_fraction > 1e-15; // :2025, SMALL
        d.layers = doLayers ? 1 : 0;
        d.bsmooth = doBoundary ? 1 : 0;
        d.normalsOn = (doLayers || doBoundary) ? 1 : 0;
        if (doLayers && dLayerLength)
        {
            const double maxLayers = prm.max_layers + 1, minLayers = prm.min_layers;
            const double edgeLen = prm.layer_edge_length < 0 ? prm.min_edge_length : prm.layer_edge_length;
            std::vector<double> len(layer.maxHop + 2, 0.0), bl(layer.maxHop + 2, 0.0);
            for (int nHops = 1; nHops <= layer.maxHop + 1; ++nHops)
            {
                const double lim = (double(nHops - 1) < maxLayers) ? double(nHops - 1) : maxLayers;
                len[nHops] = edgeLen * std::pow(prm.layer_expansion_ratio, (double)(int)lim);
                const double slope = -prm.layer_max_blending_fraction / (maxLayers - minLayers);
                const double y0 = -slope * maxLayers;
                const double y = y0 + slope * nHops;
                const double t = (y < prm.layer_max_blending_fraction) ? y : prm.layer_max_blending_fraction;
                bl[nHops] = (0.0 > t) ? 0.0 : t;
            }
            cudaMemcpy(dLayerLength, len.data(), len.size() * sizeof(double), cudaMemcpyHostToDevice);
            cudaMemcpy(dLayerBlend, bl.data(), bl.size() * sizeof(double), cudaMemcpyHostToDevice);
        }
        d.cosSmallF = (float)std::cos(d.smallAngle);
        d.cosLargeF = (float)std::cos(d.largeAngle);
        // Single-precision filter levels: every one of them converts DIFFERENCES taken in FP64
        // (relative to the edge's end point / the point's proposal / a tile-local origin), with the error budget
        // computed at run time from the actual vectors, so none has a precondition on the size or position of the
        // mesh, nor on the points staying inside the initial hull.
        const bool f32 = !getenv("SMGPU_NO_F32");
        d.faceFilter32 = d.faceFilter && f32;
        d.edgeFilter32 = d.edgeFilter && f32;
        if (noFilters)
            d.edgeFilter = d.faceFilter = d.faceFilter32 = d.edgeFilter32 = 0;
        // face-angle filter fused into k_geom_tiles_f (per-cell certificates); SMGPU_NO_FUSED_FILTER keeps the
        // per-edge kernel
        d.fusedFaceFilter = (tilesF && tilesHavePairs && d.faceFilter && f32 &&
                             !(getenv("SMGPU_NO_FUSED_FILTER") && atoi(getenv("SMGPU_NO_FUSED_FILTER")) != 0))
                                ? 1
                                : 0;
        if (d.fusedFaceFilter)
            d.faceFilter32 = 0; // the per-edge kernel is not used (smgpu_op_edge_face_angles is literal)
        d.edgeTile32 = (usePointTiles && d.edgeFilter && f32) ? 1 : 0;
        d.faceMean64 = (!d.fusedFaceFilter && d.faceFilter) ? 1 : 0;
        ensureBuffers();
    }
    // Buffers only some configurations touch, allocated when a configuration that needs them is selected:
    // the 64-byte face records (two-kernel geometry; boundary face areas for the point normals), the FP64
    // face-mean table and the 48-byte edge records of the per-edge face-angle filter.  At 368^3 they would add 21 GB.
    void ensureBuffers()
    {
        if (!d.pts)
            return; // smgpu_create has not allocated the state yet
        const bool perEdgeFilter = !d.fusedFaceFilter && d.faceFilter;
        if (!d.faceGeo && (!useTiles || anyLayerPatch || doBoundary))
            d.faceGeo = dalloc<P4>(2 * topo.F);
        if (!d.faceMean && d.faceMean64)
            d.faceMean = dalloc<P4>(topo.F);
        if (!d.edgeRec && perEdgeFilter)
        {
            sm::buildEdgeRecords(topo);
            d.edgeRec = (const int4 *)upload(topo.edgeRec);
            sm::Vec<int32_t>().swap(topo.edgeRec);
        }
    }
    bool noFilters = false; // SMGPU_NO_FILTERS=1: always take the literal path (testing aid)
    void ensureStats(int n)
    {
        if (n <= statCap)
            return;
        statCap = std::max(n, 1024);
        d.statRes = dalloc<double>(statCap);
        d.statFrozen = dalloc<long long>(statCap);
        d.statCap = statCap;
    }
    // page-locked staging buffer for point uploads / downloads (host <-> device at full PCIe rate)
    P4 *staging = nullptr;
    size_t stagingCap = 0;
    P4 *stage(size_t n)
    {
        if (n > stagingCap)
        {
            if (staging)
                cudaFreeHost(staging);
            CK(cudaMallocHost((void **)&staging, n * sizeof(P4)));
            stagingCap = n;
        }
        return staging;
    }
    // device-side staging of a pointField (24 bytes per point) and the internal-point flags for k_unpack_points
    double *dXyz = nullptr;
    size_t dXyzCap = 0;
    uint8_t *dIsInternal = nullptr;
    double *devXyz(size_t n)
    {
        if (n > dXyzCap)
        {
            dXyz = dalloc<double>(3 * n); // (an earlier, smaller buffer stays in `allocs` until destroy)
            dXyzCap = n;
        }
        return dXyz;
    }
    // Upload of the caller's pointField through the page-locked double buffer (the host copy of chunk k + 1 runs while
    // the DMA of chunk k is in flight); the 32-byte records are assembled on the device.
    void setPoints(const double *pts)
    {
        const size_t P = (size_t)topo.P;
        if (!pointOldOfNew.empty())
        { // renumbered storage: the permutation is applied on the host
            P4 *h = stage(std::max<size_t>(topo.P, topo.C));
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < topo.P; ++i)
            {
                const int64_t o = pointOldOfNew[i];
                h[i] = {pts[3 * o], pts[3 * o + 1], pts[3 * o + 2], topo.isInternal[i] ? 1.0 : 0.0};
            }
            CK(cudaMemcpyAsync(d.pts, h, P * sizeof(P4), cudaMemcpyHostToDevice, stream));
            CK(cudaStreamSynchronize(stream));
            return;
        }
        double *dx = devXyz(P);
        if (!dIsInternal)
            dIsInternal = upload(topo.isInternal);
        ensurePipe();
        CK(pipe.copy(dx, pts, 3 * P * sizeof(double), omp_get_max_threads()));
        CK(pipe.finish());
        k_unpack_points<<<grid(d.P, 256), 256, 0, stream>>>(d, dx, dIsInternal);
        CK(cudaStreamSynchronize(stream));
    }

    static int grid(int64_t n, int block) { return (int)std::max<int64_t>(1, (n + block - 1) / block); }

    // ---- kernel launches (one method per reference operator) ----
    void launchFaceGeom()
    {
        profBegin(K_FACE_GEOM);
        k_face_geom<<<grid(d.F, 256), 256, 0, stream>>>(d);
        profEnd(1);
        ++launches;
    }
    void launchCells()
    {
        profBegin(K_CELL);
        k_cell_centres<<<grid(d.C, 128), 128, 0, stream>>>(d);
        profEnd(1);
        ++launches;
    }
    // OpenFOAM's demand-driven geometry after movePoints: face centres/areas, then cell centres --
    // one fused launch over the geometry tiles, or the two per-face / per-cell kernels
    // (SMGPU_NO_TILES=1, or a mesh with a cell too large for a tile)
    void launchCellCentres()
    {
        if (useTiles && tilesF)
        {
            profBegin(K_GEOM_TILES);
            if (d.fusedFaceFilter)
                CK(cudaMemsetAsync(d.suspect, 0, (size_t)d.P, stream));
            k_geom_tiles_f<<<d.nTiles, SMK_TILE_CELLS, tileSmem, stream>>>(d);
            profEnd(1);
            ++launches;
            return;
        }
        if (useTiles)
        {
            profBegin(K_GEOM_TILES);
            if (tileMinBlocks == 3)
                k_geom_tiles<3><<<d.nTiles, SMK_TILE_CELLS, SMK_TILE_SMEM, stream>>>(d);
            else
                k_geom_tiles<2><<<d.nTiles, SMK_TILE_CELLS, SMK_TILE_SMEM, stream>>>(d);
            profEnd(1);
            ++launches;
            return;
        }
        launchFaceGeom();
        launchCells();
    }
    // start of an iteration (src/smoothMesh.C:2262-2269): geometry and, with layer treatment, the
    // boundary point normals of :2266 (they need the boundary face areas of the current mesh)
    void launchGeometry()
    {
        if (useTiles)
        {
            launchCellCentres();
            if (doLayers || doBoundary)
                launchLayerNormals();
            return;
        }
        launchFaceGeom();
        if (doLayers || doBoundary)
            launchLayerNormals();
        launchCells();
    }
    void launchPredict()
    {
        profBegin(K_PREDICT);
        if (usePointTiles)
            k_predict_tiles<<<d.nPointTiles, SMK_PT_THREADS, predictSmem, stream>>>(d);
        else
            k_predict<<<grid(d.P, 128), 128, 0, stream>>>(d);
        profEnd(1);
        ++launches;
    }
    void launchLayerNormals()
    {
        profBegin(K_LAYER);
        k_layer_normals<<<grid(d.P, 128), 128, 0, stream>>>(d, 1);
        profEnd(1);
        ++launches;
    }
    void launchLayerBlend()
    {
        profBegin(K_LAYER);
        k_layer_blend<<<grid(d.P, 128), 128, 0, stream>>>(d);
        profEnd(1);
        ++launches;
    }
    // set-up call of calculateBoundaryPointNormals + propagateOuterNeighInfo's normal copies
    // (src/smoothMesh.C:2219-2220) for the mesh currently on the device
    // device tables of the boundary point normals / layer treatment for `layer`
    void allocLayerTables()
    {
        d.normals = dalloc<P4>(topo.P);
        normalsTmp = dalloc<P4>(topo.P);
        d.hops = dHops = upload(layer.hops);
        d.pointToOuter = dPointToOuter = upload(layer.pointToOuter);
        d.normalSrc = upload(layer.normalSrc);
        d.bfOff = upload(layer.bfOff);
        d.bf = upload(layer.bf);
        const int hopCap = std::max(layer.maxHop, prm.max_layers + 1) + 2;
        dLayerLength = dalloc<double>(hopCap);
        dLayerBlend = dalloc<double>(hopCap);
        d.layerLength = dLayerLength;
        d.layerBlend = dLayerBlend;
    }
    // :2312-2355: projection of the boundary points, prismatic projection, third step clamp
    void launchBoundary()
    {
        profBegin(K_LAYER);
        if (d.nBPoints > 0)
            k_boundary_project<<<grid(d.nBPoints, 128), 128, 0, stream>>>(d);
        k_boundary_finish<<<grid(d.P, 128), 128, 0, stream>>>(d);
        profEnd(2);
        launches += 2;
    }
    void initLayerNormals()
    {
        if (!doLayers && !doBoundary)
            return;
        if (layersParallel)
        {
            // processor mesh: the set-up synchronises with the other ranks (sm::commSetupLayers)
            if (comm && sm::commIsGroupMember(comm))
                throw std::runtime_error("the layer set-up of an in-process group is collective: re-create the group "
                                         "instead of calling smgpu_set_points on one member");
            if (comm)
                sm::commSetupLayers(comm, this);
            return;
        }
        CK(cudaMemsetAsync(d.normals, 0, topo.P * sizeof(P4), stream));
        CK(cudaMemsetAsync(d.done, 0, sizeof(int), stream));
        k_face_geom<<<grid(d.F, 256), 256, 0, stream>>>(d);
        k_layer_normals<<<grid(d.P, 128), 128, 0, stream>>>(d, 1);
        CK(cudaMemcpyAsync(normalsTmp, d.normals, topo.P * sizeof(P4), cudaMemcpyDeviceToDevice, stream));
        k_layer_init_normals<<<grid(d.P, 128), 128, 0, stream>>>(d, normalsTmp);
        CK(cudaStreamSynchronize(stream));
    }
    void launchEdgeConstraints()
    {
        profBegin(K_EDGE);
        if (usePointTiles)
            k_edge_tiles<<<d.nPointTiles, SMK_PT_THREADS, edgeSmem, stream>>>(d);
        else
            k_edge_constraints<<<grid(d.P, 128), 128, 0, stream>>>(d);
        profEnd(1);
        ++launches;
    }
    void launchFaceAngle(double *dbgMin = nullptr, double *dbgMax = nullptr)
    {
        launchFaceCurrent(dbgMin, dbgMax);
        launchFaceResolve();
    }
    // first half of restrictFaceAngleDeterioration: the angles of the *current* mesh and the ordered list
    // of active points (:1252-1270, :1367-1368) -- independent of the proposed positions, so a multi-rank
    // iteration runs it while the interface exchange is in flight
    void launchFaceCurrent(double *dbgMin = nullptr, double *dbgMax = nullptr)
    {
        const int nChunks = grid(d.P, SMK_CHUNK);
        profBegin(K_FACE_CUR);
        if (d.fusedFaceFilter && !dbgMin)
            k_face_suspects<<<grid(d.P, 4 * SMK_SUSPECT_WINDOW), 128, 0, stream>>>(d); // the pairs k_geom_tiles_f did not certify
        else
            k_face_current<<<grid(d.E, 128), 128, 0, stream>>>(d, dbgMin, dbgMax);
        profEnd(1);
        profBegin(K_COMPACT);
        k_active_count<<<nChunks, 256, 0, stream>>>(d);
        k_active_scan<<<1, 256, 0, stream>>>(d, nChunks);
        k_active_fill<<<nChunks, 256, 0, stream>>>(d);
        k_face_clear<<<148, 128, 0, stream>>>(d);
        profEnd(4);
        launches += 5;
    }
    // second half: the tests on the proposed positions and the replay of the worklist (:1347-1434)
    void launchFaceResolve()
    {
        profBegin(K_FACE_TESTS);
        k_face_tests<<<148 * 4, 128, 0, stream>>>(d);
        profEnd(1);
        profBegin(K_FACE_RESOLVE);
        {
            void *args[] = {(void *)&d};
            CK(cudaLaunchCooperativeKernel((const void *)k_face_resolve, dim3(resolveBlocks), dim3(128), args, 0, stream));
        }
        profEnd(1);
        launches += 2;
    }
    void launchCommit()
    {
        profBegin(K_COMMIT);
        k_commit<<<grid(d.P, 256), 256, 0, stream>>>(d, nullptr);
        profEnd(1);
        ++launches;
    }
    void resetControl()
    {
        CK(cudaMemsetAsync(d.done, 0, sizeof(int), stream));
        CK(cudaMemsetAsync(d.iter, 0, sizeof(int), stream));
    }
    // conditions on which the reference aborts inside the loop, raised by a kernel through d.errFlag
    int checkErrFlag()
    {
        int errFlag = 0;
        CK(cudaMemcpy(&errFlag, d.errFlag, sizeof(int), cudaMemcpyDeviceToHost));
        if (!errFlag)
            return SMGPU_OK;
        static const char *msg[] = {"", "Sanity broken, outerNeighCoord is undefined for an interface point "
                                        "(src/orthogonalBoundaryBlending.C:537)",
                                    "Internal sanity check failed: Did not find any edges with the required string index "
                                    "(src/boundaryPointSmoothing.C:257)",
                                    "pointNormal is zero for a smoothing surface point (src/boundaryPointSmoothing.C:691)",
                                    "Did not find surface intersection for a boundary point (src/boundaryPointSmoothing.C:934)",
                                    "A smoothing surface point has zero point normal (src/orthogonalBoundaryBlending.C:611)"};
        CK(cudaMemset(d.errFlag, 0, sizeof(int)));
        if (errFlag == 6)
            return setErr(SMGPU_ERR_COMM, "peer-memory exchange: a rank did not deliver its records within thirty seconds "
                                          "(did every rank call smgpu_iterate with the same arguments?)");
        return setErr(SMGPU_ERR_MESH, msg[(errFlag >= 1 && errFlag <= 5) ? errFlag : 1]);
    }
    // per-iteration log of the last run (what the reference prints at src/smoothMesh.C:2396)
    int fetchStats(int64_t *n_frozen, double *residual, int32_t *iters_done)
    {
        int it = 0;
        CK(cudaMemcpy(&it, d.iter, sizeof(int), cudaMemcpyDeviceToHost));
        if (iters_done)
            *iters_done = it;
        if (it > 0 && residual)
            CK(cudaMemcpy(residual, d.statRes, it * sizeof(double), cudaMemcpyDeviceToHost));
        if (it > 0 && n_frozen)
        {
            std::vector<long long> tmp(it);
            CK(cudaMemcpy(tmp.data(), d.statFrozen, it * sizeof(long long), cudaMemcpyDeviceToHost));
            for (int i = 0; i < it; ++i)
                n_frozen[i] = tmp[i];
        }
        return SMGPU_OK;
    }
};

#include "comm_impl.cuh"

#include <condition_variable>
#include <mutex>
#include <thread>

namespace sm
{

struct HostBarrier
{
    std::mutex m;
    std::condition_variable cv;
    int n = 1, count = 0, gen = 0;
    void wait()
    {
        std::unique_lock<std::mutex> lk(m);
        const int g = gen;
        if (++count == n)
        {
            count = 0;
            ++gen;
            cv.notify_all();
        }
        else
            cv.wait(lk, [&] { return gen != g; });
    }
};

// In-process group (smgpu_group_*): the processor meshes of one decomposed case as handles of this process,
// all on one device and one stream, driven by one host thread.  The exchanges of an iteration become
// stream-ordered device copies between the members' buffers and one small reduction kernel; the collective
// host steps of the one-time layer set-up run on one short-lived host thread per member.
struct LocalGroup
{
    std::vector<smgpu_handle *> members; // rank order
    cudaStream_t stream = nullptr;
    HostBarrier bar;
    std::vector<const void *> posted; // hostSync: the send array of every member
    double **dRes = nullptr;          // device arrays of the members' redRes / redFrozen pointers
    long long **dFrozen = nullptr;
    std::vector<cudaStream_t> ownStreams; // the members' own streams, restored when the group is destroyed
};

// what the peers of a rank need to know to write into its exchange block; sits at the start of the block
struct XHeader
{
    int magic, rank, nRanks, tuple, nSlots, nNbr;
    int nbrRank[SMK_MAXNBR], nbrOff[SMK_MAXNBR + 1];
    long long offFlagT, offFlagF, offStat, offRecv, offFz;
};

// this rank's half of the plan: no communication, so a failure here cannot leave other ranks waiting
static Comm *commPrepare(smgpu_handle *h, int rank, int nRanks, const int64_t *counts, const int64_t *allGids)
{
    if (h->gid.empty() && !h->topo.procPoints.empty())
        throw std::runtime_error("mesh has processor patches but was created without point_global_id");
    Comm *cm = new Comm;
    try
    {
        std::vector<int64_t> myGids;
        for (int32_t p : h->topo.procPoints)
            myGids.push_back(h->gid[p]);
        std::vector<int64_t> cnt(counts, counts + nRanks);
        int64_t total = 0;
        for (int64_t c : cnt)
            total += c;
        std::vector<int64_t> all(allGids, allGids + total);
        cm->plan = buildExchangePlan(rank, nRanks, h->topo.procPoints, myGids, cnt, all);
        const ExchangePlan &pl = cm->plan;
        if (pl.maxCopies > SMK_MAXCOPIES)
            throw std::runtime_error("an interface point is shared by more ranks than the exchange layer supports");
        CK(cudaSetDevice(h->prm.device));
        smk::CommDev &c = cm->c;
        c.rank = rank;
        c.nSlots = (int)pl.sendPoint.size();
        c.nShared = (int)pl.sharedPoint.size();
        c.sendPoint = h->upload(pl.sendPoint);
        c.sharedPoint = h->upload(pl.sharedPoint);
        c.selfSlot = h->upload(pl.selfSlot);
        c.copyOff = h->upload(pl.copyOff);
        c.copyRank = h->upload(pl.copyRank);
        c.copySlot = h->upload(pl.copySlot);
        c.tuple = h->anyLayerPatch ? SMK_TUPLE_LAYERS : SMK_TUPLE;
        c.sendBuf = h->dalloc<double>((size_t)c.nSlots * SMK_TUPLE_BOUNDARY); // capacity for the largest record
        c.sendFz = h->dalloc<uint8_t>(c.nSlots);
        // the receive side lives in one exchange block that peers can map (smgpu_comm_p2p_*): header (who this
        // rank's neighbours are and where their records go), flag words, statistics slots, receive buffers
        XHeader hdr;
        memset(&hdr, 0, sizeof hdr);
        hdr.magic = 0x534d5832; // "SMX2"
        hdr.rank = rank;
        hdr.nRanks = nRanks;
        hdr.tuple = c.tuple;
        hdr.nSlots = c.nSlots;
        hdr.nNbr = (int)pl.nbrRank.size();
        for (int j = 0; j < hdr.nNbr && j < SMK_MAXNBR; ++j)
        {
            hdr.nbrRank[j] = pl.nbrRank[j];
            hdr.nbrOff[j] = pl.nbrOff[j];
        }
        if (hdr.nNbr <= SMK_MAXNBR)
            hdr.nbrOff[hdr.nNbr] = pl.nbrOff[hdr.nNbr];
        hdr.offFlagT = 512;
        hdr.offFlagF = hdr.offFlagT + SMK_MAXNBR * 8;
        hdr.offStat = hdr.offFlagF + SMK_MAXNBR * 8;
        hdr.offRecv = hdr.offStat + 2 * SMK_MAXRANKS * (long long)sizeof(smk::P2PStat);
        hdr.offFz = hdr.offRecv + (((long long)c.nSlots * SMK_TUPLE_BOUNDARY * 8 + 255) / 256) * 256;
        cm->xblockBytes = (size_t)hdr.offFz + (size_t)c.nSlots + 256;
        static_assert(sizeof(XHeader) <= 512, "exchange block header");
        cm->xblock = h->dalloc<unsigned char>(cm->xblockBytes);
        CK(cudaMemset(cm->xblock, 0, cm->xblockBytes));
        CK(cudaMemcpy(cm->xblock, &hdr, sizeof hdr, cudaMemcpyHostToDevice));
        c.recvBuf = reinterpret_cast<double *>(cm->xblock + hdr.offRecv);
        c.recvFz = cm->xblock + hdr.offFz;
        c.p2p = nullptr;
        c.redRes = h->dalloc<double>(1);
        c.redFrozen = h->dalloc<long long>(1);
    }
    catch (...)
    {
        delete cm;
        throw;
    }
    return cm;
}

static bool commIsGroupMember(const Comm *cm) { return cm && cm->group; }

static void commAttach(Comm *cm, smgpu_handle *h)
{
    h->d.multiRank = 1;
    h->d.locRes = cm->c.redRes;
    h->d.locFrozen = cm->c.redFrozen;
}

// the collective half over NCCL: communicator, global edge-length extrema, layer set-up
static void commConnectNccl(Comm *cm, smgpu_handle *h, int rank, int nRanks, const uint8_t id[128])
{
    CK(cudaSetDevice(h->prm.device));
    ncclUniqueId uid;
    static_assert(sizeof(uid) == 128, "ncclUniqueId size");
    memcpy(&uid, id, 128);
    NCK(nccl().CommInitRank(&cm->nccl, nRanks, uid, rank));
    CK(cudaStreamCreateWithFlags(&cm->xStream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&cm->evPacked, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&cm->evExchanged, cudaEventDisableTiming));
    commAttach(cm, h);
    // getMeshStats' returnReduce(min/max) (src/smoothMesh.C:1527-1528): the option defaults
    // must come from the global edge-length extrema, not this rank's
    double mm[2] = {-h->topo.minEdgeLength, h->topo.maxEdgeLength};
    double *dmm = h->dalloc<double>(2);
    CK(cudaMemcpy(dmm, mm, sizeof mm, cudaMemcpyHostToDevice));
    NCK(nccl().AllReduce(dmm, dmm, 2, ncclDouble, ncclMax, cm->nccl, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(mm, dmm, sizeof mm, cudaMemcpyDeviceToHost));
    h->meshMinEdge = -mm[0];
    h->meshMaxEdge = mm[1];
    h->resolveParams();
    commSetupLayers(cm, h);
}

// syncTools::syncPointList for a host field with N values of T per point (set-up only): the copies of
// every interface point are combined in ascending rank order, like the device-side merge.  NCCL ranks move
// them through the iteration's exchange buffers; the members of an in-process group (one host thread each
// during the set-up) read each other's send arrays between two barriers.
template <class T, int N, class Op> static void hostSync(Comm *cm, smgpu_handle *h, std::vector<T> &field, Op combine)
{
    const ExchangePlan &pl = cm->plan;
    const smk::CommDev &c = cm->c;
    if (c.nSlots == 0 && !cm->group)
        return;
    static_assert(sizeof(T) * N <= SMK_TUPLE * sizeof(double), "record larger than the exchange buffers");
    std::vector<T> send((size_t)c.nSlots * N), recv((size_t)c.nSlots * N);
    for (int i = 0; i < c.nSlots; ++i)
        for (int k = 0; k < N; ++k)
            send[(size_t)i * N + k] = field[(size_t)pl.sendPoint[i] * N + k];
    if (cm->group)
    {
        LocalGroup *g = cm->group;
        g->posted[pl.rank] = send.data();
        g->bar.wait();
        for (size_t j = 0; j < pl.nbrRank.size(); ++j)
        {
            const ExchangePlan &ql = g->members[pl.nbrRank[j]]->comm->plan;
            const size_t jj = std::find(ql.nbrRank.begin(), ql.nbrRank.end(), pl.rank) - ql.nbrRank.begin();
            const size_t cnt = (size_t)(pl.nbrOff[j + 1] - pl.nbrOff[j]) * N;
            if (jj >= ql.nbrRank.size() || (size_t)(ql.nbrOff[jj + 1] - ql.nbrOff[jj]) * N != cnt)
                throw std::runtime_error("internal: exchange plans of two group members disagree");
            memcpy(recv.data() + (size_t)pl.nbrOff[j] * N, (const T *)g->posted[pl.nbrRank[j]] + (size_t)ql.nbrOff[jj] * N,
                   cnt * sizeof(T));
        }
        g->bar.wait(); // every member has read before any send array goes away
    }
    else
    {
        CK(cudaMemcpyAsync(c.sendBuf, send.data(), send.size() * sizeof(T), cudaMemcpyHostToDevice, h->stream));
        haloExchange(cm, h->stream, c.sendBuf, c.recvBuf, N * sizeof(T));
        CK(cudaMemcpyAsync(recv.data(), c.recvBuf, recv.size() * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    for (size_t s = 0; s < pl.sharedPoint.size(); ++s)
    {
        const int32_t p = pl.sharedPoint[s];
        const int cb = pl.copyOff[s], nOther = pl.copyOff[s + 1] - cb;
        T v[N];
        bool first = true, selfDone = false;
        for (int k = 0, o = 0; k < nOther + 1; ++k)
        {
            const T *src;
            if (!selfDone && (o >= nOther || pl.copyRank[cb + o] > pl.rank))
            {
                src = &field[(size_t)p * N];
                selfDone = true;
            }
            else
                src = &recv[(size_t)pl.copySlot[cb + o++] * N];
            if (first)
            {
                for (int q = 0; q < N; ++q)
                    v[q] = src[q];
                first = false;
            }
            else
                combine(v, src);
        }
        for (int q = 0; q < N; ++q)
            field[(size_t)p * N + q] = v[q];
    }
}

// One-time boundary layer set-up of a processor mesh (src/smoothMesh.C:2215-2221 under -parallel).
// Collective: every rank of the communicator calls it (smgpu_comm_init, smgpu_set_points).
static void commSetupLayers(Comm *cm, smgpu_handle *h)
{
    // whenever a layer patch exists (not only while the blending fraction is positive), so that raising
    // layer_max_blending_fraction later through smgpu_set_params finds the set-up done
    if (!(h->anyLayerPatch || h->doBoundary) || !h->layersParallel)
        return;
    Dev &d = h->d;
    const int64_t P = h->topo.P;
    CK(cudaSetDevice(h->prm.device));
    // this rank's share of the set-up call of calculateBoundaryPointNormals on the current mesh
    CK(cudaMemsetAsync(d.normals, 0, P * sizeof(P4), h->stream));
    CK(cudaMemsetAsync(d.done, 0, sizeof(int), h->stream));
    k_face_geom<<<smgpu_handle::grid(d.F, 256), 256, 0, h->stream>>>(d);
    k_layer_normals<<<smgpu_handle::grid(d.P, 128), 128, 0, h->stream>>>(d, 0);
    std::vector<P4> tmp(P);
    CK(cudaMemcpyAsync(tmp.data(), d.normals, P * sizeof(P4), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    std::vector<double> normals(3 * P);
    for (int64_t p = 0; p < P; ++p)
        normals[3 * p] = tmp[p].x, normals[3 * p + 1] = tmp[p].y, normals[3 * p + 2] = tmp[p].z;
    LayerSync sync;
    sync.maxInt = [&](std::vector<int32_t> &f) {
        hostSync<int32_t, 1>(cm, h, f, [](int32_t *x, const int32_t *y) { x[0] = (x[0] > y[0]) ? x[0] : y[0]; });
    };
    sync.sumInt = [&](std::vector<int32_t> &f) {
        hostSync<int32_t, 1>(cm, h, f, [](int32_t *x, const int32_t *y) { x[0] = x[0] + y[0]; });
    };
    sync.sumVec = [&](std::vector<double> &f) {
        hostSync<double, 3>(cm, h, f, [](double *x, const double *y) {
            x[0] = x[0] + y[0], x[1] = x[1] + y[1], x[2] = x[2] + y[2];
        });
    };
    sync.maxMagSqrVec = [&](std::vector<double> &f) {
        hostSync<double, 3>(cm, h, f, [](double *x, const double *y) {
            const double mx = x[0] * x[0] + x[1] * x[1] + x[2] * x[2], my = y[0] * y[0] + y[1] * y[1] + y[2] * y[2];
            if (!(mx >= my))
                x[0] = y[0], x[1] = y[1], x[2] = y[2];
        });
    };
    h->layer = buildLayerSetupParallel(h->layerMesh, h->topo, h->patchLayerFlags, h->prm.max_layers, normals, sync);
    for (int64_t p = 0; p < P; ++p)
        tmp[p] = P4{normals[3 * p], normals[3 * p + 1], normals[3 * p + 2], 0.0};
    CK(cudaMemcpy(d.normals, tmp.data(), P * sizeof(P4), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->dHops, h->layer.hops.data(), P * sizeof(int32_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->dPointToOuter, h->layer.pointToOuter.data(), P * sizeof(int32_t), cudaMemcpyHostToDevice));
    h->layersReady = true;
    h->applyParams(); // per-hop tables for the synchronised hop counts
}

// getMeshStats' reductions for the boundary point smoothing set-up of this rank (src/smoothMesh.C:1527-1538: the
// perimeter is hi.x - lo.x + hi.y - lo.y + hi.z PLUS lo.z of the global bounding box) and the max-synchronisation
// of the hop counts.  Collective.
static void commBoundaryParallel(Comm *cm, smgpu_handle *h, const std::vector<double> &points, BoundaryParallel &par)
{
    double lo[3], hi[3];
    meshBoundingBox(h->topo, points, lo, hi);
    double box[6] = {-lo[0], -lo[1], -lo[2], hi[0], hi[1], hi[2]}; // one max-reduction
    if (cm->group)
    {
        LocalGroup *g = cm->group;
        g->posted[cm->plan.rank] = box;
        g->bar.wait();
        double all[6];
        for (int k = 0; k < 6; ++k)
            all[k] = ((const double *)g->posted[0])[k];
        for (size_t r = 1; r < g->members.size(); ++r)
            for (int k = 0; k < 6; ++k)
                all[k] = std::max(all[k], ((const double *)g->posted[r])[k]);
        g->bar.wait();
        for (int k = 0; k < 6; ++k)
            box[k] = all[k];
    }
    else
    {
        double *dbox = h->dalloc<double>(6);
        CK(cudaMemcpy(dbox, box, sizeof box, cudaMemcpyHostToDevice));
        NCK(nccl().AllReduce(dbox, dbox, 6, ncclDouble, ncclMax, cm->nccl, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaMemcpy(box, dbox, sizeof box, cudaMemcpyDeviceToHost));
    }
    par.meshMinEdgeLength = h->meshMinEdge;
    par.meshPerimeter = box[3] - (-box[0]) + box[4] - (-box[1]) + box[5] + (-box[2]);
    par.maxInt = [cm, h](std::vector<int32_t> &f) {
        hostSync<int32_t, 1>(cm, h, f, [](int32_t *x, const int32_t *y) { x[0] = (x[0] > y[0]) ? x[0] : y[0]; });
    };
}

static void commDestroy(Comm *cm)
{
    if (!cm)
        return;
    if (cm->xStream)
    {
        cudaStreamSynchronize(cm->xStream);
        cudaStreamDestroy(cm->xStream);
    }
    if (cm->evPacked)
        cudaEventDestroy(cm->evPacked);
    if (cm->evExchanged)
        cudaEventDestroy(cm->evExchanged);
    if (cm->evCommitted)
        cudaEventDestroy(cm->evCommitted);
    if (cm->evFinished)
        cudaEventDestroy(cm->evFinished);
    if (cm->nccl)
        nccl().CommDestroy(cm->nccl);
    for (void *p : cm->ipcMapped)
        cudaIpcCloseMemHandle(p);
    delete cm;
}

// Maps the peers' exchange blocks and fills the device-side table of the peer-memory exchange.  bases[r]:
// device pointer to rank r's block as seen from this process (own block for r == rank).
static void commConnectPeers(Comm *cm, smgpu_handle *h, const std::vector<unsigned char *> &bases)
{
    const ExchangePlan &pl = cm->plan;
    const int nRanks = pl.nRanks, nNbr = (int)pl.nbrRank.size();
    if (nNbr > SMK_MAXNBR || nRanks > SMK_MAXRANKS)
        throw std::runtime_error("peer-memory exchange: more neighbour ranks / ranks than its tables hold");
    smk::P2PDev x;
    memset(&x, 0, sizeof x);
    x.nNbr = nNbr;
    x.nRanks = nRanks;
    x.rank = pl.rank;
    XHeader mine;
    CK(cudaMemcpy(&mine, cm->xblock, sizeof mine, cudaMemcpyDeviceToHost));
    x.flagT = reinterpret_cast<unsigned long long *>(cm->xblock + mine.offFlagT);
    x.flagF = reinterpret_cast<unsigned long long *>(cm->xblock + mine.offFlagF);
    x.stat = reinterpret_cast<smk::P2PStat *>(cm->xblock + mine.offStat);
    std::vector<unsigned char> slotNbr(pl.sendPoint.size());
    for (int j = 0; j < nNbr; ++j)
    {
        x.nbrOff[j] = pl.nbrOff[j];
        for (int32_t i = pl.nbrOff[j]; i < pl.nbrOff[j + 1]; ++i)
            slotNbr[i] = (unsigned char)j;
    }
    x.nbrOff[nNbr] = pl.nbrOff[nNbr];
    for (int r = 0; r < nRanks; ++r)
    {
        XHeader ph;
        CK(cudaMemcpy(&ph, bases[r], sizeof ph, cudaMemcpyDefault));
        if (ph.magic != 0x534d5832 || ph.rank != r || ph.nRanks != nRanks)
            throw std::runtime_error("peer-memory exchange: the block of rank " + std::to_string(r) + " does not carry its header");
        x.peerStat[r] = reinterpret_cast<smk::P2PStat *>(bases[r] + ph.offStat);
        const auto it = std::find(pl.nbrRank.begin(), pl.nbrRank.end(), r);
        if (it == pl.nbrRank.end())
            continue;
        const int j = (int)(it - pl.nbrRank.begin());
        int jj = -1;
        for (int k = 0; k < ph.nNbr; ++k)
            if (ph.nbrRank[k] == pl.rank)
                jj = k;
        if (jj < 0 || ph.nbrOff[jj + 1] - ph.nbrOff[jj] != pl.nbrOff[j + 1] - pl.nbrOff[j])
            throw std::runtime_error("peer-memory exchange: exchange plans of two ranks disagree");
        x.peerRecv[j] = reinterpret_cast<double *>(bases[r] + ph.offRecv);
        x.peerSlot0[j] = ph.nbrOff[jj];
        x.peerRecvFz[j] = bases[r] + ph.offFz + ph.nbrOff[jj];
        x.peerFlagT[j] = reinterpret_cast<unsigned long long *>(bases[r] + ph.offFlagT) + jj;
        x.peerFlagF[j] = reinterpret_cast<unsigned long long *>(bases[r] + ph.offFlagF) + jj;
    }
    x.slotNbr = h->upload(slotNbr);
    unsigned long long one = 1;
    unsigned long long *ep = h->dalloc<unsigned long long>(1);
    CK(cudaMemcpy(ep, &one, 8, cudaMemcpyHostToDevice));
    x.epoch = ep;
    unsigned int *cnt = h->dalloc<unsigned int>(2);
    CK(cudaMemset(cnt, 0, 8));
    x.packDone = cnt;
    x.fzDone = cnt + 1;
    smk::P2PDev *dx = h->dalloc<smk::P2PDev>(1);
    CK(cudaMemcpy(dx, &x, sizeof x, cudaMemcpyHostToDevice));
    cm->c.p2p = dx;
    cm->p2p = true;
    std::vector<uint8_t> shared((size_t)h->topo.P + 8, 0);
    for (int32_t p : pl.sharedPoint)
        shared[p] = 1;
    cm->sharedFlag = h->upload(shared);
    if (!cm->evCommitted)
    {
        CK(cudaEventCreateWithFlags(&cm->evCommitted, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&cm->evFinished, cudaEventDisableTiming));
    }
}

// ---- one iteration with the interface exchanges (src/smoothMesh.C:2257-2399 under -parallel), in the
// phases between which the copies of the interface points travel ----
// A: geometry, then the local predictor tuple of every interface point.  With layer treatment the interface
// records carry the normals of the previous iteration, so they are packed before k_layer_normals replaces
// those; interface points get their normal, blend and second clamp in k_shared_merge, which overwrites
// whatever the point-wise kernels wrote for them.
static void commPhasePack(Comm *cm, smgpu_handle *h, bool ownStream = false)
{
    const smk::CommDev &c = cm->c;
    h->launchCellCentres();
    if (cm->finishPending)
    { // the previous iteration's statistics / stop flag (k_finish_iter on the exchange stream)
        CK(cudaStreamWaitEvent(h->stream, cm->evFinished, 0));
        cm->finishPending = false;
    }
    h->profBegin(smgpu_handle::K_X_PACK);
    if (ownStream)
    { // the records only need the geometry: packed (and, in peer-memory mode, delivered) beside the predictor
        CK(cudaEventRecord(cm->evPacked, h->stream));
        CK(cudaStreamWaitEvent(cm->xStream, cm->evPacked, 0));
        if (c.nSlots > 0)
            k_shared_pack<<<smgpu_handle::grid(c.nSlots, 128), 128, 0, cm->xStream>>>(h->d, c);
        CK(cudaEventRecord(cm->evExchanged, cm->xStream));
    }
    else if (c.nSlots > 0)
        k_shared_pack<<<smgpu_handle::grid(c.nSlots, 128), 128, 0, h->stream>>>(h->d, c);
    h->profEnd(1);
    h->launches += 1;
}
// B1: everything that needs nothing from the exchange (it runs while the tuples are in flight)
static void commPhaseLocal(Comm *cm, smgpu_handle *h)
{
    if (h->doLayers || h->doBoundary)
        h->launchLayerNormals(); // :2266 (the interface points get theirs, with all copies, in k_shared_merge)
    h->launchPredict();
    if (h->doLayers)
        h->launchLayerBlend();
    if (h->doBoundary)
        h->launchBoundary(); // :2307-2356; the points shared between ranks are redone by k_shared_merge
    if (h->prm.face_angle_constraint)
        h->launchFaceCurrent(); // current-mesh half of the face-angle constraint
}
// B2: merge of the copies, constraints, this rank's freeze flags of the interface points
static void commPhaseConstrain(Comm *cm, smgpu_handle *h)
{
    const smk::CommDev &c = cm->c;
    h->profBegin(smgpu_handle::K_X_MERGE);
    if (c.nShared > 0)
        k_shared_merge<<<smgpu_handle::grid(c.nShared, 64), 64, 0, h->stream>>>(h->d, c);
    h->profEnd(1);
    h->launches += 1;
    h->launchEdgeConstraints();
    if (h->prm.face_angle_constraint)
        h->launchFaceResolve();
    h->profBegin(smgpu_handle::K_X_FROZEN);
    if (c.nSlots > 0)
        k_frozen_pack<<<smgpu_handle::grid(c.nSlots, 128), 128, 0, h->stream>>>(h->d, c);
    h->profEnd(1);
    h->launches += 1;
}
// C in peer-memory mode: the points that are not shared between ranks are committed while the other ranks' freeze
// flags are still on their way; the shared points follow after the OR (:2374) and publish the statistics
static void commPhaseCommitSplit(Comm *cm, smgpu_handle *h)
{
    const smk::CommDev &c = cm->c;
    h->profBegin(smgpu_handle::K_COMMIT);
    k_commit<<<smgpu_handle::grid(h->d.P, 256), 256, 0, h->stream>>>(h->d, cm->sharedFlag);
    h->profEnd(1);
    h->profBegin(smgpu_handle::K_X_FROZEN);
    k_commit_shared<<<smgpu_handle::grid(std::max(c.nShared, 1), 256), 256, 0, h->stream>>>(h->d, c);
    h->profEnd(1);
    h->launches += 2;
}
// C: OR of the freeze flags (:2374), restore + residual + movePoints
static void commPhaseCommit(Comm *cm, smgpu_handle *h)
{
    const smk::CommDev &c = cm->c;
    h->profBegin(smgpu_handle::K_X_FROZEN);
    if (c.nSlots > 0)
        k_frozen_or<<<smgpu_handle::grid(c.nSlots, 128), 128, 0, h->stream>>>(h->d, c);
    h->profEnd(1);
    h->launches += 1;
    h->launchCommit();
}
// D: statistics and stop flag from the reduced residual / count (:2396-2405)
static void commPhaseFinish(Comm *cm, smgpu_handle *h)
{
    h->profBegin(smgpu_handle::K_X_FINISH);
    k_finish_iter<<<1, SMK_MAXRANKS, 0, h->stream>>>(h->d, cm->c);
    h->profEnd(1);
    h->launches += 1;
}

static void commIterate(Comm *cm, smgpu_handle *h)
{
    const smk::CommDev &c = cm->c;
    if (cm->p2p)
    { // peer-memory exchange: the producer kernels write into the peers' blocks, the consumer kernels wait for
      // the flags; nothing but this rank's own kernels on this rank's stream
        // with layer treatment the pack reads the normals k_layer_normals is about to replace: same stream then
        const bool beside = !h->doLayers && !h->doBoundary && cm->xStream;
        commPhasePack(cm, h, beside);
        commPhaseLocal(cm, h);
        if (beside)
            CK(cudaStreamWaitEvent(h->stream, cm->evExchanged, 0)); // this rank's own records (read by the merge)
        commPhaseConstrain(cm, h);
        commPhaseCommitSplit(cm, h);
        // the reduction over the ranks runs beside the next iteration's geometry pass (which only writes scratch
        // data and is harmless after the stop flag): its wait for the slowest rank is off the critical path
        CK(cudaEventRecord(cm->evCommitted, h->stream));
        CK(cudaStreamWaitEvent(cm->xStream, cm->evCommitted, 0));
        h->profBegin(smgpu_handle::K_X_FINISH);
        k_finish_iter<<<1, SMK_MAXRANKS, 0, cm->xStream>>>(h->d, cm->c);
        CK(cudaEventRecord(cm->evFinished, cm->xStream));
        h->profEnd(1);
        h->launches += 1;
        cm->finishPending = true;
        return;
    }
    commPhasePack(cm, h);
    // the predictor exchange runs on its own stream while the main stream works on what does not need it
    h->profBegin(smgpu_handle::K_EXCHANGE);
    CK(cudaEventRecord(cm->evPacked, h->stream));
    CK(cudaStreamWaitEvent(cm->xStream, cm->evPacked, 0));
    haloExchange(cm, cm->xStream, c.sendBuf, c.recvBuf, c.tuple * sizeof(double));
    CK(cudaEventRecord(cm->evExchanged, cm->xStream));
    h->profEnd(0);
    commPhaseLocal(cm, h);
    h->profBegin(smgpu_handle::K_EXCHANGE);
    CK(cudaStreamWaitEvent(h->stream, cm->evExchanged, 0));
    h->profEnd(0);
    commPhaseConstrain(cm, h);
    h->profBegin(smgpu_handle::K_EXCHANGE);
    haloExchange(cm, h->stream, c.sendFz, c.recvFz, 1);
    h->profEnd(0);
    commPhaseCommit(cm, h);
    h->profBegin(smgpu_handle::K_EXCHANGE);
    NCK(nccl().GroupStart());
    NCK(nccl().AllReduce(c.redRes, c.redRes, 1, ncclDouble, ncclMax, cm->nccl, h->stream));
    NCK(nccl().AllReduce(c.redFrozen, c.redFrozen, 1, ncclInt64, ncclSum, cm->nccl, h->stream));
    NCK(nccl().GroupEnd());
    h->profEnd(0);
    commPhaseFinish(cm, h);
    // no host round trip: the stop flag is global (it comes from the all-reduced residual), so every
    // rank keeps launching the same sequence and the kernels turn into no-ops once it is raised
}

// ---- in-process group ----
__global__ void k_group_reduce(int n, double *const *res, long long *const *frz)
{
    if (threadIdx.x != 0 || blockIdx.x != 0)
        return;
    double m = *res[0];
    long long s = *frz[0];
    for (int r = 1; r < n; ++r)
    {
        const double v = *res[r];
        m = (v > m) ? v : m; // returnReduce(maxOp), :1567
        s += *frz[r];        // returnReduce(sumOp), :2396
    }
    for (int r = 0; r < n; ++r)
    {
        *res[r] = m;
        *frz[r] = s;
    }
}
// which = 0: predictor tuples, 1: freeze flags
static void groupExchange(LocalGroup *g, int which)
{
    for (smgpu_handle *h : g->members)
    {
        const Comm *cm = h->comm;
        const ExchangePlan &pl = cm->plan;
        const size_t eb = which == 0 ? cm->c.tuple * sizeof(double) : 1;
        char *recv = which == 0 ? (char *)cm->c.recvBuf : (char *)cm->c.recvFz;
        for (size_t j = 0; j < pl.nbrRank.size(); ++j)
        {
            const Comm *qm = g->members[pl.nbrRank[j]]->comm;
            const ExchangePlan &ql = qm->plan;
            const size_t jj = std::find(ql.nbrRank.begin(), ql.nbrRank.end(), pl.rank) - ql.nbrRank.begin();
            const char *send = which == 0 ? (const char *)qm->c.sendBuf : (const char *)qm->c.sendFz;
            CK(cudaMemcpyAsync(recv + (size_t)pl.nbrOff[j] * eb, send + (size_t)ql.nbrOff[jj] * eb,
                               (size_t)(pl.nbrOff[j + 1] - pl.nbrOff[j]) * eb, cudaMemcpyDeviceToDevice, g->stream));
        }
    }
}
static void groupIterate(LocalGroup *g)
{
    for (smgpu_handle *h : g->members)
        commPhasePack(h->comm, h);
    groupExchange(g, 0);
    for (smgpu_handle *h : g->members)
    {
        commPhaseLocal(h->comm, h);
        commPhaseConstrain(h->comm, h);
    }
    groupExchange(g, 1);
    for (smgpu_handle *h : g->members)
        commPhaseCommit(h->comm, h);
    k_group_reduce<<<1, 32, 0, g->stream>>>((int)g->members.size(), g->dRes, g->dFrozen);
    for (smgpu_handle *h : g->members)
        commPhaseFinish(h->comm, h);
}

} // namespace sm

struct smgpu_group
{
    sm::LocalGroup g;
};

extern "C"
{

    const char *smgpu_last_error(void) { return g_err.c_str(); }
    const char *smgpu_version(void) { return "smoothmesh_b200 0.1 (sm_100a, fp64, fmad=off)"; }

    int smgpu_device_count(int32_t *n)
    {
        int c = 0;
        if (cudaGetDeviceCount(&c) != cudaSuccess)
            c = 0;
        if (n)
            *n = c;
        return c > 0 ? SMGPU_OK : setErr(SMGPU_ERR_CUDA, "no CUDA device available");
    }

    void smgpu_default_params(smgpu_params *p)
    {
        // src/smoothMesh.C:1861-1914
        p->min_edge_length = -1.0;
        p->max_step_length = -1.0;
        p->rel_step_frac = 0.5;
        p->min_angle_deg = 35.0;
        p->max_angle_deg = 160.0;
        p->rel_tol = 0.02;
        p->total_min_freeze = 0;
        p->edge_angle_constraint = 1;
        p->face_angle_constraint = 1;
        p->geometry_variant = 0;
        p->device = 0;
        p->renumber = 0;
        p->layer_max_blending_fraction = 0.3; // :1892-1905
        p->layer_edge_length = -1.0;
        p->layer_expansion_ratio = 1.3;
        p->min_layers = 1;
        p->max_layers = 4;
    }

    int smgpu_create(const smgpu_mesh_desc *md, const smgpu_params *params, smgpu_handle **out)
    {
        if (!md || !params || !out)
            return setErr(SMGPU_ERR_ARG, "null argument");
        *out = nullptr;
        int nDev = 0;
        if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0)
            return setErr(SMGPU_ERR_CUDA, "no CUDA device available (libsmgpu has no CPU fallback)");
        if (params->device < 0 || params->device >= nDev)
            return setErr(SMGPU_ERR_ARG, "device ordinal out of range");
        smgpu_handle *h = new smgpu_handle;
        // the CUDA context of the device is created on a helper thread while the host builds the connectivity
        const int device = params->device;
        std::thread contextThread([device]() {
            if (cudaSetDevice(device) == cudaSuccess)
                cudaFree(nullptr);
        });
        struct Joiner
        {
            std::thread &t;
            ~Joiner()
            {
                if (t.joinable())
                    t.join();
            }
        } contextJoiner{contextThread};
        // SMGPU_TIMING=1: wall time of the set-up phases on stderr (topology.cpp prints its own)
        const bool timing = getenv("SMGPU_TIMING") && atoi(getenv("SMGPU_TIMING")) != 0;
        auto wall = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        double tPrev = wall();
        auto tick = [&](const char *what) {
            if (!timing)
                return;
            const double tNow = wall();
            fprintf(stderr, "[smgpu create] %-28s %.3f s\n", what, tNow - tPrev);
            tPrev = tNow;
        };
        try
        {
            sm::PolyMesh m;
            sm::parCopy(m.points, md->points, 3 * (size_t)md->n_points);
            sm::parCopy(m.faceOffsets, md->face_offsets, (size_t)md->n_faces + 1);
            sm::parCopy(m.faceVerts, md->face_verts, (size_t)md->face_offsets[md->n_faces]);
            sm::parCopy(m.owner, md->owner, (size_t)md->n_faces);
            sm::parCopy(m.neighbour, md->neighbour, (size_t)md->n_internal_faces);
            m.nCells = md->n_cells;
            for (int i = 0; i < md->n_patches; ++i)
            {
                sm::Patch p;
                p.name = "patch" + std::to_string(i);
                p.type = md->patch_kind[i] == SMGPU_PATCH_PROCESSOR ? "processor"
                         : md->patch_kind[i] == SMGPU_PATCH_EMPTY   ? "empty"
                                                                    : "patch";
                p.start = md->patch_start[i];
                p.size = md->patch_size[i];
                m.patches.push_back(p);
            }
            if (md->point_global_id)
                m.pointGlobalId.assign(md->point_global_id, md->point_global_id + md->n_points);
            std::vector<int32_t> patchLayer(md->n_patches, 0);
            if (md->patch_layer)
                for (int i = 0; i < md->n_patches; ++i)
                {
                    patchLayer[i] = md->patch_layer[i] != 0;
                    h->anyLayerPatch = h->anyLayerPatch || patchLayer[i];
                }
            tick("copy of the caller's mesh");
            try
            {
                if (params->renumber)
                    m = sm::renumberMorton(m, h->pointOldOfNew, h->cellOldOfNew);
                if (md->point_global_id)
                    h->gid = m.pointGlobalId;
                h->topo = sm::buildTopology(m);
            }
            catch (const std::exception &e)
            {
                delete h;
                return setErr(SMGPU_ERR_MESH, e.what());
            }
            tick("topology (total)");
            const sm::Topology &t = h->topo;
            h->patches = m.patches;
            h->prm = *params;
            h->prmRequested = *params;
            h->meshMinEdge = t.minEdgeLength;
            h->meshMaxEdge = t.maxEdgeLength;
            for (uint8_t f : t.isInternal)
                h->nInternal += f;

            contextThread.join();
            CK(cudaSetDevice(params->device));
            CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
            CK(cudaEventCreate(&h->ev0));
            CK(cudaEventCreate(&h->ev1));
            Dev &d = h->d;
            memset(&d, 0, sizeof(d));
            d.P = (int)t.P;
            d.C = (int)t.C;
            d.E = (int)t.E;
            d.F = (int)t.F;
            tick("CUDA context, stream");
            // face records, mirrors and edge records are allocated on first need (ensureBuffers)
            d.pts = h->dalloc<P4>(t.P);
            d.newPts = h->dalloc<P4>(t.P);
            d.cellCtr = h->dalloc<P4>(t.C);
            d.frozen = h->dalloc<uint8_t>(t.P + 8);
            h->deferUploads = true; // the tables go up while the tiles are built
            d.pcOff = h->upload(t.pcOff);
            d.pc = h->upload(t.pc);
            d.ppOff = h->upload(t.ppOff);
            d.pp = h->upload(t.pp);
            d.pe = h->upload(t.pe);
            d.cornerOff = h->upload(t.cornerOff);
            d.corner = h->upload(t.corner);
            d.edge = h->upload(t.edge);
            d.efOff = h->upload(t.efOff);
            d.ef = h->upload(t.ef);
            d.ecOff = h->upload(t.ecOff);
            d.ecCell = h->upload(t.ecCell);
            d.ecPair = h->upload(t.ecPair);
            d.faceOff = h->upload(t.faceOff);
            d.faceVerts = h->upload(t.faceVerts);
            d.cfOff = h->upload(t.cfOff);
            d.cf = h->upload(t.cf);
            d.pointRec = (const int4 *)h->upload(t.pointRec);
            tick("  cudaMalloc of the tables");
            h->startUploads(params->device);
            tick("allocation of tables");
            d.uniformFaceSize = t.maxFaceSize;
            for (int64_t f = 0; f < t.F && d.uniformFaceSize; ++f)
                if (t.faceOff[f + 1] - t.faceOff[f] != d.uniformFaceSize)
                    d.uniformFaceSize = 0;
            d.uniformCellFaces = t.C ? t.cfOff[1] - t.cfOff[0] : 0;
            for (int64_t c = 0; c < t.C && d.uniformCellFaces; ++c)
                if (t.cfOff[c + 1] - t.cfOff[c] != d.uniformCellFaces)
                    d.uniformCellFaces = 0;
            d.curMin = h->dalloc<unsigned long long>(t.P);
            d.curMax = h->dalloc<unsigned long long>(t.P);
            d.activeFlag = h->dalloc<uint8_t>(t.P + 8);
            d.selfBits = h->dalloc<uint8_t>(t.P + 8);
            d.pairBits = h->dalloc<uint8_t>(t.pp.size() + 8);
            d.activeList = h->dalloc<int>(t.P);
            d.stack = h->dalloc<int>(t.P);
            d.reach = h->dalloc<int>(t.P);
            d.rootHi = h->dalloc<int>(t.P);
            d.changed = h->dalloc<int>(8);
            CK(cudaMemset(d.changed, 0, 8 * sizeof(int))); // [6]: running count of active points (k_active_count -> k_active_scan)
            {
                // cooperative launch of k_face_resolve: as many blocks as can be co-resident
                int perSm = 0, sms = 0;
                CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_face_resolve, 128, 0));
                CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, params->device));
                h->resolveBlocks = std::max(1, sms * std::min(perSm, 4));
            }
            d.blockCounts = h->dalloc<int>(smgpu_handle::grid(t.P, SMK_CHUNK) + 1);
            d.nActive = h->dalloc<int>(1);
            d.done = h->dalloc<int>(1);
            d.iter = h->dalloc<int>(1);
            d.accMaxBits = h->dalloc<unsigned long long>(1);
            d.accFrozen = h->dalloc<unsigned long long>(1);
            d.blocksDone = h->dalloc<unsigned int>(1);
            CK(cudaMemset(d.nActive, 0, sizeof(int)));
            CK(cudaMemset(d.done, 0, sizeof(int)));
            CK(cudaMemset(d.iter, 0, sizeof(int)));
            CK(cudaMemset(d.accMaxBits, 0, 8));
            CK(cudaMemset(d.accFrozen, 0, 8));
            CK(cudaMemset(d.blocksDone, 0, 4));
            CK(cudaMemset(d.frozen, 0, t.P + 8));
            CK(cudaMemset(d.activeFlag, 0, t.P + 8));
            CK(cudaMemset(d.newPts, 0, t.P * sizeof(P4)));
            h->ensureStats(1024);
            h->noFilters = getenv("SMGPU_NO_FILTERS") && atoi(getenv("SMGPU_NO_FILTERS")) != 0;
            d.nInternalFaces = (int)md->n_internal_faces;
            d.nTiles = 0;
            // The fused tile kernel has a fast path per tile (all faces quadrilaterals, all cells hexahedra) and a
            // generic one; on polyhedral meshes the generic path measured slower than the per-face / per-cell /
            // per-edge kernels (Kelvin cells: 3.62 vs 1.93 ms), so tiles are used where hexahedra dominate -- a
            // hex-dominant mesh with prisms or polyhedra keeps the fast path on every all-hex tile -- or when forced.
            const bool noTiles = getenv("SMGPU_NO_TILES") && atoi(getenv("SMGPU_NO_TILES")) != 0;
            const bool forceTiles = getenv("SMGPU_FORCE_TILES") && atoi(getenv("SMGPU_FORCE_TILES")) != 0;
            int64_t hexLike = 0;
#pragma omp parallel for schedule(static) reduction(+ : hexLike)
            for (int64_t c = 0; c < t.C; ++c)
            {
                bool hex = t.cfOff[c + 1] - t.cfOff[c] == 6;
                for (int32_t k = t.cfOff[c]; k < t.cfOff[c + 1] && hex; ++k)
                {
                    const int32_t f = t.cf[k] & 0x7fffffff;
                    hex = t.faceOff[f + 1] - t.faceOff[f] == 4;
                }
                hexLike += hex ? 1 : 0;
            }
            if (!noTiles && (forceTiles || 2 * hexLike >= t.C))
            {
                const sm::GeomTiles G = sm::buildGeomTiles(m, t, SMK_TILE_CELLS, SMK_TILE_FACES, SMK_TILE_POINTS);
                tick("  tiles built");
                h->joinUploads(); // the tables went up meanwhile; the tile arrays follow through the same pipe
                tick("  wait for the table uploads");
                if (G.nTiles > 0)
                {
                    h->useTiles = true;
                    d.nTiles = G.nTiles;
                    const bool oldTiles = getenv("SMGPU_OLD_TILES") && atoi(getenv("SMGPU_OLD_TILES")) != 0;
                    const bool allUniform = G.nUniformCells == t.C;
                    d.tileCellOff = h->upload(G.tileCellOff);
                    d.tileCells = h->upload(G.tileCells);
                    d.tileFaceOff = h->upload(G.tileFaceOff);
                    d.tileFaces = h->upload(G.tileFaces);
                    d.tilePointOff = h->upload(G.tilePointOff);
                    d.tilePoints = h->upload(G.tilePoints);
                    if (!allUniform || oldTiles)
                    { // the offset tables and variable-length references: only tiles off the fast path read them
                        d.slotOff = h->upload(G.slotOff);
                        d.slotRef = h->upload(G.slotRef);
                        d.faceRefOff = h->upload(G.faceRefOff);
                        d.faceRef = h->upload(G.faceRef);
                    }
                    d.tileUFaceOff = h->upload(G.tileUFaceOff);
                    d.tileUCellOff = h->upload(G.tileUCellOff);
                    d.uFaceRef = (const uint2 *)h->upload(G.uFaceRef);
                    d.uSlotRef = (const unsigned int *)h->upload(G.uSlotRef);
                    d.hexRec = (const uint2 *)h->upload(G.hexRec);
                    // certificates: every cell has its (edge, cell) pairs either as a canonical record (uniform
                    // tiles) or in the pair lists (the other tiles); none at all if some cell is not closed
                    bool pairsComplete = !G.tileUCellOff.empty();
                    for (int32_t k = 0; k < G.nTiles && pairsComplete; ++k)
                        pairsComplete = G.tileUCellOff[k] >= 0 || !G.cellEdgeOff.empty();
                    h->tilesHavePairs = pairsComplete;
                    h->tilesUniform = allUniform;
                    h->tileListedFaces = (int64_t)G.tileFaces.size();
                    h->tileListedPoints = (int64_t)G.tilePoints.size();
                    d.uniformCellEdges = G.uniformCellEdges;
                    d.tileSF = (G.maxTileFaces + 31) / 32 * 32;
                    d.tileSP = (G.maxTilePoints + 31) / 32 * 32;
                    d.tileSE = (G.nUniformCells > 0 && pairsComplete) ? 4 * G.maxTileCells : 0; // 32-byte record per cell
                    h->tileSmem = tileSmemBytes(d.tileSF, d.tileSP, d.tileSE);
                    if (!G.cellEdgeOff.empty())
                    {
                        d.cellEdgeOff = h->upload(G.cellEdgeOff);
                        d.cellEdgeRef = (const uint2 *)h->upload(G.cellEdgeRef);
                    }
                    d.suspect = h->dalloc<uint8_t>(t.P + 16); // k_face_suspects reads whole 16-byte words
                    CK(cudaMemset(d.suspect, 0, t.P + 16));
                    h->tilesF = !oldTiles && h->tileSmem <= 110 * 1024;
                    // the attribute belongs to the function, not to this handle: always the device's opt-in maximum,
                    // so that a later handle with smaller tiles does not lower it under an earlier handle's launches
                    int smemOptin = 0;
                    CK(cudaDeviceGetAttribute(&smemOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, params->device));
                    CK(cudaFuncSetAttribute(k_geom_tiles_f, cudaFuncAttributeMaxDynamicSharedMemorySize, smemOptin));
                    CK(cudaFuncSetAttribute(k_geom_tiles<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMK_TILE_SMEM));
                    CK(cudaFuncSetAttribute(k_geom_tiles<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMK_TILE_SMEM));
                    if (getenv("SMGPU_TILE_MINB"))
                        h->tileMinBlocks = atoi(getenv("SMGPU_TILE_MINB")) == 3 ? 3 : 2;
                }
            }
            // Brick order for the threads of the per-point gather kernels (the storage order is unchanged)
            if (!noTiles && getenv("SMGPU_POINT_ORDER") && atoi(getenv("SMGPU_POINT_ORDER")) != 0)
            {
                const sm::PointTiles T = sm::buildPointTiles(m, t, 256, 1 << 15, 1 << 15);
                if (T.nTiles > 0)
                {
                    std::vector<int32_t> order;
                    order.reserve(t.P);
                    for (int32_t k = 0; k < T.nTiles; ++k)
                        for (int32_t i = 0; i < T.ownOff[k + 1] - T.ownOff[k]; ++i)
                            order.push_back(T.halo[T.haloOff[k] + i]);
                    if ((int64_t)order.size() == t.P)
                        d.pointOrder = h->upload(order);
                }
            }
            // Per-point kernels on point tiles (k_predict_tiles / k_edge_tiles): built to the same parity bar, but at
            // 200^3 they measured slower than the per-point gather kernels (0.51 / 0.56 ms against 0.50 / 0.46 ms;
            // profiles/r2_ncu_point_tiles_n200.txt: a third of their time is the label -> point gather chain ahead of
            // the first barrier, the rest is issue-bound on ~1000 instructions per point either way), so they are an
            // option (SMGPU_POINT_TILES=1), not the default.
            if (!noTiles && getenv("SMGPU_POINT_TILES") && atoi(getenv("SMGPU_POINT_TILES")) != 0)
            {
                const sm::PointTiles T = sm::buildPointTiles(m, t, SMK_PT_THREADS, SMK_PT_ROUNDS * SMK_PT_THREADS, SMK_PT_ROUNDS * SMK_PT_THREADS);
                if (T.nTiles > 0)
                {
                    d.nPointTiles = T.nTiles;
                    d.ptOwnOff = h->upload(T.ownOff);
                    d.ptHaloOff = h->upload(T.haloOff);
                    d.ptHalo = h->upload(T.halo);
                    d.ptCellOff = h->upload(T.cellOff);
                    d.ptCell = h->upload(T.cell);
                    d.ptRec = (const uint4 *)h->upload(T.rec);
                    d.ptSH = (T.maxHalo + 31) / 32 * 32;
                    d.ptSC = (T.maxCells + 31) / 32 * 32;
                    h->predictSmem = predictTileSmem(d.ptSH, d.ptSC);
                    h->edgeSmem = edgeTileSmem(d.ptSH);
                    h->ptListedPoints = (int64_t)T.halo.size();
                    h->ptListedCells = (int64_t)T.cell.size();
                    int smemOptin = 0;
                    CK(cudaDeviceGetAttribute(&smemOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, params->device));
                    CK(cudaFuncSetAttribute(k_predict_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, smemOptin));
                    CK(cudaFuncSetAttribute(k_edge_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, smemOptin));
                    h->usePointTiles = true;
                }
            }
            tick("work space, tiles");
            h->joinUploads();
            tick("upload of tables (remainder)");
            // share-a-cell bits of the point records (kernels.cuh shareCellRec), from the tables now on the device
            if (!(getenv("SMGPU_NO_SHARE_MASK") && atoi(getenv("SMGPU_NO_SHARE_MASK")) != 0))
                k_share_mask<<<smgpu_handle::grid(d.P, 128), 128, 0, h->stream>>>(d);
            d.errFlag = h->dalloc<int>(1);
            CK(cudaMemset(d.errFlag, 0, sizeof(int)));
            if (h->anyLayerPatch)
            {
                h->layer = sm::buildLayerSetup(m, t, patchLayer, params->max_layers);
                if (!t.procPoints.empty())
                {
                    // decomposed case: hop counts, maps and set-up normals are synchronised between the
                    // ranks (smgpu_comm_init); what is uploaded below is overwritten there
                    h->layersParallel = true;
                    h->layerMesh.faceOffsets = m.faceOffsets;
                    h->layerMesh.faceVerts = m.faceVerts;
                    h->layerMesh.patches = m.patches;
                    h->patchLayerFlags = patchLayer;
                }
                h->allocLayerTables();
            }
            h->setPoints(md->points); // also fixes the single-precision mirror's origin and error bound
            h->resolveParams();
            h->initLayerNormals();
            CK(cudaDeviceSynchronize());
            tick("points, parameters");
        }
        catch (const std::exception &e)
        {
            smgpu_destroy(h);
            return setErr(SMGPU_ERR_CUDA, e.what());
        }
        *out = h;
        return SMGPU_OK;
    }

    int smgpu_enable_boundary_smoothing(smgpu_handle *h, const smgpu_boundary_geometry *g, const int32_t *patch_smoothing,
                                        double internal_smoothing_blending_fraction)
    {
        if (!h || !g || !patch_smoothing)
            return setErr(SMGPU_ERR_ARG, "null argument");
        if (!h->comm && !h->topo.procPoints.empty())
            return setErr(SMGPU_ERR_ARG, "boundary point smoothing on a processor mesh: its set-up is collective, call "
                                         "smgpu_comm_init (or create the in-process group) first");
        if (h->comm && h->comm->p2p)
            return setErr(SMGPU_ERR_ARG, "enable boundary point smoothing before smgpu_comm_p2p_connect");
        if (!h->pointOldOfNew.empty())
            return setErr(SMGPU_ERR_ARG, "boundary point smoothing cannot be combined with params.renumber");
        bool any = false;
        for (size_t i = 0; i < h->patches.size(); ++i)
            any = any || patch_smoothing[i] != 0;
        if (!any)
            return SMGPU_OK; // no smoothing patches: the feature stays off, like the reference (:2082-2086)
        try
        {
            CK(cudaSetDevice(h->prm.device));
            const sm::Topology &t = h->topo;
            Dev &d = h->d;
            // the classification works on the mesh as it is now
            std::vector<P4> cur(t.P);
            CK(cudaMemcpy(cur.data(), d.pts, t.P * sizeof(P4), cudaMemcpyDeviceToHost));
            std::vector<double> points(3 * (size_t)t.P);
            for (int64_t p = 0; p < t.P; ++p)
                points[3 * p] = cur[p].x, points[3 * p + 1] = cur[p].y, points[3 * p + 2] = cur[p].z;
            sm::PolyMesh pm; // patches only: the set-up reads faces from the topology tables
            pm.patches = h->patches;
            sm::EdgeMesh ie, te;
            ie.points.assign(g->init_points, g->init_points + 3 * g->n_init_points);
            ie.edges.assign(g->init_edges, g->init_edges + 2 * g->n_init_edges);
            ie.finish();
            te.points.assign(g->target_points, g->target_points + 3 * g->n_target_points);
            te.edges.assign(g->target_edges, g->target_edges + 2 * g->n_target_edges);
            te.finish();
            sm::TriSurface surf;
            surf.points.assign(g->surface_points, g->surface_points + 3 * g->n_surface_points);
            surf.tris.assign(g->surface_tris, g->surface_tris + 3 * g->n_surface_tris);
            std::vector<int32_t> ps(patch_smoothing, patch_smoothing + h->patches.size());
            const double layerEdgeLength = h->prm.layer_edge_length < 0 ? h->prm.min_edge_length : h->prm.layer_edge_length;
            sm::BoundarySetup B;
            try
            {
                std::vector<int32_t> cornerIO, featureIO;
                if (g->is_corner_point)
                    cornerIO.assign(g->is_corner_point, g->is_corner_point + t.P);
                if (g->is_feature_edge_point)
                    featureIO.assign(g->is_feature_edge_point, g->is_feature_edge_point + t.P);
                sm::BoundaryParallel par;
                if (h->comm)
                    sm::commBoundaryParallel(h->comm, h, points, par);
                B = sm::buildBoundarySetup(pm, t, points, ie, te, surf, ps, layerEdgeLength, cornerIO, featureIO, h->comm ? &par : nullptr);
            }
            catch (const std::exception &e)
            {
                return setErr(SMGPU_ERR_MESH, e.what());
            }
            if (!h->anyLayerPatch)
            { // boundary point normals are needed without any layer patch too (:2215-2221)
                sm::PolyMesh faces;
                faces.patches = h->patches;
                faces.faceOffsets = t.faceOff;
                faces.faceVerts = t.faceVerts;
                h->layer = sm::buildLayerSetup(faces, t, std::vector<int32_t>(h->patches.size(), 0), h->prm.max_layers);
                if (!t.procPoints.empty())
                { // decomposed case: the normals' set-up is synchronised between the ranks (sm::commSetupLayers)
                    h->layersParallel = true;
                    h->layerMesh = faces;
                    h->patchLayerFlags.assign(h->patches.size(), 0);
                }
                h->allocLayerTables();
            }
            auto toP4 = [](const std::vector<double> &xyz) {
                std::vector<P4> out(xyz.size() / 3);
                for (size_t i = 0; i < out.size(); ++i)
                    out[i] = P4{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0.0};
                return out;
            };
            std::vector<uint8_t> cls(t.P, 0);
            for (int64_t p = 0; p < t.P; ++p)
                cls[p] = (uint8_t)((B.isCorner[p] ? 1 : 0) | (B.isFeatureEdge[p] ? 2 : 0) | (B.isSmoothingSurface[p] ? 4 : 0) |
                                   (B.isConnectedToInternal[p] ? 8 : 0));
            std::vector<P4> corner(B.boundaryPoints.size());
            std::vector<int32_t> bString(B.boundaryPoints.size());
            for (size_t b = 0; b < B.boundaryPoints.size(); ++b)
            {
                const int32_t p = B.boundaryPoints[b];
                corner[b] = P4{B.cornerPoints[3 * (size_t)p], B.cornerPoints[3 * (size_t)p + 1], B.cornerPoints[3 * (size_t)p + 2], 0.0};
                bString[b] = B.pointStrings[p];
            }
            d.bClass = h->upload(cls);
            h->boundaryClass = cls;
            d.sharp = h->dalloc<uint8_t>(t.P + 8);
            CK(cudaMemset(d.sharp, 0, t.P + 8));
            d.bPoints = h->upload(B.boundaryPoints);
            d.nBPoints = (int)B.boundaryPoints.size();
            d.cornerPts = h->upload(corner);
            d.bString = h->upload(bString);
            d.bInner = h->upload(B.pointToInner);
            d.tePts = h->upload(toP4(B.targetEdges.points));
            d.teEdges = h->upload(B.targetEdges.edges);
            d.teString = h->upload(B.targetEdgeStrings);
            d.nTargetEdges = (int)B.targetEdges.nEdges();
            d.surfPts = h->upload(toP4(surf.points));
            d.surfTris = h->upload(surf.tris);
            d.nSurfTris = (int)surf.nTris();
            d.nBvhNodes = 0;
            if (surf.nTris() > 8 && !(getenv("SMGPU_NO_BVH") && atoi(getenv("SMGPU_NO_BVH")) != 0))
            { // acceleration structure of the surface ray casts (the search tree of the reference's findLine)
                const sm::TriangleBvh bvh = sm::buildTriangleBvh(surf);
                d.bvhBox = h->upload(bvh.box);
                d.bvhLeft = h->upload(bvh.left);
                d.bvhRight = h->upload(bvh.right);
                d.bvhFirst = h->upload(bvh.first);
                d.bvhCount = h->upload(bvh.count);
                d.bvhOrder = h->upload(bvh.order);
                d.nBvhNodes = (int)bvh.right.size();
            }
            d.distanceTolerance = B.distanceTolerance;
            d.internalFraction = internal_smoothing_blending_fraction;
            h->boundaryCounts[0] = B.nCorners;
            h->boundaryCounts[1] = B.nFeatureEdgePoints;
            h->boundaryCounts[2] = B.nSmoothingSurfacePoints;
            h->boundaryCounts[3] = B.nStrings + 1;
            h->doBoundary = true;
            if (h->comm)
                h->comm->c.tuple = SMK_TUPLE_BOUNDARY; // the interface records carry the feature edge sums and the inner neighbour
            h->applyParams();
            // the set-up call of calculateBoundaryPointNormals (:2219) with the sharp flags; collective on a processor mesh
            if (h->comm && h->layersParallel)
                sm::commSetupLayers(h->comm, h);
            else
                h->initLayerNormals();
            CK(cudaDeviceSynchronize());
        }
        catch (const std::exception &e)
        {
            return setErr(SMGPU_ERR_CUDA, e.what());
        }
        return SMGPU_OK;
    }

    int smgpu_get_boundary_classes(smgpu_handle *h, int32_t *is_corner_point, int32_t *is_feature_edge_point)
    {
        if (!h)
            return setErr(SMGPU_ERR_ARG, "null handle");
        if (h->boundaryClass.empty())
            return setErr(SMGPU_ERR_ARG, "boundary point smoothing is not enabled");
        for (int64_t p = 0; p < h->topo.P; ++p)
        {
            if (is_corner_point)
                is_corner_point[p] = (h->boundaryClass[p] & 1) ? 1 : 0;
            if (is_feature_edge_point)
                is_feature_edge_point[p] = (h->boundaryClass[p] & 2) ? 1 : 0;
        }
        return SMGPU_OK;
    }

    int smgpu_boundary_counts(smgpu_handle *h, int64_t out[4])
    {
        if (!h || !out)
            return setErr(SMGPU_ERR_ARG, "null argument");
        for (int i = 0; i < 4; ++i)
            out[i] = h->boundaryCounts[i];
        return SMGPU_OK;
    }

    int smgpu_destroy(smgpu_handle *h)
    {
        if (!h)
            return SMGPU_OK;
        if (h->uploadThread.joinable()) // a create that failed half-way
            h->uploadThread.join();
        cudaSetDevice(h->prm.device); // the frees below must hit this handle's context in multi-GPU processes
        if (h->pipeReady)
        {
            h->pipe.finish();
            h->pipe.destroy();
        }
        if (h->comm && !h->comm->group) // members of an in-process group are detached by smgpu_group_destroy
            sm::commDestroy(h->comm);
        for (void *p : h->allocs)
            cudaFree(p);
        if (h->staging)
            cudaFreeHost(h->staging);
        for (cudaEvent_t e : h->evPool)
            cudaEventDestroy(e);
        if (h->ev0)
            cudaEventDestroy(h->ev0);
        if (h->ev1)
            cudaEventDestroy(h->ev1);
        if (h->stream)
            cudaStreamDestroy(h->stream);
        delete h;
        return SMGPU_OK;
    }

    int smgpu_get_params(smgpu_handle *h, smgpu_params *out)
    {
        if (!h || !out)
            return setErr(SMGPU_ERR_ARG, "null argument");
        *out = h->prm;
        return SMGPU_OK;
    }

    int smgpu_set_params(smgpu_handle *h, const smgpu_params *p)
    {
        if (!h || !p)
            return setErr(SMGPU_ERR_ARG, "null argument");
        if (h->anyLayerPatch && p->max_layers != h->prmRequested.max_layers)
            return setErr(SMGPU_ERR_ARG, "max_layers fixes the layer set-up and cannot be changed after smgpu_create");
        h->prmRequested = *p;
        h->resolveParams();
        return SMGPU_OK;
    }

    int smgpu_mesh_stats(smgpu_handle *h, double *min_edge, double *max_edge, int64_t *n_internal_points,
                         int64_t *n_edges)
    {
        if (!h)
            return setErr(SMGPU_ERR_ARG, "null handle");
        if (min_edge)
            *min_edge = h->meshMinEdge;
        if (max_edge)
            *max_edge = h->meshMaxEdge;
        if (n_internal_points)
            *n_internal_points = h->nInternal;
        if (n_edges)
            *n_edges = h->topo.E;
        return SMGPU_OK;
    }

    int smgpu_iterate(smgpu_handle *h, int32_t max_iters, int64_t *n_frozen, double *residual, int32_t *iters_done)
    {
        if (!h || max_iters < 0)
            return setErr(SMGPU_ERR_ARG, "bad argument");
        if (h->doLayers && h->layersParallel && !h->layersReady)
            return setErr(SMGPU_ERR_ARG, "boundary layer treatment on a processor mesh: call smgpu_comm_init first "
                                         "(its set-up synchronises hop counts and normals between the ranks)");
        if (h->comm && h->comm->group)
            return setErr(SMGPU_ERR_ARG, "this handle is a member of an in-process group: call smgpu_group_iterate");
        if (h->comm && !h->comm->nccl)
            return setErr(SMGPU_ERR_COMM, "the communicator of this handle was prepared but never connected, or was aborted");
        try
        {
            CK(cudaSetDevice(h->prm.device));
            h->ensureStats(max_iters);
            h->launches = 0;
            h->resetControl();
            CK(cudaEventRecord(h->ev0, h->stream));
            int done = 0, launched = 0;
            {
                const int chunk = 16; // iterations launched between polls of the stop flag
                while (launched < max_iters && !done)
                {
                    const int n = std::min(chunk, max_iters - launched);
                    for (int i = 0; i < n; ++i)
                    {
                        if (h->comm)
                        { // multi-rank: the same sequence with the interface exchanges in between
                            sm::commIterate(h->comm, h);
                            continue;
                        }
                        h->launchGeometry();
                        h->launchPredict();
                        if (h->doLayers)
                            h->launchLayerBlend(); // :2283-2305
                        if (h->doBoundary)
                            h->launchBoundary(); // :2307-2356
                        h->launchEdgeConstraints();
                        if (h->prm.face_angle_constraint)
                            h->launchFaceAngle();
                        h->launchCommit();
                    }
                    launched += n;
                    if (h->comm && h->comm->finishPending)
                    { // the last iteration's reduction runs on the exchange stream: join it
                        CK(cudaStreamWaitEvent(h->stream, h->comm->evFinished, 0));
                        h->comm->finishPending = false;
                    }
                    if (launched < max_iters)
                    {
                        CK(cudaMemcpyAsync(&done, h->d.done, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
                        CK(cudaStreamSynchronize(h->stream));
                    }
                }
            }
            CK(cudaEventRecord(h->ev1, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            CK(cudaGetLastError());
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
            h->lastMs = ms;
            h->lastLaunches = h->launches;
            if (h->profiling)
                h->profCollect();
            const int rc = h->checkErrFlag();
            if (rc != SMGPU_OK)
                return rc;
            return h->fetchStats(n_frozen, residual, iters_done);
        }
        catch (const std::exception &e)
        {
            return setErr(SMGPU_ERR_CUDA, e.what());
        }
        return SMGPU_OK;
    }

    static int downloadP4(smgpu_handle *h, const P4 *src, int64_t n, double *out, const std::vector<int32_t> &oldOfNew)
    {
        try
        {
            CK(cudaSetDevice(h->prm.device));
            if (!oldOfNew.empty())
            {
                P4 *tmp = h->stage(std::max<size_t>(h->topo.P, h->topo.C));
                CK(cudaMemcpyAsync(tmp, src, n * sizeof(P4), cudaMemcpyDeviceToHost, h->stream));
                CK(cudaStreamSynchronize(h->stream));
#pragma omp parallel for schedule(static)
                for (int64_t i = 0; i < n; ++i)
                {
                    const int64_t o = oldOfNew[i];
                    out[3 * o] = tmp[i].x;
                    out[3 * o + 1] = tmp[i].y;
                    out[3 * o + 2] = tmp[i].z;
                }
                return SMGPU_OK;
            }
            // 24 bytes per point over PCIe through the page-locked double buffer: the host copy of chunk k out of it
            // runs while the DMA of chunk k + 1 is in flight
            double *dx = h->devXyz(std::max<size_t>(h->topo.P, h->topo.C));
            k_pack_points<<<smgpu_handle::grid(n, 256), 256, 0, h->stream>>>(src, dx, (int)n);
            CK(cudaStreamSynchronize(h->stream));
            h->ensurePipe();
            CK(h->pipe.copyOut(out, dx, 3 * (size_t)n * sizeof(double), omp_get_max_threads()));
        }
        catch (const std::exception &e)
        {
            return setErr(SMGPU_ERR_CUDA, e.what());
        }
        return SMGPU_OK;
    }

    int smgpu_get_points(smgpu_handle *h, double *out)
    {
        if (!h || !out)
            return setErr(SMGPU_ERR_ARG, "null argument");
        return downloadP4(h, h->d.pts, h->topo.P, out, h->pointOldOfNew);
    }

    int smgpu_set_points(smgpu_handle *h, const double *in)
    {
        if (!h || !in)
            return setErr(SMGPU_ERR_ARG, "null argument");
        try
        {
            CK(cudaSetDevice(h->prm.device));
            h->setPoints(in);
            h->applyParams();
            h->initLayerNormals(); // a fresh run starts from the set-up normals of the new mesh
        }
        catch (const std::exception &e)
        {
            return setErr(SMGPU_ERR_CUDA, e.what());
        }
        return SMGPU_OK;
    }

    int smgpu_get_frozen(smgpu_handle *h, uint8_t *out)
    {
        if (!h || !out)
            return setErr(SMGPU_ERR_ARG, "null argument");
        if (cudaSetDevice(h->prm.device) != cudaSuccess ||
            cudaMemcpy(out, h->d.frozen, h->topo.P, cudaMemcpyDeviceToHost) != cudaSuccess)
            return setErr(SMGPU_ERR_CUDA, "download failed");
        if (!h->pointOldOfNew.empty())
        {
            std::vector<uint8_t> tmp(out, out + h->topo.P);
            for (int64_t i = 0; i < h->topo.P; ++i)
                out[h->pointOldOfNew[i]] = tmp[i];
        }
        return SMGPU_OK;
    }

    int smgpu_tile_stats(smgpu_handle *h, int64_t out[8])
    {
        if (!h || !out)
            return setErr(SMGPU_ERR_ARG, "null argument");
        out[0] = h->useTiles ? h->d.nTiles : 0;
        out[1] = h->tileListedFaces;
        out[2] = h->tileListedPoints;
        out[3] = (int64_t)h->tileSmem;
        out[4] = h->usePointTiles ? h->d.nPointTiles : 0;
        out[5] = h->ptListedPoints;
        out[6] = h->ptListedCells;
        out[7] = (int64_t)h->edgeSmem;
        return SMGPU_OK;
    }

    int smgpu_filter_stats(smgpu_handle *h, int64_t out[4])
    {
        if (!h || !out)
            return setErr(SMGPU_ERR_ARG, "null argument");
        try
        {
            CK(cudaSetDevice(h->prm.device));
            const size_t P = (size_t)h->topo.P;
            std::vector<uint8_t> buf(P);
            out[0] = h->d.fusedFaceFilter;
            out[1] = 0;
            if (h->d.suspect && h->d.fusedFaceFilter)
            {
                CK(cudaMemcpy(buf.data(), h->d.suspect, P, cudaMemcpyDeviceToHost));
                for (uint8_t b : buf)
                    out[1] += b != 0;
            }
            CK(cudaMemcpy(buf.data(), h->d.activeFlag, P, cudaMemcpyDeviceToHost));
            out[2] = 0;
            for (uint8_t b : buf)
                out[2] += b != 0;
            out[3] = (h->useTiles ? 1 : 0) | (h->tilesF ? 2 : 0) | (h->tilesUniform ? 4 : 0) | (h->d.faceFilter32 ? 8 : 0) |
                     (h->d.edgeFilter32 ? 16 : 0) | (h->usePointTiles ? 32 : 0) | (h->d.edgeTile32 ? 64 : 0);
        }
        catch (const std::exception &e)
        {
            return setErr(SMGPU_ERR_CUDA, e.what());
        }
        return SMGPU_OK;
    }

    int smgpu_selftest_division(int32_t device, uint64_t seed, int64_t n, int64_t *mismatches)
    {
        if (!mismatches || n < 1)
            return setErr(SMGPU_ERR_ARG, "bad argument");
        unsigned long long *dbad = nullptr;
        if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(&dbad, 8) != cudaSuccess)
            return setErr(SMGPU_ERR_CUDA, "no CUDA device");
        cudaMemset(dbad, 0, 8);
        const int blocks = 1184, threads = 256, per = (int)std::min<int64_t>(1 << 20, (n + (int64_t)blocks * threads - 1) / ((int64_t)blocks * threads));
        k_selftest_division<<<blocks, threads>>>(seed, per, dbad);
        unsigned long long bad = 0;
        const cudaError_t e = cudaMemcpy(&bad, dbad, 8, cudaMemcpyDeviceToHost);
        cudaFree(dbad);
        if (e != cudaSuccess)
            return setErr(SMGPU_ERR_CUDA, cudaGetErrorString(e));
        *mismatches = (int64_t)bad;
        return SMGPU_OK;
    }

    int smgpu_last_timing(smgpu_handle *h, double *ms, int64_t *launches)
    {
        if (!h)
            return setErr(SMGPU_ERR_ARG, "null handle");
        if (ms)
            *ms = h->lastMs;
        if (launches)
            *launches = h->lastLaunches;
        return SMGPU_OK;
    }

    // ---- operator-level entry points ----
    static int finishOp(smgpu_handle *h)
    {
        cudaError_t e = cudaStreamSynchronize(h->stream);
        if (e == cudaSuccess)
            e = cudaGetLastError();
        if (e != cudaSuccess)
            return setErr(SMGPU_ERR_CUDA, cudaGetErrorString(e));
        return SMGPU_OK;
    }

    int smgpu_op_cell_centres(smgpu_handle *h, double *out)
    {
        if (!h)
            return setErr(SMGPU_ERR_ARG, "null handle");
        cudaSetDevice(h->prm.device);
        h->resetControl();
        h->launchCellCentres();
        int rc = finishOp(h);
        if (rc == SMGPU_OK && out)
            rc = downloadP4(h, h->d.cellCtr, h->topo.C, out, h->cellOldOfNew);
        return rc;
    }

    int smgpu_op_predict(smgpu_handle *h, double *out)
    {
        if (!h)
            return setErr(SMGPU_ERR_ARG, "null handle");
        cudaSetDevice(h->prm.device);
        h->resetControl();
        h->launchPredict();
        int rc = finishOp(h);
        if (rc == SMGPU_OK && out)
            rc = downloadP4(h, h->d.newPts, h->topo.P, out, h->pointOldOfNew);
        return rc;
    }

    int smgpu_op_layer_normals(smgpu_handle *h, double *out)
    {
        if (!h)
            return setErr(SMGPU_ERR_ARG, "null handle");
        if (!h->doLayers)
            return setErr(SMGPU_ERR_ARG, "boundary layer treatment is not enabled for this handle");
        cudaSetDevice(h->prm.device);
        h->resetControl();
        // calculateBoundaryPointNormals accumulates onto the normals of the previous call (:178), so this probe
        // works on a copy: the handle's normals (and sharp flags) are what they were, and a later smgpu_iterate
        // still follows the reference's call sequence
        const size_t P = (size_t)h->topo.P;
        std::vector<uint8_t> sharpKeep;
        if (cudaMemcpyAsync(h->normalsTmp, h->d.normals, P * sizeof(P4), cudaMemcpyDeviceToDevice, h->stream) != cudaSuccess)
            return setErr(SMGPU_ERR_CUDA, "copy failed");
        if (h->d.sharp)
        {
            sharpKeep.resize(P);
            cudaStreamSynchronize(h->stream);
            cudaMemcpy(sharpKeep.data(), h->d.sharp, P, cudaMemcpyDeviceToHost);
        }
        h->launchFaceGeom();
        h->launchLayerNormals();
        int rc = finishOp(h);
        if (rc == SMGPU_OK && out)
            rc = downloadP4(h, h->d.normals, h->topo.P, out, h->pointOldOfNew);
        cudaMemcpy(h->d.normals, h->normalsTmp, P * sizeof(P4), cudaMemcpyDeviceToDevice);
        if (h->d.sharp)
            cudaMemcpy(h->d.sharp, sharpKeep.data(), P, cudaMemcpyHostToDevice);
        return rc;
    }

    int smgpu_op_layer_blend(smgpu_handle *h, double *out)
    {
        if (!h)
            return setErr(SMGPU_ERR_ARG, "null handle");
        if (!h->doLayers)
            return setErr(SMGPU_ERR_ARG, "boundary layer treatment is not enabled for this handle");
        cudaSetDevice(h->prm.device);
        h->resetControl();
        h->launchLayerBlend();
        int rc = finishOp(h);
        if (rc == SMGPU_OK && out)
            rc = downloadP4(h, h->d.newPts, h->topo.P, out, h->pointOldOfNew);
        return rc;
    }

    int smgpu_op_edge_constraints(smgpu_handle *h, uint8_t *out)
    {
        if (!h)
            return setErr(SMGPU_ERR_ARG, "null handle");
        cudaSetDevice(h->prm.device);
        h->resetControl();
        h->launchEdgeConstraints();
        int rc = finishOp(h);
        if (rc == SMGPU_OK && out)
            rc = smgpu_get_frozen(h, out);
        return rc;
    }

    int smgpu_op_face_angle_constraint(smgpu_handle *h, uint8_t *out)
    {
        if (!h)
            return setErr(SMGPU_ERR_ARG, "null handle");
        cudaSetDevice(h->prm.device);
        h->resetControl();
        h->launchFaceAngle();
        int rc = finishOp(h);
        if (rc == SMGPU_OK && out)
            rc = smgpu_get_frozen(h, out);
        return rc;
    }

    int smgpu_op_commit(smgpu_handle *h, int64_t *n_frozen, double *residual)
    {
        if (!h)
            return setErr(SMGPU_ERR_ARG, "null handle");
        cudaSetDevice(h->prm.device);
        h->resetControl();
        h->launchCommit();
        int rc = finishOp(h);
        if (rc != SMGPU_OK)
            return rc;
        double r;
        long long nf;
        if (cudaMemcpy(&r, h->d.statRes, 8, cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(&nf, h->d.statFrozen, 8, cudaMemcpyDeviceToHost) != cudaSuccess)
            return setErr(SMGPU_ERR_CUDA, "download failed");
        if (residual)
            *residual = r;
        if (n_frozen)
            *n_frozen = nf;
        return SMGPU_OK;
    }

    int smgpu_op_edge_face_angles(smgpu_handle *h, double *min_out, double *max_out)
    {
        if (!h || !min_out || !max_out)
            return setErr(SMGPU_ERR_ARG, "null argument");
        cudaSetDevice(h->prm.device);
        double *dmin = nullptr, *dmax = nullptr;
        const size_t E = (size_t)h->topo.E;
        if (cudaMalloc(&dmin, E * 8 + 8) != cudaSuccess || cudaMalloc(&dmax, E * 8 + 8) != cudaSuccess)
            return setErr(SMGPU_ERR_CUDA, "cudaMalloc failed");
        h->resetControl();
        h->launchCellCentres(); // mesh.C() is demand-driven in the reference (:1218)
        // k_face_current only mins/maxes into curMin/curMax/activeFlag, which k_predict resets at
        // the start of every iteration
        k_face_current<<<smgpu_handle::grid(h->d.E, 128), 128, 0, h->stream>>>(h->d, dmin, dmax);
        int rc = finishOp(h);
        if (rc == SMGPU_OK && (cudaMemcpy(min_out, dmin, E * 8, cudaMemcpyDeviceToHost) != cudaSuccess ||
                               cudaMemcpy(max_out, dmax, E * 8, cudaMemcpyDeviceToHost) != cudaSuccess))
            rc = setErr(SMGPU_ERR_CUDA, "download failed");
        cudaFree(dmin);
        cudaFree(dmax);
        return rc;
    }

    int smgpu_get_edges(smgpu_handle *h, int32_t *out)
    {
        if (!h || !out)
            return setErr(SMGPU_ERR_ARG, "null argument");
        memcpy(out, h->topo.edge.data(), h->topo.edge.size() * sizeof(int32_t));
        return SMGPU_OK;
    }

    int smgpu_get_csr(smgpu_handle *h, const char *name, int32_t *offsets, int32_t *values, int64_t *n_values)
    {
        if (!h || !name)
            return setErr(SMGPU_ERR_ARG, "null argument");
        const sm::Topology &t = h->topo;
        const std::string n = name;
        const sm::Vec<int32_t> *off = nullptr, *val = nullptr;
        if (n == "pointCells")
            off = &t.pcOff, val = &t.pc;
        else if (n == "pointPoints")
            off = &t.ppOff, val = &t.pp;
        else if (n == "pointEdges")
            off = &t.ppOff, val = &t.pe;
        else if (n == "edgeFaces")
            off = &t.efOff, val = &t.ef;
        else if (n == "edgeCells")
            off = &t.ecOff, val = &t.ecCell;
        else
            return setErr(SMGPU_ERR_ARG, "unknown table " + n);
        if (n_values)
            *n_values = (int64_t)val->size();
        if (offsets)
            memcpy(offsets, off->data(), off->size() * sizeof(int32_t));
        if (values)
            memcpy(values, val->data(), val->size() * sizeof(int32_t));
        return SMGPU_OK;
    }

    int smgpu_profile(smgpu_handle *h, int32_t enable)
    {
        if (!h)
            return setErr(SMGPU_ERR_ARG, "null handle");
        h->profiling = enable != 0;
        for (int k = 0; k < smgpu_handle::K_NUM; ++k)
        {
            h->profMs[k] = 0;
            h->profLaunches[k] = 0;
        }
        return SMGPU_OK;
    }

    int smgpu_profile_get(smgpu_handle *h, int32_t *n, const char **names, double *ms_total, int64_t *launches)
    {
        static const char *kNames[smgpu_handle::K_NUM] = {"k_face_geom", "k_cell_centres", "k_predict",    "k_edge_constraints", "k_face_current",
                                                          "k_active_compact", "k_face_tests", "k_face_resolve",     "k_commit", "halo_exchange", "k_layer", "k_geom_tiles",
                                                          "x_pack", "x_merge", "x_frozen", "x_finish"};
        if (!h || !n)
            return setErr(SMGPU_ERR_ARG, "null argument");
        *n = smgpu_handle::K_NUM;
        for (int k = 0; k < smgpu_handle::K_NUM; ++k)
        {
            if (names)
                names[k] = kNames[k];
            if (ms_total)
                ms_total[k] = h->profMs[k];
            if (launches)
                launches[k] = h->profLaunches[k];
        }
        return SMGPU_OK;
    }

    int smgpu_comm_unique_id(uint8_t id_out[128])
    {
        try
        {
            ncclUniqueId uid;
            const ncclResult_t r = sm::nccl().GetUniqueId(&uid);
            if (r != ncclSuccess)
                return setErr(SMGPU_ERR_COMM, sm::nccl().GetErrorString(r));
            memcpy(id_out, &uid, 128);
        }
        catch (const std::exception &e)
        {
            return setErr(SMGPU_ERR_COMM, e.what());
        }
        return SMGPU_OK;
    }

    int smgpu_comm_local_shared(smgpu_handle *h, int64_t *n, int64_t *gids_out)
    {
        if (!h || !n)
            return setErr(SMGPU_ERR_ARG, "null argument");
        if (h->gid.empty() && !h->topo.procPoints.empty())
            return setErr(SMGPU_ERR_ARG, "mesh has processor patches but was created without point_global_id");
        *n = (int64_t)h->topo.procPoints.size();
        if (gids_out)
            for (size_t i = 0; i < h->topo.procPoints.size(); ++i)
                gids_out[i] = h->gid[h->topo.procPoints[i]];
        return SMGPU_OK;
    }

    int smgpu_comm_prepare(smgpu_handle *h, int32_t rank, int32_t n_ranks, const int64_t *counts, const int64_t *all_gids)
    {
        if (!h || !counts || (!all_gids && n_ranks > 1) || rank < 0 || rank >= n_ranks)
            return setErr(SMGPU_ERR_ARG, "bad argument");
        if (h->comm)
            return setErr(SMGPU_ERR_ARG, "communicator already initialised");
        if (h->doBoundary)
            return setErr(SMGPU_ERR_ARG, "enable boundary point smoothing after the communicator exists (its set-up is collective)");
        try
        {
            h->comm = sm::commPrepare(h, rank, n_ranks, counts, all_gids);
        }
        catch (const std::exception &e)
        {
            return setErr(SMGPU_ERR_COMM, e.what());
        }
        return SMGPU_OK;
    }

    int smgpu_comm_init(smgpu_handle *h, int32_t rank, int32_t n_ranks, const uint8_t id[128], const int64_t *counts,
                        const int64_t *all_gids)
    {
        if (!h || !id)
            return setErr(SMGPU_ERR_ARG, "null argument");
        if (h->comm && (h->comm->nccl || h->comm->group))
            return setErr(SMGPU_ERR_ARG, "communicator already initialised");
        if (!h->comm)
        { // one-step use: a local failure here leaves the other ranks waiting in ncclCommInitRank; hosts that can
          // agree on the outcome first call smgpu_comm_prepare, compare notes, and only then come here
            const int rc = smgpu_comm_prepare(h, rank, n_ranks, counts, all_gids);
            if (rc != SMGPU_OK)
                return rc;
        }
        try
        {
            sm::commConnectNccl(h->comm, h, rank, n_ranks, id);
        }
        catch (const std::exception &e)
        {
            if (h->comm->nccl)
                sm::nccl().CommAbort(h->comm->nccl);
            h->comm->nccl = nullptr;
            return setErr(SMGPU_ERR_COMM, e.what());
        }
        return SMGPU_OK;
    }

    int smgpu_comm_p2p_export(smgpu_handle *h, uint8_t handle_out[64])
    {
        if (!h || !handle_out)
            return setErr(SMGPU_ERR_ARG, "null argument");
        if (!h->comm || !h->comm->xblock)
            return setErr(SMGPU_ERR_ARG, "call smgpu_comm_prepare / smgpu_comm_init first");
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
        cudaIpcMemHandle_t mh;
        if (cudaSetDevice(h->prm.device) != cudaSuccess || cudaIpcGetMemHandle(&mh, h->comm->xblock) != cudaSuccess)
        {
            cudaGetLastError();
            return setErr(SMGPU_ERR_COMM, "cudaIpcGetMemHandle failed for the exchange block");
        }
        memcpy(handle_out, &mh, 64);
        return SMGPU_OK;
    }

    int smgpu_comm_p2p_connect(smgpu_handle *h, const uint8_t *all_handles, smgpu_handle *const *local_peers)
    {
        if (!h || (!all_handles && !local_peers))
            return setErr(SMGPU_ERR_ARG, "null argument");
        if (!h->comm || !h->comm->xblock || h->comm->group)
            return setErr(SMGPU_ERR_ARG, "the peer-memory exchange needs a prepared communicator (not an in-process group)");
        if (getenv("SMGPU_NO_P2P") && atoi(getenv("SMGPU_NO_P2P")) != 0)
            return setErr(SMGPU_ERR_COMM, "peer-memory exchange disabled by SMGPU_NO_P2P");
        sm::Comm *cm = h->comm;
        const int nRanks = cm->plan.nRanks, rank = cm->plan.rank;
        try
        {
            CK(cudaSetDevice(h->prm.device));
            std::vector<unsigned char *> bases(nRanks, nullptr);
            for (int r = 0; r < nRanks; ++r)
            {
                if (r == rank)
                {
                    bases[r] = cm->xblock;
                    continue;
                }
                smgpu_handle *lp = local_peers ? local_peers[r] : nullptr;
                if (lp)
                { // a handle of this process: its block is addressable directly once peer access is on
                    if (!lp->comm || !lp->comm->xblock)
                        throw std::runtime_error("local peer without a prepared communicator");
                    if (lp->prm.device != h->prm.device)
                    {
                        int can = 0;
                        CK(cudaDeviceCanAccessPeer(&can, h->prm.device, lp->prm.device));
                        if (!can)
                            throw std::runtime_error("no peer access between the devices of two ranks");
                        const cudaError_t e = cudaDeviceEnablePeerAccess(lp->prm.device, 0);
                        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                            throw std::runtime_error(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
                        cudaGetLastError();
                    }
                    bases[r] = lp->comm->xblock;
                    continue;
                }
                if (!all_handles)
                    throw std::runtime_error("no handle for a peer of another process");
                cudaIpcMemHandle_t mh;
                memcpy(&mh, all_handles + 64 * (size_t)r, 64);
                void *p = nullptr;
                const cudaError_t e = cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess);
                if (e != cudaSuccess)
                {
                    cudaGetLastError();
                    throw std::runtime_error(std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
                }
                cm->ipcMapped.push_back(p);
                bases[r] = (unsigned char *)p;
            }
            sm::commConnectPeers(cm, h, bases);
        }
        catch (const std::exception &e)
        {
            return setErr(SMGPU_ERR_COMM, e.what());
        }
        return SMGPU_OK;
    }

    int smgpu_comm_p2p_disable(smgpu_handle *h)
    {
        if (!h || !h->comm)
            return setErr(SMGPU_ERR_ARG, "no communicator");
        h->comm->c.p2p = nullptr;
        h->comm->p2p = false;
        // unmap the peers' blocks now: a host that tears a run down calls this on every rank, synchronises, and
        // only then destroys the handles, so no block is freed while a peer still maps it
        cudaSetDevice(h->prm.device);
        cudaStreamSynchronize(h->stream);
        for (void *p : h->comm->ipcMapped)
            cudaIpcCloseMemHandle(p);
        h->comm->ipcMapped.clear();
        return SMGPU_OK;
    }

    int smgpu_comm_abort(smgpu_handle *h)
    {
        if (!h || !h->comm || !h->comm->nccl)
            return SMGPU_OK;
        try
        {
            sm::nccl().CommAbort(h->comm->nccl); // releases kernels of this rank that wait for a peer
            h->comm->nccl = nullptr;
        }
        catch (const std::exception &e)
        {
            return setErr(SMGPU_ERR_COMM, e.what());
        }
        return SMGPU_OK;
    }

    int smgpu_group_create(smgpu_handle **handles, int32_t n, smgpu_group **out)
    {
        if (!handles || n < 1 || !out)
            return setErr(SMGPU_ERR_ARG, "bad argument");
        *out = nullptr;
        for (int r = 0; r < n; ++r)
        {
            if (!handles[r] || handles[r]->comm)
                return setErr(SMGPU_ERR_ARG, "group members must be fresh handles without a communicator");
            if (handles[r]->prm.device != handles[0]->prm.device)
                return setErr(SMGPU_ERR_ARG, "the members of an in-process group share one device");
            if (handles[r]->doBoundary)
                return setErr(SMGPU_ERR_ARG, "boundary point smoothing is single-GPU in this build");
            if (handles[r]->gid.empty() && !handles[r]->topo.procPoints.empty())
                return setErr(SMGPU_ERR_ARG, "mesh has processor patches but was created without point_global_id");
        }
        smgpu_group *grp = new smgpu_group;
        sm::LocalGroup &g = grp->g;
        try
        {
            CK(cudaSetDevice(handles[0]->prm.device));
            g.members.assign(handles, handles + n);
            g.stream = handles[0]->stream;
            g.bar.n = n;
            g.posted.assign(n, nullptr);
            // the all-gather of the processor-point lists a multi-process host does itself (smgpu_comm_local_shared)
            std::vector<int64_t> counts(n), all;
            for (int r = 0; r < n; ++r)
            {
                counts[r] = (int64_t)handles[r]->topo.procPoints.size();
                for (int32_t p : handles[r]->topo.procPoints)
                    all.push_back(handles[r]->gid[p]);
            }
            if (all.empty())
                all.push_back(0);
            double mn = 1e300, mx = 0;
            for (int r = 0; r < n; ++r)
            {
                smgpu_handle *h = handles[r];
                h->comm = sm::commPrepare(h, r, n, counts.data(), all.data());
                h->comm->group = &g;
                sm::commAttach(h->comm, h);
                CK(cudaStreamSynchronize(h->stream));
                g.ownStreams.push_back(h->stream);
                h->stream = g.stream; // one stream orders the members' kernels and the copies between them
                mn = std::min(mn, h->topo.minEdgeLength);
                mx = std::max(mx, h->topo.maxEdgeLength);
            }
            std::vector<double *> res(n);
            std::vector<long long *> frz(n);
            for (int r = 0; r < n; ++r)
            {
                res[r] = handles[r]->comm->c.redRes;
                frz[r] = handles[r]->comm->c.redFrozen;
            }
            g.dRes = (double **)handles[0]->upload(res);
            g.dFrozen = (long long **)handles[0]->upload(frz);
            for (int r = 0; r < n; ++r)
            { // getMeshStats' returnReduce(min/max), src/smoothMesh.C:1527-1528
                handles[r]->meshMinEdge = mn;
                handles[r]->meshMaxEdge = mx;
                handles[r]->resolveParams();
            }
            // the collective host steps of the layer set-up: one short-lived thread per member
            std::vector<std::string> errs(n);
            std::vector<std::thread> th;
            bool anyLayers = false;
            for (int r = 0; r < n; ++r)
                anyLayers = anyLayers || (handles[r]->anyLayerPatch && handles[r]->layersParallel);
            if (anyLayers)
            {
                for (int r = 0; r < n; ++r)
                    if (!(handles[r]->anyLayerPatch && handles[r]->layersParallel))
                        throw std::runtime_error("boundary layer treatment must be selected on every member of a group "
                                                 "(the reference evaluates -layerPatches on every rank)");
                for (int r = 0; r < n; ++r)
                    th.emplace_back([&, r] {
                        try
                        {
                            sm::commSetupLayers(handles[r]->comm, handles[r]);
                        }
                        catch (const std::exception &e)
                        {
                            errs[r] = e.what();
                        }
                    });
                for (auto &t : th)
                    t.join();
                for (int r = 0; r < n; ++r)
                    if (!errs[r].empty())
                        throw std::runtime_error(errs[r]);
            }
        }
        catch (const std::exception &e)
        {
            smgpu_group_destroy(grp);
            return setErr(SMGPU_ERR_COMM, e.what());
        }
        *out = grp;
        return SMGPU_OK;
    }

    int smgpu_group_enable_boundary_smoothing(smgpu_group *grp, const smgpu_boundary_geometry *geometry,
                                              const int32_t *const *patch_smoothing, double internal_smoothing_blending_fraction)
    {
        if (!grp || !geometry || !patch_smoothing)
            return setErr(SMGPU_ERR_ARG, "null argument");
        sm::LocalGroup &g = grp->g;
        const int n = (int)g.members.size();
        // the set-up is collective (global mesh figures, hop counts, point normals): one short-lived host thread per member
        std::vector<int> rc(n, SMGPU_OK);
        std::vector<std::string> msg(n);
        std::vector<std::thread> th;
        for (int r = 0; r < n; ++r)
            th.emplace_back([&, r] {
                rc[r] = smgpu_enable_boundary_smoothing(g.members[r], geometry, patch_smoothing[r], internal_smoothing_blending_fraction);
                if (rc[r] != SMGPU_OK)
                    msg[r] = smgpu_last_error();
            });
        for (auto &t : th)
            t.join();
        for (int r = 0; r < n; ++r)
            if (rc[r] != SMGPU_OK)
                return setErr(rc[r], msg[r]);
        return SMGPU_OK;
    }

    int smgpu_group_destroy(smgpu_group *grp)
    {
        if (!grp)
            return SMGPU_OK;
        sm::LocalGroup &g = grp->g;
        if (g.stream)
            cudaStreamSynchronize(g.stream);
        for (size_t r = 0; r < g.members.size(); ++r)
        {
            smgpu_handle *h = g.members[r];
            if (r < g.ownStreams.size())
                h->stream = g.ownStreams[r];
            if (h->comm)
            {
                sm::commDestroy(h->comm);
                h->comm = nullptr;
                h->d.multiRank = 0;
            }
        }
        delete grp;
        return SMGPU_OK;
    }

    int smgpu_group_iterate(smgpu_group *grp, int32_t max_iters, int64_t *n_frozen, double *residual, int32_t *iters_done)
    {
        if (!grp || max_iters < 0)
            return setErr(SMGPU_ERR_ARG, "bad argument");
        sm::LocalGroup &g = grp->g;
        smgpu_handle *h0 = g.members[0];
        for (smgpu_handle *h : g.members)
            if (h->doLayers && h->layersParallel && !h->layersReady)
                return setErr(SMGPU_ERR_ARG, "boundary layer treatment: the group's layer set-up has not run");
        try
        {
            CK(cudaSetDevice(h0->prm.device));
            for (smgpu_handle *h : g.members)
            {
                h->ensureStats(max_iters);
                h->launches = 0;
                h->resetControl();
            }
            CK(cudaEventRecord(h0->ev0, g.stream));
            int done = 0, launched = 0;
            const int chunk = 16;
            while (launched < max_iters && !done)
            {
                const int n = std::min(chunk, max_iters - launched);
                for (int i = 0; i < n; ++i)
                    sm::groupIterate(&g);
                launched += n;
                if (launched < max_iters)
                {
                    CK(cudaMemcpyAsync(&done, h0->d.done, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
                    CK(cudaStreamSynchronize(g.stream));
                }
            }
            CK(cudaEventRecord(h0->ev1, g.stream));
            CK(cudaStreamSynchronize(g.stream));
            CK(cudaGetLastError());
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, h0->ev0, h0->ev1));
            int64_t launches = 1;
            for (smgpu_handle *h : g.members)
            {
                launches += h->launches;
                if (h->profiling)
                    h->profCollect();
            }
            h0->lastMs = ms;
            h0->lastLaunches = launches;
            for (smgpu_handle *h : g.members)
            {
                const int rc = h->checkErrFlag();
                if (rc != SMGPU_OK)
                    return rc;
            }
            return h0->fetchStats(n_frozen, residual, iters_done);
        }
        catch (const std::exception &e)
        {
            return setErr(SMGPU_ERR_CUDA, e.what());
        }
    }

} // extern "C"
