// comm_impl.cuh -- multi-GPU exchange layer, included by smgpu.cu after smgpu_handle.
//
// One process per GPU; the mesh is partitioned by cells (decomposePar-style
// processor meshes), interface points are duplicated.  This file replaces the
// reference's MPI traffic on the hot path:
//   syncPointList(plusEqOp)        src/smoothMesh.C:134,142   centroidal sums / counts
//   syncPointList(minMagSqrEqOp)   :402,429,455               closest-point merge
//   syncPointList(orEqOp)          :472, :2374                hasCommonCell, isFrozenPoint
//   returnReduce(max / sum)        :1567, :2396               residual, nFrozenPoints
// by two grouped ncclSend/ncclRecv halo exchanges and one pair of all-reduces per
// iteration.  Every rank receives the other copies' *local* predictor tuple once
// and replays the reference's three-stage merge for all copies itself, so the
// four point syncs of the predictor collapse into one exchange.  Copies are
// combined in ascending rank order (the CPU oracle's rank emulation does the same).
#pragma once
#include "exchange.hpp"
#include "nccl_dyn.hpp"

namespace smk
{

#define SMK_TUPLE 16       /* doubles per interface-point record */
#define SMK_TUPLE_LAYERS 24 /* with boundary layer treatment: + accumulated normal, face count, outer neighbour */
#define SMK_TUPLE_BOUNDARY 32 /* with boundary point smoothing: + feature edge projections (sum, count), inner neighbour */
#define SMK_MAXCOPIES 16

#define SMK_MAXNBR 32   /* neighbour ranks of one rank the peer-memory exchange supports */
#define SMK_MAXRANKS 64 /* ranks of a run it supports */

// Peer-memory exchange (smgpu_comm_p2p_*): every rank owns one exchange block -- receive buffers, one flag
// word per neighbour and exchange, one statistics slot per rank and parity -- that its peers map (CUDA IPC
// between processes, peer access inside one process) and WRITE into directly from the kernels that produce
// the data; a flag written with release semantics after the data tells the consumer kernel, which waits for
// it with acquire loads, that the iteration's records have landed.  No collective call, no extra launch.
struct P2PStat
{
    double res;
    long long nf;
    unsigned long long epoch, pad;
};
struct P2PDev
{
    int nNbr, nRanks, rank, pad;
    const unsigned char *slotNbr; // per send slot: index of its neighbour
    int nbrOff[SMK_MAXNBR + 1];
    double *peerRecv[SMK_MAXNBR];     // neighbour j's receive buffer (in j's block) ...
    int peerSlot0[SMK_MAXNBR];        // ... and the slot of it where this rank's first record goes
    uint8_t *peerRecvFz[SMK_MAXNBR];  // (already offset to that slot)
    unsigned long long *peerFlagT[SMK_MAXNBR], *peerFlagF[SMK_MAXNBR]; // j's flag words for this rank
    unsigned long long *flagT, *flagF; // this rank's flag words, one per neighbour, written by the neighbours
    P2PStat *peerStat[SMK_MAXRANKS];   // every rank's slot array [2][nRanks]
    P2PStat *stat;                     // this rank's
    unsigned long long *epoch;         // iteration counter of the exchange (never reset)
    unsigned int *packDone, *fzDone;   // last-block counters of the two producer kernels
};
__device__ __forceinline__ void stReleaseSys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ldAcquireSys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// waits until *p >= want (want == exact for the parity-buffered statistics); a peer that never arrives raises
// errFlag 6 after about thirty seconds instead of hanging the device
__device__ __forceinline__ void spinUntil(const unsigned long long *p, unsigned long long want, int *errFlag)
{
    const long long t0 = clock64();
    while (ldAcquireSys(p) < want)
    {
        __nanosleep(64);
        if (clock64() - t0 > 60000000000ll)
        {
            *errFlag = 6;
            break;
        }
    }
}

struct CommDev
{
    int nSlots, nShared, rank;
    const P2PDev *p2p; // non-null: the exchanges go through peer memory
    int tuple; // doubles per record: SMK_TUPLE, or SMK_TUPLE_LAYERS with boundary layer treatment
    const int *sendPoint, *sharedPoint, *selfSlot, *copyOff, *copyRank, *copySlot;
    double *sendBuf, *recvBuf;
    uint8_t *sendFz, *recvFz;
    double *redRes;
    long long *redFrozen;
};

// position of boundary point p in the ascending list of boundary points (the per-boundary-point tables are indexed by it)
__device__ __forceinline__ int boundaryIndex(const Dev &d, int p)
{
    int lo = 0, hi = d.nBPoints - 1;
    while (lo < hi)
    {
        const int mid = (lo + hi) >> 1;
        if (d.bPoints[mid] < p)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}
// local predictor tuple of every send slot's point
__global__ void __launch_bounds__(128) k_shared_pack(Dev d, CommDev c)
{
    if (*d.done)
        return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < c.nSlots)
    {
        const int p = c.sendPoint[i];
        const P4 self = ld4(d.pts + p);
        const D3 x = {self.x, self.y, self.z};
        const bool internal = self.w != 0.0;
        PointLocal L;
        pointLocal(d, p, x, internal, L);
        // topology.cpp refuses meshes with a point that has fewer than two eligible neighbours (the reference's
        // "Failed to find cLabel" abort, src/smoothMesh.C:354-362), so n1/n2 are labels here; the guard keeps
        // the walk inside the table even if that invariant were ever broken
        const bool hc = (L.n1 >= 0 && L.n2 >= 0) && shareCell(d, L.n1, L.n2);
        double *r = c.sendBuf + (size_t)i * c.tuple;
        r[0] = L.sum.x, r[1] = L.sum.y, r[2] = L.sum.z;
        r[3] = (double)L.nCells;
        r[4] = L.r1.x, r[5] = L.r1.y, r[6] = L.r1.z;
        r[7] = L.r2.x, r[8] = L.r2.y, r[9] = L.r2.z;
        r[10] = L.r3.x, r[11] = L.r3.y, r[12] = L.r3.z;
        r[13] = hc ? 1.0 : 0.0;
        r[14] = r[15] = 0.0;
        if (d.normalsOn)
        {
            // calculateBoundaryPointNormals up to its synchronisation (orthogonalBoundaryBlending.C:151-182):
            // this copy's previous normal minus the unit normals of its boundary faces, and the face count
            D3 n = ld3(d.normals, p);
            const int b = d.bfOff[p], e = d.bfOff[p + 1];
            for (int k = b; k < e; ++k)
            {
                const D3 Sf = ld3(d.faceGeo, 2 * d.bf[k] + 1);
                n = n - Sf / mag(Sf);
            }
            r[16] = n.x, r[17] = n.y, r[18] = n.z;
            r[19] = (double)(e - b);
            // updateNeighCoords before its synchronisation (:472-487)
            D3 oc = {SM_GREAT, SM_GREAT, SM_GREAT};
            if (d.layers)
            {
                const int o = d.pointToOuter[p];
                if (o >= 0)
                    oc = ld3(d.pts, o);
            }
            r[20] = oc.x, r[21] = oc.y, r[22] = oc.z;
            r[23] = 0.0;
        }
        if (d.bsmooth)
        {
            // boundary point smoothing: this copy's share of calculateFeatureEdgeProjections
            // (src/boundaryPointSmoothing.C:623-656) and its inner neighbour's coordinates
            // (updateNeighCoords for the inner map, orthogonalBoundaryBlending.C:472-487)
            D3 sum = {0, 0, 0};
            int nProj = 0;
            const int cls = d.bClass[p];
            if (!internal && (cls & 2))
                featureEdgeProjectionLocal(d, p, boundaryIndex(d, p), sum, nProj);
            r[24] = sum.x, r[25] = sum.y, r[26] = sum.z;
            r[27] = (double)nProj;
            D3 ic = {SM_GREAT, SM_GREAT, SM_GREAT};
            const int inner = d.bInner[p];
            if (inner >= 0)
                ic = ld3(d.pts, inner);
            r[28] = ic.x, r[29] = ic.y, r[30] = ic.z;
            r[31] = 0.0;
        }
    }
    if (c.p2p)
    {
        // The block's records go straight into the neighbours' receive buffers over NVLink: the block copies its
        // contiguous piece of the send buffer 16 bytes per lane, consecutive lanes to consecutive addresses, so a
        // warp instruction is one 512-byte write (a neighbour boundary inside the piece only splits one of them).
        const P2PDev *x2 = c.p2p;
        __syncthreads(); // the records the threads of this block have just written
        const int s0 = blockIdx.x * blockDim.x, s1 = min(s0 + (int)blockDim.x, c.nSlots);
        const int perSlot = c.tuple / 2, pieces = (s1 - s0) * perSlot;
        for (int q = threadIdx.x; q < pieces; q += blockDim.x)
        {
            const int slot = s0 + q / perSlot, k = 2 * (q % perSlot);
            const int j = x2->slotNbr[slot];
            const double2 v = *reinterpret_cast<const double2 *>(c.sendBuf + (size_t)slot * c.tuple + k);
            *reinterpret_cast<double2 *>(x2->peerRecv[j] + (size_t)(x2->peerSlot0[j] + slot - x2->nbrOff[j]) * c.tuple + k) = v;
        }
        // the last block to finish tells every neighbour that this iteration's records are complete
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0)
        {
            const unsigned int t = atomicAdd(x2->packDone, 1u);
            if (t == gridDim.x - 1)
            {
                *x2->packDone = 0;
                __threadfence_system();
                const unsigned long long e = *x2->epoch;
                for (int j = 0; j < x2->nNbr; ++j)
                    stReleaseSys(x2->peerFlagT[j], e);
            }
        }
    }
}

// isSmallerByVectorElements / isCloserPoint, src/smoothMesh.C:222-272
__device__ __forceinline__ bool isCloserPoint(D3 a, D3 b)
{
    if (veq(a, b))
        return false;
    const double delta = mag(a) - mag(b);
    if (delta < SM_VSMALL)
        return true;
    if (fabs(delta) < SM_VSMALL)
    {
        if (a.x < b.x)
            return true;
        if (a.x > b.x)
            return false;
        if (a.y < b.y)
            return true;
        if (a.y > b.y)
            return false;
        return a.z < b.z;
    }
    return false;
}
__device__ __forceinline__ D3 minMagSqr(D3 x, D3 y) { return (magSqr(x) <= magSqr(y)) ? x : y; }

// Combine all copies of every interface point (ascending rank), replay the merge of
// findClosestPoints (:391-478) for every copy, then finish the predictor for the local copy.
__global__ void __launch_bounds__(64) k_shared_merge(Dev d, CommDev c)
{
    if (*d.done)
        return;
    if (c.p2p)
    { // the neighbours' records of this iteration have landed in this rank's receive buffer
        if (threadIdx.x < c.p2p->nNbr)
            spinUntil(c.p2p->flagT + threadIdx.x, *c.p2p->epoch, d.errFlag);
        __syncthreads();
    }
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= c.nShared)
        return;
    const int p = c.sharedPoint[s];
    const int cb = c.copyOff[s], nOther = c.copyOff[s + 1] - cb;
    const int n = nOther + 1;
    D3 cp1[SMK_MAXCOPIES], cp2[SMK_MAXCOPIES], cp3[SMK_MAXCOPIES];
    bool hc[SMK_MAXCOPIES];
    D3 sum = {0, 0, 0}, nrm = {0, 0, 0}, outer = {0, 0, 0}, fep = {0, 0, 0}, innerCoord = {0, 0, 0};
    double cnt = 0.0, nBoundaryFaces = 0.0, nFep = 0.0;
    int me = -1;
    // copies in ascending rank order, the local one inserted at its rank
    for (int k = 0, o = 0; k < n; ++k)
    {
        const double *r;
        if (me < 0 && (o >= nOther || c.copyRank[cb + o] > c.rank))
        {
            r = c.sendBuf + (size_t)c.selfSlot[s] * c.tuple;
            me = k;
        }
        else
        {
            r = c.recvBuf + (size_t)c.copySlot[cb + o] * c.tuple;
            ++o;
        }
        if (d.normalsOn)
        {
            const D3 nk = {r[16], r[17], r[18]}, ok = {r[20], r[21], r[22]};
            if (k == 0)
            {
                nrm = nk;
                nBoundaryFaces = r[19];
                outer = ok;
            }
            else
            {
                nrm = nrm + nk;                           // plusEqOp<vector>, orthogonalBoundaryBlending.C:185
                nBoundaryFaces = nBoundaryFaces + r[19];  // plusEqOp<label>, :193
                outer = minMagSqr(outer, ok);             // minMagSqrEqOp<vector>, :491
            }
        }
        if (d.bsmooth)
        {
            const D3 fk = {r[24], r[25], r[26]}, ik = {r[28], r[29], r[30]};
            if (k == 0)
            {
                fep = fk;
                nFep = r[27];
                innerCoord = ik;
            }
            else
            {
                fep = fep + fk;                          // plusEqOp<vector>, src/boundaryPointSmoothing.C:660
                nFep = nFep + r[27];                     // plusEqOp<label>, :668
                innerCoord = minMagSqr(innerCoord, ik);  // minMagSqrEqOp<vector>, orthogonalBoundaryBlending.C:491
            }
        }
        const D3 part = {r[0], r[1], r[2]};
        if (k == 0)
        {
            sum = part;
            cnt = r[3];
        }
        else
        {
            sum = sum + part; // plusEqOp<vector>, :134
            cnt = cnt + r[3]; // plusEqOp<label>, :142 (exact in double)
        }
        cp1[k] = {r[4], r[5], r[6]};
        cp2[k] = {r[7], r[8], r[9]};
        cp3[k] = {r[10], r[11], r[12]};
        hc[k] = r[13] != 0.0;
    }
    // position 1 (:395-419)
    D3 v = cp1[0];
    for (int k = 1; k < n; ++k)
        v = minMagSqr(v, cp1[k]);
    for (int k = 0; k < n; ++k)
        if (isCloserPoint(v, cp1[k]))
        {
            cp3[k] = cp2[k];
            cp2[k] = cp1[k];
            cp1[k] = v;
            hc[k] = false;
        }
    // position 2 (:424-445)
    v = cp2[0];
    for (int k = 1; k < n; ++k)
        v = minMagSqr(v, cp2[k]);
    for (int k = 0; k < n; ++k)
        if (isCloserPoint(v, cp2[k]))
        {
            cp3[k] = cp2[k];
            cp2[k] = v;
            hc[k] = false;
        }
    // position 3 (:450-469)
    v = cp3[0];
    for (int k = 1; k < n; ++k)
        v = minMagSqr(v, cp3[k]);
    for (int k = 0; k < n; ++k)
        if (isCloserPoint(v, cp3[k]))
            cp3[k] = v;
    bool anyCommon = false; // orEqOp<bool>, :472
    for (int k = 0; k < n; ++k)
        anyCommon = anyCommon || hc[k];

    const P4 self = ld4(d.pts + p);
    const D3 x = {self.x, self.y, self.z};
    const bool internal = self.w != 0.0;
    const D3 cen = (cnt != 0.0) ? sum / cnt : x; // :158-162
    double blend = 0.0;
    if (!anyCommon)
        blend = blendFraction(cp1[me], cp2[me], mag(cp1[me]), mag(cp2[me]), mag(cp3[me]), internal);
    D3 np = blendAndClamp(d, x, cen, cp1[me], cp2[me], blend);
    bool sharpNow = false;
    if (d.normalsOn)
    {
        // rest of calculateBoundaryPointNormals for this point (:200-230)
        const D3 zero = {0, 0, 0};
        if (nBoundaryFaces >= 1.0 && mag(nrm) < 0.1)
        {
            nrm = zero;
            sharpNow = true;
        }
        if (!veq(nrm, zero))
            nrm = nrm / mag(nrm);
        st4(d.normals + p, nrm, 0.0);
        if (d.sharp && nBoundaryFaces >= 1.0)
            d.sharp[p] = sharpNow ? 1 : 0;
    }
    if (d.layers)
    {
        // blendWithOrthogonalPoints + the second constrainMaxStepLength (src/smoothMesh.C:2288-2304)
        const D3 zero = {0, 0, 0};
        const int nHops = d.hops[p];
        if (!veq(nrm, zero) && internal && nHops >= 1)
        {
            const D3 undef = {SM_GREAT, SM_GREAT, SM_GREAT};
            if (veq(outer, undef))
                *d.errFlag = 1; // "Sanity broken, outerNeighCoord ..." (:537-540)
            const double length = d.layerLength[nHops], blendFrac = d.layerBlend[nHops];
            const D3 ortho = outer + length * nrm;
            np = blendFrac * ortho + (1.0 - blendFrac) * np;
        }
        const D3 stepDir = np - x;
        const double len = mag(stepDir);
        double scale = 1.0;
        if (len > d.maxStepLength)
            scale = d.maxStepLength / (len * d.relStepFrac);
        np = x + (d.relStepFrac * scale) * stepDir;
    }
    if (d.bsmooth)
    {
        // :2307-2355 for this point with the synchronised sums: projection onto corner / feature edge / target
        // surface, prismatic projection (only where THIS copy has an inner neighbour, like the reference, but
        // with the synchronised coordinates), third constrainMaxStepLength
        const int cls = d.bClass[p];
        if (!internal)
        {
            const bool frz = boundaryProjectPoint(d, boundaryIndex(d, p), cls, sharpNow, fep, (int)nFep, nrm, np);
            d.frozen[p] = frz ? 1 : 0; // k_boundary_project decided this from the previous iteration's sharp flag
            if ((cls & 4) && (cls & 8) && !(cls & 3) && !sharpNow && d.bInner[p] >= 0)
            {
                const D3 undef = {SM_GREAT, SM_GREAT, SM_GREAT};
                if (veq(innerCoord, undef))
                    *d.errFlag = 5;
                np = prismaticProjectPoint(d, np, nrm, innerCoord);
            }
        }
        const D3 stepDir = np - x;
        const double len = mag(stepDir);
        double scale = 1.0;
        if (len > d.maxStepLength)
            scale = d.maxStepLength / (len * d.relStepFrac);
        np = x + (d.relStepFrac * scale) * stepDir;
    }
    st4(d.newPts + p, np, 0.0);
}

__global__ void __launch_bounds__(128) k_frozen_pack(Dev d, CommDev c)
{
    if (*d.done)
        return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < c.nSlots)
    {
        const uint8_t f = d.frozen[c.sendPoint[i]];
        c.sendFz[i] = f;
        if (c.p2p)
        {
            const int j = c.p2p->slotNbr[i];
            c.p2p->peerRecvFz[j][i - c.p2p->nbrOff[j]] = f;
        }
    }
    if (c.p2p)
    {
        const P2PDev *x2 = c.p2p;
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0)
        {
            const unsigned int t = atomicAdd(x2->fzDone, 1u);
            if (t == gridDim.x - 1)
            {
                *x2->fzDone = 0;
                __threadfence_system();
                const unsigned long long e = *x2->epoch;
                for (int j = 0; j < x2->nNbr; ++j)
                    stReleaseSys(x2->peerFlagF[j], e);
            }
        }
    }
}
// orEqOp<bool> on isFrozenPoint, :2374
__global__ void __launch_bounds__(128) k_frozen_or(Dev d, CommDev c)
{
    if (*d.done)
        return;
    if (c.p2p)
    {
        if (threadIdx.x < c.p2p->nNbr)
            spinUntil(c.p2p->flagF + threadIdx.x, *c.p2p->epoch, d.errFlag);
        __syncthreads();
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < c.nSlots && c.recvFz[i])
        d.frozen[c.sendPoint[i]] = 1;
}
// Peer-memory mode, the points shared between ranks: wait for the neighbours' freeze flags, OR them in (orEqOp<bool>
// on isFrozenPoint, :2374), then restore / move / residual like k_commit, which has already handled every other
// point while the flags were on their way; the last block publishes this rank's statistics.
__global__ void __launch_bounds__(256) k_commit_shared(Dev d, CommDev c)
{
    if (*d.done)
        return;
    if (threadIdx.x < c.p2p->nNbr)
        spinUntil(c.p2p->flagF + threadIdx.x, *c.p2p->epoch, d.errFlag);
    __syncthreads();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    double dist = 0.0;
    unsigned int nf = 0;
    if (s < c.nShared)
    {
        const int p = c.sharedPoint[s];
        bool frz = false;
        for (int k = c.copyOff[s]; k < c.copyOff[s + 1]; ++k)
            frz = frz || c.recvFz[c.copySlot[k]] != 0;
        if (frz)
            d.frozen[p] = 1;
        commitPoint(d, p, dist, nf);
    }
    commitReduce(d, dist, nf, true);
}
// publishes the all-reduced statistics of the iteration and the stop flag (:2396-2405).  Peer-memory mode: this
// rank's (residual, nFrozen) goes into every rank's slot array first (one lane per rank), then the kernel waits
// for all slots of this iteration and reduces them in rank order (max / exact integer sum), which replaces the
// two returnReduce calls (:1567, :2396); the slots are double-buffered by iteration parity because a rank that is
// not a neighbour may already be one iteration ahead.
__global__ void __launch_bounds__(SMK_MAXRANKS) k_finish_iter(Dev d, CommDev c)
{
    if (*d.done)
        return;
    if (c.p2p)
    {
        const P2PDev *x2 = c.p2p;
        const unsigned long long e = *x2->epoch;
        const int par = (int)(e & 1), r = threadIdx.x;
        if (r < x2->nRanks)
        {
            P2PStat *dst = x2->peerStat[r] + par * x2->nRanks + x2->rank;
            dst->res = *c.redRes;
            dst->nf = *c.redFrozen;
            __threadfence_system();
            stReleaseSys(&dst->epoch, e);
            // exact match: the slot of this parity still holds iteration e - 2 until the peer has written
            const long long t0 = clock64();
            while (ldAcquireSys(&x2->stat[par * x2->nRanks + r].epoch) != e)
            {
                __nanosleep(64);
                if (clock64() - t0 > 60000000000ll)
                {
                    *d.errFlag = 6;
                    break;
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0)
        {
            double m = x2->stat[par * x2->nRanks].res;
            long long sum = x2->stat[par * x2->nRanks].nf;
            for (int q = 1; q < x2->nRanks; ++q)
            {
                const double v = x2->stat[par * x2->nRanks + q].res;
                m = (v > m) ? v : m;
                sum += x2->stat[par * x2->nRanks + q].nf;
            }
            *c.redRes = m;
            *c.redFrozen = sum;
            *x2->epoch = e + 1;
        }
        __syncthreads();
    }
    if (threadIdx.x != 0 || blockIdx.x != 0)
        return;
    const double res = *c.redRes;
    const int it = *d.iter;
    if (it < d.statCap)
    {
        d.statRes[it] = res;
        d.statFrozen[it] = *c.redFrozen;
    }
    *d.iter = it + 1;
    if (res < d.relTol)
        *d.done = 1;
}

} // namespace smk

namespace sm
{

struct LocalGroup;
struct Comm
{
    ncclComm_t nccl = nullptr;
    ExchangePlan plan;
    smk::CommDev c;
    // in-process group (smgpu_group_*): the ranks are handles of this process driven by one host thread on one
    // stream; exchanges are stream-ordered device copies between the members' buffers instead of NCCL calls
    LocalGroup *group = nullptr;
    // peer-memory exchange: this rank's exchange block (what the peers map) and the mapped blocks of the peers
    unsigned char *xblock = nullptr;
    size_t xblockBytes = 0;
    std::vector<void *> ipcMapped;
    bool p2p = false;
    const uint8_t *sharedFlag = nullptr;   // per point: shared between ranks (committed by k_commit_shared)
    cudaEvent_t evCommitted = nullptr, evFinished = nullptr;
    bool finishPending = false;            // k_finish_iter of the last iteration is in flight on xStream
    // the predictor exchange runs on its own stream, fenced by these events, so that kernels that do not
    // need its result keep the GPU busy meanwhile
    cudaStream_t xStream = nullptr;
    cudaEvent_t evPacked = nullptr, evExchanged = nullptr;
};

#define NCK(call)                                                                                                      \
    do                                                                                                                 \
    {                                                                                                                  \
        ncclResult_t r_ = (call);                                                                                      \
        if (r_ != ncclSuccess)                                                                                         \
            throw std::runtime_error(std::string("NCCL error: ") + sm::nccl().GetErrorString(r_) + " at " #call);             \
    } while (0)

static void haloExchange(Comm *cm, cudaStream_t stream, const void *send, void *recv, size_t elemBytes)
{
    const ExchangePlan &pl = cm->plan;
    if (cm->group)
        throw std::runtime_error("internal: NCCL exchange requested for a member of an in-process group");
    NCK(nccl().GroupStart());
    for (size_t j = 0; j < pl.nbrRank.size(); ++j)
    {
        const size_t off = (size_t)pl.nbrOff[j] * elemBytes, cnt = (size_t)(pl.nbrOff[j + 1] - pl.nbrOff[j]) * elemBytes;
        NCK(nccl().Send((const char *)send + off, cnt, ncclChar, pl.nbrRank[j], cm->nccl, stream));
        NCK(nccl().Recv((char *)recv + off, cnt, ncclChar, pl.nbrRank[j], cm->nccl, stream));
    }
    NCK(nccl().GroupEnd());
}

} // namespace sm
