// exchange.hpp -- host-side plan for the interface-point exchanges that replace
// syncTools::syncPointList (src/smoothMesh.C:134,142,402,429,455,472,2374) in a
// cell-decomposed multi-GPU run.  Pure CPU code: which points are shared with
// which rank, in which order, and where every copy of a shared point sits in
// the receive buffers.
//
// Semantics: every copy of a shared point is
// combined in ascending rank order and the result written back to all copies,
// so all ranks end with identical values.
#pragma once
#include <cstdint>
#include <vector>

namespace sm
{

struct ExchangePlan
{
    int rank = 0, nRanks = 1;
    // neighbour ranks (ascending) and, per neighbour, the local labels of the points shared
    // with it, ordered by global point label (both sides build the same order)
    std::vector<int> nbrRank;
    std::vector<int32_t> nbrOff;   // nbrRank.size()+1 offsets into sendPoint
    std::vector<int32_t> sendPoint; // local point label per send/recv slot
    // unique local shared points and their copies on other ranks
    std::vector<int32_t> sharedPoint; // ascending local label
    std::vector<int32_t> selfSlot;    // one send slot that holds this point's own tuple
    std::vector<int32_t> copyOff;     // sharedPoint.size()+1
    std::vector<int32_t> copyRank;    // rank of the copy (ascending within a point)
    std::vector<int32_t> copySlot;    // its slot in the receive buffer (same indexing as sendPoint)
    int maxCopies = 0;                // copies per point including the local one
};

// myGids: global labels of this rank's processor-patch points (local labels in myLocal, any
// order); counts/allGids: the same lists of every rank, concatenated in rank order.
ExchangePlan buildExchangePlan(int rank, int nRanks, const std::vector<int32_t> &myLocal,
                               const std::vector<int64_t> &myGids, const std::vector<int64_t> &counts,
                               const std::vector<int64_t> &allGids);

} // namespace sm
