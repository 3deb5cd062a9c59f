// exchange.cpp -- see exchange.hpp.
#include "exchange.hpp"

#include <algorithm>
#include <stdexcept>
#include <unordered_map>

namespace sm
{

ExchangePlan buildExchangePlan(int rank, int nRanks, const std::vector<int32_t> &myLocal,
                               const std::vector<int64_t> &myGids, const std::vector<int64_t> &counts,
                               const std::vector<int64_t> &allGids)
{
    if ((int)counts.size() != nRanks || myLocal.size() != myGids.size())
        throw std::runtime_error("buildExchangePlan: inconsistent arguments");
    ExchangePlan pl;
    pl.rank = rank;
    pl.nRanks = nRanks;
    std::unordered_map<int64_t, int32_t> mine;
    mine.reserve(myGids.size() * 2);
    for (size_t i = 0; i < myGids.size(); ++i)
        mine[myGids[i]] = myLocal[i];

    struct Copy
    {
        int rank, slot;
    };
    std::unordered_map<int32_t, std::vector<Copy>> copies; // local point -> copies elsewhere
    std::unordered_map<int32_t, int32_t> selfSlot;
    pl.nbrOff.push_back(0);
    int64_t off = 0;
    for (int r = 0; r < nRanks; ++r)
    {
        const int64_t n = counts[r];
        if (r != rank)
        {
            std::vector<std::pair<int64_t, int32_t>> common; // (gid, local label)
            for (int64_t i = 0; i < n; ++i)
            {
                auto it = mine.find(allGids[off + i]);
                if (it != mine.end())
                    common.push_back({allGids[off + i], it->second});
            }
            if (!common.empty())
            {
                std::sort(common.begin(), common.end());
                common.erase(std::unique(common.begin(), common.end()), common.end());
                pl.nbrRank.push_back(r);
                for (auto &c : common)
                {
                    const int32_t slot = (int32_t)pl.sendPoint.size();
                    pl.sendPoint.push_back(c.second);
                    copies[c.second].push_back({r, slot});
                    selfSlot.emplace(c.second, slot);
                }
                pl.nbrOff.push_back((int32_t)pl.sendPoint.size());
            }
        }
        off += n;
    }
    std::vector<int32_t> pts;
    pts.reserve(copies.size());
    for (auto &kv : copies)
        pts.push_back(kv.first);
    std::sort(pts.begin(), pts.end());
    pl.copyOff.push_back(0);
    for (int32_t p : pts)
    {
        auto &cv = copies[p];
        std::sort(cv.begin(), cv.end(), [](const Copy &a, const Copy &b) { return a.rank < b.rank; });
        pl.sharedPoint.push_back(p);
        pl.selfSlot.push_back(selfSlot[p]);
        for (auto &c : cv)
        {
            pl.copyRank.push_back(c.rank);
            pl.copySlot.push_back(c.slot);
        }
        pl.copyOff.push_back((int32_t)pl.copyRank.size());
        pl.maxCopies = std::max(pl.maxCopies, (int)cv.size() + 1);
    }
    return pl;
}

} // namespace sm
