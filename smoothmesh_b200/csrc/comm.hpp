// comm.hpp -- multi-GPU exchange layer (NCCL) replacing OpenFOAM's
// syncTools::syncPointList / returnReduce on the hot path.
#pragma once
#include <cstdint>
#include <string>

struct smgpu_handle;

namespace sm
{
struct Comm;
bool commUniqueId(uint8_t id[128], std::string &err);
Comm *commCreate(smgpu_handle *h, int rank, int nRanks, const uint8_t id[128], std::string &err);
void commDestroy(Comm *c);
// Runs iteration `it` with the interface exchanges; returns 1 when the global
// residual fell below relTol (stop), 0 otherwise.  Throws std::runtime_error.
int commIterate(Comm *c, smgpu_handle *h, int it);
} // namespace sm
