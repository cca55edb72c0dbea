// hybrid_ge.cuh -- hybrid-ML stage (GF(2) elimination on the residual stopping set).
// PLACEHOLDER for the first bring-up: the stage is not implemented yet and says so.
#pragma once
#include <string>

#include "../../include/ldpc_cuda.h"
#include "hmat.hpp"

namespace ldpc {
struct HybridScratch {
    int dummy = 0;
};
inline void hybrid_free(HybridScratch &) {}
inline int hybrid_stage(HybridScratch &, const HostCode &, const uint16_t *, int, int, int, const uint8_t *,
                        const uint32_t *, const uint8_t *, int, const uint32_t *, uint8_t *, uint8_t *,
                        unsigned long long *, long long, long long, cudaStream_t, std::string &err)
{
    err = "hybrid mode is not implemented yet";
    return LDPC_ERR_UNSUPPORTED;
}
}  // namespace ldpc
