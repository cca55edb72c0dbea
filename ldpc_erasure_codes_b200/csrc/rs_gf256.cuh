// rs_gf256.cuh -- Reed-Solomon GF(2^8) erasure codec.  PLACEHOLDER for the first bring-up.
#pragma once
#include <string>

#include "../../include/ldpc_cuda.h"

struct rs_ctx {
    int dummy;
};
namespace ldpc {
inline int rs_create_impl(rs_ctx **, int, int, int, int, int64_t, std::string &err) { err = "RS codec is not implemented yet"; return LDPC_ERR_UNSUPPORTED; }
inline int rs_destroy_impl(rs_ctx *) { return LDPC_OK; }
inline int rs_get_generator_impl(const rs_ctx *, uint8_t *, std::string &err) { err = "RS codec is not implemented yet"; return LDPC_ERR_UNSUPPORTED; }
inline int rs_encode_impl(rs_ctx *, const void *, void *, int64_t, cudaStream_t, std::string &err) { err = "RS codec is not implemented yet"; return LDPC_ERR_UNSUPPORTED; }
inline int rs_decode_impl(rs_ctx *, const void *, const uint32_t *, void *, uint8_t *, int64_t, cudaStream_t, std::string &err) { err = "RS codec is not implemented yet"; return LDPC_ERR_UNSUPPORTED; }
}  // namespace ldpc
