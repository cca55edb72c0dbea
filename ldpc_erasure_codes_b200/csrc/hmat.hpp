// hmat.hpp -- H-matrix loader and host-side code model of libldpc_cuda.
//
// Reads the reference's committed code files (Matlab/*.mat: MAT-v5, one
// zlib-compressed miMATRIX holding sparse double `H_sparse`)
// and derives everything the kernels need: check->variable rows (the
// reference's "Vlist", OpenCL/device/LDPC_Vlist_data.h:20), variable->check
// rows, and the level schedule of the back-substitution encoder
// (OpenCL/device/ldpc_erasure_encoder.cl:72-83).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace ldpc {

struct HostCode {
    int n = 0, k = 0, m = 0, nnz = 0;
    std::vector<int32_t> row_ptr;   // [m+1]
    std::vector<int32_t> col_idx;   // [nnz] ascending inside a row, 0-based
    std::vector<int32_t> col_ptr;   // [n+1]
    std::vector<int32_t> row_idx;   // [nnz] ascending inside a column
    int max_row_weight = 0, max_col_weight = 0;
    bool triangular = false;        // last entry of row r is column k + r
    int encode_levels = 0;
    // padded device tables
    int RW = 0;                     // u16 entries per check row (multiple of 8), pad 0xFFFF
    int VW = 0;                     // u16 entries per variable row (4 or 8 ...), pad 0xFFFF
    std::vector<uint16_t> cidx;     // [m][RW]
    std::vector<uint16_t> vadj;     // [n][VW]
    // static encode schedule blob (same format the peel compiler emits per codeword)
    std::vector<uint8_t> enc_blob;
};

// Schedule blob: u32 hdr[4] = {n_entries, n_levels, info0, info1}; u32 entries[n_entries]
// (variable | check << 16) sorted by level; u16 lvl_off[n_levels + 1]; padded to 16 bytes.
inline int sched_blob_max_bytes(int m)
{
    int b = 16 + 4 * m + 2 * (m + 1);
    return (b + 15) & ~15;
}

// Returns 0 or a negative LDPC_ERR_* code; `err` receives a description.
int load_mat_sparse(const std::string &path, const char *var_name, int &rows, int &cols,
                    std::vector<int32_t> &col_ptr, std::vector<int32_t> &row_idx, std::string &err);
int build_code(int rows, int cols, const std::vector<int32_t> &col_ptr,
               const std::vector<int32_t> &row_idx, HostCode &code, std::string &err);

}  // namespace ldpc
