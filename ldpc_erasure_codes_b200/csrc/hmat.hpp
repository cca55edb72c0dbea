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
    // static encode schedule (same form the peel compiler emits per codeword): entries sorted by level + level offsets
    std::vector<uint32_t> enc_entries;   // (k + r) | r << 16
    std::vector<uint16_t> enc_lvl_off;   // [encode_levels + 1]
};

// Schedule blob: u32 hdr[4] = {n_entries, n_levels | n_records << 16, info0 | n_passes << 16, info1}; u32
// entries[n_entries] (variable | check << 16) sorted by level; u16 lvl_off[n_levels + 1]; u16 passes[n_passes]: the
// executor's walk over the levels >= 2 cut into passes of at most `epw` entries of ONE level, first entry | (count - 1)
// << 11; padded to 8 bytes; u64 records[n_records], one per entry of level >= 2 in entry order: the entry's PRODUCED
// members (members of its check that earlier entries produce) as 5 x 12-bit symbol indices (padding = the executor's
// zero row), bit 63 = more than 5 (none listed); padded to 16 bytes.  n_records may be less than the number of such
// entries (the executor then applies the rest in full-row form).  The peel kernel writes hdr, entries and lvl_off
// (n_records = n_passes = 0); the executor adds passes and records in shared memory (payload_exec.cuh).  The encoder's
// static blob is complete (make_enc_blob).
inline int sched_blob_max_bytes(int m)       // what the peel kernel writes per codeword: header, entries, level offsets
{
    return (16 + 4 * m + 2 * (m + 1) + 15) & ~15;
}
inline int sched_area_bytes(int m, int records)   // a blob in the executor's shared memory: + passes + records
{
    return (sched_blob_max_bytes(m) + 2 * m + 8 * records + 15) & ~15;
}
constexpr int kSchedMaxRow = 4095;           // 12-bit symbol indices in the records
inline int sched_zero_row(int n)             // the executor's all-zero row: behind its 256-row boxes, or (n > 3840) row 4095
{
    const int z = ((n + 255) / 256) * 256;
    return z < kSchedMaxRow ? z : kSchedMaxRow;
}

// Returns 0 or a negative LDPC_ERR_* code; `err` receives a description.
int load_mat_sparse(const std::string &path, const char *var_name, int &rows, int &cols,
                    std::vector<int32_t> &col_ptr, std::vector<int32_t> &row_idx, std::string &err);
int build_code(int rows, int cols, const std::vector<int32_t> &col_ptr,
               const std::vector<int32_t> &row_idx, HostCode &code, std::string &err);
// The encoder's static schedule as a blob (empty if H is not triangular): epw > 0 -- for the executor, the rows
// list-scheduled into passes of <= epw independent entries; epw = 0 -- the plain level structure.
std::vector<uint8_t> make_enc_blob(const HostCode &code, int epw);

}  // namespace ldpc
