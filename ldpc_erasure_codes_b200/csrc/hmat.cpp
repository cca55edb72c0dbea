// hmat.cpp -- MAT-v5 reader + code-model builder (see hmat.hpp).
#include "hmat.hpp"

#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstring>

#include "../../include/ldpc_cuda.h"

namespace ldpc {
namespace {

enum { miINT8 = 1, miUINT8 = 2, miINT16 = 3, miUINT16 = 4, miINT32 = 5, miUINT32 = 6, miSINGLE = 7,
       miDOUBLE = 9, miINT64 = 12, miUINT64 = 13, miMATRIX = 14, miCOMPRESSED = 15 };

struct Elem {
    uint32_t type = 0;
    uint32_t nbytes = 0;
    const uint8_t *data = nullptr;
    size_t total = 0;  // bytes consumed including tag and padding
};

uint32_t rd32(const uint8_t *p)
{
    uint32_t v;
    memcpy(&v, p, 4);
    return v;
}

// Parses one data element tag at p (avail bytes left).  Handles the "small data element"
// packing (type in the low half-word, byte count in the high half-word, data in the tag).
bool parse_elem(const uint8_t *p, size_t avail, Elem &e)
{
    if (avail < 8) return false;
    uint32_t w0 = rd32(p);
    if (w0 >> 16) {
        e.type = w0 & 0xFFFFu;
        e.nbytes = w0 >> 16;
        if (e.nbytes > 4) return false;
        e.data = p + 4;
        e.total = 8;
        return true;
    }
    e.type = w0;
    e.nbytes = rd32(p + 4);
    e.data = p + 8;
    size_t padded = (e.type == miCOMPRESSED) ? e.nbytes : ((size_t(e.nbytes) + 7) & ~size_t(7));
    if (8 + size_t(e.nbytes) > avail) return false;
    e.total = std::min(avail, 8 + padded);
    return true;
}

size_t type_size(uint32_t t)
{
    switch (t) {
        case miINT8: case miUINT8: return 1;
        case miINT16: case miUINT16: return 2;
        case miINT32: case miUINT32: case miSINGLE: return 4;
        case miDOUBLE: case miINT64: case miUINT64: return 8;
        default: return 0;
    }
}

bool numeric_at(const Elem &e, size_t i, double &v)
{
    const uint8_t *p = e.data + i * type_size(e.type);
    switch (e.type) {
        case miINT8: v = *reinterpret_cast<const int8_t *>(p); return true;
        case miUINT8: v = *p; return true;
        case miINT16: { int16_t x; memcpy(&x, p, 2); v = x; return true; }
        case miUINT16: { uint16_t x; memcpy(&x, p, 2); v = x; return true; }
        case miINT32: { int32_t x; memcpy(&x, p, 4); v = x; return true; }
        case miUINT32: { uint32_t x; memcpy(&x, p, 4); v = x; return true; }
        case miSINGLE: { float x; memcpy(&x, p, 4); v = x; return true; }
        case miDOUBLE: { double x; memcpy(&x, p, 8); v = x; return true; }
        case miINT64: { int64_t x; memcpy(&x, p, 8); v = double(x); return true; }
        case miUINT64: { uint64_t x; memcpy(&x, p, 8); v = double(x); return true; }
        default: return false;
    }
}

bool inflate_all(const uint8_t *src, size_t n, std::vector<uint8_t> &out)
{
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit(&zs) != Z_OK) return false;
    zs.next_in = const_cast<Bytef *>(src);
    zs.avail_in = static_cast<uInt>(n);
    out.resize(std::max<size_t>(n * 8, 4096));
    size_t have = 0;
    int rc;
    do {
        if (have == out.size()) out.resize(out.size() * 2);
        zs.next_out = out.data() + have;
        zs.avail_out = static_cast<uInt>(out.size() - have);
        rc = inflate(&zs, Z_NO_FLUSH);
        have = out.size() - zs.avail_out;
    } while (rc == Z_OK);
    inflateEnd(&zs);
    if (rc != Z_STREAM_END) return false;
    out.resize(have);
    return true;
}

// Parses one miMATRIX body.  On a name match fills the CSC arrays and returns 1; returns 0
// if the variable is another one, <0 on a format error.
int parse_matrix(const uint8_t *p, size_t n, const char *var_name, int &rows, int &cols,
                 std::vector<int32_t> &col_ptr, std::vector<int32_t> &row_idx, std::string &err)
{
    size_t off = 0;
    Elem flags, dims, name;
    if (!parse_elem(p + off, n - off, flags) || flags.nbytes < 8) { err = "bad array-flags element"; return LDPC_ERR_FORMAT; }
    off += flags.total;
    if (!parse_elem(p + off, n - off, dims) || dims.type != miINT32 || dims.nbytes < 8) { err = "bad dimensions element"; return LDPC_ERR_FORMAT; }
    off += dims.total;
    if (!parse_elem(p + off, n - off, name)) { err = "bad name element"; return LDPC_ERR_FORMAT; }
    off += name.total;
    std::string nm(reinterpret_cast<const char *>(name.data), name.nbytes);
    if (var_name && nm != var_name) return 0;
    const uint32_t cls = rd32(flags.data) & 0xFFu;
    if (dims.nbytes != 8) { err = "variable '" + nm + "' is not 2-D"; return LDPC_ERR_FORMAT; }
    int32_t d0, d1;
    memcpy(&d0, dims.data, 4);
    memcpy(&d1, dims.data + 4, 4);
    if (d0 <= 0 || d1 <= 0) { err = "empty matrix"; return LDPC_ERR_FORMAT; }
    rows = d0;
    cols = d1;
    if (cls == 5) {  // mxSPARSE_CLASS: ir, jc, pr
        Elem ir, jc, pr;
        if (!parse_elem(p + off, n - off, ir)) { err = "missing ir"; return LDPC_ERR_FORMAT; }
        off += ir.total;
        if (!parse_elem(p + off, n - off, jc)) { err = "missing jc"; return LDPC_ERR_FORMAT; }
        off += jc.total;
        if (!parse_elem(p + off, n - off, pr)) { err = "missing pr"; return LDPC_ERR_FORMAT; }
        const size_t sz_ir = type_size(ir.type), sz_jc = type_size(jc.type), sz_pr = type_size(pr.type);
        if (!sz_ir || !sz_jc || !sz_pr) { err = "non-numeric sparse arrays"; return LDPC_ERR_FORMAT; }
        const size_t n_jc = jc.nbytes / sz_jc;
        if (n_jc != size_t(cols) + 1) { err = "jc length != columns + 1"; return LDPC_ERR_FORMAT; }
        col_ptr.resize(n_jc);
        for (size_t i = 0; i < n_jc; i++) { double v; numeric_at(jc, i, v); col_ptr[i] = int32_t(v); }
        const size_t nnz = size_t(col_ptr[cols]);
        if (col_ptr[0] != 0 || ir.nbytes / sz_ir < nnz || pr.nbytes / sz_pr < nnz) { err = "inconsistent ir/jc/pr sizes"; return LDPC_ERR_FORMAT; }
        row_idx.clear();
        std::vector<int32_t> cp(cols + 1, 0);
        for (int c = 0; c < cols; c++) {
            if (col_ptr[c + 1] < col_ptr[c]) { err = "jc not monotone"; return LDPC_ERR_FORMAT; }
            for (int32_t j = col_ptr[c]; j < col_ptr[c + 1]; j++) {
                double r, v;
                numeric_at(ir, size_t(j), r);
                numeric_at(pr, size_t(j), v);
                if (r < 0 || r >= rows) { err = "row index out of range"; return LDPC_ERR_FORMAT; }
                const long iv = long(v);
                if (double(iv) != v) { err = "H entry is not an integer"; return LDPC_ERR_FORMAT; }
                if (iv & 1) row_idx.push_back(int32_t(r));  // entries are taken mod 2
            }
            cp[c + 1] = int32_t(row_idx.size());
        }
        col_ptr = cp;
        return 1;
    }
    if (cls >= 6 && cls <= 15) {  // full numeric matrix, column-major
        Elem pr;
        if (!parse_elem(p + off, n - off, pr)) { err = "missing pr"; return LDPC_ERR_FORMAT; }
        const size_t sz = type_size(pr.type);
        if (!sz || pr.nbytes / sz < size_t(rows) * size_t(cols)) { err = "full matrix data too short"; return LDPC_ERR_FORMAT; }
        col_ptr.assign(cols + 1, 0);
        row_idx.clear();
        for (int c = 0; c < cols; c++) {
            for (int r = 0; r < rows; r++) {
                double v;
                numeric_at(pr, size_t(c) * rows + r, v);
                if (long(v) & 1) row_idx.push_back(r);
            }
            col_ptr[c + 1] = int32_t(row_idx.size());
        }
        return 1;
    }
    err = "variable '" + nm + "' has unsupported class " + std::to_string(cls);
    return LDPC_ERR_FORMAT;
}

}  // namespace

int load_mat_sparse(const std::string &path, const char *var_name, int &rows, int &cols,
                    std::vector<int32_t> &col_ptr, std::vector<int32_t> &row_idx, std::string &err)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) { err = "cannot open " + path; return LDPC_ERR_IO; }
    std::vector<uint8_t> buf;
    uint8_t tmp[65536];
    size_t got;
    while ((got = fread(tmp, 1, sizeof(tmp), f)) > 0) buf.insert(buf.end(), tmp, tmp + got);
    fclose(f);
    if (buf.size() < 136) { err = path + ": too short for a MAT-v5 file"; return LDPC_ERR_FORMAT; }
    if (memcmp(buf.data(), "MATLAB 5.0 MAT-file", 19) != 0) { err = path + ": not a MAT-v5 file"; return LDPC_ERR_FORMAT; }
    if (!(buf[126] == 'I' && buf[127] == 'M')) { err = path + ": big-endian MAT files are not supported"; return LDPC_ERR_FORMAT; }
    size_t off = 128;
    while (off + 8 <= buf.size()) {
        Elem e;
        if (!parse_elem(buf.data() + off, buf.size() - off, e)) { err = path + ": truncated data element"; return LDPC_ERR_FORMAT; }
        const uint8_t *body = e.data;
        size_t body_n = e.nbytes;
        std::vector<uint8_t> inflated;
        uint32_t type = e.type;
        if (type == miCOMPRESSED) {
            if (!inflate_all(e.data, e.nbytes, inflated)) { err = path + ": zlib inflate failed"; return LDPC_ERR_FORMAT; }
            Elem inner;
            if (!parse_elem(inflated.data(), inflated.size(), inner)) { err = path + ": bad compressed element"; return LDPC_ERR_FORMAT; }
            type = inner.type;
            body = inner.data;
            body_n = inner.nbytes;
        }
        if (type == miMATRIX) {
            int rc = parse_matrix(body, body_n, var_name, rows, cols, col_ptr, row_idx, err);
            if (rc < 0) { err = path + ": " + err; return rc; }
            if (rc == 1) return LDPC_OK;
        }
        off += e.total;
    }
    err = path + ": variable '" + std::string(var_name ? var_name : "?") + "' not found";
    return LDPC_ERR_FORMAT;
}

int build_code(int rows, int cols, const std::vector<int32_t> &col_ptr,
               const std::vector<int32_t> &row_idx, HostCode &code, std::string &err)
{
    code = HostCode();
    code.m = rows;
    code.n = cols;
    code.k = cols - rows;
    code.nnz = int(row_idx.size());
    if (code.k <= 0) { err = "H must have more columns than rows"; return LDPC_ERR_FORMAT; }
    code.col_ptr = col_ptr;
    code.row_idx = row_idx;
    // CSC -> CSR (columns visited ascending, so each row comes out ascending)
    code.row_ptr.assign(rows + 1, 0);
    for (int32_t r : row_idx) code.row_ptr[r + 1]++;
    for (int r = 0; r < rows; r++) code.row_ptr[r + 1] += code.row_ptr[r];
    code.col_idx.resize(row_idx.size());
    std::vector<int32_t> fill(code.row_ptr.begin(), code.row_ptr.end() - 1);
    for (int c = 0; c < cols; c++) {
        for (int32_t j = col_ptr[c]; j < col_ptr[c + 1]; j++) {
            if (j > col_ptr[c] && row_idx[j] <= row_idx[j - 1]) { err = "row indices not ascending inside a column"; return LDPC_ERR_FORMAT; }
            code.col_idx[fill[row_idx[j]]++] = c;
        }
        code.max_col_weight = std::max(code.max_col_weight, col_ptr[c + 1] - col_ptr[c]);
    }
    code.triangular = true;
    for (int r = 0; r < rows; r++) {
        const int w = code.row_ptr[r + 1] - code.row_ptr[r];
        code.max_row_weight = std::max(code.max_row_weight, w);
        if (w == 0 || code.col_idx[code.row_ptr[r + 1] - 1] != code.k + r) code.triangular = false;
    }
    if (code.n > kSchedMaxRow - 1 || code.m > 2047 || code.max_row_weight > 30) {   // (12-bit symbol indices in the peel state and the schedule records; the peel kernel's 5-bit erased-member count keeps 31 as its "fired" mark)
        err = "code outside kernel limits (n <= 4094, m <= 2047, row weight <= 30)";
        return LDPC_ERR_UNSUPPORTED;
    }
    code.RW = (code.max_row_weight + 7) & ~7;
    code.VW = code.max_col_weight <= 4 ? 4 : ((code.max_col_weight + 7) & ~7);
    code.cidx.assign(size_t(rows) * code.RW, 0xFFFFu);
    for (int r = 0; r < rows; r++)
        for (int j = code.row_ptr[r]; j < code.row_ptr[r + 1]; j++)
            code.cidx[size_t(r) * code.RW + (j - code.row_ptr[r])] = uint16_t(code.col_idx[j]);
    code.vadj.assign((size_t(cols) * code.VW + 7) & ~size_t(7), 0xFFFFu);   // padded to 16 bytes: the kernels stage it with 128-bit copies
    for (int c = 0; c < cols; c++)
        for (int j = col_ptr[c]; j < col_ptr[c + 1]; j++)
            code.vadj[size_t(c) * code.VW + (j - col_ptr[c])] = uint16_t(row_idx[j]);

    // Encoder level schedule: parity r = XOR of the row's members except its last one (the
    // diagonal), so row r depends on the rows whose parity it references.
    if (code.triangular) {
        std::vector<int> level(rows, 1);
        int nl = 0;
        for (int r = 0; r < rows; r++) {
            int l = 0;
            for (int j = code.row_ptr[r]; j < code.row_ptr[r + 1] - 1; j++) {
                const int c = code.col_idx[j];
                if (c >= code.k) l = std::max(l, level[c - code.k]);
            }
            level[r] = l + 1;
            nl = std::max(nl, level[r]);
        }
        code.encode_levels = nl;
        std::vector<uint16_t> lvl_off(nl + 1, 0);
        for (int r = 0; r < rows; r++) lvl_off[level[r]]++;
        for (int l = 1; l <= nl; l++) lvl_off[l] = uint16_t(lvl_off[l] + lvl_off[l - 1]);
        std::vector<uint32_t> entries(rows);
        std::vector<uint16_t> pos(lvl_off.begin(), lvl_off.end());
        for (int r = 0; r < rows; r++) entries[pos[level[r] - 1]++] = uint32_t(code.k + r) | (uint32_t(r) << 16);
        code.enc_entries = entries;
        code.enc_lvl_off = lvl_off;
    }
    return LDPC_OK;
}

std::vector<uint8_t> make_enc_blob(const HostCode &code, int epw)
{
    std::vector<uint8_t> blob;
    if (!code.triangular) return blob;
    const int rows = code.m, k = code.k;
    // parity members of every row (its diagonal excluded): what the row needs from earlier rows
    std::vector<std::vector<int>> deps(static_cast<size_t>(rows));
    for (int r = 0; r < rows; r++)
        for (int j = code.row_ptr[r]; j < code.row_ptr[r + 1] - 1; j++)
            if (code.col_idx[j] >= k) deps[size_t(r)].push_back(code.col_idx[j] - k);
    std::vector<uint32_t> entries;
    std::vector<uint16_t> lvl_off, passes;
    int nl = 0, n1 = rows;
    if (epw <= 0) {
        // plain level structure (what nb_exec_kernel walks)
        entries = code.enc_entries;
        lvl_off = code.enc_lvl_off;
        nl = code.encode_levels;
        n1 = nl >= 2 ? int(lvl_off[1]) : rows;
    } else {
        // The executor walks PASSES of <= epw mutually independent entries, one after the other (payload_exec.cuh), and a
        // pass costs the same whether it is full or not.  The levels of this schedule hold ~17-26 entries: cut level by
        // level they give two passes each, the second nearly empty.  Instead the rows are list-scheduled: a pass takes up
        // to epw rows whose parity members all lie in earlier passes (or need none), the rows with the longest chain of
        // dependents first.  Entry order = rows without parity members, then pass by pass (a topological order, which is
        // all the executor's bulk pass needs); the blob presents them as two "levels".
        std::vector<int> height(size_t(rows), 1), pass_of(size_t(rows), -2);
        for (int r = rows - 1; r >= 0; r--)
            for (int d : deps[size_t(r)]) height[size_t(d)] = std::max(height[size_t(d)], height[size_t(r)] + 1);
        for (int r = 0; r < rows; r++)
            if (deps[size_t(r)].empty()) { pass_of[size_t(r)] = -1; entries.push_back(uint32_t(k + r) | (uint32_t(r) << 16)); }
        n1 = int(entries.size());
        int left = rows - n1;
        for (int p = 0; left > 0; p++) {
            std::vector<int> ready;
            for (int r = 0; r < rows; r++) {
                if (pass_of[size_t(r)] != -2) continue;
                bool ok = true;
                for (int d : deps[size_t(r)]) if (pass_of[size_t(d)] == -2 || pass_of[size_t(d)] >= p) { ok = false; break; }
                if (ok) ready.push_back(r);
            }
            std::stable_sort(ready.begin(), ready.end(), [&](int a, int b2) { return height[size_t(a)] > height[size_t(b2)]; });
            const int take = std::min<int>(epw, int(ready.size()));
            passes.push_back(uint16_t(int(entries.size()) | ((take - 1) << 11)));
            for (int i = 0; i < take; i++) {
                pass_of[size_t(ready[size_t(i)])] = p;
                entries.push_back(uint32_t(k + ready[size_t(i)]) | (uint32_t(ready[size_t(i)]) << 16));
            }
            left -= take;
        }
        nl = n1 < rows ? 2 : 1;
        lvl_off.push_back(0);
        lvl_off.push_back(uint16_t(n1));
        if (nl == 2) lvl_off.push_back(uint16_t(rows));
    }
    // records: the parity members of each walked row
    const int nrec = rows - n1;
    const uint64_t z = uint64_t(sched_zero_row(code.n));
    std::vector<uint64_t> recs(static_cast<size_t>(nrec), 0);
    for (int i = n1; i < n1 + nrec; i++) {
        const int r = int(entries[size_t(i)] >> 16);
        uint64_t rec = 0;
        int nd = 0;
        for (int d : deps[size_t(r)]) {
            if (nd < 5) rec |= uint64_t(k + d) << (12 * nd);
            nd++;
        }
        for (int j = nd; j < 5; j++) rec |= z << (12 * j);
        if (nd > 5) rec = z | (z << 12) | (z << 24) | (z << 36) | (z << 48) | (1ull << 63);   // full-row form: no member listed
        recs[size_t(i - n1)] = rec;
    }
    const size_t pt_off = 16 + size_t(rows) * 4 + size_t(nl + 1) * 2;
    const size_t rec_off = (pt_off + passes.size() * 2 + 7) & ~size_t(7);
    blob.assign((rec_off + size_t(nrec) * 8 + 15) & ~size_t(15), 0);
    uint32_t hdr[4] = {uint32_t(rows), uint32_t(nl) | (uint32_t(nrec) << 16), uint32_t(passes.size()) << 16, 0u};
    memcpy(blob.data(), hdr, 16);
    memcpy(blob.data() + 16, entries.data(), size_t(rows) * 4);
    memcpy(blob.data() + 16 + size_t(rows) * 4, lvl_off.data(), size_t(nl + 1) * 2);
    if (!passes.empty()) memcpy(blob.data() + pt_off, passes.data(), passes.size() * 2);
    if (nrec) memcpy(blob.data() + rec_off, recs.data(), size_t(nrec) * 8);
    return blob;
}

}  // namespace ldpc
