// hgen.cpp -- girth-8 triangular-form H generator and short-cycle checker.
//
// Offline code-design tools of the reference, rebuilt as host C++ behind the C ABI so that the codes
// the paper mentions but does not commit -- (4080,3060), other rates -- can be made and fed to the loader:
//   * Matlab/Cycle_Finder_length4_fromroot.m:3-19   -> cycle4_from_root
//   * Matlab/Cycle_Finder_length6.m:1-76            -> cycle6_from_root
//   * Matlab/Hgen_irregularDegree_no6cycles_systematic_encoding.m:94-224 ("bit filling": check rows are filled
//     one after the other with variables drawn with probability ~ (edges the variable still needs)^3, a draw is
//     kept iff it closes no 4- or 6-cycle; row r ends with the diagonal edge (r, k + r), so the right m x m part
//     is lower triangular and the code encodes by back-substitution; the last row keeps only its diagonal edge
//     and parity columns left with one edge get a staircase edge below the diagonal, :215-224)  -> ldpc_h_generate
// The MATLAB script draws from `rand`, whose stream is not reproducible here: the draws come from a seeded
// xorshift generator instead, so parity with the reference is by PROPERTY (triangular form, degree profile,
// girth >= 8 as judged by the restated cycle finders) -- and the cycle finders themselves are deterministic and
// are checked against the reference's committed codes (no short cycles) and a Python restatement.
// Differences from the script, all stated: the row weight follows the check profile row by row (the script uses its
// first entry for every row, :49,118); indices are 0-based.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ldpc_cuda.h"

namespace {

struct Graph {
    int n = 0, m = 0;
    std::vector<std::vector<int>> vlist;   // check -> variables   (the script's Vlist)
    std::vector<std::vector<int>> clist;   // variable -> checks   (the script's Clist)
};

// Cycle_Finder_length4_fromroot.m: the variables reached through the root's checks (the root excepted) must be distinct
bool cycle4_from_root(const Graph &g, int vroot, std::vector<int> &stamp, int &stamp_id)
{
    ++stamp_id;
    for (int c : g.clist[size_t(vroot)])
        for (int v : g.vlist[size_t(c)]) {
            if (v == vroot) continue;
            if (stamp[size_t(v)] == stamp_id) return true;
            stamp[size_t(v)] = stamp_id;
        }
    return false;
}

// Cycle_Finder_length6.m: a 4-cycle through the root counts (:72-74); else the checks of tier 2 -- reached from the
// tier-1 variables through every check except the one they were reached by -- must be distinct (:44-66)
bool cycle6_from_root(const Graph &g, int vroot, std::vector<int> &vstamp, std::vector<int> &cstamp, int &stamp_id)
{
    if (cycle4_from_root(g, vroot, vstamp, stamp_id)) return true;
    ++stamp_id;
    for (int c1 : g.clist[size_t(vroot)])
        for (int v1 : g.vlist[size_t(c1)]) {
            if (v1 == vroot) continue;
            for (int c2 : g.clist[size_t(v1)]) {
                if (c2 == c1) continue;
                if (cstamp[size_t(c2)] == stamp_id) return true;
                cstamp[size_t(c2)] = stamp_id;
            }
        }
    return false;
}

struct Rng {   // xorshift64*
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull) { if (!s) s = 1; }
    uint64_t next() { s ^= s >> 12; s ^= s << 25; s ^= s >> 27; return s * 0x2545F4914F6CDD1Dull; }
    double uniform() { return double(next() >> 11) * (1.0 / 9007199254740992.0); }   // [0, 1)
};

thread_local std::string g_hgen_err;

}  // namespace

extern "C" const char *ldpc_h_last_error_string(void) { return g_hgen_err.c_str(); }

extern "C" int ldpc_h_count_short_cycles(const int32_t *row_ptr, const int32_t *col_idx, int m, int n,
                                          int64_t *vars_on_4cycles, int64_t *vars_on_6cycles)
{
    if (!row_ptr || !col_idx || m <= 0 || n <= 0) { g_hgen_err = "bad argument to ldpc_h_count_short_cycles"; return LDPC_ERR_ARG; }
    Graph g;
    g.n = n; g.m = m;
    g.vlist.resize(size_t(m));
    g.clist.resize(size_t(n));
    for (int r = 0; r < m; r++)
        for (int j = row_ptr[r]; j < row_ptr[r + 1]; j++) {
            const int v = col_idx[j];
            if (v < 0 || v >= n) { g_hgen_err = "column index out of range"; return LDPC_ERR_ARG; }
            g.vlist[size_t(r)].push_back(v);
            g.clist[size_t(v)].push_back(r);
        }
    std::vector<int> vst(size_t(n), 0), cst(size_t(m), 0);
    int id = 0;
    int64_t n4 = 0, n6 = 0;
    for (int v = 0; v < n; v++) {
        if (cycle4_from_root(g, v, vst, id)) n4++;
        if (cycle6_from_root(g, v, vst, cst, id)) n6++;       // (includes the 4-cycle roots, as the script's function does)
    }
    if (vars_on_4cycles) *vars_on_4cycles = n4;
    if (vars_on_6cycles) *vars_on_6cycles = n6;
    return LDPC_OK;
}

extern "C" int ldpc_h_generate(const int32_t *deg_c_prof, int n_c_deg, const int32_t *deg_v_prof, int n_v_deg,
                               uint64_t seed, int max_tries, int32_t dims[4], int32_t *row_ptr, int32_t *col_idx,
                               int64_t col_cap, int32_t *tries_used)
{
    if (!deg_c_prof || !deg_v_prof || n_c_deg <= 0 || n_v_deg <= 0 || !dims || max_tries <= 0) {
        g_hgen_err = "bad argument to ldpc_h_generate";
        return LDPC_ERR_ARG;
    }
    // profiles: rows of (count, degree), degrees in descending order (Hgen...m:10-14)
    long long n = 0, m = 0, ev = 0, ec = 0;
    for (int i = 0; i < n_v_deg; i++) { n += deg_v_prof[2 * i]; ev += (long long)deg_v_prof[2 * i] * deg_v_prof[2 * i + 1]; }
    for (int i = 0; i < n_c_deg; i++) { m += deg_c_prof[2 * i]; ec += (long long)deg_c_prof[2 * i] * deg_c_prof[2 * i + 1]; }
    if (n <= m || m < 2 || n > 65535) { g_hgen_err = "degree profiles give no valid (n, k)"; return LDPC_ERR_ARG; }
    if (ev != ec) { g_hgen_err = "bad degree profile: variable and check edge counts differ (Hgen...m:64-66)"; return LDPC_ERR_ARG; }
    const int N = int(n), M = int(m), K = N - M;
    std::vector<int> dv(size_t(N), 0), dc(size_t(M), 0);
    for (int i = 0, at = 0; i < n_v_deg; i++)
        for (int j = 0; j < deg_v_prof[2 * i]; j++) dv[size_t(at++)] = deg_v_prof[2 * i + 1];
    for (int i = 0, at = 0; i < n_c_deg; i++)
        for (int j = 0; j < deg_c_prof[2 * i]; j++) dc[size_t(at++)] = deg_c_prof[2 * i + 1];

    Rng rng(seed);
    Graph g;
    std::vector<int> vst(size_t(N), 0), cst(size_t(M), 0), tried(size_t(N), 0);
    int id = 0, tried_id = 0, tries = 0, best_row = 0;
    bool done = false;
    while (!done && tries < max_tries) {
        tries++;
        g = Graph();
        g.n = N; g.m = M;
        g.vlist.assign(size_t(M), {});
        g.clist.assign(size_t(N), {});
        std::vector<int> dcur(size_t(N), 0), temp_dv = dv;
        bool ok = true;
        int ii = 0;
        for (; ii < M - 1 && ok; ii++) {                              // rows 0 .. m-2 (:111); the last row keeps its diagonal only
            if (double(ii + 1) / M > 0.997)                           // the last rows may exceed the variable profile by one (:114-116)
                for (int v = 0; v < N; v++) temp_dv[size_t(v)] = dv[size_t(v)] + 1;
            const int want = dc[size_t(ii)] - 1;                     // edges left of the diagonal
            int v_count = 0;
            ++tried_id;
            while (v_count < want) {
                // candidates: variables that still need edges, left of this row's diagonal (triangle property, :127),
                // not tried for this row yet; drawn with probability ~ (edges still needed)^3 (:138-150)
                double total = 0.0;
                for (int v = 0; v < K + ii; v++)
                    if (tried[size_t(v)] != tried_id && temp_dv[size_t(v)] > dcur[size_t(v)]) {
                        const double d = double(temp_dv[size_t(v)] - dcur[size_t(v)]);
                        total += d * d * d;
                    }
                if (total <= 0.0) break;
                const double value = rng.uniform() * total;
                double cum = 0.0;
                int cur = -1;
                for (int v = 0; v < K + ii; v++)
                    if (tried[size_t(v)] != tried_id && temp_dv[size_t(v)] > dcur[size_t(v)]) {
                        const double d = double(temp_dv[size_t(v)] - dcur[size_t(v)]);
                        cum += d * d * d;
                        cur = v;
                        if (cum > value) break;
                    }
                tried[size_t(cur)] = tried_id;
                // the graph WITH the candidate edge (:163-172): keep it iff no 4- or 6-cycle runs through the variable
                g.vlist[size_t(ii)].push_back(cur);
                g.clist[size_t(cur)].push_back(ii);
                if (cycle6_from_root(g, cur, vst, cst, id)) {
                    g.vlist[size_t(ii)].pop_back();
                    g.clist[size_t(cur)].pop_back();
                } else {
                    dcur[size_t(cur)]++;
                    v_count++;
                }
            }
            if (v_count < want) ok = false;                           // this try failed (:185-187)
            g.vlist[size_t(ii)].push_back(K + ii);                    // the triangle edge (:189-194)
            g.clist[size_t(K + ii)].push_back(ii);
            dcur[size_t(K + ii)]++;
        }
        done = ok && ii == M - 1;
        best_row = std::max(best_row, ii);
        if (done) {
            // final triangle edge in the bottom right corner (:214), staircase edge for parity columns left with one edge
            // (:216-222).  The script adds these without looking for cycles; here a try whose fix-up edges close a short
            // cycle is discarded like any other failed try.
            g.vlist[size_t(M - 1)].push_back(N - 1);
            g.clist[size_t(N - 1)].push_back(M - 1);
            for (int v = K; v < N - 1; v++)
                if (g.clist[size_t(v)].size() == 1) {
                    const int r = v + 1 - K;
                    g.vlist[size_t(r)].push_back(v);
                    g.clist[size_t(v)].push_back(r);
                }
            for (int v = K; v < N && done; v++)
                if (cycle6_from_root(g, v, vst, cst, id)) done = false;
        }
    }
    if (tries_used) *tries_used = tries;
    if (!done) {
        g_hgen_err = "no girth-8 matrix with this profile in " + std::to_string(max_tries) + " tries (best try filled " +
                     std::to_string(best_row) + " of " + std::to_string(M) + " rows)";
        return LDPC_ERR_UNSUPPORTED;
    }
    long long nnz = 0;
    for (int r = 0; r < M; r++) nnz += (long long)g.vlist[size_t(r)].size();
    dims[0] = M; dims[1] = N; dims[2] = int32_t(nnz); dims[3] = 1;
    if (row_ptr && col_idx) {
        if (col_cap < nnz) { g_hgen_err = "col_idx too small"; return LDPC_ERR_ARG; }
        int32_t at = 0;
        for (int r = 0; r < M; r++) {
            row_ptr[r] = at;
            std::vector<int> row = g.vlist[size_t(r)];
            std::sort(row.begin(), row.end());
            for (int v : row) col_idx[at++] = v;
        }
        row_ptr[M] = at;
    }
    return LDPC_OK;
}
