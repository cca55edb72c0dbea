// ldpc_cuda.cu -- C ABI of libldpc_cuda (see include/ldpc_cuda.h) and the host runtime that
// stands where the reference's OpenCL host (OpenCL/host/src/main.cpp) stands: context
// creation (init_opencl :439-544), kernel launches (run :555-659), counters (data_out),
// teardown (cleanup :668-691).
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/ldpc_cuda.h"
#include "erasure_gen.cuh"
#include "hmat.hpp"
#include "fec_packets.cuh"
#include "host_inplace.cuh"
#include "hybrid_ge.cuh"
#include "nb_ldpc.cuh"
#include "payload_exec.cuh"
#include "peel_schedule.cuh"
#include "rs_gf256.cuh"

using namespace ldpc;

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local std::string g_err = "";

static int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                       \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess)                                                               \
            return fail(LDPC_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));  \
    } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to the CURRENT device's copy of the function: it is raised
// once per (device, function) to the device maximum minus the kernel's static shared memory.  Contexts on different
// GPUs may be created and used from different host threads, so the record is locked.
static int allow_max_smem(const void *func, int smem_optin)
{
    static std::mutex mu;
    static std::vector<std::pair<int, const void *>> done;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    for (const auto &d : done) if (d.first == dev && d.second == func) return LDPC_OK;
    cudaFuncAttributes a;
    CUDA_TRY(cudaFuncGetAttributes(&a, func));
    CUDA_TRY(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - int(a.sharedSizeBytes)));
    done.emplace_back(dev, func);
    return LDPC_OK;
}

// LDPC_CUDA_GE_STAGES: bit 0 = inactivation stage, bit 1 = per-warp elimination stage (default 3); the
// CTA-per-codeword kernel always runs last on what is left.  Tests use it to compare the solvers.
// LDPC_CUDA_GE_SPLIT=0: the inactivation stage solves pattern and payload in one kernel (tests compare both forms)
static bool ge_split() { const char *e = getenv("LDPC_CUDA_GE_SPLIT"); return !(e && *e == '0'); }
static int ge_stage_mask() { const char *e = getenv("LDPC_CUDA_GE_STAGES"); return e && *e ? atoi(e) : 3; }

// LDPC_CUDA_TMA4=0: the executor moves its slots box by box (3-D tensor copies only)
static const bool g_tma4 = [] { const char *e = getenv("LDPC_CUDA_TMA4"); return !(e && *e == '0'); }();

// LDPC_CUDA_EXEC_PLAIN=1: the executor applies per-codeword schedules level by level with full-row gathers (the form it
// falls back to when a blob leaves no room for its pass table; tests compare it with the bulk + walk form)
static bool exec_plain() { const char *e = getenv("LDPC_CUDA_EXEC_PLAIN"); return e && *e && *e != '0'; }
// LDPC_CUDA_ENC_SPLIT_STORE=0: the encoder stores a codeword in one piece after its first level and re-stores the walk's rows
// (the decoder's way) instead of information boxes early / parity boxes late
static bool exec_split_store() { const char *e = getenv("LDPC_CUDA_ENC_SPLIT_STORE"); return !(e && *e == '0'); }

// LDPC_CUDA_DEBUG_SYNC=1: synchronise after every launch so that a faulting kernel is named
static int debug_sync(const char *what, cudaStream_t st)
{
    static const bool on = [] { const char *e = getenv("LDPC_CUDA_DEBUG_SYNC"); return e && *e && *e != '0'; }();
    if (!on) return LDPC_OK;
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(LDPC_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    fprintf(stderr, "[ldpc_cuda] %s ok\n", what);
    return LDPC_OK;
}

extern "C" const char *ldpc_last_error_string(void) { return g_err.c_str(); }
extern "C" int ldpc_cuda_abi_version(void) { return LDPC_CUDA_ABI_VERSION; }

// ------------------------------------------------------------------------------------------
// driver entry point for tensor-map encoding (libcuda is not linked; resolved at run time)
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static int get_encode_fn(PFN_encodeTiled *fn)
{
    static const PFN_encodeTiled cached = [] {      // (initialised once, thread-safe)
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<PFN_encodeTiled>(p);
    }();
    if (!cached) return fail(LDPC_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    *fn = cached;
    return LDPC_OK;
}

// 3-D byte tensor [B][rows][S], box {W, 256, 1}
static int make_map(CUtensorMap *map, const void *base, int S, int rows, long long B, int W, bool is_load)
{
    PFN_encodeTiled enc;
    int rc = get_encode_fn(&enc);
    if (rc) return rc;
    cuuint64_t dims[3] = {cuuint64_t(S), cuuint64_t(rows), cuuint64_t(B)};
    cuuint64_t strides[2] = {cuuint64_t(S), cuuint64_t(S) * cuuint64_t(rows)};
    cuuint32_t box[3] = {cuuint32_t(W), cuuint32_t(kBoxRows), 1u};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     is_load ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(LDPC_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string(int(r)));
    return LDPC_OK;
}

// The same memory seen as [B][nfull][256][S], nfull = rows / 256 whole boxes per codeword: ONE tensor copy with box
// {W, 256, nfull, 1} moves what nfull 3-D boxes move (the TMA unit takes a few hundred cycles per request; a unit
// of the executor issued 8 + 6 of them).  Rows past nfull * 256 stay with the 3-D map, whose bounds clip them.
static int make_map4(CUtensorMap *map, const void *base, int S, int rows, long long B, int W, int nfull, bool is_load, int boxn = 0)
{
    if (boxn <= 0) boxn = nfull;      // boxes per copy (the encoder stores the whole boxes of a codeword in two parts)
    PFN_encodeTiled enc;
    int rc = get_encode_fn(&enc);
    if (rc) return rc;
    cuuint64_t dims[4] = {cuuint64_t(S), cuuint64_t(kBoxRows), cuuint64_t(nfull), cuuint64_t(B)};
    cuuint64_t strides[3] = {cuuint64_t(S), cuuint64_t(S) * kBoxRows, cuuint64_t(S) * cuuint64_t(rows)};
    cuuint32_t box[4] = {cuuint32_t(W), cuuint32_t(kBoxRows), cuuint32_t(boxn), 1u};
    cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     is_load ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(LDPC_ERR_CUDA, "cuTensorMapEncodeTiled (4-D) failed with code " + std::to_string(int(r)));
    return LDPC_OK;
}

struct ldpc_ctx;
static int cached_map(ldpc_ctx *c, CUtensorMap *out, const void *base, int rows, long long B, int W, int nfull, bool is_load, int boxn = 0);

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
struct ExecGeom {
    int W = 0, nslot = 0, slot_bytes = 0, smem_bytes = 0, sched_area = 0, msk_words = 0;
};

struct ldpc_ctx {
    HostCode code;
    int S = 0, device = 0, code_ind = -1;
    int rs_n = 0, rs_k = 0;
    long long max_batch = 0;
    int num_sms = 0, smem_optin = 0;
    int NW = 0, MW = 0, sched_stride = 0;
    // device tables
    uint16_t *d_cidx = nullptr, *d_vadj = nullptr;
    uint16_t *d_rows_dec = nullptr, *d_rows_enc = nullptr;   // check rows arranged for the executor's decode / encode geometry
    uint8_t *d_enc_blob = nullptr;
    // scratch
    uint8_t *d_sched = nullptr;
    uint32_t *d_sched_len = nullptr;
    uint32_t *d_resid = nullptr;
    uint8_t *d_fail_scratch = nullptr;
    unsigned long long *d_stats = nullptr;
    long long hybrid_batch = 0;              // codewords per hybrid-mode chunk (bounds the syndrome scratch)
    unsigned long long *d_phase = nullptr;   // LDPC_CUDA_PHASE_TIMING=1
    uint32_t *d_sim_mask = nullptr;          // ldpc_simulate_fer: masks of one chunk
    unsigned int *d_work_ctr = nullptr;      // peel kernel's codeword claim counter
    // A peel-mode batch of several max_batch chunks alternates its chunks between two internal streams (and two
    // sets of the scratch above), so that the tail of one chunk's kernels overlaps the start of the next chunk's.
    uint8_t *d_sched2 = nullptr; uint32_t *d_sched_len2 = nullptr; uint32_t *d_resid2 = nullptr;
    uint8_t *d_fail_scratch2 = nullptr; unsigned int *d_work_ctr2 = nullptr;
    cudaStream_t pstream[2] = {nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
    // geometry
    ExecGeom dec, enc;
    int force_W = 0, force_slots = 0;
    int peel_G = 8, peel_groups = 0, peel_smem = 0;
    // hybrid stage
    HybridScratch hyb;
    // host-buffer pipeline
    static constexpr int kHostStages = 3;    // stages of the host pipeline (streams + staging buffers): 3 keep the link busy while a stage decodes
    cudaStream_t hstream[kHostStages] = {};
    uint8_t *h_in[kHostStages] = {}, *h_out[kHostStages] = {};
    uint32_t *h_mask[kHostStages] = {};
    uint8_t *h_fail[kHostStages] = {};
    long long host_chunk = 0;
    int host_stages = 0;
    bool host_ready = false;
    uint8_t *h_fail_any[kHostStages] = {};
    cudaEvent_t h_ev = nullptr;              // orders the kernels of consecutive host-pipeline stages (they share the scratch)
    // tensor maps of recent executor launches (cuTensorMapEncodeTiled is a driver call on the small-batch latency path)
    struct MapKey { const void *base; int rows, W, nfull, is_load; long long B; int boxn; };
    struct MapRec { MapKey key; CUtensorMap map; };
    std::vector<MapRec> map_cache;
    // profiling
    bool prof_on = false;
    struct ProfRec { cudaEvent_t a, b; int kind; };
    std::vector<ProfRec> prof_recs;
    long long launches[LDPC_K_KINDS] = {0, 0, 0, 0, 0, 0, 0, 0};
};

// brackets one kernel launch with events when profiling is on; always counts the launch
struct ProfScope {
    ldpc_ctx *c; cudaStream_t st; int kind; cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(ldpc_ctx *c_, int kind_, cudaStream_t st_) : c(c_), st(st_), kind(kind_)
    {
        c->launches[kind]++;
        if (c->prof_on && cudaEventCreate(&a) == cudaSuccess && cudaEventCreate(&b) == cudaSuccess) cudaEventRecord(a, st);
    }
    ~ProfScope()
    {
        if (a && b) { cudaEventRecord(b, st); c->prof_recs.push_back({a, b, kind}); }
    }
};

// nfull == 0: the 3-D map; else the 4-D map over whole boxes
static int cached_map(ldpc_ctx *c, CUtensorMap *out, const void *base, int rows, long long B, int W, int nfull, bool is_load, int boxn)
{
    if (boxn <= 0) boxn = nfull;
    for (const auto &r : c->map_cache)
        if (r.key.base == base && r.key.rows == rows && r.key.B == B && r.key.W == W && r.key.nfull == nfull && r.key.is_load == int(is_load) &&
            r.key.boxn == boxn) {
            *out = r.map;
            return LDPC_OK;
        }
    ldpc_ctx::MapRec rec;
    rec.key = {base, rows, W, nfull, int(is_load), B, boxn};
    int rc = nfull ? make_map4(&rec.map, base, c->S, rows, B, W, nfull, is_load, boxn) : make_map(&rec.map, base, c->S, rows, B, W, is_load);
    if (rc) return rc;
    if (c->map_cache.size() >= 64) c->map_cache.erase(c->map_cache.begin());
    c->map_cache.push_back(rec);
    *out = rec.map;
    return LDPC_OK;
}

static const struct { const char *name; int n, k, rs_n, rs_k; } kBuiltin[] = {
    // OpenCL/device/LDPC_Vlist_data.h:10-14 (ldpc_params) + the .mat-only (4000,2000) code
    {"n2000_k1000", 2000, 1000, 250, 125},
    {"n2040_k1530", 2040, 1530, 255, 192},
    {"n4000_k2000", 4000, 2000, 250, 125},
};

static std::string codes_dir()
{
    if (const char *e = getenv("LDPC_CUDA_CODES_DIR")) return e;
    Dl_info info;
    if (dladdr(reinterpret_cast<void *>(&ldpc_cuda_abi_version), &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        size_t s = p.find_last_of('/');
        return (s == std::string::npos ? std::string(".") : p.substr(0, s)) + "/codes";
    }
    return "codes";
}

static int choose_geom(const ldpc_ctx *c, bool dynamic_sched, ExecGeom *g)
{
    const int n = c->code.n, S = c->S, m = c->code.m;
    const int nbox = (n + kBoxRows - 1) / kBoxRows;
    const int cidx_bytes = m * c->code.RW * 2;
    const int msk_words = ((nbox * kBoxRows + 8) / 32 + 1 + 3) & ~3;   // a codeword's mask + the (zero) bits of the zero rows
    const int mask_bytes = dynamic_sched ? kExecMaxGroups * msk_words * 4 : 0;
    const int fixed = cidx_bytes + mask_bytes + 64 + 128 + 128;   // check rows, masks, barriers, mailboxes, base alignment
    // a blob in shared memory: the static schedule whole; a per-codeword schedule with at least 64 records and, if
    // there is room, one per check (the peel kernel emits what fits, entries past that take the full-row form)
    ExecGeom best;
    for (int W = 64; W >= 16; W >>= 1) {         // instantiated slice widths; S is a multiple of 16, so 16 always divides
        if (W > S || S % W) continue;
        if (c->force_W && W != c->force_W) continue;
        const int slot = nbox * kBoxRows * W + kExecZeroRowBytes;
        if ((slot - kExecZeroRowBytes) / 16 > 0xFFFF) continue;   // check rows are staged as 16-bit offsets in 16-byte units
        // (a per-codeword blob must fit whole; the executor's pass table and records take what is left of the area, and a
        //  blob that leaves no room for its pass table is applied in the plain level-by-level form)
        const int blob_min = dynamic_sched ? sched_blob_max_bytes(m) : int(make_enc_blob(c->code, 32 / (W / 16)).size());
        const int blob_max = dynamic_sched ? sched_area_bytes(m, m) : blob_min;
        int nslot;
        if (dynamic_sched) nslot = (c->smem_optin - fixed) / (slot + blob_min);
        else nslot = (c->smem_optin - fixed - blob_min) / slot;
        nslot = std::min(nslot, kExecMaxGroups);
        if (c->force_slots) nslot = std::min(nslot, c->force_slots);
        if (nslot < 1) continue;
        int blob = blob_min;
        if (dynamic_sched) blob = std::min(blob_max, ((c->smem_optin - fixed - nslot * slot) / nslot) & ~15);
        ExecGeom cand;
        cand.W = W; cand.nslot = nslot; cand.slot_bytes = slot; cand.sched_area = blob; cand.msk_words = msk_words;
        cand.smem_bytes = nslot * slot + (dynamic_sched ? nslot : 1) * blob + cidx_bytes + mask_bytes + 64 + 128;
        if (!best.W || (best.nslot < 3 && cand.nslot > best.nslot)) best = cand;
        if (best.nslot >= 3) break;
    }
    if (!best.W) return fail(LDPC_ERR_UNSUPPORTED, "code too long for one shared-memory slot (n * 16 bytes must fit)");
    *g = best;
    return LDPC_OK;
}

// Check rows as the executor stages them (payload_exec.cuh, "Bank conflicts"): 16-bit offsets in 16-byte units from the
// slot base; the RW slots of a row form 128 / W groups, group a holding the members u with u mod (128 / W) == a (its
// bank class) as far as they fit, the rest in whatever slot is free; empty slots point at the zero row of their class
// behind the slot's boxes.
static int upload_exec_rows(ldpc_ctx *c, int W, uint16_t **d_rows)
{
    const HostCode &code = c->code;
    const int SL = code.RW, NCLS = 128 / W, GS = SL / NCLS, LPG = W / 16;
    const int zbase = ((code.n + kBoxRows - 1) / kBoxRows) * kBoxRows;
    std::vector<uint16_t> rows(size_t(code.m) * SL);
    std::vector<int> slot(static_cast<size_t>(SL));
    for (int r = 0; r < code.m; r++) {
        std::fill(slot.begin(), slot.end(), -1);
        std::vector<int> spill;
        for (int j = code.row_ptr[r]; j < code.row_ptr[r + 1]; j++) {
            const int u = code.col_idx[j], a = u % NCLS;
            int t = a * GS;
            while (t < (a + 1) * GS && slot[size_t(t)] >= 0) t++;
            if (t < (a + 1) * GS) slot[size_t(t)] = u; else spill.push_back(u);
        }
        for (int u : spill) {
            int t = 0;
            while (slot[size_t(t)] >= 0) t++;          // (a row has at most RW members)
            slot[size_t(t)] = u;
        }
        for (int t = 0; t < SL; t++)
            rows[size_t(r) * SL + t] = uint16_t((slot[size_t(t)] >= 0 ? slot[size_t(t)] : zbase + t / GS) * LPG);
    }
    cudaFree(*d_rows);
    *d_rows = nullptr;
    CUDA_TRY(cudaMalloc(d_rows, rows.size() * 2));
    CUDA_TRY(cudaMemcpy(*d_rows, rows.data(), rows.size() * 2, cudaMemcpyHostToDevice));
    return LDPC_OK;
}

// everything on the device that depends on the executor's geometry: the arranged check rows and the encoder's blob
static int upload_geometry_tables(ldpc_ctx *c)
{
    int rc = upload_exec_rows(c, c->dec.W, &c->d_rows_dec);
    if (!rc) rc = upload_exec_rows(c, c->enc.W, &c->d_rows_enc);
    if (rc) return rc;
    cudaFree(c->d_enc_blob);
    c->d_enc_blob = nullptr;
    if (c->code.triangular) {
        const std::vector<uint8_t> blob = make_enc_blob(c->code, 32 / (c->enc.W / 16));
        CUDA_TRY(cudaMalloc(&c->d_enc_blob, blob.size()));
        CUDA_TRY(cudaMemcpy(c->d_enc_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    }
    return LDPC_OK;
}

static void free_host_pipeline(ldpc_ctx *c);

static void free_ctx(ldpc_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaFree(c->d_cidx); cudaFree(c->d_vadj); cudaFree(c->d_rows_dec); cudaFree(c->d_rows_enc); cudaFree(c->d_enc_blob); cudaFree(c->d_sched);
    cudaFree(c->d_work_ctr);
    cudaFree(c->d_sched2); cudaFree(c->d_sched_len2); cudaFree(c->d_resid2); cudaFree(c->d_fail_scratch2); cudaFree(c->d_work_ctr2);
    for (int i = 0; i < 2; i++) {
        if (c->pstream[i]) cudaStreamDestroy(c->pstream[i]);
        if (c->ev_join[i]) cudaEventDestroy(c->ev_join[i]);
    }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    cudaFree(c->d_sched_len); cudaFree(c->d_resid); cudaFree(c->d_stats); cudaFree(c->d_fail_scratch); cudaFree(c->d_phase); cudaFree(c->d_sim_mask);
    hybrid_free(c->hyb);
    for (auto &r : c->prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    free_host_pipeline(c);
    delete c;
}

typedef void (*PeelKernel)(const PeelParams);

static PeelKernel pick_peel(int MW, int VW)
{
    const int bmw = peel_bitmap_words(MW);
    if (VW == 4) return bmw == 16 ? peel_schedule_kernel<16, 4> : (bmw == 32 ? peel_schedule_kernel<32, 4> : peel_schedule_kernel<64, 4>);
    if (VW == 8) return bmw == 16 ? peel_schedule_kernel<16, 8> : (bmw == 32 ? peel_schedule_kernel<32, 8> : peel_schedule_kernel<64, 8>);
    return nullptr;
}

static int setup_peel(ldpc_ctx *c)
{
    if (c->MW > 64 || (c->code.VW != 4 && c->code.VW != 8))
        return fail(LDPC_ERR_UNSUPPORTED, "peel kernel supports m <= 2048 and column weight <= 8");
    const int tables = ((c->code.n * c->code.VW + 7) & ~7) * 2 + c->code.m * c->code.RW * 2;
    const int per_group = peel_group_words(c->code.m, c->MW, c->NW) * 4;
    const int G = kPeelG;
    int groups = (c->smem_optin - 2048 - tables) / per_group;   // 2 KB margin: static + reserved shared memory
    groups = std::min(groups, 1024 / G);
    groups = (groups / (32 / G)) * (32 / G);
    if (groups < 32 / G) return fail(LDPC_ERR_UNSUPPORTED, "code too large for the peel kernel's shared memory");
    c->peel_G = G;
    c->peel_groups = groups;
    c->peel_smem = tables + groups * per_group;
    return allow_max_smem(reinterpret_cast<const void *>(pick_peel(c->MW, c->code.VW)), c->smem_optin);
}

extern "C" int ldpc_ctx_create(ldpc_ctx **out, const char *h_mat_path, int code_ind, int symbol_bytes, int device,
                               int64_t max_batch)
{
    if (!out) return fail(LDPC_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (symbol_bytes <= 0 || symbol_bytes % 16) return fail(LDPC_ERR_ARG, "symbol_bytes must be a positive multiple of 16");
    if (max_batch <= 0) return fail(LDPC_ERR_ARG, "max_batch must be positive");
    std::string path;
    int rs_n = 0, rs_k = 0;
    if (h_mat_path && *h_mat_path) {
        path = h_mat_path;
    } else {
        if (code_ind < 0 || code_ind >= int(sizeof(kBuiltin) / sizeof(kBuiltin[0])))
            return fail(LDPC_ERR_ARG, "code_ind out of range (0 = (2000,1000), 1 = (2040,1530), 2 = (4000,2000))");
        path = codes_dir() + "/" + kBuiltin[code_ind].name + ".mat";
        rs_n = kBuiltin[code_ind].rs_n;
        rs_k = kBuiltin[code_ind].rs_k;
    }
    int rows, cols;
    std::vector<int32_t> col_ptr, row_idx;
    std::string err;
    int rc = load_mat_sparse(path, "H_sparse", rows, cols, col_ptr, row_idx, err);
    if (rc) return fail(rc, err);
    ldpc_ctx *c = new (std::nothrow) ldpc_ctx();
    if (!c) return fail(LDPC_ERR_NOMEM, "out of host memory");
    rc = build_code(rows, cols, col_ptr, row_idx, c->code, err);
    if (rc) { delete c; return fail(rc, err); }
    if (!(h_mat_path && *h_mat_path) && (c->code.n != kBuiltin[code_ind].n || c->code.k != kBuiltin[code_ind].k)) {
        delete c;
        return fail(LDPC_ERR_FORMAT, path + ": dimensions do not match the built-in code table");
    }
    if (rs_n == 0) {  // RS-equivalent for a user code: blocks of <= 255 at the same rate (SURVEY a-2)
        for (int blk = 255; blk >= 16; blk--)
            if (c->code.n % blk == 0) { rs_n = blk; rs_k = int(std::ceil(double(blk) * c->code.k / c->code.n)); break; }
    }
    c->S = symbol_bytes; c->device = device; c->code_ind = (h_mat_path && *h_mat_path) ? -1 : code_ind;
    c->rs_n = rs_n; c->rs_k = rs_k; c->max_batch = max_batch;
    c->NW = (c->code.n + 31) / 32; c->MW = (c->code.m + 31) / 32;
    c->sched_stride = sched_blob_max_bytes(c->code.m);
    // hybrid mode keeps one syndrome set (m * S bytes) per codeword of a chunk: chunks are capped so that it stays <= 4 GiB
    c->hybrid_batch = std::min<long long>(max_batch, std::max<long long>(256, (4ll << 30) / (size_t(c->code.m) * symbol_bytes)));   // (never above max_batch: the schedule scratch is sized for it)

    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { delete c; return fail(LDPC_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e)); }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { delete c; return fail(LDPC_ERR_CUDA, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e)); }
    if (prop.major != 10) {
        delete c;
        return fail(LDPC_ERR_UNSUPPORTED, "libldpc_cuda is built for sm_100a (B200) only; device is sm_" +
                                              std::to_string(prop.major) + std::to_string(prop.minor));
    }
    c->num_sms = prop.multiProcessorCount;
    c->smem_optin = int(prop.sharedMemPerBlockOptin);

#define CTX_TRY(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            free_ctx(c);                                                                                \
            return fail(_e == cudaErrorMemoryAllocation ? LDPC_ERR_NOMEM : LDPC_ERR_CUDA,               \
                        std::string(#expr) + ": " + cudaGetErrorString(_e));                            \
        }                                                                                               \
    } while (0)
    CTX_TRY(cudaMalloc(&c->d_cidx, c->code.cidx.size() * 2));
    CTX_TRY(cudaMemcpy(c->d_cidx, c->code.cidx.data(), c->code.cidx.size() * 2, cudaMemcpyHostToDevice));
    CTX_TRY(cudaMalloc(&c->d_vadj, c->code.vadj.size() * 2));
    CTX_TRY(cudaMemcpy(c->d_vadj, c->code.vadj.data(), c->code.vadj.size() * 2, cudaMemcpyHostToDevice));
    CTX_TRY(cudaMalloc(&c->d_sched, size_t(max_batch) * c->sched_stride));
    CTX_TRY(cudaMalloc(&c->d_sched_len, size_t(max_batch) * 4));
    CTX_TRY(cudaMalloc(&c->d_resid, size_t(max_batch) * 4));
    CTX_TRY(cudaMalloc(&c->d_fail_scratch, size_t(max_batch)));
    CTX_TRY(cudaMalloc(&c->d_stats, 8 * sizeof(unsigned long long)));
    CTX_TRY(cudaMemset(c->d_stats, 0, 8 * sizeof(unsigned long long)));
    CTX_TRY(cudaMalloc(&c->d_work_ctr, sizeof(unsigned int)));
    if (const char *e = getenv("LDPC_CUDA_PHASE_TIMING")) {
        if (*e && *e != '0') {
            CTX_TRY(cudaMalloc(&c->d_phase, 24 * sizeof(unsigned long long)));
            CTX_TRY(cudaMemset(c->d_phase, 0, 24 * sizeof(unsigned long long)));
        }
    }
#undef CTX_TRY
    rc = choose_geom(c, true, &c->dec);
    if (!rc) rc = choose_geom(c, false, &c->enc);
    if (!rc) rc = setup_peel(c);
    if (!rc) rc = upload_geometry_tables(c);
    if (rc) { free_ctx(c); return rc; }
    *out = c;
    return LDPC_OK;
}

extern "C" int ldpc_read_h_file(const char *h_mat_path, int32_t dims[4], int32_t *row_ptr, int32_t *col_idx)
{
    if (!h_mat_path || !dims) return fail(LDPC_ERR_ARG, "NULL argument");
    int rows, cols;
    std::vector<int32_t> col_ptr, row_idx;
    std::string err;
    int rc = load_mat_sparse(h_mat_path, "H_sparse", rows, cols, col_ptr, row_idx, err);
    if (rc) return fail(rc, err);
    HostCode code;
    rc = build_code(rows, cols, col_ptr, row_idx, code, err);
    if (rc) return fail(rc, err);
    dims[0] = code.m; dims[1] = code.n; dims[2] = code.nnz; dims[3] = code.triangular ? 1 : 0;
    if (row_ptr) memcpy(row_ptr, code.row_ptr.data(), code.row_ptr.size() * 4);
    if (col_idx) memcpy(col_idx, code.col_idx.data(), code.col_idx.size() * 4);
    return LDPC_OK;
}

extern "C" int ldpc_ctx_destroy(ldpc_ctx *ctx)
{
    free_ctx(ctx);
    return LDPC_OK;
}

extern "C" int ldpc_ctx_info(const ldpc_ctx *c, ldpc_code_info *info)
{
    if (!c || !info) return fail(LDPC_ERR_ARG, "NULL argument");
    info->n = c->code.n; info->k = c->code.k; info->m = c->code.m; info->nnz = c->code.nnz;
    info->symbol_bytes = c->S; info->mask_words = c->NW; info->rs_n = c->rs_n; info->rs_k = c->rs_k;
    info->max_row_weight = c->code.max_row_weight; info->max_col_weight = c->code.max_col_weight;
    info->encode_levels = c->code.encode_levels; info->slice_bytes = c->dec.W; info->exec_slots = c->dec.nslot;
    info->device = c->device; info->max_batch = c->max_batch;
    return LDPC_OK;
}

extern "C" int ldpc_ctx_get_csr(const ldpc_ctx *c, int32_t *row_ptr, int32_t *col_idx)
{
    if (!c) return fail(LDPC_ERR_ARG, "NULL context");
    if (row_ptr) memcpy(row_ptr, c->code.row_ptr.data(), c->code.row_ptr.size() * 4);
    if (col_idx) memcpy(col_idx, c->code.col_idx.data(), c->code.col_idx.size() * 4);
    return LDPC_OK;
}

extern "C" int ldpc_ctx_set_exec_geometry(ldpc_ctx *c, int slice_bytes, int slots)
{
    if (!c) return fail(LDPC_ERR_ARG, "NULL context");
    if (slice_bytes && (slice_bytes % 16 || c->S % slice_bytes || slice_bytes > 64))
        return fail(LDPC_ERR_ARG, "slice_bytes must be 16, 32 or 64 and divide symbol_bytes");
    const int oldW = c->force_W, olds = c->force_slots;
    c->force_W = slice_bytes; c->force_slots = slots;
    ExecGeom d, e;
    int rc = choose_geom(c, true, &d);
    if (!rc) rc = choose_geom(c, false, &e);
    if (rc) { c->force_W = oldW; c->force_slots = olds; return rc; }
    c->dec = d; c->enc = e;
    CUDA_TRY(cudaSetDevice(c->device));
    return upload_geometry_tables(c);
}

// ------------------------------------------------------------------------------------------
// executor launch
// ------------------------------------------------------------------------------------------
typedef void (*ExecKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                           const ExecParams);

// instantiated member counts: 7 / 14 are the committed codes' maximum row weights
static ExecKernel pick_exec(int W, int max_row_weight, int *rwm_out)
{
    static const int kRwm[] = {7, 8, 14, 16, 24, 32};
    int rwm = 0;
    for (int r : kRwm) if (r >= max_row_weight) { rwm = r; break; }
    *rwm_out = rwm;
#define PICK(WW)                                                      \
    switch (rwm) {                                                    \
        case 7: return payload_exec_kernel<WW, 7>;                    \
        case 8: return payload_exec_kernel<WW, 8>;                    \
        case 14: return payload_exec_kernel<WW, 14>;                  \
        case 16: return payload_exec_kernel<WW, 16>;                  \
        case 24: return payload_exec_kernel<WW, 24>;                  \
        case 32: return payload_exec_kernel<WW, 32>;                  \
        default: return nullptr;                                      \
    }
    if (W == 16) { PICK(16) }
    if (W == 32) { PICK(32) }
    if (W == 64) { PICK(64) }
#undef PICK
    return nullptr;
}

static int launch_exec(ldpc_ctx *c, const ExecGeom &g, const void *d_in, int rows_in, void *d_out, int rows_out,
                       const uint8_t *sched, const uint32_t *sched_len, int sched_stride, long long B,
                       cudaStream_t st, uint8_t *d_synd = nullptr, const uint32_t *d_mask = nullptr)
{
    if (B <= 0) return LDPC_OK;
    CUtensorMap in_map, out_map, in4_map, out4_map;
    int rc = cached_map(c, &in_map, d_in, rows_in, B, g.W, 0, true);
    if (rc) return rc;
    rc = cached_map(c, &out_map, d_out, rows_out, B, g.W, 0, false);
    if (rc) return rc;
    const int nfull_in = g_tma4 ? rows_in / kBoxRows : 0, nfull_out = g_tma4 ? rows_out / kBoxRows : 0;
    in4_map = in_map; out4_map = out_map;      // (placeholders when there is no whole box)
    if (nfull_in) { rc = cached_map(c, &in4_map, d_in, rows_in, B, g.W, nfull_in, true); if (rc) return rc; }
    if (nfull_out) { rc = cached_map(c, &out4_map, d_out, rows_out, B, g.W, nfull_out, false); if (rc) return rc; }
    // The encoder never changes its information rows: the whole boxes below row k are stored as soon as they are loaded (outA),
    // the boxes that hold parity rows after the walk (outB and the partial box) -- no second store of the walk's rows.
    CUtensorMap outA_map = out4_map, outB_map = out4_map;
    int early = (!sched_stride && nfull_out && exec_split_store()) ? std::min(rows_in / kBoxRows, nfull_out) : 0;
    if (early) {
        rc = cached_map(c, &outA_map, d_out, rows_out, B, g.W, nfull_out, false, early);
        if (rc) return rc;
        if (nfull_out > early) { rc = cached_map(c, &outB_map, d_out, rows_out, B, g.W, nfull_out, false, nfull_out - early); if (rc) return rc; }
    }
    ExecParams p;
    p.early_boxes = early;
    p.synd = d_synd; p.mask = d_mask; p.NW = c->NW; p.nfull_in = nfull_in; p.nfull_out = nfull_out;
    p.rows = sched_stride ? c->d_rows_dec : c->d_rows_enc; p.sched = sched; p.sched_len = sched_len; p.B = B; p.sched_stride = sched_stride;
    p.sched_max = g.sched_area; p.m = c->code.m; p.RW = c->code.RW; p.rows_in = rows_in; p.rows_out = rows_out;
    p.nbox_in = (rows_in + kBoxRows - 1) / kBoxRows; p.nbox_out = (rows_out + kBoxRows - 1) / kBoxRows;
    p.slices = c->S / g.W; p.nslot = g.nslot; p.slot_bytes = g.slot_bytes;
    p.zrow = sched_zero_row(c->code.n); p.out = static_cast<uint8_t *>(d_out); p.S = c->S; p.n = c->code.n; p.msk_words = g.msk_words; p.force_plain = exec_plain() ? 1 : 0;
    p.phase_cycles = sched_stride ? c->d_phase : nullptr;
    int rwm = 0;
    ExecKernel k = pick_exec(g.W, c->code.max_row_weight, &rwm);
    if (!k || (rwm + 7) / 8 * 8 != c->code.RW) return fail(LDPC_ERR_UNSUPPORTED, "no executor instantiation for this slice width / row weight");
    rc = allow_max_smem(reinterpret_cast<const void *>(k), c->smem_optin);
    if (rc) return rc;
    const int grid = int(std::min<long long>(c->num_sms, B));
    {
        ProfScope ps(c, sched_stride ? LDPC_K_EXEC_DECODE : LDPC_K_EXEC_ENCODE, st);
        k<<<grid, g.nslot * kExecWarpsPerGroup * 32, g.smem_bytes, st>>>(in_map, out_map, in4_map, out4_map, outA_map, outB_map, p);
    }
    CUDA_TRY(cudaGetLastError());
    return debug_sync(sched_stride ? "payload_exec_kernel(decode)" : "payload_exec_kernel(encode)", st);
}

// ------------------------------------------------------------------------------------------
// encoder
// ------------------------------------------------------------------------------------------
extern "C" int ldpc_encode(ldpc_ctx *c, const void *d_info, void *d_cw, int64_t B, void *stream)
{
    if (!c || B < 0) return fail(LDPC_ERR_ARG, "bad argument to ldpc_encode");
    if (B == 0) return LDPC_OK;
    if (!d_info || !d_cw) return fail(LDPC_ERR_ARG, "NULL buffer passed to ldpc_encode");
    if (!c->code.triangular)
        return fail(LDPC_ERR_NOT_TRIANGULAR, "H is not in triangular form (last entry of row r must be column k + r)");
    CUDA_TRY(cudaSetDevice(c->device));
    // the tensor map's batch dimension is 32-bit friendly; split very large batches
    const long long step = 1ll << 22;
    for (long long b0 = 0; b0 < B; b0 += step) {
        const long long nb = std::min<long long>(step, B - b0);
        int rc = launch_exec(c, c->enc, static_cast<const uint8_t *>(d_info) + size_t(b0) * c->code.k * c->S, c->code.k,
                             static_cast<uint8_t *>(d_cw) + size_t(b0) * c->code.n * c->S, c->code.n, c->d_enc_blob,
                             nullptr, 0, nb, static_cast<cudaStream_t>(stream));
        if (rc) return rc;
    }
    return LDPC_OK;
}

// ------------------------------------------------------------------------------------------
// erasure channel
// ------------------------------------------------------------------------------------------
static void prob_threshold(double prob, uint32_t *t, int *always, int *never)
{
    *always = prob >= 1.0;
    *never = prob < 0.0;
    *t = 0;
    if (!*always && !*never) *t = uint32_t(std::floor(prob * 4294967296.0));  // v <= t  <=>  v / 2^32 <= prob
}

extern "C" int ldpc_gen_erasures(ldpc_ctx *c, const ldpc_erasure_model *model, uint32_t seed, uint64_t frame0,
                                 int64_t B, uint32_t *d_mask, void *d_payload, void *stream)
{
    if (!c || !model || !d_mask || B < 0) return fail(LDPC_ERR_ARG, "bad argument to ldpc_gen_erasures");
    if (B == 0) return LDPC_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    GenParams p;
    memset(&p, 0, sizeof(p));
    p.mask = d_mask; p.B = B; p.frame0 = frame0; p.seed = seed; p.n = c->code.n; p.NW = c->NW; p.model = model->model;
    const int threads = 256;
    if (model->model == LDPC_ERASURE_IID64 || model->model == LDPC_ERASURE_IID32) {
        if (model->model == LDPC_ERASURE_IID64) {
            if (model->per_numerator_div_64 < 0 || model->per_numerator_div_64 > 64)
                return fail(LDPC_ERR_ARG, "per_numerator_div_64 must be in 0..64");
            p.thr = uint32_t(model->per_numerator_div_64);
        } else {
            p.thr = model->threshold32;
        }
        const long long warps = B * c->NW;
        const int grid = int(std::min<long long>((warps + 7) / 8, (long long)c->num_sms * 32));
        ProfScope ps(c, LDPC_K_CHANNEL, st);
        gen_erasures_iid_kernel<<<grid, threads, 0, st>>>(p);
    } else if (model->model == LDPC_ERASURE_BURSTY) {
        if (!(model->bias > 0.0)) return fail(LDPC_ERR_ARG, "bursty model needs bias > 0");
        const double p01 = 0.1 / model->bias, p10 = 0.1;  // Bursty_Error_Channel_Model_Generator.m:14-17
        prob_threshold(model->alpha, &p.t_alpha, &p.a_alpha, &p.n_alpha);
        prob_threshold(model->beta, &p.t_beta, &p.a_beta, &p.n_beta);
        prob_threshold(p01, &p.t_p01, &p.a_p01, &p.n_p01);
        prob_threshold(p10, &p.t_p10, &p.a_p10, &p.n_p10);
        if (p.t_p01 == p.t_p10 && p.a_p01 == p.a_p10)
            return fail(LDPC_ERR_UNSUPPORTED, "bursty model with bias == 1 has no resynchronising symbols");
        const int grid = int(std::min<long long>((B + 7) / 8, (long long)c->num_sms * 32));
        ProfScope ps(c, LDPC_K_CHANNEL, st);
        gen_erasures_bursty_kernel<<<grid, threads, 0, st>>>(p);
    } else {
        return fail(LDPC_ERR_ARG, "unknown erasure model");
    }
    CUDA_TRY(cudaGetLastError());
    if (d_payload) {
        const long long warps = B * c->NW;
        const int grid = int(std::min<long long>((warps + 7) / 8, (long long)c->num_sms * 32));
        ProfScope ps(c, LDPC_K_CHANNEL, st);
        zero_erased_kernel<<<grid, threads, 0, st>>>(d_mask, static_cast<uint8_t *>(d_payload), B, c->code.n, c->NW, c->S);
        CUDA_TRY(cudaGetLastError());
    }
    return LDPC_OK;
}

extern "C" int ldpc_fill_random(void *d_dst, int64_t nbytes, uint32_t seed, uint64_t block0, int device, void *stream)
{
    if (!d_dst || nbytes < 0 || nbytes % 16) return fail(LDPC_ERR_ARG, "ldpc_fill_random: nbytes must be a multiple of 16");
    if (nbytes == 0) return LDPC_OK;
    CUDA_TRY(cudaSetDevice(device));
    const long long nblocks = nbytes / 16;
    const int grid = int(std::min<long long>((nblocks + 255) / 256, 148 * 16));
    fill_random_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<uint4 *>(d_dst), nblocks, seed, block0);
    CUDA_TRY(cudaGetLastError());
    return LDPC_OK;
}

// ------------------------------------------------------------------------------------------
// FEC packet front-ends
// ------------------------------------------------------------------------------------------
static int packetize_impl(ldpc_ctx *c, const void *d_cw, const uint16_t *d_len8, uint32_t block0, int64_t B, void *d_packets, void *stream)
{
    if (!c || B < 0) return fail(LDPC_ERR_ARG, "bad argument to ldpc_packetize");
    if (B == 0) return LDPC_OK;
    if (!d_cw || !d_packets) return fail(LDPC_ERR_ARG, "NULL buffer passed to ldpc_packetize");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long np = (long long)B * c->code.n;
    const int words = c->S / 8;
    const int lpp = words >= 32 ? 32 : (words >= 8 ? 8 : (words >= 4 ? 4 : 2));
    const int grid = int(std::min<long long>((np * lpp + 255) / 256, (long long)c->num_sms * 32));
    ProfScope ps(c, LDPC_K_CHANNEL, st);
    auto go = [&](auto kern) {
        kern<<<grid, 256, 0, st>>>(static_cast<const unsigned long long *>(d_cw), static_cast<unsigned long long *>(d_packets), np, c->code.n,
                                   words, block0, d_len8);
    };
    if (lpp == 32) go(packetize_kernel<32>);
    else if (lpp == 8) go(packetize_kernel<8>);
    else if (lpp == 4) go(packetize_kernel<4>);
    else go(packetize_kernel<2>);
    CUDA_TRY(cudaGetLastError());
    return LDPC_OK;
}

extern "C" int ldpc_packetize(ldpc_ctx *c, const void *d_cw, uint32_t block0, int64_t B, void *d_packets, void *stream)
{
    return packetize_impl(c, d_cw, nullptr, block0, B, d_packets, stream);
}

extern "C" int ldpc_packetize_var(ldpc_ctx *c, const void *d_cw, const uint16_t *d_len8, uint32_t block0, int64_t B, void *d_packets,
                                  void *stream)
{
    if (!d_len8) return fail(LDPC_ERR_ARG, "ldpc_packetize_var: d_len8 is NULL");
    return packetize_impl(c, d_cw, d_len8, block0, B, d_packets, stream);
}

// append = the block buffers already hold a partly assembled window: no zero fill, masks and counts keep their state
static int depacketize_impl(ldpc_ctx *c, const void *d_packets, const uint16_t *d_len8, int64_t n_packets, uint32_t block0, int64_t B,
                            void *d_cw, uint32_t *d_mask, uint32_t *d_counts, bool append, void *stream)
{
    if (!c || B < 0 || n_packets < 0) return fail(LDPC_ERR_ARG, "bad argument to ldpc_depacketize");
    if (B > 256) return fail(LDPC_ERR_ARG, "ldpc_depacketize: block numbers are modulo 256, a window holds at most 256 blocks");
    if (B == 0) return LDPC_OK;
    if (!d_cw || !d_mask || !d_counts || (n_packets > 0 && !d_packets)) return fail(LDPC_ERR_ARG, "NULL buffer passed to ldpc_depacketize");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int words = c->S / 8;
    if (!append) {
        CUDA_TRY(cudaMemsetAsync(d_cw, 0, size_t(B) * c->code.n * c->S, st));       // missing symbols are all-zero (receiver :59-68)
        ProfScope ps(c, LDPC_K_CHANNEL, st);
        const int grid = int(std::min<long long>((B * c->NW + 255) / 256, (long long)c->num_sms * 8));
        fec_mask_init_kernel<<<grid, 256, 0, st>>>(d_mask, B, c->code.n, c->NW, d_counts);
        CUDA_TRY(cudaGetLastError());
    }
    if (n_packets > 0) {
        ProfScope ps(c, LDPC_K_CHANNEL, st);
        const int lpp = words >= 32 ? 32 : (words >= 8 ? 8 : (words >= 4 ? 4 : 2));
        const int grid = int(std::min<long long>((n_packets * lpp + 255) / 256, (long long)c->num_sms * 32));
        auto args = [&](auto kern) {
            kern<<<grid, 256, 0, st>>>(static_cast<const unsigned long long *>(d_packets), n_packets, static_cast<unsigned long long *>(d_cw),
                                       d_mask, d_counts, B, c->code.n, c->NW, words, block0, d_len8);
        };
        if (lpp == 32) args(depacketize_kernel<32>);
        else if (lpp == 8) args(depacketize_kernel<8>);
        else if (lpp == 4) args(depacketize_kernel<4>);
        else args(depacketize_kernel<2>);
        CUDA_TRY(cudaGetLastError());
    }
    return debug_sync("depacketize_kernel", st);
}

extern "C" int ldpc_depacketize(ldpc_ctx *c, const void *d_packets, int64_t n_packets, uint32_t block0, int64_t B,
                                void *d_cw, uint32_t *d_mask, uint32_t *d_counts, void *stream)
{
    return depacketize_impl(c, d_packets, nullptr, n_packets, block0, B, d_cw, d_mask, d_counts, false, stream);
}

extern "C" int ldpc_depacketize_var(ldpc_ctx *c, const void *d_packets, const uint16_t *d_len8, int64_t n_packets, uint32_t block0, int64_t B,
                                    void *d_cw, uint32_t *d_mask, uint32_t *d_counts, void *stream)
{
    if (n_packets > 0 && !d_len8) return fail(LDPC_ERR_ARG, "ldpc_depacketize_var: d_len8 is NULL");
    return depacketize_impl(c, d_packets, d_len8, n_packets, block0, B, d_cw, d_mask, d_counts, false, stream);
}

extern "C" int ldpc_ready_to_decode(const ldpc_ctx *c, int cur_block_cnt, int next_block_cnt)
{
    if (!c) return 0;
    const int n = c->code.n, k = c->code.k, m = n - k;
    const int desired = int(double(m) * 0.8 + 0.5), minimum = int(double(m) * 0.2 + 0.5);
    return (cur_block_cnt == n) || (cur_block_cnt > k + desired && next_block_cnt > 10) || (cur_block_cnt > k + minimum && next_block_cnt > 100);
}

// ------------------------------------------------------------------------------------------
// receiver with two block buffers (the rest of it): the streaming state machine of
// OpenCL/device/ldpc_erasure_decoder_with_reordering_logic.cl:45-142 -- two block buffers {current, next}, packets
// of other blocks dropped (:105,124), arrival counters (:112-131), the hand-off rule evaluated after every packet
// (:139) -- as a host object over the packet kernel and the decoder.  The headers of a pushed batch are read back,
// the per-packet control runs on the host exactly as the reference's loop does, and the payload goes block by block
// through depacketize (appending to the window) and ldpc_decode on the GPU.  What the committed sketch leaves open
// (it stops compiling after the rule) is completed in the obvious way: the decoded block is emitted, `next` becomes
// `current` (block numbers modulo 256), its counter is kept, and the freed buffer is cleared for the new `next`.
// ------------------------------------------------------------------------------------------
struct ldpc_rx_stream {
    ldpc_ctx *c = nullptr;
    int max_iter = 50, mode = LDPC_MODE_PEEL;
    int cur = -1, next = -1;
    long long cur_cnt = 0, next_cnt = 0;
    uint8_t *d_win = nullptr;       // [2][n][S]
    uint32_t *d_mask = nullptr;     // [2][NW]
    uint32_t *d_counts = nullptr;   // [3]
    std::vector<unsigned long long> hdr;
};

extern "C" int ldpc_decode_ex(ldpc_ctx *c, const void *d_cw, const uint32_t *d_mask, void *d_out, uint8_t *d_fail,
                              uint8_t *d_fail_any, int max_iter, int mode, int64_t B, void *stream);

extern "C" int ldpc_rx_stream_create(ldpc_rx_stream **out, ldpc_ctx *c, int max_iter, int mode)
{
    if (!out || !c) return fail(LDPC_ERR_ARG, "NULL argument to ldpc_rx_stream_create");
    *out = nullptr;
    if (mode != LDPC_MODE_PEEL && mode != LDPC_MODE_HYBRID) return fail(LDPC_ERR_ARG, "unknown decode mode");
    ldpc_rx_stream *s = new (std::nothrow) ldpc_rx_stream();
    if (!s) return fail(LDPC_ERR_NOMEM, "out of host memory");
    s->c = c; s->max_iter = max_iter; s->mode = mode;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaError_t e = cudaMalloc(&s->d_win, size_t(2) * c->code.n * c->S);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_mask, size_t(2) * c->NW * 4);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_counts, 3 * 4);
    if (e != cudaSuccess) {
        cudaFree(s->d_win); cudaFree(s->d_mask); cudaFree(s->d_counts);
        delete s;
        return fail(LDPC_ERR_NOMEM, std::string("ldpc_rx_stream_create: ") + cudaGetErrorString(e));
    }
    *out = s;
    return LDPC_OK;
}

extern "C" int ldpc_rx_stream_destroy(ldpc_rx_stream *s)
{
    if (!s) return LDPC_OK;
    cudaSetDevice(s->c->device);
    cudaFree(s->d_win); cudaFree(s->d_mask); cudaFree(s->d_counts);
    delete s;
    return LDPC_OK;
}

extern "C" int ldpc_rx_stream_state(const ldpc_rx_stream *s, int32_t state[4])
{
    if (!s || !state) return fail(LDPC_ERR_ARG, "NULL argument");
    state[0] = s->cur; state[1] = s->next; state[2] = int32_t(s->cur_cnt); state[3] = int32_t(s->next_cnt);
    return LDPC_OK;
}

// decode the current block as it stands, emit it, make `next` the current block
static int rx_stream_hand_off(ldpc_rx_stream *s, void *d_out, uint8_t *d_fail, int32_t *h_blocks, int cap, int *n_dec, cudaStream_t st)
{
    ldpc_ctx *c = s->c;
    if (*n_dec >= cap) return fail(LDPC_ERR_ARG, "ldpc_rx_stream: more blocks became ready than the output holds (cap)");
    const size_t blk = size_t(c->code.n) * c->S, outb = size_t(c->code.k) * c->S;
    int rc = ldpc_decode_ex(c, s->d_win, s->d_mask, static_cast<uint8_t *>(d_out) + size_t(*n_dec) * outb, d_fail ? d_fail + *n_dec : nullptr,
                            nullptr, s->max_iter, s->mode, 1, st);
    if (rc) return rc;
    if (h_blocks) h_blocks[*n_dec] = s->cur;
    (*n_dec)++;
    CUDA_TRY(cudaMemcpyAsync(s->d_win, s->d_win + blk, blk, cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(s->d_mask, s->d_mask + c->NW, size_t(c->NW) * 4, cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemsetAsync(s->d_win + blk, 0, blk, st));
    fec_mask_init_kernel<<<1, 256, 0, st>>>(s->d_mask + c->NW, 1, c->code.n, c->NW, s->d_counts + 1);
    CUDA_TRY(cudaGetLastError());
    s->cur = s->next;
    s->next = (s->cur + 1) & 0xff;
    s->cur_cnt = s->next_cnt;
    s->next_cnt = 0;
    return LDPC_OK;
}

extern "C" int ldpc_rx_stream_push(ldpc_rx_stream *s, const void *d_packets, const uint16_t *d_len8, int64_t n_packets, void *d_out,
                                   uint8_t *d_fail, int32_t *h_blocks, int cap, int *n_decoded, void *stream)
{
    if (!s || n_packets < 0 || !n_decoded || (n_packets > 0 && !d_packets) || !d_out) return fail(LDPC_ERR_ARG, "bad argument to ldpc_rx_stream_push");
    *n_decoded = 0;
    if (n_packets == 0) return LDPC_OK;
    ldpc_ctx *c = s->c;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t ps = 8 + size_t(c->S);
    s->hdr.resize(size_t(n_packets));
    CUDA_TRY(cudaMemcpy2DAsync(s->hdr.data(), 8, d_packets, ps, 8, size_t(n_packets), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const uint8_t *pk = static_cast<const uint8_t *>(d_packets);
    int64_t seg0 = 0;
    auto place = [&](int64_t upto) -> int {      // packets [seg0, upto) into the window {cur, next}
        if (upto <= seg0 || s->cur < 0) { seg0 = upto; return LDPC_OK; }
        int rc = depacketize_impl(c, pk + size_t(seg0) * ps, d_len8 ? d_len8 + seg0 : nullptr, upto - seg0, uint32_t(s->cur), 2, s->d_win, s->d_mask,
                                  s->d_counts, true, st);
        seg0 = upto;
        return rc;
    };
    for (int64_t i = 0; i < n_packets; i++) {
        const unsigned long long h = s->hdr[size_t(i)];
        const uint32_t lo = uint32_t(h), hi = uint32_t(h >> 32);
        const bool valid = lo == hi && ((lo >> 24) & 0xffu) == kFecClass && int(lo & 0xffffu) < c->code.n;
        const int blk = int((lo >> 16) & 0xffu);
        if (s->cur < 0) {                       // the first packet names the current block (:88-91); an unusable header does not
            if (!valid) continue;
            s->cur = blk; s->next = (blk + 1) & 0xff;
            seg0 = i;
            CUDA_TRY(cudaMemsetAsync(s->d_win, 0, size_t(2) * c->code.n * c->S, st));
            fec_mask_init_kernel<<<1, 256, 0, st>>>(s->d_mask, 2, c->code.n, c->NW, s->d_counts);
            CUDA_TRY(cudaGetLastError());
        }
        if (valid && blk == s->cur) s->cur_cnt++;
        else if (valid && blk == s->next) s->next_cnt++;
        if (ldpc_ready_to_decode(c, int(s->cur_cnt), int(s->next_cnt))) {      // (:139), after every packet
            int rc = place(i + 1);
            if (!rc) rc = rx_stream_hand_off(s, d_out, d_fail, h_blocks, cap, n_decoded, st);
            if (rc) return rc;
        }
    }
    return place(n_packets);
}

extern "C" int ldpc_rx_stream_flush(ldpc_rx_stream *s, void *d_out, uint8_t *d_fail, int32_t *h_blocks, int cap, int *n_decoded, void *stream)
{
    if (!s || !n_decoded || !d_out) return fail(LDPC_ERR_ARG, "bad argument to ldpc_rx_stream_flush");
    *n_decoded = 0;
    CUDA_TRY(cudaSetDevice(s->c->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    for (int round = 0; round < 2 && s->cur >= 0 && s->cur_cnt > 0; round++) {
        int rc = rx_stream_hand_off(s, d_out, d_fail, h_blocks, cap, n_decoded, st);
        if (rc) return rc;
    }
    return LDPC_OK;
}

// ------------------------------------------------------------------------------------------
// decoder
// ------------------------------------------------------------------------------------------
// LDPC_CUDA_CHUNK_PIPELINE=0: chunks of a large batch run one after the other on the caller's stream
static bool chunk_pipeline() { const char *e = getenv("LDPC_CUDA_CHUNK_PIPELINE"); return !(e && *e == '0'); }

static int ensure_chunk_pipeline(ldpc_ctx *c)
{
    if (c->pstream[0]) return LDPC_OK;      // (set last: everything below exists)
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t r) { if (e == cudaSuccess) e = r; return e == cudaSuccess; };
    cudaStream_t s1 = nullptr, s0 = nullptr;
    if (ok(cudaMalloc(&c->d_sched2, size_t(c->max_batch) * c->sched_stride)) && ok(cudaMalloc(&c->d_sched_len2, size_t(c->max_batch) * 4)) &&
        ok(cudaMalloc(&c->d_resid2, size_t(c->max_batch) * 4)) && ok(cudaMalloc(&c->d_fail_scratch2, size_t(c->max_batch))) &&
        ok(cudaMalloc(&c->d_work_ctr2, sizeof(unsigned int))) && ok(cudaEventCreateWithFlags(&c->ev_join[0], cudaEventDisableTiming)) &&
        ok(cudaEventCreateWithFlags(&c->ev_join[1], cudaEventDisableTiming)) && ok(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming)) &&
        ok(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking)) && ok(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking))) {
        c->pstream[1] = s1;
        c->pstream[0] = s0;
        return LDPC_OK;
    }
    // a failed set-up leaves nothing behind: the next call starts clean
    cudaFree(c->d_sched2); cudaFree(c->d_sched_len2); cudaFree(c->d_resid2); cudaFree(c->d_fail_scratch2); cudaFree(c->d_work_ctr2);
    c->d_sched2 = nullptr; c->d_sched_len2 = nullptr; c->d_resid2 = nullptr; c->d_fail_scratch2 = nullptr; c->d_work_ctr2 = nullptr;
    for (int i = 0; i < 2; i++) if (c->ev_join[i]) { cudaEventDestroy(c->ev_join[i]); c->ev_join[i] = nullptr; }
    if (c->ev_fork) { cudaEventDestroy(c->ev_fork); c->ev_fork = nullptr; }
    if (s1) cudaStreamDestroy(s1);
    if (s0) cudaStreamDestroy(s0);
    return fail(e == cudaErrorMemoryAllocation ? LDPC_ERR_NOMEM : LDPC_ERR_CUDA, std::string("chunk pipeline set-up: ") + cudaGetErrorString(e));
}

// One peel -> executor (-> elimination) pass over B <= max_batch codewords.  (Cutting a batch into pieces
// on two internal streams so that one piece's executor overlaps the next one's schedule compiler was
// measured and is no faster: both kernels fill an SM's shared memory and cannot share one.)
static int decode_chunk(ldpc_ctx *c, const uint8_t *d_cw, const uint32_t *d_mask, uint8_t *d_out, uint8_t *d_fail, uint8_t *d_fail_any,
                        int max_iter, int mode, long long B, cudaStream_t st, int set = 0)
{
    uint8_t *const sched = set ? c->d_sched2 : c->d_sched;
    uint32_t *const sched_len = set ? c->d_sched_len2 : c->d_sched_len;
    PeelParams pp;
    pp.mask = d_mask; pp.sched = sched; pp.sched_len = sched_len;
    pp.work_ctr = set ? c->d_work_ctr2 : c->d_work_ctr;
    pp.fail = d_fail ? d_fail : (set ? c->d_fail_scratch2 : c->d_fail_scratch); pp.resid = set ? c->d_resid2 : c->d_resid; pp.stats = c->d_stats;
    pp.fail_any = d_fail_any;
    pp.cidx = c->d_cidx; pp.vadj = c->d_vadj; pp.B = B; pp.n = c->code.n; pp.k = c->code.k; pp.m = c->code.m;
    pp.RW = c->code.RW; pp.VW = c->code.VW; pp.NW = c->NW; pp.MW = c->MW; pp.stride = c->sched_stride;
    pp.max_iter = max_iter; pp.rs_n = c->rs_n; pp.rs_k = c->rs_k; pp.groups_per_block = c->peel_groups;
    pp.count_stats = 1;
    pp.ge_list = nullptr; pp.ge_count = nullptr;
    if (mode == LDPC_MODE_HYBRID) {
        int rch = hybrid_prepare(c->hyb, c->code, c->S, c->NW, c->MW, c->num_sms, c->smem_optin, c->hybrid_batch, g_err);
        if (rch) return rch;
        CUDA_TRY(cudaMemsetAsync(c->hyb.d_count, 0, 16 * 4, st));
        pp.ge_list = c->hyb.d_list; pp.ge_count = c->hyb.d_count;
    }
    const int grid = int(std::min<long long>(c->num_sms, (B + c->peel_groups - 1) / c->peel_groups));
    CUDA_TRY(cudaMemsetAsync(pp.work_ctr, 0, sizeof(unsigned int), st));
    {
        ProfScope ps(c, LDPC_K_PEEL, st);
        pick_peel(c->MW, c->code.VW)<<<grid, c->peel_groups * c->peel_G, c->peel_smem, st>>>(pp);
    }
    CUDA_TRY(cudaGetLastError());
    { int rcd = debug_sync("peel_schedule_kernel", st); if (rcd) return rcd; }
    const bool pattern_only = d_cw == nullptr;   // error-rate run: no payload
    if (!pattern_only) {
        int rc = launch_exec(c, c->dec, d_cw, c->code.n, d_out, c->code.k, sched, sched_len, c->sched_stride, B, st,
                             mode == LDPC_MODE_HYBRID ? c->hyb.d_synd : nullptr, d_mask);
        if (rc) return rc;
    }
    if (mode == LDPC_MODE_HYBRID) {
        GeParams gp;
        gp.mask = d_mask; gp.sched = sched; gp.list = c->hyb.d_list; gp.list_count = c->hyb.d_count;
        gp.synd = pattern_only ? nullptr : c->hyb.d_synd; gp.out = d_out; gp.fail = pp.fail; gp.fail_any = d_fail_any; gp.stats = c->d_stats; gp.cidx = c->d_cidx;
        gp.vadj = c->d_vadj; gp.VW = c->code.VW; gp.phase_cycles = c->d_phase ? c->d_phase + 8 : nullptr; gp.gmat = c->hyb.d_gmat; gp.n = c->code.n; gp.k = c->code.k; gp.m = c->code.m; gp.RW = c->code.RW; gp.NW = c->NW;
        gp.MW = c->MW; gp.S = c->S; gp.stride = c->sched_stride; gp.RSW = c->hyb.RSW;
        // warp-per-codeword stages first (inactivation decoding with typical, then worst-case slots; then plain
        // elimination per warp); what they defer falls through to the CTA-per-codeword kernel
        const unsigned int *lists[4] = {c->hyb.d_list, c->hyb.d_list2, c->hyb.d_list3, c->hyb.d_list4};
        int li = 0;
        const int stage_mask = ge_stage_mask();
        for (int sg = 0; sg < 3; sg++) {
            if (c->hyb.wpc[sg] <= 0 || !((stage_mask >> (sg == 2 ? 1 : 0)) & 1)) continue;
            GeWarpParams wp;
            wp.g = gp; wp.g.list = lists[li]; wp.g.list_count = c->hyb.d_count + li;
            wp.list_out = const_cast<unsigned int *>(lists[li + 1]); wp.count_out = c->hyb.d_count + li + 1;
            int wpc = c->hyb.wpc[sg], slot = c->hyb.slot_words[sg];
            wp.plan = nullptr; wp.plan_count = nullptr; wp.plan_words = 0; wp.apply_slot_words = 0;
            const bool split = sg < 2 && !pattern_only && c->hyb.d_plan && c->hyb.wpc_pat[sg] > 0 && ge_split();
            if (sg < 2 && pattern_only && c->hyb.wpc_pat[sg] > 0) { wpc = c->hyb.wpc_pat[sg]; slot = c->hyb.slot_words_pat[sg]; }
            if (split) {   // pattern part at full occupancy, recording plans; the payload replay follows
                wp.plan = c->hyb.d_plan; wp.plan_count = c->hyb.d_count + 8 + sg; wp.plan_words = c->hyb.plan_words;
                wp.apply_slot_words = c->hyb.slot_words[sg];
                wpc = c->hyb.wpc_pat[sg]; slot = c->hyb.slot_words_pat[sg];
            }
            wp.work_ctr = c->hyb.d_count + 4 + sg; wp.slot_words = slot;
            const size_t smem = size_t(wpc) * slot * 4;
            {
                ProfScope ps(c, sg < 2 ? LDPC_K_HYBRID : LDPC_K_HYBRID_WARP, st);
                if (sg < 2) hybrid_inact_kernel<<<c->num_sms, 32 * wpc, smem, st>>>(wp);
                else hybrid_ge_warp_kernel<<<c->num_sms, 32 * wpc, smem, st>>>(wp);
            }
            if (split) {
                GeWarpParams ap = wp;
                if (ap.g.phase_cycles) ap.g.phase_cycles += 8;
                ap.work_ctr = c->hyb.d_count + 10 + sg; ap.slot_words = c->hyb.slot_words[sg];
                ProfScope ps(c, LDPC_K_HYBRID_APPLY, st);
                hybrid_apply_kernel<<<c->num_sms, 32 * c->hyb.wpc[sg], size_t(c->hyb.wpc[sg]) * c->hyb.slot_words[sg] * 4, st>>>(ap);
            }
            li++;
        }
        gp.list = lists[li]; gp.list_count = c->hyb.d_count + li;
        {
            ProfScope ps(c, LDPC_K_HYBRID_CTA, st);
            if (gp.gmat) hybrid_ge_kernel<true><<<c->hyb.grid, kGeThreads, c->hyb.smem, st>>>(gp);
            else hybrid_ge_kernel<false><<<c->hyb.grid, kGeThreads, c->hyb.smem, st>>>(gp);
        }
        CUDA_TRY(cudaGetLastError());
        { int rcd = debug_sync("hybrid_ge_kernel", st); if (rcd) return rcd; }
    }
    return LDPC_OK;
}

extern "C" int ldpc_decode_ex(ldpc_ctx *c, const void *d_cw, const uint32_t *d_mask, void *d_out, uint8_t *d_fail,
                              uint8_t *d_fail_any, int max_iter, int mode, int64_t B, void *stream)
{
    if (!c || B < 0) return fail(LDPC_ERR_ARG, "bad argument to ldpc_decode");
    if (B == 0) return LDPC_OK;
    if (!d_cw || !d_mask || !d_out) return fail(LDPC_ERR_ARG, "NULL buffer passed to ldpc_decode");
    if (mode != LDPC_MODE_PEEL && mode != LDPC_MODE_HYBRID) return fail(LDPC_ERR_ARG, "unknown decode mode");
    if (max_iter < 0 || max_iter > 1000000) return fail(LDPC_ERR_ARG, "max_iter out of range");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const uint8_t *cw = static_cast<const uint8_t *>(d_cw);
    uint8_t *out = static_cast<uint8_t *>(d_out);
    const size_t in_cw = size_t(c->code.n) * c->S, out_cw = size_t(c->code.k) * c->S;
    const long long chunk = mode == LDPC_MODE_HYBRID ? c->hybrid_batch : c->max_batch;
    const bool pipe = mode == LDPC_MODE_PEEL && !c->prof_on && B > chunk && chunk_pipeline();
    if (pipe) {
        int rc = ensure_chunk_pipeline(c);
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(c->ev_fork, st));            // both internal streams start after what is queued on st
        for (int i = 0; i < 2; i++) CUDA_TRY(cudaStreamWaitEvent(c->pstream[i], c->ev_fork, 0));
    }
    int idx = 0;
    for (long long b0 = 0; b0 < B; b0 += chunk, idx++) {
        const long long nb = std::min<long long>(chunk, B - b0);
        int rc = decode_chunk(c, cw + size_t(b0) * in_cw, d_mask + size_t(b0) * c->NW, out + size_t(b0) * out_cw,
                              d_fail ? d_fail + b0 : nullptr, d_fail_any ? d_fail_any + b0 : nullptr, max_iter, mode, nb,
                              pipe ? c->pstream[idx & 1] : st, pipe ? (idx & 1) : 0);
        if (rc) return rc;
    }
    if (pipe)
        for (int i = 0; i < 2; i++) {                          // st continues when both are done
            CUDA_TRY(cudaEventRecord(c->ev_join[i], c->pstream[i]));
            CUDA_TRY(cudaStreamWaitEvent(st, c->ev_join[i], 0));
        }
    return LDPC_OK;
}

extern "C" int ldpc_decode(ldpc_ctx *c, const void *d_cw, const uint32_t *d_mask, void *d_out, uint8_t *d_fail,
                           int max_iter, int mode, int64_t B, void *stream)
{
    return ldpc_decode_ex(c, d_cw, d_mask, d_out, d_fail, nullptr, max_iter, mode, B, stream);
}

// ------------------------------------------------------------------------------------------
// error-rate run (pattern phase only)
// ------------------------------------------------------------------------------------------
extern "C" int ldpc_simulate_fer(ldpc_ctx *c, const ldpc_erasure_model *model, uint32_t seed, uint64_t frame0,
                                 int64_t frames, int max_iter, int mode, void *stream)
{
    if (!c || !model || frames < 0) return fail(LDPC_ERR_ARG, "bad argument to ldpc_simulate_fer");
    if (mode != LDPC_MODE_PEEL && mode != LDPC_MODE_HYBRID) return fail(LDPC_ERR_ARG, "unknown decode mode");
    if (max_iter < 0 || max_iter > 1000000) return fail(LDPC_ERR_ARG, "max_iter out of range");
    CUDA_TRY(cudaSetDevice(c->device));
    if (!c->d_sim_mask) CUDA_TRY(cudaMalloc(&c->d_sim_mask, size_t(c->max_batch) * c->NW * 4));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long sim_chunk = mode == LDPC_MODE_HYBRID ? c->hybrid_batch : c->max_batch;
    for (long long b0 = 0; b0 < frames; b0 += sim_chunk) {
        const long long nb = std::min<long long>(sim_chunk, frames - b0);
        int rc = ldpc_gen_erasures(c, model, seed, frame0 + uint64_t(b0), nb, c->d_sim_mask, nullptr, stream);
        if (rc) return rc;
        rc = decode_chunk(c, nullptr, c->d_sim_mask, nullptr, nullptr, nullptr, max_iter, mode, nb, st);
        if (rc) return rc;
    }
    return LDPC_OK;
}

// ------------------------------------------------------------------------------------------
// statistics
// ------------------------------------------------------------------------------------------
extern "C" int ldpc_get_stats(ldpc_ctx *c, ldpc_stats *out)
{
    if (!c || !out) return fail(LDPC_ERR_ARG, "NULL argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaDeviceSynchronize());
    unsigned long long h[8];
    CUDA_TRY(cudaMemcpy(h, c->d_stats, sizeof(h), cudaMemcpyDeviceToHost));
    out->frames = (int64_t)h[0]; out->ldpc_errors = (int64_t)h[1]; out->rs_errors = (int64_t)h[2];
    out->ml_attempts = (int64_t)h[3]; out->ml_failures = (int64_t)h[4]; out->ml_recovered = (int64_t)h[5];
    out->any_errors = (int64_t)h[6] - ((int64_t)h[3] - (int64_t)h[4]);   // unknown after peeling, minus the successful eliminations
    return LDPC_OK;
}

extern "C" int ldpc_reset_stats(ldpc_ctx *c)
{
    if (!c) return fail(LDPC_ERR_ARG, "NULL context");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemset(c->d_stats, 0, 8 * sizeof(unsigned long long)));
    return LDPC_OK;
}

// ------------------------------------------------------------------------------------------
// profiling
// ------------------------------------------------------------------------------------------
extern "C" int ldpc_profile_enable(ldpc_ctx *c, int on)
{
    if (!c) return fail(LDPC_ERR_ARG, "NULL context");
    c->prof_on = on != 0;
    return LDPC_OK;
}

extern "C" int ldpc_profile_read(ldpc_ctx *c, ldpc_profile *out, int reset)
{
    if (!c || !out) return fail(LDPC_ERR_ARG, "NULL argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaDeviceSynchronize());
    memset(out, 0, sizeof(*out));
    for (auto &r : c->prof_recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) out->ms[r.kind] += ms;
    }
    for (int i = 0; i < LDPC_K_KINDS; i++) out->launches[i] = c->launches[i];
    if (c->d_phase) {
        CUDA_TRY(cudaMemcpy(out->exec_phase_cycles, c->d_phase, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(out->ge_phase_cycles, c->d_phase + 8, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(out->apply_phase_cycles, c->d_phase + 16, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        if (reset) CUDA_TRY(cudaMemset(c->d_phase, 0, 24 * sizeof(unsigned long long)));
    }
    if (reset) {
        for (auto &r : c->prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
        c->prof_recs.clear();
        for (int i = 0; i < LDPC_K_KINDS; i++) c->launches[i] = 0;
    }
    return LDPC_OK;
}

// ------------------------------------------------------------------------------------------
// host-buffer entry points: chunked, staged over three streams so that the copy of chunk i+1 (and i+2)
// overlaps the kernels and the read-back of chunk i.
// ------------------------------------------------------------------------------------------
static void free_host_pipeline(ldpc_ctx *c)
{
    for (int i = 0; i < ldpc_ctx::kHostStages; i++) {
        if (c->hstream[i]) { cudaStreamDestroy(c->hstream[i]); c->hstream[i] = nullptr; }
        cudaFree(c->h_in[i]); cudaFree(c->h_out[i]); cudaFree(c->h_mask[i]); cudaFree(c->h_fail[i]); cudaFree(c->h_fail_any[i]);
        c->h_in[i] = nullptr; c->h_out[i] = nullptr; c->h_mask[i] = nullptr; c->h_fail[i] = nullptr; c->h_fail_any[i] = nullptr;
    }
    if (c->h_ev) { cudaEventDestroy(c->h_ev); c->h_ev = nullptr; }
    c->host_ready = false;
}

static int ensure_host_pipeline(ldpc_ctx *c)
{
    if (c->host_ready) return LDPC_OK;
    const size_t per_cw = size_t(c->code.n) * c->S;
    long long mb = 128;                                  // LDPC_CUDA_HOST_CHUNK_MB: bytes of input per pipeline stage (128 measured best: 299 vs 291 Gbit/s at 256)
    if (const char *e = getenv("LDPC_CUDA_HOST_CHUNK_MB")) mb = std::max(1ll, atoll(e));
    long long chunk = std::min<long long>(c->max_batch, std::max<long long>(256, (mb << 20) / (long long)per_cw));
    c->host_chunk = chunk;
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t r) { if (e == cudaSuccess) e = r; return e == cudaSuccess; };
    c->host_stages = ldpc_ctx::kHostStages;            // LDPC_CUDA_HOST_STAGES=2: round 1's double buffering (A/B comparisons)
    if (const char *es = getenv("LDPC_CUDA_HOST_STAGES")) c->host_stages = std::max(1, std::min(ldpc_ctx::kHostStages, atoi(es)));
    for (int i = 0; i < c->host_stages && e == cudaSuccess; i++) {
        ok(cudaStreamCreateWithFlags(&c->hstream[i], cudaStreamNonBlocking));
        ok(cudaMalloc(&c->h_in[i], size_t(chunk) * per_cw));
        ok(cudaMalloc(&c->h_out[i], size_t(chunk) * per_cw));
        ok(cudaMalloc(&c->h_mask[i], size_t(chunk) * c->NW * 4));
        ok(cudaMalloc(&c->h_fail[i], size_t(chunk)));
        ok(cudaMalloc(&c->h_fail_any[i], size_t(chunk)));
    }
    ok(cudaEventCreateWithFlags(&c->h_ev, cudaEventDisableTiming));
    if (e != cudaSuccess) {                              // nothing half-built survives: the next call starts clean
        free_host_pipeline(c);
        return fail(e == cudaErrorMemoryAllocation ? LDPC_ERR_NOMEM : LDPC_ERR_CUDA, std::string("host pipeline set-up: ") + cudaGetErrorString(e));
    }
    c->host_ready = true;
    return LDPC_OK;
}

extern "C" int ldpc_encode_host(ldpc_ctx *c, const void *h_info, void *h_cw, int64_t B)
{
    if (!c || !h_info || !h_cw || B < 0) return fail(LDPC_ERR_ARG, "bad argument to ldpc_encode_host");
    CUDA_TRY(cudaSetDevice(c->device));
    int rc = ensure_host_pipeline(c);
    if (rc) return rc;
    const size_t in_cw = size_t(c->code.k) * c->S, out_cw = size_t(c->code.n) * c->S;
    int i = 0;
    for (long long b0 = 0; b0 < B; b0 += c->host_chunk, i = (i + 1) % c->host_stages) {
        const long long nb = std::min<long long>(c->host_chunk, B - b0);
        cudaStream_t st = c->hstream[i];
        CUDA_TRY(cudaMemcpyAsync(c->h_in[i], static_cast<const uint8_t *>(h_info) + size_t(b0) * in_cw, size_t(nb) * in_cw,
                                 cudaMemcpyHostToDevice, st));
        rc = ldpc_encode(c, c->h_in[i], c->h_out[i], nb, st);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(static_cast<uint8_t *>(h_cw) + size_t(b0) * out_cw, c->h_out[i], size_t(nb) * out_cw,
                                 cudaMemcpyDeviceToHost, st));
    }
    for (int q = 0; q < c->host_stages; q++) CUDA_TRY(cudaStreamSynchronize(c->hstream[q]));
    return LDPC_OK;
}

// h_out != NULL: the reference's run() -- the k systematic symbols of every codeword are copied back.
// h_out == NULL (in place): only the symbols that were erased are written, by a kernel, straight into the caller's
// page-locked codeword buffer (d_hcw = that buffer as the device sees it).
static int decode_host_impl(ldpc_ctx *c, const void *h_cw, uint8_t *d_hcw, const uint32_t *h_mask, void *h_out, uint8_t *h_fail,
                            uint8_t *h_fail_any, int max_iter, int mode, int64_t B)
{
    CUDA_TRY(cudaSetDevice(c->device));
    int rc = ensure_host_pipeline(c);
    if (rc) return rc;
    const size_t in_cw = size_t(c->code.n) * c->S, out_cw = size_t(c->code.k) * c->S;
    // the scratch (schedule blobs) is shared by both streams: the peel/exec pair of chunk i+1 is ordered
    // after the pair of chunk i by an event, while the copies on either side overlap freely
    int i = 0;
    bool have_ev = false;
    // In place, the way up: gather_received_kernel fetches the received symbols itself (2 CTAs per SM measured best: 328 against
    // 313 Gbit/s with the plain copy of whole codewords; device-initiated reads reach ~44 GB/s on this link, the copy engine
    // 52, but 20 % fewer bytes cross).  LDPC_CUDA_HOST_GATHER = CTAs per SM of that kernel, 0 = plain copy.
    int gather = 2;
    if (const char *e = getenv("LDPC_CUDA_HOST_GATHER")) gather = std::max(0, std::min(8, atoi(e)));
    for (long long b0 = 0; b0 < B; b0 += c->host_chunk, i = (i + 1) % c->host_stages) {
        const long long nb = std::min<long long>(c->host_chunk, B - b0);
        cudaStream_t st = c->hstream[i];
        CUDA_TRY(cudaMemcpyAsync(c->h_mask[i], h_mask + size_t(b0) * c->NW, size_t(nb) * c->NW * 4, cudaMemcpyHostToDevice, st));
        if (d_hcw && gather) {
            ldpc::GatherParams gp;
            gp.src = d_hcw + size_t(b0) * in_cw; gp.mask = c->h_mask[i]; gp.dst = c->h_in[i]; gp.B = nb;
            gp.n = c->code.n; gp.S = c->S; gp.NW = c->NW;
            ldpc::gather_received_kernel<<<c->num_sms * gather, 256, 0, st>>>(gp);
            CUDA_TRY(cudaGetLastError());
            c->launches[LDPC_K_CHANNEL]++;
        } else {
            CUDA_TRY(cudaMemcpyAsync(c->h_in[i], static_cast<const uint8_t *>(h_cw) + size_t(b0) * in_cw, size_t(nb) * in_cw,
                                     cudaMemcpyHostToDevice, st));
        }
        if (have_ev) CUDA_TRY(cudaStreamWaitEvent(st, c->h_ev, 0));
        rc = ldpc_decode_ex(c, c->h_in[i], c->h_mask[i], c->h_out[i], c->h_fail[i], h_fail_any ? c->h_fail_any[i] : nullptr, max_iter, mode, nb, st);
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(c->h_ev, st));
        have_ev = true;
        if (h_out) {
            CUDA_TRY(cudaMemcpyAsync(static_cast<uint8_t *>(h_out) + size_t(b0) * out_cw, c->h_out[i], size_t(nb) * out_cw,
                                     cudaMemcpyDeviceToHost, st));
        } else {
            ldpc::WritebackParams wp;
            wp.out = c->h_out[i]; wp.mask = c->h_mask[i]; wp.dst = d_hcw + size_t(b0) * in_cw; wp.B = nb;
            wp.k = c->code.k; wp.n = c->code.n; wp.S = c->S; wp.NW = c->NW;
            const int grid = int(std::min<long long>((nb + 7) / 8, (long long)c->num_sms * 8));
            ldpc::writeback_erased_kernel<<<grid, 256, 0, st>>>(wp);
            CUDA_TRY(cudaGetLastError());
            c->launches[LDPC_K_CHANNEL]++;
        }
        if (h_fail) CUDA_TRY(cudaMemcpyAsync(h_fail + b0, c->h_fail[i], size_t(nb), cudaMemcpyDeviceToHost, st));
        if (h_fail_any) CUDA_TRY(cudaMemcpyAsync(h_fail_any + b0, c->h_fail_any[i], size_t(nb), cudaMemcpyDeviceToHost, st));
    }
    for (int q = 0; q < c->host_stages; q++) CUDA_TRY(cudaStreamSynchronize(c->hstream[q]));
    return LDPC_OK;
}

extern "C" int ldpc_decode_host_ex(ldpc_ctx *c, const void *h_cw, const uint32_t *h_mask, void *h_out, uint8_t *h_fail,
                                   uint8_t *h_fail_any, int max_iter, int mode, int64_t B)
{
    if (!c || !h_cw || !h_mask || !h_out || B < 0) return fail(LDPC_ERR_ARG, "bad argument to ldpc_decode_host");
    return decode_host_impl(c, h_cw, nullptr, h_mask, h_out, h_fail, h_fail_any, max_iter, mode, B);
}

// the caller's buffer as the device sees it; an error unless it is page-locked host memory
static int mapped_host_pointer(ldpc_ctx *c, void *h, const char *what, uint8_t **d)
{
    CUDA_TRY(cudaSetDevice(c->device));
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, h);
    if (e != cudaSuccess || at.type != cudaMemoryTypeHost || !at.devicePointer) {
        cudaGetLastError();
        return fail(LDPC_ERR_ARG, std::string(what) + ": the codeword buffer must be page-locked host memory (cudaHostAlloc / cudaHostRegister): "
                                                       "the device writes into it");
    }
    *d = static_cast<uint8_t *>(at.devicePointer);
    return LDPC_OK;
}

extern "C" int ldpc_decode_host_inplace(ldpc_ctx *c, void *h_cw, const uint32_t *h_mask, uint8_t *h_fail, uint8_t *h_fail_any,
                                        int max_iter, int mode, int64_t B)
{
    if (!c || !h_cw || !h_mask || B < 0) return fail(LDPC_ERR_ARG, "bad argument to ldpc_decode_host_inplace");
    if (B == 0) return LDPC_OK;
    uint8_t *d = nullptr;
    int rc = mapped_host_pointer(c, h_cw, "ldpc_decode_host_inplace", &d);
    if (rc) return rc;
    return decode_host_impl(c, h_cw, d, h_mask, nullptr, h_fail, h_fail_any, max_iter, mode, B);
}

// Encoder in place: h_cw [B][n][S] holds the information symbols in its first k rows; only those go up (a strided
// copy) and only the n-k parity rows come back.
extern "C" int ldpc_encode_host_inplace(ldpc_ctx *c, void *h_cw, int64_t B)
{
    if (!c || !h_cw || B < 0) return fail(LDPC_ERR_ARG, "bad argument to ldpc_encode_host_inplace");
    CUDA_TRY(cudaSetDevice(c->device));
    int rc = ensure_host_pipeline(c);
    if (rc) return rc;
    const size_t info_cw = size_t(c->code.k) * c->S, cw_b = size_t(c->code.n) * c->S, par_cw = cw_b - info_cw;
    int i = 0;
    for (long long b0 = 0; b0 < B; b0 += c->host_chunk, i = (i + 1) % c->host_stages) {
        const long long nb = std::min<long long>(c->host_chunk, B - b0);
        cudaStream_t st = c->hstream[i];
        uint8_t *h = static_cast<uint8_t *>(h_cw) + size_t(b0) * cw_b;
        CUDA_TRY(cudaMemcpy2DAsync(c->h_in[i], info_cw, h, cw_b, info_cw, size_t(nb), cudaMemcpyHostToDevice, st));
        rc = ldpc_encode(c, c->h_in[i], c->h_out[i], nb, st);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpy2DAsync(h + info_cw, cw_b, c->h_out[i] + info_cw, cw_b, par_cw, size_t(nb), cudaMemcpyDeviceToHost, st));
    }
    for (int q = 0; q < c->host_stages; q++) CUDA_TRY(cudaStreamSynchronize(c->hstream[q]));
    return LDPC_OK;
}

extern "C" int ldpc_decode_host(ldpc_ctx *c, const void *h_cw, const uint32_t *h_mask, void *h_out, uint8_t *h_fail,
                                int max_iter, int mode, int64_t B)
{
    return ldpc_decode_host_ex(c, h_cw, h_mask, h_out, h_fail, nullptr, max_iter, mode, B);
}

// ------------------------------------------------------------------------------------------
// multi-GPU host entry points: the batch is cut into contiguous frame ranges, GPU g gets
// [g*B/G, (g+1)*B/G); one host thread per GPU drives that GPU's host pipeline; no exchange between the GPUs
// (codewords are independent, ldpc_erasure_decoder.cl:27-104).
// ------------------------------------------------------------------------------------------
struct FanOutJob {
    const std::function<int(ldpc_ctx *, int64_t, int64_t)> *fn;
    ldpc_ctx *ctx;
    int64_t b0, nb;
    int rc;
    std::string err;
};

static void fan_out_worker(FanOutJob *job)
{
    job->rc = job->nb > 0 ? (*job->fn)(job->ctx, job->b0, job->nb) : LDPC_OK;
    if (job->rc) job->err = ldpc_last_error_string();     // (thread-local: the worker's own message)
}

static int fan_out(ldpc_ctx *const *ctxs, int n_ctx, int64_t B, const char *what,
                   const std::function<int(ldpc_ctx *, int64_t, int64_t)> &per_gpu)
{
    if (!ctxs || n_ctx <= 0 || B < 0) return fail(LDPC_ERR_ARG, std::string("bad argument to ") + what);
    for (int g = 0; g < n_ctx; g++) {
        if (!ctxs[g]) return fail(LDPC_ERR_ARG, std::string(what) + ": NULL context");
        if (ctxs[g]->code.n != ctxs[0]->code.n || ctxs[g]->code.k != ctxs[0]->code.k || ctxs[g]->S != ctxs[0]->S)
            return fail(LDPC_ERR_ARG, std::string(what) + ": the contexts must share one code and symbol size");
        for (int h = 0; h < g; h++)
            if (ctxs[h] == ctxs[g]) return fail(LDPC_ERR_ARG, std::string(what) + ": a context is listed twice (one host thread per context)");
    }
    std::vector<FanOutJob> jobs(static_cast<size_t>(n_ctx));
    std::vector<std::thread> thr;
    for (int g = 0; g < n_ctx; g++) {
        FanOutJob &j = jobs[size_t(g)];
        j.fn = &per_gpu; j.ctx = ctxs[g]; j.b0 = B * g / n_ctx; j.nb = B * (g + 1) / n_ctx - j.b0; j.rc = LDPC_OK;
    }
    for (int g = 0; g < n_ctx; g++) thr.emplace_back(fan_out_worker, &jobs[size_t(g)]);
    for (auto &t : thr) t.join();
    for (int g = 0; g < n_ctx; g++)
        if (jobs[size_t(g)].rc) return fail(jobs[size_t(g)].rc, "GPU " + std::to_string(ctxs[g]->device) + ": " + jobs[size_t(g)].err);
    return LDPC_OK;
}

extern "C" int ldpc_decode_host_multi(ldpc_ctx *const *ctxs, int n_ctx, const void *h_cw, const uint32_t *h_mask, void *h_out,
                                      uint8_t *h_fail, uint8_t *h_fail_any, int max_iter, int mode, int64_t B)
{
    if (!h_cw || !h_mask || !h_out) return fail(LDPC_ERR_ARG, "NULL buffer passed to ldpc_decode_host_multi");
    return fan_out(ctxs, n_ctx, B, "ldpc_decode_host_multi", [&](ldpc_ctx *c, int64_t b0, int64_t nb) {
        const size_t in_cw = size_t(c->code.n) * c->S, out_cw = size_t(c->code.k) * c->S;
        return ldpc_decode_host_ex(c, static_cast<const uint8_t *>(h_cw) + size_t(b0) * in_cw, h_mask + size_t(b0) * c->NW,
                                   static_cast<uint8_t *>(h_out) + size_t(b0) * out_cw, h_fail ? h_fail + b0 : nullptr,
                                   h_fail_any ? h_fail_any + b0 : nullptr, max_iter, mode, nb);
    });
}

extern "C" int ldpc_decode_host_inplace_multi(ldpc_ctx *const *ctxs, int n_ctx, void *h_cw, const uint32_t *h_mask, uint8_t *h_fail,
                                              uint8_t *h_fail_any, int max_iter, int mode, int64_t B)
{
    if (!h_cw || !h_mask) return fail(LDPC_ERR_ARG, "NULL buffer passed to ldpc_decode_host_inplace_multi");
    return fan_out(ctxs, n_ctx, B, "ldpc_decode_host_inplace_multi", [&](ldpc_ctx *c, int64_t b0, int64_t nb) {
        const size_t in_cw = size_t(c->code.n) * c->S;
        return ldpc_decode_host_inplace(c, static_cast<uint8_t *>(h_cw) + size_t(b0) * in_cw, h_mask + size_t(b0) * c->NW,
                                        h_fail ? h_fail + b0 : nullptr, h_fail_any ? h_fail_any + b0 : nullptr, max_iter, mode, nb);
    });
}

extern "C" int ldpc_encode_host_multi(ldpc_ctx *const *ctxs, int n_ctx, const void *h_info, void *h_cw, int64_t B)
{
    if (!h_info || !h_cw) return fail(LDPC_ERR_ARG, "NULL buffer passed to ldpc_encode_host_multi");
    return fan_out(ctxs, n_ctx, B, "ldpc_encode_host_multi", [&](ldpc_ctx *c, int64_t b0, int64_t nb) {
        const size_t in_cw = size_t(c->code.k) * c->S, out_cw = size_t(c->code.n) * c->S;
        return ldpc_encode_host(c, static_cast<const uint8_t *>(h_info) + size_t(b0) * in_cw, static_cast<uint8_t *>(h_cw) + size_t(b0) * out_cw, nb);
    });
}


// ------------------------------------------------------------------------------------------
// non-binary GF(256) LDPC code: Matlab/ErasureCodes_NonBinaryLDPCSim.m,
// Matlab/My_LDPC_HybridML_NonBinary_Erasure_Decoder.m.  The structure (and with it the peeling
// schedule) is the binary context's; this context adds one nonzero field element per edge.
// ------------------------------------------------------------------------------------------
struct ldpc_nb_ctx {
    ldpc_ctx *base = nullptr;
    std::vector<uint8_t> coef_csr;          // one coefficient per nonzero of H, CSR order
    uint8_t *d_coef = nullptr;              // [m][RW], pad 0
    uint8_t *d_tab = nullptr;               // log[256] | alog[512]
    uint32_t *d_m8 = nullptr;               // [256][8]
    uint8_t *d_enc_blob = nullptr;          // the encoder's static schedule
    int enc_blob_bytes = 0;
    int W = 0;                              // slice bytes of nb_exec_kernel (32, or 16 when S is not a multiple of 32)
    int smem_dec = 0, smem_enc = 0;
    // hybrid mode
    long long work_batch = 0;
    uint8_t *d_work = nullptr;              // [work_batch][n][S]
    unsigned int *d_list = nullptr, *d_count = nullptr;
    uint8_t *d_gA = nullptr, *d_gB = nullptr;
    int ge_grid = 0, ge_smem = 0, ge_small = 0;
};

static void nb_free(ldpc_nb_ctx *c)
{
    if (!c) return;
    if (c->base) cudaSetDevice(c->base->device);
    cudaFree(c->d_coef); cudaFree(c->d_tab); cudaFree(c->d_m8); cudaFree(c->d_enc_blob); cudaFree(c->d_work);
    cudaFree(c->d_list); cudaFree(c->d_count); cudaFree(c->d_gA); cudaFree(c->d_gB);
    delete c;
}

extern "C" int ldpc_nb_ctx_create(ldpc_nb_ctx **out, ldpc_ctx *base, const uint8_t *coef_csr, uint32_t coef_seed)
{
    if (!out || !base) return fail(LDPC_ERR_ARG, "NULL argument to ldpc_nb_ctx_create");
    *out = nullptr;
    const HostCode &code = base->code;
    ldpc_nb_ctx *c = new (std::nothrow) ldpc_nb_ctx();
    if (!c) return fail(LDPC_ERR_NOMEM, "out of host memory");
    c->base = base;
    c->coef_csr.resize(size_t(code.nnz));
    if (coef_csr) {
        for (int e = 0; e < code.nnz; e++) {
            if (!coef_csr[e]) { delete c; return fail(LDPC_ERR_ARG, "a coefficient of the non-binary code is zero"); }
            c->coef_csr[size_t(e)] = coef_csr[e];
        }
    } else {
        // floor((GF_SIZE-1)*rand)+1 (sim :55), drawn with Threefry4x32-20: key {3, seed}, counter = index of the nonzero
        for (int e = 0; e < code.nnz; e++) {
            const uint32_t ctr[4] = {uint32_t(e), 0u, 0u, 0u}, key[4] = {3u, coef_seed, 0u, 0u};
            uint32_t r[4];
            threefry4x32_20(ctr, key, r);
            c->coef_csr[size_t(e)] = uint8_t(1u + r[0] % 255u);
        }
    }
    std::vector<uint8_t> coef(size_t(code.m) * code.RW, 0);
    for (int r = 0; r < code.m; r++)
        for (int j = code.row_ptr[r]; j < code.row_ptr[r + 1]; j++) coef[size_t(r) * code.RW + (j - code.row_ptr[r])] = c->coef_csr[size_t(j)];
    uint8_t tab[768];
    rs_host_tables(tab, tab + 256);
    std::vector<uint32_t> m8(256 * 8);
    for (int v = 0; v < 256; v++) for (int j = 0; j < 8; j++) m8[size_t(v * 8 + j)] = ((v >> j) & 1) ? 0xFFFFFFFFu : 0u;
    const std::vector<uint8_t> blob = make_enc_blob(code, 0);      // (plain levels: nb_exec_kernel walks them with a CTA barrier each)
    c->enc_blob_bytes = int((blob.size() + 15) & ~size_t(15));
    c->W = base->S % 32 == 0 ? 32 : 16;
    const int fixed = code.m * code.RW * 2 + ((code.m * code.RW + 15) & ~15) + 256 * 8 * 4 + 768;
    c->smem_dec = code.n * c->W + fixed + sched_blob_max_bytes(code.m);
    c->smem_enc = code.n * c->W + fixed + c->enc_blob_bytes;
    if (std::max(c->smem_dec, c->smem_enc) > base->smem_optin) { delete c; return fail(LDPC_ERR_UNSUPPORTED, "code too long for the non-binary executor's shared memory"); }
    cudaError_t e = cudaSetDevice(base->device);
    auto ok = [&](cudaError_t r) { if (e == cudaSuccess) e = r; return e == cudaSuccess; };
    if (ok(cudaMalloc(&c->d_coef, coef.size())) && ok(cudaMemcpy(c->d_coef, coef.data(), coef.size(), cudaMemcpyHostToDevice)) &&
        ok(cudaMalloc(&c->d_tab, 768)) && ok(cudaMemcpy(c->d_tab, tab, 768, cudaMemcpyHostToDevice)) &&
        ok(cudaMalloc(&c->d_m8, m8.size() * 4)) && ok(cudaMemcpy(c->d_m8, m8.data(), m8.size() * 4, cudaMemcpyHostToDevice)) &&
        ok(cudaMalloc(&c->d_list, size_t(base->max_batch) * 4)) && ok(cudaMalloc(&c->d_count, 16)) && !blob.empty()) {
        std::vector<uint8_t> padded(size_t(c->enc_blob_bytes), 0);
        memcpy(padded.data(), blob.data(), blob.size());
        if (ok(cudaMalloc(&c->d_enc_blob, padded.size()))) ok(cudaMemcpy(c->d_enc_blob, padded.data(), padded.size(), cudaMemcpyHostToDevice));
    }
    if (e != cudaSuccess) { nb_free(c); return fail(e == cudaErrorMemoryAllocation ? LDPC_ERR_NOMEM : LDPC_ERR_CUDA, std::string("ldpc_nb_ctx_create: ") + cudaGetErrorString(e)); }
    *out = c;
    return LDPC_OK;
}

extern "C" int ldpc_nb_ctx_destroy(ldpc_nb_ctx *c)
{
    nb_free(c);
    return LDPC_OK;
}

extern "C" int ldpc_nb_get_coefficients(const ldpc_nb_ctx *c, uint8_t *coef_csr)
{
    if (!c || !coef_csr) return fail(LDPC_ERR_ARG, "NULL argument");
    memcpy(coef_csr, c->coef_csr.data(), c->coef_csr.size());
    return LDPC_OK;
}

static int nb_launch_exec(ldpc_nb_ctx *c, const void *d_in, int rows_in, void *d_out, int rows_out, const uint8_t *sched, int sched_stride,
                          int sched_max, int smem, long long B, cudaStream_t st)
{
    ldpc_ctx *b = c->base;
    NbExecParams p;
    p.in = static_cast<const uint8_t *>(d_in); p.out = static_cast<uint8_t *>(d_out); p.sched = sched; p.cidx = b->d_cidx; p.coef = c->d_coef;
    p.tab = c->d_tab; p.m8 = c->d_m8; p.B = B; p.sched_stride = sched_stride; p.sched_max = sched_max; p.n = b->code.n; p.m = b->code.m;
    p.RW = b->code.RW; p.S = b->S; p.rows_in = rows_in; p.rows_out = rows_out; p.slices = b->S / c->W;
    const int per_sm = std::max(1, b->smem_optin / (smem + 1024));
    const int grid = int(std::min<long long>(B * p.slices, (long long)b->num_sms * per_sm));
    auto k = c->W == 32 ? nb_exec_kernel<32> : nb_exec_kernel<16>;
    int rc = allow_max_smem(reinterpret_cast<const void *>(k), b->smem_optin);
    if (rc) return rc;
    {
        ProfScope ps(b, sched_stride ? LDPC_K_EXEC_DECODE : LDPC_K_EXEC_ENCODE, st);
        k<<<grid, kNbThreads, smem, st>>>(p);
    }
    CUDA_TRY(cudaGetLastError());
    return debug_sync("nb_exec_kernel", st);
}

extern "C" int ldpc_nb_encode(ldpc_nb_ctx *c, const void *d_info, void *d_cw, int64_t B, void *stream)
{
    if (!c || B < 0) return fail(LDPC_ERR_ARG, "bad argument to ldpc_nb_encode");
    if (B == 0) return LDPC_OK;
    if (!d_info || !d_cw) return fail(LDPC_ERR_ARG, "NULL buffer passed to ldpc_nb_encode");
    if (!c->base->code.triangular) return fail(LDPC_ERR_NOT_TRIANGULAR, "H is not in triangular form (last entry of row r must be column k + r)");
    CUDA_TRY(cudaSetDevice(c->base->device));
    return nb_launch_exec(c, d_info, c->base->code.k, d_cw, c->base->code.n, c->d_enc_blob, 0, c->enc_blob_bytes, c->smem_enc, B,
                          static_cast<cudaStream_t>(stream));
}

extern "C" int ldpc_nb_decode(ldpc_nb_ctx *c, const void *d_cw, const uint32_t *d_mask, void *d_out, uint8_t *d_fail, int max_iter,
                              int mode, int64_t B, void *stream)
{
    if (!c || B < 0) return fail(LDPC_ERR_ARG, "bad argument to ldpc_nb_decode");
    if (B == 0) return LDPC_OK;
    if (!d_cw || !d_mask || !d_out) return fail(LDPC_ERR_ARG, "NULL buffer passed to ldpc_nb_decode");
    if (mode != LDPC_MODE_PEEL && mode != LDPC_MODE_HYBRID) return fail(LDPC_ERR_ARG, "unknown decode mode");
    if (max_iter < 0 || max_iter > 1000000) return fail(LDPC_ERR_ARG, "max_iter out of range");
    ldpc_ctx *b = c->base;
    CUDA_TRY(cudaSetDevice(b->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t in_cw = size_t(b->code.n) * b->S, out_cw = size_t(b->code.k) * b->S;
    long long chunk = b->max_batch;
    if (mode == LDPC_MODE_HYBRID) {
        if (!c->d_work) {      // first hybrid call: the work buffer (whole codewords after the sweeps) and the elimination workspaces
            c->work_batch = std::max<long long>(1, std::min<long long>(b->max_batch, (1ll << 30) / (long long)in_cw));
            c->ge_grid = b->num_sms;
            c->ge_small = 768 + b->NW * 4 + ((b->NW + 1) & ~1) * 2 + ((b->code.m + 3) & ~3) + 3 * b->code.m * 2 + 64;
            c->ge_smem = b->smem_optin - 1024;       // the rest is the work area: A and b of a residual set that fits
            cudaError_t e = cudaSuccess;
            auto ok = [&](cudaError_t r) { if (e == cudaSuccess) e = r; return e == cudaSuccess; };
            if (!(ok(cudaMalloc(&c->d_work, size_t(c->work_batch) * in_cw)) && ok(cudaMalloc(&c->d_gA, size_t(c->ge_grid) * b->code.m * b->code.m)) &&
                  ok(cudaMalloc(&c->d_gB, size_t(c->ge_grid) * b->code.m * b->S)))) {
                cudaFree(c->d_work); cudaFree(c->d_gA); cudaFree(c->d_gB);
                c->d_work = nullptr; c->d_gA = nullptr; c->d_gB = nullptr;
                return fail(LDPC_ERR_NOMEM, std::string("non-binary hybrid scratch: ") + cudaGetErrorString(e));
            }
        }
        chunk = c->work_batch;
    }
    for (long long b0 = 0; b0 < B; b0 += chunk) {
        const long long nb = std::min<long long>(chunk, B - b0);
        const uint8_t *cw = static_cast<const uint8_t *>(d_cw) + size_t(b0) * in_cw;
        const uint32_t *mask = d_mask + size_t(b0) * b->NW;
        uint8_t *out = static_cast<uint8_t *>(d_out) + size_t(b0) * out_cw;
        uint8_t *failp = d_fail ? d_fail + b0 : b->d_fail_scratch;
        // pattern phase: the binary code's peel kernel (the schedule does not depend on the coefficients)
        PeelParams pp;
        pp.mask = mask; pp.sched = b->d_sched; pp.sched_len = b->d_sched_len; pp.work_ctr = b->d_work_ctr; pp.fail = failp; pp.fail_any = nullptr;
        pp.resid = b->d_resid; pp.stats = b->d_stats; pp.cidx = b->d_cidx; pp.vadj = b->d_vadj; pp.B = nb; pp.n = b->code.n; pp.k = b->code.k;
        pp.m = b->code.m; pp.RW = b->code.RW; pp.VW = b->code.VW; pp.NW = b->NW; pp.MW = b->MW; pp.stride = b->sched_stride; pp.max_iter = max_iter;
        pp.rs_n = b->rs_n; pp.rs_k = b->rs_k; pp.groups_per_block = b->peel_groups; pp.count_stats = 1;
        pp.ge_list = nullptr; pp.ge_count = nullptr;
        if (mode == LDPC_MODE_HYBRID) {
            CUDA_TRY(cudaMemsetAsync(c->d_count, 0, 16, st));
            pp.ge_list = c->d_list; pp.ge_count = c->d_count;
        }
        CUDA_TRY(cudaMemsetAsync(pp.work_ctr, 0, sizeof(unsigned int), st));
        {
            ProfScope ps(b, LDPC_K_PEEL, st);
            const int grid = int(std::min<long long>(b->num_sms, (nb + b->peel_groups - 1) / b->peel_groups));
            pick_peel(b->MW, b->code.VW)<<<grid, b->peel_groups * b->peel_G, b->peel_smem, st>>>(pp);
        }
        CUDA_TRY(cudaGetLastError());
        // payload phase
        const bool hyb = mode == LDPC_MODE_HYBRID;
        int rc = nb_launch_exec(c, cw, b->code.n, hyb ? static_cast<void *>(c->d_work) : static_cast<void *>(out), hyb ? b->code.n : b->code.k,
                                b->d_sched, b->sched_stride, sched_blob_max_bytes(b->code.m), c->smem_dec, nb, st);
        if (rc) return rc;
        if (hyb) {
            NbGeParams gp;
            gp.work = c->d_work; gp.mask = mask; gp.sched = b->d_sched; gp.list = c->d_list; gp.list_count = c->d_count; gp.fail = failp;
            gp.stats = b->d_stats; gp.cidx = b->d_cidx; gp.coef = c->d_coef; gp.tab = c->d_tab; gp.gA = c->d_gA; gp.gB = c->d_gB;
            gp.n = b->code.n; gp.k = b->code.k; gp.m = b->code.m; gp.RW = b->code.RW; gp.NW = b->NW; gp.S = b->S; gp.stride = b->sched_stride;
            gp.work_bytes = c->ge_smem - c->ge_small;
            { int rcs = allow_max_smem(reinterpret_cast<const void *>(nb_ge_kernel), b->smem_optin); if (rcs) return rcs; }
            {
                ProfScope ps(b, LDPC_K_HYBRID_CTA, st);
                nb_ge_kernel<<<c->ge_grid, kNbThreads, c->ge_smem, st>>>(gp);
            }
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaMemcpy2DAsync(out, out_cw, c->d_work, in_cw, out_cw, size_t(nb), cudaMemcpyDeviceToDevice, st));
            int rcd = debug_sync("nb_ge_kernel", st);
            if (rcd) return rcd;
        }
    }
    return LDPC_OK;
}

// ------------------------------------------------------------------------------------------
// Reed-Solomon entry points live in rs_gf256.cuh (rs_* functions below forward to it)
// ------------------------------------------------------------------------------------------
extern "C" int rs_ctx_create(rs_ctx **out, int n, int k, int symbol_bytes, int device, int64_t max_batch)
{
    return rs_create_impl(out, n, k, symbol_bytes, device, max_batch, g_err);
}
extern "C" int rs_ctx_destroy(rs_ctx *ctx) { return rs_destroy_impl(ctx); }
extern "C" int rs_ctx_get_generator(const rs_ctx *ctx, uint8_t *gsys) { return rs_get_generator_impl(ctx, gsys, g_err); }
extern "C" int rs_encode(rs_ctx *ctx, const void *d_info, void *d_cw, int64_t B, void *stream)
{
    return rs_encode_impl(ctx, d_info, d_cw, B, static_cast<cudaStream_t>(stream), g_err);
}
extern "C" int rs_decode(rs_ctx *ctx, const void *d_cw, const uint32_t *d_mask, void *d_out, uint8_t *d_fail, int64_t B,
                         void *stream)
{
    return rs_decode_impl(ctx, d_cw, d_mask, d_out, d_fail, B, static_cast<cudaStream_t>(stream), g_err);
}
