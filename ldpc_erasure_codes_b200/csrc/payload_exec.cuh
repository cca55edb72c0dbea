// payload_exec.cuh -- "payload phase": applies a recovery schedule to symbol payloads.
//
// One schedule entry (check c, symbol v) means  payload[v] = XOR of payload[u] over the other
// members u of check c.  That is the one operation both reference datapaths are made of:
//   * encoder  (OpenCL/device/ldpc_erasure_encoder.cl:72-90): parity k+r = XOR of row r's
//     members except the diagonal -- a STATIC schedule, one entry per check, levelled by the
//     rows' dependencies on earlier parities (built on the host, hmat.cpp);
//   * peeling decoder (OpenCL/device/ldpc_erasure_decoder.cl:68-90): the recovered symbol =
//     XOR of the check's other members -- a PER-CODEWORD schedule from peel_schedule.cuh.
// Entries of one level are independent; levels are separated by a barrier.
//
// Work unit = (codeword b, byte slice s): rows_in symbols x W bytes, W a multiple of 16 chosen
// so that a few units fit in shared memory (symbols' byte columns are independent and share the
// schedule).  Persistent kernel, one CTA per SM, `nslot` independent warp groups per CTA; each
// group owns one shared-memory slot and cycles
//     TMA load (cp.async.bulk.tensor, 3-D map [B][rows][S], box {W, 256, 1}, + the schedule blob
//               as a 1-D bulk copy, both completing on the slot's mbarrier)
//  -> XOR the schedule into the slot, level by level (a group of W/16 lanes owns one entry and
//     gathers the check's members with 128-bit shared loads, all issued before the first XOR)
//  -> TMA store of the first rows_out rows.
// The walk through the levels is a chain of shared-memory latencies (a level cannot start before
// the previous one has been written), so one unit alone leaves the SM idle; the groups run
// different units at different phases and fill each other's stalls, and while one group computes
// the others' TMA traffic keeps HBM busy.  HBM traffic is exactly the algorithmic bytes (+ the
// schedule blob): every input byte is read once and every output byte written once by TMA.
#pragma once
#include "device_utils.cuh"

namespace ldpc {

constexpr int kExecWarpsPerGroup = 4;
constexpr int kExecMaxGroups = 4;
constexpr int kBoxRows = 256;

struct ExecParams {
    const uint16_t *cidx;       // [m][RW] check rows, pad 0xFFFF
    const uint8_t *sched;       // per-codeword blobs (stride sched_stride) or one static blob
    const uint32_t *sched_len;  // [B] blob bytes, nullptr for a static schedule
    long long B;                // codewords in this launch
    int sched_stride;           // bytes between blobs (0 = static)
    int sched_max;              // shared bytes reserved per slot for a blob (static: size of the blob)
    int m, RW;
    int rows_in, rows_out;      // symbols loaded / stored per codeword
    int nbox_in, nbox_out;      // ceil(rows / 256)
    int slices;                 // S / W
    int nslot;                  // = number of warp groups
    int slot_bytes;             // shared bytes per slot
};

__device__ __forceinline__ void xor_acc(uint4 &a, const uint4 v)
{
    a.x ^= v.x; a.y ^= v.y; a.z ^= v.z; a.w ^= v.w;
}

template <int W, int RWM>  // W = slice bytes; RWM = members gathered per check (>= max row weight)
__global__ void __launch_bounds__(kExecMaxGroups *kExecWarpsPerGroup * 32, 1)
payload_exec_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                    const ExecParams p)
{
    constexpr int RWQ = (RWM + 7) / 8;                // uint4 chunks of a padded check row
    constexpr int LPG = W / 16;                       // lanes per entry
    constexpr int GT = kExecWarpsPerGroup * 32;       // threads per group
    constexpr int NGROUPS = GT / LPG;                 // entries a group handles per pass
    extern __shared__ __align__(1024) uint8_t smem[];
    // layout: [slots][blobs: nslot (dynamic) or 1 (static)][cidx][barriers][zero row][unit mailboxes]
    uint8_t *slots = smem;
    uint8_t *blobs = slots + size_t(p.nslot) * p.slot_bytes;
    const bool dynamic = p.sched_stride != 0;
    uint16_t *cidx_s = reinterpret_cast<uint16_t *>(blobs + size_t(dynamic ? p.nslot : 1) * p.sched_max);
    uint64_t *full = reinterpret_cast<uint64_t *>(cidx_s + size_t(p.m) * (RWQ * 8));   // [nslot]
    uint8_t *zrow = reinterpret_cast<uint8_t *>(full + 8);   // 64 zero bytes: what a skipped member reads
    int *mailbox = reinterpret_cast<int *>(zrow + 64);       // [g] unit of group g, [8] next-unit counter

    const int g = threadIdx.x / GT;          // group = slot
    const int tg = threadIdx.x % GT;         // thread in group
    const bool leader = tg == 0;

    {   // stage the check rows (and the static schedule) once per CTA
        const uint4 *src = reinterpret_cast<const uint4 *>(p.cidx);
        uint4 *dst = reinterpret_cast<uint4 *>(cidx_s);
        for (int i = threadIdx.x; i < p.m * RWQ; i += blockDim.x) dst[i] = src[i];
        if (!dynamic) {
            const uint4 *s2 = reinterpret_cast<const uint4 *>(p.sched);
            uint4 *d2 = reinterpret_cast<uint4 *>(blobs);
            for (int i = threadIdx.x; i < p.sched_max / 16; i += blockDim.x) d2[i] = s2[i];
        }
        if (threadIdx.x < 16) reinterpret_cast<uint32_t *>(zrow)[threadIdx.x] = 0u;
        if (threadIdx.x == 0) {
            for (int s = 0; s < p.nslot; s++) mbar_init(&full[s], 1);
            mailbox[8] = 0;
            mbar_fence_init();
        }
    }
    __syncthreads();

    // units of this CTA: codewords blockIdx.x, blockIdx.x + grid, ...; the slices of a codeword are
    // consecutive units, so they are in flight together (the 64-byte sectors a slice pulls into L2
    // also hold its neighbour slice)
    const long long cw_per_cta = (p.B > blockIdx.x) ? (p.B - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int n_units = int(cw_per_cta * p.slices);

    uint8_t *slot = slots + size_t(g) * p.slot_bytes;
    uint8_t *base = slot + (tg % LPG) * 16;
    const uint8_t *zero = zrow + (tg % LPG) * 16;
    uint8_t *blob = blobs + (dynamic ? size_t(g) * p.sched_max : 0);
    const int eg = tg / LPG;                 // entry lane-group inside the warp group
    const int bar_id = 1 + g;
    uint32_t phase = 0;

    while (true) {
        // ---- leader: claim a unit, start its loads ---------------------------------------
        if (leader) {
            const int j = atomicAdd(&mailbox[8], 1);
            mailbox[g] = j;
            if (j < n_units) {
                const long long b = blockIdx.x + (long long)(j / p.slices) * gridDim.x;
                const int sl = j % p.slices;
                uint32_t bytes = uint32_t(p.nbox_in) * kBoxRows * W;
                uint32_t blen = 0;
                if (dynamic) { blen = p.sched_len[b]; bytes += blen; }
                mbar_arrive_expect_tx(&full[g], bytes);
                for (int i = 0; i < p.nbox_in; i++)
                    tma_load_3d(slot + size_t(i) * kBoxRows * W, &in_map, sl * W, i * kBoxRows, int(b), &full[g]);
                if (dynamic) bulk_load_1d(blob, p.sched + b * (long long)p.sched_stride, blen, &full[g]);
            }
        }
        named_bar_sync(bar_id, GT);
        const int j = mailbox[g];
        if (j >= n_units) break;
        mbar_wait(&full[g], phase);
        phase ^= 1u;

        // ---- XOR the schedule into the slot -----------------------------------------------
        const uint32_t *hdr = reinterpret_cast<const uint32_t *>(blob);
        const int ne = int(hdr[0]);
        const int nl = int(hdr[1]);
        const uint32_t *ent = hdr + 4;
        const uint16_t *lvo = reinterpret_cast<const uint16_t *>(ent + ne);
        // software pipeline: the entry word and its check row for the NEXT level are fetched
        // before the barrier that ends the current level (they do not depend on the payload)
        int e1 = nl > 0 ? int(lvo[1]) : 0;
        int i = eg;                                   // my first entry of level 0 (starts at 0)
        uint32_t e = 0;
        uint32_t rr[RWQ * 4];
        auto fetch = [&](int idx) {
            e = ent[idx];
            const uint4 *row = reinterpret_cast<const uint4 *>(cidx_s + size_t(e >> 16) * (RWQ * 8));
#pragma unroll
            for (int qq = 0; qq < RWQ; qq++) {
                const uint4 r4 = row[qq];
                rr[qq * 4 + 0] = r4.x; rr[qq * 4 + 1] = r4.y; rr[qq * 4 + 2] = r4.z; rr[qq * 4 + 3] = r4.w;
            }
        };
        if (i < e1) fetch(i);
        for (int l = 0; l < nl; l++) {
            while (i < e1) {
                const uint32_t v = e & 0xFFFFu;
                uint4 val[RWM];
#pragma unroll
                for (int t = 0; t < RWM; t++) {
                    const uint32_t u = (rr[t >> 1] >> ((t & 1) * 16)) & 0xFFFFu;
                    const uint8_t *src = (u == v || u == 0xFFFFu) ? zero : base + size_t(u) * W;
                    val[t] = *reinterpret_cast<const uint4 *>(src);
                }
                uint4 acc = val[0];
#pragma unroll
                for (int t = 1; t < RWM; t++) xor_acc(acc, val[t]);
                *reinterpret_cast<uint4 *>(base + size_t(v) * W) = acc;
                i += NGROUPS;
                if (i < e1) fetch(i);
            }
            const int e0n = e1;                       // next level
            if (l + 1 < nl) e1 = int(lvo[l + 2]);
            i = e0n + eg;
            if (l + 1 < nl && i < e1) fetch(i);
            named_bar_sync(bar_id, GT);
        }

        // ---- store the first rows_out rows, recycle the slot --------------------------------
        fence_proxy_async_smem();
        named_bar_sync(bar_id, GT);
        if (leader) {
            const long long b = blockIdx.x + (long long)(j / p.slices) * gridDim.x;
            const int sl = j % p.slices;
            for (int q = 0; q < p.nbox_out; q++)
                tma_store_3d(&out_map, slot + size_t(q) * kBoxRows * W, sl * W, q * kBoxRows, int(b));
            bulk_commit();
            bulk_wait_read0();   // the slot's bytes are on their way to L2; it may be overwritten now
        }
    }
    if (leader) bulk_wait_all0();
}

}  // namespace ldpc
