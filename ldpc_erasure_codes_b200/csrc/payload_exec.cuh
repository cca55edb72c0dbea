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
// schedule).  Persistent kernel, one CTA per SM:
//   producer warp (one elected lane): TMA-loads unit j+NSLOT's slice [rows][W] into a free slot
//       (cp.async.bulk.tensor, 3-D map [B][rows][S], box {W, 256, 1}) together with the unit's
//       schedule blob (1-D bulk copy); when the consumers finish a unit it TMA-stores the first
//       rows_out rows of the slot to the output tensor and recycles the slot;
//   8 consumer warps: wait on the slot's mbarrier, walk the levels; a group of W/16 lanes owns
//       one entry and gathers the check's members from the slot with 128-bit shared loads.
// HBM traffic is exactly the algorithmic bytes (+ the schedule blob): each input byte is read
// once by TMA, each output byte written once by TMA; all gathers hit shared memory.
#pragma once
#include "device_utils.cuh"

namespace ldpc {

constexpr int kExecConsumerWarps = 8;
constexpr int kExecThreads = (kExecConsumerWarps + 1) * 32;
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
    int nslot;
    int slot_bytes;             // nbox_in * 256 * W
};

// 128-bit XOR accumulate of one member row
__device__ __forceinline__ void xor_acc(uint4 &a, const uint4 v)
{
    a.x ^= v.x; a.y ^= v.y; a.z ^= v.z; a.w ^= v.w;
}

template <int W, int RWQ>  // W = slice bytes; RWQ = RW / 8 (uint4 chunks of a check row)
__global__ void __launch_bounds__(kExecThreads, 1)
payload_exec_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                    const ExecParams p)
{
    constexpr int LPG = W / 16;                       // lanes per entry
    constexpr int NGROUPS = kExecConsumerWarps * 32 / LPG;
    extern __shared__ __align__(1024) uint8_t smem[];
    // layout: [slots: nslot * slot_bytes][blobs: nslot * sched_max (dynamic) or 1 * sched_max][cidx][barriers]
    uint8_t *slots = smem;
    uint8_t *blobs = slots + size_t(p.nslot) * p.slot_bytes;
    const int n_blob_areas = p.sched_stride ? p.nslot : 1;
    uint16_t *cidx_s = reinterpret_cast<uint16_t *>(blobs + size_t(n_blob_areas) * p.sched_max);
    uint64_t *bars = reinterpret_cast<uint64_t *>(cidx_s + size_t(p.m) * (RWQ * 8));
    uint64_t *full = bars;              // [nslot] TMA bytes landed
    uint64_t *done = bars + p.nslot;    // [nslot] consumers finished the unit

    const int warp = threadIdx.x >> 5;
    const bool dynamic = p.sched_stride != 0;

    {   // stage the check rows (and the static schedule) once per CTA
        const uint4 *src = reinterpret_cast<const uint4 *>(p.cidx);
        uint4 *dst = reinterpret_cast<uint4 *>(cidx_s);
        for (int i = threadIdx.x; i < p.m * RWQ; i += blockDim.x) dst[i] = src[i];
        if (!dynamic) {
            const uint4 *s2 = reinterpret_cast<const uint4 *>(p.sched);
            uint4 *d2 = reinterpret_cast<uint4 *>(blobs);
            for (int i = threadIdx.x; i < p.sched_max / 16; i += blockDim.x) d2[i] = s2[i];
        }
        if (threadIdx.x == 0) {
            for (int s = 0; s < p.nslot; s++) {
                mbar_init(&full[s], 1);
                mbar_init(&done[s], kExecConsumerWarps * 32);
            }
            mbar_fence_init();
        }
    }
    __syncthreads();

    // units of this CTA: codewords blockIdx.x, blockIdx.x + grid, ... ; all slices of a codeword
    // back to back (the second slice's sectors were pulled into L2 by the first)
    const long long cw_per_cta = (p.B > blockIdx.x) ? (p.B - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long n_units = cw_per_cta * p.slices;

    if (warp == kExecConsumerWarps) {
        // ================= producer: one lane drives TMA =================
        if ((threadIdx.x & 31) == 0) {
            auto issue_load = [&](long long j) {
                const int slot = int(j % p.nslot);
                const long long b = blockIdx.x + (j / p.slices) * (long long)gridDim.x;
                const int sl = int(j % p.slices);
                uint8_t *dst = slots + size_t(slot) * p.slot_bytes;
                uint32_t bytes = uint32_t(p.nbox_in) * kBoxRows * W;
                uint32_t blen = 0;
                if (dynamic) { blen = p.sched_len[b]; bytes += blen; }
                mbar_arrive_expect_tx(&full[slot], bytes);
                for (int i = 0; i < p.nbox_in; i++)
                    tma_load_3d(dst + size_t(i) * kBoxRows * W, &in_map, sl * W, i * kBoxRows, int(b), &full[slot]);
                if (dynamic)
                    bulk_load_1d(blobs + size_t(slot) * p.sched_max, p.sched + b * (long long)p.sched_stride, blen,
                                 &full[slot]);
            };
            const long long pre = n_units < p.nslot ? n_units : p.nslot;
            for (long long j = 0; j < pre; j++) issue_load(j);
            for (long long j = 0; j < n_units; j++) {
                const int slot = int(j % p.nslot);
                mbar_wait(&done[slot], uint32_t((j / p.nslot) & 1));
                const long long b = blockIdx.x + (j / p.slices) * (long long)gridDim.x;
                const int sl = int(j % p.slices);
                const uint8_t *src = slots + size_t(slot) * p.slot_bytes;
                for (int i = 0; i < p.nbox_out; i++)
                    tma_store_3d(&out_map, src + size_t(i) * kBoxRows * W, sl * W, i * kBoxRows, int(b));
                bulk_commit();
                bulk_wait_read0();  // slot's bytes are in flight to L2; it may be overwritten now
                if (j + p.nslot < n_units) issue_load(j + p.nslot);
            }
            bulk_wait_all0();
        }
    } else {
        // ================= consumers: XOR the schedule into the slot =================
        const int tid = threadIdx.x;             // 0 .. 255
        const int grp = tid / LPG;
        const int q = tid % LPG;
        for (long long j = 0; j < n_units; j++) {
            const int slot = int(j % p.nslot);
            uint8_t *base = slots + size_t(slot) * p.slot_bytes + q * 16;
            const uint8_t *blob = blobs + (dynamic ? size_t(slot) * p.sched_max : 0);
            mbar_wait(&full[slot], uint32_t((j / p.nslot) & 1));
            const uint32_t *hdr = reinterpret_cast<const uint32_t *>(blob);
            const int ne = int(hdr[0]);
            const int nl = int(hdr[1]);
            const uint32_t *ent = hdr + 4;
            const uint16_t *lvo = reinterpret_cast<const uint16_t *>(ent + ne);
            for (int l = 0; l < nl; l++) {
                const int e0 = lvo[l], e1 = lvo[l + 1];
                for (int i = e0 + grp; i < e1; i += NGROUPS) {
                    const uint32_t e = ent[i];
                    const uint32_t v = e & 0xFFFFu;
                    const uint4 *row = reinterpret_cast<const uint4 *>(cidx_s + size_t(e >> 16) * (RWQ * 8));
                    uint4 acc = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                    for (int qq = 0; qq < RWQ; qq++) {
                        const uint4 r4 = row[qq];
                        const uint32_t rr[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                        for (int t = 0; t < 8; t++) {
                            const uint32_t u = (rr[t >> 1] >> ((t & 1) * 16)) & 0xFFFFu;
                            if (u != v && u != 0xFFFFu)
                                xor_acc(acc, *reinterpret_cast<const uint4 *>(base + size_t(u) * W));
                        }
                    }
                    *reinterpret_cast<uint4 *>(base + size_t(v) * W) = acc;
                }
                named_bar_sync(1, kExecConsumerWarps * 32);
            }
            fence_proxy_async_smem();
            mbar_arrive(&done[slot]);
        }
    }
}

}  // namespace ldpc
