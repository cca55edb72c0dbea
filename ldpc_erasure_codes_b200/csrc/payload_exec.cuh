// payload_exec.cuh -- "payload phase": applies a recovery schedule to symbol payloads.
//
// One schedule entry (check c, symbol v) means  payload[v] = XOR of payload[u] over the other
// members u of check c.  That is the one operation both reference datapaths are made of:
//   * encoder  (OpenCL/device/ldpc_erasure_encoder.cl:72-90): parity k+r = XOR of row r's
//     members except the diagonal -- a STATIC schedule, one entry per check, levelled by the
//     rows' dependencies on earlier parities (built on the host, hmat.cpp);
//   * peeling decoder (OpenCL/device/ldpc_erasure_decoder.cl:68-90): the recovered symbol =
//     XOR of the check's other members -- a PER-CODEWORD schedule from peel_schedule.cuh.
// Entries of one level are independent; a level needs the levels below it.
//
// Work unit = (codeword b, byte slice s): rows_in symbols x W bytes, W a multiple of 16 chosen
// so that a few units fit in shared memory (symbols' byte columns are independent and share the
// schedule).  Persistent kernel, one CTA per SM, `nslot` independent warp groups per CTA; each
// group owns one shared-memory slot and cycles
//     TMA load (cp.async.bulk.tensor, 3-D / 4-D map over [B][rows][S], + the schedule blob as a
//               1-D bulk copy, all completing on the slot's mbarrier)
//  -> XOR the schedule into the slot, level by level
//  -> TMA store of the first rows_out rows.
// HBM traffic is exactly the algorithmic bytes (+ the schedule blob): every input byte is read
// once and every output byte written once by TMA; all gathers hit shared memory.
//
// The level walk.  At the code's threshold a schedule is a LONG chain of SMALL levels ((2040,1530) at
// 20 % erasures: ~330 entries in ~20 levels -- 80, 40, 25, 19, 15 entries, then ~10 per level), so the
// XOR phase is a chain of shared-memory round trips, not a bandwidth problem.  Its link is kept short:
//   * the rows a schedule produces are zeroed first and every member of the check is gathered,
//     the target included (the reference does the same: erased symbols are zero and XORed in,
//     ldpc_erasure_decoder.cl:17-20,68-75): no per-member compare/select;
//   * check rows are staged once per CTA as pre-scaled 16-byte offsets, row padding pointing at a
//     zero row behind the slot: a member's address is one shift-add;
//   * a level of more than `wide_min` entries is spread over the group's four warps and closed by the
//     group's named barrier; a run of smaller levels is walked by ONE warp alone, in program
//     order, with only __syncwarp() between levels -- no inter-warp hand-off on the chain -- while
//     the entry word and check row of the next pass are fetched ahead of the current pass's gathers.
// The other groups of the CTA run other units at other phases and keep TMA traffic flowing.
#pragma once
#include "device_utils.cuh"

namespace ldpc {

constexpr int kExecWarpsPerGroup = 4;
constexpr int kExecMaxGroups = 4;
constexpr int kBoxRows = 256;
constexpr int kExecZeroRowBytes = 128;   // behind every slot: what row padding reads

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
    int slot_bytes;             // shared bytes per slot (payload rows + the zero row)
    int wide_min;               // levels with more entries than this are spread over the group's warps
    uint8_t *synd;              // hybrid mode: [B][m][S] check syndromes of the codewords that still have erasures, or nullptr
    const uint32_t *mask;       // hybrid mode: [B][NW] erasure masks (erased rows are zeroed in the slot before the XOR phase)
    int NW;
    int nfull_in, nfull_out;    // whole 256-row boxes moved by ONE 4-D tensor copy (map [B][rows/256][256][S]); the rest box by box
    unsigned long long *phase_cycles;  // tuning aid (nullptr = off): [0] claim+issue, [1] load wait, [2] XOR, [3] store, [4] units
};

__device__ __forceinline__ void xor_acc(uint4 &a, const uint4 v)
{
    a.x ^= v.x; a.y ^= v.y; a.z ^= v.z; a.w ^= v.w;
}

__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4 v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int W, int RWM>  // W = slice bytes; RWM = members gathered per check (>= max row weight)
__global__ void __launch_bounds__(kExecMaxGroups *kExecWarpsPerGroup * 32, 1)
payload_exec_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                    const __grid_constant__ CUtensorMap in4_map, const __grid_constant__ CUtensorMap out4_map, const ExecParams p)
{
    constexpr int RWQ = (RWM + 7) / 8;                // uint4 chunks of a padded check row
    constexpr int LPG = W / 16;                       // lanes per entry
    constexpr int GT = kExecWarpsPerGroup * 32;       // threads per group
    constexpr int NGROUPS = GT / LPG;                 // entries a group handles per pass
    constexpr int EPW = 32 / LPG;                     // entries per warp and pass
    constexpr int WPGc = kExecWarpsPerGroup;
    extern __shared__ __align__(1024) uint8_t smem[];
    // layout: [slots][blobs: nslot (dynamic) or 1 (static)][check rows][barriers][unit mailboxes][masks (hybrid)]
    uint8_t *slots = smem;
    uint8_t *blobs = slots + size_t(p.nslot) * p.slot_bytes;
    const bool dynamic = p.sched_stride != 0;
    uint16_t *cidx_s = reinterpret_cast<uint16_t *>(blobs + size_t(dynamic ? p.nslot : 1) * p.sched_max);
    uint64_t *full = reinterpret_cast<uint64_t *>(cidx_s + size_t(p.m) * (RWQ * 8));   // [nslot]
    int *mailbox = reinterpret_cast<int *>(full + 8);        // [g] unit of group g, [8] next unit
    uint32_t *msk_s = reinterpret_cast<uint32_t *>(mailbox + 16) + (threadIdx.x / GT) * ((p.NW + 3) & ~3);   // hybrid mode: the unit's mask

    const int g = threadIdx.x / GT;          // group = slot
    const int tg = threadIdx.x % GT;         // thread in group
    const bool leader = tg == 0;
    const int zrow16 = (p.slot_bytes - kExecZeroRowBytes) / 16;   // the zero row, in 16-byte units from the slot base

    {   // stage the check rows once per CTA as 16-byte offsets into a slot (row u -> u * W / 16, padding -> the zero row)
        const uint32_t *src = reinterpret_cast<const uint32_t *>(p.cidx);
        uint32_t *dst = reinterpret_cast<uint32_t *>(cidx_s);
        for (int i = threadIdx.x; i < p.m * RWQ * 4; i += blockDim.x) {
            const uint32_t w = src[i];
            const uint32_t lo = w & 0xFFFFu, hi = w >> 16;
            dst[i] = (lo == 0xFFFFu ? uint32_t(zrow16) : lo * LPG) | ((hi == 0xFFFFu ? uint32_t(zrow16) : hi * LPG) << 16);
        }
        if (!dynamic) {
            const uint4 *s2 = reinterpret_cast<const uint4 *>(p.sched);
            uint4 *d2 = reinterpret_cast<uint4 *>(blobs);
            for (int i = threadIdx.x; i < p.sched_max / 16; i += blockDim.x) d2[i] = s2[i];
        }
        for (int s = threadIdx.x / 8; s < p.nslot; s += blockDim.x / 8)      // the zero rows (never touched by TMA)
            reinterpret_cast<uint4 *>(slots + size_t(s + 1) * p.slot_bytes - kExecZeroRowBytes)[threadIdx.x % 8] = make_uint4(0u, 0u, 0u, 0u);
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
    const uint32_t base_a = smem_u32(slot) + (tg % LPG) * 16;    // shared-space address of my 16-byte column
    uint8_t *blob = blobs + (dynamic ? size_t(g) * p.sched_max : 0);
    const int bar_id = 1 + g;
    uint32_t phase = 0;

    // leader state: the next unit is claimed (and its blob length fetched) one unit ahead, so that the
    // global-memory latency of that read is off the critical path
    int j_next = 0;
    uint32_t blen_next = 0;
    auto claim = [&]() {
        j_next = atomicAdd(&mailbox[8], 1);
        blen_next = 0;
        if (dynamic && j_next < n_units)
            blen_next = p.sched_len[blockIdx.x + (long long)(j_next / p.slices) * gridDim.x];
    };
    if (leader) claim();
    long long t_prev = p.phase_cycles ? clock64() : 0;
    auto lap = [&](int phase_id) {   // leader-only phase timer
        if (p.phase_cycles && leader) {
            const long long t = clock64();
            atomicAdd(&p.phase_cycles[phase_id], (unsigned long long)(t - t_prev));
            t_prev = t;
        }
    };

    const int wg = tg >> 5;                       // warp in group
    const int es = (tg & 31) / LPG;               // entry slot inside the warp

    while (true) {
        // ---- leader: start the loads of the unit claimed earlier, claim the one after -----------
        if (leader) {
            const int j = j_next;
            mailbox[g] = j;
            if (j < n_units) {
                const long long b = blockIdx.x + (long long)(j / p.slices) * gridDim.x;
                const int sl = j % p.slices;
                const uint32_t blen = blen_next;
                mbar_arrive_expect_tx(&full[g], uint32_t(p.nbox_in) * kBoxRows * W + blen);
                if (p.nfull_in) tma_load_4d(slot, &in4_map, sl * W, 0, 0, int(b), &full[g]);
                for (int i = p.nfull_in; i < p.nbox_in; i++)
                    tma_load_3d(slot + size_t(i) * kBoxRows * W, &in_map, sl * W, i * kBoxRows, int(b), &full[g]);
                if (dynamic) bulk_load_1d(blob, p.sched + b * (long long)p.sched_stride, blen, &full[g]);
                claim();
            }
        }
        named_bar_sync(bar_id, GT);
        const int j = mailbox[g];
        if (j >= n_units) break;
        lap(0);
        if (p.synd) {   // hybrid mode: fetch the codeword's erasure mask while the slot loads
            const long long b = blockIdx.x + (long long)(j / p.slices) * gridDim.x;
            for (int w = tg; w < p.NW; w += GT) msk_s[w] = p.mask[b * p.NW + w];
        }
        mbar_wait(&full[g], phase);
        phase ^= 1u;
        lap(1);

        const uint32_t *hdr = reinterpret_cast<const uint32_t *>(blob);
        const int ne = int(hdr[0]);
        const int nl = int(hdr[1]);
        const uint32_t *ent = hdr + 4;
        const uint16_t *lvo = reinterpret_cast<const uint16_t *>(ent + ne);
        const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
        if (p.synd && hdr[2] != 0u) {   // (uniform over the group)
            // Hybrid mode, the codeword keeps erasures after peeling: symbols that stay unknown must read as
            // zero when the syndromes are formed below, so every erased row is zeroed (a superset of the rows
            // the schedule produces).
            named_bar_sync(bar_id, GT);   // the mask words
            for (int u = tg / LPG; u < p.rows_in; u += NGROUPS)
                if ((msk_s[u >> 5] >> (u & 31)) & 1u) sts128(base_a + u * W, z4);
        } else {
            // the rows this schedule produces read as zero until they are written: every gather below takes
            // ALL members of its check, the target included
            for (int i = tg / LPG; i < ne; i += NGROUPS) sts128(base_a + (ent[i] & 0xFFFFu) * W, z4);
        }
        named_bar_sync(bar_id, GT);

        // ---- XOR the schedule into the slot -----------------------------------------------
        struct Prep { uint32_t e; uint32_t rr[RWQ * 4]; };
        auto fetch = [&](int idx, bool valid) -> Prep {   // entry word + its check row (payload independent)
            Prep q;
            q.e = valid ? ent[idx] : 0u;
            const uint4 *row = reinterpret_cast<const uint4 *>(cidx_s + size_t(q.e >> 16) * (RWQ * 8));
#pragma unroll
            for (int qq = 0; qq < RWQ; qq++) {
                const uint4 r4 = row[qq];
                q.rr[qq * 4 + 0] = r4.x; q.rr[qq * 4 + 1] = r4.y; q.rr[qq * 4 + 2] = r4.z; q.rr[qq * 4 + 3] = r4.w;
            }
            return q;
        };
        auto apply = [&](const Prep &q) {                 // gather every member, XOR tree, store the symbol
            uint4 val[RWM];
#pragma unroll
            for (int t = 0; t < RWM; t++) {
                const uint32_t o16 = (t & 1) ? (q.rr[t >> 1] >> 16) : (q.rr[t >> 1] & 0xFFFFu);
                val[t] = lds128(base_a + (o16 << 4));
            }
#pragma unroll
            for (int st = 1; st < RWM; st <<= 1)
#pragma unroll
                for (int t = 0; t + st < RWM; t += 2 * st) xor_acc(val[t], val[t + st]);
            sts128(base_a + (q.e & 0xFFFFu) * W, val[0]);
        };

        int L = 0;
        bool in_step = true;          // all warps of the group are at the same point of the walk
        while (L < nl) {
            const int s0 = lvo[L], s1 = lvo[L + 1];
            if (s1 - s0 > p.wide_min) {
                // a wide level: my warp takes entries s0 + wg * EPW + es, + 4 * EPW, ...
                if (!in_step) named_bar_sync(bar_id, GT);           // warp 0 has finished the small levels below
                int idx = s0 + wg * EPW + es;
                Prep cur = fetch(idx, idx < s1);
                for (int base = s0 + wg * EPW; base < s1; base += WPGc * EPW) {
                    const bool valid = idx < s1;
                    const int nidx = idx + WPGc * EPW;
                    const Prep nxt = fetch(nidx, nidx < s1);
                    if (valid) apply(cur);
                    cur = nxt;
                    idx = nidx;
                }
                named_bar_sync(bar_id, GT);
                in_step = true;
                L++;
            } else {
                // a run of small levels [L, L2): warp 0 walks it alone in program order
                int L2 = L + 1, e2 = s1;
                while (L2 < nl) {
                    const int nx = lvo[L2 + 1];
                    if (nx - e2 > p.wide_min) break;
                    e2 = nx;
                    L2++;
                }
                if (wg == 0) {
                    int lv = L, lend = s1;              // current level and its end
                    int pos = s0;                       // first entry of the current pass
                    Prep cur = fetch(pos + es, pos + es < lend);
                    bool cur_valid = pos + es < lend;
                    while (true) {
                        // next pass: the rest of this level, or the start of the next one
                        int npos = pos + EPW, nlv = lv, nlend = lend;
                        if (npos >= lend) { npos = lend; nlv = lv + 1; if (nlv < L2) nlend = lvo[nlv + 1]; }
                        const bool more = nlv < L2;
                        const bool nvalid = more && npos + es < nlend;
                        const Prep nxt = fetch(npos + es, nvalid);
                        if (cur_valid) apply(cur);
                        __syncwarp();                   // this pass's symbols are visible to the warp's next gathers
                        if (!more) break;
                        cur = nxt; cur_valid = nvalid; pos = npos; lv = nlv; lend = nlend;
                    }
                }
                in_step = false;
                L = L2;
            }
        }
        // ---- store the first rows_out rows, recycle the slot --------------------------------
        fence_proxy_async_smem();
        named_bar_sync(bar_id, GT);
        lap(2);
        if (p.synd) {
            // Hybrid mode, codeword still has erasures: the elimination stage needs, per check, the XOR of the
            // members known NOW.  They are all in the slot, so the syndromes are formed here -- one level, no
            // chain -- instead of being gathered from HBM later.  Every member is XORed in: the symbols that
            // are still unknown were zeroed above.
            if (hdr[2] != 0u) {
                const long long b = blockIdx.x + (long long)(j / p.slices) * gridDim.x;
                uint8_t *dst = p.synd + (size_t(b) * p.m) * (size_t(p.slices) * W) + size_t(j % p.slices) * W + (tg % LPG) * 16;
                for (int r = tg / LPG; r < p.m; r += NGROUPS) {
                    const uint4 *row = reinterpret_cast<const uint4 *>(cidx_s + size_t(r) * (RWQ * 8));
                    uint32_t rr[RWQ * 4];
#pragma unroll
                    for (int qq = 0; qq < RWQ; qq++) {
                        const uint4 r4 = row[qq];
                        rr[qq * 4 + 0] = r4.x; rr[qq * 4 + 1] = r4.y; rr[qq * 4 + 2] = r4.z; rr[qq * 4 + 3] = r4.w;
                    }
                    uint4 acc = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                    for (int t = 0; t < RWM; t++) {
                        const uint32_t o16 = (t & 1) ? (rr[t >> 1] >> 16) : (rr[t >> 1] & 0xFFFFu);
                        xor_acc(acc, lds128(base_a + (o16 << 4)));
                    }
                    *reinterpret_cast<uint4 *>(dst + size_t(r) * (size_t(p.slices) * W)) = acc;
                }
            }
            named_bar_sync(bar_id, GT);   // the slot is recycled by the leader below: every gather must be done
        }
        if (leader) {
            const long long b = blockIdx.x + (long long)(j / p.slices) * gridDim.x;
            const int sl = j % p.slices;
            if (p.nfull_out) tma_store_4d(&out4_map, slot, sl * W, 0, 0, int(b));
            for (int q = p.nfull_out; q < p.nbox_out; q++)
                tma_store_3d(&out_map, slot + size_t(q) * kBoxRows * W, sl * W, q * kBoxRows, int(b));
            bulk_commit();
            bulk_wait_read0();   // the slot's bytes are on their way to L2; it may be overwritten now
        }
        lap(3);
        if (p.phase_cycles && leader) atomicAdd(&p.phase_cycles[4], 1ull);
    }
    if (leader) bulk_wait_all0();
}

}  // namespace ldpc
