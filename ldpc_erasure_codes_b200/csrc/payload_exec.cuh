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
    extern __shared__ __align__(1024) uint8_t smem[];
    // layout: [slots][blobs: nslot (dynamic) or 1 (static)][cidx][barriers][zero row][unit mailboxes]
    uint8_t *slots = smem;
    uint8_t *blobs = slots + size_t(p.nslot) * p.slot_bytes;
    const bool dynamic = p.sched_stride != 0;
    uint16_t *cidx_s = reinterpret_cast<uint16_t *>(blobs + size_t(dynamic ? p.nslot : 1) * p.sched_max);
    uint64_t *full = reinterpret_cast<uint64_t *>(cidx_s + size_t(p.m) * (RWQ * 8));   // [nslot]
    uint8_t *zrow = reinterpret_cast<uint8_t *>(full + 8);   // 64 zero bytes: what a skipped member reads
    int *mailbox = reinterpret_cast<int *>(zrow + 64);       // [g] unit of group g, [8] next unit, [16 + g] level hand-off
    uint32_t *msk_s = reinterpret_cast<uint32_t *>(mailbox + 32) + (threadIdx.x / GT) * ((p.NW + 3) & ~3);   // hybrid mode: the unit's mask

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
    const uint32_t base_a = smem_u32(slot) + (tg % LPG) * 16;    // shared-space addresses
    const uint32_t zero_a = smem_u32(zrow) + (tg % LPG) * 16;
    uint8_t *blob = blobs + (dynamic ? size_t(g) * p.sched_max : 0);
    const int bar_id = 1 + g;
    const uint32_t done_a = smem_u32(&mailbox[16 + g]);          // per-group "level workers done" counter
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
    while (true) {
        // ---- leader: start the loads of the unit claimed earlier, claim the one after -----------
        if (leader) {
            const int j = j_next;
            mailbox[g] = j;
            mailbox[16 + g] = 0;
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
        if (p.synd && reinterpret_cast<const uint32_t *>(blob)[2] != 0u) {   // (uniform over the group)
            // Symbols that stay unknown must read as zero when the syndromes are formed below.  Erased symbols
            // are zero on input by contract; zeroing them here makes the decoder independent of that.
            named_bar_sync(bar_id, GT);
            const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
            for (int u = tg / LPG; u < p.rows_in; u += NGROUPS)
                if ((msk_s[u >> 5] >> (u & 31)) & 1u) sts128(base_a + u * W, z4);
            named_bar_sync(bar_id, GT);
        }

        // ---- XOR the schedule into the slot -----------------------------------------------
        // The level walk is a chain of shared-memory round trips (a level's gathers cannot start
        // before the previous level's results are written, ~600-700 cycles per level all told), so:
        //  * everything that does NOT depend on the payload -- which entry a lane group handles
        //    next, the check's member list, the members' shared addresses -- is prepared ahead;
        //  * consecutive levels go to different warps of the group (a level starts at the warp after the
        //    last worker of the level below), and there is NO group-wide barrier between levels: a warp that finishes
        //    its part of level l bumps a completion counter (release) and goes on to prepare its next
        //    task; a warp about to execute level L spins (acquire) until the counter shows that all
        //    workers of the levels below are done.
        const uint32_t *hdr = reinterpret_cast<const uint32_t *>(blob);
        const int ne = int(hdr[0]);
        const int nl = int(hdr[1]);
        const uint32_t *ent = hdr + 4;
        const uint16_t *lvo = reinterpret_cast<const uint16_t *>(ent + ne);
        constexpr int EPW = 32 / LPG;                 // entries per warp and pass
        constexpr int WPGc = kExecWarpsPerGroup;
        const int wg = tg >> 5;                       // warp in group
        const int es = (tg & 31) / LPG;               // entry slot inside the warp
        int t_idx = 0, t_end = 0;
        bool t_valid = false;
        uint32_t t_dst = 0;
        uint32_t t_src[RWM];
        auto prepare = [&]() {                        // member addresses of entry t_idx
            const uint32_t e = ent[t_idx];
            const uint32_t v = e & 0xFFFFu;
            const uint4 *row = reinterpret_cast<const uint4 *>(cidx_s + size_t(e >> 16) * (RWQ * 8));
            uint32_t rr[RWQ * 4];
#pragma unroll
            for (int qq = 0; qq < RWQ; qq++) {
                const uint4 r4 = row[qq];
                rr[qq * 4 + 0] = r4.x; rr[qq * 4 + 1] = r4.y; rr[qq * 4 + 2] = r4.z; rr[qq * 4 + 3] = r4.w;
            }
#pragma unroll
            for (int t = 0; t < RWM; t++) {
                const uint32_t u = (t & 1) ? (rr[t >> 1] >> 16) : (rr[t >> 1] & 0xFFFFu);
                t_src[t] = (u == v || u == 0xFFFFu) ? zero_a : base_a + u * W;
            }
            t_dst = base_a + v * W;
        };
        // My warp's tasks (the levels it works in) are found 32 levels at a time with one ballot: lane i
        // looks at level l0+i.  Walking the level table entry by entry costs two dependent shared loads
        // per level and warp, which is as long as the work itself when the levels are small.
        constexpr unsigned FULLM = 0xFFFFFFFFu;
        const int wl = tg & 31;
        int l0 = -32;                                 // first level of the current chunk
        unsigned todo = 0u;                           // levels of the chunk where my warp still has to work
        int c_s0 = 0, c_s1 = 0;                       // lane i: entry range of level l0+i
        uint32_t c_dt = 0, done_base = 0;             // lane i: workers of all levels below l0+i; below the next chunk
        uint32_t done_target = 0;
        auto next_task = [&]() -> bool {              // warp-uniform; sets up t_* for my lane group
            while (!todo) {
                l0 += 32;
                if (l0 >= nl) return false;
                const int ll = l0 + wl;
                c_s0 = 0; c_s1 = 0;
                if (ll < nl) { c_s0 = lvo[ll]; c_s1 = lvo[ll + 1]; }
                const int cnt = c_s1 - c_s0;
                const int nw = ll < nl ? min(WPGc, (cnt + EPW - 1) / EPW) : 0;
                int inc = nw;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(FULLM, inc, o);
                    if (wl >= o) inc += t;
                }
                c_dt = done_base + uint32_t(inc - nw);
                done_base += uint32_t(__shfl_sync(FULLM, inc, 31));
                // a level's first entries go to the warp after the last worker of the level below (c_dt counts
                // the workers so far): two-warp levels then alternate between the warp pairs, and a warp's
                // preparation for its next level overlaps the other pair's turn
                todo = __ballot_sync(FULLM, ll < nl && int((uint32_t(wg) - c_dt) % WPGc) * EPW < cnt);
            }
            const int b = __ffs(todo) - 1;
            todo &= todo - 1u;
            const int s0 = __shfl_sync(FULLM, c_s0, b);
            t_end = __shfl_sync(FULLM, c_s1, b);
            done_target = __shfl_sync(FULLM, c_dt, b);
            t_idx = s0 + int((uint32_t(wg) - done_target) % WPGc) * EPW + es;
            t_valid = t_idx < t_end;
            if (t_valid) prepare();
            return true;
        };
        bool have = next_task();
        while (have) {
            if (done_target) {                        // wait for the levels below
                uint32_t seen, spins = 0;
                do {
                    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(seen) : "r"(done_a) : "memory");
                    if (++spins > (1u << 26)) {       // a lost hand-off must trap, not hang the GPU
                        printf("libldpc_cuda: level hand-off timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
                        __trap();
                    }
                } while (seen < done_target);
            }
            while (t_valid) {
                uint4 val[RWM];
#pragma unroll
                for (int t = 0; t < RWM; t++) val[t] = lds128(t_src[t]);
#pragma unroll
                for (int st = 1; st < RWM; st <<= 1)               // XOR tree
#pragma unroll
                    for (int t = 0; t + st < RWM; t += 2 * st) xor_acc(val[t], val[t + st]);
                sts128(t_dst, val[0]);
                t_idx += NGROUPS;                      // another pass in this level?
                t_valid = t_idx < t_end;
                if (t_valid) prepare();
            }
            __syncwarp();
            if (wl == 0)
                asm volatile("red.release.cta.shared::cta.add.u32 [%0], %1;" ::"r"(done_a), "r"(1u) : "memory");
            have = next_task();                        // off the critical path: my next level, addresses prepared
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
            if (reinterpret_cast<const uint32_t *>(blob)[2] != 0u) {
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
                        const uint32_t u = (t & 1) ? (rr[t >> 1] >> 16) : (rr[t >> 1] & 0xFFFFu);
                        const uint4 v = lds128(u == 0xFFFFu ? zero_a : base_a + u * W);
                        acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
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
