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
//  -> XOR the schedule into the slot
//  -> TMA store of the first rows_out rows.
// HBM traffic is exactly the algorithmic bytes (+ the schedule blob and the erasure mask): every
// input byte is read once and every output byte written once by TMA; all gathers hit shared memory.
//
// The XOR phase.  At the code's threshold a schedule is a LONG chain of SMALL levels ((2040,1530) at
// 20 % erasures: ~330 entries in ~24 levels -- 80, 40, 25, 19, 15 entries, then ~10 per level).  Walking
// that chain with one full check-row gather (14 shared-memory loads, XOR tree, store: ~400 cycles) per
// level made the phase a latency chain of ~10-17 k cycles per unit.  It is split instead:
//   1. BULK, dependency free, all four warps of the group: the rows the schedule produces are zeroed,
//      then every entry gathers ALL members of its check (the reference does the same: erased symbols are
//      zero and XORed in, ldpc_erasure_decoder.cl:17-20,68-75) -- which yields s_i = XOR of the check's
//      RECEIVED members -- and stores s_i into its target row.
//      Rounds run from the highest level down: an entry never reads the target of an entry of the same
//      or a higher level, so a round's stores cannot disturb a later round's gathers.
//   2. WALK: entry i of level >= 2 is  row[v_i] = s_i ^ XOR of its PRODUCED members (erased originally,
//      recovered at a lower level: ~2.6 per entry at 20 %), listed in an 8-byte record per entry: 6 loads
//      and one store per pass of <= 32 / LPG entries of one level.  The records and the list of passes depend on
//      the erasure mask alone; the group builds them (one thread per entry tests the members of the entry's check
//      against the mask) in the shadow of the slot's TMA load: the blob and the mask are small and land first, on
//      their own barrier.  The encoder's static blob carries its records and passes ready made (hmat.cpp).  A warp issues in order and a pass is ~90 instructions with its bookkeeping
//      (which entries, entry word, record, addresses), so the passes rotate over the group's four warps:
//      the passes are listed in the blob; warp p % 4 prepares pass p while the three passes before it run, waits for pass p-1 on a named
//      barrier (bar.sync / bar.arrive between two warps), then only loads, XORs, stores and signals.
//      The link of the chain is one shared-memory round trip plus the hand-off.  Idle lanes rewrite a
//      zero row.  Entries with more than five produced members (1 %), or past the records the blob
//      holds, take the full-row form row[v] ^= XOR of all members (the target itself cancels).
// Bank conflicts.  A row of W bytes covers W / 4 of the 32 banks: rows u, u' collide iff u = u' mod 128 / W
// (their "class").  A 128-bit shared load is served a quarter warp at a time, i.e. 128 / W entries together:
// the host arranges every check row in 128 / W groups of slots, group a holding the members of class a, and
// entry j of a quarter warp reads its row's groups rotated by j (a rotated ADDRESS, static registers), so at
// every step the entries of a quarter warp gather from different classes.
// The other groups of the CTA run other units at other phases and keep TMA traffic flowing.
#pragma once
#include "device_utils.cuh"

namespace ldpc {

constexpr int kExecWarpsPerGroup = 4;
constexpr int kExecMaxGroups = 3;      // named barriers: 1 + g per group, 4 + 4 g + w for the walk's hand-offs (16 in all)
constexpr int kBoxRows = 256;
constexpr int kExecZeroRowBytes = 128;   // behind every slot: what row padding reads, one 16..64-byte zero row per bank class
constexpr int kExecBulkRound = 4;        // entries a thread gathers between two barriers of the bulk phase

struct ExecParams {
    const uint16_t *rows;       // [m][RW] check rows arranged for this geometry (ldpc_cuda.cu: build_exec_rows): 16-byte offsets into a slot
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
    int zrow;                   // row index of the all-zero row inside a slot (sched_zero_row)
    uint8_t *synd;              // hybrid mode: [B][m][S] check syndromes of the codewords that still have erasures, or nullptr
    const uint32_t *mask;       // per-codeword schedules: [B][NW] erasure masks
    int NW, n;
    int msk_words;              // mask words per slot: covers the bits of the zero rows behind the slot
    int force_plain;            // test aid: per-codeword schedules are applied in the plain level-by-level form, always
    uint8_t *out;               // [B][rows_out][S]: the output the tensor maps describe (the walk's symbols are re-stored with plain stores)
    int S;
    int nfull_in, nfull_out;    // whole 256-row boxes moved by ONE 4-D tensor copy (map [B][rows/256][256][S]); the rest box by box
    int early_boxes;            // static schedule (encoder): whole boxes below row rows_in, stored as soon as they are loaded (0 = off)
    unsigned long long *phase_cycles;  // tuning aid (nullptr = off): [0] claim+issue, [1] load wait, [2] level walk, [3] store, [4] units, [5] zeroing + bulk gather, [6] walk passes
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
                    const __grid_constant__ CUtensorMap in4_map, const __grid_constant__ CUtensorMap out4_map,
                    const __grid_constant__ CUtensorMap outA_map, const __grid_constant__ CUtensorMap outB_map, const ExecParams p)
{
    constexpr int RWQ = (RWM + 7) / 8;                // uint4 chunks of a padded check row
    constexpr int LPG = W / 16;                       // lanes per entry
    constexpr int LPG_SH = LPG == 1 ? 0 : (LPG == 2 ? 1 : 2);
    constexpr int W_SH = LPG_SH + 4;                  // log2(W)
    constexpr int GT = kExecWarpsPerGroup * 32;       // threads per group
    constexpr int NGROUPS = GT / LPG;                 // entries a group handles per pass
    constexpr int EPW = 32 / LPG;                     // entries per pass of the walk
    constexpr int R = kExecBulkRound;
    extern __shared__ __align__(128) uint8_t smem[];
    // layout: [slots][blobs: nslot (per-codeword schedules) or 1 (static)][check rows][masks (hybrid)][barriers][unit mailboxes]
    const bool dynamic = p.sched_stride != 0;
    uint8_t *slots = smem;
    uint8_t *blobs = slots + size_t(p.nslot) * p.slot_bytes;
    uint16_t *cidx_s = reinterpret_cast<uint16_t *>(blobs + size_t(dynamic ? p.nslot : 1) * p.sched_max);
    uint32_t *masks = reinterpret_cast<uint32_t *>(cidx_s + size_t(p.m) * (RWQ * 8));
    const int msk_words = dynamic ? p.msk_words : 0;
    uint64_t *full = reinterpret_cast<uint64_t *>(masks + size_t(p.nslot) * msk_words);   // [nslot] the slot's payload; [4 + nslot] its blob
    int *mailbox = reinterpret_cast<int *>(full + 8);        // [g] unit of group g, [8] next unit, [12 + g] blob bytes of group g's unit

    const int g = threadIdx.x / GT;          // group = slot
    const int tg = threadIdx.x % GT;         // thread in group
    const bool leader = tg == 0;
    uint32_t *msk_s = masks + g * msk_words;
    const uint32_t cidx_a = smem_u32(cidx_s);

    // a check row: SL slots in NCLS groups; my entry reads the groups rotated by its place in the quarter warp
    constexpr int SL = RWQ * 8;                       // slots (16-bit offsets) per row
    constexpr int NCLS = 128 / W;                     // bank classes = entries per quarter warp
    constexpr int GS = SL / NCLS;                     // slots per group (1, 2, 4 or 8 -> 2 .. 16 bytes)
    static_assert(GS >= 1 && GS * NCLS == SL, "row slots must split into one group per bank class");
    const uint32_t rot = (uint32_t(threadIdx.x & 31) / LPG) % NCLS;
    auto load_row = [&](uint32_t c, uint32_t (&rr)[RWQ * 4]) {   // rr: SL slots as packed pairs
        const uint32_t row_a = cidx_a + c * (SL * 2);
#pragma unroll
        for (int a = 0; a < NCLS; a++) {
            const uint32_t ga = row_a + ((uint32_t(a) + rot) % NCLS) * (GS * 2);
            if (GS % 8 == 0) {
#pragma unroll
                for (int i = 0; i < GS / 8; i++) {
                    const uint4 r4 = lds128(ga + 16 * i);
                    const int w0 = (a * GS) / 2 + 4 * i;
                    rr[w0 + 0] = r4.x; rr[w0 + 1] = r4.y; rr[w0 + 2] = r4.z; rr[w0 + 3] = r4.w;
                }
            } else if (GS % 4 == 0) {
#pragma unroll
                for (int i = 0; i < GS / 4; i++)
                    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(rr[(a * GS) / 2 + 2 * i]), "=r"(rr[(a * GS) / 2 + 2 * i + 1]) : "r"(ga + 8 * i));
            } else if (GS % 2 == 0) {
#pragma unroll
                for (int i = 0; i < GS / 2; i++)
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(rr[(a * GS) / 2 + i]) : "r"(ga + 4 * i));
            } else {
#pragma unroll
                for (int i = 0; i < GS; i++) {
                    const int t = a * GS + i;
                    unsigned short h;
                    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(ga + 2 * i));
                    if (t & 1) rr[t >> 1] |= uint32_t(h) << 16; else rr[t >> 1] = h;
                }
            }
        }
    };

    {   // stage the arranged check rows once per CTA
        const uint4 *src = reinterpret_cast<const uint4 *>(p.rows);
        uint4 *dst = reinterpret_cast<uint4 *>(cidx_s);
        for (int i = threadIdx.x; i < p.m * RWQ; i += blockDim.x) dst[i] = src[i];
        for (int i = threadIdx.x; i < p.nslot * msk_words; i += blockDim.x) masks[i] = 0u;   // (the words behind the codeword's NW stay zero)
        if (!dynamic) {
            const uint4 *s2 = reinterpret_cast<const uint4 *>(p.sched);
            uint4 *d2 = reinterpret_cast<uint4 *>(blobs);
            for (int i = threadIdx.x; i < p.sched_max / 16; i += blockDim.x) d2[i] = s2[i];
        }
        for (int s = threadIdx.x / 8; s < p.nslot; s += blockDim.x / 8) {     // the zero rows (no TMA load or entry ever writes them)
            reinterpret_cast<uint4 *>(slots + size_t(s + 1) * p.slot_bytes - kExecZeroRowBytes)[threadIdx.x % 8] = make_uint4(0u, 0u, 0u, 0u);
            const size_t off = size_t(s) * p.slot_bytes + size_t(p.zrow) * W;      // (the records' zero row: the same place unless n > 3840)
            if ((threadIdx.x % 8) * 16 < W) reinterpret_cast<uint4 *>(slots + off)[threadIdx.x % 8] = make_uint4(0u, 0u, 0u, 0u);
        }
        if (threadIdx.x == 0) {
            for (int s = 0; s < p.nslot; s++) { mbar_init(&full[s], 1); mbar_init(&full[4 + s], 1); }
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
    const int eslot = tg / LPG;                   // entry slot inside the group

    while (true) {
        // ---- leader: start the loads of the unit claimed earlier, claim the one after -----------
        if (leader) {
            const int j = j_next;
            mailbox[g] = j;
            if (j < n_units) {
                const long long b = blockIdx.x + (long long)(j / p.slices) * gridDim.x;
                const int sl = j % p.slices;
                const uint32_t blen = blen_next;
                if (dynamic) {      // the blob first, on its own barrier: the group works on it while the slot loads
                    mailbox[12 + g] = int(blen);
                    mbar_arrive_expect_tx(&full[4 + g], blen);
                    bulk_load_1d(blob, p.sched + b * (long long)p.sched_stride, blen, &full[4 + g]);
                }
                mbar_arrive_expect_tx(&full[g], uint32_t(p.nbox_in) * kBoxRows * W);
                if (p.nfull_in) tma_load_4d(slot, &in4_map, sl * W, 0, 0, int(b), &full[g]);
                for (int i = p.nfull_in; i < p.nbox_in; i++)
                    tma_load_3d(slot + size_t(i) * kBoxRows * W, &in_map, sl * W, i * kBoxRows, int(b), &full[g]);
                claim();
            }
        }
        named_bar_sync(bar_id, GT);
        const int j = mailbox[g];
        if (j >= n_units) break;
        lap(0);
        const uint32_t *hdr = reinterpret_cast<const uint32_t *>(blob);
        int ne, nl, nrec, npass, n1;
        bool simple = false;               // no room for the pass table next to this blob: plain level-by-level form
        const uint32_t *ent;
        const uint16_t *lvo;
        uint32_t pt_a, rec_a;              // shared-space addresses of the pass table and the records
        if (dynamic) {
            // the codeword's erasure mask, then -- while the slot loads -- the passes and records of the walk
            const long long b = blockIdx.x + (long long)(j / p.slices) * gridDim.x;
            for (int w = tg; w < p.NW; w += GT) {
                uint32_t x = p.mask[b * p.NW + w];
                if (w == p.NW - 1 && (p.n & 31)) x &= 0xFFFFFFFFu >> (32 - (p.n & 31));
                msk_s[w] = x;
            }
            mbar_wait(&full[4 + g], phase);
            named_bar_sync(bar_id, GT);        // the mask words
            ne = int(hdr[0]); nl = int(hdr[1] & 0xFFFFu);
            ent = hdr + 4;
            lvo = reinterpret_cast<const uint16_t *>(ent + ne);
            n1 = nl >= 2 ? int(lvo[1]) : ne;   // entries of the first level: no produced members, no record
            pt_a = smem_u32(lvo) + 2u * uint32_t(nl + 1);
            const uint32_t area_end = smem_u32(blob) + uint32_t(p.sched_max);
            // passes: level index L = 1 .. nl-1 covers entries [lvo[L], lvo[L+1]), cut into pieces of <= EPW
            // (every warp computes the count; warp 0 writes the table: first entry | (count - 1) << 11)
            npass = 0;
            for (int L0 = 1; L0 < nl; L0 += 32) {
                const int L = L0 + (tg & 31);
                const int s0 = L < nl ? int(lvo[L]) : 0, s1 = L < nl ? int(lvo[L + 1]) : 0;
                const int np = (s1 - s0 + EPW - 1) / EPW;
                int inc = np;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                    if ((tg & 31) >= o) inc += t;
                }
                if (wg == 0) {
                    uint32_t at = pt_a + 2u * uint32_t(npass + inc - np);
                    for (int pos = s0; pos < s1 && at + 2u <= area_end; pos += EPW, at += 2u) {
                        const unsigned short pw = (unsigned short)(pos | ((min(EPW, s1 - pos) - 1) << 11));
                        asm volatile("st.shared.u16 [%0], %1;" ::"r"(at), "h"(pw) : "memory");
                    }
                }
                npass += __shfl_sync(0xFFFFFFFFu, inc, 31);
            }
            simple = p.force_plain || pt_a + 2u * uint32_t(npass) > area_end;     // (a blob near its maximum size in an area sized for it alone)
            rec_a = (pt_a + 2u * uint32_t(npass) + 7u) & ~7u;
            const int room = simple ? 0 : (int(area_end) - int(rec_a)) / 8;
            nrec = max(0, min(ne - n1, room));
            // records: the members of the entry's check that were erased on arrival, the target excepted
            const unsigned long long z = (unsigned long long)uint32_t(p.zrow);
            for (int i = tg; i < nrec; i += GT) {
                const uint32_t e = ent[n1 + i];
                const uint32_t voff16 = (e & 0xFFFFu) * LPG;
                uint32_t rr[RWQ * 4];
                load_row(e >> 16, rr);
                unsigned long long rec = 0ull;
                int nd = 0;
#pragma unroll
                for (int t = 0; t < SL; t++) {
                    const uint32_t o16 = (t & 1) ? (rr[t >> 1] >> 16) : (rr[t >> 1] & 0xFFFFu);
                    const uint32_t u = o16 >> LPG_SH;
                    const bool hit = ((msk_s[u >> 5] >> (u & 31)) & 1u) && o16 != voff16;
                    if (hit) { rec = (rec << 12) | u; nd++; }
                }
#pragma unroll
                for (int q = 0; q < 5; q++)
                    if (q >= nd) rec = (rec << 12) | z;
                if (nd > 5) rec = z | (z << 12) | (z << 24) | (z << 36) | (z << 48) | (1ull << 63);
                asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(rec_a + 8u * uint32_t(i)), "r"(uint32_t(rec)), "r"(uint32_t(rec >> 32)) : "memory");
            }
        } else {
            ne = int(hdr[0]); nl = int(hdr[1] & 0xFFFFu); nrec = int(hdr[1] >> 16); npass = int(hdr[2] >> 16);
            ent = hdr + 4;
            lvo = reinterpret_cast<const uint16_t *>(ent + ne);
            n1 = nl >= 2 ? int(lvo[1]) : ne;
            pt_a = smem_u32(lvo) + 2u * uint32_t(nl + 1);
            rec_a = (pt_a + 2u * uint32_t(npass) + 7u) & ~7u;
        }
        mbar_wait(&full[g], phase);
        phase ^= 1u;
        lap(1);
        const bool split = p.early_boxes > 0;      // encoder: information boxes go out now, the boxes with parity rows after the walk
        if (split && leader) {
            const long long b = blockIdx.x + (long long)(j / p.slices) * gridDim.x;
            tma_store_4d(&outA_map, slot, (j % p.slices) * W, 0, 0, int(b));
            bulk_commit();
        }

        const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
        const bool resid = p.synd && (hdr[2] & 0xFFFFu) != 0u;     // (uniform over the group)
        if (resid) {
            // Hybrid mode, the codeword keeps erasures after peeling: symbols that stay unknown must read as
            // zero when the syndromes are formed below, so every erased row is zeroed (a superset of the rows
            // the schedule produces).
            for (int u = eslot; u < p.rows_in; u += NGROUPS)
                if ((msk_s[u >> 5] >> (u & 31)) & 1u) sts128(base_a + u * W, z4);
        } else {
            // the rows this schedule produces read as zero until they are written
            for (int i = eslot; i < ne; i += NGROUPS) sts128(base_a + ((ent[i] & 0xFFFFu) << W_SH), z4);
        }
        named_bar_sync(bar_id, GT);

        if (simple) {
            // plain form: level by level, every entry gathers ALL members of its check (its own row reads as zero), a
            // barrier of the group between two levels
            for (int L = 0; L < nl; L++) {
                for (int i = int(lvo[L]) + eslot; i < int(lvo[L + 1]); i += NGROUPS) {
                    const uint32_t e = ent[i];
                    uint32_t rr[RWQ * 4];
                    load_row(e >> 16, rr);
                    uint4 acc = z4;
#pragma unroll
                    for (int t = 0; t < SL; t++) {
                        const uint32_t o16 = (t & 1) ? (rr[t >> 1] >> 16) : (rr[t >> 1] & 0xFFFFu);
                        xor_acc(acc, lds128(base_a + (o16 << 4)));
                    }
                    sts128(base_a + ((e & 0xFFFFu) << W_SH), acc);
                }
                named_bar_sync(bar_id, GT);
            }
            npass = 0;
        }
        // ---- 1. bulk: s_i = XOR of the check's received members, records of the produced ones --------
        for (int hi = simple ? 0 : ne; hi > 0; hi -= R * NGROUPS) {
            uint4 res[R];
            uint32_t va[R];
#pragma unroll
            for (int q = 0; q < R; q++) {
                const int idx = hi - 1 - (q * NGROUPS + eslot);
                va[q] = 0u;
                res[q] = z4;
                if (idx >= 0) {
                    const uint32_t e = ent[idx];
                    uint32_t rr[RWQ * 4];
                    load_row(e >> 16, rr);
                    uint4 val[SL];
#pragma unroll
                    for (int t = 0; t < SL; t++) {
                        const uint32_t o16 = (t & 1) ? (rr[t >> 1] >> 16) : (rr[t >> 1] & 0xFFFFu);
                        val[t] = lds128(base_a + (o16 << 4));
                    }
#pragma unroll
                    for (int st = 1; st < SL; st <<= 1)
#pragma unroll
                        for (int t = 0; t + st < SL; t += 2 * st) xor_acc(val[t], val[t + st]);
                    res[q] = val[0];
                    va[q] = base_a + ((e & 0xFFFFu) << W_SH);
                }
            }
            named_bar_sync(bar_id, GT);     // every gather of the round is done
#pragma unroll
            for (int q = 0; q < R; q++)
                if (va[q]) sts128(va[q], res[q]);
        }
        fence_proxy_async_smem();           // my writes to the slot -> visible to the TMA store below
        named_bar_sync(bar_id, GT);
        if (p.phase_cycles) lap(5);
        // Every symbol the schedule does not produce, and every symbol of the first level, is final: the first
        // rows_out rows go out NOW, while the level walk runs; the symbols the walk produces (holding s_i, or whatever the
        // asynchronous read catches) are stored again below, after this store has completed.
        if (leader && !split) {
            const long long b = blockIdx.x + (long long)(j / p.slices) * gridDim.x;
            const int sl = j % p.slices;
            if (p.nfull_out) tma_store_4d(&out4_map, slot, sl * W, 0, 0, int(b));
            for (int q = p.nfull_out; q < p.nbox_out; q++)
                tma_store_3d(&out_map, slot + size_t(q) * kBoxRows * W, sl * W, q * kBoxRows, int(b));
            bulk_commit();
        }

        // ---- 2. walk the levels >= 2: pass p (<= EPW entries of one level, listed in the blob) belongs to warp p % 4 ----
        if (npass > 0) {
            const uint32_t z = uint32_t(p.zrow);
            const uint32_t zlo = z | (z << 12) | (z << 24), zhi = (z >> 8) | (z << 4) | (z << 16);
            const uint32_t ent_a = smem_u32(ent);
            const int hand_in = 4 + g * 4 + ((wg + 3) & 3), hand_out = 4 + g * 4 + wg;
            uint32_t f_e, f_lo, f_hi;
            auto fetch = [&](int pass) {       // entry word and record of my lane's entry in that pass
                f_e = z; f_lo = zlo; f_hi = zhi;                   // an idle lane rewrites the zero row with zeros
                if (pass < npass) {
                    unsigned short pw;
                    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(pw) : "r"(pt_a + 2u * uint32_t(pass)));
                    const int idx = int(pw & 0x7FFu) + es;
                    if (es <= int(pw >> 11)) {
                        const int r = idx - n1;
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(f_e) : "r"(ent_a + 4u * uint32_t(idx)));
                        if (r < nrec) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(f_lo), "=r"(f_hi) : "r"(rec_a + 8u * uint32_t(r)));
                        else f_hi |= 0x80000000u;                  // no record: full-row form
                    }
                }
            };
            fetch(wg);
            unsigned long long mine = 0;
            for (int pass = wg; pass < npass; pass += kExecWarpsPerGroup) {
                const uint32_t e = f_e;
                const uint32_t ta = base_a + ((e & 0xFFFu) << W_SH);
                const uint32_t a0 = base_a + ((f_lo & 0xFFFu) << W_SH);
                const uint32_t a1 = base_a + (((f_lo >> 12) & 0xFFFu) << W_SH);
                const uint32_t a2 = base_a + ((((f_lo >> 24) | (f_hi << 8)) & 0xFFFu) << W_SH);
                const uint32_t a3 = base_a + (((f_hi >> 4) & 0xFFFu) << W_SH);
                const uint32_t a4 = base_a + (((f_hi >> 16) & 0xFFFu) << W_SH);
                const bool slow = (f_hi >> 31) != 0u;
                const bool any_slow = __any_sync(0xFFFFFFFFu, slow);
                if (pass > 0) asm volatile("bar.sync %0, 64;" ::"r"(hand_in) : "memory");      // pass p-1 is in the slot
                uint4 acc = lds128(ta);
                const uint4 x0 = lds128(a0), x1 = lds128(a1), x2 = lds128(a2), x3 = lds128(a3), x4 = lds128(a4);
                acc.x ^= x0.x ^ x1.x ^ x2.x ^ x3.x ^ x4.x;
                acc.y ^= x0.y ^ x1.y ^ x2.y ^ x3.y ^ x4.y;
                acc.z ^= x0.z ^ x1.z ^ x2.z ^ x3.z ^ x4.z;
                acc.w ^= x0.w ^ x1.w ^ x2.w ^ x3.w ^ x4.w;
                sts128(ta, acc);
                if (any_slow) {
                    if (slow) {
                        // full-row form: b2 = s ^ (R ^ D ^ s) = R ^ D, with s = R = XOR of the received members (the
                        // target row, one of the members) and D = XOR of the produced ones
                        uint32_t rr[RWQ * 4];
                        load_row(e >> 16, rr);
                        uint4 b2 = acc;
#pragma unroll
                        for (int u = 0; u < SL; u++) {
                            const uint32_t o16 = (u & 1) ? (rr[u >> 1] >> 16) : (rr[u >> 1] & 0xFFFFu);
                            xor_acc(b2, lds128(base_a + (o16 << 4)));
                        }
                        sts128(ta, b2);
                    }
                    __syncwarp();
                }
                if (pass + 1 < npass) asm volatile("bar.arrive %0, 64;" ::"r"(hand_out) : "memory");   // my pass is in the slot
                fetch(pass + kExecWarpsPerGroup);
                mine++;
            }
            if (p.phase_cycles && (tg & 31) == 0) atomicAdd(&p.phase_cycles[6], mine);
        }
        // ---- the walk's symbols, recycle the slot --------------------------------
        if (split) {
            fence_proxy_async_smem();           // the walk's writes -> visible to the TMA store
            named_bar_sync(bar_id, GT);
            if (leader) {
                const long long b = blockIdx.x + (long long)(j / p.slices) * gridDim.x;
                const int sl = j % p.slices;
                if (p.nfull_out > p.early_boxes)
                    tma_store_4d(&outB_map, slot + size_t(p.early_boxes) * kBoxRows * W, sl * W, 0, p.early_boxes, int(b));
                for (int q = p.nfull_out; q < p.nbox_out; q++)
                    tma_store_3d(&out_map, slot + size_t(q) * kBoxRows * W, sl * W, q * kBoxRows, int(b));
                bulk_commit();
                bulk_wait_read0();
            }
            npass = 0;                          // (nothing to store again)
        } else if (leader) {
            if (npass > 0) bulk_wait_all0();    // the early store is complete: what follows overwrites it in global memory
            else bulk_wait_read0();             // the slot's bytes are on their way to L2; it may be overwritten now
        }
        named_bar_sync(bar_id, GT);
        lap(2);
        if (p.synd) {
            // Hybrid mode, codeword still has erasures: the elimination stage needs, per check, the XOR of the
            // members known NOW.  They are all in the slot, so the syndromes are formed here -- one level, no
            // chain -- instead of being gathered from HBM later.  Every member is XORed in: the symbols that
            // are still unknown were zeroed above.
            if (resid) {
                const long long b = blockIdx.x + (long long)(j / p.slices) * gridDim.x;
                uint8_t *dst = p.synd + (size_t(b) * p.m) * (size_t(p.slices) * W) + size_t(j % p.slices) * W + (tg % LPG) * 16;
                for (int r = eslot; r < p.m; r += NGROUPS) {
                    uint32_t rr[RWQ * 4];
                    load_row(uint32_t(r), rr);
                    uint4 acc = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                    for (int t = 0; t < SL; t++) {
                        const uint32_t o16 = (t & 1) ? (rr[t >> 1] >> 16) : (rr[t >> 1] & 0xFFFFu);
                        xor_acc(acc, lds128(base_a + (o16 << 4)));
                    }
                    *reinterpret_cast<uint4 *>(dst + size_t(r) * (size_t(p.slices) * W)) = acc;
                }
            }
            named_bar_sync(bar_id, GT);   // the slot is recycled by the leader below: every gather must be done
        }
        if (npass > 0) {
            const long long b = blockIdx.x + (long long)(j / p.slices) * gridDim.x;
            uint8_t *dst = p.out + size_t(b) * p.rows_out * p.S + size_t(j % p.slices) * W + (tg % LPG) * 16;
            for (int i = n1 + eslot; i < ne; i += NGROUPS) {
                const uint32_t v = ent[i] & 0xFFFu;
                if (int(v) < p.rows_out) *reinterpret_cast<uint4 *>(dst + size_t(v) * p.S) = lds128(base_a + (v << W_SH));
            }
            named_bar_sync(bar_id, GT);         // the slot is recycled by the leader: every read of it is done
        }
        lap(3);
        if (p.phase_cycles && leader) atomicAdd(&p.phase_cycles[4], 1ull);
    }
    if (leader) bulk_wait_all0();
}

}  // namespace ldpc
