// Dense fast path: an IRM with one domain and one dense binary relation R(e, e)
// (BASELINE.json config 3).  One persistent CTA per IRM / chain, resident across the IRM's
// whole move list, several CTAs per SM; every chain streams its own rows of the SHARED
// read-only slabs.
//
// HBM traffic per move = row i of slab0 (the observations (i, j)) + row i of slab1 (the
// observations (j, i)) = 2 * n bytes: exactly the algorithmic bytes of SURVEY.md 8(d).
// Everything mutable (z as uint8 slots, the K x K count table, CRP table sizes) lives in
// shared memory for the whole launch.
//
// The move order is static and the moves of one chain depend on each other only through
// the K x K count table and ONE entry of z, so the CTA is a three-role pipeline:
//
//   P  (1 warp)   TMA producer: 1-D bulk copies (cp.async.bulk + mbarrier complete_tx) of
//                 row chunks into a ring, running ahead across move boundaries.
//   A  (2 warps)  accumulate: group move r+1's 2n observations by the other entity's table
//                 (get_cluster_to_items_list, cxx/hirm.hh:388-397) WHILE move r is being decided.
//                 z[i_r] is still the old value then; the decision warps move that one
//                 observation pair to the right group afterwards (an O(1) fix-up).
//                 z lives in shared memory as bit planes (32 entities per word), the observation bytes
//                 are packed to bits with dp4a, and "ones of row i in table k" is popc(bits & mask_k):
//                 <= 8 live tables in registers, more in per-thread private shared-memory accumulators.
//   D  (2 warps)  decide move r: candidates + CRP prior (CRP::tables_weights_gibbs, :147-161),
//                 score (logp_gibbs_exact, :441-461) against the table with i's observations
//                 removed virtually, sample (log_choice, cxx/util_math.cc:58-65), and only
//                 if the table changed apply (set_cluster_assignment_gibbs, :524-551, :219-224).
//
// Tiers: the launch's shared-memory table holds kt slots.  With kt < max_tables more chains fit on an SM; a
// chain that needs slot kt stops BEFORE that move, stores its state and its position in progress[], and the
// host's next launch (kt = max_tables) resumes it.  Chains that are done exit at once.
//
// Live tables always occupy slots 0..K-1 (a dying table is replaced by the last slot), the
// slots in ascending label order are kept in s_order, and log(count) per table is cached, so
// that an unchanged move costs no bookkeeping.
#pragma once
#include "engine_types.cuh"

namespace hirm {

constexpr int DN_A_WARPS = 2;
constexpr int DN_DW_SMALL = 2, DN_DW_LARGE = 6;       // decision warps of the 8-slot tier / of the larger tiers
constexpr int DN_NA = 32 * DN_A_WARPS;
// threads of a CTA = 32 (producer warp) + DN_NA (A warps) + 32 * DW (D warps)
constexpr int DN_MAX_STAGES = 16;
constexpr int DN_LF = 256;                        // shared-memory copy of lgamma(m+1), m < 256; Stirling (error < 1e-19) beyond

// What the producer hands the decision warps for one move (through shared memory, ahead of the move's rows):
// nothing in the decision loop ever waits on a global load.
struct MoveCtl {
    int32_t item;
    uint32_t bytes;     // byte 0: observation (i, i); byte 1: (i, i_prev); byte 2: (i_prev, i); bit 24: bad move
    double u;           // the move's uniform
};
constexpr int DN_CTL = 32;                        // ring of control records (the producer is at most stages + 2 moves ahead)

struct DenseLaunch {
    int32_t n;          // entities
    int32_t npad;       // n rounded up to 16
    int32_t kt;         // table slots held in shared memory (<= max_tables: a chain that needs more stops and is
                        // resumed by the next, larger tier -- see progress[] below)
    int32_t kcap;       // max_tables of the domain (slot space of the global state)
    int32_t np;         // bit planes of z kept in shared memory: 2^np >= kt
    int32_t dw;         // decision warps of the kernel instance this layout is for
    int32_t full;       // 1: the relation has no missing cell and the domain no absent entity
    int32_t chunk;      // bytes of one row chunk (multiple of 32)
    int32_t nchunks;    // chunks per row
    int32_t stages;     // ring depth
    // byte offsets into dynamic shared memory
    int32_t off_lf, off_table, off_dbl, off_wpart, off_acc, off_ints, off_zp, off_ring, total_bytes;
};

// Shared-memory carve-up: 8-byte things first, then 4-byte, then the ring (128-aligned).
__host__ __device__ inline DenseLaunch dense_layout(int n, int kt, int kcap, int full, int chunk, int stages) {
    DenseLaunch L;
    L.dw = kt <= 8 ? DN_DW_SMALL : DN_DW_LARGE;
    L.n = n; L.npad = (n + 15) & ~15; L.kt = kt; L.kcap = kcap; L.full = full;
    L.np = kt <= 8 ? 3 : kt <= 16 ? 4 : kt <= 32 ? 5 : 6;
    chunk &= ~31;
    if (chunk > ((L.npad + 31) & ~31)) chunk = (L.npad + 31) & ~31;
    if (chunk < 32) chunk = 32;
    L.chunk = chunk; L.nchunks = (L.npad + chunk - 1) / chunk; L.stages = stages;
    int off = 8 * (2 * DN_MAX_STAGES + 4);          // mbarriers: full[S], empty[S], ready[2], go
    L.off_lf = off;    off += DN_LF * 8;
    L.off_table = off; off += kt * kt * 8;
    L.off_dbl = off;   off += 2 * (kt + 2) * 8;     // s_logcnt, s_logcntm1
    L.off_wpart = off; off += 2 * DN_A_WARPS * 4 * kt * 4;
    L.off_acc = off;   off += kt > 8 ? (full ? 1 : 2) * kt * DN_NA * 4 : 0;   // private accumulators of the > 8 tables path
    L.off_ints = off;  off += (2 * L.dw * 4 * kt + 3 * kt + kcap + 8) * 4;   // g[2][D warps][4kt], s_cnt, s_lab, s_order, s_map[kcap]
    L.off_zp = off;    off += ((n + 31) / 32) * L.np * 4;
    off = (off + 127) & ~127;
    L.off_ring = off;  off += stages * 2 * chunk;
    L.total_bytes = off;
    return L;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
// TMA 1-D bulk copy global -> shared, completion signalled on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory"); return t; }
// named barrier among the D warps only
template <int ND>
__device__ __forceinline__ void bar_d() { asm volatile("bar.sync 1, %0;" ::"n"(ND) : "memory"); }

// (n, s) of one observation byte
__device__ __forceinline__ int byte_n(uint32_t x) { return (x & 0x80u) ? 0 : 1; }
__device__ __forceinline__ int byte_s(uint32_t x) { return (x & 0x80u) ? 0 : (int)(x & 1u); }

// bytes at or beyond element n are padding: make them "missing" (or 0 where only ones are counted)
template <bool ZERO>
__device__ __forceinline__ void mask_tail(uint4 &x, int j0, int n) {
    uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int keep = n - (j0 + 4 * q);
        const uint32_t pad = keep <= 0 ? 0xFFFFFFFFu : keep < 4 ? ~((1u << (8 * keep)) - 1u) : 0u;
        w[q] = ZERO ? (w[q] & ~pad) : (w[q] | pad);
    }
    x = make_uint4(w[0], w[1], w[2], w[3]);
}

// ---- z as bit planes ---------------------------------------------------------------------------------
// The slot of every entity is kept in shared memory as NP bit planes, 32 entities per word:
// zp[g * NP + p] bit b = bit p of the slot of entity 32 g + b.  A live table's members are then a mask
// (an AND over the planes), and "how many of row i's ones fall into table k" is a popcount.
__device__ __forceinline__ int planes_get(const uint32_t *zp, int NP, int item) {
    const uint32_t *w = zp + (item >> 5) * NP;
    int slot = 0;
    for (int p = 0; p < NP; p++) slot |= (int)((w[p] >> (item & 31)) & 1u) << p;
    return slot;
}
__device__ __forceinline__ void planes_set(uint32_t *zp, int NP, int item, int slot) {
    uint32_t *w = zp + (item >> 5) * NP;
    const uint32_t bit = 1u << (item & 31);
    for (int p = 0; p < NP; p++) w[p] = ((slot >> p) & 1) ? (w[p] | bit) : (w[p] & ~bit);
}
// members of slot k among the 32 entities of a group, from its plane words P[0..NPL)
template <int NPL>
__device__ __forceinline__ uint32_t planes_mask(const uint32_t (&P)[NPL], int k) {
    uint32_t m = 0xFFFFFFFFu;
#pragma unroll
    for (int p = 0; p < NPL; p++) m &= ((k >> p) & 1) ? P[p] : ~P[p];
    return m;
}
// every entity in slot `from` moves to slot `to` (slot compaction when a table died)
__device__ __forceinline__ void planes_rename(uint32_t *zp, int NP, int G, int from, int to, int t, int nt) {
    #pragma unroll 1
    for (int g = t; g < G; g += nt) {
        uint32_t *w = zp + g * NP;
        uint32_t m = 0xFFFFFFFFu;
        for (int p = 0; p < NP; p++) m &= ((from >> p) & 1) ? w[p] : ~w[p];
        if (m) for (int p = 0; p < NP; p++) w[p] = ((to >> p) & 1) ? (w[p] | m) : (w[p] & ~m);
    }
}

// 32 observation bytes (two uint4: elements 0..15 and 16..31, each byte 0 or 1) -> 32 bits.  dp4a with the
// weights 1, 2, 4, 8 / 16, 32, 64, 128 turns eight bytes into one byte of bits.
__device__ __forceinline__ uint32_t pack32(const uint4 &lo, const uint4 &hi) {
    constexpr uint32_t W0 = 0x08040201u, W1 = 0x80402010u;
    const uint32_t b0 = __dp4a(lo.y, W1, __dp4a(lo.x, W0, 0u)), b1 = __dp4a(lo.w, W1, __dp4a(lo.z, W0, 0u));
    const uint32_t b2 = __dp4a(hi.y, W1, __dp4a(hi.x, W0, 0u)), b3 = __dp4a(hi.w, W1, __dp4a(hi.z, W0, 0u));
    return (b0 | (b1 << 8)) | ((b2 | (b3 << 8)) << 16);
}
__device__ __forceinline__ uint4 valid_bytes(const uint4 &x) {     // 1 where the byte is an observation (< 0x80)
    constexpr uint32_t ONES = 0x01010101u;
    return make_uint4((~x.x >> 7) & ONES, (~x.y >> 7) & ONES, (~x.z >> 7) & ONES, (~x.w >> 7) & ONES);
}
__device__ __forceinline__ uint4 and4(const uint4 &a, const uint4 &b) { return make_uint4(a.x & b.x, a.y & b.y, a.z & b.z, a.w & b.w); }

// ring cursor shared by the producer and the A warps (each keeps its own copy)
struct RingCursor {
    int stage; uint32_t parity;
    __device__ __forceinline__ void advance(int S) { if (++stage == S) { stage = 0; parity ^= 1u; } }
};

// shared-space loads (the hot loop addresses shared memory by 32-bit offsets: no generic pointers)
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
// try_wait with a suspend-time hint (ns): the thread sleeps in hardware until the phase completes or the time is up
__device__ __forceinline__ bool mbar_try_wait_hint_s(uint32_t bar, uint32_t parity, uint32_t ns) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(done) : "r"(bar), "r"(parity), "r"(ns) : "memory");
    return done != 0;
}
// for waits with slack (the A warps' go, the producer's empty): suspended in hardware, no issue slots taken
__device__ __forceinline__ void mbar_wait_relaxed_s(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait_hint_s(bar, parity, 20000u)) { }
}
__device__ __forceinline__ void mbar_arrive_s(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// What an A thread needs to stream a move (all by value, shared-space addresses)
struct APar {
    uint32_t full, empty, ring, zp;    // shared addresses: full[0], empty[0], ring, the bit planes
    int n, npad, chunk, nchunks, stages, np;
};

// One group of 32 elements: the row's / column's bits (and, unless the relation is fully observed, the
// "observed" bits) and the NPL plane words of the group.
template <bool FULL, int NPL>
struct Group32 {
    uint32_t sr, sc, vr, vc, P[NPL];
    __device__ __forceinline__ void load(const APar &p, uint32_t xrow, uint32_t xcol, int base, int g, bool tail) {
        uint4 r0 = lds128(xrow + 32u * g), r1 = lds128(xrow + 32u * g + 16u);
        uint4 c0 = lds128(xcol + 32u * g), c1 = lds128(xcol + 32u * g + 16u);
        const uint32_t pw = p.zp + 4u * (uint32_t)(((base >> 5) + g) * p.np);
#pragma unroll
        for (int q = 0; q < NPL; q++) P[q] = lds32(pw + 4u * q);
        if (tail) {
            const int j0 = base + 32 * g;
            mask_tail<FULL>(r0, j0, p.n); mask_tail<FULL>(r1, j0 + 16, p.n);
            mask_tail<FULL>(c0, j0, p.n); mask_tail<FULL>(c1, j0 + 16, p.n);
        }
        if (FULL) {
            sr = pack32(r0, r1); sc = pack32(c0, c1); vr = 0u; vc = 0u;
        } else {
            const uint4 a0 = valid_bytes(r0), a1 = valid_bytes(r1), b0 = valid_bytes(c0), b1 = valid_bytes(c1);
            vr = pack32(a0, a1); vc = pack32(b0, b1);
            sr = pack32(and4(r0, a0), and4(r1, a1)); sc = pack32(and4(c0, b0), and4(c1, b1));
        }
    }
};

// One move with at most KD live tables: group sums in registers; per-warp sums to wp[4k..4k+3].
template <int KD, int NPL, bool FULL>
__device__ __forceinline__ void a_move_regs(const APar p, RingCursor &cur, const int a, const int lane, int32_t *wp) {
    int rn[KD], rs[KD], cn[KD], cs[KD];
#pragma unroll
    for (int k = 0; k < KD; k++) { rn[k] = 0; rs[k] = 0; cn[k] = 0; cs[k] = 0; }
    #pragma unroll 1
    for (int q = 0; q < p.nchunks; q++) {
        const int base = q * p.chunk;
        const int bytes = min(p.chunk, p.npad - base);
        const uint32_t xrow = p.ring + (uint32_t)(cur.stage * 2 * p.chunk), xcol = xrow + (uint32_t)p.chunk;
        const int ng = (bytes + 31) >> 5;
        const bool tail = base + 32 * ng > p.n;        // the row's padding lies in this chunk (uniform)
        mbar_wait_s(p.full + 8u * cur.stage, cur.parity);
        #pragma unroll 1
        for (int g = a; g < ng; g += DN_NA) {
            Group32<FULL, NPL> G;
            G.load(p, xrow, xcol, base, g, tail);
#pragma unroll
            for (int k = 0; k < KD; k++) {
                const uint32_t m = planes_mask<NPL>(G.P, k);
                rs[k] += __popc(G.sr & m); cs[k] += __popc(G.sc & m);
                if (!FULL) { rn[k] += __popc(G.vr & m); cn[k] += __popc(G.vc & m); }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_s(p.empty + 8u * cur.stage);
        cur.advance(p.stages);
    }
#pragma unroll
    for (int k = 0; k < KD; k++) {
        const int s1 = __reduce_add_sync(0xffffffffu, rs[k]), s3 = __reduce_add_sync(0xffffffffu, cs[k]);
        if (lane == 0) { wp[4 * k + 1] = s1; wp[4 * k + 3] = s3; }
        if (!FULL) {
            const int s0 = __reduce_add_sync(0xffffffffu, rn[k]), s2 = __reduce_add_sync(0xffffffffu, cn[k]);
            if (lane == 0) { wp[4 * k] = s0; wp[4 * k + 2] = s2; }
        }
    }
}

// One move with more than 8 live tables: per-thread private accumulators in shared memory,
// acc[(k * F + f) * DN_NA + a] with 16-bit fields (rs | cs << 16, and rn | cn << 16 unless FULL).
template <int NPL, bool FULL>
__device__ __forceinline__ void a_move_acc(const APar p, RingCursor &cur, const int a, const int lane, uint32_t *acc, const int K, int32_t *wp) {
    constexpr int F = FULL ? 1 : 2;
    uint32_t *mine = acc + a;
    #pragma unroll 1
    for (int k = 0; k < K * F; k++) mine[k * DN_NA] = 0u;
    #pragma unroll 1
    for (int q = 0; q < p.nchunks; q++) {
        const int base = q * p.chunk;
        const int bytes = min(p.chunk, p.npad - base);
        const uint32_t xrow = p.ring + (uint32_t)(cur.stage * 2 * p.chunk), xcol = xrow + (uint32_t)p.chunk;
        const int ng = (bytes + 31) >> 5;
        const bool tail = base + 32 * ng > p.n;
        mbar_wait_s(p.full + 8u * cur.stage, cur.parity);
        #pragma unroll 1
        for (int g = a; g < ng; g += DN_NA) {
            Group32<FULL, NPL> G;
            G.load(p, xrow, xcol, base, g, tail);
            #pragma unroll 2
            for (int k = 0; k < K; k++) {
                const uint32_t m = planes_mask<NPL>(G.P, k);
                mine[(k * F) * DN_NA] += (uint32_t)__popc(G.sr & m) | ((uint32_t)__popc(G.sc & m) << 16);
                if (!FULL) mine[(k * F + 1) * DN_NA] += (uint32_t)__popc(G.vr & m) | ((uint32_t)__popc(G.vc & m) << 16);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_s(p.empty + 8u * cur.stage);
        cur.advance(p.stages);
    }
    #pragma unroll 1
    for (int k = 0; k < K; k++) {                          // this warp's 32 private accumulators -> wp
        const uint32_t w0 = mine[(k * F) * DN_NA];
        const int s1 = __reduce_add_sync(0xffffffffu, (int)(w0 & 0xFFFFu)), s3 = __reduce_add_sync(0xffffffffu, (int)(w0 >> 16));
        if (lane == 0) { wp[4 * k + 1] = s1; wp[4 * k + 3] = s3; }
        if (!FULL) {
            const uint32_t w1 = mine[(k * F + 1) * DN_NA];
            const int s0 = __reduce_add_sync(0xffffffffu, (int)(w1 & 0xFFFFu)), s2 = __reduce_add_sync(0xffffffffu, (int)(w1 >> 16));
            if (lane == 0) { wp[4 * k] = s0; wp[4 * k + 2] = s2; }
        }
    }
}

template <bool FULL>
__device__ __forceinline__ void a_move(const APar p, RingCursor &cur, const int a, const int lane, uint32_t *acc, const int K, int32_t *wp) {
    if (K <= 2) a_move_regs<2, 1, FULL>(p, cur, a, lane, wp);
    else if (K <= 4) a_move_regs<4, 2, FULL>(p, cur, a, lane, wp);
    else if (K <= 8) a_move_regs<8, 3, FULL>(p, cur, a, lane, wp);
    else if (K <= 16) a_move_acc<4, FULL>(p, cur, a, lane, acc, K, wp);
    else if (K <= 32) a_move_acc<5, FULL>(p, cur, a, lane, acc, K, wp);
    else a_move_acc<6, FULL>(p, cur, a, lane, acc, K, wp);
}

// DW = decision warps: 2 for the small-table tier (most chains per SM), more for the tiers whose chains have many
// tables (the score is K^2 cells: more lanes, fewer rounds).
// BYC: lay-out of a scoring round (see phase D): false = (candidate, cell) pairs flat over the lanes, true = some cells of every
// candidate per round (the instance of the last tier, whose chains have more than 32 tables).
template <int DW, int MINB, bool BYC = false>
__global__ void __launch_bounds__(32 + DN_NA + 32 * DW, MINB) dense_ree_sweep_kernel(SweepArgs args, const int32_t *blk_list, DenseLaunch L) {
    constexpr int ND = 32 * DW, NTHREADS = 32 + DN_NA + ND;
    const int k_list = blk_list ? blk_list[blockIdx.x] : blockIdx.x;
    const IrmDev &irm = args.irms[args.irm_ids[k_list]];
    const DomainDev &dd = irm.dom[0];
    const RelationDev &R = irm.rel[0];
    // everything read more than once from the descriptors goes to registers here: the loops below are full of
    // asm volatile "memory" clobbers, behind which the compiler must re-load through these references
    const uint8_t *const slab0 = R.slab[0], *const slab1 = R.slab[1];
    const int64_t pitch0 = R.pitch[0], pitch1 = R.pitch[1];
    int32_t *const z_g = dd.z;
    const int kcap_g = dd.kcap;
    const double alpha = dd.alpha;
    const bool full = L.full != 0;                  // host: the relation has no missing cell and no entity is absent
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = L.n, npad = L.npad, kt = L.kt, NS = L.stages;

    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(smem);                     // [DN_MAX_STAGES]
    uint64_t *bar_empty = bar_full + DN_MAX_STAGES;                              // [DN_MAX_STAGES]
    uint64_t *bar_ready = bar_empty + DN_MAX_STAGES;                             // [2]  A -> D: groups of move r complete
    uint64_t *bar_go = bar_ready + 2;                                            // [1]  D -> A: z is stable for move r
    double *s_lf = reinterpret_cast<double *>(smem + L.off_lf);                  // [DN_LF] lgamma(m+1)
    cell_t *table = reinterpret_cast<cell_t *>(smem + L.off_table);              // [kt*kt]  Relation::clusters, cell (a, b) at a + b*kt
    double *s_logcnt = reinterpret_cast<double *>(smem + L.off_dbl);                                         // log(count) per slot
    double *s_logcntm1 = s_logcnt + (kt + 2);                                    // log(count - 1) per slot
    int32_t *wpart = reinterpret_cast<int32_t *>(smem + L.off_wpart);            // [2][A_WARPS][4*kt] per-warp group sums
    uint32_t *acc = reinterpret_cast<uint32_t *>(smem + L.off_acc);              // [F*kt][DN_NA] (kt > 8 only)
    int32_t *g_buf = reinterpret_cast<int32_t *>(smem + L.off_ints);             // [2][D warps][4*kt]: (rn, rs, cn, cs) per other-slot, one copy per D warp
    int32_t *s_cnt = g_buf + 2 * DW * 4 * kt, *s_lab = s_cnt + kt;                        // CRP table sizes / labels per slot
    int32_t *s_order = s_lab + kt, *s_map = s_order + kt;                        // slots by ascending label; load-time slot map
    uint32_t *zp = reinterpret_cast<uint32_t *>(smem + L.off_zp);                // [G][np] bit planes of the slot of every entity
    const int NP = L.np, G = (n + 31) >> 5;
    uint8_t *ring = smem + L.off_ring;                                           // [NS][2][chunk]
    __shared__ int32_t s_K, s_KA;
    __shared__ MoveCtl s_ctl[DN_CTL];
    __shared__ double s_val[2 * ND];                   // one round of cell scores, double-buffered
    __shared__ double s_lpbuf[68];                     // the candidates' log-probabilities, from the lanes that added them up to every D warp
    __shared__ int32_t s_has_absent;
    __shared__ long long s_limit, s_issued, s_consumed, s_done;
    __shared__ long long s_ph[14];                        // profiling only   // first move NOT to be made (tier overflow); ring chunks issued / consumed

    const int64_t m0 = args.move_off[k_list], m1 = args.move_off[k_list + 1];
    const int64_t nmoves = m1 - m0;
    const int64_t r_begin = args.progress ? args.progress[blockIdx.x] : 0;   // moves already made by an earlier tier
    if (r_begin >= nmoves) return;

    // ---- load state, compacting live tables into slots 0..K-1 -----------------------------
    const int hi_g = *dd.hi;
    if (tid == 0) {
        int nt = 0;
        #pragma unroll 1
        for (int k = 0; k < hi_g; k++) nt += dd.tab_count[k] > 0;
        s_K = nt; s_KA = nt; s_has_absent = 0;
        s_limit = nmoves; s_issued = 0; s_consumed = 0;
        if (nt <= kt) {
            nt = 0;
            #pragma unroll 1
            for (int k = 0; k < hi_g; k++) {
                const int cnt = dd.tab_count[k];
                s_map[k] = cnt > 0 ? nt : 0;
                if (cnt > 0) { s_cnt[nt] = cnt; s_lab[nt] = dd.tab_label[k]; nt++; }
            }
            #pragma unroll 1
            for (int k = nt; k < kt; k++) s_cnt[k] = 0;
        }
        #pragma unroll 1
        for (int s = 0; s < NS; s++) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], DN_A_WARPS); }
        mbar_init(&bar_ready[0], DN_A_WARPS); mbar_init(&bar_ready[1], DN_A_WARPS);
        mbar_init(bar_go, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    #pragma unroll 1
    for (int k = tid; k < kt * kt; k += NTHREADS) table[k] = 0;
    #pragma unroll 1
    for (int k = tid; k < DN_LF; k += NTHREADS) s_lf[k] = g_logfact[k];
    __syncthreads();
    const int K0 = s_K;
    if (K0 > kt) return;                                       // more live tables than this tier holds: left to the next tier
    #pragma unroll 1
    for (int g = warp; g < G; g += NTHREADS / 32) {          // a warp turns 32 assignments into one word per plane
        const int j = 32 * g + lane;
        const int zj = j < n ? z_g[j] : 0;
        if (zj < 0) s_has_absent = 1;                          // rare: moves must then be checked against z = -1
        const int slot = zj >= 0 ? s_map[zj] : 0;              // an absent entity has only missing observations
        #pragma unroll 1
        for (int p = 0; p < NP; p++) {
            const uint32_t w = __ballot_sync(0xffffffffu, (slot >> p) & 1);
            if (lane == 0) zp[g * NP + p] = w;
        }
    }
    #pragma unroll 1
    for (int k = tid; k < hi_g * hi_g; k += NTHREADS) {
        const int a = k % hi_g, b = k / hi_g;
        const cell_t v = R.counts[a + b * kcap_g];
        if (v) table[s_map[a] + s_map[b] * kt] = v;            // cells of dead slots are empty
    }
    #pragma unroll 1
    for (int k = tid; k < K0; k += NTHREADS) {
        int rank = 0;
        const int lb = s_lab[k];
        #pragma unroll 1
        for (int j = 0; j < K0; j++) rank += s_lab[j] < lb;
        s_order[rank] = k;
        s_logcnt[k] = log_nl((double)s_cnt[k]);
        s_logcntm1[k] = log_nl((double)max(s_cnt[k] - 1, 1));
    }
    __syncthreads();

    if (warp == 0) {
        // =============================== P: TMA producer ===================================
        if (lane == 0) {
            RingCursor cur{0, 0u};
            int64_t issued = 0;                               // chunks issued so far
            const bool prof = args.phase_cycles != nullptr;
            long long p_wait = 0, p_t0 = prof ? clk() : 0;
            const volatile long long *limit = &s_limit;
            const bool has_absent = s_has_absent != 0;
            const uint64_t u_stream = args.uniforms ? 0 : args.stream[k_list], u_ctr0 = args.uniforms ? 0 : args.counter0[k_list];
            // software pipeline over the moves: ids two ahead, observation bytes and the injected uniform one ahead
            int it_cur = args.move_item[m0 + r_begin], dm_cur = args.move_dom[m0 + r_begin];
            int it_nxt = r_begin + 1 < nmoves ? args.move_item[m0 + r_begin + 1] : -1, dm_nxt = r_begin + 1 < nmoves ? args.move_dom[m0 + r_begin + 1] : 0;
            bool bad_cur = it_cur < 0 || it_cur >= n || dm_cur != 0;
            if (has_absent && !bad_cur) bad_cur = z_g[it_cur] < 0;
            uint32_t b_d = bad_cur ? 0xFFu : slab0[(int64_t)it_cur * pitch0 + it_cur], b_rp = 0xFFu, b_cp = 0xFFu;
            double u_cur = args.uniforms ? args.uniforms[m0 + r_begin] : 0.0;
            bool stop = false;
            #pragma unroll 1
            for (int64_t r = r_begin; r < nmoves && !stop; r++) {
                const int it = it_cur;
                // ---- this move's control record
                MoveCtl ctl;
                ctl.item = it; ctl.bytes = (b_d & 0xFFu) | ((b_rp & 0xFFu) << 8) | ((b_cp & 0xFFu) << 16) | (bad_cur ? 1u << 24 : 0u);
                ctl.u = args.uniforms ? u_cur : philox_uniform(args.seed, u_stream, u_ctr0 + (uint64_t)r);
                s_ctl[r % DN_CTL] = ctl;
                // ---- loads for the next moves (consumed one iteration later)
                const bool bad_prev = bad_cur;
                it_cur = it_nxt; dm_cur = dm_nxt;
                if (r + 2 < nmoves) { it_nxt = args.move_item[m0 + r + 2]; dm_nxt = args.move_dom[m0 + r + 2]; }
                b_d = 0xFFu; b_rp = 0xFFu; b_cp = 0xFFu;
                bad_cur = true;
                if (r + 1 < nmoves) {
                    bad_cur = it_cur < 0 || it_cur >= n || dm_cur != 0;
                    if (has_absent && !bad_cur) bad_cur = z_g[it_cur] < 0;
                    if (!bad_cur) {
                        const int64_t i2 = it_cur;
                        b_d = slab0[i2 * pitch0 + i2];
                        if (!bad_prev) { b_rp = slab0[i2 * pitch0 + it]; b_cp = slab1[i2 * pitch1 + it]; }
                    }
                    if (args.uniforms) u_cur = args.uniforms[m0 + r + 1];
                }
                const int64_t row = (it >= 0 && it < n) ? it : 0;
                const uint8_t *src0 = slab0 + row * pitch0, *src1 = slab1 + row * pitch1;
                #pragma unroll 1
                for (int q = 0; q < L.nchunks; q++) {
                    const int base = q * L.chunk;
                    const uint32_t bytes = (uint32_t)min(L.chunk, npad - base);
                    if (r >= *limit) { stop = true; break; }  // the chain stops before this move (tier overflow)
                    if (issued >= NS) {                       // the slot has been used before: wait until the A warps released it
                        const long long w0 = prof ? clk() : 0;
                        const uint32_t eb = smem_u32(&bar_empty[cur.stage]);
                        while (!mbar_try_wait_hint_s(eb, cur.parity ^ 1u, 20000u)) {
                            if (r >= *limit) { stop = true; break; }
                        }
                        if (stop) break;
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        if (prof) p_wait += clk() - w0;
                    }
                    uint8_t *dst = ring + (size_t)cur.stage * 2 * L.chunk;
                    mbar_expect_tx(&bar_full[cur.stage], 2 * bytes);
                    tma_load_1d(dst, src0 + base, bytes, &bar_full[cur.stage]);
                    tma_load_1d(dst + L.chunk, src1 + base, bytes, &bar_full[cur.stage]);
                    issued++;
                    cur.advance(NS);
                }
            }
            s_issued = issued;
            if (prof) { args.phase_cycles[blockIdx.x * 32 + 0] = clk() - p_t0; args.phase_cycles[blockIdx.x * 32 + 1] = p_wait; }
        }
    } else if (warp <= DN_A_WARPS) {
        // =============================== A: accumulate =====================================
        APar ap;
        ap.full = smem_u32(bar_full); ap.empty = smem_u32(bar_empty); ap.ring = smem_u32(ring); ap.zp = smem_u32(zp);
        ap.n = n; ap.npad = npad; ap.chunk = L.chunk; ap.nchunks = L.nchunks; ap.stages = NS; ap.np = NP;
        const int a = tid - 32, aw = warp - 1;
        const uint32_t go_s = smem_u32(bar_go), ready_s = smem_u32(bar_ready);
        RingCursor cur{0, 0u};
        const bool prof = args.phase_cycles != nullptr && a == 0;
        long long a_go = 0, a_t0 = prof ? clk() : 0;
        const volatile long long *limit = &s_limit;
        int64_t consumed = 0;
        #pragma unroll 1
        for (int64_t r = r_begin; r < nmoves; r++) {
            const int64_t rr = r - r_begin;                   // barrier phases count from this launch's first move
            const int buf = (int)(rr & 1);
            int32_t *wp = wpart + (buf * DN_A_WARPS + aw) * 4 * kt;
            const long long w0 = prof ? clk() : 0;
            mbar_wait_relaxed_s(go_s, (uint32_t)(rr & 1));    // z is stable (except z[i_{r-1}], fixed up by D)
            if (prof) a_go += clk() - w0;
            if (r >= *limit) break;                           // the chain stops before this move (tier overflow)
            const int KA = s_KA;
            if (full) a_move<true>(ap, cur, a, lane, acc, KA, wp);
            else a_move<false>(ap, cur, a, lane, acc, KA, wp);
            consumed += L.nchunks;
            __syncwarp();
            if (lane == 0) mbar_arrive_s(ready_s + 8u * buf);
        }
        if (a == 0) s_consumed = consumed;
        if (prof) { args.phase_cycles[blockIdx.x * 32 + 2] = clk() - a_t0; args.phase_cycles[blockIdx.x * 32 + 3] = a_go; }
    } else {
        // =============================== D: decide =========================================
        const int d = tid - 32 - DN_NA;                        // 0..ND-1
        const int dwarp = d >> 5;
        const double logalpha = log_nl(alpha);
        int K = K0;                                            // live slots 0..K-1 (uniform across D threads)
        int KA_prev = K0;                                      // K the A warps used for the move being decided
        bool pend_changed = false; int pend_item = -1, pend_old = 0, pend_new = 0, pend_dead = -1;
        if (d == 0) mbar_arrive(bar_go);                       // move 0 may be accumulated
        const bool prof = args.phase_cycles != nullptr && d == 0;
        long long tl = prof ? clk() : 0;
        if (prof) for (int i = 0; i < 14; i++) s_ph[i] = 0;
#define DPHASE(i) do { if (prof) { const long long _t = clk(); s_ph[i] += _t - tl; tl = _t; } } while (0)
        int64_t r_done = nmoves;                               // first move not made
        int vb = 0;                                            // s_val buffer of the next scoring round (alternates for the whole launch)
        #pragma unroll 1
        for (int64_t r = r_begin; r < nmoves; r++) {
            const int64_t m = m0 + r;
            const int64_t rr = r - r_begin;                    // barrier phases count from this launch's first move
            const int buf = (int)(rr & 1);
            int32_t *g = g_buf + (buf * DW + dwarp) * 4 * kt;     // this warp's copy of the group sums
            DPHASE(5);
            mbar_wait(&bar_ready[buf], (uint32_t)((rr >> 1) & 1));
            DPHASE(0);
            const MoveCtl ctl = s_ctl[r % DN_CTL];              // written by the producer before it issued this move's rows
            const int item = ctl.item;
            const bool bad = (ctl.bytes >> 24) & 1u;
            const uint32_t xb_ri = ctl.bytes & 0xFFu, xb_rp = (ctl.bytes >> 8) & 0xFFu, xb_cp = (ctl.bytes >> 16) & 0xFFu;
            // ---- finish move r-1: z[i_{r-1}] becomes visible, its observation pair changes group ----
            if (d == 0) {
                if (pend_changed) planes_set(zp, NP, pend_item, pend_new);
                if (pend_dead < 0) { s_KA = K; mbar_arrive(bar_go); }   // z is final: the A warps may accumulate move r+1
            }
            int c = bad ? 0 : ((pend_changed && item == pend_item) ? pend_new : planes_get(zp, NP, item));
            const int d_n = bad ? 0 : byte_n(xb_ri), d_s = bad ? 0 : byte_s(xb_ri);   // the diagonal (i, i): ONE observation
            {
                const int KG = min(K + 1, kt);                 // slot K: the fresh table (empty groups)
                const int32_t *wp0 = wpart + buf * DN_A_WARPS * 4 * kt;
                #pragma unroll 1
                for (int idx = lane; idx < 4 * KG; idx += 32) {      // every D warp computes all of them: no barrier
                    const int k = idx >> 2, f = idx & 3;
                    int v = 0;
                    if (full && !(f & 1)) {                     // no missing cell: the observation counts are the table sizes
                        g[idx] = (k < K && !bad) ? s_cnt[k] - (k == c ? 1 : 0) : 0;
                        continue;
                    }
                    if (k < KA_prev) {
#pragma unroll
                        for (int w = 0; w < DN_A_WARPS; w++) v += wp0[w * 4 * kt + idx];
                    }
                    if (pend_changed) {                        // A saw i_{r-1} in its old slot
                        const uint32_t xb = f < 2 ? xb_rp : xb_cp;
                        const int val = (f & 1) ? byte_s(xb) : byte_n(xb);
                        if (k == pend_old) v -= val;
                        if (k == pend_new) v += val;
                    }
                    if (k == c) v -= (f & 1) ? d_s : d_n;      // drop the diagonal from both groups
                    g[idx] = v;
                }
            }
            __syncwarp();
            if (pend_dead >= 0) {
                bar_d<ND>();
                // ---- the table that died in move r-1 is replaced by the last slot ------------------
                const int dead = pend_dead, last = K - 1;
                if (dead != last) {
                    planes_rename(zp, NP, G, last, dead, d, ND);
                    // row/column `last` -> `dead` (row/column `dead` is all zero)
                    #pragma unroll 1
                    for (int k = d; k < K; k += ND) {
                        if (k != last && k != dead) {
                            table[dead + k * kt] = table[last + k * kt]; table[last + k * kt] = 0;
                            table[k + dead * kt] = table[k + last * kt]; table[k + last * kt] = 0;
                        } else if (k == last) {
                            table[dead + dead * kt] = table[last + last * kt]; table[last + last * kt] = 0;
                        }
                    }
                    if (d < 4 * DW) {
                        int32_t *gw = g_buf + (buf * DW + (d >> 2)) * 4 * kt;
                        gw[4 * dead + (d & 3)] = gw[4 * last + (d & 3)]; gw[4 * last + (d & 3)] = 0;
                    }
                    if (d == 4) {
                        s_cnt[dead] = s_cnt[last]; s_lab[dead] = s_lab[last]; s_cnt[last] = 0;
                        s_logcnt[dead] = s_logcnt[last]; s_logcntm1[dead] = s_logcntm1[last];
                        #pragma unroll 1
                        for (int t = 0; t < K - 1; t++) if (s_order[t] == last) s_order[t] = dead;
                    }
                    if (c == last) c = dead;
                }
                K -= 1;
                pend_dead = -1;
                bar_d<ND>();
                if (d == 0) { s_KA = K; mbar_arrive(bar_go); }
            }
            pend_changed = false;
            DPHASE(1);
            KA_prev = K;
            if (bad) {                                        // block-uniform
                if (d == 0) {
                    raise_error(irm, KERR_BAD_MOVE);
                    if (args.out_choice) args.out_choice[m] = -1;
                    if (args.trace_cap > 0) args.trace_n[m] = 0;
                }
                continue;
            }

            DPHASE(11);
            // ---- candidates: CRP::tables_weights_gibbs (cxx/hirm.hh:147-161) -------------------------
            const bool singleton = s_cnt[c] == 1;
            const bool nofree = !singleton && K >= kt;         // no slot for the fresh-table candidate
            if (nofree && kt < kcap_g) {
                // ---- this tier's table is full: stop BEFORE the move; the next tier resumes here.  Move r+1 is
                //      already being accumulated: let it finish, then release the A warps once more so that they
                //      see the limit.
                r_done = r;
                if (d == 0) {
                    s_limit = r + 2;
                    if (r + 1 < nmoves) {
                        mbar_wait(&bar_ready[buf ^ 1], (uint32_t)(((rr + 1) >> 1) & 1));
                        if (r + 2 < nmoves) mbar_arrive(bar_go);
                    }
                }
                break;
            }
            const int ncand = K + ((singleton || nofree) ? 0 : 1);
            const int fresh_label = max(0, s_lab[s_order[K - 1]]) + 1;   // t_max starts at 0 (cxx/hirm.hh:139)

            // ---- score.  logp_gibbs_exact (cxx/hirm.hh:441-461): for every candidate t the sum over its 2K+1 target
            //      cells of S(N+n, s+sg) - S(N, s), with (N, s) the cell WITHOUT i's own observations.  The
            //      (candidate, cell) pairs are laid flat over the D lanes, ND per round; a lane evaluates one
            //      cell (branch-free, six independent log chains).  Candidate t's cells are added up in index order (a segmented
            //      reduction with a fixed order: deterministic) by one lane, the sums handed to every D warp through shared memory,
            //      so that each warp can go on to draw the table by itself.
            constexpr int MAXJ = 3;                            // candidates per lane: t = lane + 32 j (at most 64 tables + a fresh one)
            double lp[MAXJ];
#pragma unroll
            for (int j = 0; j < MAXJ; j++) lp[j] = 0.0;
            {
                // Lay-out of a round.  Tiers of up to 32 tables: the (candidate, cell) pairs flat over the lanes.
                // Last tier (BYC): QR consecutive cells of EVERY candidate (lane d: candidate d / QR, cell r QR + d mod QR), so that the
                // ordered sums of a round are QR additions per candidate, all candidates side by side -- with whole candidates per round
                // three lanes add 65 values each at K = 32 while 189 wait -- and a lane's candidate is fixed for the move.
                const int IPC = 2 * K + 1, total = ncand * IPC;
                constexpr bool BYCAND = BYC;
                const int QR = BYCAND ? ND / ncand : 1;        // >= 1: at most 65 candidates on 192 lanes
                const int t_mine = BYCAND ? d / QR : 0, q_mine = d - t_mine * QR;
                const int nround = BYCAND ? (IPC + QR - 1) / QR : (total + ND - 1) / ND;
                const int t_own = lane * DW + (d >> 5);        // the candidate this lane adds up (d = index among the D threads)
                double lp_own = 0.0;
                const cell_t rem_cc = pack_cell((uint32_t)(g[4 * c] + g[4 * c + 2] + d_n), (uint32_t)(g[4 * c + 1] + g[4 * c + 3] + d_s));
                #pragma unroll 1
                for (int r = 0; r < nround; r++) {
                    int t, q; bool valid;
                    if (BYCAND) { t = t_mine; q = r * QR + q_mine; valid = t < ncand && q < IPC; }
                    else { const int w = r * ND + d; valid = w < total; t = w / IPC; q = w - t * IPC; }
                    uint32_t N = 0u, sv = 0u, gn = 0u, gs = 0u;
                    if (valid) {                               // gather the cell's operands (cheap; lanes may diverge)
                        const int st = t < K ? s_order[t] : K;
                        cell_t cur = 0;
                        if (q == 2 * K) {                      // the (t, t) cell of the fresh slot: the diagonal only
                            if (st >= K) { gn = d_n; gs = d_s; }
                        } else {
                            const int k = q >> 1, col = q & 1;
                            if (k == st) {                     // merged target cell (t, t): row group + column group + diagonal
                                if (!col) {
                                    gn = g[4 * k] + g[4 * k + 2] + d_n; gs = g[4 * k + 1] + g[4 * k + 3] + d_s;
                                    cur = table[st + st * kt];
                                    if (st == c) cur -= rem_cc;
                                }
                            } else if (!col) {                 // (i, j), z_j = k  ->  cell (t, k)
                                gn = g[4 * k]; gs = g[4 * k + 1];
                                cur = table[st + k * kt];
                                if (st == c) cur -= pack_cell(gn, gs);
                                else if (k == c) cur -= pack_cell((uint32_t)g[4 * st + 2], (uint32_t)g[4 * st + 3]);   // i's column group with slot t sits in (t, c)
                            } else {                           // (j, i), z_j = k  ->  cell (k, t)
                                gn = g[4 * k + 2]; gs = g[4 * k + 3];
                                cur = table[k + st * kt];
                                if (st == c) cur -= pack_cell(gn, gs);
                                else if (k == c) cur -= pack_cell((uint32_t)g[4 * st], (uint32_t)g[4 * st + 1]);       // i's row group with slot t sits in (c, t)
                            }
                        }
                        N = cell_n(cur); sv = cell_s(cur);
                    }
                    __syncwarp();
                    s_val[vb * ND + d] = score_delta_bf<DN_LF>(s_lf, N, sv, gn, gs);
                    bar_d<ND>();
                    // candidate t's cells of this round, in index order: ONE lane per candidate (lane t / DW of D warp t mod DW: the
                    // candidates' sums run side by side on all D warps instead of six times over), values loaded four at a time
                    if (t_own < ncand) {
                        int w0, w1;
                        if (BYCAND) { w0 = t_own * QR; w1 = w0 + min(QR, IPC - r * QR); }
                        else { const int base = r * ND; w0 = max(t_own * IPC, base) - base; w1 = min((t_own + 1) * IPC, base + ND) - base; }
                        const double *v = s_val + vb * ND;
                        int ww = w0;
                        #pragma unroll 1
                        for (; ww + 4 <= w1; ww += 4) {
                            const double a0 = v[ww], a1 = v[ww + 1], a2 = v[ww + 2], a3 = v[ww + 3];
                            lp_own += a0; lp_own += a1; lp_own += a2; lp_own += a3;
                        }
                        #pragma unroll 1
                        for (; ww < w1; ww++) lp_own += v[ww];
                    }
                    vb ^= 1;
                }
                if (t_own < ncand) {
                    const int st = t_own < K ? s_order[t_own] : K;
                    s_lpbuf[t_own] = lp_own + (st == c ? (singleton ? logalpha : s_logcntm1[c]) : (t_own < K ? s_logcnt[st] : logalpha));
                }
                bar_d<ND>();                                   // every D warp reads all candidates' sums
#pragma unroll
                for (int j = 0; j < MAXJ; j++) {
                    const int t = lane + 32 * j;
                    if (t < ncand) lp[j] = s_lpbuf[t];
                }
            }
            DPHASE(2);

            // ---- sample: log_choice (cxx/util_math.cc:58-65); sums in index order.  Every D warp draws the (same)
            //      table by itself: registers and shuffles only, no barrier afterwards.
            int idx = 0;
            if (args.flags & 1u) {                             // HIRM_SWEEP_SCORE_ONLY: stay
#pragma unroll
                for (int j = 0; j < MAXJ; j++) {
                    const int t = lane + 32 * j;
                    if (t < ncand && (t < K ? s_order[t] : K) == c) idx = t;
                }
                idx = __reduce_max_sync(0xffffffffu, idx);
            } else if (ncand > 1) {
                double mx = -1.0 / 0.0;
#pragma unroll
                for (int j = 0; j < MAXJ; j++) if (lane + 32 * j < ncand) mx = fmax(mx, lp[j]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                double e[MAXJ];
#pragma unroll
                for (int j = 0; j < MAXJ; j++) {
                    e[j] = 0.0;
                    if (32 * j < ncand) e[j] = exp_nl(lane + 32 * j < ncand ? lp[j] - mx : -800.0);   // scale-invariant rule: one exp each; exp(-800) = 0
                }
                double total = 0.0, cum[MAXJ];                 // lane t: running sum up to t, in index order
#pragma unroll
                for (int j = 0; j < MAXJ; j++) {
                    cum[j] = 0.0;
                    if (32 * j < ncand) {
                        const int nt = min(32, ncand - 32 * j);
                        #pragma unroll 1
                        for (int t = 0; t < nt; t++) { total += __shfl_sync(0xffffffffu, e[j], t); if (t == lane) cum[j] = total; }
                    }
                }
                const double x = ctl.u * total;
                idx = ncand - 1;
#pragma unroll
                for (int j = MAXJ - 1; j >= 0; j--) {
                    if (32 * j < ncand) {
                        const unsigned hit = __ballot_sync(0xffffffffu, lane + 32 * j < ncand && cum[j] > x);
                        if (hit) idx = 32 * j + __ffs(hit) - 1;
                    }
                }
            }
            int tnew, newlabel;
            if (nofree || ncand == 0) { tnew = c; newlabel = s_lab[c]; }
            else { tnew = idx < K ? s_order[idx] : K; newlabel = idx < K ? s_lab[tnew] : fresh_label; }
            if (d == 0) {
                if (nofree) raise_error(irm, KERR_NO_FREE_SLOT);
                if (args.out_choice) args.out_choice[m] = newlabel;
                if (args.trace_cap > 0) args.trace_n[m] = ncand;
            }
            if (args.trace_cap > 0 && dwarp == 0) {
#pragma unroll
                for (int j = 0; j < MAXJ; j++) {
                    const int t = lane + 32 * j;
                    if (t < ncand && t < args.trace_cap) {
                        args.trace_labels[m * args.trace_cap + t] = t < K ? s_lab[s_order[t]] : fresh_label;
                        args.trace_logps[m * args.trace_cap + t] = lp[j];
                    }
                }
            }
            DPHASE(3);

            // ---- apply, only if the table changed ------------------------------------------------------
            if (tnew != c) {
                const bool died = singleton;
                bar_d<ND>();                                       // both D warps have drawn: the bookkeeping below may be rewritten
                #pragma unroll 1
                for (int k = d; k < K; k += ND) {          // i's observations leave row / column c
                    if (k != c) {
                        table[c + k * kt] -= pack_cell((uint32_t)g[4 * k], (uint32_t)g[4 * k + 1]);
                        table[k + c * kt] -= pack_cell((uint32_t)g[4 * k + 2], (uint32_t)g[4 * k + 3]);
                    } else {
                        table[c + c * kt] -= pack_cell((uint32_t)(g[4 * c] + g[4 * c + 2] + d_n), (uint32_t)(g[4 * c + 1] + g[4 * c + 3] + d_s));
                    }
                }
                if (d == ND - 1) {                          // Domain::set_cluster_assignment_gibbs (cxx/hirm.hh:219-224)
                    s_cnt[c] -= 1;
                    if (tnew == K) { s_cnt[K] = 1; s_lab[K] = newlabel; s_order[K] = K; }
                    else s_cnt[tnew] += 1;
                    if (died) {                                // CRP::unincorporate erases the empty table (:95-97)
                        int pos = 0;
                        #pragma unroll 1
                        for (int t = 0; t < K; t++) if (s_order[t] == c) pos = t;
                        #pragma unroll 1
                        for (int t = pos; t + 1 < K; t++) s_order[t] = s_order[t + 1];
                    }
                }
                bar_d<ND>();
                const int K2 = max(K, tnew + 1);
                #pragma unroll 1
                for (int k = d; k < K2; k += ND) {         // ... and enter row / column tnew
                    if (k != tnew) {
                        table[tnew + k * kt] += pack_cell((uint32_t)g[4 * k], (uint32_t)g[4 * k + 1]);
                        table[k + tnew * kt] += pack_cell((uint32_t)g[4 * k + 2], (uint32_t)g[4 * k + 3]);
                    } else {
                        table[tnew + tnew * kt] += pack_cell((uint32_t)(g[4 * k] + g[4 * k + 2] + d_n), (uint32_t)(g[4 * k + 1] + g[4 * k + 3] + d_s));
                    }
                }
                if (d >= ND - 4) {                          // cached log(count), log(count - 1) of the two tables
                    const int j = d - (ND - 4);
                    const int slot = (j & 2) ? tnew : c;
                    const int cnt = s_cnt[slot] - (j & 1);
                    const double v = log_nl((double)max(cnt, 1));
                    if (j & 1) s_logcntm1[slot] = v; else s_logcnt[slot] = v;
                }
                pend_changed = true; pend_item = item; pend_old = c; pend_new = tnew;
                if (tnew == K) K += 1;
                if (died) pend_dead = c;
                bar_d<ND>();                                       // the table and the cached logs are final before anyone scores again
            }
            DPHASE(4);
        }
        if (prof) for (int i = 0; i < 14; i++) args.phase_cycles[blockIdx.x * 32 + 8 + i] = s_ph[i];
#undef DPHASE
        // ---- finish the last move -----------------------------------------------------------------
        bar_d<ND>();
        if (pend_changed && d == 0) planes_set(zp, NP, pend_item, pend_new);
        bar_d<ND>();
        if (pend_dead >= 0) {
            const int dead = pend_dead, last = K - 1;
            if (dead != last) {
                planes_rename(zp, NP, G, last, dead, d, ND);
                #pragma unroll 1
                for (int k = d; k < K; k += ND) {
                    if (k != last && k != dead) {
                        table[dead + k * kt] = table[last + k * kt]; table[last + k * kt] = 0;
                        table[k + dead * kt] = table[k + last * kt]; table[k + last * kt] = 0;
                    } else if (k == last) {
                        table[dead + dead * kt] = table[last + last * kt]; table[last + last * kt] = 0;
                    }
                }
                if (d == 4) { s_cnt[dead] = s_cnt[last]; s_lab[dead] = s_lab[last]; s_cnt[last] = 0; }
            }
            K -= 1;
        }
        if (d == 0) { s_K = K; s_done = r_done; }
    }
    __syncthreads();

    // ---- bulk copies issued for moves that will not be made here must land before the CTA exits ----
    if (tid == 0) {
        #pragma unroll 1
        for (long long cidx = s_consumed; cidx < s_issued; cidx++)
            mbar_wait(&bar_full[cidx % NS], (uint32_t)((cidx / NS) & 1));
        if (args.progress) args.progress[blockIdx.x] = s_done;
    }
    __syncthreads();

    // ---- store state (slots 0..K-1, compact) ------------------------------------------------
    const int Kf = s_K;
    #pragma unroll 1
    for (int j = tid; j < n; j += NTHREADS) if (z_g[j] >= 0) z_g[j] = planes_get(zp, NP, j);
    const int span = max(hi_g, Kf);
    #pragma unroll 1
    for (int k = tid; k < span * span; k += NTHREADS) {
        const int a = k % span, b = k / span;
        R.counts[a + b * kcap_g] = (a < Kf && b < Kf) ? table[a + b * kt] : (cell_t)0;
    }
    #pragma unroll 1
    for (int k = tid; k < span; k += NTHREADS) {
        dd.tab_count[k] = k < Kf ? s_cnt[k] : 0;
        if (k < Kf) dd.tab_label[k] = s_lab[k];
    }
    if (tid == 0) *dd.hi = Kf;
}

// ---------------------------------------------------------------------------------------
// helpers for dense relations
// ---------------------------------------------------------------------------------------
// Count-table rebuild from slab0: cell (z0[i], z1[j]) += (1, x) for every observed (i, j)   (Relation::incorporate,
// cxx/hirm.hh:281-289: the table is the histogram of (observations, z)).  A CTA takes rows i = blockIdx.x, + gridDim.x, ...;
// the row is streamed with 16-byte loads and binned by z1[j] into per-thread private bins (n | s << 16 per slot, no
// atomics), the bins are folded into the CTA's copy of the table at the end of the row, the copy into the global table
// at the end.  Dynamic shared memory: zs1[n1 padded to 16] bytes, bins[k1 * RB_THREADS] words, tab[ncells] cells.
constexpr int RB_THREADS = 256;
__global__ void __launch_bounds__(RB_THREADS) rebuild_counts_dense_kernel(const IrmDev *irms, int irm_id, int rel) {
    const IrmDev &irm = irms[irm_id];
    const RelationDev *R = &irm.rel[rel];
    const DomainDev &d0 = irm.dom[R->dom[0]], &d1 = irm.dom[R->dom[1]];
    const int n0 = d0.n, n1 = d1.n, k1 = d1.kcap, n1pad = (n1 + 15) & ~15;
    const int64_t nc = R->ncells, cs0 = R->cstride[0], cs1 = R->cstride[1];
    extern __shared__ __align__(16) uint8_t rb_smem[];
    uint8_t *zs1 = rb_smem;
    uint32_t *bins = reinterpret_cast<uint32_t *>(rb_smem + n1pad);
    cell_t *tab = reinterpret_cast<cell_t *>(rb_smem + n1pad + (size_t)k1 * RB_THREADS * 4);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    bool absent = false;
    for (int j = tid; j < n1pad; j += RB_THREADS) { const int zj = j < n1 ? d1.z[j] : 0; zs1[j] = zj < 0 ? (uint8_t)0xFF : (uint8_t)zj; }
    for (int k = tid; k < k1 * RB_THREADS; k += RB_THREADS) bins[k] = 0u;
    for (int64_t k = tid; k < nc; k += RB_THREADS) tab[k] = 0;
    __syncthreads();
    uint32_t *mine = bins + tid;
    for (int i = blockIdx.x; i < n0; i += gridDim.x) {
        const int zi = d0.z[i];
        const uint8_t *row = R->slab[0] + (int64_t)i * R->pitch[0];
        for (int v = tid; v < (n1pad >> 4); v += RB_THREADS) {
            const uint4 xv = *reinterpret_cast<const uint4 *>(row + (size_t)v * 16);
            const uint4 zv = *reinterpret_cast<const uint4 *>(zs1 + (size_t)v * 16);
            const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w}, zw[4] = {zv.x, zv.y, zv.z, zv.w};
#pragma unroll
            for (int w = 0; w < 4; w++) {
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const uint32_t x = (xw[w] >> (8 * b)) & 0xFFu, slot = (zw[w] >> (8 * b)) & 0xFFu;
                    if ((x & 0x80u) || v * 16 + w * 4 + b >= n1) continue;        // missing / row padding
                    if (slot == 0xFFu || zi < 0) { absent = true; continue; }
                    mine[slot * RB_THREADS] += 1u | ((x & 1u) << 16);
                }
            }
        }
        __syncthreads();
        for (int k = warp; k < k1; k += RB_THREADS / 32) {         // fold the bins of slot k into cell (zi, k); this warp owns slot k
            uint32_t nn = 0, ss = 0;
            for (int t = lane; t < RB_THREADS; t += 32) { const uint32_t w = bins[k * RB_THREADS + t]; bins[k * RB_THREADS + t] = 0u; nn += w & 0xFFFFu; ss += w >> 16; }
            nn = __reduce_add_sync(0xffffffffu, nn); ss = __reduce_add_sync(0xffffffffu, ss);
            if (lane == 0 && nn && zi >= 0) tab[zi * cs0 + k * cs1] += pack_cell(nn, ss);
        }
        __syncthreads();
    }
    if (__syncthreads_or(absent) && tid == 0) raise_error(irm, KERR_ABSENT_ENTITY);
    for (int64_t k = tid; k < nc; k += RB_THREADS) if (tab[k]) atomicAdd(&R->counts[k], tab[k]);
}

// flag = 1 if any cell (i, j), i < rows, j < cols, is missing (byte >= 0x80)
__global__ void __launch_bounds__(256) dense_any_missing_kernel(const uint8_t *slab, int64_t rows, int64_t cols, int64_t pitch, int32_t *flag) {
    bool any = false;
    for (int64_t i = blockIdx.x; i < rows; i += gridDim.x) {
        const uint8_t *row = slab + i * pitch;
        for (int64_t j = threadIdx.x; j < cols; j += blockDim.x) any |= (row[j] & 0x80) != 0;
    }
    if (__syncthreads_or(any) && threadIdx.x == 0) atomicOr(flag, 1);
}

__global__ void transpose_u8_kernel(const uint8_t *in, int64_t rows, int64_t cols, int64_t pitch_in, uint8_t *out, int64_t pitch_out) {
    __shared__ uint8_t tile[64][65];
    const int64_t r0 = (int64_t)blockIdx.y * 64, c0 = (int64_t)blockIdx.x * 64;
    for (int y = threadIdx.y; y < 64; y += blockDim.y) {
        const int64_t r = r0 + y, c = c0 + threadIdx.x;
        tile[y][threadIdx.x] = (r < rows && c < cols) ? in[r * pitch_in + c] : (uint8_t)0xFF;
    }
    __syncthreads();
    for (int y = threadIdx.y; y < 64; y += blockDim.y) {
        const int64_t c = c0 + y, r = r0 + threadIdx.x;
        if (c < cols && r < rows) out[c * pitch_out + r] = tile[threadIdx.x][y];
    }
}

}  // namespace hirm
