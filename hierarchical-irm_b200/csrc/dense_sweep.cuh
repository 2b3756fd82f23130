// Dense fast path: an IRM with one domain and one dense binary relation R(e, e)
// (BASELINE.json config 3).  One persistent CTA per IRM / chain, resident across the
// IRM's whole move list; every chain streams its own rows of the SHARED read-only slabs.
//
// HBM traffic per move = row i of slab0 (the observations (i, j)) + row i of slab1 (the
// observations (j, i)) = 2 * n bytes: exactly the algorithmic bytes of SURVEY.md 8(d).
// Everything mutable (z as uint8 slots, the K x K count table, CRP table sizes) lives in
// shared memory for the whole launch.
//
//   stage   TMA 1-D bulk copies (cp.async.bulk + mbarrier complete_tx) fetch the two rows
//           of the moves NS-1 ahead: the move order is static, so the dependent chain
//           never waits on HBM latency.
//   A       grouping i's 2n observations by the other entity's table   -> get_cluster_to_items_list
//           (cxx/hirm.hh:388-397).  Up to 8 live tables: byte-SIMD one-hot masks of z and
//           dp4a dot products into registers (no shared-memory traffic at all).  More
//           tables: per-thread private bins hist[slot][tid] with 4 x 8-bit fields.
//   B       per-slot group sums, removed from the table in the same pass (the "rest" counts)
//   C       candidates + CRP prior: built by a dedicated warp WHILE the others accumulate
//   D       score: 6 lanes per target cell, one transcendental per lane (the move is bound
//           by fp64 dependent-issue latency, so the chain is kept one log deep)
//   E/F     sample (one exp per candidate) and apply.
// The per-move loop body is kept small (each transcendental path exists once, noinline):
// a single warp running cold code out of the instruction cache was the dominant cost of the
// first version (profiles/r01_notes.md).
#pragma once
#include "engine_types.cuh"

namespace hirm {

constexpr int DENSE_THREADS = 512;
constexpr int DENSE_WARPS = DENSE_THREADS / 32;
constexpr int DENSE_ACC_THREADS = DENSE_THREADS - 32;   // the last warp builds the candidates during A
constexpr int DENSE_ACC_WARPS = DENSE_ACC_THREADS / 32;
constexpr int DENSE_VEC_PER_THREAD = 15;   // 15 * 16 = 240 elements per thread per chunk (< 256: 8-bit fields)
constexpr int DENSE_MAX_KD = 8;            // dp4a path: live tables in slots 0..7

struct DenseLaunch {
    int32_t n;          // entities
    int32_t npad;       // n rounded up to 16
    int32_t kcap;
    int32_t stages;     // NS (<= 8)
    int32_t hist_words; // 32-bit words reserved for the private bins (>= 32 * kcap, >= 2 * kcap * kcap)
    // byte offsets into dynamic shared memory (host-computed, see dense_layout())
    int32_t off_table, off_lp, off_hist, off_ints, off_zs, off_stage, total_bytes;
};

// Shared-memory carve-up: 8-byte things first, then 4-byte, then the byte arrays (16-aligned).
__host__ __device__ inline DenseLaunch dense_layout(int n, int kcap, int stages, int hist_words) {
    DenseLaunch L;
    L.n = n; L.npad = (n + 15) & ~15; L.kcap = kcap; L.stages = stages; L.hist_words = hist_words;
    int off = 64;                                   // mbarriers
    L.off_table = off; off += kcap * kcap * 8;
    L.off_lp = off;    off += 2 * ((kcap + 2) & ~1) * 8;  // s_lp, s_tmp (unsorted log-weights / exp scratch)
    L.off_hist = off;  off += hist_words * 4;
    L.off_ints = off;  off += (11 * kcap + 16) * 4; // g_rn g_rs g_cn g_cs s_cnt s_lab s_cslot s_clabel u_slot u_label s_map
    off = (off + 15) & ~15;
    L.off_zs = off;    off += L.npad;
    L.off_stage = off; off += stages * 2 * L.npad;
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

// nibble code (bit0 valid_row, bit1 x_row, bit2 valid_col, bit3 x_col) -> one increment per 8-bit field
__device__ __forceinline__ uint32_t code_to_inc(uint32_t code) { return (code * 0x00204081u) & 0x01010101u; }

// four elements at once: per byte, the nibble code of (row byte, col byte); a byte is valid iff < 0x80
__device__ __forceinline__ uint32_t pack_codes(uint32_t xr, uint32_t xc) {
    const uint32_t vr = (~xr >> 7) & 0x01010101u, vc = (~xc >> 7) & 0x01010101u;
    return vr | ((xr & vr) << 1) | (vc << 2) | ((xc & vc) << 3);
}

// One copy of each transcendental path (see header comment).
__device__ __noinline__ double lfact_half_nl(int64_t a, int64_t m, int half) { return lfact_half(a, m, half); }
__device__ __noinline__ double log_nl(double x) { return log(x); }
__device__ __noinline__ double exp_nl(double x) { return exp(x); }

// ---- A (dp4a variant): one-hot byte masks of the slot word, dot products into registers --------
template <int KD>
__device__ __forceinline__ void acc_word(uint32_t xr, uint32_t xc, uint32_t z, int (&rn)[KD], int (&rs)[KD], int (&cn)[KD], int (&cs)[KD]) {
    constexpr uint32_t ONES = 0x01010101u;
    const uint32_t vr = (~xr >> 7) & ONES, vc = (~xc >> 7) & ONES;   // observed?
    const uint32_t sr = xr & vr, sc = xc & vc;                        // observed and 1
    const uint32_t b0 = z & ONES, b1 = (z >> 1) & ONES, b2 = (z >> 2) & ONES;
#pragma unroll
    for (int k = 0; k < KD; k++) {
        uint32_t mk = (k & 1) ? b0 : (b0 ^ ONES);
        if (KD > 2) mk &= (k & 2) ? b1 : (b1 ^ ONES);
        if (KD > 4) mk &= (k & 4) ? b2 : (b2 ^ ONES);
        rn[k] = (int)__dp4a(mk, vr, (unsigned)rn[k]); rs[k] = (int)__dp4a(mk, sr, (unsigned)rs[k]);
        cn[k] = (int)__dp4a(mk, vc, (unsigned)cn[k]); cs[k] = (int)__dp4a(mk, sc, (unsigned)cs[k]);
    }
}

template <int KD>
__device__ __forceinline__ void accumulate_dp4a(const uint8_t *xrow, const uint8_t *xcol, const uint8_t *zs, int n, int tid,
                                                int warp, int lane, int32_t *wpart /* [ACC_WARPS][4*DENSE_MAX_KD] */) {
    int rn[KD], rs[KD], cn[KD], cs[KD];
#pragma unroll
    for (int k = 0; k < KD; k++) { rn[k] = 0; rs[k] = 0; cn[k] = 0; cs[k] = 0; }
    const int nvec = n >> 4;
    #pragma unroll 1
    for (int v = tid; v < nvec; v += DENSE_ACC_THREADS) {
        const uint4 xr = *reinterpret_cast<const uint4 *>(xrow + (size_t)v * 16);
        const uint4 xc = *reinterpret_cast<const uint4 *>(xcol + (size_t)v * 16);
        const uint4 zz = *reinterpret_cast<const uint4 *>(zs + (size_t)v * 16);
        acc_word<KD>(xr.x, xc.x, zz.x, rn, rs, cn, cs);
        acc_word<KD>(xr.y, xc.y, zz.y, rn, rs, cn, cs);
        acc_word<KD>(xr.z, xc.z, zz.z, rn, rs, cn, cs);
        acc_word<KD>(xr.w, xc.w, zz.w, rn, rs, cn, cs);
    }
    const int j = (nvec << 4) + tid;                  // scalar tail, one element per thread; other bytes masked off as missing
    if (j < n) acc_word<KD>(xrow[j] | 0xFFFFFF00u, xcol[j] | 0xFFFFFF00u, zs[j], rn, rs, cn, cs);
#pragma unroll
    for (int k = 0; k < KD; k++) {
        const int a = __reduce_add_sync(0xffffffffu, rn[k]), b = __reduce_add_sync(0xffffffffu, rs[k]);
        const int c2 = __reduce_add_sync(0xffffffffu, cn[k]), d = __reduce_add_sync(0xffffffffu, cs[k]);
        if (lane == 0) {
            int32_t *p = wpart + warp * (4 * DENSE_MAX_KD) + 4 * k;
            p[0] = a; p[1] = b; p[2] = c2; p[3] = d;
        }
    }
}

__global__ void __launch_bounds__(DENSE_THREADS, 1) dense_ree_sweep_kernel(SweepArgs args, const int32_t *blk_list, DenseLaunch L) {
    const int k_list = blk_list ? blk_list[blockIdx.x] : blockIdx.x;
    const IrmDev &irm = args.irms[args.irm_ids[k_list]];
    const DomainDev &dd = irm.dom[0];
    const RelationDev &R = irm.rel[0];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = L.n, npad = L.npad, kcap = L.kcap, NS = L.stages;

    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);                        // [NS]
    cell_t *table = reinterpret_cast<cell_t *>(smem + L.off_table);             // [kcap*kcap]  Relation::clusters
    double *s_lp = reinterpret_cast<double *>(smem + L.off_lp);                 // [kcap+1]
    double *s_tmp = s_lp + ((kcap + 2) & ~1);
    uint32_t *hist = reinterpret_cast<uint32_t *>(smem + L.off_hist);           // [hist_words] private bins
    int32_t *g_rn = reinterpret_cast<int32_t *>(smem + L.off_ints);             // group sums per other-slot
    int32_t *g_rs = g_rn + kcap, *g_cn = g_rs + kcap, *g_cs = g_cn + kcap;
    int32_t *s_cnt = g_cs + kcap, *s_lab = s_cnt + kcap;                        // CRP table sizes / labels
    int32_t *s_cslot = s_lab + kcap, *s_clabel = s_cslot + (kcap + 1);          // candidates (ascending label)
    int32_t *u_slot = s_clabel + (kcap + 1), *u_label = u_slot + (kcap + 1);    // candidates before sorting
    int32_t *s_map = u_label + (kcap + 1);                                      // slot compaction map
    uint8_t *zs = smem + L.off_zs;                                              // [npad] slot of every entity
    uint8_t *stage0 = smem + L.off_stage;                                       // [NS][2][npad] staged rows
    __shared__ int32_t s_ncand, s_choice_slot, s_choice_label, s_fail, s_hi, s_holes;
    __shared__ double s_u, s_logalpha;
    __shared__ double s_part[DENSE_WARPS];
    __shared__ int32_t s_wpart[DENSE_ACC_WARPS * 4 * DENSE_MAX_KD];

    const int64_t m0 = args.move_off[k_list], m1 = args.move_off[k_list + 1];
    const int64_t nmoves = m1 - m0;
    const uint32_t row_bytes = (uint32_t)npad;
    // optional per-phase cycle accounting (thread 0 only; args.phase_cycles is null in production)
    long long ph[16] = {0}, tlast = 0;
    const bool prof = args.phase_cycles != nullptr && tid == 0;
    // BAR.SYNC does not block at issue: a clock read right after __syncthreads() would record thread 0's
    // ARRIVAL.  In profiling mode a second barrier (block-uniform condition) makes the stamp the release time.
    const bool prof_any = args.phase_cycles != nullptr;
#define PHASE(i) do { if (prof) { long long _t = clock64(); ph[i] += _t - tlast; tlast = _t; } } while (0)
#define SYNC_PHASE(i) do { __syncthreads(); if (prof_any) { __syncthreads(); PHASE(i); } } while (0)

    // ---- load state -------------------------------------------------------------------
    #pragma unroll 1
    for (int j = tid; j < npad; j += DENSE_THREADS) zs[j] = j < n ? (uint8_t)dd.z[j] : (uint8_t)0;
    #pragma unroll 1
    for (int k = tid; k < kcap * kcap; k += DENSE_THREADS) table[k] = R.counts[k];
    #pragma unroll 1
    for (int k = tid; k < kcap; k += DENSE_THREADS) { s_cnt[k] = dd.tab_count[k]; s_lab[k] = dd.tab_label[k]; }
    #pragma unroll 1
    for (int k = tid; k < L.hist_words; k += DENSE_THREADS) hist[k] = 0u;
    if (tid == 0) {
        s_hi = *dd.hi; s_holes = 1;                  // unknown: check once
        s_logalpha = log_nl(dd.alpha);
        #pragma unroll 1
        for (int s = 0; s < NS; s++) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        #pragma unroll 1
        for (int s = 0; s < NS && s < nmoves; s++) {
            const int it = args.move_item[m0 + s];
            const int64_t row = (it >= 0 && it < n) ? it : 0;
            mbar_expect_tx(&bars[s], 2 * row_bytes);
            tma_load_1d(stage0 + (size_t)s * 2 * npad, R.slab[0] + row * R.pitch[0], row_bytes, &bars[s]);
            tma_load_1d(stage0 + (size_t)s * 2 * npad + npad, R.slab[1] + row * R.pitch[1], row_bytes, &bars[s]);
        }
    }
    int stage_ctr = 0; uint32_t parity_ctr = 0;
    int item_next = nmoves > 0 ? args.move_item[m0] : 0;
    int dom_next = nmoves > 0 ? args.move_dom[m0] : 0;

    #pragma unroll 1
    for (int64_t r = 0; r < nmoves; r++) {
        const int64_t m = m0 + r;
        const int item = item_next;
        const bool bad = item < 0 || item >= n || dom_next != 0;
        // move ids are fetched one move ahead (and the TMA producer's NS ahead): never on the dependent chain
        if (r + 1 < nmoves) { item_next = args.move_item[m + 1]; dom_next = args.move_dom[m + 1]; }
        int item_tma = 0;
        if (tid == 0 && r + NS < nmoves) item_tma = args.move_item[m + NS];
        const int stage = stage_ctr;
        const uint32_t parity = parity_ctr;
        if (++stage_ctr == NS) { stage_ctr = 0; parity_ctr ^= 1u; }
        const uint8_t *xrow = stage0 + (size_t)stage * 2 * npad, *xcol = xrow + npad;

        // ---- slot compaction: keep live tables in slots 0..ntab-1 so that the dp4a path applies -------
        const int holes = s_holes;                     // every thread reads it before thread 0 may rewrite it (barrier below)
        if (holes) {
            __syncthreads();
            if (tid == 0) {
                int nt = 0;
                #pragma unroll 1
                for (int k = 0; k < s_hi; k++) { s_map[k] = nt; if (s_cnt[k] > 0) nt++; }
                s_holes = nt != s_hi ? 2 : 0;
            }
            __syncthreads();
            if (s_holes == 2) {                       // block-uniform
                const int hi0 = s_hi;
                cell_t *tmp = reinterpret_cast<cell_t *>(hist);          // hist is all-zero between moves; restored below
                #pragma unroll 1
                for (int k = tid; k < hi0 * hi0; k += DENSE_THREADS) { const int a = k % hi0, b = k / hi0; tmp[k] = table[a + b * kcap]; }
                #pragma unroll 1
                for (int j = tid; j < n; j += DENSE_THREADS) zs[j] = (uint8_t)s_map[zs[j]];
                __syncthreads();
                #pragma unroll 1
                for (int k = tid; k < hi0 * hi0; k += DENSE_THREADS) { const int a = k % hi0, b = k / hi0; table[a + b * kcap] = 0; }
                __syncthreads();
                #pragma unroll 1
                for (int k = tid; k < hi0 * hi0; k += DENSE_THREADS) {
                    const int a = k % hi0, b = k / hi0;
                    if (s_cnt[a] > 0 && s_cnt[b] > 0) table[s_map[a] + s_map[b] * kcap] = tmp[k];
                }
                __syncthreads();
                #pragma unroll 1
                for (int k = tid; k < 2 * hi0 * hi0; k += DENSE_THREADS) hist[k] = 0u;
                if (tid == 0) {
                    int nt = 0;
                    #pragma unroll 1
                    for (int k = 0; k < hi0; k++) if (s_cnt[k] > 0) { s_cnt[nt] = s_cnt[k]; s_lab[nt] = s_lab[k]; nt++; }
                    #pragma unroll 1
                    for (int k = nt; k < hi0; k++) s_cnt[k] = 0;
                    s_hi = nt; s_holes = 0;
                }
                __syncthreads();
            }
        }
        const int c = bad ? 0 : zs[item];
        const int hi = s_hi;
        if (tid == 0) s_fail = 0;
        if (tid == DENSE_THREADS - 1)                  // this move's uniform, off the critical path
            s_u = args.uniforms ? args.uniforms[m] : philox_uniform(args.seed, args.stream[k_list], args.counter0[k_list] + (uint64_t)r);

        if (prof) tlast = clock64();
        mbar_wait(&bars[stage], parity);
        PHASE(0);
        // the diagonal observation (i, i) sits in both staged rows at j = i; it is ONE observation
        const uint32_t dcode = bad ? 0u : (pack_codes(xrow[item], xcol[item]) & 0xFu);
        const int d_n = dcode & 1, d_s = (dcode >> 1) & 1;
        const bool use_dp4a = hi <= DENSE_MAX_KD;

        if (warp == DENSE_WARPS - 1) {
            // ---- C: candidates, concurrently with the accumulation done by the other warps ---------
            //      CRP::tables_weights_gibbs (cxx/hirm.hh:147-161)
            int maxlab = 0, freeslot = 1 << 30;        // t_max starts at 0 (cxx/hirm.hh:139)
            #pragma unroll 1
            for (int k = lane; k < hi; k += 32) {
                if (s_cnt[k] > 0) maxlab = max(maxlab, s_lab[k]); else freeslot = min(freeslot, k);
            }
            maxlab = __reduce_max_sync(0xffffffffu, maxlab);
            freeslot = __reduce_min_sync(0xffffffffu, freeslot);
            if (freeslot == (1 << 30)) freeslot = hi < kcap ? hi : -1;
            const bool singleton = s_cnt[c] == 1;
            int ncount = 0;
            #pragma unroll 1
            for (int base = 0; base < hi; base += 32) {
                const int k = base + lane;
                const int cnt = k < hi ? s_cnt[k] - (k == c ? 1 : 0) : 0;
                const unsigned bal = __ballot_sync(0xffffffffu, cnt > 0);
                const int pos = ncount + __popc(bal & ((1u << lane) - 1u));
                if (cnt > 0) { u_slot[pos] = k; u_label[pos] = s_lab[k]; }
                const double lw = log_nl((double)max(cnt, 1));
                if (cnt > 0) s_tmp[pos] = lw;
                ncount += __popc(bal);
            }
            if (lane == 0) {
                if (singleton) { u_slot[ncount] = c; u_label[ncount] = s_lab[c]; s_tmp[ncount] = s_logalpha; }
                else if (freeslot >= 0) { u_slot[ncount] = freeslot; u_label[ncount] = maxlab + 1; s_tmp[ncount] = s_logalpha; }
                else { raise_error(irm, KERR_NO_FREE_SLOT); s_fail = 1; }
            }
            const int nc = ncount + ((singleton || freeslot >= 0) ? 1 : 0);
            __syncwarp();
            #pragma unroll 1
            for (int j = lane; j < nc; j += 32) {      // ascending label order: rank by counting (labels are distinct)
                const int lb = u_label[j];
                int rank = 0;
                #pragma unroll 1
                for (int i = 0; i < nc; i++) rank += u_label[i] < lb;
                s_cslot[rank] = u_slot[j]; s_clabel[rank] = lb; s_lp[rank] = s_tmp[j];
            }
            if (lane == 0) s_ncand = nc;
        } else if (use_dp4a) {
            // ---- A (<= 8 live tables): register accumulation ------------------------------------------
            if (hi <= 2) accumulate_dp4a<2>(xrow, xcol, zs, n, tid, warp, lane, s_wpart);
            else if (hi <= 4) accumulate_dp4a<4>(xrow, xcol, zs, n, tid, warp, lane, s_wpart);
            else accumulate_dp4a<8>(xrow, xcol, zs, n, tid, warp, lane, s_wpart);
        }
        if (use_dp4a) {
            SYNC_PHASE(1);
            // ---- B: per-warp partials -> group sums; thread k owns slot k and removes its cells -------
            if (tid < hi) {
                const int k = tid;
                int rn = 0, rs = 0, cn = 0, cs = 0;
                #pragma unroll 1
                for (int w = 0; w < DENSE_ACC_WARPS; w++) {
                    const int32_t *p = s_wpart + w * (4 * DENSE_MAX_KD) + 4 * k;
                    rn += p[0]; rs += p[1]; cn += p[2]; cs += p[3];
                }
                if (k == c) { rn -= d_n; rs -= d_s; cn -= d_n; cs -= d_s; }       // drop the diagonal from both groups
                g_rn[k] = rn; g_rs[k] = rs; g_cn[k] = cn; g_cs[k] = cs;
                if (!bad) {
                    if (k != c) {
                        table[c + k * kcap] -= pack_cell(rn, rs);
                        table[k + c * kcap] -= pack_cell(cn, cs);
                    } else {
                        table[c + c * kcap] -= pack_cell(rn + cn + d_n, rs + cs + d_s);
                    }
                }
            }
            SYNC_PHASE(2);
        } else {
            // ---- A (> 8 live tables): per-thread private bins, diagonal included, pulled out in B -------
            const int HT = min(DENSE_ACC_THREADS, (L.hist_words / max(hi, 1)) & ~31);
            const int nvec = n >> 4;
            const int chunk_vecs = HT * DENSE_VEC_PER_THREAD;
            #pragma unroll 1
            for (int v0 = 0; v0 < nvec || v0 == 0; v0 += chunk_vecs) {
                if (tid < HT) {
                    const int vend = min(nvec, v0 + chunk_vecs);
                    uint32_t *bins = hist + tid;
                    #pragma unroll 1
                    for (int v = v0 + tid; v < vend; v += HT) {
                        const uint4 xr = *reinterpret_cast<const uint4 *>(xrow + (size_t)v * 16);
                        const uint4 xc = *reinterpret_cast<const uint4 *>(xcol + (size_t)v * 16);
                        const uint4 zz = *reinterpret_cast<const uint4 *>(zs + (size_t)v * 16);
                        const uint32_t xrw[4] = {xr.x, xr.y, xr.z, xr.w}, xcw[4] = {xc.x, xc.y, xc.z, xc.w}, zw[4] = {zz.x, zz.y, zz.z, zz.w};
#pragma unroll
                        for (int w = 0; w < 4; w++) {
                            const uint32_t codes = pack_codes(xrw[w], xcw[w]);
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                const uint32_t slot = (zw[w] >> (8 * e)) & 0xFFu;
                                bins[slot * HT] += code_to_inc((codes >> (8 * e)) & 0xFu);
                            }
                        }
                    }
                    if (v0 == 0) {                         // scalar tail: elements nvec*16 .. n-1, one per thread
                        const int j = (nvec << 4) + tid;
                        if (j < n) bins[(uint32_t)zs[j] * HT] += code_to_inc(pack_codes(xrow[j], xcol[j]) & 0xFu);
                    }
                }
                SYNC_PHASE(1);
                // ---- B: bins -> group sums per other-entity slot (accumulated over chunks) ------
                #pragma unroll 1
                for (int k = warp; k < hi; k += DENSE_WARPS) {
                    int rn = 0, rs = 0, cn = 0, cs = 0;
                    #pragma unroll 1
                    for (int j = lane; j < HT; j += 32) {
                        const uint32_t wv = hist[k * HT + j];
                        hist[k * HT + j] = 0u;
                        rn += wv & 0xFFu; rs += (wv >> 8) & 0xFFu; cn += (wv >> 16) & 0xFFu; cs += wv >> 24;
                    }
                    rn = __reduce_add_sync(0xffffffffu, rn); rs = __reduce_add_sync(0xffffffffu, rs);
                    cn = __reduce_add_sync(0xffffffffu, cn); cs = __reduce_add_sync(0xffffffffu, cs);
                    if (lane == 0) {
                        if (v0 == 0) {
                            if (k == c) { rn -= d_n; rs -= d_s; cn -= d_n; cs -= d_s; }   // drop the diagonal from both groups
                        } else { rn += g_rn[k]; rs += g_rs[k]; cn += g_cn[k]; cs += g_cs[k]; }
                        g_rn[k] = rn; g_rs[k] = rs; g_cn[k] = cn; g_cs[k] = cs;
                        // g_r[k] = (n, s) of observations (i, j), j != i, z_j = k;  g_c[k] likewise for (j, i);
                        // (d_n, d_s) = (i, i).  Last chunk: remove them from the table ("rest" counts) right here --
                        // this lane is the only one touching cells (c, k) and (k, c).
                        if (v0 + chunk_vecs >= nvec && !bad) {
                            if (k != c) {
                                table[c + k * kcap] -= pack_cell(rn, rs);
                                table[k + c * kcap] -= pack_cell(cn, cs);
                            } else {
                                table[c + c * kcap] -= pack_cell(rn + cn + d_n, rs + cs + d_s);
                            }
                        }
                    }
                }
                SYNC_PHASE(2);
            }
        }
        // stage buffer is free again: prefetch the rows of move r + NS
        if (tid == 0 && r + NS < nmoves) {
            const int64_t row = (item_tma >= 0 && item_tma < n) ? item_tma : 0;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&bars[stage], 2 * row_bytes);
            tma_load_1d(stage0 + (size_t)stage * 2 * npad, R.slab[0] + row * R.pitch[0], row_bytes, &bars[stage]);
            tma_load_1d(stage0 + (size_t)stage * 2 * npad + npad, R.slab[1] + row * R.pitch[1], row_bytes, &bars[stage]);
        }
        if (bad) {                                    // block-uniform
            if (tid == 0) {
                raise_error(irm, KERR_BAD_MOVE);
                if (args.out_choice) args.out_choice[m] = -1;
                if (args.trace_cap > 0) args.trace_n[m] = 0;
            }
            continue;
        }
        PHASE(3);
        const int ncand = s_ncand;

        // ---- D: score.  logp_gibbs_exact (cxx/hirm.hh:441-461): for every candidate t the sum over target
        //      cells of S(N+n, s+sg) - S(N, s).  Each difference is three log-factorial differences and each
        //      of those two halves of at most one transcendental: 6 lanes per target cell, so the dependent
        //      fp64 chain of a move is ONE log deep.  A candidate is owned by a team of `wpc` whole warps;
        //      lane partials are combined in a fixed order (deterministic log-probabilities).
        {
            const int nq = 2 * hi, items = 6 * nq + 6;      // +6: the (t,t) cell of a fresh slot past hi (diagonal only)
            int wpc = ncand <= DENSE_WARPS ? DENSE_WARPS / ncand : 1;
            wpc = max(1, min(wpc, (items + 31) / 32));
            const int rounds = ncand <= DENSE_WARPS ? 1 : (ncand + DENSE_WARPS - 1) / DENSE_WARPS;
            #pragma unroll 1
            for (int rd = 0; rd < rounds; rd++) {
                const int t = rounds == 1 ? warp / wpc : rd * DENSE_WARPS + warp;
                const int j = rounds == 1 ? warp - t * wpc : 0;
                double acc = 0.0;
                if (t < ncand) {
                    const int st = s_cslot[t];
                    #pragma unroll 1
                    for (int w = j * 32 + lane; w < items; w += wpc * 32) {
                        const int q = w / 6, part = w - q * 6;
                        const int term = part >> 1, half = part & 1;
                        const int k = q >> 1, col = q & 1;
                        int gn, gs; cell_t cur;
                        if (q == nq) {                     // fresh slot beyond hi: only the diagonal lands in (t, t)
                            if (st < hi) continue;
                            gn = d_n; gs = d_s; cur = 0;
                        } else if (k == st) {              // merged target cell (t, t): row group + column group + diagonal
                            if (col) continue;
                            gn = g_rn[k] + g_cn[k] + d_n; gs = g_rs[k] + g_cs[k] + d_s;
                            cur = table[st + st * kcap];
                        } else {
                            gn = col ? g_cn[k] : g_rn[k]; gs = col ? g_cs[k] : g_rs[k];
                            cur = col ? table[k + st * kcap] : table[st + k * kcap];
                        }
                        if (gn == 0) continue;
                        const int64_t N = cell_n(cur), sv = cell_s(cur);
                        // S(N+n, s+sg) - S(N, s) = D(s, sg) + D(N-s, n-sg) - D(N+1, n),  D(a, m) = log((a+m)!/a!)
                        const int64_t a = term == 0 ? sv : term == 1 ? N - sv : N + 1;
                        const int64_t mm = term == 0 ? gs : term == 1 ? gn - gs : gn;
                        const double v = lfact_half_nl(a, mm, half);
                        acc += term == 2 ? -v : v;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (rounds == 1) {
                    if (lane == 0) s_part[warp] = acc;
                } else if (lane == 0 && t < ncand) s_lp[t] += acc;
            }
            PHASE(8);
            if (rounds == 1) {
                SYNC_PHASE(9);
                if (tid < ncand) { double a = 0.0; for (int j = 0; j < wpc; j++) a += s_part[tid * wpc + j]; s_lp[tid] += a; }
            }
        }
        SYNC_PHASE(4);

        // ---- E: sample (warp 0): log_choice (cxx/util_math.cc:58-65); exp() lane-parallel, sums in index order ----
        if (warp == 0) {
            int idx = 0;
            if (args.flags & 1u) {                     // HIRM_SWEEP_SCORE_ONLY: stay
                #pragma unroll 1
                for (int t = lane; t < ncand; t += 32) if (s_cslot[t] == c) idx = t;
                idx = __reduce_max_sync(0xffffffffu, idx);
            } else if (ncand > 1) {
                double mx = -1.0 / 0.0;
                #pragma unroll 1
                for (int t = lane; t < ncand; t += 32) mx = fmax(mx, s_lp[t]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                PHASE(10);
                #pragma unroll 1
                for (int t = lane; t < ncand; t += 32) s_tmp[t] = exp_nl(s_lp[t] - mx);   // scale-invariant rule: one exp each
                __syncwarp();
                PHASE(11);
                if (lane == 0) {
                    double total = 0.0;
                    #pragma unroll 1
                    for (int t = 0; t < ncand; t++) total += s_tmp[t];
                    const double x = s_u * total;
                    double cum = 0.0;
                    idx = ncand - 1;
                    #pragma unroll 1
                    for (int t = 0; t < ncand; t++) { cum += s_tmp[t]; if (cum > x) { idx = t; break; } }
                }
                idx = __shfl_sync(0xffffffffu, idx, 0);
                PHASE(12);
            }
            if (lane == 0) {
                if (s_fail || ncand == 0) { s_choice_slot = c; s_choice_label = s_lab[c]; }
                else { s_choice_slot = s_cslot[idx]; s_choice_label = s_clabel[idx]; }
                if (args.out_choice) args.out_choice[m] = s_choice_label;
                if (args.trace_cap > 0) args.trace_n[m] = ncand;
            }
        }
        if (args.trace_cap > 0) {
            #pragma unroll 1
            for (int t = tid; t < ncand && t < args.trace_cap; t += DENSE_THREADS) {
                args.trace_labels[m * args.trace_cap + t] = s_clabel[t];
                args.trace_logps[m * args.trace_cap + t] = s_lp[t];
            }
        }
        SYNC_PHASE(5);

        // ---- F: apply: add the groups back at the chosen slot -------------------------------
        const int tnew = s_choice_slot;
        if (tid < hi) {
            const int k = tid;                         // groups are keyed by the OTHER entity's slot, which does not move
            if (k != tnew) {
                table[tnew + k * kcap] += pack_cell(g_rn[k], g_rs[k]);
                table[k + tnew * kcap] += pack_cell(g_cn[k], g_cs[k]);
            }
        }
        if (tid == DENSE_THREADS - 32) {
            int gn = d_n, gs = d_s;
            if (tnew < hi) { gn += g_rn[tnew] + g_cn[tnew]; gs += g_rs[tnew] + g_cs[tnew]; }
            table[tnew + tnew * kcap] += pack_cell(gn, gs);
        }
        if (tid == 0 && tnew != c) {
            zs[item] = (uint8_t)tnew;
            s_cnt[c] -= 1;
            const int newcnt = (tnew < hi ? s_cnt[tnew] : 0) + 1;
            s_cnt[tnew] = newcnt;
            if (newcnt == 1) s_lab[tnew] = s_choice_label;
            int nhi = max(hi, tnew + 1);
            while (nhi > 0 && s_cnt[nhi - 1] == 0) nhi--;
            s_hi = nhi;
            if (s_cnt[c] == 0 && c < nhi) s_holes = 1;   // a table died below the top: compact before the next move
        }
        SYNC_PHASE(6);
    }
    if (prof) for (int i = 0; i < 16; i++) args.phase_cycles[blockIdx.x * 16 + i] = ph[i];
#undef PHASE
#undef SYNC_PHASE

    // ---- store state ------------------------------------------------------------------
    #pragma unroll 1
    for (int j = tid; j < n; j += DENSE_THREADS) dd.z[j] = zs[j];
    #pragma unroll 1
    for (int k = tid; k < kcap * kcap; k += DENSE_THREADS) R.counts[k] = table[k];
    #pragma unroll 1
    for (int k = tid; k < kcap; k += DENSE_THREADS) { dd.tab_count[k] = s_cnt[k]; dd.tab_label[k] = s_lab[k]; }
    if (tid == 0) *dd.hi = s_hi;
}

// ---------------------------------------------------------------------------------------
// helpers for dense relations
// ---------------------------------------------------------------------------------------
// count-table rebuild from slab0: cell (z0[i], z1[j]) += (1, x) for every observed (i, j)
__global__ void __launch_bounds__(256) rebuild_counts_dense_kernel(const IrmDev *irms, int irm_id, int rel) {
    const IrmDev &irm = irms[irm_id];
    const RelationDev *R = &irm.rel[rel];
    const DomainDev &d0 = irm.dom[R->dom[0]], &d1 = irm.dom[R->dom[1]];
    extern __shared__ cell_t s_tab[];
    const int64_t nc = R->ncells;
    for (int64_t k = threadIdx.x; k < nc; k += blockDim.x) s_tab[k] = 0;
    __syncthreads();
    for (int i = blockIdx.x; i < d0.n; i += gridDim.x) {
        const int zi = d0.z[i];
        const uint8_t *row = R->slab[0] + (int64_t)i * R->pitch[0];
        for (int j = threadIdx.x; j < d1.n; j += blockDim.x) {
            const uint8_t x = row[j];
            if (x & 0x80) continue;
            const int zj = d1.z[j];
            if (zi < 0 || zj < 0) { raise_error(irm, KERR_ABSENT_ENTITY); continue; }
            atomicAdd(&s_tab[zi * R->cstride[0] + zj * R->cstride[1]], pack_cell(1u, x & 1u));
        }
    }
    __syncthreads();
    for (int64_t k = threadIdx.x; k < nc; k += blockDim.x) if (s_tab[k]) atomicAdd(&R->counts[k], s_tab[k]);
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
