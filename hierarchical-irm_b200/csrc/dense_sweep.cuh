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
//   A       every thread owns private bins hist[slot][tid] (4 x 8-bit fields: n_row, s_row,
//           n_col, s_col), so grouping i's 2n observations by the other entity's table is
//           plain LDS/IADD/STS, no atomics, no bank conflicts   -> get_cluster_to_items_list
//   B       bins -> per-slot group sums (warp REDUX), removed from the table in the same
//           pass (the "rest" counts)                             -> logp_gibbs_exact_current
//   C..F    candidates, score (one warp per candidate), sample, apply: as in the generic kernel.
#pragma once
#include "engine_types.cuh"

namespace hirm {

constexpr int DENSE_THREADS = 512;
constexpr int DENSE_WARPS = DENSE_THREADS / 32;
constexpr int DENSE_MAXK = 256;
constexpr int DENSE_VEC_PER_THREAD = 15;   // 15 * 16 = 240 elements per thread per chunk (< 256: 8-bit fields)

struct DenseLaunch {
    int32_t n;          // entities
    int32_t npad;       // n rounded up to 16
    int32_t kcap;
    int32_t stages;     // NS (<= 8)
    int32_t hist_words; // 32-bit words reserved for the private bins (>= 32 * kcap)
    // byte offsets into dynamic shared memory (host-computed, see dense_layout())
    int32_t off_table, off_lp, off_hist, off_ints, off_zs, off_stage, total_bytes;
};

// Shared-memory carve-up: 8-byte things first, then 4-byte, then the byte arrays (16-aligned).
__host__ __device__ inline DenseLaunch dense_layout(int n, int kcap, int stages, int hist_words) {
    DenseLaunch L;
    L.n = n; L.npad = (n + 15) & ~15; L.kcap = kcap; L.stages = stages; L.hist_words = hist_words;
    int off = 64;                                   // mbarriers
    L.off_table = off; off += kcap * kcap * 8;
    L.off_lp = off;    off += ((kcap + 2) & ~1) * 8;
    L.off_hist = off;  off += hist_words * 4;
    L.off_ints = off;  off += (8 * kcap + 8) * 4;   // g_rn g_rs g_cn g_cs s_cnt s_lab s_cslot s_clabel
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
    uint32_t *hist = reinterpret_cast<uint32_t *>(smem + L.off_hist);           // [hist_words] private bins
    int32_t *g_rn = reinterpret_cast<int32_t *>(smem + L.off_ints);             // group sums per other-slot
    int32_t *g_rs = g_rn + kcap, *g_cn = g_rs + kcap, *g_cs = g_cn + kcap;
    int32_t *s_cnt = g_cs + kcap, *s_lab = s_cnt + kcap;                        // CRP table sizes / labels
    int32_t *s_cslot = s_lab + kcap, *s_clabel = s_cslot + (kcap + 1);          // candidates
    uint8_t *zs = smem + L.off_zs;                                              // [npad] slot of every entity
    uint8_t *stage0 = smem + L.off_stage;                                       // [NS][2][npad] staged rows
    __shared__ int32_t s_ncand, s_choice_slot, s_choice_label, s_fail, s_hi;

    const int64_t m0 = args.move_off[k_list], m1 = args.move_off[k_list + 1];
    const int64_t nmoves = m1 - m0;
    const uint32_t row_bytes = (uint32_t)npad;

    // ---- load state -------------------------------------------------------------------
    for (int j = tid; j < npad; j += DENSE_THREADS) zs[j] = j < n ? (uint8_t)dd.z[j] : (uint8_t)0;
    for (int k = tid; k < kcap * kcap; k += DENSE_THREADS) table[k] = R.counts[k];
    for (int k = tid; k < kcap; k += DENSE_THREADS) { s_cnt[k] = dd.tab_count[k]; s_lab[k] = dd.tab_label[k]; }
    for (int k = tid; k < L.hist_words; k += DENSE_THREADS) hist[k] = 0u;
    if (tid == 0) {
        s_hi = *dd.hi;
        for (int s = 0; s < NS; s++) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < NS && s < nmoves; s++) {
            const int it = args.move_item[m0 + s];
            const int64_t row = (it >= 0 && it < n) ? it : 0;
            mbar_expect_tx(&bars[s], 2 * row_bytes);
            tma_load_1d(stage0 + (size_t)s * 2 * npad, R.slab[0] + row * R.pitch[0], row_bytes, &bars[s]);
            tma_load_1d(stage0 + (size_t)s * 2 * npad + npad, R.slab[1] + row * R.pitch[1], row_bytes, &bars[s]);
        }
    }

    for (int64_t r = 0; r < nmoves; r++) {
        const int64_t m = m0 + r;
        const int item = args.move_item[m];
        const int stage = (int)(r % NS);
        const uint32_t parity = (uint32_t)((r / NS) & 1);
        const uint8_t *xrow = stage0 + (size_t)stage * 2 * npad, *xcol = xrow + npad;
        const bool bad = item < 0 || item >= n || args.move_dom[m] != 0;
        const int c = bad ? 0 : zs[item];
        const int hi = s_hi;
        if (tid == 0) s_fail = 0;

        mbar_wait(&bars[stage], parity);
        // the diagonal observation (i, i) sits in both staged rows at j = i; it is ONE observation
        const uint32_t dcode = bad ? 0u : (pack_codes(xrow[item], xcol[item]) & 0xFu);
        const int d_n = dcode & 1, d_s = (dcode >> 1) & 1;

        // ---- A: private-bin accumulation over all j (diagonal included, pulled out in B) ------
        const int HT = min(DENSE_THREADS, (L.hist_words / max(hi, 1)) & ~31);
        const int nvec = n >> 4;
        const int chunk_vecs = HT * DENSE_VEC_PER_THREAD;
        for (int v0 = 0; v0 < nvec || v0 == 0; v0 += chunk_vecs) {
            if (tid < HT) {
                const int vend = min(nvec, v0 + chunk_vecs);
                uint32_t *bins = hist + tid;
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
            __syncthreads();
            // ---- B: bins -> group sums per other-entity slot (accumulated over chunks) ------
            for (int k = warp; k < hi; k += DENSE_WARPS) {
                int rn = 0, rs = 0, cn = 0, cs = 0;
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
                        g_rn[k] = rn; g_rs[k] = rs; g_cn[k] = cn; g_cs[k] = cs;
                    } else { g_rn[k] += rn; g_rs[k] += rs; g_cn[k] += cn; g_cs[k] += cs; }
                }
            }
            __syncthreads();
        }
        // stage buffer is free again: prefetch the rows of move r + NS
        if (tid == 0 && r + NS < nmoves) {
            const int it = args.move_item[m + NS];
            const int64_t row = (it >= 0 && it < n) ? it : 0;
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
        // g_r[k] = (n, s) of observations (i, j), j != i, z_j = k;  g_c[k] likewise for (j, i);  (d_n, d_s) = (i, i).
        // Remove them from the table ("rest" counts): thread k owns cells (c, k) and (k, c).
        if (tid < hi) {
            const int k = tid;
            if (k != c) {
                table[c + k * kcap] -= pack_cell(g_rn[k], g_rs[k]);
                table[k + c * kcap] -= pack_cell(g_cn[k], g_cs[k]);
            } else {
                table[c + c * kcap] -= pack_cell(g_rn[c] + g_cn[c] + d_n, g_rs[c] + g_cs[c] + d_s);
            }
        }
        // ---- C: candidates (one thread of another warp, concurrently with the removal) ---------
        if (tid == DENSE_THREADS - 32) {
            int ncand = 0, maxlab = 0, freeslot = -1;  // t_max starts at 0 (cxx/hirm.hh:139)
            for (int k = 0; k < hi; k++) {
                if (s_cnt[k] > 0) maxlab = max(maxlab, s_lab[k]);
                else if (freeslot < 0) freeslot = k;
            }
            if (freeslot < 0 && hi < kcap) freeslot = hi;
            const bool singleton = s_cnt[c] == 1;
            for (int k = 0; k < hi; k++) {
                const int cnt = s_cnt[k] - (k == c ? 1 : 0);
                if (cnt > 0) { s_cslot[ncand] = k; s_clabel[ncand] = s_lab[k]; s_lp[ncand] = log((double)cnt); ncand++; }
            }
            if (singleton) { s_cslot[ncand] = c; s_clabel[ncand] = s_lab[c]; s_lp[ncand] = log(dd.alpha); ncand++; }
            else if (freeslot >= 0) { s_cslot[ncand] = freeslot; s_clabel[ncand] = maxlab + 1; s_lp[ncand] = log(dd.alpha); ncand++; }
            else { raise_error(irm, KERR_NO_FREE_SLOT); s_fail = 1; }
            for (int i = 1; i < ncand; i++) {          // ascending label order
                const int sl = s_cslot[i], lb = s_clabel[i]; const double lw = s_lp[i];
                int j = i - 1;
                while (j >= 0 && s_clabel[j] > lb) { s_cslot[j + 1] = s_cslot[j]; s_clabel[j + 1] = s_clabel[j]; s_lp[j + 1] = s_lp[j]; j--; }
                s_cslot[j + 1] = sl; s_clabel[j + 1] = lb; s_lp[j + 1] = lw;
            }
            s_ncand = ncand;
        }
        __syncthreads();
        const int ncand = s_ncand;

        // ---- D: score, one warp per candidate; lanes over (other slot, row|col) ---------------
        for (int t = warp; t < ncand; t += DENSE_WARPS) {
            const int st = s_cslot[t];
            double acc = 0.0;
            for (int q = lane; q < 2 * hi; q += 32) {
                const int k = q >> 1, col = q & 1;
                if (k == st) continue;                 // merged into cell (t, t) below
                const int gn = col ? g_cn[k] : g_rn[k];
                if (gn == 0) continue;
                const int gs = col ? g_cs[k] : g_rs[k];
                const cell_t cur = col ? table[k + st * kcap] : table[st + k * kcap];
                acc += score_delta(cell_n(cur), cell_s(cur), gn, gs);
            }
            if (lane == 31) {                          // target cell (t, t): row group + column group + diagonal merge
                int gn = d_n, gs = d_s;
                if (st < hi) { gn += g_rn[st] + g_cn[st]; gs += g_rs[st] + g_cs[st]; }
                if (gn > 0) {
                    const cell_t cur = table[st + st * kcap];
                    acc += score_delta(cell_n(cur), cell_s(cur), gn, gs);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) s_lp[t] += acc;
        }
        __syncthreads();

        // ---- E: sample -----------------------------------------------------------------
        if (tid == 0) {
            int idx = 0;
            if (args.flags & 1u) {                     // HIRM_SWEEP_SCORE_ONLY: stay
                for (int t = 0; t < ncand; t++) if (s_cslot[t] == c) idx = t;
            } else {
                const double u = args.uniforms ? args.uniforms[m]
                                               : philox_uniform(args.seed, args.stream[k_list], args.counter0[k_list] + (uint64_t)r);
                idx = choose_index(s_lp, ncand, u);
            }
            if (s_fail || ncand == 0) { s_choice_slot = c; s_choice_label = s_lab[c]; }
            else { s_choice_slot = s_cslot[idx]; s_choice_label = s_clabel[idx]; }
            if (args.out_choice) args.out_choice[m] = s_choice_label;
            if (args.trace_cap > 0) args.trace_n[m] = ncand;
        }
        if (args.trace_cap > 0) {
            for (int t = tid; t < ncand && t < args.trace_cap; t += DENSE_THREADS) {
                args.trace_labels[m * args.trace_cap + t] = s_clabel[t];
                args.trace_logps[m * args.trace_cap + t] = s_lp[t];
            }
        }
        __syncthreads();

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
        }
        __syncthreads();
    }

    // ---- store state ------------------------------------------------------------------
    for (int j = tid; j < n; j += DENSE_THREADS) dd.z[j] = zs[j];
    for (int k = tid; k < kcap * kcap; k += DENSE_THREADS) R.counts[k] = table[k];
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
