// Generic sweep kernel: any mix of relations of arity <= 4 stored as per-entity CSR.
// One thread block per IRM, resident across that IRM's whole move list (moves of one
// IRM are sequentially dependent, cxx/hirm.hh:606-641; IRMs are independent).
//
// Per move (domain d, entity i, current slot c):
//   A  stream i's CSR row; every observation is (1) removed from its count-table cell
//      (the "rest" counts) and (2) added to the group accumulator keyed by the cluster
//      tuple with a SELF digit at i's own positions      -> get_cluster_to_items_list (cxx/hirm.hh:388-397)
//   B  ordered compaction of the touched accumulator cells into the group list
//   C  candidate tables + CRP prior, ascending label      -> CRP::tables_weights_gibbs (:147-161)
//   D  lp[t] = log w_t + sum over target cells of S(N+n, s+sg) - S(N, s)
//                                                         -> logp_gibbs_exact (:441-461)
//   E  sample with the injected uniform / Philox          -> log_choice (cxx/util_math.cc:58-65)
//   F  add the groups back at the chosen slot, re-seat    -> set_cluster_assignment_gibbs (:524-551, :219-224)
// All state is mutated with integer atomics / L2 (.cg) accesses, so count tables are
// bit-exact and, because the group list is ordered, log-probabilities are deterministic.
#pragma once
#include "engine_types.cuh"

namespace hirm {

constexpr int GEN_THREADS = 1024;  // maximum; launched with 1024 (at most one IRM per SM: shortest move), 256, or 128 (many IRMs: 8 CTAs per SM)
constexpr int GEN_WARPS = GEN_THREADS / 32;
constexpr int GEN_MAXK = 256;   // kcap <= 255 candidates + 1 fresh
constexpr int GEN_LF = 256;     // shared-memory copy of lgamma(m+1), m < 256; Stirling beyond (device_math.cuh)

struct GroupRef {
    const RelationDev *R;
    int a, b;            // positions of the moved domain (b = -1 if it occurs once)
    int da, db;          // digits at a, b
    int64_t rest_g;      // accumulator index without positions a, b
    int64_t rest_c;      // count-table index without positions a, b
    int kd;              // SELF digit of the moved domain
    cell_t own;
};

__device__ __forceinline__ GroupRef decode_group(const IrmDev &irm, int d, int rel, int64_t cellidx) {
    GroupRef g;
    const RelationDev *R = &irm.rel[rel];
    g.R = R; g.a = -1; g.b = -1; g.da = 0; g.db = 0; g.rest_g = 0; g.rest_c = 0;
    g.kd = irm.dom[d].kcap;
    int64_t rem = cellidx;
    for (int p = 0; p < R->arity; p++) {
        int radix = irm.dom[R->dom[p]].kcap + 1;
        int digit = (int)(rem % radix);
        rem /= radix;
        if (R->dom[p] == d && g.a < 0) { g.a = p; g.da = digit; }
        else if (R->dom[p] == d) { g.b = p; g.db = digit; }
        else { g.rest_g += digit * R->gstride[p]; g.rest_c += digit * R->cstride[p]; }
    }
    g.own = __ldcg(&R->groups[cellidx]);
    return g;
}

// Resolve the group against candidate slot t: returns false if another group owns the
// merged target cell; otherwise the merged (n, s) to add and the count-table index.
__device__ __forceinline__ bool group_target(const GroupRef &g, int t, cell_t &delta, int64_t &cidx) {
    const RelationDev *R = g.R;
    if (g.b < 0) {                                   // domain occurs once: no merging
        delta = g.own;
        cidx = g.rest_c + t * R->cstride[g.a];
        return true;
    }
    const int64_t ga = R->gstride[g.a], gb = R->gstride[g.b];
    const int64_t ca = R->cstride[g.a], cb = R->cstride[g.b];
    const int SELF = g.kd;
    if (g.da == SELF && g.db == SELF) {              // (i, i): target (t, t); absorbs row/col groups whose other end sits in t
        delta = g.own + __ldcg(&R->groups[g.rest_g + SELF * ga + t * gb]) + __ldcg(&R->groups[g.rest_g + t * ga + SELF * gb]);
        cidx = g.rest_c + t * ca + t * cb;
        return true;
    }
    if (g.da == SELF) {                              // i at position a, other end in slot db
        if (g.db != t) { delta = g.own; cidx = g.rest_c + t * ca + g.db * cb; return true; }
        if (cell_n(__ldcg(&R->groups[g.rest_g + SELF * ga + SELF * gb])) > 0) return false;
        delta = g.own + __ldcg(&R->groups[g.rest_g + t * ga + SELF * gb]);
        cidx = g.rest_c + t * ca + t * cb;
        return true;
    }
    // i at position b, other end in slot da
    if (g.da != t) { delta = g.own; cidx = g.rest_c + g.da * ca + t * cb; return true; }
    if (cell_n(__ldcg(&R->groups[g.rest_g + SELF * ga + SELF * gb])) > 0) return false;
    if (cell_n(__ldcg(&R->groups[g.rest_g + SELF * ga + t * gb])) > 0) return false;
    delta = g.own;
    cidx = g.rest_c + t * ca + t * cb;
    return true;
}

// count-table index of the group's cell when the entity sits in slot t (no merging: used
// to add the groups back, where integer adds merge by themselves)
__device__ __forceinline__ int64_t group_cell_at(const GroupRef &g, int t) {
    const RelationDev *R = g.R;
    int64_t c = g.rest_c + (g.da == g.kd ? t : g.da) * R->cstride[g.a];
    if (g.b >= 0) c += (g.db == g.kd ? t : g.db) * R->cstride[g.b];
    return c;
}

// One decoded group, built once per move (phase B2) and read by every candidate's scoring warp.
struct GroupRec {
    const RelationDev *R;
    int64_t rest_g, rest_c;   // accumulator / count-table index without the moved domain's positions
    int64_t cellidx;          // index into R->groups
    cell_t own;
    int16_t a, b, da, db, kd, pad0, pad1, pad2;
};
static_assert(sizeof(GroupRec) == 56, "GroupRec layout (engine.cu sizes the scratch by it)");

__device__ __forceinline__ GroupRef ref_of(const GroupRec &r) {
    GroupRef g;
    g.R = r.R; g.a = r.a; g.b = r.b; g.da = r.da; g.db = r.db; g.rest_g = r.rest_g; g.rest_c = r.rest_c; g.kd = r.kd; g.own = r.own;
    return g;
}

// S(N+1, s+x) - S(N, s) for a group of ONE observation (every group of a unary relation, most groups of a sparse one):
// log(s + 1) - log(N + 2) if x = 1, log(N - s + 1) - log(N + 2) if x = 0 -- two logarithms instead of the general
// form's six; gn = 0 (a lane without work) gives 0.
__device__ __forceinline__ double score_delta_unit(uint32_t N, uint32_t sv, uint32_t gn, uint32_t gs) {
    const double a = gs ? (double)sv + 1.0 : (double)(N - sv) + 1.0;
    const double v = log_core(a) - log_core((double)N + 2.0);
    return gn ? v : 0.0;
}

// Dynamic shared memory of one CTA: [rcap] group records, [rcap] touched accumulator bits, [pcap] partial sums.
struct GenericLaunch { int32_t rcap, pcap; };
__host__ __device__ inline size_t generic_smem_bytes(int rcap, int pcap) { return (size_t)rcap * (sizeof(GroupRec) + 8) + (size_t)pcap * 8; }

__global__ void __launch_bounds__(GEN_THREADS, 1) generic_sweep_kernel(SweepArgs args, const int32_t *blk_list, GenericLaunch GL) {
    const int k_list = blk_list ? blk_list[blockIdx.x] : blockIdx.x;
    const IrmDev &irm = args.irms[args.irm_ids[k_list]];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NT = blockDim.x, NW = NT >> 5;

    // CRP state of every domain (table sizes, labels, highest used slot + 1), resident for the whole move list: only
    // this CTA changes it (phase F writes shared and global copies alike)
    __shared__ int32_t s_cnt_all[MAX_DOMAINS * GEN_MAXK], s_lab_all[MAX_DOMAINS * GEN_MAXK], s_hi_all[MAX_DOMAINS];
    __shared__ int32_t s_cslot[GEN_MAXK + 1], s_clabel[GEN_MAXK + 1], u_slot[GEN_MAXK + 1], u_label[GEN_MAXK + 1];
    __shared__ double s_lp[GEN_MAXK + 1], s_tmp[GEN_MAXK + 1];
    __shared__ double s_lf[GEN_LF];
    __shared__ int32_t s_warp_tot[GEN_WARPS];
    __shared__ int32_t s_ncand, s_choice_slot, s_choice_label, s_fail;
    GroupRec *recs = reinterpret_cast<GroupRec *>(irm.gl_cell);     // [gl_cap] group records beyond the shared-memory ones (engine scratch)
    extern __shared__ __align__(16) unsigned char gen_dyn[];
    GroupRec *s_recs = reinterpret_cast<GroupRec *>(gen_dyn);                                  // [rcap]
    long long *s_bits = reinterpret_cast<long long *>(gen_dyn + (size_t)GL.rcap * sizeof(GroupRec));   // [rcap]
    double *s_part = reinterpret_cast<double *>(s_bits + GL.rcap);                              // [pcap]
    const int RCAP = GL.rcap, PCAP = GL.pcap;

    for (int k = tid; k < GEN_LF; k += NT) s_lf[k] = g_logfact[k];
    for (int dk = 0; dk < irm.n_domains; dk++) {
        const DomainDev &dq = irm.dom[dk];
        const int h = __ldcg(dq.hi);
        if (tid == 0) s_hi_all[dk] = h;
        for (int k = tid; k < h; k += NT) { s_cnt_all[dk * GEN_MAXK + k] = __ldcg(&dq.tab_count[k]); s_lab_all[dk * GEN_MAXK + k] = __ldcg(&dq.tab_label[k]); }
    }
    __syncthreads();

    // A move's inputs that do not depend on the move before it (ids, CSR row bounds) and the one that almost never does
    // (the entity's own slot) are fetched one move ahead: their L2 round trips overlap the current move.
    struct Pref { int d, item, c; int64_t beg, end; };
    const int64_t m0 = args.move_off[k_list], m1 = args.move_off[k_list + 1];
    auto fetch = [&](int64_t mm) {
        Pref f; f.d = -1; f.item = -1; f.c = -1; f.beg = 0; f.end = 0;
        if (mm >= m1) return f;
        f.d = args.move_dom[mm]; f.item = args.move_item[mm];
        if (f.d < 0 || f.d >= irm.n_domains || f.item < 0 || f.item >= irm.dom[f.d].n) { f.d = -1; return f; }
        const DomainDev &dq = irm.dom[f.d];
        f.c = __ldcg(&dq.z[f.item]);
        if (dq.ptr) { f.beg = dq.ptr[f.item]; f.end = dq.ptr[f.item + 1]; }
        return f;
    };

    Pref nxt = fetch(m0);
    // debug hook (hirm_engine_debug_phase_cycles): thread 0's SM cycles per phase, summed over the moves:
    // [0] A  [1] B1 (+C)  [2] B2  [3] D  [4] E  [5] F  -- 32 slots per block, shared with the dense kernel's layout
    const bool prof = args.phase_cycles != nullptr && tid == 0;
    long long ph_t = prof ? clock64() : 0, ph_acc[6] = {0, 0, 0, 0, 0, 0};
#define GEN_PHASE(i) do { if (prof) { const long long now = clock64(); ph_acc[i] += now - ph_t; ph_t = now; } } while (0)
    for (int64_t m = m0; m < m1; m++) {
        const Pref cur = nxt;
        nxt = fetch(m + 1);
        const bool bad = cur.d < 0;
        const int d = bad ? 0 : cur.d, item = cur.item;
        const DomainDev &dd = irm.dom[d];
        const int c = bad ? -1 : cur.c;
        if (c < 0) {                                  // uniform across the block
            if (tid == 0) {
                raise_error(irm, KERR_BAD_MOVE);
                if (args.out_choice) args.out_choice[m] = -1;
                if (args.trace_cap > 0) args.trace_n[m] = 0;
            }
            continue;
        }
        if (tid == 0) s_fail = 0;
        int32_t *s_cnt = s_cnt_all + d * GEN_MAXK, *s_lab = s_lab_all + d * GEN_MAXK;
        const int hi = s_hi_all[d];

        // ---- A: remove + group (fire-and-forget reductions; completed by the fence below) -----------
        if (dd.ptr) {
            const int64_t beg = cur.beg, end = cur.end;
            for (int64_t q = beg + tid; q < end; q += NT) {
                const uint32_t meta = dd.meta[q];
                const int rel = meta & META_REL_MASK;
                const uint32_t val = (meta >> META_VAL_SHIFT) & 1u, selfmask = (meta >> META_SELF_SHIFT) & 0xFu;
                const RelationDev *R = &irm.rel[rel];
                int64_t gidx = 0, cidx = 0;
                int oi = 0; bool first = true, absent = false;
                for (int p = 0; p < R->arity; p++) {
                    int slot, gd;
                    if ((selfmask >> p) & 1u) {
                        if (first) first = false; else oi++;
                        slot = c; gd = dd.kcap;
                    } else {
                        const int id = dd.other[(int64_t)oi * dd.nnz + q]; oi++;
                        slot = __ldcg(&irm.dom[R->dom[p]].z[id]);
                        if (slot < 0) { absent = true; slot = 0; }
                        gd = slot;
                    }
                    gidx += gd * R->gstride[p];
                    cidx += slot * R->cstride[p];
                }
                if (absent) { raise_error(irm, KERR_ABSENT_ENTITY); continue; }
                const cell_t one = pack_cell(1u, val);
                atomicAdd(&R->groups[gidx], one);
                const int64_t bit = R->gbit0 + gidx;
                atomicOr(&irm.gbitmap[bit >> 6], 1ull << (bit & 63));
                atomicAdd(&R->counts[cidx], (cell_t)0 - one);
            }
        }
        __threadfence();
        __syncthreads();
        GEN_PHASE(0);

        // ---- C: candidates, CRP::tables_weights_gibbs (cxx/hirm.hh:147-161): warp 0, lanes over the tables ----
        if (warp == 0) {
            int maxlab = 0, freeslot = 1 << 30;        // t_max starts at 0 (cxx/hirm.hh:139)
            for (int k = lane; k < hi; k += 32) {
                if (s_cnt[k] > 0) maxlab = max(maxlab, s_lab[k]); else freeslot = min(freeslot, k);
            }
            maxlab = __reduce_max_sync(0xffffffffu, maxlab);
            freeslot = __reduce_min_sync(0xffffffffu, freeslot);
            if (freeslot == (1 << 30)) freeslot = hi < dd.kcap ? hi : -1;
            const bool singleton = s_cnt[c] == 1;
            const double logalpha = log(dd.alpha);
            int ncount = 0;
            for (int base = 0; base < hi; base += 32) {
                const int k = base + lane;
                const int cnt = k < hi ? s_cnt[k] - (k == c ? 1 : 0) : 0;
                const unsigned bal = __ballot_sync(0xffffffffu, cnt > 0);
                const int pos = ncount + __popc(bal & ((1u << lane) - 1u));
                const double lw = log((double)max(cnt, 1));
                if (cnt > 0) { u_slot[pos] = k; u_label[pos] = s_lab[k]; s_tmp[pos] = lw; }
                ncount += __popc(bal);
            }
            if (lane == 0) {
                if (singleton) { u_slot[ncount] = c; u_label[ncount] = s_lab[c]; s_tmp[ncount] = logalpha; }   // own table at weight alpha, no fresh one (:152-159)
                else if (freeslot >= 0) { u_slot[ncount] = freeslot; u_label[ncount] = maxlab + 1; s_tmp[ncount] = logalpha; }   // fresh table, label max+1 (:144)
                else { raise_error(irm, KERR_NO_FREE_SLOT); s_fail = 1; }
            }
            const int nc = ncount + ((singleton || freeslot >= 0) ? 1 : 0);
            __syncwarp();
            for (int j = lane; j < nc; j += 32) {      // ascending label order: rank by counting (labels are distinct)
                const int lb = u_label[j];
                int rank = 0;
                for (int i = 0; i < nc; i++) rank += u_label[i] < lb;
                s_cslot[rank] = u_slot[j]; s_clabel[rank] = lb; s_lp[rank] = s_tmp[j];
            }
            if (lane == 0) s_ncand = nc;
        }

        // ---- B1: ordered list of the touched accumulator cells (bitmap scan: ascending bit = deterministic order) ----
        int64_t base = 0;
        for (int64_t w0 = 0; w0 < irm.gwords; w0 += NT) {
            const int64_t w = w0 + tid;
            unsigned long long bits = w < irm.gwords ? __ldcg(&irm.gbitmap[w]) : 0ull;
            const int cnt = __popcll(bits);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) s_warp_tot[warp] = incl;
            __syncthreads();
            int wv = lane < NW ? s_warp_tot[lane] : 0, wincl = wv;      // scan of the warp totals (at most 32 warps)
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int v = __shfl_up_sync(0xffffffffu, wincl, o);
                if (lane >= o) wincl += v;
            }
            const int total = __shfl_sync(0xffffffffu, wincl, 31);
            const int woff = __shfl_sync(0xffffffffu, wincl - wv, warp);
            int64_t off = base + woff + incl - cnt;
            while (bits) {
                const int bpos = __ffsll((long long)bits) - 1;
                bits &= bits - 1;
                if (off >= irm.gl_cap) raise_error(irm, KERR_GROUP_OVERFLOW);
                else if (off < RCAP) s_bits[off] = w * 64 + bpos;
                else recs[off].cellidx = w * 64 + bpos;
                off++;
            }
            base += total;
            __syncthreads();
        }
        const int64_t ngroups = base < irm.gl_cap ? base : irm.gl_cap;
        GEN_PHASE(1);
        // ---- B2: decode every group once, one group per thread (for every candidate) ---------------------------------
        for (int64_t g = tid; g < ngroups; g += NT) {
            const int64_t bit = g < RCAP ? s_bits[g] : recs[g].cellidx;
            const int rel = irm.word_rel[bit >> 6];
            const int64_t cellidx = bit - irm.rel[rel].gbit0;
            const GroupRef gr = decode_group(irm, d, rel, cellidx);
            GroupRec rc;
            rc.R = gr.R; rc.rest_g = gr.rest_g; rc.rest_c = gr.rest_c; rc.cellidx = cellidx; rc.own = gr.own;
            rc.a = (int16_t)gr.a; rc.b = (int16_t)gr.b; rc.da = (int16_t)gr.da; rc.db = (int16_t)gr.db; rc.kd = (int16_t)gr.kd;
            rc.pad0 = rc.pad1 = rc.pad2 = 0;
            if (g < RCAP) s_recs[g] = rc; else recs[g] = rc;
        }
        __syncthreads();
        GEN_PHASE(2);
        const int ncand = s_ncand;

        // ---- D: score.  Work unit = (candidate, 32 consecutive groups), one warp each, lanes over the groups; the partial
        //      sums are added per candidate in chunk order afterwards (fixed order: deterministic) -------------------------
        const int nch = (int)((ngroups + 31) >> 5);
        const bool flat = (int64_t)ncand * nch <= PCAP;
        if (flat) {
            const int nunits = ncand * nch;
            for (int unit = warp; unit < nunits; unit += NW) {
                const int t = unit / nch, ch = unit - t * nch;
                const int slot_t = s_cslot[t];
                const int64_t g = (int64_t)ch * 32 + lane;
                uint32_t N = 0u, sv = 0u, gn = 0u, gs = 0u;
                if (g < ngroups) {
                    const GroupRef gr = ref_of(g < RCAP ? s_recs[g] : recs[g]);
                    cell_t delta; int64_t cidx;
                    if (group_target(gr, slot_t, delta, cidx)) {
                        const cell_t cur = __ldcg(&gr.R->counts[cidx]);
                        N = cell_n(cur); sv = cell_s(cur); gn = cell_n(delta); gs = cell_s(delta);
                    }
                }
                __syncwarp();
                double acc = __all_sync(0xffffffffu, gn <= 1u) ? score_delta_unit(N, sv, gn, gs) : score_delta_bf<GEN_LF>(s_lf, N, sv, gn, gs);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) s_part[unit] = acc;
            }
            __syncthreads();
            for (int t = tid; t < ncand; t += NT) {
                double acc = 0.0;
                for (int ch = 0; ch < nch; ch++) acc += s_part[t * nch + ch];
                s_lp[t] += acc;
            }
        } else {
            for (int t = warp; t < ncand; t += NW) {   // very many groups: one warp per candidate walks all of them
                const int slot_t = s_cslot[t];
                double acc = 0.0;
                for (int64_t g0 = 0; g0 < ngroups; g0 += 32) {
                    const int64_t g = g0 + lane;
                    uint32_t N = 0u, sv = 0u, gn = 0u, gs = 0u;
                    if (g < ngroups) {
                        const GroupRef gr = ref_of(g < RCAP ? s_recs[g] : recs[g]);
                        cell_t delta; int64_t cidx;
                        if (group_target(gr, slot_t, delta, cidx)) {
                            const cell_t cur = __ldcg(&gr.R->counts[cidx]);
                            N = cell_n(cur); sv = cell_s(cur); gn = cell_n(delta); gs = cell_s(delta);
                        }
                    }
                    __syncwarp();
                    double v = __all_sync(0xffffffffu, gn <= 1u) ? score_delta_unit(N, sv, gn, gs) : score_delta_bf<GEN_LF>(s_lf, N, sv, gn, gs);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    acc += v;
                }
                if (lane == 0) s_lp[t] += acc;
            }
        }
        __syncthreads();
        GEN_PHASE(3);

        // ---- E: sample (warp 0): log_choice (cxx/util_math.cc:58-65); exp lane-parallel, sums in index order ----
        if (warp == 0) {
            int idx = 0;
            if (args.flags & 1u) {                     // HIRM_SWEEP_SCORE_ONLY: stay
                for (int t = lane; t < ncand; t += 32) if (s_cslot[t] == c) idx = t;
                idx = __reduce_max_sync(0xffffffffu, idx);
            } else if (ncand > 1) {
                double mx = -1.0 / 0.0;
                for (int t = lane; t < ncand; t += 32) mx = fmax(mx, s_lp[t]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                for (int t = lane; t < ncand; t += 32) s_tmp[t] = exp(s_lp[t] - mx);
                __syncwarp();
                if (lane == 0) {
                    const double u = args.uniforms ? args.uniforms[m]
                                                   : philox_uniform(args.seed, args.stream[k_list], args.counter0[k_list] + (uint64_t)(m - m0));
                    double total = 0.0;
                    for (int t = 0; t < ncand; t++) total += s_tmp[t];
                    const double x = u * total;
                    double cum = 0.0;
                    idx = ncand - 1;
                    for (int t = 0; t < ncand; t++) { cum += s_tmp[t]; if (cum > x) { idx = t; break; } }
                }
                idx = __shfl_sync(0xffffffffu, idx, 0);
            }
            if (lane == 0) {
                if (s_fail || ncand == 0) { s_choice_slot = c; s_choice_label = s_lab[c]; }
                else { s_choice_slot = s_cslot[idx]; s_choice_label = s_clabel[idx]; }
                if (args.out_choice) args.out_choice[m] = s_choice_label;
                if (args.trace_cap > 0) args.trace_n[m] = ncand;
            }
        }
        if (args.trace_cap > 0) {
            for (int t = tid; t < ncand && t < args.trace_cap; t += NT) {
                args.trace_labels[m * args.trace_cap + t] = s_clabel[t];
                args.trace_logps[m * args.trace_cap + t] = s_lp[t];
            }
        }
        __syncthreads();

        GEN_PHASE(4);
        // ---- F: apply ------------------------------------------------------------------
        const int tnew = s_choice_slot;
        for (int64_t g = tid; g < ngroups; g += NT) {
            const GroupRec rc = g < RCAP ? s_recs[g] : recs[g];
            const GroupRef gr = ref_of(rc);
            atomicAdd(&gr.R->counts[group_cell_at(gr, tnew)], gr.own);
            __stcg(&gr.R->groups[rc.cellidx], (cell_t)0);
            __stcg(&irm.gbitmap[(gr.R->gbit0 + rc.cellidx) >> 6], 0ull);
        }
        if (tid == 0 && tnew != c) {
            __stcg(&dd.z[item], tnew);
            __stcg(&dd.tab_count[c], s_cnt[c] - 1);
            const int newcnt = (tnew < hi ? s_cnt[tnew] : 0) + 1;
            __stcg(&dd.tab_count[tnew], newcnt);
            if (newcnt == 1) { __stcg(&dd.tab_label[tnew], s_choice_label); s_lab[tnew] = s_choice_label; }
            int nhi = max(hi, tnew + 1);
            if (s_cnt[c] == 1) {                       // table c emptied: CRP::unincorporate erases it (:95-97)
                while (nhi > 0 && (nhi - 1 == c || (nhi - 1 != tnew && (nhi - 1 >= hi || s_cnt[nhi - 1] == 0)))) nhi--;
            }
            __stcg(dd.hi, nhi);
            s_cnt[c] -= 1; s_cnt[tnew] = newcnt; s_hi_all[d] = nhi;      // the resident copies
        }
        if (tnew != c && nxt.d == d && nxt.item == item) nxt.c = tnew;    // the same entity again: its prefetched slot is stale
        __threadfence();
        __syncthreads();
        GEN_PHASE(5);
    }
    if (prof) for (int i = 0; i < 6; i++) args.phase_cycles[blockIdx.x * 32 + i] = ph_acc[i];
#undef GEN_PHASE
}

// ---------------------------------------------------------------------------------------
// Count-table rebuild: the table is the histogram of (observations, z)   (cxx/hirm.hh:281-289)
// ---------------------------------------------------------------------------------------
__global__ void rebuild_counts_coo_kernel(const IrmDev *irms, int irm_id, int rel) {
    const IrmDev &irm = irms[irm_id];
    const RelationDev *R = &irm.rel[rel];
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < R->nobs; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t cidx = 0; bool absent = false;
        for (int p = 0; p < R->arity; p++) {
            const int slot = irm.dom[R->dom[p]].z[R->coo_ids[e * R->arity + p]];
            if (slot < 0) absent = true;
            cidx += (slot < 0 ? 0 : slot) * R->cstride[p];
        }
        if (absent) { raise_error(irm, KERR_ABSENT_ENTITY); continue; }
        atomicAdd(&R->counts[cidx], pack_cell(1u, R->coo_val[e]));
    }
}

// Sum of BetaBernoulli::logp_score over the cells of one relation   (Relation::logp_score, cxx/hirm.hh:516-522)
// Single block, fixed reduction tree: deterministic.
__global__ void __launch_bounds__(1024) relation_score_kernel(const IrmDev *irms, int irm_id, int rel, double *out) {
    const RelationDev *R = &irms[irm_id].rel[rel];
    __shared__ double s_part[32];
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < R->ncells; i += blockDim.x) {
        const cell_t c = R->counts[i];
        if (cell_n(c)) acc += cell_score(cell_n(c), cell_s(c));
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? s_part[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) *out = v;
    }
}

}  // namespace hirm
