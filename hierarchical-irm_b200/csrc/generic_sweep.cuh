// Generic sweep kernel: any mix of relations of arity <= 4 stored as per-entity CSR.
// One thread block per IRM, resident across that IRM's whole move list (moves of one
// IRM are sequentially dependent, cxx/hirm.hh:606-641; IRMs are independent).
//
// Per move (domain d, entity i, current slot c):
//   A  stream i's CSR row; every observation is added to the group accumulator cell keyed by (the positions holding i
//      itself, the tables of the other positions)          -> get_cluster_to_items_list (cxx/hirm.hh:388-397).
//      The accumulator of a move is a small compact table (engine_types.cuh: RelDom) in SHARED memory (shared-memory
//      atomics; global scratch only if it does not fit).  Relations in which d occurs once are scored against the count
//      table with i's observations removed VIRTUALLY (the current cell minus the group's own counts): no global atomic
//      unless the entity really changes tables.  Relations in which d occurs twice remove i's observations from their
//      count table for the duration of the move (the (i,i) -> (t,t) rule merges groups, cxx/hirm.hh:324-339).
//   B  ordered list of the touched accumulator cells (bitmap scan), one group record per cell
//   C  candidate tables + CRP prior, ascending label      -> CRP::tables_weights_gibbs (:147-161)
//   D  lp[t] = log w_t + sum over target cells of S(N+n, s+sg) - S(N, s)
//                                                         -> logp_gibbs_exact (:441-461)
//   E  sample with the injected uniform / Philox          -> log_choice (cxx/util_math.cc:58-65)
//   F  move the groups to the chosen slot, re-seat        -> set_cluster_assignment_gibbs (:524-551, :219-224)
// Count tables change by integer atomics only (bit-exact); the group list is ordered, so log-probabilities are
// deterministic.
#pragma once
#include <cooperative_groups.h>
#include "engine_types.cuh"

namespace hirm {

constexpr int GEN_THREADS = 1024;  // maximum; launched with 1024 (at most one IRM per SM: shortest move), 256, or 128 (many IRMs: 8 CTAs per SM)
constexpr int GEN_WARPS = GEN_THREADS / 32;
constexpr int GEN_MAXK = 256;   // kcap <= 255 candidates + 1 fresh
constexpr int GEN_LF = 256;     // shared-memory copy of lgamma(m+1), m < 256; Stirling beyond (device_math.cuh)

// One group of the moving entity's observations, decoded once per move (phase B) and read by every candidate: everything
// the scoring of a (group, candidate) pair needs sits in the record, so that phase D's dependent chain is one shared-memory
// read and one count-table load (no relation / domain descriptor loads).
struct GroupRec {
    const RelationDev *R;     // (phase F: table updates)
    cell_t *counts;           // R->counts
    int64_t rest_c;           // count-table index without the moved domain's positions
    cell_t own;               // (observations, ones) of the group
    int32_t ca, cb;           // count-table strides of the moved domain's positions (cb = 0: the domain occurs once)
    int32_t row0, col0;       // twice-occurring domain: accumulator cell of the row / column group with the other end in slot 0
    int32_t diag, rs;         // ... of the (i, i) group; accumulator stride per other-end slot
    int32_t cell;             // accumulator cell of the group
    uint32_t hcap;            // R->hcap (0 = dense table)
    int16_t pat, dother, twice, pad;   // self pattern 0/1/2; table of the other end (patterns 0, 1 of a twice-occurring domain)
};
static_assert(sizeof(GroupRec) == 72, "GroupRec layout (engine.cu sizes the scratch by it)");
constexpr int GROUP_REC_BYTES = 72;

// The accumulator of one move: 32-bit cells (observations | ones << 16) in shared memory, fed by native 32-bit shared
// atomics (a 64-bit shared atomicAdd is a compare-and-swap spin loop: ATOMS.CAST.SPIN in the SASS), or -- when the
// domain's cell space does not fit or a row is longer than 65534 entries -- 64-bit cells in the IRM's global scratch.
struct AccView {
    const uint32_t *s32;      // shared accumulator, or null
    const cell_t *g;          // global accumulator
    __device__ __forceinline__ cell_t operator[](int32_t c) const {
        if (s32) { const uint32_t v = s32[c]; return pack_cell(v & 0xFFFFu, v >> 16); }
        return __ldcg(&g[c]);
    }
};

// Resolve the group against candidate slot t: false if another group owns the merged target cell; otherwise the
// (n, s) to add (`delta`), the count-table index of the target cell (`cidx`) and the part of that cell's CURRENT content
// that is the moving entity's own observations (`sub`): the entity is removed from the table virtually -- the count
// tables are not touched unless the entity really changes tables.  `acc` is the move's accumulator, c the entity's slot.
// Twice-occurring domain (a, b): the entity's observations sit in the cells (c, y) [row group y: i at a, other end in y],
// (x, c) [column group x] and (c, c) [(i, i)], so cell (x, y) holds  [x = c] row(y) + [y = c] col(x) + [x = y = c] diag.
__device__ __forceinline__ bool group_target(const GroupRec &g, const AccView &acc, int t, int c, cell_t &delta, int64_t &cidx, cell_t &sub) {
    const int64_t ca = g.ca;
    if (!g.twice) {                                  // domain occurs once: no merging
        delta = g.own;
        cidx = g.rest_c + t * ca;
        sub = t == c ? g.own : (cell_t)0;
        return true;
    }
    const int64_t cb = g.cb;
    const int32_t row0 = g.row0, col0 = g.col0, rs = g.rs, diag = g.diag;
    auto own_at = [&](int x, int y) -> cell_t {
        cell_t v = 0;
        if (x == c) v += acc[row0 + y * rs];
        if (y == c) v += acc[col0 + x * rs];
        if (x == c && y == c) v += acc[diag];
        return v;
    };
    if (g.pat == 2) {                                // (i, i): target (t, t); absorbs the row / column groups whose other end sits in t
        delta = g.own + acc[row0 + t * rs] + acc[col0 + t * rs];
        cidx = g.rest_c + t * ca + t * cb;
        sub = own_at(t, t);
        return true;
    }
    if (g.pat == 0) {                                // i at position a, other end in slot dother
        if (g.dother != t) { delta = g.own; cidx = g.rest_c + t * ca + g.dother * cb; sub = own_at(t, g.dother); return true; }
        if (cell_n(acc[diag]) > 0) return false;
        delta = g.own + acc[col0 + t * rs];
        cidx = g.rest_c + t * ca + t * cb;
        sub = own_at(t, t);
        return true;
    }
    // i at position b, other end in slot dother
    if (g.dother != t) { delta = g.own; cidx = g.rest_c + g.dother * ca + t * cb; sub = own_at(g.dother, t); return true; }
    if (cell_n(acc[diag]) > 0) return false;
    if (cell_n(acc[row0 + t * rs]) > 0) return false;
    delta = g.own;
    cidx = g.rest_c + t * ca + t * cb;
    sub = own_at(t, t);
    return true;
}

// count-table index of the group's cell when the entity sits in slot t (no merging: integer adds merge by themselves)
__device__ __forceinline__ int64_t group_cell_at(const GroupRec &g, int t) {
    int64_t c = g.rest_c + (int64_t)(g.pat == 1 ? (int)g.dother : t) * g.ca;
    if (g.twice) c += (int64_t)(g.pat == 0 ? (int)g.dother : t) * g.cb;
    return c;
}

// The record of accumulator cell `cell` of domain d, all but the group's own counts: a pure function of the cell (relation,
// self pattern, tables of the other positions) and of the IRM's descriptors -- precomputed per cell where the cell space is
// small enough (build_cell_templates_kernel), decoded on the fly otherwise.
__device__ __forceinline__ GroupRec decode_cell(const IrmDev &irm, int d, int32_t cell) {
    const int rel = irm.cmap[d][cell];
    const RelationDev *R = &irm.rel[rel];
    const RelDom *D = irm.rdesc + (int64_t)rel * irm.n_domains + d;
    GroupRec rc;
    rc.R = R; rc.counts = R->counts; rc.hcap = R->hcap; rc.cell = cell;
    rc.twice = (int16_t)(D->b >= 0); rc.ca = (int32_t)R->cstride[D->a]; rc.cb = D->b >= 0 ? (int32_t)R->cstride[D->b] : 0;
    rc.pat = (int16_t)(D->b >= 0 && cell >= D->base[2] ? 2 : D->b >= 0 && cell >= D->base[1] ? 1 : 0);
    const int32_t local = cell - D->base[rc.pat];
    const int32_t rest_k = local % D->rest_size;
    rc.dother = (int16_t)(local / D->rest_size);
    int64_t rest_c = 0;
    for (int p = 0; p < R->arity; p++) {
        if (p == D->a || p == D->b) continue;
        const int digit = (rest_k / D->rstride[p]) % irm.dom[R->dom[p]].kcap;
        rest_c += digit * R->cstride[p];
    }
    rc.rest_c = rest_c;
    rc.own = 0;
    rc.row0 = D->base[0] + rest_k; rc.col0 = D->base[1] + rest_k; rc.diag = D->base[2] + rest_k; rc.rs = D->rest_size;
    rc.pad = 0;
    return rc;
}
__global__ void build_cell_templates_kernel(const IrmDev *irms, int irm_id, int d, GroupRec *out) {
    const IrmDev &irm = irms[irm_id];
    for (int32_t cell = blockIdx.x * blockDim.x + threadIdx.x; cell < irm.ctot[d]; cell += gridDim.x * blockDim.x) out[cell] = decode_cell(irm, d, cell);
}

// S(N+1, s+x) - S(N, s) for a group of ONE observation (every group of a unary relation, most groups of a sparse one):
// log(s + 1) - log(N + 2) if x = 1, log(N - s + 1) - log(N + 2) if x = 0 -- two logarithms of integers instead of the
// general form's six; gn = 0 (a lane without work) gives 0.  `lt` is the table of log(m), m < LOGTAB_SIZE.
__device__ __forceinline__ double score_delta_unit(const double *lt, uint32_t N, uint32_t sv, uint32_t gn, uint32_t gs) {
    const uint32_t a = gs ? sv + 1u : N - sv + 1u;
    double v;
    if (N + 2u < LOGTAB_SIZE) v = __ldg(lt + a) - __ldg(lt + N + 2u);
    else v = log_core((double)a) - log_core((double)N + 2.0);
    return gn ? v : 0.0;
}

// Groups of a few observations (a sparse relation's typical group): the factorial ratios written out,
//   S(N+n, s+sg) - S(N, s) = sum_{j<sg} log(s+1+j) + sum_{j<n-sg} log(N-s+1+j) - sum_{j<n} log(N+2+j),
// from the table of integer logarithms (consecutive entries: one or two cache lines per sum) -- or, for counts beyond
// the table, as products: two logarithms and 2n multiplications.  `nmax` >= every lane's gn is warp-uniform and at most
// 16, so each product stays below 2^512.
__device__ __forceinline__ double score_delta_small(const double *lt, uint32_t N, uint32_t sv, uint32_t gn, uint32_t gs, int nmax) {
    const uint32_t gf = gn - gs;
    if (__all_sync(0xffffffffu, N + 18u < LOGTAB_SIZE)) {
        double num = 0.0, den = 0.0;
        for (int j = 0; j < nmax; j++) {
            if ((uint32_t)j < gs) num += __ldg(lt + sv + 1u + j);
            if ((uint32_t)j < gf) num += __ldg(lt + (N - sv) + 1u + j);
            if ((uint32_t)j < gn) den += __ldg(lt + N + 2u + j);
        }
        return num - den;
    }
    double pnum = 1.0, pden = 1.0;
    for (int j = 0; j < nmax; j++) {
        const double a = (uint32_t)j < gs ? (double)sv + (double)(j + 1) : 1.0;
        const double b = (uint32_t)j < gf ? (double)(N - sv) + (double)(j + 1) : 1.0;
        const double c = (uint32_t)j < gn ? (double)N + (double)(j + 2) : 1.0;
        pnum *= a * b;
        pden *= c;
    }
    return log_core(pnum) - log_core(pden);
}
// the cell score of a warp's 32 groups, by the cheapest form that fits all of them (warp-uniform choice)
__device__ __forceinline__ double score_delta_any(const double *lf, const double *lt, uint32_t N, uint32_t sv, uint32_t gn, uint32_t gs) {
    const int nmax = (int)__reduce_max_sync(0xffffffffu, gn);
    if (nmax <= 1) return score_delta_unit(lt, N, sv, gn, gs);
    if (nmax <= 16) return score_delta_small(lt, N, sv, gn, gs, nmax);
    return score_delta_bf<GEN_LF>(lf, N, sv, gn, gs);
}

// Dynamic shared memory of one CTA: [rcap] group records, [rcap] touched accumulator cells, [pcap] partial sums,
// [scap] 32-bit accumulator cells.
struct GenericLaunch { int32_t rcap, pcap, scap, kstride, ndom, roww, csize, ubw, adcap, headk; };   // headk: chunks (of one entry per thread) of the next move's row copied ahead   // adcap: relation descriptors of the moved domain kept in shared memory   // ubw: words per unary-block bit row (0: none)   // ndom: most domains of an IRM; roww: words per prefetched row entry (1 + other ids); csize: CTAs per IRM (thread-block cluster)
__host__ __device__ inline size_t generic_smem_bytes(int rcap, int pcap, int scap, int csize = 1) {
    return (size_t)rcap * (sizeof(GroupRec) + 4) + (size_t)pcap * 8 + (size_t)scap * 4 * (csize > 1 ? 2 : 1) + 16;   // cluster: local + merged accumulator
}
// + [2][roww][threads] words: the first `threads` entries of the next move's CSR row (meta + other ids), copied
// asynchronously one move ahead
// + [ndom][kstride] (double + 2 int32): resident CRP state of every domain
// + [threads] bitmap words and [threads] list offsets of the group-list build
// + [2][2][ubw] words: the next / current move's unary-block bit rows (observed, value), copied one move ahead
// + [adcap] ADesc
__host__ __device__ inline size_t generic_smem_bytes(int rcap, int pcap, int scap, int threads, int kstride, int ndom, int roww, int csize = 1, int ubw = 0,
                                                     int adcap = 0, int headk = 1) {
    return generic_smem_bytes(rcap, pcap, scap, csize) + (size_t)2 * headk * roww * threads * 4 + (size_t)2 * threads * 4 + (size_t)4 * ubw * 4 + 8 +
           (size_t)ndom * kstride * 16 + (size_t)adcap * sizeof(ADesc);
}

// Thread-block cluster per IRM (north_star subsystem 3).  The moves of one IRM are sequential, so an IRM that is alone
// on the GPU (an HIRM with a handful of relation clusters, a single chain) can only be made faster WITHIN a move: launched
// as a cluster of C CTAs (C SMs), the CTAs split the entity's row (phase A) and the (candidate x group) scoring units
// (phase D) and exchange through distributed shared memory:
//   A   CTA q takes every C-th chunk of the CSR row / unary columns into its LOCAL accumulator
//   --  cluster barrier; every CTA adds its non-empty cells into the MERGED accumulator of all C CTAs (remote shared
//       atomics); cluster barrier: the merged accumulators are identical
//   B   group list and records, redundantly per CTA from its merged accumulator (no exchange)
//   D   unit u is scored by CTA u mod C, the partial sum stored into every CTA's s_part[u]; cluster barrier; every
//       CTA adds the partials in unit order -- the same additions in the same order for every C, so results do not depend
//       on the cluster size
//   E   sampled redundantly (same log-probabilities, same uniform)
//   F   count-table atomics of group g by CTA g mod C; z / CRP state written by every CTA (identical values), so each
//       CTA reads back its own writes; the next cluster barrier (after the next A) orders the table updates before the
//       next scoring
// With C = 1 (no cluster attribute) the barriers are __syncthreads() and there is one accumulator.

__global__ void __launch_bounds__(GEN_THREADS, 1) generic_sweep_kernel(SweepArgs args, const int32_t *blk_list, GenericLaunch GL) {
    namespace cg = cooperative_groups;
    const int C = GL.csize > 1 ? GL.csize : 1;                 // CTAs of this IRM's cluster
    const int crank = C > 1 ? (int)cg::this_cluster().block_rank() : 0;
    const int k_blk = (int)blockIdx.x / C;
    const int k_list = blk_list ? blk_list[k_blk] : k_blk;
    const IrmDev &irm = args.irms[args.irm_ids[k_list]];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NT = blockDim.x, NW = NT >> 5;
    auto csync = [&]() { if (C > 1) cg::this_cluster().sync(); else __syncthreads(); };

    // CRP state of every domain (table sizes, labels, highest used slot + 1), resident for the whole move list: only
    // this CTA changes it (phase F writes shared and global copies alike)
    __shared__ int32_t s_hi_all[MAX_DOMAINS];     // (table sizes / labels / log sizes: dynamic shared memory, [n_domains][kstride])
    __shared__ int32_t s_cslot[GEN_MAXK + 1], s_clabel[GEN_MAXK + 1], u_slot[GEN_MAXK + 1], u_label[GEN_MAXK + 1];
    __shared__ double s_lp[GEN_MAXK + 1], s_tmp[GEN_MAXK + 1];
    __shared__ double s_lf[GEN_LF];
    __shared__ int32_t s_warp_tot[GEN_WARPS];
    __shared__ int32_t s_ncand, s_choice_slot, s_choice_label, s_fail;
    __shared__ double s_logalpha[MAX_DOMAINS], s_u;   // log(alpha); the move's uniform
    GroupRec *recs = reinterpret_cast<GroupRec *>(irm.gl_cell);     // [gl_cap] group records beyond the shared-memory ones (engine scratch)
    extern __shared__ __align__(16) unsigned char gen_dyn[];
    const int RCAP = GL.rcap, PCAP = GL.pcap, SCAP = GL.scap;
    const int64_t glcap = irm.gl_cap;
    GroupRec *s_recs = reinterpret_cast<GroupRec *>(gen_dyn);                                  // [rcap]
    double *s_part = reinterpret_cast<double *>(s_recs + RCAP);                                 // [pcap]
    uint32_t *s_acc = reinterpret_cast<uint32_t *>(s_part + PCAP);                              // [scap] observations | ones << 16 (cluster: this CTA's share of the row)
    uint32_t *s_accm = C > 1 ? s_acc + SCAP : s_acc;                                            // [scap] cluster: the merged accumulator
    int32_t *s_bits = reinterpret_cast<int32_t *>(s_accm + SCAP);                               // [rcap]
    uint32_t *s_row = reinterpret_cast<uint32_t *>(s_bits + RCAP);                              // [2][headk][roww][NT] next / current CSR row heads
    const int KS = GL.kstride;
    const int ROWW = GL.roww;
    const int HEADK = GL.headk > 0 ? GL.headk : 1;
    uint32_t *s_wbits = s_row + 2 * HEADK * ROWW * NT;                                         // [NT] bitmap words of the group-list build
    int32_t *s_woff = reinterpret_cast<int32_t *>(s_wbits + NT);                                // [NT] their list offsets
    const int UBW = GL.ubw;
    uint32_t *s_ub = reinterpret_cast<uint32_t *>(s_woff + NT);                                 // [2][2][UBW] unary-block bit rows (next / current move)
    double *s_logcnt_all = reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(s_ub + 4 * UBW) + 7) & ~(uintptr_t)7);   // [ndom][KS] log(table size)
    int32_t *s_cnt_all = reinterpret_cast<int32_t *>(s_logcnt_all + GL.ndom * KS);               // [ndom][KS]
    int32_t *s_lab_all = s_cnt_all + GL.ndom * KS;
    const int ADCAP = GL.adcap;
    ADesc *s_adesc = reinterpret_cast<ADesc *>(s_lab_all + GL.ndom * KS);                       // [ADCAP] CSR relation descriptors of the domain being moved
    __shared__ int32_t s_ad_dom;                                                                // which domain s_adesc holds
    __shared__ uint8_t *s_zptr[MAX_DOMAINS];                                                    // slot arrays of the domains (the byte shadows)
    if (tid == 0) s_ad_dom = -1;
    if (tid < MAX_DOMAINS) s_zptr[tid] = tid < irm.n_domains ? irm.dom[tid].zb : nullptr;

    const double *logtab = g_logtab;
    for (int k = tid; k < GEN_LF; k += NT) s_lf[k] = g_logfact[k];
    for (int k = tid; k < SCAP; k += NT) { s_acc[k] = 0u; s_accm[k] = 0u; }
    for (int dk = 0; dk < irm.n_domains; dk++) {
        const DomainDev &dq = irm.dom[dk];
        const int h = __ldcg(dq.hi);
        if (tid == 0) s_hi_all[dk] = h;
        for (int k = tid; k < h; k += NT) {
            const int cv = __ldcg(&dq.tab_count[k]);
            s_cnt_all[dk * KS + k] = cv; s_lab_all[dk * KS + k] = __ldcg(&dq.tab_label[k]);
            s_logcnt_all[dk * KS + k] = log((double)max(cv, 1));
        }
        if (tid == 0) s_logalpha[dk] = log(dq.alpha);
    }
    __syncthreads();
    const int NA = NT > 32 ? NT - 32 : NT;        // threads streaming CSR rows (phase A); the last warp prepares the candidates meanwhile
    const int cwarp = NT > 32 ? NW - 1 : 0;

    // A move's inputs are fetched ahead of it with asynchronous copies into a 4-deep shared-memory ring, by thread 0 and
    // at no register cost: the ids three moves ahead, the CSR row bounds two moves ahead, the entity's slot and the head of
    // the row (every thread its own entry) one move ahead.  Thread 0 waits for its copies at the end of each move,
    // before the closing barrier, so everything issued during move m is visible to the block from move m + 1 on.
    __shared__ int32_t s_pd[4], s_pi[4], s_pc[4];
    __shared__ long long s_pb[4][2];
    // this launch's piece of the IRM's move list (the whole list unless the host cut it into rounds); with `pre` the group lists
    // of the piece's moves are staged (pregroup.cuh): phases A and B1 are skipped
    const int64_t mbase = args.move_off[k_list];
    const int64_t m0 = args.seg_lo ? args.seg_lo[k_list] : mbase, m1 = args.seg_hi ? args.seg_hi[k_list] : args.move_off[k_list + 1];
    const bool pre = args.pg_pre != nullptr && args.pg_pre[k_list] != 0;
    const int pg_pitch = pre ? args.pg_pitch[k_list] : 0;
    const uint2 *pg_list = pre ? args.pg_data + args.pg_base[k_list] : nullptr;
    const int PGCAP = HEADK * ROWW * NT / 2;              // staged list entries the row buffer holds
    auto cp4 = [](void *dst, const void *src) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
    };
    auto cp8 = [](void *dst, const void *src) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
    };
    auto issue_ids = [&](int64_t mm) {            // thread 0
        const int k = (int)((mm - m0) & 3);
        if (mm < m1) { cp4(&s_pd[k], args.move_dom + mm); cp4(&s_pi[k], args.move_item + mm); }
        else s_pd[k] = -1;
    };
    auto issue_state = [&](int64_t mm) {          // thread 0; the ids of move mm are in the ring
        const int k = (int)((mm - m0) & 3);
        const int dm = s_pd[k], it = s_pi[k];
        if (mm >= m1 || dm < 0 || dm >= irm.n_domains || it < 0 || it >= irm.dom[dm].n) { s_pd[k] = -1; s_pc[k] = -1; s_pb[k][0] = 0; s_pb[k][1] = 0; return; }
        const DomainDev &dq = irm.dom[dm];
        if (dq.ptr) { cp8(&s_pb[k][0], dq.ptr + it); cp8(&s_pb[k][1], dq.ptr + it + 1); }
        else { s_pb[k][0] = 0; s_pb[k][1] = 0; }
    };
    // the entity's current slot is mutable state: read with ld.cg (L2) one move ahead, by one lane of the candidate warp,
    // and patched at the end of the move if that move re-seats the very same entity
    auto load_slot = [&](int64_t mm) {
        const int k = (int)((mm - m0) & 3);
        const int dm = s_pd[k];
        if (mm < m1 && dm >= 0) s_pc[k] = __ldcg(&irm.dom[dm].z[s_pi[k]]);
    };
    // asynchronous copy of the head of move mm's CSR row (entry q = beg + tid: meta word and other-entity ids) into row buffer `buf`
    auto prefetch_row = [&](int64_t mm, int buf) {
        if (pre) {                                        // the move's staged group list (header + groups), 8 bytes per entry
            if (mm >= m1) return;
            const uint2 *src = pg_list + (mm - m0) * (int64_t)pg_pitch;
            uint2 *dst = reinterpret_cast<uint2 *>(s_row + (size_t)buf * HEADK * ROWW * NT);
            for (int j = tid; j < pg_pitch && j < PGCAP; j += NT) cp8(dst + j, src + j);
            return;
        }
        const int k = (int)((mm - m0) & 3);
        const int dm = s_pd[k];
        if (mm >= m1 || dm < 0 || tid >= NA) return;
        const DomainDev &dq = irm.dom[dm];
        if (!dq.ptr) return;
        // this CTA's first HEADK chunks of the row (chunks crank, crank + C, ...: one entry per thread each)
        for (int j = 0; j < HEADK; j++) {
            const int64_t q = s_pb[k][0] + ((int64_t)crank + (int64_t)j * C) * NA + tid;
            if (q >= s_pb[k][1]) break;
            uint32_t *dst = s_row + (((size_t)buf * HEADK + j) * ROWW) * NT + tid;
            cp4(dst, dq.meta + q);
            for (int oi = 0; oi < dq.n_other && oi < ROWW - 1; oi++) cp4(dst + (oi + 1) * NT, dq.other + (int64_t)oi * dq.nnz + q);
        }
    };
    // ... and of the entity's unary-block bit rows (both planes, every CTA of a cluster the whole row)
    auto prefetch_ub = [&](int64_t mm, int buf) {
        const int k = (int)((mm - m0) & 3);
        const int dm = s_pd[k];
        if (mm >= m1 || dm < 0) return;
        const DomainDev &dq = irm.dom[dm];
        const int nw = dq.ub_words;
        if (tid >= 2 * nw || nw > UBW) return;
        const int plane = tid >= nw, w = tid - plane * nw;
        cp4(s_ub + (buf * 2 + plane) * UBW + w, (plane ? dq.ub_val : dq.ub_obs) + (int64_t)s_pi[k] * nw + w);
    };
    if (tid == 0) {
        issue_ids(m0); issue_ids(m0 + 1); issue_ids(m0 + 2);
        asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
        issue_state(m0); issue_state(m0 + 1);
        asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
        load_slot(m0);
    }
    __syncthreads();
    prefetch_row(m0, 0);
    prefetch_ub(m0, 0);
    asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // debug hook (hirm_engine_debug_phase_cycles): thread 0's SM cycles per phase, summed over the moves:
    // [0] A  [1] B1 (+C)  [2] B2  [3] D  [4] E  [5] F  -- 32 slots per block, shared with the dense kernel's layout
    const bool prof = args.phase_cycles != nullptr && tid == 0;
    long long ph_t = prof ? clock64() : 0, ph_acc[6] = {0, 0, 0, 0, 0, 0};
#define GEN_PHASE(i) do { if (prof) { const long long now = clock64(); ph_acc[i] += now - ph_t; ph_t = now; } } while (0)
    for (int64_t m = m0; m < m1; m++) {
        const int ring = (int)((m - m0) & 3);
        struct { int d, item, c; int64_t beg, end; } cur;
        cur.d = s_pd[ring]; cur.item = s_pi[ring]; cur.c = s_pc[ring]; cur.beg = s_pb[ring][0]; cur.end = s_pb[ring][1];
        const int rbuf = (int)((m - m0) & 1);
        asm volatile("cp.async.wait_group 0;" ::: "memory");      // this move's row head (own element: no barrier needed)
        prefetch_row(m + 1, rbuf ^ 1);
        prefetch_ub(m + 1, rbuf ^ 1);
        if (tid == 0) { issue_ids(m + 3); issue_state(m + 2); }
        asm volatile("cp.async.commit_group;" ::: "memory");
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
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            if (tid == 0) load_slot(m + 1);
            __syncthreads();
            continue;
        }
        if (dd.n_adesc > 0 && dd.n_adesc <= ADCAP && s_ad_dom != d) {     // (uniform) the moved domain changed: its relation descriptors into shared memory
            __syncthreads();
            const int32_t *src = reinterpret_cast<const int32_t *>(dd.adesc);
            int32_t *dst = reinterpret_cast<int32_t *>(s_adesc);
            for (int k = tid; k < dd.n_adesc * 16; k += NT) dst[k] = src[k];
            if (tid == 0) s_ad_dom = d;
            __syncthreads();
        }
        const ADesc *adesc = (dd.n_adesc > 0 && dd.n_adesc <= ADCAP) ? s_adesc : dd.adesc;
        int32_t *s_cnt = s_cnt_all + d * KS, *s_lab = s_lab_all + d * KS;
        const int hi = s_hi_all[d];
        const int ctot = irm.ctot[d];
        const bool acc_shared = ctot <= SCAP && dd.max_deg < 65535;   // the move's accumulator: shared memory, or the IRM's global scratch
        const AccView acc{acc_shared ? s_accm : nullptr, irm.gacc};   // (phase A adds into s_acc; with C = 1 they are the same array)

        // ---- C: candidates, CRP::tables_weights_gibbs (cxx/hirm.hh:147-161): the last warp, while the others stream the row ----
        if (warp == cwarp) {
            if (lane == 0) { s_fail = 0; load_slot(m + 1); }
            __syncwarp();
            const double *s_logcnt = s_logcnt_all + d * KS;
            int maxlab = 0, freeslot = 1 << 30;        // t_max starts at 0 (cxx/hirm.hh:139)
            for (int k = lane; k < hi; k += 32) {
                if (s_cnt[k] > 0) maxlab = max(maxlab, s_lab[k]); else freeslot = min(freeslot, k);
            }
            maxlab = __reduce_max_sync(0xffffffffu, maxlab);
            freeslot = __reduce_min_sync(0xffffffffu, freeslot);
            if (freeslot == (1 << 30)) freeslot = hi < dd.kcap ? hi : -1;
            const bool singleton = s_cnt[c] == 1;
            const double logalpha = s_logalpha[d];
            int ncount = 0;
            for (int base = 0; base < hi; base += 32) {
                const int k = base + lane;
                const int cnt = k < hi ? s_cnt[k] - (k == c ? 1 : 0) : 0;
                const unsigned bal = __ballot_sync(0xffffffffu, cnt > 0);
                const int pos = ncount + __popc(bal & ((1u << lane) - 1u));
                double lw = k < hi ? s_logcnt[k] : 0.0;
                if (k == c) lw = log((double)max(cnt, 1));      // the entity's own table counts without it
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
            if (lane == 0) {
                s_ncand = nc;
                s_u = args.uniforms ? args.uniforms[m] : philox_uniform(args.seed, args.stream[k_list], args.counter0[k_list] + (uint64_t)(m - mbase));
            }
        }

        // ---- A: group (and, for twice-occurring domains, remove) -------------------------------------------------
        if (!pre && dd.ptr && tid < NA) {
            const int64_t beg = cur.beg + (int64_t)crank * NA, end = cur.end;      // cluster: CTA q takes chunks q, q + C, q + 2C, ... of NA entries
            int jchunk = 0;
            for (int64_t q = beg + tid; q < end; q += (int64_t)NA * C, jchunk++) {
                const bool head = jchunk < HEADK;                 // prefetched into shared memory one move ahead
                const uint32_t *rowbuf = s_row + (((size_t)rbuf * HEADK + (head ? jchunk : 0)) * ROWW) * NT + tid;
                const uint32_t meta = head ? rowbuf[0] : dd.meta[q];
                const uint32_t val = (meta >> META_VAL_SHIFT) & 1u, selfmask = (meta >> META_SELF_SHIFT) & 0xFu;
                const ADesc *D = adesc + (meta & META_REL_MASK);      // the entry names its relation by ordinal
                const bool at_a = (selfmask >> D->a) & 1u, at_b = D->b >= 0 && ((selfmask >> D->b) & 1u);
                const int pat = at_a ? (at_b ? 2 : 0) : 1;
                int32_t rest_k = 0, dother = 0;
                int oi = 0; bool first = true, absent = false;
                for (int p = 0; p < D->arity; p++) {
                    int slot;
                    if ((selfmask >> p) & 1u) {
                        if (first) first = false; else oi++;
                        slot = c;
                    } else {
                        const int id = head ? (int)rowbuf[(oi + 1) * NT] : dd.other[(int64_t)oi * dd.nnz + q]; oi++;
                        slot = __ldcg(&s_zptr[D->dom[p]][id]);
                        if (slot == 0xFF) { absent = true; slot = 0; }
                        if (p == D->a || p == D->b) dother = slot; else rest_k += slot * D->rstride[p];
                    }
                }
                if (absent) { raise_error(irm, KERR_ABSENT_ENTITY); continue; }
                const cell_t one = pack_cell(1u, val);
                const int32_t cell = D->base[pat] + rest_k + (pat < 2 ? dother * D->rest_size : 0);
                if (acc_shared) atomicAdd(&s_acc[cell], 1u | (val << 16));
                else { atomicAdd(&irm.gacc[cell], one); atomicOr(&irm.gbm[cell >> 5], 1u << (cell & 31)); }
            }
        }
        // One CTA (or one cluster) owns the IRM: the barrier alone makes the global atomics and stores visible to its threads
        // (they are performed at L2, and every mutable word is read back with ld.cg).  No __threadfence(): a gpu-scope
        // fence would also invalidate L1 and send every descriptor load of the next phase back to L2.
        csync();
        if (!pre && C > 1 && acc_shared) {
            // merge: every CTA adds its non-empty cells into the merged accumulator of all C CTAs (distributed shared memory)
            cg::cluster_group cl = cg::this_cluster();
            for (int cellv = tid; cellv < ctot; cellv += NT) {
                const uint32_t v = s_acc[cellv];
                if (v) {
                    s_acc[cellv] = 0u;
                    for (int r = 0; r < C; r++) atomicAdd(cl.map_shared_rank(s_accm, r) + cellv, v);
                }
            }
            cl.sync();
        }
        GEN_PHASE(0);

        // ---- B1: ordered list of the touched accumulator cells (ascending cell = deterministic order), NT bitmap words at a
        //      time: (1) the words -- a ballot over 32 accumulator cells each, spread over the warps (no bitmap atomics in
        //      phase A), or read from the global bitmap; (2) block scan of their popcounts; (3) every non-empty cell finds its
        //      list position by itself (word offset + popcount of the lower bits) ----
        int64_t base = 0;
        const int nwords = pre ? 0 : (ctot + 31) >> 5;
        const uint2 *s_pg = reinterpret_cast<const uint2 *>(s_row + (size_t)rbuf * HEADK * ROWW * NT);      // (pre) this move's staged list
        const uint2 *g_pg = pre ? pg_list + (m - m0) * (int64_t)pg_pitch : nullptr;
        if (pre) base = (int64_t)s_pg[0].x;
        for (int w0 = 0; w0 < nwords; w0 += NT) {
            const int w = w0 + tid, nw_here = min(NT, nwords - w0);
            if (acc_shared) {
                for (int wl = warp; wl < nw_here; wl += NW) {
                    const int cellv = (w0 + wl) * 32 + lane;
                    const uint32_t b = __ballot_sync(0xffffffffu, cellv < ctot && s_accm[cellv] != 0u);
                    if (lane == 0) s_wbits[wl] = b;
                }
            } else if (tid < nw_here) s_wbits[tid] = __ldcg(&irm.gbm[w]);
            __syncthreads();
            const uint32_t bits = tid < nw_here ? s_wbits[tid] : 0u;
            const int cnt = __popc(bits);
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
            s_woff[tid] = (int)base + woff + incl - cnt;     // list position of this thread's word
            __syncthreads();
            for (int cl = tid; cl < nw_here * 32; cl += NT) {
                const int wl = cl >> 5, bit = cl & 31;
                const uint32_t bj = s_wbits[wl];
                if ((bj >> bit) & 1u) {
                    const int pos = s_woff[wl] + __popc(bj & ((1u << bit) - 1u));
                    const int cellv = (w0 + wl) * 32 + bit;
                    if (pos >= glcap) raise_error(irm, KERR_GROUP_OVERFLOW);
                    else if (pos < RCAP) s_bits[pos] = cellv;
                    else recs[pos].cell = cellv;
                }
            }
            base += total;
            __syncthreads();
        }
        const int64_t ngroups = base < glcap ? base : glcap;
        GEN_PHASE(1);
        // ---- B2: decode every group once, one group per thread (for every candidate) ---------------------------------
        const GroupRec *tmpl = reinterpret_cast<const GroupRec *>(irm.ctmpl[d]);
        for (int64_t g = tid; g < ngroups; g += NT) {
            int32_t cell; cell_t own;
            if (pre) {
                const uint2 ent = g + 1 < PGCAP ? s_pg[g + 1] : __ldcg(g_pg + g + 1);
                cell = (int32_t)ent.x; own = pack_cell(ent.y & 0xFFFFu, ent.y >> 16);
            } else {
                cell = g < RCAP ? s_bits[g] : recs[g].cell;
                own = acc[cell];
            }
            GroupRec rc = tmpl ? tmpl[cell] : decode_cell(irm, d, cell);
            rc.own = own;
            if (g < RCAP) s_recs[g] = rc; else recs[g] = rc;
        }
        __syncthreads();
        GEN_PHASE(2);
        const int ncand = s_ncand;

        // ---- D: score.  Work unit = (candidate, 32 consecutive groups), one warp each, lanes over the groups; the partial
        //      sums are added per candidate in chunk order afterwards (fixed order: deterministic) -------------------------
        // operands of one (group, candidate) pair in two steps, so that the count-table load of the NEXT unit is in flight
        // while the current unit's fp64 work runs: issue() resolves the target cell and starts the load, finish() consumes it
        struct Pending { cell_t cur, delta, sub; };
        auto issue = [&](int64_t g, int slot_t) -> Pending {
            Pending p{0ull, 0ull, 0ull};
            if (g >= ngroups) return p;
            const GroupRec gr = g < RCAP ? s_recs[g] : recs[g];
            int64_t cidx;
            if (!group_target(gr, acc, slot_t, c, p.delta, cidx, p.sub)) { p.delta = 0ull; p.sub = 0ull; return p; }
            p.cur = table_get(gr.counts, gr.hcap, cidx);   // virtual removal: finish() takes the entity's own observations out of the cell
            return p;
        };
        auto finish = [&](const Pending &p, uint32_t &N, uint32_t &sv, uint32_t &gn, uint32_t &gs) {
            const cell_t curc = p.cur - p.sub;
            N = cell_n(curc); sv = cell_s(curc); gn = cell_n(p.delta); gs = cell_s(p.delta);
            if (gn == 0u) { N = 0u; sv = 0u; gs = 0u; }
        };
        auto eval = [&](int64_t g, int slot_t, uint32_t &N, uint32_t &sv, uint32_t &gn, uint32_t &gs) {
            finish(issue(g, slot_t), N, sv, gn, gs);
        };
        // The IRM's unary-block relations are not groups: their count tables are one matrix (table slot x local column) and
        // the entity's observations two bit rows (copied into shared memory a move ahead), so a scoring unit of 32 of them
        // is one coalesced row read.  Units of a candidate: its nch group chunks, then its nuch unary chunks.
        const int nu = dd.nu, nuch = (nu + 31) >> 5;
        const bool ub_staged = dd.ub_words <= UBW;
        const uint32_t *xo = ub_staged ? s_ub + (rbuf * 2) * UBW : dd.ub_obs + (int64_t)item * dd.ub_words;
        const uint32_t *xv = ub_staged ? s_ub + (rbuf * 2 + 1) * UBW : dd.ub_val + (int64_t)item * dd.ub_words;
        auto issue_unary = [&](int ch, int slot_t) -> Pending {
            Pending p{0ull, 0ull, 0ull};
            const int j = ch * 32 + lane;
            if (j >= nu) return p;
            const int gcol = __ldg(dd.ucolg + j);
            if (!((xo[gcol >> 5] >> (gcol & 31)) & 1u)) return p;            // the relation has no observation of this entity
            p.delta = pack_cell(1u, (xv[gcol >> 5] >> (gcol & 31)) & 1u);
            p.cur = __ldcg(dd.ucount + (int64_t)slot_t * dd.ucp + j);
            if (slot_t == c) p.sub = p.delta;                                // virtual removal
            return p;
        };
        const int nch = (int)((ngroups + 31) >> 5), nchu = nch + nuch;
        auto issue_unit = [&](int t0, int unit) -> Pending {                 // unit = (candidate - t0) * nchu + chunk
            const int tb = unit / nchu, ch = unit - tb * nchu;
            return ch < nch ? issue((int64_t)ch * 32 + lane, s_cslot[t0 + tb]) : issue_unary(ch - nch, s_cslot[t0 + tb]);
        };
        // The units of ALL candidates at once where their partial sums fit the buffer; otherwise in batches of whole candidates that
        // alternate between the two halves of the buffer (a half is rewritten two batches later, behind the barrier of the batch in
        // between: one cluster barrier per batch).  An IRM early in a chain can have 40+ tables: 48 candidates x 90 units, which the
        // one-warp-per-candidate fallback below would leave to 48 of a cluster's 512 warps.
        const int HP = PCAP >> 1;
        const bool flat = nchu <= HP;
        if (flat) {
            const bool one = (int64_t)ncand * nchu <= PCAP;
            const int TB = one ? ncand : HP / nchu;            // candidates per batch
            int half = 0;
            for (int t0 = 0; t0 < ncand; t0 += TB, half ^= 1) {
                const int tbn = min(TB, ncand - t0), nunits = tbn * nchu;
                double *part = s_part + (one ? 0 : half * HP);
                const int u0 = crank + warp * C, ustep = NW * C;       // unit u belongs to CTA u mod C, warp (u / C) mod NW
                Pending pend{0ull, 0ull, 0ull};
                if (u0 < nunits) pend = issue_unit(t0, u0);
                for (int unit = u0; unit < nunits; unit += ustep) {
                    const Pending cur = pend;
                    const int unext = unit + ustep;
                    if (unext < nunits) pend = issue_unit(t0, unext);
                    uint32_t N, sv, gn, gs;
                    finish(cur, N, sv, gn, gs);
                    __syncwarp();
                    double v = score_delta_any(s_lf, logtab, N, sv, gn, gs);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (C == 1) { if (lane == 0) part[unit] = v; }
                    else if (lane < C) cg::this_cluster().map_shared_rank(part, lane)[unit] = v;     // lane r stores into CTA r's buffer
                }
                csync();
                for (int tb = tid; tb < tbn; tb += NT) {
                    double a2 = 0.0;
                    for (int ch = 0; ch < nchu; ch++) a2 += part[tb * nchu + ch];
                    s_lp[t0 + tb] += a2;
                }
            }
        } else {
            for (int t = crank + warp * C; t < ncand; t += NW * C) {   // very many groups: one warp per candidate walks all of them
                const int slot_t = s_cslot[t];
                double a2 = 0.0;
                for (int64_t g0 = 0; g0 < ngroups; g0 += 32) {
                    uint32_t N, sv, gn, gs;
                    eval(g0 + lane, slot_t, N, sv, gn, gs);
                    __syncwarp();
                    double v = score_delta_any(s_lf, logtab, N, sv, gn, gs);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    a2 += v;
                }
                for (int ch = 0; ch < nuch; ch++) {
                    uint32_t N, sv, gn, gs;
                    finish(issue_unary(ch, slot_t), N, sv, gn, gs);
                    __syncwarp();
                    double v = score_delta_unit(logtab, N, sv, gn, gs);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    a2 += v;
                }
                if (C == 1) { if (lane == 0) s_part[t] = a2; }
                else if (lane < C) cg::this_cluster().map_shared_rank(s_part, lane)[t] = a2;
            }
            csync();
            for (int t = tid; t < ncand; t += NT) s_lp[t] += s_part[t];
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
                    const double u = s_u;
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
        // ---- F: apply: groups of a once-occurring domain move only if the entity does; the others are put back ----------
        const int tnew = s_choice_slot;
        for (int64_t g = tid; g < ngroups; g += NT) {
            const GroupRec gr = g < RCAP ? s_recs[g] : recs[g];
            const bool mine = C == 1 || (int)(g % C) == crank;       // cluster: the table updates of group g are CTA (g mod C)'s
            if (tnew != c && mine) {                  // Relation::set_cluster_assignment_gibbs (:524-551): every observation leaves its cell for the new one
                table_add(irm, gr.R, group_cell_at(gr, c), (cell_t)0 - gr.own);
                table_add(irm, gr.R, group_cell_at(gr, tnew), gr.own);
            }
            if (pre) continue;                        // (no accumulator in use)
            if (acc_shared) s_accm[gr.cell] = 0u;     // every CTA clears its own (merged) accumulator
            else if (mine) { __stcg(&irm.gacc[gr.cell], (cell_t)0); __stcg(&irm.gbm[gr.cell >> 5], 0u); }
        }
        if (tnew != c) {                                   // the unary-block relations' cells of the two tables, one thread per column
            for (int j = crank * NT + tid; j < nu; j += NT * C) {
                const int gcol = __ldg(dd.ucolg + j);
                if (!((xo[gcol >> 5] >> (gcol & 31)) & 1u)) continue;
                const cell_t one = pack_cell(1u, (xv[gcol >> 5] >> (gcol & 31)) & 1u);
                atomicAdd(dd.ucount + (int64_t)c * dd.ucp + j, (cell_t)0 - one);
                atomicAdd(dd.ucount + (int64_t)tnew * dd.ucp + j, one);
            }
        }
        // z and the CRP state: written by every CTA of a cluster (identical values), so each CTA reads back its own writes
        if (tid == 0 && tnew != c) {
            __stcg(&dd.z[item], tnew);
            __stcg(&dd.zb[item], (uint8_t)tnew);
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
            s_logcnt_all[d * KS + c] = log((double)max(s_cnt[c], 1));
            s_logcnt_all[d * KS + tnew] = log((double)newcnt);
        }
        // every thread's copies for the next move (ring entries, row head, unary bit rows -- read by OTHER threads too) have
        // landed before the closing barrier
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (tid == 0) {
            const int k1 = (ring + 1) & 3;             // the same entity again: its prefetched slot is stale
            if (tnew != c && s_pd[k1] == d && s_pi[k1] == item) s_pc[k1] = tnew;
        }
        if (C > 1 && !acc_shared && !pre) cg::this_cluster().sync();   // global accumulator: cleared before any CTA's next phase A adds to it
        else __syncthreads();
        GEN_PHASE(5);
    }
    if (prof && crank == 0) for (int i = 0; i < 6; i++) args.phase_cycles[k_blk * 32 + i] += ph_acc[i];   // (summed over the rounds of a sweep)
    if (C > 1) cg::this_cluster().sync();                      // no CTA exits while another may still address its shared memory
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
        table_add(irm, R, cidx, pack_cell(1u, R->coo_val[e]));
    }
}

// Sum of BetaBernoulli::logp_score over the cells of one relation   (Relation::logp_score, cxx/hirm.hh:516-522)
// Single block, fixed reduction tree: deterministic.
__global__ void __launch_bounds__(1024) relation_score_kernel(const IrmDev *irms, int irm_id, int rel, double *out) {
    const RelationDev *R = &irms[irm_id].rel[rel];
    __shared__ double s_part[32];
    double acc = 0.0;
    for (int64_t i = threadIdx.x, ne = table_entries(R); i < ne; i += blockDim.x) {
        const cell_t c = table_entry(R, i);
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
