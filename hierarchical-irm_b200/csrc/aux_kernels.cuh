// Small batched kernels around the sweep: sweep-order generation, IRM::logp_score, CRP::transition_alpha.
// One thread block per IRM; fixed reduction trees (deterministic results).
#pragma once
#include "engine_types.cuh"

namespace hirm {

constexpr int AUX_THREADS = 256;
constexpr uint64_t ALPHA_SEED_TAG = 0x616c706861ull;   // "alpha": separates the alpha moves' Philox key from the entity moves'
constexpr uint64_t ORDER_SEED_TAG = 0x6f72646572ull;   // "order"

__device__ __forceinline__ double block_sum(double v, double *s_part) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += s_part[w];      // same order in every thread
    return t;
}

// CRP::logp_score (cxx/hirm.hh:123-132) for table sizes cnt[0..hi): K log(alpha) + sum lgamma(cnt) + lgamma(alpha) - lgamma(N + alpha)
__device__ __forceinline__ double crp_score_terms(const int32_t *cnt, int hi, int &K, int &N) {
    K = 0; N = 0;
    double t2 = 0.0;
    for (int k = 0; k < hi; k++) {
        const int c = cnt[k];
        if (c > 0) { K++; N += c; t2 += lgamma((double)c); }
    }
    return t2;
}

// IRM::logp_score (cxx/hirm.hh:730-740) = sum_d CRP::logp_score + sum_r Relation::logp_score (:516-522)
__global__ void __launch_bounds__(AUX_THREADS) logp_scores_kernel(const IrmDev *irms, const int32_t *irm_ids, double *out) {
    const IrmDev &irm = irms[irm_ids[blockIdx.x]];
    __shared__ double s_part[AUX_THREADS / 32];
    double acc = 0.0;
    for (int r = 0; r < irm.n_relations; r++) {
        const RelationDev *R = &irm.rel[r];
        for (int64_t i = threadIdx.x, ne = table_entries(R); i < ne; i += blockDim.x) {
            const cell_t c = table_entry(R, i);
            if (cell_n(c)) acc += cell_score(cell_n(c), cell_s(c));
        }
    }
    double total = block_sum(acc, s_part);
    if (threadIdx.x == 0) {
        // CRP terms of the domains some relation of this IRM uses: the reference's IRM holds no other domain
        // (IRM::add_relation creates them, remove_relation drops a domain with its last relation: cxx/hirm.hh:742-782),
        // while an IRM of the sharded HIRM carries every domain of the schema
        uint32_t used = 0u;
        for (int r = 0; r < irm.n_relations; r++)
            for (int p = 0; p < irm.rel[r].arity; p++) used |= 1u << irm.rel[r].dom[p];
        for (int d = 0; d < irm.n_domains; d++) {
            if (!((used >> d) & 1u)) continue;
            const DomainDev &dd = irm.dom[d];
            int K, N;
            const double t2 = crp_score_terms(dd.tab_count, *dd.hi, K, N);
            if (N > 0) total += K * log(dd.alpha) + t2 + lgamma(dd.alpha) - lgamma(N + dd.alpha);
        }
        out[blockIdx.x] = total;
    }
}

// CRP::transition_alpha (cxx/hirm.hh:162-173; src/hirm.py:109-116): grid-Gibbs over log_linspace(1/N, N+1, ngrid)
// (cxx/util_math.cc:17-32), the index drawn with the engine's inversion rule.  Writes alpha into the device
// descriptor and out_alpha[block * MAX_DOMAINS + d]; optional trace of the grid log-probabilities.
__global__ void __launch_bounds__(64) transition_alpha_kernel(IrmDev *irms, const int32_t *irm_ids, int ngrid, uint64_t seed,
                                                              const uint64_t *stream, const uint64_t *counter, const double *uniforms,
                                                              double *out_alpha, double *trace_lp /* nullable [blocks][MAX_DOMAINS][ngrid] */) {
    IrmDev &irm = irms[irm_ids[blockIdx.x]];
    __shared__ double s_lp[64];
    __shared__ double s_grid[64];
    for (int d = 0; d < irm.n_domains; d++) {
        DomainDev &dd = irm.dom[d];
        int K, N;
        const double t2 = crp_score_terms(dd.tab_count, *dd.hi, K, N);       // every thread: the same value
        double alpha_new = dd.alpha;
        if (N > 0) {                                                          // `if (N == 0) return;` (cxx/hirm.hh:163)
            const double lo = log(1.0 / N), hi = log((double)N + 1.0);
            const double step = (hi - lo) / (ngrid - 1);
            const int g = threadIdx.x;
            if (g < ngrid) {
                const double a = exp(lo + step * g);
                s_grid[g] = a;
                s_lp[g] = K * log(a) + t2 + lgamma(a) - lgamma(N + a);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                const double u = uniforms ? uniforms[blockIdx.x * MAX_DOMAINS + d]
                                          : philox_uniform(seed ^ ALPHA_SEED_TAG, stream[blockIdx.x], counter[blockIdx.x] * MAX_DOMAINS + d);
                alpha_new = s_grid[choose_index(s_lp, ngrid, u)];
                dd.alpha = alpha_new;
            }
            if (trace_lp && threadIdx.x < ngrid) trace_lp[((int64_t)blockIdx.x * MAX_DOMAINS + d) * ngrid + threadIdx.x] = s_lp[threadIdx.x];
            __syncthreads();
        }
        if (threadIdx.x == 0) out_alpha[blockIdx.x * MAX_DOMAINS + d] = alpha_new;
    }
}

// ---------------------------------------------------------------------------------------
// K4: Relation::logp_score of relation (src irm, rel) under the domain partitions of OTHER IRMs -- what
// HIRM::transition_cluster_assignment_relation computes for every candidate table by re-incorporating the
// relation's data there (cxx/hirm.hh:859-866; src/hirm.py:497-505).  The count table of the relation under a
// partition is a histogram, so nothing is moved: pass 1 bins the observations by the candidate's z arrays into a
// scratch table, pass 2 sums BetaBernoulli::logp_score over the cells (fixed reduction tree).
struct RelCand {
    const int32_t *z[MAX_ARITY];   // candidate's slot array per position of the relation
    int64_t stride[MAX_ARITY];     // scratch-table stride per position
    cell_t *table;                 // scratch count table (zeroed), `ncells` cells
    int64_t ncells;
};
__global__ void __launch_bounds__(AUX_THREADS) rel_hist_kernel(const int32_t *coo_ids, const uint8_t *coo_val, int64_t nobs, int arity,
                                                               const RelCand *cands, int32_t *error) {
    const RelCand &C = cands[blockIdx.y];
    for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < nobs; o += (int64_t)gridDim.x * blockDim.x) {
        int64_t idx = 0; bool absent = false;
        for (int p = 0; p < arity; p++) {
            const int slot = C.z[p][coo_ids[o * arity + p]];
            if (slot < 0) absent = true;
            idx += (slot < 0 ? 0 : slot) * C.stride[p];
        }
        if (absent) { atomicCAS(error, 0, KERR_ABSENT_ENTITY); continue; }
        atomicAdd(&C.table[idx], pack_cell(1u, coo_val[o]));
    }
}
__global__ void __launch_bounds__(AUX_THREADS) rel_score_kernel(const RelCand *cands, double *out) {
    const RelCand &C = cands[blockIdx.x];
    __shared__ double s_part[AUX_THREADS / 32];
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < C.ncells; i += blockDim.x) {
        const cell_t c = C.table[i];
        if (cell_n(c)) acc += cell_score(cell_n(c), cell_s(c));
    }
    const double total = block_sum(acc, s_part);
    if (threadIdx.x == 0) out[blockIdx.x] = total;
}

// ---------------------------------------------------------------------------------------
// K4 batched: the whole [relations x candidate partitions] matrix of Relation::logp_score values that one
// sweep of HIRM::transition_cluster_assignment_relation needs (cxx/hirm.hh:859-866), in two launches.
// The score of relation r under IRM k depends only on r's observations and k's domain partitions, never on which
// other relations sit in k, so it can be computed for every pair before the sequential re-seating of the relations.
// Observations are relation-owned COO arrays (caller's device buffers); a candidate is a set of slot arrays, one
// per domain.  Work item = (relation, run of candidates, run of observations): the COO run is read once for the
// whole run of candidates; histograms live in shared memory when (candidates x cells) fits, and are flushed to the
// global scratch tables with integer atomics (exact, order-free).  Pass 2 sums the cell scores with a fixed tree.
constexpr int RM_SH_CELLS = 4096;       // shared-memory histogram cells per work item (2 x uint32 each) when every table is small
constexpr int RM_SH_CELLS_BIG = 24576;  // ... when some table needs more (192 KB of shared memory, one CTA per SM)
struct RelDesc {
    const int32_t *ids; const uint8_t *val; int64_t nobs;
    int32_t arity; int32_t dom[MAX_ARITY];
    int64_t stride[MAX_ARITY];          // scratch-table stride per position
    int64_t ncells;
};
struct PartSet { const int32_t *z[MAX_DOMAINS]; };
struct ScoreWork {
    int32_t rel, cand0, ncand, pad;     // candidates cand0 .. cand0 + ncand - 1 of the partition-set array
    int64_t obs0, obs1;
    int64_t table0;                     // scratch offset of (rel, cand0); candidate c of the run sits at table0 + c * ncells
};
__global__ void __launch_bounds__(AUX_THREADS) relmat_hist_kernel(const RelDesc *rels, const PartSet *parts, const ScoreWork *work,
                                                                  cell_t *tables, int32_t *error, int sh_cells) {
    const ScoreWork W = work[blockIdx.x];
    const RelDesc R = rels[W.rel];
    extern __shared__ uint32_t rm_dyn[];
    uint32_t *sh_n = rm_dyn, *sh_s = rm_dyn + sh_cells;
    const int64_t run_cells = R.ncells * W.ncand;
    const bool use_sh = run_cells <= sh_cells;
    if (use_sh) {
        for (int i = threadIdx.x; i < run_cells; i += blockDim.x) { sh_n[i] = 0u; sh_s[i] = 0u; }
        __syncthreads();
    }
    for (int64_t o = W.obs0 + threadIdx.x; o < W.obs1; o += blockDim.x) {
        int32_t id[MAX_ARITY];
        for (int p = 0; p < R.arity; p++) id[p] = R.ids[o * R.arity + p];
        const uint32_t v = R.val[o];
        for (int c = 0; c < W.ncand; c++) {
            const PartSet &P = parts[W.cand0 + c];
            int64_t idx = 0; bool absent = false;
            for (int p = 0; p < R.arity; p++) {
                const int slot = P.z[R.dom[p]][id[p]];
                if (slot < 0) absent = true;
                idx += (slot < 0 ? 0 : slot) * R.stride[p];
            }
            if (absent) { atomicCAS(error, 0, KERR_ABSENT_ENTITY); continue; }
            if (use_sh) {
                const int j = (int)(c * R.ncells + idx);
                atomicAdd(&sh_n[j], 1u);
                if (v) atomicAdd(&sh_s[j], 1u);
            } else {
                atomicAdd(&tables[W.table0 + c * R.ncells + idx], pack_cell(1u, v));
            }
        }
    }
    if (use_sh) {
        __syncthreads();
        for (int i = threadIdx.x; i < run_cells; i += blockDim.x)
            if (sh_n[i]) atomicAdd(&tables[W.table0 + i], pack_cell(sh_n[i], sh_s[i]));
    }
}
struct TableRef { int64_t off, ncells, out; };   // scratch table, its size, index of the score in the output array
__global__ void __launch_bounds__(AUX_THREADS) relmat_score_kernel(const TableRef *refs, const cell_t *tables, double *out) {
    const TableRef T = refs[blockIdx.x];
    __shared__ double s_part[AUX_THREADS / 32];
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < T.ncells; i += blockDim.x) {
        const cell_t c = tables[T.off + i];
        if (cell_n(c)) acc += cell_score(cell_n(c), cell_s(c));
    }
    const double total = block_sum(acc, s_part);
    if (threadIdx.x == 0) out[T.out] = total;
}

// K4 for UNARY relations (animals.unary: all 85 relations; config 5: 2000 of 2050): no histogram at all.  A unary relation
// over a domain of n entities is two bit vectors over the entity codes (observed, value = 1); a table of a candidate
// partition is a member bit vector; then  N_t = popc(observed & members_t),  s_t = popc(value & members_t).
// CTA = (candidate, block of relations).  Tables go in passes of 32 (lane = table); the entity range in tiles whose
// member masks are built in shared memory with ballots from the candidate's slot array (coalesced) and shared by all
// relations of the block; every lane keeps (N, s) of its table for the warp's UN_RW relations in registers across the
// tiles -- no atomics, no reductions until the final sum of the 32 cell scores.
constexpr int UN_THREADS = 256, UN_WARPS = UN_THREADS / 32, UN_RW = 8;   // 64 relations per CTA
constexpr int UN_TILE = 1024;                                            // entity words per tile (32768 entities)
struct UnaryRel { const uint32_t *obs, *val; int32_t out_row, pad; };
__global__ void unary_bits_kernel(const int32_t *ids, const uint8_t *val, int64_t nobs, uint32_t *obs_bits, uint32_t *val_bits) {
    for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < nobs; o += (int64_t)gridDim.x * blockDim.x) {
        const int32_t i = ids[o];
        atomicOr(&obs_bits[i >> 5], 1u << (i & 31));
        if (val[o]) atomicOr(&val_bits[i >> 5], 1u << (i & 31));
    }
}
__global__ void __launch_bounds__(UN_THREADS) relmat_unary_kernel(const UnaryRel *rels, int n_rels, const PartSet *parts, int dom, int n_ent,
                                                                  int n_cand, double *out, int32_t *error) {
    extern __shared__ uint32_t un_masks[];                 // [32][UN_TILE + 1]
    __shared__ int32_t s_red[UN_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cand = blockIdx.x;
    const int32_t *z = parts[cand].z[dom];
    const int nwords = (n_ent + 31) >> 5;
    // highest slot in use (+1) and a check that every entity is seated
    int tmax = 0, bad = 0;
    for (int i = tid; i < n_ent; i += UN_THREADS) { const int v = z[i]; tmax = max(tmax, v + 1); bad |= v < 0; }
    tmax = __reduce_max_sync(0xffffffffu, tmax);
    bad = __reduce_max_sync(0xffffffffu, bad);
    if (lane == 0) s_red[warp] = tmax | (bad << 30);
    __syncthreads();
    tmax = 0;
    for (int w = 0; w < UN_WARPS; w++) { tmax = max(tmax, s_red[w] & ((1 << 30) - 1)); bad |= s_red[w] >> 30; }
    if (bad) { if (tid == 0) atomicCAS(error, 0, KERR_ABSENT_ENTITY); return; }
    const int r0 = (blockIdx.y * UN_WARPS + warp) * UN_RW;          // this warp's relations r0 .. r0 + UN_RW - 1
    double total[UN_RW];
#pragma unroll
    for (int k = 0; k < UN_RW; k++) total[k] = 0.0;
    for (int t0 = 0; t0 < tmax; t0 += 32) {
        const int tcount = min(32, tmax - t0);
        uint32_t nacc[UN_RW], sacc[UN_RW];
#pragma unroll
        for (int k = 0; k < UN_RW; k++) { nacc[k] = 0u; sacc[k] = 0u; }
        for (int w0 = 0; w0 < nwords; w0 += UN_TILE) {
            const int wt = min(UN_TILE, nwords - w0);
            __syncthreads();                                         // the previous tile's masks are no longer read
            for (int w = warp; w < wt; w += UN_WARPS) {              // member masks of tables t0 .. t0 + tcount - 1
                const int i = (w0 + w) * 32 + lane;
                const int zl = i < n_ent ? z[i] - t0 : -1;
                uint32_t mine = 0u;
                for (int t = 0; t < tcount; t++) {
                    const uint32_t b = __ballot_sync(0xffffffffu, zl == t);
                    if (lane == t) mine = b;
                }
                if (lane < tcount) un_masks[lane * (UN_TILE + 1) + w] = mine;
            }
            __syncthreads();
            const uint32_t *mrow = un_masks + lane * (UN_TILE + 1);
#pragma unroll
            for (int k = 0; k < UN_RW; k++) {
                if (r0 + k >= n_rels) break;
                const uint32_t *ob = rels[r0 + k].obs + w0, *vb = rels[r0 + k].val + w0;
                uint32_t na = 0u, sa = 0u;
                for (int w = 0; w < wt; w++) {
                    const uint32_t m = lane < tcount ? mrow[w] : 0u;
                    na += __popc(__ldg(ob + w) & m);
                    sa += __popc(__ldg(vb + w) & m);
                }
                nacc[k] += na; sacc[k] += sa;
            }
        }
#pragma unroll
        for (int k = 0; k < UN_RW; k++) {
            if (r0 + k >= n_rels) break;
            double v = nacc[k] ? cell_score(nacc[k], sacc[k]) : 0.0;   // BetaBernoulli::logp_score of table t0 + lane's cell
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            total[k] += v;
        }
    }
#pragma unroll
    for (int k = 0; k < UN_RW; k++)
        if (r0 + k < n_rels && lane == 0) out[(int64_t)rels[r0 + k].out_row * n_cand + cand] = total[k];
}

// A draw from the sequential CRP(alpha) prior over n customers -- what Domain::incorporate (cxx/hirm.hh:194-203 ->
// CRP::sample :101-113) does to the entities of a relation that opens a fresh IRM (:849-864) -- without the sequential
// loop: customer i opens a table with probability alpha / (i + alpha), otherwise joins the table of a uniformly drawn
// earlier customer (the two descriptions of the CRP are the same law); tables are labelled in order of appearance
// (new label = tables so far, as CRP::tables_weights :133-146 gives).  One CTA per draw: Philox per customer,
// block prefix count of the openers, in-place pointer jumping to the opener of each customer's table.
constexpr uint64_t PRIOR_SEED_TAG = 0x7072696f72ull;    // "prior"
constexpr int PRIOR_THREADS = 1024;
struct PriorDraw { int32_t *z; int32_t *cum; int32_t n; int32_t kcap; uint64_t key; double alpha; };
__global__ void __launch_bounds__(PRIOR_THREADS) crp_prior_kernel(const PriorDraw *draws, uint64_t seed, int32_t *error) {
    const PriorDraw D = draws[blockIdx.x];
    __shared__ int32_t s_warp[PRIOR_THREADS / 32];
    __shared__ int32_t s_base, s_changed;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < D.n; i0 += PRIOR_THREADS) {
        const int i = i0 + tid;
        int opener = 0;
        if (i < D.n) {
            const double u = philox_uniform(seed ^ PRIOR_SEED_TAG, D.key, 2ull * (uint64_t)i);
            const double v = philox_uniform(seed ^ PRIOR_SEED_TAG, D.key, 2ull * (uint64_t)i + 1ull);
            opener = i == 0 || u * ((double)i + D.alpha) < D.alpha;
            int par = i;
            if (!opener) { par = (int)(v * (double)i); if (par >= i) par = i - 1; }
            D.z[i] = par;
        }
        int incl = opener;
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        int woff = 0, total = 0;
        for (int w = 0; w < PRIOR_THREADS / 32; w++) { const int t = s_warp[w]; if (w < warp) woff += t; total += t; }
        if (i < D.n) D.cum[i] = s_base + woff + incl;     // openers among customers 0..i
        __syncthreads();
        if (tid == 0) s_base += total;
        __syncthreads();
    }
    if (tid == 0 && s_base > D.kcap) atomicCAS(error, 0, KERR_NO_FREE_SLOT);
    // pointer jumping: z[i] <- z[z[i]] until every customer points at an opener (pointers only move towards openers,
    // so reading a neighbour's old or new value is equally good)
    for (;;) {
        __syncthreads();
        if (tid == 0) s_changed = 0;
        __syncthreads();
        int changed = 0;
        for (int i = tid; i < D.n; i += PRIOR_THREADS) {
            const int p = ((volatile int32_t *)D.z)[i], pp = ((volatile int32_t *)D.z)[p];
            if (pp != p) { ((volatile int32_t *)D.z)[i] = pp; changed = 1; }
        }
        if (changed) s_changed = 1;
        __syncthreads();
        if (!s_changed) break;
    }
    for (int i = tid; i < D.n; i += PRIOR_THREADS) {
        const int slot = D.cum[D.z[i]] - 1;              // own entry in, own entry out: no hazard between threads
        D.z[i] = slot < D.kcap ? slot : D.kcap - 1;
    }
}

// One column of a unary block from the relation's COO arrays: bit `col` of the observed / value rows of every entity the
// relation observes.  err: 1 = code or value out of range, 2 = an entity observed twice (cxx/hirm.hh:272 asserts distinct).
__global__ void unary_block_fill_kernel(const int32_t *ids, const uint8_t *val, int64_t nobs, int32_t n_ent, int col, int words,
                                        uint32_t *obs_rows, uint32_t *val_rows, int32_t *err) {
    const uint32_t bit = 1u << (col & 31);
    for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < nobs; o += (int64_t)gridDim.x * blockDim.x) {
        const int32_t i = ids[o];
        const uint8_t v = val[o];
        if (i < 0 || i >= n_ent || v > 1) { atomicCAS(err, 0, 1); continue; }
        const int64_t w = (int64_t)i * words + (col >> 5);
        if (atomicOr(&obs_rows[w], bit) & bit) atomicCAS(err, 0, 2);
        if (v) atomicOr(&val_rows[w], bit);
    }
}

// byte shadow of a slot array (DomainDev::zb)
__global__ void slots_to_bytes_kernel(const int32_t *z, uint8_t *zb, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { const int v = z[i]; zb[i] = v < 0 ? (uint8_t)0xFF : (uint8_t)v; }
}

// CRP state of a domain from its slot array: customers per slot, label = slot, highest used slot + 1 -- after a device-side
// prior draw (hirm_engine_seat_prior).  One CTA per (irm, domain) pair listed in `pairs` (irm id, domain).
__global__ void __launch_bounds__(256) tables_from_slots_kernel(const IrmDev *irms, const int32_t *pairs) {
    const IrmDev &irm = irms[pairs[2 * blockIdx.x]];
    const DomainDev &dd = irm.dom[pairs[2 * blockIdx.x + 1]];
    __shared__ int32_t s_cnt[256];
    for (int k = threadIdx.x; k < 256; k += blockDim.x) s_cnt[k] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < dd.n; i += blockDim.x) { const int zi = dd.z[i]; if (zi >= 0) atomicAdd(&s_cnt[zi], 1); }
    __syncthreads();
    int hi = 0;
    for (int k = threadIdx.x; k < dd.kcap; k += blockDim.x) { dd.tab_count[k] = s_cnt[k]; dd.tab_label[k] = k; if (s_cnt[k] > 0) hi = k + 1; }
    hi = __reduce_max_sync(0xffffffffu, hi);
    __shared__ int32_t s_hi;
    if (threadIdx.x == 0) s_hi = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) atomicMax(&s_hi, hi);
    __syncthreads();
    if (threadIdx.x == 0) *dd.hi = s_hi;
}

// ---------------------------------------------------------------------------------------
// Per-entity CSR build on the device (Relation::data_r, cxx/hirm.hh:251,274-280): count, scan, fill.  An observation
// is listed once per entity even if the entity sits at two positions of it (set semantics of data_r).  The order of
// the entries inside a row is not specified (the sweep kernel's results do not depend on it: integer accumulators).
struct CsrRel { const int32_t *ids; const uint8_t *val; int64_t nobs; int32_t arity; int32_t rel; int32_t dom[MAX_ARITY]; int32_t n_ent[MAX_ARITY]; };
__global__ void coo_validate_kernel(CsrRel R, int32_t *error) {
    for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < R.nobs; o += (int64_t)gridDim.x * blockDim.x) {
        bool bad = R.val[o] > 1;
        for (int p = 0; p < R.arity; p++) { const int32_t v = R.ids[o * R.arity + p]; bad |= v < 0 || v >= R.n_ent[p]; }
        if (bad) atomicCAS(error, 0, 1);
    }
}
__global__ void csr_count_kernel(CsrRel R, int d, unsigned long long *deg) {
    for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < R.nobs; o += (int64_t)gridDim.x * blockDim.x) {
        for (int p = 0; p < R.arity; p++) {
            if (R.dom[p] != d) continue;
            const int32_t ent = R.ids[o * R.arity + p];
            bool dup = false;
            for (int q = 0; q < p; q++) dup |= R.dom[q] == d && R.ids[o * R.arity + q] == ent;
            if (!dup) atomicAdd(&deg[ent], 1ull);
        }
    }
}
// in-place exclusive scan of m 64-bit counts by one block; *max_out = the largest count
__global__ void __launch_bounds__(1024) scan_exclusive_kernel(unsigned long long *a, int64_t m, unsigned long long *max_out) {
    __shared__ unsigned long long s_sum[1024], s_max[1024];
    const int64_t per = (m + 1023) / 1024, lo = threadIdx.x * per, hi = lo + per < m ? lo + per : m;
    unsigned long long sum = 0, mx = 0;
    for (int64_t i = lo; i < hi; i++) { sum += a[i]; mx = a[i] > mx ? a[i] : mx; }
    s_sum[threadIdx.x] = sum; s_max[threadIdx.x] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0, gm = 0;
        for (int t = 0; t < 1024; t++) { const unsigned long long v = s_sum[t]; s_sum[t] = run; run += v; gm = s_max[t] > gm ? s_max[t] : gm; }
        *max_out = gm;
    }
    __syncthreads();
    unsigned long long run = s_sum[threadIdx.x];
    for (int64_t i = lo; i < hi; i++) { const unsigned long long v = a[i]; a[i] = run; run += v; }
}
__global__ void csr_fill_kernel(CsrRel R, int d, const int64_t *ptr, unsigned long long *cursor, uint32_t *meta, int32_t *other, int64_t nnz) {
    for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < R.nobs; o += (int64_t)gridDim.x * blockDim.x) {
        int32_t id[MAX_ARITY];
        for (int p = 0; p < R.arity; p++) id[p] = R.ids[o * R.arity + p];
        for (int p = 0; p < R.arity; p++) {
            if (R.dom[p] != d) continue;
            const int32_t ent = id[p];
            bool dup = false; uint32_t selfmask = 0;
            for (int q = 0; q < R.arity; q++)
                if (R.dom[q] == d && id[q] == ent) { if (q < p) dup = true; selfmask |= 1u << q; }
            if (dup) continue;
            const int64_t slot = ptr[ent] + (int64_t)atomicAdd(&cursor[ent], 1ull);
            meta[slot] = (uint32_t)R.rel | ((uint32_t)R.val[o] << META_VAL_SHIFT) | (selfmask << META_SELF_SHIFT);
            int oi = 0;
            for (int q = 0; q < R.arity; q++) { if (q == p) continue; other[(int64_t)oi * nnz + slot] = id[q]; oi++; }
        }
    }
}

// debug / test hook: the dense kernel's branch-free cell score on arrays of operands (one warp-converged call per element)
__global__ void __launch_bounds__(AUX_THREADS) debug_score_delta_kernel(const uint32_t *N, const uint32_t *sv, const uint32_t *gn, const uint32_t *gs,
                                                                        double *out, int n) {
    __shared__ double lf[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lf[i] = g_logfact[i];
    __syncthreads();
    const int nround = (n + gridDim.x * blockDim.x - 1) / (gridDim.x * blockDim.x);
    for (int r = 0; r < nround; r++) {
        const int i = (r * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
        const bool ok = i < n;
        const double v = score_delta_bf<256>(lf, ok ? N[i] : 0u, ok ? sv[i] : 0u, ok ? gn[i] : 0u, ok ? gs[i] : 0u);
        if (ok) out[i] = v;
    }
}

// live-table bound (highest used slot + 1) of domain 0 of every listed IRM: the dense sweep's cost estimate per chain
__global__ void gather_hi_kernel(const IrmDev *irms, const int32_t *ids, int n, int32_t *out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = *irms[ids[k]].dom[0].hi;
}

// cost estimate of a generic-kernel chain: candidates x groups a move scores, from the live-table bounds of its domains
// (sum over the domains of (tables + 1) x product of the other domains' tables, capped)
__global__ void gather_generic_cost_kernel(const IrmDev *irms, const int32_t *ids, int n, int32_t *out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const IrmDev &irm = irms[ids[k]];
    double cost = 0.0;
    for (int d = 0; d < irm.n_domains; d++) {
        double others = 1.0;
        for (int q = 0; q < irm.n_domains; q++) if (q != d) others *= (double)max(*irm.dom[q].hi, 1);
        cost += (double)(*irm.dom[d].hi + 1) * others * (double)max(irm.dom[d].n, 1);
    }
    out[k] = (int32_t)fmin(cost / 1024.0, 2.0e9);
}

// most live tables (highest used slot + 1) over the domains of every listed IRM: the cluster sizes of a sweep are dealt out by it
__global__ void gather_max_tables_kernel(const IrmDev *irms, const int32_t *ids, int n, int32_t *out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const IrmDev &irm = irms[ids[k]];
    int hi = 1;
    for (int d = 0; d < irm.n_domains; d++) hi = max(hi, *irm.dom[d].hi);
    out[k] = hi;
}

// first kernel error word of every listed IRM in one launch (and reset): one copy back instead of one per IRM
__global__ void gather_errors_kernel(const IrmDev *irms, const int32_t *ids, int n, int32_t *out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) { int32_t *w = irms[ids[k]].error; out[k] = *w; if (*w) *w = 0; }
}

__device__ __forceinline__ uint32_t gcd_u32(uint32_t a, uint32_t b) { while (b) { const uint32_t t = a % b; a = b; b = t; } return a; }

// Sweep order (the loops of cxx/hirm.cc:45-51 / IRM::transition_cluster_assignments_all, cxx/hirm.hh:590-596): for every
// sweep, every domain in index order, every entity once -- in the order j -> (a j + b) mod n with (a, b) drawn per
// (irm, sweep, domain) from Philox, gcd(a, n) = 1 (the reference's order is a hash-container artefact).
// Block k fills moves [off[k], off[k+1]).
__global__ void __launch_bounds__(AUX_THREADS) gen_moves_kernel(const IrmDev *irms, const int32_t *irm_ids, const int64_t *move_off,
                                                                int n_sweeps, uint64_t seed, const uint64_t *stream, const uint64_t *sweep0,
                                                                int32_t *move_dom, int32_t *move_item) {
    const IrmDev &irm = irms[irm_ids[blockIdx.x]];
    int64_t pos = move_off[blockIdx.x];
    __shared__ uint32_t s_a, s_b;
    for (int s = 0; s < n_sweeps; s++) {
        for (int d = 0; d < irm.n_domains; d++) {
            const uint32_t n = (uint32_t)irm.dom[d].n;
            if (n == 0) continue;
            if (threadIdx.x == 0) {
                uint32_t w[4];
                const uint64_t ctr = (sweep0[blockIdx.x] + (uint64_t)s) * MAX_DOMAINS + d, st = stream[blockIdx.x], sd = seed ^ ORDER_SEED_TAG;
                philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)st, (uint32_t)(st >> 32), (uint32_t)sd, (uint32_t)(sd >> 32), w);
                uint32_t a = n > 1 ? 1u + w[0] % (n - 1u) : 1u;
                while (gcd_u32(a, n) != 1u) a = a + 1u >= n ? 1u : a + 1u;
                s_a = a; s_b = w[1] % n;
            }
            __syncthreads();
            const uint64_t a = s_a, b = s_b;
            for (uint32_t j = threadIdx.x; j < n; j += blockDim.x) {
                move_dom[pos + j] = d;
                move_item[pos + j] = (int32_t)((a * j + b) % n);
            }
            pos += n;
            __syncthreads();
        }
    }
}

}  // namespace hirm
