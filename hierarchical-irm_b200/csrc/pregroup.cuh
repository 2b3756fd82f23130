// Pre-grouped rows.  The groups of an entity of domain d (phase A / B1 of the generic sweep kernel: its CSR row accumulated
// by the tables of the OTHER positions, get_cluster_to_items_list, cxx/hirm.hh:388-397) depend on the partition of d itself
// only where a relation holds d twice.  Everywhere else they are a function of the other domains' partitions, and those
// stand still during a run of moves of d (the reference sweeps domain by domain, cxx/hirm.hh:590-596).  So for such a run
// the rows of ALL its moves can be grouped ahead of the sequential pass, by a kernel that is throughput-bound instead of
// latency-bound: one CTA per chain with the other domains' slot bytes in SHARED memory (200 KB for two domains of 100 000
// entities: the 2000 random gathers of a row never leave the SM), rows streamed from HBM with the next row's entries in
// flight while the current one is compacted.  The sequential kernel then reads a move's group list (a few hundred bytes,
// copied into shared memory a move ahead) instead of gathering: its phases A and B1 disappear.
// Output per move: pg_pitch entries, entry 0 = {groups, 0}, then {cell, observations | ones << 16} by ascending cell --
// exactly the list phase B1 builds, so every later addition happens in the same order and results are bit-identical.
#pragma once
#include "engine_types.cuh"

namespace hirm {

constexpr int PG_THREADS = 1024;
constexpr int PG_ADCAP = 16;       // relation descriptors kept in shared memory
struct PregroupLaunch { int32_t scap, zcap; };   // accumulator cells; bytes of slot arrays one CTA can stage

// Runs of equal domain in every listed IRM's move list: out_n[w] runs (may exceed runcap: too many to record), their first
// moves and domains.  One warp per IRM.
__global__ void move_runs_kernel(const int64_t *move_off, const int32_t *move_dom, const int32_t *blk_list, int n, int runcap,
                                 int32_t *out_n, int64_t *out_start, int32_t *out_dom) {
    const int w = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (w >= n) return;
    const int k = blk_list ? blk_list[w] : w;
    const int64_t lo = move_off[k], hi = move_off[k + 1];
    int nr = 0;
    for (int64_t base = lo; base < hi; base += 32) {
        const int64_t m = base + lane;
        const int dm = m < hi ? move_dom[m] : 0;
        const bool st = m < hi && (m == lo || dm != move_dom[m - 1]);
        const unsigned bal = __ballot_sync(0xffffffffu, st);
        if (st) {
            const int pos = nr + __popc(bal & ((1u << lane) - 1u));
            if (pos < runcap) { out_start[(int64_t)w * runcap + pos] = m; out_dom[(int64_t)w * runcap + pos] = dm; }
        }
        nr += __popc(bal);
    }
    if (lane == 0) out_n[w] = nr;
}

__global__ void __launch_bounds__(PG_THREADS, 1) pregroup_rows_kernel(SweepArgs args, const int32_t *blk_list, PregroupLaunch PL) {
    const int k_list = blk_list ? blk_list[blockIdx.x] : (int)blockIdx.x;
    if (!args.pg_pre || !args.pg_pre[k_list]) return;
    const int64_t lo = args.seg_lo[k_list], hi = args.seg_hi[k_list];
    if (lo >= hi) return;
    const IrmDev &irm = args.irms[args.irm_ids[k_list]];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NT = blockDim.x, NW = NT >> 5;
    const int d = args.move_dom[lo];                       // the piece is a run of one domain (host: build_rounds)
    const DomainDev &dd = irm.dom[d];
    const int ctot = irm.ctot[d];
    const int pitch = args.pg_pitch[k_list];
    uint2 *out = args.pg_data + args.pg_base[k_list];
    extern __shared__ __align__(16) unsigned char pg_dyn[];
    uint32_t *s_acc = reinterpret_cast<uint32_t *>(pg_dyn);                // [scap] observations | ones << 16
    uint8_t *s_z = pg_dyn + (size_t)PL.scap * 4;                           // [zcap] slot bytes of the other domains
    __shared__ ADesc s_ad[PG_ADCAP];
    // per relation (by ordinal): what an entry's other-id columns are gathered from -- offset of the domain's slot bytes in shared
    // memory (-1: not staged, read from the L2), accumulator stride -- so that an entry costs a handful of instructions per column
    struct GatherPlan { int32_t zoff[3], rstride[3], dom[3], n, base, pad; };
    __shared__ GatherPlan s_plan[PG_ADCAP];
    __shared__ int32_t s_zoff[MAX_DOMAINS];
    __shared__ int32_t s_wtot[64];
    for (int k = tid; k < PL.scap; k += NT) s_acc[k] = 0u;
    if (tid == 0) {
        int used = 0;
        for (int q = 0; q < MAX_DOMAINS; q++) s_zoff[q] = -1;
        for (int o = 0; o < dd.n_adesc; o++) {
            const ADesc &A = dd.adesc[o];
            for (int p = 0; p < A.arity; p++) {
                const int q = A.dom[p];
                if (q == d || s_zoff[q] >= 0) continue;
                const int nq = irm.dom[q].n;
                if (used + nq <= PL.zcap) { s_zoff[q] = used; used += (nq + 15) & ~15; }
            }
        }
    }
    const bool ad_shared = dd.n_adesc <= PG_ADCAP;
    if (ad_shared) {
        const int32_t *src = reinterpret_cast<const int32_t *>(dd.adesc);
        int32_t *dst = reinterpret_cast<int32_t *>(s_ad);
        for (int k = tid; k < dd.n_adesc * 16; k += NT) dst[k] = src[k];
    }
    __syncthreads();
    for (int q = 0; q < irm.n_domains; q++) {
        const int zo = s_zoff[q];
        if (zo < 0) continue;
        const uint8_t *src = irm.dom[q].zb;
        const int nq = irm.dom[q].n, n16 = nq >> 4;
        for (int i = tid; i < n16; i += NT) reinterpret_cast<uint4 *>(s_z + zo)[i] = __ldcg(reinterpret_cast<const uint4 *>(src) + i);
        for (int i = (n16 << 4) + tid; i < nq; i += NT) s_z[zo + i] = __ldcg(src + i);
    }
    __syncthreads();
    const ADesc *ad = ad_shared ? s_ad : dd.adesc;
    if (ad_shared && tid < dd.n_adesc) {
        const ADesc &A = s_ad[tid];
        GatherPlan P;
        P.n = 0; P.base = A.base[0]; P.pad = 0;
        for (int j = 0; j < 3; j++) { P.zoff[j] = -1; P.rstride[j] = 0; P.dom[j] = 0; }
        for (int p = 0; p < A.arity; p++) {
            if (p == A.a) continue;                        // (the domain occurs once: every other position has an id column, in order)
            if (P.n < 3) { P.zoff[P.n] = s_zoff[A.dom[p]]; P.rstride[P.n] = A.rstride[p]; P.dom[P.n] = A.dom[p]; }
            P.n++;
        }
        s_plan[tid] = P;
    }
    __syncthreads();
    const int n_other = dd.n_other;
    const int CPT = (ctot + NT - 1) / NT;                  // accumulator cells per thread of the compaction (consecutive: list order = thread order)

    struct Entry { uint32_t meta; int32_t id0, id1, id2; bool ok; };
    auto load_entry = [&](int64_t q, int64_t end) -> Entry {
        Entry t{0u, 0, 0, 0, q < end};
        if (t.ok) {
            t.meta = __ldg(dd.meta + q);
            if (n_other > 0) t.id0 = __ldg(dd.other + q);
            if (n_other > 1) t.id1 = __ldg(dd.other + dd.nnz + q);
            if (n_other > 2) t.id2 = __ldg(dd.other + 2 * dd.nnz + q);
        }
        return t;
    };
    auto cell_of = [&](const Entry &t) -> int {           // accumulator cell of the entry | value << 30 (-1: an endpoint is not in its domain)
        if (ad_shared) {
            const GatherPlan &P = s_plan[t.meta & META_REL_MASK];
            int32_t cell = P.base; bool absent = false;
#pragma unroll
            for (int j = 0; j < 3; j++) {
                if (j >= P.n) break;
                const int id = j == 0 ? t.id0 : j == 1 ? t.id1 : t.id2, zo = P.zoff[j];
                int slot = zo >= 0 ? (int)s_z[zo + id] : (int)__ldcg(irm.dom[P.dom[j]].zb + id);
                if (slot == 0xFF) { absent = true; slot = 0; }
                cell += slot * P.rstride[j];
            }
            if (absent) { raise_error(irm, KERR_ABSENT_ENTITY); return -1; }
            return cell | (int)(((t.meta >> META_VAL_SHIFT) & 1u) << 30);
        }
        const uint32_t val = (t.meta >> META_VAL_SHIFT) & 1u, selfmask = (t.meta >> META_SELF_SHIFT) & 0xFu;
        const ADesc *D = ad + (t.meta & META_REL_MASK);
        int32_t rest_k = 0;
        int oi = 0; bool first = true, absent = false;
#pragma unroll
        for (int p = 0; p < MAX_ARITY; p++) {
            if (p >= D->arity) break;
            if ((selfmask >> p) & 1u) { if (first) first = false; else oi++; continue; }
            const int id = oi == 0 ? t.id0 : oi == 1 ? t.id1 : t.id2; oi++;
            const int q = D->dom[p], zo = s_zoff[q];
            int slot = zo >= 0 ? (int)s_z[zo + id] : (int)__ldcg(irm.dom[q].zb + id);
            if (slot == 0xFF) { absent = true; slot = 0; }
            rest_k += slot * D->rstride[p];
        }
        if (absent) { raise_error(irm, KERR_ABSENT_ENTITY); return -1; }
        return (D->base[0] + rest_k) | (int)(val << 30);
    };
    // one shared-memory atomic per distinct cell of the warp's 32 entries (a CRP partition has a dominant table: most entries of a
    // row fall into a handful of cells, and same-address shared atomics serialise; on config 4's CRP-prior partitions it measures
    // the same as plain atomics, it is the skewed partitions it is for)
    auto add_cells = [&](int key) {
        const int cell = key < 0 ? -1 : (key & 0x3FFFFFFF);
        const unsigned peers = __match_any_sync(0xffffffffu, cell);
        const unsigned ones = __ballot_sync(0xffffffffu, key >= 0 && ((key >> 30) & 1));
        if (cell >= 0 && lane == __ffs(peers) - 1) atomicAdd(&s_acc[cell], (uint32_t)__popc(peers) + ((uint32_t)__popc(peers & ones) << 16));
    };
    auto bounds = [&](int it, int64_t &b, int64_t &e) {
        if (it >= 0 && it < dd.n && dd.ptr) { b = __ldg(dd.ptr + it); e = __ldg(dd.ptr + it + 1); } else { b = 0; e = 0; }
    };
    auto item_of = [&](int64_t m) -> int { return m < hi ? __ldg(args.move_item + m) : -1; };
    // One row: X = this row's registers (entry of this thread, bounds, and the id of the row two ahead), Y = the next row's.
    // The next row's entries, the bounds of the row after it and an id four rows ahead are put in flight before this row is
    // accumulated and compacted; the two register sets swap roles from row to row (the loop below is unrolled by two), so no
    // value is touched before the row that needs it.
    auto row = [&](int64_t m, Entry &X, int64_t &bX, int64_t &eX, int &itX, Entry &Y, const int64_t &bY, const int64_t &eY) {
        Y = load_entry(bY + tid, eY);
        add_cells(X.ok ? cell_of(X) : -1);
        for (int64_t q0 = bX + NT; q0 < eX; q0 += NT) {    // (uniform trip count: the ballots inside need every lane)
            const Entry t = load_entry(q0 + tid, eX);
            add_cells(t.ok ? cell_of(t) : -1);
        }
        bounds(itX, bX, eX);                               // row m + 2
        itX = item_of(m + 4);
        __syncthreads();
        // compaction: thread t owns the CPT consecutive cells from t * CPT (list order = thread order); it reads them once (one
        // 16-byte load where CPT is 4), clears what it read, and after the block scan of the counts writes its groups from registers.
        // Two barriers per row: the cells are clear again before the second one, and the next row's first barrier keeps the scan
        // scratch apart.
        const int c0 = tid * CPT;
        uint32_t v4[4] = {0u, 0u, 0u, 0u};
        int cnt = 0;
        if (CPT == 4) {
            if (c0 < ctot) {                               // (ctot is a multiple of 4 here: host pads the accumulator to 32 cells)
                const uint4 q = *reinterpret_cast<const uint4 *>(s_acc + c0);
                v4[0] = q.x; v4[1] = q.y; v4[2] = q.z; v4[3] = q.w;
                cnt = (q.x != 0u) + (q.y != 0u) + (q.z != 0u) + (q.w != 0u);
                if (cnt) *reinterpret_cast<uint4 *>(s_acc + c0) = make_uint4(0u, 0u, 0u, 0u);
            }
        } else {
            for (int j = 0; j < CPT; j++) cnt += (c0 + j < ctot && s_acc[c0 + j] != 0u) ? 1 : 0;
        }
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const int par = (int)(m & 1);                      // (two scan scratches, alternating: no barrier after the reads)
        if (lane == 31) s_wtot[par * 32 + warp] = incl;
        __syncthreads();
        const int wv = lane < NW ? s_wtot[par * 32 + lane] : 0;
        int wincl = wv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, wincl, o);
            if (lane >= o) wincl += v;
        }
        const int total = __shfl_sync(0xffffffffu, wincl, 31);
        int pos = __shfl_sync(0xffffffffu, wincl - wv, warp) + incl - cnt;
        uint2 *slot = out + (m - lo) * (int64_t)pitch;
        if (CPT == 4) {
            if (cnt) {
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (v4[j]) { if (pos + 1 < pitch) __stcg(slot + 1 + pos, make_uint2((uint32_t)(c0 + j), v4[j])); pos++; }
            }
        } else {
            for (int j = 0; j < CPT; j++) {
                const int cell = c0 + j;
                if (cell >= ctot) break;
                const uint32_t v = s_acc[cell];
                if (v) {
                    if (pos + 1 < pitch) __stcg(slot + 1 + pos, make_uint2((uint32_t)cell, v));
                    pos++;
                    s_acc[cell] = 0u;
                }
            }
        }
        if (tid == 0) {
            __stcg(slot, make_uint2((uint32_t)min(total, pitch - 1), 0u));
            if (total > pitch - 1) raise_error(irm, KERR_GROUP_OVERFLOW);
        }
        if (CPT != 4) __syncthreads();                     // (general path: cells are cleared after the scan)
    };
    int64_t bA, eA, bB, eB;
    bounds(item_of(lo), bA, eA);
    bounds(item_of(lo + 1), bB, eB);
    int itA = item_of(lo + 2), itB = item_of(lo + 3);
    Entry A = load_entry(bA + tid, eA), B{0u, 0, 0, 0, false};
    for (int64_t m = lo; m < hi; m += 2) {
        row(m, A, bA, eA, itA, B, bB, eB);
        if (m + 1 < hi) row(m + 1, B, bB, eB, itB, A, bA, eA);
    }
}

}  // namespace hirm
