// Host side of the C ABI declared in include/hirm_engine.h: state store, CSR build,
// label<->slot maps, kernel selection and launches.  No torch, no CPU compute path for
// the move: every score/sample/apply happens in the CUDA kernels.
#include "../../include/hirm_engine.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "device_math.cuh"
#include "engine_types.cuh"
#include "generic_sweep.cuh"
#include "dense_sweep.cuh"
#include "pregroup.cuh"
#include "aux_kernels.cuh"

using namespace hirm;

static thread_local std::string g_err;
static int fail(const std::string &msg) { g_err = msg; return -1; }

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(_e));    \
    } while (0)

namespace {

template <typename T>
struct DevBuf {
    T *p = nullptr; size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept { if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        return cudaMalloc((void **)&p, std::max<size_t>(count, 1) * sizeof(T));
    }
    cudaError_t ensure(size_t count) { return count <= n && p ? cudaSuccess : alloc(count); }
};

struct RelObs {
    int kind = -1;                     // -1 unset, 0 CSR, 1 dense, 2 column of a unary block
    int ub_block = -1, ub_col = -1;    // kind 2
    int64_t nobs = 0;
    const int32_t *ids = nullptr;      // device COO [nobs * arity]: own_ids, or the caller's buffer (set_observations_device)
    const uint8_t *val = nullptr;
    DevBuf<int32_t> own_ids; DevBuf<uint8_t> own_val;
    const uint8_t *slab[2] = {nullptr, nullptr};
    int64_t pitch[2] = {0, 0};
    DevBuf<uint8_t> own_slab0, own_slab1;   // set_observations_dense_host
    int full_obs = -1;                      // dense: 1 = no missing cell, 0 = some, -1 = not scanned yet
};
struct DomObs {
    DevBuf<int64_t> ptr; DevBuf<uint32_t> meta; DevBuf<int32_t> other;
    int64_t nnz = 0; int n_other = 0; int64_t max_deg = 0;
    int ub_block = -1, n_unary = 0;    // unary-block relations of the domain (all from one block)
    std::vector<int> touching;         // CSR relations touching the domain; a CSR entry names its relation by its position here
};
// Unary relations over one domain as entity-major bit rows: relation-owned, built once, shared by every IRM.
struct UnaryBlock {
    int32_t n_ent = 0; int n_rel = 0, words = 0;
    DevBuf<uint32_t> obs, val;
    std::vector<const int32_t *> ids; std::vector<const uint8_t *> vals; std::vector<int64_t> nobs;   // the columns' COO arrays (caller-owned)
};
// Observation store: immutable after finalize(); shared by the chains of one dataset.
struct ObsStore {
    std::vector<RelObs> rel;
    std::vector<DomObs> dom;
    bool finalized = false;
};

struct Irm {
    int n_domains = 0, n_relations = 0;
    std::vector<int32_t> n_ent, kcap, arity;
    std::vector<std::array<int32_t, MAX_ARITY>> rdom;
    std::shared_ptr<ObsStore> obs;
    // device state
    std::vector<DevBuf<int32_t>> z, tab_count, tab_label, hi;
    std::vector<DevBuf<uint8_t>> zb; std::vector<bool> zb_dirty;   // byte shadow of z for the generic kernel's gathers; dirty = written by someone else
    std::vector<double> alpha;
    std::vector<DevBuf<cell_t>> counts;
    std::vector<std::array<int64_t, MAX_ARITY>> cstride;
    std::vector<int64_t> ncells;
    std::vector<int> table_mode;       // per relation: 0 auto (hashed iff the index space exceeds the dense budget), 1 dense, 2 hashed
    std::vector<uint32_t> hcap;        // per relation: entries of the hashed count table, 0 = dense
    DevBuf<uint32_t> hused;            // [n_relations] entries taken (device counters)
    bool any_hashed = false;
    // generic kernel: compact group-accumulator spaces per (relation, domain) (engine_types.cuh: RelDom)
    DevBuf<RelDom> rdesc; std::vector<DevBuf<int32_t>> cmap; std::vector<int32_t> ctot;
    std::vector<DevBuf<unsigned char>> ctmpl; bool tmpl_dirty = true;   // per domain: GroupRec template per accumulator cell
    DevBuf<cell_t> gacc; DevBuf<uint32_t> gbm;
    DevBuf<int64_t> gl_cell; int64_t gl_cap = 0;      // generic kernel: group records beyond the shared-memory ones
    DevBuf<int32_t> error;
    DevBuf<RelationDev> d_rel;
    // per domain: the count tables of this IRM's unary-block relations as ONE matrix [kcap][ucp] (local column j = relation), and
    // the block column of every local column
    std::vector<DevBuf<cell_t>> ucount; std::vector<DevBuf<int32_t>> ucolg; std::vector<int> nu, ucp;
    std::vector<cell_t *> cptr; std::vector<int> estride;   // per relation: its dense cells' address and the stride between them
    std::vector<DevBuf<ADesc>> adesc;   // per domain: phase-A descriptors of its CSR relations, by ordinal
    std::vector<bool> pregroup;         // per domain: every CSR relation holds it once (its rows can be grouped ahead of the moves)
    std::vector<bool> have_z, has_absent;
    bool counts_dirty = true, desc_dirty = true, scratch_ready = false;
    bool alive = true;
    bool dense_ree = false;            // eligible for the dense fast path
    uint64_t rng_stream = 0, move_counter = 0, sweep_counter = 0, alpha_counter = 0;   // Philox positions of sweep_all / transition_alpha
};

}  // namespace

struct hirm_engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::vector<std::unique_ptr<Irm>> irms;
    std::vector<std::unique_ptr<UnaryBlock>> ublocks;
    DevBuf<IrmDev> d_irms;
    std::vector<IrmDev> h_irms;
    // sweep scratch
    DevBuf<int32_t> s_irm_ids, s_move_dom, s_move_item, s_choice, s_trace_n, s_trace_labels, s_blk;
    DevBuf<int64_t> s_move_off; DevBuf<double> s_uniforms, s_trace_logps, s_score;
    DevBuf<uint64_t> s_stream, s_counter;
    DevBuf<long long> s_progress;
    // pre-grouped rows: runs of the move lists, the rounds' pieces, the staged group lists
    DevBuf<int32_t> s_runs_n, s_runs_dom, s_pg_pre, s_pg_pitch; DevBuf<int64_t> s_runs_start, s_seg_lo, s_seg_hi, s_pg_base, s_off_host;
    DevBuf<uint2> s_pg_data;
    DevBuf<int32_t> s_cost;
    int last_rounds = 1, last_pregrouped = 0, pg_smem_set = 0;
    DevBuf<double> s_alpha;
    DevBuf<cell_t> s_reltab; DevBuf<RelCand> s_relcand; DevBuf<int32_t> s_relerr;
    DevBuf<uint64_t> s_sweep0; DevBuf<int32_t> s_errs;
    DevBuf<unsigned char> s_job[6]; DevBuf<int32_t> s_prior;   // relation-score jobs: descriptors, fresh partitions
    // unary relations as bit vectors over the entity codes (observed, value), built once per COO buffer
    struct UnaryBits { DevBuf<uint32_t> obs, val; int64_t nobs = 0; int32_t n_ent = 0; const uint8_t *val_src = nullptr; };
    std::map<std::pair<const int32_t *, const uint8_t *>, UnaryBits> unary_cache;   // keyed by the (ids, val) buffers: relations may share an ids buffer
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int last_cluster = 1;              // largest number of CTAs per IRM of the last generic launches
    cudaStream_t side[4] = {nullptr, nullptr, nullptr, nullptr}; cudaEvent_t ev_fork = nullptr, ev_join[4] = {nullptr, nullptr, nullptr, nullptr};

    int last_launches = 0, last_kind = -1, last_occupancy = 0, last_stages = 0, last_chunk = 0, last_smem = 0; float last_ms = 0.f;
    int max_smem_optin = 0, smem_per_sm = 0, num_sms = 0, dense_static_smem = 0, dense_static_smem_large = 0, gen_regs = 64, gen_static_smem = 12 * 1024;
    long long *d_phase_cycles = nullptr;   // debug hook (hirm_engine_debug_phase_cycles)
};

// function attributes of generic_sweep_kernel are per process: remember what was set
static int g_gen_smem_set = 0;
static bool g_gen_nonportable = false;

static const char *kerr_text(int code) {
    switch (code) {
        case KERR_NO_FREE_SLOT: return "max_tables exhausted: no free table slot for the fresh-table candidate";
        case KERR_ABSENT_ENTITY: return "an observation names an entity whose assignment is -1";
        case KERR_GROUP_OVERFLOW: return "group list capacity exceeded";
        case KERR_BAD_MOVE: return "move names an entity that is not in the domain";
        case KERR_TABLE_FULL: return "hashed count table full";
        default: return "unknown kernel error";
    }
}

extern "C" const char *hirm_last_error(void) { return g_err.c_str(); }
extern "C" int hirm_engine_abi_version(void) { return 1; }

extern "C" int hirm_engine_create(hirm_engine **out, int device) {
    if (!out) return fail("out is null");
    int count = 0;
    cudaError_t ce = cudaGetDeviceCount(&count);
    if (ce != cudaSuccess || count == 0)
        return fail(std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(ce));
    if (device < 0) CK(cudaGetDevice(&device));
    if (device >= count) return fail("device ordinal out of range");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 9) return fail("this engine is built for sm_100a (Blackwell) only");
    auto e = std::make_unique<hirm_engine>();
    e->device = device;
    e->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    e->num_sms = prop.multiProcessorCount;
    e->smem_per_sm = (int)prop.sharedMemPerMultiprocessor;
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, dense_ree_sweep_kernel<DN_DW_SMALL, 5>));
    e->dense_static_smem = (int)fa.sharedSizeBytes + 64;
    CK(cudaFuncGetAttributes(&fa, dense_ree_sweep_kernel<DN_DW_LARGE, 3>));
    e->dense_static_smem_large = (int)fa.sharedSizeBytes + 64;
    CK(cudaFuncGetAttributes(&fa, generic_sweep_kernel));
    e->gen_regs = std::max(fa.numRegs, 16); e->gen_static_smem = (int)fa.sharedSizeBytes + 64;
    CK(cudaEventCreate(&e->ev0));
    CK(cudaEventCreate(&e->ev1));
    // log-factorial table from libm's lgamma (the reference's lbeta, cxx/util_math.cc:13-15)
    std::vector<double> lf(LF_SIZE);
    for (int m = 0; m < LF_SIZE; m++) lf[m] = lgamma((double)m + 1.0);
    CK(cudaMemcpyToSymbol(g_logfact, lf.data(), sizeof(double) * LF_SIZE));
    {   // log(m), m < LOGTAB_SIZE (entry 0 unused), for the scores of small groups
        std::vector<double> lt;
        if (true) {
            lt.resize(LOGTAB_SIZE);
            lt[0] = 0.0;
            for (uint32_t m = 1; m < LOGTAB_SIZE; m++) lt[m] = log((double)m);
        }
        static std::map<int, double *> tab_of_device;          // one table per device for the life of the process (engines come and go)
        double *&p = tab_of_device[device];
        if (!p) {
            CK(cudaMalloc((void **)&p, sizeof(double) * LOGTAB_SIZE));
            CK(cudaMemcpy(p, lt.data(), sizeof(double) * LOGTAB_SIZE, cudaMemcpyHostToDevice));
        }
        const double *cp = p;
        CK(cudaMemcpyToSymbol(g_logtab, &cp, sizeof(cp)));
    }
    *out = e.release();
    return 0;
}

extern "C" int hirm_engine_destroy(hirm_engine *e) {
    if (!e) return 0;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    for (auto &st : e->side) if (st) cudaStreamDestroy(st);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    for (auto &j : e->ev_join) if (j) cudaEventDestroy(j);
    delete e;
    return 0;
}

extern "C" int hirm_engine_set_stream(hirm_engine *e, void *s) {
    if (!e) return fail("engine is null");
    e->stream = (cudaStream_t)s;
    return 0;
}
extern "C" int hirm_engine_synchronize(hirm_engine *e) {
    if (!e) return fail("engine is null");
    CK(cudaStreamSynchronize(e->stream));
    return 0;
}

static Irm *get_irm(hirm_engine *e, int32_t id) {
    if (!e || id < 0 || id >= (int)e->irms.size() || !e->irms[id] || !e->irms[id]->alive) return nullptr;
    return e->irms[id].get();
}
#define GET_IRM(var, e, id)            \
    Irm *var = get_irm(e, id);         \
    if (!var) return fail("bad irm id"); \
    CK(cudaSetDevice(e->device))

extern "C" int hirm_irm_create(hirm_engine *e, int n_domains, const int32_t *n_entities, const int32_t *max_tables,
                               int n_relations, const int32_t *arity, const int32_t *rel_domains, int32_t *out_irm) {
    if (!e || !out_irm) return fail("null argument");
    CK(cudaSetDevice(e->device));
    if (n_domains < 1 || n_domains > MAX_DOMAINS) return fail("n_domains must be in [1, HIRM_MAX_DOMAINS]");
    if (n_relations < 0 || n_relations > (int)META_REL_MASK) return fail("too many relations");
    auto irm = std::make_unique<Irm>();
    irm->n_domains = n_domains; irm->n_relations = n_relations;
    for (int d = 0; d < n_domains; d++) {
        if (n_entities[d] < 0) return fail("negative n_entities");
        if (max_tables[d] < 2 || max_tables[d] > 255) return fail("max_tables must be in [2, 255]");
        irm->n_ent.push_back(n_entities[d]); irm->kcap.push_back(max_tables[d]);
    }
    irm->alpha.assign(n_domains, 1.0);                 // CRP::alpha default (cxx/hirm.hh:72)
    irm->have_z.assign(n_domains, false); irm->has_absent.assign(n_domains, false);
    irm->cstride.resize(n_relations);
    irm->ncells.resize(n_relations);
    irm->table_mode.assign(n_relations, 0); irm->hcap.assign(n_relations, 0);
    for (int r = 0; r < n_relations; r++) {
        if (arity[r] < 1 || arity[r] > MAX_ARITY) return fail("arity must be in [1, HIRM_MAX_ARITY]");
        std::array<int32_t, MAX_ARITY> ds{};
        int64_t cs = 1;
        for (int p = 0; p < arity[r]; p++) {
            int d = rel_domains[r * MAX_ARITY + p];
            if (d < 0 || d >= n_domains) return fail("relation names a domain out of range");
            ds[p] = d;
            int occ = 0;
            for (int q = 0; q <= p; q++) occ += rel_domains[r * MAX_ARITY + q] == d;
            if (occ > 2) return fail("a domain may occur at most twice in one relation");
            irm->cstride[r][p] = cs;
            cs *= irm->kcap[d];                        // <= 255^4: no overflow; beyond the dense budget the table is hashed (ensure_scratch)
        }
        irm->ncells[r] = cs;
        irm->arity.push_back(arity[r]); irm->rdom.push_back(ds);
    }
    irm->obs = std::make_shared<ObsStore>();
    irm->obs->rel.resize(n_relations);
    irm->obs->dom.resize(n_domains);
    irm->z.resize(n_domains); irm->tab_count.resize(n_domains); irm->tab_label.resize(n_domains); irm->hi.resize(n_domains);
    irm->zb.resize(n_domains); irm->zb_dirty.assign(n_domains, true);
    for (int d = 0; d < n_domains; d++) {
        CK(irm->z[d].alloc(n_entities[d]));
        CK(irm->zb[d].alloc(n_entities[d]));
        CK(cudaMemset(irm->z[d].p, 0xFF, std::max(1, n_entities[d]) * sizeof(int32_t)));
        CK(irm->tab_count[d].alloc(irm->kcap[d])); CK(cudaMemset(irm->tab_count[d].p, 0, irm->kcap[d] * sizeof(int32_t)));
        CK(irm->tab_label[d].alloc(irm->kcap[d])); CK(cudaMemset(irm->tab_label[d].p, 0, irm->kcap[d] * sizeof(int32_t)));
        CK(irm->hi[d].alloc(1)); CK(cudaMemset(irm->hi[d].p, 0, sizeof(int32_t)));
    }
    irm->counts.resize(n_relations);
    CK(irm->error.alloc(1)); CK(cudaMemset(irm->error.p, 0, sizeof(int32_t)));
    irm->rng_stream = (uint64_t)e->irms.size();
    e->irms.push_back(std::move(irm));
    *out_irm = (int32_t)e->irms.size() - 1;
    return 0;
}

extern "C" int hirm_irm_destroy(hirm_engine *e, int32_t id) {
    GET_IRM(irm, e, id);
    CK(cudaStreamSynchronize(e->stream));
    e->irms[id].reset();
    return 0;
}

extern "C" int hirm_irm_set_observations(hirm_engine *e, int32_t id, int rel, int64_t nobs, const int32_t *h_ids,
                                         const uint8_t *h_val) {
    GET_IRM(irm, e, id);
    if (rel < 0 || rel >= irm->n_relations) return fail("bad relation index");
    if (irm->obs.use_count() > 1 || irm->obs->finalized) return fail("observation store is shared or finalized");
    if (nobs < 0 || (nobs > 0 && (!h_ids || !h_val))) return fail("null observation arrays");
    const int a = irm->arity[rel];
    for (int64_t k = 0; k < nobs; k++) {
        if (h_val[k] > 1) return fail("observation values must be 0 or 1 (BetaBernoulli)");
        for (int p = 0; p < a; p++) {
            int32_t v = h_ids[k * a + p];
            if (v < 0 || v >= irm->n_ent[irm->rdom[rel][p]]) return fail("entity code out of range");
        }
    }
    RelObs &R = irm->obs->rel[rel];
    R.kind = 0; R.nobs = nobs;
    CK(R.own_ids.alloc((size_t)(nobs * a))); CK(R.own_val.alloc((size_t)nobs));
    if (nobs > 0) {
        CK(cudaMemcpy(R.own_ids.p, h_ids, (size_t)(nobs * a) * sizeof(int32_t), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(R.own_val.p, h_val, (size_t)nobs, cudaMemcpyHostToDevice));
    }
    R.ids = R.own_ids.p; R.val = R.own_val.p;
    return 0;
}

static CsrRel csr_rel_of(const Irm *irm, int r) {
    const RelObs &R = irm->obs->rel[r];
    CsrRel c;
    memset(&c, 0, sizeof(c));
    c.ids = R.ids; c.val = R.val; c.nobs = R.nobs; c.arity = irm->arity[r]; c.rel = r;
    for (int p = 0; p < c.arity; p++) { c.dom[p] = irm->rdom[r][p]; c.n_ent[p] = irm->n_ent[irm->rdom[r][p]]; }
    return c;
}
static unsigned coo_blocks(hirm_engine *e, int64_t nobs) {
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>((nobs + 255) / 256, (int64_t)e->num_sms * 8));
}

// Same with the COO arrays already in device memory (caller-owned, never copied nor freed: a relation's observations
// can be shared by every IRM the relation visits in an HIRM, cxx/hirm.hh:859-864, and by the relation-score kernels).
extern "C" int hirm_irm_set_observations_device(hirm_engine *e, int32_t id, int rel, int64_t nobs, const int32_t *d_ids,
                                                const uint8_t *d_val) {
    GET_IRM(irm, e, id);
    if (rel < 0 || rel >= irm->n_relations) return fail("bad relation index");
    if (irm->obs.use_count() > 1 || irm->obs->finalized) return fail("observation store is shared or finalized");
    if (nobs < 0 || (nobs > 0 && (!d_ids || !d_val))) return fail("null observation arrays");
    RelObs &R = irm->obs->rel[rel];
    R.kind = 0; R.nobs = nobs; R.ids = d_ids; R.val = d_val;
    R.own_ids.release(); R.own_val.release();
    if (nobs > 0) {
        CK(cudaMemsetAsync(irm->error.p, 0, sizeof(int32_t), e->stream));
        coo_validate_kernel<<<coo_blocks(e, nobs), 256, 0, e->stream>>>(csr_rel_of(irm, rel), irm->error.p);
        CK(cudaGetLastError());
        int32_t bad = 0;
        CK(cudaMemcpyAsync(&bad, irm->error.p, sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        if (bad) {
            CK(cudaMemsetAsync(irm->error.p, 0, sizeof(int32_t), e->stream));
            R.kind = -1; R.nobs = 0; R.ids = nullptr; R.val = nullptr;
            return fail("observation values must be 0 or 1 and entity codes in range");
        }
    }
    return 0;
}

// ---- unary blocks -----------------------------------------------------------------------------------------------------
// The observations of many unary relations over one domain (animals.unary: 85 features of 50 animals; config 5: 2000
// relations over 200 000 entities) as two entity-major bit matrices, built on the device from the relations' COO arrays.
// Relation::data_r of a unary relation is "entity -> its one observation" (cxx/hirm.hh:251): the row of entity i holds
// that for every relation at once, 2 bits per (entity, relation) instead of a CSR entry, and it does not depend on which
// IRM a relation currently belongs to -- an IRM selects its relations with a column mask.
extern "C" int hirm_unary_block_create(hirm_engine *e, int32_t n_entities, int n_rel, const int64_t *h_nobs,
                                       const int32_t *const *h_d_ids, const uint8_t *const *h_d_val, int32_t *out_block) {
    if (!e || !out_block) return fail("null argument");
    CK(cudaSetDevice(e->device));
    if (n_entities < 0 || n_rel < 1 || !h_nobs || !h_d_ids || !h_d_val) return fail("bad argument");
    auto B = std::make_unique<UnaryBlock>();
    B->n_ent = n_entities; B->n_rel = n_rel; B->words = (n_rel + 31) / 32;
    const size_t nw = (size_t)std::max(n_entities, 1) * B->words;
    CK(B->obs.alloc(nw)); CK(B->val.alloc(nw));
    CK(cudaMemsetAsync(B->obs.p, 0, nw * sizeof(uint32_t), e->stream));
    CK(cudaMemsetAsync(B->val.p, 0, nw * sizeof(uint32_t), e->stream));
    CK(e->s_relerr.ensure(1));
    CK(cudaMemsetAsync(e->s_relerr.p, 0, sizeof(int32_t), e->stream));
    for (int c = 0; c < n_rel; c++) {
        if (h_nobs[c] < 0 || (h_nobs[c] > 0 && (!h_d_ids[c] || !h_d_val[c]))) return fail("null observation arrays");
        B->ids.push_back(h_d_ids[c]); B->vals.push_back(h_d_val[c]); B->nobs.push_back(h_nobs[c]);
        if (h_nobs[c] == 0) continue;
        unary_block_fill_kernel<<<coo_blocks(e, h_nobs[c]), 256, 0, e->stream>>>(h_d_ids[c], h_d_val[c], h_nobs[c], n_entities, c, B->words,
                                                                              B->obs.p, B->val.p, e->s_relerr.p);
        CK(cudaGetLastError());
    }
    int32_t err = 0;
    CK(cudaMemcpyAsync(&err, e->s_relerr.p, sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (err == 1) return fail("observation values must be 0 or 1 and entity codes in range");
    if (err == 2) return fail("a unary relation observes an entity twice (observations must be distinct)");
    e->ublocks.push_back(std::move(B));
    *out_block = (int32_t)e->ublocks.size() - 1;
    return 0;
}

extern "C" int hirm_unary_block_destroy(hirm_engine *e, int32_t block) {
    if (!e || block < 0 || block >= (int)e->ublocks.size() || !e->ublocks[block]) return fail("bad unary block id");
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    e->ublocks[block].reset();
    return 0;
}

// Relation `rel` of the IRM (unary) = column `col` of the block: IRM::incorporate for all of that relation's observations.
extern "C" int hirm_irm_set_observations_unary(hirm_engine *e, int32_t id, int rel, int32_t block, int col) {
    GET_IRM(irm, e, id);
    if (rel < 0 || rel >= irm->n_relations) return fail("bad relation index");
    if (irm->arity[rel] != 1) return fail("a unary block holds unary relations");
    if (irm->obs.use_count() > 1 || irm->obs->finalized) return fail("observation store is shared or finalized");
    if (block < 0 || block >= (int)e->ublocks.size() || !e->ublocks[block]) return fail("bad unary block id");
    const UnaryBlock &B = *e->ublocks[block];
    if (col < 0 || col >= B.n_rel) return fail("bad unary block column");
    if (B.n_ent != irm->n_ent[irm->rdom[rel][0]]) return fail("the block's domain size differs from the relation's");
    RelObs &R = irm->obs->rel[rel];
    R.kind = 2; R.ub_block = block; R.ub_col = col;
    R.nobs = B.nobs[(size_t)col]; R.ids = B.ids[(size_t)col]; R.val = B.vals[(size_t)col];   // count-table builds and relation scores read the COO arrays
    R.own_ids.release(); R.own_val.release();
    return 0;
}

static int check_dense_rel(Irm *irm, int rel) {
    if (rel < 0 || rel >= irm->n_relations) return fail("bad relation index");
    if (irm->arity[rel] != 2) return fail("dense slabs are for binary relations");
    if (irm->obs.use_count() > 1 || irm->obs->finalized) return fail("observation store is shared or finalized");
    for (int p = 0; p < 2; p++)
        if (irm->kcap[irm->rdom[rel][p]] > 64) return fail("dense relations need max_tables <= 64");
    return 0;
}

extern "C" int hirm_irm_set_observations_dense(hirm_engine *e, int32_t id, int rel, const uint8_t *d_slab0, int64_t pitch0,
                                               const uint8_t *d_slab1, int64_t pitch1) {
    GET_IRM(irm, e, id);
    if (check_dense_rel(irm, rel)) return -1;
    const int n0 = irm->n_ent[irm->rdom[rel][0]], n1 = irm->n_ent[irm->rdom[rel][1]];
    if (!d_slab0 || !d_slab1) return fail("null slab");
    if (pitch0 % 16 || pitch1 % 16 || pitch0 < n1 || pitch1 < n0) return fail("pitches must be multiples of 16 and cover a row");
    if (((uintptr_t)d_slab0 | (uintptr_t)d_slab1) & 15) return fail("slabs must be 16-byte aligned");
    RelObs &R = irm->obs->rel[rel];
    R.kind = 1; R.nobs = (int64_t)n0 * n1;
    R.slab[0] = d_slab0; R.slab[1] = d_slab1; R.pitch[0] = pitch0; R.pitch[1] = pitch1;
    return 0;
}

extern "C" int hirm_engine_transpose_u8(hirm_engine *e, const uint8_t *d_in, int64_t rows, int64_t cols, int64_t pitch_in,
                                        uint8_t *d_out, int64_t pitch_out) {
    if (!e) return fail("engine is null");
    CK(cudaSetDevice(e->device));
    if (rows <= 0 || cols <= 0) return 0;
    dim3 grid((unsigned)((cols + 63) / 64), (unsigned)((rows + 63) / 64)), block(64, 8);
    transpose_u8_kernel<<<grid, block, 0, e->stream>>>(d_in, rows, cols, pitch_in, d_out, pitch_out);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int hirm_irm_set_observations_dense_host(hirm_engine *e, int32_t id, int rel, const uint8_t *h_x, int64_t h_pitch) {
    GET_IRM(irm, e, id);
    if (check_dense_rel(irm, rel)) return -1;
    const int64_t n0 = irm->n_ent[irm->rdom[rel][0]], n1 = irm->n_ent[irm->rdom[rel][1]];
    if (!h_x || h_pitch < n1) return fail("bad host matrix");
    RelObs &R = irm->obs->rel[rel];
    const int64_t p0 = (n1 + 15) & ~(int64_t)15, p1 = (n0 + 15) & ~(int64_t)15;
    CK(R.own_slab0.alloc((size_t)(n0 * p0)));
    CK(R.own_slab1.alloc((size_t)(n1 * p1)));
    CK(cudaMemsetAsync(R.own_slab0.p, 0xFF, (size_t)(n0 * p0), e->stream));
    CK(cudaMemcpy2DAsync(R.own_slab0.p, (size_t)p0, h_x, (size_t)h_pitch, (size_t)n1, (size_t)n0, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemsetAsync(R.own_slab1.p, 0xFF, (size_t)(n1 * p1), e->stream));
    if (hirm_engine_transpose_u8(e, R.own_slab0.p, n0, n1, p0, R.own_slab1.p, p1)) return -1;
    R.kind = 1; R.nobs = n0 * n1;
    R.slab[0] = R.own_slab0.p; R.slab[1] = R.own_slab1.p; R.pitch[0] = p0; R.pitch[1] = p1;
    return 0;
}

extern "C" int hirm_irm_share_observations(hirm_engine *e, int32_t dst, int32_t src) {
    GET_IRM(a, e, dst);
    Irm *b = get_irm(e, src);
    if (!b) return fail("bad source irm id");
    if (a->n_domains != b->n_domains || a->n_relations != b->n_relations || a->n_ent != b->n_ent || a->kcap != b->kcap ||
        a->arity != b->arity || a->rdom != b->rdom)
        return fail("schemas differ");
    if (!b->obs->finalized) return fail("finalize the source irm first");
    a->obs = b->obs;
    a->desc_dirty = true; a->counts_dirty = true; a->scratch_ready = false;
    return 0;
}

// Build the per-entity CSR of every domain over all CSR-stored relations (Relation::data_r,
// cxx/hirm.hh:251,274-280) on the device: degree count, scan, fill (aux_kernels.cuh).
static int finalize_store(hirm_engine *e, Irm *irm) {
    ObsStore &S = *irm->obs;
    if (S.finalized) return 0;
    cudaStream_t s0 = e->stream;
    for (int r = 0; r < irm->n_relations; r++)
        if (S.rel[r].kind < 0) { S.rel[r].kind = 0; S.rel[r].nobs = 0; }   // a relation without observations
    for (int d = 0; d < irm->n_domains; d++) {
        const int n = irm->n_ent[d];
        int max_other = 0;
        std::vector<int> touching;
        for (int r = 0; r < irm->n_relations; r++) {
            RelObs &R = S.rel[r];
            if (R.kind != 0 || R.nobs == 0) continue;
            bool touches = false;
            for (int p = 0; p < irm->arity[r]; p++) touches |= irm->rdom[r][p] == d;
            if (!touches) continue;
            touching.push_back(r);
            max_other = std::max(max_other, irm->arity[r] - 1);
        }
        DomObs &D = S.dom[d];
        D.nnz = 0; D.n_other = std::max(1, max_other); D.max_deg = 0;
        D.touching = touching;
        D.ub_block = -1; D.n_unary = 0;
        for (int r = 0; r < irm->n_relations; r++) {
            if (S.rel[r].kind != 2 || irm->rdom[r][0] != d) continue;
            if (D.ub_block >= 0 && D.ub_block != S.rel[r].ub_block) return fail("the unary-block relations of one domain must come from one block");
            D.ub_block = S.rel[r].ub_block; D.n_unary++;
        }
        if (touching.empty() || n == 0) continue;
        static_assert(sizeof(unsigned long long) == sizeof(int64_t), "ptr is scanned in place");
        CK(D.ptr.alloc((size_t)n + 1));
        CK(cudaMemsetAsync(D.ptr.p, 0, ((size_t)n + 1) * sizeof(int64_t), s0));
        DevBuf<unsigned long long> cursor;                 // [n] fill cursors + [1] largest degree
        CK(cursor.alloc((size_t)n + 1));
        CK(cudaMemsetAsync(cursor.p, 0, ((size_t)n + 1) * sizeof(unsigned long long), s0));
        for (int r : touching) {
            csr_count_kernel<<<coo_blocks(e, S.rel[r].nobs), 256, 0, s0>>>(csr_rel_of(irm, r), d, (unsigned long long *)D.ptr.p);
            CK(cudaGetLastError());
        }
        scan_exclusive_kernel<<<1, 1024, 0, s0>>>((unsigned long long *)D.ptr.p, (int64_t)n + 1, cursor.p + n);
        CK(cudaGetLastError());
        int64_t nnz = 0; unsigned long long maxdeg = 0;
        CK(cudaMemcpyAsync(&nnz, D.ptr.p + n, sizeof(int64_t), cudaMemcpyDeviceToHost, s0));
        CK(cudaMemcpyAsync(&maxdeg, cursor.p + n, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s0));
        CK(cudaStreamSynchronize(s0));
        D.nnz = nnz; D.max_deg = (int64_t)maxdeg;
        if (nnz == 0) { D.ptr.release(); continue; }
        CK(D.meta.alloc((size_t)nnz)); CK(D.other.alloc((size_t)nnz * D.n_other));
        CK(cudaMemsetAsync(D.other.p, 0, (size_t)nnz * D.n_other * sizeof(int32_t), s0));
        for (size_t o = 0; o < touching.size(); o++) {
            const int r = touching[o];
            CsrRel cr = csr_rel_of(irm, r);
            cr.rel = (int32_t)o;                           // ordinal among the domain's CSR relations (ADesc index)
            csr_fill_kernel<<<coo_blocks(e, S.rel[r].nobs), 256, 0, s0>>>(cr, d, D.ptr.p, cursor.p, D.meta.p, D.other.p, nnz);
            CK(cudaGetLastError());
        }
        CK(cudaStreamSynchronize(s0));                     // cursor goes out of scope
    }
    for (int r = 0; r < irm->n_relations; r++) {
        RelObs &R = S.rel[r];
        if (R.kind != 1) continue;                     // dense: is any cell missing?  (selects the kernel's counting scheme)
        const int64_t n0 = irm->n_ent[irm->rdom[r][0]], n1 = irm->n_ent[irm->rdom[r][1]];
        DevBuf<int32_t> flag;
        CK(flag.alloc(1)); CK(cudaMemsetAsync(flag.p, 0, sizeof(int32_t), e->stream));
        if (n0 > 0 && n1 > 0) {
            dense_any_missing_kernel<<<(unsigned)std::min<int64_t>(n0, e->num_sms * 8), 256, 0, e->stream>>>(R.slab[0], n0, n1, R.pitch[0], flag.p);
            CK(cudaGetLastError());
        }
        int32_t h = 0;
        CK(cudaMemcpyAsync(&h, flag.p, sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        R.full_obs = h ? 0 : 1;
    }
    S.finalized = true;
    return 0;
}

static const int64_t DENSE_CELL_BUDGET = (int64_t)1 << 27;   // 1 GiB of packed cells per dense table

static uint32_t pow2_ceil(uint64_t v) { uint32_t p = 1; while (p < v && p < (1u << 30)) p <<= 1; return p; }

static int ensure_scratch(hirm_engine *e, Irm *irm) {
    if (irm->scratch_ready) return 0;
    ObsStore &S = *irm->obs;
    bool any_csr = false;
    const bool env_hashed = getenv("HIRM_HASHED_TABLES") != nullptr;   // test switch: every CSR relation on the hashed table
    irm->any_hashed = false;
    CK(irm->hused.ensure((size_t)std::max(irm->n_relations, 1)));
    for (int r = 0; r < irm->n_relations; r++) {
        any_csr |= S.rel[r].kind == 0 || S.rel[r].kind == 2;
        const int mode = irm->table_mode[r];
        bool hashed = S.rel[r].kind == 0 && (mode == 2 || (mode == 0 && (env_hashed || irm->ncells[r] > DENSE_CELL_BUDGET)));
        if (mode == 1 && irm->ncells[r] > DENSE_CELL_BUDGET) return fail("dense count table too large: lower max_tables or use the hashed table");
        if (S.rel[r].kind == 2) { irm->hcap[r] = 0; irm->counts[r].release(); continue; }   // cells live in the domain's unary count matrix (below)
        if (hashed) {
            // Relation::clusters as an open-addressing table: at most one non-empty cell per observation; 4 entries per
            // possible cell (emptied cells linger until the host rebuilds), grown on demand (maintain_hashed)
            const uint64_t want = 4ull * (uint64_t)std::min<int64_t>(std::max<int64_t>(S.rel[r].nobs, 1), irm->ncells[r]);
            if (irm->hcap[r] == 0) {
                irm->hcap[r] = std::max(1024u, std::min(pow2_ceil(want), 1u << 22));
                if (const char *s = getenv("HIRM_HASH_INITIAL_ENTRIES")) irm->hcap[r] = std::max(64u, pow2_ceil((uint64_t)atoll(s)));   // test switch: exercise growth
            }
            CK(irm->counts[r].ensure((size_t)irm->hcap[r] * 2));
            irm->any_hashed = true;
        } else {
            irm->hcap[r] = 0;
            CK(irm->counts[r].ensure((size_t)irm->ncells[r]));
        }
    }
    if (any_csr) {
        // compact accumulator spaces: for every domain, the (relation, self pattern) blocks laid end to end
        const int nd = irm->n_domains;
        std::vector<RelDom> rd((size_t)irm->n_relations * nd);
        irm->cmap.resize((size_t)nd); irm->ctot.assign((size_t)nd, 0);
        int64_t max_ctot = 1;
        for (int d = 0; d < nd; d++) {
            std::vector<int32_t> cm;
            for (int r = 0; r < irm->n_relations; r++) {
                RelDom &D = rd[(size_t)r * nd + d];
                memset(&D, 0, sizeof(D));
                D.a = D.b = -1; D.rest_size = 1;
                if (S.rel[r].kind != 0) continue;          // dense slabs: own kernel; unary-block relations: scored from the count matrix, no groups
                for (int p = 0; p < irm->arity[r]; p++)
                    if (irm->rdom[r][p] == d) { if (D.a < 0) D.a = p; else D.b = p; }
                if (D.a < 0) continue;
                int64_t rest = 1;
                for (int p = 0; p < irm->arity[r]; p++) {
                    if (p == D.a || p == D.b) continue;
                    D.rstride[p] = (int32_t)rest;
                    rest *= irm->kcap[irm->rdom[r][p]];
                }
                const int64_t kd = irm->kcap[d];
                const int64_t size0 = D.b >= 0 ? rest * kd : rest, size1 = D.b >= 0 ? rest * kd : 0, size2 = D.b >= 0 ? rest : 0;
                if ((int64_t)cm.size() + size0 + size1 + size2 > ((int64_t)1 << 27)) return fail("group accumulator too large: lower max_tables");
                D.rest_size = (int32_t)rest;
                D.base[0] = (int32_t)cm.size(); D.base[1] = D.base[0] + (int32_t)size0; D.base[2] = D.base[1] + (int32_t)size1;
                cm.insert(cm.end(), (size_t)(size0 + size1 + size2), r);
            }
            irm->ctot[(size_t)d] = (int32_t)cm.size();
            max_ctot = std::max<int64_t>(max_ctot, (int64_t)cm.size());
            CK(irm->cmap[(size_t)d].ensure(cm.size()));
            if (!cm.empty()) CK(cudaMemcpy(irm->cmap[(size_t)d].p, cm.data(), cm.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        }
        CK(irm->rdesc.ensure(rd.size()));
        if (!rd.empty()) CK(cudaMemcpy(irm->rdesc.p, rd.data(), rd.size() * sizeof(RelDom), cudaMemcpyHostToDevice));
        irm->adesc.resize((size_t)nd); irm->pregroup.assign((size_t)nd, false);
        for (int d = 0; d < nd; d++) {
            const std::vector<int> &tch = S.dom[d].touching;
            std::vector<ADesc> ad(tch.size());
            bool once = true;
            for (size_t o = 0; o < tch.size(); o++) {
                const int r = tch[o];
                const RelDom &D = rd[(size_t)r * nd + d];
                ADesc &A = ad[o];
                memset(&A, 0, sizeof(A));
                A.arity = irm->arity[r];
                for (int p = 0; p < MAX_ARITY; p++) { A.dom[p] = irm->rdom[r][p]; A.rstride[p] = D.rstride[p]; }
                A.a = D.a; A.b = D.b; A.base[0] = D.base[0]; A.base[1] = D.base[1]; A.base[2] = D.base[2]; A.rest_size = D.rest_size;
                if (D.b >= 0) once = false;
            }
            // the groups of this domain's entities depend on the OTHER domains' partitions only: their rows can be grouped ahead
            // of a run of moves of the domain (pregroup_rows_kernel)
            irm->pregroup[(size_t)d] = once && !tch.empty() && irm->ctot[(size_t)d] <= PREGROUP_MAX_CELLS && S.dom[d].max_deg < 65535;
            CK(irm->adesc[(size_t)d].ensure(ad.size()));
            if (!ad.empty()) CK(cudaMemcpy(irm->adesc[(size_t)d].p, ad.data(), ad.size() * sizeof(ADesc), cudaMemcpyHostToDevice));
        }
        irm->ctmpl.resize((size_t)nd);
        for (int d = 0; d < nd; d++) {                       // cell templates where the cell space is small (filled in ensure_counts)
            if (irm->ctot[(size_t)d] > 0 && irm->ctot[(size_t)d] <= (1 << 16)) CK(irm->ctmpl[(size_t)d].ensure((size_t)irm->ctot[(size_t)d] * GROUP_REC_BYTES));
            else irm->ctmpl[(size_t)d].release();
        }
        irm->tmpl_dirty = true;
        CK(irm->gacc.ensure((size_t)max_ctot)); CK(irm->gbm.ensure((size_t)(max_ctot / 32 + 1)));
        CK(cudaMemsetAsync(irm->gacc.p, 0, (size_t)max_ctot * sizeof(cell_t), e->stream));
        CK(cudaMemsetAsync(irm->gbm.p, 0, (size_t)(max_ctot / 32 + 1) * sizeof(uint32_t), e->stream));
        int64_t cap = 1;
        for (int d = 0; d < irm->n_domains; d++) cap = std::max(cap, S.dom[d].max_deg);
        irm->gl_cap = cap;
        CK(irm->gl_cell.ensure((size_t)cap * (sizeof(GroupRec) / sizeof(int64_t))));   // one GroupRec per group
    }
    // unary-block relations: one count matrix per domain, a column per relation
    {
        const int nd = irm->n_domains;
        irm->ucount.resize((size_t)nd); irm->ucolg.resize((size_t)nd); irm->nu.assign((size_t)nd, 0); irm->ucp.assign((size_t)nd, 0);
        irm->cptr.assign((size_t)irm->n_relations, nullptr); irm->estride.assign((size_t)irm->n_relations, 1);
        for (int r = 0; r < irm->n_relations; r++) irm->cptr[(size_t)r] = irm->counts[r].p;
        for (int d = 0; d < nd; d++) {
            if (S.dom[d].ub_block < 0) continue;
            const UnaryBlock *B = e->ublocks[(size_t)S.dom[d].ub_block].get();
            if (!B) return fail("the IRM's unary block was destroyed");
            std::vector<int32_t> colg;
            std::vector<int> rels;
            std::vector<char> seen((size_t)B->n_rel, 0);
            for (int r = 0; r < irm->n_relations; r++) {
                if (S.rel[r].kind != 2 || irm->rdom[r][0] != d) continue;
                if (seen[(size_t)S.rel[r].ub_col]) return fail("two relations of one IRM name the same unary-block column");
                seen[(size_t)S.rel[r].ub_col] = 1;
                colg.push_back(S.rel[r].ub_col); rels.push_back(r);
            }
            const int nu = (int)colg.size(), ucp = (nu + 31) & ~31;
            irm->nu[(size_t)d] = nu; irm->ucp[(size_t)d] = ucp;
            CK(irm->ucount[(size_t)d].ensure((size_t)std::max(ucp, 32) * irm->kcap[(size_t)d]));
            CK(irm->ucolg[(size_t)d].ensure((size_t)std::max(nu, 1)));
            if (nu) CK(cudaMemcpy(irm->ucolg[(size_t)d].p, colg.data(), (size_t)nu * sizeof(int32_t), cudaMemcpyHostToDevice));
            for (int j = 0; j < nu; j++) { irm->cptr[(size_t)rels[(size_t)j]] = irm->ucount[(size_t)d].p + j; irm->estride[(size_t)rels[(size_t)j]] = ucp; }
        }
    }
    irm->dense_ree = irm->n_domains == 1 && irm->n_relations == 1 && S.rel[0].kind == 1 && irm->kcap[0] <= 64;
    irm->scratch_ready = true; irm->desc_dirty = true;
    return 0;
}

extern "C" int hirm_irm_finalize(hirm_engine *e, int32_t id) {
    GET_IRM(irm, e, id);
    if (finalize_store(e, irm)) return -1;
    return ensure_scratch(e, irm);
}

static int upload_desc(hirm_engine *e, int32_t id) {
    Irm *irm = e->irms[id].get();
    if (e->h_irms.size() < e->irms.size()) e->h_irms.resize(e->irms.size());
    if (e->d_irms.n < e->irms.size()) {
        size_t cap = std::max<size_t>(64, e->irms.size() * 2);
        CK(cudaStreamSynchronize(e->stream));
        CK(e->d_irms.alloc(cap));
        for (auto &p : e->irms) if (p) p->desc_dirty = true;
        e->h_irms.resize(std::max(e->h_irms.size(), e->irms.size()));
    }
    if (!irm->desc_dirty) return 0;
    ObsStore &S = *irm->obs;
    std::vector<RelationDev> rd((size_t)irm->n_relations);
    for (int r = 0; r < irm->n_relations; r++) {
        RelationDev &R = rd[r];
        memset(&R, 0, sizeof(R));
        R.arity = irm->arity[r]; R.kind = S.rel[r].kind;
        for (int p = 0; p < MAX_ARITY; p++) { R.dom[p] = irm->rdom[r][p]; R.cstride[p] = irm->cstride[r][p]; }
        R.counts = irm->cptr[(size_t)r]; R.ncells = irm->ncells[r];
        R.estride = (uint32_t)irm->estride[(size_t)r];
        if (R.estride > 1) R.cstride[0] = R.estride;       // a unary-block relation's cell of slot t sits at counts[t * pitch of the count matrix]
        R.hcap = irm->hcap[r]; R.hused = irm->hused.p + r;
        R.slab[0] = S.rel[r].slab[0]; R.slab[1] = S.rel[r].slab[1]; R.pitch[0] = S.rel[r].pitch[0]; R.pitch[1] = S.rel[r].pitch[1];
        R.coo_ids = S.rel[r].ids; R.coo_val = S.rel[r].val; R.nobs = S.rel[r].nobs;
        R.full_obs = S.rel[r].full_obs > 0 ? 1 : 0;
    }
    CK(irm->d_rel.ensure(rd.size()));
    if (!rd.empty()) CK(cudaMemcpyAsync(irm->d_rel.p, rd.data(), rd.size() * sizeof(RelationDev), cudaMemcpyHostToDevice, e->stream));
    IrmDev D;
    memset(&D, 0, sizeof(D));
    D.n_domains = irm->n_domains; D.n_relations = irm->n_relations;
    for (int d = 0; d < irm->n_domains; d++) {
        DomainDev &dd = D.dom[d];
        dd.n = irm->n_ent[d]; dd.kcap = irm->kcap[d];
        dd.z = irm->z[d].p; dd.zb = irm->zb[d].p; dd.tab_count = irm->tab_count[d].p; dd.tab_label = irm->tab_label[d].p; dd.hi = irm->hi[d].p;
        dd.alpha = irm->alpha[d];
        dd.ptr = S.dom[d].nnz ? S.dom[d].ptr.p : nullptr; dd.meta = S.dom[d].meta.p; dd.other = S.dom[d].other.p;
        dd.nnz = S.dom[d].nnz; dd.n_other = S.dom[d].n_other; dd.max_deg = S.dom[d].max_deg;
        if (d < (int)irm->adesc.size()) { dd.adesc = irm->adesc[(size_t)d].p; dd.n_adesc = (int32_t)S.dom[d].touching.size(); }
        if (S.dom[d].ub_block >= 0 && d < (int)irm->ucount.size()) {
            const UnaryBlock *B = e->ublocks[(size_t)S.dom[d].ub_block].get();
            if (!B) return fail("the IRM's unary block was destroyed");
            dd.ub_obs = B->obs.p; dd.ub_val = B->val.p; dd.ub_words = B->words;
            dd.ucolg = irm->ucolg[(size_t)d].p; dd.ucount = irm->ucount[(size_t)d].p; dd.nu = irm->nu[(size_t)d]; dd.ucp = irm->ucp[(size_t)d];
        }
    }
    D.rel = irm->d_rel.p;
    D.rdesc = irm->rdesc.p; D.gacc = irm->gacc.p; D.gbm = irm->gbm.p;
    for (int d = 0; d < irm->n_domains && d < (int)irm->cmap.size(); d++) { D.cmap[d] = irm->cmap[(size_t)d].p; D.ctot[d] = irm->ctot[(size_t)d]; }
    for (int d = 0; d < irm->n_domains && d < (int)irm->ctmpl.size(); d++) D.ctmpl[d] = irm->ctmpl[(size_t)d].p;
    D.gl_cell = irm->gl_cell.p; D.gl_cap = irm->gl_cap;
    D.error = irm->error.p;
    e->h_irms[id] = D;
    // pageable source: make the copy synchronous w.r.t. the host buffers (rd, D are locals)
    CK(cudaMemcpyAsync(e->d_irms.p + id, &e->h_irms[id], sizeof(IrmDev), cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    irm->desc_dirty = false;
    return 0;
}

static int check_kernel_error(hirm_engine *e, Irm *irm, int32_t *code_out = nullptr) {
    int32_t code = 0;
    CK(cudaMemcpyAsync(&code, irm->error.p, sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (code_out) *code_out = code;
    if (code) {
        CK(cudaMemsetAsync(irm->error.p, 0, sizeof(int32_t), e->stream));
        return fail(std::string("kernel error: ") + kerr_text(code));
    }
    return 0;
}

// error words of the IRMs of a sweep (e->s_irm_ids holds their ids): one gather launch, one copy
static int check_kernel_errors(hirm_engine *e, int n_irms) {
    std::vector<int32_t> codes((size_t)n_irms, 0);
    CK(e->s_errs.ensure((size_t)n_irms));
    gather_errors_kernel<<<(n_irms + 255) / 256, 256, 0, e->stream>>>(e->d_irms.p, e->s_irm_ids.p, n_irms, e->s_errs.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(codes.data(), e->s_errs.p, (size_t)n_irms * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    for (int32_t c : codes) if (c) return fail(std::string("kernel error: ") + kerr_text(c));
    return 0;
}

// Hashed count tables after a sweep: emptied cells linger as zero cells.  A table more than half full is rebuilt from
// (observations, z) at its next use (the histogram holds the non-empty cells only); one that is still more than half
// full right after a rebuild doubles.
static int maintain_hashed(hirm_engine *e, Irm *irm, bool after_rebuild) {
    if (!irm->any_hashed) return 0;
    std::vector<uint32_t> used((size_t)irm->n_relations);
    CK(cudaMemcpyAsync(used.data(), irm->hused.p, used.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    for (int r = 0; r < irm->n_relations; r++) {
        if (!irm->hcap[r] || used[(size_t)r] <= irm->hcap[r] / 2) continue;
        if (after_rebuild) {
            if (irm->hcap[r] >= (1u << 30)) return fail("hashed count table cannot grow further");
            irm->hcap[r] *= 2;
            CK(irm->counts[r].alloc((size_t)irm->hcap[r] * 2));
            irm->cptr[(size_t)r] = irm->counts[r].p;
            irm->desc_dirty = true;
        }
        irm->counts_dirty = true;
    }
    return 0;
}

extern "C" int hirm_irm_set_table_mode(hirm_engine *e, int32_t id, int rel, int mode) {
    GET_IRM(irm, e, id);
    if (rel < 0 || rel >= irm->n_relations || mode < 0 || mode > 2) return fail("bad relation index or mode");
    if (irm->scratch_ready) return fail("set the table mode before hirm_irm_finalize");
    irm->table_mode[rel] = mode;
    return 0;
}

// (re)build every count table from (observations, z)
static int ensure_counts(hirm_engine *e, int32_t id) {
    Irm *irm = e->irms[id].get();
    if (finalize_store(e, irm) || ensure_scratch(e, irm)) return -1;
    for (int d = 0; d < irm->n_domains; d++)
        if (!irm->have_z[d]) return fail("set_assignments has not been called for every domain");
    const bool desc_was_dirty = irm->desc_dirty;
    if (upload_desc(e, id)) return -1;
    if (irm->tmpl_dirty || desc_was_dirty) {             // the templates hold descriptor fields (count-table pointers, strides)
        for (int d = 0; d < irm->n_domains && d < (int)irm->ctmpl.size(); d++) {
            if (!irm->ctmpl[(size_t)d].p) continue;
            build_cell_templates_kernel<<<(irm->ctot[(size_t)d] + 255) / 256, 256, 0, e->stream>>>(e->d_irms.p, id, d, (GroupRec *)irm->ctmpl[(size_t)d].p);
            CK(cudaGetLastError());
        }
        irm->tmpl_dirty = false;
    }
    if (!irm->counts_dirty) return 0;
    ObsStore &S = *irm->obs;
    CK(cudaMemsetAsync(irm->hused.p, 0, (size_t)std::max(irm->n_relations, 1) * sizeof(uint32_t), e->stream));
    for (int d = 0; d < irm->n_domains && d < (int)irm->ucount.size(); d++)
        if (irm->ucount[(size_t)d].p) CK(cudaMemsetAsync(irm->ucount[(size_t)d].p, 0, irm->ucount[(size_t)d].n * sizeof(cell_t), e->stream));
    for (int r = 0; r < irm->n_relations; r++) {
        if (S.rel[r].kind != 2) {
            const size_t words = irm->hcap[r] ? (size_t)irm->hcap[r] * 2 : (size_t)irm->ncells[r];
            CK(cudaMemsetAsync(irm->counts[r].p, 0, words * sizeof(cell_t), e->stream));
        }
        if (S.rel[r].nobs == 0) continue;
        if (S.rel[r].kind != 1) {
            int blocks = (int)std::min<int64_t>((S.rel[r].nobs + 255) / 256, 148 * 8);
            rebuild_counts_coo_kernel<<<blocks, 256, 0, e->stream>>>(e->d_irms.p, id, r);
        } else {
            const int n0 = irm->n_ent[irm->rdom[r][0]], n1 = irm->n_ent[irm->rdom[r][1]], k1 = irm->kcap[irm->rdom[r][1]];
            const size_t smem = (size_t)((n1 + 15) & ~15) + (size_t)k1 * RB_THREADS * 4 + (size_t)irm->ncells[r] * sizeof(cell_t);
            if (smem > (size_t)e->max_smem_optin) return fail("dense relation: domain too large for the count-table build");
            const int per_sm = std::max(1, std::min(4, (int)((size_t)e->smem_per_sm / (smem + 1024))));
            int blocks = std::max(1, std::min(n0, e->num_sms * per_sm));
            CK(cudaFuncSetAttribute(rebuild_counts_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            rebuild_counts_dense_kernel<<<blocks, RB_THREADS, smem, e->stream>>>(e->d_irms.p, id, r);
        }
        CK(cudaGetLastError());
    }
    int32_t code = 0;
    if (check_kernel_error(e, irm, &code)) {
        if (code != KERR_TABLE_FULL || !irm->any_hashed) return -1;
        for (int r = 0; r < irm->n_relations; r++) {          // the non-empty cells alone overflow a hashed table: double and rebuild
            if (!irm->hcap[r]) continue;
            if (irm->hcap[r] >= (1u << 30)) return fail("hashed count table cannot grow further");
            irm->hcap[r] *= 2;
            CK(irm->counts[r].alloc((size_t)irm->hcap[r] * 2));
            irm->cptr[(size_t)r] = irm->counts[r].p;
        }
        irm->desc_dirty = true;
        return ensure_counts(e, id);
    }
    irm->counts_dirty = false;
    if (irm->any_hashed) {
        if (maintain_hashed(e, irm, true)) return -1;
        if (irm->counts_dirty) return ensure_counts(e, id);   // a table grew: build it again
    }
    return 0;
}

extern "C" int hirm_irm_set_assignments(hirm_engine *e, int32_t id, int dom, const int32_t *h_labels) {
    GET_IRM(irm, e, id);
    if (dom < 0 || dom >= irm->n_domains || !h_labels) return fail("bad domain index");
    const int n = irm->n_ent[dom], kcap = irm->kcap[dom];
    // slots in ascending label order (deterministic)
    std::map<int32_t, int32_t> label_slot;
    for (int i = 0; i < n; i++) if (h_labels[i] >= 0) label_slot[h_labels[i]] = 0;
    if ((int)label_slot.size() > kcap) return fail("more tables than max_tables");
    int s = 0;
    for (auto &kv : label_slot) kv.second = s++;
    std::vector<int32_t> z((size_t)std::max(n, 1), -1), cnt((size_t)kcap, 0), lab((size_t)kcap, 0);
    for (auto &kv : label_slot) lab[kv.second] = kv.first;
    for (int i = 0; i < n; i++) if (h_labels[i] >= 0) { z[i] = label_slot[h_labels[i]]; cnt[z[i]]++; }
    int32_t hi = (int32_t)label_slot.size();
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpy(irm->z[dom].p, z.data(), (size_t)std::max(n, 1) * sizeof(int32_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(irm->tab_count[dom].p, cnt.data(), kcap * sizeof(int32_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(irm->tab_label[dom].p, lab.data(), kcap * sizeof(int32_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(irm->hi[dom].p, &hi, sizeof(int32_t), cudaMemcpyHostToDevice));
    irm->have_z[dom] = true; irm->counts_dirty = true; irm->zb_dirty[dom] = true;
    irm->has_absent[dom] = false;
    for (int i = 0; i < n; i++) if (h_labels[i] < 0) irm->has_absent[dom] = true;
    return 0;
}

extern "C" int hirm_irm_get_assignments(hirm_engine *e, int32_t id, int dom, int32_t *h_labels) {
    GET_IRM(irm, e, id);
    if (dom < 0 || dom >= irm->n_domains || !h_labels) return fail("bad domain index");
    const int n = irm->n_ent[dom], kcap = irm->kcap[dom];
    std::vector<int32_t> z((size_t)std::max(n, 1)), lab((size_t)kcap);
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpy(z.data(), irm->z[dom].p, (size_t)std::max(n, 1) * sizeof(int32_t), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(lab.data(), irm->tab_label[dom].p, kcap * sizeof(int32_t), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; i++) h_labels[i] = z[i] < 0 ? -1 : lab[z[i]];
    return 0;
}

extern "C" int hirm_irm_set_alpha(hirm_engine *e, int32_t id, int dom, double alpha) {
    GET_IRM(irm, e, id);
    if (dom < 0 || dom >= irm->n_domains) return fail("bad domain index");
    if (!(alpha > 0)) return fail("alpha must be positive");
    if (irm->alpha[dom] != alpha) { irm->alpha[dom] = alpha; irm->desc_dirty = true; }
    return 0;
}

extern "C" int hirm_irm_get_counts(hirm_engine *e, int32_t id, int rel, int32_t *h_cells, int64_t cap_cells, int64_t *n_cells) {
    GET_IRM(irm, e, id);
    if (rel < 0 || rel >= irm->n_relations || !n_cells) return fail("bad relation index");
    if (ensure_counts(e, id)) return -1;
    const int a = irm->arity[rel];
    const uint32_t hcap = irm->hcap[rel];
    std::vector<cell_t> tab(hcap ? (size_t)hcap * 2 : (size_t)irm->ncells[rel]);
    if (irm->estride[(size_t)rel] > 1)
        CK(cudaMemcpy2D(tab.data(), sizeof(cell_t), irm->cptr[(size_t)rel], (size_t)irm->estride[(size_t)rel] * sizeof(cell_t), sizeof(cell_t),
                        tab.size(), cudaMemcpyDeviceToHost));
    else
        CK(cudaMemcpy(tab.data(), irm->counts[rel].p, tab.size() * sizeof(cell_t), cudaMemcpyDeviceToHost));
    std::vector<std::vector<int32_t>> labs((size_t)a);
    for (int p = 0; p < a; p++) {
        const int d = irm->rdom[rel][p];
        labs[p].resize((size_t)irm->kcap[d]);
        CK(cudaMemcpy(labs[p].data(), irm->tab_label[d].p, labs[p].size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    }
    std::vector<std::vector<int32_t>> cells;
    const int64_t n_entries = hcap ? (int64_t)hcap : irm->ncells[rel];
    for (int64_t k = 0; k < n_entries; k++) {
        int64_t rem = k;
        cell_t c = 0;
        if (hcap) { if (tab[(size_t)k * 2] == 0) continue; rem = (int64_t)tab[(size_t)k * 2] - 1; c = tab[(size_t)k * 2 + 1]; }
        else c = tab[(size_t)k];
        if (cell_n(c) == 0) continue;                  // emptied cells are erased in the reference (cxx/hirm.hh:534-537)
        std::vector<int32_t> row((size_t)a + 2);
        for (int p = 0; p < a; p++) { const int kc = irm->kcap[irm->rdom[rel][p]]; row[p] = labs[p][(size_t)(rem % kc)]; rem /= kc; }
        row[a] = (int32_t)cell_n(c); row[a + 1] = (int32_t)cell_s(c);
        cells.push_back(std::move(row));
    }
    std::sort(cells.begin(), cells.end());
    *n_cells = (int64_t)cells.size();
    if (h_cells)
        for (int64_t k = 0; k < (int64_t)cells.size() && k < cap_cells; k++)
            memcpy(h_cells + k * (a + 2), cells[(size_t)k].data(), sizeof(int32_t) * (a + 2));
    return 0;
}

extern "C" int hirm_engine_logp_scores(hirm_engine *e, int n_irms, const int32_t *h_irms, double *h_out) {
    if (!e) return fail("engine is null");
    CK(cudaSetDevice(e->device));
    if (n_irms <= 0) return 0;
    if (!h_irms || !h_out) return fail("null argument");
    for (int k = 0; k < n_irms; k++) {
        if (!get_irm(e, h_irms[k])) return fail("bad irm id");
        if (ensure_counts(e, h_irms[k])) return -1;
    }
    CK(e->s_irm_ids.ensure((size_t)n_irms)); CK(e->s_score.ensure((size_t)n_irms));
    CK(cudaMemcpyAsync(e->s_irm_ids.p, h_irms, n_irms * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    logp_scores_kernel<<<n_irms, AUX_THREADS, 0, e->stream>>>(e->d_irms.p, e->s_irm_ids.p, e->s_score.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h_out, e->s_score.p, n_irms * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return 0;
}

extern "C" int hirm_irm_logp_score(hirm_engine *e, int32_t id, double *out) {
    if (!out) return fail("out is null");
    return hirm_engine_logp_scores(e, 1, &id, out);
}

extern "C" int hirm_engine_transition_alpha(hirm_engine *e, int n_irms, const int32_t *h_irms, int ngrid, uint64_t seed,
                                            const double *h_uniforms, double *h_out_alpha, double *h_trace_lp) {
    if (!e) return fail("engine is null");
    CK(cudaSetDevice(e->device));
    if (n_irms <= 0) return 0;
    if (!h_irms) return fail("null argument");
    if (ngrid < 2 || ngrid > 64) return fail("ngrid must be in [2, 64]");
    std::vector<uint64_t> st((size_t)n_irms), ct((size_t)n_irms);
    for (int k = 0; k < n_irms; k++) {
        Irm *irm = get_irm(e, h_irms[k]);
        if (!irm) return fail("bad irm id");
        for (int j = 0; j < k; j++) if (h_irms[j] == h_irms[k]) return fail("an irm may appear once per call");
        if (ensure_counts(e, h_irms[k])) return -1;
        st[(size_t)k] = irm->rng_stream; ct[(size_t)k] = irm->alpha_counter++;
    }
    const size_t na = (size_t)n_irms * MAX_DOMAINS;
    CK(e->s_irm_ids.ensure((size_t)n_irms)); CK(e->s_stream.ensure((size_t)n_irms)); CK(e->s_counter.ensure((size_t)n_irms));
    CK(e->s_alpha.ensure(na));
    cudaStream_t s0 = e->stream;
    CK(cudaMemcpyAsync(e->s_irm_ids.p, h_irms, n_irms * sizeof(int32_t), cudaMemcpyHostToDevice, s0));
    CK(cudaMemcpyAsync(e->s_stream.p, st.data(), n_irms * sizeof(uint64_t), cudaMemcpyHostToDevice, s0));
    CK(cudaMemcpyAsync(e->s_counter.p, ct.data(), n_irms * sizeof(uint64_t), cudaMemcpyHostToDevice, s0));
    if (h_uniforms) {
        CK(e->s_uniforms.ensure(na));
        CK(cudaMemcpyAsync(e->s_uniforms.p, h_uniforms, na * sizeof(double), cudaMemcpyHostToDevice, s0));
    }
    if (h_trace_lp) CK(e->s_trace_logps.ensure(na * ngrid));
    transition_alpha_kernel<<<n_irms, 64, 0, s0>>>(e->d_irms.p, e->s_irm_ids.p, ngrid, seed, e->s_stream.p, e->s_counter.p,
                                                  h_uniforms ? e->s_uniforms.p : nullptr, e->s_alpha.p, h_trace_lp ? e->s_trace_logps.p : nullptr);
    CK(cudaGetLastError());
    std::vector<double> al(na);
    CK(cudaMemcpyAsync(al.data(), e->s_alpha.p, na * sizeof(double), cudaMemcpyDeviceToHost, s0));
    if (h_trace_lp) CK(cudaMemcpyAsync(h_trace_lp, e->s_trace_logps.p, na * ngrid * sizeof(double), cudaMemcpyDeviceToHost, s0));
    CK(cudaStreamSynchronize(s0));
    for (int k = 0; k < n_irms; k++) {                 // host mirror of CRP::alpha (the kernel wrote the device descriptors)
        Irm *irm = e->irms[h_irms[k]].get();
        for (int d = 0; d < irm->n_domains; d++) { irm->alpha[d] = al[(size_t)k * MAX_DOMAINS + d]; e->h_irms[h_irms[k]].dom[d].alpha = irm->alpha[d]; }
    }
    if (h_out_alpha) memcpy(h_out_alpha, al.data(), na * sizeof(double));
    return 0;
}

extern "C" int hirm_irm_get_alpha(hirm_engine *e, int32_t id, int dom, double *out) {
    GET_IRM(irm, e, id);
    if (dom < 0 || dom >= irm->n_domains || !out) return fail("bad domain index");
    *out = irm->alpha[dom];
    return 0;
}

extern "C" int hirm_irm_get_tables(hirm_engine *e, int32_t id, int dom, int32_t *h_labels, int32_t *h_counts, int cap, int32_t *n_tables) {
    GET_IRM(irm, e, id);
    if (dom < 0 || dom >= irm->n_domains || !n_tables) return fail("bad domain index");
    const int kcap = irm->kcap[dom];
    std::vector<int32_t> cnt((size_t)kcap), lab((size_t)kcap);
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpy(cnt.data(), irm->tab_count[dom].p, kcap * sizeof(int32_t), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(lab.data(), irm->tab_label[dom].p, kcap * sizeof(int32_t), cudaMemcpyDeviceToHost));
    std::vector<std::pair<int32_t, int32_t>> t;
    for (int k = 0; k < kcap; k++) if (cnt[(size_t)k] > 0) t.push_back({lab[(size_t)k], cnt[(size_t)k]});
    std::sort(t.begin(), t.end());
    *n_tables = (int32_t)t.size();
    for (int k = 0; k < (int)t.size() && k < cap; k++) { if (h_labels) h_labels[k] = t[(size_t)k].first; if (h_counts) h_counts[k] = t[(size_t)k].second; }
    return 0;
}

extern "C" int hirm_engine_relation_scores(hirm_engine *e, int32_t src, int rel, int n_cand, const int32_t *h_cand_irms,
                                           const int32_t *h_cand_doms, double *h_out) {
    GET_IRM(S, e, src);
    if (rel < 0 || rel >= S->n_relations || n_cand < 0 || (n_cand > 0 && (!h_cand_irms || !h_cand_doms || !h_out))) return fail("bad argument");
    if (n_cand == 0) return 0;
    if (finalize_store(e, S)) return -1;
    RelObs &R = S->obs->rel[rel];
    if (R.kind == 1) return fail("relation scores under other partitions are built for COO-backed relations only");
    const int a = S->arity[rel];
    std::vector<RelCand> cands((size_t)n_cand);
    int64_t total_cells = 0;
    for (int k = 0; k < n_cand; k++) {
        Irm *C = get_irm(e, h_cand_irms[k]);
        if (!C) return fail("bad candidate irm id");
        RelCand &rc = cands[(size_t)k];
        memset(&rc, 0, sizeof(rc));
        int64_t stride = 1;
        for (int p = 0; p < a; p++) {
            const int d = h_cand_doms[k * MAX_ARITY + p];
            if (d < 0 || d >= C->n_domains) return fail("candidate domain index out of range");
            if (C->n_ent[d] < S->n_ent[S->rdom[rel][p]]) return fail("candidate domain has fewer entity codes than the relation's");
            if (!C->have_z[d]) return fail("candidate irm without assignments");
            rc.z[p] = C->z[d].p; rc.stride[p] = stride; stride *= C->kcap[d];
        }
        rc.ncells = stride; total_cells += stride;
    }
    CK(e->s_reltab.ensure((size_t)total_cells)); CK(e->s_relcand.ensure((size_t)n_cand)); CK(e->s_score.ensure((size_t)n_cand));
    CK(e->s_relerr.ensure(1));
    int64_t off = 0;
    for (auto &rc : cands) { rc.table = e->s_reltab.p + off; off += rc.ncells; }
    cudaStream_t s0 = e->stream;
    CK(cudaMemsetAsync(e->s_reltab.p, 0, (size_t)total_cells * sizeof(cell_t), s0));
    CK(cudaMemsetAsync(e->s_relerr.p, 0, sizeof(int32_t), s0));
    CK(cudaMemcpyAsync(e->s_relcand.p, cands.data(), cands.size() * sizeof(RelCand), cudaMemcpyHostToDevice, s0));
    if (R.nobs > 0) {
        dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>((R.nobs + AUX_THREADS - 1) / AUX_THREADS, e->num_sms * 4)), (unsigned)n_cand);
        rel_hist_kernel<<<grid, AUX_THREADS, 0, s0>>>(R.ids, R.val, R.nobs, a, e->s_relcand.p, e->s_relerr.p);
        CK(cudaGetLastError());
    }
    rel_score_kernel<<<n_cand, AUX_THREADS, 0, s0>>>(e->s_relcand.p, e->s_score.p);
    CK(cudaGetLastError());
    int32_t err = 0;
    CK(cudaMemcpyAsync(h_out, e->s_score.p, n_cand * sizeof(double), cudaMemcpyDeviceToHost, s0));
    CK(cudaMemcpyAsync(&err, e->s_relerr.p, sizeof(int32_t), cudaMemcpyDeviceToHost, s0));
    CK(cudaStreamSynchronize(s0));
    if (err) return fail(std::string("kernel error: ") + kerr_text(err) + " (seat the relation's entities in the candidate first: hirm_irm_seat_entities)");
    return 0;
}

// ---- batched relation scores (K4, aux_kernels.cuh) ------------------------------------------------------------------
namespace {
struct ScoreJob {
    int sh_cells = RM_SH_CELLS;        // shared-memory histogram cells per work item of this launch
    std::vector<UnaryRel> unary[MAX_DOMAINS];   // matrix mode: unary relations per domain (relmat_unary_kernel)
    int unary_n_ent[MAX_DOMAINS] = {0}, n_cand = 0;
    size_t out_count = 0;              // scores in the output array
    std::vector<RelDesc> rels; std::vector<PartSet> parts; std::vector<ScoreWork> work; std::vector<TableRef> refs;
    std::vector<int64_t> table0;       // scratch offset of relation r's first table
    int64_t total_cells = 0;
};
constexpr int64_t RM_OBS_PER_ITEM = 65536;
}

static int job_relations(ScoreJob &J, int n_rel, const int32_t *h_arity, const int32_t *h_rel_doms, const int64_t *h_nobs,
                         const int32_t *const *h_d_ids, const uint8_t *const *h_d_val, int n_domains, const int32_t *kcap) {
    if (n_rel < 0 || (n_rel > 0 && (!h_arity || !h_rel_doms || !h_nobs || !h_d_ids || !h_d_val))) return fail("null relation arrays");
    J.rels.resize((size_t)n_rel);
    for (int r = 0; r < n_rel; r++) {
        RelDesc &R = J.rels[(size_t)r];
        memset(&R, 0, sizeof(R));
        if (h_arity[r] < 1 || h_arity[r] > MAX_ARITY) return fail("arity must be in [1, HIRM_MAX_ARITY]");
        if (h_nobs[r] < 0 || (h_nobs[r] > 0 && (!h_d_ids[r] || !h_d_val[r]))) return fail("null observation arrays");
        R.ids = h_d_ids[r]; R.val = h_d_val[r]; R.nobs = h_nobs[r]; R.arity = h_arity[r];
        int64_t stride = 1;
        for (int p = 0; p < R.arity; p++) {
            const int d = h_rel_doms[r * MAX_ARITY + p];
            if (d < 0 || d >= n_domains) return fail("relation names a domain out of range");
            R.dom[p] = d; R.stride[p] = stride; stride *= kcap[d];
            if (stride > ((int64_t)1 << 27)) return fail("count table too large: lower max_tables");
        }
        R.ncells = stride;
    }
    return 0;
}

template <typename T>
static int job_upload(hirm_engine *e, DevBuf<unsigned char> &buf, const std::vector<T> &v) {
    CK(buf.ensure(std::max<size_t>(v.size() * sizeof(T), 16)));
    if (!v.empty()) CK(cudaMemcpyAsync(buf.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, e->stream));
    return 0;
}

// histogram + score passes of a prepared job; h_out gets one value per J.refs entry
static int run_score_job(hirm_engine *e, ScoreJob &J, double *h_out) {
    cudaStream_t s0 = e->stream;
    if (J.out_count == 0) J.out_count = J.refs.size();
    CK(e->s_reltab.ensure((size_t)std::max<int64_t>(J.total_cells, 1))); CK(e->s_score.ensure(std::max<size_t>(J.out_count, 1)));
    CK(e->s_relerr.ensure(1));
    CK(cudaMemsetAsync(e->s_reltab.p, 0, (size_t)J.total_cells * sizeof(cell_t), s0));
    CK(cudaMemsetAsync(e->s_relerr.p, 0, sizeof(int32_t), s0));
    if (job_upload(e, e->s_job[0], J.rels) || job_upload(e, e->s_job[1], J.parts) || job_upload(e, e->s_job[2], J.work) ||
        job_upload(e, e->s_job[3], J.refs)) return -1;
    if (!J.work.empty()) {
        const size_t sh_bytes = (size_t)J.sh_cells * 2 * sizeof(uint32_t);
        CK(cudaFuncSetAttribute(relmat_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)RM_SH_CELLS_BIG * 2 * sizeof(uint32_t))));
        relmat_hist_kernel<<<(unsigned)J.work.size(), AUX_THREADS, sh_bytes, s0>>>((const RelDesc *)e->s_job[0].p, (const PartSet *)e->s_job[1].p,
                                                                                  (const ScoreWork *)e->s_job[2].p, e->s_reltab.p, e->s_relerr.p, J.sh_cells);
        CK(cudaGetLastError());
    }
    if (!J.refs.empty()) {
        relmat_score_kernel<<<(unsigned)J.refs.size(), AUX_THREADS, 0, s0>>>((const TableRef *)e->s_job[3].p, e->s_reltab.p, e->s_score.p);
        CK(cudaGetLastError());
    }
    std::vector<UnaryRel> all_unary;                    // keep the upload source alive until the synchronize below
    for (int d = 0; d < MAX_DOMAINS; d++) all_unary.insert(all_unary.end(), J.unary[d].begin(), J.unary[d].end());
    if (!all_unary.empty()) {
        if (job_upload(e, e->s_job[5], all_unary)) return -1;
        const size_t sh = (size_t)32 * (UN_TILE + 1) * sizeof(uint32_t);
        CK(cudaFuncSetAttribute(relmat_unary_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
        size_t off = 0;
        for (int d = 0; d < MAX_DOMAINS; d++) {
            const int nr = (int)J.unary[d].size();
            if (!nr) continue;
            dim3 grid((unsigned)J.n_cand, (unsigned)((nr + UN_WARPS * UN_RW - 1) / (UN_WARPS * UN_RW)));
            relmat_unary_kernel<<<grid, UN_THREADS, sh, s0>>>((const UnaryRel *)e->s_job[5].p + off, nr, (const PartSet *)e->s_job[1].p, d,
                                                              J.unary_n_ent[d], J.n_cand, e->s_score.p, e->s_relerr.p);
            CK(cudaGetLastError());
            off += (size_t)nr;
        }
    }
    if (J.out_count) CK(cudaMemcpyAsync(h_out, e->s_score.p, J.out_count * sizeof(double), cudaMemcpyDefault, s0));   // host or device destination
    int32_t err = 0;
    CK(cudaMemcpyAsync(&err, e->s_relerr.p, sizeof(int32_t), cudaMemcpyDeviceToHost, s0));
    CK(cudaStreamSynchronize(s0));                      // the job vectors are the sources of the async copies
    if (err) return fail(std::string("kernel error: ") + kerr_text(err));
    return 0;
}

// Relation::logp_score of every listed relation under the domain partitions of every listed IRM: the data terms of one
// whole sweep of HIRM::transition_cluster_assignment_relation (cxx/hirm.hh:833-906; the quantity of :865).
extern "C" int hirm_engine_relation_score_matrix(hirm_engine *e, int n_rel, const int32_t *h_arity, const int32_t *h_rel_doms,
                                                 const int64_t *h_nobs, const int32_t *const *h_d_ids, const uint8_t *const *h_d_val,
                                                 int n_cand, const int32_t *h_cand_irms, double *h_out) {
    if (!e) return fail("engine is null");
    CK(cudaSetDevice(e->device));
    if (n_rel <= 0 || n_cand <= 0) return 0;
    if (!h_cand_irms || !h_out) return fail("null argument");
    ScoreJob J;
    Irm *C0 = get_irm(e, h_cand_irms[0]);
    if (!C0) return fail("bad candidate irm id");
    J.parts.resize((size_t)n_cand);
    for (int c = 0; c < n_cand; c++) {
        Irm *C = get_irm(e, h_cand_irms[c]);
        if (!C) return fail("bad candidate irm id");
        if (C->n_domains != C0->n_domains || C->kcap != C0->kcap || C->n_ent != C0->n_ent)
            return fail("the candidates of one score matrix must share domains, entity codes and max_tables");
        memset(&J.parts[(size_t)c], 0, sizeof(PartSet));
        for (int d = 0; d < C->n_domains; d++) {
            if (!C->have_z[d]) return fail("candidate irm without assignments");
            if (C->has_absent[d]) return fail("candidate irm with unseated entities: seat them first (hirm_irm_seat_entities)");
            J.parts[(size_t)c].z[d] = C->z[d].p;
        }
    }
    if (job_relations(J, n_rel, h_arity, h_rel_doms, h_nobs, h_d_ids, h_d_val, C0->n_domains, C0->kcap.data())) return -1;
    for (int r = 0; r < n_rel; r++)                     // a table that does not fit the small histogram but fits the big one: big launch
        if (J.rels[(size_t)r].ncells > RM_SH_CELLS && J.rels[(size_t)r].ncells <= RM_SH_CELLS_BIG) J.sh_cells = RM_SH_CELLS_BIG;
    J.n_cand = n_cand; J.out_count = (size_t)n_rel * n_cand;
    const bool use_unary = getenv("HIRM_NO_UNARY_PATH") == nullptr;
    for (int r = 0; r < n_rel; r++) {
        const RelDesc &R = J.rels[(size_t)r];
        if (use_unary && R.arity == 1 && R.nobs > 0) {   // unary relation: bit vectors, no histogram (relmat_unary_kernel)
            const int d = R.dom[0], n_ent = C0->n_ent[d];
            auto &U = e->unary_cache[std::make_pair(R.ids, R.val)];
            if (U.nobs != R.nobs || U.n_ent != n_ent || U.val_src != R.val) {
                const size_t nw = (size_t)((n_ent + 31) / 32) + UN_TILE;   // zero padding: the kernel reads whole words only
                CK(U.obs.alloc(nw)); CK(U.val.alloc(nw));
                CK(cudaMemsetAsync(U.obs.p, 0, nw * sizeof(uint32_t), e->stream)); CK(cudaMemsetAsync(U.val.p, 0, nw * sizeof(uint32_t), e->stream));
                unary_bits_kernel<<<coo_blocks(e, R.nobs), 256, 0, e->stream>>>(R.ids, R.val, R.nobs, U.obs.p, U.val.p);
                CK(cudaGetLastError());
                U.nobs = R.nobs; U.n_ent = n_ent; U.val_src = R.val;
            }
            J.unary[d].push_back(UnaryRel{U.obs.p, U.val.p, r, 0});
            J.unary_n_ent[d] = n_ent;
            continue;
        }
        const int64_t t0 = J.total_cells;
        J.total_cells += R.ncells * n_cand;
        for (int c = 0; c < n_cand; c++) J.refs.push_back(TableRef{t0 + c * R.ncells, R.ncells, (int64_t)r * n_cand + c});
        // candidates per work item: as many tables as fit the shared histogram; a table too big for it takes global atomics,
        // a few candidates per item so that the grid stays wide
        const int run = R.ncells <= J.sh_cells ? (int)std::max<int64_t>(1, std::min<int64_t>(n_cand, J.sh_cells / R.ncells)) : std::min(n_cand, 8);
        for (int c0 = 0; c0 < n_cand; c0 += run)
            for (int64_t o = 0; o < R.nobs; o += RM_OBS_PER_ITEM)
                J.work.push_back(ScoreWork{r, c0, std::min(run, n_cand - c0), 0, o, std::min(R.nobs, o + RM_OBS_PER_ITEM), t0 + c0 * R.ncells});
    }
    return run_score_job(e, J, h_out);
}

// The unary score path keeps two bit vectors per relation, keyed by the relation's COO buffer: drop them when the caller
// frees or rewrites such buffers while the engine lives.
extern "C" int hirm_engine_clear_relation_cache(hirm_engine *e) {
    if (!e) return fail("engine is null");
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    e->unary_cache.clear();
    return 0;
}

static int launch_prior_draws(hirm_engine *e, const std::vector<PriorDraw> &draws, uint64_t seed) {
    if (draws.empty()) return 0;
    if (job_upload(e, e->s_job[4], draws)) return -1;
    CK(e->s_relerr.ensure(1));
    CK(cudaMemsetAsync(e->s_relerr.p, 0, sizeof(int32_t), e->stream));
    crp_prior_kernel<<<(unsigned)draws.size(), PRIOR_THREADS, 0, e->stream>>>((const PriorDraw *)e->s_job[4].p, seed, e->s_relerr.p);
    CK(cudaGetLastError());
    int32_t err = 0;
    CK(cudaMemcpyAsync(&err, e->s_relerr.p, sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (err) return fail("a CRP prior draw opened more tables than max_tables");
    return 0;
}

// The fresh-table candidate of every listed relation (cxx/hirm.hh:849-864): a new IRM whose domains are seated by
// sequential CRP(alpha) prior draws, one independent draw per (relation, domain) identified by its Philox stream
// h_keys[r * n_domains + d]; h_out[r] = Relation::logp_score of relation r under its own draw.  The draw can be
// reproduced later from the key (hirm_engine_crp_prior_draw) if the relation takes the fresh table.
extern "C" int hirm_engine_relation_scores_fresh(hirm_engine *e, int n_rel, const int32_t *h_arity, const int32_t *h_rel_doms,
                                                 const int64_t *h_nobs, const int32_t *const *h_d_ids, const uint8_t *const *h_d_val,
                                                 int n_domains, const int32_t *h_n_entities, const int32_t *h_max_tables, double alpha,
                                                 uint64_t seed, const uint64_t *h_keys, double *h_out) {
    if (!e) return fail("engine is null");
    CK(cudaSetDevice(e->device));
    if (n_rel <= 0) return 0;
    if (n_domains < 1 || n_domains > MAX_DOMAINS || !h_n_entities || !h_max_tables || !h_keys || !h_out) return fail("bad argument");
    if (!(alpha > 0)) return fail("alpha must be positive");
    ScoreJob J;
    if (job_relations(J, n_rel, h_arity, h_rel_doms, h_nobs, h_d_ids, h_d_val, n_domains, h_max_tables)) return -1;
    for (int r = 0; r < n_rel; r++)
        if (J.rels[(size_t)r].ncells > RM_SH_CELLS && J.rels[(size_t)r].ncells <= RM_SH_CELLS_BIG) J.sh_cells = RM_SH_CELLS_BIG;
    int64_t slots = 0;
    for (int r = 0; r < n_rel; r++) {
        bool used[MAX_DOMAINS] = {false};
        for (int p = 0; p < J.rels[(size_t)r].arity; p++) used[J.rels[(size_t)r].dom[p]] = true;
        for (int d = 0; d < n_domains; d++) if (used[d]) slots += h_n_entities[d];
    }
    CK(e->s_prior.ensure((size_t)std::max<int64_t>(2 * slots, 1)));
    std::vector<PriorDraw> draws;
    J.parts.resize((size_t)n_rel);
    int64_t off = 0;
    for (int r = 0; r < n_rel; r++) {
        const RelDesc &R = J.rels[(size_t)r];
        memset(&J.parts[(size_t)r], 0, sizeof(PartSet));
        bool used[MAX_DOMAINS] = {false};
        for (int p = 0; p < R.arity; p++) used[R.dom[p]] = true;
        for (int d = 0; d < n_domains; d++) {
            if (!used[d]) continue;
            PriorDraw D;
            D.z = e->s_prior.p + off; D.cum = e->s_prior.p + slots + off; D.n = h_n_entities[d]; D.kcap = h_max_tables[d];
            D.key = h_keys[(size_t)r * n_domains + d]; D.alpha = alpha;
            J.parts[(size_t)r].z[d] = D.z;
            draws.push_back(D);
            off += h_n_entities[d];
        }
        const int64_t t0 = J.total_cells;
        J.total_cells += R.ncells;
        J.refs.push_back(TableRef{t0, R.ncells, (int64_t)r});
        for (int64_t o = 0; o < R.nobs; o += RM_OBS_PER_ITEM)
            J.work.push_back(ScoreWork{r, r, 1, 0, o, std::min(R.nobs, o + RM_OBS_PER_ITEM), t0});
    }
    if (launch_prior_draws(e, draws, seed)) return -1;
    return run_score_job(e, J, h_out);
}

// One draw of the sequential CRP(alpha) prior over n customers, table labels in order of appearance (Domain::incorporate
// for every entity of a fresh domain, cxx/hirm.hh:194-203 with CRP::sample :101-113), Philox stream `key`.
extern "C" int hirm_engine_crp_prior_draw(hirm_engine *e, int n, double alpha, int max_tables, uint64_t seed, uint64_t key, int32_t *h_labels) {
    if (!e) return fail("engine is null");
    CK(cudaSetDevice(e->device));
    if (n < 0 || max_tables < 1 || !(alpha > 0) || (n > 0 && !h_labels)) return fail("bad argument");
    if (n == 0) return 0;
    CK(e->s_prior.ensure((size_t)2 * n));
    PriorDraw D;
    D.z = e->s_prior.p; D.cum = e->s_prior.p + n; D.n = n; D.kcap = max_tables; D.key = key; D.alpha = alpha;
    if (launch_prior_draws(e, std::vector<PriorDraw>{D}, seed)) return -1;
    CK(cudaMemcpy(h_labels, e->s_prior.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return 0;
}

// Initial seating of whole domains on the device: what Domain::incorporate does to every entity of a new model, one
// sequential CRP(alpha) prior draw per (irm, domain) (cxx/hirm.hh:194-203 with CRP::sample :101-113) -- the law of the seats
// the reference's `incorporate` loop produces, drawn by crp_prior_kernel from Philox stream (irm's stream, domain) instead
// of one `choice()` per entity.  Replaces a host loop of hirm_irm_set_assignments calls when many chains start.
extern "C" int hirm_engine_seat_prior(hirm_engine *e, int n_irms, const int32_t *h_irms, uint64_t seed) {
    if (!e) return fail("engine is null");
    CK(cudaSetDevice(e->device));
    if (n_irms <= 0) return 0;
    if (!h_irms) return fail("null argument");
    std::vector<PriorDraw> draws;
    std::vector<int32_t> pairs;
    int64_t slots = 0;
    for (int k = 0; k < n_irms; k++) {
        Irm *irm = get_irm(e, h_irms[k]);
        if (!irm) return fail("bad irm id");
        for (int d = 0; d < irm->n_domains; d++) slots += irm->n_ent[d];
    }
    CK(e->s_prior.ensure((size_t)std::max<int64_t>(slots, 1)));
    int64_t off = 0;
    for (int k = 0; k < n_irms; k++) {
        Irm *irm = e->irms[h_irms[k]].get();
        if (finalize_store(e, irm) || ensure_scratch(e, irm)) return -1;
        for (int d = 0; d < irm->n_domains; d++) {
            PriorDraw D;
            D.z = irm->z[d].p; D.cum = e->s_prior.p + off; D.n = irm->n_ent[d]; D.kcap = irm->kcap[d];
            D.key = (irm->rng_stream << 3) | (uint64_t)d; D.alpha = irm->alpha[d];
            off += irm->n_ent[d];
            if (D.n > 0) draws.push_back(D);
            pairs.push_back(h_irms[k]); pairs.push_back(d);
            irm->have_z[d] = true; irm->has_absent[d] = false; irm->zb_dirty[d] = true;
        }
        irm->counts_dirty = true;
        if (upload_desc(e, h_irms[k])) return -1;
    }
    if (launch_prior_draws(e, draws, seed)) return -1;
    CK(e->s_blk.ensure(pairs.size()));
    CK(cudaMemcpyAsync(e->s_blk.p, pairs.data(), pairs.size() * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    tables_from_slots_kernel<<<(unsigned)(pairs.size() / 2), 256, 0, e->stream>>>(e->d_irms.p, e->s_blk.p);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->stream));
    return 0;
}

// Domain::incorporate(item, table) for entities that are not yet in the domain (cxx/hirm.hh:194-203): what
// Relation::incorporate does to the domains of a candidate IRM when HIRM::transition_cluster_assignment_relation
// tries a relation there (:859-864).  The entities have no observation in this IRM, so the count tables stand.
extern "C" int hirm_irm_seat_entities(hirm_engine *e, int32_t id, int dom, int n, const int32_t *h_items, const int32_t *h_labels) {
    GET_IRM(irm, e, id);
    if (dom < 0 || dom >= irm->n_domains || n < 0 || (n > 0 && (!h_items || !h_labels))) return fail("bad argument");
    if (!irm->have_z[dom]) return fail("set_assignments first");
    if (n == 0) return 0;
    const int ne = irm->n_ent[dom], kcap = irm->kcap[dom];
    std::vector<int32_t> z((size_t)std::max(ne, 1)), cnt((size_t)kcap), lab((size_t)kcap);
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpy(z.data(), irm->z[dom].p, (size_t)ne * sizeof(int32_t), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cnt.data(), irm->tab_count[dom].p, kcap * sizeof(int32_t), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(lab.data(), irm->tab_label[dom].p, kcap * sizeof(int32_t), cudaMemcpyDeviceToHost));
    int32_t hi = 0;
    CK(cudaMemcpy(&hi, irm->hi[dom].p, sizeof(int32_t), cudaMemcpyDeviceToHost));
    for (int k = 0; k < n; k++) {
        const int it = h_items[k], lb = h_labels[k];
        if (it < 0 || it >= ne || lb < 0) return fail("entity code or label out of range");
        if (z[(size_t)it] >= 0) return fail("entity is already seated (Domain::incorporate asserts table is None)");
        int slot = -1, free_slot = -1;
        for (int s = 0; s < kcap; s++) {
            if (cnt[(size_t)s] > 0 && lab[(size_t)s] == lb) { slot = s; break; }
            if (cnt[(size_t)s] == 0 && free_slot < 0) free_slot = s;
        }
        if (slot < 0) {
            if (free_slot < 0) return fail("more tables than max_tables");
            slot = free_slot; lab[(size_t)slot] = lb;
        }
        cnt[(size_t)slot]++; z[(size_t)it] = slot; hi = std::max(hi, slot + 1);
    }
    CK(cudaMemcpy(irm->z[dom].p, z.data(), (size_t)ne * sizeof(int32_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(irm->tab_count[dom].p, cnt.data(), kcap * sizeof(int32_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(irm->tab_label[dom].p, lab.data(), kcap * sizeof(int32_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(irm->hi[dom].p, &hi, sizeof(int32_t), cudaMemcpyHostToDevice));
    irm->zb_dirty[dom] = true;
    bool absent = false;
    for (int i = 0; i < ne; i++) if (z[(size_t)i] < 0) absent = true;
    irm->has_absent[dom] = absent;
    return 0;
}

extern "C" int hirm_irm_set_rng(hirm_engine *e, int32_t id, uint64_t stream, uint64_t move_counter) {
    GET_IRM(irm, e, id);
    irm->rng_stream = stream; irm->move_counter = move_counter;
    return 0;
}

extern "C" int hirm_irm_set_rng_counters(hirm_engine *e, int32_t id, uint64_t sweep_counter, uint64_t alpha_counter) {
    GET_IRM(irm, e, id);
    irm->sweep_counter = sweep_counter; irm->alpha_counter = alpha_counter;
    return 0;
}

// ---------------------------------------------------------------------------------------
// sweep
// ---------------------------------------------------------------------------------------
struct Plan {
    std::vector<int32_t> generic_blk, dense_blk;
    std::vector<DenseLaunch> tiers;   // dense: shared-memory tiers, smallest table first; the last one holds all max_tables slots
    int n_ent = 0, kcap = 0;
    int64_t gen_maxdeg = 0;
    int gen_kmax = 0, gen_ndom = 0, gen_nother = 0;   // generic kernel: largest max_tables / domain count / other-id columns of its IRMs
    int gen_ubw = 0;                                  // ... widest unary-block bit row (words)
    int gen_adesc = 0;                                // ... most CSR relations touching one domain
    std::vector<double> gen_weight;                   // per generic IRM: entries (CSR + unary columns) an average move reads
    std::vector<int> gen_tables;                      // ... most live tables of a domain (read back when cluster sizes are dealt out)
    std::vector<std::pair<int, int>> gen_classes;     // (CTAs per IRM, number of IRMs): generic_blk is ordered by class
    int gen_ctot = 0; int64_t gen_groups = 0;         // ... largest accumulator space of a domain, most groups one move can have
    int pg_ctot = 0;                                  // ... largest accumulator space of a domain whose rows can be grouped ahead
    int nopg_ctot = 0;                                // ... of a domain whose rows cannot
};

// Shared-memory plan of one dense tier: as many chains per SM as the launch has (up to 5: the register budget) and fit; the rest of
// each CTA's share of the SM's shared memory goes to its TMA ring.
static bool plan_dense_tier(hirm_engine *e, int n, int kt, int kcap, int full, int n_dense, DenseLaunch &L) {
    const DenseLaunch probe = dense_layout(n, kt, kcap, full, 32, 0);
    const bool small = probe.dw == DN_DW_SMALL;            // kernel instance: registers allow 5 (small) / 3 (large) CTAs per SM
    int chunk = 64 * DN_NA, want = std::min(small ? 5 : 3, std::max(1, (n_dense + e->num_sms - 1) / e->num_sms));
    if (const char *s = getenv("HIRM_DENSE_CHUNK")) chunk = std::max(32, atoi(s) & ~31);
    if (const char *s = getenv("HIRM_DENSE_CTAS_PER_SM")) want = std::max(1, std::min(atoi(s), small ? 5 : 3));
    L = DenseLaunch{};
    for (; want >= 1; want--) {
        const int budget = std::min(e->max_smem_optin, e->smem_per_sm / want - 1024) - (small ? e->dense_static_smem : e->dense_static_smem_large);   // 1 KiB reserved per CTA + the kernel's static shared memory
        const DenseLaunch L0 = dense_layout(n, kt, kcap, full, chunk, 0);
        int stages = (budget - L0.total_bytes) / (2 * L0.chunk);
        stages = std::min(stages, std::min(DN_MAX_STAGES, 2 * L0.nchunks + 2));
        if (stages >= 2 || (want == 1 && stages >= 1)) { L = dense_layout(n, kt, kcap, full, chunk, stages); return true; }
    }
    return false;
}

static int launch_dense_tier(hirm_engine *e, const DenseLaunch &L, unsigned nb, const SweepArgs &a, const int32_t *blk) {
    if (L.dw == DN_DW_SMALL) {
        CK(cudaFuncSetAttribute(dense_ree_sweep_kernel<DN_DW_SMALL, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total_bytes));
        dense_ree_sweep_kernel<DN_DW_SMALL, 5><<<nb, 32 + DN_NA + 32 * DN_DW_SMALL, (size_t)L.total_bytes, e->stream>>>(a, blk, L);
    } else {
        if (L.kt > 32) {                               // the last tier: chains with more than 32 tables (scoring rounds laid out by candidate)
            CK(cudaFuncSetAttribute(dense_ree_sweep_kernel<DN_DW_LARGE, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total_bytes));
            dense_ree_sweep_kernel<DN_DW_LARGE, 3, true><<<nb, 32 + DN_NA + 32 * DN_DW_LARGE, (size_t)L.total_bytes, e->stream>>>(a, blk, L);
        } else {
            CK(cudaFuncSetAttribute(dense_ree_sweep_kernel<DN_DW_LARGE, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total_bytes));
            dense_ree_sweep_kernel<DN_DW_LARGE, 3><<<nb, 32 + DN_NA + 32 * DN_DW_LARGE, (size_t)L.total_bytes, e->stream>>>(a, blk, L);
        }
    }
    CK(cudaGetLastError());
    return 0;
}

static void assign_cluster_classes(hirm_engine *e, Plan &plan, uint32_t flags);

static int plan_sweep(hirm_engine *e, int n_irms, const int32_t *h_irms, uint32_t flags, Plan &plan) {
    bool have_dense = false;
    for (int k = 0; k < n_irms; k++) {
        Irm *irm = get_irm(e, h_irms[k]);
        if (!irm) return fail("bad irm id in sweep list");
        for (int j = 0; j < k; j++) if (h_irms[j] == h_irms[k]) return fail("an irm may appear once per sweep");
        if (irm->any_hashed && !irm->counts_dirty && maintain_hashed(e, irm, false)) return -1;   // a hashed table filling up with emptied cells is rebuilt
        if (ensure_counts(e, h_irms[k])) return -1;
        if (irm->dense_ree && !(flags & HIRM_SWEEP_FORCE_GENERIC)) {
            if (!have_dense) {
                int n_dense = 0, full = 1;
                for (int j = 0; j < n_irms; j++) {
                    Irm *o = get_irm(e, h_irms[j]);
                    if (!o || !o->dense_ree) continue;
                    n_dense++;
                    if (o->obs->rel[0].full_obs <= 0 || o->has_absent[0]) full = 0;   // one launch, one counting scheme
                }
                const int kcap = irm->kcap[0];
                std::vector<int> kts = {8, 32};
                if (const char *s = getenv("HIRM_DENSE_TIER_SLOTS")) kts = {std::max(2, atoi(s))};
                for (int kt : kts) {
                    DenseLaunch L;
                    if (kt < kcap && plan_dense_tier(e, irm->n_ent[0], kt, kcap, full, n_dense, L)) plan.tiers.push_back(L);
                }
                DenseLaunch LB;
                if (!plan_dense_tier(e, irm->n_ent[0], kcap, kcap, full, n_dense, LB))
                    return fail("domain too large for the dense fast path's shared-memory state");
                plan.tiers.push_back(LB);
                plan.n_ent = irm->n_ent[0]; plan.kcap = kcap; have_dense = true;
            } else if (irm->n_ent[0] != plan.n_ent || irm->kcap[0] != plan.kcap) {
                return fail("dense irms of one sweep must share a shape");
            }
            irm->zb_dirty[0] = true;                       // the dense kernel writes z only
            plan.dense_blk.push_back(k);
        } else {
            for (int d = 0; d < irm->n_domains; d++) {     // byte shadows of the slot arrays, where someone else wrote z
                if (!irm->zb_dirty[(size_t)d]) continue;
                if (irm->n_ent[(size_t)d] > 0) {
                    slots_to_bytes_kernel<<<std::max(1, std::min((irm->n_ent[(size_t)d] + 255) / 256, e->num_sms * 8)), 256, 0, e->stream>>>(
                        irm->z[(size_t)d].p, irm->zb[(size_t)d].p, irm->n_ent[(size_t)d]);
                    CK(cudaGetLastError());
                }
                irm->zb_dirty[(size_t)d] = false;
            }
            for (auto &R : irm->obs->rel)
                if (R.kind == 1) return fail("no kernel for this irm: dense slabs are supported for a single R(e,e) relation only");
            for (int32_t kc : irm->kcap) plan.gen_kmax = std::max(plan.gen_kmax, (int)kc);
            plan.gen_ndom = std::max(plan.gen_ndom, irm->n_domains);
            for (int d = 0; d < irm->n_domains; d++) {
                const DomObs &D = irm->obs->dom[(size_t)d];
                plan.gen_nother = std::max(plan.gen_nother, D.n_other); plan.gen_maxdeg = std::max(plan.gen_maxdeg, D.max_deg);
                const int ct = d < (int)irm->ctot.size() ? irm->ctot[(size_t)d] : 0;
                plan.gen_ctot = std::max(plan.gen_ctot, ct);
                plan.gen_adesc = std::max(plan.gen_adesc, (int)D.touching.size());
                if (D.ub_block >= 0 && e->ublocks[(size_t)D.ub_block]) plan.gen_ubw = std::max(plan.gen_ubw, e->ublocks[(size_t)D.ub_block]->words);
                plan.gen_groups = std::max(plan.gen_groups, std::min<int64_t>(D.max_deg, ct));
                if (d < (int)irm->pregroup.size() && irm->pregroup[(size_t)d]) plan.pg_ctot = std::max(plan.pg_ctot, ct);
                else plan.nopg_ctot = std::max(plan.nopg_ctot, ct);
            }
            double w = 0.0, ents = 0.0;
            for (int d = 0; d < irm->n_domains; d++) {
                const DomObs &D = irm->obs->dom[(size_t)d];
                // (a unary-block column is 1/32 of a scoring unit per candidate; a CSR entry is a gather, an atomic and its share of a group:
                // config 5's IRMs swept alone cost 0.3 us per binary relation of 100 entries and next to nothing per unary relation)
                w += (double)D.nnz + (double)D.n_unary * irm->n_ent[(size_t)d] / 16.0; ents += irm->n_ent[(size_t)d];
            }
            plan.gen_weight.push_back(w / std::max(ents, 1.0));
            plan.generic_blk.push_back(k);
        }
    }
    // live tables of the IRMs that may get a cluster: a move's scoring work grows with them as much as with the row length
    if (!plan.generic_blk.empty() && (int)plan.generic_blk.size() <= e->num_sms && !getenv("HIRM_CLUSTER_SIZE") &&
        ((flags & HIRM_SWEEP_CLUSTER) || plan.gen_maxdeg >= 1536 || plan.gen_groups >= 384)) {      // (the conditions under which assign_cluster_classes deals SMs out)
        const int ng = (int)plan.generic_blk.size();
        std::vector<int32_t> ids((size_t)ng), tabs((size_t)ng);
        for (int j = 0; j < ng; j++) ids[(size_t)j] = h_irms[plan.generic_blk[(size_t)j]];
        CK(e->s_cost.ensure((size_t)ng * 2));
        CK(cudaMemcpyAsync(e->s_cost.p, ids.data(), ng * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
        gather_max_tables_kernel<<<(ng + 255) / 256, 256, 0, e->stream>>>(e->d_irms.p, e->s_cost.p, ng, e->s_cost.p + ng);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(tabs.data(), e->s_cost.p + ng, ng * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        plan.gen_tables.assign(tabs.begin(), tabs.end());
    }
    assign_cluster_classes(e, plan, flags);
    // More single-CTA generic chains than the GPU holds at once: the launch (every round of it) ends with its slowest chain, and a
    // chain's cost per move grows with the product of its domains' live tables -- chains drawn from a CRP prior differ by 2x.
    // Heaviest first, the block scheduler fills in the light ones (longest-processing-time-first).  Results do not depend on the order.
    if ((int)plan.generic_blk.size() > 4 * e->num_sms && plan.gen_classes.size() == 1 && plan.gen_classes[0].first == 1) {
        const int ng = (int)plan.generic_blk.size();
        std::vector<int32_t> ids((size_t)ng), cost((size_t)ng);
        for (int j = 0; j < ng; j++) ids[(size_t)j] = h_irms[plan.generic_blk[(size_t)j]];
        for (int j = 0; j < ng; j++) if (upload_desc(e, ids[(size_t)j])) return -1;      // (descriptors on the device before the kernel reads them)
        CK(e->s_cost.ensure((size_t)ng * 2));              // (own scratch: the callers' buffers are already bound to the launch)
        CK(cudaMemcpyAsync(e->s_cost.p, ids.data(), ng * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
        gather_generic_cost_kernel<<<(ng + 255) / 256, 256, 0, e->stream>>>(e->d_irms.p, e->s_cost.p, ng, e->s_cost.p + ng);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(cost.data(), e->s_cost.p + ng, ng * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        std::vector<int> order((size_t)ng);
        for (int j = 0; j < ng; j++) order[(size_t)j] = j;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[(size_t)a] > cost[(size_t)b]; });
        std::vector<int32_t> sorted((size_t)ng); std::vector<double> wt((size_t)ng);
        for (int j = 0; j < ng; j++) { sorted[(size_t)j] = plan.generic_blk[(size_t)order[(size_t)j]]; wt[(size_t)j] = plan.gen_weight[(size_t)order[(size_t)j]]; }
        plan.generic_blk.swap(sorted); plan.gen_weight.swap(wt);
    }
    // More dense chains than fit on the GPU at once: the launch ends with its slowest chain, so the chains with the
    // most live tables (the cost of a move grows with them) go first and the block scheduler fills in the cheap ones
    // (longest-processing-time-first).
    if ((int)plan.dense_blk.size() > 5 * e->num_sms) {
        const int nd = (int)plan.dense_blk.size();
        std::vector<int32_t> ids((size_t)nd), hi((size_t)nd);
        for (int j = 0; j < nd; j++) ids[(size_t)j] = h_irms[plan.dense_blk[(size_t)j]];
        CK(e->s_irm_ids.ensure((size_t)nd)); CK(e->s_choice.ensure((size_t)nd));
        CK(cudaMemcpyAsync(e->s_irm_ids.p, ids.data(), nd * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
        gather_hi_kernel<<<(nd + 255) / 256, 256, 0, e->stream>>>(e->d_irms.p, e->s_irm_ids.p, nd, e->s_choice.p);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(hi.data(), e->s_choice.p, nd * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        std::vector<int> order((size_t)nd);
        for (int j = 0; j < nd; j++) order[(size_t)j] = j;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return hi[(size_t)a] > hi[(size_t)b]; });
        std::vector<int32_t> sorted((size_t)nd);
        for (int j = 0; j < nd; j++) sorted[(size_t)j] = plan.dense_blk[(size_t)order[(size_t)j]];
        plan.dense_blk.swap(sorted);
    }
    return 0;
}

// CTAs per IRM of the generic launches: a thread-block cluster per IRM when the GPU has SMs to spare and a move is heavy
// (long rows or many groups), or when the caller asks for it (HIRM_SWEEP_CLUSTER).  The launch ends with its slowest IRM,
// and the IRMs of an HIRM differ widely in size (a CRP over relations makes a few big clusters and many small ones), so the
// SMs are dealt out by weight: every IRM starts with one CTA, and the IRM with the most work per CTA doubles its cluster
// (up to 16) while SMs are left.  IRMs of equal cluster size go out in one launch, the launches run concurrently.
// Results do not depend on any of this.
static void assign_cluster_classes(hirm_engine *e, Plan &plan, uint32_t flags) {
    const int n_gen = (int)plan.generic_blk.size();
    plan.gen_classes.clear();
    if (n_gen == 0) return;
    std::vector<int> cs((size_t)n_gen, 1);
    if (const char *s = getenv("HIRM_CLUSTER_SIZE")) {
        int c = atoi(s), p = 1;
        while (p * 2 <= c && p < 16) p *= 2;
        std::fill(cs.begin(), cs.end(), std::max(1, p));
    } else {
        const bool forced = (flags & HIRM_SWEEP_CLUSTER) != 0;
        const bool heavy = plan.gen_maxdeg >= 1536 || plan.gen_groups >= 384;
        if ((forced || heavy) && n_gen <= e->num_sms) {
            // (a cluster lives inside one GPC, 16 to 20 SMs each; keeping the budget at 8 x 16 SMs made no measurable difference)
            const int budget = std::max(n_gen, std::min(e->num_sms, getenv("HIRM_CLUSTER_BUDGET") ? atoi(getenv("HIRM_CLUSTER_BUDGET")) : e->num_sms));
            // by work per CTA = entries x live tables (hirm_plan_clusters, csrc/sweep_rounds.cc: host only, covered on the CPU)
            std::vector<int32_t> tabs((size_t)n_gen, 8), out((size_t)n_gen, 1);
            for (int i = 0; i < n_gen && i < (int)plan.gen_tables.size(); i++) tabs[(size_t)i] = plan.gen_tables[(size_t)i];
            hirm_plan_clusters(n_gen, plan.gen_weight.data(), tabs.data(), budget, forced ? 1 : 0, out.data());
            for (int i = 0; i < n_gen; i++) cs[(size_t)i] = out[(size_t)i];
        }
    }
    std::vector<int> order((size_t)n_gen);
    for (int i = 0; i < n_gen; i++) order[(size_t)i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cs[(size_t)a] > cs[(size_t)b]; });
    std::vector<int32_t> blk((size_t)n_gen); std::vector<double> wt((size_t)n_gen);
    for (int i = 0; i < n_gen; i++) {
        blk[(size_t)i] = plan.generic_blk[(size_t)order[(size_t)i]]; wt[(size_t)i] = plan.gen_weight[(size_t)order[(size_t)i]];
        const int c = cs[(size_t)order[(size_t)i]];
        if (plan.gen_classes.empty() || plan.gen_classes.back().first != c) plan.gen_classes.push_back({c, 0});
        plan.gen_classes.back().second++;
    }
    plan.generic_blk.swap(blk); plan.gen_weight.swap(wt);
}

static int launch_sweep(hirm_engine *e, const Plan &plan, const SweepArgs &args, int round = 0) {
    if (round == 0) {
        e->last_launches = 0; e->last_kind = -1;
        CK(cudaEventRecord(e->ev0, e->stream));
    }
    if (!plan.generic_blk.empty() && args.pg_pre) {
        // group the rows of this round's pieces ahead of the sequential pass: one CTA per IRM, the other domains' slot bytes in
        // shared memory
        PregroupLaunch PL;
        PL.scap = std::max(256, (plan.pg_ctot + 31) & ~31);
        PL.zcap = (e->max_smem_optin - 3072 - PL.scap * 4) & ~15;
        if (PL.zcap < 0) return fail("pre-grouping: the accumulator does not fit the shared memory");
        const int smem = PL.scap * 4 + PL.zcap;
        if (smem > e->pg_smem_set) { CK(cudaFuncSetAttribute(pregroup_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); e->pg_smem_set = smem; }
        pregroup_rows_kernel<<<(unsigned)plan.generic_blk.size(), PG_THREADS, (size_t)smem, e->stream>>>(args, e->s_blk.p, PL);
        CK(cudaGetLastError());
        e->last_launches++;
    }
    if (!plan.generic_blk.empty()) {
        // at most one IRM per SM: 1024 threads each (shortest move: every phase of a move is spread over the whole SM);
        // more: 256 threads, 4 CTAs per SM (the register file's limit at 64 registers) hide one another's latency
        const int n_gen = (int)plan.generic_blk.size();
        // (measured on the scaled config 5, 2048 IRMs: 256 threads 26.4 M moves/s, 128 threads 20.2 M -- profiles/r02_notes.md)
        int gen_threads = n_gen <= e->num_sms ? 1024 : 256;
        if (const char *s = getenv("HIRM_GENERIC_THREADS")) gen_threads = std::max(64, std::min(1024, atoi(s) & ~31));
        e->last_cluster = 1;
        int done = 0, li = 0;
        const int n_classes = (int)plan.gen_classes.size();
        if (n_classes > 1) {                               // the classes run side by side: fork the side streams off the engine's stream
            for (int k = 1; k < n_classes && k <= 4; k++) if (!e->side[k - 1]) CK(cudaStreamCreateWithFlags(&e->side[k - 1], cudaStreamNonBlocking));
            if (!e->ev_fork) { CK(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming)); for (auto &j : e->ev_join) CK(cudaEventCreateWithFlags(&j, cudaEventDisableTiming)); }
            CK(cudaEventRecord(e->ev_fork, e->stream));
        }
        for (const auto &cls : plan.gen_classes) {
            int csize = cls.first;
            const int n_cls = cls.second;
            cudaStream_t st = (li == 0 || li > 4) ? e->stream : e->side[li - 1];
            if (st != e->stream) CK(cudaStreamWaitEvent(st, e->ev_fork, 0));
            GenericLaunch GL;
            GL.csize = csize;
            GL.ubw = 2 * plan.gen_ubw <= gen_threads ? plan.gen_ubw : 0;   // one thread copies one word of the two planes
            GL.adcap = plan.gen_adesc <= 64 ? plan.gen_adesc : 0;
            GL.roww = 1 + std::max(plan.gen_nother, 1);
            // chunks of the next move's row copied into shared memory a move ahead: the whole row of a CTA's share where that
            // costs at most 24 KB (phase A then waits for the slot gathers only, not for the CSR words)
            {
                const int na = gen_threads > 32 ? gen_threads - 32 : gen_threads;
                const int64_t share = (plan.gen_maxdeg + csize - 1) / csize;
                int hk = (int)std::max<int64_t>(1, std::min<int64_t>((share + na - 1) / na, 8));
                while (hk > 1 && (size_t)2 * hk * GL.roww * gen_threads * 4 > 24 * 1024) hk--;
                if (n_gen > e->num_sms) hk = 1;            // several CTAs per SM: the extra shared memory costs a resident CTA, which costs more than it saves (config 4: 12.2 -> 11.7 M moves/s)
                GL.headk = hk;
            }
            GL.pcap = 4 * gen_threads;
            GL.kstride = (std::max(plan.gen_kmax, 2) + 2) & ~1;
            GL.ndom = std::max(plan.gen_ndom, 1);
            // Shared-memory plan: the move's accumulator (4 bytes per cell of the largest per-domain cell space) and its group
            // records (76 bytes per group, at most min(longest row, cells) groups) stay on chip when they fit the CTA's share
            // of the SM; the fewer IRMs per SM, the bigger that share.  What does not fit goes to the IRM's global scratch.
            // (a pre-grouped round needs no accumulator for the domains whose rows are grouped ahead; a short run of such a domain that
            // goes through the ordinary path anyway falls back to the IRM's global accumulator)
            const int scap_want = std::max(256, ((args.pg_pre ? plan.nopg_ctot : plan.gen_ctot) + 31) & ~31);
            const int rcap_want = std::max(128, (int)((std::min<int64_t>(plan.gen_groups, 1 << 20) + 31) & ~31));
            int want_ctas = std::max(1, std::min(std::min(2048 / gen_threads, 65536 / (e->gen_regs * gen_threads)), (n_gen + e->num_sms - 1) / e->num_sms));
            const size_t fixed = generic_smem_bytes(0, GL.pcap, 0, gen_threads, GL.kstride, GL.ndom, GL.roww, 1, GL.ubw, GL.adcap, GL.headk) + (size_t)e->gen_static_smem;
            const int accs = csize > 1 ? 2 : 1;              // cluster: a local and a merged accumulator per CTA
            // the accumulator decides how many CTAs share an SM (a global accumulator costs far more than a lost CTA); the group
            // records take what is left of the CTA's share, the rest of a long group list lives in the IRM's global scratch
            for (;; want_ctas--) {
                const size_t budget = (size_t)std::min(e->max_smem_optin, e->smem_per_sm / want_ctas - 1024);
                const size_t room = budget > fixed ? budget - fixed : 0;
                if ((size_t)scap_want * 4 * accs + 128 * (GROUP_REC_BYTES + 4) <= room || want_ctas == 1) {
                    GL.scap = (int)std::min<size_t>((size_t)scap_want, room > 128 * (GROUP_REC_BYTES + 4) ? ((room - 128 * (GROUP_REC_BYTES + 4)) / (4 * accs)) & ~(size_t)31 : 0);
                    if (GL.scap < 256) GL.scap = 256;
                    const size_t left = room > (size_t)GL.scap * 4 * accs ? room - (size_t)GL.scap * 4 * accs : 0;
                    GL.rcap = (int)std::min<size_t>((size_t)rcap_want, (left / (GROUP_REC_BYTES + 4)) & ~(size_t)31);
                    if (GL.rcap < 128) GL.rcap = 128;
                    break;
                }
            }
            const size_t gen_smem = generic_smem_bytes(GL.rcap, GL.pcap, GL.scap, gen_threads, GL.kstride, GL.ndom, GL.roww, csize, GL.ubw, GL.adcap, GL.headk);
            if (gen_smem > (size_t)e->max_smem_optin) return fail("generic kernel: shared-memory plan exceeds the device limit");
            if ((int)gen_smem > g_gen_smem_set) { CK(cudaFuncSetAttribute(generic_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gen_smem)); g_gen_smem_set = (int)gen_smem; }
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.blockDim = dim3((unsigned)gen_threads); cfg.dynamicSmemBytes = gen_smem; cfg.stream = st;
            cudaLaunchAttribute attr[1];
            const int32_t *blk = e->s_blk.p + done;
            for (;; csize /= 2) {
                GL.csize = csize;
                cfg.gridDim = dim3((unsigned)n_cls * (unsigned)csize);
                cfg.numAttrs = 0; cfg.attrs = nullptr;
                if (csize > 1) {
                    if (csize > 8 && !g_gen_nonportable) { CK(cudaFuncSetAttribute(generic_sweep_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)); g_gen_nonportable = true; }
                    attr[0].id = cudaLaunchAttributeClusterDimension;
                    attr[0].val.clusterDim.x = (unsigned)csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
                    cfg.attrs = attr; cfg.numAttrs = 1;
                    int n_clusters = 0;
                    // every cluster of the class must be resident at once (the IRMs run for the whole launch: a cluster that waits
                    // for a free GPC slot doubles the launch time): halve the cluster size until the class fits
                    if (cudaOccupancyMaxActiveClusters(&n_clusters, generic_sweep_kernel, &cfg) != cudaSuccess || n_clusters < n_cls) {
                        cudaGetLastError();
                        if (getenv("HIRM_DEBUG_CLUSTERS")) fprintf(stderr, "[hirm] cluster size %d: %d clusters fit, %d wanted: halving\n", csize, n_clusters, n_cls);
                        continue;
                    }
                }
                break;
            }
            cudaEvent_t dbg0 = nullptr, dbg1 = nullptr;
            const bool dbg = getenv("HIRM_DEBUG_CLUSTERS") != nullptr;
            if (dbg) { cudaEventCreate(&dbg0); cudaEventCreate(&dbg1); cudaEventRecord(dbg0, st); }
            CK(cudaLaunchKernelEx(&cfg, generic_sweep_kernel, args, blk, GL));
            CK(cudaGetLastError());
            if (dbg) {
                cudaEventRecord(dbg1, st); cudaEventSynchronize(dbg1);       // (serialises the classes: debugging only)
                float ms = 0; cudaEventElapsedTime(&ms, dbg0, dbg1);
                double wmax = 0, wsum = 0;
                for (int i = done; i < done + n_cls; i++) { wmax = std::max(wmax, plan.gen_weight[(size_t)i]); wsum += plan.gen_weight[(size_t)i]; }
                fprintf(stderr, "[hirm] class %d: %d IRMs x %d CTAs, weight max %.0f sum %.0f, scap %d rcap %d smem %zu: %.1f ms\n", li, n_cls, csize, wmax, wsum,
                        GL.scap, GL.rcap, gen_smem, ms);
                cudaEventDestroy(dbg0); cudaEventDestroy(dbg1);
            }
            if (st != e->stream) { CK(cudaEventRecord(e->ev_join[li - 1], st)); CK(cudaStreamWaitEvent(e->stream, e->ev_join[li - 1], 0)); }
            e->last_launches++; e->last_kind = 0; e->last_cluster = std::max(e->last_cluster, csize);
            done += n_cls; li++;
        }
    }
    if (!plan.dense_blk.empty() && round == 0) {
        const unsigned nb = (unsigned)plan.dense_blk.size();
        SweepArgs a2 = args;
        if (plan.tiers.size() > 1) {                   // several tiers: per-chain progress hands the stopped chains over
            CK(e->s_progress.ensure(nb));
            CK(cudaMemsetAsync(e->s_progress.p, 0, nb * sizeof(long long), e->stream));
            a2.progress = e->s_progress.p;
        }
        for (const DenseLaunch &L : plan.tiers) {
            if (launch_dense_tier(e, L, nb, a2, e->s_blk.p + plan.generic_blk.size())) return -1;
            e->last_launches++;
        }
        e->last_kind = 1;
        const DenseLaunch &L0 = plan.tiers.front();
        int occ = 0;
        if (L0.dw == DN_DW_SMALL) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dense_ree_sweep_kernel<DN_DW_SMALL, 5>, 32 + DN_NA + 32 * DN_DW_SMALL, (size_t)L0.total_bytes);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dense_ree_sweep_kernel<DN_DW_LARGE, 3>, 32 + DN_NA + 32 * DN_DW_LARGE, (size_t)L0.total_bytes);
        e->last_occupancy = occ; e->last_stages = L0.stages; e->last_chunk = L0.chunk; e->last_smem = L0.total_bytes;
    }
    CK(cudaEventRecord(e->ev1, e->stream));
    return 0;
}

// One sweep call = one launch of every kernel, unless rows can be grouped ahead (pregroup.cuh): then the move lists are cut into
// ROUNDS -- piece r of every IRM's list -- and each round is a pre-grouping launch followed by the sequential launch.  A
// piece is at most `seg` moves of one run of a domain that every CSR relation of its IRM holds once; everything else (short
// runs, twice-held domains, lists with too many runs) goes through the ordinary path, merged into as few pieces as possible.
// h_off: the move offsets on the host, or null (read back from args.move_off).  Results do not depend on any of this.
static int run_sweep(hirm_engine *e, const Plan &plan, const SweepArgs &args, int n_irms, const int32_t *h_irms, const int64_t *h_off) {
    e->last_rounds = 1; e->last_pregrouped = 0;
    const int n_gen = (int)plan.generic_blk.size();
    int mode = -1;                                         // HIRM_PREGROUP: 0 = never, 1 = wherever possible (tests); default: many chains with long rows
    if (const char *s = getenv("HIRM_PREGROUP")) mode = atoi(s) != 0;
    bool want = n_gen > 0 && mode != 0 && plan.pg_ctot > 0 && plan.pg_ctot <= PREGROUP_MAX_CELLS;
    if (want && mode < 0) {
        want = n_gen > e->num_sms && plan.gen_maxdeg >= 128;
        for (const auto &cls : plan.gen_classes) if (cls.first > 1) want = false;
    }
    if (!want) return launch_sweep(e, plan, args);
    std::vector<int64_t> off_copy;
    if (!h_off) {
        off_copy.resize((size_t)n_irms + 1);
        CK(cudaMemcpyAsync(off_copy.data(), args.move_off, ((size_t)n_irms + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, e->stream));
        h_off = off_copy.data();
    }
    const int RUNCAP = 64;
    CK(e->s_runs_n.ensure((size_t)n_gen)); CK(e->s_runs_start.ensure((size_t)n_gen * RUNCAP)); CK(e->s_runs_dom.ensure((size_t)n_gen * RUNCAP));
    move_runs_kernel<<<(unsigned)((n_gen * 32 + 255) / 256), 256, 0, e->stream>>>(args.move_off, args.move_dom, e->s_blk.p, n_gen, RUNCAP, e->s_runs_n.p,
                                                                             e->s_runs_start.p, e->s_runs_dom.p);
    CK(cudaGetLastError());
    std::vector<int32_t> rn((size_t)n_gen), rdm((size_t)n_gen * RUNCAP);
    std::vector<int64_t> rst((size_t)n_gen * RUNCAP);
    CK(cudaMemcpyAsync(rn.data(), e->s_runs_n.p, rn.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(rdm.data(), e->s_runs_dom.p, rdm.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(rst.data(), e->s_runs_start.p, rst.size() * sizeof(int64_t), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    const int64_t MINRUN = mode == 1 ? 1 : 32;            // shorter runs are not worth a launch (forced mode: every run, for the tests)
    // which runs qualify, and the staging pitch of every IRM
    std::vector<int32_t> pitch((size_t)n_irms, 0);
    int64_t pitch_sum = 0; bool any = false;
    auto run_ok = [&](const Irm *irm, int dm, int64_t len) {
        return dm >= 0 && dm < irm->n_domains && dm < (int)irm->pregroup.size() && irm->pregroup[(size_t)dm] && len >= MINRUN;
    };
    for (int w = 0; w < n_gen; w++) {
        const int k = plan.generic_blk[(size_t)w];
        const Irm *irm = e->irms[h_irms[k]].get();
        if (rn[(size_t)w] <= 0 || rn[(size_t)w] > RUNCAP) continue;
        for (int j = 0; j < rn[(size_t)w]; j++) {
            const int64_t st = rst[(size_t)w * RUNCAP + j], en = j + 1 < rn[(size_t)w] ? rst[(size_t)w * RUNCAP + j + 1] : h_off[k + 1];
            const int dm = rdm[(size_t)w * RUNCAP + j];
            if (!run_ok(irm, dm, en - st)) continue;
            any = true;
            const int64_t groups = std::min<int64_t>(irm->obs->dom[(size_t)dm].max_deg, irm->ctot[(size_t)dm]);
            pitch[(size_t)k] = std::max(pitch[(size_t)k], (int32_t)groups + 1);
        }
        pitch_sum += pitch[(size_t)k];
    }
    if (!any) return launch_sweep(e, plan, args);
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    size_t budget = std::min<size_t>((free_b + e->s_pg_data.n * sizeof(uint2)) / 2, (size_t)8 << 30);
    if (const char *s = getenv("HIRM_PREGROUP_MB")) budget = (size_t)std::max(1, atoi(s)) << 20;
    int64_t seg = std::min<int64_t>(4096, (int64_t)(budget / sizeof(uint2)) / std::max<int64_t>(pitch_sum, 1));
    if (const char *s = getenv("HIRM_PREGROUP_SEG")) seg = std::max(1, atoi(s));
    else if (seg < 32) return launch_sweep(e, plan, args);
    // pieces (csrc/sweep_rounds.cc), over the generic IRMs in launch order; then scattered to the sweep's IRM index
    std::vector<int64_t> goff((size_t)n_gen + 1, 0);
    std::vector<int32_t> rok((size_t)n_gen * RUNCAP, 0);
    // the planner works on lists laid end to end in launch order; the runs' positions are shifted accordingly
    std::vector<int64_t> shift((size_t)n_gen, 0), rst_g(rst.size(), 0);
    for (int w = 0; w < n_gen; w++) {
        const int k = plan.generic_blk[(size_t)w];
        const Irm *irm = e->irms[h_irms[k]].get();
        goff[(size_t)w + 1] = goff[(size_t)w] + (h_off[k + 1] - h_off[k]);
        shift[(size_t)w] = goff[(size_t)w] - h_off[k];
        if (rn[(size_t)w] <= 0 || rn[(size_t)w] > RUNCAP) continue;
        for (int j = 0; j < rn[(size_t)w]; j++) {
            const int64_t st = rst[(size_t)w * RUNCAP + j], en = j + 1 < rn[(size_t)w] ? rst[(size_t)w * RUNCAP + j + 1] : h_off[k + 1];
            rok[(size_t)w * RUNCAP + j] = run_ok(irm, rdm[(size_t)w * RUNCAP + j], en - st) ? 1 : 0;
            rst_g[(size_t)w * RUNCAP + j] = st + shift[(size_t)w];
        }
    }
    int32_t nr32 = 1;
    if (hirm_plan_rounds(n_gen, goff.data(), rn.data(), RUNCAP, rst_g.data(), rok.data(), seg, &nr32, nullptr, nullptr, nullptr, 0))
        return fail("pre-grouping: inconsistent runs of the move lists");
    const size_t n_rounds = (size_t)nr32;
    std::vector<int64_t> glo(n_rounds * (size_t)n_gen), ghi(n_rounds * (size_t)n_gen);
    std::vector<int32_t> gpre(n_rounds * (size_t)n_gen);
    if (hirm_plan_rounds(n_gen, goff.data(), rn.data(), RUNCAP, rst_g.data(), rok.data(), seg, &nr32, glo.data(), ghi.data(), gpre.data(), (int)n_rounds))
        return fail("pre-grouping: inconsistent runs of the move lists");
    std::vector<int64_t> lo_all(n_rounds * (size_t)n_irms, 0), hi_all(n_rounds * (size_t)n_irms, 0), base((size_t)n_irms, 0);
    std::vector<int32_t> pre_all(n_rounds * (size_t)n_irms, 0);
    int64_t entries = 0;
    for (int k = 0; k < n_irms; k++) { base[(size_t)k] = entries; entries += (int64_t)pitch[(size_t)k] * seg; }
    for (int w = 0; w < n_gen; w++) {
        const int k = plan.generic_blk[(size_t)w];
        for (size_t r = 0; r < n_rounds; r++) {
            const int64_t a = glo[r * n_gen + w], b = ghi[r * n_gen + w];
            if (a >= b) continue;
            lo_all[r * n_irms + k] = a - shift[(size_t)w]; hi_all[r * n_irms + k] = b - shift[(size_t)w]; pre_all[r * n_irms + k] = gpre[r * n_gen + w];
        }
    }
    for (int32_t kd : plan.dense_blk) { lo_all[(size_t)kd] = h_off[kd]; hi_all[(size_t)kd] = h_off[kd + 1]; }   // (the dense kernel sweeps whole lists, in round 0)
    CK(e->s_seg_lo.ensure(lo_all.size())); CK(e->s_seg_hi.ensure(hi_all.size())); CK(e->s_pg_pre.ensure(pre_all.size()));
    CK(e->s_pg_base.ensure((size_t)n_irms)); CK(e->s_pg_pitch.ensure((size_t)n_irms));
    if (e->s_pg_data.n < (size_t)entries) {
        e->s_pg_data.release();
        if (e->s_pg_data.alloc((size_t)entries) != cudaSuccess) {      // no room for the staging buffer: the ordinary path needs none
            cudaGetLastError();
            e->s_pg_data.p = nullptr; e->s_pg_data.n = 0;
            return launch_sweep(e, plan, args);
        }
    }
    cudaStream_t st = e->stream;
    CK(cudaMemcpyAsync(e->s_seg_lo.p, lo_all.data(), lo_all.size() * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(e->s_seg_hi.p, hi_all.data(), hi_all.size() * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(e->s_pg_pre.p, pre_all.data(), pre_all.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(e->s_pg_base.p, base.data(), base.size() * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(e->s_pg_pitch.p, pitch.data(), pitch.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));                         // (the sources are locals)
    e->last_rounds = (int)n_rounds; e->last_pregrouped = 1;
    for (size_t r = 0; r < n_rounds; r++) {
        SweepArgs a = args;
        a.seg_lo = e->s_seg_lo.p + r * n_irms; a.seg_hi = e->s_seg_hi.p + r * n_irms; a.pg_pre = e->s_pg_pre.p + r * n_irms;
        a.pg_base = e->s_pg_base.p; a.pg_pitch = e->s_pg_pitch.p; a.pg_data = e->s_pg_data.p;
        if (launch_sweep(e, plan, a, (int)r)) return -1;
    }
    return 0;
}

static int upload_blk(hirm_engine *e, const Plan &plan) {
    std::vector<int32_t> blk(plan.generic_blk);
    blk.insert(blk.end(), plan.dense_blk.begin(), plan.dense_blk.end());
    CK(e->s_blk.ensure(blk.size()));
    CK(cudaMemcpyAsync(e->s_blk.p, blk.data(), blk.size() * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return 0;
}

extern "C" int hirm_engine_sweep(hirm_engine *e, int n_irms, const int32_t *h_irms, const int64_t *h_move_offsets,
                                 const int32_t *h_move_dom, const int32_t *h_move_item, const double *h_uniforms,
                                 uint64_t seed, const uint64_t *h_stream, const uint64_t *h_counter0, uint32_t flags,
                                 int32_t *h_out_choice, int trace_cap, int32_t *h_trace_n, int32_t *h_trace_labels,
                                 double *h_trace_logps) {
    if (!e) return fail("engine is null");
    CK(cudaSetDevice(e->device));
    if (n_irms <= 0) return 0;
    if (!h_irms || !h_move_offsets || !h_move_dom || !h_move_item) return fail("null argument");
    if (trace_cap > 0 && (!h_trace_n || !h_trace_labels || !h_trace_logps)) return fail("null trace buffers");
    if (h_move_offsets[0] != 0) return fail("move offsets must start at 0");
    for (int k = 0; k < n_irms; k++) if (h_move_offsets[k + 1] < h_move_offsets[k]) return fail("move offsets must be non-decreasing");
    const int64_t nm = h_move_offsets[n_irms];
    Plan plan;
    if (plan_sweep(e, n_irms, h_irms, flags, plan)) return -1;
    if (upload_blk(e, plan)) return -1;
    const size_t nmz = (size_t)std::max<int64_t>(nm, 1);
    std::vector<uint64_t> zeros((size_t)n_irms, 0), streams((size_t)n_irms);
    for (int k = 0; k < n_irms; k++) streams[(size_t)k] = h_stream ? h_stream[k] : (uint64_t)h_irms[k];
    CK(e->s_irm_ids.ensure((size_t)n_irms)); CK(e->s_move_off.ensure((size_t)n_irms + 1));
    CK(e->s_move_dom.ensure(nmz)); CK(e->s_move_item.ensure(nmz)); CK(e->s_choice.ensure(nmz));
    CK(e->s_stream.ensure((size_t)n_irms)); CK(e->s_counter.ensure((size_t)n_irms));
    cudaStream_t st = e->stream;
    CK(cudaMemcpyAsync(e->s_irm_ids.p, h_irms, n_irms * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(e->s_move_off.p, h_move_offsets, (n_irms + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(e->s_move_dom.p, h_move_dom, nm * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(e->s_move_item.p, h_move_item, nm * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(e->s_stream.p, streams.data(), n_irms * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(e->s_counter.p, h_counter0 ? h_counter0 : zeros.data(), n_irms * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    if (h_uniforms) {
        CK(e->s_uniforms.ensure(nmz));
        CK(cudaMemcpyAsync(e->s_uniforms.p, h_uniforms, nm * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    if (trace_cap > 0) {
        CK(e->s_trace_n.ensure(nmz)); CK(e->s_trace_labels.ensure(nmz * trace_cap)); CK(e->s_trace_logps.ensure(nmz * trace_cap));
        // entries beyond a move's candidates read as zero
        CK(cudaMemsetAsync(e->s_trace_labels.p, 0, nmz * trace_cap * sizeof(int32_t), st));
        CK(cudaMemsetAsync(e->s_trace_logps.p, 0, nmz * trace_cap * sizeof(double), st));
    }
    SweepArgs args;
    memset(&args, 0, sizeof(args));
    args.irms = e->d_irms.p; args.irm_ids = e->s_irm_ids.p; args.move_off = e->s_move_off.p;
    args.move_dom = e->s_move_dom.p; args.move_item = e->s_move_item.p;
    args.uniforms = h_uniforms ? e->s_uniforms.p : nullptr;
    args.seed = seed; args.stream = e->s_stream.p; args.counter0 = e->s_counter.p; args.flags = flags;
    args.out_choice = e->s_choice.p; args.phase_cycles = e->d_phase_cycles;
    args.trace_cap = trace_cap; args.trace_n = e->s_trace_n.p; args.trace_labels = e->s_trace_labels.p; args.trace_logps = e->s_trace_logps.p;
    if (run_sweep(e, plan, args, n_irms, h_irms, h_move_offsets)) return -1;
    if (h_out_choice) CK(cudaMemcpyAsync(h_out_choice, e->s_choice.p, nm * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (trace_cap > 0) {
        CK(cudaMemcpyAsync(h_trace_n, e->s_trace_n.p, nm * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(h_trace_labels, e->s_trace_labels.p, nm * trace_cap * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(h_trace_logps, e->s_trace_logps.p, nm * trace_cap * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&e->last_ms, e->ev0, e->ev1));
    return check_kernel_errors(e, n_irms);
}

static int sweep_device_impl(hirm_engine *e, int n_irms, const int32_t *h_irms, const int64_t *d_move_offsets,
                             const int32_t *d_move_dom, const int32_t *d_move_item, uint64_t seed,
                             const uint64_t *d_stream, const uint64_t *d_counter0, uint32_t flags, int32_t *d_out_choice) {
    Plan plan;
    if (plan_sweep(e, n_irms, h_irms, flags, plan)) return -1;
    if (upload_blk(e, plan)) return -1;
    CK(e->s_irm_ids.ensure((size_t)n_irms));
    CK(cudaMemcpyAsync(e->s_irm_ids.p, h_irms, n_irms * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    SweepArgs args;
    memset(&args, 0, sizeof(args));
    args.irms = e->d_irms.p; args.irm_ids = e->s_irm_ids.p; args.move_off = d_move_offsets;
    args.move_dom = d_move_dom; args.move_item = d_move_item; args.uniforms = nullptr;
    args.seed = seed; args.stream = d_stream; args.counter0 = d_counter0; args.flags = flags;
    args.out_choice = d_out_choice; args.trace_cap = 0; args.phase_cycles = e->d_phase_cycles;
    return run_sweep(e, plan, args, n_irms, h_irms, nullptr);
}

extern "C" int hirm_engine_sweep_device(hirm_engine *e, int n_irms, const int32_t *h_irms, const int64_t *d_move_offsets,
                                        const int32_t *d_move_dom, const int32_t *d_move_item, uint64_t seed,
                                        const uint64_t *d_stream, const uint64_t *d_counter0, uint32_t flags,
                                        int32_t *d_out_choice) {
    if (!e) return fail("engine is null");
    CK(cudaSetDevice(e->device));
    if (n_irms <= 0) return 0;
    if (!h_irms || !d_move_offsets || !d_move_dom || !d_move_item || !d_stream || !d_counter0) return fail("null argument");
    return sweep_device_impl(e, n_irms, h_irms, d_move_offsets, d_move_dom, d_move_item, seed, d_stream, d_counter0, flags, d_out_choice);
}

// IRM::transition_cluster_assignments_all (cxx/hirm.hh:590-596), n_sweeps times, for every listed IRM: the sweep order
// is generated on the device (gen_moves_kernel), nothing but the IRM ids crosses the bus on the way in.
extern "C" int hirm_engine_sweep_all(hirm_engine *e, int n_irms, const int32_t *h_irms, int n_sweeps, uint64_t seed, uint32_t flags,
                                     int64_t cap_moves, int32_t *h_out_dom, int32_t *h_out_item, int32_t *h_out_choice, int64_t *n_moves) {
    if (!e) return fail("engine is null");
    CK(cudaSetDevice(e->device));
    if (n_moves) *n_moves = 0;
    if (n_irms <= 0 || n_sweeps <= 0) return 0;
    if (!h_irms) return fail("null argument");
    std::vector<int64_t> off((size_t)n_irms + 1, 0);
    std::vector<uint64_t> st((size_t)n_irms), ct((size_t)n_irms), sw((size_t)n_irms);
    for (int k = 0; k < n_irms; k++) {
        Irm *irm = get_irm(e, h_irms[k]);
        if (!irm) return fail("bad irm id in sweep list");
        int64_t per = 0;
        for (int d = 0; d < irm->n_domains; d++) {
            if (!irm->have_z[d]) return fail("set_assignments has not been called for every domain");
            if (irm->has_absent[d]) return fail("sweep_all needs every entity seated: pass an explicit move list to hirm_engine_sweep");
            per += irm->n_ent[d];
        }
        off[(size_t)k + 1] = off[(size_t)k] + per * n_sweeps;
        st[(size_t)k] = irm->rng_stream; ct[(size_t)k] = irm->move_counter; sw[(size_t)k] = irm->sweep_counter;
    }
    const int64_t nm = off[(size_t)n_irms];
    if (n_moves) *n_moves = nm;
    if ((h_out_dom || h_out_item || h_out_choice) && cap_moves < nm) return fail("output buffers are too small for the generated moves");
    const size_t nmz = (size_t)std::max<int64_t>(nm, 1);
    // descriptors must be on the device before the order kernel reads the domain sizes
    for (int k = 0; k < n_irms; k++) if (ensure_counts(e, h_irms[k])) return -1;
    CK(e->s_irm_ids.ensure((size_t)n_irms)); CK(e->s_move_off.ensure((size_t)n_irms + 1));
    CK(e->s_move_dom.ensure(nmz)); CK(e->s_move_item.ensure(nmz)); CK(e->s_choice.ensure(nmz));
    CK(e->s_stream.ensure((size_t)n_irms)); CK(e->s_counter.ensure((size_t)n_irms)); CK(e->s_sweep0.ensure((size_t)n_irms));
    cudaStream_t s0 = e->stream;
    CK(cudaMemcpyAsync(e->s_irm_ids.p, h_irms, n_irms * sizeof(int32_t), cudaMemcpyHostToDevice, s0));
    CK(cudaMemcpyAsync(e->s_move_off.p, off.data(), (n_irms + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s0));
    CK(cudaMemcpyAsync(e->s_stream.p, st.data(), n_irms * sizeof(uint64_t), cudaMemcpyHostToDevice, s0));
    CK(cudaMemcpyAsync(e->s_counter.p, ct.data(), n_irms * sizeof(uint64_t), cudaMemcpyHostToDevice, s0));
    CK(cudaMemcpyAsync(e->s_sweep0.p, sw.data(), n_irms * sizeof(uint64_t), cudaMemcpyHostToDevice, s0));
    CK(cudaStreamSynchronize(s0));                      // the sources above are locals
    gen_moves_kernel<<<n_irms, AUX_THREADS, 0, s0>>>(e->d_irms.p, e->s_irm_ids.p, e->s_move_off.p, n_sweeps, seed, e->s_stream.p, e->s_sweep0.p,
                                                    e->s_move_dom.p, e->s_move_item.p);
    CK(cudaGetLastError());
    if (sweep_device_impl(e, n_irms, h_irms, e->s_move_off.p, e->s_move_dom.p, e->s_move_item.p, seed, e->s_stream.p, e->s_counter.p, flags, e->s_choice.p))
        return -1;
    e->last_launches++;                                 // the order kernel
    for (int k = 0; k < n_irms; k++) {
        Irm *irm = e->irms[h_irms[k]].get();
        irm->move_counter += (uint64_t)(off[(size_t)k + 1] - off[(size_t)k]); irm->sweep_counter += (uint64_t)n_sweeps;
    }
    if (h_out_dom) CK(cudaMemcpyAsync(h_out_dom, e->s_move_dom.p, nm * sizeof(int32_t), cudaMemcpyDeviceToHost, s0));
    if (h_out_item) CK(cudaMemcpyAsync(h_out_item, e->s_move_item.p, nm * sizeof(int32_t), cudaMemcpyDeviceToHost, s0));
    if (h_out_choice) CK(cudaMemcpyAsync(h_out_choice, e->s_choice.p, nm * sizeof(int32_t), cudaMemcpyDeviceToHost, s0));
    CK(cudaStreamSynchronize(s0));
    return check_kernel_errors(e, n_irms);
}

extern "C" int hirm_engine_last_sweep_info(hirm_engine *e, int32_t *n_launches, int32_t *kernel_kind, float *kernel_ms) {
    if (!e) return fail("engine is null");
    if (n_launches) *n_launches = e->last_launches;
    if (kernel_kind) *kernel_kind = e->last_kind;
    if (kernel_ms) {
        CK(cudaEventSynchronize(e->ev1));
        CK(cudaEventElapsedTime(&e->last_ms, e->ev0, e->ev1));
        *kernel_ms = e->last_ms;
    }
    return 0;
}

extern "C" int hirm_engine_last_cluster_size(hirm_engine *e, int32_t *ctas_per_irm) {
    if (!e || !ctas_per_irm) return fail("null argument");
    *ctas_per_irm = e->last_cluster;
    return 0;
}

extern "C" int hirm_engine_last_rounds(hirm_engine *e, int32_t *rounds, int32_t *pregrouped) {
    if (!e) return fail("engine is null");
    if (rounds) *rounds = e->last_rounds;
    if (pregrouped) *pregrouped = e->last_pregrouped;
    return 0;
}

extern "C" int hirm_engine_last_dense_plan(hirm_engine *e, int32_t *ctas_per_sm, int32_t *stages, int32_t *chunk_bytes, int32_t *smem_bytes) {
    if (!e) return fail("engine is null");
    if (ctas_per_sm) *ctas_per_sm = e->last_occupancy;
    if (stages) *stages = e->last_stages;
    if (chunk_bytes) *chunk_bytes = e->last_chunk;
    if (smem_bytes) *smem_bytes = e->last_smem;
    return 0;
}

// Debug / test hook: S(N+n, s+sg) - S(N, s) for arrays of operands with the kernels' own fp64 code (score_delta_bf).
extern "C" int hirm_engine_debug_score_delta(hirm_engine *e, int n, const uint32_t *h_N, const uint32_t *h_s, const uint32_t *h_gn,
                                             const uint32_t *h_gs, double *h_out) {
    if (!e) return fail("engine is null");
    CK(cudaSetDevice(e->device));
    if (n <= 0) return 0;
    DevBuf<uint32_t> d[4]; DevBuf<double> o;
    const uint32_t *src[4] = {h_N, h_s, h_gn, h_gs};
    for (int k = 0; k < 4; k++) { CK(d[k].alloc((size_t)n)); CK(cudaMemcpy(d[k].p, src[k], (size_t)n * 4, cudaMemcpyHostToDevice)); }
    CK(o.alloc((size_t)n));
    debug_score_delta_kernel<<<std::min((n + AUX_THREADS - 1) / AUX_THREADS, 1024), AUX_THREADS, 0, e->stream>>>(d[0].p, d[1].p, d[2].p, d[3].p, o.p, n);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpy(h_out, o.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    return 0;
}

// Debug hook: give the dense kernel a device buffer [n_blocks * 16] of int64 to fill with thread 0's
// per-phase SM cycles (wait, accumulate, reduce, candidates, score, sample, apply).  NULL switches it off.
extern "C" int hirm_engine_debug_phase_cycles(hirm_engine *e, long long *d_buf) {
    if (!e) return fail("engine is null");
    e->d_phase_cycles = d_buf;
    return 0;
}
