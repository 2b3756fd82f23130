// Device-visible descriptors of one IRM's state.  Built by the host (engine.cu),
// uploaded once per change; all pointers are device pointers.
#pragma once
#include <cstdint>
#include "device_math.cuh"

namespace hirm {

constexpr int MAX_ARITY = 4;
constexpr int MAX_DOMAINS = 8;
constexpr int SELF_NONE = -1;

// meta word of one CSR entry: relation index, observed value, mask of the positions
// at which the row's entity itself sits
constexpr int META_REL_BITS = 20;
constexpr uint32_t META_REL_MASK = (1u << META_REL_BITS) - 1;
constexpr int META_VAL_SHIFT = 20;
constexpr int META_SELF_SHIFT = 24;

// What phase A of the generic kernel needs to know about a CSR relation while an entity of domain d moves (a copy of the
// RelationDev / RelDom fields it reads, 64 bytes: kept in shared memory).  CSR entries name their relation by its ORDINAL
// among the CSR relations touching the domain.
struct ADesc {
    int32_t arity, dom[4];        // positions' domains
    int32_t a, b;                 // positions of the moved domain (b = -1: it occurs once)
    int32_t base[3], rest_size;   // accumulator space (RelDom)
    int32_t rstride[4];
    int32_t pad;
};
static_assert(sizeof(ADesc) == 64, "ADesc layout");

struct DomainDev {
    int32_t n;            // entities
    int32_t kcap;         // table slots
    int32_t *z;           // [n] slot of each entity, -1 = not in the domain        (CRP::assignments, cxx/hirm.hh:75)
    uint8_t *zb;          // [n] the same as bytes (0xFF = not in the domain): what the generic kernel's gathers read -- a quarter of the
                          //     footprint, so that the slot arrays of hundreds of chains stay in the L2
    int32_t *tab_count;   // [kcap] customers per slot, 0 = free                     (CRP::tables sizes, :74)
    int32_t *tab_label;   // [kcap] reference table label of the slot
    int32_t *hi;          // [1] highest used slot + 1
    double alpha;         // CRP::alpha (:72)
    // per-entity CSR over ALL CSR-stored relations touching this domain            (Relation::data_r, :251)
    const int64_t *ptr;   // [n+1] or null
    const uint32_t *meta; // [nnz]
    const int32_t *other; // [n_other][nnz]: entity ids at every position except the first self position
    int64_t nnz;
    int32_t n_other;
    int32_t pad;
    int64_t max_deg;      // longest CSR row of the domain + its unary-block relations (bounds a move's group count)
    // unary relations of the domain kept as a block of entity-major bit rows (hirm_unary_block_create): row i = entity i's
    // observations in every relation of the block, `ub_words` words per plane.  The block is relation-owned and shared
    // by all IRMs.  An IRM holds `nu` of the block's columns (local column j = block column ucolg[j]) and keeps THEIR count
    // tables as one matrix ucount[slot * ucp + j]: a candidate table's cells of 32 relations are one 256-byte read.
    const uint32_t *ub_obs;   // [n][ub_words] observed
    const uint32_t *ub_val;   // [n][ub_words] value
    const int32_t *ucolg;     // [nu] block column of local column j
    cell_t *ucount;           // [kcap][ucp] (N | s << 32) of (table slot, local column)
    int32_t nu, ucp;          // local columns; row pitch of ucount (nu rounded up to 32)
    int32_t ub_words;
    int32_t n_adesc;          // CSR relations touching the domain
    const ADesc *adesc;       // [n_adesc] by ordinal
};

struct RelationDev {
    int32_t arity;
    int32_t kind;                 // 0 = CSR, 1 = dense slabs, 2 = column of a unary block
    int32_t dom[MAX_ARITY];
    cell_t *counts;               // count table (Relation::clusters, :246).  Dense: `ncells` cells, index = sum_p slot_p * cstride[p].
                                  // Hashed (hcap > 0): hcap open-addressing entries {key = index + 1 (0 = empty), cell}, 16 bytes each
    int64_t cstride[MAX_ARITY];
    int64_t ncells;               // cells of the index space (prod of the domains' table slots)
    uint32_t hcap;                // hashed: entries (a power of two); 0 = dense
    uint32_t estride;             // dense table: distance between consecutive cells of the index space (1; a unary-block relation's cells are a column of its domain's count matrix)
    uint32_t *hused;              // hashed: [1] entries taken so far (emptied cells stay as zero cells until the host rebuilds)
    const uint8_t *slab[2];       // dense: rows by first / by second entity
    int64_t pitch[2];
    const int32_t *coo_ids;       // [nobs*arity] (CSR relations: kept for count-table rebuilds)
    const uint8_t *coo_val;
    int64_t nobs;
    int32_t full_obs;             // dense: 1 if no cell of the relation is missing
    int32_t pad;
};

// Group accumulators of the generic kernel.  While entity i of domain d moves, its observations in relation r fall into
// groups keyed by (which positions of the observation hold i itself, the tables of the other positions)
// (get_cluster_to_items_list, cxx/hirm.hh:388-397).  The keys that can occur for (r, d) form a small space: with a, b
// the positions of d in r (b = -1 if d occurs once) and "rest" the remaining positions,
//   pattern 0: i at a only      cell = base[0] + rest_k + (table of the entity at b) * rest_size      (b >= 0)
//   pattern 1: i at b only      cell = base[1] + rest_k + (table of the entity at a) * rest_size
//   pattern 2: i at a and b     cell = base[2] + rest_k
// rest_k = sum over the rest positions of table * rstride.  All (r, d) spaces of one domain are laid end to end
// (`ctot[d]` cells); the accumulator of a move lives in shared memory when that fits, else in the IRM's global scratch.
struct RelDom {
    int32_t a, b;                 // a < 0: the relation does not touch the domain
    int32_t base[3];
    int32_t rest_size;
    int32_t rstride[MAX_ARITY];   // of the rest positions
};

struct IrmDev {
    int32_t n_domains, n_relations;
    DomainDev dom[MAX_DOMAINS];
    const RelationDev *rel;       // [n_relations]
    // generic-kernel group accumulators
    const RelDom *rdesc;          // [n_relations * n_domains]
    const int32_t *cmap[MAX_DOMAINS];   // [ctot[d]] relation owning each accumulator cell of domain d
    const void *ctmpl[MAX_DOMAINS];     // [ctot[d]] GroupRec template of each cell (generic_sweep.cuh), or null: decoded per move
    int32_t ctot[MAX_DOMAINS];
    cell_t *gacc;                 // [max ctot] accumulator, all zero between moves       (used when ctot[d] > shared capacity)
    uint32_t *gbm;                // [max ctot / 32] one bit per non-zero accumulator cell
    int64_t *gl_cell;             // [gl_cap] group records beyond the shared-memory ones
    int64_t gl_cap;
    int32_t *error;               // [1] first error code raised by a kernel (0 = none)
};

// kernel error codes (reported through hirm_last_error)
enum : int32_t {
    KERR_NONE = 0,
    KERR_NO_FREE_SLOT = 1,     // max_tables exhausted
    KERR_ABSENT_ENTITY = 2,    // observation endpoint with z = -1
    KERR_GROUP_OVERFLOW = 3,   // group list capacity exceeded
    KERR_BAD_MOVE = 4,         // move names an entity outside the domain
    KERR_TABLE_FULL = 5,       // hashed count table: no free entry
};

__device__ __forceinline__ void raise_error(const IrmDev &irm, int32_t code) { atomicCAS(irm.error, 0, code); }

// ---- count-table access: dense array or open-addressing hash (linear probing) keyed by the cell index -----------------
// Relation::clusters is a sparse map in the reference (create on demand :540-543, erase on zero :534-537).  The hashed
// table keeps an emptied cell as a zero cell: it reads as absent (score 0, not listed by get_counts); the host drops such
// entries when it rebuilds the table.  One 16-byte entry = {key, (N | s << 32)}: a probe is one ld.cg.v2.
__device__ __forceinline__ uint32_t hash_slot(int64_t cidx, uint32_t mask) {
    return (uint32_t)(((uint64_t)cidx * 0x9E3779B97F4A7C15ull) >> 32) & mask;
}
__device__ __forceinline__ cell_t table_get(const cell_t *counts, uint32_t hcap, int64_t cidx) {
    if (!hcap) return __ldcg(&counts[cidx]);
    const uint32_t mask = hcap - 1u;
    const unsigned long long key = (unsigned long long)cidx + 1ull;
    const ulonglong2 *tab = reinterpret_cast<const ulonglong2 *>(counts);
    for (uint32_t i = hash_slot(cidx, mask), n = 0; n <= mask; i = (i + 1u) & mask, n++) {
        const ulonglong2 e = __ldcg(tab + i);
        if (e.x == key) return e.y;
        if (e.x == 0ull) return 0ull;
    }
    return 0ull;
}
__device__ __forceinline__ cell_t table_get(const RelationDev *R, int64_t cidx) { return table_get(R->counts, R->hcap, cidx); }
__device__ __forceinline__ void table_add(const IrmDev &irm, const RelationDev *R, int64_t cidx, cell_t delta) {
    if (!R->hcap) { atomicAdd(&R->counts[cidx], delta); return; }
    const uint32_t mask = R->hcap - 1u;
    const unsigned long long key = (unsigned long long)cidx + 1ull;
    for (uint32_t i = hash_slot(cidx, mask), n = 0; n <= mask; i = (i + 1u) & mask, n++) {
        unsigned long long k = __ldcg(&R->counts[2 * (size_t)i]);
        if (k == 0ull) {
            k = atomicCAS(&R->counts[2 * (size_t)i], 0ull, key);
            if (k == 0ull) { atomicAdd(R->hused, 1u); k = key; }
        }
        if (k == key) { atomicAdd(&R->counts[2 * (size_t)i + 1], delta); return; }
    }
    raise_error(irm, KERR_TABLE_FULL);
}
// iteration over the stored cells (scores, read-back): entries of the table and the cell of entry i (0 = empty)
__device__ __forceinline__ int64_t table_entries(const RelationDev *R) { return R->hcap ? (int64_t)R->hcap : R->ncells; }
__device__ __forceinline__ cell_t table_entry(const RelationDev *R, int64_t i) {
    if (!R->hcap) return R->counts[i * (int64_t)R->estride];
    return R->counts[2 * i] ? R->counts[2 * i + 1] : 0ull;
}

struct SweepArgs {
    const IrmDev *irms;           // [all irms of the engine]
    const int32_t *irm_ids;       // [n_irms] which IRM each block sweeps
    const int64_t *move_off;      // [n_irms+1]
    const int32_t *move_dom;
    const int32_t *move_item;
    const double *uniforms;       // nullable
    uint64_t seed;
    const uint64_t *stream;       // [n_irms]
    const uint64_t *counter0;     // [n_irms]
    uint32_t flags;
    int32_t *out_choice;          // nullable
    int32_t trace_cap;
    int32_t *trace_n;
    int32_t *trace_labels;
    double *trace_logps;
    long long *progress;          // nullable: [n_blocks] moves of the block's IRM already made / made so far (dense tiers)
    long long *phase_cycles;      // nullable debug hook: [n_blocks][8] SM cycles per phase (thread 0's clock64)
    // Pre-grouped rows (pregroup.cuh).  A launch may sweep a PIECE of every IRM's move list, [seg_lo, seg_hi); where pg_pre is
    // set, the piece is a run of moves of ONE domain whose groups do not depend on that domain's own partition, and the group
    // list of every move of the piece has been staged: pg_pitch entries per move, entry 0 = {groups, 0}, then
    // {accumulator cell, observations | ones << 16} in ascending cell order.
    const int64_t *seg_lo, *seg_hi;   // nullable: [n_irms]
    const int32_t *pg_pre;            // nullable: [n_irms]
    const int64_t *pg_base;           // [n_irms] first staging entry of the IRM's piece
    const int32_t *pg_pitch;          // [n_irms]
    uint2 *pg_data;
};
constexpr int PREGROUP_MAX_CELLS = 1 << 15;    // accumulator cells of a domain the pre-grouping kernel keeps in shared memory

}  // namespace hirm
