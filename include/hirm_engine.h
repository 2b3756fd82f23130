/*
 * hirm_engine.h -- C ABI of the B200 engine for HIRM's collapsed-Gibbs entity move.
 *
 * The reference (probsys/hierarchical-irm) has no FFI; its boundary for this path is the
 * method surface of IRM / Relation / Domain / CRP.  Each entry point below names the
 * reference interface it replaces (paths relative to the reference tree).  Plain C
 * types only: no torch / STL types cross this boundary.  INTEGRATION.md shows the
 * ctypes binding used by the Python facade and the C++ shim a maintainer would add to
 * cxx/hirm.cc.
 *
 * Conventions
 *  - every function returns 0 on success, <0 on error; hirm_last_error() gives the text
 *    (thread local).  Nothing aborts or throws across the boundary (the reference
 *    asserts/aborts: cxx/hirm.hh:35,205,293).
 *  - h_ = host pointer, d_ = device pointer (caller allocated, never freed by the engine).
 *  - entities are int32 codes 0..n-1 per domain (cxx/util_io.cc:63-96 encode_observations),
 *    tables are identified by int32 *labels* exactly as in the reference (new label =
 *    max label + 1, never compacted: cxx/hirm.hh:133-161); the engine keeps a private
 *    label<->slot map.
 *  - one engine per GPU and host thread; one engine holds many IRMs (the relation
 *    clusters of an HIRM, or independent chains) and sweeps them in one launch.
 *  - there is no CPU fallback: every call fails with an error if CUDA is unusable.
 */
#ifndef HIRM_ENGINE_H
#define HIRM_ENGINE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HIRM_MAX_ARITY 4      /* positions per relation */
#define HIRM_MAX_DOMAINS 8    /* domains per IRM */
#define HIRM_DENSE_MISSING 0xFF

typedef struct hirm_engine hirm_engine;

/* sweep flags */
#define HIRM_SWEEP_SCORE_ONLY 1u   /* score + trace, do not sample/apply (parity hook) */
#define HIRM_SWEEP_FORCE_GENERIC 2u /* use the CSR/generic kernel even where a fast path exists */
#define HIRM_SWEEP_CLUSTER 4u      /* one thread-block cluster per IRM (up to 16 CTAs share a move's row and scoring units through
                                      distributed shared memory) even when the engine's heuristic would not choose it; the engine uses
                                      clusters by itself when few IRMs with heavy moves are swept.  Results do not depend on it. */

const char *hirm_last_error(void);
int hirm_engine_abi_version(void);

/* Engine lifetime.  device < 0 selects the current CUDA device. */
int hirm_engine_create(hirm_engine **out, int device);
int hirm_engine_destroy(hirm_engine *e);
/* run all engine work on this CUDA stream (a cudaStream_t cast to void*); NULL = default stream */
int hirm_engine_set_stream(hirm_engine *e, void *cuda_stream);
int hirm_engine_synchronize(hirm_engine *e);

/* IRM::IRM(schema) + IRM::add_relation   (cxx/hirm.hh:570-575,742-757; src/hirm.py:379-386,435-446)
 * rel_domains: [n_relations * HIRM_MAX_ARITY] domain index per position.
 * max_tables:  [n_domains] table slots kept on the device per domain (>= tables ever alive + 1). */
int hirm_irm_create(hirm_engine *e, int n_domains, const int32_t *n_entities, const int32_t *max_tables,
                    int n_relations, const int32_t *arity, const int32_t *rel_domains, int32_t *out_irm);
int hirm_irm_destroy(hirm_engine *e, int32_t irm);

/* Bulk IRM::incorporate(r, items, value)   (cxx/hirm.hh:582 -> Relation::incorporate :271-290;
 * cxx/util_io.cc:98-112 incorporate_observations).  COO on the host: h_ids [nobs*arity], h_val
 * [nobs] in {0,1}.  Builds the per-entity CSR index that replaces Relation::data_r
 * (cxx/hirm.hh:251,274-280): one entry per (entity, observation), an observation listed once per
 * entity even if the entity sits at two positions.  Observations must be distinct
 * (cxx/hirm.hh:272).  Call for every relation, then hirm_irm_finalize(). */
int hirm_irm_set_observations(hirm_engine *e, int32_t irm, int rel, int64_t nobs,
                              const int32_t *h_ids, const uint8_t *h_val);
/* Same with the COO arrays already in device memory.  The buffers stay owned by the caller, are never copied and must
 * outlive the IRM: in an HIRM one copy of a relation's observations serves every IRM the relation visits
 * (HIRM::transition_cluster_assignment_relation re-incorporates the data into each candidate, cxx/hirm.hh:859-864). */
int hirm_irm_set_observations_device(hirm_engine *e, int32_t irm, int rel, int64_t nobs, const int32_t *d_ids,
                                     const uint8_t *d_val);
/* Many unary relations over one domain as a relation-owned block of entity-major bit rows (observed, value): row i holds
 * entity i's observation in every relation of the block -- Relation::data_r of a unary relation is "entity -> its one
 * observation" (cxx/hirm.hh:251), 2 bits here instead of a CSR entry.  Built once on the device from the relations' COO
 * arrays (h_d_ids[c] / h_d_val[c]: DEVICE pointers, caller-owned, must outlive the block), shared by every IRM: moving a
 * relation between the IRMs of an HIRM (cxx/hirm.hh:859-864 re-incorporates its data) changes a column mask, not data. */
int hirm_unary_block_create(hirm_engine *e, int32_t n_entities, int n_rel, const int64_t *h_nobs,
                            const int32_t *const *h_d_ids, const uint8_t *const *h_d_val, int32_t *out_block);
int hirm_unary_block_destroy(hirm_engine *e, int32_t block);
/* IRM::incorporate of every observation of unary relation `rel` = column `col` of the block (instead of
 * hirm_irm_set_observations*).  All unary-block relations of one domain of an IRM must come from the same block. */
int hirm_irm_set_observations_unary(hirm_engine *e, int32_t irm, int rel, int32_t block, int col);
/* Dense binary relation R(d0,d1) as two byte slabs in device memory (0, 1, HIRM_DENSE_MISSING):
 *   d_slab0[i*pitch0 + j] = value of (i, j)   i in d0, j in d1   ("rows by first entity")
 *   d_slab1[j*pitch1 + i] = value of (i, j)                      ("rows by second entity")
 * pitches in bytes, multiples of 16; buffers stay owned by the caller and may be shared by many IRMs. */
int hirm_irm_set_observations_dense(hirm_engine *e, int32_t irm, int rel,
                                    const uint8_t *d_slab0, int64_t pitch0,
                                    const uint8_t *d_slab1, int64_t pitch1);
/* Same, from a host matrix h_x[i*h_pitch + j]; the engine uploads it and builds the transpose on the
 * device (engine-owned buffers). */
int hirm_irm_set_observations_dense_host(hirm_engine *e, int32_t irm, int rel, const uint8_t *h_x, int64_t h_pitch);
/* Make irm_dst use irm_src's observation store (same schema; independent chains over one dataset). */
int hirm_irm_share_observations(hirm_engine *e, int32_t irm_dst, int32_t irm_src);
/* Storage of Relation::clusters (cxx/hirm.hh:246; a sparse map with create-on-demand :540-543 and erase-on-zero :534-537)
 * for a CSR-stored relation: HIRM_TABLE_DENSE = a dense array over the index space prod(max_tables) (at most 2^27 cells),
 * HIRM_TABLE_HASHED = an open-addressing table of the non-empty cells keyed by the cluster tuple (64-bit CAS insert, 64-bit
 * add of (N | s << 32); an emptied cell reads as absent and is dropped when the engine rebuilds the table; the table grows
 * on demand), HIRM_TABLE_AUTO (default) = hashed iff the index space exceeds the dense budget.  Results are identical
 * either way.  Call before hirm_irm_finalize. */
#define HIRM_TABLE_AUTO 0
#define HIRM_TABLE_DENSE 1
#define HIRM_TABLE_HASHED 2
int hirm_irm_set_table_mode(hirm_engine *e, int32_t irm, int rel, int mode);
int hirm_irm_finalize(hirm_engine *e, int32_t irm);

/* Domain::incorporate(item, table) for every entity   (cxx/hirm.hh:194-203; forced seats as in
 * from_txt cxx/util_io.cc:328-337).  h_labels [n_entities[dom]], -1 = entity not in the domain.
 * The count tables (Relation::clusters, cxx/hirm.hh:246) are rebuilt from (observations, z) on the
 * device at the next use: they are a histogram (cxx/hirm.hh:281-289). */
int hirm_irm_set_assignments(hirm_engine *e, int32_t irm, int dom, const int32_t *h_labels);
/* Domain::incorporate for every entity of every domain of the listed IRMs, seats drawn on the device: one sequential
 * CRP(alpha) prior draw per (irm, domain) (cxx/hirm.hh:194-203 + CRP::sample :101-113), Philox stream = (the irm's stream
 * (hirm_irm_set_rng) << 3) | domain.  The initial state the reference's incorporate loop gives a new model. */
int hirm_engine_seat_prior(hirm_engine *e, int n_irms, const int32_t *h_irms, uint64_t seed);
/* CRP::assignments read-back (cxx/hirm.hh:75) */
int hirm_irm_get_assignments(hirm_engine *e, int32_t irm, int dom, int32_t *h_labels);
/* CRP::alpha (cxx/hirm.hh:72); set by the host after CRP::transition_alpha (:162-173) */
int hirm_irm_set_alpha(hirm_engine *e, int32_t irm, int dom, double alpha);
/* Relation::clusters read-back (cxx/hirm.hh:246): non-empty cells, lexicographically sorted,
 * (arity + 2) int32 per cell: table labels..., N, s.  *n_cells gets the count even if > cap. */
int hirm_irm_get_counts(hirm_engine *e, int32_t irm, int rel, int32_t *h_cells, int64_t cap_cells, int64_t *n_cells);
/* IRM::logp_score   (cxx/hirm.hh:730-740) */
int hirm_irm_logp_score(hirm_engine *e, int32_t irm, double *out);

/* The hot path.  For each listed IRM, in order, the moves
 *     IRM::transition_cluster_assignment_item(d, item)     (cxx/hirm.hh:606-641; src/hirm.py:401-421)
 * i.e. CRP::tables_weights_gibbs (:147) + Relation::logp_gibbs_exact (:441) + log_choice
 * (cxx/util_math.cc:58) + Relation::set_cluster_assignment_gibbs (:524) + Domain::set_cluster_assignment_gibbs (:219),
 * in the explicit order given (the sweep loops cxx/hirm.hh:590-604, cxx/hirm.cc:45-51,73-81).
 * IRMs are independent and run concurrently; moves of one IRM are sequential.
 *   h_move_offsets [n_irms+1]  moves of irm k are [off[k], off[k+1])
 *   h_move_dom / h_move_item   per move
 *   h_uniforms     nullable; per move uniform in [0,1) (parity: "same uniforms => same assignment")
 *                  when NULL move m of irm k uses Philox4x32-10(key=seed, ctr=(h_counter0[k]+m, h_stream[k]))
 *   h_out_choice   nullable; per move chosen table label
 *   trace_cap>0:   per move candidate labels/log-probabilities (ascending label order):
 *                  h_trace_n [n_moves], h_trace_labels/logps [n_moves*trace_cap]
 */
int hirm_engine_sweep(hirm_engine *e, int n_irms, const int32_t *h_irms, const int64_t *h_move_offsets,
                      const int32_t *h_move_dom, const int32_t *h_move_item, const double *h_uniforms,
                      uint64_t seed, const uint64_t *h_stream, const uint64_t *h_counter0, uint32_t flags,
                      int32_t *h_out_choice, int trace_cap, int32_t *h_trace_n, int32_t *h_trace_labels,
                      double *h_trace_logps);

/* Same launch with every array already resident on the device and no host synchronisation:
 * the call bench.py times for the HBM-resident figure.  d_out_choice nullable. */
int hirm_engine_sweep_device(hirm_engine *e, int n_irms, const int32_t *h_irms, const int64_t *d_move_offsets,
                             const int32_t *d_move_dom, const int32_t *d_move_item, uint64_t seed,
                             const uint64_t *d_stream, const uint64_t *d_counter0, uint32_t flags,
                             int32_t *d_out_choice);

/* IRM::transition_cluster_assignments_all (cxx/hirm.hh:590-596; src/hirm.py:392-400; the item loops of
 * cxx/hirm.cc:45-51,73-81), n_sweeps times for every listed IRM.  The sweep order is engine-defined and generated
 * on the device: per sweep every domain in index order, its entities in the order j -> (a j + b) mod n with (a, b)
 * drawn per (irm, sweep, domain) from Philox (the reference's order is a hash-container artefact).  Uniforms:
 * Philox4x32-10(key = seed, ctr = (irm's move counter + m, irm's stream)), see hirm_irm_set_rng.
 * Outputs nullable: the generated moves and chosen labels (cap_moves entries each); *n_moves = moves made. */
int hirm_engine_sweep_all(hirm_engine *e, int n_irms, const int32_t *h_irms, int n_sweeps, uint64_t seed, uint32_t flags,
                          int64_t cap_moves, int32_t *h_out_dom, int32_t *h_out_item, int32_t *h_out_choice, int64_t *n_moves);
/* Philox stream id (default: the irm id) and move counter used by hirm_engine_sweep_all / _transition_alpha */
int hirm_irm_set_rng(hirm_engine *e, int32_t irm, uint64_t stream, uint64_t move_counter);
/* the two other Philox positions of an IRM: sweeps made by hirm_engine_sweep_all (keys the sweep order) and calls of
 * hirm_engine_transition_alpha; set when an IRM is re-created or moved to another GPU in the middle of a chain */
int hirm_irm_set_rng_counters(hirm_engine *e, int32_t irm, uint64_t sweep_counter, uint64_t alpha_counter);

/* IRM::logp_score for many IRMs in one launch   (cxx/hirm.hh:730-740) */
int hirm_engine_logp_scores(hirm_engine *e, int n_irms, const int32_t *h_irms, double *h_out);

/* CRP::transition_alpha for every domain of every listed IRM   (cxx/hirm.hh:162-173: 20-point grid; src/hirm.py:109-116:
 * 30 points).  Grid = log_linspace(1/N, N+1, ngrid) (cxx/util_math.cc:17-32), index drawn with the engine's inversion
 * rule from h_uniforms[k * HIRM_MAX_DOMAINS + d] or, when NULL, Philox.  h_out_alpha [n_irms * HIRM_MAX_DOMAINS] and
 * h_trace_lp [n_irms * HIRM_MAX_DOMAINS * ngrid] (grid log-probabilities) are nullable. */
int hirm_engine_transition_alpha(hirm_engine *e, int n_irms, const int32_t *h_irms, int ngrid, uint64_t seed,
                                 const double *h_uniforms, double *h_out_alpha, double *h_trace_lp);
int hirm_irm_get_alpha(hirm_engine *e, int32_t irm, int dom, double *out);
/* CRP::tables read-back (cxx/hirm.hh:74): labels ascending and customers per table */
int hirm_irm_get_tables(hirm_engine *e, int32_t irm, int dom, int32_t *h_labels, int32_t *h_counts, int cap, int32_t *n_tables);

/* Relation::logp_score of relation (src_irm, rel) under the domain partitions of other IRMs: the data term of every
 * candidate table in HIRM::transition_cluster_assignment_relation (cxx/hirm.hh:859-866; src/hirm.py:497-505), without
 * moving any data (the count table under a partition is a histogram).  h_cand_doms [n_cand * HIRM_MAX_ARITY]: for
 * each candidate, the index IN THE CANDIDATE of the domain at each position of the relation; the candidate's domains
 * must use the same entity codes.  Entities of the relation that are absent from a candidate's domain must be seated
 * first (hirm_irm_seat_entities), as the reference's trial incorporate does. */
int hirm_engine_relation_scores(hirm_engine *e, int32_t src_irm, int rel, int n_cand, const int32_t *h_cand_irms,
                                const int32_t *h_cand_doms, double *h_out);
/* One whole sweep of relation moves at once: Relation::logp_score (cxx/hirm.hh:516-522) of EVERY listed relation under the
 * domain partitions of EVERY listed IRM -- the quantity of cxx/hirm.hh:865 for all (relation, candidate table) pairs.  It
 * depends only on the relation's observations and the candidate's partitions, so the matrix can be computed (and, across
 * GPUs, all-gathered) before the sequential re-seating of the relations.  Relations are given as device COO arrays
 * (h_d_ids[r] / h_d_val[r] are DEVICE pointers, h_rel_doms [n_rel * HIRM_MAX_ARITY] the candidates' domain index per
 * position); the candidates must share domains, entity codes and max_tables, with every entity seated.
 * h_out [n_rel * n_cand], row-major; it may also be a DEVICE buffer (unified addressing): across GPUs the matrix is then
 * all-gathered (ncclAllGather) straight from there, without a round trip through the host. */
int hirm_engine_relation_score_matrix(hirm_engine *e, int n_rel, const int32_t *h_arity, const int32_t *h_rel_doms,
                                      const int64_t *h_nobs, const int32_t *const *h_d_ids, const uint8_t *const *h_d_val,
                                      int n_cand, const int32_t *h_cand_irms, double *h_out);
/* Unary relations are scored from bit vectors built once per COO buffer pair (keyed by the addresses h_d_ids[r], h_d_val[r]).
 * Lifetime rule: such buffers must stay allocated and unchanged while the engine may score them -- call this after freeing
 * or rewriting one (an allocator may hand the same address to new data), and when an engine is reused for a new dataset. */
int hirm_engine_clear_relation_cache(hirm_engine *e);
/* The fresh-table candidate of every listed relation (cxx/hirm.hh:849-864: `new IRM({}, prng)`, whose domains are seated by
 * sequential CRP-prior draws when the relation is incorporated, :194-203): one independent CRP(alpha) draw per (relation,
 * domain the relation uses), Philox stream h_keys[r * n_domains + d]; h_out[r] = the relation's logp_score under its draw. */
int hirm_engine_relation_scores_fresh(hirm_engine *e, int n_rel, const int32_t *h_arity, const int32_t *h_rel_doms,
                                      const int64_t *h_nobs, const int32_t *const *h_d_ids, const uint8_t *const *h_d_val,
                                      int n_domains, const int32_t *h_n_entities, const int32_t *h_max_tables, double alpha,
                                      uint64_t seed, const uint64_t *h_keys, double *h_out);
/* The same draw again, to the host: sequential CRP(alpha) prior over n customers (Domain::incorporate for every entity of a
 * fresh domain, cxx/hirm.hh:194-203 + CRP::sample :101-113), labels in order of appearance, Philox stream `key`. */
int hirm_engine_crp_prior_draw(hirm_engine *e, int n, double alpha, int max_tables, uint64_t seed, uint64_t key, int32_t *h_labels);
/* Domain::incorporate(item, table) for entities not yet in the domain (cxx/hirm.hh:194-203) */
int hirm_irm_seat_entities(hirm_engine *e, int32_t irm, int dom, int n, const int32_t *h_items, const int32_t *h_labels);

/* statistics of the last sweep: kernel launches made, which kernel ran (0 generic, 1 dense) */
int hirm_engine_last_sweep_info(hirm_engine *e, int32_t *n_launches, int32_t *kernel_kind, float *kernel_ms);

/* CTAs per IRM (thread-block cluster size) of the last generic launch */
int hirm_engine_last_cluster_size(hirm_engine *e, int32_t *ctas_per_irm);

/* Rounds of the last sweep call and whether rows were grouped ahead of the moves.  Where every CSR relation of an IRM holds
 * a domain once, the groups of that domain's entities (get_cluster_to_items_list, cxx/hirm.hh:388-397) do not depend on the
 * domain's own partition, so for a run of moves of the domain the engine groups the rows of ALL its moves first -- a
 * bandwidth-bound kernel with the other domains' table slots in shared memory -- and the sequential pass reads the group
 * lists.  The move lists are then swept in rounds (piece r of every list: a pre-grouping launch + the sequential launch).
 * Chosen by the engine (many IRMs with long rows); HIRM_PREGROUP=0/1 in the environment forces it off / on wherever
 * possible, HIRM_PREGROUP_SEG caps the moves per piece.  Results are bit-identical either way. */
int hirm_engine_last_rounds(hirm_engine *e, int32_t *rounds, int32_t *pregrouped);

/* The host part of that (csrc/sweep_rounds.cc, no CUDA): cut n_lists move lists -- list k = moves [off[k], off[k+1]), run j of its
 * n_runs[k] runs of equal domain starting at run_start[k * runcap + j] and qualifying for pre-grouping iff run_ok[k * runcap + j] -- into
 * pieces: at most `seg` moves of one qualifying run (pre = 1), or everything between two such runs, merged (pre = 0).  n_runs[k] <= 0 or
 * > runcap: the list is one ordinary piece.  Pass 1 (lo == NULL) returns the number of rounds (most pieces of a list, at least 1);
 * pass 2 fills lo / hi / pre [round * n_lists + k] (rounds beyond a list's last piece: empty). */
int hirm_plan_rounds(int n_lists, const int64_t *off, const int32_t *n_runs, int runcap, const int64_t *run_start, const int32_t *run_ok,
                     int64_t seg, int32_t *n_rounds, int64_t *lo, int64_t *hi, int32_t *pre, int max_rounds);

/* Cluster sizes of a generic launch (host only): n IRMs on `budget` SMs, entries[i] = CSR entries (+ unary columns / 16) an average move of
 * IRM i reads, tables[i] = its live tables (NULL: 8 each).  Every IRM starts with one CTA; the IRM with the most work per CTA (entries x
 * max(1, tables / 8)) doubles its cluster, up to 16, while SMs are left -- unless `forced`, not below 192 units of work per CTA.  Returns
 * the CTAs in use. */
int hirm_plan_clusters(int n, const double *entries, const int32_t *tables, int budget, int forced, int32_t *ctas);

/* shared-memory plan of the last dense launch: resident CTAs (chains) per SM, TMA ring stages, bytes per
 * row chunk, dynamic shared memory per CTA */
int hirm_engine_last_dense_plan(hirm_engine *e, int32_t *ctas_per_sm, int32_t *stages, int32_t *chunk_bytes, int32_t *smem_bytes);

/* device helper: out[j*pitch_out + i] = in[i*pitch_in + j] for a rows x cols byte matrix */
int hirm_engine_transpose_u8(hirm_engine *e, const uint8_t *d_in, int64_t rows, int64_t cols, int64_t pitch_in,
                             uint8_t *d_out, int64_t pitch_out);

/* test hook: BetaBernoulli score difference S(N+n, s+sg) - S(N, s) (cxx/hirm.hh:52-56) for n operand quadruples,
 * evaluated by the sweep kernels' own fp64 code */
int hirm_engine_debug_score_delta(hirm_engine *e, int n, const uint32_t *h_N, const uint32_t *h_s, const uint32_t *h_gn,
                                  const uint32_t *h_gs, double *h_out);

/* debug hook: device buffer [n_blocks * 16] int64 that the dense kernel fills with SM cycles per role
 * (producer, accumulate warps, decision warps: see profiles/phase_breakdown.py); NULL switches it off */
int hirm_engine_debug_phase_cycles(hirm_engine *e, long long *d_buf);

/* ---- HIRM relation moves, host part (no CUDA device needed) ---------------------------------------------------------------
 * The sequential CRP re-seating of one sweep of relation moves (HIRM::transition_cluster_assignment_relation,
 * cxx/hirm.hh:833-906) from the [relations x candidate tables] matrix of data terms that
 * hirm_engine_relation_score_matrix / _relation_scores_fresh computed (and, across GPUs, one all-gather made known to every
 * rank): CRP::tables_weights_gibbs (:147-161) + log_choice with the engine's inversion rule, relation after relation.
 * labels[n_tables] ascending; sizes[n_tables] relations per table and rel_table[n_rel] are updated in place (a table whose
 * size reaches 0 is deleted, :886-890); scores[r * ld + j] = data term of relation r under table j, scores[r * ld +
 * fresh_col] under its fresh table.  Processing starts at order[*pos].  Returns 0 when every listed relation has moved;
 * 1 when the relation at order[*pos] chose the FRESH table *fresh_label (not applied: the caller makes that IRM and its
 * score column, then calls again with the new table and *pos + 1); < 0 on a bad argument. */
int hirm_relation_reseat(int n_rel, int n_tables, const int32_t *labels, int32_t *sizes, int32_t *rel_table,
                         const double *scores, int64_t ld, int fresh_col, const int32_t *order, const double *uniforms,
                         int n_order, double alpha, int32_t *pos, int32_t *n_moved, int32_t *fresh_label);

/* ---- ingest: the reference's text formats, natively (host only; no CUDA device needed) ------------------------------
 * load_schema (cxx/util_io.cc:11-34) + load_observations (:36-61) + encode_observations (:63-96): <base>.schema and
 * <base>.obs -> per relation a COO array of integer entity codes (per domain 0, 1, 2... in order of first appearance, lines
 * in file order, positions left to right) and a value array, ready for hirm_irm_set_observations.  Relations keep the
 * schema file's order, domains the order of their first appearance in it.  Pointers returned by the accessors stay
 * owned by the dataset (hirm_dataset_entity: until the next call). */
typedef struct hirm_dataset hirm_dataset;
const char *hirm_dataset_last_error(void);
int hirm_dataset_load(const char *schema_path, const char *obs_path, hirm_dataset **out);
int hirm_dataset_free(hirm_dataset *ds);
int hirm_dataset_counts(const hirm_dataset *ds, int32_t *n_domains, int32_t *n_relations);
int hirm_dataset_domain(const hirm_dataset *ds, int dom, const char **name, int32_t *n_entities);
int hirm_dataset_relation(const hirm_dataset *ds, int rel, const char **name, int32_t *arity, int32_t *doms, int64_t *nobs);
int hirm_dataset_observations(const hirm_dataset *ds, int rel, const int32_t **ids, const uint8_t **val);
int hirm_dataset_entity(hirm_dataset *ds, int dom, int32_t code, const char **name);

/* ---- the same reader on the GPU, for files of 1e8 lines and more (csrc/ingest_gpu.cu) -------------------------------------
 * The .obs file is copied to the device as it is; lines are split, tokens hashed, entity names numbered per domain in order
 * of first appearance (exactly the codes of cxx/util_io.cc:63-96 / hirm_dataset_load) and the COO arrays of every relation
 * written by streaming kernels.  The observations of a relation come out in an unspecified order (the engine's stores do
 * not depend on it).  Errors carry the 1-based line number, as the host reader's do: values other than 0 / 1, unknown
 * relation, wrong number of entities, an observation listed twice (cxx/hirm.hh:272).  The returned COO arrays are DEVICE
 * buffers owned by the dataset (valid until hirm_gpu_dataset_free); names are kept on the host for the writers. */
typedef struct hirm_gpu_dataset hirm_gpu_dataset;
const char *hirm_gpu_dataset_last_error(void);
int hirm_gpu_dataset_load(const char *schema_path, const char *obs_path, int device, hirm_gpu_dataset **out);
int hirm_gpu_dataset_free(hirm_gpu_dataset *ds);
int hirm_gpu_dataset_counts(const hirm_gpu_dataset *ds, int32_t *n_domains, int32_t *n_relations);
int hirm_gpu_dataset_domain(const hirm_gpu_dataset *ds, int dom, const char **name, int32_t *n_entities);
int hirm_gpu_dataset_relation(const hirm_gpu_dataset *ds, int rel, const char **name, int32_t *arity, int32_t *doms, int64_t *nobs,
                              const int32_t **d_ids, const uint8_t **d_val);
int hirm_gpu_dataset_entity(const hirm_gpu_dataset *ds, int dom, int32_t code, const char **name);
/* D2D copy of one relation's arrays into the caller's device buffers: ids [nobs * arity] int32, val [nobs] uint8 */
int hirm_gpu_dataset_copy_relation(const hirm_gpu_dataset *ds, int rel, int32_t *d_ids_out, uint8_t *d_val_out);
/* seconds spent: [0] host-to-device copy of the text, [1] line split + parse, [2] dictionaries, [3] COO arrays + duplicate check */
int hirm_gpu_dataset_timings(const hirm_gpu_dataset *ds, double *seconds4);

#ifdef __cplusplus
}
#endif
#endif
