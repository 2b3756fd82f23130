/* TEST INFRASTRUCTURE ONLY -- the CPU oracle.  Nothing in the product path may
 * include, link or call this (only tests/, __graft_entry__.smoke() and the
 * cpu_baseline leg of bench.py).
 *
 * Plain-C restatement of the reference's collapsed-Gibbs entity move
 *   IRM::transition_cluster_assignment_item      /root/reference/cxx/hirm.hh:606-641
 *                                                /root/reference/src/hirm.py:401-421
 * and the pieces it calls (see hirm_oracle.c for the per-function citations).
 * Pinned against the reference itself: tests/golden/*.json are traces produced by
 * oracle/_ref/ref_trace (the unmodified reference headers driven move by move).
 */
#ifndef HIRM_ORACLE_H
#define HIRM_ORACLE_H
#include <stdint.h>

#define HIRM_ORACLE_MAX_ARITY 4

typedef struct hirm_oracle hirm_oracle;

hirm_oracle *hirm_oracle_create(int n_domains, const int32_t *n_entities, int n_relations,
                                const int32_t *arity, const int32_t *rel_domains /* [n_relations*4] */);
void hirm_oracle_destroy(hirm_oracle *o);
/* ids: [nobs*arity] entity codes, val: [nobs] in {0,1} */
int hirm_oracle_set_observations(hirm_oracle *o, int rel, int64_t nobs, const int32_t *ids, const uint8_t *val);
/* labels: [n_entities[dom]] table labels (>=0), -1 = entity not in the domain */
int hirm_oracle_set_assignments(hirm_oracle *o, int dom, const int32_t *labels);
int hirm_oracle_get_assignments(hirm_oracle *o, int dom, int32_t *labels);
int hirm_oracle_set_alpha(hirm_oracle *o, int dom, double alpha);
/* candidate labels (ascending) and un-normalised log-probabilities; returns count or <0 */
int hirm_oracle_score(hirm_oracle *o, int dom, int item, int32_t *out_labels, double *out_logps, int cap);
/* inversion rule shared with the engine: returns the chosen index */
int hirm_oracle_choose(const double *logps, int n, double u);
/* one move with an injected uniform; returns the chosen label or <0 */
int hirm_oracle_move(hirm_oracle *o, int dom, int item, double u);
/* uniforms == NULL: u_m = philox(seed, stream, offset+m) */
int hirm_oracle_sweep(hirm_oracle *o, int64_t n_moves, const int32_t *move_dom, const int32_t *move_item,
                      const double *uniforms, uint64_t seed, uint64_t stream, uint64_t offset);
/* non-empty cells, lexicographically sorted, (arity + 2) int32 per cell: labels..., N, s */
int64_t hirm_oracle_get_counts(hirm_oracle *o, int rel, int32_t *out, int64_t cap_cells);
double hirm_oracle_logp_score(hirm_oracle *o);
/* CRP::transition_alpha with an injected uniform; returns (and sets) the new alpha */
double hirm_oracle_transition_alpha(hirm_oracle *o, int dom, int ngrid, double u, double *out_grid, double *out_lp);
void hirm_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double hirm_oracle_uniform(uint64_t seed, uint64_t stream, uint64_t counter);
#endif
