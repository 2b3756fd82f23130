/* TEST INFRASTRUCTURE ONLY -- see hirm_oracle.h.
 *
 * CPU restatement, in plain C, of the reference's entity move.  Every function
 * cites the reference lines it follows (paths relative to /root/reference).
 *
 * One deliberate, documented difference (DESIGN.md "Q1"):
 * for a relation that uses the moved entity's domain at more than one position
 * the reference's logp_gibbs_exact groups observations by their full current
 * cluster tuple but takes the candidate's target cell from the first group
 * member only (cxx/hirm.hh:421, guard assert commented out at :428), which is
 * order dependent and not the Gibbs conditional.  This oracle scores the exact
 * remove-then-score conditional (observations merged per *target cell*), which
 * coincides with the reference to rounding wherever the domain occurs once, and
 * with brute force over the reference's own Relation::logp_score everywhere
 * (both are checked in tests/test_oracle_golden.py against oracle/_ref traces).
 */
#include "hirm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MAXA HIRM_ORACLE_MAX_ARITY

typedef struct { int32_t z[MAXA]; } cellkey;
typedef struct { cellkey k; int32_t N, s; int used; } cell;

typedef struct {
    cell *slots; int64_t cap, count; int arity;
} celltable;

static uint64_t key_hash(const cellkey *k, int arity) {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (int p = 0; p < arity; p++) {
        h ^= (uint64_t)(uint32_t)k->z[p] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
        h *= 0xBF58476D1CE4E5B9ull;
    }
    return h ^ (h >> 31);
}
static int key_eq(const cellkey *a, const cellkey *b, int arity) {
    for (int p = 0; p < arity; p++) if (a->z[p] != b->z[p]) return 0;
    return 1;
}
static void ct_init(celltable *t, int arity, int64_t cap) {
    t->arity = arity; t->cap = cap; t->count = 0;
    t->slots = (cell *)calloc((size_t)cap, sizeof(cell));
}
static void ct_free(celltable *t) { free(t->slots); t->slots = NULL; }
static cell *ct_find(celltable *t, const cellkey *k, int create);
static void ct_grow(celltable *t) {
    celltable n; ct_init(&n, t->arity, t->cap * 2);
    for (int64_t i = 0; i < t->cap; i++) if (t->slots[i].used) {
        cell *c = ct_find(&n, &t->slots[i].k, 1);
        c->N = t->slots[i].N; c->s = t->slots[i].s;
    }
    ct_free(t); *t = n;
}
static cell *ct_find(celltable *t, const cellkey *k, int create) {
    if (create && (t->count + 1) * 2 > t->cap) ct_grow(t);
    uint64_t i = key_hash(k, t->arity) & (uint64_t)(t->cap - 1);
    for (;;) {
        cell *c = &t->slots[i];
        if (!c->used) {
            if (!create) return NULL;
            c->used = 1; c->k = *k; c->N = 0; c->s = 0; t->count++;
            return c;
        }
        if (key_eq(&c->k, k, t->arity)) return c;
        i = (i + 1) & (uint64_t)(t->cap - 1);
    }
}
static void ct_clear(celltable *t) { memset(t->slots, 0, (size_t)t->cap * sizeof(cell)); t->count = 0; }

typedef struct {
    int arity; int dom[MAXA];
    int64_t nobs; int32_t *ids; uint8_t *val;
    celltable counts;            /* Relation::clusters (cxx/hirm.hh:246): cell -> (N, s); N==0 means absent */
    /* Relation::data_r (cxx/hirm.hh:251): per distinct domain, per entity, the
       observations that mention it -- each observation once even if the entity
       sits at two positions (:274-280 insert into a set keyed by domain name) */
    int n_adj; int adj_dom[MAXA]; int64_t *adj_ptr[MAXA]; int64_t *adj_obs[MAXA];
} relation;

struct hirm_oracle {
    int n_domains; int32_t *n_entities;
    int n_relations; relation *rel;
    int32_t **z; double *alpha;
    /* CRP::tables (cxx/hirm.hh:74) as label -> size, small unsorted arrays */
    int *n_tab; int *cap_tab; int32_t **tab_label; int32_t **tab_count;
    celltable tmp;
};

hirm_oracle *hirm_oracle_create(int n_domains, const int32_t *n_entities, int n_relations,
                                const int32_t *arity, const int32_t *rel_domains) {
    hirm_oracle *o = (hirm_oracle *)calloc(1, sizeof(*o));
    o->n_domains = n_domains; o->n_relations = n_relations;
    o->n_entities = (int32_t *)malloc(sizeof(int32_t) * n_domains);
    o->z = (int32_t **)calloc(n_domains, sizeof(int32_t *));
    o->alpha = (double *)malloc(sizeof(double) * n_domains);
    o->n_tab = (int *)calloc(n_domains, sizeof(int)); o->cap_tab = (int *)calloc(n_domains, sizeof(int));
    o->tab_label = (int32_t **)calloc(n_domains, sizeof(int32_t *));
    o->tab_count = (int32_t **)calloc(n_domains, sizeof(int32_t *));
    for (int d = 0; d < n_domains; d++) {
        o->n_entities[d] = n_entities[d];
        o->z[d] = (int32_t *)malloc(sizeof(int32_t) * (n_entities[d] > 0 ? n_entities[d] : 1));
        for (int i = 0; i < n_entities[d]; i++) o->z[d][i] = -1;
        o->alpha[d] = 1.0;   /* CRP::alpha default (cxx/hirm.hh:72) */
    }
    o->rel = (relation *)calloc(n_relations, sizeof(relation));
    for (int r = 0; r < n_relations; r++) {
        relation *R = &o->rel[r];
        R->arity = arity[r];
        if (R->arity < 1 || R->arity > MAXA) { hirm_oracle_destroy(o); return NULL; }
        for (int p = 0; p < R->arity; p++) R->dom[p] = rel_domains[r * MAXA + p];
        ct_init(&R->counts, R->arity, 64);
    }
    ct_init(&o->tmp, MAXA, 64);
    return o;
}

void hirm_oracle_destroy(hirm_oracle *o) {
    if (!o) return;
    for (int r = 0; r < o->n_relations; r++) {
        relation *R = &o->rel[r];
        free(R->ids); free(R->val); ct_free(&R->counts);
        for (int a = 0; a < R->n_adj; a++) { free(R->adj_ptr[a]); free(R->adj_obs[a]); }
    }
    for (int d = 0; d < o->n_domains; d++) { free(o->z[d]); free(o->tab_label[d]); free(o->tab_count[d]); }
    free(o->rel); free(o->z); free(o->alpha); free(o->n_entities);
    free(o->n_tab); free(o->cap_tab); free(o->tab_label); free(o->tab_count);
    ct_free(&o->tmp); free(o);
}

int hirm_oracle_set_observations(hirm_oracle *o, int rel, int64_t nobs, const int32_t *ids, const uint8_t *val) {
    relation *R = &o->rel[rel];
    free(R->ids); free(R->val);
    for (int a = 0; a < R->n_adj; a++) { free(R->adj_ptr[a]); free(R->adj_obs[a]); }
    R->nobs = nobs;
    R->ids = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nobs * R->arity + 1));
    R->val = (uint8_t *)malloc((size_t)nobs + 1);
    memcpy(R->ids, ids, sizeof(int32_t) * (size_t)(nobs * R->arity));
    memcpy(R->val, val, (size_t)nobs);
    for (int64_t e = 0; e < nobs; e++) {
        if (val[e] > 1) return -1;                       /* BetaBernoulli::incorporate assert (cxx/hirm.hh:35) */
        for (int p = 0; p < R->arity; p++) {
            int32_t id = ids[e * R->arity + p];
            if (id < 0 || id >= o->n_entities[R->dom[p]]) return -2;
        }
    }
    /* distinct domains of the relation, in position order */
    R->n_adj = 0;
    for (int p = 0; p < R->arity; p++) {
        int seen = 0;
        for (int a = 0; a < R->n_adj; a++) if (R->adj_dom[a] == R->dom[p]) seen = 1;
        if (!seen) R->adj_dom[R->n_adj++] = R->dom[p];
    }
    for (int a = 0; a < R->n_adj; a++) {
        int d = R->adj_dom[a]; int n = o->n_entities[d];
        int64_t *ptr = (int64_t *)calloc((size_t)n + 1, sizeof(int64_t));
        for (int pass = 0; pass < 2; pass++) {
            int64_t *fill = NULL;
            if (pass == 1) {
                for (int i = 0; i < n; i++) ptr[i + 1] += ptr[i];
                R->adj_obs[a] = (int64_t *)malloc(sizeof(int64_t) * (size_t)(ptr[n] + 1));
                fill = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
                memcpy(fill, ptr, sizeof(int64_t) * (size_t)n);
            }
            for (int64_t e = 0; e < nobs; e++) {
                for (int p = 0; p < R->arity; p++) {
                    if (R->dom[p] != d) continue;
                    int32_t id = ids[e * R->arity + p];
                    int dup = 0;      /* same entity at an earlier position of the same domain */
                    for (int q = 0; q < p; q++) if (R->dom[q] == d && ids[e * R->arity + q] == id) dup = 1;
                    if (dup) continue;
                    if (pass == 0) ptr[id + 1]++; else R->adj_obs[a][fill[id]++] = e;
                }
            }
            free(fill);
        }
        R->adj_ptr[a] = ptr;
    }
    return 0;
}

/* Relation::get_cluster_assignment (cxx/hirm.hh:315-322) */
static void obs_key(const hirm_oracle *o, const relation *R, int64_t e, cellkey *k) {
    memset(k, 0, sizeof(*k));
    for (int p = 0; p < R->arity; p++) k->z[p] = o->z[R->dom[p]][R->ids[e * R->arity + p]];
}

/* The count table is a histogram of (observations, z): Relation::incorporate
   (cxx/hirm.hh:281-289) applied to every observation. */
static void rebuild_counts(hirm_oracle *o, relation *R) {
    ct_clear(&R->counts);
    for (int64_t e = 0; e < R->nobs; e++) {
        cellkey k; obs_key(o, R, e, &k);
        int absent = 0;
        for (int p = 0; p < R->arity; p++) if (k.z[p] < 0) absent = 1;
        if (absent) continue;
        cell *c = ct_find(&R->counts, &k, 1);
        c->N += 1; c->s += R->val[e];
    }
}

static int tab_index(hirm_oracle *o, int d, int32_t label) {
    for (int k = 0; k < o->n_tab[d]; k++) if (o->tab_label[d][k] == label) return k;
    return -1;
}
static void tab_add(hirm_oracle *o, int d, int32_t label, int delta) {
    int k = tab_index(o, d, label);
    if (k < 0) {
        if (o->n_tab[d] == o->cap_tab[d]) {
            o->cap_tab[d] = o->cap_tab[d] ? 2 * o->cap_tab[d] : 16;
            o->tab_label[d] = (int32_t *)realloc(o->tab_label[d], sizeof(int32_t) * o->cap_tab[d]);
            o->tab_count[d] = (int32_t *)realloc(o->tab_count[d], sizeof(int32_t) * o->cap_tab[d]);
        }
        k = o->n_tab[d]++;
        o->tab_label[d][k] = label; o->tab_count[d][k] = 0;
    }
    o->tab_count[d][k] += delta;
    if (o->tab_count[d][k] == 0) {      /* CRP::unincorporate erases an empty table (cxx/hirm.hh:95-97) */
        o->n_tab[d]--;
        o->tab_label[d][k] = o->tab_label[d][o->n_tab[d]];
        o->tab_count[d][k] = o->tab_count[d][o->n_tab[d]];
    }
}

int hirm_oracle_set_assignments(hirm_oracle *o, int dom, const int32_t *labels) {
    memcpy(o->z[dom], labels, sizeof(int32_t) * (size_t)o->n_entities[dom]);
    o->n_tab[dom] = 0;
    for (int i = 0; i < o->n_entities[dom]; i++) if (labels[i] >= 0) tab_add(o, dom, labels[i], 1);
    for (int r = 0; r < o->n_relations; r++) {
        int touches = 0;
        for (int p = 0; p < o->rel[r].arity; p++) if (o->rel[r].dom[p] == dom) touches = 1;
        if (touches && o->rel[r].nobs > 0) rebuild_counts(o, &o->rel[r]);
    }
    return 0;
}
int hirm_oracle_get_assignments(hirm_oracle *o, int dom, int32_t *labels) {
    memcpy(labels, o->z[dom], sizeof(int32_t) * (size_t)o->n_entities[dom]);
    return 0;
}
int hirm_oracle_set_alpha(hirm_oracle *o, int dom, double alpha) { o->alpha[dom] = alpha; return 0; }

/* BetaBernoulli::logp_score with alpha = beta = 1 (cxx/hirm.hh:52-56) through
   lbeta(int,int) = lgamma(z) + lgamma(w) - lgamma(z+w) (cxx/util_math.cc:13-15);
   lbeta(1,1) is exactly 0. */
static double cell_score(int32_t N, int32_t s) {
    return lgamma((double)s + 1.0) + lgamma((double)(N - s) + 1.0) - lgamma((double)N + 2.0);
}

static int cmp_i32(const void *a, const void *b) {
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}

/* CRP::tables_weights_gibbs (cxx/hirm.hh:147-161; src/hirm.py:101-108) with the
   candidates put in ascending label order (the engine's convention). */
static int candidates(hirm_oracle *o, int d, int item, int32_t *labels, double *logw, int cap) {
    int32_t cur = o->z[d][item];
    if (cur < 0) return -1;
    int K = o->n_tab[d];
    if (K + 1 > cap) return -2;
    int32_t tmax = 0;                               /* t_max starts at 0 (cxx/hirm.hh:139) */
    int n = 0, singleton = 0;
    for (int k = 0; k < K; k++) {
        int32_t lab = o->tab_label[d][k];
        if (lab > tmax) tmax = lab;
        if (lab == cur && o->tab_count[d][k] == 1) { singleton = 1; }
        labels[n++] = lab;
    }
    if (!singleton) labels[n++] = tmax + 1;         /* fresh table (cxx/hirm.hh:144); erased again for a singleton (:152-159) */
    qsort(labels, (size_t)n, sizeof(int32_t), cmp_i32);
    for (int k = 0; k < n; k++) {
        int idx = tab_index(o, d, labels[k]);
        double w;
        if (idx < 0) w = o->alpha[d];                                  /* the fresh table */
        else if (labels[k] == cur) w = singleton ? o->alpha[d] : (double)(o->tab_count[d][idx] - 1);
        else w = (double)o->tab_count[d][idx];
        logw[k] = log(w);                           /* cxx/hirm.hh:612-615 */
    }
    return n;
}

static int rel_adj_index(const relation *R, int d) {
    for (int a = 0; a < R->n_adj; a++) if (R->adj_dom[a] == d) return a;
    return -1;
}

/* add (+1) or remove (-1) the entity's observations at its *current* label:
   what Relation::set_cluster_assignment_gibbs does per observation
   (cxx/hirm.hh:528-548), split into its two halves */
static void apply_item(hirm_oracle *o, int d, int item, int sign) {
    for (int r = 0; r < o->n_relations; r++) {
        relation *R = &o->rel[r];
        int a = rel_adj_index(R, d);
        if (a < 0 || R->nobs == 0) continue;
        for (int64_t q = R->adj_ptr[a][item]; q < R->adj_ptr[a][item + 1]; q++) {
            int64_t e = R->adj_obs[a][q];
            cellkey k; obs_key(o, R, e, &k);
            cell *c = ct_find(&R->counts, &k, 1);
            c->N += sign; c->s += sign * (int)R->val[e];
        }
    }
}

int hirm_oracle_score(hirm_oracle *o, int d, int item, int32_t *out_labels, double *out_logps, int cap) {
    int n = candidates(o, d, item, out_labels, out_logps, cap);
    if (n < 0) return n;
    int32_t cur = o->z[d][item];
    /* remove the entity's observations: the "rest" counts every candidate is scored against
       (logp_gibbs_exact_current un-incorporates them, cxx/hirm.hh:404-408) */
    apply_item(o, d, item, -1);
    for (int t = 0; t < n; t++) {
        o->z[d][item] = out_labels[t];               /* get_cluster_assignment_gibbs (cxx/hirm.hh:324-339) */
        double lp = 0;
        for (int r = 0; r < o->n_relations; r++) {   /* has_observation gate (cxx/hirm.hh:618) */
            relation *R = &o->rel[r];
            int a = rel_adj_index(R, d);
            if (a < 0 || R->nobs == 0) continue;
            int64_t q0 = R->adj_ptr[a][item], q1 = R->adj_ptr[a][item + 1];
            if (q0 == q1) continue;
            ct_clear(&o->tmp); o->tmp.arity = R->arity;
            for (int64_t q = q0; q < q1; q++) {      /* group by target cell (cf. :388-397) */
                int64_t e = R->adj_obs[a][q];
                cellkey k; obs_key(o, R, e, &k);
                cell *g = ct_find(&o->tmp, &k, 1);
                g->N += 1; g->s += R->val[e];
            }
            double lr = 0;
            for (int64_t i = 0; i < o->tmp.cap; i++) {
                cell *g = &o->tmp.slots[i];
                if (!g->used) continue;
                cell *c = ct_find(&R->counts, &g->k, 0);
                int32_t N0 = c ? c->N : 0, s0 = c ? c->s : 0;      /* aux cell (0,0) if absent (:423-424) */
                lr += cell_score(N0 + g->N, s0 + g->s) - cell_score(N0, s0);   /* :426-438 */
            }
            lp += lr;
        }
        out_logps[t] += lp;                          /* cxx/hirm.hh:622-624 */
    }
    o->z[d][item] = cur;
    apply_item(o, d, item, +1);
    return n;
}

/* log_choice (cxx/util_math.cc:58-65) with the inversion convention of
   random.Random.choices (src/util_math.py:24-28): bisect_right(cum, u*total)
   capped at n-1.  SURVEY.md 8a-Q2: the engine fixes this one rule. */
int hirm_oracle_choose(const double *lp, int n, double u) {
    double m = lp[0];
    for (int i = 1; i < n; i++) if (lp[i] > m) m = lp[i];
    double s = 0;
    for (int i = 0; i < n; i++) s += exp(lp[i] - m);
    double Z = log(s) + m;                            /* logsumexp (cxx/util_math.cc:43-50) */
    double total = 0;
    for (int i = 0; i < n; i++) total += exp(lp[i] - Z);
    double x = u * total, c = 0;
    for (int i = 0; i < n; i++) {
        c += exp(lp[i] - Z);
        if (c > x) return i;
    }
    return n - 1;
}

int hirm_oracle_move(hirm_oracle *o, int d, int item, double u) {
    int32_t labels[4096]; double lp[4096];
    int n = hirm_oracle_score(o, d, item, labels, lp, 4096);
    if (n < 0) return n;
    int32_t choice = labels[hirm_oracle_choose(lp, n, u)];
    int32_t cur = o->z[d][item];
    if (choice != cur) {                              /* cxx/hirm.hh:632-640 */
        apply_item(o, d, item, -1);
        o->z[d][item] = choice;
        apply_item(o, d, item, +1);
        tab_add(o, d, cur, -1);
        tab_add(o, d, choice, +1);
    }
    return choice;
}

/* Philox4x32-10 (Salmon et al., SC'11); known-answer vectors in tests/test_oracle_golden.py */
void hirm_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; round++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
double hirm_oracle_uniform(uint64_t seed, uint64_t stream, uint64_t counter) {
    uint32_t ctr[4] = {(uint32_t)counter, (uint32_t)(counter >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, w[4];
    hirm_oracle_philox4x32_10(ctr, key, w);
    uint64_t hi = w[0] >> 5, lo = w[1] >> 6;          /* 27 + 26 = 53 bits */
    return (double)((hi << 26) | lo) * (1.0 / 9007199254740992.0);
}

int hirm_oracle_sweep(hirm_oracle *o, int64_t n_moves, const int32_t *move_dom, const int32_t *move_item,
                      const double *uniforms, uint64_t seed, uint64_t stream, uint64_t offset) {
    for (int64_t m = 0; m < n_moves; m++) {
        double u = uniforms ? uniforms[m] : hirm_oracle_uniform(seed, stream, offset + (uint64_t)m);
        int rc = hirm_oracle_move(o, move_dom[m], move_item[m], u);
        if (rc < 0) return rc;
    }
    return 0;
}

static int cmp_cells(const void *a, const void *b, void *arg) {
    int w = *(int *)arg;
    const int32_t *x = (const int32_t *)a, *y = (const int32_t *)b;
    for (int i = 0; i < w; i++) if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1;
    return 0;
}
static int g_w;
static int cmp_cells_g(const void *a, const void *b) { return cmp_cells(a, b, &g_w); }

int64_t hirm_oracle_get_counts(hirm_oracle *o, int rel, int32_t *out, int64_t cap_cells) {
    relation *R = &o->rel[rel];
    int w = R->arity + 2; int64_t n = 0;
    for (int64_t i = 0; i < R->counts.cap; i++) {
        cell *c = &R->counts.slots[i];
        if (!c->used || c->N == 0) continue;          /* emptied cells are erased (cxx/hirm.hh:534-537) */
        if (n < cap_cells) {
            for (int p = 0; p < R->arity; p++) out[n * w + p] = c->k.z[p];
            out[n * w + R->arity] = c->N; out[n * w + R->arity + 1] = c->s;
        }
        n++;
    }
    if (n <= cap_cells) { g_w = w; qsort(out, (size_t)n, sizeof(int32_t) * (size_t)w, cmp_cells_g); }
    return n;
}

/* IRM::logp_score (cxx/hirm.hh:730-740) = sum CRP::logp_score (:123-132) + sum Relation::logp_score (:516-522) */
double hirm_oracle_logp_score(hirm_oracle *o) {
    double crp = 0;
    for (int d = 0; d < o->n_domains; d++) {
        /* only the domains some relation uses: the reference's IRM holds no other (cxx/hirm.hh:742-782) */
        int used = 0;
        for (int r = 0; r < o->n_relations; r++)
            for (int p = 0; p < o->rel[r].arity; p++) used |= o->rel[r].dom[p] == d;
        if (!used) continue;
        int N = 0;
        for (int k = 0; k < o->n_tab[d]; k++) N += o->tab_count[d][k];
        if (N == 0) continue;
        double t2 = 0;
        for (int k = 0; k < o->n_tab[d]; k++) t2 += lgamma((double)o->tab_count[d][k]);
        crp += o->n_tab[d] * log(o->alpha[d]) + t2 + lgamma(o->alpha[d]) - lgamma(N + o->alpha[d]);
    }
    double rel = 0;
    for (int r = 0; r < o->n_relations; r++) {
        relation *R = &o->rel[r];
        for (int64_t i = 0; i < R->counts.cap; i++) {
            cell *c = &R->counts.slots[i];
            if (c->used && c->N > 0) rel += cell_score(c->N, c->s);
        }
    }
    return crp + rel;
}

/* CRP::transition_alpha (cxx/hirm.hh:162-173; 30 grid points in src/hirm.py:109-116): grid-Gibbs over
 * log_linspace(1/N, N+1, ngrid) (cxx/util_math.cc:17-32, endpoint included), each grid point scored with
 * CRP::logp_score (cxx/hirm.hh:123-132); the index is drawn with the shared inversion rule from the injected
 * uniform.  out_grid / out_lp (nullable) receive the grid and its log-probabilities.  Returns the new alpha
 * (which it also sets), or the unchanged alpha if the domain is empty (:163). */
double hirm_oracle_transition_alpha(hirm_oracle *o, int dom, int ngrid, double u, double *out_grid, double *out_lp) {
    int N = 0;
    for (int k = 0; k < o->n_tab[dom]; k++) N += o->tab_count[dom][k];
    if (N == 0 || ngrid < 2 || ngrid > 64) return o->alpha[dom];
    double t2 = 0;
    for (int k = 0; k < o->n_tab[dom]; k++) t2 += lgamma((double)o->tab_count[dom][k]);
    double grid[64], lp[64];
    const double lo = log(1.0 / N), hi = log((double)N + 1.0);
    const double step = (hi - lo) / (ngrid - 1);           /* linspace(start, stop, num, endpoint = true) */
    for (int g = 0; g < ngrid; g++) {
        grid[g] = exp(lo + step * g);
        lp[g] = o->n_tab[dom] * log(grid[g]) + t2 + lgamma(grid[g]) - lgamma(N + grid[g]);
        if (out_grid) out_grid[g] = grid[g];
        if (out_lp) out_lp[g] = lp[g];
    }
    o->alpha[dom] = grid[hirm_oracle_choose(lp, ngrid, u)];
    return o->alpha[dom];
}
