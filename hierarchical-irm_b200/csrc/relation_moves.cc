// Host side of the HIRM relation move: the sequential CRP re-seating of the relations of one sweep, from the gathered
// [relations x candidate tables] matrix of data terms (HIRM::transition_cluster_assignment_relation, cxx/hirm.hh:833-906,
// src/hirm.py:482-519).  The matrix comes from the device (hirm_engine_relation_score_matrix, all-gathered across GPUs);
// what is left is O(R K) scalar work that every rank repeats identically from shared uniforms -- here in C++ instead of a
// Python loop.  No CUDA in this file.
#include "../../include/hirm_engine.h"

#include <algorithm>
#include <cmath>
#include <vector>

// CRP::tables_weights_gibbs (cxx/hirm.hh:147-161) + the data terms, then the engine's inversion rule (first index whose
// running sum of exp(lp - max) exceeds u * total; cxx/util_math.cc:58-65 with Random.choices' convention).
//
//   n_tables, labels[n_tables] ascending, sizes[n_tables]  relations per table (updated in place; 0 = table deleted, :886-890)
//   rel_table[n_rel]     current table label of every relation (updated in place)
//   scores[n_rel * ld]   row r: data term of relation r under table j at column j (j < n_tables), fresh-table term at
//                        column fresh_col
//   order[n_order], uniforms[n_order]   the relations to move, in order, and their uniforms; processing starts at *pos
// The loop stops BEFORE applying a move to the fresh table (its partitions and its score column for the relations still to
// move have to be made first): returns 1 with *pos = index in `order` of that relation and *fresh_label = its new label;
// returns 0 when every relation has moved; < 0 on error.  *n_moved counts the relations that changed tables.
extern "C" int hirm_relation_reseat(int n_rel, int n_tables, const int32_t *labels, int32_t *sizes, int32_t *rel_table,
                                    const double *scores, int64_t ld, int fresh_col, const int32_t *order, const double *uniforms,
                                    int n_order, double alpha, int32_t *pos, int32_t *n_moved, int32_t *fresh_label) {
    if (n_rel < 0 || n_tables < 1 || !labels || !sizes || !rel_table || !scores || !order || !uniforms || !pos || !n_moved || !fresh_label)
        return -1;
    if (!(alpha > 0) || fresh_col < n_tables || fresh_col >= ld) return -1;
    std::vector<double> lp((size_t)n_tables + 1);
    std::vector<int> cand((size_t)n_tables + 1);
    const double log_alpha = std::log(alpha);
    for (int k = *pos; k < n_order; k++) {
        const int r = order[k];
        if (r < 0 || r >= n_rel) return -1;
        const int32_t cur = rel_table[r];
        int jcur = -1, nc = 0, jmax = -1;
        for (int j = 0; j < n_tables; j++) {
            if (sizes[j] <= 0) continue;                        // deleted table
            if (labels[j] == cur) jcur = j;
            jmax = j;                                           // labels ascending: the last live one is the largest
            const int n = sizes[j] - (labels[j] == cur ? 1 : 0);
            lp[(size_t)nc] = (n > 0 ? std::log((double)n) : log_alpha) + scores[(size_t)r * ld + j];   // a singleton keeps its table at weight alpha (:152-159)
            cand[(size_t)nc++] = j;
        }
        if (jcur < 0) return -2;
        const bool offer_fresh = sizes[jcur] > 1;               // otherwise no fresh table is offered
        if (offer_fresh) { lp[(size_t)nc] = log_alpha + scores[(size_t)r * ld + fresh_col]; cand[(size_t)nc++] = -1; }
        double mx = lp[0];
        for (int i = 1; i < nc; i++) mx = std::max(mx, lp[(size_t)i]);
        double total = 0.0;
        for (int i = 0; i < nc; i++) { lp[(size_t)i] = std::exp(lp[(size_t)i] - mx); total += lp[(size_t)i]; }
        const double x = uniforms[k] * total;
        double cum = 0.0;
        int pick = nc - 1;
        for (int i = 0; i < nc; i++) { cum += lp[(size_t)i]; if (cum > x) { pick = i; break; } }
        const int j = cand[(size_t)pick];
        if (j < 0) { *pos = k; *fresh_label = labels[jmax] + 1; return 1; }
        if (j != jcur) {
            sizes[jcur]--; sizes[j]++;
            rel_table[r] = labels[j];
            (*n_moved)++;
        }
    }
    *pos = n_order;
    return 0;
}
