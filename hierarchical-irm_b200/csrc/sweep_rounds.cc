// Host side of the rows-grouped-ahead sweep (csrc/pregroup.cuh): cutting every IRM's move list into PIECES.  A piece is either
// a stretch of at most `seg` moves inside one run of a domain whose rows can be grouped ahead (pre = 1), or everything between
// two such runs, merged (pre = 0: the ordinary path, which handles any mix of domains).  Round r of a sweep call = piece r of
// every list.  No CUDA in this file: the logic is covered on the CPU (tests/test_sweep_rounds.py).
#include "../../include/hirm_engine.h"

#include <algorithm>
#include <vector>

// n_lists move lists, list k = moves [off[k], off[k+1]).  n_runs[k] runs of equal domain (<= 0 or > runcap: unknown -- the list
// becomes one ordinary piece), run j of list k starts at run_start[k * runcap + j] and qualifies iff run_ok[k * runcap + j] != 0.
// Pass 1 (lo == NULL): *n_rounds = most pieces of any list (at least 1).  Pass 2: lo / hi / pre [round * n_lists + k], rounds
// beyond a list's last piece empty (lo = hi = 0, pre = 0); max_rounds = capacity in rounds.  Returns 0, or -1 on bad arguments.
extern "C" int hirm_plan_rounds(int n_lists, const int64_t *off, const int32_t *n_runs, int runcap, const int64_t *run_start,
                                const int32_t *run_ok, int64_t seg, int32_t *n_rounds, int64_t *lo, int64_t *hi, int32_t *pre,
                                int max_rounds) {
    if (n_lists < 0 || !off || !n_runs || runcap < 1 || !run_start || !run_ok || seg < 1 || !n_rounds) return -1;
    if (lo && (!hi || !pre || max_rounds < 1)) return -1;
    if (lo) {
        std::fill(lo, lo + (size_t)max_rounds * n_lists, (int64_t)0);
        std::fill(hi, hi + (size_t)max_rounds * n_lists, (int64_t)0);
        std::fill(pre, pre + (size_t)max_rounds * n_lists, (int32_t)0);
    }
    struct Piece { int64_t lo, hi; int pre; };
    std::vector<Piece> P;
    int rounds = 1;
    for (int k = 0; k < n_lists; k++) {
        P.clear();
        const int64_t b = off[k], e = off[k + 1];
        if (e < b) return -1;
        if (b < e) {
            const int nr = n_runs[k];
            if (nr <= 0 || nr > runcap) P.push_back({b, e, 0});
            else {
                for (int j = 0; j < nr; j++) {
                    const int64_t st = run_start[(size_t)k * runcap + j], en = j + 1 < nr ? run_start[(size_t)k * runcap + j + 1] : e;
                    if (st < b || en > e || en <= st || (j == 0 && st != b)) return -1;
                    if (run_ok[(size_t)k * runcap + j]) {
                        for (int64_t a = st; a < en; a += seg) P.push_back({a, std::min(en, a + seg), 1});
                    } else if (!P.empty() && !P.back().pre) P.back().hi = en;
                    else P.push_back({st, en, 0});
                }
            }
        }
        rounds = std::max(rounds, (int)P.size());
        if (lo) {
            if ((int)P.size() > max_rounds) return -1;
            for (size_t r = 0; r < P.size(); r++) {
                lo[r * n_lists + k] = P[r].lo; hi[r * n_lists + k] = P[r].hi; pre[r * n_lists + k] = P[r].pre;
            }
        }
    }
    *n_rounds = rounds;
    return 0;
}

// CTAs per IRM (thread-block cluster sizes) of a generic launch of n IRMs on `budget` SMs.  Work of a move ~ the entries it reads
// (entries[i]: CSR entries + unary columns / 16 of an average move) x the live tables it scores them against (tables[i], or NULL:
// 8 each): config 5's IRMs swept alone on 16 CTAs cost 25 us per move with 6000 entries and 9 tables, 42 us with 3400 entries and
// 47 tables (profiles/r02_c5_irm_costs.txt).  The launch ends with its slowest IRM: every IRM starts with one CTA, and the IRM with
// the most work per CTA doubles its cluster (up to 16, a power of two) while SMs are left; unless `forced`, not below 192 units of
// work per CTA (too little to be worth a cluster barrier).  Returns the CTAs in use, or -1 on bad arguments.
extern "C" int hirm_plan_clusters(int n, const double *entries, const int32_t *tables, int budget, int forced, int32_t *ctas) {
    if (n < 0 || (n > 0 && (!entries || !ctas))) return -1;
    auto work = [&](int i) { return entries[i] * std::max(1.0, (tables ? (double)tables[i] : 8.0) / 8.0); };
    for (int i = 0; i < n; i++) ctas[i] = 1;
    int total = n;
    for (;;) {
        int best = -1;
        for (int i = 0; i < n; i++) {
            if (ctas[i] >= 16 || total + ctas[i] > budget) continue;
            if (!forced && work(i) / ctas[i] < 192.0) continue;
            if (best < 0 || work(i) / ctas[i] > work(best) / ctas[best]) best = i;
        }
        if (best < 0) break;
        total += ctas[best]; ctas[best] *= 2;
    }
    return total;
}
