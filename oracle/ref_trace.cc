// TEST INFRASTRUCTURE ONLY -- never linked into the product.
//
// Trace harness around the UNMODIFIED reference headers (compiled in place from
// /root/reference/cxx by oracle/Makefile; output binary lives in oracle/_ref/).
// It drives the reference's own public members
//   Domain::tables_weights_gibbs          (cxx/hirm.hh:228)
//   Relation::logp_gibbs_exact            (cxx/hirm.hh:441)
//   Relation::set_cluster_assignment_gibbs(cxx/hirm.hh:524)
//   Domain::set_cluster_assignment_gibbs  (cxx/hirm.hh:219)
//   Relation::logp_score                  (cxx/hirm.hh:516)
// one entity move at a time, in an explicit (domain, entity) order, and records
// for every move: the candidate table labels, the reference's candidate
// log-probabilities, brute-force log-probabilities (really re-seat the entity
// and difference the reference's logp_score -- SURVEY.md 8a-Q1 / 8g), a uniform
// variate and the table chosen from it with the engine's inversion rule
// (SURVEY.md 8a-Q2).  State snapshots (assignments, count tables, logp_score)
// are taken at the start and after every sweep.  The JSON it prints is what
// tests/golden/*.json hold; tests/golden/make_golden.sh regenerates them.
//
// usage: ref_trace irm  <path_base> <seed> <sweeps> [shuffle]
//        ref_trace hirm <path_base> <seed> <burn_iters>
//
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <iostream>
#include <map>
#include <random>
#include <string>
#include <vector>

#include "globals.hh"
#include "hirm.hh"
#include "util_io.hh"
#include "util_math.hh"

using std::map;
using std::string;
using std::vector;

static bool g_first_move = true;

static void jdoubles(const vector<double> &v) {
    printf("[");
    for (size_t i = 0; i < v.size(); i++) printf("%s%.17g", i ? "," : "", v[i]);
    printf("]");
}
static void jints(const vector<int> &v) {
    printf("[");
    for (size_t i = 0; i < v.size(); i++) printf("%s%d", i ? "," : "", v[i]);
    printf("]");
}

// The engine's sampling convention (DESIGN.md "sampling rule"): candidates in
// ascending label order, p = exp(lp - logsumexp(lp)), first index whose running
// sum exceeds u * sum(p), capped at n-1.
static int choose(const vector<double> &lp, double u) {
    double m = *std::max_element(lp.begin(), lp.end());
    double s = 0;
    for (double x : lp) s += exp(x - m);
    double Z = log(s) + m;
    vector<double> p(lp.size());
    double total = 0;
    for (size_t i = 0; i < lp.size(); i++) { p[i] = exp(lp[i] - Z); total += p[i]; }
    double x = u * total, c = 0;
    for (size_t i = 0; i < lp.size(); i++) {
        c += p[i];
        if (c > x) return (int)i;
    }
    return (int)lp.size() - 1;
}

static void reseat(IRM &irm, const string &d, int item, int table) {
    Domain *domain = irm.domains.at(d);
    if (domain->get_cluster_assignment(item) == table) return;
    for (const auto &r : irm.domain_to_relations.at(d)) {
        auto relation = irm.relations.at(r);
        if (relation->has_observation(*domain, item)) {
            relation->set_cluster_assignment_gibbs(*domain, item, table);
        }
    }
    domain->set_cluster_assignment_gibbs(item, table);
}

static double relation_scores(IRM &irm, const string &d) {
    // sorted order so that the floating-point sum is reproducible
    vector<string> rs(irm.domain_to_relations.at(d).begin(), irm.domain_to_relations.at(d).end());
    std::sort(rs.begin(), rs.end());
    double s = 0;
    for (const auto &r : rs) s += irm.relations.at(r)->logp_score();
    return s;
}

static void dump_state(IRM &irm, const vector<string> &dnames, const vector<string> &rnames) {
    printf("{\"z\":{");
    for (size_t k = 0; k < dnames.size(); k++) {
        Domain *dom = irm.domains.at(dnames[k]);
        int n = 0;
        for (auto it : dom->items) n = std::max(n, it + 1);
        vector<int> z(n, -1);
        for (auto it : dom->items) z[it] = dom->get_cluster_assignment(it);
        printf("%s\"%s\":", k ? "," : "", dnames[k].c_str());
        jints(z);
    }
    printf("},\"alpha\":{");
    for (size_t k = 0; k < dnames.size(); k++)
        printf("%s\"%s\":%.17g", k ? "," : "", dnames[k].c_str(), irm.domains.at(dnames[k])->crp.alpha);
    printf("},\"counts\":{");
    for (size_t k = 0; k < rnames.size(); k++) {
        Relation *rel = irm.relations.at(rnames[k]);
        vector<vector<int>> cells;
        for (const auto &[z, cl] : rel->clusters) {
            vector<int> c(z.begin(), z.end());
            c.push_back(cl->N);
            c.push_back(((BetaBernoulli *)cl)->s);
            cells.push_back(c);
        }
        std::sort(cells.begin(), cells.end());
        printf("%s\"%s\":[", k ? "," : "", rnames[k].c_str());
        for (size_t j = 0; j < cells.size(); j++) { if (j) printf(","); jints(cells[j]); }
        printf("]");
    }
    printf("},\"logp_score\":%.17g}", irm.logp_score());
}

// One traced move.  Returns nothing; mutates the reference model by forcing
// the table chosen with the engine's rule from the brute-force (exact) logps.
static void traced_move(IRM &irm, const string &d, int item, PRNG &urng) {
    Domain *domain = irm.domains.at(d);
    int cur = domain->get_cluster_assignment(item);
    auto crp_dist = domain->tables_weights_gibbs(item);
    vector<int> labels;
    for (const auto &[t, w] : crp_dist) labels.push_back(t);
    std::sort(labels.begin(), labels.end());
    vector<double> prior(labels.size());
    for (size_t k = 0; k < labels.size(); k++) prior[k] = log(crp_dist.at(labels[k]));

    // (1) the reference's own candidate log-probabilities (cxx/hirm.hh:608-626)
    vector<double> ref_lp = prior;
    bool repeated = false;
    {
        vector<string> rs(irm.domain_to_relations.at(d).begin(), irm.domain_to_relations.at(d).end());
        std::sort(rs.begin(), rs.end());
        for (const auto &r : rs) {
            auto relation = irm.relations.at(r);
            int occ = 0;
            for (auto dd : relation->domains) occ += (dd->name == d);
            if (relation->has_observation(*domain, item)) {
                if (occ > 1) repeated = true;
                auto lp = relation->logp_gibbs_exact(*domain, item, labels);
                for (size_t k = 0; k < labels.size(); k++) ref_lp[k] += lp[k];
            }
        }
    }
    // (2) brute force: re-seat for real, difference the reference's logp_score
    vector<double> bf_lp(labels.size());
    double base = relation_scores(irm, d);
    for (size_t k = 0; k < labels.size(); k++) {
        reseat(irm, d, item, labels[k]);
        bf_lp[k] = prior[k] + (relation_scores(irm, d) - base);
        reseat(irm, d, item, cur);
    }
    double u = std::generate_canonical<double, 53>(urng);
    int idx = choose(bf_lp, u);
    int choice = labels[idx];

    printf("%s{\"d\":\"%s\",\"i\":%d,\"cur\":%d,\"alpha\":%.17g,\"repeated\":%s,\"labels\":",
           g_first_move ? "" : ",\n", d.c_str(), item, cur, domain->crp.alpha, repeated ? "true" : "false");
    g_first_move = false;
    jints(labels);
    printf(",\"ref_lp\":"); jdoubles(ref_lp);
    printf(",\"bf_lp\":"); jdoubles(bf_lp);
    printf(",\"u\":%.17g,\"choice\":%d}", u, choice);

    reseat(irm, d, item, choice);
}

static void dump_irm_header(IRM &irm, vector<string> &dnames, vector<string> &rnames) {
    dnames.clear(); rnames.clear();
    for (const auto &[d, dom] : irm.domains) dnames.push_back(d);
    for (const auto &[r, rel] : irm.relations) rnames.push_back(r);
    std::sort(dnames.begin(), dnames.end());
    std::sort(rnames.begin(), rnames.end());
    printf("\"domains\":{");
    for (size_t k = 0; k < dnames.size(); k++) {
        int n = 0;
        for (auto it : irm.domains.at(dnames[k])->items) n = std::max(n, it + 1);
        printf("%s\"%s\":%d", k ? "," : "", dnames[k].c_str(), n);
    }
    printf("},\n\"relations\":[");
    for (size_t k = 0; k < rnames.size(); k++) {
        Relation *rel = irm.relations.at(rnames[k]);
        printf("%s{\"name\":\"%s\",\"domains\":[", k ? "," : "", rnames[k].c_str());
        for (size_t j = 0; j < rel->domains.size(); j++)
            printf("%s\"%s\"", j ? "," : "", rel->domains[j]->name.c_str());
        printf("],\"obs\":[");
        vector<vector<int>> obs;
        for (const auto &[items, value] : rel->data) {
            vector<int> o(items.begin(), items.end());
            o.push_back((int)value);
            obs.push_back(o);
        }
        std::sort(obs.begin(), obs.end());
        for (size_t j = 0; j < obs.size(); j++) { if (j) printf(","); jints(obs[j]); }
        printf("]}");
    }
    printf("],\n");
}

static void traced_sweeps(IRM &irm, int sweeps, bool shuffle, bool alpha_moves, PRNG &urng) {
    vector<string> dnames, rnames;
    dump_irm_header(irm, dnames, rnames);
    printf("\"init\":"); dump_state(irm, dnames, rnames);
    printf(",\n\"moves\":[\n");
    g_first_move = true;
    int n_moves = 0;
    // state snapshots are emitted inline in the move list as {"ckpt": <moves so far>, ...}
    for (int s = 0; s < sweeps; s++) {
        for (const auto &d : dnames) {
            Domain *dom = irm.domains.at(d);
            vector<int> items(dom->items.begin(), dom->items.end());
            std::sort(items.begin(), items.end());
            if (shuffle) std::shuffle(items.begin(), items.end(), urng);
            for (int it : items) { traced_move(irm, d, it, urng); n_moves++; }
        }
        printf(",\n{\"ckpt\":%d,\"state\":", n_moves);
        dump_state(irm, dnames, rnames);
        printf("}");
        if (alpha_moves) {
            // the reference's alpha grid-Gibbs (cxx/hirm.hh:162-173) with its own grid (log_linspace,
            // cxx/util_math.cc:26) and its own CRP::logp_score (:123) per grid point; the index is drawn with
            // the engine's inversion rule from a traced uniform, so that later sweeps
            // exercise log(alpha) != 0 for the fresh-table candidate
            for (const auto &d : dnames) {
                CRP &crp = irm.domains.at(d)->crp;
                if (crp.N == 0) continue;
                auto grid = log_linspace(1. / crp.N, crp.N + 1, 20, true);
                vector<double> lps;
                for (const auto &g : grid) { crp.alpha = g; lps.push_back(crp.logp_score()); }
                double u = std::generate_canonical<double, 53>(urng);
                int idx = choose(lps, u);
                crp.alpha = grid[idx];
                printf(",\n{\"amove\":\"%s\",\"N\":%d,\"ngrid\":20,\"grid\":", d.c_str(), crp.N);
                jdoubles(grid);
                printf(",\"lp\":"); jdoubles(lps);
                printf(",\"u\":%.17g,\"idx\":%d,\"alpha\":%.17g}", u, idx, crp.alpha);
            }
        }
    }
    printf("\n]");
}


// ---- HIRM relation moves (cxx/hirm.hh:833-906), traced -------------------------------------------------------
// The body of HIRM::transition_cluster_assignment_relation restated around the reference's own calls
// (crp.tables_weights_gibbs, IRM::add_relation / incorporate, Relation::logp_score, IRM::remove_relation), with the
// table drawn by the engine's inversion rule from a traced uniform instead of log_choice(prng).  Dumps, per move,
// the candidate tables (ascending), the reference's log-probabilities, and -- for the fresh IRM -- the seats the
// reference's CRP-prior draws gave the relation's entities.
static void dump_hirm_state(HIRM &hirm) {
    vector<std::pair<string, int>> rt;
    for (const auto &[r, rc] : hirm.relation_to_code) rt.push_back({r, hirm.crp.assignments.at(rc)});
    std::sort(rt.begin(), rt.end());
    printf("{\"relation_table\":{");
    for (size_t k = 0; k < rt.size(); k++) printf("%s\"%s\":%d", k ? "," : "", rt[k].first.c_str(), rt[k].second);
    printf("},\"irm_z\":{");
    vector<int> tables;
    for (const auto &[t, irm] : hirm.irms) tables.push_back(t);
    std::sort(tables.begin(), tables.end());
    for (size_t k = 0; k < tables.size(); k++) {
        IRM *irm = hirm.irms.at(tables[k]);
        printf("%s\"%d\":{", k ? "," : "", tables[k]);
        vector<string> dn;
        for (const auto &[d, dom] : irm->domains) dn.push_back(d);
        std::sort(dn.begin(), dn.end());
        for (size_t j = 0; j < dn.size(); j++) {
            Domain *dom = irm->domains.at(dn[j]);
            vector<std::pair<int, int>> za;
            for (auto it : dom->items) za.push_back({it, dom->get_cluster_assignment(it)});
            std::sort(za.begin(), za.end());
            printf("%s\"%s\":{\"alpha\":%.17g,\"items\":[", j ? "," : "", dn[j].c_str(), dom->crp.alpha);
            for (size_t q = 0; q < za.size(); q++) printf("%s[%d,%d]", q ? "," : "", za[q].first, za[q].second);
            printf("]}");
        }
        printf("}");
    }
    printf("},\"logp_score\":%.17g}", hirm.logp_score());
}

static void traced_relation_moves(HIRM &hirm, PRNG &urng) {
    printf("\"hirm\":{\"init\":");
    dump_hirm_state(hirm);
    printf(",\n\"rmoves\":[\n");
    vector<string> rs;
    for (const auto &[r, rc] : hirm.relation_to_code) rs.push_back(r);
    std::sort(rs.begin(), rs.end());
    bool first = true;
    for (const auto &r : rs) {
        auto rc = hirm.relation_to_code.at(r);
        auto table_current = hirm.crp.assignments.at(rc);
        auto relation = hirm.get_relation(r);
        auto crp_dist = hirm.crp.tables_weights_gibbs(table_current);
        vector<string> domains;
        for (const auto &d : relation->domains) domains.push_back(d->name);
        vector<int> tables;
        for (const auto &[table, w] : crp_dist) tables.push_back(table);
        std::sort(tables.begin(), tables.end());
        vector<double> logps, lp_datas;
        int table_aux = -1; IRM *irm_aux = NULL;
        for (int table : tables) {
            IRM *irm;
            if (hirm.irms.count(table) == 0) { irm = new IRM({}, hirm.prng); table_aux = table; irm_aux = irm; }
            else irm = hirm.irms.at(table);
            if (table != table_current) {
                irm->add_relation(r, domains);
                for (const auto &[items, value] : relation->data) irm->incorporate(r, items, value);
            }
            double lp_data = irm->relations.at(r)->logp_score();
            lp_datas.push_back(lp_data);
            logps.push_back(log(crp_dist.at(table)) + lp_data);
        }
        double u = std::generate_canonical<double, 53>(urng);
        int choice = tables[choose(logps, u)];
        printf("%s{\"r\":\"%s\",\"cur\":%d,\"tables\":", first ? "" : ",\n", r.c_str(), table_current);
        first = false;
        jints(tables);
        printf(",\"lp\":"); jdoubles(logps);
        printf(",\"lp_data\":"); jdoubles(lp_datas);
        printf(",\"u\":%.17g,\"choice\":%d,\"fresh\":%d", u, choice, table_aux);
        if (irm_aux) {                                 // the seats the reference drew for the fresh IRM
            printf(",\"fresh_z\":{");
            bool fd = true;
            for (const auto &[d, dom] : irm_aux->domains) {
                vector<std::pair<int, int>> za;
                for (auto it : dom->items) za.push_back({it, dom->get_cluster_assignment(it)});
                std::sort(za.begin(), za.end());
                printf("%s\"%s\":[", fd ? "" : ",", d.c_str());
                fd = false;
                for (size_t q = 0; q < za.size(); q++) printf("%s[%d,%d]", q ? "," : "", za[q].first, za[q].second);
                printf("]");
            }
            printf("}");
        }
        printf("}");
        // the rest of cxx/hirm.hh:878-901, verbatim in effect
        for (const auto &[table, customers] : hirm.crp.tables) {
            auto irm = hirm.irms.at(table);
            if (table != choice) irm->remove_relation(r);
            if (irm->relations.size() == 0) { hirm.irms.erase(table); delete irm; }
        }
        if (irm_aux && choice == table_aux) hirm.irms[choice] = irm_aux; else delete irm_aux;
        hirm.crp.unincorporate(rc);
        hirm.crp.incorporate(rc, choice);
    }
    printf("\n],\n\"final\":");
    dump_hirm_state(hirm);
    printf("}");
}

int main(int argc, char **argv) {
    if (argc < 5) {
        fprintf(stderr, "usage: ref_trace irm|hirm <path_base> <seed> <n> [shuffle] [noalpha]\n");
        return 2;
    }
    string mode = argv[1], base = argv[2];
    int seed = atoi(argv[3]), n = atoi(argv[4]);
    bool shuffle = false, alpha_moves = true;
    for (int a = 5; a < argc; a++) {
        if (string(argv[a]) == "shuffle") shuffle = true;
        if (string(argv[a]) == "noalpha") alpha_moves = false;
    }
    PRNG prng(seed);
    PRNG urng(seed + 7919);
    auto schema = load_schema(base + ".schema");
    auto observations = load_observations(base + ".obs");
    auto encoding = encode_observations(schema, observations);

    if (mode == "irm") {
        IRM irm(schema, &prng);
        incorporate_observations(irm, encoding, observations);
        printf("{\"mode\":\"irm\",\"seed\":%d,\"irms\":[{\"table\":0,\n", seed);
        traced_sweeps(irm, n, shuffle, alpha_moves, urng);
        printf("}]}\n");
        return 0;
    }
    if (mode == "hirm") {
        HIRM hirm(schema, &prng);
        incorporate_observations(hirm, encoding, observations);
        // burn-in with the reference's own, unmodified inference loop body
        // (cxx/hirm.cc:61-90): relation moves, entity moves, alpha moves.
        for (int i = 0; i < n; i++) {
            for (const auto &[r, rc] : hirm.relation_to_code) hirm.transition_cluster_assignment_relation(r);
            for (const auto &[t, irm] : hirm.irms) {
                for (const auto &[d, domain] : irm->domains)
                    for (auto item : vector<int>(domain->items.begin(), domain->items.end()))
                        irm->transition_cluster_assignment_item(d, item);
                for (const auto &[d, domain] : irm->domains) domain->crp.transition_alpha();
            }
        }
        vector<int> tables;
        for (const auto &[t, irm] : hirm.irms) tables.push_back(t);
        std::sort(tables.begin(), tables.end());
        printf("{\"mode\":\"hirm\",\"seed\":%d,\"irms\":[", seed);
        for (size_t k = 0; k < tables.size(); k++) {
            printf("%s{\"table\":%d,\n", k ? ",\n" : "", tables[k]);
            traced_sweeps(*hirm.irms.at(tables[k]), 1, shuffle, false, urng);
            printf("}");
        }
        printf("],\n");
        traced_relation_moves(hirm, urng);
        printf("}\n");
        return 0;
    }
    fprintf(stderr, "unknown mode %s\n", mode.c_str());
    return 2;
}
