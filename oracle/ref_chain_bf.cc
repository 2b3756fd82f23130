// TEST INFRASTRUCTURE ONLY -- never linked into the product.
//
// IRM chains of the UNMODIFIED reference (headers compiled in place from /root/reference/cxx by oracle/Makefile into
// oracle/_ref/ref_chain_bf) with ONE substitution: the candidate log-probabilities of an entity move come from the brute
// force over the reference's own scoring -- really re-seat the entity with Relation::set_cluster_assignment_gibbs /
// Domain::set_cluster_assignment_gibbs (cxx/hirm.hh:524-551,219-224) and difference Relation::logp_score (:516-522) --
// instead of Relation::logp_gibbs_exact, whose target cell for a group of a REPEATED domain is taken from the group's
// first member only (:421).  Everything else is the reference's: the loop of inference_irm
// (cxx/hirm.cc:39-59), CRP::tables_weights_gibbs (:147-161), log_choice with its mt19937 (cxx/util_math.cc:58-65),
// CRP::transition_alpha (:162-173).  This is the chain the engine's exact conditional must match statistically on
// R(e, e) data (BASELINE config 3's shape).
//
// usage: ref_chain_bf <path_base> <seed> <iters>      prints "iter <i> <IRM::logp_score> <tables of the first domain>"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "globals.hh"
#include "hirm.hh"
#include "util_io.hh"
#include "util_math.hh"

static void reseat(IRM &irm, const std::string &d, int item, int table) {
    Domain *domain = irm.domains.at(d);
    if (domain->get_cluster_assignment(item) == table) return;
    for (const auto &r : irm.domain_to_relations.at(d)) {
        auto relation = irm.relations.at(r);
        if (relation->has_observation(*domain, item)) relation->set_cluster_assignment_gibbs(*domain, item, table);
    }
    domain->set_cluster_assignment_gibbs(item, table);
}
static double relation_scores(IRM &irm, const std::string &d) {
    double s = 0;
    for (const auto &r : irm.domain_to_relations.at(d)) s += irm.relations.at(r)->logp_score();
    return s;
}

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: ref_chain_bf <path_base> <seed> <iters>\n"); return 1; }
    const std::string base = argv[1];
    const int seed = atoi(argv[2]), iters = atoi(argv[3]);
    PRNG prng(seed);
    auto schema = load_schema(base + ".schema");
    auto observations = load_observations(base + ".obs");
    auto encoding = encode_observations(schema, observations);
    IRM irm(schema, &prng);
    incorporate_observations(irm, encoding, observations);
    for (int i = 0; i < iters; i++) {
        for (const auto &[d, domain] : irm.domains) {
            std::vector<int> items(domain->items.begin(), domain->items.end());   // the loop below re-seats: iterate a copy
            for (int item : items) {
                const int cur = domain->get_cluster_assignment(item);
                auto crp_dist = domain->tables_weights_gibbs(item);
                std::vector<int> tables;
                std::vector<double> logps;
                const double base_score = relation_scores(irm, d);
                for (const auto &[t, w] : crp_dist) {
                    tables.push_back(t);
                    reseat(irm, d, item, t);
                    logps.push_back(log(w) + relation_scores(irm, d) - base_score);
                    reseat(irm, d, item, cur);
                }
                const int choice = tables[log_choice(logps, &prng)];
                reseat(irm, d, item, choice);
            }
        }
        for (auto const &[d, domain] : irm.domains) domain->crp.transition_alpha();
        printf("iter %d %.17g %d\n", i, irm.logp_score(), (int)irm.domains.begin()->second->crp.tables.size());
    }
    return 0;
}
