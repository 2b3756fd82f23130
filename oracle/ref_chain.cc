// TEST INFRASTRUCTURE ONLY -- never linked into the product.
//
// Chain driver around the UNMODIFIED reference headers (compiled in place from /root/reference/cxx by oracle/Makefile
// into oracle/_ref/ref_chain).  It builds an HIRM the way the reference's own main does (cxx/hirm.cc:140-181:
// load_schema, load_observations, encode_observations, HIRM(schema, &prng), incorporate_observations), runs the
// iteration of inference_hirm (cxx/hirm.cc:61-90: every relation's transition_cluster_assignment_relation, then every
// IRM's entity moves and CRP::transition_alpha) through the reference's public members, and prints
// HIRM::logp_score (cxx/hirm.hh:998-1005) and the number of IRMs at the end of every iteration.  (The reference CLI's
// own --verbose output mixes HIRM and per-IRM scores, cxx/hirm.cc:70,79,85, so it cannot be parsed for this.)
//
// usage: ref_chain <path_base> <seed> <iters>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "globals.hh"
#include "hirm.hh"
#include "util_io.hh"
#include "util_math.hh"

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: ref_chain <path_base> <seed> <iters>\n"); return 1; }
    const std::string base = argv[1];
    const int seed = atoi(argv[2]), iters = atoi(argv[3]);
    PRNG prng(seed);
    auto schema = load_schema(base + ".schema");
    auto observations = load_observations(base + ".obs");
    auto encoding = encode_observations(schema, observations);
    HIRM hirm(schema, &prng);
    incorporate_observations(hirm, encoding, observations);
    for (int i = 0; i < iters; i++) {
        for (const auto &[r, rc] : hirm.relation_to_code) hirm.transition_cluster_assignment_relation(r);
        for (const auto &[t, irm] : hirm.irms) {
            for (const auto &[d, domain] : irm->domains)
                for (auto item : domain->items) irm->transition_cluster_assignment_item(d, item);
            for (const auto &[d, domain] : irm->domains) domain->crp.transition_alpha();
        }
        printf("iter %d %.17g %d\n", i, hirm.logp_score(), (int)hirm.irms.size());
    }
    return 0;
}
