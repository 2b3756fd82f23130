// TEST / BASELINE INFRASTRUCTURE ONLY -- never linked into the product.
//
// Times the UNMODIFIED reference (headers compiled in place from /root/reference/cxx by
// oracle/Makefile) on the hot path itself: IRM::transition_cluster_assignment_item
// (cxx/hirm.hh:606-641), reached through the reference's public API exactly as its CLI
// does (IRM::IRM, IRM::incorporate, then the item loop of cxx/hirm.cc:45-51).  The
// dataset is BASELINE.json config 3's generator (planted two-block 'bernoulli R e e',
// dense) at a size the reference can hold: it needs ~420 B per observation, so the
// full n = 20000 (4e8 cells, ~168 GB, >0.2 s per move) is out of reach.
//
// usage: ref_bench <n> <seed> <rounds> <moves_per_round> [max_seconds_per_round]
// prints one line per round: "round <r> moves <m> seconds <t> tables <K>"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

#include "globals.hh"
#include "hirm.hh"

int main(int argc, char **argv) {
    if (argc < 5) { fprintf(stderr, "usage: ref_bench n seed rounds moves_per_round [max_s]\n"); return 2; }
    const int n = atoi(argv[1]), seed = atoi(argv[2]), rounds = atoi(argv[3]), per = atoi(argv[4]);
    const double max_s = argc > 5 ? atof(argv[5]) : 1e9;
    PRNG prng(seed);
    std::mt19937_64 data_rng(12345);          // same data for every seed (chains differ, data does not)
    std::uniform_real_distribution<double> U(0, 1);
    T_schema schema;
    schema["R"] = {"e", "e"};
    IRM irm(schema, &prng);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            bool zi = i >= n / 2, zj = j >= n / 2;
            double p = (zi != zj) ? 0.9 : 0.1;
            irm.incorporate("R", {i, j}, U(data_rng) < p ? 1.0 : 0.0);
        }
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    std::shuffle(order.begin(), order.end(), prng);
    size_t pos = 0;
    for (int r = 0; r < rounds; r++) {
        auto t0 = std::chrono::steady_clock::now();
        int done = 0;
        double el = 0;
        for (; done < per; done++) {
            irm.transition_cluster_assignment_item("e", order[pos]);
            pos = (pos + 1) % order.size();
            el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (el > max_s) { done++; break; }
        }
        printf("round %d moves %d seconds %.6f tables %zu\n", r, done, el, irm.domains.at("e")->crp.tables.size());
        fflush(stdout);
    }
    return 0;
}
