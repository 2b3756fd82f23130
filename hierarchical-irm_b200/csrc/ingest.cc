// Native ingest of the reference's text formats (host C++, no CUDA): <base>.schema and <base>.obs -> relation-owned COO
// arrays of integer entity codes, ready for hirm_irm_set_observations / hirm.sharded.  Replaces, for large files,
//   load_schema          cxx/util_io.cc:11-34     (one relation per line: <distribution> <relation> <domain>...)
//   load_observations    cxx/util_io.cc:36-61     (one observation per line: <value> <relation> <entity>...)
//   encode_observations  cxx/util_io.cc:63-96     (per domain, entities coded 0, 1, 2... in order of first appearance,
//                                                  lines in file order, positions left to right)
// The file is read once into memory; entity names are views into that buffer (no per-token allocation).
#include "../../include/hirm_engine.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <string_view>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace {
thread_local std::string g_ingest_err;
int ifail(const std::string &m) { g_ingest_err = m; return -1; }

bool read_file(const char *path, std::vector<char> &buf) {
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize((size_t)(n < 0 ? 0 : n) + 1);
    const size_t got = n > 0 ? fread(buf.data(), 1, (size_t)n, f) : 0;
    fclose(f);
    buf[got] = '\n';
    buf.resize(got + 1);
    return true;
}

// next whitespace-delimited token of the line [p, end); empty view at the end of the line
inline std::string_view next_token(const char *&p, const char *end) {
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\r')) p++;
    const char *b = p;
    while (p < end && *p != ' ' && *p != '\t' && *p != '\r') p++;
    return std::string_view(b, (size_t)(p - b));
}
}  // namespace

struct hirm_dataset {
    std::vector<char> schema_buf, obs_buf;
    std::vector<std::string> dom_names, rel_names;
    std::vector<std::vector<int32_t>> rel_doms;
    std::vector<std::unordered_map<std::string_view, int32_t>> codes;   // per domain: entity -> code
    std::vector<std::vector<std::string_view>> names;                   // per domain: code -> entity
    std::vector<std::vector<int32_t>> ids;                              // per relation: [nobs * arity]
    std::vector<std::vector<uint8_t>> val;                              // per relation: [nobs]
    std::string scratch;
};

extern "C" const char *hirm_dataset_last_error(void) { return g_ingest_err.c_str(); }

extern "C" int hirm_dataset_load(const char *schema_path, const char *obs_path, hirm_dataset **out) {
    if (!schema_path || !obs_path || !out) return ifail("null argument");
    auto ds = new hirm_dataset();
    std::unique_ptr<hirm_dataset> guard(ds);
    if (!read_file(schema_path, ds->schema_buf)) return ifail(std::string("cannot read ") + schema_path);
    if (!read_file(obs_path, ds->obs_buf)) return ifail(std::string("cannot read ") + obs_path);
    std::unordered_map<std::string_view, int> rel_index, dom_index;
    // ---- schema
    for (const char *p = ds->schema_buf.data(), *fe = p + ds->schema_buf.size(); p < fe;) {
        const char *le = (const char *)memchr(p, '\n', (size_t)(fe - p));
        if (!le) le = fe;
        const char *q = p;
        const std::string_view dist = next_token(q, le);
        if (!dist.empty()) {
            if (dist != "bernoulli") return ifail("only bernoulli relations are supported (README.md:122-123), got " + std::string(dist));
            const std::string_view rel = next_token(q, le);
            if (rel.empty()) return ifail("schema line without a relation name");
            if (rel_index.count(rel)) return ifail("relation listed twice in the schema: " + std::string(rel));
            std::vector<int32_t> doms;
            for (std::string_view d = next_token(q, le); !d.empty(); d = next_token(q, le)) {
                auto it = dom_index.find(d);
                if (it == dom_index.end()) {
                    it = dom_index.emplace(d, (int)ds->dom_names.size()).first;
                    ds->dom_names.emplace_back(d);
                }
                doms.push_back(it->second);
            }
            if (doms.empty() || (int)doms.size() > HIRM_MAX_ARITY) return ifail("a relation needs 1 to HIRM_MAX_ARITY domains: " + std::string(rel));
            rel_index.emplace(rel, (int)ds->rel_names.size());
            ds->rel_names.emplace_back(rel);
            ds->rel_doms.push_back(std::move(doms));
        }
        p = le + 1;
    }
    const size_t nd = ds->dom_names.size(), nr = ds->rel_names.size();
    ds->codes.resize(nd); ds->names.resize(nd); ds->ids.resize(nr); ds->val.resize(nr);
    // ---- observations
    // an observation may be listed once: Relation::incorporate asserts `data.count(items) == 0` (cxx/hirm.hh:272)
    std::vector<std::unordered_set<std::string>> seen(nr);
    int64_t line_no = 0;
    for (const char *p = ds->obs_buf.data(), *fe = p + ds->obs_buf.size(); p < fe;) {
        const char *le = (const char *)memchr(p, '\n', (size_t)(fe - p));
        if (!le) le = fe;
        line_no++;
        const char *q = p;
        const std::string_view v = next_token(q, le);
        if (!v.empty()) {
            int value;
            if (v == "1") value = 1;
            else if (v == "0") value = 0;
            else {                                       // `stream >> value` reads a double (cxx/util_io.cc:49)
                const std::string tmp(v);
                char *endp = nullptr;
                const double x = strtod(tmp.c_str(), &endp);
                if (endp == tmp.c_str() || (x != 0.0 && x != 1.0)) return ifail("line " + std::to_string(line_no) + ": observation values must be 0 or 1");
                value = x != 0.0;
            }
            const std::string_view rel = next_token(q, le);
            auto rit = rel_index.find(rel);
            if (rit == rel_index.end()) return ifail("line " + std::to_string(line_no) + ": unknown relation " + std::string(rel));
            const int r = rit->second;
            const std::vector<int32_t> &doms = ds->rel_doms[(size_t)r];
            for (size_t k = 0; k < doms.size(); k++) {
                const std::string_view ent = next_token(q, le);
                if (ent.empty()) return ifail("line " + std::to_string(line_no) + ": too few entities for relation " + std::string(rel));
                auto &cm = ds->codes[(size_t)doms[k]];
                auto it = cm.find(ent);
                if (it == cm.end()) {
                    it = cm.emplace(ent, (int32_t)cm.size()).first;
                    ds->names[(size_t)doms[k]].push_back(ent);
                }
                ds->ids[(size_t)r].push_back(it->second);
            }
            if (!next_token(q, le).empty()) return ifail("line " + std::to_string(line_no) + ": too many entities for relation " + std::string(rel));
            const std::vector<int32_t> &all = ds->ids[(size_t)r];
            const std::string key((const char *)(all.data() + all.size() - doms.size()), doms.size() * sizeof(int32_t));
            if (!seen[(size_t)r].insert(key).second)
                return ifail("line " + std::to_string(line_no) + ": observation listed twice for relation " + std::string(rel));
            ds->val[(size_t)r].push_back((uint8_t)value);
        }
        p = le + 1;
    }
    *out = guard.release();
    return 0;
}

extern "C" int hirm_dataset_free(hirm_dataset *ds) { delete ds; return 0; }

extern "C" int hirm_dataset_counts(const hirm_dataset *ds, int32_t *n_domains, int32_t *n_relations) {
    if (!ds) return ifail("dataset is null");
    if (n_domains) *n_domains = (int32_t)ds->dom_names.size();
    if (n_relations) *n_relations = (int32_t)ds->rel_names.size();
    return 0;
}
extern "C" int hirm_dataset_domain(const hirm_dataset *ds, int dom, const char **name, int32_t *n_entities) {
    if (!ds || dom < 0 || dom >= (int)ds->dom_names.size()) return ifail("bad domain index");
    if (name) *name = ds->dom_names[(size_t)dom].c_str();
    if (n_entities) *n_entities = (int32_t)ds->names[(size_t)dom].size();
    return 0;
}
extern "C" int hirm_dataset_relation(const hirm_dataset *ds, int rel, const char **name, int32_t *arity, int32_t *doms, int64_t *nobs) {
    if (!ds || rel < 0 || rel >= (int)ds->rel_names.size()) return ifail("bad relation index");
    if (name) *name = ds->rel_names[(size_t)rel].c_str();
    if (arity) *arity = (int32_t)ds->rel_doms[(size_t)rel].size();
    if (doms) for (size_t k = 0; k < ds->rel_doms[(size_t)rel].size(); k++) doms[k] = ds->rel_doms[(size_t)rel][k];
    if (nobs) *nobs = (int64_t)ds->val[(size_t)rel].size();
    return 0;
}
extern "C" int hirm_dataset_observations(const hirm_dataset *ds, int rel, const int32_t **ids, const uint8_t **val) {
    if (!ds || rel < 0 || rel >= (int)ds->rel_names.size()) return ifail("bad relation index");
    if (ids) *ids = ds->ids[(size_t)rel].data();
    if (val) *val = ds->val[(size_t)rel].data();
    return 0;
}
extern "C" int hirm_dataset_entity(hirm_dataset *ds, int dom, int32_t code, const char **name) {
    if (!ds || dom < 0 || dom >= (int)ds->dom_names.size() || !name) return ifail("bad domain index");
    if (code < 0 || code >= (int32_t)ds->names[(size_t)dom].size()) return ifail("entity code out of range");
    ds->scratch.assign(ds->names[(size_t)dom][(size_t)code]);
    *name = ds->scratch.c_str();
    return 0;
}
