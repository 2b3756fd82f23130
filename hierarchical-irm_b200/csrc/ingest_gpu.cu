// GPU ingest of the reference's text formats: <base>.schema + <base>.obs -> per relation a COO array of integer entity
// codes and a value array in DEVICE memory, ready for hirm_irm_set_observations_device / hirm_unary_block_create.
// Replaces, for files of 1e8 lines and more, the serial host reader (csrc/ingest.cc; the reference's own loader is
// load_schema + load_observations + encode_observations, cxx/util_io.cc:11-96, one hash-map insert per token on one core).
//
// Same encoding as the reference: per domain, entities are numbered 0, 1, 2 ... in order of FIRST APPEARANCE in the .obs
// file (lines top to bottom, positions left to right; cxx/util_io.cc:63-96).  The observations of a relation come out in an
// unspecified order (the engine's stores are order-free: CSR build and count tables use integer atomics).
//
// Pipeline (every pass is a streaming kernel over the text or over per-line records):
//   1. the file is mapped and copied to the device as it is;
//   2. newline count per 1 KiB tile (one warp each) -> host scan of the tile counts -> line start offsets;
//   3. one thread per line: value, relation (looked up in a device table of the schema's names, bytes compared), entity
//      tokens hashed (FNV-1a 64 + mix) -> per-line records;
//   4. per domain an open-addressing table keyed by the token hash: atomicCAS insert, atomicMin of the occurrence index
//      (line * 4 + position) = first appearance; a second pass compares every occurrence's bytes with the first one's
//      (a 64-bit collision is reported, not ignored);
//   5. the distinct entities (first occurrence, table slot) go to the host, are sorted by first occurrence (that IS the
//      reference's code order) and the codes go back into the table;
//   6. one thread per line appends its observation to its relation's COO arrays (warp-aggregated cursor);
//   7. duplicates: a table keyed by a hash of (relation, codes) holding the line index; a hit with equal codes is an error
//      (Relation::incorporate asserts distinct observations, cxx/hirm.hh:272).
// The host keeps the entity NAMES (slices of the mapped file) for the writers.
#include "../../include/hirm_engine.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

#include <cuda_runtime.h>

namespace {

thread_local std::string g_gerr;
int gfail(const std::string &m) { g_gerr = m; return -1; }
#define GCK(call)                                                                                   \
    do {                                                                                            \
        cudaError_t _e = (call);                                                                    \
        if (_e != cudaSuccess) return gfail(std::string(#call) + ": " + cudaGetErrorString(_e));    \
    } while (0)

constexpr int MAXA = HIRM_MAX_ARITY;
constexpr int TILE = 1024;                 // bytes per warp in the newline passes
enum : int { ERR_NONE = 0, ERR_VALUE = 1, ERR_RELATION = 2, ERR_FEW = 3, ERR_MANY = 4, ERR_COLLISION = 5, ERR_DUPLICATE = 6, ERR_TABLE = 7 };

struct RelEntry { uint64_t hash; int32_t name_off, name_len, arity, pad; int32_t dom[MAXA]; };

__host__ __device__ inline uint64_t fnv1a(const char *p, int n) {
    uint64_t h = 0xcbf29ce484222325ull;
    for (int i = 0; i < n; i++) { h ^= (unsigned char)p[i]; h *= 0x100000001b3ull; }
    h ^= h >> 32; h *= 0x9E3779B97F4A7C15ull; h ^= h >> 29;
    return h ? h : 1ull;                   // 0 marks an empty table slot
}
__device__ __forceinline__ bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n'; }

__global__ void count_newlines_kernel(const char *text, int64_t size, int32_t *tile_counts, int64_t ntiles) {
    const int lane = threadIdx.x & 31;
    for (int64_t tile = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); tile < ntiles; tile += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        const int64_t p0 = tile * TILE + lane * 32;
        int c = 0;
        if (p0 + 32 <= size) {
            const uint4 *q = reinterpret_cast<const uint4 *>(text + p0);
            for (int k = 0; k < 2; k++) {
                const uint4 v = q[k];
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                for (int j = 0; j < 4; j++)
                    for (int b = 0; b < 4; b++) c += ((w[j] >> (8 * b)) & 0xFFu) == (uint32_t)'\n';
            }
        } else {
            for (int64_t p = p0; p < size && p < p0 + 32; p++) c += text[p] == '\n';
        }
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0) tile_counts[tile] = c;
    }
}
// line_start[l] = offset of the first byte of line l (line 0 starts at 0); tile_base[tile] = newlines before the tile
__global__ void line_starts_kernel(const char *text, int64_t size, const int64_t *tile_base, int64_t ntiles, int64_t *line_start) {
    const int lane = threadIdx.x & 31;
    for (int64_t tile = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); tile < ntiles; tile += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        const int64_t p0 = tile * TILE + lane * 32;
        int c = 0;
        for (int64_t p = p0; p < size && p < p0 + 32; p++) c += text[p] == '\n';
        int incl = c;
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        int64_t k = tile_base[tile] + incl - c;
        for (int64_t p = p0; p < size && p < p0 + 32; p++)
            if (text[p] == '\n') line_start[++k] = p + 1;
    }
}

struct LineRec { int32_t rel; uint8_t val, nent, pad0, pad1; };
// one thread per line
__global__ void parse_lines_kernel(const char *text, int64_t size, const int64_t *line_start, int64_t nlines, const RelEntry *rels, int rel_cap,
                                   const char *names, LineRec *recs, uint64_t *ent_hash, int64_t *ent_off, int32_t *ent_len,
                                   int ndom, unsigned long long *rel_count, unsigned long long *dom_tokens, unsigned long long *err) {
    const int lane = threadIdx.x & 31;
    for (int64_t l0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) - lane; l0 < nlines; l0 += (int64_t)gridDim.x * blockDim.x) {   // warp-uniform trip count
      const int64_t l = l0 + lane;
      LineRec R{-1, 0, 0, 0, 0};
      int my_dom[MAXA] = {-1, -1, -1, -1};
      if (l < nlines) do {
        int64_t p = line_start[l];
        const int64_t e = l + 1 < nlines ? line_start[l + 1] - 1 : size;    // the newline (or end of file)
        auto token = [&](int64_t &a, int64_t &b) {
            while (p < e && is_space(text[p])) p++;
            a = p;
            while (p < e && !is_space(text[p])) p++;
            b = p;
            return b > a;
        };
        auto fail = [&](int code) { atomicMin(err, ((unsigned long long)l << 8) | (unsigned long long)code); };
        int64_t a, b;
        if (!token(a, b)) break;                                            // blank line
        // value: 0 or 1, optionally written as a decimal (0.0, 1., 1.000)
        bool ok = (text[a] == '0' || text[a] == '1');
        for (int64_t q = a + 1; q < b && ok; q++) ok = (q == a + 1 && text[q] == '.') || (q > a + 1 && text[a + 1] == '.' && text[q] == '0');
        if (!ok) { fail(ERR_VALUE); break; }
        R.val = text[a] == '1';
        if (!token(a, b)) { fail(ERR_RELATION); break; }
        const uint64_t h = fnv1a(text + a, (int)(b - a));
        const RelEntry *E = nullptr;
        for (uint32_t i = (uint32_t)h & (rel_cap - 1), n = 0; n < (uint32_t)rel_cap; i = (i + 1) & (rel_cap - 1), n++) {
            const RelEntry &C = rels[i];
            if (C.hash == 0) break;
            if (C.hash == h && C.name_len == (int)(b - a)) {
                bool same = true;
                for (int k = 0; k < C.name_len && same; k++) same = names[C.name_off + k] == text[a + k];
                if (same) { E = &C; break; }
            }
        }
        if (!E) { fail(ERR_RELATION); break; }
        int ne = 0;
        bool bad = false;
        while (token(a, b)) {
            if (ne == E->arity) { fail(ERR_MANY); bad = true; break; }
            ent_hash[l * MAXA + ne] = fnv1a(text + a, (int)(b - a));
            ent_off[l * MAXA + ne] = a; ent_len[l * MAXA + ne] = (int32_t)(b - a);
            ne++;
        }
        if (!bad && ne < E->arity) { fail(ERR_FEW); bad = true; }
        if (bad) break;
        R.rel = E->pad; R.nent = (uint8_t)ne;
        for (int k = 0; k < ne; k++) my_dom[k] = E->dom[k];
      } while (0);
      if (l < nlines) recs[l] = R;
      // counters, one atomic per warp and distinct value: lines per relation, entity tokens per domain
      const unsigned peers = __match_any_sync(0xffffffffu, R.rel);
      if (R.rel >= 0 && lane == __ffs((int)peers) - 1) atomicAdd(&rel_count[R.rel], (unsigned long long)__popc(peers));
      for (int d = 0; d < ndom; d++) {
          int c = 0;
          for (int k = 0; k < MAXA; k++) c += my_dom[k] == d;
          c = __reduce_add_sync(0xffffffffu, c);
          if (lane == 0 && c) atomicAdd(&dom_tokens[d], (unsigned long long)c);
      }
    }
}

struct DomTable { uint64_t *key; unsigned long long *first; int32_t *code; uint32_t mask; };
__device__ __forceinline__ uint32_t dom_slot(const DomTable &T, uint64_t h, bool insert, unsigned long long *err, int64_t l) {
    for (uint32_t i = (uint32_t)(h >> 20) & T.mask, n = 0; n <= T.mask; i = (i + 1) & T.mask, n++) {
        uint64_t k = T.key[i];
        if (k == 0 && insert) { k = atomicCAS((unsigned long long *)&T.key[i], 0ull, (unsigned long long)h); if (k == 0) k = h; }
        if (k == h) return i;
        if (k == 0) break;
    }
    atomicMin(err, ((unsigned long long)l << 8) | (unsigned long long)ERR_TABLE);
    return 0;
}
__global__ void dict_insert_kernel(const LineRec *recs, int64_t nlines, const int32_t *rel_dom /* [nrel][MAXA] */, const uint64_t *ent_hash,
                                   const DomTable *tabs, unsigned long long *err) {
    for (int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < nlines; l += (int64_t)gridDim.x * blockDim.x) {
        const LineRec R = recs[l];
        if (R.rel < 0) continue;
        for (int k = 0; k < R.nent; k++) {
            const DomTable &T = tabs[rel_dom[R.rel * MAXA + k]];
            const uint32_t i = dom_slot(T, ent_hash[l * MAXA + k], true, err, l);
            atomicMin(&T.first[i], (unsigned long long)(l * MAXA + k));
        }
    }
}
// every occurrence spells the same name as the first occurrence of its hash (else: a 64-bit collision)
__global__ void dict_verify_kernel(const char *text, const LineRec *recs, int64_t nlines, const int32_t *rel_dom, const uint64_t *ent_hash,
                                   const int64_t *ent_off, const int32_t *ent_len, const DomTable *tabs, unsigned long long *err) {
    for (int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < nlines; l += (int64_t)gridDim.x * blockDim.x) {
        const LineRec R = recs[l];
        if (R.rel < 0) continue;
        for (int k = 0; k < R.nent; k++) {
            const DomTable &T = tabs[rel_dom[R.rel * MAXA + k]];
            const uint32_t i = dom_slot(T, ent_hash[l * MAXA + k], false, err, l);
            const unsigned long long f = T.first[i];
            if (f == (unsigned long long)(l * MAXA + k)) continue;
            const int len = ent_len[l * MAXA + k];
            bool same = ent_len[f] == len;
            for (int q = 0; q < len && same; q++) same = text[ent_off[f] + q] == text[ent_off[l * MAXA + k] + q];
            if (!same) atomicMin(err, ((unsigned long long)l << 8) | (unsigned long long)ERR_COLLISION);
        }
    }
}
__global__ void dict_list_kernel(DomTable T, unsigned long long *out_first, uint32_t *out_slot, unsigned long long *n_out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= T.mask; i += gridDim.x * blockDim.x) {
        if (T.key[i] == 0) continue;
        const unsigned long long k = atomicAdd(n_out, 1ull);
        out_first[k] = T.first[i]; out_slot[k] = i;
    }
}
__global__ void dict_codes_kernel(DomTable T, const uint32_t *slots, int64_t n) {
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n; c += (int64_t)gridDim.x * blockDim.x) T.code[slots[c]] = (int32_t)c;
}
__global__ void gather_occurrences_kernel(const int64_t *occ, int64_t n, const int64_t *ent_off, const int32_t *ent_len, int64_t *out_off, int32_t *out_len) {
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n; c += (int64_t)gridDim.x * blockDim.x) { out_off[c] = ent_off[occ[c]]; out_len[c] = ent_len[occ[c]]; }
}
struct RelOut { int32_t *ids; uint8_t *val; unsigned long long *cursor; int32_t arity, pad; };
__global__ void write_coo_kernel(const LineRec *recs, int64_t nlines, const int32_t *rel_dom, const uint64_t *ent_hash, const DomTable *tabs,
                                 const RelOut *outs, int64_t *line_pos, unsigned long long *err) {
    const int lane = threadIdx.x & 31;
    for (int64_t l0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) - lane; l0 < nlines; l0 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t l = l0 + lane;
        LineRec R{-1, 0, 0, 0, 0};
        if (l < nlines) R = recs[l];
        // one cursor atomic per warp and relation
        const unsigned peers = __match_any_sync(0xffffffffu, R.rel);
        const int leader = __ffs((int)peers) - 1;
        unsigned long long base = 0;
        if (R.rel >= 0 && lane == leader) base = atomicAdd(outs[R.rel].cursor, (unsigned long long)__popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (l >= nlines) continue;
        if (R.rel < 0) { line_pos[l] = -1; continue; }
        const RelOut &O = outs[R.rel];
        const int64_t pos = (int64_t)base + __popc(peers & ((1u << lane) - 1u));
        line_pos[l] = pos;
        for (int k = 0; k < R.nent; k++) {
            const DomTable &T = tabs[rel_dom[R.rel * MAXA + k]];
            O.ids[pos * O.arity + k] = T.code[dom_slot(T, ent_hash[l * MAXA + k], false, err, l)];
        }
        O.val[pos] = R.val;
    }
}
// duplicate observations: table keyed by a hash of (relation, codes) holding the first line that inserted it
__global__ void duplicates_kernel(const LineRec *recs, int64_t nlines, const RelOut *outs, const int64_t *line_pos, uint64_t *keys,
                                  long long *lines, uint32_t mask, unsigned long long *err) {
    for (int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < nlines; l += (int64_t)gridDim.x * blockDim.x) {
        const LineRec R = recs[l];
        if (R.rel < 0) continue;
        const RelOut &O = outs[R.rel];
        const int32_t *mine = O.ids + line_pos[l] * O.arity;
        uint64_t h = 0x9E3779B97F4A7C15ull * (uint64_t)(R.rel + 1);
        for (int k = 0; k < R.nent; k++) { h ^= (uint64_t)(uint32_t)mine[k] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); h *= 0xBF58476D1CE4E5B9ull; }
        h ^= h >> 31;
        if (h == 0) h = 1;
        bool placed = false;
        for (uint32_t i = (uint32_t)(h >> 16) & mask, n = 0; n <= mask && !placed; i = (i + 1) & mask, n++) {
            uint64_t k = keys[i];
            if (k == 0) {
                k = atomicCAS((unsigned long long *)&keys[i], 0ull, (unsigned long long)h);
                if (k == 0) { atomicExch((unsigned long long *)&lines[i], (unsigned long long)l); placed = true; break; }
            }
            if (k != h) continue;
            long long other;
            while ((other = *(volatile long long *)&lines[i]) < 0) { }     // the inserting thread is about to publish its line
            const LineRec R2 = recs[other];
            bool same = R2.rel == R.rel;
            const int32_t *theirs = outs[R2.rel].ids + line_pos[other] * outs[R2.rel].arity;
            for (int q = 0; q < R.nent && same; q++) same = theirs[q] == mine[q];
            if (same) { atomicMin(err, ((unsigned long long)(l > other ? l : other) << 8) | (unsigned long long)ERR_DUPLICATE); placed = true; }
        }
        if (!placed) atomicMin(err, ((unsigned long long)l << 8) | (unsigned long long)ERR_TABLE);
    }
}

uint32_t pow2_at_least(uint64_t v) { uint32_t p = 1024; while (p < v && p < (1u << 30)) p <<= 1; return p; }
unsigned grid_for(int64_t n, int per = 256) { return (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + per - 1) / per, 148 * 32)); }

struct Mapped {
    const char *p = nullptr; size_t n = 0; int fd = -1;
    ~Mapped() { if (p && n) munmap((void *)p, n); if (fd >= 0) close(fd); }
    bool open_file(const char *path) {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) return false;
        n = (size_t)st.st_size;
        if (n == 0) return true;
        void *q = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (q == MAP_FAILED) { p = nullptr; return false; }
        p = (const char *)q;
        return true;
    }
};

template <typename T>
struct Dev {
    T *p = nullptr;
    ~Dev() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n, bool zero = false) {
        cudaError_t e = cudaMalloc((void **)&p, std::max<size_t>(n, 1) * sizeof(T));
        if (e == cudaSuccess && zero) e = cudaMemset(p, 0, std::max<size_t>(n, 1) * sizeof(T));
        return e;
    }
    T *release() { T *q = p; p = nullptr; return q; }
};

}  // namespace

struct hirm_gpu_dataset {
    int device = 0;
    std::vector<std::string> dom_names, rel_names;
    std::vector<std::vector<int32_t>> rel_doms;
    std::vector<int64_t> nobs;
    std::vector<int32_t *> d_ids; std::vector<uint8_t *> d_val;         // device, owned
    std::vector<std::vector<std::string>> names;                        // [domain][code]
    double seconds[4] = {0, 0, 0, 0};                                   // copy, split + parse, dictionary, COO
    ~hirm_gpu_dataset() {
        for (auto p : d_ids) if (p) cudaFree(p);
        for (auto p : d_val) if (p) cudaFree(p);
    }
};

extern "C" const char *hirm_gpu_dataset_last_error(void) { return g_gerr.c_str(); }

extern "C" int hirm_gpu_dataset_load(const char *schema_path, const char *obs_path, int device, hirm_gpu_dataset **out) {
    if (!schema_path || !obs_path || !out) return gfail("null argument");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return gfail("no usable CUDA device (the GPU reader has no CPU fallback: use hirm_dataset_load)");
    if (device < 0) GCK(cudaGetDevice(&device));
    GCK(cudaSetDevice(device));
    auto ds = std::make_unique<hirm_gpu_dataset>();
    ds->device = device;
    // ---- schema (host; cxx/util_io.cc:11-34) ---------------------------------------------------------------------------
    Mapped sch;
    if (!sch.open_file(schema_path)) return gfail(std::string("cannot read ") + schema_path);
    std::unordered_map<std::string, int> dom_index, rel_index;
    {
        const char *p = sch.p, *fe = sch.p + sch.n;
        while (p < fe) {
            const char *le = (const char *)memchr(p, '\n', (size_t)(fe - p));
            if (!le) le = fe;
            std::vector<std::string> w;
            for (const char *q = p; q < le;) {
                while (q < le && (*q == ' ' || *q == '\t' || *q == '\r')) q++;
                const char *b = q;
                while (q < le && !(*q == ' ' || *q == '\t' || *q == '\r')) q++;
                if (q > b) w.emplace_back(b, q);
            }
            if (!w.empty()) {
                if (w[0] != "bernoulli") return gfail("only bernoulli relations are supported (README.md:122-123), got " + w[0]);
                if (w.size() < 3 || w.size() > 2 + MAXA) return gfail("a relation needs 1 to HIRM_MAX_ARITY domains: " + (w.size() > 1 ? w[1] : std::string("?")));
                if (rel_index.count(w[1])) return gfail("relation listed twice in the schema: " + w[1]);
                std::vector<int32_t> doms;
                for (size_t k = 2; k < w.size(); k++) {
                    auto it = dom_index.find(w[k]);
                    if (it == dom_index.end()) { it = dom_index.emplace(w[k], (int)ds->dom_names.size()).first; ds->dom_names.push_back(w[k]); }
                    doms.push_back(it->second);
                }
                rel_index.emplace(w[1], (int)ds->rel_names.size());
                ds->rel_names.push_back(w[1]); ds->rel_doms.push_back(doms);
            }
            p = le + 1;
        }
    }
    const int nrel = (int)ds->rel_names.size(), ndom = (int)ds->dom_names.size();
    if (nrel == 0) return gfail("the schema lists no relation");
    ds->nobs.assign((size_t)nrel, 0); ds->d_ids.assign((size_t)nrel, nullptr); ds->d_val.assign((size_t)nrel, nullptr);
    ds->names.resize((size_t)ndom);
    // device table of the relation names
    const int rel_cap = (int)pow2_at_least((uint64_t)nrel * 2);
    std::vector<RelEntry> h_rels((size_t)rel_cap);
    memset(h_rels.data(), 0, h_rels.size() * sizeof(RelEntry));
    std::string pool;
    std::vector<int32_t> h_rel_dom((size_t)nrel * MAXA, 0);
    for (int r = 0; r < nrel; r++) {
        const std::string &nm = ds->rel_names[(size_t)r];
        const uint64_t h = fnv1a(nm.data(), (int)nm.size());
        uint32_t i = (uint32_t)h & (rel_cap - 1);
        while (h_rels[i].hash) i = (i + 1) & (rel_cap - 1);
        RelEntry &E = h_rels[i];
        E.hash = h; E.name_off = (int32_t)pool.size(); E.name_len = (int32_t)nm.size(); E.arity = (int32_t)ds->rel_doms[(size_t)r].size(); E.pad = r;
        for (int k = 0; k < E.arity; k++) { E.dom[k] = ds->rel_doms[(size_t)r][(size_t)k]; h_rel_dom[(size_t)r * MAXA + k] = E.dom[k]; }
        pool += nm;
    }
    // ---- 1. the text, as it is ---------------------------------------------------------------------------------------------
    cudaEvent_t ev[5];
    for (auto &e : ev) GCK(cudaEventCreate(&e));
    Mapped obs;
    if (!obs.open_file(obs_path)) return gfail(std::string("cannot read ") + obs_path);
    const int64_t size = (int64_t)obs.n;
    Dev<char> d_text, d_names; Dev<RelEntry> d_rels; Dev<int32_t> d_rel_dom;
    GCK(d_text.alloc((size_t)size + 64, false));
    GCK(cudaEventRecord(ev[0]));
    if (size) GCK(cudaMemcpy(d_text.p, obs.p, (size_t)size, cudaMemcpyHostToDevice));
    GCK(cudaMemset(d_text.p + size, '\n', 64));
    GCK(d_names.alloc(pool.size() + 1)); GCK(cudaMemcpy(d_names.p, pool.data(), pool.size(), cudaMemcpyHostToDevice));
    GCK(d_rels.alloc(h_rels.size())); GCK(cudaMemcpy(d_rels.p, h_rels.data(), h_rels.size() * sizeof(RelEntry), cudaMemcpyHostToDevice));
    GCK(d_rel_dom.alloc(h_rel_dom.size())); GCK(cudaMemcpy(d_rel_dom.p, h_rel_dom.data(), h_rel_dom.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    GCK(cudaEventRecord(ev[1]));
    // ---- 2. lines ------------------------------------------------------------------------------------------------------------
    const int64_t ntiles = (size + TILE - 1) / TILE;
    Dev<int32_t> d_tc; Dev<int64_t> d_tb, d_ls;
    GCK(d_tc.alloc((size_t)ntiles, true));
    if (ntiles) count_newlines_kernel<<<grid_for(ntiles, 8), 256>>>(d_text.p, size, d_tc.p, ntiles);
    GCK(cudaGetLastError());
    std::vector<int32_t> h_tc((size_t)ntiles);
    if (ntiles) GCK(cudaMemcpy(h_tc.data(), d_tc.p, (size_t)ntiles * sizeof(int32_t), cudaMemcpyDeviceToHost));
    std::vector<int64_t> h_tb((size_t)ntiles);
    int64_t newlines = 0;
    for (int64_t t = 0; t < ntiles; t++) { h_tb[(size_t)t] = newlines; newlines += h_tc[(size_t)t]; }
    const int64_t nlines = size == 0 ? 0 : newlines + (obs.p[size - 1] == '\n' ? 0 : 1);
    GCK(d_tb.alloc((size_t)ntiles)); GCK(d_ls.alloc((size_t)newlines + 2, true));
    if (ntiles) {
        GCK(cudaMemcpy(d_tb.p, h_tb.data(), (size_t)ntiles * sizeof(int64_t), cudaMemcpyHostToDevice));
        line_starts_kernel<<<grid_for(ntiles, 8), 256>>>(d_text.p, size, d_tb.p, ntiles, d_ls.p);
        GCK(cudaGetLastError());
    }
    // ---- 3. per-line records ---------------------------------------------------------------------------------------------------
    Dev<LineRec> d_recs; Dev<uint64_t> d_eh; Dev<int64_t> d_eo, d_pos; Dev<int32_t> d_el; Dev<unsigned long long> d_relc, d_domt, d_err;
    GCK(d_recs.alloc((size_t)nlines)); GCK(d_eh.alloc((size_t)nlines * MAXA)); GCK(d_eo.alloc((size_t)nlines * MAXA)); GCK(d_el.alloc((size_t)nlines * MAXA));
    GCK(d_relc.alloc((size_t)nrel, true)); GCK(d_domt.alloc((size_t)ndom, true)); GCK(d_err.alloc(1));
    GCK(cudaMemset(d_err.p, 0xFF, sizeof(unsigned long long)));
    if (nlines) parse_lines_kernel<<<grid_for(nlines), 256>>>(d_text.p, size, d_ls.p, nlines, d_rels.p, rel_cap, d_names.p, d_recs.p, d_eh.p, d_eo.p, d_el.p,
                                                             ndom, d_relc.p, d_domt.p, d_err.p);
    GCK(cudaGetLastError());
    GCK(cudaEventRecord(ev[2]));
    auto check = [&](const char *stage) -> int {
        unsigned long long e = 0;
        GCK(cudaMemcpy(&e, d_err.p, sizeof(e), cudaMemcpyDeviceToHost));
        if (e == ~0ull) return 0;
        const long long line = (long long)(e >> 8) + 1;
        static const char *text[] = {"", "observation values must be 0 or 1", "unknown relation", "too few entities for the relation",
                                     "too many entities for the relation", "two entity names share a 64-bit hash", "observation listed twice",
                                     "internal table full"};
        return gfail(std::string("line ") + std::to_string(line) + ": " + text[e & 0xFF] + " (" + stage + ")");
    };
    if (check("parse")) return -1;
    std::vector<unsigned long long> h_relc((size_t)nrel), h_domt((size_t)ndom);
    GCK(cudaMemcpy(h_relc.data(), d_relc.p, (size_t)nrel * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    GCK(cudaMemcpy(h_domt.data(), d_domt.p, (size_t)ndom * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    // ---- 4./5. dictionaries -------------------------------------------------------------------------------------------------------
    std::vector<DomTable> h_tabs((size_t)ndom);
    std::vector<Dev<uint64_t>> keys((size_t)ndom); std::vector<Dev<unsigned long long>> firsts((size_t)ndom); std::vector<Dev<int32_t>> codes((size_t)ndom);
    for (int d = 0; d < ndom; d++) {
        const uint32_t cap = pow2_at_least(std::min<uint64_t>(2 * h_domt[(size_t)d] + 16, 1ull << 28));
        GCK(keys[(size_t)d].alloc(cap, true)); GCK(firsts[(size_t)d].alloc(cap)); GCK(codes[(size_t)d].alloc(cap, true));
        GCK(cudaMemset(firsts[(size_t)d].p, 0xFF, (size_t)cap * sizeof(unsigned long long)));
        h_tabs[(size_t)d] = DomTable{keys[(size_t)d].p, firsts[(size_t)d].p, codes[(size_t)d].p, cap - 1};
    }
    Dev<DomTable> d_tabs;
    GCK(d_tabs.alloc((size_t)ndom)); GCK(cudaMemcpy(d_tabs.p, h_tabs.data(), (size_t)ndom * sizeof(DomTable), cudaMemcpyHostToDevice));
    if (nlines) {
        dict_insert_kernel<<<grid_for(nlines), 256>>>(d_recs.p, nlines, d_rel_dom.p, d_eh.p, d_tabs.p, d_err.p);
        GCK(cudaGetLastError());
        dict_verify_kernel<<<grid_for(nlines), 256>>>(d_text.p, d_recs.p, nlines, d_rel_dom.p, d_eh.p, d_eo.p, d_el.p, d_tabs.p, d_err.p);
        GCK(cudaGetLastError());
    }
    if (check("dictionary")) return -1;
    for (int d = 0; d < ndom; d++) {
        const uint32_t cap = h_tabs[(size_t)d].mask + 1;
        Dev<unsigned long long> d_first, d_n; Dev<uint32_t> d_slot;
        GCK(d_first.alloc(cap)); GCK(d_slot.alloc(cap)); GCK(d_n.alloc(1, true));
        dict_list_kernel<<<grid_for(cap), 256>>>(h_tabs[(size_t)d], d_first.p, d_slot.p, d_n.p);
        GCK(cudaGetLastError());
        unsigned long long n = 0;
        GCK(cudaMemcpy(&n, d_n.p, sizeof(n), cudaMemcpyDeviceToHost));
        std::vector<unsigned long long> hf((size_t)n); std::vector<uint32_t> hs((size_t)n);
        if (n) {
            GCK(cudaMemcpy(hf.data(), d_first.p, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            GCK(cudaMemcpy(hs.data(), d_slot.p, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        }
        std::vector<uint32_t> order((size_t)n);
        for (size_t k = 0; k < (size_t)n; k++) order[k] = (uint32_t)k;
        std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return hf[a] < hf[b]; });   // first appearance = the reference's code order
        std::vector<uint32_t> slots_sorted((size_t)n);
        for (size_t c = 0; c < (size_t)n; c++) slots_sorted[c] = hs[order[c]];
        if (n) {
            GCK(cudaMemcpy(d_slot.p, slots_sorted.data(), (size_t)n * sizeof(uint32_t), cudaMemcpyHostToDevice));
            dict_codes_kernel<<<grid_for((int64_t)n), 256>>>(h_tabs[(size_t)d], d_slot.p, (int64_t)n);
            GCK(cudaGetLastError());
        }
        // names: the first occurrence's token, read from the mapped file (offsets / lengths of those occurrences come back)
        std::vector<int64_t> occ((size_t)n);
        for (size_t c = 0; c < (size_t)n; c++) occ[c] = (int64_t)hf[order[c]];
        auto &nm = ds->names[(size_t)d];
        nm.resize((size_t)n);
        // gather (offset, length) of the n first occurrences: small strided reads
        std::vector<int64_t> h_off((size_t)n); std::vector<int32_t> h_len((size_t)n);
        if (n) {
            Dev<int64_t> d_occ, d_off; Dev<int32_t> d_len;
            GCK(d_occ.alloc((size_t)n)); GCK(d_off.alloc((size_t)n)); GCK(d_len.alloc((size_t)n));
            GCK(cudaMemcpy(d_occ.p, occ.data(), (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice));
            gather_occurrences_kernel<<<grid_for((int64_t)n), 256>>>(d_occ.p, (int64_t)n, d_eo.p, d_el.p, d_off.p, d_len.p);
            GCK(cudaGetLastError());
            GCK(cudaMemcpy(h_off.data(), d_off.p, (size_t)n * sizeof(int64_t), cudaMemcpyDeviceToHost));
            GCK(cudaMemcpy(h_len.data(), d_len.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost));
        }
        for (size_t c = 0; c < (size_t)n; c++) nm[c].assign(obs.p + h_off[c], (size_t)h_len[c]);
    }
    GCK(cudaEventRecord(ev[3]));
    // ---- 6./7. COO arrays, duplicates ----------------------------------------------------------------------------------------------
    std::vector<RelOut> h_outs((size_t)nrel);
    Dev<unsigned long long> d_cursor;
    GCK(d_cursor.alloc((size_t)nrel, true));
    for (int r = 0; r < nrel; r++) {
        const int a = (int)ds->rel_doms[(size_t)r].size();
        ds->nobs[(size_t)r] = (int64_t)h_relc[(size_t)r];
        GCK(cudaMalloc((void **)&ds->d_ids[(size_t)r], std::max<size_t>((size_t)h_relc[(size_t)r] * a, 1) * sizeof(int32_t)));
        GCK(cudaMalloc((void **)&ds->d_val[(size_t)r], std::max<size_t>((size_t)h_relc[(size_t)r], 1)));
        h_outs[(size_t)r] = RelOut{ds->d_ids[(size_t)r], ds->d_val[(size_t)r], d_cursor.p + r, a, 0};
    }
    Dev<RelOut> d_outs;
    GCK(d_outs.alloc((size_t)nrel)); GCK(cudaMemcpy(d_outs.p, h_outs.data(), (size_t)nrel * sizeof(RelOut), cudaMemcpyHostToDevice));
    GCK(d_pos.alloc((size_t)nlines));
    if (nlines) {
        write_coo_kernel<<<grid_for(nlines), 256>>>(d_recs.p, nlines, d_rel_dom.p, d_eh.p, d_tabs.p, d_outs.p, d_pos.p, d_err.p);
        GCK(cudaGetLastError());
        const uint32_t cap = pow2_at_least(std::min<uint64_t>(2 * (uint64_t)nlines + 16, 1ull << 29));
        Dev<uint64_t> d_k; Dev<long long> d_l;
        GCK(d_k.alloc(cap, true)); GCK(d_l.alloc(cap));
        GCK(cudaMemset(d_l.p, 0xFF, (size_t)cap * sizeof(long long)));
        duplicates_kernel<<<grid_for(nlines), 256>>>(d_recs.p, nlines, d_outs.p, d_pos.p, d_k.p, d_l.p, cap - 1, d_err.p);
        GCK(cudaGetLastError());
    }
    if (check("observations")) return -1;
    GCK(cudaEventRecord(ev[4]));
    GCK(cudaEventSynchronize(ev[4]));
    for (int k = 0; k < 4; k++) { float ms = 0; cudaEventElapsedTime(&ms, ev[k], ev[k + 1]); ds->seconds[k] = ms * 1e-3; }
    for (auto &e : ev) cudaEventDestroy(e);
    *out = ds.release();
    return 0;
}

extern "C" int hirm_gpu_dataset_free(hirm_gpu_dataset *ds) {
    if (ds) { cudaSetDevice(ds->device); delete ds; }
    return 0;
}
extern "C" int hirm_gpu_dataset_counts(const hirm_gpu_dataset *ds, int32_t *n_domains, int32_t *n_relations) {
    if (!ds) return gfail("dataset is null");
    if (n_domains) *n_domains = (int32_t)ds->dom_names.size();
    if (n_relations) *n_relations = (int32_t)ds->rel_names.size();
    return 0;
}
extern "C" int hirm_gpu_dataset_domain(const hirm_gpu_dataset *ds, int dom, const char **name, int32_t *n_entities) {
    if (!ds || dom < 0 || dom >= (int)ds->dom_names.size()) return gfail("bad domain index");
    if (name) *name = ds->dom_names[(size_t)dom].c_str();
    if (n_entities) *n_entities = (int32_t)ds->names[(size_t)dom].size();
    return 0;
}
extern "C" int hirm_gpu_dataset_relation(const hirm_gpu_dataset *ds, int rel, const char **name, int32_t *arity, int32_t *doms, int64_t *nobs,
                                         const int32_t **d_ids, const uint8_t **d_val) {
    if (!ds || rel < 0 || rel >= (int)ds->rel_names.size()) return gfail("bad relation index");
    if (name) *name = ds->rel_names[(size_t)rel].c_str();
    if (arity) *arity = (int32_t)ds->rel_doms[(size_t)rel].size();
    if (doms) for (size_t k = 0; k < ds->rel_doms[(size_t)rel].size(); k++) doms[k] = ds->rel_doms[(size_t)rel][k];
    if (nobs) *nobs = ds->nobs[(size_t)rel];
    if (d_ids) *d_ids = ds->d_ids[(size_t)rel];
    if (d_val) *d_val = ds->d_val[(size_t)rel];
    return 0;
}
extern "C" int hirm_gpu_dataset_entity(const hirm_gpu_dataset *ds, int dom, int32_t code, const char **name) {
    if (!ds || dom < 0 || dom >= (int)ds->dom_names.size() || !name) return gfail("bad domain index");
    if (code < 0 || code >= (int32_t)ds->names[(size_t)dom].size()) return gfail("entity code out of range");
    *name = ds->names[(size_t)dom][(size_t)code].c_str();
    return 0;
}
// copy one relation's COO arrays into the caller's device buffers (e.g. torch tensors), ids [nobs * arity], val [nobs]
extern "C" int hirm_gpu_dataset_copy_relation(const hirm_gpu_dataset *ds, int rel, int32_t *d_ids_out, uint8_t *d_val_out) {
    if (!ds || rel < 0 || rel >= (int)ds->rel_names.size()) return gfail("bad relation index");
    GCK(cudaSetDevice(ds->device));
    const size_t n = (size_t)ds->nobs[(size_t)rel], a = ds->rel_doms[(size_t)rel].size();
    if (n && (!d_ids_out || !d_val_out)) return gfail("null output buffer");
    if (n) {
        GCK(cudaMemcpy(d_ids_out, ds->d_ids[(size_t)rel], n * a * sizeof(int32_t), cudaMemcpyDeviceToDevice));
        GCK(cudaMemcpy(d_val_out, ds->d_val[(size_t)rel], n, cudaMemcpyDeviceToDevice));
    }
    return 0;
}
extern "C" int hirm_gpu_dataset_timings(const hirm_gpu_dataset *ds, double *seconds4) {
    if (!ds || !seconds4) return gfail("null argument");
    for (int k = 0; k < 4; k++) seconds4[k] = ds->seconds[k];
    return 0;
}
