// Device arithmetic shared by the sweep kernels: Philox4x32-10 uniforms, integer
// log-factorial differences (the Beta-Bernoulli cell score with alpha = beta = 1),
// and the sampling rule.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace hirm {

// ---------------------------------------------------------------------------------
// log-factorial table LF[m] = lgamma(m+1), m < LF_SIZE, filled once per device by the
// host with libm's lgamma -- the same function the reference's lbeta calls
// (cxx/util_math.cc:13-15), so small-count scores agree with it to the last bit.
// ---------------------------------------------------------------------------------
constexpr int LF_SIZE = 4096;
__device__ double g_logfact[LF_SIZE];

__device__ __forceinline__ double stirling_corr(double y) {
    // lgamma(y) - [(y-1/2) ln y - y + ln(2 pi)/2] for y >= LF_SIZE: three terms are < 1e-30 off
    double r = 1.0 / y, r2 = r * r;
    return r * (1.0 / 12.0 - r2 * (1.0 / 360.0 - r2 * (1.0 / 1260.0)));
}
__device__ __forceinline__ double lgamma_large(double y) {
    return (y - 0.5) * log(y) - y + 0.91893853320467274178 + stirling_corr(y);
}
// log((a+m)!) - log(a!) for integers a >= 0, m >= 0, without cancellation for large a.
__device__ __forceinline__ double lfact_diff(int64_t a, int64_t m) {
    if (m == 0) return 0.0;
    if (a + m < LF_SIZE) return g_logfact[a + m] - g_logfact[a];
    if (m == 1) return log((double)(a + 1));
    if (a < LF_SIZE) return lgamma_large((double)(a + m + 1)) - g_logfact[a];
    double y = (double)(a + 1), dm = (double)m;
    // g(y+m) - g(y),  g = Stirling:  (y-1/2) log1p(m/y) + m ln(y+m) - m + corr(y+m) - corr(y)
    return (y - 0.5) * log1p(dm / y) + dm * log(y + dm) - dm + (stirling_corr(y + dm) - stirling_corr(y));
}

// 1/y for y >= 1: MUFU.RCP seed + two Newton steps in fp64 (the fp64 divide is a long dependent chain)
__device__ __forceinline__ double fast_rcp(double y) {
    double r = (double)__frcp_rn((float)y);
    r = r * (2.0 - y * r);
    return r * (2.0 - y * r);
}
__device__ __forceinline__ double stirling_corr_r(double r) {   // same series, argument r = 1/y
    double r2 = r * r;
    return r * (1.0 / 12.0 - r2 * (1.0 / 360.0 - r2 * (1.0 / 1260.0)));
}
// lfact_diff(a, m) split into two halves that each cost at most ONE transcendental, so that two lanes
// can evaluate them concurrently (fp64 dependent-issue latency is what bounds a move):
//   lfact_half(a, m, 0) + lfact_half(a, m, 1) == lfact_diff(a, m)   (up to rounding)
__device__ __forceinline__ double lfact_half(int64_t a, int64_t m, int half) {
    if (m == 0) return 0.0;
    if (a + m < LF_SIZE) return half ? 0.0 : g_logfact[a + m] - g_logfact[a];
    if (m == 1) return half ? 0.0 : log((double)(a + 1));
    if (a < LF_SIZE) {
        if (half) return 0.0;
        const double y2 = (double)(a + m + 1);
        return (y2 - 0.5) * log(y2) - y2 + 0.91893853320467274178 + stirling_corr_r(fast_rcp(y2)) - g_logfact[a];
    }
    const double y = (double)(a + 1), dm = (double)m;
    if (half) return dm * log(y + dm);
    const double ry = fast_rcp(y);
    return (y - 0.5) * log1p(dm * ry) - dm + (stirling_corr_r(fast_rcp(y + dm)) - stirling_corr_r(ry));
}

// Single out-of-line copies of the fp64 transcendentals for the dense sweep kernel: its decision warps
// run a long, once-per-move code path, and every inlined copy of log/log1p/exp (hundreds of bytes of
// SASS each) is instruction-cache traffic on that path.
__device__ __noinline__ double log_nl(double x) { return log(x); }
__device__ __noinline__ double log1p_nl(double x) { return log1p(x); }
__device__ __noinline__ double exp_nl(double x) { return exp(x); }

// BetaBernoulli::logp_score (cxx/hirm.hh:52-56), alpha = beta = 1:
//   S(N, s) = lgamma(s+1) + lgamma(N-s+1) - lgamma(N+2)
// score_delta = S(N+n, s+sg) - S(N, s): what logp_gibbs_exact_variant returns for one
// group (cxx/hirm.hh:418-439) and, with (N, s) the rest counts, logp_gibbs_exact_current (:399-416).
__device__ __forceinline__ double score_delta(int64_t N, int64_t s, int64_t n, int64_t sg) {
    return lfact_diff(s, sg) + lfact_diff(N - s, n - sg) - lfact_diff(N + 1, n);
}
__device__ __forceinline__ double cell_score(int64_t N, int64_t s) { return score_delta(0, 0, N, s); }

// packed count cell: N in the low 32 bits, s in the high 32 bits.  Adding or subtracting
// packed values is exact integer arithmetic on both fields as long as neither goes negative.
typedef unsigned long long cell_t;
__device__ __host__ __forceinline__ cell_t pack_cell(uint32_t n, uint32_t s) { return (cell_t)n | ((cell_t)s << 32); }
__device__ __host__ __forceinline__ uint32_t cell_n(cell_t c) { return (uint32_t)c; }
__device__ __host__ __forceinline__ uint32_t cell_s(cell_t c) { return (uint32_t)(c >> 32); }

// ---------------------------------------------------------------------------------
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11).  One counter per move.
// ---------------------------------------------------------------------------------
__device__ __host__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int round = 0; round < 10; round++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __host__ __forceinline__ double philox_uniform(uint64_t seed, uint64_t stream, uint64_t counter) {
    uint32_t w[4];
    philox4x32_10((uint32_t)counter, (uint32_t)(counter >> 32), (uint32_t)stream, (uint32_t)(stream >> 32),
                  (uint32_t)seed, (uint32_t)(seed >> 32), w);
    uint64_t hi = w[0] >> 5, lo = w[1] >> 6;   // 27 + 26 = 53 bits
    return (double)((hi << 26) | lo) * (1.0 / 9007199254740992.0);
}

// log_choice (cxx/util_math.cc:58-65) with the inversion rule of random.Random.choices
// (src/util_math.py:24-28): p = exp(lp - logsumexp(lp)); first index whose running sum
// exceeds u * sum(p); capped at n-1.  Candidates are in ascending label order.
// The rule is scale invariant -- sum_{j<=i} p_j > u sum_j p_j  <=>  sum_{j<=i} e_j > u sum_j e_j with
// e_j = exp(lp_j - max lp) -- so one exp per candidate suffices (no log, no second exp on the
// dependent chain); it picks the same index except when u falls within rounding of a CDF step.
__device__ __forceinline__ int choose_index(const double *lp, int n, double u) {
    double m = lp[0];
    for (int i = 1; i < n; i++) m = fmax(m, lp[i]);
    double total = 0;
    for (int i = 0; i < n; i++) total += exp(lp[i] - m);
    double x = u * total, c = 0;
    for (int i = 0; i < n; i++) {
        c += exp(lp[i] - m);
        if (c > x) return i;
    }
    return n - 1;
}

}  // namespace hirm
