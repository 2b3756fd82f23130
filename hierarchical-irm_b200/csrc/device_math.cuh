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
// log(m) for integers m < LOGTAB_SIZE (libm's log, filled once per device by the engine; 8 MB, L2-resident): the score of a
// group of a few observations is a short sum of logarithms of integer counts -- table look-ups instead of ~45 dependent
// fp64 instructions per logarithm.  Counts beyond the table take log_core.
constexpr uint32_t LOGTAB_SIZE = 1u << 20;
__device__ const double *g_logtab;

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

// Single out-of-line copies of the library transcendentals for the dense sweep kernel's rare paths
// (every inlined copy is instruction-cache traffic on a once-per-move code path).
__device__ __noinline__ double log_nl(double x) { return log(x); }
__device__ __noinline__ double exp_nl(double x) { return exp(x); }

// ---------------------------------------------------------------------------------
// Branch-free fp64 building blocks for the dense kernel's scoring.  The move's critical path is a chain of
// DEPENDENT fp64 instructions (DFMA issues back to back only every ~23 cycles on sm_100), so the score of a
// target cell is written as straight-line code whose six logarithms are independent chains the scheduler can
// interleave -- no calls, no special-case branches (arguments are positive, finite and >= 1 by construction).
// ---------------------------------------------------------------------------------
// 1/x for a normal positive x: MUFU.RCP64H seed and two Newton steps (error ~1 ulp)
__device__ __forceinline__ double rcp_nr(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
// ln(v) for a normal v >= 1.  Argument reduction v = 2^k m, m in [sqrt(1/2), sqrt(2)); with f = m - 1 and
// s = f / (2 + f):  ln(m) = f - f^2/2 + s (f^2/2 + R(s^2)),  R the degree-7 minimax polynomial of W. Kahan's
// classic scheme (as in 4.3BSD libm / fdlibm's log); error < 1 ulp, checked against mpmath in tests.
__device__ __forceinline__ double log_core(double v) {
    int hi = __double2hiint(v);
    const int lo = __double2loint(v);
    int k = (hi >> 20) - 1023;
    hi &= 0x000fffff;
    const int i = (hi + 0x95f64) & 0x100000;          // mantissa >= sqrt(2): use m / 2 and k + 1
    hi |= i ^ 0x3ff00000;
    k += i >> 20;
    const double f = __hiloint2double(hi, lo) - 1.0;
    const double s = f * rcp_nr(2.0 + f);
    const double dk = (double)k;
    const double z = s * s, w = z * z;
    const double t1 = w * fma(w, fma(w, 1.531383769920937332e-01, 2.222219843214978396e-01), 3.999999999940941908e-01);
    const double t2 = z * fma(w, fma(w, fma(w, 1.479819860511658591e-01, 1.818357216161805012e-01), 2.857142874366239149e-01), 6.666666666666735130e-01);
    const double R = t2 + t1;
    const double hfsq = 0.5 * f * f;
    return dk * 6.93147180369123816490e-01 - ((hfsq - fma(s, hfsq + R, dk * 1.90821492927058770002e-10)) - f);
}
// log((a+m)! / a!) against a table lf[LFS] of lgamma(j+1), branch-free per lane (selects only; one warp-uniform
// early exit); must be called by all 32 lanes of a converged warp (lanes without work pass m = 0):
//   a + m < LFS          lf[a+m] - lf[a]
//   a < LFS <= a + m     Stirling(a+m+1) - lf[a]
//   a >= LFS             (y+m-1/2) log1p(m/y) + m (ln y - 1) + corr(y+m) - corr(y),  y = a + 1   (no cancellation)
// log1p(x) = ln(u) + (x - (u - 1)) / u with u = 1 + x.
template <int LFS>
__device__ __forceinline__ double lfact_diff_bf(const double *lf, uint32_t a, uint32_t m) {
    const uint32_t am = a + m;
    const bool small = am < (uint32_t)LFS, big = a >= (uint32_t)LFS;
    const double lfa = lf[min(a, (uint32_t)LFS - 1u)];
    const double tab = lf[min(am, (uint32_t)LFS - 1u)] - lfa;
    // the callers' warps are converged: when no lane is beyond the table the six logarithms are skipped
    if (!__any_sync(0xffffffffu, m != 0u && !small)) return m == 0u ? 0.0 : tab;
    const double y = (double)a + 1.0, dm = (double)m, ym = y + dm;
    const double ry = rcp_nr(y), rym = rcp_nr(ym);
    const double x = dm * ry, u = 1.0 + x;
    const double L = log_core(big ? y : ym);
    const double P = log_core(u) + (x - (u - 1.0)) * rcp_nr(u);
    const double c2 = stirling_corr_r(rym);
    const double rb = fma(ym - 0.5, P, dm * (L - 1.0)) + (c2 - stirling_corr_r(ry));
    const double rm = fma(ym - 0.5, L, -ym) + 0.91893853320467274178 + c2 - lfa;
    return m == 0u ? 0.0 : small ? tab : big ? rb : rm;
}
// S(N+n, s+sg) - S(N, s) = D(s, sg) + D(N-s, n-sg) - D(N+1, n)   (see score_delta below), branch-free.
// The three terms share ONE warp-uniform early exit (no lane of the warp beyond the table), so that beyond it their
// logarithm chains are one straight-line block the scheduler interleaves (three separate exits serialise them: each chain is
// ~40 dependent fp64 instructions at ~23 cycles).  Term by term the values are those of lfact_diff_bf.
template <int LFS>
__device__ __forceinline__ double lfact_diff_heavy(const double *lf, uint32_t a, uint32_t m, double tab, double lfa) {
    const uint32_t am = a + m;
    const bool small = am < (uint32_t)LFS, big = a >= (uint32_t)LFS;
    const double y = (double)a + 1.0, dm = (double)m, ym = y + dm;
    const double ry = rcp_nr(y), rym = rcp_nr(ym);
    const double x = dm * ry, u = 1.0 + x;
    const double L = log_core(big ? y : ym);
    const double P = log_core(u) + (x - (u - 1.0)) * rcp_nr(u);
    const double c2 = stirling_corr_r(rym);
    const double rb = fma(ym - 0.5, P, dm * (L - 1.0)) + (c2 - stirling_corr_r(ry));
    const double rm = fma(ym - 0.5, L, -ym) + 0.91893853320467274178 + c2 - lfa;
    return m == 0u ? 0.0 : small ? tab : big ? rb : rm;
}
template <int LFS>
__device__ __forceinline__ double score_delta_bf(const double *lf, uint32_t N, uint32_t sv, uint32_t gn, uint32_t gs) {
    const uint32_t a1 = sv, m1 = gs, a2 = N - sv, m2 = gn - gs, a3 = N + 1u, m3 = gn;
    const double lfa1 = lf[min(a1, (uint32_t)LFS - 1u)], lfa2 = lf[min(a2, (uint32_t)LFS - 1u)], lfa3 = lf[min(a3, (uint32_t)LFS - 1u)];
    const double tab1 = lf[min(a1 + m1, (uint32_t)LFS - 1u)] - lfa1, tab2 = lf[min(a2 + m2, (uint32_t)LFS - 1u)] - lfa2,
                 tab3 = lf[min(a3 + m3, (uint32_t)LFS - 1u)] - lfa3;
    const bool beyond = (m1 != 0u && a1 + m1 >= (uint32_t)LFS) || (m2 != 0u && a2 + m2 >= (uint32_t)LFS) || (m3 != 0u && a3 + m3 >= (uint32_t)LFS);
    if (!__any_sync(0xffffffffu, beyond)) return ((m1 == 0u ? 0.0 : tab1) + (m2 == 0u ? 0.0 : tab2)) - (m3 == 0u ? 0.0 : tab3);
    const double d1 = lfact_diff_heavy<LFS>(lf, a1, m1, tab1, lfa1), d2 = lfact_diff_heavy<LFS>(lf, a2, m2, tab2, lfa2),
                 d3 = lfact_diff_heavy<LFS>(lf, a3, m3, tab3, lfa3);
    return (d1 + d2) - d3;
}

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
