// Microbenchmark: cycles per dependent call of the dense kernel's lfact_diff_warp_s (one warp), and of its parts.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../hierarchical-irm_b200/csrc lfact_latency.cu -o lfact_latency
#include <cstdio>
#include <cmath>
#include "dense_sweep.cuh"
using namespace hirm;
#define REP 128
template <int OP> __global__ void k(double *out, long long *cyc, uint32_t a0, uint32_t m0, int nwarps_active) {
    __shared__ double lf[DN_LF];
    for (int i = threadIdx.x; i < DN_LF; i += blockDim.x) lf[i] = lgamma(i + 1.0);
    __syncthreads();
    uint32_t a = a0 + threadIdx.x * 977u, m = m0 + (threadIdx.x & 7);
    double acc = 0;
    long long t0 = clk();
#pragma unroll 1
    for (int i = 0; i < REP; i++) {
        double v;
        if (OP == 0) v = lfact_diff_warp_s(lf, a, m);
        if (OP == 1) v = log_nl((double)a + 1.0);
        if (OP == 2) v = log1p_nl((double)m * fast_rcp((double)a + 1.0));
        if (OP == 3) v = stirling_corr_r(fast_rcp((double)a + 1.0 + (double)m));
        if (OP == 4) v = exp_nl(-(double)m * 1e-3);
        acc += v;
        a += (uint32_t)(v * 1e-9) + 1u;      // dependence
    }
    long long t1 = clk();
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int OP> void run(const char *name, uint32_t a0, uint32_t m0, int threads) {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
    k<OP><<<1, threads>>>(out, cyc, a0, m0, 0); cudaDeviceSynchronize();
    k<OP><<<1, threads>>>(out, cyc, a0, m0, 0); cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-34s threads=%4d a=%u m=%u: %8.1f cycles per call (%s)\n", name, threads, a0, m0, (double)h / REP, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    std::vector<double> lfh(LF_SIZE);
    for (int m = 0; m < LF_SIZE; m++) lfh[m] = lgamma(m + 1.0);
    cudaMemcpyToSymbol(g_logfact, lfh.data(), sizeof(double) * LF_SIZE);
    for (int threads : {32, 64, 256}) {
        run<0>("lfact_diff_warp_s big", 100000000u, 5000u, threads);
        run<0>("lfact_diff_warp_s mid (a<256<=a+m)", 100u, 5000u, threads);
        run<0>("lfact_diff_warp_s small", 10u, 50u, threads);
        run<1>("log_nl", 100000000u, 5000u, threads);
        run<2>("rcp + log1p_nl", 100000000u, 5000u, threads);
        run<3>("rcp + corr", 100000000u, 5000u, threads);
        run<4>("exp_nl", 100000000u, 5000u, threads);
    }
    return 0;
}
