// Microbenchmark: cycles per dependent call of the dense kernel's cell score (score_delta_bf) and its parts, one
// warp / several warps.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../hierarchical-irm_b200/csrc lfact_latency.cu
#include <cstdio>
#include <cmath>
#include <vector>
#include "device_math.cuh"
using namespace hirm;
#define REP 128
__device__ __forceinline__ long long clk2() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory"); return t; }
template <int OP> __global__ void k(double *out, long long *cyc, uint32_t a0, uint32_t m0) {
    __shared__ double lf[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lf[i] = lgamma(i + 1.0);
    __syncthreads();
    uint32_t a = a0 + threadIdx.x * 977u, m = m0 + (threadIdx.x & 7);
    double acc = 0;
    long long t0 = clk2();
#pragma unroll 1
    for (int i = 0; i < REP; i++) {
        double v;
        if (OP == 0) v = score_delta_bf<256>(lf, a, a / 3u, m, m / 2u);
        if (OP == 1) v = lfact_diff_bf<256>(lf, a, m);
        if (OP == 2) v = log_core((double)a + 1.0);
        if (OP == 3) v = rcp_nr((double)a + 1.0);
        if (OP == 4) v = exp_nl(-(double)(a & 1023u) * 1e-3);
        acc += v;
        a += (uint32_t)(v * 1e-9) + 1u;      // dependence
    }
    long long t1 = clk2();
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int OP> void run(const char *name, uint32_t a0, uint32_t m0, int threads) {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
    k<OP><<<1, threads>>>(out, cyc, a0, m0); cudaDeviceSynchronize();
    k<OP><<<1, threads>>>(out, cyc, a0, m0); cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-34s threads=%4d a=%u m=%u: %8.1f cycles per call (%s)\n", name, threads, a0, m0, (double)h / REP, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    std::vector<double> lfh(LF_SIZE);
    for (int m = 0; m < LF_SIZE; m++) lfh[m] = lgamma(m + 1.0);
    cudaMemcpyToSymbol(g_logfact, lfh.data(), sizeof(double) * LF_SIZE);
    for (int threads : {32, 64, 256, 512}) {
        run<0>("score_delta_bf (cell, 3 terms)", 100000000u, 5000u, threads);
        run<1>("lfact_diff_bf (1 term)", 100000000u, 5000u, threads);
        run<2>("log_core", 100000000u, 5000u, threads);
        run<3>("rcp_nr", 100000000u, 5000u, threads);
        run<4>("exp_nl", 100000000u, 5000u, threads);
    }
    return 0;
}
