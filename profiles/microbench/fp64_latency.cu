// Microbenchmark: dependent-issue latency (SM cycles, one warp / one lane active) of the fp64
// operations on the move's critical path.  nvcc -arch=sm_100a -O3 fp64_latency.cu && ./a.out
#include <cstdio>
#include <cuda_runtime.h>
#define REP 256
template <int OP> __device__ double step(double x, double c) {
    if (OP == 0) return fma(x, c, 1e-9);
    if (OP == 1) return log(x) + c;
    if (OP == 2) return exp(x * 1e-3) ;
    if (OP == 3) return log1p(x * 1e-3) + c;
    if (OP == 4) return 1.0 / x + c;
    if (OP == 5) { double r = (double)__frcp_rn((float)x); r = r * (2.0 - x * r); r = r * (2.0 - x * r); return r + c; }
    if (OP == 6) return (double)((float)x * 1.0001f) + c;      // fp32 fma + conversions
    if (OP == 7) return (double)__logf((float)x) + c;
    if (OP == 8) return x + c;
    if (OP == 9) return lgamma(x) * 1e-3 + c;
    return x;
}
template <int OP> __global__ void k(double *out, long long *cyc, double x0, double c, int active_lanes) {
    double x = x0 + threadIdx.x * 1e-6;
    long long t0 = 0, t1 = 0;
    if ((threadIdx.x & 31) < active_lanes) {
        t0 = clock64();
#pragma unroll 1
        for (int i = 0; i < REP; i++) x = step<OP>(x, c);
        t1 = clock64();
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void smem_rmw(long long *cyc, int *out, int stride) {
    __shared__ unsigned int h[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) h[i] = 0;
    __syncthreads();
    unsigned idx = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < REP; i++) { h[idx] += i; idx = (idx + stride) & 4095; }
    long long t1 = clock64();
    __syncthreads();
    out[threadIdx.x] = h[threadIdx.x];
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void syncs(long long *cyc) {
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < REP; i++) __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int OP> void run(const char *name, double x0, double c, int threads, int lanes) {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 1024 * 8 * 4); cudaMalloc(&cyc, 8);
    k<OP><<<1, threads>>>(out, cyc, x0, c, lanes); cudaDeviceSynchronize();
    k<OP><<<1, threads>>>(out, cyc, x0, c, lanes); cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s threads=%4d lanes=%2d : %7.1f cycles per dependent op\n", name, threads, lanes, (double)h / REP);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int lanes : {1, 32}) {
        run<0>("DFMA", 1.0, 0.999, 32, lanes);
        run<8>("DADD", 1.0, 0.5, 32, lanes);
        run<1>("log(double)", 5.0, 3.0, 32, lanes);
        run<2>("exp(double)", 1.0, 0.0, 32, lanes);
        run<3>("log1p(double)", 5.0, 3.0, 32, lanes);
        run<4>("1.0/x (double)", 5.0, 3.0, 32, lanes);
        run<5>("fast_rcp (rcp+2 Newton)", 5.0, 3.0, 32, lanes);
        run<6>("fp32 mul + 2 cvt", 5.0, 0.0, 32, lanes);
        run<7>("__logf + 2 cvt", 5.0, 3.0, 32, lanes);
        run<9>("lgamma(double)", 50.0, 30.0, 32, lanes);
    }
    run<0>("DFMA 16 warps", 1.0, 0.999, 512, 32);
    run<1>("log 16 warps", 5.0, 3.0, 512, 32);
    long long *cyc, h; int *o; cudaMalloc(&cyc, 8); cudaMalloc(&o, 4096);
    for (int thr : {32, 512}) {
        smem_rmw<<<1, thr>>>(cyc, o, 32); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("smem RMW chain (LDS+IADD+STS) threads=%d: %.1f cycles per RMW\n", thr, (double)h / REP);
        syncs<<<1, thr>>>(cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("__syncthreads threads=%d: %.1f cycles\n", thr, (double)h / REP);
    }
    return 0;
}
