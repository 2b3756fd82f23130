// Accuracy of rcp.approx.ftz.f64 + k Newton steps, and of log_core / lfact_diff_bf against the host's libm.
#include <cstdio>
#include <cmath>
#include <vector>
#include "device_math.cuh"
using namespace hirm;
__global__ void k(const double *x, double *o0, double *o1, double *o2, double *o3, double *ol, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
    double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x[i]));
    o0[i] = r;
    double e = fma(-x[i], r, 1.0); r = fma(r, e, r); o1[i] = r;
    e = fma(-x[i], r, 1.0); r = fma(r, e, r); o2[i] = r;
    e = fma(-x[i], r, 1.0); r = fma(r, e, r); o3[i] = r;
    ol[i] = log_core(x[i]);
}
int main() {
    const int n = 1 << 16;
    std::vector<double> x(n);
    for (int i = 0; i < n; i++) x[i] = 1.0 + (double)(rand() % 1000000) * pow(10.0, (rand() % 9) - 3);
    double *dx, *d[5]; cudaMalloc(&dx, n * 8); for (auto &p : d) cudaMalloc(&p, n * 8);
    cudaMemcpy(dx, x.data(), n * 8, cudaMemcpyHostToDevice);
    k<<<n / 256, 256>>>(dx, d[0], d[1], d[2], d[3], d[4], n);
    std::vector<double> o(n);
    const char *names[5] = {"rcp.approx.ftz.f64", "+1 Newton", "+2 Newton", "+3 Newton", "log_core vs libm log"};
    for (int j = 0; j < 5; j++) {
        cudaMemcpy(o.data(), d[j], n * 8, cudaMemcpyDeviceToHost);
        double mx = 0;
        for (int i = 0; i < n; i++) { double ref = j < 4 ? 1.0 / x[i] : log(x[i]); double er = fabs(o[i] - ref) / fabs(ref); if (ref != 0 && er > mx) mx = er; }
        printf("%-24s max relative error %.3e\n", names[j], mx);
    }
    return 0;
}
