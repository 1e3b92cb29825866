// max ulp error of mlmcb::exp_bounded against libdevice exp (run on the GPU box)
#include "../multilevel_mc_sdes_b200/csrc/mlmcb_kernels.cuh"
#include <cstdio>
#include <cmath>
using namespace mlmcb;
__global__ void k(const __grid_constant__ LevelConsts C, int n, double lo, double hi, double* worst, double* worst_x) {
    double w = 0, wx = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double x = lo + (hi - lo) * ((i + 0.37) / n);
        const double a = exp_bounded(C, x), b = exp(x);
        const double ulp = fabs(a - b) / (b * 0x1p-52);
        if (ulp > w) { w = ulp; wx = x; }
    }
    worst[blockIdx.x * blockDim.x + threadIdx.x] = w;
    worst_x[blockIdx.x * blockDim.x + threadIdx.x] = wx;
}
int main() {
    LevelConsts C = {};
    C.ln2_hi = 6.93147180369123816490e-01; C.ln2_lo = 1.90821492927058770002e-10;
    { double f = 1.0; for (int kk = 2; kk <= 13; ++kk) { f *= (double)kk; C.ex[kk - 2] = 1.0 / f; } }
    const int T = 256 * 128;
    double *w, *wx;
    cudaMallocManaged(&w, T * 8); cudaMallocManaged(&wx, T * 8);
    const double ranges[][2] = {{-2, 2}, {-40, 40}, {-700, 700}, {-1e-3, 1e-3}, {-720, -690}, {690, 720}};
    for (auto& r : ranges) {
        k<<<256, 128>>>(C, 1 << 26, r[0], r[1], w, wx);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("cuda error\n"); return 1; }
        double m = 0, mx = 0;
        for (int i = 0; i < T; ++i) if (w[i] > m) { m = w[i]; mx = wx[i]; }
        printf("x in [%g, %g]: max error %.3f ulp at x = %.17g\n", r[0], r[1], m, mx);
    }
    return 0;
}
