// Micro-benchmark: does a DFMA (half-rate FP64 pipe, 16 lanes/SMSP) leave the issue slot of the
// following cycle free for another pipe?  Times per-SM cycles for N DFMA alone, N*R FFMA/IMAD alone,
// and both interleaved.  nvcc -arch=sm_100a -O3 -o issue_fp64 issue_fp64.cu && ./issue_fp64
#include <cstdio>
#include <cuda_runtime.h>

template <int NF64, int NF32, int NINT, int NLDS>
__global__ void __launch_bounds__(1024, 1) k(double *out, long long *cycles, int iters)
{
    __shared__ float sm[4096];
    double d[8]; float f[8]; int n[8];
    for (int i = 0; i < 8; ++i) { d[i] = 1.0 + threadIdx.x * 1e-9 + i; f[i] = 1.0f + threadIdx.x * 1e-6f + i; n[i] = threadIdx.x + i; }
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
    __syncthreads();
    const double a = 1.0000001, b = 1e-9; const float fa = 1.0001f, fb = 1e-6f;
    float ls = 0.f; int idx = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int k = 0; k < NF64; ++k) d[(r + k) & 7] = fma(d[(r + k) & 7], a, b);
#pragma unroll
            for (int k = 0; k < NF32; ++k) f[(r + k) & 7] = fmaf(f[(r + k) & 7], fa, fb);
#pragma unroll
            for (int k = 0; k < NINT; ++k) n[(r + k) & 7] = n[(r + k) & 7] * 3 + 7;
#pragma unroll
            for (int k = 0; k < NLDS; ++k) { ls += sm[(idx + 32 * (r + k)) & 4095]; }
        }
    }
    long long t1 = clock64();
    double s = ls; for (int i = 0; i < 8; ++i) s += d[i] + f[i] + n[i];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int A, int B, int C, int D>
void run(const char *name, int iters)
{
    double *out; long long *cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 148 * 8);
    k<A, B, C, D><<<148, 1024>>>(out, cyc, 10);
    k<A, B, C, D><<<148, 1024>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    // per SMSP: 8 warps; instructions per warp per iter = 8*(A+B+C+D)
    const double per_group = avg / iters / 8.0 / 8.0;   // cycles per (one r-step of one warp) on its SMSP, i.e. per A+B+C+D warp-instructions
    printf("%-28s f64=%d f32=%d int=%d lds=%d : %.2f cycles per group of %d warp-instructions (%.2f cyc/inst)\n", name, A, B, C, D,
           per_group, A + B + C + D, per_group / (A + B + C + D));
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    const int it = 2000;
    run<1, 0, 0, 0>("DFMA only", it);
    run<0, 1, 0, 0>("FFMA only", it);
    run<0, 0, 1, 0>("IMAD only", it);
    run<1, 1, 0, 0>("DFMA + FFMA", it);
    run<1, 0, 1, 0>("DFMA + IMAD", it);
    run<1, 1, 1, 0>("DFMA + FFMA + IMAD", it);
    run<1, 2, 0, 0>("DFMA + 2 FFMA", it);
    run<1, 2, 1, 0>("DFMA + 2 FFMA + IMAD", it);
    run<2, 1, 1, 0>("2 DFMA + FFMA + IMAD", it);
    run<0, 0, 0, 1>("LDS only", it);
    run<1, 0, 0, 1>("DFMA + LDS", it);
    run<2, 1, 1, 1>("2 DFMA + FFMA + IMAD + LDS", it);
    return 0;
}
