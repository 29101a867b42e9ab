// mmc_common.cuh — shared device helpers and the device-resident state of one handle.
// Target: sm_100a only (B200). FP64 throughout: the reference is Float64 end to end and the
// parity bound is 1e-10 relative.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MMC_MAX_SITES 16   // sites per molecule supported on the per-move path
#define MMC_MAX_TYPES 8    // LJ atom types (vdwTable is nt x nt)
#define MMC_MAX_NK 8       // largest |k| component the k-space kernels are sized for
#define MMC_NSCAL 8        // scalar slots at the head of the partial-sum vector

// ---------------------------------------------------------------------------------------
// Device-resident state (HBM layout).
//   site[n_sites]  double4 {x, y, z, q}   32 B, one LDG.128 pair per site, coalesced
//   com[n_mol]     double4 {x, y, z, -}
//   mol[n_mol]     int2    {first site (0-based), site count}
//   atype[n_sites] int     0-based LJ type
//   eps/sig        nt*nt   column-major like vdwTable.ϵᵢⱼ / σᵢⱼ
//   kvec[NK]       int4    {kx, ky, kz, -};  cfac[NK];  rhok[2][NK] double2 = ρ(k) Old/New
// Everything for 256k SPC/E molecules is 768k*32 B + 256k*40 B ≈ 35 MB: L2-resident (126 MB).
// ---------------------------------------------------------------------------------------
struct DevSystem {
    int n_mol, n_sites, max_sites, n_types;
    int uni;                 // > 0: uniform topology, molecule m owns sites [m*uni, (m+1)*uni)
    double4 *site;
    double4 *com;
    int2 *mol;
    int *atype;
    double eps[MMC_MAX_TYPES * MMC_MAX_TYPES];
    double sig[MMC_MAX_TYPES * MMC_MAX_TYPES];
    double box, rc_lj, rc_qq;
    double kappa, factor;
    int nk, nkvecs;
    int4 *kvec;
    double *cfac;
    double2 *rhok[2];
};

// bits of the upload kernels' validation word (kernels_upload.cuh, kernels_peer.cuh)
enum { REPACK_BAD_ATYPE = 1, REPACK_BAD_RANGE = 2, REPACK_TOO_MANY_SITES = 4, REPACK_COM_OUTSIDE = 8, REPACK_PEER_TIMEOUT = 16 };

// LJ-active site pair (a, b) of the uniform molecule (eps_ab > 0.001, Ewald/energy.jl:270)
struct LJActive { int a, b; double eps, sig; };

struct DevAtoms {
    int n;
    double4 *r;      // {x, y, z, -}
    double2 *es;     // {eps_j, sig_j}
    double box, rc;
};

// Ewald/boundaries.jl:8-14 vector1D(c1, c2, L), branch-free.
//   c1 <  c2: (c2-c1) < (c1-c2+L) ? (c2-c1) : (c2-c1-L)
//   c1 >= c2: (c1-c2) < (c2-c1+L) ? (c2-c1) : (c2-c1+L)
// With d = c2-c1 both tests read |d| < (L - |d|) (c1-c2 == -d exactly), so the selected
// value has the same bits as the reference's.
__device__ __forceinline__ double min_image(double c1, double c2, double L)
{
    const double d = c2 - c1;
    const double ad = fabs(d);
    const double w = (d > 0.0) ? d - L : d + L;
    return (ad < (L - ad)) ? d : w;
}

// L2 (cache-global) load of a double4 written by another CTA of the same launch
__device__ __forceinline__ double4 ldcg4(const double4 *p)
{
    const double2 a = __ldcg(reinterpret_cast<const double2 *>(p));
    const double2 b = __ldcg(reinterpret_cast<const double2 *>(p) + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block sum of NV values per thread; result valid in thread 0.
// scratch must hold NV * (BLOCK/32) doubles.
template <int NV, int BLOCK>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = BLOCK / 32;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i] = warp_sum(v[i]);
        if (lane == 0) scratch[i * NW + warp] = v[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double s = 0.0;
            for (int w = 0; w < NW; ++w) s += scratch[i * NW + w];
            v[i] = s;
        }
    }
    __syncthreads();
}

// Deterministic block sum whose result is valid in EVERY thread.
template <int NV, int BLOCK>
__device__ __forceinline__ void block_sum_all(double (&v)[NV], double *scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = BLOCK / 32;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i] = warp_sum(v[i]);
        if (lane == 0) scratch[i * NW + warp] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = 0.0;
        for (int w = 0; w < NW; ++w) s += scratch[i * NW + w];
        v[i] = s;
    }
}

struct cplx { double re, im; };
// Julia Base: Complex(a.re*b.re - a.im*b.im, a.re*b.im + a.im*b.re)
__device__ __forceinline__ cplx cmul(cplx a, cplx b)
{
    cplx c;
    c.re = a.re * b.re - a.im * b.im;
    c.im = a.re * b.im + a.im * b.re;
    return c;
}
__device__ __forceinline__ cplx cconj_if(cplx a, bool neg)
{
    cplx c;
    c.re = a.re;
    c.im = neg ? -a.im : a.im;
    return c;
}

// LJ 12-6 term of Ewald/energy.jl:270-282 for one site pair; rij = COM minimum-image vector.
__device__ __forceinline__ void lj_pair(double eps, double sig, double r2, double dx, double dy,
                                        double dz, double rijx, double rijy, double rijz,
                                        double &pot, double &vir)
{
    const double s2 = sig * sig / r2;
    const double s6 = s2 * s2 * s2;
    const double s12 = s6 * s6;
    pot += eps * (s12 - s6);
    const double virab = eps * (2.0 * s12 - s6);
    const double fx = dx * virab * s2, fy = dy * virab * s2, fz = dz * virab * s2;
    vir += rijx * fx + rijy * fy + rijz * fz;
}
