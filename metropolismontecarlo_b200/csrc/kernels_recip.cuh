// kernels_recip.cuh — full structure-factor rebuild, RecipLong (Ewald/ewalds.jl:538-604),
// SURVEY.md §8a row a6.
//
//   ρ(k) = Σ_l q_l e^{ikx x_l} e^{iky y_l} e^{ikz z_l},   E = Σ_k cfac_k |ρ(k)|²
//
// Decomposition: sites are split over CTAs (and over ranks when sharded); inside a CTA the
// sites stream through shared memory in sub-chunks whose per-site tables (q·e^{ikx}, e^{iky},
// e^{ikz}, k = 0..nk; negative k by conjugation, ewalds.jl:566-567,581-582) are built by the
// same recurrence the reference uses; every thread owns KPT k-vectors and keeps their complex
// accumulators in registers.  Per-CTA partial sums go to HBM and are folded in CTA order by
// k_rhok_reduce, so the result is deterministic.  FP64 accumulation throughout.
#pragma once
#include "mmc_common.cuh"

#define RHOK_BLOCK 384
#define RHOK_SITES 32

struct RhokArgs {
    const double4 *site;     // {x, y, z, q}
    int s_begin, s_end;      // this rank's site range
    int per_block;           // sites per CTA (multiple of RHOK_SITES)
    int nk, nkvecs;
    const int4 *kvec;
    double box;
    double2 *partial;        // [gridDim.x][nkvecs]
};

template <int KPT>
__global__ void __launch_bounds__(RHOK_BLOCK) k_rhok_partial(RhokArgs A)
{
    __shared__ cplx s_e[RHOK_SITES][3][MMC_MAX_NK + 1];
    const int tid = threadIdx.x;
    const int c0 = A.s_begin + blockIdx.x * A.per_block;
    const int c1 = min(A.s_end, c0 + A.per_block);
    const double twopi = 2.0 * 3.141592653589793;
    int kx[KPT], ky[KPT], kz[KPT];
    bool ny[KPT], nz[KPT];
    double are[KPT], aim[KPT];
#pragma unroll
    for (int u = 0; u < KPT; ++u) {
        const int k = tid + u * RHOK_BLOCK;
        int4 kv = make_int4(0, 0, 0, 0);
        if (k < A.nkvecs) kv = A.kvec[k];
        kx[u] = kv.x; ky[u] = abs(kv.y); kz[u] = abs(kv.z);
        ny[u] = kv.y < 0; nz[u] = kv.z < 0;
        are[u] = 0.0; aim[u] = 0.0;
    }
    for (int base = c0; base < c1; base += RHOK_SITES) {
        __syncthreads();
        if (tid < RHOK_SITES * 3) {
            const int l = tid / 3, d = tid - 3 * l;
            double x = 0.0, q = 0.0;
            if (base + l < c1) {
                const double4 s = A.site[base + l];
                x = d == 0 ? s.x : (d == 1 ? s.y : s.z);
                q = s.w;
            }
            const double sc = (d == 0) ? q : 1.0;      // charge folded into the x table: (q*ex)*ey*ez
            cplx e1;
            sincos(twopi * x / A.box, &e1.im, &e1.re);  // ewalds.jl:561-564
            cplx e; e.re = 1.0; e.im = 0.0;
            cplx st; st.re = sc * e.re; st.im = sc * e.im;
            s_e[l][d][0] = st;
            e = e1;
            for (int k = 1; k <= A.nk; ++k) {
                st.re = sc * e.re; st.im = sc * e.im;
                s_e[l][d][k] = st;
                e = cmul(e, e1);                         // ewalds.jl:573-575
            }
        }
        __syncthreads();
#pragma unroll 4
        for (int l = 0; l < RHOK_SITES; ++l) {
#pragma unroll
            for (int u = 0; u < KPT; ++u) {
                const cplx t = cmul(cmul(s_e[l][0][kx[u]], cconj_if(s_e[l][1][ky[u]], ny[u])),
                                    cconj_if(s_e[l][2][kz[u]], nz[u]));
                are[u] += t.re;
                aim[u] += t.im;
            }
        }
    }
#pragma unroll
    for (int u = 0; u < KPT; ++u) {
        const int k = tid + u * RHOK_BLOCK;
        if (k < A.nkvecs) A.partial[(size_t)blockIdx.x * A.nkvecs + k] = make_double2(are[u], aim[u]);
    }
}

// ρ(k) = Σ_b partial[b][k] in CTA order → out[k] (a slot of the partial-sum vector)
__global__ void k_rhok_reduce(const double2 *partial, int nb, int nkvecs, double2 *out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nkvecs) return;
    double re = 0.0, im = 0.0;
    for (int b = 0; b < nb; ++b) {
        const double2 p = partial[(size_t)b * nkvecs + k];
        re += p.x; im += p.y;
    }
    out[k] = make_double2(re, im);
}

// E = Σ_k cfac_k |ρ(k)|² (un-scaled, ewalds.jl:599) and ρ(k) stored to both buffers (:600-601)
#define RHOKE_BLOCK 256
__global__ void __launch_bounds__(RHOKE_BLOCK)
k_rhok_energy(const double2 *rho, const double *cfac, int nkvecs, double2 *dst0, double2 *dst1,
              double *energy_out)
{
    __shared__ double s_red[RHOKE_BLOCK / 32];
    double acc[1] = {0.0};
    for (int k = threadIdx.x; k < nkvecs; k += RHOKE_BLOCK) {
        const double2 s = rho[k];
        acc[0] += cfac[k] * (s.x * s.x + s.y * s.y);
        if (dst0) dst0[k] = s;
        if (dst1) dst1[k] = s;
    }
    block_sum<1, RHOKE_BLOCK>(acc, s_red);
    if (threadIdx.x == 0) *energy_out = acc[0];
}
