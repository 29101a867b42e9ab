// kernels_upload.cuh — device-side conversion of the reference's Julia array layouts
// (xyz-interleaved coords, Float64 charges, Int64 1-based atype / firstAtom /
// lastAtom, COM) into the engine's HBM layout (double4 {x,y,z,q} per site, double4 COM, int2
// {first, count}, int type), with the argument validation folded in.  The raw arrays are DMA'd
// as they are (no host-side repacking pass); this kernel touches every byte once.
#pragma once
#include "mmc_common.cuh"

struct RepackArgs {
    const double *coords, *charge, *com;
    const long long *atype, *first_atom, *last_atom;
    int n_mol, n_sites, n_types;
    double box;
    double4 *site, *dcom;
    int2 *mol;
    int *dtype;
    int *info;     // [0] error bits, [1] max sites per molecule, [2] non-uniform flag, [3] charges differ between molecules
};


static __global__ void k_repack(RepackArgs A)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const long long f0 = A.first_atom[0], l0 = A.last_atom[0];
    const int US = (int)(l0 - f0 + 1);
    if (t < A.n_sites) {
        const long long ty = A.atype[t];
        if (ty < 1 || ty > A.n_types) atomicOr(&A.info[0], REPACK_BAD_ATYPE);
        A.dtype[t] = (int)(ty - 1);
        A.site[t] = make_double4(A.coords[3 * (size_t)t], A.coords[3 * (size_t)t + 1], A.coords[3 * (size_t)t + 2], A.charge[t]);
        if (US > 0 && ty != A.atype[t % US]) atomicOr(&A.info[2], 1);   // type sequence differs from molecule 1
        if (US > 0 && A.charge[t] != A.charge[t % US]) atomicOr(&A.info[3], 1);
    }
    if (t < A.n_mol) {
        const long long f = A.first_atom[t], l = A.last_atom[t];
        int cnt = 0;
        if (f < 1 || l < f || l > A.n_sites) atomicOr(&A.info[0], REPACK_BAD_RANGE);
        else {
            cnt = (int)(l - f + 1);
            if (cnt > MMC_MAX_SITES) atomicOr(&A.info[0], REPACK_TOO_MANY_SITES);
        }
        A.mol[t] = make_int2((int)(f - 1), cnt);
        atomicMax(&A.info[1], cnt);
        if (cnt != US || f - 1 != (long long)t * US) atomicOr(&A.info[2], 1);
        const double x = A.com[3 * (size_t)t], y = A.com[3 * (size_t)t + 1], z = A.com[3 * (size_t)t + 2];
        if (!(x >= 0.0 && x <= A.box && y >= 0.0 && y <= A.box && z >= 0.0 && z <= A.box))
            atomicOr(&A.info[0], REPACK_COM_OUTSIDE);
        A.dcom[t] = make_double4(x, y, z, 0.0);
    }
}

// positions only (mmc_upload_positions): coords and COM of an already uploaded system, charges/topology untouched
static __global__ void k_repack_positions(const double *coords, const double *com, int n_mol, int n_sites, double box,
                                   double4 *site, double4 *dcom, int *info)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_sites) {
        double4 s = site[t];
        s.x = coords[3 * (size_t)t]; s.y = coords[3 * (size_t)t + 1]; s.z = coords[3 * (size_t)t + 2];
        site[t] = s;
    }
    if (t < n_mol) {
        const double x = com[3 * (size_t)t], y = com[3 * (size_t)t + 1], z = com[3 * (size_t)t + 2];
        if (!(x >= 0.0 && x <= box && y >= 0.0 && y <= box && z >= 0.0 && z <= box)) atomicOr(&info[0], REPACK_COM_OUTSIDE);
        dcom[t] = make_double4(x, y, z, 0.0);
    }
}

// the two halves of k_repack_positions, for the pipelined end-to-end path (mmc_potential_host): COMs first (binning
// needs nothing else), sites chunk by chunk as their copies land
static __global__ void k_repack_com(const double *com, int n_mol, double box, double4 *dcom, int *info)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_mol) return;
    const double x = com[3 * (size_t)t], y = com[3 * (size_t)t + 1], z = com[3 * (size_t)t + 2];
    if (!(x >= 0.0 && x <= box && y >= 0.0 && y <= box && z >= 0.0 && z <= box)) atomicOr(&info[0], REPACK_COM_OUTSIDE);
    dcom[t] = make_double4(x, y, z, 0.0);
}

static __global__ void k_repack_sites(const double *coords, int s0, int s1, double4 *site)
{
    const int t = s0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= s1) return;
    double4 s = site[t];
    s.x = coords[3 * (size_t)t]; s.y = coords[3 * (size_t)t + 1]; s.z = coords[3 * (size_t)t + 2];
    site[t] = s;
}

// sites of the molecule blocks (256 molecules each) a rank of a domain-decomposed evaluation has received
static __global__ void k_repack_sites_blocks(const double *coords, const unsigned char *__restrict__ need, int US, int n_sites, double4 *site)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_sites || !need[(t / US) >> 8]) return;
    double4 s = site[t];
    s.x = coords[3 * (size_t)t]; s.y = coords[3 * (size_t)t + 1]; s.z = coords[3 * (size_t)t + 2];
    site[t] = s;
}

// Σq and Σq² in two deterministic stages (EwaldSelf ewalds.jl:829-833, Wolf constants energy.jl:924-934)
static __global__ void __launch_bounds__(256) k_charge_partial(const double4 *site, int n, double2 *part)
{
    __shared__ double s_red[2 * 8];
    double acc[2] = {0.0, 0.0};
    for (int l = blockIdx.x * 256 + threadIdx.x; l < n; l += gridDim.x * 256) {
        const double q = site[l].w;
        acc[0] += q; acc[1] += q * q;
    }
    block_sum<2, 256>(acc, s_red);
    if (threadIdx.x == 0) part[blockIdx.x] = make_double2(acc[0], acc[1]);
}

static __global__ void __launch_bounds__(256) k_charge_final(const double2 *part, int nb, double *out)
{
    __shared__ double s_red[2 * 8];
    double acc[2] = {0.0, 0.0};
    for (int l = threadIdx.x; l < nb; l += 256) { acc[0] += part[l].x; acc[1] += part[l].y; }
    block_sum<2, 256>(acc, s_red);
    if (threadIdx.x == 0) { out[0] = acc[0]; out[1] = acc[1]; }
}

// mmc_potential_host, windowed pair evaluation: need[w] = the last site chunk that holds a molecule of window w's layers
// (home layers [ncd·w/nwin, ncd·(w+1)/nwin) plus the layer above them, periodic).  Chunk c = sites [n_sites·c/n_chunks, n_sites·(c+1)/n_chunks).
static __global__ void k_window_need(const int *__restrict__ cell_of, int n_mol, int S, int ncd, int nwin, int n_sites, int n_chunks, int *need)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_mol) return;
    const int z = cell_of[m] / (ncd * ncd);
    const long long last_site = (long long)m * S + S - 1;
    int c = (int)(last_site * n_chunks / n_sites);
    while (c + 1 < n_chunks && (long long)n_sites * (c + 1) / n_chunks <= last_site) ++c;
    while (c > 0 && (long long)n_sites * c / n_chunks > last_site) --c;
    for (int w = 0; w < nwin; ++w) {
        const int zlo = (int)((long long)ncd * w / nwin), zhi = (int)((long long)ncd * (w + 1) / nwin);
        if ((z >= zlo && z < zhi) || z == zhi % ncd) atomicMax(&need[w], c);
    }
}

// Intramolecular Ewald correction (opt-in, not in the reference): Σ_mol Σ_{a<b} q_a q_b erf(κ r_ab)/r_ab, un-scaled, in two
// deterministic stages like the charge sums.  r_ab by the reference's minimum image in the RESIDENT box (a rigid shift of the
// molecule — the volume scaling of a trial — does not change it).
static __global__ void __launch_bounds__(256) k_intra_partial(const double4 *site, const int2 *mol, int n_mol, double kappa, double box, double *part)
{
    __shared__ double s_red[8];
    double acc[1] = {0.0};
    for (int m = blockIdx.x * 256 + threadIdx.x; m < n_mol; m += gridDim.x * 256) {
        const int2 mi = mol[m];
        double e = 0.0;
        for (int a = 0; a < mi.y; ++a) {
            const double4 sa = site[mi.x + a];
            for (int b = a + 1; b < mi.y; ++b) {
                const double4 sb = site[mi.x + b];
                const double dx = min_image(sa.x, sb.x, box), dy = min_image(sa.y, sb.y, box), dz = min_image(sa.z, sb.z, box);
                const double r = sqrt(dx * dx + dy * dy + dz * dz);
                e += sa.w * sb.w * erf(kappa * r) / r;
            }
        }
        acc[0] += e;
    }
    block_sum<1, 256>(acc, s_red);
    if (threadIdx.x == 0) part[blockIdx.x] = acc[0];
}

static __global__ void __launch_bounds__(256) k_intra_final(const double *part, int nb, double *out)
{
    __shared__ double s_red[8];
    double acc[1] = {0.0};
    for (int l = threadIdx.x; l < nb; l += 256) acc[0] += part[l];
    block_sum<1, 256>(acc, s_red);
    if (threadIdx.x == 0) out[0] = acc[0];
}

// {m·S, S} for every molecule: the molecule table of a copy padded to S slots per molecule (mixed topologies)
static __global__ void k_mol_packed(int2 *mol, int n_mol, int S)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m < n_mol) mol[m] = make_int2(m * S, S);
}
