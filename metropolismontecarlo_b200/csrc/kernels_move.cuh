// kernels_move.cuh — the per-trial-move kernels.
//
// One launch evaluates, for molecule i:
//   cfg 0: LJ_poly_ΔU + EwaldReal against the resident state        (Ewald/main.jl:491,501)
//   cfg 1: the same with molecule i at its trial position            (Ewald/main.jl:557,566)
//   recip: RecipMove's ρ(k) delta update and ΔE                      (Ewald/main.jl:581)
// and the last CTA to finish folds the per-CTA partials in a fixed order and writes the
// scalars straight into a mapped pinned host slot (one PCIe write, no memcpy, no second launch).
//
// Work decomposition per pair CTA: (1) one thread per partner molecule does the COM gate
// (|COM_ij|² < rc², strict, Ewald/energy.jl:250 / ewalds.jl:337), survivors are compacted in
// index order into shared memory; (2) the n_in × n_a × n_b site pairs are spread over all
// threads, so the ~10³ erfc evaluations of a water move run ≈1 per thread instead of 9 in series.
// These launches are latency-bound by construction (≈2×10⁵ flop per move at N=750).
#pragma once
#include "mmc_common.cuh"

#define MOVE_BLOCK 128

struct MoveArgs {
    int i;            // 0-based molecule
    int n_cfg;        // 1: resident state only, 2: resident + trial
    int tiles;        // pair CTAs per cfg
    int recip_blocks; // CTAs doing the ρ(k) delta (0: none)
    int want_lj, want_qq;
    int ignore_overlap; // rows for the overlap fix-up of potential(): skip offending site pairs
    int cur;          // index of the "Old" ρ(k) buffer
    unsigned long long seq;
    int recip_ns;     // sites in the ρ(k) delta (0: take molecule i's)
    int recip_from_args; // 1: old sites and charges come from site_old/q (mmc_recip_move)
    double com_new[3];
    double site_new[MMC_MAX_SITES * 3];
    double site_old[MMC_MAX_SITES * 3];
    double q[MMC_MAX_SITES];
};

// what the last CTA writes to the host
struct MoveOut {
    double lj_pot[2], lj_vir[2];   // already ×4 and ×24/3
    double qq[2];                  // EwaldReal, un-scaled, 0 if overlap
    double d_recip;                // ×factor
    int overlap[2];
    unsigned long long seq;        // written last
};

struct MoveScratch {
    double4 *partial;        // [blocks] {lj_pot, lj_vir, coul, overlap}
    unsigned int *ticket;
    MoveOut *out;            // mapped pinned host memory (device alias)
};

__device__ __forceinline__ void move_pair_block(const DevSystem &S, const MoveArgs &A, int cfg,
                                                int tile0, double (&acc)[4])
{
    __shared__ double4 s_isite[MMC_MAX_SITES];
    __shared__ int s_itype[MMC_MAX_SITES];
    __shared__ double4 s_rij[MOVE_BLOCK];   // {rij.x, rij.y, rij.z, flags}
    __shared__ int s_j[MOVE_BLOCK];
    __shared__ int s_wcount[MOVE_BLOCK / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i = A.i;
    const int2 mi = S.mol[i];
    const int nsi = mi.y;
    const double L = S.box;
    double cx, cy, cz;
    if (cfg == 0) {
        const double4 c = S.com[i];
        cx = c.x; cy = c.y; cz = c.z;
    } else {
        cx = A.com_new[0]; cy = A.com_new[1]; cz = A.com_new[2];
    }
    if (tid < nsi) {
        double4 s = S.site[mi.x + tid];
        if (cfg == 1) {
            s.x = A.site_new[3 * tid]; s.y = A.site_new[3 * tid + 1]; s.z = A.site_new[3 * tid + 2];
        }
        s_isite[tid] = s;
        s_itype[tid] = S.atype[mi.x + tid];
    }
    const double rc_lj2 = S.rc_lj * S.rc_lj, rc_qq2 = S.rc_qq * S.rc_qq;
    const int nt = S.n_types;
    const int SM = S.max_sites;
    const int per_mol = nsi * SM;

    for (int tile = tile0; tile * MOVE_BLOCK < S.n_mol; tile += A.tiles) {
        __syncthreads();   // s_isite ready / previous tile's queue fully consumed
        // ---- phase 1: COM gate, ordered compaction
        const int j = tile * MOVE_BLOCK + tid;
        bool in = false;
        double rx = 0, ry = 0, rz = 0;
        int flags = 0;
        if (j < S.n_mol && j != i) {
            const double4 cj = S.com[j];
            rx = min_image(cx, cj.x, L);
            ry = min_image(cy, cj.y, L);
            rz = min_image(cz, cj.z, L);
            const double r2 = rx * rx + ry * ry + rz * rz;
            if (A.want_lj && r2 < rc_lj2) flags |= 1;
            if (A.want_qq && r2 < rc_qq2) flags |= 2;
            in = flags != 0;
        }
        const unsigned m = __ballot_sync(0xffffffffu, in);
        if (lane == 0) s_wcount[warp] = __popc(m);
        __syncthreads();
        int off = 0, n_in = 0;
#pragma unroll
        for (int w = 0; w < MOVE_BLOCK / 32; ++w) {
            if (w < warp) off += s_wcount[w];
            n_in += s_wcount[w];
        }
        if (in) {
            const int p = off + __popc(m & ((1u << lane) - 1u));
            s_rij[p] = make_double4(rx, ry, rz, (double)flags);
            s_j[p] = j;
        }
        __syncthreads();
        // ---- phase 2: site pairs spread over the CTA
        const int items = n_in * per_mol;
        for (int w = tid; w < items; w += MOVE_BLOCK) {
            const int jj = w / per_mol;
            const int rem = w - jj * per_mol;
            const int a = rem / SM, b = rem - a * SM;
            const int2 mj = S.mol[s_j[jj]];
            if (b >= mj.y) continue;
            const double4 sb = S.site[mj.x + b];
            const double4 sa = s_isite[a];
            const double4 rij = s_rij[jj];
            const int fl = (int)rij.w;
            const double dx = min_image(sa.x, sb.x, L);
            const double dy = min_image(sa.y, sb.y, L);
            const double dz = min_image(sa.z, sb.z, L);
            const double r2 = dx * dx + dy * dy + dz * dz;
            if (fl & 1) {
                const int ta = s_itype[a], tb = S.atype[mj.x + b];
                const double eps = S.eps[ta + tb * nt];
                if (r2 < (rc_lj2 + 100) && eps > 0.001)
                    lj_pair(eps, S.sig[ta + tb * nt], r2, dx, dy, dz, rij.x, rij.y, rij.z, acc[0], acc[1]);
            }
            if (fl & 2) {
                const double qq = sa.w * sb.w;
                if ((r2 < 0.5) && (qq < 0)) {
                    acc[3] = 1.0;                                    // ewalds.jl:359-360
                } else if (r2 < (rc_qq2 + 100)) {
                    const double r = sqrt(r2);
                    acc[2] += qq * erfc(S.kappa * r) / r;            // ewalds.jl:366-367
                }
            }
        }
    }
}

// RecipMove (Ewald/ewalds.jl:718-826) for the ns sites of the moved molecule.
// k-vectors are spread over the recip CTAs' threads; each CTA rebuilds the tiny e^{ik·r}
// tables (2 × ns × 3 sincos + recurrences) in shared memory.
__device__ __forceinline__ void move_recip_block(const DevSystem &S, const MoveArgs &A, int rb,
                                                 double (&acc)[4])
{
    __shared__ cplx s_e[2][MMC_MAX_SITES][3][MMC_MAX_NK + 1];
    __shared__ double s_q[MMC_MAX_SITES];
    const int tid = threadIdx.x;
    int2 mi = make_int2(0, A.recip_ns);
    if (!A.recip_from_args) mi = S.mol[A.i];
    const int ns = mi.y, nk = S.nk;
    const double L = S.box;
    const double twopi = 2.0 * 3.141592653589793;
    for (int t = tid; t < 2 * ns * 3; t += MOVE_BLOCK) {
        const int cfg = t / (ns * 3), rem = t - cfg * ns * 3, l = rem / 3, d = rem - 3 * l;
        double x;
        if (cfg == 0 && A.recip_from_args) {
            x = A.site_old[3 * l + d];
            if (d == 0) s_q[l] = A.q[l];
        } else if (cfg == 0) {
            const double4 s = S.site[mi.x + l];
            x = d == 0 ? s.x : (d == 1 ? s.y : s.z);
            if (d == 0) s_q[l] = s.w;
        } else {
            x = A.site_new[3 * l + d];
        }
        cplx e1;
        sincos(twopi * x / L, &e1.im, &e1.re);      // ewalds.jl:770-781: cos(twopi*x/L) + sin(...)im
        cplx e; e.re = 1.0; e.im = 0.0;
        s_e[cfg][l][d][0] = e;
        e = e1;
        s_e[cfg][l][d][1] = e;
        for (int k = 2; k <= nk; ++k) { e = cmul(e, e1); s_e[cfg][l][d][k] = e; }  // :790-795
    }
    __syncthreads();
    const double2 *Sold = S.rhok[A.cur];
    double2 *Snew = S.rhok[A.cur ^ 1];
    for (int k = rb * MOVE_BLOCK + tid; k < S.nkvecs; k += A.recip_blocks * MOVE_BLOCK) {
        const int4 kv = S.kvec[k];
        const int aky = abs(kv.y), akz = abs(kv.z);
        const bool ny = kv.y < 0, nz = kv.z < 0;
        const double2 so = Sold[k];
        double nr = so.x, ni = so.y;
        for (int l = 0; l < ns; ++l) {
            const cplx tn = cmul(cmul(s_e[1][l][0][kv.x], cconj_if(s_e[1][l][1][aky], ny)),
                                 cconj_if(s_e[1][l][2][akz], nz));
            const cplx to = cmul(cmul(s_e[0][l][0][kv.x], cconj_if(s_e[0][l][1][aky], ny)),
                                 cconj_if(s_e[0][l][2][akz], nz));
            nr += s_q[l] * (tn.re - to.re);                          // ewalds.jl:804-813
            ni += s_q[l] * (tn.im - to.im);
        }
        Snew[k] = make_double2(nr, ni);
        acc[0] += S.cfac[k] * ((nr * nr + ni * ni) - (so.x * so.x + so.y * so.y));  // :817-821
    }
}

__global__ void __launch_bounds__(MOVE_BLOCK)
k_move(const __grid_constant__ DevSystem S, const __grid_constant__ MoveArgs A, MoveScratch W)
{
    __shared__ double s_red[4 * (MOVE_BLOCK / 32)];
    __shared__ bool s_last;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const int b = blockIdx.x;
    const int n_pair = A.n_cfg * A.tiles;
    if (b < n_pair) move_pair_block(S, A, b / A.tiles, b % A.tiles, acc);
    else move_recip_block(S, A, b - n_pair, acc);
    block_sum<4, MOVE_BLOCK>(acc, s_red);
    if (threadIdx.x == 0) {
        W.partial[b] = make_double4(acc[0], acc[1], acc[2], acc[3]);
        __threadfence();
        const unsigned t = atomicAdd(W.ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    // ---- last CTA: fold partials in block order (deterministic), publish to the host
    __threadfence();
    if (threadIdx.x == 0) {
        MoveOut o;
        bool any_ovl = false;
        for (int cfg = 0; cfg < 2; ++cfg) {
            double lp = 0, lv = 0, cq = 0, ov = 0;
            if (cfg < A.n_cfg)
                for (int t = 0; t < A.tiles; ++t) {
                    const double4 p = ldcg4(&W.partial[cfg * A.tiles + t]);
                    lp += p.x; lv += p.y; cq += p.z; ov += p.w;
                }
            const bool ovl = (ov > 0.0) && !A.ignore_overlap;
            any_ovl |= ovl;
            o.lj_pot[cfg] = lp * 4;              // energy.jl:289  pot * 4
            o.lj_vir[cfg] = lv * 24 / 3.0;       //                vir * 24 / 3.0
            o.qq[cfg] = ovl ? 0.0 : cq;          // ewalds.jl:360  return 0.0, true
            o.overlap[cfg] = (ov > 0.0) ? 1 : 0;
        }
        double dr = 0.0;
        for (int t = 0; t < A.recip_blocks; ++t) dr += ldcg4(&W.partial[n_pair + t]).x;
        o.d_recip = any_ovl ? 0.0 : dr * S.factor;   // main.jl:580-590; ewalds.jl:825
        o.seq = 0;
        *W.out = o;
        *W.ticket = 0;
        __threadfence_system();
        *((volatile unsigned long long *)&W.out->seq) = A.seq;
    }
}

// commit of an accepted move: main.jl:527,552 made permanent (the ρ(k) part is a pointer swap)
__global__ void k_set_molecule(DevSystem S, int i, double cx, double cy, double cz, MoveArgs A)
{
    const int2 mi = S.mol[i];
    const int t = threadIdx.x;
    if (t == 0) S.com[i] = make_double4(cx, cy, cz, 0.0);
    if (t < mi.y) {
        double4 s = S.site[mi.x + t];
        s.x = A.site_new[3 * t]; s.y = A.site_new[3 * t + 1]; s.z = A.site_new[3 * t + 2];
        S.site[mi.x + t] = s;
    }
}

// ------------------------------------------------------------------------------ monatomic
#define ATOM_BLOCK 256

struct AtomArgs {
    int i, n_cfg, blocks;
    unsigned long long seq;
    double r_new[3];
};

// Monatomic/mainMonatomic.jl:227-272 LJ_ΔU for atom i at its resident (cfg 0) and trial (cfg 1)
// position in one pass over the partner atoms: r_j, ε_j, σ_j are loaded once for both.
__global__ void __launch_bounds__(ATOM_BLOCK)
k_move_atom(DevAtoms S, const __grid_constant__ AtomArgs A, MoveScratch W)
{
    __shared__ double s_red[4 * (ATOM_BLOCK / 32)];
    __shared__ bool s_last;
    const double4 r0 = S.r[A.i];
    const double L = S.box, rc2 = S.rc * S.rc;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int j = blockIdx.x * ATOM_BLOCK + threadIdx.x; j < S.n; j += gridDim.x * ATOM_BLOCK) {
        if (j == A.i) continue;
        const double4 rj = S.r[j];
        const double2 es = S.es[j];
        {
            const double dx = min_image(r0.x, rj.x, L), dy = min_image(r0.y, rj.y, L),
                         dz = min_image(r0.z, rj.z, L);
            const double r2 = dx * dx + dy * dy + dz * dz;
            if (!(r2 > rc2)) {                                   // mainMonatomic.jl:249
                const double sr2 = es.y * es.y / r2, sr6 = sr2 * sr2 * sr2, sr12 = sr6 * sr6;
                acc[0] += es.x * (sr12 - sr6);
                acc[1] += es.x * (2 * sr12 - sr6);
            }
        }
        if (A.n_cfg > 1) {
            const double dx = min_image(A.r_new[0], rj.x, L), dy = min_image(A.r_new[1], rj.y, L),
                         dz = min_image(A.r_new[2], rj.z, L);
            const double r2 = dx * dx + dy * dy + dz * dz;
            if (!(r2 > rc2)) {
                const double sr2 = es.y * es.y / r2, sr6 = sr2 * sr2 * sr2, sr12 = sr6 * sr6;
                acc[2] += es.x * (sr12 - sr6);
                acc[3] += es.x * (2 * sr12 - sr6);
            }
        }
    }
    block_sum<4, ATOM_BLOCK>(acc, s_red);
    if (threadIdx.x == 0) {
        W.partial[blockIdx.x] = make_double4(acc[0], acc[1], acc[2], acc[3]);
        __threadfence();
        s_last = (atomicAdd(W.ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last || threadIdx.x != 0) return;
    __threadfence();
    double p0 = 0, v0 = 0, p1 = 0, v1 = 0;
    for (int t = 0; t < (int)gridDim.x; ++t) {
        const double4 p = ldcg4(&W.partial[t]);
        p0 += p.x; v0 += p.y; p1 += p.z; v1 += p.w;
    }
    MoveOut o;
    o.lj_pot[0] = p0 * 4.0; o.lj_vir[0] = v0 * 24.0 / 3.0;     // mainMonatomic.jl:271
    o.lj_pot[1] = p1 * 4.0; o.lj_vir[1] = v1 * 24.0 / 3.0;
    o.qq[0] = o.qq[1] = 0.0; o.d_recip = 0.0; o.overlap[0] = o.overlap[1] = 0; o.seq = 0;
    *W.out = o;
    *W.ticket = 0;
    __threadfence_system();
    *((volatile unsigned long long *)&W.out->seq) = A.seq;
}

__global__ void k_set_atom(DevAtoms S, int i, double x, double y, double z)
{
    S.r[i] = make_double4(x, y, z, 0.0);
}
