// kernels_move.cuh — the per-trial-move kernels.
//
// One launch evaluates, for molecule i:
//   cfg 0: LJ_poly_ΔU + EwaldReal against the resident state        (Ewald/main.jl:491,501)
//   cfg 1: the same with molecule i at its trial position            (Ewald/main.jl:557,566)
//   recip: RecipMove's ρ(k) delta update and ΔE                      (Ewald/main.jl:581)
// These launches are latency bound by construction (≈2×10⁵ flop per move at N = 750 ≈ 6 ns of
// FP64 peak), so everything here is about the number of DEPENDENT memory round trips:
//   * the trial coordinates, and the previous accepted move still to be committed, travel as
//     kernel parameters (no memcpy, no separate commit launch: every reader substitutes the
//     pending molecule from the parameters and the last CTA writes it back to HBM);
//   * a pair CTA loads COM + sites + types of its 128 partner molecules in ONE round trip
//     (speculatively, before the COM gate), gates, compacts the survivors in index order into
//     shared memory and spreads the n_in × n_a × n_b site pairs over all threads;
//   * erfc(κr)/r uses the per-box erf polynomial (erf_poly.h) — a short dependency chain;
//   * every CTA publishes its partial sums straight into its own 64-byte slot of a mapped pinned
//     host array as four self-validating 16-byte {value, seq} packets; the host polls the slots
//     and folds them in CTA order.  No tickets, no last-block pass, no fence, no D2H copy, no
//     cudaStreamSynchronize.
#pragma once
#include "erf_poly.h"
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
    int commit_i;     // ≥ 0: accepted move not yet written to HBM (molecule index), -1: none
    int commit_ns;
    int i_first, i_ns; // molecule i's first site and site count
    double com_new[3];
    double site_new[MMC_MAX_SITES * 3];
    double site_old[MMC_MAX_SITES * 3];
    double q[MMC_MAX_SITES];
    double commit_com[3];
    double commit_site[MMC_MAX_SITES * 3];
};

// One 64-byte slot per CTA in mapped pinned host memory: four 16-byte packets {value, seq}.
// A packet is ONE 16-byte store, so it reaches host memory whole: the host accepts a packet when
// its seq matches and needs no fence between "data" and "flag" — which removes the
// system-scope membar (≈1-2 µs over PCIe) from the critical path of every move.
struct MovePacket { double v; unsigned long long seq; };
struct MoveSlot { MovePacket p[4]; };   // pair CTA: lj_pot, lj_vir, coul, overlap; recip CTA: ΔE (un-scaled)

struct MoveScratch {
    MoveSlot *slots;               // device alias of the mapped host array
};

// lanes 0..3 of the calling warp each store one packet (one coalesced 64-byte write)
__device__ __forceinline__ void publish(MoveSlot *slot, const double (&acc)[4], unsigned long long seq)
{
    const int l = threadIdx.x;
    if (l < 4) {
        const double v = l == 0 ? acc[0] : (l == 1 ? acc[1] : (l == 2 ? acc[2] : acc[3]));
        const ulonglong2 pk = make_ulonglong2((unsigned long long)__double_as_longlong(v), seq);
        asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(&slot->p[l]), "l"(pk.x), "l"(pk.y) : "memory");
    }
}

// q_a q_b erfc(κ r)/r with the reference's overlap rule and r² < r_cut²+100 test (ewalds.jl:359-367)
__device__ __forceinline__ void move_coul(const DevSystem &S, const ErfPoly &P, double r2, double qq, double cut2,
                                          double (&acc)[4])
{
    if ((r2 < 0.5) && (qq < 0)) {
        acc[3] = 1.0;
    } else if (r2 < cut2) {
        if (P.deg > 0) {
            const double rinv = rsqrt(r2);
            const double sv = fma(r2 * P.kappa2, P.scale, -1.0);
            double pv = P.c[P.deg];
#pragma unroll 1
            for (int k = P.deg - 1; k >= 0; --k) pv = fma(pv, sv, P.c[k]);
            acc[2] = fma(qq, fma(-P.kappa, pv, rinv), acc[2]);
        } else {
            const double r = sqrt(r2);
            acc[2] += qq * erfc(S.kappa * r) / r;
        }
    }
}

// SP > 0: every molecule has at most SP sites and the partner sites are prefetched with the COMs
template <int SP>
__device__ __forceinline__ void move_pair_block(const DevSystem &S, const MoveArgs &A, const ErfPoly &P, int cfg,
                                                int tile0, double (&acc)[4])
{
    constexpr int SQ = SP > 0 ? SP : 1;
    __shared__ double4 s_isite[MMC_MAX_SITES];
    __shared__ int s_itype[MMC_MAX_SITES];
    __shared__ double4 s_rij[MOVE_BLOCK];   // {rij.x, rij.y, rij.z, flags}
    __shared__ int s_j[MOVE_BLOCK];
    __shared__ int s_wcount[MOVE_BLOCK / 32];
    __shared__ double4 s_qsite[SP > 0 ? MOVE_BLOCK * SQ : 1];
    __shared__ int s_qtype[SP > 0 ? MOVE_BLOCK * SQ : 1];
    __shared__ int s_qns[SP > 0 ? MOVE_BLOCK : 1];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i = A.i, ci = A.commit_i;
    const int2 mi = make_int2(A.i_first, A.i_ns);      // host knows molecule i's layout: no dependent load
    const int nsi = mi.y;
    const double L = S.box;
    double cx, cy, cz;
    if (cfg == 1) { cx = A.com_new[0]; cy = A.com_new[1]; cz = A.com_new[2]; }
    else if (i == ci) { cx = A.commit_com[0]; cy = A.commit_com[1]; cz = A.commit_com[2]; }
    else { const double4 c = S.com[i]; cx = c.x; cy = c.y; cz = c.z; }
    if (tid < nsi) {
        double4 s = S.site[mi.x + tid];
        if (cfg == 1) { s.x = A.site_new[3 * tid]; s.y = A.site_new[3 * tid + 1]; s.z = A.site_new[3 * tid + 2]; }
        else if (i == ci) { s.x = A.commit_site[3 * tid]; s.y = A.commit_site[3 * tid + 1]; s.z = A.commit_site[3 * tid + 2]; }
        s_isite[tid] = s;
        s_itype[tid] = S.atype[mi.x + tid];
    }
    const double rc_lj2 = S.rc_lj * S.rc_lj, rc_qq2 = S.rc_qq * S.rc_qq;
    const int nt = S.n_types;
    const int SM = S.max_sites;
    const int per_mol = nsi * SM;

    for (int tile = tile0; tile * MOVE_BLOCK < S.n_mol; tile += A.tiles) {
        __syncthreads();   // s_isite ready / previous tile's queue fully consumed
        // ---- phase 1: one round trip for COM (+ sites + types), COM gate, ordered compaction
        const int j = tile * MOVE_BLOCK + tid;
        bool in = false;
        double rx = 0, ry = 0, rz = 0;
        int flags = 0;
        double4 pre[SQ];
        int pty[SQ];
        int2 mj = make_int2(0, 0);
        if (j < S.n_mol && j != i) {
            double4 cj = S.com[j];
            if (SP > 0) {
                mj = S.uni > 0 ? make_int2(j * S.uni, S.uni) : S.mol[j];   // uniform topology: no dependent load
#pragma unroll
                for (int k = 0; k < SQ; ++k) {
                    pre[k] = (k < mj.y) ? S.site[mj.x + k] : make_double4(0, 0, 0, 0);
                    pty[k] = (k < mj.y) ? S.atype[mj.x + k] : 0;
                }
            }
            if (j == ci) {                                    // accepted move still pending in the parameters
                cj.x = A.commit_com[0]; cj.y = A.commit_com[1]; cj.z = A.commit_com[2];
                if (SP > 0) {
#pragma unroll
                    for (int k = 0; k < SQ; ++k)
                        if (k < mj.y) { pre[k].x = A.commit_site[3 * k]; pre[k].y = A.commit_site[3 * k + 1]; pre[k].z = A.commit_site[3 * k + 2]; }
                }
            }
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
            if (SP > 0) {
                s_qns[p] = mj.y;
#pragma unroll
                for (int k = 0; k < SQ; ++k) { s_qsite[p * SQ + k] = pre[k]; s_qtype[p * SQ + k] = pty[k]; }
            }
        }
        __syncthreads();
        // ---- phase 2: site pairs spread over the CTA
        const int items = n_in * per_mol;
        for (int w = tid; w < items; w += MOVE_BLOCK) {
            const int jj = w / per_mol;
            const int rem = w - jj * per_mol;
            const int a = rem / SM, b = rem - a * SM;
            double4 sb;
            int tb;
            if (SP > 0) {
                if (b >= s_qns[jj]) continue;
                sb = s_qsite[jj * SQ + b];
                tb = s_qtype[jj * SQ + b];
            } else {
                const int jm = s_j[jj];
                const int2 mjj = S.mol[jm];
                if (b >= mjj.y) continue;
                sb = S.site[mjj.x + b];
                tb = S.atype[mjj.x + b];
                if (jm == ci) { sb.x = A.commit_site[3 * b]; sb.y = A.commit_site[3 * b + 1]; sb.z = A.commit_site[3 * b + 2]; }
            }
            const double4 sa = s_isite[a];
            const double4 rij = s_rij[jj];
            const int fl = (int)rij.w;
            const double dx = min_image(sa.x, sb.x, L);
            const double dy = min_image(sa.y, sb.y, L);
            const double dz = min_image(sa.z, sb.z, L);
            const double r2 = dx * dx + dy * dy + dz * dz;
            if (fl & 1) {
                const int ta = s_itype[a];
                const double eps = S.eps[ta + tb * nt];
                if (r2 < (rc_lj2 + 100) && eps > 0.001)
                    lj_pair(eps, S.sig[ta + tb * nt], r2, dx, dy, dz, rij.x, rij.y, rij.z, acc[0], acc[1]);
            }
            if (fl & 2) move_coul(S, P, r2, sa.w * sb.w, rc_qq2 + 100, acc);
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
    if (!A.recip_from_args) mi = make_int2(A.i_first, A.i_ns);
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
            if (A.i == A.commit_i) x = A.commit_site[3 * l + d];
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

template <int SP>
static __global__ void __launch_bounds__(MOVE_BLOCK)
k_move(const __grid_constant__ DevSystem S, const __grid_constant__ MoveArgs A,
       const __grid_constant__ ErfPoly P, MoveScratch W)
{
    __shared__ double s_red[4 * (MOVE_BLOCK / 32)];
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const int b = blockIdx.x;
    const int n_pair = A.n_cfg * A.tiles;
    if (b < n_pair) move_pair_block<SP>(S, A, P, b / A.tiles, b % A.tiles, acc);
    else move_recip_block(S, A, b - n_pair, acc);
    block_sum_all<4, MOVE_BLOCK>(acc, s_red);
    publish(&W.slots[b], acc, A.seq);
    // the last CTA makes the pending accepted move permanent (main.jl:527,552); nobody in this
    // launch reads these addresses: every reader took molecule commit_i from the parameters
    if (b == gridDim.x - 1 && A.commit_i >= 0) {
        const int2 mc = S.mol[A.commit_i];
        const int t = threadIdx.x;
        if (t == 0) S.com[A.commit_i] = make_double4(A.commit_com[0], A.commit_com[1], A.commit_com[2], 0.0);
        if (t < mc.y) {
            double4 s = S.site[mc.x + t];
            s.x = A.commit_site[3 * t]; s.y = A.commit_site[3 * t + 1]; s.z = A.commit_site[3 * t + 2];
            S.site[mc.x + t] = s;
        }
    }
}

// explicit write of one molecule (mmc_set_molecule, and the flush of a pending accepted move)
static __global__ void k_set_molecule(DevSystem S, int i, double cx, double cy, double cz, MoveArgs A)
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
    int commit_i;                  // ≥ 0: accepted move not yet written to HBM
    unsigned long long seq;
    double r_new[3];
    double commit_r[3];
};

// Monatomic/mainMonatomic.jl:227-272 LJ_ΔU for atom i at its resident (cfg 0) and trial (cfg 1)
// position in one pass over the partner atoms: r_j, ε_j, σ_j are loaded once for both.
static __global__ void __launch_bounds__(ATOM_BLOCK)
k_move_atom(DevAtoms S, const __grid_constant__ AtomArgs A, MoveScratch W)
{
    __shared__ double s_red[4 * (ATOM_BLOCK / 32)];
    double4 r0 = S.r[A.i];
    if (A.i == A.commit_i) { r0.x = A.commit_r[0]; r0.y = A.commit_r[1]; r0.z = A.commit_r[2]; }
    const double L = S.box, rc2 = S.rc * S.rc;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int j = blockIdx.x * ATOM_BLOCK + threadIdx.x; j < S.n; j += gridDim.x * ATOM_BLOCK) {
        if (j == A.i) continue;
        double4 rj = S.r[j];
        const double2 es = S.es[j];
        if (j == A.commit_i) { rj.x = A.commit_r[0]; rj.y = A.commit_r[1]; rj.z = A.commit_r[2]; }
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
    block_sum_all<4, ATOM_BLOCK>(acc, s_red);
    publish(&W.slots[blockIdx.x], acc, A.seq);
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0 && A.commit_i >= 0)
        S.r[A.commit_i] = make_double4(A.commit_r[0], A.commit_r[1], A.commit_r[2], 0.0);
}

static __global__ void k_set_atom(DevAtoms S, int i, double x, double y, double z)
{
    S.r[i] = make_double4(x, y, z, 0.0);
}
