// kernels_chain.cuh — a block of trial moves evaluated back to back on the device.
//
// The per-move entry points (kernels_move.cuh) pay one kernel launch and one PCIe round trip per
// trial move: ≈12 µs for ≈2×10⁵ flop.  A Markov chain is sequential, so the only way to take that
// latency out is to keep the WHOLE accept/reject loop of Ewald/main.jl:487-651 next to the data for
// a block of moves: the caller hands over the stretch of its uniform random stream that the block
// will consume (the reference's own draw order) and gets back the accept/reject
// record, the per-move deltas and the final state — one launch per block (e.g. per sweep) instead
// of one per move.  Same decisions, same order, same stream.
//
//   one CTA of 512 threads, one SM; the state of the (small) system lives in SHARED memory:
//     site[N·S] double4, com[N] double4, ρ(k) Old/New, k-vectors, cfac        (N = 750: 127 KB)
//   per move:
//     step 0  lane 0 of warp 0 draws the move exactly like the host driver (mmc_driver.inl; every
//             product and sum rounded separately, as the reference's Julia does), from a ring of
//             uniforms that warp 0 prefetches one move ahead;
//     step 1  all threads: COM gate of molecule i at its old AND trial position against all j,
//             ordered compaction (ballot/popc) of the partners inside either cut-off; the last warp
//             builds the 2·S·3 e^{ik·r} recurrence tables of RecipMove (ewalds.jl:770-795);
//     step 2  all threads: (partner, old/new, site a) items, S site pairs each in lock-step
//             (LJ energy.jl:270-282, erfc Coulomb ewalds.jl:359-367 via the erf polynomial), and the
//             ρ(k) delta update ewalds.jl:804-821 for k = tid;
//     step 3  ordered block reduction, Metropolis (auxillary.jl:106-114) and, on accept, the state
//             update + ρ(k) Old/New flip (main.jl:599-629) by lane 0.
//   5 block barriers per move; nothing leaves the SM until the block of moves is done.
#pragma once
#include "kernels_move.cuh"

#define CHAIN_THREADS 512
#define CHAIN_WARPS (CHAIN_THREADS / 32)
#define CHAIN_MAXIT 4            // N <= CHAIN_MAXIT * CHAIN_THREADS
#define CHAIN_RING 64

struct ChainOut {                // mirrors mmc_loop_stats + return code + final ρ(k) index
    long long n_moves, n_accepted, n_overlap, uniforms_used;
    long long trans_attempt, trans_accept, rot_attempt, rot_accept;
    double dr_max, dphi_max, total_energy, total_virial;
    int ret, cur;
    long long phase_cycles[6];   // lane 0 of warp 0: step 0, step 1 + B2, compaction + B3, step 2, warp sums + B4 wait, reduce + decision + B5
};

struct ChainArgs {
    long long n_moves, n_uniforms;
    int style_qq, style_recip;   // Coulomb on (EWALD/WOLF); ρ(k) on (EWALD)
    int adjust, cur;
    double temperature, inv_temperature, dr_max, dphi_max, p_trans, p_rot, e0, v0;
    const double *uniforms;      // [n_uniforms]
    double *quat;                // [N][4]
    const double *db;            // [n_sites][3] body-fixed site vectors
    unsigned char *accepted;     // [n_moves] or NULL
    double *delta;               // [n_moves] or NULL
    ChainOut *out;
};

// n_mol: molecules held by ONE CTA (all of them for k_chain, a slice's capacity (N + C - 1) / C + 1 for k_chains)
__host__ __device__ inline size_t chain_smem_bytes(int n_mol, int S, int nk)
{
    return sizeof(double4) * ((size_t)n_mol * S + n_mol) + (size_t)nk * (2 * sizeof(double2) + sizeof(int4) + sizeof(double)) +
           sizeof(int2) * (size_t)n_mol;
}

namespace chain {
// separately rounded arithmetic: what the host driver (no FMA contraction) and Julia compute
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }

// Ewald/adjust.jl:1-83
struct MoveStat { long long naccepp, naccept, attempp, attempt; double set_value, d_max; };
__device__ inline void adjust_step(MoveStat &m, double L)
{
    if (m.attempp == 0) { m.naccepp = m.naccept; m.attempp = m.attempt; return; }
    const double ratio = __ddiv_rn((double)(m.naccept - m.naccepp), (double)(m.attempt - m.attempp));
    const double old = m.d_max;
    m.d_max = __ddiv_rn(mul(m.d_max, ratio), m.set_value);
    const double r = __ddiv_rn(m.d_max, old);
    if (r > 1.5) m.d_max = mul(old, 1.5);
    if (r < 0.5) m.d_max = mul(old, 0.5);
    if (m.d_max > __ddiv_rn(L, 2.0)) m.d_max = __ddiv_rn(L, 2.0);
    m.naccepp = m.naccept; m.attempp = m.attempt;
}
}  // namespace chain

// Driver state of the block (what Loop() keeps in local variables): lives in shared memory and is touched
// by lane 0 of warp 0 only, so that it costs the 511 other threads no registers in the pair loops.
struct ChainDriver {
    long long n_acc, n_ovl, n_done;
    chain::MoveStat tr, ro;
    double dr_max, dphi_max, tot_e, tot_v;
    double ei[4], nq[4], ndb[4 * 3];
    int ret, is_trans, cur;
};

// MUFU.RSQ64H seed + one cubic Newton step (see kernels_pairs_v3.cuh fast_rsqrt)
__device__ __forceinline__ double chain_rsqrt(double x)
{
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    const double e = fma(x, -(y0 * y0), 1.0);
    return fma(fma(e, 0.375, 0.5), y0 * e, y0);
}

// DEG > 0: padded degree of the erf polynomial at compile time; 0: erfc(); -1: run-time degree
template <int S, int DEG>
static __global__ void __launch_bounds__(CHAIN_THREADS, 1)
k_chain(const __grid_constant__ DevSystem Sy, const __grid_constant__ ChainArgs A, const __grid_constant__ ErfPoly P)
{
    using namespace chain;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = Sy.n_mol, NK = A.style_recip ? Sy.nkvecs : 0;
    double4 *s_site = reinterpret_cast<double4 *>(smem_raw);
    double4 *s_com = s_site + (size_t)N * S;
    double2 *s_rhok[2];
    s_rhok[0] = reinterpret_cast<double2 *>(s_com + N);
    s_rhok[1] = s_rhok[0] + NK;
    int4 *s_kvec = reinterpret_cast<int4 *>(s_rhok[1] + NK);
    double *s_cfac = reinterpret_cast<double *>(s_kvec + NK);
    int2 *s_list = reinterpret_cast<int2 *>(s_cfac + NK);

    __shared__ cplx s_tab[2][S][3][MMC_MAX_NK + 1];
    __shared__ int s_type[S];
    __shared__ int s_wcount[CHAIN_MAXIT * CHAIN_WARPS];
    __shared__ double s_red[9 * CHAIN_WARPS];
    __shared__ double s_u[CHAIN_RING];
    __shared__ double s_tcom[3], s_tsite[S][3];
    __shared__ int s_stop, s_cur;
    __shared__ ChainDriver D;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const double L = Sy.box;
    const double rc_lj2 = Sy.rc_lj * Sy.rc_lj, rc_qq2 = Sy.rc_qq * Sy.rc_qq;
    const int nt = Sy.n_types, nk = Sy.nk;
    const int nit = (N + CHAIN_THREADS - 1) / CHAIN_THREADS;
    const double twopi = 2.0 * 3.141592653589793;

    // ---- the resident state comes on chip once
    for (int t = tid; t < N * S; t += CHAIN_THREADS) s_site[t] = Sy.site[t];
    for (int t = tid; t < N; t += CHAIN_THREADS) s_com[t] = Sy.com[t];
    for (int t = tid; t < NK; t += CHAIN_THREADS) {
        s_rhok[A.cur][t] = Sy.rhok[A.cur][t]; s_rhok[A.cur ^ 1][t] = Sy.rhok[A.cur][t];
        s_kvec[t] = Sy.kvec[t]; s_cfac[t] = Sy.cfac[t];
    }
    if (tid < S) s_type[tid] = Sy.atype[tid];
    if (tid == 0) {
        s_stop = 0; s_cur = A.cur;
        D.n_acc = 0; D.n_ovl = 0; D.n_done = 0;
        D.tr = MoveStat{0, 0, 0, 0, 0.5, A.dr_max}; D.ro = MoveStat{0, 0, 0, 0, 0.5, A.dphi_max};
        D.dr_max = A.dr_max; D.dphi_max = A.dphi_max; D.tot_e = A.e0; D.tot_v = A.v0;
        D.ret = 0; D.is_trans = 1; D.cur = A.cur;
        for (int k = 0; k < 4; ++k) { D.ei[k] = 0.0; D.nq[k] = A.quat[k]; }
        for (int k = 0; k < S * 3; ++k) D.ndb[k] = A.db[k];
    }
    if (warp == 0) {                                        // first fill of the uniform ring
        const long long lim = A.n_uniforms < CHAIN_RING ? A.n_uniforms : CHAIN_RING;
        if (lane < lim) s_u[lane] = A.uniforms[lane];
        if (lane + 32 < lim) s_u[lane + 32] = A.uniforms[lane + 32];
    }
    // prefetch registers of warp 0: uniforms (two per lane) and the next molecule's quaternion / body frame (one per lane)
    double pre0 = 0.0, pre1 = 0.0, pfq = 0.0;
    int pre_cnt = 0, pf_on = 0;
    long long pos = 0, ring_end = A.n_uniforms < CHAIN_RING ? A.n_uniforms : CHAIN_RING;   // stream position (lane 0), ring fill (warp 0)
    bool dry = false;
    __syncthreads();

    auto next_u = [&]() -> double {                         // UStream::next of the host driver (lane 0 of warp 0)
        if (pos >= A.n_uniforms) { dry = true; return 0.5; }
        const double v = (pos < ring_end) ? s_u[pos & (CHAIN_RING - 1)] : A.uniforms[pos];
        ++pos;
        return v;
    };

    long long pc[6] = {0, 0, 0, 0, 0, 0};
    int cur = A.cur;
    long long pos_m = 0;                                    // stream position at the start of the current move (lane 0 of warp 0)
    for (long long m = 0; m < A.n_moves; ++m) {
        const int i = (int)(m % N);                         // sweep order i = 1..N (main.jl:490)
        long long tc0 = clock64();
        pos_m = pos;
        // ================= step 0: the trial move (main.jl:514-552), lane 0 of warp 0
        if (warp == 0) {
            if (pre_cnt > 0) {                              // uniforms requested during the previous move have landed
                if (lane < pre_cnt) s_u[(ring_end + lane) & (CHAIN_RING - 1)] = pre0;
                if (lane + 32 < pre_cnt) s_u[(ring_end + 32 + lane) & (CHAIN_RING - 1)] = pre1;
                ring_end += pre_cnt;
                pre_cnt = 0;
            }
            if (pf_on) {                                    // ... and so have the next molecule's quaternion and body frame
                if (lane >= 1 && lane <= 4) D.nq[lane - 1] = pfq;
                if (lane >= 5 && lane < 5 + S * 3) D.ndb[lane - 5] = pfq;
                pf_on = 0;
            }
            __syncwarp();
            if (lane == 0) {
                const double dr_max = D.dr_max, dphi_max = D.dphi_max;
                const double nq0 = D.nq[0], nq1 = D.nq[1], nq2 = D.nq[2], nq3 = D.nq[3];
                const double4 c0 = s_com[i];
                double rnew[3] = {c0.x, c0.y, c0.z};
                double e0 = nq0, e1 = nq1, e2 = nq2, e3 = nq3;
                int ret = 0;
                const double chose = next_u();              // main.jl:516
                if (chose < A.p_trans) {                    // main.jl:519-529, auxillary.jl:94-103
                    D.is_trans = 1; D.tr.attempt += 1;
                    const double z0 = next_u(), z1 = next_u(), z2 = next_u();
                    rnew[0] = add(rnew[0], mul(sub(z0, 0.5), dr_max));
                    rnew[1] = add(rnew[1], mul(sub(z1, 0.5), dr_max));
                    rnew[2] = add(rnew[2], mul(sub(z2, 0.5), dr_max));
#pragma unroll
                    for (int k = 0; k < 3; ++k) {           // boundaries.jl:16-26
                        if (rnew[k] > L) rnew[k] = sub(rnew[k], L);
                        if (rnew[k] < 0) rnew[k] = add(rnew[k], L);
                    }
                } else if (chose <= A.p_rot) {              // main.jl:530-538, quaternions.jl:52-73,94-120,158-182
                    D.is_trans = 0; D.ro.attempt += 1;
                    if (fabs(sub(add(add(add(mul(nq0, nq0), mul(nq1, nq1)), mul(nq2, nq2)), mul(nq3, nq3)), 1.0)) > 1.e-6) ret = 2;
                    double ax0, ax1, ax2, nrm;
                    for (;;) {
                        ax0 = sub(mul(2.0, next_u()), 1.0); ax1 = sub(mul(2.0, next_u()), 1.0); ax2 = sub(mul(2.0, next_u()), 1.0);
                        nrm = add(add(mul(ax0, ax0), mul(ax1, ax1)), mul(ax2, ax2));
                        if (nrm < 1.0 || dry) break;
                    }
                    const double sn = __dsqrt_rn(nrm);
                    ax0 = __ddiv_rn(ax0, sn); ax1 = __ddiv_rn(ax1, sn); ax2 = __ddiv_rn(ax2, sn);
                    const double zeta = next_u();
                    const double angle = mul(sub(mul(2.0, zeta), 1.0), dphi_max);
                    double sh, ch;
                    sincos(mul(0.5, angle), &sh, &ch);
                    const double r0 = ch, r1 = mul(sh, ax0), r2q = mul(sh, ax1), r3 = mul(sh, ax2);
                    e0 = sub(sub(sub(mul(r0, nq0), mul(r1, nq1)), mul(r2q, nq2)), mul(r3, nq3));   // quatmul(rot, old)
                    e1 = add(sub(add(mul(r1, nq0), mul(r0, nq1)), mul(r3, nq2)), mul(r2q, nq3));
                    e2 = sub(add(add(mul(r2q, nq0), mul(r3, nq1)), mul(r0, nq2)), mul(r1, nq3));
                    e3 = add(add(sub(mul(r3, nq0), mul(r2q, nq1)), mul(r1, nq2)), mul(r0, nq3));
                } else ret = 3;                             // main.jl:539-541
                if (ret == 0 && fabs(sub(add(add(add(mul(e0, e0), mul(e1, e1)), mul(e2, e2)), mul(e3, e3)), 1.0)) > 1.e-6) ret = 2;
                D.ei[0] = e0; D.ei[1] = e1; D.ei[2] = e2; D.ei[3] = e3;
                // quaternions.jl:37-50 — rows as written in the reference, [2,3] = 2(q2 q4 + q1 q2)
                const double q1 = e0, q2 = e1, q3 = e2, q4 = e3;
                double a[3][3];
                a[0][0] = sub(sub(add(mul(q1, q1), mul(q2, q2)), mul(q3, q3)), mul(q4, q4));
                a[0][1] = mul(2, add(mul(q2, q3), mul(q1, q4))); a[0][2] = mul(2, sub(mul(q2, q4), mul(q1, q3)));
                a[1][0] = mul(2, sub(mul(q2, q3), mul(q1, q4)));
                a[1][1] = sub(add(sub(mul(q1, q1), mul(q2, q2)), mul(q3, q3)), mul(q4, q4));
                a[1][2] = mul(2, add(mul(q2, q4), mul(q1, q2)));
                a[2][0] = mul(2, add(mul(q2, q4), mul(q1, q3))); a[2][1] = mul(2, sub(mul(q3, q4), mul(q1, q2)));
                a[2][2] = add(sub(sub(mul(q1, q1), mul(q2, q2)), mul(q3, q3)), mul(q4, q4));
#pragma unroll
                for (int s = 0; s < S; ++s) {               // main.jl:545-548: COM + MATMUL(ai, db)
                    const double d0 = D.ndb[3 * s], d1 = D.ndb[3 * s + 1], d2 = D.ndb[3 * s + 2];
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        s_tsite[s][c] = add(rnew[c], add(add(mul(d0, a[0][c]), mul(d1, a[1][c])), mul(d2, a[2][c])));
                }
                s_tcom[0] = rnew[0]; s_tcom[1] = rnew[1]; s_tcom[2] = rnew[2];
                if (ret) { D.ret = ret; s_stop = ret; }
            }
            __syncwarp();
            {   // requests for the NEXT move, consumed at the top of its step 0: uniforms, quaternion, body frame
                const long long p0 = __shfl_sync(0xffffffffu, pos, 0), re = ring_end;
                long long lim = p0 + CHAIN_RING;            // slot x may be overwritten once x - RING < pos
                if (lim > A.n_uniforms) lim = A.n_uniforms;
                const long long want = lim - re;
                pre_cnt = want > 0 ? (int)want : 0;
                if (lane < pre_cnt) pre0 = A.uniforms[re + lane];
                if (lane + 32 < pre_cnt) pre1 = A.uniforms[re + 32 + lane];
                const int inx = (i + 1 == N) ? 0 : i + 1;   // move m cannot change molecule i+1
                if (lane >= 1 && lane <= 4) pfq = A.quat[4 * inx + lane - 1];
                if (lane >= 5 && lane < 5 + S * 3) pfq = A.db[(size_t)inx * S * 3 + lane - 5];
                pf_on = 1;
            }
        }
        { const long long t = clock64(); pc[0] += t - tc0; tc0 = t; }
        __syncthreads();                                    // B1
        if (s_stop) break;

        // ================= step 1: COM gate for the old and the trial position
        const double4 co = s_com[i];
        const double cnx = s_tcom[0], cny = s_tcom[1], cnz = s_tcom[2];
        int myfl[CHAIN_MAXIT]; unsigned mymask[CHAIN_MAXIT];
#pragma unroll
        for (int it = 0; it < CHAIN_MAXIT; ++it) {
            myfl[it] = 0; mymask[it] = 0;
            if (it < nit) {
                const int j = it * CHAIN_THREADS + tid;
                int fl = 0;
                if (j < N && j != i) {
                    const double4 cj = s_com[j];
                    {
                        const double rx = min_image(co.x, cj.x, L), ry = min_image(co.y, cj.y, L), rz = min_image(co.z, cj.z, L);
                        const double r2 = add(add(mul(rx, rx), mul(ry, ry)), mul(rz, rz));
                        if (r2 < rc_lj2) fl |= 1;
                        if (A.style_qq && r2 < rc_qq2) fl |= 2;
                    }
                    {
                        const double rx = min_image(cnx, cj.x, L), ry = min_image(cny, cj.y, L), rz = min_image(cnz, cj.z, L);
                        const double r2 = add(add(mul(rx, rx), mul(ry, ry)), mul(rz, rz));
                        if (r2 < rc_lj2) fl |= 4;
                        if (A.style_qq && r2 < rc_qq2) fl |= 8;
                    }
                }
                const unsigned mk = __ballot_sync(0xffffffffu, fl != 0);
                if (lane == 0) s_wcount[it * CHAIN_WARPS + warp] = __popc(mk);
                myfl[it] = fl; mymask[it] = mk;
            }
        }
        if (A.style_recip && warp == CHAIN_WARPS - 1 && lane < 2 * S * 3) {   // ewalds.jl:770-795
            const int cfg = lane / (S * 3), rem = lane - cfg * S * 3, l = rem / 3, d = rem - 3 * l;
            double x;
            if (cfg == 0) { const double4 s = s_site[i * S + l]; x = d == 0 ? s.x : (d == 1 ? s.y : s.z); }
            else x = s_tsite[l][d];
            cplx e1;
            sincos(twopi * x / L, &e1.im, &e1.re);
            cplx e; e.re = 1.0; e.im = 0.0;
            s_tab[cfg][l][d][0] = e;
            e = e1;
            s_tab[cfg][l][d][1] = e;
            for (int k = 2; k <= nk; ++k) { e = cmul(e, e1); s_tab[cfg][l][d][k] = e; }
        }
        __syncthreads();                                    // B2
        { const long long t = clock64(); pc[1] += t - tc0; tc0 = t; }
        int n_in;
        {   // exclusive prefix of the per-(iteration, warp) counts, in index order
            const int nidx = nit * CHAIN_WARPS;
            const int c0 = lane < nidx ? s_wcount[lane] : 0, c1 = lane + 32 < nidx ? s_wcount[lane + 32] : 0;
            int i0 = c0, i1 = c1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v0 = __shfl_up_sync(0xffffffffu, i0, o), v1 = __shfl_up_sync(0xffffffffu, i1, o);
                if (lane >= o) { i0 += v0; i1 += v1; }
            }
            const int t0 = __shfl_sync(0xffffffffu, i0, 31), t1 = __shfl_sync(0xffffffffu, i1, 31);
            n_in = t0 + t1;
            const int e0 = i0 - c0, e1 = t0 + i1 - c1;       // exclusive prefixes of entries lane and lane+32
#pragma unroll
            for (int it = 0; it < CHAIN_MAXIT; ++it) {
                if (it < nit) {
                    const int idx = it * CHAIN_WARPS + warp;
                    const int b0 = __shfl_sync(0xffffffffu, e0, idx & 31), b1 = __shfl_sync(0xffffffffu, e1, idx & 31);
                    const int base = idx < 32 ? b0 : b1;
                    if (myfl[it]) s_list[base + __popc(mymask[it] & lt)] = make_int2(it * CHAIN_THREADS + tid, myfl[it]);
                }
            }
        }
        __syncthreads();                                    // B3
        { const long long t = clock64(); pc[2] += t - tc0; tc0 = t; }

        // ================= step 2: site pairs of (partner, cfg, site a) items + ρ(k) delta
        double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        const int items = n_in * 2 * S;
        for (int w = tid; w < items; w += CHAIN_THREADS) {
            const int jj = w / (2 * S), rem = w - jj * 2 * S, cfg = rem / S, a = rem - cfg * S;
            const int2 e = s_list[jj];
            const int j = e.x, fl = (e.y >> (2 * cfg)) & 3;
            if (!fl) continue;
            double4 sa = s_site[i * S + a];
            double cix = co.x, ciy = co.y, ciz = co.z;
            if (cfg) { sa.x = s_tsite[a][0]; sa.y = s_tsite[a][1]; sa.z = s_tsite[a][2]; cix = cnx; ciy = cny; ciz = cnz; }
            double r2[S], dx[S], dy[S], dz[S], qq[S], ri[S];
#pragma unroll
            for (int b = 0; b < S; ++b) {
                const double4 sb = s_site[j * S + b];
                dx[b] = min_image(sa.x, sb.x, L); dy[b] = min_image(sa.y, sb.y, L); dz[b] = min_image(sa.z, sb.z, L);
                r2[b] = dx[b] * dx[b] + dy[b] * dy[b] + dz[b] * dz[b];
                qq[b] = sa.w * sb.w;
                ri[b] = chain_rsqrt(r2[b]);
            }
            double l0 = 0.0, l1 = 0.0, l2 = 0.0, l3 = 0.0;
            if (fl & 1) {                                   // energy.jl:270-282
                const int ta = s_type[a];
                const double4 cj = s_com[j];
                const double rijx = min_image(cix, cj.x, L), rijy = min_image(ciy, cj.y, L), rijz = min_image(ciz, cj.z, L);
#pragma unroll
                for (int b = 0; b < S; ++b) {
                    const int tb = s_type[b];
                    const double eps = Sy.eps[ta + tb * nt];
                    if (r2[b] < (rc_lj2 + 100) && eps > 0.001) {          // σ²/r² formed as σ²·(1/√r²)²: no division on the critical path
                        const double sig = Sy.sig[ta + tb * nt];
                        const double s2 = sig * sig * (ri[b] * ri[b]), s6 = s2 * s2 * s2, s12 = s6 * s6;
                        l0 += eps * (s12 - s6);
                        const double virab = eps * (2.0 * s12 - s6) * s2;
                        l1 += rijx * (dx[b] * virab) + rijy * (dy[b] * virab) + rijz * (dz[b] * virab);
                    }
                }
            }
            if (fl & 2) {                                   // ewalds.jl:359-367
                bool use[S];
#pragma unroll
                for (int b = 0; b < S; ++b) {
                    use[b] = false;
                    if ((r2[b] < 0.5) && (qq[b] < 0)) l3 = 1.0;
                    else if (r2[b] < rc_qq2 + 100) use[b] = true;
                }
                if (DEG != 0) {                             // erfc(κr)/r = 1/r − κ·E(κ²r²), S chains in lock-step
                    double sv[S], pv[S];
                    const int deg = DEG > 0 ? DEG : P.deg;
#pragma unroll
                    for (int b = 0; b < S; ++b) { sv[b] = fma(r2[b] * P.kappa2, P.scale, -1.0); pv[b] = P.c[deg]; }
                    if (DEG > 0) {
#pragma unroll
                        for (int k = DEG - 1; k >= 0; --k)
#pragma unroll
                            for (int b = 0; b < S; ++b) pv[b] = fma(pv[b], sv[b], P.c[k]);
                    } else {
#pragma unroll 1
                        for (int k = deg - 1; k >= 0; --k) {
                            const double ck = P.c[k];
#pragma unroll
                            for (int b = 0; b < S; ++b) pv[b] = fma(pv[b], sv[b], ck);
                        }
                    }
#pragma unroll
                    for (int b = 0; b < S; ++b) if (use[b]) l2 = fma(qq[b], fma(-P.kappa, pv[b], ri[b]), l2);
                } else {
#pragma unroll
                    for (int b = 0; b < S; ++b) if (use[b]) { const double r = sqrt(r2[b]); l2 += qq[b] * erfc(Sy.kappa * r) / r; }
                }
            }
            if (cfg) { acc[4] += l0; acc[5] += l1; acc[6] += l2; acc[7] += l3; }
            else { acc[0] += l0; acc[1] += l1; acc[2] += l2; acc[3] += l3; }
        }
        if (A.style_recip) {                                // ewalds.jl:804-821
            const double2 *Sold = s_rhok[cur];
            double2 *Snew = s_rhok[cur ^ 1];
            for (int k = CHAIN_THREADS - 1 - tid; k < NK; k += CHAIN_THREADS) {   // from the far end: the low warps carry more pair items
                const int4 kv = s_kvec[k];
                const int aky = abs(kv.y), akz = abs(kv.z);
                const bool ny = kv.y < 0, nz = kv.z < 0;
                const double2 so = Sold[k];
                double nr = so.x, ni = so.y;
#pragma unroll
                for (int l = 0; l < S; ++l) {
                    const cplx tn = cmul(cmul(s_tab[1][l][0][kv.x], cconj_if(s_tab[1][l][1][aky], ny)), cconj_if(s_tab[1][l][2][akz], nz));
                    const cplx to = cmul(cmul(s_tab[0][l][0][kv.x], cconj_if(s_tab[0][l][1][aky], ny)), cconj_if(s_tab[0][l][2][akz], nz));
                    const double q = s_site[i * S + l].w;
                    nr += q * (tn.re - to.re);
                    ni += q * (tn.im - to.im);
                }
                Snew[k] = make_double2(nr, ni);
                acc[8] += s_cfac[k] * ((nr * nr + ni * ni) - (so.x * so.x + so.y * so.y));
            }
        }
        { const long long t = clock64(); pc[3] += t - tc0; tc0 = t; }
        // ================= step 3: ordered reduction, decision, state update
#pragma unroll
        for (int v = 0; v < 9; ++v) {
            acc[v] = warp_sum(acc[v]);
            if (lane == 0) s_red[v * CHAIN_WARPS + warp] = acc[v];
        }
        __syncthreads();                                    // B4
        { const long long t = clock64(); pc[4] += t - tc0; tc0 = t; }
        if (warp == 0) {
            double tot[9];
#pragma unroll
            for (int v = 0; v < 9; ++v) {
                double x = lane < CHAIN_WARPS ? s_red[v * CHAIN_WARPS + lane] : 0.0;
#pragma unroll
                for (int o = CHAIN_WARPS / 2; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
                tot[v] = x;
            }
            // launch_move_on's folding (energy.jl:289 pot*4, vir*24/3; ewalds.jl:360; main.jl:580-590) + mmc_trial_move.
            // Every lane holds the totals; the six divisions run in six lanes at once, lane 0 keeps the critical path
            // delta → delta/T → exp.
#pragma unroll
            for (int v = 0; v < 9; ++v) tot[v] = __shfl_sync(0xffffffffu, tot[v], 0);
            const bool ovl0 = tot[3] > 0.0, ovl1 = tot[7] > 0.0, overlap = ovl0 || ovl1;
            const double lj_old = mul(tot[0], 4), lj_new = mul(tot[4], 4);
            const double qq_old = mul(ovl0 ? 0.0 : tot[2], Sy.factor), qq_new = mul(ovl1 ? 0.0 : tot[6], Sy.factor);
            const double d_recip = (overlap || !A.style_recip) ? 0.0 : mul(tot[8], Sy.factor);
            double old_e = lj_old, new_e = lj_new;
            if (A.style_qq) { old_e = add(old_e, qq_old); new_e = add(new_e, qq_new); }     // main.jl:501-505, 566-570
            const double delta = add(sub(new_e, old_e), d_recip);                             // main.jl:593
            double num = delta, den = A.temperature;                                          // lane 0: delta / T
            if (lane == 1) { num = mul(tot[1], 24); den = 3.0; }                              // lj_vir_old
            if (lane == 2) { num = mul(tot[5], 24); den = 3.0; }                              // lj_vir_new
            if (lane == 3) { num = qq_old; den = 3.0; }                                       // ewalds.jl:907 virial = E/3
            if (lane == 4) { num = qq_new; den = 3.0; }
            if (lane == 5) { num = d_recip; den = 3.0; }
            const double quo = __ddiv_rn(num, den);
            const double ljv_old = __shfl_sync(0xffffffffu, quo, 1), ljv_new = __shfl_sync(0xffffffffu, quo, 2);
            const double qqv_old = __shfl_sync(0xffffffffu, quo, 3), qqv_new = __shfl_sync(0xffffffffu, quo, 4);
            const double recv = __shfl_sync(0xffffffffu, quo, 5);
            if (lane == 0) {
                double old_v = ljv_old, new_v = ljv_new;
                if (A.style_qq) { old_v = add(old_v, qqv_old); new_v = add(new_v, qqv_new); }
                const double x = quo;
                if (overlap) D.n_ovl += 1;
                bool okm = true;
                if (!(x < 0.0)) okm = exp(-x) > next_u();                            // auxillary.jl:106-114
                // a move whose draws ran past the end of the caller's stream never happened: nothing is committed or recorded,
                // its counters are taken back and the stream position returns to the start of the move (resume point)
                const bool accd = okm && !overlap && !dry;                            // main.jl:598
                if (accd) {
                    D.tot_e = add(D.tot_e, delta);
                    D.tot_v = add(D.tot_v, add(sub(new_v, old_v), recv));
                    D.n_acc += 1;
                    if (D.is_trans) D.tr.naccept += 1; else D.ro.naccept += 1;
                    s_com[i] = make_double4(s_tcom[0], s_tcom[1], s_tcom[2], 0.0);
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        double4 t = s_site[i * S + s];
                        t.x = s_tsite[s][0]; t.y = s_tsite[s][1]; t.z = s_tsite[s][2];
                        s_site[i * S + s] = t;
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) A.quat[4 * i + k] = D.ei[k];
                    if (A.style_recip && !overlap) s_cur = cur ^ 1;                   // main.jl:621 as an index flip
                }
                if (dry) {
                    D.ret = 1; s_stop = 1; pos = pos_m;
                    if (D.is_trans) D.tr.attempt -= 1; else D.ro.attempt -= 1;
                    if (overlap) D.n_ovl -= 1;
                } else {
                    if (A.accepted) A.accepted[m] = accd ? 1 : 0;
                    if (A.delta) A.delta[m] = delta;
                    if (A.adjust && i == N - 1) {                                     // main.jl:645-651
                        D.tr.d_max = D.dr_max; adjust_step(D.tr, L); D.dr_max = D.tr.d_max;
                        D.ro.d_max = D.dphi_max; adjust_step(D.ro, L); D.dphi_max = D.ro.d_max;
                    }
                    D.n_done = m + 1;
                }
            }
        }
        __syncthreads();                                    // B5
        { const long long t = clock64(); pc[5] += t - tc0; tc0 = t; }
        cur = s_cur;
        if (s_stop) break;
    }
    __syncthreads();
    // ---- the state goes back to HBM; ρ(k) of the final state into the buffer the handle will call "Old"
    for (int t = tid; t < N * S; t += CHAIN_THREADS) Sy.site[t] = s_site[t];
    for (int t = tid; t < N; t += CHAIN_THREADS) Sy.com[t] = s_com[t];
    for (int t = tid; t < NK; t += CHAIN_THREADS) Sy.rhok[cur][t] = s_rhok[cur][t];
    if (tid == 0) {
        ChainOut o;
        o.n_moves = D.n_done; o.n_accepted = D.n_acc; o.n_overlap = D.n_ovl; o.uniforms_used = pos;
        o.trans_attempt = D.tr.attempt; o.trans_accept = D.tr.naccept; o.rot_attempt = D.ro.attempt; o.rot_accept = D.ro.naccept;
        o.dr_max = D.dr_max; o.dphi_max = D.dphi_max; o.total_energy = D.tot_e; o.total_virial = D.tot_v;
        o.ret = D.ret; o.cur = cur;
        for (int k = 0; k < 6; ++k) o.phase_cycles[k] = pc[k];
        *A.out = o;
    }
}

#include <cooperative_groups.h>
#include <cstdio>
#define CHAINC_THREADS 256
#define CHAINC_WARPS (CHAINC_THREADS / 32)
#define CHAINC_MAXC 8

// ---------------------------------------------------------------------------------------------------------
// k_chains — the same block of moves on a thread-block CLUSTER (one CTA per SM, distributed shared memory).
// Every CTA keeps a full replica of the state and runs the (cheap, deterministic) driver redundantly: same
// uniforms, same trial move, same decision, same state update — so nothing but nine partial sums per move ever
// crosses between SMs.  What is split is the work that made the single-CTA kernel slow: CTA r gates and evaluates
// only the partner molecules j in its slice [r·N/C, (r+1)·N/C) and owns the slice [r·NK/C, (r+1)·NK/C) of ρ(k).
// Per move: the CTA's ordered partial sums go into slot r of every CTA's exchange buffer (st.shared::cluster),
// one cluster barrier (barrier.cluster arrive.release / wait.acquire), then every CTA adds the C slots by the same
// fixed tree — bit-identical totals everywhere — and decides.  The exchange buffer is double-buffered by move
// parity, so one cluster barrier per move is enough.  Rank 0 alone writes the accept/reject record.
//
// The trial move is off the critical path.  Per-phase counters of the first cluster version: drawing the move
// (lane 0: 930 cycles for a translation, 2400 for a quaternion rotation, plus the e^{ik·r} tables) was the
// largest serial piece of a move.  The trial move of move
// m+1 depends on nothing move m decides except (a) whether Metropolis(m) consumed a uniform — the stream position
// is P or P+1 — and (b) the step sizes, which only change at the end of a sweep.  So two generator warps (6, 7)
// build BOTH candidates for move m+1 (stream position P and P+1, trial coordinates and tables) while the six
// worker warps evaluate move m; the decision of move m selects one.  At a sweep end with step-size adaptation,
// and for move 0, the move is drawn after the decision instead (one candidate, exact position).  Everything a
// candidate needs comes from data move m cannot change (molecule i+1's COM, sites, quaternion, body frame).
// Cluster barrier: generators arrive as soon as the move starts (they publish nothing) and wait at its end.
#define CHAINS_WORKERS 192
#define CHAINS_WWARPS (CHAINS_WORKERS / 32)
#define CHAINS_MAXIT 8      // partner molecules per CTA <= CHAINS_MAXIT * CHAINS_WORKERS

template <int S>
struct ChainCand {
    double com[3], site[S][3], ei[4];            // trial COM, trial sites, trial quaternion
    double ocom[3], osite[S][3], q[S];           // molecule's resident COM, sites and charges (from L2: its owner is another SM)
    long long pos_begin, pos_end;                // stream position before / after drawing this candidate
    int is_trans, ret, dry, pad;
};

__device__ __forceinline__ void chain_worker_bar() { asm volatile("bar.sync 1, %0;" ::"n"(CHAINS_WORKERS) : "memory"); }

// SLICED = false: every CTA holds the whole system (fastest: the molecule about to move is on chip); true: a slice per CTA
template <int S, int DEG, bool SLICED>
static __global__ void __launch_bounds__(CHAINC_THREADS, 1)
k_chains(const __grid_constant__ DevSystem Sy, const __grid_constant__ ChainArgs A, const __grid_constant__ ErfPoly P)
{
    using namespace chain;
    namespace cg = cooperative_groups;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = Sy.n_mol, NK = A.style_recip ? Sy.nkvecs : 0;
    namespace cgx = cooperative_groups;
    const int C0 = (int)cgx::this_cluster().num_blocks();
    const int n_cap = SLICED ? (N + C0 - 1) / C0 + 1 : N;               // molecules held by this CTA (chain_smem_bytes: same formula)
    double4 *s_site = reinterpret_cast<double4 *>(smem_raw);            // SLICED: this CTA's slice of the molecules only
    double4 *s_com = s_site + (size_t)n_cap * S;
    double2 *s_rhok[2];
    s_rhok[0] = reinterpret_cast<double2 *>(s_com + n_cap);
    s_rhok[1] = s_rhok[0] + NK;
    int4 *s_kvec = reinterpret_cast<int4 *>(s_rhok[1] + NK);
    double *s_cfac = reinterpret_cast<double *>(s_kvec + NK);
    int2 *s_list = reinterpret_cast<int2 *>(s_cfac + NK);

    __shared__ ChainCand<S> s_cand[2][2];                              // [move parity][candidate]
    __shared__ cplx s_tabc[2][2][2][S][3][MMC_MAX_NK + 1];             // [move parity][candidate][old/new][site][xyz][power]
    __shared__ double s_ug[2][32], s_gq[2][4], s_gdb[2][S * 3];        // per generator: uniform window, quaternion, body frame
    __shared__ int s_type[S];
    __shared__ int s_wcount[CHAINS_MAXIT * CHAINS_WWARPS];
    __shared__ double s_red[9 * 8];
    __shared__ double s_xchg[2][9][CHAINC_MAXC];                       // [move parity][value][source rank], written remotely
    __shared__ double s_tot[9];
    __shared__ int s_stop, s_cur, s_sel;
    __shared__ ChainDriver D;
    __shared__ long long s_pos;                                        // stream position after the last decision

    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    double *myquat = A.quat + (size_t)rank * 4 * Sy.n_mol;             // this CTA's replica of the quaternion array

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool worker = warp < CHAINS_WWARPS;
    const int gen = warp - CHAINS_WWARPS;                              // 0, 1 for the generator warps
    const unsigned lt = (1u << lane) - 1u;
    const double L = Sy.box;
    const double rc_lj2 = Sy.rc_lj * Sy.rc_lj, rc_qq2 = Sy.rc_qq * Sy.rc_qq;
    const int nt = Sy.n_types, nk = Sy.nk;
    const int j_lo = (int)((long long)N * rank / C), j_hi = (int)((long long)N * (rank + 1) / C);
    const int k_lo = (int)((long long)NK * rank / C), k_hi = (int)((long long)NK * (rank + 1) / C);
    const int n_loc = j_hi - j_lo;
    const int s_lo = SLICED ? j_lo : 0;                                 // first molecule held in shared memory
    const double twopi = 2.0 * 3.141592653589793;

    {
        const int n_hold = SLICED ? n_loc : N;
        for (int t = tid; t < n_hold * S; t += CHAINC_THREADS) s_site[t] = Sy.site[(size_t)s_lo * S + t];
        for (int t = tid; t < n_hold; t += CHAINC_THREADS) s_com[t] = Sy.com[s_lo + t];
    }
    for (int t = tid; t < NK; t += CHAINC_THREADS) {
        s_rhok[A.cur][t] = Sy.rhok[A.cur][t]; s_rhok[A.cur ^ 1][t] = Sy.rhok[A.cur][t];
        s_kvec[t] = Sy.kvec[t]; s_cfac[t] = Sy.cfac[t];
    }
    if (tid < S) s_type[tid] = Sy.atype[tid];
    for (int k = tid; k < 2 * 9 * CHAINC_MAXC; k += CHAINC_THREADS) (&s_xchg[0][0][0])[k] = 0.0;
    if (tid < 9 * 8) s_red[tid] = 0.0;
    if (tid == 0) {
        s_stop = 0; s_cur = A.cur; s_sel = 0; s_pos = 0;
        D.n_acc = 0; D.n_ovl = 0; D.n_done = 0;
        D.tr = MoveStat{0, 0, 0, 0, 0.5, A.dr_max}; D.ro = MoveStat{0, 0, 0, 0, 0.5, A.dphi_max};
        D.dr_max = A.dr_max; D.dphi_max = A.dphi_max; D.tot_e = A.e0; D.tot_v = A.v0;
        D.ret = 0; D.is_trans = 1; D.cur = A.cur;
    }
    __syncthreads();

    // ---- one candidate for move `mv`, drawn from stream position `p0` (one warp; SURVEY A.5, mmc_driver.inl)
    auto generate = [&](int g, long long mv, long long p0, int par) {
        const int i = (int)(mv % N);
        {   // everything this candidate reads from HBM in one round trip: 32 uniforms, the quaternion, the body frame
            const long long x = p0 + lane;
            const double uv = x < A.n_uniforms ? A.uniforms[x] : 0.5;
            double qv = 0.0;
            if (lane < 4) qv = myquat[4 * i + lane];
            else if (lane < 4 + S * 3) qv = A.db[(size_t)i * S * 3 + lane - 4];
            // the molecule's resident COM and sites come from L2 (its owner is another SM): lanes 16..16+S
            double4 rv = make_double4(0, 0, 0, 0);
            if (lane == 16) rv = SLICED ? ldcg4(&Sy.com[i]) : s_com[i];
            else if (lane > 16 && lane <= 16 + S) rv = SLICED ? ldcg4(&Sy.site[(size_t)i * S + lane - 17]) : s_site[i * S + lane - 17];
            s_ug[g][lane] = uv;
            if (lane < 4) s_gq[g][lane] = qv;
            else if (lane < 4 + S * 3) s_gdb[g][lane - 4] = qv;
            ChainCand<S> &T0 = s_cand[par][g];
            if (lane == 16) { T0.ocom[0] = rv.x; T0.ocom[1] = rv.y; T0.ocom[2] = rv.z; }
            else if (lane > 16 && lane <= 16 + S) { T0.osite[lane - 17][0] = rv.x; T0.osite[lane - 17][1] = rv.y; T0.osite[lane - 17][2] = rv.z; T0.q[lane - 17] = rv.w; }
        }
        __syncwarp();
        ChainCand<S> &T = s_cand[par][g];
        if (lane == 0) {
            long long pos = p0;
            bool dry = false;
            auto next_u = [&]() -> double {                 // UStream::next of the host driver
                if (pos >= A.n_uniforms) { dry = true; return 0.5; }
                const long long k = pos - p0;
                const double v = k < 32 ? s_ug[g][k] : A.uniforms[pos];
                ++pos;
                return v;
            };
            const double dr_max = D.dr_max, dphi_max = D.dphi_max;
            const double nq0 = s_gq[g][0], nq1 = s_gq[g][1], nq2 = s_gq[g][2], nq3 = s_gq[g][3];
            const double4 c0 = make_double4(T.ocom[0], T.ocom[1], T.ocom[2], 0.0);
            double rnew[3] = {c0.x, c0.y, c0.z};
            double e0 = nq0, e1 = nq1, e2 = nq2, e3 = nq3;
            int ret = 0, is_trans = 1;
            const double chose = next_u();                  // main.jl:516
            if (chose < A.p_trans) {                        // main.jl:519-529, auxillary.jl:94-103
                const double z0 = next_u(), z1 = next_u(), z2 = next_u();
                rnew[0] = add(rnew[0], mul(sub(z0, 0.5), dr_max));
                rnew[1] = add(rnew[1], mul(sub(z1, 0.5), dr_max));
                rnew[2] = add(rnew[2], mul(sub(z2, 0.5), dr_max));
#pragma unroll
                for (int k = 0; k < 3; ++k) {               // boundaries.jl:16-26
                    if (rnew[k] > L) rnew[k] = sub(rnew[k], L);
                    if (rnew[k] < 0) rnew[k] = add(rnew[k], L);
                }
            } else if (chose <= A.p_rot) {                  // main.jl:530-538, quaternions.jl:52-73,94-120,158-182
                is_trans = 0;
                if (fabs(sub(add(add(add(mul(nq0, nq0), mul(nq1, nq1)), mul(nq2, nq2)), mul(nq3, nq3)), 1.0)) > 1.e-6) ret = 2;
                double ax0, ax1, ax2, nrm;
                for (;;) {
                    ax0 = sub(mul(2.0, next_u()), 1.0); ax1 = sub(mul(2.0, next_u()), 1.0); ax2 = sub(mul(2.0, next_u()), 1.0);
                    nrm = add(add(mul(ax0, ax0), mul(ax1, ax1)), mul(ax2, ax2));
                    if (nrm < 1.0 || dry) break;
                }
                const double sn = __dsqrt_rn(nrm);
                ax0 = __ddiv_rn(ax0, sn); ax1 = __ddiv_rn(ax1, sn); ax2 = __ddiv_rn(ax2, sn);
                const double zeta = next_u();
                const double angle = mul(sub(mul(2.0, zeta), 1.0), dphi_max);
                double sh, ch;
                sincos(mul(0.5, angle), &sh, &ch);
                const double r0 = ch, r1 = mul(sh, ax0), r2q = mul(sh, ax1), r3 = mul(sh, ax2);
                e0 = sub(sub(sub(mul(r0, nq0), mul(r1, nq1)), mul(r2q, nq2)), mul(r3, nq3));   // quatmul(rot, old)
                e1 = add(sub(add(mul(r1, nq0), mul(r0, nq1)), mul(r3, nq2)), mul(r2q, nq3));
                e2 = sub(add(add(mul(r2q, nq0), mul(r3, nq1)), mul(r0, nq2)), mul(r1, nq3));
                e3 = add(add(sub(mul(r3, nq0), mul(r2q, nq1)), mul(r1, nq2)), mul(r0, nq3));
            } else ret = 3;                                 // main.jl:539-541
            if (ret == 0 && fabs(sub(add(add(add(mul(e0, e0), mul(e1, e1)), mul(e2, e2)), mul(e3, e3)), 1.0)) > 1.e-6) ret = 2;
            T.ei[0] = e0; T.ei[1] = e1; T.ei[2] = e2; T.ei[3] = e3;
            // quaternions.jl:37-50 — rows as written in the reference, [2,3] = 2(q2 q4 + q1 q2)
            const double q1 = e0, q2 = e1, q3 = e2, q4 = e3;
            double a[3][3];
            a[0][0] = sub(sub(add(mul(q1, q1), mul(q2, q2)), mul(q3, q3)), mul(q4, q4));
            a[0][1] = mul(2, add(mul(q2, q3), mul(q1, q4))); a[0][2] = mul(2, sub(mul(q2, q4), mul(q1, q3)));
            a[1][0] = mul(2, sub(mul(q2, q3), mul(q1, q4)));
            a[1][1] = sub(add(sub(mul(q1, q1), mul(q2, q2)), mul(q3, q3)), mul(q4, q4));
            a[1][2] = mul(2, add(mul(q2, q4), mul(q1, q2)));
            a[2][0] = mul(2, add(mul(q2, q4), mul(q1, q3))); a[2][1] = mul(2, sub(mul(q3, q4), mul(q1, q2)));
            a[2][2] = add(sub(sub(mul(q1, q1), mul(q2, q2)), mul(q3, q3)), mul(q4, q4));
#pragma unroll
            for (int s = 0; s < S; ++s) {                   // main.jl:545-548: COM + MATMUL(ai, db)
                const double d0 = s_gdb[g][3 * s], d1 = s_gdb[g][3 * s + 1], d2 = s_gdb[g][3 * s + 2];
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    T.site[s][c] = add(rnew[c], add(add(mul(d0, a[0][c]), mul(d1, a[1][c])), mul(d2, a[2][c])));
            }
            T.com[0] = rnew[0]; T.com[1] = rnew[1]; T.com[2] = rnew[2];
            T.pos_begin = p0; T.pos_end = pos; T.is_trans = is_trans; T.ret = ret; T.dry = dry ? 1 : 0;
        }
        __syncwarp();
        if (A.style_recip && lane < 2 * S * 3) {            // ewalds.jl:770-795 for the old and the trial sites
            const int cfg = lane / (S * 3), rem = lane - cfg * S * 3, l = rem / 3, d = rem - 3 * l;
            double x;
            if (cfg == 0) x = T.osite[l][d];
            else x = T.site[l][d];
            cplx e1;
            sincos(twopi * x / L, &e1.im, &e1.re);
            cplx e; e.re = 1.0; e.im = 0.0;
            s_tabc[par][g][cfg][l][d][0] = e;
            e = e1;
            s_tabc[par][g][cfg][l][d][1] = e;
            for (int k = 2; k <= nk; ++k) { e = cmul(e, e1); s_tabc[par][g][cfg][l][d][k] = e; }
        }
    };

    if (A.n_moves > 0 && gen == 0) generate(0, 0, 0, 0);    // move 0: one candidate at the exact position
    __syncthreads();

    int cur = A.cur;
    for (long long m = 0; m < A.n_moves; ++m) {
        const int i = (int)(m % N), par = (int)(m & 1), sel = s_sel;
        const ChainCand<S> &T = s_cand[par][sel];
        if (tid == 0 && T.ret != 3) { if (T.is_trans) D.tr.attempt += 1; else D.ro.attempt += 1; }
        if (T.ret) { if (tid == 0) { D.ret = T.ret; s_pos = T.pos_end; } break; }   // quaternion-norm error / no move selected: stop before the move
        // the move after this one is drawn while this one is evaluated, unless the step sizes may change in between
        const bool have_next = m + 1 < A.n_moves;
        const bool late = A.adjust && i == N - 1;
        double u_metro = 0.5;                                // the uniform Metropolis(m) would draw: requested now, used at the decision
        if (tid == 0) {
            D.is_trans = T.is_trans;
            if (T.pos_end < A.n_uniforms) u_metro = A.uniforms[T.pos_end];
        }
        if (!worker) {
            cluster.barrier_arrive();                        // generators publish nothing across SMs
            if (have_next && !late) generate(gen, m + 1, T.pos_end + gen, par ^ 1);
            cluster.barrier_wait();
        } else {
            // ================= step 1: COM gate for the old and the trial position (this CTA's partners)
            const double4 co = make_double4(T.ocom[0], T.ocom[1], T.ocom[2], 0.0);
            const double cnx = T.com[0], cny = T.com[1], cnz = T.com[2];
            int myfl[CHAINS_MAXIT]; unsigned mymask[CHAINS_MAXIT];
#pragma unroll
            for (int it = 0; it < CHAINS_MAXIT; ++it) {
                myfl[it] = 0; mymask[it] = 0;
                if (it * CHAINS_WORKERS < n_loc) {
                    const int j = j_lo + it * CHAINS_WORKERS + tid;
                    int fl = 0;
                    if (j < j_hi && j != i) {
                        const double4 cj = s_com[j - s_lo];
                        {
                            const double rx = min_image(co.x, cj.x, L), ry = min_image(co.y, cj.y, L), rz = min_image(co.z, cj.z, L);
                            const double r2 = add(add(mul(rx, rx), mul(ry, ry)), mul(rz, rz));
                            if (r2 < rc_lj2) fl |= 1;
                            if (A.style_qq && r2 < rc_qq2) fl |= 2;
                        }
                        {
                            const double rx = min_image(cnx, cj.x, L), ry = min_image(cny, cj.y, L), rz = min_image(cnz, cj.z, L);
                            const double r2 = add(add(mul(rx, rx), mul(ry, ry)), mul(rz, rz));
                            if (r2 < rc_lj2) fl |= 4;
                            if (A.style_qq && r2 < rc_qq2) fl |= 8;
                        }
                    }
                    const unsigned mk = __ballot_sync(0xffffffffu, fl != 0);
                    if (lane == 0) s_wcount[it * CHAINS_WWARPS + warp] = __popc(mk);
                    myfl[it] = fl; mymask[it] = mk;
                } else if (lane == 0) s_wcount[it * CHAINS_WWARPS + warp] = 0;
            }
            chain_worker_bar();
            int n_in = 0;
            {   // ordered compaction: survivors of iteration `it`, warp `warp` go after everything with a smaller (it, warp)
#pragma unroll
                for (int it = 0; it < CHAINS_MAXIT; ++it) {
                    if (it * CHAINS_WORKERS < n_loc) {
                        int below = 0, tot = 0;
#pragma unroll
                        for (int w = 0; w < CHAINS_WWARPS; ++w) { const int c = s_wcount[it * CHAINS_WWARPS + w]; tot += c; if (w < warp) below += c; }
                        if (myfl[it]) s_list[n_in + below + __popc(mymask[it] & lt)] = make_int2(j_lo + it * CHAINS_WORKERS + tid, myfl[it]);
                        n_in += tot;
                    }
                }
            }
            chain_worker_bar();

            // ================= step 2: site pairs of (partner, cfg, site a) items + ρ(k) delta of this CTA's k-vectors
            double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            const int items = n_in * 2 * S;
            for (int w = tid; w < items; w += CHAINS_WORKERS) {
                const int jj = w / (2 * S), rem = w - jj * 2 * S, cfg = rem / S, a = rem - cfg * S;
                const int2 e = s_list[jj];
                const int j = e.x, fl = (e.y >> (2 * cfg)) & 3;
                if (!fl) continue;
                double4 sa = make_double4(T.osite[a][0], T.osite[a][1], T.osite[a][2], T.q[a]);
                double cix = co.x, ciy = co.y, ciz = co.z;
                if (cfg) { sa.x = T.site[a][0]; sa.y = T.site[a][1]; sa.z = T.site[a][2]; cix = cnx; ciy = cny; ciz = cnz; }
                double r2[S], dx[S], dy[S], dz[S], qq[S], ri[S];
#pragma unroll
                for (int b = 0; b < S; ++b) {
                    const double4 sb = s_site[(j - s_lo) * S + b];
                    dx[b] = min_image(sa.x, sb.x, L); dy[b] = min_image(sa.y, sb.y, L); dz[b] = min_image(sa.z, sb.z, L);
                    r2[b] = dx[b] * dx[b] + dy[b] * dy[b] + dz[b] * dz[b];
                    qq[b] = sa.w * sb.w;
                    ri[b] = chain_rsqrt(r2[b]);
                }
                double l0 = 0.0, l1 = 0.0, l2 = 0.0, l3 = 0.0;
                if (fl & 1) {                               // energy.jl:270-282
                    const int ta = s_type[a];
                    const double4 cj = s_com[j - s_lo];
                    const double rijx = min_image(cix, cj.x, L), rijy = min_image(ciy, cj.y, L), rijz = min_image(ciz, cj.z, L);
#pragma unroll
                    for (int b = 0; b < S; ++b) {
                        const int tb = s_type[b];
                        const double eps = Sy.eps[ta + tb * nt];
                        if (r2[b] < (rc_lj2 + 100) && eps > 0.001) {      // σ²/r² formed as σ²·(1/√r²)²: no division
                            const double sig = Sy.sig[ta + tb * nt];
                            const double s2 = sig * sig * (ri[b] * ri[b]), s6 = s2 * s2 * s2, s12 = s6 * s6;
                            l0 += eps * (s12 - s6);
                            const double virab = eps * (2.0 * s12 - s6) * s2;
                            l1 += rijx * (dx[b] * virab) + rijy * (dy[b] * virab) + rijz * (dz[b] * virab);
                        }
                    }
                }
                if (fl & 2) {                               // ewalds.jl:359-367
                    bool use[S];
#pragma unroll
                    for (int b = 0; b < S; ++b) {
                        use[b] = false;
                        if ((r2[b] < 0.5) && (qq[b] < 0)) l3 = 1.0;
                        else if (r2[b] < rc_qq2 + 100) use[b] = true;
                    }
                    if (DEG != 0) {                         // erfc(κr)/r = 1/r − κ·E(κ²r²), S chains in lock-step
                        double sv[S], pv[S];
                        const int deg = DEG > 0 ? DEG : P.deg;
#pragma unroll
                        for (int b = 0; b < S; ++b) { sv[b] = fma(r2[b] * P.kappa2, P.scale, -1.0); pv[b] = P.c[deg]; }
                        if (DEG > 0) {
#pragma unroll
                            for (int k = DEG - 1; k >= 0; --k)
#pragma unroll
                                for (int b = 0; b < S; ++b) pv[b] = fma(pv[b], sv[b], P.c[k]);
                        } else {
#pragma unroll 1
                            for (int k = deg - 1; k >= 0; --k) {
                                const double ck = P.c[k];
#pragma unroll
                                for (int b = 0; b < S; ++b) pv[b] = fma(pv[b], sv[b], ck);
                            }
                        }
#pragma unroll
                        for (int b = 0; b < S; ++b) if (use[b]) l2 = fma(qq[b], fma(-P.kappa, pv[b], ri[b]), l2);
                    } else {
#pragma unroll
                        for (int b = 0; b < S; ++b) if (use[b]) { const double r = sqrt(r2[b]); l2 += qq[b] * erfc(Sy.kappa * r) / r; }
                    }
                }
                if (cfg) { acc[4] += l0; acc[5] += l1; acc[6] += l2; acc[7] += l3; }
                else { acc[0] += l0; acc[1] += l1; acc[2] += l2; acc[3] += l3; }
            }
            if (A.style_recip) {                            // ewalds.jl:804-821
                const double2 *Sold = s_rhok[cur];
                double2 *Snew = s_rhok[cur ^ 1];
                for (int k = k_lo + (CHAINS_WORKERS - 1 - tid); k < k_hi; k += CHAINS_WORKERS) {   // from the far end: the low warps carry the pair items
                    const int4 kv = s_kvec[k];
                    const int aky = abs(kv.y), akz = abs(kv.z);
                    const bool ny = kv.y < 0, nz = kv.z < 0;
                    const double2 so = Sold[k];
                    double nr = so.x, ni = so.y;
#pragma unroll
                    for (int l = 0; l < S; ++l) {
                        const cplx tn = cmul(cmul(s_tabc[par][sel][1][l][0][kv.x], cconj_if(s_tabc[par][sel][1][l][1][aky], ny)), cconj_if(s_tabc[par][sel][1][l][2][akz], nz));
                        const cplx to = cmul(cmul(s_tabc[par][sel][0][l][0][kv.x], cconj_if(s_tabc[par][sel][0][l][1][aky], ny)), cconj_if(s_tabc[par][sel][0][l][2][akz], nz));
                        const double q = T.q[l];
                        nr += q * (tn.re - to.re);
                        ni += q * (tn.im - to.im);
                    }
                    Snew[k] = make_double2(nr, ni);
                    acc[8] += s_cfac[k] * ((nr * nr + ni * ni) - (so.x * so.x + so.y * so.y));
                }
            }
            // ================= step 3: ordered reduction, exchange, decision, state update
#pragma unroll
            for (int v = 0; v < 9; ++v) {
                acc[v] = warp_sum(acc[v]);
                if (lane == 0) s_red[v * 8 + warp] = acc[v];
            }
            chain_worker_bar();
            if (warp == 0) {   // partial sums (s_red[v][warp], slots 6, 7 are zero) folded by a fixed shuffle tree, then into every CTA's buffer
                double x0 = s_red[lane], x1 = s_red[lane + 32], x2 = lane < 8 ? s_red[lane + 64] : 0.0;
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) {
                    x0 += __shfl_xor_sync(0xffffffffu, x0, o); x1 += __shfl_xor_sync(0xffffffffu, x1, o); x2 += __shfl_xor_sync(0xffffffffu, x2, o);
                }
                const int v0 = lane >> 3, d = lane & 7;       // every lane of a group holds the group's sum: lane (v0, d) serves CTA d
                if (d < C) {
                    double *dst = cluster.map_shared_rank(&s_xchg[par][0][0], d);
                    dst[v0 * CHAINC_MAXC + rank] = x0;
                    dst[(4 + v0) * CHAINC_MAXC + rank] = x1;
                    if (v0 == 0) dst[8 * CHAINC_MAXC + rank] = x2;
                }
            }
            cluster.barrier_arrive();
            cluster.barrier_wait();                          // every CTA's nine sums are in every CTA's buffer
            if (warp == 0) {
                double tot[9];
                {   // the C slots of every value added by the same fixed tree in every CTA: bit-identical totals everywhere
                    const double *xb = &s_xchg[par][0][0];
                    double x0 = xb[lane], x1 = xb[lane + 32], x2 = lane < 8 ? xb[lane + 64] : 0.0;
#pragma unroll
                    for (int o = 1; o < 8; o <<= 1) {
                        x0 += __shfl_xor_sync(0xffffffffu, x0, o); x1 += __shfl_xor_sync(0xffffffffu, x1, o); x2 += __shfl_xor_sync(0xffffffffu, x2, o);
                    }
                    if ((lane & 7) == 0) { s_tot[lane >> 3] = x0; s_tot[4 + (lane >> 3)] = x1; if (lane == 0) s_tot[8] = x2; }
                    __syncwarp();
#pragma unroll
                    for (int v = 0; v < 9; ++v) tot[v] = s_tot[v];
                    __syncwarp();
                }
                // launch_move_on's folding (energy.jl:289 pot*4, vir*24/3; ewalds.jl:360; main.jl:580-590) + mmc_trial_move
                const bool ovl0 = tot[3] > 0.0, ovl1 = tot[7] > 0.0, overlap = ovl0 || ovl1;
                const double lj_old = mul(tot[0], 4), lj_new = mul(tot[4], 4);
                const double qq_old = mul(ovl0 ? 0.0 : tot[2], Sy.factor), qq_new = mul(ovl1 ? 0.0 : tot[6], Sy.factor);
                const double d_recip = (overlap || !A.style_recip) ? 0.0 : mul(tot[8], Sy.factor);
                double old_e = lj_old, new_e = lj_new;
                if (A.style_qq) { old_e = add(old_e, qq_old); new_e = add(new_e, qq_new); }     // main.jl:501-505, 566-570
                const double delta = add(sub(new_e, old_e), d_recip);                             // main.jl:593
                double num = delta, den = A.temperature, rden = A.inv_temperature;                // lane 0: delta / T
                if (lane == 1) { num = mul(tot[1], 24); den = 3.0; rden = 1.0 / 3.0; }            // lj_vir_old
                if (lane == 2) { num = mul(tot[5], 24); den = 3.0; rden = 1.0 / 3.0; }            // lj_vir_new
                if (lane == 3) { num = qq_old; den = 3.0; rden = 1.0 / 3.0; }                     // ewalds.jl:907 virial = E/3
                if (lane == 4) { num = qq_new; den = 3.0; rden = 1.0 / 3.0; }
                if (lane == 5) { num = d_recip; den = 3.0; rden = 1.0 / 3.0; }
                const double q0 = mul(num, rden);                                                 // rounded quotient via the rounded reciprocal
                const double quo = fma(fma(-q0, den, num), rden, q0);                             // and one exact-remainder correction
                const double ljv_old = __shfl_sync(0xffffffffu, quo, 1), ljv_new = __shfl_sync(0xffffffffu, quo, 2);
                const double qqv_old = __shfl_sync(0xffffffffu, quo, 3), qqv_new = __shfl_sync(0xffffffffu, quo, 4);
                const double recv = __shfl_sync(0xffffffffu, quo, 5);
                if (lane == 0) {
                    double old_v = ljv_old, new_v = ljv_new;
                    if (A.style_qq) { old_v = add(old_v, qqv_old); new_v = add(new_v, qqv_new); }
                    const double x = quo;
                    if (overlap) D.n_ovl += 1;
                    long long pos = T.pos_end;
                    bool dry = T.dry != 0;
                    bool okm = true;
                    int drew = 0;
                    if (!(x < 0.0)) {                                                         // auxillary.jl:106-114
                        if (pos >= A.n_uniforms) { dry = true; u_metro = 0.5; } else { ++pos; drew = 1; }
                        okm = exp(-x) > u_metro;
                    }
                    // a move whose draws ran past the end of the stream never happened (see k_chain)
                    const bool accd = okm && !overlap && !dry;                                // main.jl:598
                    if (accd) {
                        D.tot_e = add(D.tot_e, delta);
                        D.tot_v = add(D.tot_v, add(sub(new_v, old_v), recv));
                        D.n_acc += 1;
                        if (T.is_trans) D.tr.naccept += 1; else D.ro.naccept += 1;
                        const bool owner = i >= j_lo && i < j_hi;
                        if (!SLICED || owner) {           // on-chip copy: every replica, or the owner's slice
                            s_com[i - s_lo] = make_double4(T.com[0], T.com[1], T.com[2], 0.0);
#pragma unroll
                            for (int s = 0; s < S; ++s) s_site[(i - s_lo) * S + s] = make_double4(T.site[s][0], T.site[s][1], T.site[s][2], T.q[s]);
                        }
                        if (owner) {                      // the resident arrays (L2) are kept current by the owner of molecule i
                            Sy.com[i] = make_double4(T.com[0], T.com[1], T.com[2], 0.0);
#pragma unroll
                            for (int s = 0; s < S; ++s) Sy.site[(size_t)i * S + s] = make_double4(T.site[s][0], T.site[s][1], T.site[s][2], T.q[s]);
                            // no fence here: the readers are generator warps of this cluster, at least N-1 moves later, and every
                            // move passes barrier.cluster arrive(release) / wait(acquire), which orders these stores for them
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) myquat[4 * i + k] = T.ei[k];                  // this CTA's own replica
                        if (A.style_recip && !overlap) s_cur = cur ^ 1;                           // main.jl:621 as an index flip
                    }
                    s_pos = pos;
                    s_sel = drew;                                                             // which candidate of move m+1 is the real one
                    if (dry) {
                        D.ret = 1; s_stop = 1; s_pos = T.pos_begin;
                        if (T.is_trans) D.tr.attempt -= 1; else D.ro.attempt -= 1;
                        if (overlap) D.n_ovl -= 1;
                    } else {
                        if (A.accepted && rank == 0) A.accepted[m] = accd ? 1 : 0;
                        if (A.delta && rank == 0) A.delta[m] = delta;
                        if (late) {                                                           // main.jl:645-651
                            D.tr.d_max = D.dr_max; adjust_step(D.tr, L); D.dr_max = D.tr.d_max;
                            D.ro.d_max = D.dphi_max; adjust_step(D.ro, L); D.dphi_max = D.ro.d_max;
                        }
                        D.n_done = m + 1;
                    }
                }
            }
        }
        __syncthreads();                                    // B5: decision, state and (speculative) candidates are in place
        cur = s_cur;
        if (s_stop) break;
        if (have_next && late) {                            // step sizes may have changed: draw move m+1 now, at the exact position
            if (gen == 0) generate(0, m + 1, s_pos, par ^ 1);
            if (tid == 0) s_sel = 0;
            __syncthreads();
        }
    }
    __syncthreads();
    // sites and COMs: the owners kept the resident arrays current; ρ(k): every CTA writes its slice
    for (int t = k_lo + tid; t < k_hi; t += CHAINC_THREADS) Sy.rhok[cur][t] = s_rhok[cur][t];
    __threadfence();
    cluster.sync();                                         // nobody leaves while a neighbour may still write into its buffer
    if (tid == 0 && rank == 0) {
        ChainOut o;
        o.n_moves = D.n_done; o.n_accepted = D.n_acc; o.n_overlap = D.n_ovl; o.uniforms_used = s_pos;
        o.trans_attempt = D.tr.attempt; o.trans_accept = D.tr.naccept; o.rot_attempt = D.ro.attempt; o.rot_accept = D.ro.naccept;
        o.dr_max = D.dr_max; o.dphi_max = D.dphi_max; o.total_energy = D.tot_e; o.total_virial = D.tot_v;
        o.ret = D.ret; o.cur = cur;
        for (int k = 0; k < 6; ++k) o.phase_cycles[k] = 0;
        *A.out = o;
    }
}

// ---------------------------------------------------------------------------------------------------------
// k_chain_atoms — Monatomic/mainMonatomic.jl:373-413 for a block of moves on a thread-block cluster.
// The reference scans all N atoms twice per move (LJ_ΔU old, LJ_ΔU new).  Here CTA r of the cluster keeps a FLOAT
// copy of its slice of the positions in shared memory (16 B per atom: 32 000 atoms = 64 KB per CTA at C = 8) and
// scans it with a conservative FP32 distance gate for the old and the trial position at once (min-image by rintf);
// the few atoms that pass (≈50 of 32 000) are re-tested and evaluated in FP64 from the resident arrays (exact rule
// `!(r² > r_cut²)`, σ_j²/r² by division as in the reference).  Four partial sums per CTA cross the cluster through
// DSMEM; one cluster barrier per move; every CTA takes the same decision redundantly; the owner of atom i updates
// the resident position (L2) and its float copy.  r_old(i+1) and the uniforms are requested one move ahead.
#define CHAINA_THREADS 512
#define CHAINA_WARPS (CHAINA_THREADS / 32)
#define CHAINA_MAXC 16       // CTAs per cluster (16 is the non-portable maximum of sm_100; 8 is the fallback)

struct ChainAtomArgs {
    long long n_moves, n_uniforms;
    double temperature, inv_temperature, dr_max, e0, v0;
    float gate_rc2f;             // r_cut² + 4× the worst-case FP32 error of the gate
    const double *uniforms;
    unsigned char *accepted;     // [n_moves] or NULL
    double *delta;               // [n_moves] or NULL
    ChainOut *out;
};

static __global__ void __launch_bounds__(CHAINA_THREADS, 1)
k_chain_atoms(DevAtoms At, const __grid_constant__ ChainAtomArgs A)
{
    using namespace chain;
    namespace cg = cooperative_groups;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *s_f = reinterpret_cast<float4 *>(smem_raw);                  // float copy of this CTA's slice
    __shared__ double s_red[4 * CHAINA_WARPS];
    __shared__ double s_xchg[2][4][CHAINA_MAXC];                         // [move parity][value][source rank]; unused ranks stay 0
    __shared__ double s_tot[4];
    __shared__ double s_u[CHAIN_RING];
    __shared__ double s_r0[3], s_rn[3];
    __shared__ int s_stop;

    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = At.n;
    const int a_lo = (int)((long long)N * rank / C), a_hi = (int)((long long)N * (rank + 1) / C), n_loc = a_hi - a_lo;
    const double L = At.box, rc2 = At.rc * At.rc;
    const float Lf = (float)L, invLf = (float)(1.0 / L), rc2f = A.gate_rc2f;

    for (int t = tid; t < n_loc; t += CHAINA_THREADS) {
        const double4 r = At.r[a_lo + t];
        s_f[t] = make_float4((float)r.x, (float)r.y, (float)r.z, 0.f);
    }
    for (int k = tid; k < 2 * 4 * CHAINA_MAXC; k += CHAINA_THREADS) (&s_xchg[0][0][0])[k] = 0.0;
    if (tid == 0) s_stop = 0;

    // driver state: lane 0 of warp 0 (a handful of registers)
    long long pos = 0, ring_end = A.n_uniforms < CHAIN_RING ? A.n_uniforms : CHAIN_RING;
    bool dry = false;
    double tot_e = A.e0, tot_v = A.v0;
    long long n_acc = 0, n_tacc = 0, n_done = 0;
    int ret = 0;
    double pre0 = 0.0, pre1 = 0.0; int pre_cnt = 0;
    double4 rnext = make_double4(0, 0, 0, 0);                              // position of the next move's atom, requested a move ahead
    if (warp == 0) {
        const long long lim = ring_end;
        if (lane < lim) s_u[lane] = A.uniforms[lane];
        if (lane + 32 < lim) s_u[lane + 32] = A.uniforms[lane + 32];
        if (lane == 0 && A.n_moves > 0) rnext = ldcg4(&At.r[0]);
    }
    __syncthreads();
    auto next_u = [&]() -> double {
        if (pos >= A.n_uniforms) { dry = true; return 0.5; }
        const double v = (pos < ring_end) ? s_u[pos & (CHAIN_RING - 1)] : A.uniforms[pos];
        ++pos;
        return v;
    };

    long long pos_m = 0;                                     // stream position at the start of the current move
    for (long long m = 0; m < A.n_moves; ++m) {
        const int i = (int)(m % N), par = (int)(m & 1);
        pos_m = pos;
        // ================= step 0: the trial position (mainMonatomic.jl:375-380), lane 0 of warp 0
        if (warp == 0) {
            if (pre_cnt > 0) {
                if (lane < pre_cnt) s_u[(ring_end + lane) & (CHAIN_RING - 1)] = pre0;
                if (lane + 32 < pre_cnt) s_u[(ring_end + 32 + lane) & (CHAIN_RING - 1)] = pre1;
                ring_end += pre_cnt;
                pre_cnt = 0;
            }
            __syncwarp();
            if (lane == 0) {
                const double z0 = next_u(), z1 = next_u(), z2 = next_u();
                double rn[3] = {add(rnext.x, mul(sub(z0, 0.5), A.dr_max)), add(rnext.y, mul(sub(z1, 0.5), A.dr_max)),
                                add(rnext.z, mul(sub(z2, 0.5), A.dr_max))};
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if (rn[k] > L) rn[k] = sub(rn[k], L);
                    if (rn[k] < 0) rn[k] = add(rn[k], L);
                }
                s_r0[0] = rnext.x; s_r0[1] = rnext.y; s_r0[2] = rnext.z;
                s_rn[0] = rn[0]; s_rn[1] = rn[1]; s_rn[2] = rn[2];
                const int inx = (i + 1 == N) ? 0 : i + 1;   // move m cannot change atom i+1; L2 load: its owner is another SM
                rnext = ldcg4(&At.r[inx]);
            }
            {
                const long long p0 = __shfl_sync(0xffffffffu, pos, 0);
                long long lim = p0 + CHAIN_RING;
                if (lim > A.n_uniforms) lim = A.n_uniforms;
                const long long want = lim - ring_end;
                pre_cnt = want > 0 ? (int)want : 0;
                if (lane < pre_cnt) pre0 = A.uniforms[ring_end + lane];
                if (lane + 32 < pre_cnt) pre1 = A.uniforms[ring_end + 32 + lane];
            }
        }
        __syncthreads();
        // ================= step 1: FP32 gate over this CTA's slice, FP64 evaluation of what passes
        const double x0 = s_r0[0], y0 = s_r0[1], z0 = s_r0[2], xn = s_rn[0], yn = s_rn[1], zn = s_rn[2];
        const float fx0 = (float)x0, fy0 = (float)y0, fz0 = (float)z0, fxn = (float)xn, fyn = (float)yn, fzn = (float)zn;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int t = tid; t < n_loc; t += CHAINA_THREADS) {
            const float4 f = s_f[t];
            float dx = f.x - fx0, dy = f.y - fy0, dz = f.z - fz0;
            dx -= Lf * rintf(dx * invLf); dy -= Lf * rintf(dy * invLf); dz -= Lf * rintf(dz * invLf);
            const float d2o = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            float ex = f.x - fxn, ey = f.y - fyn, ez = f.z - fzn;
            ex -= Lf * rintf(ex * invLf); ey -= Lf * rintf(ey * invLf); ez -= Lf * rintf(ez * invLf);
            const float d2n = fmaf(ez, ez, fmaf(ey, ey, ex * ex));
            if ((d2o <= rc2f || d2n <= rc2f) && a_lo + t != i) {
                const double4 rj = At.r[a_lo + t];
                const double2 es = At.es[a_lo + t];
                {
                    const double ax = min_image(x0, rj.x, L), ay = min_image(y0, rj.y, L), az = min_image(z0, rj.z, L);
                    const double r2 = ax * ax + ay * ay + az * az;
                    if (!(r2 > rc2)) {                       // mainMonatomic.jl:249
                        const double sr2 = es.y * es.y / r2, sr6 = sr2 * sr2 * sr2, sr12 = sr6 * sr6;
                        acc[0] += es.x * (sr12 - sr6);
                        acc[1] += es.x * (2 * sr12 - sr6);
                    }
                }
                {
                    const double ax = min_image(xn, rj.x, L), ay = min_image(yn, rj.y, L), az = min_image(zn, rj.z, L);
                    const double r2 = ax * ax + ay * ay + az * az;
                    if (!(r2 > rc2)) {
                        const double sr2 = es.y * es.y / r2, sr6 = sr2 * sr2 * sr2, sr12 = sr6 * sr6;
                        acc[2] += es.x * (sr12 - sr6);
                        acc[3] += es.x * (2 * sr12 - sr6);
                    }
                }
            }
        }
        // ================= step 2: ordered reduction, exchange, decision
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            acc[v] = warp_sum(acc[v]);
            if (lane == 0) s_red[v * CHAINA_WARPS + warp] = acc[v];
        }
        __syncthreads();
        if (warp == 0) {   // s_red[v][16 warps]: fixed shuffle tree over groups of 16 lanes, then into every CTA's buffer
            double xa = s_red[lane], xb = s_red[lane + 32];
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) { xa += __shfl_xor_sync(0xffffffffu, xa, o); xb += __shfl_xor_sync(0xffffffffu, xb, o); }
            const int d = lane & 15, vh = lane >> 4;          // lane (vh, d): values vh and 2 + vh for CTA d
            if (d < C) {
                double *dst = cluster.map_shared_rank(&s_xchg[par][0][0], d);
                dst[vh * CHAINA_MAXC + rank] = xa;
                dst[(2 + vh) * CHAINA_MAXC + rank] = xb;
            }
        }
        cluster.sync();
        if (warp == 0) {
            // 4 values x 16 rank slots = 64 doubles: two per lane, groups of 16 lanes (same fixed tree in every CTA)
            double x0 = (&s_xchg[par][0][0])[lane], x1 = (&s_xchg[par][0][0])[lane + 32];
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) { x0 += __shfl_xor_sync(0xffffffffu, x0, o); x1 += __shfl_xor_sync(0xffffffffu, x1, o); }
            if ((lane & 15) == 0) { s_tot[lane >> 4] = x0; s_tot[2 + (lane >> 4)] = x1; }
            __syncwarp();
            if (lane == 0) {
                // launch_move_atom's folding (mainMonatomic.jl:271): pot*4, vir*24/3
                const double lj_old = mul(s_tot[0], 4.0), v_old = __ddiv_rn(mul(s_tot[1], 24.0), 3.0);
                const double lj_new = mul(s_tot[2], 4.0), v_new = __ddiv_rn(mul(s_tot[3], 24.0), 3.0);
                const double delta = sub(lj_new, lj_old);
                const double q0 = mul(delta, A.inv_temperature);
                const double x = fma(fma(-q0, A.temperature, delta), A.inv_temperature, q0);   // delta / T
                bool okm = true;
                if (!(x < 0.0)) okm = exp(-x) > next_u();    // Metropolis
                if (dry) okm = false;                        // the stream ran dry inside this move: it never happened
                if (okm) {
                    tot_e = add(tot_e, delta);
                    tot_v = add(tot_v, sub(v_new, v_old));
                    n_acc += 1;
                    if (i >= a_lo && i < a_hi) {             // the owner of atom i updates the resident position and its float copy
                        At.r[i] = make_double4(s_rn[0], s_rn[1], s_rn[2], 0.0);
                        s_f[i - a_lo] = make_float4((float)s_rn[0], (float)s_rn[1], (float)s_rn[2], 0.f);
                    }
                }
                if (dry) { ret = 1; s_stop = 1; pos = pos_m; }
                else {
                    if (A.accepted && rank == 0) A.accepted[m] = okm ? 1 : 0;
                    if (A.delta && rank == 0) A.delta[m] = delta;
                    n_done = m + 1; if (okm) n_tacc += 1;
                }
            }
        }
        __syncthreads();
        if (s_stop) break;
    }
    __threadfence();
    cluster.sync();
    if (tid == 0 && rank == 0) {
        ChainOut o;
        o.n_moves = n_done; o.n_accepted = n_acc; o.n_overlap = 0; o.uniforms_used = pos;
        o.trans_attempt = n_done; o.trans_accept = n_tacc; o.rot_attempt = 0; o.rot_accept = 0;
        o.dr_max = A.dr_max; o.dphi_max = 0.0; o.total_energy = tot_e; o.total_virial = tot_v;
        o.ret = ret; o.cur = 0;
        for (int k = 0; k < 6; ++k) o.phase_cycles[k] = 0;
        *A.out = o;
    }
}
