// kernels_pairs_v4.cuh — production pair kernel for cell mode, 3-site molecules with identical
// per-site charges (SPC/E, TIP3P): same reference semantics as k_pairs_v3 (COM gate
// energy.jl:250 / ewalds.jl:337, nine erfc site pairs ewalds.jl:359-367, O–O LJ energy.jl:270-282,
// overlap flags), rebuilt around what ncu showed on v3: its consume phase gathers 256 B of shared
// memory per lane and molecule pair with random addresses (≈40 B/clk/SM effective), which is as
// expensive as the FP64 arithmetic itself, and its FP64 COM gate takes 31 % of the warp time.
//
//   gate   = FP32 and conservative: COMs relative to the home cell's origin as float4; a pair
//            passes when d² < r_cut²·(1+margin) with margin ≥ 8× the worst-case FP32 error, so no
//            true pair is ever lost; the exact FP64 test `|COM_ij|² < r_cut²` (strict, on the same
//            doubles as the reference) is repeated in the consume phase on the survivors, where
//            the FP64 COMs are needed anyway for the virial.  False positives (~3e-5) idle a lane.
//   order  = row-major: a warp walks its rows p of the home cell and, per row, all columns q, so
//            the queue holds runs of entries with the same p and a consume round reads its A
//            molecule by broadcast (≤ 2-3 distinct rows per round) — only B is gathered;
//   rows   = 12 doubles (3 sites xyz + COM xyz) = 96 B per molecule instead of 128 B + 32 B COM:
//            charges are per-site constants of the uniform molecule (kernel parameters), which
//            also removes nine DMULs per molecule pair;
//   CTAs   = 4 warps, 4 CTAs/SM: a barrier now waits for 3 other warps instead of 7, and three
//            other CTAs compute while one stages its tiles.
#pragma once
#include "kernels_pairs_v3.cuh"

#define V4_BLOCK 128
#define V4_WARPS (V4_BLOCK / 32)
#define V4_ACAP 64
#define V4_SLOTS 5
#define V4_BCAP (V4_SLOTS * V4_ACAP)
#define V4_ROW 12      // doubles per staged molecule: O xyz, H1 xyz, H2 xyz, COM xyz
#define V4_QCAP 1024   // 16-bit entries per warp (power of two); drained when fewer than V4_BCAP are free

constexpr size_t V4_SMEM = (size_t)(V4_ACAP + V4_BCAP) * V4_ROW * sizeof(double) +
                           (size_t)(V4_ACAP + V4_BCAP) * sizeof(float4) +
                           (size_t)V4_WARPS * V4_QCAP * sizeof(unsigned short);

template <int DEG>
__global__ void __launch_bounds__(V4_BLOCK, 4) k_pairs_v4(const __grid_constant__ PairArgs A, const int4 *__restrict__ slots)
{
    constexpr int S = 3;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_rowA = reinterpret_cast<double *>(smem_raw);
    double *s_rowB = s_rowA + V4_ACAP * V4_ROW;
    float4 *s_fA = reinterpret_cast<float4 *>(s_rowB + V4_BCAP * V4_ROW);
    float4 *s_fB = s_fA + V4_ACAP;
    unsigned short *s_queue = reinterpret_cast<unsigned short *>(s_fB + V4_BCAP);
    __shared__ int s_boff[V4_SLOTS + 1], s_bglob[V4_SLOTS];
    __shared__ double s_red[4 * V4_WARPS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned short *q = s_queue + warp * V4_QCAP;
    const double L = A.L;
    {   // the always-true cut-off tests (energy.jl:270, ewalds.jl:362) must really be always true for this state
        const double reach = sqrt(A.rc_qq2) + 2.0 * (*A.max_dev);
        if (!(reach * reach < A.rc_qq2 + 100.0) && tid == 0 && blockIdx.x == 0) atomicExch(A.err_flag, 1u);
    }
    const double lj_eps = A.lj_eps_tab[0], lj_sig2 = A.lj_sig_tab[0] * A.lj_sig_tab[0];
    const double edge = L / (double)A.ncd;
    const float rc2f = A.gate_rc2f;
    const unsigned lt = (1u << lane) - 1u;
    // value of one Coulomb term at r² = 1 (what an overlapping site pair is replaced by, then removed again)
    double f_one;
    {
        const double sv1 = fma(A.ep.kappa2, A.ep.scale, -1.0);
        double pv1 = A.ep.c[DEG];
#pragma unroll
        for (int k = DEG - 1; k >= 0; --k) pv1 = fma(pv1, sv1, A.ep.c[k]);
        f_one = fma(-A.ep.kappa, pv1, fast_rsqrt(1.0));
    }

    double acc_lj = 0.0, acc_vir = 0.0, acc_q = 0.0;
    unsigned long long my_pairs = 0;

    for (long long u = A.unit_begin + blockIdx.x; u < A.unit_end; u += gridDim.x) {
        const int c = (int)(u / V3_GROUPS), g = (int)(u - (long long)c * V3_GROUPS);
        const int sl0 = c_v3_group_begin[g], sl1 = c_v3_group_begin[g + 1];
        const int a_lo = A.cell_start[c];
        int nA = A.cell_start[c + 1] - a_lo;
        const int n = A.ncd;
        const double ox = (double)(c % n) * edge, oy = (double)((c / n) % n) * edge, oz = (double)(c / (n * n)) * edge;
        __syncthreads();                                   // everyone is done with the previous unit's tiles
        if (tid == 0) {
            int off = 0;
            for (int s = sl0; s < sl1; ++s) {
                const int4 si = slots[c * 14 + s];
                s_boff[s - sl0] = off; s_bglob[s - sl0] = si.x;
                off += si.y;
            }
            s_boff[sl1 - sl0] = off;
        }
        bool bad = nA > V4_ACAP;
        if (bad) nA = 0;
        // ---- stage A (home cell): 3 site rows + 1 COM row per molecule
        for (int t = tid; t < nA * S; t += V4_BLOCK) {
            const double4 v = A.site[(size_t)a_lo * S + t];
            const int m = t / S, k = t - m * S;
            double *d = s_rowA + m * V4_ROW + 3 * k;
            d[0] = v.x; d[1] = v.y; d[2] = v.z;
        }
        for (int t = tid; t < nA; t += V4_BLOCK) {
            const double4 v = A.com[a_lo + t];
            double *d = s_rowA + t * V4_ROW + 9;
            d[0] = v.x; d[1] = v.y; d[2] = v.z;
            s_fA[t] = make_float4((float)(v.x - ox), (float)(v.y - oy), (float)(v.z - oz), 0.f);
        }
        // ---- stage B (the group's neighbour cells), translated by the slot's periodic shift
        int nB = 0;
        for (int s = sl0; s < sl1; ++s) {
            const int4 si = slots[c * 14 + s];
            const int cnt = si.y;
            if (cnt > V4_ACAP) { bad = true; continue; }
            const int cx = si.z & 3, cy = (si.z >> 2) & 3, cz = (si.z >> 4) & 3;
            const double shx = cx == 1 ? L : (cx == 2 ? -L : 0.0), shy = cy == 1 ? L : (cy == 2 ? -L : 0.0),
                         shz = cz == 1 ? L : (cz == 2 ? -L : 0.0);
            for (int t = tid; t < cnt * S; t += V4_BLOCK) {
                const double4 v = A.site[(size_t)si.x * S + t];
                const int m = t / S, k = t - m * S;
                double *d = s_rowB + (nB + m) * V4_ROW + 3 * k;
                d[0] = v.x + shx; d[1] = v.y + shy; d[2] = v.z + shz;
            }
            for (int t = tid; t < cnt; t += V4_BLOCK) {
                const double4 v = A.com[si.x + t];
                const double x = v.x + shx, y = v.y + shy, z = v.z + shz;
                double *d = s_rowB + (nB + t) * V4_ROW + 9;
                d[0] = x; d[1] = y; d[2] = z;
                s_fB[nB + t] = make_float4((float)(x - ox), (float)(y - oy), (float)(z - oz), 0.f);
            }
            nB += cnt;
        }
        if (bad) { if (tid == 0) atomicExch(A.err_flag, 1u); nA = 0; }
        __syncthreads();
        const int self_n = (g == 0) ? nA : 0;              // slot 0 of group 0 is the home cell itself: keep q > p

        int head = 0, tail = 0;                            // warp-private ring window [head, tail)
        auto consume = [&](int base, int count) {          // `count` queued molecule pairs, one per lane
            const bool have = lane < count;
            const unsigned e = have ? q[(base + lane) & (V4_QCAP - 1)] : 0u;
            const int p = e & 63u, qi = e >> 6;
            const double2 *ra = reinterpret_cast<const double2 *>(s_rowA + p * V4_ROW);
            const double2 *rb = reinterpret_cast<const double2 *>(s_rowB + qi * V4_ROW);
            const double2 a0 = ra[0], a1 = ra[1], a2 = ra[2], a3 = ra[3], a4 = ra[4], a5 = ra[5];
            const double2 b0 = rb[0], b1 = rb[1], b2 = rb[2], b3 = rb[3], b4 = rb[4], b5 = rb[5];
            const double ax[S] = {a0.x, a1.y, a3.x}, ay[S] = {a0.y, a2.x, a3.y}, az[S] = {a1.x, a2.y, a4.x};
            const double bx[S] = {b0.x, b1.y, b3.x}, by[S] = {b0.y, b2.x, b3.y}, bz[S] = {b1.x, b2.y, b4.x};
            // exact gate on the FP64 COMs (strict <, energy.jl:250 / ewalds.jl:337)
            const double rx = b4.y - a4.y, ry = b5.x - a5.x, rz = b5.y - a5.y;
            // un-contracted, left to right, like Julia evaluates rij[1]*rij[1] + rij[2]*rij[2] + rij[3]*rij[3]
            const double r2com = __dadd_rn(__dadd_rn(__dmul_rn(rx, rx), __dmul_rn(ry, ry)), __dmul_rn(rz, rz));
            const bool act = have && (__double_as_longlong(r2com) < A.rcqq_bits);
            const unsigned am = __ballot_sync(0xffffffffu, act);
            if (lane == 0) my_pairs += __popc(am);
            if (act) {
                double r2[S * S], sv[S * S], pv[S * S], ri[S * S];
                double ddx = 0, ddy = 0, ddz = 0;            // O–O separation for the LJ term
                unsigned ovl = 0;
#pragma unroll
                for (int a = 0; a < S; ++a)
#pragma unroll
                    for (int b = 0; b < S; ++b) {
                        const int j = a * S + b;
                        const double dx = bx[b] - ax[a], dy = by[b] - ay[a], dz = bz[b] - az[a];
                        if (j == 0) { ddx = dx; ddy = dy; ddz = dz; }
                        r2[j] = dx * dx + dy * dy + dz * dz;
                        // overlap rule r² < 0.5 && q_a q_b < 0 (ewalds.jl:359): sign of q_a q_b is a launch constant
                        if (((A.qq_negmask >> j) & 1u) && __double2hiint(r2[j]) < 0x3FE00000) ovl |= 1u << j;
                    }
                if (ovl) {
#pragma unroll
                    for (int j = 0; j < S * S; ++j) if ((ovl >> j) & 1u) r2[j] = 1.0;
                }
#pragma unroll
                for (int j = 0; j < S * S; ++j) {
                    ri[j] = fast_rsqrt(r2[j]);
                    sv[j] = fma(r2[j] * A.ep.kappa2, A.ep.scale, -1.0);
                    pv[j] = A.ep.c[DEG];
                }
#pragma unroll
                for (int k = DEG - 1; k >= 0; --k)
#pragma unroll
                    for (int j = 0; j < S * S; ++j) pv[j] = fma(pv[j], sv[j], A.ep.c[k]);
#pragma unroll
                for (int j = 0; j < S * S; ++j) acc_q = fma(A.qq_tab[j], fma(-A.ep.kappa, pv[j], ri[j]), acc_q);   // ewalds.jl:366-367
                {   // LJ 12-6 on the O–O pair (energy.jl:270-282), virial with the COM separation
                    const double rinv2 = ri[0] * ri[0];
                    const double s2 = lj_sig2 * rinv2, s6 = s2 * s2 * s2, s12 = s6 * s6;
                    acc_lj += lj_eps * (s12 - s6);
                    const double w = lj_eps * (2.0 * s12 - s6) * s2;
                    acc_vir += w * (rx * ddx + ry * ddy + rz * ddz);
                }
                if (ovl) {                                                           // ewalds.jl:359-360
#pragma unroll
                    for (int j = 0; j < S * S; ++j) if ((ovl >> j) & 1u) acc_q = fma(-A.qq_tab[j], f_one, acc_q);
                    int slot_i = 0;
                    while (slot_i + 1 < sl1 - sl0 && qi >= s_boff[slot_i + 1]) ++slot_i;
                    const int qglob = s_bglob[slot_i] + (qi - s_boff[slot_i]);
                    if (atomicExch(&A.ovl[a_lo + p], 1u) == 0u) atomicAdd(A.n_ovl, 1u);
                    if (atomicExch(&A.ovl[qglob], 1u) == 0u) atomicAdd(A.n_ovl, 1u);
                }
            }
        };

        // warp w owns rows p ≡ w + u (mod 4); per row, all columns: entries of one row are contiguous in the ring.
        // Gate everything first, then consume in full rounds (one inlined copy of the consume body).
        for (int p = (warp + (int)u) & (V4_WARPS - 1);; p += V4_WARPS) {
            const bool last = p >= nA;
            if (!last) {
                const float4 fa = s_fA[p];
                const int q_first = (p < self_n) ? ((p + 1) & ~31) : 0;   // home cell: only q > p
                for (int qb = q_first; qb < nB; qb += 64) {
                    const int q0 = qb + lane, q1 = q0 + 32;
                    const float4 f0 = s_fB[q0 < nB ? q0 : 0];
                    const float4 f1 = s_fB[q1 < nB ? q1 : 0];
                    const float dx0 = f0.x - fa.x, dy0 = f0.y - fa.y, dz0 = f0.z - fa.z;
                    const float dx1 = f1.x - fa.x, dy1 = f1.y - fa.y, dz1 = f1.z - fa.z;
                    const float d0 = fmaf(dz0, dz0, fmaf(dy0, dy0, dx0 * dx0));
                    const float d1 = fmaf(dz1, dz1, fmaf(dy1, dy1, dx1 * dx1));
                    const bool pass0 = (d0 < rc2f) && (q0 < nB) && (q0 >= self_n || q0 > p);
                    const bool pass1 = (d1 < rc2f) && (q1 < nB) && (q1 >= self_n || q1 > p);
                    const unsigned m0 = __ballot_sync(0xffffffffu, pass0);
                    const unsigned m1 = __ballot_sync(0xffffffffu, pass1);
                    if (pass0) q[(tail + __popc(m0 & lt)) & (V4_QCAP - 1)] = (unsigned short)(p | (q0 << 6));
                    tail += __popc(m0);
                    if (pass1) q[(tail + __popc(m1 & lt)) & (V4_QCAP - 1)] = (unsigned short)(p | (q1 << 6));
                    tail += __popc(m1);
                }
            }
            if (last || tail - head > V4_QCAP - V4_BCAP) {   // end of the unit, or the ring is nearly full
                __syncwarp();
                while (tail - head >= 32 || (last && tail > head)) {
                    const int cnt = min(32, tail - head);    // only the unit's last round is partial
                    consume(head, cnt);
                    head += cnt;
                }
            }
            if (last) break;
        }
    }
    __syncthreads();
    double accp[4] = {acc_lj, acc_vir, acc_q, (double)my_pairs};
    block_sum<4, V4_BLOCK>(accp, s_red);
    if (tid == 0) A.partial[blockIdx.x] = make_double4(accp[0], accp[1], accp[2], accp[3]);
}
