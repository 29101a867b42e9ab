// kernels_pairs_v5.cuh — production pair kernel for cell mode, 3-site molecules with identical
// per-site charges (SPC/E, TIP3P).  Reference semantics: (COM gate energy.jl:250 /
// ewalds.jl:337, nine erfc site pairs ewalds.jl:359-367, O–O LJ energy.jl:270-282, overlap flags)
// and the unit / queue / row layout of its predecessors (round 1: k_pairs_v3/v4, removed); what changed is what the ncu
// source page of v4 showed (profiles/r01_v4_ncu_full_pairs.txt): only 31 % of its 480 M warp instructions were FP64
// arithmetic — the COM gate took ≈80 instructions per 64 tests and the consume phase ≈100
// non-FP64 instructions per round of 32 molecule pairs.
//
//   gate     = the B gate coordinates are padded with far-away sentinels to a multiple of 64, the
//              "home cell against itself: only q > p" rule is one integer compare against
//              q_min = min(p, self_n − 1), queue entries are formed by one add from a per-row base
//              and stored by predicated STS — no bounds tests, no divergent branches: ≈35
//              instructions per 64 tests;
//   overlap  = one integer min over the nine r² high words decides whether the sign-aware rule
//              (ewalds.jl:359) has to be looked at at all (never, in a sane configuration);
//   erf      = DIRECT: for big boxes (v_max = κ²(r_cut²+100) ≲ 1) the smooth part −κ·E(κ²r²) is one
//              polynomial in r² itself (κ folded into the coefficients on the host, exact degree):
//              7 DFMA per site pair at config E instead of 2 + 8 + 1; otherwise the mapped
//              Chebyshev form with σκ² folded into one constant.
#pragma once
#include "kernels_pairs.cuh"

// ---- shared by the cell-mode water kernels: unit = (home cell, group of 5/5/4 half-shell slots)
#define V3_ACAP 64
#define V3_SLOTS 5
#define V3_BCAP (V3_SLOTS * V3_ACAP)
#define V3_QCAP 512    // ring buffer per warp (power of two): ≤ 31 left-overs + 8 rows x 32 new entries
#define V3_GROUPS 3

static __constant__ int c_v3_group_begin[V3_GROUPS + 1] = {0, 5, 10, 14};

// per (cell, slot): where the neighbour cell's molecules are and how they are shifted
static __global__ void k_slots_build(PairArgs A, int4 *slots, int ncell)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncell * 14) return;
    const int c = t / 14, slot = t - 14 * c;
    const int n = A.ncd;
    const int cx = c % n, cy = (c / n) % n, cz = c / (n * n);
    int nx = cx + c_half_shell[slot][0], ny = cy + c_half_shell[slot][1], nz = cz + c_half_shell[slot][2];
    int code = 0;
    if (nx >= n) { nx -= n; code |= 1 << 0; } else if (nx < 0) { nx += n; code |= 2 << 0; }
    if (ny >= n) { ny -= n; code |= 1 << 2; } else if (ny < 0) { ny += n; code |= 2 << 2; }
    if (nz >= n) { nz -= n; code |= 1 << 4; } else if (nz < 0) { nz += n; code |= 2 << 4; }
    const int cb = nx + n * (ny + n * nz);
    const int b_lo = A.cell_start[cb];
    slots[t] = make_int4(b_lo, A.cell_start[cb + 1] - b_lo, code, 0);
}

// MUFU.RSQ64H seed + one cubic Newton step (relative error ≈ 1e-19·… → correctly rounded to ~1 ulp for
// normal positive r²; r² = 0 gives +inf like 1/sqrt(0)): the CUDA rsqrt() sequence without its
// special-value slow path, which this kernel never needs.
__device__ __forceinline__ double fast_rsqrt(double x)
{
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    const double e = fma(x, -(y0 * y0), 1.0);
    return fma(fma(e, 0.375, 0.5), y0 * e, y0);
}


#define V5_BLOCK 128
#define V5_WARPS (V5_BLOCK / 32)
#define V5_ACAP 64
#define V5_SLOTS 5
#define V5_BCAP (V5_SLOTS * V5_ACAP)
#define V5_ROW 12      // doubles per staged molecule: O xyz, H1 xyz, H2 xyz, COM xyz
#define V5_QCAP 1024   // 16-bit entries per warp (power of two); drained when fewer than V5_BCAP are free

constexpr size_t V5_SMEM = (size_t)(V5_ACAP + V5_BCAP) * V5_ROW * sizeof(double) +
                           (size_t)(V5_ACAP + V5_BCAP) * sizeof(float4) +
                           (size_t)V5_WARPS * V5_QCAP * sizeof(unsigned short);

// smooth part of one Coulomb site pair, −κ·E(κ² r²): DIRECT → Horner in r², else Horner in s = σκ² r² − 1
template <int DEG, bool DIRECT>
__device__ __forceinline__ void v5_poly9(const PairArgs &A, const double (&r2)[9], double (&pv)[9])
{
    double x[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        x[j] = DIRECT ? r2[j] : fma(r2[j], A.pk2s, -1.0);
        pv[j] = A.pc[DEG];
    }
#pragma unroll
    for (int k = DEG - 1; k >= 0; --k)
#pragma unroll
        for (int j = 0; j < 9; ++j) pv[j] = fma(pv[j], x[j], A.pc[k]);
}

template <int DEG, bool DIRECT>
static __global__ void __launch_bounds__(V5_BLOCK, 4) k_pairs_v5(const __grid_constant__ PairArgs A, const int4 *__restrict__ slots)
{
    constexpr int S = 3;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_rowA = reinterpret_cast<double *>(smem_raw);
    double *s_rowB = s_rowA + V5_ACAP * V5_ROW;
    float4 *s_fA = reinterpret_cast<float4 *>(s_rowB + V5_BCAP * V5_ROW);
    float4 *s_fB = s_fA + V5_ACAP;
    unsigned short *s_queue = reinterpret_cast<unsigned short *>(s_fB + V5_BCAP);
    __shared__ int s_boff[V5_SLOTS + 1], s_bglob[V5_SLOTS];
    __shared__ double s_red[4 * V5_WARPS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned short *q = s_queue + warp * V5_QCAP;
    const double L = A.L;
    {   // the always-true cut-off tests (energy.jl:270, ewalds.jl:362) must really be always true for this state
        const double reach = sqrt(A.rc_qq2) + 2.0 * (*A.max_dev);
        if (!(reach * reach < A.rc_qq2 + 100.0) && tid == 0 && blockIdx.x == 0) atomicExch(A.err_flag, 1u);
    }
    const double lj_eps = A.lj_eps_tab[0], lj_sig2 = A.lj_sig_tab[0] * A.lj_sig_tab[0];
    const double edge = L / (double)A.ncd;
    const float rc2f = A.gate_rc2f;
    const unsigned lt = (1u << lane) - 1u;

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
        bool bad = nA > V5_ACAP;
        if (bad) nA = 0;
        // ---- stage A (home cell): 3 site rows + 1 COM row per molecule
        for (int t = tid; t < nA * S; t += V5_BLOCK) {
            const double4 v = A.site[(size_t)a_lo * S + t];
            const int m = t / S, k = t - m * S;
            double *d = s_rowA + m * V5_ROW + 3 * k;
            d[0] = v.x; d[1] = v.y; d[2] = v.z;
        }
        for (int t = tid; t < nA; t += V5_BLOCK) {
            const double4 v = A.com[a_lo + t];
            double *d = s_rowA + t * V5_ROW + 9;
            d[0] = v.x; d[1] = v.y; d[2] = v.z;
            s_fA[t] = make_float4((float)(v.x - ox), (float)(v.y - oy), (float)(v.z - oz), 0.f);
        }
        // ---- stage B (the group's neighbour cells), translated by the slot's periodic shift
        int nB = 0;
        for (int s = sl0; s < sl1; ++s) {
            const int4 si = slots[c * 14 + s];
            const int cnt = si.y;
            if (cnt > V5_ACAP) { bad = true; continue; }
            const int cx = si.z & 3, cy = (si.z >> 2) & 3, cz = (si.z >> 4) & 3;
            const double shx = cx == 1 ? L : (cx == 2 ? -L : 0.0), shy = cy == 1 ? L : (cy == 2 ? -L : 0.0),
                         shz = cz == 1 ? L : (cz == 2 ? -L : 0.0);
            for (int t = tid; t < cnt * S; t += V5_BLOCK) {
                const double4 v = A.site[(size_t)si.x * S + t];
                const int m = t / S, k = t - m * S;
                double *d = s_rowB + (nB + m) * V5_ROW + 3 * k;
                d[0] = v.x + shx; d[1] = v.y + shy; d[2] = v.z + shz;
            }
            for (int t = tid; t < cnt; t += V5_BLOCK) {
                const double4 v = A.com[si.x + t];
                const double x = v.x + shx, y = v.y + shy, z = v.z + shz;
                double *d = s_rowB + (nB + t) * V5_ROW + 9;
                d[0] = x; d[1] = y; d[2] = z;
                s_fB[nB + t] = make_float4((float)(x - ox), (float)(y - oy), (float)(z - oz), 0.f);
            }
            nB += cnt;
        }
        if (bad) { if (tid == 0) atomicExch(A.err_flag, 1u); nA = 0; }
        // sentinels: the gate walks B in steps of 64 without bounds tests (V5_BCAP is a multiple of 64)
        if (tid < 64 && nB + tid < ((nB + 63) & ~63)) s_fB[nB + tid] = make_float4(1e18f, 1e18f, 1e18f, 0.f);
        __syncthreads();
        const int self_n = (g == 0) ? nA : 0;              // slot 0 of group 0 is the home cell itself: keep q > p

        int head = 0, tail = 0;                            // warp-private ring window [head, tail)
        auto consume = [&](int base, int count) {          // `count` queued molecule pairs, one per lane
            const bool have = lane < count;
            const unsigned e = have ? q[(base + lane) & (V5_QCAP - 1)] : 0u;
            const int p = e & 63u, qi = e >> 6;
            const double2 *ra = reinterpret_cast<const double2 *>(s_rowA + p * V5_ROW);
            const double2 *rb = reinterpret_cast<const double2 *>(s_rowB + qi * V5_ROW);
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
                double r2[S * S], pv[S * S], ri[S * S];
                double ddx = 0, ddy = 0, ddz = 0;            // O–O separation for the LJ term
                int hmin = 0x7fffffff;
#pragma unroll
                for (int a = 0; a < S; ++a)
#pragma unroll
                    for (int b = 0; b < S; ++b) {
                        const int j = a * S + b;
                        const double dx = bx[b] - ax[a], dy = by[b] - ay[a], dz = bz[b] - az[a];
                        if (j == 0) { ddx = dx; ddy = dy; ddz = dz; }
                        r2[j] = dx * dx + dy * dy + dz * dz;
                        hmin = min(hmin, __double2hiint(r2[j]));
                    }
                unsigned ovl = 0;
                if (hmin < 0x3FE00000) {   // some site pair has r² < 0.5: apply the sign rule q_a q_b < 0 (ewalds.jl:359)
#pragma unroll
                    for (int j = 0; j < S * S; ++j)
                        if (((A.qq_negmask >> j) & 1u) && __double2hiint(r2[j]) < 0x3FE00000) { ovl |= 1u << j; r2[j] = 1.0; }
                }
#pragma unroll
                for (int j = 0; j < S * S; ++j) ri[j] = fast_rsqrt(r2[j]);
                v5_poly9<DEG, DIRECT>(A, r2, pv);
#pragma unroll
                for (int j = 0; j < S * S; ++j) acc_q = fma(A.qq_tab[j], ri[j] + pv[j], acc_q);   // ewalds.jl:366-367
                {   // LJ 12-6 on the O–O pair (energy.jl:270-282), virial with the COM separation
                    const double rinv2 = ri[0] * ri[0];
                    const double s2 = lj_sig2 * rinv2, s6 = s2 * s2 * s2, s12 = s6 * s6;
                    acc_lj += lj_eps * (s12 - s6);
                    const double w = lj_eps * (2.0 * s12 - s6) * s2;
                    acc_vir += w * (rx * ddx + ry * ddy + rz * ddz);
                }
                if (ovl) {                                                           // ewalds.jl:359-360
                    // an overlapping site pair was evaluated at r² = 1 and is removed again: 1/√1 + P(1)
                    double f_one = A.pc[DEG];
                    const double x1 = DIRECT ? 1.0 : fma(1.0, A.pk2s, -1.0);
#pragma unroll
                    for (int k = DEG - 1; k >= 0; --k) f_one = fma(f_one, x1, A.pc[k]);
                    f_one = fast_rsqrt(1.0) + f_one;
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
        for (int p = (warp + (int)u) & (V5_WARPS - 1);; p += V5_WARPS) {
            const bool last = p >= nA;
            if (!last) {
                const float4 fa = s_fA[p];
                const int qmin = min(p, self_n - 1) - lane;      // home cell against itself: only q > p
                unsigned e0 = (unsigned)p | ((unsigned)lane << 6);
                const float4 *fb = s_fB + lane;
                for (int qb = 0; qb < nB; qb += 64, e0 += 64u << 6) {
                    const float4 f0 = fb[qb], f1 = fb[qb + 32];
                    const float dx0 = f0.x - fa.x, dy0 = f0.y - fa.y, dz0 = f0.z - fa.z;
                    const float dx1 = f1.x - fa.x, dy1 = f1.y - fa.y, dz1 = f1.z - fa.z;
                    const float d0 = fmaf(dz0, dz0, fmaf(dy0, dy0, dx0 * dx0));
                    const float d1 = fmaf(dz1, dz1, fmaf(dy1, dy1, dx1 * dx1));
                    const bool pass0 = (d0 < rc2f) && (qb > qmin);
                    const bool pass1 = (d1 < rc2f) && (qb + 32 > qmin);
                    const unsigned m0 = __ballot_sync(0xffffffffu, pass0);
                    const unsigned m1 = __ballot_sync(0xffffffffu, pass1);
                    const int i0 = (tail + __popc(m0 & lt)) & (V5_QCAP - 1);
                    tail += __popc(m0);
                    const int i1 = (tail + __popc(m1 & lt)) & (V5_QCAP - 1);
                    tail += __popc(m1);
                    if (pass0) q[i0] = (unsigned short)e0;
                    if (pass1) q[i1] = (unsigned short)(e0 + (32u << 6));
                }
            }
            if (last || tail - head > V5_QCAP - V5_BCAP) {   // end of the unit, or the ring is nearly full
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
    block_sum<4, V5_BLOCK>(accp, s_red);
    if (tid == 0) A.partial[blockIdx.x] = make_double4(accp[0], accp[1], accp[2], accp[3]);
}
