// kernels_pairs_v6.cuh — k_pairs_v5 with the staging done by the bulk-copy engine.
//
// ncu on v5 (profiles/r01_ncu_full_pairs_v5.txt): the per-unit staging loops were 16 % of the
// instructions but 45 % of all warp stall samples — every thread ran LDG → DADD → STS chains behind
// a dependent descriptor load, slot after slot, and the other warps sat in the two barriers around
// them.  Here k_gather additionally writes the cell-sorted state in the kernel's own row format
//     rows[p] = {O xyz, H1 xyz, H2 xyz, COM xyz}   (12 doubles = 96 B, contiguous per cell)
//     gf[p]   = COM − origin of its own cell as float4 (gate coordinates)
// so that a unit's tiles are ≤ 12 contiguous byte ranges: lanes 0–5 of warp 0 issue one
// `cp.async.bulk.shared::cluster.global` pair each (UBLKCP) against one mbarrier, with the
// descriptors prefetched during the previous unit; nobody executes a staging loop.  What is left
// after the data has landed is a short fix-up: gate coordinates of B get their slot's cell offset
// (cell-local + k·edge, so the periodic wrap is implicit), and only the slots that cross the
// periodic boundary get ±L added to their rows (same `x ± L` doubles as v4/v5).
#pragma once
#include "kernels_pairs_v5.cuh"

#define V6_BLOCK 128
#define V6_WARPS (V6_BLOCK / 32)
#define V6_ACAP 64
#define V6_SLOTS 5
#define V6_BCAP 256    // B tile: the group's ≤5 neighbour cells normally fit (5 x 37 at liquid density); else two passes
#define V6_ROW 12
#define V6_QCAP 512    // 16-bit entries per warp (power of two); drained when fewer than V6_BCAP are free

constexpr size_t V6_SMEM = (size_t)(V6_ACAP + V6_BCAP) * V6_ROW * sizeof(double) +
                           (size_t)(V6_ACAP + V6_BCAP) * sizeof(float4) +
                           (size_t)V6_WARPS * V6_QCAP * sizeof(unsigned short);

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct V6Extra {
    const double *rows;     // [n_mol][12], cell-sorted
    const float4 *gf;       // [n_mol], COM relative to its own cell's origin
    // dynamic unit scheduling (NULL: units dealt round-robin, one partial per CTA).  Cells hold 9..64 molecules on the lattice
    // start, units differ 10x in cost, and a static deal leaves the slowest CTA 8 % (one GPU) to 50 % (one rank of eight: 3.5
    // units per CTA) behind the mean.  With tickets a CTA takes the next unit when it is done with its own; every unit's sums
    // are written per (unit, warp) and folded in unit order afterwards, so the result does not depend on which CTA ran what.
    unsigned int *ticket;   // zeroed before the launch
    double4 *unit_partial;  // [units of this rank][V6_WARPS]
};

template <int DEG, bool DIRECT>
static __global__ void __launch_bounds__(V6_BLOCK, 5) k_pairs_v6(const __grid_constant__ PairArgs A, const int4 *__restrict__ slots,
                                                           const V6Extra X)
{
    constexpr int S = 3;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_rowA = reinterpret_cast<double *>(smem_raw);
    double *s_rowB = s_rowA + V6_ACAP * V6_ROW;
    float4 *s_fA = reinterpret_cast<float4 *>(s_rowB + V6_BCAP * V6_ROW);
    float4 *s_fB = s_fA + V6_ACAP;
    unsigned short *s_queue = reinterpret_cast<unsigned short *>(s_fB + V6_BCAP);
    __shared__ int s_boff[V6_SLOTS + 1], s_bglob[V6_SLOTS], s_code[V6_SLOTS], s_nA, s_nB, s_pass_end;
    __shared__ double s_red[4 * V6_WARPS];
    __shared__ __align__(8) unsigned long long s_mbar;
    __shared__ long long s_unext;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned short *q = s_queue + warp * V6_QCAP;
    const double L = A.L;
    {   // the always-true cut-off tests (energy.jl:270, ewalds.jl:362) must really be always true for this state
        const double reach = sqrt(A.rc_qq2) + 2.0 * (*A.max_dev);
        if (!(reach * reach < A.rc_qq2 + 100.0) && tid == 0 && blockIdx.x == 0) atomicExch(A.err_flag, 1u);
    }
    const double lj_eps = A.lj_eps_tab[0], lj_sig2 = A.lj_sig_tab[0] * A.lj_sig_tab[0];
    const float edge_f = (float)(L / (double)A.ncd);
    const float rc2f = A.gate_rc2f;
    const unsigned lt = (1u << lane) - 1u;
    if (tid == 0) mbar_init(&s_mbar, 1);
    unsigned phase = 0;

    double acc_lj = 0.0, acc_vir = 0.0, acc_q = 0.0;
    unsigned long long my_pairs = 0;

    // piece descriptor of the unit (warp 0 only): lanes 0..4 = the group's slots {b_lo, count, code}, lane 5 = home cell
    auto fetch_desc = [&](long long u) -> int4 {
        int4 d = make_int4(0, 0, 0, 0);
        if (warp == 0 && u < A.unit_end) {
            const int c = (int)(u / V3_GROUPS), g = (int)(u - (long long)c * V3_GROUPS);
            const int sl0 = c_v3_group_begin[g], nsl = c_v3_group_begin[g + 1] - sl0;
            if (lane < nsl) d = slots[c * 14 + sl0 + lane];
            else if (lane == 5) { const int a_lo = A.cell_start[c]; d = make_int4(a_lo, A.cell_start[c + 1] - a_lo, 0, 0); }
        }
        return d;
    };
    long long u = A.unit_begin + blockIdx.x;
    int4 desc = fetch_desc(u);
    const bool dyn = X.ticket != nullptr;
    // one ticket is always in flight (drawn a unit ahead), so the atomic's round trip is never waited for
    unsigned tk_pending = 0;
    if (dyn && tid == 0) tk_pending = atomicAdd(X.ticket, 1u);
    long long u_next = u + gridDim.x;

    for (; u < A.unit_end; u = u_next) {
        const int c = (int)(u / V3_GROUPS), g = (int)(u - (long long)c * V3_GROUPS);
        const int sl0 = c_v3_group_begin[g], nsl = c_v3_group_begin[g + 1] - sl0;
      // a group whose neighbour cells hold more than V6_BCAP molecules together is evaluated in two passes
      for (int s_begin = 0, s_end = 0; s_begin < nsl; s_begin = s_end) {
        fence_proxy_async_smem();                          // our generic-proxy writes (fix-up) before the engine overwrites
        __syncthreads();                                   // everyone is done with the previous tiles
        if (warp == 0) {
            const int cnt = desc.y;
            int incl = (lane >= s_begin && lane < V6_SLOTS) ? cnt : 0;   // inclusive prefix of the B counts from s_begin on
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            // slots [s_begin, se) fit the B tile (every count is <= V6_ACAP, so at least four do)
            const int se = min(nsl, __popc(__ballot_sync(0xffffffffu, lane < V6_SLOTS && incl <= V6_BCAP)));
            const int nBt = __shfl_sync(0xffffffffu, incl, se - 1);
            const int nAt = __shfl_sync(0xffffffffu, cnt, 5);
            const bool bad = __any_sync(0xffffffffu, lane < 6 && cnt > V6_ACAP);
            const bool mine = (lane >= s_begin && lane < se) || (lane == 5 && s_begin == 0);
            if (lane < V6_SLOTS) { s_boff[lane] = incl - cnt; s_bglob[lane] = desc.x; s_code[lane] = desc.z; }
            __syncwarp();
            if (lane == 0) {
                s_boff[se] = bad ? 0 : nBt;                // one past the pass's last slot
                s_nA = bad ? 0 : nAt;
                s_nB = bad ? 0 : nBt;
                s_pass_end = bad ? nsl : se;
                if (bad) atomicExch(A.err_flag, 1u);
                mbar_arrive_expect_tx(&s_mbar, bad ? 0u : (unsigned)(nBt + (s_begin == 0 ? nAt : 0)) * (V6_ROW * 8 + 16));
            }
            if (!bad && mine && cnt > 0) {
                const int off = incl - cnt;
                double *drow = (lane == 5) ? s_rowA : s_rowB + off * V6_ROW;
                float4 *dgf = (lane == 5) ? s_fA : s_fB + off;
                bulk_g2s(drow, X.rows + (size_t)desc.x * V6_ROW, (unsigned)cnt * V6_ROW * 8, &s_mbar);
                bulk_g2s(dgf, X.gf + desc.x, (unsigned)cnt * 16, &s_mbar);
            }
            if (bad || se == nsl) {                        // next unit's descriptors travel while this one is evaluated
                long long un = u + gridDim.x;
                if (dyn) {
                    const unsigned tk = __shfl_sync(0xffffffffu, tk_pending, 0);
                    un = A.unit_begin + (long long)gridDim.x + tk;
                    if (lane == 0 && un < A.unit_end) tk_pending = atomicAdd(X.ticket, 1u);
                }
                if (lane == 0) s_unext = un;
                desc = fetch_desc(un);
            }
        }
        mbar_wait(&s_mbar, phase);
        phase ^= 1u;
        s_end = s_pass_end;
        const int nA = s_nA, nB = s_nB;
        // ---- fix-up: B gate coordinates get the slot's cell offset; rows of wrapped slots get ±L
        for (int t = tid; t < ((nB + 63) & ~63); t += V6_BLOCK) {
            if (t < nB) {
                int sl = s_begin;
#pragma unroll
                for (int k = 1; k < V6_SLOTS; ++k) sl += (k > s_begin && k < s_end && t >= s_boff[k]) ? 1 : 0;
                float4 f = s_fB[t];
                f.x += (float)c_half_shell[sl0 + sl][0] * edge_f;
                f.y += (float)c_half_shell[sl0 + sl][1] * edge_f;
                f.z += (float)c_half_shell[sl0 + sl][2] * edge_f;
                s_fB[t] = f;
            } else {
                s_fB[t] = make_float4(1e18f, 1e18f, 1e18f, 0.f);   // sentinels: the gate walks B in steps of 64
            }
        }
        for (int sl = s_begin; sl < s_end && nB > 0; ++sl) {   // (nB == 0: the unit was declined — offsets are not meaningful)
            const int code = s_code[sl];
            if (code == 0) continue;                       // uniform branch: only cells on the box faces
            const int cx = code & 3, cy = (code >> 2) & 3, cz = (code >> 4) & 3;
            const double sh[3] = {cx == 1 ? L : (cx == 2 ? -L : 0.0), cy == 1 ? L : (cy == 2 ? -L : 0.0),
                                  cz == 1 ? L : (cz == 2 ? -L : 0.0)};
            double *base = s_rowB + s_boff[sl] * V6_ROW;
            const int nval = (s_boff[sl + 1] - s_boff[sl]) * V6_ROW;
            for (int t = tid; t < nval; t += V6_BLOCK) {
                const int k = t % 3;
                base[t] = base[t] + (k == 0 ? sh[0] : (k == 1 ? sh[1] : sh[2]));
            }
        }
        __syncthreads();
        u_next = s_unext;                                  // written by warp 0 in the unit's last pass (earlier passes: stale, unused)
        const int self_n = (g == 0 && s_begin == 0) ? nA : 0;   // slot 0 of group 0 is the home cell itself: keep q > p

        int head = 0, tail = 0;                            // warp-private ring window [head, tail)
        auto consume = [&](int base, int count) {          // `count` queued molecule pairs, one per lane
            const bool have = lane < count;
            const unsigned e = have ? q[(base + lane) & (V6_QCAP - 1)] : 0u;
            const int p = e & 63u, qi = e >> 6;
            const double2 *ra = reinterpret_cast<const double2 *>(s_rowA + p * V6_ROW);
            const double2 *rb = reinterpret_cast<const double2 *>(s_rowB + qi * V6_ROW);
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
                    int slot_i = s_begin;
                    while (slot_i + 1 < s_end && qi >= s_boff[slot_i + 1]) ++slot_i;
                    const int qglob = s_bglob[slot_i] + (qi - s_boff[slot_i]);
                    if (atomicExch(&A.ovl[A.cell_start[c] + p], 1u) == 0u) atomicAdd(A.n_ovl, 1u);
                    if (atomicExch(&A.ovl[qglob], 1u) == 0u) atomicAdd(A.n_ovl, 1u);
                }
            }
        };

        // warp w owns rows p ≡ w + u (mod 4); per row, all columns: entries of one row are contiguous in the ring.
        // Gate everything first, then consume in full rounds (one inlined copy of the consume body).
        for (int p = (warp + (int)u) & (V6_WARPS - 1);; p += V6_WARPS) {
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
                    const int i0 = (tail + __popc(m0 & lt)) & (V6_QCAP - 1);
                    tail += __popc(m0);
                    const int i1 = (tail + __popc(m1 & lt)) & (V6_QCAP - 1);
                    tail += __popc(m1);
                    if (pass0) q[i0] = (unsigned short)e0;
                    if (pass1) q[i1] = (unsigned short)(e0 + (32u << 6));
                }
            }
            if (last || tail - head > V6_QCAP - V6_BCAP) {   // end of the unit, or the ring is nearly full
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
      if (dyn) {   // this unit's sums, per warp, at a place that depends on the unit only
          const double w0 = warp_sum(acc_lj), w1 = warp_sum(acc_vir), w2 = warp_sum(acc_q);
          if (lane == 0) X.unit_partial[(size_t)(u - A.unit_begin) * V6_WARPS + warp] = make_double4(w0, w1, w2, (double)my_pairs);
          acc_lj = 0.0; acc_vir = 0.0; acc_q = 0.0; my_pairs = 0;
      }
    }
    if (dyn) return;
    __syncthreads();
    double accp[4] = {acc_lj, acc_vir, acc_q, (double)my_pairs};
    block_sum<4, V6_BLOCK>(accp, s_red);
    if (tid == 0) A.partial[blockIdx.x] = make_double4(accp[0], accp[1], accp[2], accp[3]);
}

// stage 1 of the fold of the per-(unit, warp) sums: CTA b adds its contiguous share in a fixed order
static __global__ void __launch_bounds__(256) k_unit_fold(const double4 *__restrict__ unit_partial, long long n, double4 *out)
{
    __shared__ double s_red[4 * 8];
    const long long per = (n + gridDim.x - 1) / gridDim.x;
    const long long lo = per * blockIdx.x, hi = (lo + per < n) ? lo + per : n;
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    for (long long i = lo + threadIdx.x; i < hi; i += 256) { const double4 p = unit_partial[i]; v[0] += p.x; v[1] += p.y; v[2] += p.z; v[3] += p.w; }
    block_sum<4, 256>(v, s_red);
    if (threadIdx.x == 0) out[blockIdx.x] = make_double4(v[0], v[1], v[2], v[3]);
}
