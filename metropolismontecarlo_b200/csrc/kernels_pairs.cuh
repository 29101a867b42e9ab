// kernels_pairs.cuh — full-system pair energy for potential() / volume moves
// (Ewald/energy.jl:946-1032 and :864-943; Ewald/volumeChange.jl:91-111), SURVEY.md §8a a9/a10/a12.
//
// The reference evaluates Σ_i LJ_poly_ΔU(i)/2 and Σ_i EwaldReal(i)/2: every molecule pair twice,
// with an O(N) COM scan per molecule.  Here every unordered molecule pair is visited once:
//   * cell mode  (box ≥ 3 r_cut): molecules are binned on their COM into cells of edge ≥ r_cut
//     and stored cell-sorted; a work unit is (home cell, one of 14 half-shell neighbour slots);
//   * tile mode  (small boxes):   a work unit is a (64 x 64) tile pair I ≤ J of the molecule list.
// Units are dealt statically to persistent CTAs (no atomics → deterministic sums).  Inside a
// unit each warp (a) runs the COM gate |COM_ij|² < r_cut² (strict, energy.jl:250, ewalds.jl:337)
// for its share of the molecule pairs and compacts the survivors into its own shared-memory
// queue with ballot/popc, then (b) spreads the queue's n_a·n_b site pairs over its 32 lanes, so
// the FP64 pipe sees full warps of erfc work, and (c) runs the LJ-active site pairs (O–O only
// for water) as a second, equally dense pass.
//
// Minimum image: the reference wraps every site pair separately (vector1D, boundaries.jl:8-14).
// In cell mode the image of a surviving pair is fixed by the cell offset (shift ∈ {-L,0,+L} per
// axis) whenever r_cut + 2·max|site-COM| < L/2; d = (x_b - x_a) ± L is then formed in the same
// order as the reference, so the bits agree.  If that bound does not hold the kernel falls back
// to the literal per-pair vector1D (uniform branch on a device-side flag).
#pragma once
#include "mmc_common.cuh"
#include "erf_poly.h"

#define PAIR_BLOCK 256
#define PAIR_WARPS (PAIR_BLOCK / 32)
#define PAIR_TILE 64
#define PAIR_QCAP (PAIR_TILE * PAIR_TILE / PAIR_WARPS)

struct PairArgs {
    const double4 *com;        // cell-sorted (cell mode) / original order (tile mode)
    const double4 *site;       // S sites per molecule, site[m*S + a] = {x,y,z,q}
    const int *cell_start;     // [ncell+1] (cell mode)
    int ncd, S, n_mol, mode;   // mode 0: cells, 1: tiles
    int n_tiles;
    long long unit_begin, unit_end;
    double L, rc_lj2, rc_qq2, kappa;
    int want_lj, want_qq;
    int nlj;
    const LJActive *lj;
    double4 *partial;          // [gridDim.x]
    unsigned int *ovl;         // [n_mol] overlap flags (index space of `com`)
    unsigned int *n_ovl;
    const double *max_dev;     // max |site - COM| component, written by k_gather
    unsigned int *err_flag;    // set when a cell does not fit the fast kernel's tile
    const int4 *units;         // fast kernel: {a_lo, b_lo, nA | nB<<16, self | codes<<1} per unit (k_units_build)
    long long rclj_bits, rcqq_bits, cutlj_bits, cutqq_bits;   // bit patterns of r_cut² and r_cut²+100
    double lj_eps_tab[16], lj_sig_tab[16];   // v3: LJ table by (site a, site b) of the uniform molecule, 0 = inactive
    double qq_tab[9];          // v4: q_a q_b of the uniform 3-site molecule (launch constants)
    unsigned qq_negmask;       // v4: bit j set when qq_tab[j] < 0 (the overlap rule only fires there)
    float gate_rc2f;           // v4: conservative FP32 COM-gate threshold (>= r_cut² + worst-case FP32 error)
    ErfPoly ep;                // smooth part of erfc(κr)/r as one polynomial (deg 0: use erfc())
    double pc[MMC_ERF_MAXDEG + 1];   // v5: −κ·(coefficients), in r² (DIRECT, κ^2k folded) or in s = pk2s·r² − 1
    double pk2s;               // v5: σ κ²
    double *per_mol;           // k_pairs<ST, true>: [n_mol x 3] per-molecule rows {Σlj_pot, Σlj_vir, Σcoul} (index space of `com`)
    // mixed topologies (k_pairs<0, *> only): every molecule padded to S site slots, stype[m*S + a] = 0-based LJ type of the
    // slot or −1 for padding; `lj` then lists the active TYPE pairs (a, b = types) instead of site-slot pairs
    const signed char *stype;
};

// ---- one Coulomb site pair: q_a q_b erfc(κ r)/r with the overlap rule (ewalds.jl:359-367).
// Comparisons of non-negative doubles are done on their bit patterns (integer pipe) so the FP64
// pipe only sees arithmetic.  POLY: erfc(κr)/r = 1/r − κ·E(κ²r²), E from the fitted polynomial.
template <int DEG>   // DEG > 0: padded degree of the erf polynomial (compile time → straight DFMA run); 0: erfc()
__device__ __forceinline__ bool coul_pair(const PairArgs &A, double r2, double qq, long long cut_bits, double &acc)
{
    const long long r2b = __double_as_longlong(r2);
    if (r2b < 0x3FE0000000000000LL && __double2hiint(qq) < 0 && qq != 0.0) return true;   // r² < 0.5 && qq < 0
    if (r2b < cut_bits) {
        if (DEG > 0) {
            const double rinv = rsqrt(r2);
            const double sv = fma(r2 * A.ep.kappa2, A.ep.scale, -1.0);
            double pv = A.ep.c[DEG];
#pragma unroll
            for (int k = DEG - 1; k >= 0; --k) pv = fma(pv, sv, A.ep.c[k]);   // coefficients: constant-bank operands
            acc = fma(qq, fma(-A.ep.kappa, pv, rinv), acc);
        } else {
            const double r = sqrt(r2);
            acc += qq * erfc(A.kappa * r) / r;
        }
    }
    return false;
}

static __constant__ int c_half_shell[14][3] = {
    {0, 0, 0},
    {1, 0, 0},
    {-1, 1, 0}, {0, 1, 0}, {1, 1, 0},
    {-1, -1, 1}, {0, -1, 1}, {1, -1, 1},
    {-1, 0, 1}, {0, 0, 1}, {1, 0, 1},
    {-1, 1, 1}, {0, 1, 1}, {1, 1, 1}};

// PM: every evaluated molecule pair is also credited to the rows of BOTH molecules (mmc_energy_all: LJ_poly_ΔU(i) and
// EwaldReal(i) for all i from one pass over the unique pairs; FP64 atomics, so the row sums are order-free to ~1e-13).
template <int ST, bool PM>   // ST = sites per molecule at compile time (0: runtime A.S)
static __global__ void __launch_bounds__(PAIR_BLOCK, 2) k_pairs(const __grid_constant__ PairArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = ST ? ST : A.S;
    const int SS = S * S;
    double4 *s_comA = reinterpret_cast<double4 *>(smem_raw);
    double4 *s_comB = s_comA + PAIR_TILE;
    double4 *s_siteA = s_comB + PAIR_TILE;
    double4 *s_siteB = s_siteA + PAIR_TILE * S;
    unsigned int *s_queue = reinterpret_cast<unsigned int *>(s_siteB + PAIR_TILE * S);
    const bool mix = (ST == 0) && A.stype != nullptr;
    signed char *s_typeA = reinterpret_cast<signed char *>(s_queue + PAIR_WARPS * PAIR_QCAP);     // mix only (the host sizes smem)
    signed char *s_typeB = s_typeA + PAIR_TILE * S;
    __shared__ LJActive s_lj[64];
    __shared__ double s_red[4 * PAIR_WARPS];
    __shared__ double2 s_tab[MMC_MAX_TYPES * MMC_MAX_TYPES];   // mix: {eps, sig} by type pair, eps = 0 for the inactive ones

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned int *q = s_queue + warp * PAIR_QCAP;
    const int nlj = min(A.nlj, 64);
    if (tid < nlj) s_lj[tid] = A.lj[tid];
    if (mix) {
        if (tid < MMC_MAX_TYPES * MMC_MAX_TYPES) s_tab[tid] = make_double2(0.0, 0.0);
        __syncthreads();
        if (tid < nlj) s_tab[A.lj[tid].a * MMC_MAX_TYPES + A.lj[tid].b] = make_double2(A.lj[tid].eps, A.lj[tid].sig);
    }
    const double L = A.L;
    const bool cells = (A.mode == 0);
    double rcmax = sqrt(fmax(A.rc_lj2, A.rc_qq2));
    const bool fast = cells && (rcmax + 2.0 * (*A.max_dev) < 0.5 * L);
    const long long cut_bits = A.cutqq_bits;

    // static, contiguous share of this rank's units
    const long long n_units = A.unit_end - A.unit_begin;
    const long long per = (n_units + gridDim.x - 1) / gridDim.x;
    const long long u0 = A.unit_begin + per * blockIdx.x;
    const long long u1 = (u0 + per < A.unit_end) ? u0 + per : A.unit_end;

    double acc[3] = {0.0, 0.0, 0.0};   // lj_pot, lj_vir, coul
    unsigned long long my_pairs = 0;   // molecule pairs that passed the gate (counted by lane 0 of each warp)

    for (long long u = u0; u < u1; ++u) {
        int a_lo, a_hi, b_lo, b_hi;
        double shx = 0.0, shy = 0.0, shz = 0.0;
        bool self;
        if (cells) {
            const int c = (int)(u / 14), slot = (int)(u - 14 * (long long)c);
            const int n = A.ncd;
            const int cx = c % n, cy = (c / n) % n, cz = c / (n * n);
            int nx = cx + c_half_shell[slot][0], ny = cy + c_half_shell[slot][1],
                nz = cz + c_half_shell[slot][2];
            if (nx >= n) { nx -= n; shx = L; } else if (nx < 0) { nx += n; shx = -L; }
            if (ny >= n) { ny -= n; shy = L; } else if (ny < 0) { ny += n; shy = -L; }
            if (nz >= n) { nz -= n; shz = L; } else if (nz < 0) { nz += n; shz = -L; }
            const int cb = nx + n * (ny + n * nz);
            a_lo = A.cell_start[c]; a_hi = A.cell_start[c + 1];
            b_lo = A.cell_start[cb]; b_hi = A.cell_start[cb + 1];
            self = (slot == 0);
        } else {
            // unit → (I, J) with I ≤ J, row-major over the upper triangle
            long long r = u; int I = 0, rowlen = A.n_tiles;
            while (r >= rowlen) { r -= rowlen; ++I; --rowlen; }
            const int J = I + (int)r;
            a_lo = I * PAIR_TILE; a_hi = min(A.n_mol, a_lo + PAIR_TILE);
            b_lo = J * PAIR_TILE; b_hi = min(A.n_mol, b_lo + PAIR_TILE);
            self = (I == J);
        }
        for (int a0 = a_lo; a0 < a_hi; a0 += PAIR_TILE) {
            const int nA = min(PAIR_TILE, a_hi - a0);
            for (int b0 = b_lo; b0 < b_hi; b0 += PAIR_TILE) {
                if (self && b0 + PAIR_TILE <= a0) continue;   // strictly lower sub-tile
                const int nB = min(PAIR_TILE, b_hi - b0);
                __syncthreads();
                for (int t = tid; t < nA; t += PAIR_BLOCK) s_comA[t] = A.com[a0 + t];
                for (int t = tid; t < nB; t += PAIR_BLOCK) s_comB[t] = A.com[b0 + t];
                for (int t = tid; t < nA * S; t += PAIR_BLOCK) s_siteA[t] = A.site[(size_t)a0 * S + t];
                for (int t = tid; t < nB * S; t += PAIR_BLOCK) s_siteB[t] = A.site[(size_t)b0 * S + t];
                if (mix) {
                    for (int t = tid; t < nA * S; t += PAIR_BLOCK) s_typeA[t] = A.stype[(size_t)a0 * S + t];
                    for (int t = tid; t < nB * S; t += PAIR_BLOCK) s_typeB[t] = A.stype[(size_t)b0 * S + t];
                }
                __syncthreads();
                // ---- (a) COM gate, warp-private ordered compaction
                int npairs = 0;
                const int ntests = nA * nB;
                for (int t0 = warp * 32; t0 < ntests; t0 += PAIR_BLOCK) {
                    const int t = t0 + lane;
                    int fl = 0, p = 0, qq_ = 0;
                    if (t < ntests) {
                        p = t / nB; qq_ = t - p * nB;
                        if (!self || (b0 + qq_ > a0 + p)) {
                            const double4 ca = s_comA[p], cb = s_comB[qq_];
                            double dx, dy, dz;
                            if (cells) {
                                dx = (cb.x - ca.x) + shx; dy = (cb.y - ca.y) + shy; dz = (cb.z - ca.z) + shz;
                            } else {
                                dx = min_image(ca.x, cb.x, L); dy = min_image(ca.y, cb.y, L);
                                dz = min_image(ca.z, cb.z, L);
                            }
                            const double r2 = dx * dx + dy * dy + dz * dz;
                            if (A.want_lj && r2 < A.rc_lj2) fl |= 1;
                            if (A.want_qq && r2 < A.rc_qq2) fl |= 2;
                        }
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, fl != 0);
                    if (fl) q[npairs + __popc(m & ((1u << lane) - 1u))] =
                        (unsigned)p | ((unsigned)qq_ << 8) | ((unsigned)fl << 16);
                    npairs += __popc(m);
                }
                if (lane == 0) my_pairs += npairs;
                __syncwarp();
                // ---- (b) Coulomb: all S x S site pairs of the queued molecule pairs
                if (A.want_qq) {
                    const int items = npairs * SS;
                    for (int w = lane; w < items; w += 32) {
                        const int pr = w / SS, ab = w - pr * SS;
                        const unsigned e = q[pr];
                        if (!(e & (2u << 16))) continue;
                        const int p = e & 255u, qi = (e >> 8) & 255u;
                        const int a = ab / S, b = ab - a * S;
                        if (mix && (s_typeA[p * S + a] < 0 || s_typeB[qi * S + b] < 0)) continue;      // padding slot
                        const double4 sa = s_siteA[p * S + a], sb = s_siteB[qi * S + b];
                        double dx, dy, dz;
                        if (fast) {
                            dx = (sb.x - sa.x) + shx; dy = (sb.y - sa.y) + shy; dz = (sb.z - sa.z) + shz;
                        } else {
                            dx = min_image(sa.x, sb.x, L); dy = min_image(sa.y, sb.y, L);
                            dz = min_image(sa.z, sb.z, L);
                        }
                        const double r2 = dx * dx + dy * dy + dz * dz;
                        double c1 = 0.0;
                        if (coul_pair<0>(A, r2, sa.w * sb.w, cut_bits, PM ? c1 : acc[2])) {           // ewalds.jl:359
                            if (atomicExch(&A.ovl[a0 + p], 1u) == 0u) atomicAdd(A.n_ovl, 1u);
                            if (atomicExch(&A.ovl[b0 + qi], 1u) == 0u) atomicAdd(A.n_ovl, 1u);
                        }
                        if (PM && c1 != 0.0) {
                            acc[2] += c1;
                            atomicAdd(&A.per_mol[3 * (size_t)(a0 + p) + 2], c1);
                            atomicAdd(&A.per_mol[3 * (size_t)(b0 + qi) + 2], c1);
                        }
                    }
                }
                // ---- (c) LJ: only the site-type combinations with ε_ij > 0.001 (energy.jl:270)
                if (A.want_lj) {
                    const int per = mix ? SS : nlj;       // mixed: every slot pair, resolved through the type table
                    const int items = npairs * per;
                    for (int w = lane; w < items; w += 32) {
                        const int pr = w / per, li = w - pr * per;
                        const unsigned e = q[pr];
                        if (!(e & (1u << 16))) continue;
                        const int p = e & 255u, qi = (e >> 8) & 255u;
                        LJActive lj;
                        if (mix) {
                            lj.a = li / S; lj.b = li - lj.a * S;
                            const int ta = s_typeA[p * S + lj.a], tb = s_typeB[qi * S + lj.b];
                            if (ta < 0 || tb < 0) continue;
                            const double2 es = s_tab[ta * MMC_MAX_TYPES + tb];
                            if (es.x == 0.0) continue;                                  // ε_ij <= 0.001 (energy.jl:270)
                            lj.eps = es.x; lj.sig = es.y;
                        } else
                            lj = s_lj[li];
                        const double4 sa = s_siteA[p * S + lj.a], sb = s_siteB[qi * S + lj.b];
                        const double4 ca = s_comA[p], cb = s_comB[qi];
                        double dx, dy, dz, rx, ry, rz;
                        if (cells) {
                            rx = (cb.x - ca.x) + shx; ry = (cb.y - ca.y) + shy; rz = (cb.z - ca.z) + shz;
                        } else {
                            rx = min_image(ca.x, cb.x, L); ry = min_image(ca.y, cb.y, L);
                            rz = min_image(ca.z, cb.z, L);
                        }
                        if (fast) {
                            dx = (sb.x - sa.x) + shx; dy = (sb.y - sa.y) + shy; dz = (sb.z - sa.z) + shz;
                        } else {
                            dx = min_image(sa.x, sb.x, L); dy = min_image(sa.y, sb.y, L);
                            dz = min_image(sa.z, sb.z, L);
                        }
                        const double r2 = dx * dx + dy * dy + dz * dz;
                        if (r2 < (A.rc_lj2 + 100)) {
                            if (PM) {
                                double l0 = 0.0, l1 = 0.0;
                                lj_pair(lj.eps, lj.sig, r2, dx, dy, dz, rx, ry, rz, l0, l1);
                                acc[0] += l0; acc[1] += l1;
                                atomicAdd(&A.per_mol[3 * (size_t)(a0 + p)], l0); atomicAdd(&A.per_mol[3 * (size_t)(a0 + p) + 1], l1);
                                atomicAdd(&A.per_mol[3 * (size_t)(b0 + qi)], l0); atomicAdd(&A.per_mol[3 * (size_t)(b0 + qi) + 1], l1);
                            } else
                                lj_pair(lj.eps, lj.sig, r2, dx, dy, dz, rx, ry, rz, acc[0], acc[1]);
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    double accp[4] = {acc[0], acc[1], acc[2], (double)my_pairs};
    block_sum<4, PAIR_BLOCK>(accp, s_red);
    if (tid == 0) A.partial[blockIdx.x] = make_double4(accp[0], accp[1], accp[2], accp[3]);
}

// ------------------------------------------------------------------------------------------
// k_pairs_fast — the production variant for uniform S-site molecules whose cells fit one tile.
// Same unit → gate → queue → site-pair scheme as k_pairs, plus:
//   * units dealt round-robin (u = blockIdx.x + k·gridDim.x): cells of unequal population
//     (27…64 molecules on the lattice start) average out across CTAs, still deterministic;
//   * the two tiles of the NEXT unit stream into the other shared-memory buffer with cp.async
//     (LDGSTS, 16 B per thread) while the current unit is evaluated: one barrier per unit and
//     no exposed L2 latency;
//   * 16-bit queue entries, reciprocal-multiply index math, bit-pattern compares.
// A cell with more than TILE molecules raises err_flag; the host then re-runs k_pairs.
// ------------------------------------------------------------------------------------------
struct UnitDesc { int a_lo, b_lo, nA, nB, self; double shx, shy, shz; };

__device__ __forceinline__ void unpack_unit(const int4 d, double L, UnitDesc &U)
{
    U.a_lo = d.x; U.b_lo = d.y; U.nA = d.z & 0xffff; U.nB = (d.z >> 16) & 0xffff; U.self = d.w & 1;
    const int cx = (d.w >> 1) & 3, cy = (d.w >> 3) & 3, cz = (d.w >> 5) & 3;
    U.shx = cx == 1 ? L : (cx == 2 ? -L : 0.0);
    U.shy = cy == 1 ? L : (cy == 2 ? -L : 0.0);
    U.shz = cz == 1 ? L : (cz == 2 ? -L : 0.0);
}

// one thread per unit: resolves (cell, slot) / (I, J) into tile ranges and wrap codes once per binning
static __global__ void k_units_build(PairArgs A, int4 *units, long long n_units)
{
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_units) return;
    int a_lo, nA, b_lo, nB, self, code = 0;
    if (A.mode == 0) {
        const int c = (int)(u / 14), slot = (int)(u - 14 * (long long)c);
        const int n = A.ncd;
        const int cx = c % n, cy = (c / n) % n, cz = c / (n * n);
        int nx = cx + c_half_shell[slot][0], ny = cy + c_half_shell[slot][1], nz = cz + c_half_shell[slot][2];
        if (nx >= n) { nx -= n; code |= 1 << 0; } else if (nx < 0) { nx += n; code |= 2 << 0; }
        if (ny >= n) { ny -= n; code |= 1 << 2; } else if (ny < 0) { ny += n; code |= 2 << 2; }
        if (nz >= n) { nz -= n; code |= 1 << 4; } else if (nz < 0) { nz += n; code |= 2 << 4; }
        const int cb = nx + n * (ny + n * nz);
        a_lo = A.cell_start[c]; nA = A.cell_start[c + 1] - a_lo;
        b_lo = A.cell_start[cb]; nB = A.cell_start[cb + 1] - b_lo;
        self = (slot == 0);
    } else {
        long long r = u; int I = 0, rowlen = A.n_tiles;
        while (r >= rowlen) { r -= rowlen; ++I; --rowlen; }
        const int J = I + (int)r;
        a_lo = I * PAIR_TILE; nA = min(A.n_mol, a_lo + PAIR_TILE) - a_lo;
        b_lo = J * PAIR_TILE; nB = min(A.n_mol, b_lo + PAIR_TILE) - b_lo;
        self = (I == J);
    }
    units[u] = make_int4(a_lo, b_lo, nA | (nB << 16), self | (code << 1));
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}

// gate + queue + Coulomb + LJ for one unit; WRAP: the unit crosses a periodic boundary
template <int S, int TILE, int DEG, bool WRAP>
__device__ __forceinline__ void pair_unit_body(const PairArgs &A, const UnitDesc &U, const double4 *s_comA,
                                               const double4 *s_comB, const double4 *s_siteA, const double4 *s_siteB,
                                               unsigned short *q, const LJActive *s_lj, int nlj, bool cells, bool fast,
                                               int lane, int warp, double (&acc)[3], unsigned long long &my_pairs)
{
    constexpr int SS = S * S;
    const double L = A.L;
    const int nA = U.nA, nB = U.nB;
    const double shx = U.shx, shy = U.shy, shz = U.shz;
    // ---- (a) COM gate, warp-private ordered compaction
    int npairs = 0;
    const int ntests = nA * nB;
    const unsigned inv_nB = nB > 0 ? (0xFFFFFFFFu / (unsigned)nB) + 1u : 0u;   // t/nB = umulhi(t, inv) for t < 2^16
    for (int t0 = warp * 32; t0 < ntests; t0 += PAIR_BLOCK) {
        const int t = t0 + lane;
        int fl = 0, p = 0, qq_ = 0;
        if (t < ntests) {
            p = (int)__umulhi((unsigned)t, inv_nB); qq_ = t - p * nB;
            if (!U.self || (U.b_lo + qq_ > U.a_lo + p)) {
                const double4 ca = s_comA[p], cb = s_comB[qq_];
                double dx, dy, dz;
                if (cells) {
                    dx = cb.x - ca.x; dy = cb.y - ca.y; dz = cb.z - ca.z;
                    if (WRAP) { dx += shx; dy += shy; dz += shz; }
                } else {
                    dx = min_image(ca.x, cb.x, L); dy = min_image(ca.y, cb.y, L); dz = min_image(ca.z, cb.z, L);
                }
                const long long r2b = __double_as_longlong(dx * dx + dy * dy + dz * dz);
                if (A.want_lj && r2b < A.rclj_bits) fl |= 1;
                if (A.want_qq && r2b < A.rcqq_bits) fl |= 2;
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, fl != 0);
        if (fl) q[npairs + __popc(m & ((1u << lane) - 1u))] = (unsigned short)(p | (qq_ << 7) | (fl << 14));
        npairs += __popc(m);
    }
    if (lane == 0) my_pairs += npairs;
    __syncwarp();
    // ---- (b) Coulomb: S x S site pairs of the queued molecule pairs, one per lane.
    // item w = pr*SS + ab advances by 32 per iteration: (pr, ab) are stepped incrementally.
    if (A.want_qq) {
        const int items = npairs * SS;
        int pr = lane / SS, ab = lane - pr * SS;
        constexpr int DP = 32 / SS, DA = 32 - DP * SS;
        for (int w = lane; w < items; w += 32) {
            const unsigned e = q[pr];
            const int p = e & 127u, qi = (e >> 7) & 127u;
            const int a = (S == 3) ? ((ab * 11) >> 5) : ab / S, b = ab - a * S;
            const bool on = (e & (2u << 14)) != 0;
            pr += DP; ab += DA;
            if (ab >= SS) { ab -= SS; ++pr; }
            if (!on) continue;
            const double4 sa = s_siteA[p * S + a], sb = s_siteB[qi * S + b];
            double dx, dy, dz;
            if (fast) {
                dx = sb.x - sa.x; dy = sb.y - sa.y; dz = sb.z - sa.z;
                if (WRAP) { dx += shx; dy += shy; dz += shz; }
            } else {
                dx = min_image(sa.x, sb.x, L); dy = min_image(sa.y, sb.y, L); dz = min_image(sa.z, sb.z, L);
            }
            const double r2 = dx * dx + dy * dy + dz * dz;
            if (coul_pair<DEG>(A, r2, sa.w * sb.w, A.cutqq_bits, acc[2])) {                  // ewalds.jl:359
                if (atomicExch(&A.ovl[U.a_lo + p], 1u) == 0u) atomicAdd(A.n_ovl, 1u);
                if (atomicExch(&A.ovl[U.b_lo + qi], 1u) == 0u) atomicAdd(A.n_ovl, 1u);
            }
        }
    }
    // ---- (c) LJ: only the site-type combinations with ε_ij > 0.001 (energy.jl:270)
    if (A.want_lj) {
        const int items = npairs * nlj;
        for (int w = lane; w < items; w += 32) {
            const int pr = (nlj == 1) ? w : w / nlj, li = (nlj == 1) ? 0 : w - pr * nlj;
            const unsigned e = q[pr];
            if (!(e & (1u << 14))) continue;
            const int p = e & 127u, qi = (e >> 7) & 127u;
            const LJActive lj = s_lj[li];
            const double4 sa = s_siteA[p * S + lj.a], sb = s_siteB[qi * S + lj.b];
            const double4 ca = s_comA[p], cb = s_comB[qi];
            double dx, dy, dz, rx, ry, rz;
            if (cells) {
                rx = cb.x - ca.x; ry = cb.y - ca.y; rz = cb.z - ca.z;
                if (WRAP) { rx += shx; ry += shy; rz += shz; }
            } else {
                rx = min_image(ca.x, cb.x, L); ry = min_image(ca.y, cb.y, L); rz = min_image(ca.z, cb.z, L);
            }
            if (fast) {
                dx = sb.x - sa.x; dy = sb.y - sa.y; dz = sb.z - sa.z;
                if (WRAP) { dx += shx; dy += shy; dz += shz; }
            } else {
                dx = min_image(sa.x, sb.x, L); dy = min_image(sa.y, sb.y, L); dz = min_image(sa.z, sb.z, L);
            }
            const double r2 = dx * dx + dy * dy + dz * dz;
            if (__double_as_longlong(r2) < A.cutlj_bits)
                lj_pair(lj.eps, lj.sig, r2, dx, dy, dz, rx, ry, rz, acc[0], acc[1]);
        }
    }
}

template <int ST, int TILE, int DEG>
static __global__ void __launch_bounds__(PAIR_BLOCK, (TILE <= 64 ? 4 : 2)) k_pairs_fast(const __grid_constant__ PairArgs A)
{
    constexpr int S = ST;
    constexpr int QCAP = TILE * TILE / PAIR_WARPS;
    constexpr int TILE_D4 = 2 * TILE + 2 * TILE * S;        // double4 per buffer: comA comB siteA siteB
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double4 *s_buf = reinterpret_cast<double4 *>(smem_raw);
    unsigned short *s_queue = reinterpret_cast<unsigned short *>(s_buf + 2 * TILE_D4);
    __shared__ int4 s_unit[2];
    __shared__ LJActive s_lj[64];
    __shared__ double s_red[4 * PAIR_WARPS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned short *q = s_queue + warp * QCAP;
    const int nlj = min(A.nlj, 64);
    if (tid < nlj) s_lj[tid] = A.lj[tid];
    const bool cells = (A.mode == 0);
    const double rcmax = sqrt(fmax(A.rc_lj2, A.rc_qq2));
    const bool fast = cells && (rcmax + 2.0 * (*A.max_dev) < 0.5 * A.L);

    auto prefetch = [&](int b, long long u) {
        int4 d = A.units[u];
        int nA = d.z & 0xffff, nB = (d.z >> 16) & 0xffff;
        if (nA > TILE || nB > TILE) { if (tid == 0) atomicExch(A.err_flag, 1u); nA = 0; nB = 0; d.z = 0; }
        if (tid == 0) s_unit[b] = d;
        double4 *base = s_buf + b * TILE_D4;
        // 16-byte chunks: comA | comB | siteA | siteB
        const char *gA = reinterpret_cast<const char *>(A.com + d.x), *gB = reinterpret_cast<const char *>(A.com + d.y);
        const char *gSA = reinterpret_cast<const char *>(A.site + (size_t)d.x * S);
        const char *gSB = reinterpret_cast<const char *>(A.site + (size_t)d.y * S);
        char *dA = reinterpret_cast<char *>(base), *dB = reinterpret_cast<char *>(base + TILE);
        char *dSA = reinterpret_cast<char *>(base + 2 * TILE), *dSB = reinterpret_cast<char *>(base + 2 * TILE + TILE * S);
        for (int t = tid; t < 2 * nA; t += PAIR_BLOCK) cp_async16(dA + 16 * t, gA + 16 * t);
        for (int t = tid; t < 2 * nB; t += PAIR_BLOCK) cp_async16(dB + 16 * t, gB + 16 * t);
        for (int t = tid; t < 2 * nA * S; t += PAIR_BLOCK) cp_async16(dSA + 16 * t, gSA + 16 * t);
        for (int t = tid; t < 2 * nB * S; t += PAIR_BLOCK) cp_async16(dSB + 16 * t, gSB + 16 * t);
        asm volatile("cp.async.commit_group;\n" ::);
    };

    double acc[3] = {0.0, 0.0, 0.0};
    unsigned long long my_pairs = 0;
    long long u = A.unit_begin + blockIdx.x;
    int it = 0;
    if (u < A.unit_end) prefetch(0, u);
    for (; u < A.unit_end; u += gridDim.x, it ^= 1) {
        asm volatile("cp.async.wait_all;\n" ::);
        __syncthreads();                       // unit u has landed; everyone is done with unit u - G
        if (u + gridDim.x < A.unit_end) prefetch(it ^ 1, u + gridDim.x);
        UnitDesc U;
        unpack_unit(s_unit[it], A.L, U);
        const double4 *s_comA = s_buf + it * TILE_D4, *s_comB = s_comA + TILE;
        const double4 *s_siteA = s_comA + 2 * TILE, *s_siteB = s_siteA + TILE * S;
        if ((s_unit[it].w >> 1) != 0)
            pair_unit_body<S, TILE, DEG, true>(A, U, s_comA, s_comB, s_siteA, s_siteB, q, s_lj, nlj, cells, fast, lane, warp, acc, my_pairs);
        else
            pair_unit_body<S, TILE, DEG, false>(A, U, s_comA, s_comB, s_siteA, s_siteB, q, s_lj, nlj, cells, fast, lane, warp, acc, my_pairs);
    }
    __syncthreads();
    double accp[4] = {acc[0], acc[1], acc[2], (double)my_pairs};
    block_sum<4, PAIR_BLOCK>(accp, s_red);
    if (tid == 0) A.partial[blockIdx.x] = make_double4(accp[0], accp[1], accp[2], accp[3]);
}

// fold the per-CTA partials in CTA order into the head of the partial-sum vector:
// out[0] = Σ lj_pot, out[1] = Σ lj_vir, out[2] = Σ coul (un-scaled), out[3] = #overlapped molecules
static __global__ void k_pair_reduce(const double4 *partial, int nb, const unsigned int *n_ovl, const int *max_count,
                              const unsigned int *err_flag, double *out)
{
    // fixed order: thread t adds partials t, t+256, ... ; then the 256 thread sums are folded by block_sum (warp
    // shuffles, then warp order) — the same tree every run, so the result is reproducible
    __shared__ double s_red[4 * 8];
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < nb; i += 256) { const double4 p = partial[i]; v[0] += p.x; v[1] += p.y; v[2] += p.z; v[3] += p.w; }
    block_sum<4, 256>(v, s_red);
    if (threadIdx.x != 0) return;
    out[0] = v[0]; out[1] = v[1]; out[2] = v[2]; out[3] = (double)(*n_ovl); out[5] = v[3];
    out[6] = (double)(*max_count); out[7] = (double)(*err_flag);
}

// ------------------------------------------------------------------ cell binning + gather
// per-molecule rows (evaluation order) → the caller's arrays in molecule order, scaled as LJ_poly_ΔU (energy.jl:289:
// 4·pot, 24·vir/3) and EwaldShort (ewalds.jl:905: pot·factor) return them; a flagged molecule reports (0, overlap)
// like EwaldReal's early return (ewalds.jl:359-360)
struct PerMolArgs {
    const double *rows; const int *perm; const unsigned int *ovl;
    int n_mol, want_qq; double factor;
    double *lj_pot, *lj_vir, *coul; int *overlap;
};

static __global__ void k_permol_scatter(PerMolArgs A)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= A.n_mol) return;
    const int m = A.perm ? A.perm[p] : p;
    A.lj_pot[m] = 4 * A.rows[3 * (size_t)p];
    A.lj_vir[m] = 24 * A.rows[3 * (size_t)p + 1] / 3;
    const bool ov = A.want_qq && A.ovl[p] != 0u;
    A.coul[m] = (A.want_qq && !ov) ? A.rows[3 * (size_t)p + 2] * A.factor : 0.0;
    A.overlap[m] = ov ? 1 : 0;
}

struct CellArgs {
    const double4 *com;
    int n_mol, ncd;
    double inv_cell;     // ncd / L
    int *cell_of;        // [n_mol]
    int *count;          // [ncell]  (zeroed)
    int *start;          // [ncell+1]
    int *fill;           // [ncell]  (zeroed)
    int *perm;           // [n_mol] sorted position -> molecule
    int *max_count;      // largest cell population
    int zl_lo, zl_cnt;   // sharded evaluation: only the z-layers [zl_lo, zl_lo + zl_cnt) (mod ncd) are needed by this rank
};

__device__ __forceinline__ int cell_coord(double x, double inv_cell, int n)
{
    int c = (int)floor(x * inv_cell);
    return c < 0 ? 0 : (c >= n ? n - 1 : c);
}

static __global__ void k_cell_count(CellArgs A)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= A.n_mol) return;
    const double4 c = A.com[m];
    const int cx = cell_coord(c.x, A.inv_cell, A.ncd), cy = cell_coord(c.y, A.inv_cell, A.ncd),
              cz = cell_coord(c.z, A.inv_cell, A.ncd);
    const int id = cx + A.ncd * (cy + A.ncd * cz);
    A.cell_of[m] = id;
    atomicAdd(&A.count[id], 1);
}

// exclusive scan of count[0..ncell) into start[0..ncell], single CTA
static __global__ void k_cell_scan(CellArgs A, int ncell)
{
    __shared__ int s_warp[32];
    const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int per = (ncell + nth - 1) / nth;
    const int lo = tid * per, hi = min(ncell, lo + per);
    int s = 0, mx = 0;
    for (int i = lo; i < hi; ++i) { const int c = A.count[i]; s += c; mx = max(mx, c); }
    // block-wide exclusive scan of the per-thread sums: shuffle scan inside each warp, then across the warp totals
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 31) s_warp[warp] = incl;
    if (lane == 0 && mx > 0) atomicMax(A.max_count, mx);
    __syncthreads();
    if (warp == 0) {
        const int nw = nth >> 5;
        int w = lane < nw ? s_warp[lane] : 0, wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += v; }
        s_warp[lane] = wi - w;                               // exclusive prefix of the warp totals
        if (lane == 31) A.start[ncell] = wi;
    }
    __syncthreads();
    int run = s_warp[warp] + incl - s;
    for (int i = lo; i < hi; ++i) { A.start[i] = run; run += A.count[i]; }
}

static __global__ void k_cell_fill(CellArgs A)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= A.n_mol) return;
    const int id = A.cell_of[m];
    const int pos = atomicAdd(&A.fill[id], 1);
    A.perm[A.start[id] + pos] = m;
}

// one warp per cell: order the cell's molecules by index so that the summation order (and so
// every bit of the result) is independent of the atomics' arrival order above
static __global__ void k_cell_sort(CellArgs A, int ncell)
{
    const int cell = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (cell >= ncell) return;
    {   // a rank of a sharded evaluation only orders the cells it will read
        const int z = cell / (A.ncd * A.ncd);
        int dz = z - A.zl_lo; if (dz < 0) dz += A.ncd;
        if (dz >= A.zl_cnt) return;
    }
    const int lo = A.start[cell], n = A.start[cell + 1] - lo;
    // rank sort through registers: n is a few dozen
    int mine[8], rank[8];
    for (int u = 0; u < 8; ++u) {
        const int t = lane + 32 * u;
        mine[u] = t < n ? A.perm[lo + t] : 0x7fffffff;
        rank[u] = 0;
    }
    if (n > 256) return;   // pathological density: keep arrival order (still correct, not bit-stable)
    for (int j = 0; j < n; ++j) {
        const int v = A.perm[lo + j];
        for (int u = 0; u < 8; ++u) rank[u] += (v < mine[u]);
    }
    __syncwarp();
    for (int u = 0; u < 8; ++u) {
        const int t = lane + 32 * u;
        if (t < n) A.perm[lo + rank[u]] = mine[u];
    }
}

struct GatherArgs {
    const double4 *com, *site;
    const int *perm;        // NULL: identity
    int n_mol, S;
    double f;               // box_new / box (1.0: no volume change)
    double4 *scom, *ssite;
    unsigned long long *max_dev_bits;   // atomicMax over the bits of max |site-COM| component
    unsigned int *ovl;                  // per-molecule overlap flags, cleared here
    // optional (k_pairs_v6): the same state as 96-byte rows {site xyz x3, COM xyz} + cell-local float gate coordinates
    double *rows;
    float4 *gf;
    const int *cell_of;     // [n_mol] cell of each molecule (original index)
    int ncd;
    double edge;            // box_new / ncd
    int zl_lo, zl_cnt;      // cell mode: gather only molecules whose cell lies in z-layers [zl_lo, zl_lo + zl_cnt) (mod ncd)
    // mixed topologies: molecule m owns sites [mol[m].x, mol[m].x + mol[m].y); the copy is padded to S slots per molecule,
    // padding = {scaled COM, q = 0}, stype = −1
    const int2 *mol;
    const int *atype;
    signed char *stype;
};

// cell-sorted copy of the state; for a volume trial the COMs are scaled by f and the sites
// rigidly shifted (Ewald/volumeChange.jl:62-80: coords_new = f*coords; change = coords_new -
// coords; atom_XYZ = atom_coords + change). f == 1 reproduces the state bit for bit.
static __global__ void k_gather(GatherArgs A)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    // no early return: every lane takes part in the warp reduction of max |site - COM| below (inactive lanes contribute 0)
    bool active = p < A.n_mol;
    const int m = active ? (A.perm ? A.perm[p] : p) : 0;
    if (active && A.perm && A.zl_cnt < A.ncd) {      // sharded evaluation: this rank reads its unit range and one layer above it
        const int z = A.cell_of[m] / (A.ncd * A.ncd);
        int dz = z - A.zl_lo; if (dz < 0) dz += A.ncd;
        if (dz >= A.zl_cnt) active = false;
    }
    double dev = 0.0;
    if (active) {
    A.ovl[p] = 0u;
    const double4 c = A.com[m];
    double4 cn = c;
    cn.x = A.f * c.x; cn.y = A.f * c.y; cn.z = A.f * c.z;
    const double chx = cn.x - c.x, chy = cn.y - c.y, chz = cn.z - c.z;
    A.scom[p] = cn;
    const int2 mi = A.mol ? A.mol[m] : make_int2(m * A.S, A.S);
    for (int a = 0; a < A.S; ++a) {
        if (a >= mi.y) {
            A.ssite[(size_t)p * A.S + a] = make_double4(cn.x, cn.y, cn.z, 0.0);
            A.stype[(size_t)p * A.S + a] = (signed char)-1;
            continue;
        }
        double4 s = A.site[(size_t)mi.x + a];
        if (A.stype) A.stype[(size_t)p * A.S + a] = (signed char)A.atype[mi.x + a];
        dev = fmax(dev, fmax(fabs(s.x - c.x), fmax(fabs(s.y - c.y), fabs(s.z - c.z))));
        s.x = s.x + chx; s.y = s.y + chy; s.z = s.z + chz;
        A.ssite[(size_t)p * A.S + a] = s;
        if (A.rows) { double *r = A.rows + (size_t)p * 12 + 3 * a; r[0] = s.x; r[1] = s.y; r[2] = s.z; }
    }
    if (A.rows) {
        double *r = A.rows + (size_t)p * 12 + 9;
        r[0] = cn.x; r[1] = cn.y; r[2] = cn.z;
        const int id = A.cell_of[m], n = A.ncd;
        A.gf[p] = make_float4((float)(cn.x - (double)(id % n) * A.edge), (float)(cn.y - (double)((id / n) % n) * A.edge),
                              (float)(cn.z - (double)(id / (n * n)) * A.edge), 0.f);
    }
    }
    // warp max, then one atomic per warp (non-negative doubles order like their bit patterns)
    for (int o = 16; o > 0; o >>= 1) dev = fmax(dev, __shfl_xor_sync(0xffffffffu, dev, o));
    if ((threadIdx.x & 31) == 0) atomicMax(A.max_dev_bits, (unsigned long long)__double_as_longlong(dev));
}

// accept of a volume move: the scaled state becomes the resident one (volumeChange.jl:141-144)
static __global__ void k_apply_scale(DevSystem S, double f)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= S.n_mol) return;
    const double4 c = S.com[m];
    double4 cn = c;
    cn.x = f * c.x; cn.y = f * c.y; cn.z = f * c.z;
    const double chx = cn.x - c.x, chy = cn.y - c.y, chz = cn.z - c.z;
    S.com[m] = cn;
    const int2 mi = S.mol[m];
    for (int a = 0; a < mi.y; ++a) {
        double4 s = S.site[mi.x + a];
        s.x = s.x + chx; s.y = s.y + chy; s.z = s.z + chz;
        S.site[mi.x + a] = s;
    }
}

// Σq and Σq² (EwaldSelf ewalds.jl:829-833, Wolf constants energy.jl:924-934), single CTA, ordered
static __global__ void k_charge_sums(const double4 *site, int n, double *out)
{
    __shared__ double s_red[2 * 8];
    double acc[2] = {0.0, 0.0};
    for (int l = threadIdx.x; l < n; l += 256) { const double q = site[l].w; acc[0] += q; acc[1] += q * q; }
    block_sum<2, 256>(acc, s_red);
    if (threadIdx.x == 0) { out[0] = acc[0]; out[1] = acc[1]; }
}

// ------------------------------------------------------------------- monatomic potential
// Monatomic/mainMonatomic.jl:275-289: Σ_i LJ_ΔU(i) / 2, rows over CTAs (double counted like the
// reference: the per-j ε_j, σ_j make the (i,j) and (j,i) terms different).
static __global__ void __launch_bounds__(256) k_atoms_rows(DevAtoms S, double2 *rows)
{
    __shared__ double s_red[2 * 8];
    const double L = S.box, rc2 = S.rc * S.rc;
    for (int i = blockIdx.x; i < S.n; i += gridDim.x) {
        const double4 r0 = S.r[i];
        double acc[2] = {0.0, 0.0};
        for (int j = threadIdx.x; j < S.n; j += 256) {
            if (j == i) continue;
            const double4 rj = S.r[j];
            const double dx = min_image(r0.x, rj.x, L), dy = min_image(r0.y, rj.y, L),
                         dz = min_image(r0.z, rj.z, L);
            const double r2 = dx * dx + dy * dy + dz * dz;
            if (!(r2 > rc2)) {
                const double2 es = S.es[j];
                const double sr2 = es.y * es.y / r2, sr6 = sr2 * sr2 * sr2, sr12 = sr6 * sr6;
                acc[0] += es.x * (sr12 - sr6);
                acc[1] += es.x * (2 * sr12 - sr6);
            }
        }
        block_sum<2, 256>(acc, s_red);
        if (threadIdx.x == 0) rows[i] = make_double2(acc[0] * 4.0, acc[1] * 24.0 / 3.0);
    }
}

static __global__ void __launch_bounds__(256) k_rows_sum(const double2 *rows, int n, double *out)
{
    __shared__ double s_red[2 * 8];
    double acc[2] = {0.0, 0.0};
    for (int i = threadIdx.x; i < n; i += 256) { acc[0] += rows[i].x; acc[1] += rows[i].y; }
    block_sum<2, 256>(acc, s_red);
    if (threadIdx.x == 0) { out[0] = acc[0] / 2; out[1] = acc[1] / 2; }
}

// ------------------------------------------------------------------- FP64 peak probe
// 8 independent DFMA chains per thread; reports 2 flop per DFMA.
static __global__ void __launch_bounds__(256) k_dfma_probe(double *out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
           a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 0.123) out[0] = a0;
}
