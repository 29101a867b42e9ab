// kernels_recip.cuh — full structure-factor rebuild, RecipLong (Ewald/ewalds.jl:538-604),
// SURVEY.md §8a row a6.
//
//   ρ(k) = Σ_l q_l e^{ikx x_l} e^{iky y_l} e^{ikz z_l},   E = Σ_k cfac_k |ρ(k)|²
//
// Decomposition: sites are split over CTAs (and over ranks when sharded); inside a CTA the
// sites stream through shared memory in sub-chunks whose per-site tables (q·e^{ikx}, e^{iky},
// e^{ikz}, k = 0..nk; negative k by conjugation, ewalds.jl:566-567,581-582) are built by the
// same recurrence the reference uses; a lane owns one (kx, |ky|) pair and keeps the accumulators of
// its kz values in registers (k_rhok_pairs for nk <= 6, k_rhok_big beyond).  Per-CTA partial sums go
// to HBM and are folded in CTA order, so the result is deterministic.  FP64 accumulation throughout.
#pragma once
#include <type_traits>
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
    // volume trial on the resident (unscaled) state: site' = site + (f·COM − COM) of its molecule (volumeChange.jl:62-80);
    // com == NULL: the sites are used as they are
    const double4 *com;
    double f;
    int US;                  // sites per molecule (uniform topology)
};

// coordinate d of site l as the k-space sum sees it
__device__ __forceinline__ double rhok_coord(const double4 &s, int d, const double4 *com, double f, int US, int l)
{
    double x = d == 0 ? s.x : (d == 1 ? s.y : s.z);
    if (com) {
        const double4 c = com[l / US];
        const double cd = d == 0 ? c.x : (d == 1 ? c.y : c.z);
        x = x + (f * cd - cd);
    }
    return x;
}

// ρ(k) = Σ_b partial[b][k] → out[k].  blockDim = (32 k, 32 slices): each slice folds a contiguous 1/32 of the
// CTAs in CTA order, the 32 slice sums are added in slice order (deterministic); many short dependent chains
// instead of four long ones (the kernel is pure load latency).
static __global__ void k_rhok_reduce(const double2 *partial, int nb, int nkvecs, double2 *out)
{
    __shared__ double2 s_s[32][33];
    const int k = blockIdx.x * 32 + threadIdx.x, sl = threadIdx.y;
    const int b0 = (int)((long long)nb * sl / 32), b1 = (int)((long long)nb * (sl + 1) / 32);
    double re = 0.0, im = 0.0;
    if (k < nkvecs)
        for (int b = b0; b < b1; ++b) {
            const double2 p = partial[(size_t)b * nkvecs + k];
            re += p.x; im += p.y;
        }
    s_s[sl][threadIdx.x] = make_double2(re, im);
    __syncthreads();
    if (sl == 0 && k < nkvecs) {
        double2 t = s_s[0][threadIdx.x];
        for (int j = 1; j < 32; ++j) { t.x += s_s[j][threadIdx.x].x; t.y += s_s[j][threadIdx.x].y; }
        out[k] = t;
    }
}

// E = Σ_k cfac_k |ρ(k)|² (un-scaled, ewalds.jl:599) and ρ(k) stored to both buffers (:600-601)
#define RHOKE_BLOCK 256
static __global__ void __launch_bounds__(RHOKE_BLOCK)
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

// ------------------------------------------------------------------------------------------
// k_rhok_pairs<NK> — ρ(k) rebuild, second version (the first one was LSU bound: three LDS.128 per
// eight FP64 instructions, 28 % FP64 pipe).
//
// A lane owns one (kx, |ky|) pair (28 pairs for nk = 5, k² < 27: one warp-width) and keeps the
// accumulators of ALL its kz in registers.  Per site it forms the four products
//   P1 = cx·cy  P2 = sx·sy  P3 = sx·cy  P4 = cx·sy   (cx, sx carry the charge, ewalds.jl:592-596)
// → U = e^{i(kx x + ky y)} = (P1−P2, P3+P4),  V = e^{i(kx x − ky y)} = (P1+P2, P3−P4), and for every
// kz > 0 the eight sums Σ U_r c_z, Σ U_i s_z, Σ U_r s_z, Σ U_i c_z (and the same for V), from which
//   ρ(kx, +ky, ±kz) = (A1 ∓ A2) + i(A4 ± A3)      (negative k by conjugation, ewalds.jl:566-567)
// are assembled once at the end: 2 DFMA per (site, k) instead of 8 FP64 instructions, and the
// e^{ikz z} reads are warp-wide broadcasts.  Warps of a CTA take different sites of a 64-site
// sub-chunk whose tables are built cooperatively with the reference's recurrence; per-CTA partials
// are folded in CTA order by k_rhok_reduce as before (deterministic).
// ------------------------------------------------------------------------------------------
#define RHOK2_BLOCK 256
#define RHOK2_SITES 80     // sites per sub-chunk: 3 x 80 = 240 table rows, one per thread in a single pass, ten sites per warp

struct Rhok2Args {
    const double4 *site;
    int s_begin, s_end, per_block;
    int nkvecs, npairs;          // lanes in use
    const int2 *pairs;           // [npairs] {kx, |ky|}
    const int *kindex;           // [(nk+1) x (2nk+1) x (2nk+1)] index of (kx, ky, kz) in the k list or -1
    double box;
    double2 *partial;            // [gridDim.x][nkvecs]
    const double4 *com;          // as RhokArgs: volume trial on the resident state (NULL: sites as they are)
    double f;
    int US;
};

template <int NK>
static __global__ void __launch_bounds__(RHOK2_BLOCK, 2) k_rhok_pairs(Rhok2Args A)
{
    constexpr int NW = RHOK2_BLOCK / 32;
    constexpr int NACC = 4 + 8 * NK;
    // (cos, sin) of k·x, k·y, k·z; x carries q.  Row stride 16·(NK+1 | 1) bytes: an odd number of 16-byte words, so the table
    // build (one row per thread, 16-byte stores) is free of bank conflicts (a 96-byte stride is 8-way conflicted)
    __shared__ double2 s_t[RHOK2_SITES][3][(NK + 1) | 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c0 = A.s_begin + blockIdx.x * A.per_block;
    const int c1 = min(A.s_end, c0 + A.per_block);
    const double twopi = 2.0 * 3.141592653589793;
    const int pl = min(lane, A.npairs - 1);
    const int2 pr = A.pairs[pl];
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.0;

    for (int base = c0; base < c1; base += RHOK2_SITES) {
        __syncthreads();
        for (int t = tid; t < RHOK2_SITES * 3; t += RHOK2_BLOCK) {
            const int l = t / 3, d = t - 3 * l;
            double x = 0.0, q = 0.0;
            if (base + l < c1) {
                const double4 s = A.site[base + l];
                x = rhok_coord(s, d, A.com, A.f, A.US, base + l);
                q = s.w;
            }
            const double sc = (d == 0) ? q : 1.0;
            cplx e1;
            sincos(twopi * x / A.box, &e1.im, &e1.re);          // ewalds.jl:561-564
            cplx e; e.re = 1.0; e.im = 0.0;
            s_t[l][d][0] = make_double2(sc, 0.0);
            e = e1;
#pragma unroll
            for (int k = 1; k <= NK; ++k) {
                s_t[l][d][k] = make_double2(sc * e.re, sc * e.im);
                e = cmul(e, e1);                                 // ewalds.jl:573-575
            }
        }
        __syncthreads();
        for (int l = warp; l < RHOK2_SITES; l += NW) {
            const double2 ex = s_t[l][0][pr.x], ey = s_t[l][1][pr.y];
            const double p1 = ex.x * ey.x, p2 = ex.y * ey.y, p3 = ex.y * ey.x, p4 = ex.x * ey.y;
            const double ur = p1 - p2, ui = p3 + p4, vr = p1 + p2, vi = p3 - p4;
            acc[0] += ur; acc[1] += ui; acc[2] += vr; acc[3] += vi;       // kz = 0
#pragma unroll
            for (int kz = 1; kz <= NK; ++kz) {
                const double2 ez = s_t[l][2][kz];                          // same address in every lane: broadcast
                double *a = acc + 4 + 8 * (kz - 1);
                a[0] = fma(ur, ez.x, a[0]); a[1] = fma(ui, ez.y, a[1]); a[2] = fma(ur, ez.y, a[2]); a[3] = fma(ui, ez.x, a[3]);
                a[4] = fma(vr, ez.x, a[4]); a[5] = fma(vi, ez.y, a[5]); a[6] = fma(vr, ez.y, a[6]); a[7] = fma(vi, ez.x, a[7]);
            }
        }
    }
    // ---- fold the 8 warps in warp order, then assemble ρ(k) for this lane's k-vectors
    __syncthreads();
    __shared__ double s_acc[NW * 32 * 4];
    const int W = 2 * NK + 1;
    double2 *out = A.partial + (size_t)blockIdx.x * A.nkvecs;
    // accumulators go through shared memory four at a time
#pragma unroll
    for (int s0 = 0; s0 < NACC; s0 += 4) {
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) s_acc[(warp * 32 + lane) * 4 + j] = acc[s0 + j];
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                double t = 0.0;
                for (int w = 0; w < NW; ++w) t += s_acc[(w * 32 + lane) * 4 + j];
                acc[s0 + j] = t;
            }
        }
    }
    if (warp != 0 || lane >= A.npairs) return;
    const int kx = pr.x, ky = pr.y;
    auto put = [&](int sy, int sz, int kz, double re, double im) {
        const int idx = A.kindex[(kx * W + (sy * ky + NK)) * W + (sz * kz + NK)];
        if (idx >= 0) out[idx] = make_double2(re, im);
    };
    put(+1, +1, 0, acc[0], acc[1]);
    if (ky > 0) put(-1, +1, 0, acc[2], acc[3]);
#pragma unroll
    for (int kz = 1; kz <= NK; ++kz) {
        const double *a = acc + 4 + 8 * (kz - 1);
        put(+1, +1, kz, a[0] - a[1], a[2] + a[3]);
        put(+1, -1, kz, a[0] + a[1], a[3] - a[2]);
        if (ky > 0) {
            put(-1, +1, kz, a[4] - a[5], a[6] + a[7]);
            put(-1, -1, kz, a[4] + a[5], a[7] - a[6]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// k_rhok_big — the (kx, |ky|)-per-lane register blocking of k_rhok_pairs for LARGE k-sets (nk up to 16: a converged
// Ewald sum at κ·r_cut ≈ 3.2 needs nk ≈ 12, 3.6 k k-vectors; SURVEY §8d "E2").  The (nk+1)² pairs no longer fit one
// warp-width and the 8·nk accumulators no longer fit the register file, so the k-set is cut into COMBOS = (tile of 32
// pairs) x (tile of RHOKB_ZT kz values): one combo per warp, accumulators in registers (4 + 8·ZT doubles), every warp of
// the CTA walks ALL sites of the CTA's chunk (the e^{ik·r} tables of a sub-chunk are built once, cooperatively, by the
// reference's recurrence, ewalds.jl:556-585, and read by all eight combos).  gridDim.y covers the combos in groups of eight;
// CTAs of different y rebuild the tables of the same sites — 3 sincos + 3·nk complex products per site against
// 8 x 32 x (8·ZT + 8) DFMA: under 2 %.  A CTA writes complete per-chunk partials for ITS k-vectors only, so the partial
// array [chunk][k] is the same as k_rhok_pairs' and the same tail folds it.  Sharding by k-RANGES (north star: "k-vector
// ranges split per GPU") is a range of combo groups per rank: every rank then sums all sites for its k-vectors, and the
// vectors the ranks exchange have disjoint support.
// ------------------------------------------------------------------------------------------
#define RHOKB_BLOCK 256
#define RHOKB_WARPS (RHOKB_BLOCK / 32)
#define RHOKB_SITES 64     // sites per sub-chunk: tables in dynamic shared memory, 64 x 3 x ((nk+1)|1) x 16 B (40 KB at nk = 12)
#define RHOKB_ZT_MAX 5      // kz values per combo: 4 or 5, whichever wastes less (six need 52 accumulators: spills at 128 registers)
#define MMC_MAX_NK_FULL 16     // the full-energy k-space path (per-move kernels: MMC_MAX_NK)

struct RhokBigArgs {
    const double4 *site;
    int s_begin, s_end, per_block;
    int nk, nkvecs;
    const int *kindex;           // [(nk+1) x (2nk+1) x (2nk+1)] index of (kx, ky, kz) in the k list or -1
    double box;
    double2 *partial;            // [gridDim.x][nkvecs]
    const double4 *com;          // volume trial on the resident state (NULL: sites as they are), as RhokArgs
    double f;
    int US;
    const int2 *pairs; int npairs;       // (kx, |ky|) pairs that own a k-vector, sorted by kx² + ky²
    const int2 *combos; int n_combos;    // {pair tile, kz tile} combos that hold at least one k-vector
    int group_begin;             // first combo group of this launch (k-range sharding: a rank's share of the groups)
};

// blockDim.x = 32 x (warps per CTA, <= 8): the host picks the warp count that divides the combos evenly over gridDim.y
template <int ZT>
__device__ __forceinline__ void rhok_big_body(const RhokBigArgs &A)
{
    constexpr int NACC = 4 + 8 * ZT;
    extern __shared__ __align__(16) double2 s_tab[];        // [RHOKB_SITES][3][TS], TS = (nk+1)|1 (odd: conflict-free row builds)
    const int TS = (A.nk + 1) | 1;
    auto s_t = [&](int l, int d, int k) -> double2 & { return s_tab[(l * 3 + d) * TS + k]; };
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c0 = A.s_begin + blockIdx.x * A.per_block;
    const int c1 = min(A.s_end, c0 + A.per_block);
    const double twopi = 2.0 * 3.141592653589793;
    const int NK = A.nk;
    const int combo = (A.group_begin + blockIdx.y) * (int)(blockDim.x >> 5) + warp;
    const bool live = combo < A.n_combos;
    const int2 cb = live ? A.combos[combo] : make_int2(0, 0);
    const int pt = cb.x, zt = cb.y;
    const int pi = pt * 32 + lane;
    const bool lane_on = live && pi < A.npairs;
    const int2 pr = A.pairs[lane_on ? pi : 0];
    const int kx = pr.x, ky = pr.y;
    const int kz0 = zt * ZT + 1;                 // this tile's kz = kz0 .. kz0 + ZT − 1 (kz = 0 rides with tile 0)
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.0;

    for (int base = c0; base < c1; base += RHOKB_SITES) {
        __syncthreads();
        for (int t = tid; t < RHOKB_SITES * 3; t += blockDim.x) {
            const int l = t / 3, d = t - 3 * l;
            double x = 0.0, q = 0.0;
            if (base + l < c1) {
                const double4 s = A.site[base + l];
                x = rhok_coord(s, d, A.com, A.f, A.US, base + l);
                q = s.w;
            }
            const double sc = (d == 0) ? q : 1.0;      // charge folded into the x table: (q·e_x)·e_y·e_z, ewalds.jl:592-596
            cplx e1;
            sincos(twopi * x / A.box, &e1.im, &e1.re);  // ewalds.jl:561-564
            s_t(l, d, 0) = make_double2(sc, 0.0);
            cplx e = e1;
            for (int k = 1; k <= NK; ++k) {
                s_t(l, d, k) = make_double2(sc * e.re, sc * e.im);
                e = cmul(e, e1);                         // ewalds.jl:573-575
            }
        }
        __syncthreads();
        if (!live) continue;
        // row pointers advanced per site (the lambda's index arithmetic was 40 % of the loop's instructions)
        const double2 *px = s_tab + kx, *py = s_tab + TS + ky, *pz = s_tab + 2 * TS + kz0;
        const int nz = min(ZT, NK - kz0 + 1);
        auto sweep = [&](auto full_tag) {                                  // FULL: all ZT kz of this tile exist (no predicates in the loop)
            constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll 2
            for (int l = 0; l < RHOKB_SITES; ++l, px += 3 * TS, py += 3 * TS, pz += 3 * TS) {
                const double2 ex = *px, ey = *py;
                const double p1 = ex.x * ey.x, p2 = ex.y * ey.y, p3 = ex.y * ey.x, p4 = ex.x * ey.y;
                const double ur = p1 - p2, ui = p3 + p4, vr = p1 + p2, vi = p3 - p4;
                if (zt == 0) { acc[0] += ur; acc[1] += ui; acc[2] += vr; acc[3] += vi; }      // kz = 0 rides with tile 0 (uniform)
#pragma unroll
                for (int z = 0; z < ZT; ++z) {
                    if (FULL || z < nz) {                                      // uniform
                        const double2 ez = pz[z];                              // same address in every lane: broadcast
                        double *a = acc + 4 + 8 * z;
                        a[0] = fma(ur, ez.x, a[0]); a[1] = fma(ui, ez.y, a[1]); a[2] = fma(ur, ez.y, a[2]); a[3] = fma(ui, ez.x, a[3]);
                        a[4] = fma(vr, ez.x, a[4]); a[5] = fma(vi, ez.y, a[5]); a[6] = fma(vr, ez.y, a[6]); a[7] = fma(vi, ez.x, a[7]);
                    }
                }
            }
        };
        if (nz == ZT) sweep(std::true_type{}); else sweep(std::false_type{});
    }
    if (!lane_on) return;
    const int W = 2 * NK + 1;
    double2 *out = A.partial + (size_t)blockIdx.x * A.nkvecs;
    auto put = [&](int sy, int sz, int kz, double re, double im) {
        const int idx = A.kindex[(kx * W + (sy * ky + NK)) * W + (sz * kz + NK)];
        if (idx >= 0) out[idx] = make_double2(re, im);
    };
    if (zt == 0) {
        put(+1, +1, 0, acc[0], acc[1]);
        if (ky > 0) put(-1, +1, 0, acc[2], acc[3]);
    }
#pragma unroll
    for (int z = 0; z < ZT; ++z) {
        const int kz = kz0 + z;
        if (kz > NK) break;
        const double *a = acc + 4 + 8 * z;
        put(+1, +1, kz, a[0] - a[1], a[2] + a[3]);      // negative k by conjugation, ewalds.jl:566-567
        put(+1, -1, kz, a[0] + a[1], a[3] - a[2]);
        if (ky > 0) {
            put(-1, +1, kz, a[4] - a[5], a[6] + a[7]);
            put(-1, -1, kz, a[4] + a[5], a[7] - a[6]);
        }
    }
}

// blockDim.x = 128 or 256 (registers are allocated per four warps): four or two CTAs per SM
template <int ZT>
static __global__ void __launch_bounds__(256, 2) k_rhok_big(const __grid_constant__ RhokBigArgs A) { rhok_big_body<ZT>(A); }
