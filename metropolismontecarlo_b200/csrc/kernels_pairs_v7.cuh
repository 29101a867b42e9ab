// kernels_pairs_v7.cuh — full pair energy for potential() / volume moves (Ewald/energy.jl:946-1032, :864-943;
// Ewald/volumeChange.jl:91-111), cell mode, 3-site molecules with identical per-site charges (SPC/E, TIP3P).
// Reference semantics as before: COM gate |COM_ij|² < r_cut² (strict, energy.jl:250 / ewalds.jl:337), nine
// q_a q_b erfc(κr)/r site pairs (ewalds.jl:359-367), O–O LJ with the COM-separation virial (energy.jl:270-282),
// every unordered molecule pair once.
//
// What changed against k_pairs_v6 and why (ncu source page of v6, profiles/r02_v6_sass_regions.txt: 405 M warp
// instructions per launch, of which 31 % FP64 arithmetic, 38 % the COM gate, 17 % consume bookkeeping, 14 % per-unit
// staging / fix-up / spin):
//
//   state   = a GHOST-EXTENDED cell grid in a fixed-capacity layout.  k_bin7 drops every molecule into the bucket of
//             its cell (64 slots per cell, one atomic per molecule, no prefix sum); k_gather7 (one warp per cell of the
//             (ncd+2) x (ncd+2) x (ncd+1) extended grid) orders a cell's members by molecule index and writes 96-byte rows
//             {3 sites xyz, COM xyz} + cell-local FP32 gate coordinates, the ghost cells already translated by ±L.  The
//             pair kernel therefore has no periodic-wrap logic, no slot descriptors to build (a neighbour is `cell + const`)
//             and nothing to fix up after a tile has landed.
//   gate    = a lane OWNS B molecules (one per 32-wide chunk of the B tile, coordinates in registers) and walks the
//             warp's A rows, which arrive by shared-memory broadcast as {−2a, r_c² − |a|²} (coordinates relative to the box centre,
//             |·|² precomputed by the gather):
//                 |a − b|² < r_c²   ⇔   |b|² − 2a·b < r_c² − |a|²          3 FFMA + 1 FSETP + 1 predicated LOP per test
//             into a per-lane 16-bit row mask — no ballot, no popc, no store per test: 6 instructions per 32 tests
//             instead of 26.  (FP32, conservative: the threshold carries the worst-case rounding error; the exact FP64
//             test on the same doubles as the reference is repeated on the survivors.)
//   queue   = once per unit: popc over the masks, one warp scan, every lane appends its own survivors.
//   pipeline= a fifth PRODUCER warp runs the unit sequence ahead of the four consumer warps: it draws the ticket, reads
//             the cell populations, and issues the cp.async.bulk copies — gate coordinates of unit j+1 (two-stage ring)
//             while unit j is evaluated, the rows of unit j while unit j is being gated.  Consumers never execute a
//             CTA-wide barrier and only ever wait on an mbarrier whose copy was issued a phase earlier.
#pragma once
#include "kernels_pairs.cuh"

// ---- unit = (home cell, group of 5/5/4 half-shell slots)
#define V3_GROUPS 3
static __constant__ int c_v3_group_begin[V3_GROUPS + 1] = {0, 5, 10, 14};

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

// ---- mbarrier / bulk-copy (TMA engine, SASS UBLKCP) primitives
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
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// (Measured: the poll loop is not what limits the kernel.  __nanosleep between polls and try_wait's suspend-time hint both
// return after ~20 ns on this part and leave the kernel time unchanged; a polling warp yields its issue slot.)
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

#define V7_CONSUMERS 4
#define V7_BLOCK (32 * (V7_CONSUMERS + 1))
#define V7_CAP 64          // molecules per cell (fixed-capacity layout); a denser cell makes the path decline
#define V7_BCAP 256        // rows of the B tile; a group with more neighbours is split into two sub-units
#define V7_ROW 14          // doubles per stored molecule: 12 of payload {3 sites xyz, COM xyz} + 2 of padding.  The 112-byte stride is an
                           // ODD number of 16-byte words, so the LDS.128 of eight consecutive rows hit eight different bank groups;
                           // the 96-byte stride of v6 folds every fourth row onto the same banks (ncu: 63 % of the shared-memory
                           // wavefronts of the first v7 build were conflict replays, the LSU pipe 80 % busy)
#define V7_QCAP 1024       // queue entries per consumer warp

constexpr size_t V7_SMEM = (size_t)(V7_CAP + V7_BCAP) * V7_ROW * sizeof(double) +        // rows: A | B
                           2 * (size_t)(V7_CAP + V7_BCAP) * sizeof(float4) +              // gate coordinates, two stages
                           (size_t)V7_CONSUMERS * V7_QCAP * sizeof(unsigned short);

// ---- extended grid: real cell (cx, cy, cz) is extended cell (cx+1, cy+1, cz); ghosts at x,y = 0 / ncd+1 and z = ncd
struct V7Grid {
    int ncd, EX, EY;       // EX = EY = ncd + 2; extended layers ez = 0 .. ncd
    int rank, world;
    // This rank's home cells are the contiguous range [range[rank], range[rank + 1]) of the real cells in (z, y, x) order —
    // k_partition7 cuts the order into `world` ranges of equal estimated pair work from the cell populations, on the device,
    // identically on every rank.  (One rank: {0, ncd³}.)
    const int *range;
    double edge;           // box_new / ncd
    double box_new;
};
// extended row (ez * EY + ey) of a real cell index, and which extended rows a rank with home cells [c0, c1) reads: the rows
// of its home cells and the next one (slots (·,0,0), (·,+1,0)), and the three rows around them one layer up (slots (·,−1..+1,+1))
__device__ __forceinline__ int v7_ext_row(const V7Grid &G, int c) { const int n = G.ncd; return (c / (n * n)) * G.EY + ((c / n) % n) + 1; }
__device__ __forceinline__ bool v7_row_needed(const V7Grid &G, int R, int c0, int c1)
{
    if (c1 <= c0) return false;
    const int Ra = v7_ext_row(G, c0), Rb = v7_ext_row(G, c1 - 1);
    return (R >= Ra && R <= Rb + 1) || (R >= Ra + G.EY - 1 && R <= Rb + G.EY + 1);
}
__host__ __device__ __forceinline__ int v7_ext_cells(const V7Grid &G) { return G.EX * G.EY * (G.ncd + 1); }

struct Bin7Args {
    const double4 *com;
    int n_mol;
    double inv_cell;       // ncd / box (of the resident coordinates: fractional positions do not change with the box)
    V7Grid G;
    int *count;            // [ncd³] (zeroed)
    int *bucket;           // [ncd³][V7_CAP]
    int *cell_of;          // optional [n_mol]: the molecule's cell (k_need7)
};

// one thread per molecule: cell of its COM, slot by arrival (k_gather7 orders the members afterwards)
static __global__ void k_bin7(Bin7Args A)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= A.n_mol) return;
    const double4 c = A.com[m];
    const int n = A.G.ncd;
    const int cx = cell_coord(c.x, A.inv_cell, n), cy = cell_coord(c.y, A.inv_cell, n), cz = cell_coord(c.z, A.inv_cell, n);
    const int id = cx + n * (cy + n * cz);
    if (A.cell_of) A.cell_of[m] = id;
    const int pos = atomicAdd(&A.count[id], 1);
    if (pos < V7_CAP) A.bucket[(size_t)id * V7_CAP + pos] = m;
}

// Cuts the (z, y, x) order of the cells into `world` contiguous ranges of equal estimated pair work: the work of home cell c is
// n_c · (n_c + Σ populations of its 13 half-shell neighbours) — the gate tests, to which the survivors are proportional at uniform
// density.  One CTA; deterministic, so every rank computes the same boundaries from the same (replicated) COMs.
__device__ __forceinline__ void partition7_body(const int *__restrict__ count, int n, int world, int *range)
{
    __shared__ unsigned long long s_part[1024];
    __shared__ unsigned long long s_total;
    const int tid = threadIdx.x, ncell = n * n * n;
    const int per = (ncell + 1023) / 1024;
    const int lo = min(ncell, tid * per), hi = min(ncell, lo + per);
    auto cost = [&](int c) -> unsigned long long {
        const int cx = c % n, cy = (c / n) % n, cz = c / (n * n);
        const int nc = min(count[c], V7_CAP);
        int nb = 0;
        for (int s = 0; s < 14; ++s) {
            int x = cx + c_half_shell[s][0], y = cy + c_half_shell[s][1], z = cz + c_half_shell[s][2];
            x = x < 0 ? x + n : (x >= n ? x - n : x); y = y < 0 ? y + n : (y >= n ? y - n : y); z = z >= n ? z - n : z;
            nb += min(count[x + n * (y + n * z)], V7_CAP);
        }
        return (unsigned long long)nc * nb + 64ull;     // (+ a constant per cell: the per-unit overhead of nearly empty cells)
    };
    unsigned long long mine = 0;
    for (int c = lo; c < hi; ++c) mine += cost(c);
    // exclusive prefix over the 1024 threads: warp scans + the 32 warp totals (a serial scan by one thread was 16 us of the 40)
    __shared__ unsigned long long s_wsum[32];
    unsigned long long incl = mine;
    for (int o = 1; o < 32; o <<= 1) { const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += v; }
    if ((tid & 31) == 31) s_wsum[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
        const unsigned long long w = s_wsum[tid];
        unsigned long long wi = w;
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long v = __shfl_up_sync(0xffffffffu, wi, o); if (tid >= o) wi += v; }
        s_wsum[tid] = wi - w;
        if (tid == 31) { s_total = wi; range[0] = 0; range[world] = ncell; }
    }
    __syncthreads();
    s_part[tid] = s_wsum[tid >> 5] + incl - mine;
    __syncthreads();
    // boundary r is the first cell whose prefix reaches total · r / world: found by the thread whose share contains it
    unsigned long long run = s_part[tid];
    for (int c = lo; c < hi; ++c) {
        const unsigned long long nxt = run + cost(c);
        for (int r = 1; r < world; ++r) {
            const unsigned long long target = s_total / world * r;
            if (run < target && nxt >= target) range[r] = c + 1;
        }
        run = nxt;
    }
}

// Order in which the persistent pair CTAs draw the units of one segment (a rank's or a window's home range): most expensive
// first, so that the last tickets are the cheap units and the CTAs finish together (cell populations on the lattice start range
// from 27 to 64, unit costs over a factor of five; in lattice order the slowest CTA ends one big unit after the others).
// Cost of unit (c, g) = n_c · Σ populations of the group's cells (half of n_c for the home cell itself).  Counting sort on 1024
// cost classes; the order inside a class is arbitrary — every unit writes its own partial slot, so results do not depend on it.
// One CTA per segment: segment s = blockIdx.x + seg0 holds the cells [range[s], range[s + 1]); order[3·range[s] + t] = global unit.
__device__ __forceinline__ void order7_body(const int *__restrict__ count, int n, const int *range, int seg, int *__restrict__ order)
{
    __shared__ unsigned s_hist[1024];
    const int tid = threadIdx.x;
    const int c0 = __ldcg(range + seg), c1 = __ldcg(range + seg + 1);
    const int u0 = V3_GROUPS * c0, nu = V3_GROUPS * (c1 - c0);
    auto cost = [&](int ug) -> unsigned {
        const int c = ug / V3_GROUPS, g = ug - c * V3_GROUPS;
        const int cx = c % n, cy = (c / n) % n, cz = c / (n * n);
        const int nc = min(count[c], V7_CAP);
        int nb2 = 0;                                            // twice the partner count
        for (int s = c_v3_group_begin[g]; s < c_v3_group_begin[g + 1]; ++s) {
            int x = cx + c_half_shell[s][0], y = cy + c_half_shell[s][1], z = cz + c_half_shell[s][2];
            x = x < 0 ? x + n : (x >= n ? x - n : x); y = y < 0 ? y + n : (y >= n ? y - n : y); z = z >= n ? z - n : z;
            const int k = min(count[x + n * (y + n * z)], V7_CAP);
            nb2 += s == 0 ? k : 2 * k;
        }
        return (unsigned)(nc * nb2);
    };
    s_hist[tid] = 0u;
    __syncthreads();
    // classes on the absolute scale (cost <= 64 · 2 · 5 · 64): no pass for the maximum
    auto cls = [&](unsigned cst) -> int { return 1023 - (int)min(1023u, cst / 40u); };   // 0 = most expensive
    for (int t = tid; t < nu; t += 1024) atomicAdd(&s_hist[cls(cost(u0 + t))], 1u);
    __syncthreads();
    // exclusive scan of the 1024 class sizes (one value per thread: warp scans + the 32 warp totals)
    __shared__ unsigned s_wsum[32];
    const unsigned mine = s_hist[tid];
    unsigned incl = mine;
    for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += v; }
    if ((tid & 31) == 31) s_wsum[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
        unsigned w = s_wsum[tid], wi = w;
        for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, wi, o); if (tid >= o) wi += v; }
        s_wsum[tid] = wi - w;
    }
    __syncthreads();
    s_hist[tid] = s_wsum[tid >> 5] + incl - mine;
    __syncthreads();
    for (int t = tid; t < nu; t += 1024) {
        const unsigned pos = atomicAdd(&s_hist[cls(cost(u0 + t))], 1u);
        order[u0 + pos] = u0 + t;
    }
}

static __global__ void __launch_bounds__(1024) k_partition7(const int *__restrict__ count, int n, int world, int *range)
{
    partition7_body(count, n, world, range);
}
static __global__ void __launch_bounds__(1024) k_order7(const int *__restrict__ count, int n, const int *__restrict__ range, int seg0, int *__restrict__ order)
{
    order7_body(count, n, range, blockIdx.x + seg0, order);
}
// both in one launch (one CTA): the ranks' ranges, then the draw order of this rank's units
static __global__ void __launch_bounds__(1024) k_partition_order7(const int *__restrict__ count, int n, int world, int *range, int rank, int *__restrict__ order)
{
    partition7_body(count, n, world, range);
    __threadfence();
    __syncthreads();
    order7_body(count, n, range, rank, order);
}

// mmc_potential_host on one GPU, windowed: the home cells are cut into `nwin` contiguous ranges (range[0 .. nwin]) that are
// gathered and evaluated one after the other while later site chunks are still on the bus; need[w] = the last chunk (sites
// [n_sites·c/n_chunks, n_sites·(c+1)/n_chunks)) that holds a molecule window w reads.
static __global__ void k_window_need7(const int *__restrict__ cell_of, int n_mol, int US, int n_sites, int n_chunks, V7Grid G, int nwin, int *need)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_mol) return;
    const int n = G.ncd, c = cell_of[m];
    const int cy = (c / n) % n, cz = c / (n * n);
    const long long last_site = (long long)m * US + US - 1;
    int ch = (int)(last_site * n_chunks / n_sites);
    while (ch + 1 < n_chunks && (long long)n_sites * (ch + 1) / n_chunks <= last_site) ++ch;
    while (ch > 0 && (long long)n_sites * ch / n_chunks > last_site) --ch;
    for (int w = 0; w < nwin; ++w) {
        const int c0 = G.range[w], c1 = G.range[w + 1];
        bool nd = false;
#pragma unroll
        for (int gz = 0; gz < 2; ++gz) {
            if (gz == 1 && cz != 0) continue;
            const int ez = gz ? n : cz;
            nd |= v7_row_needed(G, ez * G.EY + cy + 1, c0, c1);
            if (cy == 0) nd |= v7_row_needed(G, ez * G.EY + n + 1, c0, c1);
            if (cy == n - 1) nd |= v7_row_needed(G, ez * G.EY + 0, c0, c1);
        }
        if (nd) atomicMax(&need[w], ch);
    }
}

// domain-decomposed host evaluation: which blocks of 256 molecules hold a molecule this rank reads (its home range, the
// half shell around it, the ghost images included)
static __global__ void k_need7(const int *__restrict__ cell_of, int n_mol, V7Grid G, unsigned char *need)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_mol) return;
    const int c0 = G.range[G.rank], c1 = G.range[G.rank + 1];
    const int n = G.ncd, c = cell_of[m];
    const int cy = (c / n) % n, cz = c / (n * n);
    bool nd = false;
    // the molecule's rows in the extended grid: its own, the y-ghost (cy = 0 -> row n+1, cy = n-1 -> row 0), the z-ghost layer n
#pragma unroll
    for (int gz = 0; gz < 2; ++gz) {
        if (gz == 1 && cz != 0) continue;
        const int ez = gz ? n : cz;
        nd |= v7_row_needed(G, ez * G.EY + cy + 1, c0, c1);
        if (cy == 0) nd |= v7_row_needed(G, ez * G.EY + n + 1, c0, c1);
        if (cy == n - 1) nd |= v7_row_needed(G, ez * G.EY + 0, c0, c1);
    }
    if (nd) need[m >> 8] = 1;
}

struct Gather7Args {
    const double4 *com, *site;     // resident state (3 sites per molecule)
    const int *count, *bucket;
    V7Grid G;
    double f;                      // box_new / box (1.0: no volume change)
    double *rows;                  // [ext cells][V7_CAP][12]
    float4 *gf;                    // [ext cells][V7_CAP]
    int *ecount;                   // [ext cells]
    unsigned long long *max_dev_bits;
    unsigned int *err_flag;
    int *max_count;
};

// one warp per extended cell: members ordered by molecule index (so every sum downstream is independent of the
// atomics' arrival order), scaled for a volume trial (volumeChange.jl:62-80: COM' = f·COM, sites shifted rigidly),
// ghosts translated by ±L'
static __global__ void __launch_bounds__(256) k_gather7(Gather7Args A)
{
    const int lane = threadIdx.x & 31;
    const int wcell = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const V7Grid &G = A.G;
    const int per_layer = G.EX * G.EY;
    if (wcell >= per_layer * (G.ncd + 1)) return;
    const int ez = wcell / per_layer, rem = wcell - ez * per_layer;
    const int ey = rem / G.EX, ex = rem - ey * G.EX;
    if (G.world > 1 && !v7_row_needed(G, ez * G.EY + ey, G.range[G.rank], G.range[G.rank + 1])) return;
    const int e = wcell;
    const int n = G.ncd;
    int rx = ex - 1, ry = ey - 1, rz = ez;
    double shx = 0.0, shy = 0.0, shz = 0.0;
    if (rx < 0) { rx += n; shx = -G.box_new; } else if (rx >= n) { rx -= n; shx = G.box_new; }
    if (ry < 0) { ry += n; shy = -G.box_new; } else if (ry >= n) { ry -= n; shy = G.box_new; }
    if (rz >= n) { rz -= n; shz = G.box_new; }
    const int c = rx + n * (ry + n * rz);
    int cnt = A.count[c];
    if (lane == 0) atomicMax(A.max_count, cnt);
    if (cnt > V7_CAP) { if (lane == 0) atomicExch(A.err_flag, 1u); cnt = 0; }
    if (lane == 0) A.ecount[e] = cnt;
    const int *b = A.bucket + (size_t)c * V7_CAP;
    const int v0 = lane < cnt ? b[lane] : 0x7fffffff, v1 = lane + 32 < cnt ? b[lane + 32] : 0x7fffffff;
    int r0 = 0, r1 = 0;
    for (int t = 0; t < cnt; ++t) {
        const int v = t < 32 ? __shfl_sync(0xffffffffu, v0, t) : __shfl_sync(0xffffffffu, v1, t - 32);
        r0 += v < v0; r1 += v < v1;
    }
    const double hb = 0.5 * G.box_new;                    // gate coordinates are relative to the box centre
    double dev = 0.0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int m = h ? v1 : v0, r = h ? r1 : r0;
        if (m == 0x7fffffff) continue;
        const double4 cm = A.com[m];
        const double cnx = A.f * cm.x, cny = A.f * cm.y, cnz = A.f * cm.z;
        const double chx = cnx - cm.x, chy = cny - cm.y, chz = cnz - cm.z;
        double v[12];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double4 s = A.site[(size_t)m * 3 + a];
            dev = fmax(dev, fmax(fabs(s.x - cm.x), fmax(fabs(s.y - cm.y), fabs(s.z - cm.z))));
            v[3 * a] = (s.x + chx) + shx; v[3 * a + 1] = (s.y + chy) + shy; v[3 * a + 2] = (s.z + chz) + shz;
        }
        v[9] = cnx + shx; v[10] = cny + shy; v[11] = cnz + shz;
        double2 *dst = reinterpret_cast<double2 *>(A.rows + ((size_t)e * V7_CAP + r) * V7_ROW);
#pragma unroll
        for (int k = 0; k < 6; ++k) dst[k] = make_double2(v[2 * k], v[2 * k + 1]);
        const float gx = (float)(v[9] - hb), gy = (float)(v[10] - hb), gz = (float)(v[11] - hb);
        A.gf[(size_t)e * V7_CAP + r] = make_float4(gx, gy, gz, fmaf(gz, gz, fmaf(gy, gy, gx * gx)));
    }
    for (int o = 16; o > 0; o >>= 1) dev = fmax(dev, __shfl_xor_sync(0xffffffffu, dev, o));
    if (lane == 0) atomicMax(A.max_dev_bits, (unsigned long long)__double_as_longlong(dev));
}

struct V7Args {
    V7Grid G;                      // units = 3 groups x the home cells [G.range[rank], G.range[rank + 1])
    const double *rows;
    const float4 *gf;
    const int *ecount;
    float gate_rc2f;               // conservative FP32 gate threshold (>= r_cut² + worst-case rounding of the dot form)
    double rc_qq2;
    long long rcqq_bits;
    double qq_tab[9];
    unsigned qq_negmask;
    double lj_eps, lj_sig2;
    double pc[MMC_ERF_MAXDEG + 1];
    double pk2s;
    const double *max_dev;
    unsigned int *err_flag;
    unsigned int *n_ovl;
    unsigned int *ticket;          // zeroed before the launch
    const int *order;              // [3 ncd³] k_order7: the t-th unit this rank draws is order[3·range[rank] + t] (global unit id)
    double4 *unit_partial;         // [3 ncd³][V7_CONSUMERS], indexed by the GLOBAL unit 3·cell + group
};

// what the producer publishes per sub-unit (ring of four, guarded by the gate-coordinate barriers)
struct V7Desc {
    int valid;                     // 0: this CTA's unit sequence has ended
    int nA, nB, nsl;
    int self_n;                    // nA when slot 0 of this sub-unit is the home cell itself (keep q > p), else 0
    int rot;                       // rotation of the warps over the home cell's row blocks (spreads the short last block over the warps)
    long long unit;                // partial slot of this sub-unit's unit
    int first;                     // 1: first sub-unit of its unit (store the sums), 0: add to what is there
};

template <int DEG, bool DIRECT>
static __global__ void __launch_bounds__(V7_BLOCK, 4) k_pairs_v7(const __grid_constant__ V7Args A)
{
    constexpr int S = 3;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_rowA = reinterpret_cast<double *>(smem_raw);
    double *s_rowB = s_rowA + V7_CAP * V7_ROW;
    float4 *s_gf = reinterpret_cast<float4 *>(s_rowB + V7_BCAP * V7_ROW);          // [2][V7_CAP + V7_BCAP]
    unsigned short *s_queue = reinterpret_cast<unsigned short *>(s_gf + 2 * (V7_CAP + V7_BCAP));
    __shared__ V7Desc s_desc[4];
    __shared__ float4 s_a2[V7_CONSUMERS][16];
    __shared__ __align__(8) unsigned long long s_full_gf[2], s_empty_gf[2], s_full_rows, s_empty_rows;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&s_full_gf[0], 1); mbar_init(&s_full_gf[1], 1);
        mbar_init(&s_empty_gf[0], V7_CONSUMERS); mbar_init(&s_empty_gf[1], V7_CONSUMERS);
        mbar_init(&s_full_rows, 1); mbar_init(&s_empty_rows, V7_CONSUMERS);
        // the always-true cut-off tests (energy.jl:270, ewalds.jl:362) must really be always true for this state
        const double reach = sqrt(A.rc_qq2) + 2.0 * (*A.max_dev);
        if (!(reach * reach < A.rc_qq2 + 100.0) && blockIdx.x == 0) atomicExch(A.err_flag, 1u);
    }
    __syncthreads();                                       // the only CTA-wide barrier of the kernel

    if (warp == V7_CONSUMERS) {
        // ======================================================================== producer warp
        const int n = A.G.ncd, EX = A.G.EX, EY = A.G.EY;
        const int c_first = A.G.range[A.G.rank];
        const long long n_units = (long long)V3_GROUPS * (A.G.range[A.G.rank + 1] - c_first);
        long long u = blockIdx.x;
        unsigned seq = 0;
        while (u < n_units) {
            const long long ug = A.order[(long long)V3_GROUPS * c_first + u];            // global unit
            const int c = (int)(ug / V3_GROUPS), g = (int)(ug - (long long)c * V3_GROUPS);
            const int cz = c / (n * n), r2 = c - cz * n * n, cy = r2 / n, cx = r2 - cy * n;
            const int e = (cx + 1) + EX * ((cy + 1) + EY * cz);
            const int sl0 = c_v3_group_begin[g], nsl_all = c_v3_group_begin[g + 1] - sl0;
            int en = e, cnt = 0;
            if (lane < nsl_all) {
                const int s = sl0 + lane;
                en = e + c_half_shell[s][0] + EX * (c_half_shell[s][1] + EY * c_half_shell[s][2]);
                cnt = A.ecount[en];
            } else if (lane == 5) cnt = A.ecount[e];
            const int nA = __shfl_sync(FULL, cnt, 5);
            // a unit without work (an empty home cell or only empty neighbours) never reaches the consumers: its slots are zeroed here
            const bool any_b = __any_sync(FULL, lane < nsl_all && cnt > 0);
            if ((nA == 0 || !any_b) && lane < V7_CONSUMERS) A.unit_partial[(size_t)ug * V7_CONSUMERS + lane] = make_double4(0.0, 0.0, 0.0, 0.0);
            int pass = 0;
            for (int s_begin = 0; s_begin < nsl_all && nA > 0;) {
                int incl = (lane >= s_begin && lane < nsl_all) ? cnt : 0;          // inclusive prefix of the B counts from s_begin on
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) { const int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
                // slots [s_begin, se) fit the B tile (every count is <= V7_CAP, so at least four do)
                const int se = min(nsl_all, __popc(__ballot_sync(FULL, lane < 5 && incl <= V7_BCAP)));
                const int nB = __shfl_sync(FULL, incl, se - 1);
                if (nB > 0) {
                    const int sg = seq & 1;
                    V7Desc &D = s_desc[seq & 3];
                    mbar_wait(&s_empty_gf[sg], ((seq >> 1) & 1u) ^ 1u);         // consumers are done gating sub-unit seq − 2
                    const bool mine = lane >= s_begin && lane < se;
                    if (lane == 0) {
                        D.valid = 1; D.nA = nA; D.nB = nB; D.nsl = se - s_begin;
                        D.self_n = (g == 0 && s_begin == 0) ? nA : 0;
                        D.rot = (int)(ug & 3); D.unit = ug; D.first = pass == 0 ? 1 : 0;
                    }
                    __syncwarp();
                    float4 *gfA = s_gf + sg * (V7_CAP + V7_BCAP), *gfB = gfA + V7_CAP;
                    if (lane == 0) mbar_arrive_expect_tx(&s_full_gf[sg], (unsigned)(nA + nB) * 16u);
                    __syncwarp();
                    if (mine && cnt > 0) bulk_g2s(gfB + (incl - cnt), A.gf + (size_t)en * V7_CAP, (unsigned)cnt * 16u, &s_full_gf[sg]);
                    if (lane == 5) bulk_g2s(gfA, A.gf + (size_t)e * V7_CAP, (unsigned)nA * 16u, &s_full_gf[sg]);
                    // the rows follow as soon as the consumers have left the previous sub-unit: they land while this one is gated
                    mbar_wait(&s_empty_rows, (seq & 1u) ^ 1u);
                    if (lane == 0) mbar_arrive_expect_tx(&s_full_rows, (unsigned)(nA + nB) * (V7_ROW * 8u));
                    __syncwarp();
                    if (mine && cnt > 0)
                        bulk_g2s(s_rowB + (size_t)(incl - cnt) * V7_ROW, A.rows + (size_t)en * V7_CAP * V7_ROW, (unsigned)cnt * (V7_ROW * 8u), &s_full_rows);
                    if (lane == 5) bulk_g2s(s_rowA, A.rows + (size_t)e * V7_CAP * V7_ROW, (unsigned)nA * (V7_ROW * 8u), &s_full_rows);
                    ++seq; ++pass;
                }
                s_begin = se;
            }
            // the next unit is claimed only now — when this one's rows are on their way, i.e. the consumers are a whole unit
            // behind: a ticket drawn further ahead is a unit another CTA cannot take when the queue runs dry (sharded ranks
            // have four units per CTA; three claimed ahead left SMs idle for a third of the kernel)
            unsigned tk = 0;
            if (lane == 0) tk = atomicAdd(A.ticket, 1u);
            u = (long long)gridDim.x + __shfl_sync(FULL, tk, 0);
        }
        // end marker
        const int sg = seq & 1;
        mbar_wait(&s_empty_gf[sg], ((seq >> 1) & 1u) ^ 1u);
        if (lane == 0) { s_desc[seq & 3].valid = 0; mbar_arrive(&s_full_gf[sg]); }
        return;
    }

    // ============================================================================ consumer warps
    unsigned short *q = s_queue + warp * V7_QCAP;
    float4 *a2 = s_a2[warp];
    const float rc2f = A.gate_rc2f;
    const double lj_eps = A.lj_eps, lj_sig2 = A.lj_sig2;
    double acc_lj = 0.0, acc_vir = 0.0, acc_q = 0.0;
    unsigned long long my_pairs = 0;

    auto consume = [&](unsigned e, bool have) {          // one queued molecule pair per lane
        const int p = e & 63u, qi = e >> 6;
        const double2 *ra = reinterpret_cast<const double2 *>(s_rowA + p * V7_ROW);
        const double2 *rb = reinterpret_cast<const double2 *>(s_rowB + qi * V7_ROW);
        const double2 a4 = ra[4], a5 = ra[5], b4 = rb[4], b5 = rb[5];
        // exact gate on the FP64 COMs (strict <, energy.jl:250 / ewalds.jl:337); un-contracted, left to right, like Julia
        // evaluates rij[1]*rij[1] + rij[2]*rij[2] + rij[3]*rij[3]
        const double rx = b4.y - a4.y, ry = b5.x - a5.x, rz = b5.y - a5.y;
        const double r2com = __dadd_rn(__dadd_rn(__dmul_rn(rx, rx), __dmul_rn(ry, ry)), __dmul_rn(rz, rz));
        const bool act = have && (__double_as_longlong(r2com) < A.rcqq_bits);
        const unsigned am = __ballot_sync(FULL, act);
        if (lane == 0) my_pairs += __popc(am);
        if (act) {
            const double2 a0 = ra[0], a1 = ra[1], a2r = ra[2], a3 = ra[3];
            const double2 b0 = rb[0], b1 = rb[1], b2 = rb[2], b3 = rb[3];
            const double ax[S] = {a0.x, a1.y, a3.x}, ay[S] = {a0.y, a2r.x, a3.y}, az[S] = {a1.x, a2r.y, a4.x};
            const double bx[S] = {b0.x, b1.y, b3.x}, by[S] = {b0.y, b2.x, b3.y}, bz[S] = {b1.x, b2.y, b4.x};
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
            if (hmin < 0x3FE00000) {   // some site pair has r² < 0.5: the sign rule q_a q_b < 0 (ewalds.jl:359) decides
                bool ov = false;
#pragma unroll
                for (int j = 0; j < S * S; ++j) ov |= ((A.qq_negmask >> j) & 1u) && __double2hiint(r2[j]) < 0x3FE00000;
                if (ov) atomicAdd(A.n_ovl, 1u);   // the overlap rule zeroes whole rows: the host re-evaluates on the general path
            }
#pragma unroll
            for (int j = 0; j < S * S; ++j) ri[j] = fast_rsqrt(r2[j]);
            {   // smooth part of erfc(κr)/r as one polynomial: Horner in r² (DIRECT) or in s = σκ²r² − 1
                double x[S * S];
#pragma unroll
                for (int j = 0; j < S * S; ++j) { x[j] = DIRECT ? r2[j] : fma(r2[j], A.pk2s, -1.0); pv[j] = A.pc[DEG]; }
#pragma unroll
                for (int k = DEG - 1; k >= 0; --k)
#pragma unroll
                    for (int j = 0; j < S * S; ++j) pv[j] = fma(pv[j], x[j], A.pc[k]);
            }
#pragma unroll
            for (int j = 0; j < S * S; ++j) acc_q = fma(A.qq_tab[j], ri[j] + pv[j], acc_q);   // ewalds.jl:366-367
            {   // LJ 12-6 on the O–O pair (energy.jl:270-282), virial with the COM separation
                const double rinv2 = ri[0] * ri[0];
                const double s2 = lj_sig2 * rinv2, s6 = s2 * s2 * s2, s12 = s6 * s6;
                acc_lj += lj_eps * (s12 - s6);
                const double wv = lj_eps * (2.0 * s12 - s6) * s2;
                acc_vir += wv * (rx * ddx + ry * ddy + rz * ddz);
            }
        }
    };

    for (unsigned seq = 0;; ++seq) {
        const int sg = seq & 1;
        mbar_wait(&s_full_gf[sg], (seq >> 1) & 1u);
        const V7Desc &D = s_desc[seq & 3];
        if (!D.valid) break;
        const int nA = D.nA, nB = D.nB, self_n = D.self_n;
        // warp w owns the contiguous rows [pb, pb + nr) of the home cell (blocks of ceil(nA / 4), rotated by unit)
        const int nrw = (nA + V7_CONSUMERS - 1) / V7_CONSUMERS;
        const int pb = ((warp + D.rot) & (V7_CONSUMERS - 1)) * nrw;
        const int nr = max(0, min(nrw, nA - pb));
        const float4 *gfA = s_gf + sg * (V7_CAP + V7_BCAP), *gfB = gfA + V7_CAP;
        // ---- this warp's A rows as {−2a, r_c² − |a|²}; the tail of the table never passes
        if (lane < 16) {
            float4 t = make_float4(0.f, 0.f, 0.f, -INFINITY);
            if (lane < nr) {
                const float4 a = gfA[pb + lane];
                t = make_float4(-2.f * a.x, -2.f * a.y, -2.f * a.z, rc2f - a.w);
            }
            a2[lane] = t;
        }
        __syncwarp();
        const int nch = nr > 0 ? (nB + 31) >> 5 : 0;
        const int nr4 = (nr + 3) & ~3;
        bool rows_in = false, gf_released = false;
        // Gate chunk after chunk (lane owns B molecule t = 32 j + lane; 16-bit mask over the warp's rows), each chunk's survivors
        // appended to the queue behind those of the lanes below (one warp scan per chunk).  Normally the whole sub-unit is queued
        // before the first round is evaluated; a sub-unit with more survivors than the queue holds is evaluated in several goes.
        for (int j = 0;;) {
            int tail = 0;
            for (; j < nch; ++j) {
                const int t = 32 * j + lane;
                const bool valid = t < nB;
                const float4 g = gfB[valid ? t : 0];
                const float bx = g.x, by = g.y, bz = g.z;
                const float bw = valid ? g.w : INFINITY;
                unsigned mask = 0u;
                for (int r4 = nr4 - 4; r4 >= 0; r4 -= 4) {          // four rows per step, highest rows first: mask = mask << 4 | nibble
                    unsigned nib = 0u;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float4 a = a2[r4 + k];
                        const float tt = fmaf(a.z, bz, fmaf(a.y, by, fmaf(a.x, bx, bw)));
                        if (tt < a.w) nib |= 1u << k;
                    }
                    mask = (mask << 4) | nib;
                }
                if (t < self_n) {
                    // home cell against itself: every unordered pair {p, t} once, and about half of its pairs to every row
                    // (so that the contiguous row blocks of the four warps carry equal work): row p < t takes the pair when
                    // p + t is even, row p > t when p + t is odd.  p = pb + r.
                    const int nv = t - pb;                          // rows r < nv are below t, r == nv is t itself
                    const unsigned lo = nv >= 16 ? 0xffffu : (nv > 0 ? ((1u << nv) - 1u) : 0u);
                    const unsigned hi = nv >= 15 ? 0u : (nv < 0 ? 0xffffu : (0xffffu << (nv + 1)) & 0xffffu);
                    const unsigned even = ((pb + t) & 1) ? 0xaaaau : 0x5555u;      // rows r with pb + r + t even
                    mask &= (lo & even) | (hi & ~even);
                }
                const int cnt = __popc(mask);
                int incl = cnt;
#pragma unroll
                for (int o2 = 1; o2 < 32; o2 <<= 1) { const int v = __shfl_up_sync(FULL, incl, o2); if (lane >= o2) incl += v; }
                const int tot = __shfl_sync(FULL, incl, 31);
                if (tail + tot > V7_QCAP) break;                     // no room: evaluate what is queued, then gate this chunk again
                int base = tail + incl - cnt;
                const unsigned eb = ((unsigned)t << 6) | (unsigned)pb;
                while (mask) {
                    const int r = __ffs(mask) - 1;
                    mask &= mask - 1u;
                    q[base++] = (unsigned short)(eb + (unsigned)r);
                }
                tail += tot;
            }
            __syncwarp();
            if (j >= nch && !gf_released) { if (lane == 0) mbar_arrive(&s_empty_gf[sg]); gf_released = true; }   // the producer may refill this gate stage
            if (!rows_in) { mbar_wait(&s_full_rows, seq & 1u); rows_in = true; }
            for (int b = 0; b < tail; b += 32) {
                const bool have = b + lane < tail;
                consume(have ? q[b + lane] : 0u, have);
            }
            __syncwarp();
            if (j >= nch) break;
        }
        // ---- this sub-unit's sums, per warp, at a place that depends on the unit only (folded in unit order afterwards)
        const double w0 = warp_sum(acc_lj), w1 = warp_sum(acc_vir), w2 = warp_sum(acc_q);
        if (lane == 0) {
            double4 *slot = A.unit_partial + (size_t)D.unit * V7_CONSUMERS + warp;
            double4 v = make_double4(w0, w1, w2, (double)my_pairs);
            if (!D.first) { const double4 o = *slot; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }   // (this lane wrote it a sub-unit ago)
            *slot = v;
        }
        acc_lj = 0.0; acc_vir = 0.0; acc_q = 0.0; my_pairs = 0;
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty_rows);             // the producer may refill the rows
    }
}

// ------------------------------------------------------------------------------------------------------------
// k_eval_tail — everything after the two big kernels in one launch.  Block b folds its contiguous share of the
// per-(unit, warp) pair sums (fixed order) and its share of the k-vectors of the ρ(k) partials (CTA order); the last
// block to finish (ticket) adds the block sums in block order, forms E_recip = Σ cfac |ρ(k)|² (ewalds.jl:599),
// stores ρ(k) to the resident buffers (:600-601), and either publishes the scalars to the host (one rank) or pushes
// this rank's vector into every peer's exchange buffer (sharded evaluation, kernels_peer.cuh).
// ------------------------------------------------------------------------------------------------------------
#define TAIL_BLOCKS 128
#define TAIL_THREADS 256

struct TailArgs {
    const double4 *unit_partial;   // pair sums [units][V7_CONSUMERS], units = 3 x (range[rank + 1] − range[rank])
    const int *range; int rank;
    const double2 *rhok_partial; int rhok_blocks, nkvecs;   // ρ(k) partials [rhok_blocks][nkvecs] (nkvecs == 0: no k-space)
    double2 *rhok_scratch;         // [16][nkvecs] slice sums
    double4 *block_sums;           // [TAIL_BLOCKS]
    unsigned int *done;            // ticket of the tail blocks (zeroed)
    const unsigned int *n_ovl, *err_flag; const int *max_count;
    double *vec;                   // [MMC_NSCAL + 2 nkvecs]: the rank's partial-sum vector (sharding.py layout)
    // one rank: finish here
    int finish;                    // 1: E_recip, resident ρ(k), host result
    const double *cfac; double2 *dst0, *dst1;
    double *host_out;              // mapped pinned: [0..7] vec head, [8] sequence number (written last)
    unsigned long long seq;
    // sharded evaluation: the last block stores the vector into slot `rank` of every rank's exchange buffer (kernels_peer.cuh)
    int push;
    PeerArgs peer;
    // ... and then waits for every rank's vector, adds them in rank order and finishes (k_peer_sum_finish's body)
    int peer_finish;
    PeerFinishArgs fin;
    double *fin_out;               // [nvec] the summed vector
};

static __global__ void __launch_bounds__(TAIL_THREADS) k_eval_tail(const __grid_constant__ TailArgs A)
{
    __shared__ double s_red[4 * (TAIL_THREADS / 32)];
    __shared__ double2 s_s[8][33];
    __shared__ int s_last;
    const int tid = threadIdx.x, b = blockIdx.x;
    {   // pair sums: contiguous share; four independent running sums per thread (the loads are in flight together), fixed tree
        const long long p0 = (long long)V3_GROUPS * V7_CONSUMERS * A.range[A.rank], p1 = (long long)V3_GROUPS * V7_CONSUMERS * A.range[A.rank + 1];
        const long long per = (p1 - p0 + gridDim.x - 1) / gridDim.x;
        const long long lo = p0 + per * b < p1 ? p0 + per * b : p1, hi = lo + per < p1 ? lo + per : p1;
        double v[4] = {0.0, 0.0, 0.0, 0.0}, w[4] = {0.0, 0.0, 0.0, 0.0};
        long long i = lo + tid;
        for (; i + TAIL_THREADS < hi; i += 2 * TAIL_THREADS) {
            const double4 p = A.unit_partial[i], q = A.unit_partial[i + TAIL_THREADS];
            v[0] += p.x; v[1] += p.y; v[2] += p.z; v[3] += p.w;
            w[0] += q.x; w[1] += q.y; w[2] += q.z; w[3] += q.w;
        }
        if (i < hi) { const double4 p = A.unit_partial[i]; v[0] += p.x; v[1] += p.y; v[2] += p.z; v[3] += p.w; }
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] += w[k];
        block_sum<4, TAIL_THREADS>(v, s_red);
        if (tid == 0) A.block_sums[b] = make_double4(v[0], v[1], v[2], v[3]);
    }
    // ρ(k) partials [rhok_blocks][nkvecs]: the blocks form a (k-group of 32) x (slice of the CTA partials) grid — every thread
    // adds a handful of partials (the kernel is pure load latency: the more loads in flight, the better), a block folds its eight
    // sub-slices in order into scratch[slice][k], and the last block adds the slices in slice order.  Deterministic.
    const int n_groups = (A.nkvecs + 31) / 32;
    const int n_slices = n_groups > 0 ? max(1, min(16, (int)gridDim.x / n_groups)) : 1;
    for (int w = b; w < n_groups * n_slices; w += gridDim.x) {
        const int kg = w % n_groups, sl = w / n_groups;
        const int k = 32 * kg + (tid & 31), sub = tid >> 5;
        const int s0 = (int)((long long)A.rhok_blocks * sl / n_slices), s1 = (int)((long long)A.rhok_blocks * (sl + 1) / n_slices);
        const int c0 = s0 + (int)((long long)(s1 - s0) * sub / 8), c1 = s0 + (int)((long long)(s1 - s0) * (sub + 1) / 8);
        double re = 0.0, im = 0.0, re2 = 0.0, im2 = 0.0;
        if (k < A.nkvecs) {
            int c = c0;
            for (; c + 1 < c1; c += 2) {
                const double2 p = A.rhok_partial[(size_t)c * A.nkvecs + k], q = A.rhok_partial[(size_t)(c + 1) * A.nkvecs + k];
                re += p.x; im += p.y; re2 += q.x; im2 += q.y;
            }
            if (c < c1) { const double2 p = A.rhok_partial[(size_t)c * A.nkvecs + k]; re += p.x; im += p.y; }
            re += re2; im += im2;
        }
        __syncthreads();
        s_s[sub][tid & 31] = make_double2(re, im);
        __syncthreads();
        if (sub == 0 && k < A.nkvecs) {
            double2 t = s_s[0][tid];
            for (int j = 1; j < 8; ++j) { t.x += s_s[j][tid].x; t.y += s_s[j][tid].y; }
            A.rhok_scratch[(size_t)sl * A.nkvecs + k] = t;
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(A.done, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    if (tid < (int)gridDim.x) { const double4 p = ldcg4(&A.block_sums[tid]); v[0] = p.x; v[1] = p.y; v[2] = p.z; v[3] = p.w; }
    block_sum<4, TAIL_THREADS>(v, s_red);
    double er[1] = {0.0};
    for (int k = tid; k < A.nkvecs; k += TAIL_THREADS) {
        double2 s = __ldcg(&A.rhok_scratch[k]);
        for (int j = 1; j < n_slices; ++j) { const double2 t = __ldcg(&A.rhok_scratch[(size_t)j * A.nkvecs + k]); s.x += t.x; s.y += t.y; }
        A.vec[MMC_NSCAL + 2 * k] = s.x; A.vec[MMC_NSCAL + 2 * k + 1] = s.y;
        if (A.finish) {
            er[0] += A.cfac[k] * (s.x * s.x + s.y * s.y);
            if (A.dst0) A.dst0[k] = s;
            if (A.dst1) A.dst1[k] = s;
        }
    }
    if (A.finish && A.nkvecs > 0) block_sum<1, TAIL_THREADS>(er, s_red);
    __shared__ double s_h[MMC_NSCAL];
    if (tid == 0) {
        s_h[0] = v[0]; s_h[1] = v[1]; s_h[2] = v[2]; s_h[3] = (double)(*A.n_ovl); s_h[4] = er[0]; s_h[5] = v[3];
        s_h[6] = (double)(*A.max_count); s_h[7] = (double)(*A.err_flag);
        for (int i = 0; i < MMC_NSCAL; ++i) A.vec[i] = s_h[i];
        if (A.finish && A.host_out) {
            for (int i = 0; i < MMC_NSCAL; ++i) A.host_out[i] = s_h[i];
            __threadfence_system();
            *reinterpret_cast<volatile unsigned long long *>(A.host_out + MMC_NSCAL) = A.seq;
        }
        *A.done = 0u;                                     // ready for the next evaluation
    }
    if (A.push) {   // k_peer_push, by the block that holds the finished vector
        __syncthreads();
        const PeerArgs &P = A.peer;
        for (int d = 0; d < P.world; ++d) {
            double *dst = P.slot[d] + ((size_t)P.parity * P.world + P.rank) * P.nvec_cap;
            for (int t = tid; t < P.nvec; t += TAIL_THREADS) dst[t] = t < MMC_NSCAL ? s_h[t] : A.vec[t];
        }
        __threadfence_system();
        __syncthreads();
        if (tid < P.world) {
            volatile unsigned long long *f = P.flag[tid] + (size_t)P.parity * P.world + P.rank;
            *f = P.epoch;
        }
        if (A.peer_finish) {
            __syncthreads();
            peer_sum_finish_block(P, A.fin_out, A.fin);
        }
    }
}
