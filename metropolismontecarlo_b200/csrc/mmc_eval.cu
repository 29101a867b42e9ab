// mmc_eval.cu — C ABI of libmmc_b200.so, part 2: the full-system energy (potential(), Ewald/energy.jl:946-1032 and
// :864-943), its sharded forms, the host-array form and the volume move (Ewald/volumeChange.jl:50-147) on top of the
// pair kernels (kernels_pairs_v7.cuh; kernels_pairs.cuh for general topologies), the rho(k) rebuild (kernels_recip.cuh)
// and the peer exchange (kernels_peer.cuh).
#include "mmc_handle.h"
#include "kernels_pairs.cuh"
#include "kernels_pairs_v7.cuh"
#include "kernels_recip.cuh"
#include "kernels_upload.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <immintrin.h>

using namespace mmc_detail;

namespace {

// ---------------------------------------------------------------- full-energy evaluation
struct EvalCtx {
    double f, box, kappa;          // scale factor, box and kappa the energy is evaluated at
    const double *d_cfac;
    int rank, world;
    cudaEvent_t wait_sites = nullptr;   // mmc_potential_host: the sites arrive on the side stream; wait for them before the gather
    bool rhok_external = false;         //                     ... and the ρ(k) partials are produced there, chunk by chunk
    double *per_mol = nullptr;          // mmc_energy_all: [n_mol x 3] per-molecule rows (general kernel, evaluation order)
    int rhok_blocks = 0;                // rhok_external: CTAs whose partials sit in d_rhok_partial ...
    cudaEvent_t rhok_done = nullptr;    //                ... once this event has fired
    int n_windows = 1;                  // mmc_potential_host, one GPU: home-cell windows evaluated as their sites arrive ...
    const cudaEvent_t *win_wait = nullptr;   //                      ... each after this event
    bool partial_state = false;         // domain-decomposed host evaluation: only this rank's slab is resident, so a fall-back to
                                        // an unsharded evaluation is the caller's business (finalize_host returns 2)
};

}  // namespace

namespace mmc_detail {

// out == nullptr: partials only, written from block `block0` on (the caller reduces all blocks later); *nb_out = blocks used
int rhok_launch(mmc_handle *h, const double4 *site, int s_begin, int s_end, double box, double2 *out, cudaStream_t st,
                int block0, int *nb_out, int cap_blocks, const double4 *com, double f)
{
    if (!st) st = h->stream;
    const int n = s_end - s_begin;
    const int nkv = h->S.nkvecs;
    const bool v2 = h->n_kpairs <= 32 && h->S.nk <= 6 && h->use_rhok_v2;
    const bool big = !v2;                       // large k-sets: combos of (32 pairs) x (4 or 5 kz) per warp, k_rhok_big
    const int chunk = v2 ? RHOK2_SITES : RHOKB_SITES;
    // combos of (32 (kx,|ky|) pairs) x (ZT kz values) that hold k-vectors (list built by mmc_ewald_prepare); warps per CTA: four or
    // eight (registers are allocated per four warps), whichever costs less — dead warps in the last CTA row against one more
    // rebuild of the e^{ik·r} tables per row (≈ 1200 warp instructions per 64 sites; a warp spends 64 x (12 + 8·ZT) on them)
    const int ZT = h->k_zt, n_combos = h->n_kcombos;
    auto row_cost = [&](int w) { return (long long)((n_combos + w - 1) / w) * (w * 64LL * (12 + 8 * ZT) + 1200); };
    const int big_warps = row_cost(4) <= row_cost(8) ? 4 : 8;
    const int n_groups = (n_combos + big_warps - 1) / big_warps;
    int g0 = 0, g1 = n_groups;
    if (big && h->rhok_kshard && h->cfg.world > 1 && com == nullptr && s_begin == 0 && s_end == h->S.n_sites) {
        // k-RANGE sharding (north star): this rank sums ALL sites for its share of the combo groups
        g0 = (int)((long long)n_groups * h->cfg.rank / h->cfg.world);
        g1 = (int)((long long)n_groups * (h->cfg.rank + 1) / h->cfg.world);
    }
    const int resident = (big && big_warps == 4) ? 4 : 2;          // CTAs per SM
    const int waves = std::max(1, resident * h->sm_count * std::max(1, h->rhok_split) / (big ? std::max(1, g1 - g0) : 1));
    int per = std::max(2 * chunk, (n + waves - 1) / waves);
    per = (per + chunk - 1) / chunk * chunk;
    const int nb = std::max(1, (n + per - 1) / per);
    if (nb_out) *nb_out = nb;
    const int need = std::max(block0 + nb, cap_blocks);
    if (need > h->rhok_grid_cap) {
        if (block0 > 0) FAIL(MMC_ECUDA, "rho(k) partial buffer too small for a chunked rebuild (internal)");
        dfree(h->d_rhok_partial);
        CK(cudaMalloc(&h->d_rhok_partial, (size_t)need * nkv * sizeof(double2)));
        h->rhok_grid_cap = need;
    }
    double2 *part = h->d_rhok_partial + (size_t)block0 * nkv;
    const int US = h->US > 0 ? h->US : 1;
    if (h->tm.full()) cudaEventRecord(h->tm.ev[2], st);
    if (v2) {
        Rhok2Args R{site, s_begin, s_end, per, nkv, h->n_kpairs, h->d_kpairs, h->d_kindex, box, part, com, f, US};
        switch (h->S.nk) {
            case 1: k_rhok_pairs<1><<<nb, RHOK2_BLOCK, 0, st>>>(R); break;
            case 2: k_rhok_pairs<2><<<nb, RHOK2_BLOCK, 0, st>>>(R); break;
            case 3: k_rhok_pairs<3><<<nb, RHOK2_BLOCK, 0, st>>>(R); break;
            case 4: k_rhok_pairs<4><<<nb, RHOK2_BLOCK, 0, st>>>(R); break;
            case 5: k_rhok_pairs<5><<<nb, RHOK2_BLOCK, 0, st>>>(R); break;
            default: k_rhok_pairs<6><<<nb, RHOK2_BLOCK, 0, st>>>(R); break;
        }
    } else {
        if (h->S.nk > MMC_MAX_NK_FULL) FAIL(MMC_EINVAL, "nk too large for the rebuild kernel (<= 16)");
        // every CTA writes only its own k-vectors: the others of this launch's k-share must read as zero
        if (g1 - g0 < n_groups) CK(cudaMemsetAsync(part, 0, (size_t)nb * nkv * sizeof(double2), st));
        RhokBigArgs R{site, s_begin, s_end, per, h->S.nk, nkv, h->d_kindex, box, part, com, f, US, h->d_kpairs, h->n_kpairs, h->d_kcombos, n_combos, g0};
        const size_t smem = (size_t)RHOKB_SITES * 3 * ((h->S.nk + 1) | 1) * sizeof(double2);
        if (g1 > g0) {
            const dim3 grid(nb, g1 - g0);
            if (ZT == 4) k_rhok_big<4><<<grid, 32 * big_warps, smem, st>>>(R);
            else k_rhok_big<5><<<grid, 32 * big_warps, smem, st>>>(R);
        }
    }
    LAUNCH_CHECK();
    if (h->tm.full()) cudaEventRecord(h->tm.ev[3], st);
    if (out) {
        k_rhok_reduce<<<(nkv + 31) / 32, dim3(32, 32), 0, st>>>(h->d_rhok_partial, nb, nkv, out);
        LAUNCH_CHECK();
    }
    return MMC_OK;
}

}  // namespace mmc_detail

namespace {

// k_pairs_fast instantiations: water (3 sites) x tile {64, 128} x padded polynomial degree
#ifdef MMC_DEV_FEW_DEGS      // development builds: the degrees of configs A, D, E only (full list: 3x shorter compile)
#define MMC_FOR_DEGS(X) X(0) X(16) X(20)
#define MMC_FOR_POS_DEGS(X) X(16) X(20)
#define MMC_FOR_DIRECT_DEGS(X) X(7)
#else
#define MMC_FOR_DEGS(X) X(0) X(8) X(12) X(16) X(20) X(24) X(32) X(44)
#define MMC_FOR_POS_DEGS(X) X(8) X(12) X(16) X(20) X(24) X(32) X(44)
#define MMC_FOR_DIRECT_DEGS(X) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12)
#endif
// (false: no instantiation for this degree — the caller fails loudly instead of returning zeros)
bool launch_pairs_fast(int tile, int deg, int grid, size_t smem, cudaStream_t st, const PairArgs &P)
{
#define X(D)                                                                                   \
    if (deg == D) {                                                                            \
        if (tile == 64) k_pairs_fast<3, 64, D><<<grid, PAIR_BLOCK, smem, st>>>(P);             \
        else k_pairs_fast<3, 128, D><<<grid, PAIR_BLOCK, smem, st>>>(P);                       \
        return true;                                                                           \
    }
    MMC_FOR_DEGS(X)
#undef X
    return false;
}
bool launch_pairs_v7(int deg, bool direct, int grid, cudaStream_t st, const V7Args &A)
{
    if (direct) {
#define X(D) if (deg == D) { k_pairs_v7<D, true><<<grid, V7_BLOCK, V7_SMEM, st>>>(A); return true; }
        MMC_FOR_DIRECT_DEGS(X)
#undef X
    } else {
#define X(D) if (deg == D) { k_pairs_v7<D, false><<<grid, V7_BLOCK, V7_SMEM, st>>>(A); return true; }
        MMC_FOR_POS_DEGS(X)
#undef X
    }
    return false;
}
}  // namespace

void mmc_detail::eval_set_attributes()
{
    cudaFuncSetAttribute(k_pairs<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(k_pairs<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(k_pairs<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(k_pairs<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int big_smem = RHOKB_SITES * 3 * 17 * (int)sizeof(double2);
    cudaFuncSetAttribute(k_rhok_big<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, big_smem);
    cudaFuncSetAttribute(k_rhok_big<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, big_smem);
    cudaFuncSetAttribute(k_rhok_big<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);   // four CTAs x 41 KB
    cudaFuncSetAttribute(k_rhok_big<5>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
#define X(D) cudaFuncSetAttribute(k_pairs_v7<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V7_SMEM);
    MMC_FOR_DIRECT_DEGS(X)
#undef X
#define X(D) cudaFuncSetAttribute(k_pairs_v7<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V7_SMEM);
    MMC_FOR_POS_DEGS(X)
#undef X
#define X(D)                                                                                                   \
    cudaFuncSetAttribute(k_pairs_fast<3, 64, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);      \
    cudaFuncSetAttribute(k_pairs_fast<3, 128, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    MMC_FOR_DEGS(X)
#undef X
}

namespace {

void bind_flags(mmc_handle *h, int ncell)
{
    h->d_maxdev = reinterpret_cast<double *>(h->d_flags);
    h->d_novl = reinterpret_cast<unsigned *>(h->d_flags + 2);
    h->d_errflag = reinterpret_cast<unsigned *>(h->d_flags + 3);
    h->d_maxcount = h->d_flags + 4;
    h->d_count = h->d_flags + 8;
    h->d_fill = h->d_flags + 8 + ncell;
}

// A pair kernel declined the state (level 0 k_pairs_v7: a cell with more than 64 molecules, a molecule reaching outside
// the reference's +100 window, or an overlap — whose whole-row rule the general path implements; level 1 k_pairs_fast: a
// cell that does not fit its tile): next level.
bool escalate_pair_level(mmc_handle *h)
{
    h->pair_level += 1;
    return h->pair_level <= 2;
}

// sharded ρ(k) by k-ranges instead of sites (large k-sets only; mmc_debug_set "rhok_kshard")
bool kshard_on(const mmc_handle *h, const EvalCtx &E)
{
    const bool v2 = h->n_kpairs <= 32 && h->S.nk <= 6 && h->use_rhok_v2;
    return h->rhok_kshard && !v2 && E.world > 1 && !E.rhok_external && E.f == 1.0;
}

int grid_cells(const mmc_handle *h, int style, double box)
{
    const bool want_qq = (style == MMC_STYLE_EWALD || style == MMC_STYLE_WOLF);
    const double rcmax = std::max(h->S.rc_lj, want_qq ? h->S.rc_qq : 0.0);
    int ncd = (int)std::floor(box / rcmax);
    return ncd > 128 ? 128 : ncd;
}

// ---------------------------------------------------------------------------------------------------- k_pairs_v7 path
// Serves: cell mode (box >= 3 r_cut), uniform 3-site molecules with identical per-site charges, LJ on site pair (0,0) only,
// equal cut-offs, Coulomb on (EWALD / WOLF), a usable erf polynomial.
bool v7_eligible(mmc_handle *h, int style, const EvalCtx &E, ErfPoly &ep)
{
    const DevSystem &S = h->S;
    const bool want_qq = (style == MMC_STYLE_EWALD || style == MMC_STYLE_WOLF);
    if (h->pair_level != 0 || E.per_mol || !h->uniform || h->US != 3 || !h->uniform_q || !want_qq) return false;
    if (S.rc_lj != S.rc_qq || h->lj.size() != 1 || h->lj[0].a != 0 || h->lj[0].b != 0) return false;
    if (grid_cells(h, style, E.box) < 3) return false;
    get_erf_poly(h, E.kappa, S.rc_qq * S.rc_qq + 100, ep);
    return ep.deg > 0;
}

V7Grid v7_grid(const mmc_handle *h, int style, const EvalCtx &E)
{
    V7Grid G{};
    G.ncd = grid_cells(h, style, E.box); G.EX = G.ncd + 2; G.EY = G.ncd + 2;
    G.rank = E.rank; G.world = E.world;
    G.range = h->d7_range + (E.world > 1 ? 2 : 0);      // [0, 1]: the whole grid (one rank); [2 ..]: the partition of a sharded evaluation
    G.edge = E.box / G.ncd; G.box_new = E.box;
    return G;
}

int v7_alloc(mmc_handle *h, int ncd, int EXY)
{
    if (!h->d7_flags) {
        CK(cudaMalloc(&h->d7_flags, 16 * sizeof(int)));
        CK(cudaMalloc(&h->d7_block_sums, TAIL_BLOCKS * sizeof(double4)));
        CK(cudaMalloc(&h->d7_range, 32 * sizeof(int)));
    }
    V7Grid G{};
    G.ncd = ncd; G.EX = EXY; G.EY = EXY;
    const long long units = (long long)V3_GROUPS * ncd * ncd * ncd;
    if (G.ncd != h->d7_ncd) {
        dfree(h->d7_count); dfree(h->d7_bucket); dfree(h->d7_ecount); dfree(h->d7_rows); dfree(h->d7_gf);
        const size_t ncell = (size_t)G.ncd * G.ncd * G.ncd, next = (size_t)v7_ext_cells(G);
        CK(cudaMalloc(&h->d7_count, ncell * sizeof(int)));
        CK(cudaMalloc(&h->d7_bucket, ncell * V7_CAP * sizeof(int)));
        CK(cudaMalloc(&h->d7_ecount, next * sizeof(int)));
        CK(cudaMalloc(&h->d7_rows, next * V7_CAP * V7_ROW * sizeof(double)));
        CK(cudaMalloc(&h->d7_gf, next * V7_CAP * sizeof(float4)));
        h->d7_ncd = G.ncd;
        h->bin_version = 0;
        const int whole[2] = {0, (int)ncell};
        CK(cudaMemcpyAsync(h->d7_range, whole, sizeof(whole), cudaMemcpyHostToDevice, h->stream));
    }
    const size_t need = (size_t)std::max(1LL, units) * V7_CONSUMERS;
    if (need > h->d7_partial_cap) {
        dfree(h->d7_unit_partial); dfree(h->d7_order);
        CK(cudaMalloc(&h->d7_unit_partial, need * sizeof(double4)));
        CK(cudaMalloc(&h->d7_order, (size_t)std::max(1LL, units) * sizeof(int)));
        h->d7_partial_cap = need;
        h->bin_version = 0;
    }
    return MMC_OK;
}

// MMC_TRACE_HOST=1: a timeline of mmc_potential_host (timing events on every stream, printed when the call returns)
struct HostTrace {
    bool on = false;
    int n = 0;
    cudaEvent_t ev[48] = {};
    const char *name[48] = {};
    void mark(cudaStream_t st, const char *what)
    {
        if (!on || n >= 48) return;
        if (!ev[n]) cudaEventCreate(&ev[n]);
        cudaEventRecord(ev[n], st);
        name[n++] = what;
    }
    void dump()
    {
        if (!on) return;
        cudaDeviceSynchronize();
        for (int i = 1; i < n; ++i) { float ms = 0; cudaEventElapsedTime(&ms, ev[0], ev[i]); std::fprintf(stderr, "  %8.3f ms  %s\n", ms, name[i]); }
        n = 0;
    }
};
HostTrace g_trace;

// Enqueues one evaluation on the v7 path and leaves this rank's partial-sum vector in d_vec.  finish: one rank — E_recip,
// resident ρ(k) and the scalars in the mapped host slot come out of the same tail kernel.
int eval_v7(mmc_handle *h, int style, const EvalCtx &E, const ErfPoly &ep, double *d_vec, bool finish, double2 *dst0, double2 *dst1,
            const PeerArgs *push = nullptr, const PeerFinishArgs *fin = nullptr)
{
    const DevSystem &S = h->S;
    const bool ewald = style == MMC_STYLE_EWALD;
    int rc = v7_alloc(h, grid_cells(h, style, E.box), grid_cells(h, style, E.box) + 2);
    if (rc) return rc;
    const V7Grid G = v7_grid(h, style, E);
    if (h->tm.full()) cudaEventRecord(h->tm.ev[4], h->stream);
    // ---- ρ(k) rebuild (RecipLong, ewalds.jl:538-604) of this rank's share of the sites: depends on nothing the pair path
    // produces (a volume trial scales the resident sites inside the kernel).  overlap_rhok: 0 = on this stream before the pair
    // path; 1 = on the low-priority side stream, made eligible TOGETHER with the pair kernel (after the gather): the persistent
    // pair CTAs take the SMs first and the rebuild's CTAs (128 registers x 256 threads, they cannot co-reside with four pair
    // CTAs) fill the SMs as the ticket queue drains — the rebuild hides the pair kernel's tail instead of slowing its bulk;
    // 2 = side stream from the start of the evaluation (round 1's placement: both kernels run 20 % slower side by side).
    const long long ns_all = S.n_sites;
    int rs0 = (int)(ns_all * E.rank / E.world), rs1 = (int)(ns_all * (E.rank + 1) / E.world);
    if (kshard_on(h, E)) { rs0 = 0; rs1 = (int)ns_all; }       // k-range sharding: all sites, this rank's k-vectors (rhok_launch)
    int rhok_blocks = E.rhok_blocks;
    bool forked = false;
    const bool rhok_here = ewald && !E.rhok_external;
    const bool rhok_side = rhok_here && h->overlap_rhok && (long long)(rs1 - rs0) * S.nkvecs > 10000000LL;
    // (overlap_rhok == 1 additionally runs the first rhok_early_pct % of the sites at once: the binning and the gather are
    //  latency-bound and leave the SMs nearly idle for ~35 us, which is about that share of the rebuild)
    const bool split = rhok_side && h->overlap_rhok == 1 && h->rhok_early_pct > 0 && !kshard_on(h, E);
    const int rsm = split ? rs0 + (int)((long long)(rs1 - rs0) * h->rhok_early_pct / 100) / 64 * 64 : rs0;
    int blocks_early = 0;
    auto launch_rhok = [&](bool side, int a, int b, int block0, int *nb) -> int {
        if (side) {
            CK(cudaEventRecord(h->ev_fork, h->stream));
            CK(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
        }
        int rcr = rhok_launch(h, S.site, a, b, E.box, nullptr, side ? h->side : h->stream, block0, nb, 4 * h->sm_count * std::max(1, h->rhok_split) + 8,
                              E.f == 1.0 ? nullptr : S.com, E.f);
        if (rcr) return rcr;
        if (side) { CK(cudaEventRecord(h->ev_join, h->side)); forked = true; }
        return MMC_OK;
    };
    if (rhok_here && (!rhok_side || h->overlap_rhok == 2) && (rc = launch_rhok(rhok_side, rs0, rs1, 0, &rhok_blocks))) return rc;
    if (split && rsm > rs0 && (rc = launch_rhok(true, rs0, rsm, 0, &blocks_early))) return rc;
    // ---- binning (fractional COM coordinates do not change with the box: the buckets of an unchanged state are reused,
    // e.g. by consecutive volume trials) and the gather into the extended grid
    CK(cudaMemsetAsync(h->d7_flags, 0, 16 * sizeof(int), h->stream));
    const int tb = 256;
    if (h->bin_version != h->state_version || h->bin_ncd != G.ncd || h->bin_world != E.world) {
        CK(cudaMemsetAsync(h->d7_count, 0, sizeof(int) * (size_t)G.ncd * G.ncd * G.ncd, h->stream));
        Bin7Args B{S.com, S.n_mol, (double)G.ncd / S.box, G, h->d7_count, h->d7_bucket, nullptr};
        k_bin7<<<(S.n_mol + tb - 1) / tb, tb, 0, h->stream>>>(B); LAUNCH_CHECK();
        if (E.world > 1) {       // the ranks' home ranges: equal estimated pair work, from the cell populations
            k_partition7<<<1, 1024, 0, h->stream>>>(h->d7_count, G.ncd, E.world, h->d7_range + 2); LAUNCH_CHECK();
        }
        k_order7<<<1, 1024, 0, h->stream>>>(h->d7_count, G.ncd, G.range, G.rank, h->d7_order); LAUNCH_CHECK();
        h->bin_version = h->state_version; h->bin_ncd = G.ncd; h->bin_world = E.world;
    }
    if (E.wait_sites) CK(cudaStreamWaitEvent(h->stream, E.wait_sites, 0));      // binning needed the COMs only; the gather needs the sites
    unsigned int *fl = reinterpret_cast<unsigned int *>(h->d7_flags);
    V7Args A{};
    {
        const double rcut = S.rc_qq, edge = G.edge;
        // conservative FP32 gate in the dot form |b|² − 2a·b < r_c² − |a|² on coordinates relative to the box centre
        // (components <= M = L/2 + edge, ghosts included): eight roundings of numbers <= 3M² and the input roundings
        // (4·sqrt(3)·r_c·M·2^-24), times four.  Config E: 0.07 Å² on r_c² = 100, i.e. 0.1 % more survivors for the exact test.
        const double M = 0.5 * E.box + edge;
        const double margin = 4.0 * (8.0 * 3.0 * M * M + 7.0 * rcut * M) / 16777216.0;
        A.gate_rc2f = std::nextafterf((float)(rcut * rcut + margin), INFINITY);
        A.rc_qq2 = rcut * rcut;
        std::memcpy(&A.rcqq_bits, &A.rc_qq2, 8);
        A.qq_negmask = 0;
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {
                A.qq_tab[a * 3 + b] = h->q_site[a] * h->q_site[b];
                if (A.qq_tab[a * 3 + b] < 0.0) A.qq_negmask |= 1u << (a * 3 + b);
            }
        A.lj_eps = h->lj[0].eps; A.lj_sig2 = h->lj[0].sig * h->lj[0].sig;
    }
    // −κ folded into the coefficients; DIRECT: also κ^2k, so the kernel runs Horner in r² itself
    const bool direct = ep.ddeg > 0;
    const int deg = direct ? ep.ddeg : ep.deg;
    {
        double k2k = 1.0;
        for (int k = 0; k <= deg; ++k) {
            A.pc[k] = direct ? -E.kappa * ep.a[k] * k2k : -E.kappa * ep.c[k];
            k2k *= E.kappa * E.kappa;
        }
        A.pk2s = ep.kappa2 * ep.scale;
    }
    A.rows = h->d7_rows; A.gf = h->d7_gf; A.ecount = h->d7_ecount;
    A.max_dev = reinterpret_cast<const double *>(h->d7_flags);
    A.n_ovl = fl + 2; A.err_flag = fl + 3;
    A.unit_partial = h->d7_unit_partial; A.order = h->d7_order;
    const int nwin = (E.world == 1 && E.n_windows > 1) ? E.n_windows : 1;
    for (int w = 0; w < nwin; ++w) {
        V7Grid Gw = G;
        if (nwin > 1) {          // a window is a "rank" of the window partition: same kernels, same range mechanism
            Gw.range = h->d7_range + 16; Gw.rank = w; Gw.world = nwin;
            if (E.win_wait) CK(cudaStreamWaitEvent(h->stream, E.win_wait[w], 0));
        }
        Gather7Args Ga{S.com, S.site, h->d7_count, h->d7_bucket, Gw, E.f, h->d7_rows, h->d7_gf, h->d7_ecount,
                       reinterpret_cast<unsigned long long *>(h->d7_flags), fl + 3, h->d7_flags + 4};
        const int warps = G.EX * G.EY * (G.ncd + 1);
        k_gather7<<<(warps + 7) / 8, 256, 0, h->stream>>>(Ga); LAUNCH_CHECK();
        g_trace.mark(h->stream, "window gathered");
        if (w == 0 && rhok_side && h->overlap_rhok == 1) {
            int nb_late = 0;
            if ((rc = launch_rhok(true, rsm, rs1, blocks_early, &nb_late))) return rc;
            rhok_blocks = blocks_early + nb_late;
        }
        if (w == 0 && h->tm.on) { if (h->tm.full()) cudaEventRecord(h->tm.ev[5], h->stream); cudaEventRecord(h->tm.ev[0], h->stream); }
        A.G = Gw; A.ticket = fl + 8 + w;
        const long long units = (long long)V3_GROUPS * G.ncd * G.ncd * G.ncd / (E.world * nwin) + 1;      // (about: the grid size only)
        const int grid = (int)std::max(1LL, std::min<long long>((long long)h->v7_ctas_per_sm * h->sm_count, units));
        if (!launch_pairs_v7(deg, direct, grid, h->stream, A)) FAIL(MMC_ECUDA, "k_pairs_v7: no instantiation for this polynomial degree (internal)");
        LAUNCH_CHECK();
        g_trace.mark(h->stream, "window pairs done");
    }
    if (h->tm.on) cudaEventRecord(h->tm.ev[1], h->stream);
    if (forked) CK(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    if (E.rhok_external && E.rhok_done) CK(cudaStreamWaitEvent(h->stream, E.rhok_done, 0));
    // ---- everything else in one launch
    TailArgs T{};
    T.unit_partial = h->d7_unit_partial; T.range = G.range; T.rank = E.rank;
    T.rhok_partial = h->d_rhok_partial; T.rhok_blocks = rhok_blocks; T.nkvecs = ewald ? S.nkvecs : 0;
    if (ewald && (size_t)S.nkvecs > h->d7_scratch_cap) {
        dfree(h->d7_rhok_scratch);
        CK(cudaMalloc(&h->d7_rhok_scratch, sizeof(double2) * 16 * (size_t)S.nkvecs));
        h->d7_scratch_cap = (size_t)S.nkvecs;
    }
    T.rhok_scratch = h->d7_rhok_scratch;
    T.block_sums = h->d7_block_sums; T.done = fl + 6;
    T.n_ovl = fl + 2; T.err_flag = fl + 3; T.max_count = h->d7_flags + 4;
    T.vec = d_vec;
    T.finish = finish ? 1 : 0; T.cfac = E.d_cfac; T.dst0 = dst0; T.dst1 = dst1;
    T.host_out = h->d7_res; T.seq = finish ? ++h->res_seq : 0;
    T.push = push ? 1 : 0;
    if (push) T.peer = *push;
    T.peer_finish = (push && fin) ? 1 : 0;
    if (push && fin) { T.fin = *fin; T.fin_out = h->d_peer_total; }
    k_eval_tail<<<TAIL_BLOCKS, TAIL_THREADS, 0, h->stream>>>(T); LAUNCH_CHECK();
    if (h->tm.full()) cudaEventRecord(h->tm.ev[6], h->stream);
    h->last_fast = 7; h->last_mode = 0; h->last_ncd = G.ncd;
    return MMC_OK;
}

// wait for the tail kernel of evaluation `res_seq` to publish its scalars in the mapped host slot
int v7_wait(mmc_handle *h, double *hv)
{
    if (h->cfg.sync_mode == 1) CK(cudaStreamSynchronize(h->stream));
    volatile unsigned long long *flag = reinterpret_cast<volatile unsigned long long *>(h->h7_res + MMC_NSCAL);
    unsigned spins = 0;
    while (*flag != h->res_seq) {
        _mm_pause();
        if ((++spins & 0xfffu) == 0) {
            cudaError_t q = cudaStreamQuery(h->stream);
            if (q != cudaSuccess && q != cudaErrorNotReady) { h->err = std::string("evaluation: ") + cudaGetErrorString(q); return MMC_ECUDA; }
            if (q == cudaSuccess && *flag != h->res_seq) FAIL(MMC_ECUDA, "evaluation finished without publishing its result");
        }
    }
    for (int i = 0; i < MMC_NSCAL; ++i) hv[i] = h->h7_res[i];
    return MMC_OK;
}

// ---------------------------------------------------------------------------------------------------- general path
// Leaves this rank's partial sums in d_vec: [0] Σlj_pot [1] Σlj_vir [2] Σcoul [3] #overlap
// [MMC_NSCAL ..) ρ(k) partial (re,im).  Cell lists + k_pairs_fast (3 sites) / k_pairs (any uniform topology), tile pairs
// for boxes below 3 r_cut.
int eval_partials(mmc_handle *h, int style, const EvalCtx &E, double *d_vec)
{
    const bool force_general = h->pair_level >= 2 || E.per_mol != nullptr;
    const DevSystem &S = h->S;
    const int US = h->ES;
    if (US < 1 || US > MMC_MAX_SITES) FAIL(MMC_EINVAL, "pair kernel: 1..16 sites per molecule (internal)");
    const bool want_qq = (style == MMC_STYLE_EWALD || style == MMC_STYLE_WOLF);
    const int ncd = grid_cells(h, style, E.box);
    const bool cells = ncd >= 3;
    CK(cudaMemsetAsync(d_vec, 0, (MMC_NSCAL + 2 * (size_t)std::max(S.nkvecs, 1)) * sizeof(double), h->stream));
    if (h->tm.full()) cudaEventRecord(h->tm.ev[4], h->stream);
    // ρ(k) reads the resident sites, or — for a volume trial — the scaled copy (padded slots of a mixed topology carry q = 0)
    const long long ns_all = (E.f != 1.0 && h->mixed) ? (long long)S.n_mol * US : S.n_sites;
    int rs0 = (int)(ns_all * E.rank / E.world), rs1 = (int)(ns_all * (E.rank + 1) / E.world);
    if (kshard_on(h, E) && E.f == 1.0) { rs0 = 0; rs1 = (int)ns_all; }
    const int tb = 256, gm = (S.n_mol + tb - 1) / tb;
    long long n_units;
    int zl_lo = 0, zl_cnt = 1 << 30;
    if (cells) {
        const int ncell = ncd * ncd * ncd;
        if (ncell > h->ncell_cap) {
            // one block: [flags(8 ints) | count(ncell) | fill(ncell)] so that one memset clears it all
            dfree(h->d_flags); dfree(h->d_start);
            CK(cudaMalloc(&h->d_flags, sizeof(int) * (8 + 2 * (size_t)ncell)));
            CK(cudaMalloc(&h->d_start, sizeof(int) * (ncell + 1)));
            h->ncell_cap = ncell;
        }
        bind_flags(h, ncell);
        CK(cudaMemsetAsync(h->d_flags, 0, sizeof(int) * (8 + 2 * (size_t)ncell), h->stream));
        // fractional COM coordinates are invariant under the volume scaling: bin the resident state.
        // A rank of a sharded evaluation reads the home cells of its unit range and their half-shell neighbours: z-layers
        // [z(first home cell), z(last home cell) + 1]; only those are ordered and gathered.  Units are (cell, slot) tuples,
        // 14 per cell, dealt in contiguous ranges: the first home cell of rank r is floor(ncell·r/world) and the last one is
        // at most ceil(ncell·(r+1)/world) − 1.
        zl_lo = 0; zl_cnt = ncd;
        if (E.world > 1 && E.f == 1.0) {
            const int c0 = (int)((long long)ncell * E.rank / E.world);
            const int c1 = (int)(((long long)ncell * (E.rank + 1) + E.world - 1) / E.world) - 1;
            if (c1 >= c0) {
                zl_lo = c0 / (ncd * ncd);
                zl_cnt = std::min(ncd, c1 / (ncd * ncd) - zl_lo + 2);
            }
        }
        CellArgs C{S.com, S.n_mol, ncd, (double)ncd / S.box, h->d_cell_of, h->d_count, h->d_start,
                   h->d_fill, h->d_perm, h->d_maxcount, zl_lo, zl_cnt};
        k_cell_count<<<gm, tb, 0, h->stream>>>(C); LAUNCH_CHECK();
        k_cell_scan<<<1, 1024, 0, h->stream>>>(C, ncell); LAUNCH_CHECK();
        k_cell_fill<<<gm, tb, 0, h->stream>>>(C); LAUNCH_CHECK();
        k_cell_sort<<<(ncell * 32 + tb - 1) / tb, tb, 0, h->stream>>>(C, ncell); LAUNCH_CHECK();
        n_units = 14LL * ncell;
        if (h->max_cell_cached < 0) {   // unknown density: one synchronous read-back, cached afterwards
            int mc = 0;
            CK(cudaMemcpyAsync(&mc, h->d_maxcount, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            h->max_cell_cached = mc;
        }
    } else {
        if (!h->d_flags) CK(cudaMalloc(&h->d_flags, sizeof(int) * 8));
        bind_flags(h, 0);
        CK(cudaMemsetAsync(h->d_flags, 0, sizeof(int) * 8, h->stream));
        const long long nt = (S.n_mol + PAIR_TILE - 1) / PAIR_TILE;
        n_units = nt * (nt + 1) / 2;
    }
    GatherArgs G{S.com, S.site, cells ? h->d_perm : nullptr, S.n_mol, US, E.f, h->d_scom, h->d_ssite,
                 reinterpret_cast<unsigned long long *>(h->d_maxdev), h->d_ovl, nullptr, nullptr, h->d_cell_of, ncd, E.box / ncd,
                 zl_lo, std::min(zl_cnt, ncd), h->mixed ? S.mol : nullptr, h->mixed ? S.atype : nullptr, h->mixed ? h->d_stype : nullptr};
    if (E.wait_sites) CK(cudaStreamWaitEvent(h->stream, E.wait_sites, 0));      // binning needed the COMs only; the gather needs the sites
    k_gather<<<gm, tb, 0, h->stream>>>(G); LAUNCH_CHECK();
    if (h->tm.full()) cudaEventRecord(h->tm.ev[5], h->stream);

    PairArgs P{};
    P.com = h->d_scom; P.site = h->d_ssite; P.cell_start = h->d_start;
    P.ncd = ncd; P.S = US; P.n_mol = S.n_mol; P.mode = cells ? 0 : 1;
    P.n_tiles = (S.n_mol + PAIR_TILE - 1) / PAIR_TILE;
    P.L = E.box; P.rc_lj2 = S.rc_lj * S.rc_lj; P.rc_qq2 = S.rc_qq * S.rc_qq; P.kappa = E.kappa;
    P.want_lj = 1; P.want_qq = want_qq ? 1 : 0;
    P.nlj = (int)h->lj.size(); P.lj = h->d_lj;
    P.partial = h->d_pair_partial; P.ovl = h->d_ovl; P.n_ovl = h->d_novl; P.max_dev = h->d_maxdev;
    P.err_flag = h->d_errflag;
    P.per_mol = E.per_mol;
    P.stype = h->mixed ? h->d_stype : nullptr;
    P.rclj_bits = 0; P.rcqq_bits = 0; P.cutlj_bits = 0; P.cutqq_bits = 0;
    { double v;
      v = P.rc_lj2; std::memcpy(&P.rclj_bits, &v, 8); v = P.rc_qq2; std::memcpy(&P.rcqq_bits, &v, 8);
      v = P.rc_lj2 + 100; std::memcpy(&P.cutlj_bits, &v, 8); v = P.rc_qq2 + 100; std::memcpy(&P.cutqq_bits, &v, 8); }
    P.ep = ErfPoly{};
    if (want_qq) get_erf_poly(h, E.kappa, S.rc_qq * S.rc_qq + 100, P.ep);
    const int max_cell = cells ? h->max_cell_cached : PAIR_TILE;
    const int tile = (US == 3 && !force_general && !h->mixed) ? (max_cell <= 64 ? 64 : (max_cell <= 128 ? 128 : 0)) : 0;
    P.unit_begin = n_units * E.rank / E.world;
    P.unit_end = n_units * (E.rank + 1) / E.world;
    const long long my_units = P.unit_end - P.unit_begin;
    int grid;
    if (h->tm.on) cudaEventRecord(h->tm.ev[0], h->stream);
    if (tile) {
        if (n_units > h->units_cap) {
            dfree(h->d_units);
            CK(cudaMalloc(&h->d_units, sizeof(int4) * n_units));
            h->units_cap = n_units;
        }
        P.units = h->d_units;
        k_units_build<<<(unsigned)((n_units + 255) / 256), 256, 0, h->stream>>>(P, h->d_units, n_units);
        LAUNCH_CHECK();
        if (h->tm.on) cudaEventRecord(h->tm.ev[0], h->stream);
        const size_t smem = 2 * (2 * (size_t)tile + 2 * (size_t)tile * US) * sizeof(double4) +
                            (size_t)tile * tile * sizeof(unsigned short);
        grid = (int)std::max(1LL, std::min<long long>((tile == 64 ? 4 : 2) * h->sm_count, my_units));
        if (!launch_pairs_fast(tile, P.ep.deg, grid, smem, h->stream, P)) FAIL(MMC_ECUDA, "k_pairs_fast: no instantiation for this polynomial degree (internal)");
    } else {
        const size_t smem = (2 * PAIR_TILE + 2 * PAIR_TILE * (size_t)US) * sizeof(double4) +
                            (size_t)PAIR_WARPS * PAIR_QCAP * sizeof(unsigned) + (h->mixed ? 2 * PAIR_TILE * (size_t)US : 0);
        grid = (int)std::max(1LL, std::min<long long>(2 * h->sm_count, my_units));
        if (h->mixed) {
            if (P.per_mol) k_pairs<0, true><<<grid, PAIR_BLOCK, smem, h->stream>>>(P);
            else k_pairs<0, false><<<grid, PAIR_BLOCK, smem, h->stream>>>(P);
        } else if (P.per_mol) {
            if (US == 3) k_pairs<3, true><<<grid, PAIR_BLOCK, smem, h->stream>>>(P);
            else k_pairs<0, true><<<grid, PAIR_BLOCK, smem, h->stream>>>(P);
        } else if (US == 3) k_pairs<3, false><<<grid, PAIR_BLOCK, smem, h->stream>>>(P);
        else k_pairs<0, false><<<grid, PAIR_BLOCK, smem, h->stream>>>(P);
    }
    LAUNCH_CHECK();
    if (h->tm.on) cudaEventRecord(h->tm.ev[1], h->stream);
    k_pair_reduce<<<1, 256, 0, h->stream>>>(h->d_pair_partial, grid, h->d_novl, h->d_maxcount, h->d_errflag, d_vec);
    LAUNCH_CHECK();
    h->last_fast = tile;
    h->last_mode = cells ? 0 : 1;
    h->last_ncd = ncd;
    if (style == MMC_STYLE_EWALD && !E.rhok_external) {
        // resident sites when the box is unchanged (a sharded rank gathers only its layers), scaled copy otherwise
        int rc = rhok_launch(h, E.f == 1.0 ? S.site : h->d_ssite, rs0, rs1, E.box, reinterpret_cast<double2 *>(d_vec + MMC_NSCAL));
        if (rc) return rc;
    }
    return MMC_OK;
}

// single-molecule Coulomb row on the (scaled, sorted) evaluation copy, overlap pairs skipped
int overlap_row(mmc_handle *h, const EvalCtx &E, int sorted_index, double *row)
{
    ErfPoly poly{};                      // rare path: plain erfc()
    DevSystem V = h->S;
    V.site = h->d_ssite; V.com = h->d_scom; V.mol = h->d_mol_uniform;
    if (h->mixed) {      // the padded copy is a uniform system of ES-slot molecules (padding: q = 0 — contributes nothing, overlaps nothing)
        V.uni = h->ES; V.max_sites = h->ES; V.n_sites = V.n_mol * h->ES; V.atype = h->d_atype_pad;
    }
    V.box = E.box; V.kappa = E.kappa;
    MoveArgs A{};
    A.i = sorted_index; A.n_cfg = 1; A.tiles = move_tiles(h); A.recip_blocks = 0;
    A.want_lj = 0; A.want_qq = 1; A.ignore_overlap = 1; A.cur = h->cur;
    int rc = launch_move_on(h, V, A, poly, false);
    if (rc) return rc;
    *row = h->h_out->qq[0];
    return MMC_OK;
}

// Properties in the reference's order (energy.jl:972-1021) from the eight scalars of a (rank-summed) partial vector:
// hv[0] Σlj_pot, [1] Σlj_vir, [2] Σcoul (overlap rows already removed), [4] E_recip un-scaled, [3] #overlapped molecules
int assemble(mmc_handle *h, int style, const EvalCtx &E, double lj_pot, double lj_vir, double coul, double recip_raw,
             long long novl, mmc_properties *out)
{
    const DevSystem &S = h->S;
    std::memset(out, 0, sizeof(*out));
    const double factor = S.factor;
    out->lj = lj_pot * 4;                       // Σ_i(4 pot_i)/2 over unique pairs
    const double vir_lj = lj_vir * 24 / 3.0;
    out->energy = out->lj;
    out->virial = vir_lj;
    out->overlaps = novl;
    if (style == MMC_STYLE_EWALD || style == MMC_STYLE_WOLF) {
        const double totReal = coul * factor;   // (Σ_i row_i) * factor / 2
        out->real = totReal;
        out->energy += totReal;
        out->coulomb += totReal;
        if (style == MMC_STYLE_EWALD) {
            out->virial += totReal / 3.0;
            const double recipEnergy = recip_raw * factor;
            out->recip = recipEnergy;
            out->energy += recipEnergy;
            out->coulomb += recipEnergy;
            out->virial += recipEnergy / 3.0;
            const double selfEnergy = -E.kappa * h->sum_q2 / std::sqrt(M_PI) * factor;   // ewalds.jl:829-833
            out->self_ = selfEnergy;
            out->energy += selfEnergy;
            out->coulomb += selfEnergy;
            out->virial += selfEnergy / 3.0;
            if (h->intramolecular && !E.partial_state) {      // opt-in, not in the reference (include/mmc_b200.h: mmc_set_intramolecular)
                double raw = 0.0;
                int rc = intra_energy(h, E.kappa, &raw);
                if (rc) return rc;
                out->intra = -raw * factor;
                out->energy += out->intra;
                out->coulomb += out->intra;
                out->virial += out->intra / 3.0;
            }
        } else {
            // energy.jl:924-934 with Σ_iΣ_j q_i q_j = (Σq)² in closed form; r_cut = LJ_rcut (:874)
            const double r_cut = S.rc_lj;
            const double ec = std::erfc(E.kappa * r_cut);
            const double prefactor = -(h->sum_q * h->sum_q) * ec / r_cut;
            const double prefactor2 = (ec / 2 / r_cut + E.kappa / std::sqrt(M_PI)) * h->sum_q2;
            out->wolf_const = (prefactor - prefactor2) * factor;
            out->energy += out->wolf_const;
            out->coulomb += out->wolf_const;
        }
    }
    return MMC_OK;
}

void read_timings(mmc_handle *h, int style)
{
    if (!h->tm.on) return;
    cudaEventElapsedTime(&h->tm.ms[0], h->tm.ev[0], h->tm.ev[1]);
    if (!h->tm.full()) return;
    if (style == MMC_STYLE_EWALD) cudaEventElapsedTime(&h->tm.ms[1], h->tm.ev[2], h->tm.ev[3]);
    cudaEventElapsedTime(&h->tm.ms[2], h->tm.ev[4], h->tm.ev[5]);
    cudaEventElapsedTime(&h->tm.ms[3], h->tm.ev[4], h->tm.ev[6]);
}

// v7, one rank: the tail kernel has done the device part; 1 = the kernel declined the state (caller escalates)
int finish_v7(mmc_handle *h, int style, const EvalCtx &E, mmc_properties *out)
{
    double hv[MMC_NSCAL];
    int rc = v7_wait(h, hv);
    if (rc) return rc;
    if (h->tm.full()) { CK(cudaStreamSynchronize(h->stream)); if (style == MMC_STYLE_EWALD) CK(cudaStreamSynchronize(h->side)); }
    if (hv[7] != 0.0) return 1;
    if (hv[3] != 0.0) { h->v7_left_for_overlap = true; return 1; }     // back to k_pairs_v7 once the overlaps are gone
    h->last_pairs = (long long)hv[5];
    if ((rc = assemble(h, style, E, hv[0], hv[1], hv[2], hv[4], 0, out))) return rc;
    read_timings(h, style);
    h->cnt.full_energy_evals++;
    return MMC_OK;
}

int evaluate_unsharded(mmc_handle *h, int style, EvalCtx E, double2 *dst0, double2 *dst1, mmc_properties *out);

// d_vec holds the (already rank-summed) partials; computes E_recip on the device, brings the
// scalars to the host and assembles Properties in the reference's order (energy.jl:972-1021).
// Returns 1 when the pair kernel that produced the partials declined the state (one rank: the caller escalates).
int finalize_host(mmc_handle *h, int style, const EvalCtx &E, const double *hv, double2 *dst0, double2 *dst1, mmc_properties *out);

int finalize(mmc_handle *h, int style, const EvalCtx &E, double *d_vec, double2 *dst0, double2 *dst1,
             mmc_properties *out)
{
    const DevSystem &S = h->S;
    if (style == MMC_STYLE_EWALD) {
        k_rhok_energy<<<1, RHOKE_BLOCK, 0, h->stream>>>(reinterpret_cast<const double2 *>(d_vec + MMC_NSCAL),
                                                       E.d_cfac, S.nkvecs, dst0, dst1, d_vec + 4);
        LAUNCH_CHECK();
    }
    CK(cudaMemcpyAsync(h->h_vec, d_vec, MMC_NSCAL * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (h->tm.full() && h->last_fast != 7) cudaEventRecord(h->tm.ev[6], h->stream);
    CK(cudaStreamSynchronize(h->stream));
    if (h->tm.full() && style == MMC_STYLE_EWALD) CK(cudaStreamSynchronize(h->side));
    return finalize_host(h, style, E, h->h_vec, dst0, dst1, out);
}

// the host part: hv = the eight scalars of the (rank-summed) vector
int finalize_host(mmc_handle *h, int style, const EvalCtx &E, const double *hv, double2 *dst0, double2 *dst1, mmc_properties *out)
{
    const DevSystem &S = h->S;
    const bool coulomb = style == MMC_STYLE_EWALD || style == MMC_STYLE_WOLF;
    const long long novl = (long long)hv[3];
    if (E.world == 1) {
        if (h->last_mode == 0 && h->last_fast != 7) h->max_cell_cached = (int)hv[6];
        if (hv[7] != 0.0) return 1;      // a cell outgrew the kernel's tile: caller re-runs
        if (h->last_fast == 7 && novl > 0) return 1;
    } else if (hv[7] != 0.0 || (novl > 0 && coulomb)) {
        // Summed over ranks, so every rank takes this branch together: a rank's pair kernel declined the state, or
        // molecules overlap (ewalds.jl:359-360 zeroes whole rows, which needs a molecule's complete neighbourhood).  The
        // state is replicated, so every rank evaluates the whole system on the general path and gets the same Properties:
        // slow, rare, and no retry protocol leaks to the caller.
        if (hv[7] != 0.0 || h->last_fast == 7) { if (!escalate_pair_level(h)) FAIL(MMC_ECUDA, "pair kernel fallback chain exhausted (internal)"); }
        if (novl > 0 && hv[7] == 0.0) h->v7_left_for_overlap = true;
        h->max_cell_cached = -1;
        if (E.partial_state) return 2;
        EvalCtx E1 = E;
        E1.rank = 0; E1.world = 1; E1.wait_sites = nullptr; E1.rhok_external = false; E1.rhok_done = nullptr;
        return evaluate_unsharded(h, style, E1, dst0, dst1, out);
    }
    double lj_pot = hv[0], lj_vir = hv[1], coul = hv[2];
    const double recip_raw = hv[4];
    if (novl > 0 && coulomb) {
        // reference semantics: a molecule whose EwaldReal row hits the overlap rule contributes
        // 0 for its whole row (ewalds.jl:359-360 inside energy.jl:991-1001): U - ½ Σ_flagged row_i
        std::vector<unsigned> fl(S.n_mol);
        CK(cudaMemcpy(fl.data(), h->d_ovl, sizeof(unsigned) * S.n_mol, cudaMemcpyDeviceToHost));
        for (int p = 0; p < S.n_mol; ++p)
            if (fl[p]) {
                double row;
                int rc = overlap_row(h, E, p, &row);
                if (rc) return rc;
                coul -= row / 2;
            }
        h->cnt.overlap_events += novl;
    }
    h->last_pairs = (long long)hv[5];
    if (h->v7_left_for_overlap && novl == 0 && coulomb) { h->v7_left_for_overlap = false; h->pair_level = h->pair_floor; }
    { int rca = assemble(h, style, E, lj_pot, lj_vir, coul, recip_raw, novl, out); if (rca) return rca; }
    read_timings(h, style);
    h->cnt.full_energy_evals++;
    return MMC_OK;
}

// One complete evaluation on this GPU alone: k_pairs_v7 when it serves the system, the general path otherwise; a kernel
// that declines the state hands over to the next level.
int evaluate_unsharded(mmc_handle *h, int style, EvalCtx E, double2 *dst0, double2 *dst1, mmc_properties *out)
{
    for (;;) {
        ErfPoly ep{};
        int rc;
        if (v7_eligible(h, style, E, ep)) {
            if ((rc = eval_v7(h, style, E, ep, h->d_vec, true, dst0, dst1))) return rc;
            rc = finish_v7(h, style, E, out);
        } else {
            if ((rc = eval_partials(h, style, E, h->d_vec))) return rc;
            if (E.rhok_external && style == MMC_STYLE_EWALD) {       // mmc_potential_host: the chunks' ρ(k) partials, folded in CTA order
                if (E.rhok_done) CK(cudaStreamWaitEvent(h->stream, E.rhok_done, 0));
                k_rhok_reduce<<<(h->S.nkvecs + 31) / 32, dim3(32, 32), 0, h->stream>>>(h->d_rhok_partial, E.rhok_blocks, h->S.nkvecs,
                                                                                      reinterpret_cast<double2 *>(h->d_vec + MMC_NSCAL));
                LAUNCH_CHECK();
            }
            rc = finalize(h, style, E, h->d_vec, dst0, dst1, out);
        }
        if (rc != 1) return rc;
        if (!escalate_pair_level(h)) FAIL(MMC_ECUDA, "pair kernel fallback chain exhausted (internal)");
        h->max_cell_cached = -1;
        if (E.n_windows > 1 || E.wait_sites) { CK(cudaStreamSynchronize(h->side)); CK(cudaStreamSynchronize(h->rk)); }
        E.wait_sites = nullptr; E.n_windows = 1; E.win_wait = nullptr;      // the state is on the device by now
    }
}

// this rank's share of a sharded evaluation, vector left in d_vec
int evaluate_partial(mmc_handle *h, int style, const EvalCtx &E, double *d_vec)
{
    ErfPoly ep{};
    if (v7_eligible(h, style, E, ep)) return eval_v7(h, style, E, ep, d_vec, false, nullptr, nullptr);
    return eval_partials(h, style, E, d_vec);
}

// debug (pair_level 3): literal Σ_i rows / 2 through the single-molecule kernel, any topology — N launches
int potential_rows(mmc_handle *h, int style, mmc_properties *out)
{
    const DevSystem &S = h->S;
    const bool want_qq = (style == MMC_STYLE_EWALD || style == MMC_STYLE_WOLF);
    double lj = 0, vir = 0, real = 0;
    long long novl = 0;
    for (int i = 0; i < S.n_mol; ++i) {
        MoveArgs A{};
        A.i = i; A.n_cfg = 1; A.tiles = move_tiles(h); A.want_lj = 1; A.want_qq = want_qq; A.cur = h->cur;
        int rc = launch_move_on(h, h->S, A, h->move_poly, true);
        if (rc) return rc;
        lj += h->h_out->lj_pot[0]; vir += h->h_out->lj_vir[0]; real += h->h_out->qq[0];
        novl += h->h_out->overlap[0];
    }
    std::memset(out, 0, sizeof(*out));
    out->lj = lj / 2; out->energy = lj / 2; out->virial = vir / 2; out->overlaps = novl;
    if (want_qq) {
        const double totReal = real * S.factor / 2;
        out->real = totReal; out->energy += totReal; out->coulomb += totReal;
        if (style == MMC_STYLE_EWALD) {
            out->virial += totReal / 3.0;
            int rc = ensure_vec(h);
            if (rc) return rc;
            rc = rhok_launch(h, S.site, 0, S.n_sites, S.box, reinterpret_cast<double2 *>(h->d_vec + MMC_NSCAL));
            if (rc) return rc;
            k_rhok_energy<<<1, RHOKE_BLOCK, 0, h->stream>>>(reinterpret_cast<const double2 *>(h->d_vec + MMC_NSCAL),
                                                           S.cfac, S.nkvecs, S.rhok[0], S.rhok[1], h->d_vec + 4);
            LAUNCH_CHECK();
            CK(cudaMemcpyAsync(h->h_vec, h->d_vec, MMC_NSCAL * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            const double recipEnergy = h->h_vec[4] * S.factor;
            out->recip = recipEnergy; out->energy += recipEnergy; out->coulomb += recipEnergy;
            out->virial += recipEnergy / 3.0;
            const double selfEnergy = -S.kappa * h->sum_q2 / std::sqrt(M_PI) * S.factor;
            out->self_ = selfEnergy; out->energy += selfEnergy; out->coulomb += selfEnergy;
            out->virial += selfEnergy / 3.0;
            if (h->intramolecular) {
                double raw = 0.0;
                if ((rc = intra_energy(h, S.kappa, &raw))) return rc;
                out->intra = -raw * S.factor; out->energy += out->intra; out->coulomb += out->intra; out->virial += out->intra / 3.0;
            }
            h->new_valid = false;
        } else {
            const double ec = std::erfc(S.kappa * S.rc_lj);
            out->wolf_const = (-(h->sum_q * h->sum_q) * ec / S.rc_lj -
                               (ec / 2 / S.rc_lj + S.kappa / std::sqrt(M_PI)) * h->sum_q2) * S.factor;
            out->energy += out->wolf_const; out->coulomb += out->wolf_const;
        }
    }
    h->cnt.full_energy_evals++;
    return MMC_OK;
}

}  // namespace

extern "C" {

int mmc_recip_long(mmc_handle *h, double *energy)
{
    if (!h) return MMC_EINVAL;
    if (!h->has_system || !h->has_ewald) FAIL(MMC_ESTATE, "system and Ewald tables required");
    { int rcf = flush_pending(h); if (rcf) return rcf; }
    int rc = rhok_launch(h, h->S.site, 0, h->S.n_sites, h->S.box, reinterpret_cast<double2 *>(h->d_vec + MMC_NSCAL));
    if (rc) return rc;
    k_rhok_energy<<<1, RHOKE_BLOCK, 0, h->stream>>>(reinterpret_cast<const double2 *>(h->d_vec + MMC_NSCAL),
                                                   h->S.cfac, h->S.nkvecs, h->S.rhok[0], h->S.rhok[1], h->d_vec + 4);
    LAUNCH_CHECK();
    CK(cudaMemcpyAsync(h->h_vec, h->d_vec, MMC_NSCAL * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->new_valid = false;
    if (energy) *energy = h->h_vec[4];
    return MMC_OK;
}

int mmc_partial_count(mmc_handle *h, int64_t *n_doubles)
{
    if (!h || !n_doubles) return MMC_EINVAL;
    *n_doubles = MMC_NSCAL + 2 * (int64_t)std::max(h->S.nkvecs, 1);
    return MMC_OK;
}

int mmc_potential_partial(mmc_handle *h, int32_t style, double *d_partials)
{
    if (!h) return MMC_EINVAL;
    int rc = style_check(h, style);
    if (rc) return rc;
    { int rcf = flush_pending(h); if (rcf) return rcf; }
    if (style == MMC_STYLE_LJ_ATOMS || !d_partials) FAIL(MMC_EINVAL, "sharded evaluation is for molecular systems");
    if ((rc = ensure_vec(h))) return rc;
    EvalCtx E{1.0, h->S.box, h->S.kappa, h->S.cfac, h->cfg.rank, h->cfg.world};
    return evaluate_partial(h, style, E, d_partials);
}

int mmc_potential_finalize(mmc_handle *h, int32_t style, const double *d_partials, mmc_properties *out)
{
    if (!h) return MMC_EINVAL;
    int rc = style_check(h, style);
    if (rc) return rc;
    if (!d_partials || !out) FAIL(MMC_EINVAL, "null argument");
    EvalCtx E{1.0, h->S.box, h->S.kappa, h->S.cfac, h->cfg.rank, h->cfg.world};
    rc = finalize(h, style, E, const_cast<double *>(d_partials), h->S.rhok[0], h->S.rhok[1], out);
    if (rc == 1) {       // (world == 1) the pair kernel declined the state: next level, evaluated here
        if (!escalate_pair_level(h)) FAIL(MMC_ECUDA, "pair kernel fallback chain exhausted (internal)");
        rc = evaluate_unsharded(h, style, E, h->S.rhok[0], h->S.rhok[1], out);
    }
    if (style == MMC_STYLE_EWALD) h->new_valid = false;
    return rc;
}

// ---- sharded evaluation with the exchange over NVLink peer memory (kernels_peer.cuh) ----------------------
// exchange buffer (doubles): [2][world][nvec_cap] vector slots | [2][world] epoch flags | header: COM staging capacity (molecules)
// | [world] COM-slice flags | [3 x capacity] COM staging area (raw layout of moa.COM)
static size_t peer_flag_offset_doubles(const mmc_handle *h) { return 2 * (size_t)h->cfg.world * h->peer_nvec_cap; }
static size_t peer_hdr_offset_doubles(const mmc_handle *h) { return peer_flag_offset_doubles(h) + 2 * (size_t)h->cfg.world; }
static size_t peer_comflag_offset_doubles(const mmc_handle *h) { return peer_hdr_offset_doubles(h) + 2; }    // (+2: the staging area stays 16-byte aligned)
static size_t peer_stage_offset_doubles(const mmc_handle *h) { return peer_comflag_offset_doubles(h) + MMC_PEER_MAX; }

int mmc_peer_export(mmc_handle *h, void *handle64)
{
    if (!h || !handle64) return MMC_EINVAL;
    if (h->cfg.world < 1 || h->cfg.world > MMC_PEER_MAX) FAIL(MMC_EINVAL, "peer exchange supports up to 8 ranks");
    CK(cudaSetDevice(h->cfg.device));
    if (!h->d_peer_buf) {
        h->peer_nvec_cap = MMC_NSCAL + 2 * 16384;      // (nk = 16, k² < 257: 8.6 k k-vectors)
        // a system uploaded before the export sizes the COM staging area (the all-gather of mmc_potential_host, kernels_peer.cuh)
        const unsigned long long stage_cap = h->has_system ? (unsigned long long)h->S.n_mol : 0ull;
        const size_t doubles = peer_stage_offset_doubles(h) + 3 * (size_t)stage_cap;
        CK(cudaMalloc(&h->d_peer_buf, doubles * sizeof(double)));
        CK(cudaMemset(h->d_peer_buf, 0, doubles * sizeof(double)));
        CK(cudaMemcpy(h->d_peer_buf + peer_hdr_offset_doubles(h), &stage_cap, sizeof(stage_cap), cudaMemcpyHostToDevice));
        h->peer_stage_cap = stage_cap;
        CK(cudaMalloc(&h->d_peer_total, h->peer_nvec_cap * sizeof(double)));
        CK(cudaHostAlloc((void **)&h->h_peer_status, sizeof(int), cudaHostAllocMapped));
        *h->h_peer_status = 0;
        CK(cudaHostGetDevicePointer((void **)&h->d_peer_status, h->h_peer_status, 0));
        h->peer_base[h->cfg.rank] = h->d_peer_buf;
        h->peer_ready = 1;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t ih;
    CK(cudaIpcGetMemHandle(&ih, h->d_peer_buf));
    std::memcpy(handle64, &ih, 64);
    return MMC_OK;
}

int mmc_peer_import(mmc_handle *h, int32_t peer_rank, const void *handle64)
{
    if (!h || !handle64) return MMC_EINVAL;
    if (!h->d_peer_buf) FAIL(MMC_ESTATE, "mmc_peer_export first");
    if (peer_rank < 0 || peer_rank >= h->cfg.world) FAIL(MMC_EINVAL, "peer rank out of range");
    if (peer_rank == h->cfg.rank || h->peer_base[peer_rank]) return MMC_OK;
    CK(cudaSetDevice(h->cfg.device));
    cudaIpcMemHandle_t ih;
    std::memcpy(&ih, handle64, 64);
    void *p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess));
    h->peer_base[peer_rank] = p; h->peer_opened[peer_rank] = true;
    h->peer_ready += 1;
    {   // every rank ends with the same staging capacity: the smallest one
        unsigned long long cap = 0;
        CK(cudaMemcpy(&cap, reinterpret_cast<double *>(p) + peer_hdr_offset_doubles(h), sizeof(cap), cudaMemcpyDeviceToHost));
        h->peer_stage_cap = std::min(h->peer_stage_cap, cap);
    }
    return MMC_OK;
}

// same-process form (emulated ranks in one process, tests): the peer's buffer by device pointer
int mmc_peer_import_ptr(mmc_handle *h, int32_t peer_rank, void *peer_buffer)
{
    if (!h || !peer_buffer) return MMC_EINVAL;
    if (!h->d_peer_buf) FAIL(MMC_ESTATE, "mmc_peer_export first");
    if (peer_rank < 0 || peer_rank >= h->cfg.world) FAIL(MMC_EINVAL, "peer rank out of range");
    if (peer_rank == h->cfg.rank || h->peer_base[peer_rank]) return MMC_OK;
    h->peer_base[peer_rank] = peer_buffer;
    h->peer_ready += 1;
    h->peer_same_process = true;      // emulated ranks: streams of ONE context — no kernel may spin on a peer early in the call
    {
        unsigned long long cap = 0;
        CK(cudaMemcpy(&cap, reinterpret_cast<double *>(peer_buffer) + peer_hdr_offset_doubles(h), sizeof(cap), cudaMemcpyDeviceToHost));
        h->peer_stage_cap = std::min(h->peer_stage_cap, cap);
    }
    return MMC_OK;
}

int mmc_peer_buffer(mmc_handle *h, void **buffer)
{
    if (!h || !buffer) return MMC_EINVAL;
    if (!h->d_peer_buf) FAIL(MMC_ESTATE, "mmc_peer_export first");
    *buffer = h->d_peer_buf;
    return MMC_OK;
}

static PeerArgs peer_args(mmc_handle *h)
{
    PeerArgs P{};
    const size_t fo = peer_flag_offset_doubles(h);
    for (int q = 0; q < h->cfg.world; ++q) {
        P.slot[q] = reinterpret_cast<double *>(h->peer_base[q]);
        P.flag[q] = reinterpret_cast<unsigned long long *>(reinterpret_cast<double *>(h->peer_base[q]) + fo);
    }
    P.world = h->cfg.world; P.rank = h->cfg.rank;
    P.nvec = (int)(MMC_NSCAL + 2 * (size_t)std::max(h->S.nkvecs, 1)); P.nvec_cap = (int)h->peer_nvec_cap;
    P.epoch = h->peer_epoch; P.parity = (int)(h->peer_epoch & 1);
    return P;
}

}  // extern "C"

namespace {
// mmc_potential_host on a sharded handle: DOMAIN DECOMPOSITION.  Rank r owns the z-slab [z0, z1) of the cell grid; it
// receives all COMs (24 B per molecule: they decide who needs what), bins them, and then copies from the caller's site
// array only the blocks of 256 molecules that hold a molecule of its slab, of the layer above it (the half shell) or of its
// share of the sites for rho(k) — (1/world + 1/ncd) of the 72 B per molecule instead of all of it when the caller's order is
// spatially coherent (a lattice start, a sorted restart), everything when it is not.  The partial vectors are exchanged over
// NVLink as in mmc_potential_sharded.  Afterwards only those blocks are current on this GPU (partial_resident).
int potential_host_sharded(mmc_handle *h, const double *coords, const double *com, int style, const ErfPoly &ep, mmc_properties *out)
{
    DevSystem &S = h->S;
    int rc;
    if ((rc = ensure_vec(h))) return rc;
    h->pair_level = h->pair_floor;
    if (h->pend_kind == 1) h->pend_kind = 0;
    h->trial_pending = false; h->vol_pending = false; h->new_valid = false;
    double *d_coords = reinterpret_cast<double *>(h->d_raw);
    double *d_com = d_coords + 4 * (size_t)S.n_sites;
    const bool ewald = style == MMC_STYLE_EWALD;
    const int US = h->US;
    const int nblk = (S.n_mol + 255) >> 8;
    if (nblk > h->need_cap) {
        dfree(h->d7_need);
        if (h->h7_need) cudaFreeHost(h->h7_need);
        CK(cudaMalloc(&h->d7_need, (nblk + 3) / 4 * 4));
        CK(cudaHostAlloc((void **)&h->h7_need, nblk, cudaHostAllocMapped));
        CK(cudaHostGetDevicePointer((void **)&h->d7_need_host, h->h7_need, 0));
        h->need_cap = nblk;
    }
    EvalCtx E{1.0, S.box, S.kappa, S.cfac, h->cfg.rank, h->cfg.world};
    E.partial_state = true;
    if ((rc = v7_alloc(h, grid_cells(h, style, E.box), grid_cells(h, style, E.box) + 2))) return rc;
    const V7Grid G = v7_grid(h, style, E);
    // ---- COMs, binning, and which molecule blocks this rank reads
    g_trace.on = std::getenv("MMC_TRACE_HOST") != nullptr && h->cfg.rank == 0;
    g_trace.mark(h->stream, "start");
    // all ranks hold the same system and the same (smallest) staging capacity, so they take the same branch
    const bool com_gather = h->com_allgather && !h->peer_same_process && h->peer_stage_cap >= (unsigned long long)S.n_mol && S.n_mol >= E.world;
    ComGatherArgs CG{};
    if (com_gather) {      // 1/world of the COMs over this rank's PCIe link, the rest over NVLink (k_repack_com_gather below)
        const int m0 = com_slice_begin(S.n_mol, E.world, E.rank), m1 = com_slice_begin(S.n_mol, E.world, E.rank + 1);
        double *stage = h->d_peer_buf + peer_stage_offset_doubles(h);
        CK(cudaMemcpyAsync(stage + 3 * (size_t)m0, com + 3 * (size_t)m0, sizeof(double) * 3 * (size_t)(m1 - m0), cudaMemcpyHostToDevice, h->stream));
        for (int q = 0; q < E.world; ++q) {
            CG.stage[q] = reinterpret_cast<double *>(h->peer_base[q]) + peer_stage_offset_doubles(h);
            CG.flag[q] = reinterpret_cast<unsigned long long *>(reinterpret_cast<double *>(h->peer_base[q]) + peer_comflag_offset_doubles(h));
        }
        CG.epoch = ++h->com_epoch; CG.world = E.world; CG.rank = E.rank; CG.n_mol = S.n_mol; CG.box = S.box; CG.dcom = S.com; CG.info = h->d_info;
        CG.zero_words = reinterpret_cast<unsigned int *>(h->d7_need); CG.n_zero_words = (nblk + 3) / 4;
        CG.zero_count = h->d7_count; CG.n_zero_count = G.ncd * G.ncd * G.ncd;
        CK(cudaEventRecord(h->ev_copy[1], h->stream));
        k_com_publish<<<1, 32, 0, h->stream>>>(CG); LAUNCH_CHECK();       // the peers wait for this: out first (the host is the
    } else {                                                               // bottleneck of this call's first 100 us)
        CK(cudaMemcpyAsync(d_com, com, sizeof(double) * 3 * S.n_mol, cudaMemcpyHostToDevice, h->stream));
        CK(cudaEventRecord(h->ev_copy[1], h->stream));
    }
    g_trace.mark(h->stream, "COM copy done");
    // copies the runs of molecule blocks that `want` selects (the copy stream carries nothing but copies); returns the bytes
    auto copy_runs = [&](auto want) -> long long {
        long long nbytes = 0;
        for (int b = 0; b < nblk;) {
            if (!want(b)) { ++b; continue; }
            int e2 = b;
            while (e2 < nblk && want(e2)) ++e2;
            const long long s0 = (long long)b * 256 * US, s1 = std::min<long long>((long long)e2 * 256 * US, S.n_sites);
            if (cudaMemcpyAsync(d_coords + 3 * s0, coords + 3 * s0, sizeof(double) * 3 * (size_t)(s1 - s0), cudaMemcpyHostToDevice, h->copy) != cudaSuccess) return -1;
            nbytes += (long long)sizeof(double) * 3 * (s1 - s0);
            b = e2;
        }
        return nbytes;
    };
    // Which blocks this rank reads is known only after the binning — but it changes slowly from one evaluation to the next (a
    // Monte Carlo move displaces one molecule by a fraction of an Å), so the blocks the PREVIOUS call needed start crossing the
    // bus right behind the COMs, while the binning kernels run; whatever turns out to be missing afterwards is copied then.
    const bool spec = (int)h->need_prev.size() == nblk && h->need_prev_world == E.world && h->dd_speculate;
    long long bytes = com_gather ? (long long)sizeof(double) * 3 * (com_slice_begin(S.n_mol, E.world, E.rank + 1) - com_slice_begin(S.n_mol, E.world, E.rank))
                                 : (long long)sizeof(double) * 3 * S.n_mol;
    // (host order = priority order: the gather first — two launches — then the speculative copies, then the binning kernels)
    if (com_gather) {
        k_com_wait<<<1, 256, 0, h->stream>>>(CG); LAUNCH_CHECK();           // (+ clears the validation word and the need flags)
        k_repack_com_gather<<<(S.n_mol + 255) / 256, 256, 0, h->stream>>>(CG); LAUNCH_CHECK();     // (+ clears the cell populations)
        g_trace.mark(h->stream, "COMs gathered over NVLink");
    }
    if (spec) {
        CK(cudaStreamWaitEvent(h->copy, h->ev_copy[1], 0));       // behind the COMs, not beside them: the binning waits for those
        const long long nb = copy_runs([&](int b) { return h->need_prev[b] != 0; });
        if (nb < 0) FAIL(MMC_ECUDA, "cudaMemcpyAsync (site blocks)");
        bytes += nb;
        g_trace.mark(h->copy, "site blocks of the previous call's slab copied");
    }
    if (!com_gather) {
        CK(cudaMemsetAsync(h->d_info, 0, 4 * sizeof(int), h->stream));
        k_repack_com<<<(S.n_mol + 255) / 256, 256, 0, h->stream>>>(d_com, S.n_mol, S.box, S.com, h->d_info); LAUNCH_CHECK();
        CK(cudaMemsetAsync(h->d7_need, 0, nblk, h->stream));
        CK(cudaMemsetAsync(h->d7_count, 0, sizeof(int) * (size_t)G.ncd * G.ncd * G.ncd, h->stream));
    }
    h->state_version++;
    Bin7Args B{S.com, S.n_mol, (double)G.ncd / S.box, G, h->d7_count, h->d7_bucket, h->d_cell_of};
    k_bin7<<<(S.n_mol + 255) / 256, 256, 0, h->stream>>>(B); LAUNCH_CHECK();
    k_partition_order7<<<1, 1024, 0, h->stream>>>(h->d7_count, G.ncd, E.world, h->d7_range + 2, G.rank, h->d7_order); LAUNCH_CHECK();
    k_need7<<<(S.n_mol + 255) / 256, 256, 0, h->stream>>>(h->d_cell_of, S.n_mol, G, h->d7_need); LAUNCH_CHECK();
    h->bin_version = h->state_version; h->bin_ncd = G.ncd; h->bin_world = E.world;
    k_bytes_to_host<<<1, 256, 0, h->stream>>>(h->d7_need, nblk, h->d_info, 4, h->d7_need_host, h->d_up->info); LAUNCH_CHECK();
    CK(cudaEventRecord(h->ev_fork, h->stream));
    g_trace.mark(h->stream, "binned, partitioned, needs on the host");
    CK(cudaStreamSynchronize(h->stream));
    if (h->h_up->info[0] & REPACK_PEER_TIMEOUT) FAIL(MMC_ENCCL, "COM all-gather: a rank's slice did not arrive");
    if (h->h_up->info[0] & REPACK_COM_OUTSIDE) FAIL(MMC_EINVAL, "a COM lies outside [0, box] (the reference's PBC keeps COMs inside)");
    // ---- this rank's share of the sites for rho(k) (by site index, as in the resident sharded evaluation)
    const long long ns_all = S.n_sites;
    const int rs0 = (int)(ns_all * E.rank / E.world), rs1 = (int)(ns_all * (E.rank + 1) / E.world);
    if (ewald)
        for (int b = (rs0 / US) >> 8; b <= (((rs1 - 1) / US) >> 8) && b < nblk; ++b) h->h7_need[b] = 1;
    // ---- copy the runs of needed blocks (those the speculative copy did not bring)
    CK(cudaStreamWaitEvent(h->copy, h->ev_fork, 0));
    {
        const long long nb = spec ? copy_runs([&](int b) { return h->h7_need[b] && !h->need_prev[b]; }) : copy_runs([&](int b) { return h->h7_need[b] != 0; });
        if (nb < 0) FAIL(MMC_ECUDA, "cudaMemcpyAsync (site blocks)");
        bytes += nb;
    }
    if (spec) {      // next call's guess = this call's needs; the repack below takes every block that was copied
        for (int b = 0; b < nblk; ++b) { const unsigned char now = h->h7_need[b]; h->h7_need[b] = now | h->need_prev[b]; h->need_prev[b] = now; }
    } else {
        h->need_prev.assign(h->h7_need, h->h7_need + nblk);
        h->need_prev_world = E.world;
    }
    h->last_h2d_bytes = bytes;
    CK(cudaMemcpyAsync(h->d7_need, h->h7_need, nblk, cudaMemcpyHostToDevice, h->copy));      // (with the rho(k) blocks added)
    CK(cudaEventRecord(h->ev_copy[0], h->copy));
    g_trace.mark(h->copy, "site blocks copied");
    CK(cudaStreamWaitEvent(h->side, h->ev_copy[0], 0));
    k_repack_sites_blocks<<<(S.n_sites + 255) / 256, 256, 0, h->side>>>(d_coords, h->d7_need, US, S.n_sites, S.site); LAUNCH_CHECK();
    CK(cudaEventRecord(h->ev_sites, h->side));
    int blocks = 0;
    if (ewald && (rc = rhok_launch(h, S.site, rs0, rs1, S.box, nullptr, h->side, 0, &blocks))) return rc;
    CK(cudaEventRecord(h->ev_join, h->side));
    g_trace.mark(h->side, "rho(k) partials done");
    h->partial_resident = true;
    E.wait_sites = h->ev_sites; E.rhok_external = true; E.rhok_blocks = blocks; E.rhok_done = h->ev_join;
    h->peer_epoch += 1;
    const PeerArgs P = peer_args(h);
    const PeerFinishArgs F{ewald ? S.nkvecs : 0, E.d_cfac, S.rhok[0], S.rhok[1], h->d7_res, ++h->res_seq};
    if ((rc = eval_v7(h, style, E, ep, h->d_vec, false, nullptr, nullptr, &P, &F))) return rc;      // push + sum + finish in the tail
    double hv[MMC_NSCAL];
    if ((rc = v7_wait(h, hv))) return rc;
    g_trace.mark(h->stream, "tail: pushed, peers summed, result published");
    g_trace.dump();
    if (h->h7_res[MMC_NSCAL + 1] != 0.0) FAIL(MMC_ENCCL, "peer exchange: a rank's partial sums did not arrive");
    rc = finalize_host(h, style, E, hv, S.rhok[0], S.rhok[1], out);
    if (rc == 2) {     // declined / overlapping molecules (every rank takes this branch): the whole state, one GPU, general path
        if ((rc = mmc_upload_positions(h, coords, com))) return rc;
        EvalCtx E1{1.0, S.box, S.kappa, S.cfac, 0, 1};
        rc = evaluate_unsharded(h, style, E1, S.rhok[0], S.rhok[1], out);
    }
    return rc;
}

}  // namespace

extern "C" {

// this rank's partial sums, pushed into every rank's exchange buffer (asynchronous: returns after the launches)
int mmc_potential_sharded_begin(mmc_handle *h, int32_t style)
{
    if (!h) return MMC_EINVAL;
    int rc = style_check(h, style);
    if (rc) return rc;
    if (style == MMC_STYLE_LJ_ATOMS) FAIL(MMC_EINVAL, "sharded evaluation is for molecular systems");
    if (h->peer_ready != h->cfg.world) FAIL(MMC_ESTATE, "peer exchange not set up: mmc_peer_export / mmc_peer_import for every rank");
    if (h->sharded_pending) FAIL(MMC_ESTATE, "mmc_potential_sharded_end has not been called");
    if ((size_t)(MMC_NSCAL + 2 * std::max(h->S.nkvecs, 1)) > h->peer_nvec_cap) FAIL(MMC_EINVAL, "too many k-vectors for the exchange buffer");
    { int rcf = flush_pending(h); if (rcf) return rcf; }
    if ((rc = ensure_vec(h))) return rc;
    EvalCtx E{1.0, h->S.box, h->S.kappa, h->S.cfac, h->cfg.rank, h->cfg.world};
    h->peer_epoch += 1;
    const PeerArgs P = peer_args(h);
    ErfPoly ep{};
    h->sharded_fused = false;
    if (v7_eligible(h, style, E, ep)) {          // the tail kernel of the evaluation pushes the vector itself, waits for the peers' and finishes
        const PeerFinishArgs F{style == MMC_STYLE_EWALD ? h->S.nkvecs : 0, E.d_cfac, h->S.rhok[0], h->S.rhok[1], h->d7_res, ++h->res_seq};
        if ((rc = eval_v7(h, style, E, ep, h->d_vec, false, nullptr, nullptr, &P, &F))) return rc;
        h->sharded_fused = true;
    } else {
        if ((rc = eval_partials(h, style, E, h->d_vec))) return rc;
        k_peer_push<<<h->cfg.world, 256, 0, h->stream>>>(P, h->d_vec);
        LAUNCH_CHECK();
    }
    h->sharded_pending = true; h->sharded_style = style;
    return MMC_OK;
}

// wait for every rank's push, add the slots in rank order, finalise (one launch; the scalars arrive in the mapped host slot)
int mmc_potential_sharded_end(mmc_handle *h, mmc_properties *out)
{
    if (!h || !out) return MMC_EINVAL;
    if (!h->sharded_pending) FAIL(MMC_ESTATE, "mmc_potential_sharded_begin first");
    h->sharded_pending = false;
    const int style = h->sharded_style;
    const PeerArgs P = peer_args(h);
    EvalCtx E{1.0, h->S.box, h->S.kappa, h->S.cfac, h->cfg.rank, h->cfg.world};
    if (!h->sharded_fused) {
        PeerFinishArgs F{style == MMC_STYLE_EWALD ? h->S.nkvecs : 0, E.d_cfac, h->S.rhok[0], h->S.rhok[1], h->d7_res, ++h->res_seq};
        k_peer_sum_finish<<<1, 256, 0, h->stream>>>(P, h->d_peer_total, F);
        LAUNCH_CHECK();
    }
    if (h->tm.full() && h->last_fast != 7) cudaEventRecord(h->tm.ev[6], h->stream);
    double hv[MMC_NSCAL];
    int rc = v7_wait(h, hv);
    if (rc) return rc;
    if (h->tm.full()) { CK(cudaStreamSynchronize(h->stream)); if (style == MMC_STYLE_EWALD) CK(cudaStreamSynchronize(h->side)); }
    if (h->h7_res[MMC_NSCAL + 1] != 0.0) FAIL(MMC_ENCCL, "peer exchange: a rank's partial sums did not arrive");
    rc = finalize_host(h, style, E, hv, h->S.rhok[0], h->S.rhok[1], out);
    if (style == MMC_STYLE_EWALD) h->new_valid = false;
    return rc;
}

// all ranks call this together: begin + end
int mmc_potential_sharded(mmc_handle *h, int32_t style, mmc_properties *out)
{
    int rc = mmc_potential_sharded_begin(h, style);
    if (rc) return rc;
    return mmc_potential_sharded_end(h, out);
}

int mmc_potential(mmc_handle *h, int32_t style, mmc_properties *out)
{
    if (!h) return MMC_EINVAL;
    int rc = style_check(h, style);
    if (rc) return rc;
    { int rcf = flush_pending(h); if (rcf) return rcf; }
    if (!out) FAIL(MMC_EINVAL, "null argument");
    if (style == MMC_STYLE_LJ_ATOMS) {
        k_atoms_rows<<<std::min(h->At.n, 8 * h->sm_count), 256, 0, h->stream>>>(h->At, h->d_rows); LAUNCH_CHECK();
        k_rows_sum<<<1, 256, 0, h->stream>>>(h->d_rows, h->At.n, h->d_atoms_out); LAUNCH_CHECK();
        double r[2];
        CK(cudaMemcpyAsync(r, h->d_atoms_out, sizeof(r), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        std::memset(out, 0, sizeof(*out));
        out->energy = r[0]; out->lj = r[0]; out->virial = r[1];
        h->cnt.full_energy_evals++;
        return MMC_OK;
    }
    if (h->pair_level >= 3) return potential_rows(h, style, out);        // debug: the literal Σ_i rows / 2 through k_move
    if ((rc = ensure_vec(h))) return rc;
    EvalCtx E{1.0, h->S.box, h->S.kappa, h->S.cfac, 0, 1};
    rc = evaluate_unsharded(h, style, E, h->S.rhok[0], h->S.rhok[1], out);
    if (style == MMC_STYLE_EWALD) h->new_valid = false;
    return rc;
}

// a2 + a3/a4 for EVERY molecule in one evaluation: what Σ_i in potential() iterates over (energy.jl:966-1001), kept per i.
// One pass over the unique in-cutoff pairs (cell lists), each pair credited to both molecules.
int mmc_energy_all(mmc_handle *h, int32_t style, double *lj_pot, double *lj_vir, double *coul, int32_t *overlap)
{
    if (!h) return MMC_EINVAL;
    int rc = style_check(h, style);
    if (rc) return rc;
    if (style == MMC_STYLE_LJ_ATOMS) FAIL(MMC_EINVAL, "mmc_energy_all is for molecular systems");
    { int rcf = flush_pending(h); if (rcf) return rcf; }
    if ((rc = ensure_vec(h))) return rc;
    const DevSystem &S = h->S;
    const size_t n = (size_t)S.n_mol;
    if (!h->d_permol) {
        CK(cudaMalloc(&h->d_permol, sizeof(double) * 3 * n));
        CK(cudaMalloc(&h->d_permol_out, sizeof(double) * 3 * n + sizeof(int) * n));
    }
    CK(cudaMemsetAsync(h->d_permol, 0, sizeof(double) * 3 * n, h->stream));
    EvalCtx E{1.0, S.box, S.kappa, S.cfac, 0, 1};
    E.rhok_external = true;          // pair part only: ρ(k) is not a per-molecule quantity
    E.per_mol = h->d_permol;
    if ((rc = eval_partials(h, style, E, h->d_vec))) return rc;
    const bool want_qq = (style == MMC_STYLE_EWALD || style == MMC_STYLE_WOLF);
    double *o = h->d_permol_out;
    PerMolArgs A{h->d_permol, h->last_mode == 0 ? h->d_perm : nullptr, h->d_ovl, S.n_mol, want_qq ? 1 : 0, S.factor,
                 o, o + n, o + 2 * n, reinterpret_cast<int *>(o + 3 * n)};
    k_permol_scatter<<<(S.n_mol + 255) / 256, 256, 0, h->stream>>>(A); LAUNCH_CHECK();
    if (lj_pot) CK(cudaMemcpyAsync(lj_pot, o, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
    if (lj_vir) CK(cudaMemcpyAsync(lj_vir, o + n, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
    if (coul) CK(cudaMemcpyAsync(coul, o + 2 * n, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
    if (overlap) CK(cudaMemcpyAsync(overlap, o + 3 * n, sizeof(int) * n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return MMC_OK;
}

// End to end in one call: positions from HOST arrays (pointer(soa.coords), pointer(moa.COM)) → Properties on the host, with the
// copies overlapped with the work that does not need them yet.  COMs go first (the cell binning needs nothing else); the sites
// follow in chunks on the side stream, each chunk repacked and fed to the ρ(k) rebuild as it lands; the gather and the pair kernel
// start when the last chunk is in.  Same result as mmc_upload_positions + mmc_potential.
int mmc_potential_host(mmc_handle *h, const double *coords, const double *com, int32_t style, mmc_properties *out)
{
    if (!h) return MMC_EINVAL;
    int rc = style_check(h, style, false);
    if (rc) return rc;
    if (!coords || !com || !out || style == MMC_STYLE_LJ_ATOMS) FAIL(MMC_EINVAL, "bad arguments");
    DevSystem &S = h->S;
    const int nchunk = h->host_chunks;
    CK(cudaSetDevice(h->cfg.device));
    if (h->cfg.world > 1 && h->peer_ready == h->cfg.world && S.n_sites >= 100000) {       // all ranks call this together
        EvalCtx Ed{1.0, S.box, S.kappa, S.cfac, h->cfg.rank, h->cfg.world};
        ErfPoly ep{};
        h->pair_level = h->pair_floor;
        if (v7_eligible(h, style, Ed, ep)) return potential_host_sharded(h, coords, com, style, ep, out);
    }
    if (!h->uniform || h->cfg.world != 1 || S.n_sites < 64 * nchunk || S.n_sites < 100000) {          // small or general systems: the plain sequence
        if ((rc = mmc_upload_positions(h, coords, com))) return rc;
        return h->cfg.world > 1 && h->peer_ready == h->cfg.world ? mmc_potential_sharded(h, style, out) : mmc_potential(h, style, out);
    }
    if ((rc = ensure_vec(h))) return rc;
    h->pair_level = h->pair_floor;
    if (h->pend_kind == 1) h->pend_kind = 0;
    h->trial_pending = false; h->vol_pending = false; h->new_valid = false;
    double *d_coords = reinterpret_cast<double *>(h->d_raw);
    double *d_com = d_coords + 4 * (size_t)S.n_sites;
    const bool ewald = style == MMC_STYLE_EWALD;
    EvalCtx E{1.0, S.box, S.kappa, S.cfac, 0, 1};
    ErfPoly ep{};
    const int ncd = grid_cells(h, style, S.box);
    const int nwin = (h->host_windows > 1 && v7_eligible(h, style, E, ep) && ncd >= 2 * h->host_windows) ? h->host_windows : 1;
    g_trace.on = std::getenv("MMC_TRACE_HOST") != nullptr;
    g_trace.mark(h->stream, "start");
    // ---- main stream: COMs first (the cell binning needs nothing else; issued before the site chunks so that it is not queued
    // behind them in the host->device copy engine)
    CK(cudaMemcpyAsync(d_com, com, sizeof(double) * 3 * S.n_mol, cudaMemcpyHostToDevice, h->stream));
    g_trace.mark(h->stream, "COM copy done");
    CK(cudaMemsetAsync(h->d_info, 0, 4 * sizeof(int), h->stream));
    k_repack_com<<<(S.n_mol + 255) / 256, 256, 0, h->stream>>>(d_com, S.n_mol, S.box, S.com, h->d_info); LAUNCH_CHECK();
    // ---- copy stream: the site array in chunks, nothing but copies (they follow the COMs through the copy engine; nothing of
    // this call has to be waited for: the previous evaluation has returned); side stream: each chunk repacked as it lands — its
    // stream order makes ev_chunk[c] mean "chunks 0..c are resident"; rk stream: (Ewald) the chunk's ρ(k) partials.
    int blocks = 0, cap = 0;
    if (ewald) {   // blocks a chunk needs (same formula as rhok_launch), to size the partial buffer once
        const bool v2 = h->n_kpairs <= 32 && S.nk <= 6 && h->use_rhok_v2;
        const int ck = v2 ? RHOK2_SITES : RHOKB_SITES;
        for (int c = 0; c < nchunk; ++c) {
            const int n = (int)((long long)S.n_sites * (c + 1) / nchunk) - (int)((long long)S.n_sites * c / nchunk);
            const int waves = 2 * h->sm_count * std::max(1, h->rhok_split);
            int per = std::max(2 * ck, (n + waves - 1) / waves);
            per = (per + ck - 1) / ck * ck;
            cap += std::max(1, (n + per - 1) / per);
        }
    }
    for (int c = 0; c < nchunk; ++c) {
        const int s0 = (int)((long long)S.n_sites * c / nchunk), s1 = (int)((long long)S.n_sites * (c + 1) / nchunk);
        CK(cudaMemcpyAsync(d_coords + 3 * (size_t)s0, coords + 3 * (size_t)s0, sizeof(double) * 3 * (size_t)(s1 - s0), cudaMemcpyHostToDevice, h->copy));
        CK(cudaEventRecord(h->ev_copy[c], h->copy));
        g_trace.mark(h->copy, "chunk copy done");
        CK(cudaStreamWaitEvent(h->side, h->ev_copy[c], 0));
        k_repack_sites<<<(s1 - s0 + 255) / 256, 256, 0, h->side>>>(d_coords, s0, s1, S.site); LAUNCH_CHECK();
        g_trace.mark(h->side, "chunk repacked");
        cudaEvent_t evc = c == nchunk - 1 ? h->ev_sites : h->ev_chunk[c];
        CK(cudaEventRecord(evc, h->side));
        if (ewald) {
            int nb = 0;
            CK(cudaStreamWaitEvent(h->rk, evc, 0));
            if ((rc = rhok_launch(h, S.site, s0, s1, S.box, nullptr, h->rk, blocks, &nb, cap))) return rc;
            g_trace.mark(h->rk, "chunk rho(k) partial done");
            blocks += nb;
        }
    }
    CK(cudaEventRecord(h->ev_join, ewald ? h->rk : h->side));
    h->state_version++;                                         // new positions: the cell buckets are rebuilt
    h->partial_resident = false;
    h->last_h2d_bytes = (long long)sizeof(double) * 3 * ((long long)S.n_sites + S.n_mol);
    // (a device->host COPY here would queue in the copy engine behind the site chunks and hold this stream back with it)
    if (h->host_mailbox) { k_bytes_to_host<<<1, 32, 0, h->stream>>>(nullptr, 0, h->d_info, 4, nullptr, h->d_up->info); LAUNCH_CHECK(); }
    else CK(cudaMemcpyAsync(h->h_up->info, h->d_info, 4 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    E.rhok_external = true; E.rhok_blocks = blocks; E.rhok_done = h->ev_join;
    cudaEvent_t win_ev[4];
    if (nwin > 1) {
        // The home cells are cut into nwin contiguous ranges; a window's gather and its pair kernel start as soon as the chunk
        // that completes the rows it reads has landed, while later chunks are still on the bus.  Which chunk that is follows
        // from the cell of every molecule — known now: the COMs are in.  (A spatially coherent order — a lattice start, a sorted
        // restart — lets the windows start early; for a random order every window needs the last chunk.)
        if ((rc = v7_alloc(h, ncd, ncd + 2))) return rc;
        const V7Grid G = v7_grid(h, style, E);
        const int ncell = ncd * ncd * ncd;
        if (h->win_ncd != ncd || h->win_n != nwin) {
            int wr[8];
            for (int w = 0; w <= nwin; ++w) wr[w] = (int)((long long)ncell * w / nwin);
            CK(cudaMemcpyAsync(h->d7_range + 16, wr, sizeof(int) * (nwin + 1), cudaMemcpyHostToDevice, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            h->win_ncd = ncd; h->win_n = nwin;
        }
        CK(cudaMemsetAsync(h->d7_count, 0, sizeof(int) * (size_t)ncell, h->stream));
        Bin7Args B{S.com, S.n_mol, (double)ncd / S.box, G, h->d7_count, h->d7_bucket, h->d_cell_of};
        k_bin7<<<(S.n_mol + 255) / 256, 256, 0, h->stream>>>(B); LAUNCH_CHECK();
        h->bin_version = h->state_version; h->bin_ncd = ncd; h->bin_world = 1;
        CK(cudaMemsetAsync(h->d7_range + 24, 0, 4 * sizeof(int), h->stream));
        V7Grid Gw = G; Gw.range = h->d7_range + 16; Gw.world = nwin;
        k_order7<<<nwin, 1024, 0, h->stream>>>(h->d7_count, ncd, Gw.range, 0, h->d7_order); LAUNCH_CHECK();
        k_window_need7<<<(S.n_mol + 255) / 256, 256, 0, h->stream>>>(h->d_cell_of, S.n_mol, h->US, S.n_sites, nchunk, Gw, nwin, h->d7_range + 24); LAUNCH_CHECK();
        if (h->host_mailbox) { k_bytes_to_host<<<1, 32, 0, h->stream>>>(nullptr, 0, h->d7_range + 24, 4, nullptr, h->d_up->win_need); LAUNCH_CHECK(); }
        else CK(cudaMemcpyAsync(h->h_up->win_need, h->d7_range + 24, 4 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        g_trace.mark(h->stream, "binned, window needs known");
        CK(cudaStreamSynchronize(h->stream));          // ~0.15 ms into the call; the site chunks are in flight on the copy stream meanwhile
        int need[4] = {h->h_up->win_need[0], h->h_up->win_need[1], h->h_up->win_need[2], h->h_up->win_need[3]};
        for (int w = 0; w < nwin; ++w) {
            int c = std::max(0, std::min(need[w], nchunk - 1));
            if (w > 0) c = std::max(c, std::max(0, std::min(need[w - 1], nchunk - 1)));
            need[w] = c;
            win_ev[w] = (c == nchunk - 1) ? h->ev_sites : h->ev_chunk[c];
        }
        E.n_windows = nwin; E.win_wait = win_ev;
    } else {
        E.wait_sites = h->ev_sites;
    }
    rc = evaluate_unsharded(h, style, E, S.rhok[0], S.rhok[1], out);
    g_trace.mark(h->stream, "result published");
    g_trace.dump();
    if (rc < 0) return rc;
    CK(cudaStreamSynchronize(h->stream));
    if (h->h_up->info[0] & REPACK_COM_OUTSIDE) FAIL(MMC_EINVAL, "a COM lies outside [0, box] (the reference's PBC keeps COMs inside)");
    return rc;
}

int mmc_last_host_bytes(mmc_handle *h, int64_t *h2d_bytes)
{
    if (!h || !h2d_bytes) return MMC_EINVAL;
    *h2d_bytes = h->last_h2d_bytes;
    return MMC_OK;
}

int mmc_volume_trial(mmc_handle *h, double box_new, double kappa_new, int32_t style, mmc_properties *out)
{
    if (!h) return MMC_EINVAL;
    int rc = style_check(h, style);
    if (rc) return rc;
    { int rcf = flush_pending(h); if (rcf) return rcf; }
    if (style == MMC_STYLE_LJ_ATOMS) FAIL(MMC_EINVAL, "volume trial is implemented for molecular systems");
    if (!out || !(box_new > 0)) FAIL(MMC_EINVAL, "bad arguments");
    if ((rc = ensure_vec(h))) return rc;
    const bool coul = (style == MMC_STYLE_EWALD || style == MMC_STYLE_WOLF);
    if (coul && !(kappa_new > 0)) FAIL(MMC_EINVAL, "kappa_new must be positive");
    if (style == MMC_STYLE_EWALD) {
        fill_cfac(h->kxyz, kappa_new, box_new, h->cfac_trial);          // PrepareEwaldVariables at L'
        CK(cudaMemcpyAsync(h->d_cfac_trial, h->cfac_trial.data(), sizeof(double) * h->S.nkvecs,
                           cudaMemcpyHostToDevice, h->stream));
    }
    const double f = box_new / h->S.box;                                 // volumeChange.jl:62
    EvalCtx E{f, box_new, coul ? kappa_new : h->S.kappa, h->d_cfac_trial, 0, 1};
    if ((rc = evaluate_unsharded(h, style, E, h->d_rhok_trial, nullptr, out))) return rc;
    h->vol_pending = true; h->vol_box = box_new; h->vol_kappa = E.kappa; h->vol_f = f; h->vol_style = style;
    return MMC_OK;
}

int mmc_volume_accept(mmc_handle *h)
{
    if (!h) return MMC_EINVAL;
    if (!h->vol_pending) FAIL(MMC_ESTATE, "mmc_volume_accept without a pending volume trial");
    { int rcf = flush_pending(h); if (rcf) return rcf; }
    k_apply_scale<<<(h->S.n_mol + 255) / 256, 256, 0, h->stream>>>(h->S, h->vol_f);
    LAUNCH_CHECK();
    h->S.box = h->vol_box;
    h->state_version++;
    if (h->vol_style == MMC_STYLE_EWALD || h->vol_style == MMC_STYLE_WOLF) h->S.kappa = h->vol_kappa;
    if (h->vol_style == MMC_STYLE_EWALD) {
        std::swap(h->S.cfac, h->d_cfac_trial);
        h->cfac.swap(h->cfac_trial);
        std::swap(h->S.rhok[h->cur], h->d_rhok_trial);
    } else if (h->has_ewald) {
        fill_cfac(h->kxyz, h->S.kappa, h->S.box, h->cfac);
        CK(cudaMemcpyAsync(h->S.cfac, h->cfac.data(), sizeof(double) * h->S.nkvecs, cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    if (h->has_ewald) get_erf_poly(h, h->S.kappa, h->S.rc_qq * h->S.rc_qq + 100, h->move_poly);
    h->vol_pending = false; h->new_valid = false; h->trial_pending = false;
    h->cnt.commits++;
    return MMC_OK;
}

int mmc_volume_reject(mmc_handle *h)
{
    if (!h) return MMC_EINVAL;
    if (!h->vol_pending) FAIL(MMC_ESTATE, "mmc_volume_reject without a pending volume trial");
    h->vol_pending = false;
    return MMC_OK;
}

int mmc_measure_fp64_peak(mmc_handle *h, double *tflops)
{
    if (!h || !tflops) return MMC_EINVAL;
    double *d = nullptr;
    CK(cudaMalloc(&d, sizeof(double)));
    const int blocks = h->sm_count * 8, iters = 16384;
    k_dfma_probe<<<blocks, 256, 0, h->stream>>>(d, 256);   // warm-up
    LAUNCH_CHECK();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(h->tm.ev[0], h->stream);
        k_dfma_probe<<<blocks, 256, 0, h->stream>>>(d, iters);
        LAUNCH_CHECK();
        cudaEventRecord(h->tm.ev[1], h->stream);
        CK(cudaStreamSynchronize(h->stream));
        float ms = 0;
        cudaEventElapsedTime(&ms, h->tm.ev[0], h->tm.ev[1]);
        best = std::min(best, ms);
    }
    cudaFree(d);
    const double flop = 2.0 * 8.0 * (double)iters * 256.0 * blocks;
    *tflops = flop / (best * 1e-3) / 1e12;
    return MMC_OK;
}

}  // extern "C"
