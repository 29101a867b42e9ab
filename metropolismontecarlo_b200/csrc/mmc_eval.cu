// mmc_eval.cu — C ABI of libmmc_b200.so, part 2: the full-system energy (potential(), Ewald/energy.jl:946-1032 and
// :864-943), its sharded forms, the host-array form and the volume move (Ewald/volumeChange.jl:50-147) on top of the
// pair kernels (kernels_pairs*.cuh), the rho(k) rebuild (kernels_recip.cuh) and the peer exchange (kernels_peer.cuh).
#include "mmc_handle.h"
#include "kernels_pairs.cuh"
#include "kernels_pairs_v5.cuh"
#include "kernels_pairs_v6.cuh"
#include "kernels_recip.cuh"
#include "kernels_upload.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

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
    const cudaEvent_t *chunk_ev = nullptr;   // mmc_potential_host: event of every site chunk, in upload order (the last one == wait_sites)
    int n_chunks = 0;
};


}  // namespace

namespace mmc_detail {

// out == nullptr: partials only, written from block `block0` on (the caller reduces all blocks later); *nb_out = blocks used
int rhok_launch(mmc_handle *h, const double4 *site, int s_begin, int s_end, double box, double2 *out, cudaStream_t st,
                int block0, int *nb_out, int cap_blocks)
{
    if (!st) st = h->stream;
    const int n = s_end - s_begin;
    const int nkv = h->S.nkvecs;
    const bool v2 = h->n_kpairs <= 32 && h->S.nk <= 6 && h->use_rhok_v2;
    const int chunk = v2 ? RHOK2_SITES : RHOK_SITES;
    const int waves = 2 * h->sm_count * std::max(1, h->rhok_split);
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
    if (h->tm.on) cudaEventRecord(h->tm.ev[2], st);
    if (v2) {
        Rhok2Args R{site, s_begin, s_end, per, nkv, h->n_kpairs, h->d_kpairs, h->d_kindex, box, part};
        switch (h->S.nk) {
            case 1: k_rhok_pairs<1><<<nb, RHOK2_BLOCK, 0, st>>>(R); break;
            case 2: k_rhok_pairs<2><<<nb, RHOK2_BLOCK, 0, st>>>(R); break;
            case 3: k_rhok_pairs<3><<<nb, RHOK2_BLOCK, 0, st>>>(R); break;
            case 4: k_rhok_pairs<4><<<nb, RHOK2_BLOCK, 0, st>>>(R); break;
            case 5: k_rhok_pairs<5><<<nb, RHOK2_BLOCK, 0, st>>>(R); break;
            default: k_rhok_pairs<6><<<nb, RHOK2_BLOCK, 0, st>>>(R); break;
        }
    } else {
        RhokArgs R{site, s_begin, s_end, per, h->S.nk, nkv, h->S.kvec, box, part};
        const int kpt = (nkv + RHOK_BLOCK - 1) / RHOK_BLOCK;
        if (kpt <= 1) k_rhok_partial<1><<<nb, RHOK_BLOCK, 0, st>>>(R);
        else if (kpt <= 2) k_rhok_partial<2><<<nb, RHOK_BLOCK, 0, st>>>(R);
        else if (kpt <= 4) k_rhok_partial<4><<<nb, RHOK_BLOCK, 0, st>>>(R);
        else if (kpt <= 8) k_rhok_partial<8><<<nb, RHOK_BLOCK, 0, st>>>(R);
        else FAIL(MMC_EINVAL, "too many k-vectors for the rebuild kernel (nk too large)");
    }
    LAUNCH_CHECK();
    if (h->tm.on) cudaEventRecord(h->tm.ev[3], st);
    if (out) {
        k_rhok_reduce<<<(nkv + 31) / 32, dim3(32, 32), 0, st>>>(h->d_rhok_partial, nb, nkv, out);
        LAUNCH_CHECK();
    }
    return MMC_OK;
}

}  // namespace mmc_detail

namespace {

// k_pairs_fast instantiations: water (3 sites) x tile {64, 128} x padded polynomial degree
#define MMC_FOR_DEGS(X) X(0) X(8) X(12) X(16) X(20) X(24) X(32) X(44)
#define MMC_FOR_POS_DEGS(X) X(8) X(12) X(16) X(20) X(24) X(32) X(44)
void launch_pairs_fast(int tile, int deg, int grid, size_t smem, cudaStream_t st, const PairArgs &P)
{
#define X(D)                                                                                   \
    if (deg == D) {                                                                            \
        if (tile == 64) k_pairs_fast<3, 64, D><<<grid, PAIR_BLOCK, smem, st>>>(P);             \
        else k_pairs_fast<3, 128, D><<<grid, PAIR_BLOCK, smem, st>>>(P);                       \
        return;                                                                                \
    }
    MMC_FOR_DEGS(X)
#undef X
}
#define MMC_FOR_DIRECT_DEGS(X) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12)
void launch_pairs_v5(int deg, bool direct, int grid, cudaStream_t st, const PairArgs &P, const int4 *slots)
{
    if (direct) {
#define X(D) if (deg == D) { k_pairs_v5<D, true><<<grid, V5_BLOCK, V5_SMEM, st>>>(P, slots); return; }
        MMC_FOR_DIRECT_DEGS(X)
#undef X
    } else {
#define X(D) if (deg == D) { k_pairs_v5<D, false><<<grid, V5_BLOCK, V5_SMEM, st>>>(P, slots); return; }
        MMC_FOR_POS_DEGS(X)
#undef X
    }
}
void launch_pairs_v6(int deg, bool direct, int grid, cudaStream_t st, const PairArgs &P, const int4 *slots, const V6Extra &X)
{
    if (direct) {
#define X_(D) if (deg == D) { k_pairs_v6<D, true><<<grid, V6_BLOCK, V6_SMEM, st>>>(P, slots, X); return; }
        MMC_FOR_DIRECT_DEGS(X_)
#undef X_
    } else {
#define X_(D) if (deg == D) { k_pairs_v6<D, false><<<grid, V6_BLOCK, V6_SMEM, st>>>(P, slots, X); return; }
        MMC_FOR_POS_DEGS(X_)
#undef X_
    }
}
}  // namespace

void mmc_detail::eval_set_attributes()
{
    cudaFuncSetAttribute(k_pairs<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(k_pairs<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(k_pairs<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(k_pairs<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
#define X(D) cudaFuncSetAttribute(k_pairs_v6<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V6_SMEM);
    MMC_FOR_DIRECT_DEGS(X)
#undef X
#define X(D) cudaFuncSetAttribute(k_pairs_v6<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V6_SMEM);
    MMC_FOR_POS_DEGS(X)
#undef X
#define X(D) cudaFuncSetAttribute(k_pairs_v5<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V5_SMEM);
    MMC_FOR_DIRECT_DEGS(X)
#undef X
#define X(D) cudaFuncSetAttribute(k_pairs_v5<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V5_SMEM);
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

// A pair kernel declined the state.  The water kernels (levels 0-1: v6, v5) share their preconditions (cell
// population <= 64, site reach inside the reference's +100 window), so a decline by one of them goes straight to
// k_pairs_fast (level 2); after that the general kernel (level 3).
bool escalate_pair_level(mmc_handle *h)
{
    h->pair_level = h->pair_level < 2 ? 2 : h->pair_level + 1;
    return h->pair_level <= 3;
}

// Leaves this rank's partial sums in d_vec: [0] Σlj_pot [1] Σlj_vir [2] Σcoul [3] #overlap
// [MMC_NSCAL ..) ρ(k) partial (re,im).
int eval_partials(mmc_handle *h, int style, const EvalCtx &E, double *d_vec)
{
    const bool force_general = h->pair_level >= 3 || E.per_mol != nullptr;
    if (!h->uniform) FAIL(MMC_EINVAL, "pair kernel needs a uniform topology (internal)");
    const DevSystem &S = h->S;
    const int US = h->US;
    const bool want_qq = (style == MMC_STYLE_EWALD || style == MMC_STYLE_WOLF);
    const double rcmax = std::max(S.rc_lj, want_qq ? S.rc_qq : 0.0);
    int ncd = (int)std::floor(E.box / rcmax);
    if (ncd > 128) ncd = 128;
    const bool cells = ncd >= 3;
    CK(cudaMemsetAsync(d_vec, 0, (MMC_NSCAL + 2 * (size_t)std::max(S.nkvecs, 1)) * sizeof(double), h->stream));
    if (h->tm.on) cudaEventRecord(h->tm.ev[4], h->stream);
    // The ρ(k) rebuild (RecipLong) depends on nothing the pair path produces when the box is unchanged (f == 1): it
    // reads the resident sites in their own order and runs on the side stream while binning, gather and the pair
    // kernel run here.  For a volume trial it needs the scaled sites and forks after the gather instead.
    const long long ns_all = S.n_sites;
    const int rs0 = (int)(ns_all * E.rank / E.world), rs1 = (int)(ns_all * (E.rank + 1) / E.world);
    bool rhok_forked = false;
    // (small systems: the rebuild is a few µs of work, the fork/join events would cost more than they hide)
    const bool rhok_side = h->overlap_rhok && !E.rhok_external && (long long)(rs1 - rs0) * S.nkvecs > 10000000LL;
    auto fork_rhok_resident = [&]() -> int {
        CK(cudaEventRecord(h->ev_fork, h->stream));
        CK(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
        int rcr = rhok_launch(h, S.site, rs0, rs1, E.box, reinterpret_cast<double2 *>(d_vec + MMC_NSCAL), h->side);
        if (rcr) return rcr;
        CK(cudaEventRecord(h->ev_join, h->side));
        rhok_forked = true;
        return MMC_OK;
    };
    if (style == MMC_STYLE_EWALD && rhok_side && h->overlap_rhok == 1 && E.f == 1.0) {
        int rcr = fork_rhok_resident();
        if (rcr) return rcr;
    }
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
        // fractional COM coordinates are invariant under the volume scaling: bin the resident state
        // a rank of a sharded evaluation reads the home cells of its unit range and their half-shell neighbours: z-layers
        // [z(first home cell), z(last home cell) + 1]; only those are ordered and gathered (units are (cell, group) triples)
        // Units are (cell, group) or (cell, slot) tuples, U per cell, dealt in contiguous ranges: for every U the first home
        // cell of rank r is floor(ncell·r/world) and the last one is at most ceil(ncell·(r+1)/world) − 1.
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
        if (style == MMC_STYLE_EWALD && rhok_side && h->overlap_rhok == 3 && E.f == 1.0) {   // fork after the (tiny, launch-bound) binning kernels
            int rcr = fork_rhok_resident();
            if (rcr) return rcr;
        }
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
    // k_pairs_v6 reads the state as 96-byte rows + float gate coordinates (written by the same gather)
    const bool want_rows = cells && US == 3 && h->pair_level == 0 && h->uniform_q && want_qq;
    if (want_rows && !h->d_mrows) {
        CK(cudaMalloc(&h->d_mrows, sizeof(double) * 12 * (size_t)S.n_mol));
        CK(cudaMalloc(&h->d_gf, sizeof(float4) * (size_t)S.n_mol));
    }
    GatherArgs G{S.com, S.site, cells ? h->d_perm : nullptr, S.n_mol, US, E.f, h->d_scom, h->d_ssite,
                 reinterpret_cast<unsigned long long *>(h->d_maxdev), h->d_ovl,
                 want_rows ? h->d_mrows : nullptr, want_rows ? h->d_gf : nullptr, h->d_cell_of, ncd, E.box / ncd,
                 zl_lo, std::min(zl_cnt, ncd)};
    // Sites still arriving from the host (mmc_potential_host): the home cells are cut into z-layer windows; a window's gather and its
    // share of the pair kernel start as soon as the chunk that completes its layers (+ one layer above, the half shell) has landed,
    // while later chunks are still on the bus.  Which chunk that is comes from the cell of every molecule (known: the COMs are in).
    int nwin = 1, win_need[4] = {0, 0, 0, 0};
    auto win_z = [&](int w) { return (int)((long long)ncd * w / nwin); };
    auto gather_window = [&](int w) -> int {          // layers not gathered by an earlier window: [zlo + (w > 0), zhi], zhi wraps to 0 for the last
        const int lo = win_z(w) + (w > 0 ? 1 : 0), hi = std::min(win_z(w + 1), ncd - 1);
        if (hi < lo) return MMC_OK;
        GatherArgs Gw = G; Gw.zl_lo = lo; Gw.zl_cnt = hi - lo + 1;
        k_gather<<<gm, tb, 0, h->stream>>>(Gw); LAUNCH_CHECK();
        return MMC_OK;
    };
    if (E.chunk_ev && E.n_chunks > 1 && cells && want_rows && E.world == 1 && E.f == 1.0 && h->v6_dynamic && h->host_windows > 1 &&
        ncd >= 4 * h->host_windows) {
        nwin = std::min(4, h->host_windows);
        if (!h->d_winneed) CK(cudaMalloc(&h->d_winneed, 4 * sizeof(int)));
        CK(cudaMemsetAsync(h->d_winneed, 0, 4 * sizeof(int), h->stream));
        k_window_need<<<gm, tb, 0, h->stream>>>(h->d_cell_of, S.n_mol, US, ncd, nwin, S.n_sites, E.n_chunks, h->d_winneed); LAUNCH_CHECK();
        CK(cudaMemcpyAsync(win_need, h->d_winneed, 4 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));          // ~0.15 ms into the call; the site chunks are in flight on the copy stream meanwhile
        for (int w = 0; w < nwin; ++w) win_need[w] = std::max(0, std::min(win_need[w], E.n_chunks - 1));
        for (int w = 1; w < nwin; ++w) win_need[w] = std::max(win_need[w], win_need[w - 1]);
        CK(cudaStreamWaitEvent(h->stream, E.chunk_ev[win_need[0]], 0));
        { int rcw = gather_window(0); if (rcw) return rcw; }
    } else {
        if (E.wait_sites) CK(cudaStreamWaitEvent(h->stream, E.wait_sites, 0));      // binning needed the COMs only; the gather needs the sites
        k_gather<<<gm, tb, 0, h->stream>>>(G); LAUNCH_CHECK();
    }
    if (h->tm.on) cudaEventRecord(h->tm.ev[5], h->stream);
    if (style == MMC_STYLE_EWALD && rhok_side && !rhok_forked) {      // volume trial: scaled, sorted sites
        CK(cudaEventRecord(h->ev_fork, h->stream));
        CK(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
        int rcr = rhok_launch(h, h->d_ssite, rs0, rs1, E.box, reinterpret_cast<double2 *>(d_vec + MMC_NSCAL), h->side);
        if (rcr) return rcr;
        CK(cudaEventRecord(h->ev_join, h->side));
        rhok_forked = true;
    }

    PairArgs P{};
    P.com = h->d_scom; P.site = h->d_ssite; P.cell_start = h->d_start;
    P.ncd = ncd; P.S = US; P.n_mol = S.n_mol; P.mode = cells ? 0 : 1;
    P.n_tiles = (S.n_mol + PAIR_TILE - 1) / PAIR_TILE;
    P.L = E.box; P.rc_lj2 = S.rc_lj * S.rc_lj; P.rc_qq2 = S.rc_qq * S.rc_qq; P.kappa = E.kappa;
    P.want_lj = 1; P.want_qq = want_qq ? 1 : 0;
    P.nlj = (int)h->lj.size(); P.lj = h->d_lj;
    for (int k = 0; k < 16; ++k) { P.lj_eps_tab[k] = 0.0; P.lj_sig_tab[k] = 0.0; }
    if (US <= 4) for (const LJActive &e : h->lj) { P.lj_eps_tab[e.a * US + e.b] = e.eps; P.lj_sig_tab[e.a * US + e.b] = e.sig; }
    P.partial = h->d_pair_partial; P.ovl = h->d_ovl; P.n_ovl = h->d_novl; P.max_dev = h->d_maxdev;
    P.err_flag = h->d_errflag;
    P.per_mol = E.per_mol;
    P.rclj_bits = 0; P.rcqq_bits = 0; P.cutlj_bits = 0; P.cutqq_bits = 0;
    { double v;
      v = P.rc_lj2; std::memcpy(&P.rclj_bits, &v, 8); v = P.rc_qq2; std::memcpy(&P.rcqq_bits, &v, 8);
      v = P.rc_lj2 + 100; std::memcpy(&P.cutlj_bits, &v, 8); v = P.rc_qq2 + 100; std::memcpy(&P.cutqq_bits, &v, 8); }
    P.ep = ErfPoly{};
    if (want_qq) get_erf_poly(h, E.kappa, S.rc_qq * S.rc_qq + 100, P.ep);
    const int max_cell = cells ? h->max_cell_cached : PAIR_TILE;
    // the water kernels serve: 3 sites, LJ only on site pair (0,0), equal cut-offs, Coulomb on, polynomial erf, identical
    // per-site charges in every molecule (they become launch constants)
    const bool water = cells && US == 3 && !force_general && h->pair_level <= 1 && max_cell <= V3_ACAP && want_qq &&
                       S.rc_lj == S.rc_qq && P.ep.deg > 0 && h->lj.size() == 1 && h->lj[0].a == 0 && h->lj[0].b == 0 && h->uniform_q;
    const bool v6 = water && h->pair_level == 0 && want_rows;
    const bool v5 = water && !v6;
    if (nwin > 1 && !v6) {       // another kernel serves this state: it wants the whole gathered copy
        if (E.wait_sites) CK(cudaStreamWaitEvent(h->stream, E.wait_sites, 0));
        k_gather<<<gm, tb, 0, h->stream>>>(G); LAUNCH_CHECK();
        nwin = 1;
    }
    const int tile = (US == 3 && !force_general && !v5 && !v6) ? (max_cell <= 64 ? 64 : (max_cell <= 128 ? 128 : 0)) : 0;
    if (v5 || v6) n_units = (long long)V3_GROUPS * ncd * ncd * ncd;
    if (v5 || v6) {
        P.qq_negmask = 0;
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {
                P.qq_tab[a * 3 + b] = h->q_site[a] * h->q_site[b];
                if (P.qq_tab[a * 3 + b] < 0.0) P.qq_negmask |= 1u << (a * 3 + b);
            }
        // conservative FP32 gate: |d²_f32 - d²| <= 2*sqrt(3)*rc*delta + 3*delta² + 3*2^-23*rc², delta = 6*2^-24*edge
        // (roundings to float of cell-local coordinates < edge, of their sum with the slot offset <= 2*edge, one
        // float subtraction of <= 3*edge: 7 half-ulps of edge at most); 4x safety
        const double edge = E.box / ncd, rc = S.rc_qq, delta = 8.0 * edge / 16777216.0;
        const double margin = 4.0 * (2.0 * 1.7320508075688772 * rc * delta + 3.0 * delta * delta + 3.6e-7 * rc * rc);
        P.gate_rc2f = std::nextafterf((float)(rc * rc + margin), INFINITY);
    }
    bool v5_direct = false;
    int v5_deg = 0;
    if (v5 || v6) {   // −κ folded into the coefficients; DIRECT: also κ^2k, so the kernel runs Horner in r² itself
        v5_direct = P.ep.ddeg > 0;
        v5_deg = v5_direct ? P.ep.ddeg : P.ep.deg;
        double k2k = 1.0;
        for (int k = 0; k <= v5_deg; ++k) {
            P.pc[k] = v5_direct ? -E.kappa * P.ep.a[k] * k2k : -E.kappa * P.ep.c[k];
            k2k *= E.kappa * E.kappa;
        }
        P.pk2s = P.ep.kappa2 * P.ep.scale;
    }
    P.unit_begin = n_units * E.rank / E.world;
    P.unit_end = n_units * (E.rank + 1) / E.world;
    const long long my_units = P.unit_end - P.unit_begin;
    int grid;
    long long v6_units = 0;      // > 0: k_pairs_v6 ran with tickets and left per-(unit, warp) sums
    if (h->tm.on) cudaEventRecord(h->tm.ev[0], h->stream);
    if (v5 || v6) {
        {
            const long long nslots = 14LL * ncd * ncd * ncd;
            if (nslots > h->slots_cap) {
                dfree(h->d_slots);
                CK(cudaMalloc(&h->d_slots, sizeof(int4) * nslots));
                h->slots_cap = nslots;
            }
            k_slots_build<<<(unsigned)((nslots + 255) / 256), 256, 0, h->stream>>>(P, h->d_slots, ncd * ncd * ncd);
            LAUNCH_CHECK();
            if (h->tm.on) cudaEventRecord(h->tm.ev[0], h->stream);
        }
        if (v6) {
            grid = (int)std::max(1LL, std::min<long long>(h->v6_ctas_per_sm * h->sm_count, my_units));
            V6Extra X{h->d_mrows, h->d_gf, nullptr, nullptr};
            if (h->v6_dynamic && E.world == 1) {   // measured on config E: -2.3 % on one GPU, but +5..25 µs on a rank's share of a 2..8-rank
                                                    // evaluation (few units per CTA: the greedy order ends on expensive units), so ranks keep the static deal
                const size_t need = (size_t)my_units * V6_WARPS;
                if (need > h->unit_partial_cap) {
                    dfree(h->d_unit_partial);
                    CK(cudaMalloc(&h->d_unit_partial, sizeof(double4) * need));
                    h->unit_partial_cap = need;
                }
                X.ticket = reinterpret_cast<unsigned int *>(h->d_flags + 5);      // cleared with the flags at the start of the evaluation
                X.unit_partial = h->d_unit_partial;
                v6_units = my_units;
            }
            if (nwin > 1) {          // one launch per z-layer window, each as soon as its sites are in
                const long long per_layer = (long long)V3_GROUPS * ncd * ncd;
                for (int w = 0; w < nwin; ++w) {
                    if (w > 0) {
                        CK(cudaStreamWaitEvent(h->stream, E.chunk_ev[win_need[w]], 0));
                        int rcw = gather_window(w); if (rcw) return rcw;
                        CK(cudaMemsetAsync(X.ticket, 0, sizeof(unsigned int), h->stream));
                    }
                    PairArgs Pw = P;
                    Pw.unit_begin = per_layer * win_z(w); Pw.unit_end = per_layer * win_z(w + 1);
                    V6Extra Xw = X; Xw.unit_partial = X.unit_partial + (size_t)Pw.unit_begin * V6_WARPS;
                    const int gw = (int)std::max(1LL, std::min<long long>(h->v6_ctas_per_sm * h->sm_count, Pw.unit_end - Pw.unit_begin));
                    launch_pairs_v6(v5_deg, v5_direct, gw, h->stream, Pw, h->d_slots, Xw);
                }
            } else
            launch_pairs_v6(v5_deg, v5_direct, grid, h->stream, P, h->d_slots, X);
        } else {
            grid = (int)std::max(1LL, std::min<long long>(4 * h->sm_count, my_units));
            launch_pairs_v5(v5_deg, v5_direct, grid, h->stream, P, h->d_slots);
        }
    } else if (tile) {
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
        launch_pairs_fast(tile, P.ep.deg, grid, smem, h->stream, P);
    } else {
        const size_t smem = (2 * PAIR_TILE + 2 * PAIR_TILE * (size_t)US) * sizeof(double4) +
                            (size_t)PAIR_WARPS * PAIR_QCAP * sizeof(unsigned);
        grid = (int)std::max(1LL, std::min<long long>(2 * h->sm_count, my_units));
        if (P.per_mol) {
            if (US == 3) k_pairs<3, true><<<grid, PAIR_BLOCK, smem, h->stream>>>(P);
            else k_pairs<0, true><<<grid, PAIR_BLOCK, smem, h->stream>>>(P);
        } else if (US == 3) k_pairs<3, false><<<grid, PAIR_BLOCK, smem, h->stream>>>(P);
        else k_pairs<0, false><<<grid, PAIR_BLOCK, smem, h->stream>>>(P);
    }
    LAUNCH_CHECK();
    if (h->tm.on) cudaEventRecord(h->tm.ev[1], h->stream);
    if (v6_units > 0) {          // fold the unit sums in unit order: 64 contiguous shares, then the usual final fold
        grid = (int)std::min<long long>(64, h->pair_grid);
        k_unit_fold<<<grid, 256, 0, h->stream>>>(h->d_unit_partial, v6_units * V6_WARPS, h->d_pair_partial); LAUNCH_CHECK();
    }
    k_pair_reduce<<<1, 256, 0, h->stream>>>(h->d_pair_partial, grid, h->d_novl, h->d_maxcount, h->d_errflag, d_vec);
    LAUNCH_CHECK();
    h->last_fast = v6 ? 6 : (v5 ? 5 : tile);
    h->last_mode = cells ? 0 : 1;
    h->last_ncd = ncd;

    if (style == MMC_STYLE_EWALD && !E.rhok_external) {
        if (rhok_forked) CK(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
        else {   // same stream: resident sites when the box is unchanged (a sharded rank gathers only its layers), scaled copy otherwise
            int rc = rhok_launch(h, E.f == 1.0 ? S.site : h->d_ssite, rs0, rs1, E.box, reinterpret_cast<double2 *>(d_vec + MMC_NSCAL));
            if (rc) return rc;
        }
    }
    return MMC_OK;
}

// single-molecule Coulomb row on the (scaled, sorted) evaluation copy, overlap pairs skipped
int overlap_row(mmc_handle *h, const EvalCtx &E, int sorted_index, double *row)
{
    ErfPoly poly{};                      // rare path: plain erfc()
    DevSystem V = h->S;
    V.site = h->d_ssite; V.com = h->d_scom; V.mol = h->d_mol_uniform;
    V.box = E.box; V.kappa = E.kappa;
    MoveArgs A{};
    A.i = sorted_index; A.n_cfg = 1; A.tiles = move_tiles(h); A.recip_blocks = 0;
    A.want_lj = 0; A.want_qq = 1; A.ignore_overlap = 1; A.cur = h->cur;
    int rc = launch_move_on(h, V, A, poly, false);
    if (rc) return rc;
    *row = h->h_out->qq[0];
    return MMC_OK;
}

// d_vec holds the (already rank-summed) partials; computes E_recip on the device, brings the
// scalars to the host and assembles Properties in the reference's order (energy.jl:972-1021).
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
    if (h->tm.on) cudaEventRecord(h->tm.ev[6], h->stream);
    CK(cudaStreamSynchronize(h->stream));
    if (E.world == 1) {
        if (h->last_mode == 0) h->max_cell_cached = (int)h->h_vec[6];
        if (h->h_vec[7] != 0.0) return 1;      // a cell outgrew the fast kernel's tile: caller re-runs
    } else if (h->h_vec[7] != 0.0) {           // summed over ranks: every rank takes the same branch
        if (!escalate_pair_level(h)) FAIL(MMC_ECUDA, "pair kernel fallback chain exhausted (internal)");
        h->max_cell_cached = -1;
        return 1;                               // MMC_RETRY: caller repeats partial + all-reduce + finalize
    }
    double lj_pot = h->h_vec[0], lj_vir = h->h_vec[1], coul = h->h_vec[2];
    const long long novl = (long long)h->h_vec[3];
    const double recip_raw = h->h_vec[4];
    if (novl > 0 && (style == MMC_STYLE_EWALD || style == MMC_STYLE_WOLF)) {
        // reference semantics: a molecule whose EwaldReal row hits the overlap rule contributes
        // 0 for its whole row (ewalds.jl:359-360 inside energy.jl:991-1001): U - ½ Σ_flagged row_i
        if (E.world > 1) FAIL(MMC_ESTATE, "overlap in a sharded evaluation: re-run unsharded (mmc_potential)");
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
    std::memset(out, 0, sizeof(*out));
    const double factor = S.factor;
    out->lj = lj_pot * 4;                       // Σ_i(4 pot_i)/2 over unique pairs
    const double vir_lj = lj_vir * 24 / 3.0;
    out->energy = out->lj;
    out->virial = vir_lj;
    out->overlaps = novl;
    h->last_pairs = (long long)h->h_vec[5];
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
    if (h->tm.on) {
        cudaEventElapsedTime(&h->tm.ms[0], h->tm.ev[0], h->tm.ev[1]);
        if (style == MMC_STYLE_EWALD) cudaEventElapsedTime(&h->tm.ms[1], h->tm.ev[2], h->tm.ev[3]);
        cudaEventElapsedTime(&h->tm.ms[2], h->tm.ev[4], h->tm.ev[5]);
        cudaEventElapsedTime(&h->tm.ms[3], h->tm.ev[4], h->tm.ev[6]);
    }
    h->cnt.full_energy_evals++;
    return MMC_OK;
}

// non-uniform topologies: literal Σ_i rows / 2 through the single-molecule kernel
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
    if (!h->uniform) FAIL(MMC_EINVAL, "sharded evaluation needs a uniform topology");
    if ((rc = ensure_vec(h))) return rc;
    EvalCtx E{1.0, h->S.box, h->S.kappa, h->S.cfac, h->cfg.rank, h->cfg.world};
    return eval_partials(h, style, E, d_partials);
}

int mmc_potential_finalize(mmc_handle *h, int32_t style, const double *d_partials, mmc_properties *out)
{
    if (!h) return MMC_EINVAL;
    int rc = style_check(h, style);
    if (rc) return rc;
    if (!d_partials || !out) FAIL(MMC_EINVAL, "null argument");
    EvalCtx E{1.0, h->S.box, h->S.kappa, h->S.cfac, h->cfg.rank, h->cfg.world};
    rc = finalize(h, style, E, const_cast<double *>(d_partials), h->S.rhok[0], h->S.rhok[1], out);
    if (style == MMC_STYLE_EWALD) h->new_valid = false;
    return rc;
}

// ---- sharded evaluation with the exchange over NVLink peer memory (kernels_peer.cuh) ----------------------
static size_t peer_flag_offset_doubles(const mmc_handle *h) { return 2 * (size_t)h->cfg.world * h->peer_nvec_cap; }

int mmc_peer_export(mmc_handle *h, void *handle64)
{
    if (!h || !handle64) return MMC_EINVAL;
    if (h->cfg.world < 1 || h->cfg.world > MMC_PEER_MAX) FAIL(MMC_EINVAL, "peer exchange supports up to 8 ranks");
    CK(cudaSetDevice(h->cfg.device));
    if (!h->d_peer_buf) {
        h->peer_nvec_cap = MMC_NSCAL + 2 * 4096;
        const size_t doubles = peer_flag_offset_doubles(h) + 2 * (size_t)h->cfg.world;
        CK(cudaMalloc(&h->d_peer_buf, doubles * sizeof(double)));
        CK(cudaMemset(h->d_peer_buf, 0, doubles * sizeof(double)));
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

// this rank's partial sums, pushed into every rank's exchange buffer (asynchronous: returns after the launches)
int mmc_potential_sharded_begin(mmc_handle *h, int32_t style)
{
    if (!h) return MMC_EINVAL;
    int rc = style_check(h, style);
    if (rc) return rc;
    if (style == MMC_STYLE_LJ_ATOMS) FAIL(MMC_EINVAL, "sharded evaluation is for molecular systems");
    if (!h->uniform) FAIL(MMC_EINVAL, "sharded evaluation needs a uniform topology");
    if (h->peer_ready != h->cfg.world) FAIL(MMC_ESTATE, "peer exchange not set up: mmc_peer_export / mmc_peer_import for every rank");
    if (h->sharded_pending) FAIL(MMC_ESTATE, "mmc_potential_sharded_end has not been called");
    if ((size_t)(MMC_NSCAL + 2 * std::max(h->S.nkvecs, 1)) > h->peer_nvec_cap) FAIL(MMC_EINVAL, "too many k-vectors for the exchange buffer");
    { int rcf = flush_pending(h); if (rcf) return rcf; }
    if ((rc = ensure_vec(h))) return rc;
    EvalCtx E{1.0, h->S.box, h->S.kappa, h->S.cfac, h->cfg.rank, h->cfg.world};
    if ((rc = eval_partials(h, style, E, h->d_vec))) return rc;
    h->peer_epoch += 1;
    const PeerArgs P = peer_args(h);
    k_peer_push<<<h->cfg.world, 256, 0, h->stream>>>(P, h->d_vec);
    LAUNCH_CHECK();
    h->sharded_pending = true; h->sharded_style = style;
    return MMC_OK;
}

// wait for every rank's push, add the slots in rank order, finalise.  MMC_RETRY as mmc_potential_finalize.
int mmc_potential_sharded_end(mmc_handle *h, mmc_properties *out)
{
    if (!h || !out) return MMC_EINVAL;
    if (!h->sharded_pending) FAIL(MMC_ESTATE, "mmc_potential_sharded_begin first");
    h->sharded_pending = false;
    const int style = h->sharded_style;
    const PeerArgs P = peer_args(h);
    k_peer_sum<<<1, 256, 0, h->stream>>>(P, h->d_peer_total, h->d_peer_status);
    LAUNCH_CHECK();
    EvalCtx E{1.0, h->S.box, h->S.kappa, h->S.cfac, h->cfg.rank, h->cfg.world};
    int rc = finalize(h, style, E, h->d_peer_total, h->S.rhok[0], h->S.rhok[1], out);
    if (style == MMC_STYLE_EWALD) h->new_valid = false;
    if (rc < 0) return rc;
    if (*(volatile int *)h->h_peer_status) FAIL(MMC_ENCCL, "peer exchange: a rank's partial sums did not arrive");   // finalize synchronised the stream
    return rc;
}

// all ranks call this together: begin + end, repeated while the pair kernel chain escalates (MMC_RETRY)
int mmc_potential_sharded(mmc_handle *h, int32_t style, mmc_properties *out)
{
    for (int attempt = 0; attempt < 8; ++attempt) {
        int rc = mmc_potential_sharded_begin(h, style);
        if (rc) return rc;
        rc = mmc_potential_sharded_end(h, out);
        if (rc != MMC_RETRY) return rc;
    }
    FAIL(MMC_ECUDA, "sharded potential did not converge on a pair kernel (internal)");
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
    if (!h->uniform) return potential_rows(h, style, out);
    if ((rc = ensure_vec(h))) return rc;
    EvalCtx E{1.0, h->S.box, h->S.kappa, h->S.cfac, 0, 1};
    if ((rc = eval_partials(h, style, E, h->d_vec))) return rc;
    rc = finalize(h, style, E, h->d_vec, h->S.rhok[0], h->S.rhok[1], out);
    while (rc == 1) {    // the chosen pair kernel declined this state (dense cell / wrapped molecules): next level
        if (!escalate_pair_level(h)) FAIL(MMC_ECUDA, "pair kernel fallback chain exhausted (internal)");
        if ((rc = eval_partials(h, style, E, h->d_vec))) return rc;
        rc = finalize(h, style, E, h->d_vec, h->S.rhok[0], h->S.rhok[1], out);
    }
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
    if (!h->uniform) FAIL(MMC_EINVAL, "mmc_energy_all needs a uniform topology");
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
    int rc = style_check(h, style);
    if (rc) return rc;
    if (!coords || !com || !out || style == MMC_STYLE_LJ_ATOMS) FAIL(MMC_EINVAL, "bad arguments");
    DevSystem &S = h->S;
    const int nchunk = h->host_chunks;
    if (!h->uniform || h->cfg.world != 1 || S.n_sites < 64 * nchunk) {          // small or general systems: the plain sequence
        if ((rc = mmc_upload_positions(h, coords, com))) return rc;
        return mmc_potential(h, style, out);
    }
    CK(cudaSetDevice(h->cfg.device));
    if ((rc = ensure_vec(h))) return rc;
    h->pair_level = h->pair_floor;
    if (h->pend_kind == 1) h->pend_kind = 0;
    h->trial_pending = false; h->vol_pending = false; h->new_valid = false;
    double *d_coords = reinterpret_cast<double *>(h->d_raw);
    double *d_com = d_coords + 4 * (size_t)S.n_sites;
    const bool ewald = style == MMC_STYLE_EWALD;
    // main stream: COMs
    CK(cudaMemcpyAsync(d_com, com, sizeof(double) * 3 * S.n_mol, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(h->d_info, 0, 4 * sizeof(int), h->stream));
    k_repack_com<<<(S.n_mol + 255) / 256, 256, 0, h->stream>>>(d_com, S.n_mol, S.box, S.com, h->d_info); LAUNCH_CHECK();
    CK(cudaEventRecord(h->ev_fork, h->stream));
    // copy stream: site chunks, each repacked as it lands; side stream: (Ewald) ρ(k) partials of a chunk as soon as it is in
    CK(cudaStreamWaitEvent(h->copy, h->ev_fork, 0));
    CK(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
    int blocks = 0, cap = 0;
    if (ewald) {   // blocks a chunk needs (same formula as rhok_launch), to size the partial buffer once
        const bool v2 = h->n_kpairs <= 32 && S.nk <= 6 && h->use_rhok_v2;
        const int ck = v2 ? RHOK2_SITES : RHOK_SITES;
        for (int c = 0; c < nchunk; ++c) {
            const int n = (int)((long long)S.n_sites * (c + 1) / nchunk) - (int)((long long)S.n_sites * c / nchunk);
            int per = std::max(2 * ck, (n + 2 * h->sm_count - 1) / (2 * h->sm_count));
            per = (per + ck - 1) / ck * ck;
            cap += std::max(1, (n + per - 1) / per);
        }
    }
    for (int c = 0; c < nchunk; ++c) {
        const int s0 = (int)((long long)S.n_sites * c / nchunk), s1 = (int)((long long)S.n_sites * (c + 1) / nchunk);
        // the copy stream carries nothing but copies: a repack kernel in it would hold the next copy back whenever the SMs are
        // taken by a window of the pair kernel (persistent CTAs).  Repack + ρ(k) partials of the chunk follow on the side stream.
        CK(cudaMemcpyAsync(d_coords + 3 * (size_t)s0, coords + 3 * (size_t)s0, sizeof(double) * 3 * (size_t)(s1 - s0), cudaMemcpyHostToDevice, h->copy));
        CK(cudaEventRecord(h->ev_copy[c], h->copy));
        CK(cudaStreamWaitEvent(h->side, h->ev_copy[c], 0));
        k_repack_sites<<<(s1 - s0 + 255) / 256, 256, 0, h->side>>>(d_coords, s0, s1, S.site); LAUNCH_CHECK();
        CK(cudaEventRecord(c == nchunk - 1 ? h->ev_sites : h->ev_chunk[c], h->side));
        if (ewald) {
            int nb = 0;
            if ((rc = rhok_launch(h, S.site, s0, s1, S.box, nullptr, h->side, blocks, &nb, cap))) return rc;
            blocks += nb;
        }
    }
    CK(cudaEventRecord(h->ev_join, h->side));
    EvalCtx E{1.0, S.box, S.kappa, S.cfac, 0, 1};
    E.wait_sites = h->ev_sites; E.rhok_external = true;
    cudaEvent_t chunk_events[8];
    for (int c = 0; c < nchunk; ++c) chunk_events[c] = (c == nchunk - 1) ? h->ev_sites : h->ev_chunk[c];
    E.chunk_ev = chunk_events; E.n_chunks = nchunk;
    for (;;) {
        if ((rc = eval_partials(h, style, E, h->d_vec))) return rc;
        CK(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
        if (ewald) {
            k_rhok_reduce<<<(S.nkvecs + 31) / 32, dim3(32, 32), 0, h->stream>>>(h->d_rhok_partial, blocks, S.nkvecs,
                                                                               reinterpret_cast<double2 *>(h->d_vec + MMC_NSCAL));
            LAUNCH_CHECK();
        }
        CK(cudaMemcpyAsync(h->h_up->info, h->d_info, 4 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        rc = finalize(h, style, E, h->d_vec, S.rhok[0], S.rhok[1], out);
        if (rc != 1) break;
        if (!escalate_pair_level(h)) FAIL(MMC_ECUDA, "pair kernel fallback chain exhausted (internal)");
        E.wait_sites = nullptr; E.chunk_ev = nullptr;            // the state is on the device now
    }
    if (rc < 0) return rc;
    if (h->h_up->info[0] & REPACK_COM_OUTSIDE) FAIL(MMC_EINVAL, "a COM lies outside [0, box] (the reference's PBC keeps COMs inside)");
    return rc;
}

int mmc_volume_trial(mmc_handle *h, double box_new, double kappa_new, int32_t style, mmc_properties *out)
{
    if (!h) return MMC_EINVAL;
    int rc = style_check(h, style);
    if (rc) return rc;
    { int rcf = flush_pending(h); if (rcf) return rcf; }
    if (style == MMC_STYLE_LJ_ATOMS) FAIL(MMC_EINVAL, "volume trial is implemented for molecular systems");
    if (!out || !(box_new > 0)) FAIL(MMC_EINVAL, "bad arguments");
    if (!h->uniform) FAIL(MMC_EINVAL, "volume trial needs a uniform topology");
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
    if ((rc = eval_partials(h, style, E, h->d_vec))) return rc;
    rc = finalize(h, style, E, h->d_vec, h->d_rhok_trial, nullptr, out);
    while (rc == 1) {
        if (!escalate_pair_level(h)) FAIL(MMC_ECUDA, "pair kernel fallback chain exhausted (internal)");
        if ((rc = eval_partials(h, style, E, h->d_vec))) return rc;
        rc = finalize(h, style, E, h->d_vec, h->d_rhok_trial, nullptr, out);
    }
    if (rc) return rc;
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
