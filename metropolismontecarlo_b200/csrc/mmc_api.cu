// mmc_api.cu — C ABI of libmmc_b200.so (include/mmc_b200.h), part 1: lifetime, upload, k-space tables and the
// per-move entry points on top of the sm_100a kernels in kernels_move.cuh / kernels_upload.cuh.
// No CPU fallback: every energy returned here was computed by a kernel in this directory.
#include "mmc_handle.h"
#include "kernels_upload.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <immintrin.h>

namespace {
std::string g_create_error;

void free_system(mmc_handle *h)
{
    dfree(h->S.site); dfree(h->S.com); dfree(h->S.mol); dfree(h->S.atype);
    dfree(h->d_intra); dfree(h->d_stype); dfree(h->d_atype_pad); dfree(h->d_lj); dfree(h->d_mol_uniform); dfree(h->d_qsums); dfree(h->d_raw); dfree(h->d_info); dfree(h->d_qpart);
    h->raw_bytes = 0; h->cap_mol = 0; h->cap_sites = 0;
    dfree(h->d_cell_of); dfree(h->d_start); dfree(h->d_perm); dfree(h->d_flags);
    h->d_count = h->d_fill = nullptr; h->d_maxcount = nullptr; h->d_novl = h->d_errflag = nullptr; h->d_maxdev = nullptr;
    dfree(h->d_chain); h->chain_bytes = 0;
    dfree(h->d_permol); dfree(h->d_permol_out);
    dfree(h->d_scom); dfree(h->d_ssite); dfree(h->d_pair_partial); dfree(h->d_ovl);
    dfree(h->d_rhok_partial); dfree(h->d_units);
    h->units_cap = 0;
    dfree(h->d7_flags); dfree(h->d7_count); dfree(h->d7_bucket); dfree(h->d7_ecount); dfree(h->d7_rows); dfree(h->d7_gf);
    dfree(h->d7_unit_partial); dfree(h->d7_order); dfree(h->d7_block_sums); dfree(h->d7_need); dfree(h->d7_range); dfree(h->d7_rhok_scratch); h->d7_scratch_cap = 0;
    if (h->h7_need) { cudaFreeHost(h->h7_need); h->h7_need = nullptr; }
    h->need_cap = 0;
    h->d7_ncd = 0; h->d7_partial_cap = 0; h->bin_version = 0;
    h->max_cell_cached = -1;
    h->has_system = false;
}

void free_ewald(mmc_handle *h)
{
    dfree(h->S.kvec); dfree(h->S.cfac); dfree(h->S.rhok[0]); dfree(h->S.rhok[1]);
    dfree(h->d_rhok_trial); dfree(h->d_cfac_trial); dfree(h->d_vec); dfree(h->d_kpairs); dfree(h->d_kcombos); dfree(h->d_kindex);
    if (h->h_vec) cudaFreeHost(h->h_vec);
    h->h_vec = nullptr;
    h->has_ewald = false;
}

void free_atoms(mmc_handle *h)
{
    dfree(h->At.r); dfree(h->At.es); dfree(h->d_rows); dfree(h->d_atoms_out);
    h->has_atoms = false;
}

// wait until every CTA of launch `seq` has published its four packets
int wait_slots(mmc_handle *h, int nblocks)
{
    if (h->cfg.sync_mode == 1) CK(cudaStreamSynchronize(h->stream));
    unsigned spins = 0;
    for (int b = 0; b < nblocks; ++b)
        for (int k = 0; k < 4; ++k) {
            volatile unsigned long long *flag = &h->h_slots[b].p[k].seq;
            while (*flag != h->seq) {
                _mm_pause();
                if ((++spins & 0xfffffu) == 0 || h->cfg.sync_mode == 1) {   // make sure the kernel is still alive
                    cudaError_t q = cudaStreamQuery(h->stream);
                    if (q != cudaSuccess && q != cudaErrorNotReady) {
                        h->err = std::string("move kernel: ") + cudaGetErrorString(q);
                        return MMC_ECUDA;
                    }
                    if (q == cudaSuccess && *flag != h->seq)
                        FAIL(MMC_ECUDA, "move kernel finished without publishing its result");
                }
            }
        }
    return MMC_OK;
}

}  // namespace

namespace mmc_detail {

int ensure_vec(mmc_handle *h)
{
    if (h->d_vec) return MMC_OK;
    const size_t n = MMC_NSCAL + 2 * (size_t)std::max(h->S.nkvecs, 1);
    CK(cudaMalloc(&h->d_vec, n * sizeof(double)));
    CK(cudaMemsetAsync(h->d_vec, 0, n * sizeof(double), h->stream));
    CK(cudaHostAlloc(&h->h_vec, n * sizeof(double), cudaHostAllocDefault));
    return MMC_OK;
}

int move_tiles(const mmc_handle *h)
{
    const int t = (h->S.n_mol + MOVE_BLOCK - 1) / MOVE_BLOCK;
    return std::max(1, std::min(t, h->sm_count));
}

// writes a pending accepted move to HBM with its own tiny launch (only needed when the next call
// is not a move launch, which would carry it in its parameters)
int flush_pending(mmc_handle *h)
{
    if (h->pend_kind == 1) {
        MoveArgs T{};
        std::memcpy(T.site_new, h->pend_site, sizeof(double) * 3 * h->pend_ns);
        k_set_molecule<<<1, 32, 0, h->stream>>>(h->S, h->pend_i, h->pend_com[0], h->pend_com[1], h->pend_com[2], T);
        LAUNCH_CHECK();
    } else if (h->pend_kind == 2) {
        k_set_atom<<<1, 1, 0, h->stream>>>(h->At, h->pend_i, h->pend_com[0], h->pend_com[1], h->pend_com[2]);
        LAUNCH_CHECK();
    }
    h->pend_kind = 0;
    return MMC_OK;
}

// launch k_move on system `sys`, wait for its CTAs and fold their partials in CTA order
// (energy.jl:289 pot*4, vir*24/3; ewalds.jl:360 "return 0.0, true"; main.jl:580-590 skip on overlap)
int launch_move_on(mmc_handle *h, const DevSystem &sys, MoveArgs &A, const ErfPoly &poly, bool carry_commit)
{
    A.seq = ++h->seq;
    A.commit_i = -1;
    if (carry_commit && h->pend_kind == 1) {
        A.commit_i = h->pend_i; A.commit_ns = h->pend_ns;
        std::memcpy(A.commit_com, h->pend_com, sizeof(A.commit_com));
        std::memcpy(A.commit_site, h->pend_site, sizeof(double) * 3 * h->pend_ns);
        h->pend_kind = 0;                       // this launch writes it back
    }
    { const int2 mi = (&sys == &h->S) ? mol_of(h, A.i) : make_int2(A.i * h->ES, h->ES)      /* the evaluation copy: ES slots per molecule */; A.i_first = mi.x; A.i_ns = mi.y; }
    const int blocks = A.n_cfg * A.tiles + A.recip_blocks;
    if (blocks <= 0 || blocks > h->max_slots) FAIL(MMC_EINVAL, "bad move launch size");
    if (sys.max_sites <= 3) k_move<3><<<blocks, MOVE_BLOCK, 0, h->stream>>>(sys, A, poly, h->W);
    else if (sys.max_sites == 4) k_move<4><<<blocks, MOVE_BLOCK, 0, h->stream>>>(sys, A, poly, h->W);
    else k_move<0><<<blocks, MOVE_BLOCK, 0, h->stream>>>(sys, A, poly, h->W);
    LAUNCH_CHECK();
    int rc = wait_slots(h, blocks);
    if (rc) return rc;
    MoveOut &o = h->mout;
    bool any_ovl = false;
    for (int cfg = 0; cfg < 2; ++cfg) {
        double lp = 0, lv = 0, cq = 0, ov = 0;
        if (cfg < A.n_cfg)
            for (int t = 0; t < A.tiles; ++t) {
                const MoveSlot &sl = h->h_slots[cfg * A.tiles + t];
                lp += sl.p[0].v; lv += sl.p[1].v; cq += sl.p[2].v; ov += sl.p[3].v;
            }
        const bool ovl = (ov > 0.0) && !A.ignore_overlap;
        any_ovl |= ovl;
        o.lj_pot[cfg] = lp * 4;
        o.lj_vir[cfg] = lv * 24 / 3.0;
        o.qq[cfg] = ovl ? 0.0 : cq;
        o.overlap[cfg] = (ov > 0.0) ? 1 : 0;
    }
    double dr = 0.0;
    for (int t = 0; t < A.recip_blocks; ++t) dr += h->h_slots[A.n_cfg * A.tiles + t].p[0].v;
    o.d_recip = any_ovl ? 0.0 : dr * sys.factor;
    return MMC_OK;
}

int launch_move(mmc_handle *h, MoveArgs &A) { return launch_move_on(h, h->S, A, h->move_poly, true); }

void fill_cfac(const std::vector<int32_t> &kxyz, double kappa, double box, std::vector<double> &cfac)
{
    const double b = 1.0 / 4.0 / kappa / kappa / box / box;
    const double twopi = 2.0 * M_PI, twopi_sq = twopi * twopi;
    const size_t n = kxyz.size() / 3;
    cfac.resize(n);
    for (size_t i = 0; i < n; ++i) {
        const int kx = kxyz[3 * i], ky = kxyz[3 * i + 1], kz = kxyz[3 * i + 2];
        const double kr_sq = twopi_sq * (double)(kx * kx + ky * ky + kz * kz);
        double c = twopi * std::exp(-b * kr_sq) / kr_sq / box;
        if (kx > 0) c = c * 2.0;
        cfac[i] = c;
    }
}

// smooth part of erfc(κr)/r on the domain r² < r_cut²+100 the reference imposes (ewalds.jl:362);
// fits are cached on a geometric grid of domain ends so NPT box changes reuse them
void get_erf_poly(mmc_handle *h, double kappa, double r2_max, ErfPoly &P)
{
    const double vg = erfpoly::grid_vmax(kappa * kappa * r2_max);
    for (auto &e : h->poly_cache)
        if (e.first == vg) { P = e.second; P.kappa = kappa; P.kappa2 = kappa * kappa; return; }
    ErfPoly Q{};
    Q.kappa = kappa; Q.kappa2 = kappa * kappa;
    erfpoly::fit(vg, Q);
    if (h->poly_cache.size() >= 32) h->poly_cache.erase(h->poly_cache.begin());
    h->poly_cache.emplace_back(vg, Q);
    P = Q;
}

int style_check(mmc_handle *h, int style, bool need_full_state)
{
    if (need_full_state && h->partial_resident)
        FAIL(MMC_ESTATE, "after mmc_potential_host on a sharded handle only this rank's slab of the sites is resident: mmc_upload_positions first");
    if (style == MMC_STYLE_LJ_ATOMS) {
        if (!h->has_atoms) FAIL(MMC_ESTATE, "no atomic system uploaded");
        return MMC_OK;
    }
    if (style != MMC_STYLE_EWALD && style != MMC_STYLE_WOLF && style != MMC_STYLE_LJ_ONLY)
        FAIL(MMC_EINVAL, "unknown style");
    if (!h->has_system) FAIL(MMC_ESTATE, "no molecular system uploaded");
    if ((style == MMC_STYLE_EWALD || style == MMC_STYLE_WOLF) && !h->has_ewald)
        FAIL(MMC_ESTATE, "mmc_ewald_prepare has not been called");
    return MMC_OK;
}

// Σ_mol Σ_{a<b} q_a q_b erf(κ r_ab)/r_ab on the resident state (one small evaluation: n_mol threads)
int intra_energy(mmc_handle *h, double kappa, double *e_unscaled)
{
    if (!h->d_intra) CK(cudaMalloc(&h->d_intra, 257 * sizeof(double)));
    const int nb = std::max(1, std::min(256, (h->S.n_mol + 255) / 256));
    k_intra_partial<<<nb, 256, 0, h->stream>>>(h->S.site, h->S.mol, h->S.n_mol, kappa, h->S.box, h->d_intra); LAUNCH_CHECK();
    k_intra_final<<<1, 256, 0, h->stream>>>(h->d_intra, nb, h->d_intra + 256); LAUNCH_CHECK();
    CK(cudaMemcpyAsync(e_unscaled, h->d_intra + 256, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return MMC_OK;
}

}  // namespace mmc_detail

namespace {
using namespace mmc_detail;

int launch_move_atom(mmc_handle *h, AtomArgs &A)
{
    A.seq = ++h->seq;
    A.commit_i = -1;
    if (h->pend_kind == 2) {
        A.commit_i = h->pend_i;
        std::memcpy(A.commit_r, h->pend_com, sizeof(A.commit_r));
        h->pend_kind = 0;
    }
    if (A.blocks > h->max_slots) FAIL(MMC_EINVAL, "bad move launch size");
    k_move_atom<<<A.blocks, ATOM_BLOCK, 0, h->stream>>>(h->At, A, h->W);
    LAUNCH_CHECK();
    int rc = wait_slots(h, A.blocks);
    if (rc) return rc;
    double p0 = 0, v0 = 0, p1 = 0, v1 = 0;
    for (int t = 0; t < A.blocks; ++t) {
        const MoveSlot &sl = h->h_slots[t];
        p0 += sl.p[0].v; v0 += sl.p[1].v; p1 += sl.p[2].v; v1 += sl.p[3].v;
    }
    MoveOut &o = h->mout;
    o.lj_pot[0] = p0 * 4.0; o.lj_vir[0] = v0 * 24.0 / 3.0;     // mainMonatomic.jl:271
    o.lj_pot[1] = p1 * 4.0; o.lj_vir[1] = v1 * 24.0 / 3.0;
    o.qq[0] = o.qq[1] = 0.0; o.d_recip = 0.0; o.overlap[0] = o.overlap[1] = 0;
    return MMC_OK;
}

int check_mol_index(mmc_handle *h, int64_t i)
{
    if (!h->has_system) FAIL(MMC_ESTATE, "no molecular system uploaded");
    if (h->partial_resident)
        FAIL(MMC_ESTATE, "after mmc_potential_host on a sharded handle only this rank's slab of the sites is resident: mmc_upload_positions first");
    if (i < 1 || i > h->S.n_mol) FAIL(MMC_EINVAL, "molecule index out of range (1-based)");
    return MMC_OK;
}

void fill_kvectors(int nk, int k_sq_max, std::vector<int32_t> &kxyz)
{
    kxyz.clear();
    for (int kx = 0; kx <= nk; ++kx)           // Ewald/ewalds.jl:70-91 loop order
        for (int ky = -nk; ky <= nk; ++ky)
            for (int kz = -nk; kz <= nk; ++kz) {
                const int k_sq = kx * kx + ky * ky + kz * kz;
                if (k_sq < k_sq_max && k_sq != 0) { kxyz.push_back(kx); kxyz.push_back(ky); kxyz.push_back(kz); }
            }
}

}  // namespace

// ============================================================================ C ABI
extern "C" {

int mmc_version(void) { return MMC_VERSION; }

const char *mmc_last_error(const mmc_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int mmc_create(const mmc_config *cfg, mmc_handle **out)
{
    if (!cfg || !out) { g_create_error = "null argument"; return MMC_EINVAL; }
    if (cfg->world < 1 || cfg->rank < 0 || cfg->rank >= cfg->world) { g_create_error = "bad rank/world"; return MMC_EINVAL; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) +
                         " (libmmc_b200 has no CPU fallback)";
        return MMC_ECUDA;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { g_create_error = "device ordinal out of range"; return MMC_EINVAL; }
    if ((e = cudaSetDevice(cfg->device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return MMC_ECUDA; }
    mmc_handle *h = new mmc_handle();
    h->cfg = *cfg;
    auto fail = [&](const char *what, cudaError_t ce) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(ce);
        delete h;
        return MMC_ECUDA;
    };
    if (cfg->stream) h->stream = (cudaStream_t)cfg->stream;
    else {
        int prio_lo = 0, prio_hi = 0;                       // main stream at the highest priority, side stream at the lowest:
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);   // the ρ(k) rebuild fills what binning/gather/pairs leave free
        if ((e = cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, prio_hi)) != cudaSuccess) return fail("stream", e);
        h->own_stream = true;
    }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, cfg->device)) != cudaSuccess) return fail("props", e);
    h->sm_count = prop.multiProcessorCount;
    h->max_slots = 2 * h->sm_count + 64 + 1024;
    if ((e = cudaHostAlloc((void **)&h->h_slots, sizeof(MoveSlot) * h->max_slots, cudaHostAllocMapped)) != cudaSuccess) return fail("hostalloc", e);
    std::memset(h->h_slots, 0, sizeof(MoveSlot) * h->max_slots);
    if ((e = cudaHostGetDevicePointer((void **)&h->W.slots, h->h_slots, 0)) != cudaSuccess) return fail("mapped ptr", e);
    if ((e = cudaHostAlloc((void **)&h->h7_res, sizeof(double) * (MMC_NSCAL + 8), cudaHostAllocMapped)) != cudaSuccess) return fail("hostalloc", e);
    std::memset(h->h7_res, 0, sizeof(double) * (MMC_NSCAL + 8));
    if ((e = cudaHostGetDevicePointer((void **)&h->d7_res, h->h7_res, 0)) != cudaSuccess) return fail("mapped ptr", e);
    for (auto &ev : h->tm.ev)
        if ((e = cudaEventCreate(&ev)) != cudaSuccess) return fail("event", e);
    {
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        if ((e = cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, prio_lo)) != cudaSuccess) return fail("side stream", e);
    }
    if ((e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming)) != cudaSuccess) return fail("event", e);
    if ((e = cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming)) != cudaSuccess) return fail("event", e);
    if ((e = cudaEventCreateWithFlags(&h->ev_sites, cudaEventDisableTiming)) != cudaSuccess) return fail("event", e);
    if ((e = cudaStreamCreateWithFlags(&h->copy, cudaStreamNonBlocking)) != cudaSuccess) return fail("copy stream", e);
    {
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        if ((e = cudaStreamCreateWithPriority(&h->rk, cudaStreamNonBlocking, prio_lo)) != cudaSuccess) return fail("rk stream", e);
    }
    if ((e = cudaEventCreateWithFlags(&h->ev_rk, cudaEventDisableTiming)) != cudaSuccess) return fail("event", e);
    for (auto &ev : h->ev_chunk)
        if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return fail("event", e);
    for (auto &ev : h->ev_copy)
        if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return fail("event", e);
    eval_set_attributes();
    *out = h;
    return MMC_OK;
}

int mmc_destroy(mmc_handle *h)
{
    if (!h) return MMC_OK;
    cudaSetDevice(h->cfg.device);
    cudaStreamSynchronize(h->stream);
    free_system(h); free_ewald(h); free_atoms(h);
    if (h->h_slots) cudaFreeHost(h->h_slots);
    if (h->h7_res) cudaFreeHost(h->h7_res);
    if (h->h_up) cudaFreeHost(h->h_up);
    for (auto &ev : h->tm.ev) cudaEventDestroy(ev);
    for (int q = 0; q < MMC_PEER_MAX; ++q) if (h->peer_opened[q] && h->peer_base[q]) cudaIpcCloseMemHandle(h->peer_base[q]);
    dfree(h->d_peer_buf); dfree(h->d_peer_total);
    if (h->h_peer_status) cudaFreeHost(h->h_peer_status);
    if (h->side) { cudaStreamSynchronize(h->side); cudaStreamDestroy(h->side); }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->ev_sites) cudaEventDestroy(h->ev_sites);
    if (h->copy) { cudaStreamSynchronize(h->copy); cudaStreamDestroy(h->copy); }
    if (h->rk) { cudaStreamSynchronize(h->rk); cudaStreamDestroy(h->rk); }
    if (h->ev_rk) cudaEventDestroy(h->ev_rk);
    for (auto &ev : h->ev_chunk) if (ev) cudaEventDestroy(ev);
    for (auto &ev : h->ev_copy) if (ev) cudaEventDestroy(ev);
    if (h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
    return MMC_OK;
}

int mmc_upload_system(mmc_handle *h, int64_t n_mol, int64_t n_sites, const double *coords,
                      const double *charge, const int64_t *atype, const int64_t *first_atom,
                      const int64_t *last_atom, const double *com, int32_t n_types, const double *eps,
                      const double *sig, double box, double rc_lj, double rc_qq)
{
    if (!h) return MMC_EINVAL;
    if (!coords || !charge || !atype || !first_atom || !last_atom || !com || !eps || !sig) FAIL(MMC_EINVAL, "null array");
    if (n_mol < 1 || n_sites < n_mol || n_mol > (1 << 30) || n_sites > (1LL << 30)) FAIL(MMC_EINVAL, "bad sizes");
    if (n_types < 1 || n_types > MMC_MAX_TYPES) FAIL(MMC_EINVAL, "n_types out of range (1..8)");
    if (!(box > 0) || !(rc_lj > 0) || !(rc_qq > 0)) FAIL(MMC_EINVAL, "box and cutoffs must be positive");
    CK(cudaSetDevice(h->cfg.device));
    DevSystem &S = h->S;
    const double box_before = h->has_system ? S.box : 0.0;
    const bool realloc_needed = !h->has_system || h->cap_mol != (int)n_mol || h->cap_sites != (int)n_sites;
    if (realloc_needed) {
        const DevSystem keep = S;
        const bool had_ewald = h->has_ewald;
        free_system(h);
        S = DevSystem{};
        if (had_ewald) {   // k-space tables survive a re-upload of coordinates
            S.kappa = keep.kappa; S.factor = keep.factor; S.nk = keep.nk; S.nkvecs = keep.nkvecs;
            S.kvec = keep.kvec; S.cfac = keep.cfac; S.rhok[0] = keep.rhok[0]; S.rhok[1] = keep.rhok[1];
        }
        CK(cudaMalloc(&S.site, sizeof(double4) * n_sites));
        CK(cudaMalloc(&S.com, sizeof(double4) * n_mol));
        CK(cudaMalloc(&S.mol, sizeof(int2) * n_mol));
        CK(cudaMalloc(&S.atype, sizeof(int) * n_sites));
        h->raw_bytes = sizeof(double) * (size_t)(4 * n_sites + 3 * n_mol) + sizeof(int64_t) * (size_t)(n_sites + 2 * n_mol);
        CK(cudaMalloc(&h->d_raw, h->raw_bytes));
        CK(cudaMalloc(&h->d_info, 4 * sizeof(int)));
        CK(cudaMalloc(&h->d_qpart, 256 * sizeof(double2)));
        CK(cudaMalloc(&h->d_qsums, 2 * sizeof(double)));
        CK(cudaMalloc(&h->d_mol_uniform, sizeof(int2) * n_mol));
        CK(cudaMalloc(&h->d_lj, sizeof(LJActive) * 64));
        CK(cudaMalloc(&h->d_cell_of, sizeof(int) * n_mol));
        CK(cudaMalloc(&h->d_perm, sizeof(int) * n_mol));
        CK(cudaMalloc(&h->d_scom, sizeof(double4) * n_mol));
        CK(cudaMalloc(&h->d_ssite, sizeof(double4) * n_sites));
        h->ssite_cap = (size_t)n_sites;
        h->pair_grid = 5 * h->sm_count;
        CK(cudaMalloc(&h->d_pair_partial, sizeof(double4) * h->pair_grid));
        CK(cudaMalloc(&h->d_ovl, sizeof(unsigned) * n_mol));
        h->ncell_cap = 0; h->rhok_grid_cap = 0;
        h->cap_mol = (int)n_mol; h->cap_sites = (int)n_sites;
        if (!h->h_up) {
            CK(cudaHostAlloc((void **)&h->h_up, sizeof(*h->h_up), cudaHostAllocMapped));
            CK(cudaHostGetDevicePointer((void **)&h->d_up, h->h_up, 0));
        }
    }
    h->has_system = false;
    // ---- DMA the caller's arrays as they are (pinned sources go at full PCIe rate), repack on the device
    unsigned char *r = h->d_raw;
    double *d_coords = reinterpret_cast<double *>(r); r += sizeof(double) * 3 * n_sites;
    double *d_charge = reinterpret_cast<double *>(r); r += sizeof(double) * n_sites;
    double *d_com = reinterpret_cast<double *>(r); r += sizeof(double) * 3 * n_mol;
    long long *d_atype = reinterpret_cast<long long *>(r); r += sizeof(int64_t) * n_sites;
    long long *d_first = reinterpret_cast<long long *>(r); r += sizeof(int64_t) * n_mol;
    long long *d_last = reinterpret_cast<long long *>(r);
    CK(cudaMemcpyAsync(d_coords, coords, sizeof(double) * 3 * n_sites, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_charge, charge, sizeof(double) * n_sites, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_com, com, sizeof(double) * 3 * n_mol, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_atype, atype, sizeof(int64_t) * n_sites, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_first, first_atom, sizeof(int64_t) * n_mol, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_last, last_atom, sizeof(int64_t) * n_mol, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(h->d_info, 0, 4 * sizeof(int), h->stream));
    RepackArgs R{d_coords, d_charge, d_com, d_atype, d_first, d_last, (int)n_mol, (int)n_sites, n_types, box,
                 S.site, S.com, S.mol, S.atype, h->d_info};
    k_repack<<<(unsigned)((n_sites + 255) / 256), 256, 0, h->stream>>>(R);
    LAUNCH_CHECK();
    const int qb = (int)std::min<int64_t>(256, (n_sites + 255) / 256);
    k_charge_partial<<<qb, 256, 0, h->stream>>>(S.site, (int)n_sites, h->d_qpart); LAUNCH_CHECK();
    k_charge_final<<<1, 256, 0, h->stream>>>(h->d_qpart, qb, h->d_qsums); LAUNCH_CHECK();
    CK(cudaMemcpyAsync(h->h_up->info, h->d_info, 4 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(h->h_up->qs, h->d_qsums, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    const int err = h->h_up->info[0];
    if (err & REPACK_BAD_ATYPE) FAIL(MMC_EINVAL, "atype out of range (1-based)");
    if (err & REPACK_BAD_RANGE) FAIL(MMC_EINVAL, "first_atom/last_atom out of range");
    if (err & REPACK_TOO_MANY_SITES) FAIL(MMC_EINVAL, "more than 16 sites in a molecule");
    if (err & REPACK_COM_OUTSIDE) FAIL(MMC_EINVAL, "a COM lies outside [0, box] (the reference's PBC keeps COMs inside)");
    S.n_mol = (int)n_mol; S.n_sites = (int)n_sites; S.max_sites = h->h_up->info[1]; S.n_types = n_types;
    S.box = box; S.rc_lj = rc_lj; S.rc_qq = rc_qq;
    for (int a = 0; a < n_types * n_types; ++a) { S.eps[a] = eps[a]; S.sig[a] = sig[a]; }
    h->sum_q = h->h_up->qs[0]; h->sum_q2 = h->h_up->qs[1];
    bool uniform = h->h_up->info[2] == 0;
    const int US = (int)(last_atom[0] - first_atom[0] + 1);
    h->uniform_q = uniform && h->h_up->info[3] == 0 && US <= MMC_MAX_SITES;
    if (h->uniform_q) for (int a = 0; a < US; ++a) h->q_site[a] = charge[a];
    // LJ-active site-type combinations of the uniform molecule (ε_ij > 0.001, energy.jl:270)
    h->lj.clear();
    if (uniform)
        for (int a = 0; a < US; ++a)
            for (int b = 0; b < US; ++b) {
                const int ta = (int)atype[a] - 1, tb = (int)atype[b] - 1;
                const double e_ = eps[ta + tb * n_types];
                if (e_ > 0.001) h->lj.push_back(LJActive{a, b, e_, sig[ta + tb * n_types]});
            }
    if (h->lj.size() > 64) uniform = false;
    h->uniform = uniform; h->US = uniform ? US : 0;
    S.uni = h->US;
    // any other topology (molecules of different size or type sequence, ≤ 16 sites each): the cell path works on a copy padded
    // to max_sites slots per molecule and resolves LJ through the active TYPE pairs
    h->mixed = !uniform; h->ES = uniform ? US : S.max_sites;
    if (h->mixed) {
        h->lj.clear();
        for (int ta = 0; ta < n_types; ++ta)
            for (int tb = 0; tb < n_types; ++tb)
                if (eps[ta + tb * n_types] > 0.001) h->lj.push_back(LJActive{ta, tb, eps[ta + tb * n_types], sig[ta + tb * n_types]});
        const size_t need = (size_t)n_mol * h->ES;
        if (need > h->ssite_cap) {
            dfree(h->d_ssite);
            CK(cudaMalloc(&h->d_ssite, sizeof(double4) * need));
            h->ssite_cap = need;
        }
        dfree(h->d_stype); dfree(h->d_atype_pad);
        CK(cudaMalloc(&h->d_stype, need));
        CK(cudaMalloc(&h->d_atype_pad, need * sizeof(int)));       // types of the padded copy as k_move wants them (Coulomb rows only: all 0)
        CK(cudaMemsetAsync(h->d_atype_pad, 0, need * sizeof(int), h->stream));
        k_mol_packed<<<(unsigned)((n_mol + 255) / 256), 256, 0, h->stream>>>(h->d_mol_uniform, (int)n_mol, h->ES); LAUNCH_CHECK();
        if (!h->lj.empty())
            CK(cudaMemcpyAsync(h->d_lj, h->lj.data(), sizeof(LJActive) * h->lj.size(), cudaMemcpyHostToDevice, h->stream));
    }
    h->h_mol.clear();
    if (uniform) {
        if (realloc_needed || true) {
            // packed {m*US, US}: identical to S.mol for a uniform topology
            CK(cudaMemcpyAsync(h->d_mol_uniform, S.mol, sizeof(int2) * n_mol, cudaMemcpyDeviceToDevice, h->stream));
        }
        if (!h->lj.empty())
            CK(cudaMemcpyAsync(h->d_lj, h->lj.data(), sizeof(LJActive) * h->lj.size(), cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    } else {
        h->h_mol.resize(n_mol);
        for (int64_t m = 0; m < n_mol; ++m) h->h_mol[m] = make_int2((int)(first_atom[m] - 1), (int)(last_atom[m] - first_atom[m] + 1));
    }
    h->max_cell_cached = -1;
    h->pair_level = h->pair_floor; h->v7_left_for_overlap = false;
    if (h->pend_kind == 1) h->pend_kind = 0;
    if (h->has_ewald) {
        get_erf_poly(h, S.kappa, rc_qq * rc_qq + 100, h->move_poly);
        if (box != box_before) {
            // the k-space tables survive a re-upload, but cfac depends on the box (PrepareEwaldVariables, ewalds.jl:52,78-83):
            // rebuilt here for the new box with the resident kappa, and the resident rho(k) (a sum over the old positions)
            // is cleared like after mmc_ewald_prepare.  A caller that follows the reference's kappa = alpha/box convention
            // calls mmc_ewald_prepare again for the new kappa.
            fill_cfac(h->kxyz, S.kappa, box, h->cfac);
            CK(cudaMemcpyAsync(S.cfac, h->cfac.data(), sizeof(double) * S.nkvecs, cudaMemcpyHostToDevice, h->stream));
            CK(cudaMemsetAsync(S.rhok[0], 0, sizeof(double2) * S.nkvecs, h->stream));
            CK(cudaMemsetAsync(S.rhok[1], 0, sizeof(double2) * S.nkvecs, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            h->cur = 0;
        }
    }
    h->has_system = true;
    h->partial_resident = false;
    h->state_version++;
    h->trial_pending = false; h->vol_pending = false; h->new_valid = false;
    return MMC_OK;
}

int mmc_upload_atoms(mmc_handle *h, int64_t n, const double *r, const double *eps_j, const double *sig_j,
                     double box, double r_cut)
{
    if (!h) return MMC_EINVAL;
    if (!r || !eps_j || !sig_j || n < 1 || n > (1 << 30)) FAIL(MMC_EINVAL, "bad arguments");
    if (!(box > 0) || !(r_cut > 0)) FAIL(MMC_EINVAL, "box and cutoff must be positive");
    CK(cudaSetDevice(h->cfg.device));
    free_atoms(h);
    std::vector<double4> hr(n);
    std::vector<double2> he(n);
    for (int64_t i = 0; i < n; ++i) {
        hr[i] = make_double4(r[3 * i], r[3 * i + 1], r[3 * i + 2], 0.0);
        he[i] = make_double2(eps_j[i], sig_j[i]);
    }
    h->At.n = (int)n; h->At.box = box; h->At.rc = r_cut;
    CK(cudaMalloc(&h->At.r, sizeof(double4) * n));
    CK(cudaMalloc(&h->At.es, sizeof(double2) * n));
    CK(cudaMalloc(&h->d_rows, sizeof(double2) * n));
    CK(cudaMalloc(&h->d_atoms_out, 2 * sizeof(double)));
    CK(cudaMemcpyAsync(h->At.r, hr.data(), sizeof(double4) * n, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->At.es, he.data(), sizeof(double2) * n, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->has_atoms = true;
    if (h->pend_kind == 2) h->pend_kind = 0;
    h->trial_pending = false;
    return MMC_OK;
}

// All positions of an uploaded system at once (soa.coords, moa.COM after the caller changed them itself: a
// checkpoint, its own volume scaling, ...): the bulk form of mmc_set_molecule (Ewald/main.jl:527,552).  Charges,
// types and topology stay; the resident ρ(k) is NOT rebuilt (call mmc_recip_long / mmc_potential as after an upload).
int mmc_upload_positions(mmc_handle *h, const double *coords, const double *com)
{
    if (!h) return MMC_EINVAL;
    if (!h->has_system) FAIL(MMC_ESTATE, "no molecular system uploaded");
    if (!coords || !com) FAIL(MMC_EINVAL, "null array");
    CK(cudaSetDevice(h->cfg.device));
    DevSystem &S = h->S;
    double *d_coords = reinterpret_cast<double *>(h->d_raw);
    double *d_com = d_coords + 4 * (size_t)S.n_sites;      // same slots as in mmc_upload_system's staging block
    CK(cudaMemcpyAsync(d_coords, coords, sizeof(double) * 3 * S.n_sites, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_com, com, sizeof(double) * 3 * S.n_mol, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(h->d_info, 0, 4 * sizeof(int), h->stream));
    k_repack_positions<<<(unsigned)((S.n_sites + 255) / 256), 256, 0, h->stream>>>(d_coords, d_com, S.n_mol, S.n_sites, S.box,
                                                                                  S.site, S.com, h->d_info);
    LAUNCH_CHECK();
    h->state_version++;
    h->partial_resident = false;
    CK(cudaMemcpyAsync(h->h_up->info, h->d_info, 4 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (h->h_up->info[0] & REPACK_COM_OUTSIDE) FAIL(MMC_EINVAL, "a COM lies outside [0, box] (the reference's PBC keeps COMs inside)");
    h->pair_level = h->pair_floor; h->v7_left_for_overlap = false;
    if (h->pend_kind == 1) h->pend_kind = 0;
    h->trial_pending = false; h->vol_pending = false; h->new_valid = false;
    return MMC_OK;
}

int mmc_download_system(mmc_handle *h, double *coords, double *com)
{
    if (!h) return MMC_EINVAL;
    if (!h->has_system) FAIL(MMC_ESTATE, "no molecular system uploaded");
    { int rcf = flush_pending(h); if (rcf) return rcf; }
    std::vector<double4> hs(h->S.n_sites), hc(h->S.n_mol);
    CK(cudaMemcpyAsync(hs.data(), h->S.site, sizeof(double4) * hs.size(), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(hc.data(), h->S.com, sizeof(double4) * hc.size(), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (coords) for (size_t s = 0; s < hs.size(); ++s) { coords[3 * s] = hs[s].x; coords[3 * s + 1] = hs[s].y; coords[3 * s + 2] = hs[s].z; }
    if (com) for (size_t m = 0; m < hc.size(); ++m) { com[3 * m] = hc[m].x; com[3 * m + 1] = hc[m].y; com[3 * m + 2] = hc[m].z; }
    return MMC_OK;
}

int mmc_download_atoms(mmc_handle *h, double *r)
{
    if (!h || !r) return MMC_EINVAL;
    if (!h->has_atoms) FAIL(MMC_ESTATE, "no atomic system uploaded");
    { int rcf = flush_pending(h); if (rcf) return rcf; }
    std::vector<double4> hr(h->At.n);
    CK(cudaMemcpyAsync(hr.data(), h->At.r, sizeof(double4) * hr.size(), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (size_t i = 0; i < hr.size(); ++i) { r[3 * i] = hr[i].x; r[3 * i + 1] = hr[i].y; r[3 * i + 2] = hr[i].z; }
    return MMC_OK;
}

int mmc_ewald_prepare(mmc_handle *h, double kappa, int32_t nk, int32_t k_sq_max, double factor, int32_t *nkvecs)
{
    if (!h) return MMC_EINVAL;
    if (!h->has_system) FAIL(MMC_ESTATE, "upload the system before mmc_ewald_prepare (cfac depends on the box)");
    if (!(kappa > 0) || nk < 1 || nk > 16 || k_sq_max < 2) FAIL(MMC_EINVAL, "bad Ewald parameters (1 <= nk <= 16)");
    CK(cudaSetDevice(h->cfg.device));
    free_ewald(h);
    fill_kvectors(nk, k_sq_max, h->kxyz);
    const int n = (int)(h->kxyz.size() / 3);
    if (n < 1) FAIL(MMC_EINVAL, "no k-vectors inside k_sq_max");
    fill_cfac(h->kxyz, kappa, h->S.box, h->cfac);
    std::vector<int4> kv(n);
    for (int i = 0; i < n; ++i) kv[i] = make_int4(h->kxyz[3 * i], h->kxyz[3 * i + 1], h->kxyz[3 * i + 2], 0);
    DevSystem &S = h->S;
    S.kappa = kappa; S.factor = factor; S.nk = nk; S.nkvecs = n;
    h->k_sq_max = k_sq_max;
    CK(cudaMalloc(&S.kvec, sizeof(int4) * n));
    CK(cudaMalloc(&S.cfac, sizeof(double) * n));
    CK(cudaMalloc(&h->d_cfac_trial, sizeof(double) * n));
    CK(cudaMalloc(&S.rhok[0], sizeof(double2) * n));
    CK(cudaMalloc(&S.rhok[1], sizeof(double2) * n));
    CK(cudaMalloc(&h->d_rhok_trial, sizeof(double2) * n));
    CK(cudaMemcpyAsync(S.kvec, kv.data(), sizeof(int4) * n, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(S.cfac, h->cfac.data(), sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
    {   // lookup tables of the pair-wise rebuild kernel: (kx,|ky|) pairs that own at least one k-vector, index cube
        const int W = 2 * nk + 1;
        std::vector<int> kindex((size_t)(nk + 1) * W * W, -1);
        std::vector<char> used((size_t)(nk + 1) * (nk + 1), 0);
        for (int i = 0; i < n; ++i) {
            const int kx = h->kxyz[3 * i], ky = h->kxyz[3 * i + 1], kz = h->kxyz[3 * i + 2];
            kindex[((size_t)kx * W + (ky + nk)) * W + (kz + nk)] = i;
            used[(size_t)kx * (nk + 1) + std::abs(ky)] = 1;
        }
        // sorted by kx² + ky² (stable): a tile of 32 consecutive pairs then shares its largest useful kz, and k_rhok_big
        // skips the (pair tile, kz tile) combos that hold no k-vector at all
        std::vector<int2> kp;
        std::vector<int> kzmax((size_t)(nk + 1) * (nk + 1), 0);
        for (int i = 0; i < n; ++i) {
            int &m = kzmax[(size_t)h->kxyz[3 * i] * (nk + 1) + std::abs(h->kxyz[3 * i + 1])];
            m = std::max(m, std::abs(h->kxyz[3 * i + 2]));
        }
        for (int kx = 0; kx <= nk; ++kx)
            for (int ky = 0; ky <= nk; ++ky)
                if (used[(size_t)kx * (nk + 1) + ky]) kp.push_back(make_int2(kx, ky));
        std::stable_sort(kp.begin(), kp.end(), [](const int2 &p, const int2 &q) { return p.x * p.x + p.y * p.y < q.x * q.x + q.y * q.y; });
        h->n_kpairs = (int)kp.size();
        {   // combos for ZT = 4 and 5 kz values per warp: the cheaper list wins (per combo and site: 8 + 8·ZT DFMA slots)
            std::vector<int2> best; int best_zt = 4; long long best_cost = -1;
            for (int zt = 4; zt <= 5; ++zt) {
                std::vector<int2> cb;
                for (int pt = 0; pt * 32 < (int)kp.size(); ++pt) {
                    int mz = 0;
                    for (int j = pt * 32; j < std::min((int)kp.size(), pt * 32 + 32); ++j) mz = std::max(mz, kzmax[(size_t)kp[j].x * (nk + 1) + kp[j].y]);
                    const int nzt = std::max(1, (mz + zt - 1) / zt);          // (tile 0 of a pair tile also carries kz = 0)
                    for (int z = 0; z < nzt; ++z) cb.push_back(make_int2(pt, z));
                }
                const long long cost = (long long)cb.size() * (8 + 8 * zt);
                if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = cb; best_zt = zt; }
            }
            h->k_zt = best_zt; h->n_kcombos = (int)best.size();
            dfree(h->d_kcombos);
            CK(cudaMalloc(&h->d_kcombos, sizeof(int2) * best.size()));
            CK(cudaMemcpyAsync(h->d_kcombos, best.data(), sizeof(int2) * best.size(), cudaMemcpyHostToDevice, h->stream));
        }
        CK(cudaMalloc(&h->d_kpairs, sizeof(int2) * kp.size()));
        CK(cudaMalloc(&h->d_kindex, sizeof(int) * kindex.size()));
        CK(cudaMemcpyAsync(h->d_kpairs, kp.data(), sizeof(int2) * kp.size(), cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->d_kindex, kindex.data(), sizeof(int) * kindex.size(), cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    CK(cudaMemsetAsync(S.rhok[0], 0, sizeof(double2) * n, h->stream));   // zeros(ComplexF64, NKVECS), ewalds.jl:98-99
    CK(cudaMemsetAsync(S.rhok[1], 0, sizeof(double2) * n, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->cur = 0; h->new_valid = false; h->rhok_grid_cap = 0; dfree(h->d_rhok_partial);
    h->has_ewald = true;
    get_erf_poly(h, kappa, h->S.rc_qq * h->S.rc_qq + 100, h->move_poly);
    int rc = ensure_vec(h);
    if (rc) return rc;
    if (nkvecs) *nkvecs = n;
    return MMC_OK;
}

int mmc_get_kvectors(mmc_handle *h, int32_t *kxyz, double *cfac)
{
    if (!h) return MMC_EINVAL;
    if (!h->has_ewald) FAIL(MMC_ESTATE, "mmc_ewald_prepare has not been called");
    if (kxyz) std::memcpy(kxyz, h->kxyz.data(), sizeof(int32_t) * h->kxyz.size());
    if (cfac) CK(cudaMemcpy(cfac, h->S.cfac, sizeof(double) * h->S.nkvecs, cudaMemcpyDeviceToHost));
    return MMC_OK;
}

int mmc_get_rhok(mmc_handle *h, double *sum_old, double *sum_new)
{
    if (!h) return MMC_EINVAL;
    if (!h->has_ewald) FAIL(MMC_ESTATE, "mmc_ewald_prepare has not been called");
    CK(cudaStreamSynchronize(h->stream));
    const size_t bytes = sizeof(double2) * h->S.nkvecs;
    if (sum_old) CK(cudaMemcpy(sum_old, h->S.rhok[h->cur], bytes, cudaMemcpyDeviceToHost));
    if (sum_new) CK(cudaMemcpy(sum_new, h->S.rhok[h->new_valid ? (h->cur ^ 1) : h->cur], bytes, cudaMemcpyDeviceToHost));
    return MMC_OK;
}

int mmc_lj_mol(mmc_handle *h, int64_t i, double *pot, double *vir)
{
    if (!h) return MMC_EINVAL;
    int rc = check_mol_index(h, i);
    if (rc) return rc;
    MoveArgs A{};
    A.i = (int)(i - 1); A.n_cfg = 1; A.tiles = move_tiles(h); A.want_lj = 1; A.cur = h->cur;
    if ((rc = launch_move(h, A))) return rc;
    if (pot) *pot = h->h_out->lj_pot[0];
    if (vir) *vir = h->h_out->lj_vir[0];
    return MMC_OK;
}

int mmc_ewald_real(mmc_handle *h, int64_t i, double *pot, int32_t *overlap)
{
    if (!h) return MMC_EINVAL;
    int rc = check_mol_index(h, i);
    if (rc) return rc;
    if (!h->has_ewald) FAIL(MMC_ESTATE, "mmc_ewald_prepare has not been called");
    MoveArgs A{};
    A.i = (int)(i - 1); A.n_cfg = 1; A.tiles = move_tiles(h); A.want_qq = 1; A.cur = h->cur;
    if ((rc = launch_move(h, A))) return rc;
    if (pot) *pot = h->h_out->qq[0];
    if (overlap) *overlap = h->h_out->overlap[0];
    return MMC_OK;
}

int mmc_ewald_short(mmc_handle *h, int64_t i, double *e, double *v, int32_t *overlap)
{
    double pot = 0.0;
    int rc = mmc_ewald_real(h, i, &pot, overlap);
    if (rc) return rc;
    const double realEwald = pot * h->S.factor;      // ewalds.jl:905-907
    if (e) *e = realEwald;
    if (v) *v = realEwald / 3;
    return MMC_OK;
}

int mmc_set_molecule(mmc_handle *h, int64_t i, const double com[3], const double *sites)
{
    if (!h) return MMC_EINVAL;
    int rc = check_mol_index(h, i);
    if (rc) return rc;
    if (!com || !sites) FAIL(MMC_EINVAL, "null array");
    if ((rc = flush_pending(h))) return rc;
    MoveArgs A{};
    A.i = (int)(i - 1);
    std::memcpy(A.site_new, sites, sizeof(double) * 3 * mol_of(h, i - 1).y);
    k_set_molecule<<<1, 32, 0, h->stream>>>(h->S, A.i, com[0], com[1], com[2], A);
    LAUNCH_CHECK();
    h->state_version++;
    h->trial_pending = false;
    return MMC_OK;
}

int mmc_recip_move(mmc_handle *h, const double *r_old, const double *r_new, const double *q, int32_t n, double *dE)
{
    if (!h) return MMC_EINVAL;
    if (!h->has_system || !h->has_ewald) FAIL(MMC_ESTATE, "system and Ewald tables required");
    if (!r_old || !r_new || !q || n < 1 || n > MMC_MAX_SITES) FAIL(MMC_EINVAL, "bad arguments (1 <= n <= 16)");
    if (h->S.nk > MMC_MAX_NK) FAIL(MMC_EINVAL, "the per-move k-space kernels are sized for nk <= 8 (the full-energy path takes nk <= 16)");
    MoveArgs A{};
    A.i = 0; A.n_cfg = 0; A.tiles = 0; A.cur = h->cur;
    A.recip_blocks = (h->S.nkvecs + MOVE_BLOCK - 1) / MOVE_BLOCK;
    A.recip_ns = n; A.recip_from_args = 1;
    std::memcpy(A.site_new, r_new, sizeof(double) * 3 * n);
    std::memcpy(A.site_old, r_old, sizeof(double) * 3 * n);
    std::memcpy(A.q, q, sizeof(double) * n);
    int rc = launch_move(h, A);
    if (rc) return rc;
    h->new_valid = true;
    if (dE) *dE = h->h_out->d_recip;
    return MMC_OK;
}

int mmc_recip_commit(mmc_handle *h)
{
    if (!h) return MMC_EINVAL;
    if (!h->has_ewald) FAIL(MMC_ESTATE, "mmc_ewald_prepare has not been called");
    if (h->new_valid) { h->cur ^= 1; h->new_valid = false; h->cnt.commits++; }
    return MMC_OK;
}

int mmc_recip_rollback(mmc_handle *h)
{
    if (!h) return MMC_EINVAL;
    if (!h->has_ewald) FAIL(MMC_ESTATE, "mmc_ewald_prepare has not been called");
    h->new_valid = false;
    return MMC_OK;
}

int mmc_ewald_self(mmc_handle *h, double *energy)
{
    if (!h) return MMC_EINVAL;
    if (!h->has_system || !h->has_ewald) FAIL(MMC_ESTATE, "system and Ewald tables required");
    if (energy) *energy = -h->S.kappa * h->sum_q2 / std::sqrt(M_PI) * h->S.factor;   // Σq² reduced on the device at upload
    return MMC_OK;
}

int mmc_lj_atom(mmc_handle *h, int64_t i, double *pot, double *vir)
{
    if (!h) return MMC_EINVAL;
    if (!h->has_atoms) FAIL(MMC_ESTATE, "no atomic system uploaded");
    if (i < 1 || i > h->At.n) FAIL(MMC_EINVAL, "atom index out of range (1-based)");
    AtomArgs A{};
    A.i = (int)(i - 1); A.n_cfg = 1;
    A.blocks = std::max(1, std::min((h->At.n + ATOM_BLOCK - 1) / ATOM_BLOCK, 2 * h->sm_count));
    int rc = launch_move_atom(h, A);
    if (rc) return rc;
    if (pot) *pot = h->h_out->lj_pot[0];
    if (vir) *vir = h->h_out->lj_vir[0];
    return MMC_OK;
}

int mmc_set_atom(mmc_handle *h, int64_t i, const double r[3])
{
    if (!h) return MMC_EINVAL;
    if (!h->has_atoms) FAIL(MMC_ESTATE, "no atomic system uploaded");
    if (i < 1 || i > h->At.n || !r) FAIL(MMC_EINVAL, "bad arguments");
    { int rc = flush_pending(h); if (rc) return rc; }
    k_set_atom<<<1, 1, 0, h->stream>>>(h->At, (int)(i - 1), r[0], r[1], r[2]);
    LAUNCH_CHECK();
    h->trial_pending = false;
    return MMC_OK;
}

int mmc_trial_move(mmc_handle *h, int64_t i, const double com_new[3], const double *sites_new, int32_t style,
                   mmc_trial_result *out)
{
    if (!h) return MMC_EINVAL;
    int rc = check_mol_index(h, i);
    if (rc) return rc;
    if ((rc = style_check(h, style))) return rc;
    if (style == MMC_STYLE_LJ_ATOMS || !com_new || !sites_new || !out) FAIL(MMC_EINVAL, "bad arguments");
    if (style == MMC_STYLE_EWALD && h->S.nk > MMC_MAX_NK) FAIL(MMC_EINVAL, "the per-move k-space kernels are sized for nk <= 8 (the full-energy path takes nk <= 16)");
    MoveArgs &A = h->last;
    A = MoveArgs{};
    A.i = (int)(i - 1); A.n_cfg = 2; A.tiles = move_tiles(h);
    A.want_lj = 1; A.want_qq = (style != MMC_STYLE_LJ_ONLY);
    A.recip_blocks = (style == MMC_STYLE_EWALD) ? (h->S.nkvecs + MOVE_BLOCK - 1) / MOVE_BLOCK : 0;
    A.cur = h->cur;
    A.com_new[0] = com_new[0]; A.com_new[1] = com_new[1]; A.com_new[2] = com_new[2];
    std::memcpy(A.site_new, sites_new, sizeof(double) * 3 * mol_of(h, i - 1).y);
    if ((rc = launch_move(h, A))) return rc;
    const MoveOut &o = *h->h_out;
    const double factor = h->S.factor;
    out->lj_old = o.lj_pot[0]; out->lj_vir_old = o.lj_vir[0];
    out->lj_new = o.lj_pot[1]; out->lj_vir_new = o.lj_vir[1];
    out->qq_old = o.qq[0] * factor; out->qq_vir_old = out->qq_old / 3;     // ewalds.jl:905-907
    out->qq_new = o.qq[1] * factor; out->qq_vir_new = out->qq_new / 3;
    out->d_recip = o.d_recip;
    out->overlap_old = o.overlap[0]; out->overlap_new = o.overlap[1];
    h->last_overlap = o.overlap[0] || o.overlap[1];
    h->trial_pending = true; h->trial_kind = 1; h->trial_style = style;
    h->new_valid = (style == MMC_STYLE_EWALD) && !h->last_overlap;
    h->cnt.trial_moves++;
    if (h->last_overlap) h->cnt.overlap_events++;
    return MMC_OK;
}

int mmc_trial_atom(mmc_handle *h, int64_t i, const double r_new[3], mmc_trial_result *out)
{
    if (!h) return MMC_EINVAL;
    if (!h->has_atoms) FAIL(MMC_ESTATE, "no atomic system uploaded");
    if (i < 1 || i > h->At.n || !r_new || !out) FAIL(MMC_EINVAL, "bad arguments");
    AtomArgs &A = h->last_atom;
    A = AtomArgs{};
    A.i = (int)(i - 1); A.n_cfg = 2;
    A.blocks = std::max(1, std::min((h->At.n + ATOM_BLOCK - 1) / ATOM_BLOCK, 2 * h->sm_count));
    A.r_new[0] = r_new[0]; A.r_new[1] = r_new[1]; A.r_new[2] = r_new[2];
    int rc = launch_move_atom(h, A);
    if (rc) return rc;
    std::memset(out, 0, sizeof(*out));
    out->lj_old = h->h_out->lj_pot[0]; out->lj_vir_old = h->h_out->lj_vir[0];
    out->lj_new = h->h_out->lj_pot[1]; out->lj_vir_new = h->h_out->lj_vir[1];
    h->trial_pending = true; h->trial_kind = 2;
    h->cnt.trial_moves++;
    return MMC_OK;
}

int mmc_accept(mmc_handle *h)
{
    if (!h) return MMC_EINVAL;
    if (!h->trial_pending) FAIL(MMC_ESTATE, "mmc_accept without a pending trial move");
    { int rc = flush_pending(h); if (rc) return rc; }      // at most one accepted move rides along
    if (h->trial_kind == 1) {
        // the write-back itself is deferred: the next move launch carries it in its parameters
        const MoveArgs &A = h->last;
        h->state_version++;
        h->pend_kind = 1; h->pend_i = A.i; h->pend_ns = mol_of(h, A.i).y;
        std::memcpy(h->pend_com, A.com_new, sizeof(h->pend_com));
        std::memcpy(h->pend_site, A.site_new, sizeof(double) * 3 * h->pend_ns);
        if (h->trial_style == MMC_STYLE_EWALD && !h->last_overlap) h->cur ^= 1;   // main.jl:621 as a pointer swap
    } else {
        const AtomArgs &A = h->last_atom;
        h->pend_kind = 2; h->pend_i = A.i;
        std::memcpy(h->pend_com, A.r_new, sizeof(h->pend_com));
    }
    h->trial_pending = false; h->new_valid = false;
    h->cnt.commits++;
    return MMC_OK;
}

int mmc_reject(mmc_handle *h)
{
    if (!h) return MMC_EINVAL;
    if (!h->trial_pending) FAIL(MMC_ESTATE, "mmc_reject without a pending trial move");
    h->trial_pending = false; h->new_valid = false;    // main.jl:623-628: resident state was never touched
    return MMC_OK;
}

int mmc_set_intramolecular(mmc_handle *h, int32_t enabled)
{
    if (!h) return MMC_EINVAL;
    h->intramolecular = enabled ? 1 : 0;
    return MMC_OK;
}

int mmc_get_counters(mmc_handle *h, mmc_counters *out)
{
    if (!h || !out) return MMC_EINVAL;
    *out = h->cnt;
    return MMC_OK;
}

int mmc_set_timing(mmc_handle *h, int32_t enabled)
{
    if (!h) return MMC_EINVAL;
    h->tm.on = enabled == 2 ? 2 : (enabled != 0 ? 1 : 0);
    return MMC_OK;
}

int mmc_last_timings(mmc_handle *h, float *ms4)
{
    if (!h || !ms4) return MMC_EINVAL;
    for (int i = 0; i < 4; ++i) ms4[i] = h->tm.ms[i];
    return MMC_OK;
}

int mmc_last_eval_info(mmc_handle *h, int64_t *pairs_in_cutoff, int32_t *mode, int32_t *cells_per_dim,
                       int32_t *pair_kernel)
{
    if (h && pair_kernel) *pair_kernel = h->last_fast;
    if (!h) return MMC_EINVAL;
    if (pairs_in_cutoff) *pairs_in_cutoff = h->last_pairs;
    if (mode) *mode = h->last_mode;
    if (cells_per_dim) *cells_per_dim = h->last_ncd;
    return MMC_OK;
}

int mmc_debug_set(mmc_handle *h, const char *key, int64_t value)
{
    if (!h || !key) return MMC_EINVAL;
    const std::string k(key);
    if (k == "chain_cluster_atoms") { h->chain_cluster_atoms = (int)value; return MMC_OK; }
    if (k == "chain_cluster") { if (value < 1 || value > 8) FAIL(MMC_EINVAL, "chain_cluster must be 1..8"); h->chain_cluster = (int)value; return MMC_OK; }
    if (k == "host_mailbox") { h->host_mailbox = value != 0; return MMC_OK; }
    if (k == "dd_speculate") { h->dd_speculate = value != 0; return MMC_OK; }
    if (k == "com_allgather") { h->com_allgather = value != 0; return MMC_OK; }
    if (k == "overlap_rhok") { h->overlap_rhok = (int)value; return MMC_OK; }   // 0: one stream, 1: fork at the start, 2: fork after the gather
    if (k == "rhok_early_pct") { if (value < 0 || value > 90) FAIL(MMC_EINVAL, "rhok_early_pct must be 0..90"); h->rhok_early_pct = (int)value; return MMC_OK; }
    if (k == "rhok_kshard") { h->rhok_kshard = value != 0; return MMC_OK; }
    if (k == "host_windows") { if (value < 1 || value > 4) FAIL(MMC_EINVAL, "host_windows must be 1..4"); h->host_windows = (int)value; return MMC_OK; }
    if (k == "v7_ctas_per_sm") { if (value < 1 || value > 4) FAIL(MMC_EINVAL, "v7_ctas_per_sm must be 1..4"); h->v7_ctas_per_sm = (int)value; return MMC_OK; }
    if (k == "host_chunks") { if (value < 1 || value > 8) FAIL(MMC_EINVAL, "host_chunks must be 1..8"); h->host_chunks = (int)value; return MMC_OK; }
    if (k == "rhok_split") {
        if (value < 1 || value > 64) FAIL(MMC_EINVAL, "rhok_split must be 1..64");
        h->rhok_split = (int)value;
        return MMC_OK;
    }
    if (k == "pair_level") {                 // first pair kernel the fallback chain may use (0 v7, 1 fast, 2 general; 3: mmc_potential as Σ_i rows / 2 via k_move)
        if (value < 0 || value > 3) FAIL(MMC_EINVAL, "pair_level must be 0..3");
        h->pair_floor = (int)value; h->pair_level = (int)value;
        return MMC_OK;
    }
    FAIL(MMC_EINVAL, "unknown debug key");
}

}  // extern "C"
