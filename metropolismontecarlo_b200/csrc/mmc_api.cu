// mmc_api.cu — C ABI of libmmc_b200.so (include/mmc_b200.h) on top of the sm_100a kernels.
// No CPU fallback: every energy returned here was computed by a kernel in this directory.
#include "../../include/mmc_b200.h"
#include "kernels_move.cuh"
#include "kernels_chain.cuh"
#include "kernels_pairs.cuh"
#include "kernels_pairs_v3.cuh"
#include "kernels_pairs_v4.cuh"
#include "kernels_pairs_v5.cuh"
#include "kernels_pairs_v6.cuh"
#include "kernels_peer.cuh"
#include "kernels_recip.cuh"
#include "kernels_upload.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <immintrin.h>

namespace {

std::string g_create_error;

// host-side fold of one move launch: exactly the numbers Loop() gets from its calls
struct MoveOut { double lj_pot[2], lj_vir[2], qq[2], d_recip; int overlap[2]; };

struct Timers { cudaEvent_t ev[8]; bool on = false; float ms[4] = {0, 0, 0, 0}; };

}  // namespace

struct mmc_handle {
    mmc_config cfg{};
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t side = nullptr;         // the ρ(k) rebuild of a full evaluation runs here, beside binning + pair kernel
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_sites = nullptr;
    cudaStream_t copy = nullptr;         // mmc_potential_host: host->device chunks + repack; ρ(k) partials follow on `side`
    cudaEvent_t ev_chunk[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int host_windows = 3;                // ... and z-layer windows the pair evaluation is cut into while they arrive (1: wait for all sites)
    int *d_winneed = nullptr;
    cudaEvent_t ev_copy[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int host_chunks = 6;                 // mmc_potential_host: pieces the site array is uploaded in (mmc_debug_set "host_chunks", 1..8)
    int overlap_rhok = 1;                // mmc_debug_set "overlap_rhok": 0 = everything on one stream
    std::string err;

    // ---- molecular system
    bool has_system = false;
    DevSystem S{};
    std::vector<int2> h_mol;     // host mirror of S.mol
    unsigned char *d_raw = nullptr;     // device staging for the caller's arrays in their own layout
    size_t raw_bytes = 0;
    int cap_mol = 0, cap_sites = 0;     // sizes the resident buffers were allocated for
    int *d_info = nullptr;
    double2 *d_qpart = nullptr;
    struct UploadResult { int info[4]; double qs[2]; } *h_up = nullptr;   // pinned
    bool uniform = false;        // every molecule: same site count, same type sequence, packed
    int US = 0;                  // uniform sites per molecule
    std::vector<LJActive> lj;
    LJActive *d_lj = nullptr;
    int2 *d_mol_uniform = nullptr;
    double sum_q = 0.0, sum_q2 = 0.0;
    double *d_qsums = nullptr;

    // ---- ewald
    bool has_ewald = false;
    int k_sq_max = 0;
    std::vector<int32_t> kxyz;
    std::vector<double> cfac;
    int cur = 0;                 // index of the Old ρ(k) buffer
    bool new_valid = false;
    double2 *d_rhok_trial = nullptr;
    int2 *d_kpairs = nullptr;
    int *d_kindex = nullptr;
    int n_kpairs = 0;
    double *d_cfac_trial = nullptr;
    std::vector<double> cfac_trial;

    // ---- move scratch
    MoveScratch W{};
    MoveSlot *h_slots = nullptr;  // mapped pinned: one slot per CTA of a move launch
    int max_slots = 0;
    MoveOut mout{};               // host-side fold of the slots of the last move launch
    MoveOut *h_out = &mout;
    ErfPoly move_poly{};          // erf polynomial of the resident box for the per-move kernels
    // ---- sharded evaluation over peer memory (mmc_peer_*, mmc_potential_sharded_begin/end)
    double *d_peer_buf = nullptr;                 // [2][world][peer_nvec_cap] doubles, then [2][world] flags
    size_t peer_nvec_cap = 0;
    void *peer_base[MMC_PEER_MAX] = {nullptr};    // mapped exchange buffers of all ranks (own: d_peer_buf)
    bool peer_opened[MMC_PEER_MAX] = {false};     // opened through cudaIpcOpenMemHandle (to be closed)
    int peer_ready = 0;                           // number of imported ranks
    unsigned long long peer_epoch = 0;
    double *d_peer_total = nullptr;               // summed vector
    int *h_peer_status = nullptr;                 // mapped pinned host word written by k_peer_sum (no extra copy to read it)
    int *d_peer_status = nullptr;                 // its device alias
    bool sharded_pending = false;
    int sharded_style = 0;
    // ---- device-resident block of moves (mmc_loop_run_device)
    unsigned char *d_chain = nullptr;   // [uniforms | quat | db | delta | out | accepted]
    int chain_cluster = 8;              // CTAs (SMs) per cluster for mmc_loop_run_device; 1 = single-CTA kernel
    int chain_cluster_atoms = 16;       // ... for mmc_loop_run_atoms_device (16 = non-portable cluster size, falls back to 8)
    size_t chain_bytes = 0;
    int pend_kind = 0;            // accepted move not yet written to HBM: 0 none, 1 molecule, 2 atom
    int pend_i = 0, pend_ns = 0;
    double pend_com[3] = {0, 0, 0};
    double pend_site[3 * MMC_MAX_SITES] = {0};
    unsigned long long seq = 0;
    bool trial_pending = false;
    int trial_kind = 0;          // 1 molecule, 2 atom
    int trial_style = 0;
    MoveArgs last{};
    AtomArgs last_atom{};
    bool last_overlap = false;

    // ---- full-energy scratch
    int *d_cell_of = nullptr, *d_count = nullptr, *d_start = nullptr, *d_fill = nullptr, *d_perm = nullptr;
    int ncell_cap = 0;
    double4 *d_scom = nullptr, *d_ssite = nullptr;
    double *d_mrows = nullptr;   // k_pairs_v6: cell-sorted state as rows of 12 doubles
    double *d_permol = nullptr, *d_permol_out = nullptr;   // mmc_energy_all: per-molecule rows (evaluation order) and the scaled output arrays
    float4 *d_gf = nullptr;      //             and cell-local float COMs
    double4 *d_pair_partial = nullptr;
    int pair_grid = 0;
    unsigned int *d_ovl = nullptr, *d_novl = nullptr;
    double *d_maxdev = nullptr;
    double2 *d_rhok_partial = nullptr;
    int rhok_grid_cap = 0;
    double *d_vec = nullptr;     // internal partial-sum vector (MMC_NSCAL + 2*NK doubles)
    double *h_vec = nullptr;     // pinned
    int last_mode = -1;          // 0 cells, 1 tiles, 2 rows
    int last_ncd = 0;
    int max_cell_cached = -1;    // largest cell population seen at the last binning (-1: unknown)
    int *d_maxcount = nullptr;
    int *d_flags = nullptr;      // [maxdev(2) | novl | errflag | maxcount | 3 spare | count(ncell) | fill(ncell)]
    int4 *d_units = nullptr;
    int4 *d_slots = nullptr;
    long long slots_cap = 0;
    int use_rhok_v2 = 1;
    int v3_ctas_per_sm = 2;
    int v6_ctas_per_sm = 5;
    int use_v3 = 1;              // 0 disables the v3 pair kernel (A/B testing)
    int pair_level = 0;          // first pair kernel allowed: 0 k_pairs_v6, 1 k_pairs_v5, 2 k_pairs_v4, 3 k_pairs_v3, 4 k_pairs_fast, 5 general k_pairs
                                 // (raised when a kernel declines the state)
    int v6_dynamic = 1;          // k_pairs_v6 draws units by ticket (0: static round-robin deal)
    double4 *d_unit_partial = nullptr;
    size_t unit_partial_cap = 0;
    int rhok_split = 1;          // ρ(k) rebuild: CTAs per resident slot (short CTAs let higher-priority kernels in between)
    int pair_floor = 0;          // lowest level the chain may start from (mmc_debug_set "pair_level": A/B tests)
    bool uniform_q = false;      // every molecule carries the charges of molecule 1 (per site index)
    double q_site[MMC_MAX_SITES] = {0};
    long long units_cap = 0;
    unsigned int *d_errflag = nullptr;
    std::vector<std::pair<double, ErfPoly>> poly_cache;
    int last_fast = 0;           // tile size of the fast pair kernel used last (0: general kernel)
    long long last_pairs = 0;    // molecule pairs inside the cutoff in the last evaluation (all ranks)

    // ---- volume trial
    bool vol_pending = false;
    double vol_box = 0, vol_kappa = 0, vol_f = 1;
    int vol_style = 0;

    // ---- atoms
    bool has_atoms = false;
    DevAtoms At{};
    double2 *d_rows = nullptr;
    double *d_atoms_out = nullptr;

    int sm_count = 148;
    mmc_counters cnt{};
    Timers tm;
};

namespace {

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                      \
            return MMC_ECUDA;                                                                 \
        }                                                                                     \
    } while (0)

#define FAIL(code, msg)                                                                       \
    do {                                                                                      \
        h->err = (msg);                                                                       \
        return (code);                                                                        \
    } while (0)

#define LAUNCH_CHECK()                                                                        \
    do {                                                                                      \
        h->cnt.kernel_launches++;                                                             \
        cudaError_t e_ = cudaGetLastError();                                                  \
        if (e_ != cudaSuccess) {                                                              \
            h->err = std::string("kernel launch: ") + cudaGetErrorString(e_);                 \
            return MMC_ECUDA;                                                                 \
        }                                                                                     \
    } while (0)

template <typename T>
void dfree(T *&p)
{
    if (p) cudaFree(p);
    p = nullptr;
}

void free_system(mmc_handle *h)
{
    dfree(h->S.site); dfree(h->S.com); dfree(h->S.mol); dfree(h->S.atype);
    dfree(h->d_lj); dfree(h->d_mol_uniform); dfree(h->d_qsums); dfree(h->d_raw); dfree(h->d_info); dfree(h->d_qpart);
    h->raw_bytes = 0; h->cap_mol = 0; h->cap_sites = 0;
    dfree(h->d_cell_of); dfree(h->d_start); dfree(h->d_perm); dfree(h->d_flags);
    h->d_count = h->d_fill = nullptr; h->d_maxcount = nullptr; h->d_novl = h->d_errflag = nullptr; h->d_maxdev = nullptr;
    dfree(h->d_mrows); dfree(h->d_gf); dfree(h->d_chain); h->chain_bytes = 0;
    dfree(h->d_winneed); dfree(h->d_permol); dfree(h->d_permol_out); dfree(h->d_unit_partial); h->unit_partial_cap = 0;
    dfree(h->d_scom); dfree(h->d_ssite); dfree(h->d_pair_partial); dfree(h->d_ovl);
    dfree(h->d_rhok_partial); dfree(h->d_units); dfree(h->d_slots);
    h->units_cap = 0; h->slots_cap = 0;
    h->max_cell_cached = -1;
    h->has_system = false;
}

void free_ewald(mmc_handle *h)
{
    dfree(h->S.kvec); dfree(h->S.cfac); dfree(h->S.rhok[0]); dfree(h->S.rhok[1]);
    dfree(h->d_rhok_trial); dfree(h->d_cfac_trial); dfree(h->d_vec); dfree(h->d_kpairs); dfree(h->d_kindex);
    if (h->h_vec) cudaFreeHost(h->h_vec);
    h->h_vec = nullptr;
    h->has_ewald = false;
}

void free_atoms(mmc_handle *h)
{
    dfree(h->At.r); dfree(h->At.es); dfree(h->d_rows); dfree(h->d_atoms_out);
    h->has_atoms = false;
}

int ensure_vec(mmc_handle *h)
{
    if (h->d_vec) return MMC_OK;
    const size_t n = MMC_NSCAL + 2 * (size_t)std::max(h->S.nkvecs, 1);
    CK(cudaMalloc(&h->d_vec, n * sizeof(double)));
    CK(cudaMemsetAsync(h->d_vec, 0, n * sizeof(double), h->stream));
    CK(cudaHostAlloc(&h->h_vec, n * sizeof(double), cudaHostAllocDefault));
    return MMC_OK;
}

inline int2 mol_of(const mmc_handle *h, int64_t i0)
{
    return h->uniform ? make_int2((int)(i0 * h->US), h->US) : h->h_mol[i0];
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
    { const int2 mi = (&sys == &h->S) ? mol_of(h, A.i) : make_int2(A.i * h->US, h->US); A.i_first = mi.x; A.i_ns = mi.y; }
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

// cfac table of Ewald/ewalds.jl:52,78-83 (host-side table, uploaded once per box size)
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

// out == nullptr: partials only, written from block `block0` on (the caller reduces all blocks later); *nb_out = blocks used
int rhok_launch(mmc_handle *h, const double4 *site, int s_begin, int s_end, double box, double2 *out, cudaStream_t st = nullptr,
                int block0 = 0, int *nb_out = nullptr, int cap_blocks = 0)
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
constexpr size_t V3_SMEM = (V3_ACAP * 4 + V3_BCAP * 4) * sizeof(double4) + PAIR_WARPS * V3_QCAP * sizeof(unsigned);
void launch_pairs_v3(int deg, int grid, cudaStream_t st, const PairArgs &P, const int4 *slots)
{
#define X(D) if (deg == D) { k_pairs_v3<D><<<grid, PAIR_BLOCK, V3_SMEM, st>>>(P, slots); return; }
    MMC_FOR_POS_DEGS(X)
#undef X
}
void launch_pairs_v4(int deg, int grid, cudaStream_t st, const PairArgs &P, const int4 *slots)
{
#define X(D) if (deg == D) { k_pairs_v4<D><<<grid, V4_BLOCK, V4_SMEM, st>>>(P, slots); return; }
    MMC_FOR_POS_DEGS(X)
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
void pairs_fast_set_attributes()
{
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
#define X(D) cudaFuncSetAttribute(k_pairs_v4<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V4_SMEM);
    MMC_FOR_POS_DEGS(X)
#undef X
#define X(D) cudaFuncSetAttribute(k_pairs_v3<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V3_SMEM);
    MMC_FOR_POS_DEGS(X)
#undef X
#define X(D)                                                                                                   \
    cudaFuncSetAttribute(k_pairs_fast<3, 64, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);      \
    cudaFuncSetAttribute(k_pairs_fast<3, 128, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    MMC_FOR_DEGS(X)
#undef X
}

void bind_flags(mmc_handle *h, int ncell)
{
    h->d_maxdev = reinterpret_cast<double *>(h->d_flags);
    h->d_novl = reinterpret_cast<unsigned *>(h->d_flags + 2);
    h->d_errflag = reinterpret_cast<unsigned *>(h->d_flags + 3);
    h->d_maxcount = h->d_flags + 4;
    h->d_count = h->d_flags + 8;
    h->d_fill = h->d_flags + 8 + ncell;
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

// A pair kernel declined the state.  The water kernels (levels 0-3: v6, v5, v4, v3) share their
// preconditions (cell population <= 64, site reach inside the reference's +100 window), so a decline
// by one of them goes straight to k_pairs_fast; after that one level at a time.
bool escalate_pair_level(mmc_handle *h)
{
    h->pair_level = h->pair_level < 4 ? 4 : h->pair_level + 1;
    return h->pair_level <= 5;
}

// Leaves this rank's partial sums in d_vec: [0] Σlj_pot [1] Σlj_vir [2] Σcoul [3] #overlap
// [MMC_NSCAL ..) ρ(k) partial (re,im).
int eval_partials(mmc_handle *h, int style, const EvalCtx &E, double *d_vec)
{
    const bool force_general = h->pair_level >= 5 || E.per_mol != nullptr;
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
    // v3 serves water-like molecules: 3 sites, LJ only on site pair (0,0), equal cut-offs, Coulomb on, polynomial erf
    const bool water = cells && US == 3 && !force_general && h->pair_level <= 3 && max_cell <= V3_ACAP && want_qq &&
                       S.rc_lj == S.rc_qq && P.ep.deg > 0 && h->lj.size() == 1 && h->lj[0].a == 0 && h->lj[0].b == 0;
    // v4 additionally needs identical per-site charges (they become launch constants)
    const bool v6 = water && h->pair_level == 0 && h->uniform_q && want_rows;
    const bool v5 = water && !v6 && h->pair_level <= 1 && h->uniform_q;
    const bool v4 = water && !v5 && !v6 && h->pair_level <= 2 && h->uniform_q;
    const bool v3 = water && !v4 && !v5 && !v6 && h->use_v3;
    if (nwin > 1 && !v6) {       // another kernel serves this state: it wants the whole gathered copy
        if (E.wait_sites) CK(cudaStreamWaitEvent(h->stream, E.wait_sites, 0));
        k_gather<<<gm, tb, 0, h->stream>>>(G); LAUNCH_CHECK();
        nwin = 1;
    }
    const int tile = (US == 3 && !force_general && !v3 && !v4 && !v5 && !v6) ? (max_cell <= 64 ? 64 : (max_cell <= 128 ? 128 : 0)) : 0;
    if (v3 || v4 || v5 || v6) n_units = (long long)V3_GROUPS * ncd * ncd * ncd;
    if (v4 || v5 || v6) {
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
    if (v3 || v4 || v5 || v6) {
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
        } else if (v5) {
            grid = (int)std::max(1LL, std::min<long long>(4 * h->sm_count, my_units));
            launch_pairs_v5(v5_deg, v5_direct, grid, h->stream, P, h->d_slots);
        } else if (v4) {
            grid = (int)std::max(1LL, std::min<long long>(4 * h->sm_count, my_units));
            launch_pairs_v4(P.ep.deg, grid, h->stream, P, h->d_slots);
        } else {
            grid = (int)std::max(1LL, std::min<long long>(h->v3_ctas_per_sm * h->sm_count, my_units));
            launch_pairs_v3(P.ep.deg, grid, h->stream, P, h->d_slots);
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
    h->last_fast = v6 ? 6 : (v5 ? 5 : (v4 ? 4 : (v3 ? 3 : tile)));
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
        int rc = launch_move(h, A);
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

int style_check(mmc_handle *h, int style)
{
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
    for (auto &ev : h->ev_chunk)
        if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return fail("event", e);
    for (auto &ev : h->ev_copy)
        if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return fail("event", e);
    cudaFuncSetAttribute(k_pairs<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(k_pairs<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(k_pairs<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(k_pairs<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    pairs_fast_set_attributes();
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
        h->pair_grid = 5 * h->sm_count;
        CK(cudaMalloc(&h->d_pair_partial, sizeof(double4) * h->pair_grid));
        CK(cudaMalloc(&h->d_ovl, sizeof(unsigned) * n_mol));
        h->ncell_cap = 0; h->rhok_grid_cap = 0;
        h->cap_mol = (int)n_mol; h->cap_sites = (int)n_sites;
        if (!h->h_up) CK(cudaHostAlloc((void **)&h->h_up, sizeof(*h->h_up), cudaHostAllocDefault));
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
    h->pair_level = h->pair_floor;
    if (h->pend_kind == 1) h->pend_kind = 0;
    if (h->has_ewald) get_erf_poly(h, S.kappa, rc_qq * rc_qq + 100, h->move_poly);
    h->has_system = true;
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
    CK(cudaMemcpyAsync(h->h_up->info, h->d_info, 4 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (h->h_up->info[0] & REPACK_COM_OUTSIDE) FAIL(MMC_EINVAL, "a COM lies outside [0, box] (the reference's PBC keeps COMs inside)");
    h->pair_level = h->pair_floor;
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
    if (!(kappa > 0) || nk < 1 || nk > MMC_MAX_NK || k_sq_max < 2) FAIL(MMC_EINVAL, "bad Ewald parameters (1 <= nk <= 8)");
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
        std::vector<int2> kp;
        for (int kx = 0; kx <= nk; ++kx)
            for (int ky = 0; ky <= nk; ++ky)
                if (used[(size_t)kx * (nk + 1) + ky]) kp.push_back(make_int2(kx, ky));
        h->n_kpairs = (int)kp.size();
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
    h->trial_pending = false;
    return MMC_OK;
}

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

int mmc_recip_move(mmc_handle *h, const double *r_old, const double *r_new, const double *q, int32_t n, double *dE)
{
    if (!h) return MMC_EINVAL;
    if (!h->has_system || !h->has_ewald) FAIL(MMC_ESTATE, "system and Ewald tables required");
    if (!r_old || !r_new || !q || n < 1 || n > MMC_MAX_SITES) FAIL(MMC_EINVAL, "bad arguments (1 <= n <= 16)");
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

int mmc_trial_move(mmc_handle *h, int64_t i, const double com_new[3], const double *sites_new, int32_t style,
                   mmc_trial_result *out)
{
    if (!h) return MMC_EINVAL;
    int rc = check_mol_index(h, i);
    if (rc) return rc;
    if ((rc = style_check(h, style))) return rc;
    if (style == MMC_STYLE_LJ_ATOMS || !com_new || !sites_new || !out) FAIL(MMC_EINVAL, "bad arguments");
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

int mmc_get_counters(mmc_handle *h, mmc_counters *out)
{
    if (!h || !out) return MMC_EINVAL;
    *out = h->cnt;
    return MMC_OK;
}

int mmc_set_timing(mmc_handle *h, int32_t enabled)
{
    if (!h) return MMC_EINVAL;
    h->tm.on = enabled != 0;
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
    if (k == "chain_cluster") { if (value < 1 || value > CHAINC_MAXC) FAIL(MMC_EINVAL, "chain_cluster must be 1..8"); h->chain_cluster = (int)value; return MMC_OK; }
    if (k == "overlap_rhok") { h->overlap_rhok = (int)value; return MMC_OK; }   // 0: one stream, 1: fork at the start, 2: fork after the gather
    if (k == "v6_ctas_per_sm") { if (value < 1 || value > 5) FAIL(MMC_EINVAL, "v6_ctas_per_sm must be 1..5"); h->v6_ctas_per_sm = (int)value; return MMC_OK; }
    if (k == "host_windows") { if (value < 1 || value > 4) FAIL(MMC_EINVAL, "host_windows must be 1..4"); h->host_windows = (int)value; return MMC_OK; }
    if (k == "host_chunks") { if (value < 1 || value > 8) FAIL(MMC_EINVAL, "host_chunks must be 1..8"); h->host_chunks = (int)value; return MMC_OK; }
    if (k == "v6_dynamic") { h->v6_dynamic = value != 0; return MMC_OK; }
    if (k == "rhok_split") {
        if (value < 1 || value > 64) FAIL(MMC_EINVAL, "rhok_split must be 1..64");
        h->rhok_split = (int)value;
        return MMC_OK;
    }
    if (k == "pair_level") {                 // first pair kernel the fallback chain may use (0 v6 .. 5 general)
        if (value < 0 || value > 5) FAIL(MMC_EINVAL, "pair_level must be 0..5");
        h->pair_floor = (int)value; h->pair_level = (int)value;
        return MMC_OK;
    }
    FAIL(MMC_EINVAL, "unknown debug key");
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

#include "mmc_driver.inl"
