// mmc_handle.h — internal: the state behind an mmc_handle and the helpers shared by the translation units of
// libmmc_b200.so (mmc_api.cu: lifetime, upload, per-move entry points; mmc_eval.cu: full-energy evaluation, sharding,
// volume moves; mmc_loop.cu: the Loop() stand-in and the block-of-moves kernels).  Not part of the public ABI.
#pragma once
#include "../../include/mmc_b200.h"
#include "erf_poly.h"
#include "kernels_move.cuh"
#include "kernels_peer.cuh"
#include "mmc_common.cuh"

#include <string>
#include <utility>
#include <vector>

// host-side fold of one move launch: exactly the numbers Loop() gets from its calls
struct MoveOut { double lj_pot[2], lj_vir[2], qq[2], d_recip; int overlap[2]; };
struct Timers { cudaEvent_t ev[8]; int on = 0; float ms[4] = {0, 0, 0, 0}; bool full() const { return on == 1; } };   // on: 0 off, 1 every phase, 2 pair kernel only

struct mmc_handle {
    mmc_config cfg{};
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t side = nullptr;         // the ρ(k) rebuild of a full evaluation runs here, beside binning + pair kernel
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_sites = nullptr;
    cudaStream_t copy = nullptr;         // mmc_potential_host: host->device chunks (nothing else); repack follows on `side`,
    cudaStream_t rk = nullptr;           //                     the chunks' ρ(k) partials on `rk` (they cannot co-reside with the pair kernel
                                         //                     and must not hold back the "chunk resident" events behind them)
    cudaEvent_t ev_rk = nullptr;
    cudaEvent_t ev_chunk[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_copy[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int host_windows = 3;                // mmc_potential_host, one GPU: home-cell windows evaluated as their sites arrive (1: wait for all sites)
    int win_ncd = 0, win_n = 0;
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
    struct UploadResult { int info[4]; double qs[2]; int win_need[4]; } *h_up = nullptr;   // pinned (mapped: kernels write it directly)
    UploadResult *d_up = nullptr;     // its device alias
    bool uniform = false;        // every molecule: same site count, same type sequence, packed
    int US = 0;                  // uniform sites per molecule
    bool mixed = false;          // molecules of different size / type sequence: the cell path evaluates a copy padded to ES slots
    int ES = 0;                  // site slots per molecule of the evaluation copy (US, or max_sites when mixed)
    int *d_atype_pad = nullptr;       // mixed: [n_mol x ES] zeros — the type array the overlap rows' k_move indexes on the padded copy
    signed char *d_stype = nullptr;   // mixed: LJ type per slot of the evaluation copy (−1 = padding)
    size_t ssite_cap = 0;        // capacity of d_ssite in sites
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
    int2 *d_kcombos = nullptr; int n_kcombos = 0, k_zt = 4;   // k_rhok_big: (pair tile, kz tile) combos that hold k-vectors; kz per tile
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
    unsigned long long peer_stage_cap = 0;        // COM staging capacity (molecules) every rank has: the smallest one
    unsigned long long com_epoch = 0;
    bool peer_same_process = false;               // peers imported by pointer (ranks emulated in one process)
    int host_mailbox = 1;                         // mmc_debug_set "host_mailbox": small results through mapped memory (1) or device->host copies (0)
    int dd_speculate = 1;                         // mmc_debug_set "dd_speculate": 0 = no speculative site-block copy
    int com_allgather = 1;                        // mmc_debug_set "com_allgather": 0 = every rank copies all COMs itself
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
    double *d_permol = nullptr, *d_permol_out = nullptr;   // mmc_energy_all: per-molecule rows (evaluation order) and the scaled output arrays
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
    int use_rhok_v2 = 1;
    int rhok_early_pct = 0;      // overlap_rhok == 1: share of the sites whose ρ(k) partials run beside binning + gather
    int rhok_kshard = 0;         // sharded ρ(k) rebuild of a large k-set: 1 = split the k-vectors (combo groups) per rank, 0 = the sites
    int pair_level = 0;          // first pair kernel allowed: 0 k_pairs_v7, 1 k_pairs_fast, 2 general k_pairs
                                 // (raised when a kernel declines the state)
    // ---- k_pairs_v7 path (kernels_pairs_v7.cuh): ghost-extended cell grid in a fixed-capacity layout
    int *d7_flags = nullptr;     // [8]: max |site-COM| (2 ints) | #overlap | err | max cell population | unit ticket | tail ticket | spare
    int *d7_count = nullptr, *d7_bucket = nullptr, *d7_ecount = nullptr;
    double *d7_rows = nullptr;
    float4 *d7_gf = nullptr;
    double4 *d7_unit_partial = nullptr, *d7_block_sums = nullptr;
    bool sharded_fused = false;                     // the pending sharded evaluation finishes inside its own tail kernel
    int *d7_order = nullptr;                        // [3 ncd³] k_order7: draw order of the units, expensive first
    int d7_ncd = 0;
    size_t d7_partial_cap = 0;
    double *h7_res = nullptr, *d7_res = nullptr;     // mapped pinned result slot [MMC_NSCAL + 1]: scalars + sequence number
    unsigned long long res_seq = 0;
    unsigned long long state_version = 1;           // bumped whenever resident positions change
    unsigned long long bin_version = 0;             // state the buckets were built from (0: none)
    int bin_ncd = 0, bin_world = 0;
    double2 *d7_rhok_scratch = nullptr;             // [16][nkvecs] slice sums of the ρ(k) partials (k_eval_tail)
    size_t d7_scratch_cap = 0;
    int *d7_range = nullptr;                        // [0,1] = {0, ncd³}; [2 .. 2+world] = home-cell boundaries of the ranks (k_partition7)
    int v7_ctas_per_sm = 4;
    unsigned char *d7_need_host = nullptr;                    // device alias of h7_need (mapped pinned)
    unsigned char *d7_need = nullptr, *h7_need = nullptr;     // domain-decomposed host evaluation: molecule blocks this rank reads
    int need_cap = 0;
    std::vector<unsigned char> need_prev;     // the blocks the previous domain-decomposed call needed (copied speculatively by the next one)
    int need_prev_world = 0;
    bool partial_resident = false;                  // after it only those blocks' sites are current on this GPU
    long long last_h2d_bytes = 0;                   // bytes the last mmc_potential_host moved host -> device on this rank
    int intramolecular = 0;                         // mmc_set_intramolecular
    double *d_intra = nullptr;                      // [256 partials | result]
    bool v7_left_for_overlap = false;               // k_pairs_v7 handed the state over because molecules overlapped
    int v7_rhok_blocks = 0;                         // CTAs of the last ρ(k) partial launch
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
inline void dfree(T *&p)
{
    if (p) cudaFree(p);
    p = nullptr;
}

inline int2 mol_of(const mmc_handle *h, int64_t i0)
{
    return h->uniform ? make_int2((int)(i0 * h->US), h->US) : h->h_mol[i0];
}

namespace mmc_detail {
// mmc_api.cu
int ensure_vec(mmc_handle *h);
int move_tiles(const mmc_handle *h);
int flush_pending(mmc_handle *h);
int launch_move_on(mmc_handle *h, const DevSystem &sys, MoveArgs &A, const ErfPoly &poly, bool carry_commit);
int style_check(mmc_handle *h, int style, bool need_full_state = true);
int intra_energy(mmc_handle *h, double kappa, double *e_unscaled);
void fill_cfac(const std::vector<int32_t> &kxyz, double kappa, double box, std::vector<double> &cfac);
void get_erf_poly(mmc_handle *h, double kappa, double r2_max, ErfPoly &P);
// mmc_eval.cu
void eval_set_attributes();
int rhok_launch(mmc_handle *h, const double4 *site, int s_begin, int s_end, double box, double2 *out, cudaStream_t st = nullptr,
                int block0 = 0, int *nb_out = nullptr, int cap_blocks = 0, const double4 *com = nullptr, double f = 1.0);
}  // namespace mmc_detail
