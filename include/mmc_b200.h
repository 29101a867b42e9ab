/*
 * mmc_b200.h — C ABI of libmmc_b200.so, the B200-native (sm_100a CUDA, FP64) energy engine
 * that replaces the energy routines of BradenDKelly/MetropolisMonteCarlo.
 *
 * The reference is pure Julia and has no FFI boundary of its own; the boundary is the set of
 * Julia call sites in Loop()/potential() (Ewald/main.jl:491,502,557,567,581,408 and
 * Monatomic/mainMonatomic.jl:377,381).  Each entry point below names the reference function
 * it replaces (file:line into the reference tree).  Julia binds these with `ccall`
 * (see INTEGRATION.md and julia/MMCB200.jl); plain pointers and sizes only.
 *
 * Conventions (the reference's own):
 *   - molecule / atom indices are 1-based, first_atom/last_atom inclusive (Julia arrays);
 *   - coordinates are xyz-interleaved doubles (Vector{SVector{3,Float64}} memory);
 *   - eps/sig tables are column-major nt x nt (Julia Matrix); lengths in Angstrom, energies in K;
 *   - every function returns 0 on success or a negative mmc_status; physical overlap is a
 *     result flag, never an error; no exception or abort crosses the ABI.
 *   - one host thread per handle; every call returns after its scalar results are on the host.
 *
 * There is NO CPU fallback: every energy is computed by the CUDA kernels in
 * metropolismontecarlo_b200/csrc.  Without a CUDA device mmc_create fails with MMC_ECUDA.
 */
#ifndef MMC_B200_H
#define MMC_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMC_VERSION 100

typedef enum {
    MMC_OK = 0,
    MMC_EINVAL = -1,   /* bad argument / system not uploaded                         */
    MMC_ECUDA = -2,    /* CUDA runtime error (message in mmc_last_error)             */
    MMC_ENCCL = -3,    /* collective error                                           */
    MMC_ESTATE = -4    /* call out of order (accept without trial, ...)              */
} mmc_status;

/* coulombStyle / Wolf switches of Ewald/main.jl:74-75 */
typedef enum {
    MMC_STYLE_EWALD = 0,    /* real-space erfc + k-space + self  (energy.jl:946-1032)        */
    MMC_STYLE_WOLF = 1,     /* real-space erfc + Wolf constants   (energy.jl:864-943)        */
    MMC_STYLE_LJ_ONLY = 2,  /* molecules, no Coulomb                                         */
    MMC_STYLE_LJ_ATOMS = 3  /* monatomic LJ (Monatomic/mainMonatomic.jl:227-289)             */
} mmc_style;

typedef struct mmc_handle mmc_handle;

typedef struct {
    int32_t device;     /* CUDA device ordinal                                                */
    int32_t rank;       /* shard index for the full-energy path (0 when unsharded)            */
    int32_t world;      /* number of shards (1 = unsharded); each rank is one process, one GPU */
    int32_t sync_mode;  /* 0: host polls a mapped pinned result slot; 1: cudaStreamSynchronize */
    void *stream;       /* optional cudaStream_t to enqueue on (NULL: the handle's own)        */
} mmc_config;

/* Ewald/auxillary.jl:37-45 `Properties` as filled by potential(): energy, virial, coulomb;
 * plus the components the reference prints (energy.jl:979,1003,1012,1018). */
typedef struct {
    double energy, virial, coulomb;
    double lj, real, recip, self_;
    double wolf_const;      /* (prefactor - prefactor2)*factor, energy.jl:924-934 */
    int64_t overlaps;       /* molecules whose EwaldReal row hit the overlap rule */
    double intra;           /* intramolecular Ewald correction; 0 unless mmc_set_intramolecular is on (not in the reference) */
} mmc_properties;

/* result of one fused trial move: exactly the numbers Loop() gets from its five calls */
typedef struct {
    double lj_old, lj_vir_old;       /* LJ_poly_ΔU before the move   (main.jl:491)  */
    double lj_new, lj_vir_new;       /* LJ_poly_ΔU after the move    (main.jl:557)  */
    double qq_old, qq_vir_old;       /* EwaldShort before (factor included) (:501)  */
    double qq_new, qq_vir_new;       /* EwaldShort after             (main.jl:566)  */
    double d_recip;                  /* RecipMove, 0 if overlap or style != EWALD (:580-590) */
    int32_t overlap_old, overlap_new;
} mmc_trial_result;

/* ---- lifetime -------------------------------------------------------------------------- */
int mmc_create(const mmc_config *cfg, mmc_handle **out);
int mmc_destroy(mmc_handle *h);
const char *mmc_last_error(const mmc_handle *h);   /* h may be NULL: last create error */
int mmc_version(void);

/* ---- upload: replaces the in-place Julia state soa/moa/vdwTable ------------------------- */
/* Ewald/setup.jl:447-537 (MakeAtomArrays "kmc") + :546-673 (MakeTables) layouts, A.6.
 * COMs must lie in [0, box] (the reference's PBC keeps them there, boundaries.jl:16-26). */
int mmc_upload_system(mmc_handle *h, int64_t n_mol, int64_t n_sites,
                      const double *coords, const double *charge, const int64_t *atype,
                      const int64_t *first_atom, const int64_t *last_atom, const double *com,
                      int32_t n_types, const double *eps, const double *sig,
                      double box, double rc_lj, double rc_qq);
/* All positions of an uploaded system at once (pointer(soa.coords), pointer(moa.COM) after the caller changed
 * them itself): the bulk form of mmc_set_molecule (Ewald/main.jl:527,552).  Charges, types and topology stay; the
 * resident rho(k) is not rebuilt (mmc_recip_long / mmc_potential do that, as after mmc_upload_system). */
int mmc_upload_positions(mmc_handle *h, const double *coords, const double *com);
/* mmc_upload_positions + mmc_potential in one call with the host->device copies overlapped with the work that does
 * not need them yet (COMs first: cell binning; sites in chunks on a side stream, each fed to the rho(k) rebuild as it
 * lands; gather + pair kernel when the last chunk is in).  Host arrays in, Properties out.
 * On a sharded handle (world > 1, peer exchange set up; all ranks call it together with the same arrays) the evaluation is
 * domain-decomposed: a rank copies all COMs but only the blocks of the site array that hold molecules of its own slab of
 * the cell grid (+ one layer, + its share of the sites for rho(k)); afterwards ONLY those are current on that GPU, and
 * every entry point that needs the whole state returns MMC_ESTATE until mmc_upload_positions / mmc_upload_system.
 * mmc_last_host_bytes: bytes the last mmc_potential_host copied host -> device on this rank. */
int mmc_potential_host(mmc_handle *h, const double *coords, const double *com, int32_t style, mmc_properties *out);
int mmc_last_host_bytes(mmc_handle *h, int64_t *h2d_bytes);

/* Monatomic/mainMonatomic.jl:140-146 Requirements(r, eps, sig, box, r_cut) */
int mmc_upload_atoms(mmc_handle *h, int64_t n, const double *r, const double *eps_j,
                     const double *sig_j, double box, double r_cut);
/* read the resident state back (checkpointing; the reference mutates soa/moa in place) */
int mmc_download_system(mmc_handle *h, double *coords, double *com);
int mmc_download_atoms(mmc_handle *h, double *r);

/* ---- k-space setup: Ewald/ewalds.jl:45-103 PrepareEwaldVariables ------------------------ */
/* The tables belong to (kappa, box).  They survive mmc_upload_system / mmc_upload_positions; when a re-upload changes the
 * box, cfac is rebuilt for the new box with the resident kappa and the resident rho(k) is cleared (the reference rebuilds
 * both in PrepareEwaldVariables per box) — call mmc_ewald_prepare again for a new kappa (= alpha / box, main.jl:290). */
int mmc_ewald_prepare(mmc_handle *h, double kappa, int32_t nk, int32_t k_sq_max,
                      double factor, int32_t *nkvecs);
int mmc_get_kvectors(mmc_handle *h, int32_t *kxyz /* nkvecs x 3 */, double *cfac);
/* ewald.sumQExpOld / sumQExpNew (re,im interleaved); either pointer may be NULL */
int mmc_get_rhok(mmc_handle *h, double *sum_old, double *sum_new);

/* Opt-in (default off = the reference's semantics): add the intramolecular correction of the Ewald sum,
 *     E_intra = -factor * sum_molecules sum_{a<b in molecule} q_a q_b erf(kappa r_ab) / r_ab,
 * to the EWALD energies of mmc_potential*, mmc_potential_host and mmc_volume_trial (energy, coulomb, virial += E_intra/3 like the
 * other Coulomb terms, and Properties.intra).  The reference omits it (Ewald/energy.jl:1008-1021 adds the reciprocal and the
 * self term only); because kappa = alpha/L changes with the box, NPT volume moves need it for physically meaningful energies
 * (SURVEY 8-f4).  Rigid-body trial moves leave it unchanged, so the per-move entry points do not carry it; neither does
 * mmc_potential_host on a sharded handle (only a slab of the sites is resident there). */
int mmc_set_intramolecular(mmc_handle *h, int32_t enabled);

/* ---- per-function drop-ins (molecules) ------------------------------------------------- */
/* Ewald/energy.jl:209-290  LJ_poly_ΔU(i, moa, soa, vdwTable, r_cut, box) -> (4 pot, 24 vir/3) */
int mmc_lj_mol(mmc_handle *h, int64_t i, double *pot, double *vir);
/* Ewald/ewalds.jl:293-376  EwaldReal(i, moa, soa, ewald, r_cut, box) -> (pot un-scaled, overlap) */
int mmc_ewald_real(mmc_handle *h, int64_t i, double *pot, int32_t *overlap);
/* Ewald/ewalds.jl:892-910  EwaldShort(i, moa, soa, sim_props, ewald, box) -> (e, e/3, overlap) */
int mmc_ewald_short(mmc_handle *h, int64_t i, double *e, double *v, int32_t *overlap);
/* the writes at Ewald/main.jl:527,552 (trial) and :623-624 (restore) */
int mmc_set_molecule(mmc_handle *h, int64_t i, const double com[3], const double *sites);
/* Ewald/ewalds.jl:538-604  RecipLong -> un-scaled energy; fills both S(k) buffers */
int mmc_recip_long(mmc_handle *h, double *energy);
/* Ewald/ewalds.jl:718-826  RecipMove(box, ewald, r_old, r_new, q) -> energy*factor.
 * Device semantics: S_new = S_old + delta — identical to the reference's in-place "+=" wherever
 * S_new == S_old on entry, which main.jl:621,628 guarantee. */
int mmc_recip_move(mmc_handle *h, const double *r_old, const double *r_new, const double *q,
                   int32_t n, double *dE);
int mmc_recip_commit(mmc_handle *h);    /* Ewald/main.jl:621 (Old <- New; pointer swap)     */
int mmc_recip_rollback(mmc_handle *h);  /* Ewald/main.jl:628 (New <- Old; nothing to copy)  */
/* Ewald/ewalds.jl:829-833 EwaldSelf */
int mmc_ewald_self(mmc_handle *h, double *energy);

/* ---- per-function drop-ins (monatomic) ------------------------------------------------- */
/* Monatomic/mainMonatomic.jl:227-272 LJ_ΔU(i, system) */
int mmc_lj_atom(mmc_handle *h, int64_t i, double *pot, double *vir);
int mmc_set_atom(mmc_handle *h, int64_t i, const double r[3]);

/* ---- full-system energy: potential(...) ------------------------------------------------ */
/* Ewald/energy.jl:946-1032 (EWALD), :864-943 (WOLF), Monatomic/mainMonatomic.jl:275-289 (LJ_ATOMS).
 * EWALD also rebuilds S(k) into both buffers like RecipLong does. Unsharded handles only.
 * Any topology the reference's firstAtom/lastAtom tables can describe with <= 16 sites per molecule (Ewald/energy.jl:219-226):
 * rigid three-site water takes the k_pairs_v7 pipeline, other uniform molecules k_pairs_fast / k_pairs, and mixtures of
 * different molecules k_pairs on an evaluation copy padded to the largest molecule (LJ by the active type pairs). */
int mmc_potential(mmc_handle *h, int32_t style, mmc_properties *out);

/* LJ_poly_ΔU(i) (Ewald/energy.jl:209-290) and EwaldShort(i) (Ewald/ewalds.jl:892-910) for EVERY molecule i from one
 * evaluation — the rows that potential() sums at energy.jl:966-1001, kept per molecule.  Arrays of n_mol
 * entries in molecule order, any of them may be NULL: lj_pot[i] = 4·pot, lj_vir[i] = 24·vir/3, coul[i] = pot·factor
 * (0 with overlap[i] = 1 when the overlap rule of ewalds.jl:359 fires for molecule i; coul is 0 for LJ_ONLY).  Any
 * topology, unsharded handles.  Row sums are accumulated with FP64 atomics: equal to the per-i calls to ~1e-13 relative. */
int mmc_energy_all(mmc_handle *h, int32_t style, double *lj_pot, double *lj_vir, double *coul, int32_t *overlap);

/* sharded variant (world > 1): each rank evaluates its share of the molecule-pair work and of
 * the sites of S(k) and leaves MMC_NPARTIAL doubles in device memory at d_partials; the caller
 * sums that buffer across ranks (NCCL all-reduce over NVLink) and calls finalize on every rank.
 * world == 1 works too (partial == total). */
int mmc_partial_count(mmc_handle *h, int64_t *n_doubles);
int mmc_potential_partial(mmc_handle *h, int32_t style, double *d_partials);
int mmc_potential_finalize(mmc_handle *h, int32_t style, const double *d_partials,
                           mmc_properties *out);
/* (If a rank's pair kernel declined the state, or molecules overlap — ewalds.jl:359-360 zeroes whole rows, which needs a
 * molecule's complete neighbourhood — every rank sees it in the summed vector and evaluates the whole (replicated) system
 * itself on the general path: same Properties everywhere, slower, no retry protocol for the caller.) */

/* ---- the same exchange over NVLink peer memory instead of a library all-reduce ------------ */
/* Every rank owns an exchange buffer that its peers map through CUDA IPC (one process per GPU on one node).
 * mmc_peer_export: allocate it and return its 64-byte cudaIpcMemHandle_t; pass every rank's handle to every
 * other rank (any transport: torch.distributed all_gather_object, MPI, a file) and mmc_peer_import them.
 * mmc_potential_sharded (all ranks together) = partial sums -> each rank stores its vector into slot `rank` of
 * every rank's buffer + an epoch flag (k_peer_push) -> wait for the `world` flags, add the slots in rank order
 * (k_peer_sum) -> finalise.  Same Properties on every rank, bit-identical totals; no collective library call.
 * _begin/_end are the two halves (begin returns after the launches); mmc_peer_import_ptr/mmc_peer_buffer are
 * the same-process form used to emulate ranks in one process (tests).
 * Export AFTER mmc_upload_system: the buffer then also holds a staging area for the centres of mass, and mmc_potential_host
 * on the sharded handles copies only 1/world of the COM array over each rank's PCIe link — the other slices are read from
 * the peers' staging areas over NVLink (k_repack_com_gather).  Exported earlier, every rank copies all COMs itself.
 * (Ranks emulated in ONE process — mmc_peer_import_ptr — always copy all COMs themselves: the all-gather's kernels spin on the
 * peers' flags early in the call, which deadlocks streams that share one CUDA context.  One process per GPU has no such coupling.) */
int mmc_peer_export(mmc_handle *h, void *ipc_handle_64_bytes);
int mmc_peer_import(mmc_handle *h, int32_t peer_rank, const void *ipc_handle_64_bytes);
int mmc_peer_import_ptr(mmc_handle *h, int32_t peer_rank, void *peer_buffer);
int mmc_peer_buffer(mmc_handle *h, void **buffer);
int mmc_potential_sharded_begin(mmc_handle *h, int32_t style);
int mmc_potential_sharded_end(mmc_handle *h, mmc_properties *out);
int mmc_potential_sharded(mmc_handle *h, int32_t style, mmc_properties *out);

/* ---- fused fast path: one launch, one host sync per trial move ------------------------- */
/* Same numbers as mmc_lj_mol + mmc_ewald_short (old), mmc_set_molecule, mmc_lj_mol +
 * mmc_ewald_short (new), mmc_recip_move — Ewald/main.jl:491-590. The resident state is not
 * changed until mmc_accept. */
int mmc_trial_move(mmc_handle *h, int64_t i, const double com_new[3], const double *sites_new,
                   int32_t style, mmc_trial_result *out);
int mmc_accept(mmc_handle *h);   /* Ewald/main.jl:599-621 state part */
int mmc_reject(mmc_handle *h);   /* Ewald/main.jl:622-629 state part */
/* Monatomic/mainMonatomic.jl:377-381: LJ_ΔU(old), move, LJ_ΔU(new) */
int mmc_trial_atom(mmc_handle *h, int64_t i, const double r_new[3], mmc_trial_result *out);

/* ---- volume move: Ewald/volumeChange.jl:50-147 (spec in a docstring) -------------------- */
/* COMs scaled by f = box_new/box, sites rigidly shifted, kappa_new (= alpha/box_new in the
 * reference's convention, main.jl:290) and cfac rebuilt, full energy at box_new; S(k) of the
 * trial box kept aside until accept. */
int mmc_volume_trial(mmc_handle *h, double box_new, double kappa_new, int32_t style,
                     mmc_properties *out);
int mmc_volume_accept(mmc_handle *h);
int mmc_volume_reject(mmc_handle *h);

/* ---- host-side stand-in for the Julia driver ------------------------------------------- */
/* Ewald/main.jl:487-651 Loop(), restated in C++ above the entry points above (one
 * mmc_trial_move + mmc_accept/mmc_reject per move), fed by a caller-supplied stream of
 * uniforms in the reference's draw order.  This is what Julia would do
 * through ccall; it exists because Julia is not available where this library is tested. */
typedef struct {
    double temperature, dr_max, dphi_max;
    double p_trans, p_rot;          /* cumulative thresholds, main.jl:97-107 */
    int32_t style;                  /* MMC_STYLE_EWALD / WOLF / LJ_ONLY      */
    int32_t adjust;                 /* Adjust!/Adjust_rot! after every sweep */
} mmc_loop_params;

typedef struct {
    int64_t n_moves, n_accepted, n_overlap, uniforms_used;
    int64_t trans_attempt, trans_accept, rot_attempt, rot_accept;
    double dr_max, dphi_max;
    double total_energy, total_virial;
} mmc_loop_stats;

/* com/quat/db are host arrays owned by the caller and updated in place (moa.COM, moa.quat);
 * returns 0, or 1 if the uniform stream ran dry, 2 on a quaternion-norm error, <0 on mmc errors */
int mmc_loop_run(mmc_handle *h, const mmc_loop_params *p, double *com, double *quat,
                 const double *db, const double *uniforms, int64_t n_uniforms, int64_t n_moves,
                 double e0, double v0, uint8_t *accepted, double *delta, mmc_loop_stats *stats);
/* The same block of moves evaluated on the device in ONE launch (csrc/kernels_chain.cuh): the state of a
 * small system (uniform topology, <= 4 sites per molecule, fits one SM's shared memory: up to ~1600
 * three-site molecules) stays on chip, the uniform stream is consumed in the same order, and the
 * accept/reject record, deltas, statistics, com/quat and the library's resident state (sites, COMs,
 * rho(k)) come back when the block is done.  What a Julia Loop() would call once per sweep/block
 * instead of five times per move (Ewald/main.jl:487-651).  Same arguments and return codes as
 * mmc_loop_run; MMC_EINVAL when the system does not qualify. */
int mmc_loop_run_device(mmc_handle *h, const mmc_loop_params *p, double *com, double *quat,
                        const double *db, const double *uniforms, int64_t n_uniforms, int64_t n_moves,
                        double e0, double v0, uint8_t *accepted, double *delta, mmc_loop_stats *stats);
/* Monatomic/mainMonatomic.jl:373-413 */
int mmc_loop_run_atoms(mmc_handle *h, double temperature, double dr_max, double *r,
                       const double *uniforms, int64_t n_uniforms, int64_t n_moves,
                       double e0, double v0, uint8_t *accepted, double *delta,
                       mmc_loop_stats *stats);

/* mmc_loop_run_atoms for a block of moves in one launch (csrc/kernels_chain.cuh k_chain_atoms: the atoms are
 * sliced over the CTAs of a 16-SM (or 8-SM) cluster, FP32 distance gate on chip, FP64 evaluation of what passes). */
int mmc_loop_run_atoms_device(mmc_handle *h, double temperature, double dr_max, double *r,
                              const double *uniforms, int64_t n_uniforms, int64_t n_moves,
                              double e0, double v0, uint8_t *accepted, double *delta,
                              mmc_loop_stats *stats);

/* The reference's random stream without Julia: the first `n` values, after skipping `skip`, that
 * `Random.seed!(seed); rand()` yields in the Julia the reference targets (1.x <= 1.6, global
 * MersenneTwister = dSFMT-19937 seeded by init_by_array(make_seed(seed))); `rand(Float64,3)`
 * (Ewald/auxillary.jl:99, quaternions.jl:64) consumes three consecutive values of the same
 * stream.  seed = 11234 is the reference's own (Ewald/main.jl:36, Monatomic/mainMonatomic.jl:15).
 * Host code; feeds mmc_loop_run / mmc_loop_run_atoms. */
int mmc_julia_rand(uint64_t seed, int64_t skip, double *out, int64_t n);

/* Page-lock caller-owned host arrays (e.g. the storage behind soa.coords and moa.COM, Ewald/main.jl:306-330) so that
 * mmc_upload_positions / mmc_potential_host DMA them at full PCIe rate and really asynchronously; pageable memory is
 * staged by the driver at a fraction of that.  Register once after the arrays are allocated, unregister before they
 * are freed or resized.  Thin wrappers over cudaHostRegister / cudaHostUnregister; no handle needed. */
int mmc_host_register(void *ptr, size_t bytes);
int mmc_host_unregister(void *ptr);

/* ---- instrumentation ------------------------------------------------------------------- */
typedef struct {
    int64_t kernel_launches;     /* kernels of this library launched on this handle */
    int64_t trial_moves, commits, overlap_events, full_energy_evals;
} mmc_counters;
int mmc_get_counters(mmc_handle *h, mmc_counters *out);
/* device times [ms] of the most recent full-energy evaluation, measured with CUDA events on the
 * handle's stream when timing is enabled: ms4[0] pair kernel, ms4[1] rho(k) rebuild kernel,
 * ms4[2] binning + gather, ms4[3] whole evaluation up to the result copy.  enabled = 1: every phase (the events and the
 * stream synchronisations they need cost ~35 us per evaluation); enabled = 2: the pair kernel only (two event records, no
 * synchronisation: what bench.py leaves on inside its timed region) */
int mmc_set_timing(mmc_handle *h, int32_t enabled);
int mmc_last_timings(mmc_handle *h, float *ms4);
/* what the last full-energy evaluation did: molecule pairs inside the cutoff (summed over ranks
 * after finalize), path taken (0 = cell list, 1 = tile pairs, 2 = per-molecule rows), cells per
 * box edge, and the pair kernel used (7 = k_pairs_v7, 64/128 = k_pairs_fast tile, 0 = k_pairs) */
int mmc_last_eval_info(mmc_handle *h, int64_t *pairs_in_cutoff, int32_t *mode, int32_t *cells_per_dim,
                       int32_t *pair_kernel);
/* test/profiling knobs. key "pair_level": first full-energy pair kernel the fallback chain may use
 * (0 = k_pairs_v7, 1 = k_pairs_fast, 2 = general k_pairs); results are identical
 * within rounding, only speed differs. */
int mmc_debug_set(mmc_handle *h, const char *key, int64_t value);
/* FP64 DFMA-chain microbenchmark: measured FP64 FMA throughput of the device [TFLOP/s] */
int mmc_measure_fp64_peak(mmc_handle *h, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* MMC_B200_H */
