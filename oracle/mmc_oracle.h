/*
 * mmc_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C, Float64, line-by-line restatement of the energy hot path of
 * BradenDKelly/MetropolisMonteCarlo (a pure-Julia code).  Every function cites
 * the reference file:line it follows (paths are relative to /root/reference).
 *
 * Who may use this: tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs.  The product library
 * (metropolismontecarlo_b200/libmmc_b200.so) never links, loads or calls it.
 *
 * Pinning status: the reference cannot be executed here (no Julia in the image,
 * no manifest for its packages), so there is no oracle/_ref.  The oracle is
 * pinned against (1) the reference's own known answers in Ewald/tests.jl
 * (test_LJ: -0.381860031778575 and -0.320336594278575; the two-triangle LJ sum),
 * (2) the external NIST SPC/E reference energies for the four configurations the
 * reference bundles (Ewald/spce_sample_config_periodic{1..4}.txt): E_fourier and
 * E_self to all six published digits, and (3) an independent numpy restatement
 * (oracle/numpy_ref.py).  Real-space Ewald / Wolf totals have no in-repo known
 * answer in the reference: for those, parity is "unpinned" beyond (3).
 *
 * erfc: the reference calls SpecialFunctions.erfc (un-vendored, un-pinned Julia
 * package; forwards to openlibm erfc for Float64).  Here: glibc erfc (<1 ulp).
 *
 * Build: gcc -O2 -ffp-contract=off (no FMA contraction: Julia does not contract).
 */
#ifndef MMC_ORACLE_H
#define MMC_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Ewald/auxillary.jl:37-45 (mutable struct Properties) + diagnostic parts */
typedef struct {
    double energy, virial, coulomb;      /* the fields potential() fills       */
    double lj, real, recip, self_;       /* components (printed by reference)  */
    double wolf_const;                   /* (prefactor-prefactor2)*factor       */
    int64_t overlaps;                    /* molecules whose EwaldReal overlapped */
} ora_properties;

/* A system in the reference's own (Julia) memory layout, SURVEY Appendix A.6 */
typedef struct {
    int64_t n_mol, n_sites;
    double *coords;          /* soa.coords  n_sites x 3, xyz interleaved        */
    const double *charge;    /* soa.charge  n_sites                             */
    const int64_t *atype;    /* soa.atype   n_sites, 1-based                    */
    const int64_t *first_atom, *last_atom; /* moa.firstAtom/lastAtom, 1-based incl */
    double *com;             /* moa.COM     n_mol x 3                           */
    int64_t n_types;
    const double *eps, *sig; /* vdwTable.eps_ij / sig_ij, column-major nt x nt   */
} ora_system;

/* Ewald/ewalds.jl:9-19 (mutable struct EWALD) */
typedef struct {
    double kappa;
    int64_t nk, k_sq_max, nkvecs;
    int32_t *kxyz;           /* nkvecs x 3                                      */
    double *cfac;            /* nkvecs                                          */
    double *sum_old, *sum_new; /* nkvecs x 2 (re,im interleaved = ComplexF64)   */
    double factor;
} ora_ewald;

double ora_vector1D(double c1, double c2, double box);
void   ora_PBC(double v[3], double box);

void ora_LJ_poly_dU(int64_t i, const ora_system *s, double r_cut, double box,
                    double *pot, double *vir);
void ora_EwaldReal(int64_t i, const ora_system *s, double kappa, double r_cut,
                   double box, double *pot, int *overlap);
void ora_EwaldShort(int64_t i, const ora_system *s, const ora_ewald *ew,
                    double qq_rcut, double box, double *e, double *v, int *overlap);

int64_t ora_count_kvecs(int64_t nk, int64_t k_sq_max);
void ora_PrepareEwaldVariables(ora_ewald *ew, double box);
double ora_RecipLong(ora_ewald *ew, int64_t n, const double *r, const double *q, double box);
double ora_RecipMove(double box, ora_ewald *ew, int64_t n, const double *r_old,
                     const double *r_new, const double *q);
double ora_EwaldSelf(const ora_ewald *ew, int64_t n, const double *q);
double ora_EwaldIntraBox(const ora_system *s, double kappa, double factor, double box);
void ora_recip_commit(ora_ewald *ew);    /* Ewald/main.jl:621 */
void ora_recip_rollback(ora_ewald *ew);  /* Ewald/main.jl:628 */

void ora_potential_ewald(const ora_system *s, ora_ewald *ew, double lj_rcut,
                         double qq_rcut, double box, int n_threads, ora_properties *out);
void ora_potential_wolf(const ora_system *s, const ora_ewald *ew, double lj_rcut,
                        double qq_rcut, double box, int n_threads, ora_properties *out);
/* bounded-sample timing helper: per-molecule rows for i in [i0,i1) only */
void ora_potential_rows(const ora_system *s, double kappa, double lj_rcut, double qq_rcut,
                        double box, int64_t i0, int64_t i1, int n_threads,
                        double *lj_sum, double *vir_sum, double *real_sum, int64_t *overlaps);

/* Monatomic/mainMonatomic.jl */
void ora_LJ_dU_atom(int64_t i, int64_t n, const double *r, const double *eps,
                    const double *sig, double box, double r_cut, double *pot, double *vir);
void ora_potential_atoms(int64_t n, const double *r, const double *eps, const double *sig,
                         double box, double r_cut, int n_threads, double *energy, double *virial);

/* Ewald/volumeChange.jl:50-147 (spec in a docstring) */
void ora_volume_scale(ora_system *s, double box_old, double box_new);

/* ---- driver restatement (Ewald/main.jl:487-651) for trajectory parity ---- */
typedef struct {
    double temperature, dr_max, dphi_max, p_trans, p_rot; /* thresholds main.jl:97-107 */
    double lj_rcut, qq_rcut, box;
    int style;               /* 0 = ewald, 1 = wolf, 2 = LJ only (no coulomb)   */
    int adjust;              /* 1: Adjust!/Adjust_rot! after every sweep         */
} ora_loop_params;

typedef struct {
    int64_t n_moves, n_accepted, n_overlap, uniforms_used;
    int64_t trans_attempt, trans_accept, rot_attempt, rot_accept;
    double dr_max, dphi_max;   /* after adaptation                              */
    double total_energy, total_virial; /* running totals (Sum of accepted deltas) */
} ora_loop_stats;

/* returns 0 ok, 1 ran out of uniforms, 2 quaternion normalisation error */
int ora_loop(ora_system *s, ora_ewald *ew, const double *db /* n_sites x 3 body frame */,
             double *quat /* n_mol x 4 */, const ora_loop_params *p,
             const double *uniforms, int64_t n_uniforms, int64_t n_moves,
             double e0, double v0,
             uint8_t *accepted /* n_moves */, double *delta /* n_moves */,
             ora_loop_stats *stats);

/* Monatomic/mainMonatomic.jl:373-413 */
int ora_loop_atoms(int64_t n, double *r, const double *eps, const double *sig, double box,
                   double r_cut, double temperature, double dr_max,
                   const double *uniforms, int64_t n_uniforms, int64_t n_moves,
                   double e0, double v0, uint8_t *accepted, double *delta,
                   ora_loop_stats *stats);

void ora_q_to_a(const double q[4], double a[9] /* row-major */);
void ora_MATMUL(const double a[9], const double db[3], double out[3]);
void ora_quatmul(const double a[4], const double b[4], double c[4]);

#ifdef __cplusplus
}
#endif
#endif
