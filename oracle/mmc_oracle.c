/*
 * mmc_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 * See mmc_oracle.h for scope, pinning status and who may call this.
 *
 * Plain C restatement, in Float64 and in the reference's own evaluation order, of
 * the Julia energy routines of BradenDKelly/MetropolisMonteCarlo.  Citations are
 * file:line into /root/reference.  Nothing here is optimised: the O(N) scan per
 * call, the double-counted O(N^2) potential() and the k-outer RecipLong loop are
 * kept because this file is also the timed "port" CPU baseline.
 */
#include "mmc_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---------------------------------------------------------------- geometry */

/* Ewald/boundaries.jl:8-14 (duplicate at Ewald/ewalds.jl:30-38,
 * Monatomic/mainMonatomic.jl:217-225) */
double ora_vector1D(double c1, double c2, double box)
{
    if (c1 < c2)
        return (c2 - c1) < (c1 - c2 + box) ? (c2 - c1) : (c2 - c1 - box);
    else
        return (c1 - c2) < (c2 - c1 + box) ? (c2 - c1) : (c2 - c1 + box);
}

/* Ewald/boundaries.jl:16-26 */
void ora_PBC(double v[3], double box)
{
    for (int k = 0; k < 3; ++k) {
        if (v[k] > box) v[k] -= box;
        if (v[k] < 0)   v[k] += box;
    }
}

/* ------------------------------------------------- single-molecule energies */

/* Ewald/energy.jl:209-290  LJ_poly_ΔU(i, moa, soa, vdwTable, r_cut, box)
 * i is 1-based like the reference. */
void ora_LJ_poly_dU(int64_t i, const ora_system *s, double r_cut, double box,
                    double *pot_out, double *vir_out)
{
    const double *ri = s->com + 3 * (i - 1);
    const int64_t startAtom = s->first_atom[i - 1], endAtom = s->last_atom[i - 1];
    const double diameter = 0;
    const double rm_cut_box = (r_cut + diameter);
    const double rm_cut_box_sq = rm_cut_box * rm_cut_box;
    const double r_cut_sq = r_cut * r_cut;
    const int64_t nt = s->n_types;
    double pot = 0.0, vir = 0.0;

    for (int64_t j = 1; j <= s->n_mol; ++j) {
        if (j == i) continue;
        const double *rj = s->com + 3 * (j - 1);
        double rij[3];
        for (int k = 0; k < 3; ++k) rij[k] = ora_vector1D(ri[k], rj[k], box);
        double rij2 = rij[0] * rij[0] + rij[1] * rij[1] + rij[2] * rij[2];
        if (rij2 < rm_cut_box_sq) {
            for (int64_t a = startAtom; a <= endAtom; ++a) {
                const double *ra = s->coords + 3 * (a - 1);
                for (int64_t b = s->first_atom[j - 1]; b <= s->last_atom[j - 1]; ++b) {
                    const double *rb = s->coords + 3 * (b - 1);
                    double rab[3];
                    for (int k = 0; k < 3; ++k) rab[k] = ora_vector1D(ra[k], rb[k], box);
                    double rab2 = rab[0] * rab[0] + rab[1] * rab[1] + rab[2] * rab[2];
                    /* table.ϵᵢⱼ[moli_type[a], list_type[b]] — column-major */
                    int64_t ta = s->atype[a - 1], tb = s->atype[b - 1];
                    double eij = s->eps[(ta - 1) + (tb - 1) * nt];
                    if (rab2 < (r_cut_sq + 100) && eij > 0.001) {
                        double sij = s->sig[(ta - 1) + (tb - 1) * nt];
                        double s2 = sij * sij / rab2;      /* σᵢⱼ^2 / rab²  */
                        double s6 = s2 * s2 * s2;          /* σ²^3          */
                        double s12 = s6 * s6;              /* σ⁶^2          */
                        pot += eij * (s12 - s6);
                        double virab = eij * (2.0 * s12 - s6);
                        double fab[3];
                        for (int k = 0; k < 3; ++k) fab[k] = rab[k] * virab * s2;
                        vir += rij[0] * fab[0] + rij[1] * fab[1] + rij[2] * fab[2];
                    }
                }
            }
        }
    }
    *pot_out = pot * 4;
    *vir_out = vir * 24 / 3.0;
}

/* Ewald/ewalds.jl:293-376  EwaldReal(chosenOne, moa, soa, ewald, r_cut, box) */
void ora_EwaldReal(int64_t i, const ora_system *s, double kappa, double r_cut,
                   double box, double *pot_out, int *overlap_out)
{
    const double *ri = s->com + 3 * (i - 1);
    const int64_t start_a = s->first_atom[i - 1], end_a = s->last_atom[i - 1];
    const double diameter = 0;
    const double rm_cut_box = (r_cut + diameter);
    const double rm_cut_box_sq = rm_cut_box * rm_cut_box;
    const double r_cut_sq = r_cut * r_cut;
    const double ovr = 0.5;
    double pot = 0.0;

    for (int64_t j = 1; j <= s->n_mol; ++j) {
        if (j == i) continue;
        const double *rj = s->com + 3 * (j - 1);
        double rij[3];
        for (int k = 0; k < 3; ++k) rij[k] = ora_vector1D(ri[k], rj[k], box);
        double rij2 = rij[0] * rij[0] + rij[1] * rij[1] + rij[2] * rij[2];
        if (rij2 < rm_cut_box_sq) {
            for (int64_t a = start_a; a <= end_a; ++a) {
                const double *ra = s->coords + 3 * (a - 1);
                for (int64_t b = s->first_atom[j - 1]; b <= s->last_atom[j - 1]; ++b) {
                    const double *rb = s->coords + 3 * (b - 1);
                    double rab[3];
                    for (int k = 0; k < 3; ++k) rab[k] = ora_vector1D(ra[k], rb[k], box);
                    double rab2 = rab[0] * rab[0] + rab[1] * rab[1] + rab[2] * rab[2];
                    if ((rab2 < ovr) && (s->charge[a - 1] * s->charge[b - 1] < 0)) {
                        *pot_out = 0.0;          /* return 0.0, true  (ewalds.jl:360) */
                        *overlap_out = 1;
                        return;
                    } else if (rab2 < (r_cut_sq + 100)) {
                        double rab_mag = sqrt(rab2);
                        pot += s->charge[a - 1] * s->charge[b - 1] * erfc(kappa * rab_mag) / rab_mag;
                    } else {
                        pot += 0.0;
                    }
                }
            }
        }
    }
    *pot_out = pot;
    *overlap_out = 0;
}

/* Ewald/ewalds.jl:892-910  EwaldShort(i, moa, soa, sim_props, ewald, box) */
void ora_EwaldShort(int64_t i, const ora_system *s, const ora_ewald *ew,
                    double qq_rcut, double box, double *e, double *v, int *overlap)
{
    double partial_e = 0.0, partial_v = 0.0, realEwald;
    ora_EwaldReal(i, s, ew->kappa, qq_rcut, box, &realEwald, overlap);
    realEwald *= ew->factor;
    partial_e += realEwald;
    partial_v += (realEwald / 3);
    *e = partial_e;
    *v = partial_v;
}

/* ------------------------------------------------------------------ k-space */

/* Ewald/ewalds.jl:54-65 (count pass) */
int64_t ora_count_kvecs(int64_t nk, int64_t k_sq_max)
{
    int64_t n = 0;
    for (int64_t kx = 0; kx <= nk; ++kx)
        for (int64_t ky = -nk; ky <= nk; ++ky)
            for (int64_t kz = -nk; kz <= nk; ++kz) {
                int64_t k_sq = kx * kx + ky * ky + kz * kz;
                if ((k_sq < k_sq_max) && (k_sq != 0)) n += 1;
            }
    return n;
}

/* Ewald/ewalds.jl:45-103  PrepareEwaldVariables(ewald, boxSize).
 * ew->kappa/nk/k_sq_max/factor must be set; kxyz, cfac, sum_old, sum_new must
 * have room for ora_count_kvecs() entries.  The reference's
 * "@assert k_sq_max == 27" is not restated (parameter check, not arithmetic). */
void ora_PrepareEwaldVariables(ora_ewald *ew, double box)
{
    const double kappa = ew->kappa;
    const int64_t nk = ew->nk, k_sq_max = ew->k_sq_max;
    const double b = 1.0 / 4.0 / kappa / kappa / box / box;
    const double twopi = 2.0 * M_PI;
    const double twopi_sq = twopi * twopi;
    int64_t NKVECS = 0;
    for (int64_t kx = 0; kx <= nk; ++kx)
        for (int64_t ky = -nk; ky <= nk; ++ky)
            for (int64_t kz = -nk; kz <= nk; ++kz) {
                int64_t k_sq = kx * kx + ky * ky + kz * kz;
                if ((k_sq < k_sq_max) && (k_sq != 0)) {
                    ew->kxyz[3 * NKVECS + 0] = (int32_t)kx;
                    ew->kxyz[3 * NKVECS + 1] = (int32_t)ky;
                    ew->kxyz[3 * NKVECS + 2] = (int32_t)kz;
                    double kr_sq = twopi_sq * (double)k_sq;
                    ew->cfac[NKVECS] = twopi * exp(-b * kr_sq) / kr_sq / box;
                    if (kx > 0) ew->cfac[NKVECS] = ew->cfac[NKVECS] * 2.0;
                    NKVECS += 1;
                }
            }
    ew->nkvecs = NKVECS;
    memset(ew->sum_old, 0, sizeof(double) * 2 * NKVECS);
    memset(ew->sum_new, 0, sizeof(double) * 2 * NKVECS);
}

typedef struct { double re, im; } cplx;
/* Julia Base complex multiply: Complex(a.re*b.re - a.im*b.im, a.re*b.im + a.im*b.re) */
static inline cplx cmul(cplx a, cplx b)
{
    cplx c = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re };
    return c;
}
static inline cplx cconj(cplx a) { cplx c = { a.re, -a.im }; return c; }
static inline cplx rmul(double q, cplx a) { cplx c = { q * a.re, q * a.im }; return c; }

/* builds the per-site e^{ikx}, e^{iky}, e^{ikz} tables of ewalds.jl:556-585.
 * ex: n x (nk+1) indexed [j*(nk+1)+k], ey/ez: n x (2nk+1) indexed [j*(2nk+1)+(k+nk)] */
static void build_eik(int64_t n, int64_t nk, const double *r, double L,
                      cplx *ex, cplx *ey, cplx *ez)
{
    const double twopi = 2.0 * M_PI;
    const int64_t wx = nk + 1, wy = 2 * nk + 1;
    const cplx one = { 1.0, 0.0 };
    for (int64_t j = 0; j < n; ++j) {
        ex[j * wx + 0] = one;
        ey[j * wy + nk] = one;
        ez[j * wy + nk] = one;
        if (nk >= 1) {
            cplx e1;
            e1.re = cos(twopi * (r[3 * j + 0]) / L); e1.im = sin(twopi * (r[3 * j + 0]) / L);
            ex[j * wx + 1] = e1;
            e1.re = cos(twopi * (r[3 * j + 1]) / L); e1.im = sin(twopi * (r[3 * j + 1]) / L);
            ey[j * wy + nk + 1] = e1;
            e1.re = cos(twopi * (r[3 * j + 2]) / L); e1.im = sin(twopi * (r[3 * j + 2]) / L);
            ez[j * wy + nk + 1] = e1;
            ey[j * wy + nk - 1] = cconj(ey[j * wy + nk + 1]);
            ez[j * wy + nk - 1] = cconj(ez[j * wy + nk + 1]);
        }
    }
    for (int64_t k = 2; k <= nk; ++k)
        for (int64_t j = 0; j < n; ++j) {
            ex[j * wx + k] = cmul(ex[j * wx + k - 1], ex[j * wx + 1]);
            ey[j * wy + nk + k] = cmul(ey[j * wy + nk + k - 1], ey[j * wy + nk + 1]);
            ez[j * wy + nk + k] = cmul(ez[j * wy + nk + k - 1], ez[j * wy + nk + 1]);
            ey[j * wy + nk - k] = cconj(ey[j * wy + nk + k]);
            ez[j * wy + nk - k] = cconj(ez[j * wy + nk + k]);
        }
}

/* Ewald/ewalds.jl:538-604  RecipLong(ewald, r, q, box) -> un-scaled energy;
 * stores S(k) into both sumQExpNew and sumQExpOld (ewalds.jl:600-601). */
double ora_RecipLong(ora_ewald *ew, int64_t n, const double *r, const double *q, double box)
{
    const int64_t nk = ew->nk, wx = nk + 1, wy = 2 * nk + 1;
    cplx *ex = (cplx *)malloc(sizeof(cplx) * n * wx);
    cplx *ey = (cplx *)malloc(sizeof(cplx) * n * wy);
    cplx *ez = (cplx *)malloc(sizeof(cplx) * n * wy);
    build_eik(n, nk, r, box, ex, ey, ez);
    double energy = 0.0;
    for (int64_t i = 0; i < ew->nkvecs; ++i) {          /* k outer, site inner */
        const int64_t kx = ew->kxyz[3 * i], ky = ew->kxyz[3 * i + 1], kz = ew->kxyz[3 * i + 2];
        cplx term = { 0.0, 0.0 };
        for (int64_t l = 0; l < n; ++l) {
            cplx t = cmul(cmul(rmul(q[l], ex[l * wx + kx]), ey[l * wy + nk + ky]), ez[l * wy + nk + kz]);
            term.re += t.re;
            term.im += t.im;
        }
        energy += ew->cfac[i] * (term.re * term.re - (-term.im) * term.im); /* real(conj(t)*t) */
        ew->sum_new[2 * i] = term.re; ew->sum_new[2 * i + 1] = term.im;
        ew->sum_old[2 * i] = term.re; ew->sum_old[2 * i + 1] = term.im;
    }
    free(ex); free(ey); free(ez);
    return energy;
}

/* Ewald/ewalds.jl:718-826  RecipMove(box, ewalds, r_old, r_new, q) -> energy*factor.
 * sumQExpNew is updated IN PLACE (+=) exactly like the reference (ewalds.jl:804-813);
 * the reference asserts n==3, nk==5, k_sq_max==27 — parameter checks, not restated. */
double ora_RecipMove(double box, ora_ewald *ew, int64_t n, const double *r_old,
                     const double *r_new, const double *q)
{
    const int64_t nk = ew->nk, wx = nk + 1, wy = 2 * nk + 1;
    cplx *buf = (cplx *)malloc(sizeof(cplx) * n * (wx + 2 * wy) * 2);
    cplx *exn = buf, *eyn = exn + n * wx, *ezn = eyn + n * wy;
    cplx *exo = ezn + n * wy, *eyo = exo + n * wx, *ezo = eyo + n * wy;
    build_eik(n, nk, r_new, box, exn, eyn, ezn);
    build_eik(n, nk, r_old, box, exo, eyo, ezo);
    double energy = 0.0;
    for (int64_t i = 0; i < ew->nkvecs; ++i) {
        const int64_t kx = ew->kxyz[3 * i], ky = ew->kxyz[3 * i + 1], kz = ew->kxyz[3 * i + 2];
        for (int64_t l = 0; l < n; ++l) {
            cplx tn = cmul(cmul(exn[l * wx + kx], eyn[l * wy + nk + ky]), ezn[l * wy + nk + kz]);
            cplx to = cmul(cmul(exo[l * wx + kx], eyo[l * wy + nk + ky]), ezo[l * wy + nk + kz]);
            cplx d = { tn.re - to.re, tn.im - to.im };
            ew->sum_new[2 * i]     += q[l] * d.re;
            ew->sum_new[2 * i + 1] += q[l] * d.im;
        }
        double nr = ew->sum_new[2 * i], ni = ew->sum_new[2 * i + 1];
        double orr = ew->sum_old[2 * i], oi = ew->sum_old[2 * i + 1];
        energy += ew->cfac[i] * ((nr * nr - (-ni) * ni) - (orr * orr - (-oi) * oi));
    }
    free(buf);
    return energy * ew->factor;
}

/* Ewald/ewalds.jl:829-833  EwaldSelf: -kappa * sum(q.^2) / sqrt(pi) * factor */
double ora_EwaldSelf(const ora_ewald *ew, int64_t n, const double *q)
{
    double s = 0.0;
    for (int64_t l = 0; l < n; ++l) s += q[l] * q[l];
    return -ew->kappa * s / sqrt(M_PI) * ew->factor;
}

/* NOT in the reference (Ewald/energy.jl:1008-1021 adds recip + self and nothing else): the intramolecular correction of the
 * Ewald sum, E_intra = -factor * sum_molecules sum_{a<b in molecule} q_a q_b erf(kappa r_ab) / r_ab, which removes the
 * interaction of a site with the screening clouds of its own molecule's sites that the k-space term contains.  Twin of the
 * engine's opt-in mmc_set_intramolecular; r_ab by the reference's own minimum image (vector1D), so a molecule
 * stored across the periodic boundary (the NIST files) is handled like one stored whole. */
double ora_EwaldIntraBox(const ora_system *s, double kappa, double factor, double box)
{
    double tot = 0.0;
    for (int64_t m = 0; m < s->n_mol; ++m) {
        double e = 0.0;
        for (int64_t a = s->first_atom[m] - 1; a < s->last_atom[m]; ++a)
            for (int64_t b = a + 1; b < s->last_atom[m]; ++b) {
                const double dx = ora_vector1D(s->coords[3 * a], s->coords[3 * b], box), dy = ora_vector1D(s->coords[3 * a + 1], s->coords[3 * b + 1], box),
                             dz = ora_vector1D(s->coords[3 * a + 2], s->coords[3 * b + 2], box);
                const double r = sqrt(dx * dx + dy * dy + dz * dz);
                e += s->charge[a] * s->charge[b] * erf(kappa * r) / r;
            }
        tot += e;
    }
    return -tot * factor;
}

/* Ewald/main.jl:621  ewald.sumQExpOld = [item for item in ewald.sumQExpNew] */
void ora_recip_commit(ora_ewald *ew)
{ memcpy(ew->sum_old, ew->sum_new, sizeof(double) * 2 * ew->nkvecs); }
/* Ewald/main.jl:628  ewald.sumQExpNew = [item for item in ewald.sumQExpOld] */
void ora_recip_rollback(ora_ewald *ew)
{ memcpy(ew->sum_new, ew->sum_old, sizeof(double) * 2 * ew->nkvecs); }

/* ------------------------------------------------------------ total energies */

/* per-molecule rows for i in [i0,i1) (0-based half-open), summed in index order.
 * This is the body of the two O(N^2) loops at energy.jl:972-977 and :991-1001.
 * With n_threads>1 rows are computed in parallel and then summed serially in the
 * reference's order, so the result is bit-identical to the serial loop. */
void ora_potential_rows(const ora_system *s, double kappa, double lj_rcut, double qq_rcut,
                        double box, int64_t i0, int64_t i1, int n_threads,
                        double *lj_sum, double *vir_sum, double *real_sum, int64_t *overlaps)
{
    const int64_t m = i1 - i0;
    double *e = (double *)malloc(sizeof(double) * m * 3);
    int *ov = (int *)malloc(sizeof(int) * m);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 8) num_threads(n_threads > 0 ? n_threads : 1)
#endif
    for (int64_t t = 0; t < m; ++t) {
        ora_LJ_poly_dU(i0 + t + 1, s, lj_rcut, box, &e[3 * t], &e[3 * t + 1]);
        if (kappa >= 0.0)
            ora_EwaldReal(i0 + t + 1, s, kappa, qq_rcut, box, &e[3 * t + 2], &ov[t]);
        else { e[3 * t + 2] = 0.0; ov[t] = 0; }
    }
    double a = 0.0, b = 0.0, c = 0.0; int64_t no = 0;
    for (int64_t t = 0; t < m; ++t) { a += e[3 * t]; b += e[3 * t + 1]; c += e[3 * t + 2]; no += ov[t]; }
    *lj_sum = a; *vir_sum = b; *real_sum = c; *overlaps = no;
    free(e); free(ov);
    (void)n_threads;
}

/* Ewald/energy.jl:946-1032  potential(moa, soa, tot, ewalds, vdwTable, sim_props, "ewald")
 * (the reference passes the global totProps.qq_rcut to EwaldReal, energy.jl:994) */
void ora_potential_ewald(const ora_system *s, ora_ewald *ew, double lj_rcut,
                         double qq_rcut, double box, int n_threads, ora_properties *out)
{
    double lj, vir, totReal; int64_t nov;
    ora_potential_rows(s, ew->kappa, lj_rcut, qq_rcut, box, 0, s->n_mol, n_threads,
                       &lj, &vir, &totReal, &nov);
    memset(out, 0, sizeof(*out));
    out->energy = lj / 2;
    out->virial = vir / 2;
    out->lj = lj / 2;
    totReal *= ew->factor / 2;
    out->energy += totReal;
    out->coulomb += totReal;
    out->virial += totReal / 3.0;
    out->real = totReal;
    double recipEnergy = ora_RecipLong(ew, s->n_sites, s->coords, s->charge, box);
    recipEnergy *= ew->factor;
    out->energy += recipEnergy;
    out->coulomb += recipEnergy;
    out->virial += recipEnergy / 3.0;
    out->recip = recipEnergy;
    double selfEnergy = ora_EwaldSelf(ew, s->n_sites, s->charge);
    out->energy += selfEnergy;
    out->coulomb += selfEnergy;
    out->virial += selfEnergy / 3.0;
    out->self_ = selfEnergy;
    out->overlaps = nov;
}

/* Ewald/energy.jl:864-943  potential(...) Wolf variant.  The O(n_s^2) prefactor
 * loop of :924-930 is restated literally (erfc(kappa*r_cut)/r_cut is loop
 * invariant, so hoisting the call gives the same bits). */
void ora_potential_wolf(const ora_system *s, const ora_ewald *ew, double lj_rcut,
                        double qq_rcut, double box, int n_threads, ora_properties *out)
{
    double lj, vir, totReal; int64_t nov;
    ora_potential_rows(s, ew->kappa, lj_rcut, qq_rcut, box, 0, s->n_mol, n_threads,
                       &lj, &vir, &totReal, &nov);
    memset(out, 0, sizeof(*out));
    out->energy = lj / 2;
    out->virial = vir / 2;
    out->lj = lj / 2;
    totReal *= ew->factor / 2;
    out->energy += totReal;
    out->coulomb += totReal;
    out->real = totReal;
    const double r_cut = lj_rcut;                     /* energy.jl:874 */
    const double ec = erfc(ew->kappa * r_cut);
    double prefactor = 0.0;
    for (int64_t i = 0; i < s->n_sites; ++i)
        for (int64_t j = 0; j < s->n_sites; ++j)
            prefactor += s->charge[i] * s->charge[j] * ec / r_cut;
    prefactor *= -1;
    double qq = 0.0;
    for (int64_t i = 0; i < s->n_sites; ++i) qq += s->charge[i] * s->charge[i];
    double prefactor2 = (ec / 2 / r_cut + ew->kappa / sqrt(M_PI)) * qq;
    out->wolf_const = (prefactor - prefactor2) * ew->factor;
    out->energy += (prefactor - prefactor2) * ew->factor;
    out->coulomb += (prefactor - prefactor2) * ew->factor;
    out->overlaps = nov;
}

/* ----------------------------------------------------------------- monatomic */

/* Monatomic/mainMonatomic.jl:227-272  LJ_ΔU(i, system); i 1-based */
void ora_LJ_dU_atom(int64_t i, int64_t n, const double *r, const double *eps,
                    const double *sig, double box, double r_cut, double *pot_out, double *vir_out)
{
    const double rcut_sq = r_cut * r_cut;
    double pot = 0.0, vir = 0.0;
    const double *ri = r + 3 * (i - 1);
    for (int64_t j = 1; j <= n; ++j) {
        if (j == i) continue;
        const double *atom = r + 3 * (j - 1);
        double d[3];
        for (int k = 0; k < 3; ++k) d[k] = ora_vector1D(ri[k], atom[k], box);
        double rij_sq = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        if (rij_sq > rcut_sq) {
            pot += 0.0;
            vir += 0.0;
        } else {
            double sr2 = sig[j - 1] * sig[j - 1] / rij_sq;
            double sr6 = sr2 * sr2 * sr2;
            double sr12 = sr6 * sr6;
            pot += eps[j - 1] * (sr12 - sr6);
            vir += eps[j - 1] * (2 * sr12 - sr6);
        }
    }
    *pot_out = pot * 4.0;
    *vir_out = vir * 24.0 / 3.0;
}

/* Monatomic/mainMonatomic.jl:275-289  potential(system, tot) — double counts, halves */
void ora_potential_atoms(int64_t n, const double *r, const double *eps, const double *sig,
                         double box, double r_cut, int n_threads, double *energy, double *virial)
{
    double *e = (double *)malloc(sizeof(double) * 2 * n);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 64) num_threads(n_threads > 0 ? n_threads : 1)
#endif
    for (int64_t i = 1; i <= n; ++i)
        ora_LJ_dU_atom(i, n, r, eps, sig, box, r_cut, &e[2 * (i - 1)], &e[2 * (i - 1) + 1]);
    double en = 0.0, vi = 0.0;
    for (int64_t i = 0; i < n; ++i) { en += e[2 * i]; vi += e[2 * i + 1]; }
    *energy = en / 2;
    *virial = vi / 2;
    free(e);
    (void)n_threads;
}

/* ------------------------------------------------------------- volume change */

/* Ewald/volumeChange.jl:62-80: f = L_new/L_old; coords_new = f*coords;
 * change = coords_new - coords; atom_XYZ = atom_coords + change(mol) */
void ora_volume_scale(ora_system *s, double box_old, double box_new)
{
    const double f = box_new / box_old;
    for (int64_t i = 0; i < s->n_mol; ++i) {
        double change[3];
        for (int k = 0; k < 3; ++k) {
            double cn = f * s->com[3 * i + k];
            change[k] = cn - s->com[3 * i + k];
            s->com[3 * i + k] = cn;
        }
        for (int64_t a = s->first_atom[i]; a <= s->last_atom[i]; ++a)
            for (int k = 0; k < 3; ++k) s->coords[3 * (a - 1) + k] = s->coords[3 * (a - 1) + k] + change[k];
    }
}

/* -------------------------------------------------- driver pieces (the caller) */

/* Ewald/quaternions.jl:11-50  q_to_a — rows exactly as written in the reference,
 * including the a[2,3] element 2*(q2*q4 + q1*q2) at :42-44. q is (q1..q4). */
void ora_q_to_a(const double q[4], double a[9])
{
    const double q1 = q[0], q2 = q[1], q3 = q[2], q4 = q[3];
    a[0] = q1 * q1 + q2 * q2 - q3 * q3 - q4 * q4;
    a[1] = 2 * (q2 * q3 + q1 * q4);
    a[2] = 2 * (q2 * q4 - q1 * q3);
    a[3] = 2 * (q2 * q3 - q1 * q4);
    a[4] = q1 * q1 - q2 * q2 + q3 * q3 - q4 * q4;
    a[5] = 2 * (q2 * q4 + q1 * q2);
    a[6] = 2 * (q2 * q4 + q1 * q3);
    a[7] = 2 * (q3 * q4 - q1 * q2);
    a[8] = q1 * q1 - q2 * q2 - q3 * q3 + q4 * q4;
}

/* Ewald/auxillary.jl:154-159  MATMUL(ai, db) = (db·ai[:,1], db·ai[:,2], db·ai[:,3]) */
void ora_MATMUL(const double a[9], const double db[3], double out[3])
{
    for (int c = 0; c < 3; ++c)
        out[c] = db[0] * a[0 + c] + db[1] * a[3 + c] + db[2] * a[6 + c];
}

/* Ewald/quaternions.jl:75-92 */
void ora_quatmul(const double a[4], const double b[4], double c[4])
{
    c[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    c[1] = a[1] * b[0] + a[0] * b[1] - a[3] * b[2] + a[2] * b[3];
    c[2] = a[2] * b[0] + a[3] * b[1] + a[0] * b[2] - a[1] * b[3];
    c[3] = a[3] * b[0] - a[2] * b[1] + a[1] * b[2] + a[0] * b[3];
}

typedef struct { const double *u; int64_t n, pos; int dry; } ustream;
static inline double urand(ustream *s)
{
    if (s->pos >= s->n) { s->dry = 1; return 0.5; }
    return s->u[s->pos++];
}

/* Ewald/auxillary.jl:106-114 */
static int metropolis(double delta, ustream *us)
{
    if (delta < 0.0) return 1;
    else if (exp(-delta) > urand(us)) return 1;
    else return 0;
}

/* Ewald/adjust.jl:1-41 (Adjust!) and :43-83 (Adjust_rot!) — identical bodies */
typedef struct { int64_t naccepp, naccept, attempp, attempt; double set_value, d_max; } moves_t;
static void adjust(moves_t *m, double L)
{
    if (m->attempp == 0) {
        m->naccepp = m->naccept;
        m->attempp = m->attempt;
    } else {
        double ratio = (double)(m->naccept - m->naccepp) / (double)(m->attempt - m->attempp);
        double dr_old = m->d_max;
        m->d_max = m->d_max * ratio / m->set_value;
        double dr_ratio = m->d_max / dr_old;
        if (dr_ratio > 1.5) m->d_max = dr_old * 1.5;
        if (dr_ratio < 0.5) m->d_max = dr_old * 0.5;
        if (m->d_max > L / 2) m->d_max = L / 2;
        m->naccepp = m->naccept;
        m->attempp = m->attempt;
    }
}

/* Ewald/main.jl:487-651  Loop(...) — one call = n_moves trial moves, molecules
 * visited in sweep order i = 1..N, Adjust! after each full sweep (:645-651).
 * Draw order per move: SURVEY.md Appendix A.5. */
int ora_loop(ora_system *s, ora_ewald *ew, const double *db, double *quat,
             const ora_loop_params *p, const double *uniforms, int64_t n_uniforms,
             int64_t n_moves, double e0, double v0, uint8_t *accepted, double *delta_out,
             ora_loop_stats *st)
{
    ustream us = { uniforms, n_uniforms, 0, 0 };
    moves_t trans = { 0, 0, 0, 0, 0.5, p->dr_max }, rot = { 0, 0, 0, 0, 0.5, p->dphi_max };
    double dr_max = p->dr_max, dphi_max = p->dphi_max;
    const double box = p->box;
    memset(st, 0, sizeof(*st));
    st->total_energy = e0; st->total_virial = v0;
    double ra_old[3 * 64], ra_new[3 * 64];
    int rc = 0;

    for (int64_t m = 0; m < n_moves; ++m) {
        const int64_t i = (m % s->n_mol) + 1;
        const int64_t fa = s->first_atom[i - 1], la = s->last_atom[i - 1];
        const int64_t npm = la - fa + 1;
        double partial_old_e, partial_old_v, pe, pv; int overlap1 = 0, overlap2 = 0;
        ora_LJ_poly_dU(i, s, p->lj_rcut, box, &partial_old_e, &partial_old_v);     /* :491 */
        if (p->style != 2) {
            ora_EwaldShort(i, s, ew, p->qq_rcut, box, &pe, &pv, &overlap1);        /* :501 */
            partial_old_v += pv;
            partial_old_e += pe;
        }
        double rm_old[3];
        memcpy(rm_old, s->com + 3 * (i - 1), sizeof(rm_old));                      /* :514 */
        memcpy(ra_old, s->coords + 3 * (fa - 1), sizeof(double) * 3 * npm);        /* :515 */
        const int64_t pos_m = us.pos;   /* resume point if the caller's stream ends inside this move */
        double chose_move = urand(&us);                                            /* :516 */
        double ei[4], ai[9];
        int is_trans;
        if (chose_move < p->p_trans) {                                             /* :519 */
            is_trans = 1;
            trans.attempt += 1;
            /* auxillary.jl:94-103 random_translate_vector */
            double rnew[3];
            double z0 = urand(&us), z1 = urand(&us), z2 = urand(&us);
            rnew[0] = s->com[3 * (i - 1) + 0] + (z0 - 0.5) * dr_max;
            rnew[1] = s->com[3 * (i - 1) + 1] + (z1 - 0.5) * dr_max;
            rnew[2] = s->com[3 * (i - 1) + 2] + (z2 - 0.5) * dr_max;
            ora_PBC(rnew, box);
            memcpy(s->com + 3 * (i - 1), rnew, sizeof(rnew));
            memcpy(ei, quat + 4 * (i - 1), sizeof(ei));
        } else if (chose_move <= p->p_rot) {                                       /* :530 */
            is_trans = 0;
            rot.attempt += 1;
            /* quaternions.jl:158-182 random_rotate_quaternion */
            const double *old = quat + 4 * (i - 1);
            double nrm = old[0] * old[0] + old[1] * old[1] + old[2] * old[2] + old[3] * old[3];
            if (fabs(nrm - 1.0) > 1.e-6) { rc = 2; st->n_moves = m; goto done; }
            /* quaternions.jl:52-73 random_vector */
            double e[3], norm;
            for (;;) {
                e[0] = 2.0 * urand(&us) - 1.0;
                e[1] = 2.0 * urand(&us) - 1.0;
                e[2] = 2.0 * urand(&us) - 1.0;
                norm = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
                if (norm < 1.0 || us.dry) break;
            }
            double sn = sqrt(norm);
            e[0] = e[0] / sn; e[1] = e[1] / sn; e[2] = e[2] / sn;
            double zeta = urand(&us);
            double angle = (2.0 * zeta - 1.0) * dphi_max;
            /* quaternions.jl:94-120 rotate_quaternion */
            double rotq[4];
            rotq[0] = cos(0.5 * angle);
            rotq[1] = sin(0.5 * angle) * e[0];
            rotq[2] = sin(0.5 * angle) * e[1];
            rotq[3] = sin(0.5 * angle) * e[2];
            ora_quatmul(rotq, old, ei);
        } else {
            rc = 3; st->n_moves = m; goto done;                                    /* :539-541 */
        }
        if (us.dry) {
            /* The reference's RNG never ends; a finite recorded stream can.  A move whose draws ran past the end never
             * happened: state, counters and stream position go back to the start of the move (same rule in the engine's
             * drivers, csrc/mmc_loop.cu and kernels_chain.cuh), so a caller who refills the stream resumes exactly here. */
            memcpy(s->com + 3 * (i - 1), rm_old, sizeof(rm_old));
            if (is_trans) trans.attempt -= 1; else rot.attempt -= 1;
            us.pos = pos_m; rc = 1; st->n_moves = m; goto done;
        }
        {   /* q_to_a aborts on |q·q-1| > 1e-6 (quaternions.jl:20-25) */
            double nrm = ei[0] * ei[0] + ei[1] * ei[1] + ei[2] * ei[2] + ei[3] * ei[3];
            if (fabs(nrm - 1.0) > 1.e-6) { rc = 2; st->n_moves = m; goto done; }
        }
        ora_q_to_a(ei, ai);
        for (int64_t a = 0; a < npm; ++a) {                                        /* :545-548 */
            double d[3];
            ora_MATMUL(ai, db + 3 * (fa - 1 + a), d);
            for (int k = 0; k < 3; ++k) ra_new[3 * a + k] = s->com[3 * (i - 1) + k] + d[k];
        }
        memcpy(s->coords + 3 * (fa - 1), ra_new, sizeof(double) * 3 * npm);        /* :552 */

        double partial_new_e, partial_new_v;
        ora_LJ_poly_dU(i, s, p->lj_rcut, box, &partial_new_e, &partial_new_v);     /* :557 */
        if (p->style != 2) {
            ora_EwaldShort(i, s, ew, p->qq_rcut, box, &pe, &pv, &overlap2);        /* :566 */
            partial_new_v += pv;
            partial_new_e += pe;
        }
        int overlap = (overlap1 || overlap2);
        double deltaRecip = 0.0;
        if (!overlap && p->style == 0)                                             /* :580 */
            deltaRecip = ora_RecipMove(box, ew, npm, ra_old, ra_new, s->charge + (fa - 1));
        double delta = (partial_new_e) - (partial_old_e) + deltaRecip;             /* :593 */
        int acc = (metropolis(delta / p->temperature, &us) && overlap == 0);       /* :598 */
        if (us.dry) {   /* Metropolis found the stream empty: the move never happened (see above) */
            memcpy(s->com + 3 * (i - 1), rm_old, sizeof(rm_old));
            memcpy(s->coords + 3 * (fa - 1), ra_old, sizeof(double) * 3 * npm);
            if (p->style == 0) ora_recip_rollback(ew);
            if (is_trans) trans.attempt -= 1; else rot.attempt -= 1;
            us.pos = pos_m; rc = 1; st->n_moves = m; goto done;
        }
        if (overlap) st->n_overlap += 1;
        if (acc) {
            st->total_energy += delta;
            st->total_virial += (partial_new_v - partial_old_v) + deltaRecip / 3;
            st->n_accepted += 1;
            if (is_trans) trans.naccept += 1; else rot.naccept += 1;
            memcpy(quat + 4 * (i - 1), ei, sizeof(ei));                            /* :620 */
            if (p->style == 0) ora_recip_commit(ew);                               /* :621 */
        } else {
            memcpy(s->com + 3 * (i - 1), rm_old, sizeof(rm_old));                  /* :623 */
            memcpy(s->coords + 3 * (fa - 1), ra_old, sizeof(double) * 3 * npm);    /* :624 */
            if (p->style == 0) ora_recip_rollback(ew);                             /* :628 */
        }
        if (accepted) accepted[m] = (uint8_t)acc;
        if (delta_out) delta_out[m] = delta;
        if (p->adjust && i == s->n_mol) {                                          /* :645-651 */
            trans.d_max = dr_max; adjust(&trans, box); dr_max = trans.d_max;
            rot.d_max = dphi_max; adjust(&rot, box); dphi_max = rot.d_max;
        }
        st->n_moves = m + 1;
    }
done:
    st->uniforms_used = us.pos;
    st->trans_attempt = trans.attempt; st->trans_accept = trans.naccept;
    st->rot_attempt = rot.attempt; st->rot_accept = rot.naccept;
    st->dr_max = dr_max; st->dphi_max = dphi_max;
    return rc;
}

/* Monatomic/mainMonatomic.jl:373-413 — translation only, no move-type draw,
 * no step-size adaptation. */
int ora_loop_atoms(int64_t n, double *r, const double *eps, const double *sig, double box,
                   double r_cut, double temperature, double dr_max,
                   const double *uniforms, int64_t n_uniforms, int64_t n_moves,
                   double e0, double v0, uint8_t *accepted, double *delta_out,
                   ora_loop_stats *st)
{
    ustream us = { uniforms, n_uniforms, 0, 0 };
    memset(st, 0, sizeof(*st));
    st->total_energy = e0; st->total_virial = v0;
    int rc = 0;
    for (int64_t m = 0; m < n_moves; ++m) {
        const int64_t i = (m % n) + 1;
        double eo, vo, en, vn, rold[3], rnew[3];
        ora_LJ_dU_atom(i, n, r, eps, sig, box, r_cut, &eo, &vo);
        memcpy(rold, r + 3 * (i - 1), sizeof(rold));
        const int64_t pos_m = us.pos;
        double z0 = urand(&us), z1 = urand(&us), z2 = urand(&us);
        if (us.dry) { us.pos = pos_m; rc = 1; st->n_moves = m; break; }   /* the move never happened (see ora_loop) */
        rnew[0] = rold[0] + (z0 - 0.5) * dr_max;
        rnew[1] = rold[1] + (z1 - 0.5) * dr_max;
        rnew[2] = rold[2] + (z2 - 0.5) * dr_max;
        ora_PBC(rnew, box);
        memcpy(r + 3 * (i - 1), rnew, sizeof(rnew));
        ora_LJ_dU_atom(i, n, r, eps, sig, box, r_cut, &en, &vn);
        double delta = en - eo;
        int acc = metropolis(delta / temperature, &us);
        if (us.dry) { memcpy(r + 3 * (i - 1), rold, sizeof(rold)); us.pos = pos_m; rc = 1; st->n_moves = m; break; }
        if (acc) {
            st->total_energy += delta;
            st->total_virial += (vn - vo);
            st->n_accepted += 1;
        } else {
            memcpy(r + 3 * (i - 1), rold, sizeof(rold));
        }
        if (accepted) accepted[m] = (uint8_t)acc;
        if (delta_out) delta_out[m] = delta;
        st->n_moves = m + 1;
        st->trans_attempt += 1; st->trans_accept += acc;
    }
    st->uniforms_used = us.pos;
    st->dr_max = dr_max;
    return rc;
}
