// erf_poly.h — host-side fit of the smooth part of the Ewald real-space kernel.
//
//   erfc(κ r)/r = 1/r − κ·E(κ² r²),     E(v) = erf(√v)/√v = (2/√π) Σ_n (−v)ⁿ / (n! (2n+1))
//
// E is entire in v, so on the bounded domain the reference itself imposes on every site pair
// (rab² < r_cut² + 100, Ewald/ewalds.jl:362) it is approximated to double rounding by ONE
// polynomial whose coefficients are the same for every lane: they sit in the kernel's constant
// bank and cost no loads, no table, no branches.  The degree adapts to the box: κ = 5.6/L makes
// v_max = κ²(r_cut²+100) ≈ 0.16 for the 256k-molecule box (degree ≈ 9) and ≈ 7 for L = 30 Å
// (degree ≈ 26).  The fit is a Chebyshev interpolant computed in long double (glibc erfl) and
// re-expanded in the mapped variable s = 2v/v_max − 1 ∈ [−1, 1]; its measured error against erfl
// on a dense grid, evaluated with the same fma-Horner the device uses, picks the degree.
#pragma once
#include <cmath>
#include <vector>

#define MMC_ERF_MAXDEG 44
#define MMC_ERF_DIRECT_MAXDEG 12

struct ErfPoly {
    int deg;                 // 0: no usable fit, kernels fall back to erfc()
    double kappa, kappa2;    // κ, κ²
    double scale;            // 2 / v_max
    double c[MMC_ERF_MAXDEG + 1];
    // "direct" form for small domains (big boxes: v_max = κ²(r_cut²+100) ≲ 3): E(v) ≈ Σ a_k v^k by
    // plain Horner in v, so a kernel can fold κ into the coefficients and run Horner in r² itself
    // (no mapped variable, exact degree, no padding).  ddeg = 0: not accurate enough, use c[].
    int ddeg;
    double a[MMC_ERF_DIRECT_MAXDEG + 1];
};

namespace erfpoly {

inline long double E_ref(long double v)
{
    if (v < 1e-8L) return 1.1283791670955125738961589031215452L * (1.0L - v / 3.0L + v * v / 10.0L);
    const long double x = sqrtl(v);
    return erfl(x) / x;
}

// monomial coefficients (in s) of the degree-D Chebyshev interpolant of E on [0, vmax]
inline void cheb_fit(int D, long double vmax, std::vector<long double> &mono)
{
    const int N = D + 1;
    const long double PI = 3.14159265358979323846264338327950288L;
    std::vector<long double> f(N), a(N);
    for (int k = 0; k < N; ++k) {
        const long double s = cosl(PI * (k + 0.5L) / N);
        f[k] = E_ref((s + 1.0L) * vmax * 0.5L);
    }
    for (int j = 0; j < N; ++j) {
        long double acc = 0.0L;
        for (int k = 0; k < N; ++k) acc += f[k] * cosl(PI * j * (k + 0.5L) / N);
        a[j] = acc * 2.0L / N;
    }
    a[0] *= 0.5L;
    // T_0 = 1, T_1 = s, T_{n+1} = 2 s T_n − T_{n−1}
    std::vector<long double> t0(N, 0.0L), t1(N, 0.0L), t2(N, 0.0L);
    mono.assign(N, 0.0L);
    t0[0] = 1.0L;
    mono[0] += a[0];
    if (N > 1) { t1[1] = 1.0L; mono[1] += a[1]; }
    for (int n = 2; n < N; ++n) {
        for (int i = 0; i < N; ++i) t2[i] = (i > 0 ? 2.0L * t1[i - 1] : 0.0L) - t0[i];
        for (int i = 0; i < N; ++i) mono[i] += a[n] * t2[i];
        t0 = t1; t1 = t2;
    }
}

inline double eval(const ErfPoly &P, double v)
{
    const double s = std::fma(v, P.scale, -1.0);
    double p = P.c[P.deg];
    for (int k = P.deg - 1; k >= 0; --k) p = std::fma(p, s, P.c[k]);
    return p;
}

inline double eval_direct(const ErfPoly &P, double v)
{
    double p = P.a[P.ddeg];
    for (int k = P.ddeg - 1; k >= 0; --k) p = std::fma(p, v, P.a[k]);
    return p;
}

// Direct (monomial-in-v) form: the degree-D Chebyshev interpolant re-expanded from s = σv − 1 to v in
// long double; accepted when the double-precision Horner in v meets `tol` on the dense grid.
inline void fit_direct(long double vmax, ErfPoly &P, const std::vector<double> &vs, const std::vector<long double> &ref,
                       double tol)
{
    P.ddeg = 0;
    const long double sigma = 2.0L / vmax;
    std::vector<long double> mono;
    for (int D = 3; D <= MMC_ERF_DIRECT_MAXDEG; ++D) {
        cheb_fit(D, vmax, mono);
        ErfPoly Q = P;
        Q.ddeg = D;
        for (int k = 0; k <= D; ++k) {
            long double acc = 0.0L, binom = 1.0L;      // C(i, k) for i = k, k+1, ...
            for (int i = k; i <= D; ++i) {
                acc += mono[i] * binom * (((i - k) & 1) ? -1.0L : 1.0L);
                binom = binom * (long double)(i + 1) / (long double)(i + 1 - k);
            }
            Q.a[k] = (double)(acc * powl(sigma, (long double)k));
        }
        double err = 0.0;
        for (size_t i = 0; i < vs.size(); ++i) {
            const double e = std::fabs((double)((long double)eval_direct(Q, vs[i]) - ref[i]));
            if (e > err) err = e;
        }
        if (err <= tol) { for (int k = 0; k <= D; ++k) P.a[k] = Q.a[k]; P.ddeg = D; return; }
    }
}

// Picks the smallest degree whose measured max |E_poly − E| is below `tol`; returns that error.
// vmax is the upper end of the fitted domain in v = κ² r² (the coefficients depend on vmax only).
inline double fit(double vmax_d, ErfPoly &P, double tol = 2.5e-16)
{
    const long double vmax = (long double)vmax_d;
    P.scale = (double)(2.0L / vmax);
    P.deg = 0;
    const int M = 1200;
    std::vector<double> vs(M + 1);
    std::vector<long double> ref(M + 1);
    for (int i = 0; i <= M; ++i) { vs[i] = (double)(vmax * i / M); ref[i] = E_ref((long double)vs[i]); }
    double best = 1e300;
    ErfPoly bestP = P;
    std::vector<long double> mono;
    for (int D = 4; D <= MMC_ERF_MAXDEG; ++D) {
        cheb_fit(D, vmax, mono);
        ErfPoly Q = P;
        Q.deg = D;
        for (int i = 0; i <= D; ++i) Q.c[i] = (double)mono[i];
        double err = 0.0;
        for (int i = 0; i <= M; ++i) {
            const double e = std::fabs((double)((long double)eval(Q, vs[i]) - ref[i]));
            if (e > err) err = e;
        }
        if (err < best) { best = err; bestP = Q; }
        if (err <= tol) break;
        if (D > 12 && err > 4.0 * best) break;   // past the conditioning floor
    }
    if (best <= 1e-15) {
        const double k = P.kappa, k2 = P.kappa2;
        P = bestP; P.kappa = k; P.kappa2 = k2;
        // pad to the next instantiated kernel length (8, 12, 16, 20, 24, 32, 44): zero leading coefficients are exact no-ops
        static const int lens[] = {8, 12, 16, 20, 24, 32, 44};
        int padded = MMC_ERF_MAXDEG;
        for (int l : lens) if (P.deg <= l) { padded = l; break; }
        for (int i = P.deg + 1; i <= padded; ++i) P.c[i] = 0.0;
        P.deg = padded;
    }
    fit_direct(vmax, P, vs, ref, tol);
    return best;
}

// Domain ends are rounded up to a geometric grid (ratio 2^(1/4)) so that the small box-size
// changes of NPT volume moves reuse a cached fit: a fit on [0, v_grid] covers every v_max below it.
inline double grid_vmax(double vmax)
{
    const double e = std::ceil(std::log2(vmax) * 4.0) / 4.0;
    return std::exp2(e);
}

}  // namespace erfpoly
