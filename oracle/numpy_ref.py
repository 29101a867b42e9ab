"""Independent numpy restatement of the reference formulas (ORACLE / test infrastructure).

Purpose: a second, differently-structured implementation to pin oracle/mmc_oracle.c
against.  It follows SURVEY.md Appendix A (formulas), not the C code: vectorised
minimum image via the same compare, direct cos/sin of k·r for S(k) (no recurrence).
Agreement is expected to ~1e-12 relative, not bitwise.  Never imported by the product.
"""
from __future__ import annotations

import numpy as np
from scipy.special import erfc


def vector1d(c1, c2, box):
    """Ewald/boundaries.jl:8-14, vectorised."""
    d = c2 - c1
    lt = c1 < c2
    keep_a = d < (c1 - c2 + box)
    keep_b = (c1 - c2) < (d + box)
    return np.where(lt, np.where(keep_a, d, d - box), np.where(keep_b, d, d + box))


def _partners(i, com, r_cut, box):
    n = com.shape[0]
    rij = vector1d(com[i - 1][None, :], com, box)
    r2 = (rij * rij).sum(axis=1)
    mask = r2 < r_cut * r_cut
    mask[i - 1] = False
    return np.nonzero(mask)[0], rij


def lj_poly(i, s, r_cut, box):
    """Ewald/energy.jl:209-290; s has coords, atype, first_atom, last_atom, com, eps, sig (nt,nt)."""
    js, rij = _partners(i, s.com, r_cut, box)
    pot = vir = 0.0
    ia = np.arange(s.first_atom[i - 1] - 1, s.last_atom[i - 1])
    for j in js:
        jb = np.arange(s.first_atom[j] - 1, s.last_atom[j])
        rab = vector1d(s.coords[ia][:, None, :], s.coords[jb][None, :, :], box)
        r2 = (rab * rab).sum(axis=2)
        e = s.eps[np.ix_(s.atype[ia] - 1, s.atype[jb] - 1)]
        sg = s.sig[np.ix_(s.atype[ia] - 1, s.atype[jb] - 1)]
        ok = (r2 < r_cut * r_cut + 100) & (e > 0.001)
        s2 = np.where(ok, sg * sg / r2, 0.0)
        s6 = s2 ** 3
        s12 = s6 ** 2
        pot += (e * (s12 - s6))[ok].sum()
        virab = e * (2.0 * s12 - s6) * s2
        vir += ((rab * virab[:, :, None]) @ rij[j])[ok].sum()
    return 4 * pot, 24 * vir / 3.0


def ewald_real(i, s, kappa, r_cut, box):
    """Ewald/ewalds.jl:293-376."""
    js, _ = _partners(i, s.com, r_cut, box)
    ia = np.arange(s.first_atom[i - 1] - 1, s.last_atom[i - 1])
    pot = 0.0
    for j in js:
        jb = np.arange(s.first_atom[j] - 1, s.last_atom[j])
        rab = vector1d(s.coords[ia][:, None, :], s.coords[jb][None, :, :], box)
        r2 = (rab * rab).sum(axis=2)
        qq = s.charge[ia][:, None] * s.charge[jb][None, :]
        if np.any((r2 < 0.5) & (qq < 0)):
            return 0.0, True
        r = np.sqrt(r2)
        pot += (qq * erfc(kappa * r) / r)[r2 < r_cut * r_cut + 100].sum()
    return pot, False


def kvectors(kappa, nk, k_sq_max, box):
    """Ewald/ewalds.jl:45-103."""
    ks, cf = [], []
    b = 1.0 / 4.0 / kappa / kappa / box / box
    for kx in range(0, nk + 1):
        for ky in range(-nk, nk + 1):
            for kz in range(-nk, nk + 1):
                k2 = kx * kx + ky * ky + kz * kz
                if 0 < k2 < k_sq_max:
                    kr2 = (2 * np.pi) ** 2 * k2
                    c = 2 * np.pi * np.exp(-b * kr2) / kr2 / box
                    ks.append((kx, ky, kz))
                    cf.append(2 * c if kx > 0 else c)
    return np.array(ks, dtype=np.int32), np.array(cf)


def structure_factor(kxyz, r, q, box):
    """S(k) = Σ q e^{i 2π k·r / L}, direct evaluation."""
    out = np.zeros(len(kxyz), dtype=np.complex128)
    for lo in range(0, len(kxyz), 64):
        k = kxyz[lo:lo + 64].astype(np.float64)
        ph = 2 * np.pi * (k @ r.T) / box
        out[lo:lo + 64] = (q[None, :] * (np.cos(ph) + 1j * np.sin(ph))).sum(axis=1)
    return out


def recip_energy(cfac, S):
    return float((cfac * (S.real ** 2 + S.imag ** 2)).sum())


def ewald_self(kappa, factor, q):
    return -kappa * (q * q).sum() / np.sqrt(np.pi) * factor


def lj_atom(i, r, eps, sig, box, r_cut):
    """Monatomic/mainMonatomic.jl:227-272 — pair kept iff not (r² > rc²)."""
    d = vector1d(r[i - 1][None, :], r, box)
    r2 = (d * d).sum(axis=1)
    ok = ~(r2 > r_cut * r_cut)
    ok[i - 1] = False
    sr2 = sig[ok] ** 2 / r2[ok]
    sr6 = sr2 ** 3
    sr12 = sr6 ** 2
    return 4.0 * (eps[ok] * (sr12 - sr6)).sum(), 24.0 * (eps[ok] * (2 * sr12 - sr6)).sum() / 3.0
