"""Independent pin of the real-space rows: LJ_poly_ΔU(i) and EwaldReal(i) for every molecule of coord750 in 40-digit arithmetic.

Written from the JULIA source (Ewald/boundaries.jl:8-14 vector1D; Ewald/energy.jl:209-290 LJ_poly_ΔU; Ewald/ewalds.jl:293-376
EwaldReal), not from oracle/mmc_oracle.c: COM gate `rij2 < r_cut^2` on the minimum-image COM vector, every site pair with its own
per-component minimum image, `rab2 < r_cut^2 + 100` on the site pairs, eps > 0.001 for LJ, early return (0.0, true) on
`rab2 < 0.5 && q_a q_b < 0`, results (4 pot, 24 vir / 3) and the un-scaled Coulomb sum.  Inputs are the float64 arrays the engine
and the oracle receive (systems.load_nist(4) = Ewald/coord750.txt shifted as main.jl:247-275), converted exactly to mpmath numbers;
all arithmetic, sqrt and erfc in 40 digits, so the result is the exact value of the reference's formula on those inputs, rounded
once to float64 (error <= 1 ulp; the float64 oracle/engine differ from it by their own accumulated rounding, ~1e-15 relative).
Decisions (`<`) are taken on the 40-digit values; the script reports the closest call so that a float64 decision cannot differ.

    python tests/golden/make_realspace_pin.py        ->  tests/golden/realspace_pin_coord750.npz   (about 3 minutes)
"""
import sys
from pathlib import Path

import numpy as np
from mpmath import mp, mpf, erfc, sqrt

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from metropolismontecarlo_b200 import systems  # noqa: E402

mp.dps = 40


def vector1D(c1, c2, box):
    if c1 < c2:
        return (c2 - c1) if (c2 - c1) < (c1 - c2 + box) else (c2 - c1 - box)
    return (c2 - c1) if (c1 - c2) < (c2 - c1 + box) else (c2 - c1 + box)


def main():
    ms = systems.load_nist(4)
    n = ms.n_mol
    box, r_cut = mpf(ms.box), mpf(10.0)
    kappa = mpf(float(systems.ALPHA / ms.box))            # the float64 kappa = 5.6 / L the engine receives
    com = [[mpf(float(v)) for v in row] for row in ms.com]
    xyz = [[mpf(float(v)) for v in row] for row in ms.coords]
    q = [mpf(float(v)) for v in ms.charge]
    atype = ms.atype
    eps = [[mpf(float(v)) for v in row] for row in ms.eps]
    sig = [[mpf(float(v)) for v in row] for row in ms.sig]
    first, last = ms.first_atom, ms.last_atom
    rc2 = r_cut * r_cut
    lj_pot = np.zeros(n); lj_vir = np.zeros(n); qq = np.zeros(n); ovl = np.zeros(n, dtype=np.int8)
    closest_gate, closest_ovr = mpf(1), mpf(1)
    for i in range(n):
        pot, vir, cpot = mpf(0), mpf(0), mpf(0)
        overlap = False
        for j in range(n):
            if j == i:
                continue
            rij = [vector1D(com[i][k], com[j][k], box) for k in range(3)]
            rij2 = rij[0] * rij[0] + rij[1] * rij[1] + rij[2] * rij[2]
            closest_gate = min(closest_gate, abs(rij2 - rc2) / rc2)
            if not (rij2 < rc2):
                continue
            for a in range(first[i] - 1, last[i]):
                for b in range(first[j] - 1, last[j]):
                    rab = [vector1D(xyz[a][k], xyz[b][k], box) for k in range(3)]
                    rab2 = rab[0] * rab[0] + rab[1] * rab[1] + rab[2] * rab[2]
                    e = eps[atype[a] - 1][atype[b] - 1]
                    if rab2 < rc2 + 100 and e > mpf("0.001"):
                        s = sig[atype[a] - 1][atype[b] - 1]
                        s2 = s * s / rab2
                        s6 = s2 ** 3
                        s12 = s6 ** 2
                        pot += e * (s12 - s6)
                        virab = e * (2 * s12 - s6)
                        vir += sum(rij[k] * rab[k] * virab * s2 for k in range(3))
                    qaqb = q[a] * q[b]
                    closest_ovr = min(closest_ovr, abs(rab2 - mpf("0.5")))
                    if not overlap:
                        if rab2 < mpf("0.5") and qaqb < 0:
                            overlap = True          # EwaldReal returns (0.0, true) here; LJ_poly_ΔU has no such rule
                        elif rab2 < rc2 + 100:
                            r = sqrt(rab2)
                            cpot += qaqb * erfc(kappa * r) / r
        lj_pot[i] = float(pot * 4)
        lj_vir[i] = float(vir * 24 / 3)
        qq[i] = 0.0 if overlap else float(cpot)
        ovl[i] = 1 if overlap else 0
        if i % 50 == 0:
            print(i, lj_pot[i], qq[i], flush=True)
    out = Path(__file__).resolve().parent / "realspace_pin_coord750.npz"
    np.savez_compressed(out, lj_pot=lj_pot, lj_vir=lj_vir, qq_pot=qq, overlap=ovl,
                        closest_gate_rel=float(closest_gate), closest_overlap_abs=float(closest_ovr),
                        r_cut=10.0, kappa=float(kappa), box=float(box), digits=mp.dps)
    print("wrote", out, "closest COM-gate call (relative):", float(closest_gate), " closest r²-0.5:", float(closest_ovr))
    print("totals/2: LJ", lj_pot.sum() / 2, " real (un-scaled)", qq.sum() / 2)


if __name__ == "__main__":
    main()
