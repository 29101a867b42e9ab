"""CPU tests: pin the oracle (oracle/mmc_oracle.c) against the reference's own known
answers, the NIST SPC/E reference energies for the bundled configurations, and the
independent numpy restatement.  No GPU, no product code under test here."""
from pathlib import Path

import numpy as np
import pytest

from metropolismontecarlo_b200 import systems
from oracle import numpy_ref as npr
from oracle import oracle as ora
from tests.util import ora_ewald, ora_system, rel


def LennardJones(rij):  # Monatomic/mainMonatomic.jl:335-337
    return 4 * 1 * ((1 / rij) ** 12 - (1 / rij) ** 6)


def test_factor_constant():
    # Ewald/constants.jl:24-28 → 167100.9566... K·Å/e²
    assert abs(ora.factor() - 167100.95663229248) < 1e-6
    assert ora.factor() == systems.FACTOR


def test_vector1D_minimum_image():
    # Ewald/boundaries.jl:8-14
    assert ora.vector1D(0.0, 4.0, 5.0) == -1.0
    assert ora.vector1D(4.0, 0.0, 5.0) == 1.0
    assert ora.vector1D(0.0, 2.0, 5.0) == 2.0
    assert ora.vector1D(1.0, 1.0, 5.0) == 0.0
    # tie |d| == L/2 is not "<" so it wraps (c1<c2 branch: d - L)
    assert ora.vector1D(0.0, 2.5, 5.0) == -2.5
    assert ora.vector1D(2.5, 0.0, 5.0) == 2.5
    rng = np.random.default_rng(1)
    a, b = rng.random(1000) * 7, rng.random(1000) * 7
    want = npr.vector1d(a, b, 7.0)
    got = np.array([ora.vector1D(x, y, 7.0) for x, y in zip(a, b)])
    assert np.array_equal(want, got)


def test_reference_test_LJ_known_answers():
    # Ewald/tests.jl:127-161 / Monatomic/mainMonatomic.jl:292-330
    box, rc = 5.0, 2.5
    r = np.array([[0, 0, 0], [0, 0, 2], [0, 1.5, 0]], dtype=np.float64)
    e, _ = ora.LJ_dU_atom(1, r, np.ones(3), np.ones(3), box, rc)
    assert abs(e - (-0.381860031778575)) < 1e-14
    assert abs(e - (LennardJones(2.0) + LennardJones(1.5))) < 1e-14
    r[1] = [0, 0, 4]
    e, _ = ora.LJ_dU_atom(1, r, np.ones(3), np.ones(3), box, rc)
    assert abs(e - (-0.320336594278575)) < 1e-14
    assert abs(e - (LennardJones(1.0) + LennardJones(1.5))) < 1e-3   # the reference's own tolerance


def test_reference_two_LJ_triangles():
    # Ewald/tests.jl:8-82: two A&T triangles 2σ apart along z, all sites LJ(ε=σ=1), box 1000
    alpha2 = 75.0 * np.pi / 180.0 / 2.0
    db = np.array([[-np.sin(alpha2), 0.0, -np.cos(alpha2) / 3.0],
                   [0.0, 0.0, 2 * np.cos(alpha2) / 3.0],
                   [np.sin(alpha2), 0.0, -np.cos(alpha2) / 3.0]])
    a, b = db, db + np.array([0, 0, 2.0])
    ra = np.vstack([a, b])
    com = np.vstack([a.mean(axis=0), b.mean(axis=0)])
    s = ora.System(ra, np.zeros(6), np.ones(6, dtype=np.int64), [1, 4], [3, 6], com,
                   np.ones((1, 1)), np.ones((1, 1)))
    want = sum(LennardJones(np.sqrt(((ra[i] - ra[j]) ** 2).sum())) for i in range(3) for j in range(3, 6))
    got, _ = ora.LJ_poly_dU(1, s, 500.0, 1000.0)
    assert abs(got - want) < 1e-4          # the reference's own tolerance (tests.jl:72)
    assert rel(got, want) < 1e-13


# NIST SPC/E reference calculations (10 Å cutoff), E/k_B in K, six published digits.
NIST = {
    1: dict(n=100, L=20.0, fourier=6.27009e3, self_=-2.84469e6),
    2: dict(n=200, L=20.0, fourier=6.03495e3, self_=-5.68938e6),
    3: dict(n=300, L=20.0, fourier=5.24461e3, self_=-8.53407e6),
    4: dict(n=750, L=30.0, fourier=7.58785e3, self_=-1.42235e7),
}


@pytest.mark.parametrize("cfg", [1, 2, 3, 4])
def test_nist_fourier_and_self(cfg):
    ms = systems.load_nist(cfg)
    assert ms.n_mol == NIST[cfg]["n"] and ms.box == NIST[cfg]["L"]
    ew = ora_ewald(ms.box)
    assert ew.nkvecs == 337
    e_recip = ora.RecipLong(ew, ms.coords, ms.charge, ms.box) * ew.factor
    e_self = ora.EwaldSelf(ew, ms.charge)
    assert abs(e_recip / NIST[cfg]["fourier"] - 1) < 5e-6
    assert abs(e_self / NIST[cfg]["self_"] - 1) < 5e-6
    # S(k) left in both buffers (ewalds.jl:600-601)
    assert np.array_equal(ew.sum_old, ew.sum_new) and np.abs(ew.sum_new).max() > 0


def test_nist_config4_atom_cutoff_contrast():
    """Validates fixture + constants: with NIST's ATOM cutoff the numpy formulas give NIST's
    E_disp=4.48593e5 and E_real=-3.57226e6; the reference (COM cutoff) deliberately does not."""
    ms = systems.load_nist(4)
    kappa, L = systems.ALPHA / ms.box, ms.box
    O = ms.coords[0::3]
    d = O[:, None, :] - O[None, :, :]
    d -= L * np.rint(d / L)
    r2 = (d * d).sum(-1)
    iu = np.triu_indices(len(O), 1)
    r2u = r2[iu]
    m = r2u < 100.0
    s6 = (systems.SPCE_SIGMA_O ** 2 / r2u[m]) ** 3
    e_disp = (4 * systems.SPCE_EPS_O * (s6 * s6 - s6)).sum()
    assert abs(e_disp / 4.48593e5 - 1) < 5e-6
    from scipy.special import erfc
    mol = np.repeat(np.arange(ms.n_mol), 3)
    e_real = 0.0
    for lo in range(0, ms.n_sites, 250):
        d = ms.coords[lo:lo + 250, None, :] - ms.coords[None, :, :]
        d -= L * np.rint(d / L)
        r = np.sqrt((d * d).sum(-1))
        qq = ms.charge[lo:lo + 250, None] * ms.charge[None, :]
        ok = (r < 10.0) & (mol[lo:lo + 250, None] != mol[None, :])
        e_real += (qq[ok] * erfc(kappa * r[ok]) / r[ok]).sum()
    e_real *= systems.FACTOR / 2
    assert abs(e_real / -3.57226e6 - 1) < 5e-6


def test_kvectors_match_numpy():
    ew = ora_ewald(30.0)
    k, c = npr.kvectors(systems.ALPHA / 30.0, 5, 27, 30.0)
    assert np.array_equal(k, ew.kxyz)
    assert np.allclose(c, ew.cfac, rtol=1e-15, atol=0)
    assert tuple(ew.kxyz[0]) == (0, -5, -1) and tuple(ew.kxyz[-1]) == (5, 1, 0)


@pytest.mark.parametrize("cfg", [1, 4])
def test_single_molecule_vs_numpy(cfg):
    ms = systems.load_nist(cfg)
    s = ora_system(ms)
    kappa = systems.ALPHA / ms.box
    rc = 10.0 if cfg == 4 else 9.0
    for i in [1, 2, 17, ms.n_mol // 2, ms.n_mol]:
        e, v = ora.LJ_poly_dU(i, s, rc, ms.box)
        e2, v2 = npr.lj_poly(i, ms, rc, ms.box)
        assert rel(e, e2) < 1e-12 and rel(v, v2) < 1e-11
        p, ov = ora.EwaldReal(i, s, kappa, rc, ms.box)
        p2, ov2 = npr.ewald_real(i, ms, kappa, rc, ms.box)
        assert ov == ov2 and rel(p, p2) < 1e-12


def test_recip_vs_numpy_and_delta_consistency():
    ms = systems.load_nist(1)
    s = ora_system(ms)
    ew = ora_ewald(ms.box)
    e = ora.RecipLong(ew, ms.coords, ms.charge, ms.box)
    S = npr.structure_factor(ew.kxyz, ms.coords, ms.charge, ms.box)
    assert np.allclose(ew.sum_new[:, 0] + 1j * ew.sum_new[:, 1], S, rtol=0, atol=1e-11)
    assert rel(e, npr.recip_energy(ew.cfac, S)) < 1e-12
    # move molecule 7: delta update == fresh rebuild (ewalds.jl:434-436 comment)
    rng = np.random.default_rng(3)
    i = 7
    sl = slice(3 * (i - 1), 3 * i)
    r_old = ms.coords[sl].copy()
    r_new = r_old + (rng.random(3) - 0.5)
    dE = ora.RecipMove(ms.box, ew, r_old, r_new, ms.charge[sl])
    coords2 = ms.coords.copy()
    coords2[sl] = r_new
    ew2 = ora_ewald(ms.box)
    e2 = ora.RecipLong(ew2, coords2, ms.charge, ms.box)
    assert abs(dE - (e2 - e) * ew.factor) < 1e-9 * abs(e * ew.factor)
    assert np.allclose(ew.sum_new, ew2.sum_new, rtol=0, atol=1e-11)
    assert not np.array_equal(ew.sum_new, ew.sum_old)
    ora.recip_rollback(ew)
    assert np.array_equal(ew.sum_new, ew.sum_old)


def test_potential_coord750_reference_semantics():
    """coord750.txt under the reference's COM-cutoff semantics (SURVEY §7 step 1 numbers)."""
    ms = systems.load_nist(4)
    s = ora_system(ms)
    ew = ora_ewald(ms.box)
    p = ora.potential_ewald(s, ew, 10.0, 10.0, ms.box, n_threads=4)
    assert abs(p.lj / 4.488629e5 - 1) < 2e-6
    assert abs(p.real / -3.492756e6 - 1) < 2e-6
    assert abs(p.recip / 7.58785e3 - 1) < 5e-6
    assert abs(p.self_ / -1.42235e7 - 1) < 5e-6
    assert p.overlaps == 0
    assert rel(p.energy, p.lj + p.real + p.recip + p.self_) < 1e-14
    assert rel(p.coulomb, p.real + p.recip + p.self_) < 1e-14
    # threads do not change bits (rows summed in reference order)
    p1 = ora.potential_ewald(s, ora_ewald(ms.box), 10.0, 10.0, ms.box, n_threads=1)
    assert p1.energy == p.energy and p1.virial == p.virial
    # Wolf variant: same LJ and real parts; constants in closed form (neutral box)
    w = ora.potential_wolf(s, ew, 10.0, 10.0, ms.box, n_threads=4)
    assert w.lj == p.lj and w.real == p.real
    from scipy.special import erfc
    kappa = ew.kappa
    qq = (ms.charge ** 2).sum()
    closed = -(erfc(kappa * 10.0) / 20.0 + kappa / np.sqrt(np.pi)) * qq * ew.factor
    assert rel(w.wolf_const, closed) < 1e-9
    # Wolf potential() adds no Coulomb virial (energy.jl:905-912)
    assert rel(w.virial, p.virial - (p.real + p.recip + p.self_) / 3.0) < 1e-9


def test_realspace_rows_against_independent_40_digit_pin():
    """The oracle's LJ_poly_ΔU(i) and EwaldReal(i) for ALL 750 molecules of coord750 against tests/golden/
    realspace_pin_coord750.npz: a 40-digit mpmath evaluation written from the Julia source (COM gate, per-site minimum image,
    +100 window, overlap early return; tests/golden/make_realspace_pin.py), i.e. the exact value of the reference's formula on
    these float64 inputs.  This pins the real-space erfc and LJ rows — which the reference itself holds no golden numbers
    for — independently of the C restatement.  1e-13 relative to each row's own magnitude (rows are sums of ~10³ terms of
    mixed sign; the closest COM-gate call in the file is 4e-6 relative, so no float64 decision can differ)."""
    g = np.load(Path(__file__).resolve().parent / "golden" / "realspace_pin_coord750.npz")
    assert float(g["closest_gate_rel"]) > 1e-9 and float(g["closest_overlap_abs"]) > 1e-6
    ms = systems.load_nist(4)
    s = ora_system(ms)
    kappa = systems.ALPHA / ms.box
    assert kappa == float(g["kappa"]) and ms.box == float(g["box"])
    lj = np.empty(750); vir = np.empty(750); qq = np.empty(750)
    for i in range(1, 751):
        lj[i - 1], vir[i - 1] = ora.LJ_poly_dU(i, s, 10.0, ms.box)
        qq[i - 1], ov = ora.EwaldReal(i, s, kappa, 10.0, ms.box)
        assert ov == bool(g["overlap"][i - 1])
    # a row is a sum of ~1000 terms of mixed sign: compare on the scale of the terms (Σ|term| ~ 50x the typical |row|)
    for got, want in ((lj, g["lj_pot"]), (vir, g["lj_vir"]), (qq, g["qq_pot"])):
        scale = np.maximum(np.abs(want), 0.02 * np.abs(want).max())
        assert (np.abs(got - want) / scale).max() < 1e-13
    # totals of potential(): Σ rows / 2
    assert rel(lj.sum() / 2, g["lj_pot"].sum() / 2) < 1e-13
    assert rel(qq.sum() / 2, g["qq_pot"].sum() / 2) < 1e-12
    p = ora.potential_ewald(s, ora_ewald(ms.box), 10.0, 10.0, ms.box)
    assert rel(p.lj, g["lj_pot"].sum() / 2) < 1e-13 and rel(p.real, g["qq_pot"].sum() / 2 * systems.FACTOR) < 1e-12


def test_monatomic_potential_vs_numpy():
    at = systems.lj_lattice(343, 0.75, 2.5)
    rng = np.random.default_rng(5)
    at.r[:] = (at.r + rng.normal(0, 0.05, at.r.shape)) % at.box
    e, v = ora.potential_atoms(at.r, at.eps, at.sig, at.box, at.r_cut, 2)
    es = sum(npr.lj_atom(i, at.r, at.eps, at.sig, at.box, at.r_cut)[0] for i in range(1, at.n + 1)) / 2
    assert rel(e, es) < 1e-12


def test_loop_invariant_sum_of_deltas_equals_recompute():
    """Poly/main.jl:232-235 invariant: running energy == fresh potential() after a block of moves."""
    ms = systems.load_nist(1)           # 100 molecules, L = 20, use rc = 9 (< L/2)
    s = ora_system(ms)
    ew = ora_ewald(ms.box)
    p0 = ora.potential_ewald(s, ew, 9.0, 9.0, ms.box)
    prm = ora.LoopParams(298.15, 0.316555789, 0.05, 0.5, 1.0, 9.0, 9.0, ms.box, 0, 1)
    quat = ms.quat.copy()
    u = np.random.default_rng(11234).random(8000)
    rc, acc, delta, st = ora.loop(s, ew, ms.db, quat, prm, u, 500, p0.energy, p0.virial)
    assert rc == 0 and st.n_moves == 500
    assert 0 < st.n_accepted < 500
    assert st.rot_attempt > 0 and st.trans_attempt > 0
    p1 = ora.potential_ewald(s, ora_ewald(ms.box), 9.0, 9.0, ms.box)
    assert abs(st.total_energy - p1.energy) < 1e-3          # the reference's own tolerance
    assert rel(st.total_energy, p1.energy) < 1e-11
    # resident S(k) after 500 delta updates == fresh rebuild
    ew2 = ora_ewald(ms.box)
    ora.RecipLong(ew2, s.coords, s.charge, ms.box)
    assert np.allclose(ew.sum_old, ew2.sum_old, rtol=0, atol=1e-9)


@pytest.mark.parametrize("n_short", [700, 701, 702, 703, 704, 705])
def test_loop_short_stream_resumes_exactly(n_short):
    """A finite uniform stream that ends inside a move: the move never happened (state, counters, records untouched),
    uniforms_used is the position at the start of that move, and resuming there with the rest of the stream gives the
    record of a one-shot run.  (The reference's RNG is endless; this is the rule for recorded streams — ADVICE r1.)"""
    ms = systems.load_nist(1)
    prm = ora.LoopParams(298.15, 0.316555789, 0.05, 0.5, 1.0, 9.0, 9.0, ms.box, 0, 0)
    u = np.random.default_rng(7).random(3000)

    def fresh():
        s, ew = ora_system(ms), ora_ewald(ms.box)
        ora.RecipLong(ew, s.coords, s.charge, ms.box)
        return s, ew, ms.quat.copy()
    s1, ew1, q1 = fresh()
    rc, acc1, del1, st1 = ora.loop(s1, ew1, ms.db, q1, prm, u, 400, 0.0, 0.0)
    assert rc == 0 and st1.n_moves == 400
    s2, ew2, q2 = fresh()
    rc, acc2, del2, st2 = ora.loop(s2, ew2, ms.db, q2, prm, u[:n_short], 400, 0.0, 0.0)
    assert rc == 1 and 0 < st2.n_moves < 400 and st2.uniforms_used <= n_short
    n_a, used = st2.n_moves, st2.uniforms_used
    assert st2.trans_attempt + st2.rot_attempt == n_a          # the abandoned move is not counted
    # resume: molecules continue in sweep order, so the second block starts at molecule n_a % N; rotate the system view
    rc, acc3, del3, st3 = _resume(ora, s2, ew2, ms, q2, prm, u[used:], 400 - n_a, n_a)
    assert rc == 0
    assert np.array_equal(np.concatenate([acc2[:n_a], acc3]), acc1)
    # (the rotated molecule order changes the summation order of the j loop: deltas agree to rounding)
    assert np.abs(np.concatenate([del2[:n_a], del3]) - del1).max() < 1e-9
    assert np.array_equal(s2.coords, s1.coords) and np.array_equal(q2, q1)
    assert st2.n_accepted + st3.n_accepted == st1.n_accepted


def _resume(ora, s, ew, ms, quat, prm, u, n_moves, first):
    """Continue a block of moves at move index `first` (sweep order i = first % N + 1): the oracle's Loop always starts at
    molecule 1, so the molecules are rotated by `first` for the call and rotated back afterwards."""
    n = ms.n_mol
    k = first % n
    roll = lambda a, per: np.roll(a.reshape(n, per, -1), -k, axis=0).reshape(a.shape).copy()
    s.coords[:] = roll(s.coords, 3); s.com[:] = roll(s.com, 1); quat[:] = roll(quat, 1)
    db = roll(ms.db, 3)
    out = ora.loop(s, ew, db, quat, prm, u, n_moves, 0.0, 0.0)
    unroll = lambda a, per: np.roll(a.reshape(n, per, -1), k, axis=0).reshape(a.shape).copy()
    s.coords[:] = unroll(s.coords, 3); s.com[:] = unroll(s.com, 1); quat[:] = unroll(quat, 1)
    return out


# ---- the reference's random stream (Julia MersenneTwister = dSFMT-19937) -------------------

def test_julia_rng_known_answers():
    """Values Julia (1.x <= 1.6) prints: the manual's `rand(MersenneTwister(1234), 2)` and
    `Random.seed!(0); rand(4)`.  One matching Float64 already fixes seeding, recursion and masks."""
    from oracle.julia_rng import julia_rand
    assert julia_rand(1234, 3).tolist() == [0.5908446386657102, 0.7667970365022592, 0.5662374165061859]
    assert julia_rand(0, 4).tolist() == [0.8236475079774124, 0.9103565379264364, 0.16456579813368521,
                                         0.17732884646626457]


def test_julia_rng_library_matches_oracle():
    """mmc_julia_rand (host C in the library's driver stand-in) == the oracle's Python restatement,
    across several state regenerations (382 doubles each), a 64-bit seed, and with `skip`."""
    from metropolismontecarlo_b200.energy import julia_rand as lib_rand
    from oracle.julia_rng import julia_rand
    for seed in (0, 1234, 11234, (1 << 32) + 5, (1 << 63) + 12345):
        want = julia_rand(seed, 1500)
        got = lib_rand(seed, 1500)
        assert np.array_equal(want, got), seed
        assert np.array_equal(lib_rand(seed, 400, skip=381), want[381:781])
    u = lib_rand(11234, 200000)
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 5e-3


def test_top_and_pdb_readers(tmp_path):
    """GROMACS .top / .pdb subset readers (Ewald/setup.jl ReadTopFile / ReadPDB) and the vdwTable conversion of
    Ewald/main.jl:183-186, on a file in the reference's own format; the committed TIP3P fixture was parsed by the
    same readers from the reference's water.top + tip3p.pdb (tests/golden/make_fixtures.py)."""
    top = tmp_path / "w.top"
    top.write_text("""; comment
[ defaults ]
  1   2   yes   0.5   0.8333

[ atomtypes ]
; name   mass      charge     ptype  sigma(nm)     epsilon(kJ/mol)
OX        OX    15.99940   -0.8340      A   0.315061     0.6364000
HX        HX     1.008000   0.4170       A   0.0000       0.0000000

[ moleculetype ]
WAT\t3

[ atoms ]
     1  OX   1    WAT      OW      1      -0.8340\t\t15.99940
     2  HX   1    WAT      HW      1       0.4170\t\t1.0080
     3  HX   1    WAT      HW      1       0.4170\t\t1.0080

#ifndef FLEXIBLE
[ settles ]
1\t1\t0.09572\t0.15139
#endif

[ molecules ]
WAT 1000
""")
    pdb = tmp_path / "w.pdb"
    pdb.write_text("TITLE x\nATOM      1  OW   WAT    1      -4.369   0.061  -0.042  0.00  0.00       WAT O\n"
                   "ATOM      2  HW   WAT    1      -3.370   0.049   0.000  0.00  0.00       WAT H\n"
                   "ATOM      3  HW   WAT    1      -4.743  -0.180   0.854  0.00  0.00       WAT H\nEND\n")
    m = systems.model_from_files(top, pdb)
    ref = systems.tip3p_model()
    assert m["eps_kj"] == ref["eps_kj"] and m["sig_nm"] == ref["sig_nm"]
    assert [tuple(a) for a in m["atoms"]] == [tuple(a) for a in ref["atoms"]] and m["xyz"] == ref["xyz"]
    assert systems.read_top(top)["counts"] == [("WAT", 1000)]
    eps, sig = systems.tables_from_types(m["eps_kj"], m["sig_nm"])
    assert abs(eps[0, 0] - 76.5413) < 1e-3 and abs(sig[0, 0] - 3.15061) < 1e-12 and eps[0, 1] == 0.0
    real = Path("/root/reference/water.top")
    if real.exists():                                  # build container only
        got = systems.model_from_files(real, real.parent / "tip3p.pdb", "WAT")
        assert got["xyz"] == ref["xyz"] and got["eps_kj"] == ref["eps_kj"] and [list(a) for a in got["atoms"]] == ref["atoms"]


def test_tip3p_lattice_is_rigid_and_neutral():
    ms = systems.rigid_lattice(systems.tip3p_model(), 125)
    assert abs(ms.charge.sum()) < 1e-12 and ms.n_sites == 375
    d_oh = np.linalg.norm(ms.db[1] - ms.db[0])
    assert abs(d_oh - 0.99996) < 1e-3                   # tip3p.pdb geometry (Å)
    mass = np.array([15.9994, 1.008, 1.008])
    assert np.abs((ms.db[:3] * mass[:, None]).sum(axis=0)).max() < 1e-12      # body frame centred on the COM
    s = ora_system(ms)
    e = ora.potential_ewald(s, ora_ewald(ms.box), 9.0, 9.0, ms.box, 2)
    assert np.isfinite(e.energy) and e.lj != 0.0


def test_closed_form_terms_against_40_digit_arithmetic():
    """The terms of potential() that are closed forms of the charges — Wolf's two constants (Ewald/energy.jl:925-934), EwaldSelf
    (Ewald/ewalds.jl:829-833) — and the opt-in intramolecular correction, evaluated in 40-digit mpmath arithmetic on the float64
    inputs of coord750 (erfc, erf, sqrt(pi) from mpmath, not from libm): pins what the NIST energies and the row pin leave open."""
    mpmath = pytest.importorskip("mpmath")
    from mpmath import mp, mpf
    mp.dps = 40
    ms = systems.load_nist(4)
    s = ora_system(ms)
    ew = ora_ewald(ms.box)
    kappa, factor, rc = mpf(float(ew.kappa)), mpf(float(ew.factor)), mpf(10.0)
    q = [mpf(float(v)) for v in ms.charge]
    sq, sq2 = sum(q), sum(v * v for v in q)
    # prefactor = −Σ_i Σ_j q_i q_j erfc(κ r_c)/r_c ;  prefactor2 = (erfc(κ r_c)/(2 r_c) + κ/√π) Σ q²
    wolf = (-(sq * sq) * mpmath.erfc(kappa * rc) / rc - (mpmath.erfc(kappa * rc) / 2 / rc + kappa / mpmath.sqrt(mpmath.pi)) * sq2) * factor
    w = ora.potential_wolf(s, ew, 10.0, 10.0, ms.box, n_threads=4)
    assert abs(mpf(w.wolf_const) - wolf) < mpf(1e-12) * abs(wolf)
    self_ = -kappa / mpmath.sqrt(mpmath.pi) * sq2 * factor
    assert abs(mpf(ora.EwaldSelf(ew, ms.charge)) - self_) < mpf(1e-13) * abs(self_)
    # −factor Σ_mol Σ_{a<b} q_a q_b erf(κ r_ab)/r_ab, r_ab by the reference's minimum image (the NIST files store wrapped molecules)
    def v1d(c1, c2, box):
        if c1 < c2:
            return (c2 - c1) if (c2 - c1) < (c1 - c2 + box) else (c2 - c1 - box)
        return (c2 - c1) if (c1 - c2) < (c2 - c1 + box) else (c2 - c1 + box)
    box = mpf(float(ms.box))
    xyz = [[mpf(float(v)) for v in row] for row in ms.coords]
    intra = mpf(0)
    for m in range(ms.n_mol):
        a0, a1 = int(ms.first_atom[m]) - 1, int(ms.last_atom[m])
        for a in range(a0, a1):
            for b in range(a + 1, a1):
                d = [v1d(xyz[a][k], xyz[b][k], box) for k in range(3)]
                r = mpmath.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
                intra += q[a] * q[b] * mpmath.erf(kappa * r) / r
    intra *= -factor
    got = ora.EwaldIntra(s, float(ew.kappa), float(ew.factor), ms.box)
    assert got > 0 and abs(mpf(got) - intra) < mpf(1e-12) * abs(intra)


def test_recip_long_against_30_digit_arithmetic():
    """RecipLong (Ewald/ewalds.jl:538-604) and the cfac table of PrepareEwaldVariables (:45-103) on the smallest NIST box (100
    molecules, L = 20 Å): every ρ(k) = Σ_l q_l e^{i 2π k·r_l / L} and E = Σ_k cfac_k |ρ(k)|² in 30-digit mpmath arithmetic with
    direct exponentials — the exact value of the formula the reference evaluates by recurrence in float64."""
    mpmath = pytest.importorskip("mpmath")
    from mpmath import mp, mpf
    mp.dps = 30
    ms = systems.load_nist(1)
    ew = ora_ewald(ms.box)
    got_e = ora.RecipLong(ew, ms.coords, ms.charge, ms.box)
    L, kappa = mpf(float(ms.box)), mpf(float(ew.kappa))
    b = 1 / (4 * kappa * kappa * L * L)
    twopi = 2 * mpmath.pi
    q = [mpf(float(v)) for v in ms.charge]
    ph = [[twopi * mpf(float(v)) / L for v in row] for row in ms.coords]
    # e^{i k x} for k = -5..5 per site and axis, from one exponential each (exact to 30 digits)
    tab = [[[mpmath.expj(k * p) for k in range(-5, 6)] for p in row] for row in ph]
    energy, worst = mpf(0), mpf(0)
    terms = []
    kx_ky_kz = [(kx, ky, kz) for kx in range(0, 6) for ky in range(-5, 6) for kz in range(-5, 6) if 0 < kx * kx + ky * ky + kz * kz < 27]
    assert len(kx_ky_kz) == ew.nkvecs
    qsum = sum(abs(v) for v in q)
    for i, (kx, ky, kz) in enumerate(kx_ky_kz):
        term = sum(q[l] * tab[l][0][kx + 5] * tab[l][1][ky + 5] * tab[l][2][kz + 5] for l in range(len(q)))
        k_sq = kx * kx + ky * ky + kz * kz
        kr_sq = twopi * twopi * k_sq
        cfac = twopi * mpmath.exp(-b * kr_sq) / kr_sq / L * (2 if kx > 0 else 1)
        terms.append((term, cfac))
        energy += cfac * (term.real * term.real + term.imag * term.imag)
        worst = max(worst, abs(mpf(float(ew.sum_new[i, 0])) - term.real), abs(mpf(float(ew.sum_new[i, 1])) - term.imag))
        assert abs(mpf(float(ew.cfac[i])) - cfac) < mpf(1e-14) * cfac
    assert worst < mpf(1e-13) * qsum
    assert abs(mpf(got_e) - energy) < mpf(1e-12) * energy
    # RecipMove (ewalds.jl:718-826) for a rigid displacement of molecule 7: Σ_k cfac (|ρ_new|² − |ρ_old|²) · factor, ρ_new = ρ_old + Δ
    sl = slice(int(ms.first_atom[6]) - 1, int(ms.last_atom[6]))
    r_old = ms.coords[sl].copy()
    r_new = r_old + np.array([0.21, -0.13, 0.08])
    got_d = ora.RecipMove(ms.box, ew, r_old, r_new, ms.charge[sl])
    def e3(r, kx, ky, kz):
        return mpmath.expj(twopi * (kx * mpf(float(r[0])) + ky * mpf(float(r[1])) + kz * mpf(float(r[2]))) / L)
    delta = mpf(0)
    for (kx, ky, kz), (term, cfac) in zip(kx_ky_kz, terms):
        new = term + sum(mpf(float(qa)) * (e3(rn, kx, ky, kz) - e3(ro, kx, ky, kz)) for qa, ro, rn in zip(ms.charge[sl], r_old, r_new))
        delta += cfac * ((new.real * new.real + new.imag * new.imag) - (term.real * term.real + term.imag * term.imag))
    delta *= mpf(float(ew.factor))
    assert abs(mpf(got_d) - delta) < mpf(1e-10) * max(abs(delta), mpf(1))
