"""Restart / output files and tail corrections: host-side, CPU only."""
import math

import numpy as np

from metropolismontecarlo_b200 import checkpoint as ck
from metropolismontecarlo_b200.systems import load_nist, spce_lattice


def test_print_pdb_layout(tmp_path):
    """Line layout of PrintPDB (Ewald/initialConfigurations.jl:160-181): the strings below are the reference's
    @sprintf patterns filled in by hand for the first atom of the lattice."""
    ms = spce_lattice(8)
    path = ck.print_pdb(ms, ms.box, 7, str(tmp_path / "final"), atom_names=["O", "H"])
    assert path.endswith("final_7.pdb")
    lines = open(path).read().splitlines()
    assert len(lines) == 1 + ms.n_sites
    b = "%7.3f" % ms.box
    assert lines[0] == f"CRYST1  {b} {b} {b} 90.00  90.00  90.00 P 1           1 "
    x, y, z = ("%7.3f" % v for v in ms.coords[0])
    assert lines[1] == f"ATOM      1   O  SOL     1     {x} {y} {z}  1.00  0.00 "
    assert lines[6].split()[:5] == ["ATOM", "6", "H", "SOL", "2"]


def test_print_output_round_trip(tmp_path):
    ms = spce_lattice(27)
    path = ck.print_output(ms, ms.quat, ms.box, 3, str(tmp_path / "xyz_quat"))
    lines = open(path).read().splitlines()
    assert lines[1] == " Molecular coordinates and quaternions"
    assert lines[2] == " #, mol name, atom Start, atom End, x, y, z, q0, q1, q2, q3"
    assert lines[3 + ms.n_mol] == "Atom coordinates"
    box, com, quat, fa, la, q, r = ck.read_output(path)
    assert abs(box - ms.box) < 1e-3
    assert np.array_equal(fa, ms.first_atom) and np.array_equal(la, ms.last_atom)
    assert np.allclose(com, ms.com, atol=5.1e-4) and np.allclose(quat, ms.quat, atol=5.1e-4)   # %7.3f
    assert np.allclose(q, ms.charge, atol=5.1e-4) and np.allclose(r, ms.coords, atol=5.1e-4)


def test_cnf_round_trip_exact(tmp_path):
    """ReadCNF layout (initialConfigurations.jl:239-280): count, box, then x y z q0 q1 q2 q3; exact=True is lossless."""
    ms = spce_lattice(64)
    p = ck.write_cnf(str(tmp_path / "cnf_input.inp"), ms.com, ms.quat, ms.box, exact=True)
    rm, quat, box = ck.read_cnf(p)
    assert box == ms.box and np.array_equal(rm, ms.com) and np.array_equal(quat, ms.quat)
    p = ck.write_cnf(str(tmp_path / "cnf8.inp"), ms.com, ms.quat, ms.box)
    rm, quat, box = ck.read_cnf(p)
    assert np.allclose(rm, ms.com, atol=1e-8) and np.allclose(quat, ms.quat, atol=1e-8)


def test_monatomic_tail_corrections():
    """Ewald/auxillary.jl:16-35 at rho* = 0.75, rc = 2.5 (config C): closed forms worked by hand."""
    rho, rc = 0.75, 2.5
    sr3 = 1 / 15.625
    assert math.isclose(ck.potential_lrc(rho, rc), math.pi * (8 / 9 * sr3 ** 3 - 8 / 3 * sr3) * rho, rel_tol=1e-15)
    assert math.isclose(ck.pressure_lrc(rho, rc), math.pi * (32 / 9 * sr3 ** 3 - 16 / 3 * sr3) * rho ** 2, rel_tol=1e-15)
    assert math.isclose(ck.pressure_delta(rho, rc), math.pi * 8 / 3 * (sr3 ** 3 - sr3) * rho ** 2, rel_tol=1e-15)
    # the energy correction is the integral of 4(r^-12 - r^-6) 2 pi rho r^2 from rc to infinity
    from scipy.integrate import quad
    val, _ = quad(lambda r: 4 * (r ** -12 - r ** -6) * 2 * math.pi * rho * r * r, rc, np.inf)
    assert math.isclose(ck.potential_lrc(rho, rc), val, rel_tol=1e-9)


def test_polyatomic_tail_corrections_reduce_to_monatomic():
    """ener_corr / press_corr (Ewald/energy.jl:514-603) with one type, eps = sig = 1 equal N * potential_lrc and
    pressure_lrc: the two families of the reference agree where both apply."""
    n, rho, rc = 500, 0.6, 3.0
    box = (n / rho) ** (1 / 3)
    e = ck.ener_corr(np.ones((1, 1)), np.ones((1, 1)), [n], rc, box)
    p = ck.press_corr(np.ones((1, 1)), np.ones((1, 1)), [n], rc, box)
    assert math.isclose(e, n * ck.potential_lrc(rho, rc), rel_tol=1e-12)
    assert math.isclose(p, ck.pressure_lrc(rho, rc), rel_tol=1e-12)
    # SPC/E water (only O-O carries epsilon): negative, a few K per molecule at the NIST density
    ms = load_nist(1)
    ew = ck.ener_corr(ms.eps, ms.sig, [ms.n_mol, 2 * ms.n_mol], 10.0, ms.box)
    assert -40.0 * ms.n_mol < ew < 0.0
