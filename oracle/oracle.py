"""ctypes loader for the CPU ORACLE (oracle/mmc_oracle.c).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by the product package
(metropolismontecarlo_b200/).  See oracle/mmc_oracle.h for pinning status.

All indices follow the reference (Julia): molecule index i and atom indices are
1-based; firstAtom/lastAtom are inclusive.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)
c_int32_p = C.POINTER(C.c_int32)
c_uint8_p = C.POINTER(C.c_uint8)


def build(force: bool = False) -> Path:
    """gcc-compile the oracle in place (seconds). gcc exists here and on the GPU box."""
    so = _HERE / "libmmc_oracle.so"
    src = _HERE / "mmc_oracle.c"
    hdr = _HERE / "mmc_oracle.h"
    if force or not so.exists() or so.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-fopenmp", "-std=gnu11", "-shared",
               "-o", str(so), str(src), "-lm"]
        subprocess.run(cmd, check=True, cwd=str(_HERE))
    return so


class Properties(C.Structure):
    _fields_ = [("energy", C.c_double), ("virial", C.c_double), ("coulomb", C.c_double),
                ("lj", C.c_double), ("real", C.c_double), ("recip", C.c_double),
                ("self_", C.c_double), ("wolf_const", C.c_double), ("overlaps", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class _System(C.Structure):
    _fields_ = [("n_mol", C.c_int64), ("n_sites", C.c_int64), ("coords", c_double_p),
                ("charge", c_double_p), ("atype", c_int64_p), ("first_atom", c_int64_p),
                ("last_atom", c_int64_p), ("com", c_double_p), ("n_types", C.c_int64),
                ("eps", c_double_p), ("sig", c_double_p)]


class _Ewald(C.Structure):
    _fields_ = [("kappa", C.c_double), ("nk", C.c_int64), ("k_sq_max", C.c_int64),
                ("nkvecs", C.c_int64), ("kxyz", c_int32_p), ("cfac", c_double_p),
                ("sum_old", c_double_p), ("sum_new", c_double_p), ("factor", C.c_double)]


class LoopParams(C.Structure):
    _fields_ = [("temperature", C.c_double), ("dr_max", C.c_double), ("dphi_max", C.c_double),
                ("p_trans", C.c_double), ("p_rot", C.c_double), ("lj_rcut", C.c_double),
                ("qq_rcut", C.c_double), ("box", C.c_double), ("style", C.c_int), ("adjust", C.c_int)]


class LoopStats(C.Structure):
    _fields_ = [("n_moves", C.c_int64), ("n_accepted", C.c_int64), ("n_overlap", C.c_int64),
                ("uniforms_used", C.c_int64), ("trans_attempt", C.c_int64),
                ("trans_accept", C.c_int64), ("rot_attempt", C.c_int64), ("rot_accept", C.c_int64),
                ("dr_max", C.c_double), ("dphi_max", C.c_double), ("total_energy", C.c_double),
                ("total_virial", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        L = C.CDLL(str(so))
        L.ora_vector1D.restype = C.c_double
        L.ora_vector1D.argtypes = [C.c_double] * 3
        L.ora_RecipLong.restype = C.c_double
        L.ora_RecipMove.restype = C.c_double
        L.ora_EwaldSelf.restype = C.c_double
        L.ora_count_kvecs.restype = C.c_int64
        L.ora_count_kvecs.argtypes = [C.c_int64, C.c_int64]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _ip(a):
    return a.ctypes.data_as(c_int64_p)


class System:
    """The reference's soa/moa + vdwTable, as contiguous numpy arrays (A.6 layouts)."""

    def __init__(self, coords, charge, atype, first_atom, last_atom, com, eps, sig):
        self.coords = np.ascontiguousarray(coords, dtype=np.float64).reshape(-1, 3).copy()
        self.charge = np.ascontiguousarray(charge, dtype=np.float64).copy()
        self.atype = np.ascontiguousarray(atype, dtype=np.int64).copy()
        self.first_atom = np.ascontiguousarray(first_atom, dtype=np.int64).copy()
        self.last_atom = np.ascontiguousarray(last_atom, dtype=np.int64).copy()
        self.com = np.ascontiguousarray(com, dtype=np.float64).reshape(-1, 3).copy()
        eps = np.asarray(eps, dtype=np.float64)
        self.n_types = eps.shape[0]
        # column-major nt x nt like a Julia Matrix
        self.eps = np.asfortranarray(eps).ravel(order="F").copy()
        self.sig = np.asfortranarray(np.asarray(sig, dtype=np.float64)).ravel(order="F").copy()
        self.n_mol = self.com.shape[0]
        self.n_sites = self.coords.shape[0]

    def copy(self):
        nt = self.n_types
        return System(self.coords, self.charge, self.atype, self.first_atom, self.last_atom,
                      self.com, self.eps.reshape(nt, nt, order="F"), self.sig.reshape(nt, nt, order="F"))

    def c(self):
        return _System(self.n_mol, self.n_sites, _dp(self.coords), _dp(self.charge),
                       _ip(self.atype), _ip(self.first_atom), _ip(self.last_atom), _dp(self.com),
                       self.n_types, _dp(self.eps), _dp(self.sig))


class Ewald:
    """mutable struct EWALD (Ewald/ewalds.jl:9-19) after PrepareEwaldVariables."""

    def __init__(self, kappa, nk, k_sq_max, factor, box):
        self.kappa, self.nk, self.k_sq_max, self.factor = float(kappa), int(nk), int(k_sq_max), float(factor)
        n = lib().ora_count_kvecs(self.nk, self.k_sq_max)
        self.kxyz = np.zeros((n, 3), dtype=np.int32)
        self.cfac = np.zeros(n, dtype=np.float64)
        self.sum_old = np.zeros((n, 2), dtype=np.float64)
        self.sum_new = np.zeros((n, 2), dtype=np.float64)
        self.nkvecs = n
        e = self.c()
        lib().ora_PrepareEwaldVariables(C.byref(e), C.c_double(box))
        assert e.nkvecs == n

    def c(self):
        return _Ewald(self.kappa, self.nk, self.k_sq_max, self.nkvecs,
                      self.kxyz.ctypes.data_as(c_int32_p), _dp(self.cfac), _dp(self.sum_old),
                      _dp(self.sum_new), self.factor)


# Ewald/constants.jl:24-28
def factor() -> float:
    kb1 = 1.3806488e-23
    e01 = 8.854187817e-12
    e01 *= 1e-10
    e1 = 1.602176565e-19
    return e1 ** 2 / e01 / 4 / np.pi / kb1


def vector1D(c1, c2, box):
    return lib().ora_vector1D(c1, c2, box)


def LJ_poly_dU(i, s: System, r_cut, box):
    pot, vir = C.c_double(), C.c_double()
    cs = s.c()
    lib().ora_LJ_poly_dU(C.c_int64(i), C.byref(cs), C.c_double(r_cut), C.c_double(box),
                         C.byref(pot), C.byref(vir))
    return pot.value, vir.value


def EwaldReal(i, s: System, kappa, r_cut, box):
    pot, ov = C.c_double(), C.c_int()
    cs = s.c()
    lib().ora_EwaldReal(C.c_int64(i), C.byref(cs), C.c_double(kappa), C.c_double(r_cut),
                        C.c_double(box), C.byref(pot), C.byref(ov))
    return pot.value, bool(ov.value)


def EwaldShort(i, s: System, ew: Ewald, qq_rcut, box):
    e, v, ov = C.c_double(), C.c_double(), C.c_int()
    cs, ce = s.c(), ew.c()
    lib().ora_EwaldShort(C.c_int64(i), C.byref(cs), C.byref(ce), C.c_double(qq_rcut),
                         C.c_double(box), C.byref(e), C.byref(v), C.byref(ov))
    return e.value, v.value, bool(ov.value)


def EwaldIntra(s: System, kappa, factor, box):
    """Intramolecular correction (not in the reference; twin of the engine's opt-in flag): -factor Σ_mol Σ_{a<b} q_a q_b erf(κ r)/r,
    r by the reference's minimum image."""
    cs = s.c()
    f = lib().ora_EwaldIntraBox
    f.restype = C.c_double
    return f(C.byref(cs), C.c_double(kappa), C.c_double(factor), C.c_double(box))


def RecipLong(ew: Ewald, r, q, box):
    r = np.ascontiguousarray(r, dtype=np.float64)
    q = np.ascontiguousarray(q, dtype=np.float64)
    ce = ew.c()
    return lib().ora_RecipLong(C.byref(ce), C.c_int64(q.shape[0]), _dp(r), _dp(q), C.c_double(box))


def RecipMove(box, ew: Ewald, r_old, r_new, q):
    r_old = np.ascontiguousarray(r_old, dtype=np.float64)
    r_new = np.ascontiguousarray(r_new, dtype=np.float64)
    q = np.ascontiguousarray(q, dtype=np.float64)
    ce = ew.c()
    return lib().ora_RecipMove(C.c_double(box), C.byref(ce), C.c_int64(q.shape[0]),
                               _dp(r_old), _dp(r_new), _dp(q))


def EwaldSelf(ew: Ewald, q):
    q = np.ascontiguousarray(q, dtype=np.float64)
    ce = ew.c()
    return lib().ora_EwaldSelf(C.byref(ce), C.c_int64(q.shape[0]), _dp(q))


def recip_commit(ew: Ewald):
    ew.sum_old[:] = ew.sum_new


def recip_rollback(ew: Ewald):
    ew.sum_new[:] = ew.sum_old


def potential_ewald(s: System, ew: Ewald, lj_rcut, qq_rcut, box, n_threads=1):
    out = Properties()
    cs, ce = s.c(), ew.c()
    lib().ora_potential_ewald(C.byref(cs), C.byref(ce), C.c_double(lj_rcut), C.c_double(qq_rcut),
                              C.c_double(box), C.c_int(n_threads), C.byref(out))
    return out


def potential_wolf(s: System, ew: Ewald, lj_rcut, qq_rcut, box, n_threads=1):
    out = Properties()
    cs, ce = s.c(), ew.c()
    lib().ora_potential_wolf(C.byref(cs), C.byref(ce), C.c_double(lj_rcut), C.c_double(qq_rcut),
                             C.c_double(box), C.c_int(n_threads), C.byref(out))
    return out


def potential_rows(s: System, kappa, lj_rcut, qq_rcut, box, i0, i1, n_threads=1):
    """Rows i0..i1-1 (0-based) of the two O(N^2) loops in potential(); kappa<0 skips Coulomb."""
    lj, vir, real, nov = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
    cs = s.c()
    lib().ora_potential_rows(C.byref(cs), C.c_double(kappa), C.c_double(lj_rcut),
                             C.c_double(qq_rcut), C.c_double(box), C.c_int64(i0), C.c_int64(i1),
                             C.c_int(n_threads), C.byref(lj), C.byref(vir), C.byref(real), C.byref(nov))
    return lj.value, vir.value, real.value, nov.value


def LJ_dU_atom(i, r, eps, sig, box, r_cut):
    r = np.ascontiguousarray(r, dtype=np.float64)
    eps = np.ascontiguousarray(eps, dtype=np.float64)
    sig = np.ascontiguousarray(sig, dtype=np.float64)
    pot, vir = C.c_double(), C.c_double()
    lib().ora_LJ_dU_atom(C.c_int64(i), C.c_int64(eps.shape[0]), _dp(r), _dp(eps), _dp(sig),
                         C.c_double(box), C.c_double(r_cut), C.byref(pot), C.byref(vir))
    return pot.value, vir.value


def potential_atoms(r, eps, sig, box, r_cut, n_threads=1):
    r = np.ascontiguousarray(r, dtype=np.float64)
    eps = np.ascontiguousarray(eps, dtype=np.float64)
    sig = np.ascontiguousarray(sig, dtype=np.float64)
    e, v = C.c_double(), C.c_double()
    lib().ora_potential_atoms(C.c_int64(eps.shape[0]), _dp(r), _dp(eps), _dp(sig), C.c_double(box),
                              C.c_double(r_cut), C.c_int(n_threads), C.byref(e), C.byref(v))
    return e.value, v.value


def volume_scale(s: System, box_old, box_new):
    cs = s.c()
    lib().ora_volume_scale(C.byref(cs), C.c_double(box_old), C.c_double(box_new))


def loop(s: System, ew: Ewald, db, quat, params: LoopParams, uniforms, n_moves, e0=0.0, v0=0.0):
    """Ewald/main.jl Loop restatement. Mutates s, ew, quat. Returns (rc, accepted, delta, stats)."""
    db = np.ascontiguousarray(db, dtype=np.float64)
    assert quat.dtype == np.float64 and quat.flags.c_contiguous
    uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
    acc = np.zeros(n_moves, dtype=np.uint8)
    delta = np.zeros(n_moves, dtype=np.float64)
    st = LoopStats()
    cs, ce = s.c(), ew.c()
    rc = lib().ora_loop(C.byref(cs), C.byref(ce), _dp(db), _dp(quat), C.byref(params),
                        _dp(uniforms), C.c_int64(uniforms.shape[0]), C.c_int64(n_moves),
                        C.c_double(e0), C.c_double(v0), acc.ctypes.data_as(c_uint8_p), _dp(delta),
                        C.byref(st))
    return rc, acc, delta, st


def loop_atoms(r, eps, sig, box, r_cut, temperature, dr_max, uniforms, n_moves, e0=0.0, v0=0.0):
    assert r.dtype == np.float64 and r.flags.c_contiguous
    eps = np.ascontiguousarray(eps, dtype=np.float64)
    sig = np.ascontiguousarray(sig, dtype=np.float64)
    uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
    acc = np.zeros(n_moves, dtype=np.uint8)
    delta = np.zeros(n_moves, dtype=np.float64)
    st = LoopStats()
    rc = lib().ora_loop_atoms(C.c_int64(eps.shape[0]), _dp(r), _dp(eps), _dp(sig), C.c_double(box),
                              C.c_double(r_cut), C.c_double(temperature), C.c_double(dr_max),
                              _dp(uniforms), C.c_int64(uniforms.shape[0]), C.c_int64(n_moves),
                              C.c_double(e0), C.c_double(v0), acc.ctypes.data_as(c_uint8_p),
                              _dp(delta), C.byref(st))
    return rc, acc, delta, st


def q_to_a(q):
    q = np.ascontiguousarray(q, dtype=np.float64)
    a = np.zeros(9)
    lib().ora_q_to_a(_dp(q), _dp(a))
    return a.reshape(3, 3)
