"""Host-side mirror of the reference's energy interface on top of libmmc_b200.so.

The reference is Julia; its toolchain is absent here, so this module plays the Julia side of
the drop-in boundary in Python: same function names, argument meaning, 1-based indices and
return tuples as the reference's energy routines (citations into /root/reference), every one
of them a thin ctypes call into the C ABI (include/mmc_b200.h).  The Julia state ``soa/moa/
ewald`` that the reference mutates in place lives in HBM inside an ``Engine``.

No energy is ever computed in Python; if the CUDA library or a GPU is missing these calls raise.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import (STYLE_EWALD, STYLE_LJ_ATOMS, STYLE_LJ_ONLY, STYLE_WOLF, Counters, LoopParams,
                   LoopStats, MMCError, Properties, TrialResult)
from .systems import ALPHA, FACTOR, K_SQ_MAX, NK, AtomicSystem, MolecularSystem

_dp = lambda a: a.ctypes.data_as(_lib.c_double_p)  # noqa: E731
_ip = lambda a: a.ctypes.data_as(_lib.c_int64_p)   # noqa: E731

_STYLES = {"ewald": STYLE_EWALD, "wolf": STYLE_WOLF, "lj": STYLE_LJ_ONLY, "atoms": STYLE_LJ_ATOMS}


def _style(s):
    return _STYLES[s] if isinstance(s, str) else int(s)


class Engine:
    """One handle = one system resident on one B200 (mmc_create / mmc_destroy)."""

    def __init__(self, device: int = 0, rank: int = 0, world: int = 1, sync_mode: int = 0, stream: int | None = None):
        self.lib = _lib.load()
        cfg = _lib.Config(device, rank, world, sync_mode, stream)
        self.h = _lib.H()
        rc = self.lib.mmc_create(C.byref(cfg), C.byref(self.h))
        if rc != 0:
            raise MMCError(rc, self.lib.mmc_last_error(None).decode())
        self.nkvecs = 0
        self.box = None

    def _ck(self, rc):
        if rc < 0:
            raise MMCError(rc, self.lib.mmc_last_error(self.h).decode())
        return rc

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.mmc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- upload (MakeAtomArrays/MakeTables layouts, Ewald/setup.jl:447-673)
    def upload_system(self, ms: MolecularSystem, rc_lj: float, rc_qq: float | None = None):
        rc_qq = rc_lj if rc_qq is None else rc_qq
        coords = np.ascontiguousarray(ms.coords, dtype=np.float64)
        charge = np.ascontiguousarray(ms.charge, dtype=np.float64)
        atype = np.ascontiguousarray(ms.atype, dtype=np.int64)
        fa = np.ascontiguousarray(ms.first_atom, dtype=np.int64)
        la = np.ascontiguousarray(ms.last_atom, dtype=np.int64)
        com = np.ascontiguousarray(ms.com, dtype=np.float64)
        eps = np.ascontiguousarray(np.asarray(ms.eps, dtype=np.float64).ravel(order="F"))
        sig = np.ascontiguousarray(np.asarray(ms.sig, dtype=np.float64).ravel(order="F"))
        self._ck(self.lib.mmc_upload_system(self.h, ms.n_mol, ms.n_sites, _dp(coords), _dp(charge), _ip(atype),
                                            _ip(fa), _ip(la), _dp(com), ms.eps.shape[0], _dp(eps), _dp(sig),
                                            ms.box, rc_lj, rc_qq))
        self.box = float(ms.box)
        self.n_mol, self.n_sites = ms.n_mol, ms.n_sites

    def upload_positions(self, coords, com):
        """All site coordinates and COMs of the uploaded system at once (bulk mmc_set_molecule, main.jl:527,552)."""
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        com = np.ascontiguousarray(com, dtype=np.float64)
        assert coords.shape == (self.n_sites, 3) and com.shape == (self.n_mol, 3)
        self._ck(self.lib.mmc_upload_positions(self.h, _dp(coords), _dp(com)))

    def potential_host(self, coords, com, style="ewald") -> Properties:
        """upload_positions + potential in one call, copies overlapped with compute (mmc_potential_host)."""
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        com = np.ascontiguousarray(com, dtype=np.float64)
        assert coords.shape == (self.n_sites, 3) and com.shape == (self.n_mol, 3)
        out = Properties()
        self._ck(self.lib.mmc_potential_host(self.h, _dp(coords), _dp(com), _style(style), C.byref(out)))
        return out

    def last_host_bytes(self) -> int:
        """Bytes the last potential_host copied host -> device on this rank."""
        n = C.c_int64()
        self._ck(self.lib.mmc_last_host_bytes(self.h, C.byref(n)))
        return n.value

    def upload_atoms(self, at: AtomicSystem):
        r = np.ascontiguousarray(at.r, dtype=np.float64)
        e = np.ascontiguousarray(at.eps, dtype=np.float64)
        s = np.ascontiguousarray(at.sig, dtype=np.float64)
        self._ck(self.lib.mmc_upload_atoms(self.h, at.n, _dp(r), _dp(e), _dp(s), at.box, at.r_cut))
        self.n_atoms = at.n

    def download_system(self):
        coords = np.empty((self.n_sites, 3))
        com = np.empty((self.n_mol, 3))
        self._ck(self.lib.mmc_download_system(self.h, _dp(coords), _dp(com)))
        return coords, com

    def download_atoms(self):
        r = np.empty((self.n_atoms, 3))
        self._ck(self.lib.mmc_download_atoms(self.h, _dp(r)))
        return r

    # ---- Ewald/ewalds.jl:45-103
    def PrepareEwaldVariables(self, kappa: float, nk: int = NK, k_sq_max: int = K_SQ_MAX, factor: float = FACTOR):
        n = C.c_int32()
        self._ck(self.lib.mmc_ewald_prepare(self.h, kappa, nk, k_sq_max, factor, C.byref(n)))
        self.nkvecs = n.value
        self.kappa, self.factor = kappa, factor
        return n.value

    def kvectors(self):
        k = np.empty((self.nkvecs, 3), dtype=np.int32)
        c = np.empty(self.nkvecs)
        self._ck(self.lib.mmc_get_kvectors(self.h, k.ctypes.data_as(_lib.c_int32_p), _dp(c)))
        return k, c

    def rhok(self):
        """(sumQExpOld, sumQExpNew) as complex arrays."""
        o = np.empty((self.nkvecs, 2))
        n = np.empty((self.nkvecs, 2))
        self._ck(self.lib.mmc_get_rhok(self.h, _dp(o), _dp(n)))
        return o[:, 0] + 1j * o[:, 1], n[:, 0] + 1j * n[:, 1]

    # ---- single-molecule drop-ins
    def LJ_poly_ΔU(self, i: int):
        """Ewald/energy.jl:209-290 → (energy, virial)."""
        p, v = C.c_double(), C.c_double()
        self._ck(self.lib.mmc_lj_mol(self.h, i, C.byref(p), C.byref(v)))
        return p.value, v.value

    LJ_poly_dU = LJ_poly_ΔU

    def EwaldReal(self, i: int):
        """Ewald/ewalds.jl:293-376 → (pot un-scaled, overlap)."""
        p, o = C.c_double(), C.c_int32()
        self._ck(self.lib.mmc_ewald_real(self.h, i, C.byref(p), C.byref(o)))
        return p.value, bool(o.value)

    def EwaldShort(self, i: int):
        """Ewald/ewalds.jl:892-910 → (partial_e, partial_v, overlap)."""
        e, v, o = C.c_double(), C.c_double(), C.c_int32()
        self._ck(self.lib.mmc_ewald_short(self.h, i, C.byref(e), C.byref(v), C.byref(o)))
        return e.value, v.value, bool(o.value)

    def set_molecule(self, i: int, com, sites):
        """The in-place writes of Ewald/main.jl:527,552 / :623-624."""
        com = np.ascontiguousarray(com, dtype=np.float64)
        sites = np.ascontiguousarray(sites, dtype=np.float64)
        self._ck(self.lib.mmc_set_molecule(self.h, i, _dp(com), _dp(sites)))

    def RecipLong(self):
        """Ewald/ewalds.jl:538-604 → un-scaled energy (ρ(k) stored to Old and New)."""
        e = C.c_double()
        self._ck(self.lib.mmc_recip_long(self.h, C.byref(e)))
        return e.value

    def RecipMove(self, r_old, r_new, q):
        """Ewald/ewalds.jl:718-826 → energy·factor."""
        r_old = np.ascontiguousarray(r_old, dtype=np.float64)
        r_new = np.ascontiguousarray(r_new, dtype=np.float64)
        q = np.ascontiguousarray(q, dtype=np.float64)
        e = C.c_double()
        self._ck(self.lib.mmc_recip_move(self.h, _dp(r_old), _dp(r_new), _dp(q), q.shape[0], C.byref(e)))
        return e.value

    def recip_commit(self):
        self._ck(self.lib.mmc_recip_commit(self.h))

    def recip_rollback(self):
        self._ck(self.lib.mmc_recip_rollback(self.h))

    def EwaldSelf(self):
        """Ewald/ewalds.jl:829-833."""
        e = C.c_double()
        self._ck(self.lib.mmc_ewald_self(self.h, C.byref(e)))
        return e.value

    # ---- monatomic
    def LJ_ΔU(self, i: int):
        """Monatomic/mainMonatomic.jl:227-272 → (energy, virial)."""
        p, v = C.c_double(), C.c_double()
        self._ck(self.lib.mmc_lj_atom(self.h, i, C.byref(p), C.byref(v)))
        return p.value, v.value

    LJ_dU = LJ_ΔU

    def set_atom(self, i: int, r):
        r = np.ascontiguousarray(r, dtype=np.float64)
        self._ck(self.lib.mmc_set_atom(self.h, i, _dp(r)))

    # ---- totals
    def potential(self, style="ewald") -> Properties:
        """Ewald/energy.jl:946-1032 ("ewald"), :864-943 ("wolf"), mainMonatomic.jl:275-289 ("atoms")."""
        out = Properties()
        self._ck(self.lib.mmc_potential(self.h, _style(style), C.byref(out)))
        return out

    def energy_all(self, style="ewald"):
        """LJ_poly_ΔU(i) (energy.jl:209-290) and EwaldShort(i)[1] (ewalds.jl:892-910) for every molecule at once:
        (lj_pot[n_mol], lj_vir[n_mol], coul[n_mol], overlap[n_mol]) in molecule order."""
        n = self.n_mol
        lj, vir, qq = (np.empty(n) for _ in range(3))
        ov = np.empty(n, dtype=np.int32)
        self._ck(self.lib.mmc_energy_all(self.h, _style(style), _dp(lj), _dp(vir), _dp(qq),
                                         ov.ctypes.data_as(C.POINTER(C.c_int32))))
        return lj, vir, qq, ov

    def partial_count(self) -> int:
        n = C.c_int64()
        self._ck(self.lib.mmc_partial_count(self.h, C.byref(n)))
        return n.value

    def potential_partial(self, style, d_partials_ptr: int):
        self._ck(self.lib.mmc_potential_partial(self.h, _style(style), C.c_void_p(d_partials_ptr)))

    def potential_finalize(self, style, d_partials_ptr: int):
        out = Properties()
        self._ck(self.lib.mmc_potential_finalize(self.h, _style(style), C.c_void_p(d_partials_ptr), C.byref(out)))
        return out

    # ---- the sharded evaluation with the exchange over NVLink peer memory (include/mmc_b200.h mmc_peer_*)
    def peer_export(self) -> bytes:
        buf = C.create_string_buffer(64)
        self._ck(self.lib.mmc_peer_export(self.h, C.cast(buf, C.c_void_p)))
        return buf.raw

    def peer_import(self, rank: int, handle: bytes):
        buf = C.create_string_buffer(handle, 64)
        self._ck(self.lib.mmc_peer_import(self.h, rank, C.cast(buf, C.c_void_p)))

    def peer_buffer(self) -> int:
        p = C.c_void_p()
        self._ck(self.lib.mmc_peer_buffer(self.h, C.byref(p)))
        return p.value

    def peer_import_ptr(self, rank: int, ptr: int):
        self._ck(self.lib.mmc_peer_import_ptr(self.h, rank, C.c_void_p(ptr)))

    def potential_sharded_begin(self, style):
        self._ck(self.lib.mmc_potential_sharded_begin(self.h, _style(style)))

    def potential_sharded_end(self):
        out = Properties()
        self._ck(self.lib.mmc_potential_sharded_end(self.h, C.byref(out)))
        return out

    def potential_sharded(self, style) -> Properties:
        out = Properties()
        self._ck(self.lib.mmc_potential_sharded(self.h, _style(style), C.byref(out)))
        return out

    # ---- fused trial move
    def trial_move(self, i: int, com_new, sites_new, style="ewald") -> TrialResult:
        com_new = np.ascontiguousarray(com_new, dtype=np.float64)
        sites_new = np.ascontiguousarray(sites_new, dtype=np.float64)
        out = TrialResult()
        self._ck(self.lib.mmc_trial_move(self.h, i, _dp(com_new), _dp(sites_new), _style(style), C.byref(out)))
        return out

    def trial_atom(self, i: int, r_new) -> TrialResult:
        r_new = np.ascontiguousarray(r_new, dtype=np.float64)
        out = TrialResult()
        self._ck(self.lib.mmc_trial_atom(self.h, i, _dp(r_new), C.byref(out)))
        return out

    def accept(self):
        self._ck(self.lib.mmc_accept(self.h))

    def reject(self):
        self._ck(self.lib.mmc_reject(self.h))

    # ---- volume move (Ewald/volumeChange.jl:50-147)
    def volume_trial(self, box_new: float, kappa_new: float, style="ewald") -> Properties:
        out = Properties()
        self._ck(self.lib.mmc_volume_trial(self.h, box_new, kappa_new, _style(style), C.byref(out)))
        return out

    def volume_accept(self):
        self._ck(self.lib.mmc_volume_accept(self.h))

    def volume_reject(self):
        self._ck(self.lib.mmc_volume_reject(self.h))

    # ---- the Loop stand-in (Ewald/main.jl:487-651) in the library's C++ host driver
    def loop_run(self, params: LoopParams, com, quat, db, uniforms, n_moves, e0=0.0, v0=0.0, device=False):
        """Loop() of Ewald/main.jl:487-651 over the C ABI: one fused launch per move (device=False, the
        ccall protocol) or the whole block of moves in one launch with the state on chip (device=True)."""
        fn = self.lib.mmc_loop_run_device if device else self.lib.mmc_loop_run
        assert com.dtype == np.float64 and com.flags.c_contiguous
        assert quat.dtype == np.float64 and quat.flags.c_contiguous
        db = np.ascontiguousarray(db, dtype=np.float64)
        uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
        acc = np.zeros(n_moves, dtype=np.uint8)
        delta = np.zeros(n_moves)
        st = LoopStats()
        rc = self._ck(fn(self.h, C.byref(params), _dp(com), _dp(quat), _dp(db), _dp(uniforms),
                                            uniforms.shape[0], n_moves, e0, v0,
                                            acc.ctypes.data_as(_lib.c_uint8_p), _dp(delta), C.byref(st)))
        return rc, acc, delta, st

    def loop_run_atoms(self, temperature, dr_max, r, uniforms, n_moves, e0=0.0, v0=0.0, device=False):
        fn = self.lib.mmc_loop_run_atoms_device if device else self.lib.mmc_loop_run_atoms
        assert r.dtype == np.float64 and r.flags.c_contiguous
        uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
        acc = np.zeros(n_moves, dtype=np.uint8)
        delta = np.zeros(n_moves)
        st = LoopStats()
        rc = self._ck(fn(self.h, temperature, dr_max, _dp(r), _dp(uniforms),
                                                  uniforms.shape[0], n_moves, e0, v0,
                                                  acc.ctypes.data_as(_lib.c_uint8_p), _dp(delta), C.byref(st)))
        return rc, acc, delta, st

    def set_intramolecular(self, on: bool):
        """Opt-in intramolecular Ewald correction (mmc_set_intramolecular; the reference omits it, energy.jl:1008-1021)."""
        self._ck(self.lib.mmc_set_intramolecular(self.h, int(on)))

    # ---- instrumentation
    def counters(self) -> Counters:
        c = Counters()
        self._ck(self.lib.mmc_get_counters(self.h, C.byref(c)))
        return c

    def set_timing(self, on: bool):
        self._ck(self.lib.mmc_set_timing(self.h, int(on)))

    def last_timings(self):
        ms = (C.c_float * 4)()
        self._ck(self.lib.mmc_last_timings(self.h, ms))
        return {"pairs_ms": ms[0], "rhok_ms": ms[1], "bin_gather_ms": ms[2], "total_ms": ms[3]}

    def last_eval_info(self):
        n, m, c, k = C.c_int64(), C.c_int32(), C.c_int32(), C.c_int32()
        self._ck(self.lib.mmc_last_eval_info(self.h, C.byref(n), C.byref(m), C.byref(c), C.byref(k)))
        kern = {7: "k_pairs_v7", 64: "k_pairs_fast<64>", 128: "k_pairs_fast<128>", 0: "k_pairs"}.get(k.value, str(k.value))
        return {"pairs_in_cutoff": n.value, "mode": ("cells", "tiles", "rows")[m.value] if m.value >= 0 else None,
                "cells_per_dim": c.value, "pair_kernel": kern}

    def debug_set(self, key: str, value: int):
        self._ck(self.lib.mmc_debug_set(self.h, key.encode(), C.c_int64(value)))

    def measure_fp64_peak(self) -> float:
        t = C.c_double()
        self._ck(self.lib.mmc_measure_fp64_peak(self.h, C.byref(t)))
        return t.value


def host_register(a: np.ndarray):
    """Page-lock a caller-owned array (mmc_host_register) so uploads from it run at full PCIe rate."""
    rc = _lib.load().mmc_host_register(C.c_void_p(a.ctypes.data), C.c_size_t(a.nbytes))
    if rc < 0:
        raise _lib.MMCError(rc, "mmc_host_register failed")


def host_unregister(a: np.ndarray):
    rc = _lib.load().mmc_host_unregister(C.c_void_p(a.ctypes.data))
    if rc < 0:
        raise _lib.MMCError(rc, "mmc_host_unregister failed")


def julia_rand(seed: int, n: int, skip: int = 0) -> np.ndarray:
    """`Random.seed!(seed); [rand() for _ in 1:n]` of the reference's Julia (MersenneTwister /
    dSFMT-19937) via mmc_julia_rand — the uniform stream Loop() consumes (Ewald/main.jl:36,516)."""
    out = np.empty(n, dtype=np.float64)
    rc = _lib.load().mmc_julia_rand(C.c_uint64(seed), C.c_int64(skip), out.ctypes.data_as(_lib.c_double_p), C.c_int64(n))
    if rc:
        raise _lib.MMCError(rc, "mmc_julia_rand: bad arguments")
    return out


def water_engine(ms: MolecularSystem, r_cut: float = 10.0, alpha: float = ALPHA, nk: int = NK,
                 k_sq_max: int = K_SQ_MAX, **kw) -> Engine:
    """Engine set up like Ewald/main.jl:285-303: kappa = alpha/box, nk = 5, k² < 27."""
    eng = Engine(**kw)
    eng.upload_system(ms, r_cut, r_cut)
    eng.PrepareEwaldVariables(alpha / ms.box, nk, k_sq_max, FACTOR)
    return eng
