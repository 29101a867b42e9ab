"""ctypes binding of libmmc_b200.so — the same C ABI Julia binds with ccall (include/mmc_b200.h).

Loading fails loudly when the in-tree library is missing: there is no Python/CPU fallback for
any energy routine.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libmmc_b200.so"

MMC_OK, MMC_EINVAL, MMC_ECUDA, MMC_ENCCL, MMC_ESTATE = 0, -1, -2, -3, -4
STYLE_EWALD, STYLE_WOLF, STYLE_LJ_ONLY, STYLE_LJ_ATOMS = 0, 1, 2, 3

c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)
c_int32_p = C.POINTER(C.c_int32)
c_uint8_p = C.POINTER(C.c_uint8)
c_float_p = C.POINTER(C.c_float)


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32),
                ("sync_mode", C.c_int32), ("stream", C.c_void_p)]


class Properties(C.Structure):
    _fields_ = [("energy", C.c_double), ("virial", C.c_double), ("coulomb", C.c_double),
                ("lj", C.c_double), ("real", C.c_double), ("recip", C.c_double),
                ("self_", C.c_double), ("wolf_const", C.c_double), ("overlaps", C.c_int64), ("intra", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class TrialResult(C.Structure):
    _fields_ = [("lj_old", C.c_double), ("lj_vir_old", C.c_double), ("lj_new", C.c_double),
                ("lj_vir_new", C.c_double), ("qq_old", C.c_double), ("qq_vir_old", C.c_double),
                ("qq_new", C.c_double), ("qq_vir_new", C.c_double), ("d_recip", C.c_double),
                ("overlap_old", C.c_int32), ("overlap_new", C.c_int32)]


class LoopParams(C.Structure):
    _fields_ = [("temperature", C.c_double), ("dr_max", C.c_double), ("dphi_max", C.c_double),
                ("p_trans", C.c_double), ("p_rot", C.c_double), ("style", C.c_int32), ("adjust", C.c_int32)]


class LoopStats(C.Structure):
    _fields_ = [("n_moves", C.c_int64), ("n_accepted", C.c_int64), ("n_overlap", C.c_int64),
                ("uniforms_used", C.c_int64), ("trans_attempt", C.c_int64), ("trans_accept", C.c_int64),
                ("rot_attempt", C.c_int64), ("rot_accept", C.c_int64), ("dr_max", C.c_double),
                ("dphi_max", C.c_double), ("total_energy", C.c_double), ("total_virial", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Counters(C.Structure):
    _fields_ = [("kernel_launches", C.c_int64), ("trial_moves", C.c_int64), ("commits", C.c_int64),
                ("overlap_events", C.c_int64), ("full_energy_evals", C.c_int64)]


H = C.c_void_p
# name -> (restype, argtypes); every symbol include/mmc_b200.h declares
SIGNATURES = {
    "mmc_create": (C.c_int, [C.POINTER(Config), C.POINTER(H)]),
    "mmc_destroy": (C.c_int, [H]),
    "mmc_last_error": (C.c_char_p, [H]),
    "mmc_version": (C.c_int, []),
    "mmc_upload_system": (C.c_int, [H, C.c_int64, C.c_int64, c_double_p, c_double_p, c_int64_p, c_int64_p,
                                    c_int64_p, c_double_p, C.c_int32, c_double_p, c_double_p,
                                    C.c_double, C.c_double, C.c_double]),
    "mmc_upload_atoms": (C.c_int, [H, C.c_int64, c_double_p, c_double_p, c_double_p, C.c_double, C.c_double]),
    "mmc_download_system": (C.c_int, [H, c_double_p, c_double_p]),
    "mmc_download_atoms": (C.c_int, [H, c_double_p]),
    "mmc_ewald_prepare": (C.c_int, [H, C.c_double, C.c_int32, C.c_int32, C.c_double, c_int32_p]),
    "mmc_get_kvectors": (C.c_int, [H, c_int32_p, c_double_p]),
    "mmc_get_rhok": (C.c_int, [H, c_double_p, c_double_p]),
    "mmc_lj_mol": (C.c_int, [H, C.c_int64, c_double_p, c_double_p]),
    "mmc_ewald_real": (C.c_int, [H, C.c_int64, c_double_p, c_int32_p]),
    "mmc_ewald_short": (C.c_int, [H, C.c_int64, c_double_p, c_double_p, c_int32_p]),
    "mmc_set_molecule": (C.c_int, [H, C.c_int64, c_double_p, c_double_p]),
    "mmc_recip_long": (C.c_int, [H, c_double_p]),
    "mmc_recip_move": (C.c_int, [H, c_double_p, c_double_p, c_double_p, C.c_int32, c_double_p]),
    "mmc_recip_commit": (C.c_int, [H]),
    "mmc_recip_rollback": (C.c_int, [H]),
    "mmc_ewald_self": (C.c_int, [H, c_double_p]),
    "mmc_lj_atom": (C.c_int, [H, C.c_int64, c_double_p, c_double_p]),
    "mmc_set_atom": (C.c_int, [H, C.c_int64, c_double_p]),
    "mmc_potential": (C.c_int, [H, C.c_int32, C.POINTER(Properties)]),
    "mmc_partial_count": (C.c_int, [H, c_int64_p]),
    "mmc_potential_partial": (C.c_int, [H, C.c_int32, C.c_void_p]),
    "mmc_potential_finalize": (C.c_int, [H, C.c_int32, C.c_void_p, C.POINTER(Properties)]),
    "mmc_trial_move": (C.c_int, [H, C.c_int64, c_double_p, c_double_p, C.c_int32, C.POINTER(TrialResult)]),
    "mmc_accept": (C.c_int, [H]),
    "mmc_reject": (C.c_int, [H]),
    "mmc_trial_atom": (C.c_int, [H, C.c_int64, c_double_p, C.POINTER(TrialResult)]),
    "mmc_volume_trial": (C.c_int, [H, C.c_double, C.c_double, C.c_int32, C.POINTER(Properties)]),
    "mmc_volume_accept": (C.c_int, [H]),
    "mmc_volume_reject": (C.c_int, [H]),
    "mmc_peer_export": (C.c_int, [H, C.c_void_p]),
    "mmc_peer_import": (C.c_int, [H, C.c_int32, C.c_void_p]),
    "mmc_peer_import_ptr": (C.c_int, [H, C.c_int32, C.c_void_p]),
    "mmc_peer_buffer": (C.c_int, [H, C.POINTER(C.c_void_p)]),
    "mmc_potential_sharded_begin": (C.c_int, [H, C.c_int32]),
    "mmc_potential_sharded_end": (C.c_int, [H, C.POINTER(Properties)]),
    "mmc_potential_sharded": (C.c_int, [H, C.c_int32, C.POINTER(Properties)]),
    "mmc_upload_positions": (C.c_int, [H, c_double_p, c_double_p]),
    "mmc_energy_all": (C.c_int, [H, C.c_int32, c_double_p, c_double_p, c_double_p, c_int32_p]),
    "mmc_potential_host": (C.c_int, [H, c_double_p, c_double_p, C.c_int32, C.POINTER(Properties)]),
    "mmc_last_host_bytes": (C.c_int, [H, c_int64_p]),
    "mmc_loop_run": (C.c_int, [H, C.POINTER(LoopParams), c_double_p, c_double_p, c_double_p, c_double_p,
                               C.c_int64, C.c_int64, C.c_double, C.c_double, c_uint8_p, c_double_p,
                               C.POINTER(LoopStats)]),
    "mmc_loop_run_device": (C.c_int, [H, C.POINTER(LoopParams), c_double_p, c_double_p, c_double_p, c_double_p,
                                      C.c_int64, C.c_int64, C.c_double, C.c_double, c_uint8_p, c_double_p,
                                      C.POINTER(LoopStats)]),
    "mmc_loop_run_atoms": (C.c_int, [H, C.c_double, C.c_double, c_double_p, c_double_p, C.c_int64, C.c_int64,
                                     C.c_double, C.c_double, c_uint8_p, c_double_p, C.POINTER(LoopStats)]),
    "mmc_loop_run_atoms_device": (C.c_int, [H, C.c_double, C.c_double, c_double_p, c_double_p, C.c_int64, C.c_int64,
                                            C.c_double, C.c_double, c_uint8_p, c_double_p, C.POINTER(LoopStats)]),
    "mmc_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "mmc_host_unregister": (C.c_int, [C.c_void_p]),
    "mmc_julia_rand": (C.c_int, [C.c_uint64, C.c_int64, c_double_p, C.c_int64]),
    "mmc_set_intramolecular": (C.c_int, [H, C.c_int32]),
    "mmc_get_counters": (C.c_int, [H, C.POINTER(Counters)]),
    "mmc_set_timing": (C.c_int, [H, C.c_int32]),
    "mmc_last_timings": (C.c_int, [H, c_float_p]),
    "mmc_last_eval_info": (C.c_int, [H, c_int64_p, c_int32_p, c_int32_p, c_int32_p]),
    "mmc_debug_set": (C.c_int, [H, C.c_char_p, C.c_int64]),
    "mmc_measure_fp64_peak": (C.c_int, [H, c_double_p]),
}

_lib = None


class MMCError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libmmc_b200 error {code}: {msg}")
        self.code = code


def load() -> C.CDLL:
    """dlopen the in-tree libmmc_b200.so (built by __graft_entry__.build() / csrc/build.sh)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or metropolismontecarlo_b200/csrc/build.sh. There is no CPU fallback.")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)      # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
