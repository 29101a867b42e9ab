"""Shared helpers for the parity tests: one system description → oracle structs and engine uploads."""
import numpy as np

from metropolismontecarlo_b200 import systems
from oracle import oracle as ora


def ora_system(ms: systems.MolecularSystem) -> ora.System:
    return ora.System(ms.coords, ms.charge, ms.atype, ms.first_atom, ms.last_atom, ms.com, ms.eps, ms.sig)


def ora_ewald(box, alpha=systems.ALPHA, nk=systems.NK, k_sq_max=systems.K_SQ_MAX) -> ora.Ewald:
    return ora.Ewald(alpha / box, nk, k_sq_max, systems.FACTOR, box)


def rel(a, b):
    return abs(a - b) / max(abs(a), abs(b), 1e-300)
