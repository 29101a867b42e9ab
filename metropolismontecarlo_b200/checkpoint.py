"""Configuration output / restart files of the reference driver and its closed-form tail corrections.

Host-side text I/O and scalar formulas only — nothing here touches the energy path.

Writers follow the reference's `@sprintf` layouts character for character:
  * `print_pdb`     — `PrintPDB(soa, moa, boxSize, step, filename)`   Ewald/initialConfigurations.jl:160-181
  * `print_output`  — `PrintOutput(system, totProps, …, "xyz_quat")`  Ewald/initialConfigurations.jl:183-237
  * `read_cnf`      — `ReadCNF("cnf_input.inp")`                      Ewald/initialConfigurations.jl:239-280
The reference writes three decimals (`%7.3f`), so its own restart is lossy; `write_cnf(..., exact=True)` (an extension)
keeps every bit (`%.17g`) and `read_cnf` reads both.

Tail corrections (LJ units for the monatomic forms, table units for the polyatomic ones):
  * `potential_lrc`, `pressure_lrc`, `pressure_delta` — Ewald/auxillary.jl:16-35 (same in Poly/auxillary.jl:7-26)
  * `ener_corr`, `press_corr`                          — Ewald/energy.jl:514-614
"""
from __future__ import annotations

import math

import numpy as np

from .systems import MolecularSystem


# ---------------------------------------------------------------------------------------------- writers / readers
def print_pdb(ms: MolecularSystem, box, step: int = 1, filename: str = "pdbOutput", atom_names=None,
              mol_name: str = "SOL") -> str:
    """Ewald/initialConfigurations.jl:160-181.  `atom_names[t-1]` is the name of atom type t (1-based like soa.atype); returns the path."""
    box = np.broadcast_to(np.asarray(box, dtype=np.float64), (3,))
    if atom_names is None:
        atom_names = [f"T{t + 1}" for t in range(int(ms.atype.max()))]
    path = f"{filename}_{step}.pdb"
    mol_of_site = np.repeat(np.arange(1, ms.n_mol + 1), ms.last_atom - ms.first_atom + 1)
    with open(path, "w") as f:
        f.write("%-7s %7.3f %7.3f %7.3f %30s \n" % ("CRYST1", box[0], box[1], box[2], "90.00  90.00  90.00 P 1           1"))
        for i in range(ms.n_sites):
            x, y, z = ms.coords[i]
            f.write("%-6s %4d %3s %4s %5d %3s %7.3f %7.3f %7.3f %5.2f %5.2f \n" % (
                "ATOM", i + 1, atom_names[int(ms.atype[i]) - 1], mol_name, mol_of_site[i], " ", x, y, z, 1.00, 0.00))
    return path


def print_output(ms: MolecularSystem, quat: np.ndarray, box: float, step: int = 1, filename: str = "xyz_quat",
                 atom_names=None, atom_types=None, mol_name: str = "SOL") -> str:
    """Ewald/initialConfigurations.jl:183-237: molecule block (COM + quaternion) then atom block (charge + position)."""
    if atom_names is None:
        atom_names = [f"T{int(t)}" for t in ms.atype]
    if atom_types is None:
        atom_types = atom_names
    path = f"{filename}_{step}.pdb"
    with open(path, "w") as f:
        f.write("%-7s %7.3f %7.3f %7.3f\n" % ("Output", box, box, box))
        f.write(" Molecular coordinates and quaternions\n")
        f.write(" #, mol name, atom Start, atom End, x, y, z, q0, q1, q2, q3\n")
        for i in range(ms.n_mol):
            f.write("%4d %-7s %4d %4d %7.3f %7.3f %7.3f %7.3f %7.3f %7.3f %7.3f\n" % (
                i + 1, mol_name, ms.first_atom[i], ms.last_atom[i], *ms.com[i], *quat[i]))
        f.write("Atom coordinates\n")
        f.write("#, name, type, charge, x, y, z\n")
        for i in range(ms.n_sites):
            f.write("%4d %-7s %-7s %7.3f %7.3f %7.3f %7.3f \n" % (
                i + 1, atom_names[i], atom_types[i], ms.charge[i], *ms.coords[i]))
    return path


def read_output(path: str):
    """Reads a `print_output` file back: (box, com[n_mol,3], quat[n_mol,4], first_atom, last_atom, charge[n_s], coords[n_s,3])."""
    with open(path) as f:
        lines = f.read().splitlines()
    box = float(lines[0].split()[1])
    k = 3
    com, quat, fa, la = [], [], [], []
    while not lines[k].startswith("Atom coordinates"):
        t = lines[k].split()
        fa.append(int(t[2])); la.append(int(t[3]))
        com.append([float(v) for v in t[4:7]]); quat.append([float(v) for v in t[7:11]])
        k += 1
    k += 2
    q, r = [], []
    for ln in lines[k:]:
        t = ln.split()
        if len(t) < 7:
            continue
        q.append(float(t[3])); r.append([float(v) for v in t[4:7]])
    return (box, np.array(com), np.array(quat), np.array(fa, dtype=np.int64), np.array(la, dtype=np.int64),
            np.array(q), np.array(r))


def write_cnf(path: str, com: np.ndarray, quat: np.ndarray, box: float, exact: bool = False) -> str:
    """The layout `ReadCNF` expects (Ewald/initialConfigurations.jl:239-280): line 1 = molecule count, line 2 = box,
    then `x y z q0 q1 q2 q3` per molecule."""
    fmt = "%.17g" if exact else "%15.8f"
    with open(path, "w") as f:
        f.write(f"{len(com)}\n")
        f.write((fmt % box) + "\n")
        for c, q in zip(com, quat):
            f.write(" ".join(fmt % v for v in (*c, *q)) + "\n")
    return path


def read_cnf(path: str = "cnf_input.inp"):
    """Ewald/initialConfigurations.jl:239-280 — returns (rm[n,3], quat[n,4], box) like the reference's `r, e, box1`."""
    r, e, box1 = [], [], 0.0
    with open(path) as f:
        for i, line in enumerate(f, start=1):
            if i == 2:
                box1 = float(line.strip())
            if i >= 3:
                lin = line.split()
                if not lin:
                    continue
                r.append([float(v) for v in lin[0:3]])
                e.append([float(v) for v in lin[3:7]])
    return np.array(r), np.array(e), box1


# ---------------------------------------------------------------------------------------------- tail corrections
def potential_lrc(rho: float, r_cut: float) -> float:
    """LJ long-range energy correction per atom, reduced units (Ewald/auxillary.jl:16-21)."""
    sr3 = 1.0 / r_cut ** 3
    return math.pi * ((8.0 / 9.0) * sr3 ** 3 - (8.0 / 3.0) * sr3) * rho


def pressure_lrc(rho: float, r_cut: float) -> float:
    """LJ long-range pressure correction, reduced units (Ewald/auxillary.jl:23-28)."""
    sr3 = 1.0 / r_cut ** 3
    return math.pi * ((32.0 / 9.0) * sr3 ** 3 - (16.0 / 3.0) * sr3) * rho ** 2


def pressure_delta(rho: float, r_cut: float) -> float:
    """Pressure correction for the discontinuity of the cut potential at r_cut (Ewald/auxillary.jl:30-35)."""
    sr3 = 1.0 / r_cut ** 3
    return math.pi * (8.0 / 3.0) * (sr3 ** 3 - sr3) * rho ** 2


def ener_corr(eps: np.ndarray, sig: np.ndarray, counts, r_cut: float, box: float) -> float:
    """Polyatomic LJ tail correction to the energy (Ewald/energy.jl:565-603): eps/sig are the mixed nt×nt tables,
    counts[t] the number of atoms of type t (the reference's `b`)."""
    eps = np.asarray(eps, dtype=np.float64); sig = np.asarray(sig, dtype=np.float64)
    vol = box ** 3
    coru = 0.0
    for i in range(len(counts)):
        for j in range(len(counts)):
            sig3 = sig[i, j] * sig[i, j] * sig[i, j]
            sigor3 = sig3 / (r_cut * r_cut * r_cut)
            sigor9 = sigor3 * sigor3 * sigor3
            coru += counts[i] * counts[j] * eps[i, j] * sig3 * ((1.0 / 3.0) * sigor9 - sigor3)
    return 8.0 * math.pi / (3.0 * vol) * coru


def press_corr(eps: np.ndarray, sig: np.ndarray, counts, r_cut: float, box: float) -> float:
    """Polyatomic LJ tail correction to the pressure (Ewald/energy.jl:514-562)."""
    eps = np.asarray(eps, dtype=np.float64); sig = np.asarray(sig, dtype=np.float64)
    vol = box ** 3
    corp = 0.0
    for i in range(len(counts)):
        for j in range(len(counts)):
            sig3 = sig[i, j] * sig[i, j] * sig[i, j]
            sigor3 = sig3 / (r_cut * r_cut * r_cut)
            sigor9 = sigor3 * sigor3 * sigor3
            corp += counts[i] * counts[j] * eps[i, j] * sig3 * ((2.0 / 3.0) * sigor9 - sigor3)
    return 16.0 * math.pi / (3.0 * vol * vol) * corp
