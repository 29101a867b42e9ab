"""Host-side system construction in the reference's own array layouts.

This is the input side of the drop-in boundary: the
arrays built here are exactly what Julia's ``soa``/``moa``/``vdwTable`` hold and what
``mmc_upload_system`` takes.  Nothing in this file touches the GPU.

Reference behaviour mirrored (file:line into /root/reference):
  * NIST SPC/E config → soa/moa : Ewald/initialConfigurations.jl:282-355 (ReadNIST, COM with
    masses [15.99, 1.009, 1.009]) and the shift into [0, L) at Ewald/main.jl:247-275.
  * lattice start              : Ewald/initialConfigurations.jl:10-53 (InitCubicGrid).
  * random orientations         : Ewald/quaternions.jl:122-156 (random_quaternion),
    sites = COM + MATMUL(q_to_a(q), db) (Ewald/main.jl:545-548, quaternions.jl:11-50).
  * constants                   : Ewald/constants.jl:24-28 (factor), Ewald/main.jl:242-245 (SPC/E LJ).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

# Ewald/constants.jl:24-28
_kb1 = 1.3806488e-23
_eps01 = 8.854187817e-12
_eps01 *= 1e-10
_e1 = 1.602176565e-19
FACTOR = _e1 ** 2 / _eps01 / 4 / np.pi / _kb1  # 167100.9566... K·Å/e²

# Ewald/main.jl:242-245 and initialConfigurations.jl:316,329
SPCE_SIGMA_O = 0.316555789 * 10.0
SPCE_EPS_O = 78.1974311
SPCE_Q_O = -2 * 0.42380
SPCE_Q_H = 0.42380
SPCE_MASS = np.array([15.99, 1.009, 1.009])
ALPHA = 5.6       # Ewald/main.jl:285-289 — kappa = alpha / box
NK = 5            # Ewald/main.jl:293
K_SQ_MAX = 27     # Ewald/main.jl:294


@dataclass
class MolecularSystem:
    """soa + moa + vdwTable (+ body frames / quaternions for the driver)."""
    coords: np.ndarray        # soa.coords  (n_sites, 3) float64
    charge: np.ndarray        # soa.charge  (n_sites,)
    atype: np.ndarray         # soa.atype   (n_sites,) int64, 1-based
    first_atom: np.ndarray    # moa.firstAtom (n_mol,) int64, 1-based
    last_atom: np.ndarray     # moa.lastAtom  (n_mol,) int64, inclusive
    com: np.ndarray           # moa.COM     (n_mol, 3)
    eps: np.ndarray           # vdwTable.ϵᵢⱼ (nt, nt) [K]
    sig: np.ndarray           # vdwTable.σᵢⱼ (nt, nt) [Å]
    box: float
    db: np.ndarray = field(default=None)     # body-fixed site vectors (n_sites, 3)
    quat: np.ndarray = field(default=None)   # moa.quat (n_mol, 4)

    @property
    def n_mol(self):
        return self.com.shape[0]

    @property
    def n_sites(self):
        return self.coords.shape[0]

    def copy(self):
        return MolecularSystem(*(None if v is None else (v.copy() if isinstance(v, np.ndarray) else v)
                                 for v in (self.coords, self.charge, self.atype, self.first_atom,
                                           self.last_atom, self.com, self.eps, self.sig, self.box,
                                           self.db, self.quat)))


def q_to_a(q):
    """Rotation matrix rows exactly as written at Ewald/quaternions.jl:37-50
    (element [2,3] is 2(q2 q4 + q1 q2) in the reference)."""
    q1, q2, q3, q4 = q
    return np.array([
        [q1 * q1 + q2 * q2 - q3 * q3 - q4 * q4, 2 * (q2 * q3 + q1 * q4), 2 * (q2 * q4 - q1 * q3)],
        [2 * (q2 * q3 - q1 * q4), q1 * q1 - q2 * q2 + q3 * q3 - q4 * q4, 2 * (q2 * q4 + q1 * q2)],
        [2 * (q2 * q4 + q1 * q3), 2 * (q3 * q4 - q1 * q2), q1 * q1 - q2 * q2 - q3 * q3 + q4 * q4],
    ])


def matmul_ref(ai, db):
    """Ewald/auxillary.jl:154-159: (db·ai[:,1], db·ai[:,2], db·ai[:,3])."""
    return np.array([db[0] * ai[0, c] + db[1] * ai[1, c] + db[2] * ai[2, c] for c in range(3)])


def _water_tables():
    eps = np.array([[SPCE_EPS_O, 0.0], [0.0, 0.0]])
    sig = np.array([[SPCE_SIGMA_O, SPCE_SIGMA_O / 2], [SPCE_SIGMA_O / 2, 0.0]])
    return eps, sig


def _water_topology(n_mol):
    charge = np.tile(np.array([SPCE_Q_O, SPCE_Q_H, SPCE_Q_H]), n_mol)
    atype = np.tile(np.array([1, 2, 2], dtype=np.int64), n_mol)
    first = np.arange(n_mol, dtype=np.int64) * 3 + 1
    last = first + 2
    return charge, atype, first, last


def spce_from_nist(xyz: np.ndarray, box: float) -> MolecularSystem:
    """NIST SPC/E configuration (O,H,H rows) → reference layouts, shifted like main.jl:247-275."""
    xyz = np.asarray(xyz, dtype=np.float64).reshape(-1, 3)
    n_mol = xyz.shape[0] // 3
    m = SPCE_MASS
    total = (m[0] + m[1]) + m[2]
    a = xyz.reshape(n_mol, 3, 3)
    com = ((a[:, 0, :] * m[0] + a[:, 1, :] * m[1]) + a[:, 2, :] * m[2]) / total
    shift = np.abs(com.min(axis=0))
    com = com + shift
    coords = xyz + shift
    charge, atype, first, last = _water_topology(n_mol)
    eps, sig = _water_tables()
    db = coords - np.repeat(com, 3, axis=0)   # body frame for quaternion (1,0,0,0): A = I
    quat = np.tile(np.array([1.0, 0.0, 0.0, 0.0]), (n_mol, 1))
    return MolecularSystem(coords, charge, atype, first, last, com, eps, sig, float(box), db, quat)


def load_nist(config: int) -> MolecularSystem:
    """config 1..4 from the committed fixture tests/golden/nist_spce.npz (4 == coord750.txt)."""
    p = Path(__file__).resolve().parent.parent / "tests" / "golden" / "nist_spce.npz"
    z = np.load(p)
    return spce_from_nist(z[f"xyz{config}"], float(z[f"box{config}"]))


def init_cubic_grid(n: int, rho: float) -> tuple[np.ndarray, float]:
    """Ewald/initialConfigurations.jl:10-53."""
    L = (n / rho) ** (1.0 / 3.0)
    ncube = 2
    while ncube ** 3 < n:
        ncube += 1
    idx = np.arange(n)
    posit = np.stack([idx % ncube, (idx // ncube) % ncube, idx // (ncube * ncube)], axis=1)
    return (posit + 0.01) * (L / ncube), L


def random_quaternions(n: int, rng: np.random.Generator) -> np.ndarray:
    """Ewald/quaternions.jl:122-156 (Marsaglia), vectorised with rejection."""
    def disk(k):
        out = np.empty((k, 2))
        nrm = np.empty(k)
        todo = np.arange(k)
        while todo.size:
            z = 2.0 * rng.random((todo.size, 2)) - 1.0
            s = (z * z).sum(axis=1)
            ok = s < 1.0
            out[todo[ok]] = z[ok]
            nrm[todo[ok]] = s[ok]
            todo = todo[~ok]
        return out, nrm
    z1, n1 = disk(n)
    z2, n2 = disk(n)
    f = np.sqrt((1.0 - n1) / n2)
    return np.column_stack([z1[:, 0], z1[:, 1], z2[:, 0] * f, z2[:, 1] * f])


def spce_body_frame() -> np.ndarray:
    """SPC/E geometry (r_OH = 1 Å, HOH = 109.47°) with the COM at the origin."""
    half = np.deg2rad(109.47) / 2.0
    r = np.array([[0.0, 0.0, 0.0],
                  [np.sin(half), 0.0, np.cos(half)],
                  [-np.sin(half), 0.0, np.cos(half)]])
    com = (r * SPCE_MASS[:, None]).sum(axis=0) / SPCE_MASS.sum()
    return r - com


def spce_lattice(n_mol: int, rho: float = 0.033101144, seed: int = 11234) -> MolecularSystem:
    """Synthetic SPC/E box (configs D/E of SURVEY §8d): COMs on InitCubicGrid at density rho
    (Ewald/main.jl:63), random unit quaternions, sites = COM + MATMUL(A(q), db)."""
    rng = np.random.default_rng(seed)
    com, L = init_cubic_grid(n_mol, rho)
    quat = random_quaternions(n_mol, rng)
    dbm = spce_body_frame()
    q1, q2, q3, q4 = quat.T
    A = np.empty((n_mol, 3, 3))
    A[:, 0, 0] = q1 * q1 + q2 * q2 - q3 * q3 - q4 * q4
    A[:, 0, 1] = 2 * (q2 * q3 + q1 * q4)
    A[:, 0, 2] = 2 * (q2 * q4 - q1 * q3)
    A[:, 1, 0] = 2 * (q2 * q3 - q1 * q4)
    A[:, 1, 1] = q1 * q1 - q2 * q2 + q3 * q3 - q4 * q4
    A[:, 1, 2] = 2 * (q2 * q4 + q1 * q2)          # as written in the reference
    A[:, 2, 0] = 2 * (q2 * q4 + q1 * q3)
    A[:, 2, 1] = 2 * (q3 * q4 - q1 * q2)
    A[:, 2, 2] = q1 * q1 - q2 * q2 - q3 * q3 + q4 * q4
    # MATMUL(ai, db)[c] = sum_r db[r] * ai[r, c]
    d = np.einsum("ar,mrc->mac", dbm, A)
    coords = (com[:, None, :] + d).reshape(-1, 3)
    charge, atype, first, last = _water_topology(n_mol)
    eps, sig = _water_tables()
    db = np.tile(dbm, (n_mol, 1))
    return MolecularSystem(coords, charge, atype, first, last, com, eps, sig, float(L), db, quat)


# ------------------------------------------------------------------ GROMACS .top / .pdb subset (input formats)
R_GAS = 8.3144621e-3      # kJ mol^-1 K^-1, Ewald/constants.jl:11 (ε[kJ/mol] / R → K, main.jl:185)


def read_top(path) -> dict:
    """The subset of a GROMACS topology the reference's ReadTopFile uses (Ewald/setup.jl:30-390) for rigid
    molecules: [ atomtypes ] (name, mass, charge, σ [nm], ε [kJ/mol]) and, per [ moleculetype ], its [ atoms ]
    (type, charge, mass) and the [ molecules ] counts.  Comments (;) and preprocessor lines (#) are skipped."""
    section, types, mols, counts, cur = None, {}, {}, [], None
    for raw in Path(path).read_text().splitlines():
        line = raw.split(";", 1)[0].strip()
        if not line or line.startswith("#"):
            continue
        if line.startswith("["):
            section = line.strip("[] \t").lower()
            continue
        t = line.split()
        if section == "atomtypes":            # name bond_type mass charge ptype sigma epsilon   (7 columns, as water.top)
            types[t[0]] = {"mass": float(t[-5]), "charge": float(t[-4]), "sigma_nm": float(t[-2]), "eps_kj": float(t[-1])}
        elif section == "moleculetype":
            cur = t[0]
            mols[cur] = []
        elif section == "atoms" and cur is not None:   # nr type resnr residue atom cgnr charge mass
            mols[cur].append({"type": t[1], "name": t[4], "charge": float(t[6]), "mass": float(t[7])})
        elif section == "molecules":
            counts.append((t[0], int(t[1])))
    return {"atomtypes": types, "molecules": mols, "counts": counts}


def read_pdb(path) -> np.ndarray:
    """ATOM/HETATM coordinates [Å] of a PDB file (columns 31-54), as the reference's ReadPDB (Ewald/setup.jl) uses them."""
    xyz = []
    for line in Path(path).read_text().splitlines():
        if line.startswith(("ATOM", "HETATM")):
            xyz.append([float(line[30:38]), float(line[38:46]), float(line[46:54])])
    return np.array(xyz, dtype=np.float64)


def tables_from_types(eps_kj, sig_nm):
    """vdwTable of Ewald/main.jl:183-186: geometric ε, arithmetic σ (structs.jl:342-346), then ε/R → K, σ·10 → Å."""
    e = np.asarray(eps_kj, dtype=np.float64)
    g = np.asarray(sig_nm, dtype=np.float64)
    eps = np.sqrt(e[:, None] * e[None, :]) / R_GAS
    sig = (g[:, None] + g[None, :]) / 2 * 10.0
    return eps, sig


def rigid_lattice(model: dict, n_mol: int, rho: float = 0.033101144, seed: int = 11234) -> MolecularSystem:
    """A box of one rigid molecule type the way the reference's "crystal" start builds it (Ewald/main.jl:156-189):
    body frame = PDB coordinates shifted to the COM (main.jl:163-164), COMs on InitCubicGrid, random quaternions,
    sites = COM + MATMUL(A(q), db).  `model`: {"types": [names], "eps_kj", "sig_nm" per type, "atoms": [(type
    index, charge, mass)], "xyz": body coordinates in Å}."""
    rng = np.random.default_rng(seed)
    atoms = model["atoms"]
    S = len(atoms)
    mass = np.array([a[2] for a in atoms])
    xyz = np.asarray(model["xyz"], dtype=np.float64)
    dbm = xyz - (xyz * mass[:, None]).sum(axis=0) / mass.sum()
    com, L = init_cubic_grid(n_mol, rho)
    quat = random_quaternions(n_mol, rng)
    coords = np.empty((n_mol * S, 3))
    for m in range(n_mol):
        coords[S * m:S * m + S] = com[m] + matmul_ref(q_to_a(quat[m]), dbm)
    charge = np.tile(np.array([a[1] for a in atoms]), n_mol)
    atype = np.tile(np.array([a[0] + 1 for a in atoms], dtype=np.int64), n_mol)
    first = np.arange(n_mol, dtype=np.int64) * S + 1
    eps, sig = tables_from_types(model["eps_kj"], model["sig_nm"])
    return MolecularSystem(coords, charge, atype, first, first + S - 1, com, eps, sig, float(L), np.tile(dbm, (n_mol, 1)), quat)


def model_from_files(top_path, pdb_path, molname=None) -> dict:
    """{types, eps_kj, sig_nm, atoms, xyz} of one molecule type from a .top and its .pdb (what main.jl:156-171 reads)."""
    top = read_top(top_path)
    molname = molname or next(iter(top["molecules"]))
    names = list(top["atomtypes"])
    atoms = [(names.index(a["type"]), a["charge"], a["mass"]) for a in top["molecules"][molname]]
    return {"types": names, "eps_kj": [top["atomtypes"][n]["eps_kj"] for n in names],
            "sig_nm": [top["atomtypes"][n]["sigma_nm"] for n in names], "atoms": atoms,
            "xyz": read_pdb(pdb_path)[:len(atoms)].tolist()}


def tip3p_model() -> dict:
    """TIP3P as the reference ships it (water.top + tip3p.pdb at the repository root), from the committed fixture."""
    import json
    return json.loads((Path(__file__).resolve().parent.parent / "tests" / "golden" / "tip3p_model.json").read_text())


@dataclass
class AtomicSystem:
    """Monatomic/mainMonatomic.jl Requirements(r, ϵ, σ, box, r_cut)."""
    r: np.ndarray
    eps: np.ndarray
    sig: np.ndarray
    box: float
    r_cut: float

    @property
    def n(self):
        return self.r.shape[0]


def lj_lattice(n: int, rho: float = 0.75, r_cut: float = 2.5) -> AtomicSystem:
    """Config C: InitCubicGrid at rho*, ε=σ=1 (Monatomic/mainMonatomic.jl:343-356)."""
    r, L = init_cubic_grid(n, rho)
    return AtomicSystem(np.ascontiguousarray(r), np.ones(n), np.ones(n), float(L), float(r_cut))


def water_ion_mixture(n_mol: int, n_ions: int, rho: float = 0.033101144, seed: int = 11234) -> MolecularSystem:
    """A mixed topology (what the reference reads from Ewald/topol.top + mea.pdb style inputs: molecules with different site
    counts, firstAtom/lastAtom per molecule, Ewald/energy.jl:219-226): SPC/E water with `n_ions` of the molecules replaced by
    single-site ions of alternating charge ±1 e at the molecule's COM.  Three LJ types (O, H, ion; Lorentz-Berthelot mixing as
    Ewald/structs.jl:342-346, ion: σ = 2.35 Å, ε/k_B = 65 K)."""
    w = spce_lattice(n_mol, rho, seed)
    ion_of = np.zeros(n_mol, dtype=bool)
    ion_of[np.linspace(0, n_mol - 1, n_ions).astype(int)] = True
    coords, charge, atype, first, last, db = [], [], [], [], [], []
    sign = 1.0
    for m in range(n_mol):
        first.append(len(charge) + 1)
        if ion_of[m]:
            coords.append(w.com[m]); charge.append(sign); atype.append(3); db.append(np.zeros(3))
            sign = -sign
        else:
            for a in range(3):
                coords.append(w.coords[3 * m + a]); charge.append(w.charge[3 * m + a]); atype.append(w.atype[3 * m + a]); db.append(w.db[3 * m + a])
        last.append(len(charge))
    e = np.array([SPCE_EPS_O, 0.0, 65.0]); sg = np.array([SPCE_SIGMA_O, 0.0, 2.35])
    eps = np.sqrt(np.outer(e, e)); sig = 0.5 * (sg[:, None] + sg[None, :])
    quat = w.quat.copy()
    quat[ion_of] = np.array([1.0, 0.0, 0.0, 0.0])
    return MolecularSystem(np.array(coords), np.array(charge), np.array(atype, dtype=np.int64), np.array(first, dtype=np.int64),
                           np.array(last, dtype=np.int64), w.com.copy(), eps, sig, w.box, np.array(db), quat)
