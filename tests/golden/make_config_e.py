"""Golden Properties of config E (BASELINE.json: synthetic SPC/E box, 256 000 molecules, 768 000 sites, L = 197.757 Å) from ONE full
run of the oracle's potential(…, "ewald") — the reference's own O(N²) double-counted algorithm (Ewald/energy.jl:946-1032), OpenMP
over rows, everything else serial as in the reference.  About 1-3 minutes on 8-16 host cores.

    python tests/golden/make_config_e.py     ->  tests/golden/config_e_properties.json

The GPU test (tests/test_gpu_fullsize.py) compares mmc_potential's seven Properties with these numbers at 1e-10 relative.
"""
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from metropolismontecarlo_b200 import systems  # noqa: E402
from oracle import oracle as ora  # noqa: E402


def main():
    ora.build()
    n = 256000
    ms = systems.spce_lattice(n)
    s = ora.System(ms.coords, ms.charge, ms.atype, ms.first_atom, ms.last_atom, ms.com, ms.eps, ms.sig)
    ew = ora.Ewald(systems.ALPHA / ms.box, systems.NK, systems.K_SQ_MAX, systems.FACTOR, ms.box)
    threads = os.cpu_count() or 1
    t0 = time.time()
    p = ora.potential_ewald(s, ew, 10.0, 10.0, ms.box, threads)
    dt = time.time() - t0
    t0 = time.time()
    w = ora.potential_wolf(s, ew, 10.0, 10.0, ms.box, threads)
    dtw = time.time() - t0
    out = {"n_molecules": n, "box": ms.box, "r_cut": 10.0, "kappa": systems.ALPHA / ms.box, "nk": systems.NK, "k_sq_max": systems.K_SQ_MAX,
           "generator": "systems.spce_lattice(256000): InitCubicGrid at rho = 0.033101144 A^-3, random quaternions, seed 11234",
           "oracle_seconds": {"ewald": dt, "wolf": dtw, "threads": threads},
           "ewald": {k: getattr(p, k) for k in ("energy", "virial", "coulomb", "lj", "real", "recip", "self_", "overlaps")},
           "wolf": {k: getattr(w, k) for k in ("energy", "virial", "coulomb", "lj", "real", "wolf_const", "overlaps")}}
    path = Path(__file__).resolve().parent / "config_e_properties.json"
    path.write_text(json.dumps(out, indent=1) + "\n")
    print("wrote", path, f"({dt:.0f} s + {dtw:.0f} s on {threads} threads)")
    print(json.dumps(out["ewald"], indent=1))


if __name__ == "__main__":
    main()
