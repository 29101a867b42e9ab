"""Minimal driver for ncu: upload config E (or --molecules N) and run a few full-energy evaluations."""
import argparse, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from metropolismontecarlo_b200 import systems
from metropolismontecarlo_b200.energy import water_engine

ap = argparse.ArgumentParser()
ap.add_argument("--molecules", type=int, default=256000)
ap.add_argument("--evals", type=int, default=4)
ap.add_argument("--style", default="ewald")
ap.add_argument("--pair-level", type=int, default=0, help="0 v6, 1 v5, 2 v4, 3 v3, 4 fast, 5 general")
ap.add_argument("--overlap", type=int, default=1)
ap.add_argument("--rows", action="store_true", help="also time mmc_energy_all (per-molecule rows, general kernel with FP64 atomics)")
a = ap.parse_args()
ms = systems.spce_lattice(a.molecules) if a.molecules != 750 else systems.load_nist(4)
eng = water_engine(ms, 10.0)
eng.set_timing(True)
eng.debug_set("pair_level", a.pair_level)
eng.debug_set("overlap_rhok", a.overlap)
for k in range(a.evals):
    t0 = time.perf_counter()
    p = eng.potential(a.style)
    dt = time.perf_counter() - t0
    print(k, "E/N", p.energy / ms.n_mol, "wall ms", dt * 1e3, eng.last_timings(), eng.last_eval_info())
if a.rows:
    for k in range(3):
        t0 = time.perf_counter()
        lj, vir, qq, ov = eng.energy_all(a.style)
        dt = time.perf_counter() - t0
        print("energy_all", k, "wall ms", dt * 1e3, "pair kernel ms", eng.last_timings()["pairs_ms"], "sum/2", lj.sum() / 2, qq.sum() / 2)
eng.close()
