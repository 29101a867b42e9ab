"""Config E, one GPU: device times of one potential() with the rho(k) rebuild beside the pair kernel and alone, for profiling under
ncu (python tools/prof_eval.py [--molecules N] [--reps R] [--overlap 0|1] [--pair-level L])"""
import argparse, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from metropolismontecarlo_b200 import systems
from metropolismontecarlo_b200.energy import water_engine

ap = argparse.ArgumentParser()
ap.add_argument("--molecules", type=int, default=256000)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--overlap", type=int, default=-1)
ap.add_argument("--pair-level", dest="pair_level", type=int, default=0)
ap.add_argument("--ctas", type=int, default=0)
a = ap.parse_args()
ms = systems.spce_lattice(a.molecules)
eng = water_engine(ms, 10.0)
eng.debug_set("pair_level", a.pair_level)
if a.ctas:
    eng.debug_set("v7_ctas_per_sm", a.ctas)
eng.set_timing(True)
for ov in ((0, 1) if a.overlap < 0 else (a.overlap,)):
    eng.debug_set("overlap_rhok", ov)
    rows = []
    for k in range(a.reps + 2):
        t0 = time.perf_counter()
        p = eng.potential("ewald")
        w = time.perf_counter() - t0
        t = eng.last_timings()
        if k >= 2:
            rows.append((t["pairs_ms"], t["rhok_ms"], t["bin_gather_ms"], t["total_ms"], 1e3 * w))
    r = np.median(np.array(rows), axis=0)
    print(f"overlap_rhok={ov} {eng.last_eval_info()['pair_kernel']}: pairs {r[0]:.4f} rhok {r[1]:.4f} bin+gather {r[2]:.4f} total(dev) {r[3]:.4f} wall {r[4]:.4f} ms  E/N={p.energy/ms.n_mol:.9f}")
eng.close()
