"""Moves/s micro-benchmark for the per-move path (configs A/B/C), optional sync_mode."""
import argparse, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from metropolismontecarlo_b200 import systems
from metropolismontecarlo_b200.energy import Engine, LoopParams, water_engine

ap = argparse.ArgumentParser()
ap.add_argument("--moves", type=int, default=10000)
ap.add_argument("--sync", type=int, default=0)
ap.add_argument("--which", default="ABC")
a = ap.parse_args()
u = np.random.default_rng(11234).random(8 * a.moves + 100)
ms = systems.load_nist(4)
for name, style, sid in (("A", "ewald", 0), ("B", "wolf", 1)):
    if name not in a.which:
        continue
    eng = water_engine(ms, 10.0, sync_mode=a.sync)
    p0 = eng.potential(style)
    com, quat = ms.com.copy(), ms.quat.copy()
    eng.loop_run(LoopParams(298.15, 0.316555789, 0.05, 0.5, 1.0, sid, 1), com, quat, ms.db, u, 300, p0.energy, p0.virial)
    eng.upload_system(ms, 10.0, 10.0); p0 = eng.potential(style)
    com, quat = ms.com.copy(), ms.quat.copy()
    t0 = time.perf_counter()
    rc, acc, d, st = eng.loop_run(LoopParams(298.15, 0.316555789, 0.05, 0.5, 1.0, sid, 1), com, quat, ms.db, u, a.moves, p0.energy, p0.virial)
    dt = time.perf_counter() - t0
    print(name, style, "sync", a.sync, "us/move %.2f" % (1e6 * dt / a.moves), "moves/s %.0f" % (a.moves / dt), "acc", st.n_accepted)
    eng.close()
if "C" in a.which:
    at = systems.lj_lattice(32000, 0.75, 2.5)
    eng = Engine(sync_mode=a.sync); eng.upload_atoms(at); p0 = eng.potential("atoms")
    r = at.r.copy(); eng.loop_run_atoms(1.0, at.box / 30, r, u, 300, p0.energy, p0.virial)
    eng.upload_atoms(at); r = at.r.copy()
    t0 = time.perf_counter()
    rc, acc, d, st = eng.loop_run_atoms(1.0, at.box / 30, r, u, a.moves, p0.energy, p0.virial)
    dt = time.perf_counter() - t0
    print("C lj32k sync", a.sync, "us/move %.2f" % (1e6 * dt / a.moves), "moves/s %.0f" % (a.moves / dt), "acc", st.n_accepted)
