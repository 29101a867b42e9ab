"""Config D volume trial (4000 SPC/E molecules): host call -> Properties on the host, a different box every trial.  python tools/prof_volume.py"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from metropolismontecarlo_b200 import systems
from metropolismontecarlo_b200.energy import water_engine
ms = systems.spce_lattice(4000)
eng = water_engine(ms, 10.0)
rng = np.random.default_rng(3)
V = ms.box ** 3
boxes = (V + (rng.random(600) - 0.5) * 0.01 * V) ** (1.0 / 3.0)
for rep in range(3):
    t0 = time.perf_counter()
    for L in boxes[rep * 200:(rep + 1) * 200]:
        p = eng.volume_trial(L, systems.ALPHA / L, "ewald")
        eng.volume_reject()
    dt = (time.perf_counter() - t0) / 200
    print(f"volume trial: {1e3 * dt:.4f} ms  ({eng.last_eval_info()['pair_kernel']})  E = {p.energy:.6f}")
from oracle import oracle as ora
from tests.util import ora_system
s = ora_system(ms)
L = boxes[-1]
ora.volume_scale(s, ms.box, L)
w = ora.potential_ewald(s, ora.Ewald(systems.ALPHA / L, 5, 27, systems.FACTOR, L), 10.0, 10.0, L, 8)
print("last trial vs oracle: rel", abs(p.energy - w.energy) / abs(w.energy))
eng.close()
