"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): every new kernel once."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from metropolismontecarlo_b200 import systems
from metropolismontecarlo_b200.energy import Engine, LoopParams, water_engine
ms = systems.spce_lattice(4000)
eng = water_engine(ms, 10.0)
p = eng.potential("ewald")
print("k_pairs path:", eng.last_eval_info(), p.energy)
v = eng.volume_trial(ms.box * 1.01, systems.ALPHA / (ms.box * 1.01), "ewald")
eng.volume_reject()
eng.close()
ms = systems.load_nist(1)
u = np.random.default_rng(1).random(4000)
for cluster in (1, 8):
    eng = water_engine(ms, 9.0)
    eng.debug_set("chain_cluster", cluster)
    g0 = eng.potential("ewald")
    com, quat = ms.com.copy(), ms.quat.copy()
    rc, acc, delta, st = eng.loop_run(LoopParams(298.15, 0.3, 0.05, 0.5, 1.0, 0, 1), com, quat, ms.db, u, 300, g0.energy, g0.virial, device=True)
    print("chain cluster", cluster, rc, st.n_accepted)
    eng.close()
at = systems.lj_lattice(1000, 0.75, 2.5)
eng = Engine()
eng.upload_atoms(at)
g0 = eng.potential("atoms")
r = at.r.copy()
rc, acc, delta, st = eng.loop_run_atoms(1.0, at.box / 100, r, u, 300, g0.energy, g0.virial, device=True)
print("atoms chain", rc, st.n_accepted)
eng.close()
