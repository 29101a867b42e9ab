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
# ---- round 2: mixed topologies (padded copy, overlap rows), all-molecule rows, large k-sets, sharded partials, host path
ms = systems.water_ion_mixture(4096, 300)
ions = np.flatnonzero(ms.last_atom == ms.first_atom)
b = ions[1]
ms.coords[ms.first_atom[b] - 1] = ms.com[b] = np.clip(ms.com[ions[0]] + np.array([0.3, 0.2, 0.0]), 0.0, ms.box)
eng = water_engine(ms, 10.0)
p = eng.potential("ewald")
print("mixed:", eng.last_eval_info(), p.overlaps)
eng.energy_all("ewald")
v = eng.volume_trial(ms.box * 1.01, systems.ALPHA / (ms.box * 1.01), "ewald")
eng.volume_reject()
eng.set_intramolecular(True); eng.potential("ewald"); eng.set_intramolecular(False)
eng.close()
ms = systems.spce_lattice(4000)
eng = Engine()
eng.upload_system(ms, 10.0, 10.0)
for nk, k2 in ((12, 145), (7, 50)):
    eng.PrepareEwaldVariables(0.32, nk, k2)
    print("large k:", nk, eng.potential("ewald").recip)
eng.close()
import torch
for r in range(2):
    e = water_engine(ms, 10.0, rank=r, world=2)
    vec = torch.zeros(e.partial_count(), dtype=torch.float64, device="cuda")
    e.potential_partial("ewald", vec.data_ptr())
    torch.cuda.synchronize()
    e.close()
ms = systems.spce_lattice(40000)
eng = water_engine(ms, 10.0)
print("host path:", eng.potential_host(ms.coords, ms.com, "ewald").energy)
eng.close()
