"""SURVEY 8d "E2": config E with a converged k-space sum (kappa = 0.32 /A, i.e. kappa r_cut = 3.2; nk = 12, k^2 < 145: 3.6 k k-vectors).
python tools/prof_e2.py [--ncu]"""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from metropolismontecarlo_b200 import systems
from metropolismontecarlo_b200.energy import Engine
ms = systems.spce_lattice(256000)
eng = Engine()
eng.upload_system(ms, 10.0, 10.0)
nkv = eng.PrepareEwaldVariables(0.32, 12, 145)
eng.set_timing(True)
eng.debug_set("overlap_rhok", 0)
rows = []
for k in range(2 if "--ncu" in sys.argv else 8):
    t0 = time.perf_counter(); p = eng.potential("ewald"); w = time.perf_counter() - t0
    t = eng.last_timings(); rows.append((t["pairs_ms"], t["rhok_ms"], t["total_ms"], 1e3 * w))
r = np.median(np.array(rows[1:]), axis=0)
flops = (14 * nkv + 168) * ms.n_sites + 6 * nkv
print(f"E2: {nkv} k-vectors, {eng.last_eval_info()['pair_kernel']}: pairs {r[0]:.3f} ms, rho(k) rebuild {r[1]:.3f} ms = {flops/r[1]/1e9:.1f} TFLOP/s algorithmic, total {r[2]:.3f} ms, wall {r[3]:.3f} ms, E/N = {p.energy/ms.n_mol:.6f} K")
eng.close()
