"""mmc_potential_host on config E, one GPU: wall time per call for window / chunk / rho(k)-granularity settings"""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from metropolismontecarlo_b200 import systems
from metropolismontecarlo_b200.energy import water_engine
ms = systems.spce_lattice(256000)
eng = water_engine(ms, 10.0)
co = torch.from_numpy(ms.coords).pin_memory().numpy(); cm = torch.from_numpy(ms.com).pin_memory().numpy()
ref = eng.potential("ewald")
for split in (1, 4, 16):
    eng.debug_set("rhok_split", split)
    for win, ch in ((1, 1), (2, 8), (3, 8), (4, 8), (2, 4), (3, 6)):
        eng.debug_set("host_windows", win); eng.debug_set("host_chunks", ch)
        ts = []
        for k in range(25):
            t0 = time.perf_counter(); p = eng.potential_host(co, cm, "ewald"); ts.append(time.perf_counter() - t0)
        assert abs(p.energy - ref.energy) < 1e-11 * abs(ref.energy)
        print(f"rhok_split {split:2d} windows {win} chunks {ch}: median {1e3*np.median(ts[5:]):.3f} ms  min {1e3*min(ts):.3f} ms  -> {1/np.median(ts[5:]):.0f} evals/s", flush=True)
eng.close()
