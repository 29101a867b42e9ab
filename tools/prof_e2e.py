"""Where the end-to-end evaluation spends its time: copies, repack, evaluation (config E, pinned host arrays)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from metropolismontecarlo_b200 import systems
from metropolismontecarlo_b200.energy import water_engine
ms = systems.spce_lattice(256000)
eng = water_engine(ms, 10.0)
coords = torch.from_numpy(np.ascontiguousarray(ms.coords)).pin_memory().numpy()
com = torch.from_numpy(np.ascontiguousarray(ms.com)).pin_memory().numpy()
def t(fn, n=10):
    fn(); fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("potential (resident)        ms %.3f" % t(lambda: eng.potential("ewald")))
print("upload_positions            ms %.3f" % t(lambda: eng.upload_positions(coords, com)))
print("upload + potential          ms %.3f" % t(lambda: (eng.upload_positions(coords, com), eng.potential("ewald"))))
print("potential_host (pipelined)  ms %.3f" % t(lambda: eng.potential_host(coords, com, "ewald")))
print("potential_host wolf         ms %.3f" % t(lambda: eng.potential_host(coords, com, "wolf")))
d = torch.empty(coords.size, dtype=torch.float64, device="cuda"); h = torch.from_numpy(coords.reshape(-1))
print("raw H2D 18.4 MB pinned      ms %.3f" % t(lambda: d.copy_(h, non_blocking=True)))
