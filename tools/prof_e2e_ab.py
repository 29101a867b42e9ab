"""mmc_potential_host on one GPU, A/B of library switches in one process (same box, same PCIe link): python tools/prof_e2e_ab.py"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from metropolismontecarlo_b200 import systems
from metropolismontecarlo_b200.energy import water_engine
ms = systems.spce_lattice(256000)
eng = water_engine(ms, 10.0)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
coords, com = pin(ms.coords), pin(ms.com)
def rate(n=40):
    for _ in range(5):
        eng.potential_host(coords, com, "ewald")
    t0 = time.perf_counter()
    for _ in range(n):
        eng.potential_host(coords, com, "ewald")
    return n / (time.perf_counter() - t0)
for rep in range(2):
    for mb in (0, 1):
        eng.debug_set("host_mailbox", mb)
        for win, ch in ((3, 6), (1, 1), (2, 4), (4, 8)):
            eng.debug_set("host_windows", win); eng.debug_set("host_chunks", ch)
            print(f"host_mailbox {mb} windows {win} chunks {ch}: {rate():7.1f} evals/s", flush=True)
# plain H2D bandwidth of this box for the 24.6 MB site array
d = torch.empty(coords.size, dtype=torch.float64, device="cuda")
src = torch.from_numpy(coords)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20):
    d.copy_(src.reshape(-1), non_blocking=True)
torch.cuda.synchronize()
print(f"H2D 24.6 MB: {20 * coords.nbytes / (time.perf_counter() - t0) / 1e9:.1f} GB/s")
eng.close()
