import sys, time, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from metropolismontecarlo_b200 import systems
from metropolismontecarlo_b200.energy import water_engine
ms = systems.spce_lattice(256000)
eng = water_engine(ms, 10.0)
co = torch.from_numpy(ms.coords).pin_memory().numpy(); cm = torch.from_numpy(ms.com).pin_memory().numpy()
for win, ch in ((4, 8), (1, 1)):
    eng.debug_set("host_windows", win); eng.debug_set("host_chunks", ch)
    for k in range(4): eng.potential_host(co, cm, "ewald")
    os.environ["MMC_TRACE_HOST"] = "1"
    print(f"---- windows {win} chunks {ch}", file=sys.stderr, flush=True)
    eng.potential_host(co, cm, "ewald")
    del os.environ["MMC_TRACE_HOST"]
