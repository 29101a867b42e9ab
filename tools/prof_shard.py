"""Timings of ONE rank's share of a sharded full-energy evaluation (no NCCL: the partial vector is finalised as is)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from metropolismontecarlo_b200 import systems
from metropolismontecarlo_b200.energy import water_engine
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ms = systems.spce_lattice(256000)
eng = water_engine(ms, 10.0, rank=0, world=world)
eng.set_timing(True)
vec = torch.zeros(eng.partial_count(), dtype=torch.float64, device="cuda")
for k in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.potential_partial("ewald", vec.data_ptr())
    t1 = time.perf_counter()
    p = eng.potential_finalize("ewald", vec.data_ptr())
    t2 = time.perf_counter()
    print(k, "enqueue ms %.3f" % ((t1 - t0) * 1e3), "finalize(wait) ms %.3f" % ((t2 - t1) * 1e3), eng.last_timings())
eng.close()
