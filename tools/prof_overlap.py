"""Schedule experiments for the full evaluation on config E: where the rho(k) rebuild forks (overlap_rhok 0/1/3) and how many
CTAs it is cut into (rhok_split); prints the median step time of each combination (timing events of the library)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from metropolismontecarlo_b200 import systems
from metropolismontecarlo_b200.energy import water_engine

ms = systems.spce_lattice(256000)
eng = water_engine(ms, 10.0)
eng.set_timing(True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for ov in (1, 3, 0):
    for split in (1, 2, 4, 8, 16):
        eng.debug_set("overlap_rhok", ov)
        eng.debug_set("rhok_split", split)
        ts, ps = [], []
        for k in range(12):
            flush.fill_(k)
            torch.cuda.synchronize()
            p = eng.potential("ewald")
            t = eng.last_timings()
            ts.append(t["total_ms"]); ps.append(t["pairs_ms"])
        print(f"overlap {ov} split {split:2d}: total {np.median(ts[2:]):.4f} ms  pairs {np.median(ps[2:]):.4f} ms  E/N {p.energy / ms.n_mol:.9f}")
eng.close()
