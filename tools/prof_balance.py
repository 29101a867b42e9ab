"""k_pairs_v6 unit scheduling: static round-robin deal (v6_dynamic 0) against tickets (1), on one GPU and on one rank's share of
a 2/4/8-rank evaluation (emulated: the partial pass of rank world/2; wall clock around the call + device sync)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from metropolismontecarlo_b200 import systems
from metropolismontecarlo_b200.energy import water_engine

ms = systems.spce_lattice(256000)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for world in (1, 2, 4, 8):
    eng = water_engine(ms, 10.0, rank=world // 2, world=world)
    eng.set_timing(True)
    buf = torch.zeros(eng.partial_count(), dtype=torch.float64, device="cuda")
    for dyn in (0, 1, 0, 1):
        eng.debug_set("v6_dynamic", dyn)
        ts, ps, ws = [], [], []
        for k in range(14):
            flush.fill_(k)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if world == 1:
                p = eng.potential("ewald")
                e = p.energy / ms.n_mol
            else:
                eng.potential_partial("ewald", buf.data_ptr())
                torch.cuda.synchronize()
                e = float(buf[2].item())
            ws.append((time.perf_counter() - t0) * 1e3)
            t = eng.last_timings()
            ts.append(t["total_ms"]); ps.append(t["pairs_ms"])
        print(f"world {world} dynamic {dyn}: pairs {np.median(ps[2:]):.4f} ms  total {np.median(ts[2:]):.4f} ms  wall {np.median(ws[2:]):.4f} ms  check {e!r}", flush=True)
    eng.close()
