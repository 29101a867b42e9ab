import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from metropolismontecarlo_b200 import systems
from metropolismontecarlo_b200.energy import water_engine
ms = systems.spce_lattice(256000)
eng = water_engine(ms, 10.0)
eng.set_timing(True)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
ref = None
for mode, pct in ((0, 0), (2, 0), (1, 0), (1, 15), (1, 25), (1, 35), (1, 50)):
    eng.debug_set("overlap_rhok", mode); eng.debug_set("rhok_early_pct", pct)
    rows = []
    for k in range(14):
        flush.fill_(k); torch.cuda.synchronize()
        p = eng.potential("ewald"); t = eng.last_timings()
        rows.append((t["pairs_ms"], t["total_ms"]))
    if ref is None: ref = p
    assert abs(p.energy - ref.energy) < 1e-12 * abs(ref.energy) and abs(p.recip - ref.recip) < 1e-11 * abs(ref.recip)
    r = np.median(np.array(rows[4:]), axis=0)
    print(f"overlap_rhok {mode} early {pct:2d}%: pairs {r[0]:.4f} ms  total(dev, cold L2) {r[1]:.4f} ms", flush=True)
