"""ONE rank's share of a sharded full-energy evaluation on one GPU (rank r of `world`, no exchange): kernel times for
v7_ctas_per_sm = 4, 3, 2 and with / without the rho(k) rebuild beside the pair kernel.  python tools/prof_shard.py [world] [rank]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from metropolismontecarlo_b200 import systems
from metropolismontecarlo_b200.energy import water_engine
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ms = systems.spce_lattice(256000)
eng = water_engine(ms, 10.0, rank=rank, world=world)
eng.set_timing(True)
vec = torch.zeros(eng.partial_count(), dtype=torch.float64, device="cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for ov in (0, 1):
    eng.debug_set("overlap_rhok", ov)
    for ctas in (4, 3, 2, 1):
        eng.debug_set("v7_ctas_per_sm", ctas)
        rows = []
        for k in range(12):
            flush.fill_(k); torch.cuda.synchronize()
            eng.potential_partial("ewald", vec.data_ptr())
            eng.potential_finalize("ewald", vec.data_ptr())      # (this rank's vector alone: only the timings matter here)
            t = eng.last_timings()
            rows.append((t["pairs_ms"], t["rhok_ms"], t["bin_gather_ms"], t["total_ms"]))
        r = np.median(np.array(rows[3:]), axis=0)
        print(f"world {world} rank {rank} overlap {ov} ctas/SM {ctas}: pairs {r[0]*1e3:6.1f} us  rhok {r[1]*1e3:6.1f}  bin+gather {r[2]*1e3:6.1f}  total(dev) {r[3]*1e3:6.1f} us", flush=True)
eng.debug_set("overlap_rhok", 1); eng.debug_set("v7_ctas_per_sm", 4)
for pct in (0, 30, 50, 70):
    eng.debug_set("rhok_early_pct", pct)
    rows = []
    for k in range(12):
        flush.fill_(k); torch.cuda.synchronize()
        eng.potential_partial("ewald", vec.data_ptr())
        eng.potential_finalize("ewald", vec.data_ptr())
        t = eng.last_timings()
        rows.append((t["pairs_ms"], t["rhok_ms"], t["bin_gather_ms"], t["total_ms"]))
    r = np.median(np.array(rows[3:]), axis=0)
    print(f"world {world} rank {rank} rhok_early_pct {pct}: pairs {r[0]*1e3:6.1f} us  total(dev) {r[3]*1e3:6.1f} us", flush=True)
eng.close()
