"""Concurrent block-of-moves chains on one GPU: aggregate moves/s for R = 1..16 replicas (python tools/prof_replicas.py)"""
import sys, time, threading
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from metropolismontecarlo_b200 import systems
from metropolismontecarlo_b200.energy import LoopParams, water_engine

ms = systems.load_nist(4)
prm = LoopParams(298.15, 0.316555789, 0.05, 0.5, 1.0, 0, 1)
n_moves = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
cluster = int(sys.argv[2]) if len(sys.argv) > 2 else 8
Rs = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else (1, 2, 4, 8, 12, 16, 18)
for R in Rs:
    engs = [water_engine(ms, 10.0) for _ in range(R)]
    for e in engs:
        e.debug_set("chain_cluster", cluster)
    us = [np.random.default_rng(11234 + r).random(8 * n_moves) for r in range(R)]
    best = 1e9
    for rep in range(3):
        p0 = []
        for e in engs:
            e.upload_system(ms, 10.0, 10.0); p0.append(e.potential("ewald"))
        bar = threading.Barrier(R + 1)
        def run(r):
            com, quat = ms.com.copy(), ms.quat.copy()
            bar.wait()
            engs[r].loop_run(prm, com, quat, ms.db, us[r], n_moves, p0[r].energy, p0[r].virial, device=True)
        th = [threading.Thread(target=run, args=(r,)) for r in range(R)]
        for t in th: t.start()
        bar.wait(); t0 = time.perf_counter()
        for t in th: t.join()
        if rep: best = min(best, time.perf_counter() - t0)
    print(f"cluster {cluster} R={R:2d}: {R*n_moves/best/1e3:8.1f} k moves/s aggregate, {1e6*best/n_moves:6.2f} us/move per replica", flush=True)
    for e in engs: e.close()
