"""Time mmc_loop_run_device against mmc_loop_run on coord750 (configs A/B)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from metropolismontecarlo_b200 import systems
from metropolismontecarlo_b200.energy import LoopParams, water_engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
ms = systems.load_nist(4)
u = np.random.default_rng(11234).random(8 * n)
for style, sid in (("ewald", 0), ("wolf", 1)):
    for device in (False, 1, 8):
        eng = water_engine(ms, 10.0)
        if device:
            eng.debug_set("chain_cluster", device)
        p0 = eng.potential(style)
        com, quat = ms.com.copy(), ms.quat.copy()
        eng.loop_run(LoopParams(298.15, 0.316555789, 0.05, 0.5, 1.0, sid, 1), com, quat, ms.db, u, 300, p0.energy, p0.virial, device=bool(device))
        eng.upload_system(ms, 10.0, 10.0)
        p0 = eng.potential(style)
        com, quat = ms.com.copy(), ms.quat.copy()
        t0 = time.perf_counter()
        rc, acc, delta, st = eng.loop_run(LoopParams(298.15, 0.316555789, 0.05, 0.5, 1.0, sid, 1), com, quat, ms.db, u, n, p0.energy, p0.virial, device=bool(device))
        dt = time.perf_counter() - t0
        print(style, ("device C=%d" % device) if device else "host      ", "rc", rc, "moves/s %.0f" % (n / dt), "us/move %.2f" % (1e6 * dt / n), "accepted", st.n_accepted,
              "E", st.total_energy, "fresh", eng.potential(style).energy)
        eng.close()
