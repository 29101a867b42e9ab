"""Time mmc_loop_run_atoms_device against mmc_loop_run_atoms on config C (32 000 LJ atoms)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from metropolismontecarlo_b200 import systems
from metropolismontecarlo_b200.energy import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
at = systems.lj_lattice(32000, 0.75, 2.5)
u = np.random.default_rng(11234).random(5 * n)
for div in (30.0, 150.0):
    outs = []
    for device in (False, True):
        eng = Engine()
        eng.upload_atoms(at)
        p0 = eng.potential("atoms")
        r = at.r.copy()
        eng.loop_run_atoms(1.0, at.box / div, r, u, 300, p0.energy, p0.virial, device=device)
        eng.upload_atoms(at)
        r = at.r.copy()
        t0 = time.perf_counter()
        rc, acc, delta, st = eng.loop_run_atoms(1.0, at.box / div, r, u, n, p0.energy, p0.virial, device=device)
        dt = time.perf_counter() - t0
        outs.append(acc)
        print("dr=L/%g" % div, "device" if device else "host  ", "rc", rc, "moves/s %.0f" % (n / dt), "us/move %.2f" % (1e6 * dt / n),
              "accepted", st.n_accepted, "E", st.total_energy, "fresh", eng.potential("atoms").energy)
        eng.close()
    print("  same record:", np.array_equal(outs[0], outs[1]))
