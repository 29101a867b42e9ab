"""k_pairs_v7 against the general kernels and the oracle on a range of boxes (GPU): python tools/check_v7.py"""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from metropolismontecarlo_b200 import systems
from metropolismontecarlo_b200.energy import water_engine
from oracle import oracle as ora

def rel(a, b): return abs(a - b) / max(abs(a), abs(b), 1e-300)
F = ("energy", "virial", "coulomb", "lj", "real", "recip", "self_")
def cmp(tag, a, b, tol):
    worst = max(rel(getattr(a, f), getattr(b, f)) for f in F)
    print(f"{tag}: worst rel {worst:.2e}", "OK" if worst < tol else "FAIL", flush=True)
    return worst < tol

ok = True
for name, ms in (("coord750", systems.load_nist(4)), ("D4000", systems.spce_lattice(4000)), ("lat3200", systems.spce_lattice(3200)),
                 ("E256k", systems.spce_lattice(256000))):
    eng = water_engine(ms, 10.0)
    res = {}
    for level in (0, 1, 2):
        eng.debug_set("pair_level", level)
        t0 = time.perf_counter()
        res[level] = (eng.potential("ewald"), eng.potential("wolf"), eng.last_eval_info())
        print(name, level, res[level][2], f"{time.perf_counter()-t0:.3f}s", flush=True)
    ok &= cmp(f"{name} ewald v7 vs fast", res[0][0], res[1][0], 1e-11)
    ok &= cmp(f"{name} wolf v7 vs fast", res[0][1], res[1][1], 1e-11)
    ok &= cmp(f"{name} ewald v7 vs general", res[0][0], res[2][0], 1e-11)
    assert res[0][2]["pairs_in_cutoff"] == res[2][2]["pairs_in_cutoff"], (res[0][2], res[2][2])
    eng.debug_set("pair_level", 0)
    # volume trial vs legacy
    L2 = ms.box * 1.013
    v0 = eng.volume_trial(L2, systems.ALPHA / L2, "ewald"); eng.volume_reject()
    eng.debug_set("pair_level", 1)
    v1 = eng.volume_trial(L2, systems.ALPHA / L2, "ewald"); eng.volume_reject()
    eng.debug_set("pair_level", 0)
    ok &= cmp(f"{name} volume trial v7 vs fast", v0, v1, 1e-11)
    # repeatability (bit-identical)
    a = eng.potential("ewald"); b = eng.potential("ewald")
    print(name, "bit-identical repeat:", all(getattr(a, f) == getattr(b, f) for f in F), flush=True)
    if ms.n_mol <= 4000:
        s = ora.System(ms.coords, ms.charge, ms.atype, ms.first_atom, ms.last_atom, ms.com, ms.eps, ms.sig)
        ew = ora.Ewald(systems.ALPHA / ms.box, systems.NK, systems.K_SQ_MAX, systems.FACTOR, ms.box)
        want = ora.potential_ewald(s, ew, 10.0, 10.0, ms.box, 8)
        ok &= cmp(f"{name} v7 vs oracle", a, want, 1e-10)
    if ms.n_mol == 256000:
        eng.set_timing(True)
        for k in range(5):
            eng.potential("ewald"); print("timings", eng.last_timings(), flush=True)
        ts = []
        for k in range(20):
            t0 = time.perf_counter(); eng.potential("ewald"); ts.append(time.perf_counter() - t0)
        print("wall per eval (ms): min %.3f median %.3f" % (1e3 * min(ts), 1e3 * float(np.median(ts))), flush=True)
    eng.close()
print("ALL OK" if ok else "SOME FAILED")
