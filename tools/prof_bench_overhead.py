"""What the bench's CUDA-event interval around one evaluation contains beyond the kernels: library timing events on/off,
last_timings() inside the loop or not.  python tools/prof_bench_overhead.py"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from metropolismontecarlo_b200 import systems
from metropolismontecarlo_b200.energy import Engine
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
eng = Engine(device=0, stream=stream.cuda_stream)
ms = systems.spce_lattice(256000)
eng.upload_system(ms, 10.0, 10.0)
eng.PrepareEwaldVariables(systems.ALPHA / ms.box)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for timing in (True, False):
    for read in (True, False):
        if read and not timing:
            continue
        eng.set_timing(timing)
        n = 60
        ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        walls = []
        torch.cuda.synchronize()
        for k in range(n):
            flush.fill_(k & 0xff)
            ev0[k].record()
            t0 = time.perf_counter()
            eng.potential("ewald")
            walls.append(time.perf_counter() - t0)
            ev1[k].record()
            if read:
                eng.last_timings()
        torch.cuda.synchronize()
        ms_ev = np.median([a.elapsed_time(b) for a, b in zip(ev0, ev1)][10:])
        print(f"timing {timing} read_in_loop {read}: event interval {ms_ev:.4f} ms, host wall of the call {1e3*np.median(walls[10:]):.4f} ms")
eng.close()
