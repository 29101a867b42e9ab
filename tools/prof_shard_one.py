"""One rank of a sharded evaluation, a few times (for ncu / launch lists).  python tools/prof_shard_one.py [world] [rank] [reps]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from metropolismontecarlo_b200 import systems
from metropolismontecarlo_b200.energy import water_engine
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 3
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
ms = systems.spce_lattice(256000)
eng = water_engine(ms, 10.0, rank=rank, world=world)
vec = torch.zeros(eng.partial_count(), dtype=torch.float64, device="cuda")
for k in range(reps):
    eng.potential_partial("ewald", vec.data_ptr())
    torch.cuda.synchronize()
eng.close()
