"""Two processes, two GPUs: the sharded full-energy evaluation with the exchange over NVLink peer memory
(CUDA IPC handles carried by torch.distributed, k_peer_push / k_peer_sum) against the unsharded evaluation and
against the NCCL all-reduce form.  Skipped on a box with fewer than two GPUs."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from metropolismontecarlo_b200 import systems
    from metropolismontecarlo_b200.energy import Engine
    from metropolismontecarlo_b200.sharding import setup_peer_exchange, sharded_potential
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ms = systems.spce_lattice(40000)
    eng = Engine(device=rank, rank=rank, world=world, stream=stream.cuda_stream)
    eng.upload_system(ms, 10.0, 10.0)
    eng.PrepareEwaldVariables(systems.ALPHA / ms.box)
    setup_peer_exchange(eng, world)
    res = []
    for _ in range(3):                               # repeated: epoch / parity protocol
        p = eng.potential_sharded("ewald")
        res.append((p.energy, p.virial, p.recip, p.real))
    vec = torch.zeros(eng.partial_count(), dtype=torch.float64, device=dev)
    q = sharded_potential(eng, "ewald", vec, world)  # the NCCL form of the same evaluation
    res.append((q.energy, q.virial, q.recip, q.real))
    for _ in range(2):                               # the domain-decomposed host-array form (each rank copies its slab only)
        d = eng.potential_host(ms.coords, ms.com, "ewald")
        res.append((d.energy, d.virial, d.recip, d.real))
    b_gather = float(eng.last_host_bytes())
    # a different state (the molecules re-ordered: the same energy): the previous call's site blocks go speculatively, the rest after
    perm = np.roll(np.arange(ms.n_mol), ms.n_mol // 3)
    d = eng.potential_host(ms.coords.reshape(-1, 3, 3)[perm].reshape(-1, 3).copy(), ms.com[perm].copy(), "ewald")
    res.append((d.energy, d.virial, d.recip, d.real))
    eng.debug_set("com_allgather", 0)                # every rank copies all COMs itself (both ranks switch together)
    d = eng.potential_host(ms.coords, ms.com, "ewald")
    res.append((d.energy, d.virial, d.recip, d.real))
    d = eng.potential_host(ms.coords, ms.com, "ewald")
    res.append((d.energy, d.virial, d.recip, d.real))
    res.append((b_gather, float(eng.last_host_bytes()), 0.0, 0.0))
    np.save(f"{out}.{rank}.npy", np.array(res))
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_peer_exchange_two_processes(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from metropolismontecarlo_b200 import systems
    from metropolismontecarlo_b200.energy import water_engine
    out = str(tmp_path / "res")
    mp.spawn(_worker, args=(2, 29541, out), nprocs=2, join=True)
    r0, r1 = np.load(out + ".0.npy"), np.load(out + ".1.npy")
    assert np.array_equal(r0[:3], r1[:3])            # bit-identical totals on both ranks, every repetition
    assert np.array_equal(r0[0], r0[1]) and np.array_equal(r0[0], r0[2])
    ms = systems.spce_lattice(40000)
    eng = water_engine(ms, 10.0)
    want = eng.potential("ewald")
    eng.close()
    for k, w in enumerate((want.energy, want.virial, want.recip, want.real)):
        assert abs(r0[0][k] - w) <= 1e-12 * abs(w)
        assert abs(r0[3][k] - w) <= 1e-12 * abs(w)   # NCCL form agrees too
        assert abs(r0[4][k] - w) <= 1e-12 * abs(w) and abs(r0[5][k] - w) <= 1e-12 * abs(w)   # and the domain-decomposed host form
    assert np.array_equal(r0[4:9], r1[4:9])
    for j in (6, 7, 8):                              # re-ordered state; COMs copied by every rank instead of gathered over NVLink
        for k, w in enumerate((want.energy, want.virial, want.recip, want.real)):
            assert abs(r0[j][k] - w) <= 1e-11 * abs(w), (j, k)
    full = 24 * (ms.n_sites + ms.n_mol)
    assert max(r0[9][1], r1[9][1]) < 0.85 * full     # each rank copied its slab, not the whole site array
    # with the all-gather a rank copies only its slice of the COMs (slices begin on multiples of 256 molecules)
    cut = ms.n_mol // 2 // 256 * 256
    assert r0[9][1] - r0[9][0] == 24 * (ms.n_mol - cut) and r1[9][1] - r1[9][0] == 24 * cut
