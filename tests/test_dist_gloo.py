"""CPU, world_size 2, gloo: the host-side logic of the sharded full-energy path.

Each rank builds the partial-sum vector for ITS share of the molecules (pair rows) and of the
sites (ρ(k)) with the oracle, the vectors are all-reduced exactly like bench.py does with NCCL,
and the sum must equal the unsharded oracle evaluation.  This covers shard_range (the integer
partition the library also uses), the vector layout and the collective plumbing; the GPU kernels
behind potential_partial are covered by tests/test_gpu_parity.py::test_sharded_partials_sum_to_unsharded.
"""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from metropolismontecarlo_b200 import sharding, systems
    from oracle import oracle as ora
    ms = systems.load_nist(1)
    rc = 9.0
    s = ora.System(ms.coords, ms.charge, ms.atype, ms.first_atom, ms.last_atom, ms.com, ms.eps, ms.sig)
    ew = ora.Ewald(systems.ALPHA / ms.box, systems.NK, systems.K_SQ_MAX, systems.FACTOR, ms.box)
    vec = torch.zeros(sharding.partial_len(ew.nkvecs), dtype=torch.float64)
    m0, m1 = sharding.shard_range(ms.n_mol, rank, world)
    lj, vir, real, nov = ora.potential_rows(s, ew.kappa, rc, rc, ms.box, m0, m1)
    vec[0], vec[1], vec[2], vec[3] = lj / 8, vir * 3 / 48, real / 2, nov      # rows → unique-pair sums
    s0, s1 = sharding.shard_range(ms.n_sites, rank, world)
    ora.RecipLong(ew, ms.coords[s0:s1], ms.charge[s0:s1], ms.box)
    vec[sharding.NSCAL:] = torch.from_numpy(ew.sum_new.ravel().copy())
    dist.all_reduce(vec)
    if rank == 0:
        np.save(out, vec.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    from metropolismontecarlo_b200.sharding import shard_range
    for n in (1, 7, 96026, 768000):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(180)
def test_two_rank_partials_allreduce_to_unsharded(tmp_path):
    out = tmp_path / "vec.npy"
    mp.spawn(_worker, args=(2, 29533, str(out)), nprocs=2, join=True)
    vec = np.load(out)
    from metropolismontecarlo_b200 import sharding, systems
    from oracle import oracle as ora
    ms = systems.load_nist(1)
    s = ora.System(ms.coords, ms.charge, ms.atype, ms.first_atom, ms.last_atom, ms.com, ms.eps, ms.sig)
    ew = ora.Ewald(systems.ALPHA / ms.box, systems.NK, systems.K_SQ_MAX, systems.FACTOR, ms.box)
    p = ora.potential_ewald(s, ew, 9.0, 9.0, ms.box)
    assert abs(4 * vec[0] - p.lj) < 1e-12 * abs(p.lj)                       # energy.jl:289 / :977
    assert abs(vec[2] * ew.factor - p.real) < 1e-12 * abs(p.real)
    S = vec[sharding.NSCAL::2] + 1j * vec[sharding.NSCAL + 1::2]
    want = ew.sum_new[:, 0] + 1j * ew.sum_new[:, 1]
    assert np.abs(S - want).max() < 1e-11
    e_recip = float((ew.cfac * (S.real ** 2 + S.imag ** 2)).sum()) * ew.factor
    assert abs(e_recip - p.recip) < 1e-11 * abs(p.recip)


def test_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (CPU only: the oracle port on a bounded sample) prints ONE JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    out = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--molecules", "4000", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "full_ewald_energy_evals_per_sec" and d["value"] > 0 and d["steps"] == 2
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
