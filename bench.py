#!/usr/bin/env python
"""bench.py — full-Ewald-energy evaluations/s (BASELINE.json metric) on the synthetic 256 000-molecule
SPC/E box (config E), sharded over N B200s, plus MC moves/s for configs A/B/C at N=1.

  python bench.py --gpus 1 --steps K --warmup W            (N>1: under torch.distributed.run)
  python bench.py --impl reference ...                      the reference algorithm on host cores

One "step" = one potential(…, "ewald") evaluation of the whole system (what every volume move
needs): molecule-pair LJ + real-space erfc Coulomb over unique pairs + ρ(k) rebuild + E_recip
+ E_self.  `value` times the device-resident path (state already in HBM); `e2e` times the
reference-facing call with HOST arrays (mmc_upload_system from the Julia-layout arrays +
mmc_potential + result on the host) — see DESIGN.md §Measurement.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "full_ewald_energy_evals_per_sec"
UNIT = "evals/s"
N_MOL_E = 256_000
RC = 10.0
FLOP_PER_PAIR = 624          # 9 x 67 + 21 per in-cutoff water-water pair
FLOP_PER_SITE_K = 14         # rho(k) rebuild, 14 flop per (site, k) + 168 per site + 6 per k
# From the committed `ncu --set full` capture of the same command line (profiles/r02_ncu_full_eval_kernels.txt): DRAM bytes of ONE
# launch (dram__bytes_read.sum + dram__bytes_write.sum) and the executed-instruction counters.  A profiler number is never taken
# inside bench.py; these are the capture's values, quoted so that the line carries them, and only for the configuration the
# capture was made on (config E, one GPU).  Algorithmic bytes: 43 MB of rows + gate coordinates read once (L2-resident).
NCU_CAPTURE = {
    "k_pairs_v7": {"dram_bytes": 44249856, "fp64_pipe_busy_pct": 49.4, "issue_slots_busy_pct": 61.4, "fp64_instructions_per_pair": 203,
                   "warp_instructions": 312.8e6, "source": "profiles/r02_ncu_full_eval_kernels.txt"},
    "k_rhok_pairs": {"dram_bytes": 24606976, "fp64_pipe_busy_pct": 54.8, "issue_slots_busy_pct": 39.6,
                     "source": "profiles/r02_ncu_full_eval_kernels.txt"},
}


def workload_config(n_mol):
    return {
        "workload": f"config E: synthetic SPC/E box, {n_mol} molecules ({3 * n_mol} sites) on InitCubicGrid at "
                    "rho=0.033101144 A^-3 with random quaternions (seed 11234); full Ewald potential(): "
                    "r_cut=10 A COM cutoff, kappa=5.6/L, nk=5, k^2<27 (337 k-vectors)",
        "n_molecules": n_mol,
        "l2": "flushed before every timed evaluation (256 MiB device write); state itself (35 MB) is L2-sized",
        "step": "one step = --evals-per-step evaluations (default 10), each timed with CUDA events on the launching stream",
        "sharding": "contiguous cost-balanced ranges of the cell grid (pair work) and rho(k) sites split per rank; the 682-double partial "
                    "vectors are exchanged peer to peer over NVLink (CUDA IPC buffers; the evaluation's tail kernel pushes its vector, waits "
                    "for the peers', adds them in rank order and finishes) or, with --collective nccl, by one all-reduce",
    }


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML (pynvml, ~1 ms period) when available, else
    nvidia-smi (the recipe's clocks line, ~0.2 s period)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self._stop = index, [], threading.Event()
        self.t = threading.Thread(target=self._run, daemon=True)
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _run_nvml(self):
        n = self.nvml
        bits = {"hw_slowdown": n.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": n.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": n.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": n.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop.is_set():
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                rs = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.rows.append([str(sm), str(self.max_sm)] + ["Active" if rs & bits[k] else "Not Active"
                                                                for k in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")])
            except Exception:
                pass
            self._stop.wait(0.001)

    def _run(self):
        if self.nvml is not None:
            return self._run_nvml()
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self.t.start()

    def stop(self):
        self._stop.set()
        self.t.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------- CPU baseline
def cpu_full_energy_sample(ms, n_rows, site_frac, threads):
    """Reference algorithm (oracle port) on a bounded sample: n_rows of the 2 x O(N^2) row loops of
    potential() (energy.jl:972-1001) and RecipLong on 1/site_frac of the sites; extrapolated linearly."""
    from oracle import oracle as ora
    from metropolismontecarlo_b200 import systems
    s = ora.System(ms.coords, ms.charge, ms.atype, ms.first_atom, ms.last_atom, ms.com, ms.eps, ms.sig)
    kappa = systems.ALPHA / ms.box
    rng = np.random.default_rng(1)
    n_rows = min(n_rows, ms.n_mol)
    i0 = int(rng.integers(0, ms.n_mol - n_rows)) if ms.n_mol > n_rows else 0
    t0 = time.perf_counter()
    ora.potential_rows(s, kappa, RC, RC, ms.box, i0, i0 + n_rows, threads)
    t_rows = time.perf_counter() - t0
    ns = ms.n_sites // site_frac
    ew = ora.Ewald(kappa, systems.NK, systems.K_SQ_MAX, systems.FACTOR, ms.box)
    t0 = time.perf_counter()
    ora.RecipLong(ew, ms.coords[:ns], ms.charge[:ns], ms.box)
    t_recip = time.perf_counter() - t0
    t_eval = t_rows * (ms.n_mol / n_rows) + t_recip * site_frac
    return 1.0 / t_eval, t_rows, t_recip


def run_reference(args, rank):
    """--impl reference: the reference's own CPU algorithm (oracle port; Julia is not installed, so
    there is no oracle/_ref) on all host cores, on a bounded sample of the same workload."""
    if rank != 0:
        return
    from oracle import oracle as ora
    from metropolismontecarlo_b200 import systems
    ora.build()
    threads = os.cpu_count() or 1
    ms = systems.spce_lattice(args.molecules)
    n_rows, frac = 4096, 16      # ≈ 4 s per step on the GPU box's 16 host cores: 20 + 3 steps end within two minutes
    for _ in range(args.warmup):
        cpu_full_energy_sample(ms, 256, 64, threads)
    vals = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        vals.append(cpu_full_energy_sample(ms, n_rows, frac, threads)[0])
    wall = time.perf_counter() - t0
    v = float(np.mean(vals))
    sample = (f"{n_rows} of {ms.n_mol} rows of the two O(N^2) loops ({threads} OpenMP threads over rows) + RecipLong on "
              f"1/{frac} of the sites (serial, as the reference), extrapolated linearly to one full evaluation")
    emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": dict(workload_config(ms.n_mol), evals_per_step=args.evals_per_step),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "sample_wall_s": wall,
    })


# ------------------------------------------------------------------------------------ moves/s (N=1)
def moves_benchmarks(n_moves=10_000):
    """Configs A, B, C: moves/s through the library's host driver (one fused launch + one host wait
    per move — the ccall protocol), and the oracle's Loop on one host core for a bounded sample."""
    from metropolismontecarlo_b200 import systems
    from metropolismontecarlo_b200.energy import Engine, LoopParams, water_engine
    from oracle import oracle as ora
    out = {}
    ms = systems.load_nist(4)
    u = np.random.default_rng(11234).random(8 * n_moves)
    for name, style, sid in (("A_spce750_ewald", "ewald", 0), ("B_spce750_wolf", "wolf", 1)):
        eng = water_engine(ms, RC)
        p0 = eng.potential(style)
        com, quat = ms.com.copy(), ms.quat.copy()
        eng.loop_run(LoopParams(298.15, 0.316555789, 0.05, 0.5, 1.0, sid, 1), com, quat, ms.db, u, 500, p0.energy, p0.virial)
        eng.upload_system(ms, RC, RC)
        p0 = eng.potential(style)
        com, quat = ms.com.copy(), ms.quat.copy()
        l0 = eng.counters().kernel_launches
        t0 = time.perf_counter()
        rc, acc, delta, st = eng.loop_run(LoopParams(298.15, 0.316555789, 0.05, 0.5, 1.0, sid, 1), com, quat, ms.db,
                                          u, n_moves, p0.energy, p0.virial)
        dt = time.perf_counter() - t0
        launches = eng.counters().kernel_launches - l0
        s = ora.System(ms.coords, ms.charge, ms.atype, ms.first_atom, ms.last_atom, ms.com, ms.eps, ms.sig)
        ew = ora.Ewald(systems.ALPHA / ms.box, 5, 27, systems.FACTOR, ms.box)
        if sid == 0:
            ora.RecipLong(ew, ms.coords, ms.charge, ms.box)
        n_cpu = 2000
        prm = ora.LoopParams(298.15, 0.316555789, 0.05, 0.5, 1.0, RC, RC, ms.box, sid, 1)
        t0 = time.perf_counter()
        ora.loop(s, ew, ms.db, ms.quat.copy(), prm, u, n_cpu, 0.0, 0.0)
        dtc = time.perf_counter() - t0
        # the same block of moves in ONE launch (mmc_loop_run_device: state in shared memory); one untimed block first
        # (first use loads the kernel and allocates the block's device buffers)
        eng.upload_system(ms, RC, RC)
        p0 = eng.potential(style)
        com, quat = ms.com.copy(), ms.quat.copy()
        eng.loop_run(LoopParams(298.15, 0.316555789, 0.05, 0.5, 1.0, sid, 1), com, quat, ms.db, u, 500, p0.energy, p0.virial, device=True)
        eng.upload_system(ms, RC, RC)
        p0 = eng.potential(style)
        com, quat = ms.com.copy(), ms.quat.copy()
        t0 = time.perf_counter()
        rc_d, acc_d, delta_d, st_d = eng.loop_run(LoopParams(298.15, 0.316555789, 0.05, 0.5, 1.0, sid, 1), com, quat, ms.db,
                                                   u, n_moves, p0.energy, p0.virial, device=True)
        dt_dev = time.perf_counter() - t0
        assert np.array_equal(acc, acc_d), "device block of moves diverged from the per-move protocol"
        out[name] = {"moves_per_s": n_moves / dt, "us_per_move": 1e6 * dt / n_moves, "accepted": int(st.n_accepted), "_acc_sum": int(acc_d.sum()),
                     "gpu_launches": int(launches), "cpu_port_moves_per_s_1core": n_cpu / dtc,
                     "flop_per_move": 2.1e5 if sid == 0 else 1.75e5,
                     "block_offload": {"moves_per_s": n_moves / dt_dev, "us_per_move": 1e6 * dt_dev / n_moves,
                                       "gpu_launches": 1, "what": "mmc_loop_run_device: 10^4 moves in one launch, "
                                       "uniforms H2D + results D2H inside the timed region, same accept/reject record"}}
        eng.close()
    # ---- config D: 4000 SPC/E, NPT with Ewald — a volume trial is one full pair + k-space recompute at the new box
    # (Ewald/volumeChange.jl:50-147: scale COMs, kappa = 5.6/L', cfac and rho(k) rebuilt); molecule moves in between go
    # through the per-move protocol (4000 molecules do not fit one SM's shared memory)
    ms_d = systems.spce_lattice(4000)
    eng = water_engine(ms_d, RC)
    p0 = eng.potential("ewald")
    rng = np.random.default_rng(11234)
    vol0 = ms_d.box ** 3
    boxes = [(vol0 + (rng.random() - 0.5) * 0.01 * vol0) ** (1.0 / 3.0) for _ in range(60)]   # vmax = 0.01 V
    for L in boxes[:10]:
        eng.volume_trial(L, systems.ALPHA / L, "ewald"); eng.volume_reject()
    l0 = eng.counters().kernel_launches
    t0 = time.perf_counter()
    for L in boxes[10:]:
        pv = eng.volume_trial(L, systems.ALPHA / L, "ewald")
        eng.volume_reject()
    dt_v = (time.perf_counter() - t0) / 50
    launches_v = (eng.counters().kernel_launches - l0) / 50
    info_d = eng.last_eval_info()
    com, quat = ms_d.com.copy(), ms_d.quat.copy()
    n_d = 2000
    t0 = time.perf_counter()
    rc, acc, delta, st = eng.loop_run(LoopParams(298.15, 0.316555789, 0.05, 0.5, 1.0, 0, 1), com, quat, ms_d.db, u, n_d,
                                      p0.energy, p0.virial)
    dt_m = (time.perf_counter() - t0) / n_d
    # block offload: one sweep (4000 moves) per launch; 4000 molecules are sliced over the cluster's 8 SMs
    eng.upload_system(ms_d, RC, RC)
    p0 = eng.potential("ewald")
    com, quat = ms_d.com.copy(), ms_d.quat.copy()
    eng.loop_run(LoopParams(298.15, 0.316555789, 0.05, 0.5, 1.0, 0, 1), com, quat, ms_d.db, u, 500, p0.energy, p0.virial, device=True)
    eng.upload_system(ms_d, RC, RC)
    p0 = eng.potential("ewald")
    com, quat = ms_d.com.copy(), ms_d.quat.copy()
    t0 = time.perf_counter()
    rc_b, acc_b, delta_b, st_b = eng.loop_run(LoopParams(298.15, 0.316555789, 0.05, 0.5, 1.0, 0, 1), com, quat, ms_d.db, u, n_moves,
                                              p0.energy, p0.virial, device=True)
    dt_b = (time.perf_counter() - t0) / n_moves
    assert np.array_equal(acc, acc_b[:n_d]), "device block of moves diverged from the per-move protocol"
    sweep = ms_d.n_mol * dt_m + dt_v                      # one NPT cycle: N molecule moves + one volume trial
    sweep_b = ms_d.n_mol * dt_b + dt_v
    out["D_spce4000_npt_ewald"] = {"volume_trial_evals_per_s": 1.0 / dt_v, "ms_per_volume_trial": 1e3 * dt_v,
                                   "gpu_launches_per_volume_trial": launches_v, "pair_kernel": info_d["pair_kernel"],
                                   "moves_per_s": 1.0 / dt_m, "us_per_move": 1e6 * dt_m,
                                   "blended_moves_per_s": (ms_d.n_mol + 1) / sweep,
                                   "block_offload": {"moves_per_s": 1.0 / dt_b, "us_per_move": 1e6 * dt_b,
                                                     "blended_moves_per_s": (ms_d.n_mol + 1) / sweep_b,
                                                     "what": "mmc_loop_run_device, state sliced over the 8 CTAs of the cluster"},
                                   "what": "volume trial = host call to result on host (scale + full Ewald energy at the new box, "
                                           "random box per trial, vmax = 0.01 V); blended = 4000 molecule moves + 1 volume trial"}
    eng.close()
    at = systems.lj_lattice(32_000, 0.75, 2.5)
    eng = Engine()
    eng.upload_atoms(at)
    p0 = eng.potential("atoms")
    r = at.r.copy()
    eng.loop_run_atoms(1.0, at.box / 30, r, u, 500, p0.energy, p0.virial)
    eng.upload_atoms(at)
    r = at.r.copy()
    t0 = time.perf_counter()
    rc, acc, delta, st = eng.loop_run_atoms(1.0, at.box / 30, r, u, n_moves, p0.energy, p0.virial)
    dt = time.perf_counter() - t0
    acc = acc.copy()
    n_cpu = 500
    t0 = time.perf_counter()
    ora.loop_atoms(at.r.copy(), at.eps, at.sig, at.box, at.r_cut, 1.0, at.box / 30, u, n_cpu, 0.0, 0.0)
    dtc = time.perf_counter() - t0
    eng.upload_atoms(at)
    r = at.r.copy()
    eng.loop_run_atoms(1.0, at.box / 30, r, u, 500, p0.energy, p0.virial, device=True)      # untimed first block
    eng.upload_atoms(at)
    r = at.r.copy()
    t0 = time.perf_counter()
    rc_d, acc_d, delta_d, st_d = eng.loop_run_atoms(1.0, at.box / 30, r, u, n_moves, p0.energy, p0.virial, device=True)
    dt_dev = time.perf_counter() - t0
    assert np.array_equal(acc, acc_d), "device block of moves diverged from the per-move protocol"
    out["C_lj32000"] = {"moves_per_s": n_moves / dt, "us_per_move": 1e6 * dt / n_moves, "accepted": int(st.n_accepted),
                        "cpu_port_moves_per_s_1core": n_cpu / dtc, "flop_per_move": 1.35e6,
                        "block_offload": {"moves_per_s": n_moves / dt_dev, "us_per_move": 1e6 * dt_dev / n_moves, "gpu_launches": 1,
                                          "what": "mmc_loop_run_atoms_device: 10^4 moves in one launch on an 8-SM cluster"}}
    eng.close()
    return out


# ------------------------------------------------------------------------------------------- ours
def replica_moves(local_rank, world, dev, n_moves=10_000):
    """Per-move paths do not shard (a Markov chain is sequential, SURVEY §8e: "replicas only"): at N GPUs the moves/s figure is
    N independent replicas of config A, one per GPU, each running its block of moves in one launch; aggregate = moves of all
    replicas / max time over ranks.  Every replica gets its own uniform stream."""
    import torch
    import torch.distributed as dist
    from metropolismontecarlo_b200 import systems
    from metropolismontecarlo_b200.energy import LoopParams, water_engine
    ms = systems.load_nist(4)
    eng = water_engine(ms, RC, device=local_rank)
    u = np.random.default_rng(11234 + int(os.environ.get("RANK", "0"))).random(8 * n_moves)
    prm = LoopParams(298.15, 0.316555789, 0.05, 0.5, 1.0, 0, 1)
    best = None
    for rep in range(3):      # the first block is untimed (kernel load, buffer allocation)
        eng.upload_system(ms, RC, RC)
        p0 = eng.potential("ewald")
        com, quat = ms.com.copy(), ms.quat.copy()
        dist.barrier()
        t0 = time.perf_counter()
        rc, acc, delta, st = eng.loop_run(prm, com, quat, ms.db, u, n_moves, p0.energy, p0.virial, device=True)
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rep > 0:
            best = float(t.item()) if best is None else min(best, float(t.item()))
    fresh = eng.potential("ewald")
    assert abs(st.total_energy - fresh.energy) <= 1e-9 * abs(fresh.energy)
    eng.close()
    return {"moves_per_s": world * n_moves / best, "replicas": world, "us_per_move_per_replica": 1e6 * best / n_moves,
            "what": "config A (750 SPC/E, Ewald), mmc_loop_run_device, one independent replica per GPU; "
                    "uniforms H2D + results D2H inside the timed region, max over ranks"}


def e2_large_k(ms, fp64_peak):
    """SURVEY 8d "E2" (labelled extension): config E with a converged k-space sum — kappa = 0.32 /A (kappa r_cut = 3.2), nk = 12,
    k^2 < 145: 3796 k-vectors instead of the reference's 337 — through k_rhok_big."""
    from metropolismontecarlo_b200.energy import Engine
    eng = Engine()
    eng.upload_system(ms, RC, RC)
    nkv = eng.PrepareEwaldVariables(0.32, 12, 145)
    eng.set_timing(True)
    eng.debug_set("overlap_rhok", 0)
    rows = []
    for _ in range(8):
        t0 = time.perf_counter()
        p = eng.potential("ewald")
        w = time.perf_counter() - t0
        t = eng.last_timings()
        rows.append((t["pairs_ms"], t["rhok_ms"], 1e3 * w))
    r = np.median(np.array(rows[2:]), axis=0)
    flops = (FLOP_PER_SITE_K * nkv + 168) * ms.n_sites + 6 * nkv
    eng.close()
    return {"what": "extension, not a reference configuration: kappa = 0.32 /A (kappa r_cut = 3.2), nk = 12, k^2 < 145", "k_vectors": nkv,
            "evals_per_s": 1e3 / r[2], "ms_per_eval_wall": float(r[2]), "kernel_ms": {"pairs": float(r[0]), "rhok_rebuild": float(r[1])},
            "rhok_kernel": "k_rhok_big", "rhok_algorithmic_tflops": flops / (r[1] * 1e-3) / 1e12,
            "rhok_frac_of_fp64_peak": (flops / (r[1] * 1e-3) / 1e12 / fp64_peak) if fp64_peak else None,
            "rhok_note": "algorithmic = the reference's 14 flop per (site, k) (SURVEY 8d); the kernel forms e^{i(kx x +- ky y)} once per (kx, |ky|) "
                         "and gets the four sign combinations of (ky, kz) from 8 DFMA, i.e. it executes ~4 flop per (site, k), so the algorithmic "
                         "rate may exceed the DFMA peak; executed: FP64 pipe 63 % busy (profiles/r02_ncu_rhok_big.txt)",
            "energy_per_molecule_K": p.energy / ms.n_mol}


def replicas_one_gpu(n_rep=15, n_moves=10_000, cluster=8):
    """Per-move paths do not shard, and one chain uses 8 of the 148 SMs (k_chains: one 8-CTA cluster).  So ONE GPU runs R
    independent replicas of config A at once: R handles, R streams, R host threads (what R Julia processes sharing a GPU would
    do); aggregate moves/s = R x n_moves / wall time of the slowest.  Replica 0 runs the stream of the single-chain leg and must
    give the same accept/reject record."""
    import threading
    from metropolismontecarlo_b200 import systems
    from metropolismontecarlo_b200.energy import LoopParams, water_engine
    ms = systems.load_nist(4)
    prm = LoopParams(298.15, 0.316555789, 0.05, 0.5, 1.0, 0, 1)
    engs = [water_engine(ms, RC) for _ in range(n_rep)]
    for e in engs:
        e.debug_set("chain_cluster", cluster)
    us = [np.random.default_rng(11234).random(8 * n_moves) if r == 0 else np.random.default_rng(11234 + r).random(8 * n_moves) for r in range(n_rep)]
    out = [None] * n_rep
    walls = []
    for rep in range(3):                       # first round untimed (kernel load, buffer allocation)
        p0s = []
        for e in engs:
            e.upload_system(ms, RC, RC)
            p0s.append(e.potential("ewald"))
        start = threading.Barrier(n_rep + 1)

        def run(r):
            com, quat = ms.com.copy(), ms.quat.copy()
            start.wait()
            out[r] = engs[r].loop_run(prm, com, quat, ms.db, us[r], n_moves, p0s[r].energy, p0s[r].virial, device=True)
        th = [threading.Thread(target=run, args=(r,)) for r in range(n_rep)]
        for t in th:
            t.start()
        start.wait()
        t0 = time.perf_counter()
        for t in th:
            t.join()
        if rep > 0:
            walls.append(time.perf_counter() - t0)
    for r in range(n_rep):
        fresh = engs[r].potential("ewald")
        assert out[r][0] == 0 and abs(out[r][3].total_energy - fresh.energy) <= 1e-9 * abs(fresh.energy)
    acc0 = out[0][1].copy()
    for e in engs:
        e.close()
    best = min(walls)
    return {"moves_per_s": n_rep * n_moves / best, "replicas": n_rep, "us_per_move_per_replica": 1e6 * best / n_moves,
            "sms_per_chain": cluster,
            "what": f"config A, {n_rep} independent chains on ONE GPU (k_chains: a {cluster}-SM cluster each), one host thread and one "
                    "stream per chain; uniforms H2D + results D2H inside the timed region"}, acc0


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from metropolismontecarlo_b200 import systems
    from metropolismontecarlo_b200.energy import Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libmmc_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # One explicit (non-default) stream for everything: the engine's kernels, torch's NCCL
    # all-reduce and the timing events are all ordered on it.  (A NULL stream in mmc_config means
    # "the handle's own stream", so torch's legacy default stream cannot be shared.)
    stream = torch.cuda.Stream(device=dev, priority=-1)   # high priority: the library's side stream (rho(k) rebuild) runs below it
    torch.cuda.set_stream(stream)
    eng = Engine(device=local_rank, rank=rank, world=world, stream=stream.cuda_stream)
    ms = systems.spce_lattice(args.molecules)
    eng.upload_system(ms, RC, RC)
    eng.PrepareEwaldVariables(systems.ALPHA / ms.box)
    if args.overlap_rhok >= 0:
        eng.debug_set("overlap_rhok", args.overlap_rhok)
    for kv in args.debug:                           # library tuning switches for experiments: --debug key=value
        k_, v_ = kv.split("=")
        eng.debug_set(k_, int(v_))
    nvec = eng.partial_count()
    vec = torch.zeros(nvec, dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    from metropolismontecarlo_b200.sharding import setup_peer_exchange, sharded_potential

    p2p = world > 1 and args.collective == "p2p"
    if p2p:        # every rank maps every rank's exchange buffer (CUDA IPC); NCCL only carries the 64-byte handles
        ok = 1
        try:
            setup_peer_exchange(eng, world)
        except Exception as e:            # no CUDA IPC between these processes: all ranks fall back to the NCCL all-reduce
            print(f"rank {rank}: peer exchange unavailable ({e}); using the NCCL all-reduce", file=sys.stderr)
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        p2p = bool(flag.item())
        if not p2p:
            args.collective = "nccl"

    def step():
        if world == 1:
            return eng.potential("ewald")
        if p2p:    # partial -> the tail kernel stores the 682 doubles into every peer's buffer over NVLink -> ordered sum -> Properties
            return eng.potential_sharded("ewald")
        # partial -> NCCL all-reduce over NVLink (8 scalars + 337 complex rho(k)) -> finalize
        return sharded_potential(eng, "ewald", vec, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        props = step()
    fp64_peak, peak_clock = (0.0, None)
    if rank == 0:
        cs = ClockSampler(local_rank)
        cs.start()
        fp64_peak = eng.measure_fp64_peak()
        peak_clock = cs.stop()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    # inside the timed region only the pair kernel is bracketed by the library's events (two records, no synchronisation); the
    # per-phase instrumentation costs ~35 us per evaluation and runs in a separate pass afterwards
    eng.set_timing(2)
    inner = max(1, args.evals_per_step)
    n_ev = args.steps * inner
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(n_ev)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(n_ev)]
    l0 = eng.counters().kernel_launches
    pair_ms, rhok_ms = [], []
    barrier()
    if sampler:
        sampler.start()
    for k in range(n_ev):
        flush.fill_(k & 0xff)                          # evict the 126 MB L2 before every timed evaluation
        ev0[k].record()
        props = step()
        ev1[k].record()
        pair_ms.append(eng.last_timings()["pairs_ms"])
    barrier()
    clocks = sampler.stop() if sampler else None
    launches = eng.counters().kernel_launches - l0
    eng.set_timing(True)                               # every phase, same placement of the rebuild, outside the timed region
    for k in range(7):
        flush.fill_(k)
        step()
        rhok_ms.append(eng.last_timings()["rhok_ms"])
    barrier()
    # the two big kernels alone (rho(k) rebuild NOT running beside the pair kernel), outside the timed region: explains `roofline`
    eng.debug_set("overlap_rhok", 0)
    iso, iso_r = [], []
    for _ in range(7):
        step()
        iso.append(eng.last_timings()["pairs_ms"])
        iso_r.append(eng.last_timings()["rhok_ms"])
    eng.debug_set("overlap_rhok", args.overlap_rhok if args.overlap_rhok >= 0 else 1)
    pair_ms_isolated, rhok_ms_isolated = float(np.median(iso)), float(np.median(iso_r))
    eng.set_timing(False)
    total_ms = sum(a.elapsed_time(b) for a, b in zip(ev0, ev1))
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)       # max over ranks, device-timed
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = 1e3 * inner / ms_per_step
    info = eng.last_eval_info()
    # sharded == unsharded, outside the timed region: every rank evaluates the whole (replicated) system itself once
    sharded_rel = None
    if world > 1:
        whole = eng.potential("ewald")
        sharded_rel = max(abs(getattr(props, f) - getattr(whole, f)) / abs(getattr(whole, f)) for f in ("energy", "virial", "lj", "real", "recip"))
        assert sharded_rel < 1e-11, f"sharded evaluation differs from the unsharded one: {sharded_rel}"

    # ---- e2e: host arrays in, host scalars out, every step.  The inputs are the reference's own array layouts (pointer(soa.coords),
    # pointer(moa.COM)) in PINNED host memory; mmc_potential_host DMAs them as they are.  One rank: copies overlapped with binning and
    # the rho(k) rebuild.  N ranks: domain decomposition — a rank copies all COMs but only the site blocks of its slab of the cell grid.
    ms_pin = ms.copy()
    for name in ("coords", "com"):
        setattr(ms_pin, name, torch.from_numpy(np.ascontiguousarray(getattr(ms, name))).pin_memory().numpy())
    # (warm-up: W steps of evals_per_step calls like the device-timed leg — the first few dozen copies out of a freshly pinned
    #  buffer run at 60-70 % of the link's rate on some boxes, tools/prof_e2e_ab.py)
    for _ in range(max(args.warmup, 3) * inner):
        p2 = eng.potential_host(ms_pin.coords, ms_pin.com, "ewald")
    h2d = eng.last_host_bytes()
    barrier()
    t0 = time.perf_counter()
    n_e2e = max(10, min(n_ev, 50))
    for _ in range(n_e2e):
        p2 = eng.potential_host(ms_pin.coords, ms_pin.com, "ewald")
    barrier()
    e2e_s = (time.perf_counter() - t0) / n_e2e
    t = torch.tensor([e2e_s, float(h2d)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s, h2d_max = float(t[0].item()), int(t[1].item())
    assert abs(p2.energy - props.energy) <= 1e-12 * abs(props.energy)
    eng.upload_positions(ms.coords, ms.com)            # (a sharded handle holds only its slab after potential_host)

    replicas = replica_moves(local_rank, world, dev) if (world > 1 and not args.no_moves) else None
    if rank == 0:
        pairs = info["pairs_in_cutoff"]
        t_pair = float(np.mean(pair_ms)) * 1e-3
        # algorithmic FP64 flops of the dominant kernel for THIS rank's share of the pairs
        alg_flops = FLOP_PER_PAIR * pairs / world
        achieved = alg_flops / t_pair / 1e12
        nk = eng.nkvecs
        rhok_flops = (FLOP_PER_SITE_K * nk + 168) * ms.n_sites / world + 6 * nk
        cap = NCU_CAPTURE.get(info["pair_kernel"]) if (world == 1 and ms.n_mol == N_MOL_E) else None
        cap_r = NCU_CAPTURE.get("k_rhok_pairs") if (world == 1 and ms.n_mol == N_MOL_E) else None
        peak_src = ("live DFMA-chain probe on this GPU before the timed region (MEASURED_PEAKS.json has no FP64 figure); nominal "
                    "148 SM x 64 lanes x 2 x 1.965 GHz = 37.2 TFLOP/s")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(ms.n_mol), evals_per_step=inner),
            "ms_per_eval": ms_per_step / inner,
            "energy_per_molecule_K": props.energy / ms.n_mol,
            "sharded_vs_unsharded_rel": sharded_rel,
            "pairs_in_cutoff": pairs, "path": info["mode"], "cells_per_dim": info["cells_per_dim"],
            "kernel_ms": {"pairs": float(np.mean(pair_ms)), "rhok_rebuild": float(np.mean(rhok_ms))},
            "roofline": {"bound": "fp64", "kernel": info["pair_kernel"], "achieved": achieved, "peak": fp64_peak,
                         "unit": "TFLOP/s", "frac": achieved / fp64_peak if fp64_peak else None,
                         "peak_source": peak_src, "peak_probe_clocks": peak_clock,
                         "algorithmic_flop_per_launch": alg_flops,
                         "note": "achieved/frac are in situ (timed region; the rho(k) rebuild is queued behind the pair kernel on a low-priority stream "
                                 "and fills the SMs as the ticket queue drains)",
                         "isolated": {"ms": pair_ms_isolated, "achieved": alg_flops / (pair_ms_isolated * 1e-3) / 1e12,
                                      "frac": (alg_flops / (pair_ms_isolated * 1e-3) / 1e12 / fp64_peak) if fp64_peak else None},
                         "executed": cap,
                         "traffic": cap["dram_bytes"] if cap else None,
                         "traffic_unit": "bytes of DRAM per launch (ncu capture named in `executed.source`; null when this run is not the captured configuration)"},
            "roofline_rhok": {"bound": "fp64", "kernel": "k_rhok_pairs",
                              "achieved": rhok_flops / (rhok_ms_isolated * 1e-3) / 1e12 if rhok_ms_isolated > 0 else None,
                              "peak": fp64_peak, "unit": "TFLOP/s",
                              "frac": (rhok_flops / (rhok_ms_isolated * 1e-3) / 1e12 / fp64_peak) if (fp64_peak and rhok_ms_isolated > 0) else None,
                              "algorithmic_flop_per_launch": rhok_flops,
                              "note": "kernel timed alone (rho(k) rebuild before the pair path on the same stream, outside the timed region): in the timed "
                                      "region it is queued behind the pair kernel on a low-priority stream and fills the SMs as the pair kernel's ticket "
                                      "queue drains, so its in-situ event time (kernel_ms.rhok_rebuild) contains the wait for the pair kernel",
                              "in_situ_ms": float(np.mean(rhok_ms)),
                              "executed": cap_r, "traffic": cap_r["dram_bytes"] if cap_r else None},
            "e2e": {"value": 1.0 / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d_max, "d2h_bytes_per_step": 72,
                    "evaluations_timed": n_e2e,
                    "what": "mmc_potential_host: pinned host soa.coords + moa.COM in, Properties out, wall clock, max over ranks.  N=1: all "
                            "24.6 MB, copies overlapped with binning and the rho(k) rebuild.  N>1: domain decomposition, every rank copies all "
                            "COMs and only the site blocks of its slab of the cell grid (h2d_bytes_per_step = the largest rank's bytes)"},
            "gpu_launches": int(launches) * world,
            "collective": (args.collective if world > 1 else None),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            from oracle import oracle as ora
            ora.build()
            threads = os.cpu_count() or 1
            v, t_rows, t_recip = cpu_full_energy_sample(ms, 4096, 8, threads)      # ≈ 20 core-seconds
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": f"4096 of {ms.n_mol} rows of the reference's two O(N^2) loops ({threads} OpenMP threads, "
                          f"{t_rows:.1f} s) + RecipLong on 1/8 of the sites (serial, {t_recip:.1f} s), extrapolated "
                          "linearly; Julia is not installed, so this is the C restatement of the reference algorithm"}
            if not args.no_moves:
                line["moves"] = moves_benchmarks()
                line["moves"]["B_spce750_wolf"].pop("_acc_sum", None)
                try:
                    rep, acc0 = replicas_one_gpu(15, 10_000, 8)
                    rep["same_record_as_single_chain"] = bool(line["moves"]["A_spce750_ewald"].pop("_acc_sum") == int(acc0.sum()))
                    line["moves"]["A_spce750_ewald"]["replicas_one_gpu"] = rep
                    line["moves"]["A_spce750_ewald"]["replicas_one_gpu_2sm"] = replicas_one_gpu(64, 10_000, 2)[0]
                except Exception as e:           # (never lose the line to the optional leg)
                    line["moves"]["A_spce750_ewald"]["replicas_one_gpu"] = {"error": str(e)}
                line["moves"]["A_spce750_ewald"].pop("_acc_sum", None)
        if replicas is not None:
            line["moves"] = {"A_spce750_ewald_replicas": replicas}
        if world == 1 and ms.n_mol == N_MOL_E and not args.no_moves:
            try:
                line["e2_large_k"] = e2_large_k(ms, fp64_peak)
            except Exception as e:
                line["e2_large_k"] = {"error": str(e)}
        emit(line)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    # exactly ONE line on stdout (the JSON): libraries that print banners there (NCCL's version line under
    # NCCL_DEBUG=VERSION) are sent to stderr while the benchmark runs
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        _main()
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        if _RESULT_LINES:
            os.write(1, ("\n".join(_RESULT_LINES) + "\n").encode())


_RESULT_LINES = []


def emit(line: dict):
    _RESULT_LINES.append(json.dumps(line))


def _main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--molecules", type=int, default=N_MOL_E)
    ap.add_argument("--evals-per-step", dest="evals_per_step", type=int, default=10,
                    help="evaluations per step (each preceded by an L2 flush and timed with its own pair of CUDA events)")
    ap.add_argument("--collective", default="p2p", choices=["p2p", "nccl"],
                    help="exchange of the partial sums at N > 1: NVLink peer-memory kernels (default) or an NCCL all-reduce")
    ap.add_argument("--overlap-rhok", dest="overlap_rhok", type=int, default=-1, help="placement of the rho(k) rebuild (library default when < 0)")
    ap.add_argument("--debug", action="append", default=[], help="mmc_debug_set key=value (experiments)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-moves", action="store_true", help="skip the moves/s legs (configs A/B/C)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("for --gpus N > 1 launch under torch.distributed.run (one rank per GPU)")
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
