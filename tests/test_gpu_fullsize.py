"""Config E at BASELINE.json's full size (256 000 SPC/E molecules, 768 000 sites, L = 197.757 Å) on one B200.

The oracle's potential() is O(N²) (about a minute on 16 host cores), so the full-size checks are the size-independent
properties of the path plus every oracle piece that is O(N) or O(n_s · NK):
  * per-molecule rows: mmc_energy_all against the oracle's LJ_poly_ΔU(i) / EwaldShort(i) for sampled i (each is one
    O(N) scan, energy.jl:209-290, ewalds.jl:293-376), and Σ_i rows / 2 == potential()'s LJ and real-space terms
    (energy.jl:966-1001) — together these pin the pair part of the full evaluation to the oracle;
  * RecipLong and EwaldSelf against the oracle directly (ewalds.jl:538-604, 829-833);
  * the three pair kernels (k_pairs_v7, k_pairs_fast, k_pairs) agree; sharded partial sums over 8 emulated ranks == unsharded;
  * a volume trial at f = 1 reproduces potential(); ρ(k) after delta updates == a fresh rebuild.
Tolerance 1e-10 relative (north star), stated per assertion.
"""
import numpy as np
import pytest

from metropolismontecarlo_b200 import systems
from oracle import oracle as ora
from tests.util import ora_ewald, ora_system, rel  # noqa: F401

pytestmark = pytest.mark.gpu

N_E = 256000


@pytest.fixture(scope="module")
def cfg_e():
    from metropolismontecarlo_b200.energy import water_engine
    ms = systems.spce_lattice(N_E)
    eng = water_engine(ms, 10.0)
    yield ms, eng
    eng.close()


def test_full_size_rows_and_totals_against_oracle(cfg_e):
    ms, eng = cfg_e
    assert ms.n_sites == 768000 and abs(ms.box - 197.757) < 1e-2
    p = eng.potential("ewald")
    info = eng.last_eval_info()
    assert info["mode"] == "cells" and info["cells_per_dim"] == 19 and info["pair_kernel"] == "k_pairs_v7"
    lj, vir, qq, ov = eng.energy_all("ewald")
    assert not ov.any() and p.overlaps == 0
    assert rel(lj.sum() / 2, p.lj) < 1e-10 and rel(qq.sum() / 2, p.real) < 1e-10
    s = ora_system(ms)
    ew = ora_ewald(ms.box)
    rng = np.random.default_rng(11234)
    sample = np.unique(np.concatenate([[1, 2, N_E - 1, N_E], rng.integers(1, N_E + 1, 60)]))
    for i in sample:
        e0, v0 = ora.LJ_poly_dU(int(i), s, 10.0, ms.box)
        c0, _, ov0 = ora.EwaldShort(int(i), s, ew, 10.0, ms.box)
        assert not ov0
        assert rel(lj[i - 1], e0) < 1e-10 and rel(qq[i - 1], c0) < 1e-10, i
        assert abs(vir[i - 1] - v0) < 1e-10 * max(1.0, abs(v0), abs(e0)), i
    # k-space and self term: the oracle itself at full size
    e_recip = ora.RecipLong(ew, ms.coords, ms.charge, ms.box) * systems.FACTOR
    assert rel(p.recip, e_recip) < 1e-10
    assert rel(p.self_, ora.EwaldSelf(ew, ms.charge)) < 1e-10     # Σq² over 768 000 sites: serial (oracle) vs tree order
    assert rel(p.energy, p.lj + p.real + p.recip + p.self_) < 1e-14
    assert rel(p.virial, vir.sum() / 2 + (p.real + p.recip + p.self_) / 3) < 1e-10       # energy.jl:1019-1021
    old, new = eng.rhok()
    assert np.array_equal(old, new)                       # RecipLong writes both buffers (ewalds.jl:598-599)
    assert abs(old - (ew.sum_new[:, 0] + 1j * ew.sum_new[:, 1])).max() < 1e-12 * 0.8476 * ms.n_sites


def test_full_size_pair_kernels_agree(cfg_e):
    ms, eng = cfg_e
    ref = None
    try:
        for level in (0, 1, 2):
            eng.debug_set("pair_level", level)
            p = eng.potential("ewald")
            assert eng.last_eval_info()["pair_kernel"] == ("k_pairs_v7", "k_pairs_fast<64>", "k_pairs")[level]
            if ref is None:
                ref = p
                assert eng.last_eval_info()["pairs_in_cutoff"] > 17_000_000
            else:
                assert rel(p.lj, ref.lj) < 1e-11 and rel(p.real, ref.real) < 1e-11 and rel(p.virial, ref.virial) < 1e-11, level
                # the v7 tail folds the rho(k) CTA partials in (slice, sub-slice) order, the legacy k_rhok_reduce in CTA order
                assert rel(p.recip, ref.recip) < 1e-13
    finally:
        eng.debug_set("pair_level", 0)


def test_full_size_sharded_sum_equals_unsharded(cfg_e):
    import torch
    from metropolismontecarlo_b200.energy import water_engine
    ms, eng = cfg_e
    ref = eng.potential("ewald")
    world = 8
    engs = [water_engine(ms, 10.0, rank=r, world=world) for r in range(world)]
    n = engs[0].partial_count()
    bufs = [torch.zeros(n, dtype=torch.float64, device="cuda") for _ in range(world)]
    for e, b in zip(engs, bufs):
        e.potential_partial("ewald", b.data_ptr())
    torch.cuda.synchronize()
    total = torch.stack(bufs).sum(0)
    got = engs[3].potential_finalize("ewald", total.clone().data_ptr())
    for f in ("energy", "virial", "coulomb", "lj", "real", "recip", "self_"):
        assert rel(getattr(got, f), getattr(ref, f)) < 1e-11, f
    for e in engs:
        e.close()


def test_full_size_volume_identity_and_rhok_delta(cfg_e):
    ms, eng = cfg_e
    ref = eng.potential("ewald")
    v = eng.volume_trial(ms.box, systems.ALPHA / ms.box, "ewald")
    for f in ("energy", "virial", "coulomb", "lj", "real", "recip"):
        assert rel(getattr(v, f), getattr(ref, f)) < 1e-12, f
    eng.volume_reject()
    # 20 accepted single-molecule moves through the delta update, then a fresh rebuild
    rng = np.random.default_rng(5)
    coords, com = ms.coords.copy(), ms.com.copy()
    for k in range(20):
        i = int(rng.integers(1, N_E + 1))
        d = rng.uniform(-0.3, 0.3, 3)
        a = 3 * (i - 1)
        t = eng.trial_move(i, com[i - 1] + d, coords[a:a + 3] + d, "ewald")
        assert not t.overlap_new
        eng.accept()
        coords[a:a + 3] += d
        com[i - 1] += d
    old_delta, _ = eng.rhok()
    eng.RecipLong()
    fresh, _ = eng.rhok()
    assert abs(old_delta - fresh).max() < 1e-12 * 0.8476 * ms.n_sites
    eng.upload_system(ms, 10.0, 10.0)


def test_full_size_potential_host(cfg_e):
    """mmc_potential_host at full size: COMs first (binning), sites in chunks with the rho(k) partials of each chunk computed as
    it lands, the home cells cut into windows that are gathered and evaluated while later chunks are still on the bus.  Same
    Properties as upload + potential() for the lattice order (windows really start early) and for a random molecule order (every
    window needs the last chunk), with 1..4 windows and 1..8 chunks."""
    from metropolismontecarlo_b200.energy import water_engine
    ms, eng = cfg_e
    ref = eng.potential("ewald")
    fields = ("energy", "virial", "coulomb", "lj", "real", "recip", "self_")
    for windows, chunks in ((4, 8), (1, 4), (4, 4), (3, 5), (2, 2), (4, 1)):
        eng.debug_set("host_windows", windows)
        eng.debug_set("host_chunks", chunks)
        for style in ("ewald", "wolf"):
            got = eng.potential_host(ms.coords, ms.com, style)
            want = ref if style == "ewald" else eng.potential("wolf")
            for f in fields:
                assert rel(getattr(got, f), getattr(want, f)) < 1e-12, (windows, chunks, style, f)
            assert got.overlaps == 0
    eng.debug_set("host_windows", 3)
    eng.debug_set("host_chunks", 6)
    # a random molecule order: same energy (to summation order)
    perm = np.random.default_rng(3).permutation(N_E)
    mp = ms.copy()
    mp.com = ms.com[perm].copy()
    mp.coords = ms.coords.reshape(N_E, 3, 3)[perm].reshape(-1, 3).copy()
    e2 = water_engine(mp, 10.0)
    got = e2.potential_host(mp.coords, mp.com, "ewald")
    for f in fields:
        assert rel(getattr(got, f), getattr(ref, f)) < 1e-11, f
    want = e2.potential("ewald")
    for f in fields:
        assert rel(getattr(got, f), getattr(want, f)) < 1e-12, f
    e2.close()


def test_config_e_against_full_oracle_golden(cfg_e):
    """mmc_potential's Properties at config E against ONE full run of the oracle's O(N²) potential() on the same inputs
    (tests/golden/config_e_properties.json, written by tests/golden/make_config_e.py).  1e-10 relative."""
    import json
    from pathlib import Path
    g = json.loads((Path(__file__).resolve().parent / "golden" / "config_e_properties.json").read_text())
    ms, eng = cfg_e
    assert g["n_molecules"] == ms.n_mol and g["box"] == ms.box and g["kappa"] == systems.ALPHA / ms.box
    p = eng.potential("ewald")
    assert eng.last_eval_info()["pair_kernel"] == "k_pairs_v7"
    for f in ("energy", "virial", "coulomb", "lj", "real", "recip", "self_"):
        assert rel(getattr(p, f), g["ewald"][f]) < 1e-10, (f, getattr(p, f), g["ewald"][f])
    assert p.overlaps == g["ewald"]["overlaps"] == 0
    w = eng.potential("wolf")
    for f in ("energy", "virial", "coulomb", "lj", "real", "wolf_const"):
        assert rel(getattr(w, f), g["wolf"][f]) < 1e-10, (f, getattr(w, f), g["wolf"][f])
    for level in (1, 2):                              # the general kernels against the same golden
        eng.debug_set("pair_level", level)
        q = eng.potential("ewald")
        for f in ("energy", "virial", "lj", "real"):
            assert rel(getattr(q, f), g["ewald"][f]) < 1e-10, (level, f)
    eng.debug_set("pair_level", 0)


def test_config_c_32000_atoms_against_oracle():
    """Config C at its stated size (BASELINE.json: monatomic LJ, 32 000 atoms, rho* = 0.75, r_cut = 2.5, T* = 1, dr_max = L/30;
    Monatomic/mainMonatomic.jl:227-289, 373-413): LJ_ΔU rows for sampled atoms, potential(), and a 10⁴-move trajectory through
    the per-move protocol AND the block-of-moves kernel against the oracle's loop on the reference's own random stream
    (identical accept/reject record, deltas, final positions)."""
    from metropolismontecarlo_b200.energy import Engine, julia_rand
    at = systems.lj_lattice(32000, 0.75, 2.5)
    assert at.n == 32000 and abs(at.box - 34.9432) < 1e-3
    rng = np.random.default_rng(11234)
    at.r[:] = (at.r + rng.normal(0, 0.05, at.r.shape)) % at.box          # off the perfect lattice: a generic configuration
    eng = Engine()
    eng.upload_atoms(at)
    sample = np.unique(np.concatenate([[1, 2, 31999, 32000], rng.integers(1, 32001, 64)]))
    for i in sample:
        e, v = eng.LJ_ΔU(int(i))
        e0, v0 = ora.LJ_dU_atom(int(i), at.r, at.eps, at.sig, at.box, at.r_cut)
        assert rel(e, e0) < 1e-12 and rel(v, v0) < 1e-12, i
    p = eng.potential("atoms")
    e0, v0 = ora.potential_atoms(at.r, at.eps, at.sig, at.box, at.r_cut, 16)
    assert rel(p.energy, e0) < 1e-11 and rel(p.virial, v0) < 1e-11
    n_moves = 10_000
    u = julia_rand(11234, 5 * n_moves)
    r_o = at.r.copy()
    rc_o, acc_o, del_o, st_o = ora.loop_atoms(r_o, at.eps, at.sig, at.box, at.r_cut, 1.0, at.box / 30, u, n_moves, e0, v0)
    assert rc_o == 0 and 0 < st_o.n_accepted < n_moves
    for device in (False, True):
        eng.upload_atoms(at)
        r_g = at.r.copy()
        rc_g, acc_g, del_g, st_g = eng.loop_run_atoms(1.0, at.box / 30, r_g, u, n_moves, p.energy, p.virial, device=device)
        assert rc_g == 0 and np.array_equal(acc_g, acc_o), device
        assert st_g.uniforms_used == st_o.uniforms_used and st_g.n_accepted == st_o.n_accepted
        assert np.abs(del_g - del_o).max() < 1e-10 * max(1.0, np.abs(del_o).max())
        assert np.abs(r_g - r_o).max() < 1e-12 and np.abs(eng.download_atoms() - r_o).max() < 1e-12
        assert rel(st_g.total_energy, st_o.total_energy) < 1e-10
    eng.close()


def test_mixed_topology_at_full_size():
    """A 256 000-molecule water + ion mixture (2000 single-site ions, three LJ types; 764 000 sites) through the cell-list
    pair kernel on the padded evaluation copy: totals = Σ_i rows / 2, sampled rows (ions and waters) against the oracle's
    O(N) per-molecule functions, RecipLong and EwaldSelf against the oracle, and a volume trial at f = 1 == potential()."""
    from metropolismontecarlo_b200.energy import water_engine
    ms = systems.water_ion_mixture(N_E, 2000)
    assert ms.n_sites == 3 * (N_E - 2000) + 2000
    eng = water_engine(ms, 10.0)
    p = eng.potential("ewald")
    info = eng.last_eval_info()
    assert info["mode"] == "cells" and info["pair_kernel"] == "k_pairs" and p.overlaps == 0
    lj, vir, qq, ov = eng.energy_all("ewald")
    assert not ov.any()
    assert rel(lj.sum() / 2, p.lj) < 1e-10 and rel(qq.sum() / 2, p.real) < 1e-10
    s = ora_system(ms)
    ew = ora_ewald(ms.box)
    ions = np.flatnonzero(ms.last_atom == ms.first_atom) + 1
    assert len(ions) == 2000
    rng = np.random.default_rng(5)
    sample = np.unique(np.concatenate([[1, N_E], ions[:6], ions[-3:], rng.integers(1, N_E + 1, 20)]))
    for i in sample:
        e0, v0 = ora.LJ_poly_dU(int(i), s, 10.0, ms.box)
        c0, _, ov0 = ora.EwaldShort(int(i), s, ew, 10.0, ms.box)
        assert not ov0
        assert rel(lj[i - 1], e0) < 1e-10 and rel(qq[i - 1], c0) < 1e-10, i
        assert abs(vir[i - 1] - v0) < 1e-10 * max(1.0, abs(v0), abs(e0)), i
    assert rel(p.recip, ora.RecipLong(ew, ms.coords, ms.charge, ms.box) * systems.FACTOR) < 1e-10
    assert rel(p.self_, ora.EwaldSelf(ew, ms.charge)) < 1e-10
    v = eng.volume_trial(ms.box, systems.ALPHA / ms.box, "ewald")
    eng.volume_reject()
    for f in ("energy", "virial", "lj", "real", "recip"):
        assert rel(getattr(v, f), getattr(p, f)) < 1e-12, f
    eng.close()
