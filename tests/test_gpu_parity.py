"""GPU parity tests: every entry point of the C ABI against the CPU oracle on identical inputs.

Tolerances (north star): 1e-10 relative in FP64 for energies; S(k) element-wise 1e-12·|q|·n_sites
absolute.  Bit-exactness is not expected: sums run in parallel order and FMA contraction differs.
"""
import numpy as np
import pytest

from metropolismontecarlo_b200 import systems
from oracle import oracle as ora
from tests.util import ora_ewald, ora_system, rel

pytestmark = pytest.mark.gpu

RTOL = 1e-10


@pytest.fixture(scope="module")
def c750():
    from metropolismontecarlo_b200.energy import water_engine
    ms = systems.load_nist(4)
    eng = water_engine(ms, 10.0)
    yield ms, eng
    eng.close()


def test_kvectors_and_cfac(c750):
    ms, eng = c750
    ew = ora_ewald(ms.box)
    k, c = eng.kvectors()
    assert eng.nkvecs == 337
    assert np.array_equal(k, ew.kxyz)
    assert np.allclose(c, ew.cfac, rtol=1e-15, atol=0)


def test_single_molecule_all_i(c750):
    """LJ_poly_ΔU and EwaldReal for all 750 molecules of coord750.txt."""
    ms, eng = c750
    s = ora_system(ms)
    kappa = systems.ALPHA / ms.box
    worst = 0.0
    for i in range(1, ms.n_mol + 1):
        e, v = eng.LJ_poly_ΔU(i)
        e0, v0 = ora.LJ_poly_dU(i, s, 10.0, ms.box)
        p, ov = eng.EwaldReal(i)
        p0, ov0 = ora.EwaldReal(i, s, kappa, 10.0, ms.box)
        assert ov == ov0
        worst = max(worst, rel(e, e0), rel(p, p0))
        assert rel(e, e0) < 1e-12 and rel(p, p0) < 1e-11, i
        assert abs(v - v0) < 1e-11 * max(1.0, abs(v0), abs(e0)), i
    es, vs, ovs = eng.EwaldShort(7)
    e0, v0, _ = ora.EwaldShort(7, s, ora_ewald(ms.box), 10.0, ms.box)
    assert rel(es, e0) < 1e-11 and rel(vs, v0) < 1e-11 and not ovs
    print("worst single-molecule rel err", worst)


def test_single_molecule_rows_against_independent_pin(c750):
    """The engine's LJ_poly_ΔU(i) / EwaldReal(i) for all 750 molecules against the 40-digit mpmath evaluation of the Julia
    formulas (tests/golden/realspace_pin_coord750.npz, tests/golden/make_realspace_pin.py) — a pin that does not pass through
    the C oracle.  Per-move kernel (mmc_lj_mol / mmc_ewald_real) and the all-rows evaluation (mmc_energy_all)."""
    from pathlib import Path
    g = np.load(Path(__file__).resolve().parent / "golden" / "realspace_pin_coord750.npz")
    ms, eng = c750
    lj = np.empty(750); vir = np.empty(750); qq = np.empty(750)
    for i in range(1, 751):
        lj[i - 1], vir[i - 1] = eng.LJ_poly_ΔU(i)
        qq[i - 1], ov = eng.EwaldReal(i)
        assert ov == bool(g["overlap"][i - 1])
    for got, want in ((lj, g["lj_pot"]), (vir, g["lj_vir"]), (qq, g["qq_pot"])):
        scale = np.maximum(np.abs(want), 0.02 * np.abs(want).max())
        assert (np.abs(got - want) / scale).max() < 1e-12
    lj2, vir2, qq2, ov2 = eng.energy_all("ewald")
    for got, want in ((lj2, g["lj_pot"]), (vir2, g["lj_vir"]), (qq2 / systems.FACTOR, g["qq_pot"])):
        scale = np.maximum(np.abs(want), 0.02 * np.abs(want).max())
        assert (np.abs(got - want) / scale).max() < 1e-11
    p = eng.potential("ewald")
    assert rel(p.lj, g["lj_pot"].sum() / 2) < 1e-12 and rel(p.real, g["qq_pot"].sum() / 2 * systems.FACTOR) < 1e-11


def test_reupload_with_a_new_box_rebuilds_cfac():
    """the k-space tables survive mmc_upload_system, so a re-upload with a different box must rebuild cfac
    (2π·exp(−b k²)/k²/L with b = 1/(4κ²L²), ewalds.jl:52,78-83) for the new box and drop the resident rho(k)."""
    from metropolismontecarlo_b200.energy import water_engine
    ms = systems.load_nist(1)
    eng = water_engine(ms, 9.0)
    kappa = systems.ALPHA / ms.box
    eng.potential("ewald")
    ms2 = ms.copy()
    f = 1.04
    ms2.box = ms.box * f
    ms2.com = ms.com * f
    ms2.coords = ms.coords + np.repeat(ms2.com - ms.com, 3, axis=0)
    for n_sites_change in (False, True):
        if n_sites_change:                       # a different molecule count takes the re-allocating branch of the upload,
            keep = ms2.n_mol - 3                 # and the box changes once more
            g = 0.98
            com2 = ms2.com[:keep] * g
            ms2 = systems.MolecularSystem(ms2.coords[:3 * keep] + np.repeat(com2 - ms2.com[:keep], 3, axis=0), ms2.charge[:3 * keep].copy(),
                                          ms2.atype[:3 * keep].copy(), ms2.first_atom[:keep].copy(), ms2.last_atom[:keep].copy(), com2,
                                          ms2.eps, ms2.sig, ms2.box * g, ms2.db[:3 * keep].copy(), ms2.quat[:keep].copy())
        eng.upload_system(ms2, 9.0, 9.0)
        ew = ora.Ewald(kappa, systems.NK, systems.K_SQ_MAX, systems.FACTOR, ms2.box)      # same kappa, new box
        k, c = eng.kvectors()
        assert np.allclose(c, ew.cfac, rtol=1e-15, atol=0)
        old, new = eng.rhok()
        assert not old.any() and not new.any()
        want = ora.potential_ewald(ora_system(ms2), ew, 9.0, 9.0, ms2.box)
        got = eng.potential("ewald")
        _check_props(got, want)
        e0 = ora.RecipLong(ew, ms2.coords, ms2.charge, ms2.box)
        assert rel(eng.RecipLong(), e0) < 1e-11
        i = 7
        sl = slice(3 * (i - 1), 3 * i)
        t = eng.trial_move(i, ms2.com[i - 1] + 0.1, ms2.coords[sl] + 0.1, "ewald")
        assert abs(t.d_recip - ora.RecipMove(ms2.box, ew, ms2.coords[sl], ms2.coords[sl] + 0.1, ms2.charge[sl])) < 1e-9 * max(1.0, abs(t.d_recip))
        eng.reject()
    eng.close()


def test_recip_long_and_self(c750):
    ms, eng = c750
    ew = ora_ewald(ms.box)
    e0 = ora.RecipLong(ew, ms.coords, ms.charge, ms.box)
    e = eng.RecipLong()
    assert rel(e, e0) < RTOL
    old, new = eng.rhok()
    want = ew.sum_new[:, 0] + 1j * ew.sum_new[:, 1]
    tol = 1e-12 * np.abs(ms.charge).max() * ms.n_sites
    assert np.abs(old - want).max() < tol and np.abs(new - want).max() < tol
    assert rel(eng.EwaldSelf(), ora.EwaldSelf(ew, ms.charge)) < 1e-13


def test_recip_move_commit_rollback(c750):
    ms, eng = c750
    ew = ora_ewald(ms.box)
    ora.RecipLong(ew, ms.coords, ms.charge, ms.box)
    eng.RecipLong()
    rng = np.random.default_rng(7)
    coords = ms.coords.copy()
    for step in range(6):
        i = int(rng.integers(1, ms.n_mol + 1))
        sl = slice(3 * (i - 1), 3 * i)
        r_old = coords[sl].copy()
        r_new = r_old + (rng.random(3) - 0.5) * 0.6
        d0 = ora.RecipMove(ms.box, ew, r_old, r_new, ms.charge[sl])
        d = eng.RecipMove(r_old, r_new, ms.charge[sl])
        assert abs(d - d0) < 1e-9 * max(1.0, abs(d0))
        old, new = eng.rhok()
        assert np.abs(new - (ew.sum_new[:, 0] + 1j * ew.sum_new[:, 1])).max() < 1e-9
        assert np.abs(old - (ew.sum_old[:, 0] + 1j * ew.sum_old[:, 1])).max() < 1e-9
        if step % 2 == 0:
            ora.recip_commit(ew)
            eng.recip_commit()
            coords[sl] = r_new
        else:
            ora.recip_rollback(ew)
            eng.recip_rollback()
        old, new = eng.rhok()
        assert np.array_equal(old, new)
        assert np.abs(old - (ew.sum_old[:, 0] + 1j * ew.sum_old[:, 1])).max() < 1e-9


def _check_props(p, q, tol=RTOL):
    for k in ("energy", "virial", "coulomb", "lj", "real", "recip", "self_", "wolf_const"):
        a, b = getattr(p, k), getattr(q, k)
        assert abs(a - b) <= tol * max(abs(b), abs(q.energy) * 1e-3, 1e-30), (k, a, b)
    assert p.overlaps == q.overlaps


def test_potential_coord750_cell_mode(c750):
    """Config A: potential(…, "ewald") and the Wolf variant on coord750.txt (L = 3 r_cut → 27 cells)."""
    ms, eng = c750
    s = ora_system(ms)
    want = ora.potential_ewald(s, ora_ewald(ms.box), 10.0, 10.0, ms.box, 4)
    got = eng.potential("ewald")
    _check_props(got, want)
    assert abs(got.lj / 4.488629e5 - 1) < 2e-6 and abs(got.real / -3.492756e6 - 1) < 2e-6
    assert abs(got.recip / 7.58785e3 - 1) < 5e-6 and abs(got.self_ / -1.42235e7 - 1) < 5e-6   # NIST
    w0 = ora.potential_wolf(s, ora_ewald(ms.box), 10.0, 10.0, ms.box, 4)
    w = eng.potential("wolf")
    _check_props(w, w0)
    lj = eng.potential("lj")
    assert rel(lj.energy, want.lj) < RTOL and lj.coulomb == 0.0


@pytest.mark.parametrize("cfg,rc", [(1, 9.0), (2, 9.5), (3, 8.0)])
def test_potential_small_boxes_tile_mode(cfg, rc):
    """NIST configs 1-3 (L = 20 Å < 3 r_cut): brute-force tile path with literal vector1D."""
    from metropolismontecarlo_b200.energy import water_engine
    ms = systems.load_nist(cfg)
    eng = water_engine(ms, rc)
    s = ora_system(ms)
    want = ora.potential_ewald(s, ora_ewald(ms.box), rc, rc, ms.box, 2)
    got = eng.potential("ewald")
    _check_props(got, want)
    old, _ = eng.rhok()
    ew = ora_ewald(ms.box)
    ora.RecipLong(ew, ms.coords, ms.charge, ms.box)
    assert np.abs(old - (ew.sum_old[:, 0] + 1j * ew.sum_old[:, 1])).max() < 1e-10
    eng.close()


def test_fused_trial_move_equals_five_calls(c750):
    ms, eng = c750
    eng.upload_system(ms, 10.0, 10.0)
    eng.RecipLong()
    s = ora_system(ms)
    ew = ora_ewald(ms.box)
    ora.RecipLong(ew, ms.coords, ms.charge, ms.box)
    rng = np.random.default_rng(21)
    for style in ("ewald", "wolf", "lj"):
        for _ in range(8):
            i = int(rng.integers(1, ms.n_mol + 1))
            sl = slice(3 * (i - 1), 3 * i)
            d = (rng.random(3) - 0.5) * 0.5
            com_new, sites_new = s.com[i - 1] + d, s.coords[sl] + d
            t = eng.trial_move(i, com_new, sites_new, style)
            lo, vo = ora.LJ_poly_dU(i, s, 10.0, ms.box)
            qo, qvo, ovo = ora.EwaldShort(i, s, ew, 10.0, ms.box)
            r_old, com_old = s.coords[sl].copy(), s.com[i - 1].copy()
            s.coords[sl], s.com[i - 1] = sites_new, com_new
            ln, vn = ora.LJ_poly_dU(i, s, 10.0, ms.box)
            qn, qvn, ovn = ora.EwaldShort(i, s, ew, 10.0, ms.box)
            assert rel(t.lj_old, lo) < 1e-12 and rel(t.lj_new, ln) < 1e-12
            assert abs(t.lj_vir_old - vo) < 1e-10 * max(1, abs(vo)) and abs(t.lj_vir_new - vn) < 1e-10 * max(1, abs(vn))
            if style != "lj":
                assert rel(t.qq_old, qo) < 1e-11 and rel(t.qq_new, qn) < 1e-11
                assert rel(t.qq_vir_old, qvo) < 1e-11 and rel(t.qq_vir_new, qvn) < 1e-11
                assert (t.overlap_old, t.overlap_new) == (int(ovo), int(ovn))
            if style == "ewald":
                d0 = ora.RecipMove(ms.box, ew, r_old, sites_new, ms.charge[sl])
                assert abs(t.d_recip - d0) < 1e-9 * max(1.0, abs(d0))
            else:
                assert t.d_recip == 0.0
            if rng.random() < 0.5:
                eng.accept()
                if style == "ewald":
                    ora.recip_commit(ew)
            else:
                eng.reject()
                s.coords[sl], s.com[i - 1] = r_old, com_old
                if style == "ewald":
                    ora.recip_rollback(ew)
    coords, com = eng.download_system()
    assert np.array_equal(coords, s.coords) and np.array_equal(com, s.com)
    old, new = eng.rhok()
    assert np.abs(old - (ew.sum_old[:, 0] + 1j * ew.sum_old[:, 1])).max() < 1e-9


def test_overlap_rule(c750):
    """r² < 0.5 Å² with q_a q_b < 0 → EwaldReal returns (0.0, true) (ewalds.jl:359-360);
    potential() drops the whole row of both molecules; a trial move reports the flag."""
    ms0, eng = c750
    ms = ms0.copy()
    # put molecule 2 so that its first H sits 0.3 Å from the O of molecule 1
    shift = (ms.coords[0] + np.array([0.3, 0.0, 0.0])) - ms.coords[4]
    ms.coords[3:6] += shift
    ms.com[1] += shift
    ms.com[1] = np.clip(ms.com[1], 0.0, ms.box)
    eng.upload_system(ms, 10.0, 10.0)
    s = ora_system(ms)
    kappa = systems.ALPHA / ms.box
    for i in (1, 2, 3):
        p, ov = eng.EwaldReal(i)
        p0, ov0 = ora.EwaldReal(i, s, kappa, 10.0, ms.box)
        assert ov == ov0 and rel(p, p0) < 1e-11 if p0 != 0 else p == 0.0
    assert eng.EwaldReal(1) == (0.0, True) and eng.EwaldReal(2) == (0.0, True)
    want = ora.potential_ewald(s, ora_ewald(ms.box), 10.0, 10.0, ms.box, 4)
    got = eng.potential("ewald")
    assert want.overlaps >= 2
    _check_props(got, want)
    t = eng.trial_move(3, ms.com[2], ms.coords[6:9], "ewald")
    assert t.overlap_old == 0 and t.overlap_new == 0
    t = eng.trial_move(1, ms.com[0], ms.coords[0:3], "ewald")
    assert t.overlap_old == 1 and t.overlap_new == 1 and t.qq_old == 0.0 and t.d_recip == 0.0
    eng.reject()
    eng.upload_system(ms0, 10.0, 10.0)


def test_volume_trial_matches_scaled_recompute(c750):
    """Ewald/volumeChange.jl:50-147: scale COMs, rigid-shift sites, κ = α/L', full energy at L'."""
    ms, eng = c750
    eng.upload_system(ms, 10.0, 10.0)
    e_before = eng.potential("ewald").energy
    for box_new in (30.4, 29.8):      # 29.8 < 3 r_cut → the trial falls back to tile mode
        s = ora_system(ms)
        ora.volume_scale(s, ms.box, box_new)
        ew = ora.Ewald(systems.ALPHA / box_new, 5, 27, systems.FACTOR, box_new)
        want = ora.potential_ewald(s, ew, 10.0, 10.0, box_new, 4)
        got = eng.volume_trial(box_new, systems.ALPHA / box_new, "ewald")
        _check_props(got, want)
        eng.volume_reject()
        assert rel(eng.potential("ewald").energy, e_before) < 1e-13
    got = eng.volume_trial(30.4, systems.ALPHA / 30.4, "ewald")
    eng.volume_accept()
    again = eng.potential("ewald")
    assert rel(again.energy, got.energy) < 1e-12
    coords, com = eng.download_system()
    s = ora_system(ms)
    ora.volume_scale(s, ms.box, 30.4)
    assert np.array_equal(coords, s.coords) and np.array_equal(com, s.com)
    # a molecule move after the accepted volume move uses the new box, κ, cfac and ρ(k)
    ew = ora.Ewald(systems.ALPHA / 30.4, 5, 27, systems.FACTOR, 30.4)
    ora.RecipLong(ew, s.coords, s.charge, 30.4)
    i = 11
    sl = slice(30, 33)
    t = eng.trial_move(i, s.com[i - 1] + 0.2, s.coords[sl] + 0.2, "ewald")
    d0 = ora.RecipMove(30.4, ew, s.coords[sl], s.coords[sl] + 0.2, s.charge[sl])
    qo = ora.EwaldShort(i, s, ew, 10.0, 30.4)[0]
    assert abs(t.d_recip - d0) < 1e-9 * max(1.0, abs(d0)) and rel(t.qq_old, qo) < 1e-11
    eng.reject()
    eng.upload_system(ms, 10.0, 10.0)
    eng.PrepareEwaldVariables(systems.ALPHA / ms.box)


def test_sharded_partials_sum_to_unsharded(c750):
    """§8e: R ranks emulated as R handles run one after another on one GPU; partial vectors are
    summed (what the NCCL all-reduce does) and finalised on every 'rank'."""
    import torch
    from metropolismontecarlo_b200.energy import water_engine
    ms, eng = c750
    eng.upload_system(ms, 10.0, 10.0)
    ref = eng.potential("ewald")
    for world in (2, 4, 8):
        engs = [water_engine(ms, 10.0, rank=r, world=world) for r in range(world)]
        n = engs[0].partial_count()
        bufs = [torch.zeros(n, dtype=torch.float64, device="cuda") for _ in range(world)]
        for e, b in zip(engs, bufs):
            e.potential_partial("ewald", b.data_ptr())
        torch.cuda.synchronize()
        total = torch.stack(bufs).sum(0)
        props = []
        for e in engs:      # (coord750 has molecules wrapped across the box: k_pairs_v7 declines, every rank sees it in the
            buf = total.clone()     # summed vector and evaluates the replicated state on the general path — no retry protocol)
            props.append(e.potential_finalize("ewald", buf.data_ptr()))
        for e, p in zip(engs, props):
            _check_props(p, ref, 1e-12)
            old, _ = e.rhok()
            assert np.abs(old - eng.rhok()[0]).max() < 1e-9
        for e in engs:
            e.close()


def test_config_d_4000_molecules():
    """Config D: 4000 SPC/E on the cubic lattice with random quaternions (cell mode, 4³ cells)."""
    from metropolismontecarlo_b200.energy import water_engine
    ms = systems.spce_lattice(4000)
    eng = water_engine(ms, 10.0)
    s = ora_system(ms)
    want = ora.potential_ewald(s, ora_ewald(ms.box), 10.0, 10.0, ms.box, 8)
    got = eng.potential("ewald")
    _check_props(got, want)
    for i in (1, 1999, 4000):
        assert rel(eng.LJ_poly_ΔU(i)[0], ora.LJ_poly_dU(i, s, 10.0, ms.box)[0]) < 1e-12
    box_new = ms.box * 1.01
    ora.volume_scale(s, ms.box, box_new)
    w2 = ora.potential_ewald(s, ora.Ewald(systems.ALPHA / box_new, 5, 27, systems.FACTOR, box_new), 10.0, 10.0, box_new, 8)
    _check_props(eng.volume_trial(box_new, systems.ALPHA / box_new, "ewald"), w2)
    eng.close()


def npr_pairs_in_cutoff(com, box, rc):
    """unordered molecule pairs with |COM_ij|² < rc² under the reference's minimum image (numpy, O(N²) in blocks)"""
    from oracle import numpy_ref as npr
    n, tot = len(com), 0
    for lo in range(0, n, 256):
        d = npr.vector1d(com[lo:lo + 256, None, :], com[None, :, :], box)
        r2 = (d * d).sum(-1)
        idx = np.arange(lo, min(lo + 256, n))[:, None]
        tot += int(((r2 < rc * rc) & (np.arange(n)[None, :] > idx)).sum())
    return tot


@pytest.mark.parametrize("amp", [0.35, 1.2])
def test_pair_kernel_variants_agree_with_oracle(amp):
    """Every pair kernel (k_pairs_v7, k_pairs_fast tiles, general k_pairs) on a disordered box: 2197 SPC/E molecules, lattice
    COMs displaced by up to `amp` Å, so cells are unevenly filled.  amp = 1.2 Å makes 59 molecules overlap (ewalds.jl:359-360
    zeroes their whole rows): k_pairs_v7 must detect that and hand the state to the general path, which implements the rule."""
    from metropolismontecarlo_b200.energy import water_engine
    ms = systems.spce_lattice(2197)
    rng = np.random.default_rng(9)
    d = rng.uniform(-amp, amp, ms.com.shape)
    newcom = np.clip(ms.com + d, 0.0, ms.box)
    ms.coords = ms.coords + np.repeat(newcom - ms.com, 3, axis=0)
    ms.com = newcom
    s = ora_system(ms)
    want = ora.potential_ewald(s, ora_ewald(ms.box), 10.0, 10.0, ms.box, 8)
    assert (want.overlaps > 0) == (amp > 1.0)
    want_pairs = npr_pairs_in_cutoff(ms.com, ms.box, 10.0)
    eng = water_engine(ms, 10.0)
    want_wolf = ora.potential_wolf(s, ora_ewald(ms.box), 10.0, 10.0, ms.box, 8)
    for level, name in ((0, "k_pairs_v7" if want.overlaps == 0 else "k_pairs_fast<64>"), (1, "k_pairs_fast<64>"), (2, "k_pairs")):
        eng.upload_system(ms, 10.0, 10.0)              # (an upload resets the kernel chain to the requested level)
        eng.debug_set("pair_level", level)
        got = eng.potential("ewald")
        assert eng.last_eval_info()["pair_kernel"] == name, (level, eng.last_eval_info())
        assert eng.last_eval_info()["pairs_in_cutoff"] == want_pairs
        assert got.overlaps == want.overlaps
        _check_props(got, want)
        _check_props(eng.potential("wolf"), want_wolf)
    eng.debug_set("pair_level", 0)
    lj = eng.potential("lj")                       # no Coulomb: served by k_pairs_fast
    assert eng.last_eval_info()["pair_kernel"].startswith("k_pairs_fast") and rel(lj.energy, want.lj) < RTOL
    eng.close()
    eng = water_engine(ms, 10.0)
    eng.upload_system(ms, 10.0, 9.0)               # unequal cut-offs: not v3 territory
    s9 = ora.potential_ewald(s, ora_ewald(ms.box), 10.0, 9.0, ms.box, 8)
    g9 = eng.potential("ewald")
    assert eng.last_eval_info()["pair_kernel"].startswith("k_pairs_fast")
    _check_props(g9, s9)
    eng.close()


def test_monatomic_known_answers_and_parity():
    from metropolismontecarlo_b200.energy import Engine
    eng = Engine()
    # the reference's own test_LJ (Ewald/tests.jl:127-161)
    r = np.array([[0, 0, 0], [0, 0, 2], [0, 1.5, 0]], dtype=np.float64)
    eng.upload_atoms(systems.AtomicSystem(r, np.ones(3), np.ones(3), 5.0, 2.5))
    assert abs(eng.LJ_ΔU(1)[0] - (-0.381860031778575)) < 1e-14
    eng.set_atom(2, [0, 0, 4.0])
    assert abs(eng.LJ_ΔU(1)[0] - (-0.320336594278575)) < 1e-14
    # lattice + noise, 2197 atoms, heterogeneous eps/sig to exercise the per-j parameters
    at = systems.lj_lattice(2197, 0.75, 2.5)
    rng = np.random.default_rng(5)
    at.r[:] = (at.r + rng.normal(0, 0.08, at.r.shape)) % at.box
    at.eps[:] = 0.8 + 0.4 * rng.random(at.n)
    at.sig[:] = 0.95 + 0.1 * rng.random(at.n)
    eng.upload_atoms(at)
    for i in (1, 2, 1000, 2197):
        e, v = eng.LJ_ΔU(i)
        e0, v0 = ora.LJ_dU_atom(i, at.r, at.eps, at.sig, at.box, at.r_cut)
        assert rel(e, e0) < 1e-12 and rel(v, v0) < 1e-12
    p = eng.potential("atoms")
    e0, v0 = ora.potential_atoms(at.r, at.eps, at.sig, at.box, at.r_cut, 4)
    assert rel(p.energy, e0) < 1e-12 and rel(p.virial, v0) < 1e-12
    t = eng.trial_atom(17, at.r[16] + 0.05)
    r2 = at.r.copy()
    r2[16] += 0.05
    assert rel(t.lj_old, ora.LJ_dU_atom(17, at.r, at.eps, at.sig, at.box, at.r_cut)[0]) < 1e-12
    assert rel(t.lj_new, ora.LJ_dU_atom(17, r2, at.eps, at.sig, at.box, at.r_cut)[0]) < 1e-12
    eng.accept()
    assert np.array_equal(eng.download_atoms(), r2)
    eng.close()


def test_error_behaviour():
    from metropolismontecarlo_b200.energy import Engine, MMCError
    eng = Engine()
    with pytest.raises(MMCError):
        eng.potential("ewald")              # nothing uploaded
    ms = systems.load_nist(1)
    eng.upload_system(ms, 9.0)
    with pytest.raises(MMCError):
        eng.EwaldReal(1)                    # no Ewald tables yet
    with pytest.raises(MMCError):
        eng.LJ_poly_ΔU(0)                   # indices are 1-based
    with pytest.raises(MMCError):
        eng.accept()                        # no pending trial
    assert eng.LJ_poly_ΔU(1)[0] != 0.0
    eng.close()


def test_upload_positions_equals_full_upload(c750):
    """mmc_upload_positions (bulk set_molecule) leaves the same state as a full mmc_upload_system of the moved system."""
    ms, eng = c750
    eng.upload_system(ms, 10.0, 10.0)
    rng = np.random.default_rng(4)
    ms2 = ms.copy()
    d = rng.uniform(-0.2, 0.2, ms.com.shape)
    ms2.com = np.clip(ms.com + d, 0.0, ms.box)
    ms2.coords = ms.coords + np.repeat(ms2.com - ms.com, 3, axis=0)
    eng.upload_positions(ms2.coords, ms2.com)
    got = eng.potential("ewald")
    coords, com = eng.download_system()
    assert np.array_equal(coords, ms2.coords) and np.array_equal(com, ms2.com)
    eng.upload_system(ms2, 10.0, 10.0)
    want = eng.potential("ewald")
    _check_props(got, want, 1e-13)
    bad = ms2.com.copy(); bad[3, 0] = ms.box + 1.0
    with pytest.raises(Exception):
        eng.upload_positions(ms2.coords, bad)
    eng.upload_system(ms, 10.0, 10.0)


def test_peer_exchange_emulated_ranks(c750):
    """The NVLink peer-memory exchange (k_peer_push / k_peer_sum) with R ranks emulated as R handles in one process
    (same-process import by pointer): begin on every rank, then end on every rank; totals equal the unsharded
    evaluation and are bit-identical across ranks; repeated evaluations exercise the epoch/parity protocol."""
    from metropolismontecarlo_b200.energy import water_engine
    ms, eng = c750
    eng.upload_system(ms, 10.0, 10.0)
    ref = eng.potential("ewald")
    ms_big = systems.spce_lattice(4000)
    for msx, world in ((ms, 2), (ms, 8), (ms_big, 4)):
        engs = [water_engine(msx, 10.0, rank=r, world=world) for r in range(world)]
        for e in engs:
            e.peer_export()
        for e in engs:
            for r, o in enumerate(engs):
                e.peer_import_ptr(r, o.peer_buffer())
        want = ref if msx is ms else None
        for rep in range(3):
            for e in engs:
                e.potential_sharded_begin("ewald")
            props = [e.potential_sharded_end() for e in engs]
            if want is None:
                single = water_engine(msx, 10.0)
                want = single.potential("ewald")
                single.close()
            for p in props:
                _check_props(p, want, 1e-12)
                assert p.energy == props[0].energy and p.recip == props[0].recip     # same order of summation on every rank
        for e in engs:
            e.close()


def test_tip3p_1000_molecules_reference_run():
    """The reference's own shipped run (Ewald/main.jl "crystal" branch): 1000 TIP3P molecules (water.top + tip3p.pdb) on
    InitCubicGrid at rho = 0.033101144, random orientations, r_cut = 10, kappa = 5.6/L, 337 k-vectors."""
    from metropolismontecarlo_b200.energy import LoopParams, julia_rand, water_engine
    ms = systems.rigid_lattice(systems.tip3p_model(), 1000)
    s = ora_system(ms)
    ew = ora_ewald(ms.box)
    eng = water_engine(ms, 10.0)
    for style, want in (("ewald", ora.potential_ewald(s, ew, 10.0, 10.0, ms.box, 8)),
                        ("wolf", ora.potential_wolf(s, ora_ewald(ms.box), 10.0, 10.0, ms.box, 8))):
        _check_props(eng.potential(style), want)
    for i in (1, 17, 500, 1000):
        lj, vir = eng.LJ_poly_ΔU(i)
        e, v, ovl = eng.EwaldShort(i)
        wl = ora.LJ_poly_dU(i, s, 10.0, ms.box)
        we = ora.EwaldShort(i, s, ew, 10.0, ms.box)
        assert rel(lj, wl[0]) < 1e-12 and rel(vir, wl[1]) < 1e-11 and rel(e, we[0]) < 1e-12 and bool(ovl) == bool(we[2])
    # 3000 moves of the reference loop: per-move protocol, block offload and oracle agree move by move
    u = julia_rand(11234, 8 * 3000)
    p0 = ora.potential_ewald(s, ew, 10.0, 10.0, ms.box, 8)
    prm = ora.LoopParams(298.15, 0.316555789, 0.05, 0.5, 1.0, 10.0, 10.0, ms.box, 0, 1)
    quat_o = ms.quat.copy()
    rc_o, acc_o, del_o, st_o = ora.loop(s, ew, ms.db, quat_o, prm, u, 3000, p0.energy, p0.virial)
    for device in (False, True):
        eng.upload_system(ms, 10.0, 10.0)
        g0 = eng.potential("ewald")
        com, quat = ms.com.copy(), ms.quat.copy()
        rc_g, acc_g, del_g, st_g = eng.loop_run(LoopParams(298.15, 0.316555789, 0.05, 0.5, 1.0, 0, 1), com, quat, ms.db, u, 3000,
                                                g0.energy, g0.virial, device=device)
        assert rc_g == 0 and np.array_equal(acc_g, acc_o) and st_g.uniforms_used == st_o.uniforms_used
        assert np.abs(com - s.com).max() < 1e-11
        assert rel(st_g.total_energy, eng.potential("ewald").energy) < 1e-10
    eng.close()


def test_pairs_v7_two_pass_groups():
    """k_pairs_v7's two-pass path: 3200 SPC/E molecules on the 15³ lattice give 4 cells of 11.5 Å per edge holding
    27 … 64 molecules (3 or 4 lattice planes per cell and direction), so 5-slot groups of up to 288 molecules exceed the
    256-row B tile and are evaluated in two passes.  Totals and pair count against the oracle."""
    from metropolismontecarlo_b200.energy import water_engine
    ms = systems.spce_lattice(3200)
    eng = water_engine(ms, 10.0)
    got = eng.potential("ewald")
    info = eng.last_eval_info()
    assert info["pair_kernel"] == "k_pairs_v7" and info["cells_per_dim"] == 4, info
    s = ora_system(ms)
    want = ora.potential_ewald(s, ora_ewald(ms.box), 10.0, 10.0, ms.box, 8)
    _check_props(got, want)
    assert info["pairs_in_cutoff"] == npr_pairs_in_cutoff(ms.com, ms.box, 10.0)
    eng.close()


def test_pair_kernel_declines_dense_cells_with_stale_density():
    """A cell above 64 molecules that the cached density does not know about (positions re-sent with
    mmc_upload_positions): k_pairs_v7 must decline cleanly (regression: its boundary fix-up ran on a declined unit)
    and the chain must end on a kernel that gives the oracle's answer."""
    from metropolismontecarlo_b200.energy import water_engine
    ms = systems.spce_lattice(4000)                 # 4 cells per edge, full cells hold exactly 64 molecules
    eng = water_engine(ms, 10.0)
    eng.potential("ewald")
    assert eng.last_eval_info()["pair_kernel"] == "k_pairs_v7"
    rng = np.random.default_rng(2)
    newcom = np.clip(ms.com + rng.uniform(-1.5, 1.5, ms.com.shape), 0.0, ms.box)
    ms2 = ms.copy()
    ms2.coords = ms.coords + np.repeat(newcom - ms.com, 3, axis=0)
    ms2.com = newcom
    s = ora_system(ms2)
    want = ora.potential_ewald(s, ora_ewald(ms.box), 10.0, 10.0, ms.box, 8)
    for rep in range(3):
        eng.upload_system(ms, 10.0, 10.0)
        eng.potential("ewald")
        eng.upload_positions(ms2.coords, ms2.com)
        got = eng.potential("ewald")
        assert eng.last_eval_info()["pair_kernel"] != "k_pairs_v7"
        _check_props(got, want)
    eng.close()


def test_pairs_v7_sparse_box_with_empty_cells():
    """Low density: 6 cells per edge with a handful of molecules each and some empty ones (zero-size tiles), molecules
    clustered in one corner so that whole neighbour cells are empty."""
    from metropolismontecarlo_b200.energy import water_engine
    ms = systems.spce_lattice(600, rho=0.0025)
    rng = np.random.default_rng(8)
    newcom = ms.com.copy()
    g = np.stack(np.meshgrid(*(np.arange(6),) * 3, indexing="ij"), -1).reshape(-1, 3)[:200]
    newcom[:200] = 0.5 + 3.2 * g + rng.uniform(-0.2, 0.2, (200, 3))   # a dense corner (3.2 Å grid: no overlapping molecules) ...
    newcom[200:] = np.clip(ms.com[200:] + rng.uniform(-2, 2, (400, 3)), 0.0, ms.box)
    ms.coords = ms.coords + np.repeat(newcom - ms.com, 3, axis=0)
    ms.com = newcom
    eng = water_engine(ms, 10.0)
    got = eng.potential("ewald")
    info = eng.last_eval_info()
    assert info["pair_kernel"] == "k_pairs_v7" and info["cells_per_dim"] >= 5, info
    s = ora_system(ms)
    want = ora.potential_ewald(s, ora_ewald(ms.box), 10.0, 10.0, ms.box, 8)
    assert info["pairs_in_cutoff"] == npr_pairs_in_cutoff(ms.com, ms.box, 10.0)
    assert got.overlaps == want.overlaps
    if want.overlaps == 0:
        _check_props(got, want)
    eng.close()


def test_potential_host_pipelined_equals_upload_then_potential():
    """mmc_potential_host (COMs first, sites in chunks on the side stream feeding the rho(k) rebuild, gather + pairs last)
    against mmc_upload_positions + mmc_potential, Ewald / Wolf / LJ, twice (the state it leaves is the uploaded one)."""
    from metropolismontecarlo_b200.energy import water_engine
    ms = systems.spce_lattice(8000)
    eng = water_engine(ms, 10.0)
    eng.potential("ewald")
    rng = np.random.default_rng(12)
    for rep in range(2):
        newcom = np.clip(ms.com + rng.uniform(-0.3, 0.3, ms.com.shape), 0.0, ms.box)
        coords = ms.coords + np.repeat(newcom - ms.com, 3, axis=0)
        for style in ("ewald", "wolf", "lj"):
            got = eng.potential_host(coords, newcom, style)
            c2, m2 = eng.download_system()
            assert np.array_equal(c2, coords) and np.array_equal(m2, newcom)
            eng.upload_positions(coords, newcom)
            want = eng.potential(style)
            _check_props(got, want, 1e-13)
        if rep == 0:
            assert np.abs(eng.rhok()[0]).max() > 0
    bad = newcom.copy(); bad[7, 2] = -1.0
    with pytest.raises(Exception):
        eng.potential_host(coords, bad, "ewald")
    eng.close()


def test_energy_all_rows_equal_per_molecule_calls(c750):
    """mmc_energy_all: LJ_poly_ΔU(i) and EwaldShort(i) for every i from ONE pass over the unique pairs,
    against the oracle's per-molecule functions (energy.jl:209-290, ewalds.jl:293-376, 892-910) on coord750 (cell mode),
    on a small NIST box (tile mode), with the overlap rule, and for the Wolf / LJ-only styles."""
    from metropolismontecarlo_b200.energy import water_engine
    ms, eng = c750
    for rc, m in ((10.0, ms), (9.0, systems.load_nist(1))):      # cell mode (3^3 cells) and tile mode (L = 20)
        eng.upload_system(m, rc, rc)
        eng.PrepareEwaldVariables(systems.ALPHA / m.box)
        s = ora_system(m)
        ew = ora_ewald(m.box)
        lj, vir, qq, ov = eng.energy_all("ewald")
        assert not ov.any()
        for i in range(1, m.n_mol + 1):
            e0, v0 = ora.LJ_poly_dU(i, s, rc, m.box)
            c0, _, ov0 = ora.EwaldShort(i, s, ew, rc, m.box)
            assert not ov0
            assert rel(lj[i - 1], e0) < 1e-11 and rel(qq[i - 1], c0) < 1e-10, (rc, i)
            assert abs(vir[i - 1] - v0) < 1e-10 * max(1.0, abs(v0), abs(e0)), (rc, i)
        # Σ_i rows / 2 is what potential() reports (energy.jl:966-1001)
        p = eng.potential("ewald")
        assert rel(lj.sum() / 2, p.lj) < 1e-11 and rel(qq.sum() / 2, p.real) < 1e-11
        ljw, _, qqw, _ = eng.energy_all("wolf")
        assert np.allclose(ljw, lj, rtol=1e-12) and np.allclose(qqw, qq, rtol=1e-12)
        ljo, _, qqo, ovo = eng.energy_all("lj")
        assert np.allclose(ljo, lj, rtol=1e-12) and not qqo.any() and not ovo.any()
    # overlap: both molecules of the offending pair report (0, overlap) like EwaldReal's early return
    mo = ms.copy()
    shift = (mo.coords[0] + np.array([0.3, 0.0, 0.0])) - mo.coords[4]
    mo.coords[3:6] += shift
    mo.com[1] = np.clip(mo.com[1] + shift, 0.0, mo.box)
    eng.upload_system(mo, 10.0, 10.0)
    eng.PrepareEwaldVariables(systems.ALPHA / mo.box)
    s = ora_system(mo)
    ew = ora_ewald(mo.box)
    lj, vir, qq, ov = eng.energy_all("ewald")
    assert ov[0] == 1 and ov[1] == 1 and ov.sum() == 2 and qq[0] == 0.0 and qq[1] == 0.0
    for i in (1, 2, 3, 17, 400):
        c0, _, ov0 = ora.EwaldShort(i, s, ew, 10.0, mo.box)
        assert bool(ov[i - 1]) == bool(ov0) and (rel(qq[i - 1], c0) < 1e-10 if c0 != 0 else qq[i - 1] == 0.0)
    eng.upload_system(ms, 10.0, 10.0)
    # config D (4000 molecules, 4^3+ cells): rows against the engine's own per-molecule kernel
    md = systems.spce_lattice(4000)
    ed = water_engine(md, 10.0)
    lj, vir, qq, ov = ed.energy_all("ewald")
    for i in (1, 2, 1999, 4000):
        e, v = ed.LJ_poly_ΔU(i)
        c, _, o = ed.EwaldShort(i)
        assert rel(lj[i - 1], e) < 1e-11 and rel(qq[i - 1], c) < 1e-10 and abs(vir[i - 1] - v) < 1e-9 * max(1.0, abs(v))
    p = ed.potential("ewald")
    assert rel(lj.sum() / 2, p.lj) < 1e-11 and rel(qq.sum() / 2, p.real) < 1e-10
    ed.close()


def _subsystem(ms, n):
    out = ms.copy()
    out.coords, out.charge, out.atype = ms.coords[:3 * n].copy(), ms.charge[:3 * n].copy(), ms.atype[:3 * n].copy()
    out.first_atom, out.last_atom, out.com = ms.first_atom[:n].copy(), ms.last_atom[:n].copy(), ms.com[:n].copy()
    out.db, out.quat = ms.db[:3 * n].copy(), ms.quat[:n].copy()
    return out


@pytest.mark.parametrize("n", [1, 2, 3, 5])
def test_tiny_systems(n):
    """Ragged / minimum sizes: 1–5 molecules (a lone molecule has no pair term at all; fewer molecules than lanes, than
    CTAs of the move kernel, than k-space site chunks) through every entry point that takes the system as a whole, in a
    cell-mode box (L = 30) and a tile-mode box (L = 20)."""
    from metropolismontecarlo_b200.energy import LoopParams, water_engine
    for cfg, rc in ((4, 10.0), (1, 9.0)):
        ms = _subsystem(systems.load_nist(cfg), n)
        s = ora_system(ms)
        ew = ora_ewald(ms.box)
        eng = water_engine(ms, rc)
        want = ora.potential_ewald(s, ew, rc, rc, ms.box, 1)
        _check_props(eng.potential("ewald"), want)
        _check_props(eng.potential("wolf"), ora.potential_wolf(s, ew, rc, rc, ms.box, 1))
        if n == 1:
            assert want.lj == 0.0 and want.real == 0.0 and want.recip != 0.0
        lj, vir, qq, ov = eng.energy_all("ewald")
        for i in range(1, n + 1):
            e0, v0 = ora.LJ_poly_dU(i, s, rc, ms.box)
            c0, _, _ = ora.EwaldShort(i, s, ew, rc, ms.box)
            assert abs(lj[i - 1] - e0) <= 1e-11 * abs(e0) and abs(qq[i - 1] - c0) <= 1e-10 * abs(c0)
            assert abs(eng.LJ_poly_ΔU(i)[0] - e0) <= 1e-12 * abs(e0)
            assert abs(eng.EwaldShort(i)[0] - c0) <= 1e-11 * abs(c0)
        # a trial move of the last molecule: ΔU == U(new) − U(old) from two full evaluations
        d = np.array([0.11, -0.07, 0.05])
        t = eng.trial_move(n, ms.com[n - 1] + d, ms.coords[3 * (n - 1):] + d, "ewald")
        eng.accept()
        s.coords[3 * (n - 1):] += d
        s.com[n - 1] += d
        after = ora.potential_ewald(s, ora_ewald(ms.box), rc, rc, ms.box, 1)
        du = (t.lj_new - t.lj_old) + (t.qq_new - t.qq_old) + t.d_recip
        assert abs(du - (after.energy - want.energy)) < 1e-9 * max(1.0, abs(want.energy) * 1e-3)
        _check_props(eng.potential("ewald"), after)
        # block of moves on the device == the per-move protocol on the same uniforms
        u = np.random.default_rng(n).random(4000)
        outs = []
        for device in (False, True):
            e2 = water_engine(ms, rc)
            g0 = e2.potential("ewald")
            com, quat = ms.com.copy(), ms.quat.copy()
            r, acc, delta, st = e2.loop_run(LoopParams(298.15, 0.3, 0.05, 0.5, 1.0, 0, 1), com, quat, ms.db, u, 300,
                                            g0.energy, g0.virial, device=device)
            fresh = e2.potential("ewald")
            assert r == 0 and abs(st.total_energy - fresh.energy) < 1e-9 * abs(fresh.energy)
            outs.append((acc.copy(), st.uniforms_used, com.copy()))
            e2.close()
        assert np.array_equal(outs[0][0], outs[1][0]) and outs[0][1] == outs[1][1]
        assert np.abs(outs[0][2] - outs[1][2]).max() < 1e-12
        eng.close()


def test_host_register_pins_caller_arrays():
    """mmc_host_register / mmc_host_unregister: a caller's pageable arrays page-locked in place; uploads from them give
    the same result; registering twice is accepted, unregistering an unknown pointer is MMC_ESTATE."""
    from metropolismontecarlo_b200.energy import MMCError, host_register, host_unregister, water_engine
    ms = systems.spce_lattice(8000)
    eng = water_engine(ms, 10.0)
    ref = eng.potential("ewald")
    coords, com = ms.coords.copy(), ms.com.copy()
    host_register(coords); host_register(com); host_register(com)
    got = eng.potential_host(coords, com, "ewald")
    _check_props(got, ref, 1e-13)
    host_unregister(coords); host_unregister(com)
    with pytest.raises(MMCError):
        host_unregister(com)
    eng.close()


def test_mixed_topology_water_and_ions():
    """A non-uniform system (3-site SPC/E + 1-site ions, three LJ types; the reference's routines take per-molecule
    firstAtom/lastAtom, Ewald/energy.jl:219-226) through mmc_potential (Ewald, Wolf, LJ), the single-molecule entry points,
    mmc_trial_move and a block of moves of mmc_loop_run, all against the oracle."""
    from metropolismontecarlo_b200.energy import LoopParams, julia_rand, water_engine
    ms = systems.water_ion_mixture(512, 40)
    assert ms.n_sites == 3 * 472 + 40 and ms.eps.shape == (3, 3)
    s = ora_system(ms)
    ew = ora_ewald(ms.box)
    eng = water_engine(ms, 10.0)
    for style, want in (("ewald", ora.potential_ewald(s, ew, 10.0, 10.0, ms.box, 8)), ("wolf", ora.potential_wolf(s, ew, 10.0, 10.0, ms.box, 8))):
        got = eng.potential(style)
        _check_props(got, want)
        assert got.overlaps == want.overlaps
    kappa = systems.ALPHA / ms.box
    for i in (1, 2, 13, 14, 256, 511, 512):
        e, v = eng.LJ_poly_ΔU(i)
        e0, v0 = ora.LJ_poly_dU(i, s, 10.0, ms.box)
        p, ov = eng.EwaldReal(i)
        p0, ov0 = ora.EwaldReal(i, s, kappa, 10.0, ms.box)
        assert rel(e, e0) < 1e-11 and abs(v - v0) < 1e-10 * max(1.0, abs(v0), abs(e0)) and rel(p, p0) < 1e-10 and ov == ov0, i
    # a block of moves through the per-move protocol (ions: translations; rotations act on a single site at the COM)
    p0 = ora.potential_ewald(s, ew, 10.0, 10.0, ms.box, 8)
    g0 = eng.potential("ewald")
    u = julia_rand(11234, 8 * 1500)
    q_o = ms.quat.copy()
    prm = ora.LoopParams(298.15, 0.3, 0.05, 0.5, 1.0, 10.0, 10.0, ms.box, 0, 1)
    rc_o, acc_o, del_o, st_o = ora.loop(s, ew, ms.db, q_o, prm, u, 1500, p0.energy, p0.virial)
    com, quat = ms.com.copy(), ms.quat.copy()
    rc_g, acc_g, del_g, st_g = eng.loop_run(LoopParams(298.15, 0.3, 0.05, 0.5, 1.0, 0, 1), com, quat, ms.db, u, 1500, g0.energy, g0.virial)
    assert rc_o == 0 and rc_g == 0 and np.array_equal(acc_g, acc_o) and st_g.uniforms_used == st_o.uniforms_used
    assert np.abs(del_g - del_o).max() < 1e-9 * max(1.0, np.abs(del_o).max())
    assert np.abs(com - s.com).max() < 1e-12
    fresh = eng.potential("ewald")
    assert rel(st_g.total_energy, fresh.energy) < 1e-9
    eng.close()


@pytest.mark.parametrize("n_mol,n_ions", [(4096, 300), (512, 40)])
def test_mixed_topology_on_the_pair_kernels(n_mol, n_ions):
    """a non-uniform topology (3-site water + 1-site ions, three LJ types) is evaluated by the cell / tile
    pair kernel k_pairs on a copy padded to max_sites slots per molecule (LJ resolved through the active type pairs) instead of
    N k_move launches.  4096 molecules: cell mode (4³ cells); 512: tile mode (box < 3 r_cut).  Against the oracle's potential(),
    the literal Σ_i rows / 2 (debug pair_level 3), mmc_energy_all against the per-molecule calls, a volume trial against the
    scaled recompute, and the sharded partials (2 and 3 emulated ranks) against the unsharded evaluation."""
    import torch
    from metropolismontecarlo_b200.energy import water_engine
    ms = systems.water_ion_mixture(n_mol, n_ions)
    s = ora_system(ms)
    ew = ora_ewald(ms.box)
    eng = water_engine(ms, 10.0)
    wants = {"ewald": ora.potential_ewald(s, ew, 10.0, 10.0, ms.box, 8), "wolf": ora.potential_wolf(s, ew, 10.0, 10.0, ms.box, 8)}
    for style, want in wants.items():
        got = eng.potential(style)
        info = eng.last_eval_info()
        assert info["pair_kernel"] == "k_pairs" and info["mode"] == ("cells" if n_mol == 4096 else "tiles"), info
        _check_props(got, want)
        assert got.overlaps == want.overlaps == 0
    got = eng.potential("ewald")
    eng.debug_set("pair_level", 3)
    rows = eng.potential("ewald")
    eng.debug_set("pair_level", 0)
    _check_props(got, rows, 1e-11)
    lj = eng.potential("lj")
    assert rel(lj.energy, wants["ewald"].lj) < 1e-11
    # every molecule's rows from the one pass, ions included
    ljr, virr, qqr, ovr = eng.energy_all("ewald")
    assert not ovr.any()
    for i in (1, 2, 14, 15, n_mol // 2, n_mol):
        e0, v0 = ora.LJ_poly_dU(i, s, 10.0, ms.box)
        c0, _, ov0 = ora.EwaldShort(i, s, ew, 10.0, ms.box)
        assert not ov0 and rel(ljr[i - 1], e0) < 1e-11 and abs(virr[i - 1] - v0) < 1e-10 * max(1.0, abs(v0), abs(e0)) and rel(qqr[i - 1], c0) < 1e-10, i
    assert rel(ljr.sum() / 2, got.lj) < 1e-11 and rel(qqr.sum() / 2, got.real) < 1e-11
    # volume trial: COMs scaled, sites shifted rigidly, κ = α/L'
    box_new = ms.box * 1.01
    s2 = ora_system(ms)
    ora.volume_scale(s2, ms.box, box_new)
    w2 = ora.potential_ewald(s2, ora.Ewald(systems.ALPHA / box_new, 5, 27, systems.FACTOR, box_new), 10.0, 10.0, box_new, 8)
    _check_props(eng.volume_trial(box_new, systems.ALPHA / box_new, "ewald"), w2)
    eng.volume_reject()
    assert eng.potential("ewald").energy == got.energy
    # sharded: partial vectors of R ranks summed = the unsharded evaluation
    for world in (2, 3):
        engs = [water_engine(ms, 10.0, rank=r, world=world) for r in range(world)]
        n = engs[0].partial_count()
        bufs = [torch.zeros(n, dtype=torch.float64, device="cuda") for _ in range(world)]
        for e, b in zip(engs, bufs):
            e.potential_partial("ewald", b.data_ptr())
        torch.cuda.synchronize()
        total = torch.stack(bufs).sum(0)
        for e in engs:
            _check_props(e.potential_finalize("ewald", total.clone().data_ptr()), got, 1e-12)
            e.close()
    eng.close()


@pytest.mark.parametrize("n_mol,n_ions", [(4096, 300), (512, 40)])
def test_mixed_topology_overlap_rule(n_mol, n_ions):
    """The overlap rule (ewalds.jl:359-360: r² < 0.5 with q_a q_b < 0 zeroes the molecule's whole real-space row) on a mixed
    topology: an anion moved onto a cation, and a water hydrogen onto an anion — potential() (cell mode and tile mode, padded
    evaluation copy + overlap rows through k_move) and mmc_energy_all against the oracle."""
    from metropolismontecarlo_b200.energy import water_engine
    ms = systems.water_ion_mixture(n_mol, n_ions)
    ions = np.flatnonzero(ms.last_atom == ms.first_atom)
    a, b = ions[0], ions[1]                                   # charges +1, −1
    assert ms.charge[ms.first_atom[a] - 1] * ms.charge[ms.first_atom[b] - 1] < 0
    new_b = np.clip(ms.com[a] + np.array([0.3, 0.2, 0.0]), 0.0, ms.box)
    ms.coords[ms.first_atom[b] - 1] = new_b
    ms.com[b] = new_b
    c = ions[3]                                               # an anion; bring a water's first H onto it
    w = next(m for m in range(n_mol) if ms.last_atom[m] - ms.first_atom[m] == 2 and abs(m - c) > 5)
    fa = ms.first_atom[w] - 1
    shift = (ms.coords[ms.first_atom[c] - 1] + np.array([0.0, 0.25, 0.1])) - ms.coords[fa + 1]
    ms.coords[fa:fa + 3] += shift
    ms.com[w] = np.clip(ms.com[w] + shift, 0.0, ms.box)
    s = ora_system(ms)
    ew = ora_ewald(ms.box)
    want = ora.potential_ewald(s, ew, 10.0, 10.0, ms.box, 8)
    assert want.overlaps >= 4
    eng = water_engine(ms, 10.0)
    got = eng.potential("ewald")
    assert eng.last_eval_info()["pair_kernel"] == "k_pairs"
    _check_props(got, want)
    assert got.overlaps == want.overlaps
    lj, vir, qq, ov = eng.energy_all("ewald")
    for i in (a + 1, b + 1, c + 1, w + 1, 1, n_mol):
        c0, _, ov0 = ora.EwaldShort(int(i), s, ew, 10.0, ms.box)
        e0, _ = ora.LJ_poly_dU(int(i), s, 10.0, ms.box)
        assert bool(ov[i - 1]) == bool(ov0) and (rel(qq[i - 1], c0) < 1e-10 if c0 != 0 else qq[i - 1] == 0.0), i
        assert rel(lj[i - 1], e0) < 1e-10, i
    assert ov[a] and ov[b] and ov[c] and ov[w]
    eng.close()


@pytest.mark.parametrize("world,order", [(2, "lattice"), (4, "lattice"), (3, "random")])
def test_domain_decomposed_host_evaluation_emulated_ranks(world, order):
    """mmc_potential_host on sharded handles (ranks emulated as threads of one process on one GPU): every rank copies all COMs
    but only the site blocks of its own slab of the cell grid (+ one layer, + its rho(k) share) — less than the whole array for
    the lattice order, all of it for a random order — and every rank returns the Properties of the unsharded evaluation,
    bit-identical across ranks.  Afterwards the state is only partially resident: entry points that need all of it refuse."""
    import threading
    from metropolismontecarlo_b200.energy import MMCError, water_engine
    ms = systems.spce_lattice(40000)
    if order == "random":
        perm = np.random.default_rng(4).permutation(ms.n_mol)
        ms.com = ms.com[perm].copy()
        ms.coords = ms.coords.reshape(-1, 3, 3)[perm].reshape(-1, 3).copy()
    single = water_engine(ms, 10.0)
    want = single.potential("ewald")
    want_w = single.potential("wolf")
    assert single.last_eval_info()["pair_kernel"] == "k_pairs_v7"
    single.close()
    engs = [water_engine(ms, 10.0, rank=r, world=world) for r in range(world)]
    for e in engs:
        e.peer_export()
    for e in engs:
        for r, o in enumerate(engs):
            e.peer_import_ptr(r, o.peer_buffer())
    full = 24 * (ms.n_sites + ms.n_mol)
    # (ranks emulated in one process copy all COMs themselves: the COM all-gather over peer memory needs one process per GPU — its
    # kernels spin on the peers' flags, which deadlocks streams of one context — and is covered by tests/test_gpu_multi.py)
    for style, ref in (("ewald", want), ("wolf", want_w), ("ewald", want)):
        res = [None] * world

        def run(r):
            res[r] = engs[r].potential_host(ms.coords, ms.com, style)
        th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
        for t in th:
            t.start()
        for t in th:
            t.join(timeout=120)
        assert all(p is not None for p in res)
        for p in res:
            _check_props(p, ref, 1e-12)
            assert p.energy == res[0].energy and p.recip == res[0].recip and p.real == res[0].real
        got = [e.last_host_bytes() for e in engs]
        if order == "lattice":
            assert max(got) < 0.85 * full and all(g >= 24 * ms.n_mol for g in got), (got, full)
        else:
            assert max(got) <= full
    with pytest.raises(MMCError):
        engs[0].potential("ewald")                      # only a slab of the sites is resident
    # a DIFFERENT state in the next call (the molecules re-ordered: the same energy, other blocks): the site blocks of the previous
    # call's slab are copied speculatively behind the COMs, the missing ones after the binning — same Properties
    perm2 = np.roll(np.arange(ms.n_mol), ms.n_mol // 3)
    com2 = ms.com[perm2].copy()
    coords2 = ms.coords.reshape(-1, 3, 3)[perm2].reshape(-1, 3).copy()
    for _ in range(2):
        res = [None] * world

        def run2(r):
            res[r] = engs[r].potential_host(coords2, com2, "ewald")
        th = [threading.Thread(target=run2, args=(r,)) for r in range(world)]
        for t in th:
            t.start()
        for t in th:
            t.join(timeout=120)
        assert all(p is not None for p in res)
        for p in res:
            _check_props(p, want, 1e-11)
            assert p.energy == res[0].energy
    engs[0].upload_positions(ms.coords, ms.com)
    _check_props(engs[0].potential("ewald"), want, 1e-12)
    for e in engs:
        e.close()


def test_intramolecular_correction_is_opt_in():
    """the intramolecular Ewald correction the reference omits (Ewald/energy.jl:1008-1021), as an opt-in flag.
    Default off = the reference's Properties; on: Properties.intra = −factor Σ_mol Σ_{a<b} q_a q_b erf(κ r)/r (oracle twin
    ora_EwaldIntra), added to energy / coulomb and, like the other Coulomb terms, /3 to the virial — for potential() on the v7
    path, the general path, a mixed topology, and a volume trial (κ = α/L' changes with the box)."""
    from metropolismontecarlo_b200.energy import water_engine
    for ms in (systems.spce_lattice(4000), systems.load_nist(4), systems.water_ion_mixture(512, 40)):
        eng = water_engine(ms, 10.0)
        s = ora_system(ms)
        kappa = systems.ALPHA / ms.box
        ref = eng.potential("ewald")
        assert ref.intra == 0.0
        eng.set_intramolecular(True)
        got = eng.potential("ewald")
        want = ora.EwaldIntra(s, kappa, systems.FACTOR, ms.box)
        assert want > 0 and rel(got.intra, want) < 1e-12       # (q_O q_H < 0 dominates: the correction raises the energy)
        # (E_intra nearly cancels E_self: compare on the scale of the terms)
        assert abs(got.energy - (ref.energy + want)) < 1e-12 * abs(ref.energy) and abs(got.coulomb - (ref.coulomb + want)) < 1e-12 * abs(ref.coulomb)
        assert abs(got.virial - (ref.virial + want / 3)) < 1e-12 * max(abs(ref.virial), abs(want))
        for f in ("lj", "real", "recip", "self_"):
            assert getattr(got, f) == getattr(ref, f)
        assert eng.potential("wolf").intra == 0.0           # an Ewald-sum term only
        if ms.n_mol == 4000:
            L2 = ms.box * 1.02
            v = eng.volume_trial(L2, systems.ALPHA / L2, "ewald")
            eng.volume_reject()
            assert rel(v.intra, ora.EwaldIntra(s, systems.ALPHA / L2, systems.FACTOR, ms.box)) < 1e-12      # rigid shift: same r_ab, new κ
        eng.set_intramolecular(False)
        assert eng.potential("ewald").energy == ref.energy
        eng.close()


@pytest.mark.parametrize("nk,k_sq_max", [(12, 145), (7, 50), (16, 257)])
def test_large_k_sets_rebuild_and_k_range_sharding(nk, k_sq_max):
    """SURVEY §8d "E2" / north star "k-vector ranges split per GPU": ρ(k) rebuilds beyond the reference's nk = 5 (a converged
    Ewald sum at κ·r_cut ≈ 3.2 needs nk ≈ 12: 3.6 k k-vectors) through k_rhok_big — RecipLong energy and every ρ(k) element
    against the oracle (Ewald/ewalds.jl:538-604, whose arithmetic is general in nk), the full potential(), a volume trial (the
    kernel scales the resident sites itself), and the sharded evaluation with the k-space work split by SITES and by K-RANGES
    (emulated ranks), both equal to the unsharded one."""
    from metropolismontecarlo_b200.energy import Engine, MMCError
    ms = systems.spce_lattice(4000)
    kappa = 0.32
    ew = ora.Ewald(kappa, nk, k_sq_max, systems.FACTOR, ms.box)
    eng = Engine()
    eng.upload_system(ms, 10.0, 10.0)
    n = eng.PrepareEwaldVariables(kappa, nk, k_sq_max)
    assert n == ew.nkvecs and (nk != 12 or n > 3000)
    k, c = eng.kvectors()
    assert np.array_equal(k, ew.kxyz) and np.allclose(c, ew.cfac, rtol=1e-14, atol=0)
    e0 = ora.RecipLong(ew, ms.coords, ms.charge, ms.box)
    assert rel(eng.RecipLong(), e0) < 1e-11
    old, new = eng.rhok()
    want_rho = ew.sum_new[:, 0] + 1j * ew.sum_new[:, 1]
    assert np.abs(old - want_rho).max() < 1e-12 * 0.8476 * ms.n_sites and np.array_equal(old, new)
    s = ora_system(ms)
    want = ora.potential_ewald(s, ora.Ewald(kappa, nk, k_sq_max, systems.FACTOR, ms.box), 10.0, 10.0, ms.box, 8)
    got = eng.potential("ewald")
    _check_props(got, want)
    # volume trial == a fresh engine on the scaled configuration (volumeChange.jl:62-80)
    L2 = ms.box * 1.015
    v = eng.volume_trial(L2, kappa, "ewald")
    eng.volume_reject()
    ms2 = ms.copy()
    ms2.box = L2
    ms2.com = ms.com * (L2 / ms.box)
    ms2.coords = ms.coords + np.repeat(ms2.com - ms.com, 3, axis=0)
    want2 = ora.potential_ewald(ora_system(ms2), ora.Ewald(kappa, nk, k_sq_max, systems.FACTOR, L2), 10.0, 10.0, L2, 8)
    _check_props(v, want2)
    if nk > 8:          # the per-move kernels are sized for nk <= 8 and say so
        with pytest.raises(MMCError):
            eng.trial_move(1, ms.com[0], ms.coords[:3], "ewald")
    eng.close()
    for world, kshard in ((2, 0), (3, 1), (4, 1)):
        engs = []
        for r in range(world):
            e = Engine(rank=r, world=world)
            e.upload_system(ms, 10.0, 10.0)
            e.PrepareEwaldVariables(kappa, nk, k_sq_max)
            e.debug_set("rhok_kshard", kshard)
            engs.append(e)
        for e in engs:
            e.peer_export()
        for e in engs:
            for r, o in enumerate(engs):
                e.peer_import_ptr(r, o.peer_buffer())
        for rep in range(2):
            for e in engs:
                e.potential_sharded_begin("ewald")
            props = [e.potential_sharded_end() for e in engs]
            for p in props:
                _check_props(p, want, 1e-11)
                assert p.energy == props[0].energy and p.recip == props[0].recip
        r0, _ = engs[-1].rhok()
        assert np.abs(r0 - want_rho).max() < 1e-12 * 0.8476 * ms.n_sites
        for e in engs:
            e.close()
