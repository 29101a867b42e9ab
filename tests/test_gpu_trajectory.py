"""Trajectory parity: the library's host driver (mmc_loop_run: one fused trial-move launch +
accept/reject per move) against the oracle's restatement of Ewald/main.jl Loop(), both fed the
same stream of uniforms in the reference's draw order.  The stream is the
reference's own: Julia's MersenneTwister seeded with 11234 (Ewald/main.jl:36), reproduced by
mmc_julia_rand and, independently, by oracle/julia_rng.py (pinned to Julia's printed values).

Bar (north star): identical accept/reject sequence for the first 10^4 moves; per-move deltas
within 1e-10 relative of the energy scale; Σ accepted deltas == fresh potential() (the reference's
own block invariant, Poly/main.jl:232-235, tolerance 1e-3 there)."""
import numpy as np
import pytest

from metropolismontecarlo_b200 import systems
from metropolismontecarlo_b200.energy import LoopParams, julia_rand
from oracle import oracle as ora
from tests.util import ora_ewald, ora_system, rel

pytestmark = pytest.mark.gpu

N_MOVES = 10_000


@pytest.mark.parametrize("style,sid", [("ewald", 0), ("wolf", 1)])
def test_accept_reject_sequence_coord750(style, sid):
    from metropolismontecarlo_b200.energy import water_engine
    ms = systems.load_nist(4)
    rc, T = 10.0, 298.15
    u = julia_rand(11234, 8 * N_MOVES)              # Random.seed!(11234), Ewald/main.jl:36
    from oracle.julia_rng import julia_rand as ora_rand
    assert np.array_equal(u[:2000], ora_rand(11234, 2000))
    # ---- oracle
    s = ora_system(ms)
    ew = ora_ewald(ms.box)
    p0 = ora.potential_ewald(s, ew, rc, rc, ms.box, 4) if sid == 0 else ora.potential_wolf(s, ew, rc, rc, ms.box, 4)
    prm = ora.LoopParams(T, 0.316555789, 0.05, 0.5, 1.0, rc, rc, ms.box, sid, 1)
    quat_o = ms.quat.copy()
    rc_o, acc_o, del_o, st_o = ora.loop(s, ew, ms.db, quat_o, prm, u, N_MOVES, p0.energy, p0.virial)
    assert rc_o == 0
    # ---- engine
    eng = water_engine(ms, rc)
    g0 = eng.potential(style)
    assert rel(g0.energy, p0.energy) < 1e-10
    com, quat = ms.com.copy(), ms.quat.copy()
    rc_g, acc_g, del_g, st_g = eng.loop_run(LoopParams(T, 0.316555789, 0.05, 0.5, 1.0, sid, 1), com, quat, ms.db,
                                            u, N_MOVES, g0.energy, g0.virial)
    assert rc_g == 0 and st_g.n_moves == N_MOVES
    assert np.array_equal(acc_g, acc_o), f"first divergence at move {int(np.argmax(acc_g != acc_o))}"
    assert st_g.uniforms_used == st_o.uniforms_used
    assert st_g.n_accepted == st_o.n_accepted and st_g.rot_accept == st_o.rot_accept
    assert st_g.dr_max == st_o.dr_max and st_g.dphi_max == st_o.dphi_max
    scale = max(np.abs(del_o).max(), 1.0)
    assert np.abs(del_g - del_o).max() < 1e-10 * abs(p0.energy) and np.abs(del_g - del_o).max() < 1e-6 * scale
    assert np.array_equal(com, s.com) and np.array_equal(quat, quat_o)
    coords, com_d = eng.download_system()
    assert np.array_equal(coords, s.coords) and np.array_equal(com_d, s.com)
    # running total == fresh recompute, on both sides
    fresh = eng.potential(style)
    assert abs(st_g.total_energy - fresh.energy) < 1e-3
    assert rel(st_g.total_energy, fresh.energy) < 1e-10
    assert rel(st_g.total_energy, st_o.total_energy) < 1e-11
    if sid == 0:   # resident ρ(k) after ~5·10³ delta updates == oracle's
        old_before = ew.sum_old[:, 0] + 1j * ew.sum_old[:, 1]
        # potential() above rebuilt ρ(k); compare against a fresh oracle rebuild too
        ew2 = ora_ewald(ms.box)
        ora.RecipLong(ew2, s.coords, s.charge, ms.box)
        assert np.abs(eng.rhok()[0] - (ew2.sum_old[:, 0] + 1j * ew2.sum_old[:, 1])).max() < 1e-9
        assert np.abs(old_before - (ew2.sum_old[:, 0] + 1j * ew2.sum_old[:, 1])).max() < 1e-8
    print(style, "accepted", st_g.n_accepted, "of", N_MOVES, "overlaps", st_g.n_overlap)
    eng.close()


def test_rhok_resident_after_moves_matches_rebuild():
    from metropolismontecarlo_b200.energy import water_engine
    ms = systems.load_nist(1)
    eng = water_engine(ms, 9.0)
    g0 = eng.potential("ewald")
    u = np.random.default_rng(3).random(20000)
    com, quat = ms.com.copy(), ms.quat.copy()
    rc, acc, delta, st = eng.loop_run(LoopParams(298.15, 0.3, 0.05, 0.5, 1.0, 0, 1), com, quat, ms.db, u, 2000,
                                      g0.energy, g0.virial)
    assert rc == 0 and 0 < st.n_accepted < 2000
    resident = eng.rhok()[0].copy()
    eng.RecipLong()
    assert np.abs(resident - eng.rhok()[0]).max() < 1e-9
    eng.close()


def test_monatomic_lj_trajectory():
    """Monatomic/mainMonatomic.jl:373-413 with 1000 atoms, rho*=0.75, T*=1, dr_max = L/30."""
    from metropolismontecarlo_b200.energy import Engine
    at = systems.lj_lattice(1000, 0.75, 2.5)
    u = julia_rand(11234, 5 * N_MOVES)              # Random.seed!(11234), mainMonatomic.jl:15
    e0, v0 = ora.potential_atoms(at.r, at.eps, at.sig, at.box, at.r_cut, 4)
    r_o = at.r.copy()
    rc_o, acc_o, del_o, st_o = ora.loop_atoms(r_o, at.eps, at.sig, at.box, at.r_cut, 1.0, at.box / 30, u, N_MOVES, e0, v0)
    eng = Engine()
    eng.upload_atoms(at)
    g0 = eng.potential("atoms")
    assert rel(g0.energy, e0) < 1e-12
    r_g = at.r.copy()
    rc_g, acc_g, del_g, st_g = eng.loop_run_atoms(1.0, at.box / 30, r_g, u, N_MOVES, g0.energy, g0.virial)
    assert rc_o == 0 and rc_g == 0
    assert np.array_equal(acc_g, acc_o)
    assert np.abs(del_g - del_o).max() < 1e-10 * max(1.0, np.abs(del_o).max())
    assert np.array_equal(r_g, r_o) and np.array_equal(eng.download_atoms(), r_o)
    assert rel(st_g.total_energy, eng.potential("atoms").energy) < 1e-10
    eng.close()


@pytest.mark.parametrize("style,sid,cluster", [("ewald", 0, 8), ("wolf", 1, 8), ("ewald", 0, 1), ("wolf", 1, 4)])
def test_device_loop_matches_oracle_and_host_driver(style, sid, cluster):
    """mmc_loop_run_device: the whole block of 10^4 moves in one launch (state in shared memory).
    Same accept/reject record, uniform consumption and step-size adaptation as the oracle's Loop();
    deltas to rounding; the state it leaves in the library equals the host-driven run's."""
    from metropolismontecarlo_b200.energy import water_engine
    ms = systems.load_nist(4)
    rc, T = 10.0, 298.15
    u = julia_rand(11234, 8 * N_MOVES)
    s = ora_system(ms)
    ew = ora_ewald(ms.box)
    p0 = ora.potential_ewald(s, ew, rc, rc, ms.box, 4) if sid == 0 else ora.potential_wolf(s, ew, rc, rc, ms.box, 4)
    prm = ora.LoopParams(T, 0.316555789, 0.05, 0.5, 1.0, rc, rc, ms.box, sid, 1)
    quat_o = ms.quat.copy()
    rc_o, acc_o, del_o, st_o = ora.loop(s, ew, ms.db, quat_o, prm, u, N_MOVES, p0.energy, p0.virial)
    eng = water_engine(ms, rc)
    eng.debug_set("chain_cluster", cluster)         # 1: one CTA (k_chain); > 1: thread-block cluster (k_chainc)
    g0 = eng.potential(style)
    com, quat = ms.com.copy(), ms.quat.copy()
    rc_g, acc_g, del_g, st_g = eng.loop_run(LoopParams(T, 0.316555789, 0.05, 0.5, 1.0, sid, 1), com, quat, ms.db,
                                            u, N_MOVES, g0.energy, g0.virial, device=True)
    assert rc_g == 0 and st_g.n_moves == N_MOVES
    assert np.array_equal(acc_g, acc_o), f"first divergence at move {int(np.argmax(acc_g != acc_o))}"
    assert st_g.uniforms_used == st_o.uniforms_used
    assert st_g.n_accepted == st_o.n_accepted and st_g.rot_accept == st_o.rot_accept and st_g.n_overlap == st_o.n_overlap
    assert abs(st_g.dr_max - st_o.dr_max) < 1e-12 and abs(st_g.dphi_max - st_o.dphi_max) < 1e-12
    scale = max(np.abs(del_o).max(), 1.0)
    assert np.abs(del_g - del_o).max() < 1e-10 * abs(p0.energy) and np.abs(del_g - del_o).max() < 1e-6 * scale
    # device cos/sin may differ from libm in the last bit: positions to 1e-12 Å instead of bit equality
    assert np.abs(com - s.com).max() < 1e-12 and np.abs(quat - quat_o).max() < 1e-12
    coords, com_d = eng.download_system()
    assert np.abs(coords - s.coords).max() < 1e-11 and np.array_equal(com_d, com)
    fresh = eng.potential(style)                    # running total == fresh recompute (Poly/main.jl:232-235)
    assert rel(st_g.total_energy, fresh.energy) < 1e-10
    assert rel(st_g.total_energy, st_o.total_energy) < 1e-11
    # the per-move entry points continue from the state the block left behind
    i = 7
    t = eng.trial_move(i, com[i - 1] + 0.05, coords[3 * (i - 1):3 * i] + 0.05, style)
    e_old = ora.LJ_poly_dU(i, s, rc, ms.box)[0] + ora.EwaldShort(i, s, ew, rc, ms.box)[0]
    assert rel(t.lj_old + t.qq_old, e_old) < 1e-9
    eng.reject()
    eng.close()


def test_device_loop_short_stream_and_small_system():
    """The uniform stream ends inside a move (every draw position of the last moves is tried): return code 1, the move in
    progress never happened — nothing committed or recorded, counters taken back, uniforms_used = position at the start
    of that move — identically in the oracle, the host driver and the device block (no fabricated draw is ever
    acted upon).  Resuming there with the rest of the stream reproduces the one-shot record.  100-molecule box."""
    from metropolismontecarlo_b200.energy import water_engine
    ms = systems.load_nist(1)
    u = np.random.default_rng(5).random(4000)
    prm = LoopParams(298.15, 0.3, 0.05, 0.5, 1.0, 0, 0)
    full = {}
    for device in (False, True):
        eng = water_engine(ms, 9.0)
        g0 = eng.potential("ewald")
        com, quat = ms.com.copy(), ms.quat.copy()
        rc, acc, delta, st = eng.loop_run(prm, com, quat, ms.db, u, 300, g0.energy, g0.virial, device=device)
        assert rc == 0
        full[device] = (acc.copy(), com.copy(), quat.copy(), st.n_accepted)
        eng.close()
    assert np.array_equal(full[False][0], full[True][0])
    for n_short in (1500, 1501, 1502, 1503, 1504, 1505, 1506, 3):
        outs = []
        s = ora_system(ms)
        ew = ora_ewald(ms.box)
        p0 = ora.potential_ewald(s, ew, 9.0, 9.0, ms.box)
        q_o = ms.quat.copy()
        oprm = ora.LoopParams(298.15, 0.3, 0.05, 0.5, 1.0, 9.0, 9.0, ms.box, 0, 0)
        rc_o, acc_o, _, st_o = ora.loop(s, ew, ms.db, q_o, oprm, u[:n_short], 2000, p0.energy, p0.virial)
        for device in (False, True):
            eng = water_engine(ms, 9.0)
            g0 = eng.potential("ewald")
            com, quat = ms.com.copy(), ms.quat.copy()
            rc, acc, delta, st = eng.loop_run(prm, com, quat, ms.db, u[:n_short], 2000, g0.energy, g0.virial, device=device)
            outs.append((rc, acc.copy(), st.n_moves, st.uniforms_used, st.n_accepted, com.copy(), st.trans_attempt + st.rot_attempt))
            assert rc == 1 and rc_o == 1
            assert st.n_moves == st_o.n_moves and st.uniforms_used == st_o.uniforms_used <= n_short
            assert st.n_accepted == st_o.n_accepted and np.array_equal(acc[:st.n_moves], acc_o[:st.n_moves])
            assert st.trans_attempt == st_o.trans_attempt and st.rot_attempt == st_o.rot_attempt
            assert st.trans_attempt + st.rot_attempt == st.n_moves
            assert np.abs(com - s.com).max() < 1e-12
            # the resident state is the state after n_moves complete moves: a fresh evaluation equals the running total
            assert rel(eng.potential("ewald").energy, st.total_energy) < 1e-10
            n_a = st.n_moves
            if n_a < 300 and n_a % ms.n_mol == 0 and n_a > 0:       # resume at a sweep boundary with the rest of the stream
                rc2, acc2, _, st2 = eng.loop_run(prm, com, quat, ms.db, u[st.uniforms_used:], 300 - n_a, st.total_energy,
                                                 st.total_virial, device=device)
                assert rc2 == 0 and np.array_equal(np.concatenate([acc[:n_a], acc2]), full[device][0])
            eng.close()
        assert outs[0][2:5] == outs[1][2:5] and outs[0][6] == outs[1][6]
        assert np.array_equal(outs[0][1], outs[1][1])
        assert np.abs(outs[0][5] - outs[1][5]).max() < 1e-12


def _resite(ms, keep):
    """A uniform system with other site counts, derived from an SPC/E box: keep = list of (source site, charge scale)
    per molecule; a fourth site copies H2 with a small opposite charge so the box stays neutral-ish (the engine does
    not need neutrality).  Body frames and COMs stay those of the water molecules."""
    n = ms.n_mol
    S = len(keep)
    idx = np.array([[3 * m + a for a, _ in keep] for m in range(n)]).reshape(-1)
    sc = np.tile(np.array([s for _, s in keep]), n)
    out = ms.copy()
    out.coords = ms.coords[idx].copy()
    out.charge = ms.charge[idx] * sc
    out.atype = ms.atype[idx].copy()
    out.db = ms.db[idx].copy()
    out.first_atom = (np.arange(n) * S + 1).astype(np.int64)
    out.last_atom = (np.arange(n) * S + S).astype(np.int64)
    return out


@pytest.mark.parametrize("keep,style,sid", [
    ([(0, 1.0)], "lj", 2),                                        # one LJ site per molecule
    ([(0, 1.0), (1, 2.0)], "wolf", 1),                            # two sites (O, H with doubled charge)
    ([(0, 1.0), (1, 1.0), (2, 0.5), (2, 0.5)], "ewald", 0),       # four sites (second H split in two)
])
def test_device_loop_other_site_counts_match_host_driver(keep, style, sid):
    """k_chains<S> / k_chain<S> for S = 1, 2, 4 against the per-move protocol (which is oracle-tested)."""
    from metropolismontecarlo_b200.energy import water_engine
    ms = _resite(systems.load_nist(1), keep)
    u = julia_rand(11234, 8 * 3000)
    outs = []
    for device, cluster in ((False, 1), (True, 1), (True, 8)):
        eng = water_engine(ms, 9.0)
        eng.debug_set("chain_cluster", cluster)
        g0 = eng.potential(style)
        com, quat = ms.com.copy(), ms.quat.copy()
        rc, acc, delta, st = eng.loop_run(LoopParams(298.15, 0.4, 0.08, 0.5, 1.0, sid, 1), com, quat, ms.db, u, 3000,
                                          g0.energy, g0.virial, device=device)
        assert rc == 0
        fresh = eng.potential(style)
        assert rel(st.total_energy, fresh.energy) < 1e-9
        outs.append((acc.copy(), delta.copy(), st.uniforms_used, st.dr_max, com.copy()))
        eng.close()
    for o in outs[1:]:
        assert np.array_equal(o[0], outs[0][0])
        assert o[2] == outs[0][2] and abs(o[3] - outs[0][3]) < 1e-12
        assert np.abs(o[1] - outs[0][1]).max() < 1e-6 * max(1.0, np.abs(outs[0][1]).max())
        assert np.abs(o[4] - outs[0][4]).max() < 1e-11


@pytest.mark.parametrize("n_atoms,dr_div", [(1000, 30.0), (4096, 150.0)])
def test_monatomic_device_loop(n_atoms, dr_div):
    """mmc_loop_run_atoms_device (k_chain_atoms: atoms sliced over an 8-SM cluster, FP32 gate + FP64 evaluation)
    against the oracle's loop: identical accept/reject record and positions; dr_max = L/150 gives ~50 % acceptance."""
    from metropolismontecarlo_b200.energy import Engine
    at = systems.lj_lattice(n_atoms, 0.75, 2.5)
    u = julia_rand(11234, 5 * N_MOVES)
    e0, v0 = ora.potential_atoms(at.r, at.eps, at.sig, at.box, at.r_cut, 4)
    r_o = at.r.copy()
    rc_o, acc_o, del_o, st_o = ora.loop_atoms(r_o, at.eps, at.sig, at.box, at.r_cut, 1.0, at.box / dr_div, u, N_MOVES, e0, v0)
    eng = Engine()
    eng.upload_atoms(at)
    g0 = eng.potential("atoms")
    r_g = at.r.copy()
    rc_g, acc_g, del_g, st_g = eng.loop_run_atoms(1.0, at.box / dr_div, r_g, u, N_MOVES, g0.energy, g0.virial, device=True)
    assert rc_o == 0 and rc_g == 0 and st_g.n_moves == N_MOVES
    assert np.array_equal(acc_g, acc_o), f"first divergence at move {int(np.argmax(acc_g != acc_o))}"
    assert st_g.uniforms_used == st_o.uniforms_used and st_g.n_accepted == st_o.n_accepted
    assert np.abs(del_g - del_o).max() < 1e-10 * max(1.0, np.abs(del_o).max())
    assert np.array_equal(r_g, r_o) and np.array_equal(eng.download_atoms(), r_o)
    assert rel(st_g.total_energy, eng.potential("atoms").energy) < 1e-10
    print("atoms", n_atoms, "accepted", st_g.n_accepted)
    eng.close()


def test_device_loop_config_d_4000_molecules():
    """4000 SPC/E molecules do not fit one SM: k_chains keeps a slice per CTA (the resident arrays in L2 are the truth for
    the molecule about to move).  Same record as the per-move protocol; running total == fresh potential()."""
    from metropolismontecarlo_b200.energy import water_engine
    ms = systems.spce_lattice(4000)
    u = julia_rand(11234, 8 * 6000)
    outs = []
    for device in (False, True):
        eng = water_engine(ms, 10.0)
        g0 = eng.potential("ewald")
        com, quat = ms.com.copy(), ms.quat.copy()
        rc, acc, delta, st = eng.loop_run(LoopParams(298.15, 0.316555789, 0.05, 0.5, 1.0, 0, 1), com, quat, ms.db, u, 6000,
                                          g0.energy, g0.virial, device=device)
        assert rc == 0 and st.n_moves == 6000
        fresh = eng.potential("ewald")
        assert rel(st.total_energy, fresh.energy) < 1e-10
        coords, com_d = eng.download_system()
        assert np.array_equal(com_d, com)
        outs.append((acc.copy(), delta.copy(), st.uniforms_used, st.dr_max, com.copy(), coords.copy()))
        eng.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and outs[0][2] == outs[1][2] and abs(outs[0][3] - outs[1][3]) < 1e-12
    assert np.abs(outs[0][1] - outs[1][1]).max() < 1e-6 * max(1.0, np.abs(outs[0][1]).max())
    assert np.abs(outs[0][4] - outs[1][4]).max() < 1e-11 and np.abs(outs[0][5] - outs[1][5]).max() < 1e-11


def test_concurrent_replicas_on_one_gpu_match_their_oracle_runs():
    """Per-move paths do not shard (replicas only), and one chain uses 8 of 148 SMs: several independent chains run
    at once on ONE GPU (one handle, stream and host thread each).  Every replica's accept/reject record, stream consumption and
    final COMs must equal the oracle's Loop() on that replica's own uniform stream."""
    import threading
    from metropolismontecarlo_b200.energy import water_engine
    ms = systems.load_nist(4)
    n_rep, n_moves = 6, 3000
    us = [julia_rand(11234 + r, 8 * n_moves) for r in range(n_rep)]
    engs = [water_engine(ms, 10.0) for _ in range(n_rep)]
    p0 = [e.potential("ewald") for e in engs]
    res = [None] * n_rep

    def run(r):
        com, quat = ms.com.copy(), ms.quat.copy()
        out = engs[r].loop_run(LoopParams(298.15, 0.316555789, 0.05, 0.5, 1.0, 0, 1), com, quat, ms.db, us[r], n_moves,
                               p0[r].energy, p0[r].virial, device=True)
        res[r] = (out, com)
    th = [threading.Thread(target=run, args=(r,)) for r in range(n_rep)]
    for t_ in th:
        t_.start()
    for t_ in th:
        t_.join(timeout=300)
    for r in range(n_rep):
        s = ora_system(ms)
        ew = ora_ewald(ms.box)
        w0 = ora.potential_ewald(s, ew, 10.0, 10.0, ms.box)
        prm = ora.LoopParams(298.15, 0.316555789, 0.05, 0.5, 1.0, 10.0, 10.0, ms.box, 0, 1)
        rc_o, acc_o, del_o, st_o = ora.loop(s, ew, ms.db, ms.quat.copy(), prm, us[r], n_moves, w0.energy, w0.virial)
        (rc_g, acc_g, del_g, st_g), com = res[r]
        assert rc_g == 0 and rc_o == 0 and np.array_equal(acc_g, acc_o), r
        assert st_g.uniforms_used == st_o.uniforms_used and st_g.n_accepted == st_o.n_accepted
        assert np.abs(com - s.com).max() < 1e-12
        assert rel(st_g.total_energy, engs[r].potential("ewald").energy) < 1e-9
    assert len({tuple(res[r][0][1][:200]) for r in range(n_rep)}) == n_rep       # different streams, different records
    for e in engs:
        e.close()
