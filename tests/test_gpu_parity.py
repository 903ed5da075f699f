"""Parity of the CUDA path (through the C ABI) against the oracle -- runs on the B200 box (-m gpu).

Tier 1 = lock-step: each GPU phase is fed the oracle's state at the start of that step.
  integer keys (cell index, flat index, loss flags, alive counts) ............ bit-exact
  per-ring z, v after one step, FAST arithmetic ............................... rel <= 1e-14
  per-ring z, v after one step, EXACT arithmetic ............................... bit-exact
  node field from identical potentials ......................................... bit-exact
  RHS (= -rho/eps0), fp64 accumulation ......................................... rel-L2 <= 1e-12
  RHS, fixed-point accumulation (2^-F quantisation of the weights) ............. rel-L2 <= 1e-11, bitwise reproducible
  phi_self / phi_trap (direct solver vs LU oracle) ............................. rel-L2 <= 1e-10
Tier 2 = free-running C1 for 175 steps (5 plasma periods): counts equal, z rel-L2 <= 1e-7 (e-) / 1e-9 (pbar),
  RHS rel-L2 <= 1e-5 (solver-level differences are amplified ~1e5x over that horizon, SURVEY A-8).
"""
import importlib
import json
import math
import os

import numpy as np
import pytest

from conftest import GOLDEN, rel_l2
from oracle import port, ref

ptp = importlib.import_module("pic-trapped-plasma_b200")
pytestmark = pytest.mark.gpu

KAT = json.load(open(os.path.join(GOLDEN, "trap_kat.json")))


@pytest.fixture(scope="module")
def trap():
    t = ptp.default_trap()
    yield t
    t.close()


def _fresh_c1(kat, deposit_mode=ptp.PTP_DEPOSIT_FP64, arith=ptp.PTP_ARITH_FAST):
    t = ptp.default_trap()
    t.set_deposit_mode(deposit_mode)
    t.set_arith_mode(arith)
    el = ptp.Plasma(t, "Electrons", ptp.massE, -ptp.ePos)
    ap = ptp.Plasma(t, "Antiprotons", ptp.massP, -ptp.ePos)
    for tag, p in (("e", el), ("p", ap)):
        p.upload(kat[f"{tag}_r0"], kat[f"{tag}_z0"], kat[f"{tag}_v0"], float(kat[f"{tag}_chargeMacro"]))
        assert p.macroChargeDensity == float(kat[f"{tag}_mcd"])
    return t, el, ap


def _by_id(p):
    r, z, v, ids = p.download()
    o = np.argsort(ids)
    return r[o], z[o], v[o], ids[o]


# ------------------------------------------------------------------------------------------ solver (a6, a7, a8)
def test_trap_potential_matches_reference(trap, c1_kat):
    phi = trap.phi()
    assert rel_l2(phi, c1_kat["phi_trap"]) < 1e-10
    assert phi[293] == pytest.approx(KAT["phi_r0_k293"], rel=1e-10)
    # extractTrapLaplacian convention: A phi - RHS ~ 0
    pt = port.default_trap()
    rhs = pt.wall_rhs()
    assert np.linalg.norm(trap.apply(phi) - rhs) / np.linalg.norm(rhs) < 1e-13
    assert np.array_equal(trap.wallPotential(), pt.wall_potential())
    x = np.random.default_rng(0).standard_normal(trap.G)
    assert rel_l2(trap.apply(x), pt.apply(x)) < 1e-15
    assert rel_l2(trap.solve(x), pt.solve(x)) < 1e-10
    pt.close()


def test_set_potential_every_step_protocol(trap, c1_kat):
    pt = port.default_trap()
    for v in (-51.0, -63.5, -70.0):
        trap.setPotential(1, v)
        pt.set_potential(1, v)
        assert rel_l2(trap.phi(), pt.phi) < 1e-10
    assert rel_l2(trap.phi(), c1_kat["phi_trap"]) < 1e-10
    pt.close()


def test_steps_leave_trap_potential_untouched(c1_kat):
    """Regression: the fused inverse-transform + node-field kernel once wrote E = 0 one element in front of the node-field
    array (column -1 of its first tile), which is the last node of phi_trap when the two allocations are adjacent."""
    t, el, ap = _fresh_c1(c1_kat)
    el.solvePoisson()
    ap.solvePoisson()
    before = t.phi()
    t.movePlasmas(float(c1_kat["dt"]), 3)
    assert np.array_equal(t.phi(), before)
    t.close()


def test_self_potential_of_reference_rhs(trap, c1_kat):
    for tag in ("e", "p"):
        assert rel_l2(trap.solve(c1_kat[f"{tag}_rhs0"]), c1_kat[f"{tag}_phi0"]) < 1e-10


def test_sor_cross_check_small_grid():
    el = [ptp.Electrode(0.01, 5.0), ptp.Electrode(0.02, -40.0), ptp.Electrode(0.015, 3.0)]
    t = ptp.PenningTrap(0.02, el, [0.002, 0.001], 57, 9)
    direct = t.phi()
    t.set_solver(ptp.PTP_SOLVER_SOR, 1e-13, 40000)
    t.set_phi(np.zeros(t.G))
    t.solveLaplace()
    assert rel_l2(t.phi(), direct) < 1e-9
    t.close()


# ------------------------------------------------------------------------------------------ deposit (a5)
@pytest.mark.parametrize("mode,tol", [(ptp.PTP_DEPOSIT_FP64, 1e-12), (ptp.PTP_DEPOSIT_FIXED64, 1e-11)])
def test_deposit_lockstep(c1_kat, mode, tol):
    t, el, ap = _fresh_c1(c1_kat, mode)
    pt = port.default_trap()
    for tag, p in (("e", el), ("p", ap)):
        p.updateRHS()
        assert rel_l2(p.rhs(), c1_kat[f"{tag}_rhs0"]) < tol
        # integer keys: bit-exact
        op = pt.plasma(tag, p.mass, p.charge)
        op.set_rings(c1_kat[f"{tag}_r0"], c1_kat[f"{tag}_z0"], c1_kat[f"{tag}_v0"], float(c1_kat[f"{tag}_chargeMacro"]))
        k_o, idx_o = op.cell_index()
        _, _, _, ids = p.download()
        k_g, idx_g = p.cell_index()
        assert np.array_equal(k_g, k_o[ids]) and np.array_equal(idx_g, idx_o[ids])
        assert p.getNumMacro() == len(ids) == 4001
    pt.close()
    t.close()


def test_deposit_reproduces_reference_data_file():
    """Diagnostics/Charge_Density-0.txt through the CUDA deposit (the reference's only golden data)."""
    d = np.load(os.path.join(GOLDEN, "fixture_rings_r0.npz"))
    fixture = np.loadtxt(os.path.join(GOLDEN, "charge_density_0.txt"))
    t = ptp.default_trap()
    for mode, tol in ((ptp.PTP_DEPOSIT_FP64, 1e-12), (ptp.PTP_DEPOSIT_FIXED64, 1e-11)):
        t.set_deposit_mode(mode)
        p = ptp.Plasma(t, "Electrons", ptp.massE, -ptp.ePos)
        z = d["z"]
        p.upload(np.zeros(len(z), np.int32), z, np.zeros(len(z)), float(d["chargeMacro"]))
        p.updateRHS()
        assert rel_l2(p.rhs()[275:311] * ptp.epsilon / ptp.ePos, fixture[:, 1]) < tol
    t.close()


# ------------------------------------------------------------------------------------------ node field (a4)
def test_node_field_bit_exact(c1_kat):
    t, el, ap = _fresh_c1(c1_kat)
    t.set_phi(c1_kat["phi_trap"])
    el.set_self_potential(c1_kat["e_phi0"])
    ap.set_self_potential(c1_kat["p_phi0"])
    assert np.array_equal(t.enodes(), c1_kat["enodes0"])
    t.close()


# ------------------------------------------------------------------------------------------ push (a2, a3) + step (a1)
@pytest.mark.parametrize("arith", [ptp.PTP_ARITH_FAST, ptp.PTP_ARITH_EXACT])
@pytest.mark.parametrize("mode", [ptp.PTP_DEPOSIT_FP64, ptp.PTP_DEPOSIT_FIXED64])
def test_step_lockstep(c1_kat, arith, mode):
    t, el, ap = _fresh_c1(c1_kat, mode, arith)
    t.set_phi(c1_kat["phi_trap"])
    el.set_self_potential(c1_kat["e_phi0"])
    ap.set_self_potential(c1_kat["p_phi0"])
    dt = float(c1_kat["dt"])
    t.push_deposit(dt)
    for tag, p in (("e", el), ("p", ap)):
        r, z, v, ids = _by_id(p)
        assert np.array_equal(ids, np.arange(4001)) and np.array_equal(r, c1_kat[f"{tag}_r0"])
        z1, v1 = c1_kat[f"{tag}_z1"], c1_kat[f"{tag}_v1"]
        if arith == ptp.PTP_ARITH_EXACT:
            assert np.array_equal(z, z1) and np.array_equal(v, v1)
        else:
            assert np.max(np.abs(z - z1) / np.abs(z1)) <= 1e-14
            assert np.max(np.abs(v - v1)) <= 1e-14 * np.max(np.abs(v1)) and rel_l2(v, v1) <= 1e-14
        assert rel_l2(p.rhs(), c1_kat[f"{tag}_rhs1"]) < (1e-12 if mode == ptp.PTP_DEPOSIT_FP64 else 1e-11)
    t.solve_fields()
    for tag, p in (("e", el), ("p", ap)):
        assert rel_l2(p.selfPotential(), c1_kat[f"{tag}_phi1"]) < 1e-10
    assert rel_l2(t.enodes(), c1_kat["enodes1"]) < 1e-9
    t.close()


def test_free_running_175_steps(c1_kat):
    t, el, ap = _fresh_c1(c1_kat)
    el.solvePoisson()
    ap.solvePoisson()
    assert rel_l2(el.selfPotential(), c1_kat["e_phi0"]) < 1e-10
    t.movePlasmas(float(c1_kat["dt"]), 175)
    assert el.getNumMacro() == KAT["c1_count_e_175"] and ap.getNumMacro() == KAT["c1_count_p_175"]
    _, ze, _, _ = _by_id(el)
    _, zp, _, _ = _by_id(ap)
    assert rel_l2(ze, c1_kat["e_z175"]) < 1e-7 and rel_l2(zp, c1_kat["p_z175"]) < 1e-9
    assert rel_l2(el.rhs(), c1_kat["e_rhs175"]) < 1e-5
    # potential energy against the oracle's getPotentialEnergy on the same state (Source/Plasma.cpp:244-252)
    pe = el.getPotentialEnergy() + ap.getPotentialEnergy()
    assert pe == pytest.approx(KAT["c1_PE_175"], rel=1e-8)
    t.close()


def test_driver_d_loss_counts(c1_kat):
    """Diagnostics/D) Useless Boundary Test.txt:118-136: per-step setPotential + e-kick losses, integer KAT."""
    gold = json.load(open(os.path.join(GOLDEN, "driver_d_counts.json")))
    dt = gold["dt"]

    def changed_voltage(Vi, Vf, duration, compression, tt):
        return (Vf - Vi) * (1 + math.exp(-(tt - duration / 2) * compression * 2 / duration)) ** -1 + Vi

    t, el, ap = _fresh_c1(c1_kat)
    el.solvePoisson()
    ap.solvePoisson()
    counts = [[el.getNumMacro(), ap.getNumMacro()]]

    def step():
        t.movePlasmas(dt)
        counts.append([el.getNumMacro(), ap.getNumMacro()])

    i = 1
    while i * dt <= 10e-9:
        t.setPotential(1, changed_voltage(-70, -51, 10e-9, 4.5, i * dt))
        step()
        i += 1
    t.setPotential(1, -51)
    i = 1
    while i * dt <= 80e-9:
        step()
        i += 1
    i = 1
    while i * dt <= 10e-9:
        t.setPotential(1, changed_voltage(-51, -70, 10e-9, 4.5, i * dt))
        step()
        i += 1
    assert counts == gold["counts"]
    pt = port.default_trap()
    left, right = pt.limits()
    assert el.getNumMacroCentralWell(left, right) == gold["central_well_e"]
    assert ap.getNumMacroCentralWell(left, right) == gold["central_well_p"]
    pt.close()
    # the survivors are the same rings (ids) as in the oracle run
    t.sort()
    assert el.getNumMacro() == 3965
    t.close()


def test_electrode_programme_matches_per_step_solves(c1_kat):
    """the driver-D voltage ramp run as a device-side programme (phi_trap = sum_i V_i phi_i over the electrode
    basis, one axpy per step, no host round trip) against the same ramp with setPotential + Laplace solve before every
    step: phi_trap to rounding, loss counts and surviving rings identical, the golden loss counts reproduced."""
    gold = json.load(open(os.path.join(GOLDEN, "driver_d_counts.json")))
    dt = gold["dt"]

    def changed_voltage(Vi, Vf, duration, compression, tt):
        return (Vf - Vi) * (1 + math.exp(-(tt - duration / 2) * compression * 2 / duration)) ** -1 + Vi

    ramp = []
    i = 1
    while i * dt <= 10e-9:
        ramp.append(changed_voltage(-70, -51, 10e-9, 4.5, i * dt))
        i += 1
    # reference protocol: one solve per step
    ta, ea, aa = _fresh_c1(c1_kat)
    ea.solvePoisson()
    aa.solvePoisson()
    for v in ramp:
        ta.setPotential(1, v)
        ta.movePlasmas(dt)
    # programme: basis fields, all steps queued in one call
    tb, eb, ab = _fresh_c1(c1_kat)
    eb.solvePoisson()
    ab.solvePoisson()
    tb.useElectrodeBasis()
    assert rel_l2(tb.phi(), c1_kat["phi_trap"]) < 1e-12
    sched = np.tile(np.array([0.0, -70.0, -15.0, -70.0, 0.0]), (len(ramp), 1))
    sched[:, 1] = ramp
    tb.movePlasmasProgramme(dt, sched)
    assert rel_l2(tb.phi(), ta.phi()) < 1e-13
    assert [eb.getNumMacro(), ab.getNumMacro()] == [ea.getNumMacro(), aa.getNumMacro()] == gold["counts"][len(ramp)]
    for pa, pb in ((ea, eb), (aa, ab)):
        ra, za, va, ia = _by_id(pa)
        rb, zb, vb, ib = _by_id(pb)
        assert np.array_equal(ia, ib)
        assert np.max(np.abs(za - zb) / za) < 1e-11
    # single setPotential calls go through the basis too
    tb.setPotential(1, -51.0)
    ta.setPotential(1, -51.0)
    assert rel_l2(tb.phi(), ta.phi()) < 1e-13
    ta.close()
    tb.close()


@pytest.mark.parametrize("hot", [0, 1])
def test_losses_match_oracle_ring_by_ring(hot):
    """Small odd grid, hot rings: loss flags and survivors bit-exact against the restated swap-pop loop.
    hot = 1: through the per-warp-bin form of the push kernel (ptp_plasma_set_hot) - rings in arbitrary order."""
    args = (0.02, [0.01, 0.02, 0.015], [5.0, -40.0, 3.0], [0.002, 0.001], 57, 9)
    pt = port.PortTrap(*args)
    t = ptp.PenningTrap(args[0], [ptp.Electrode(a, b) for a, b in zip(args[1], args[2])], args[3], args[4], args[5])
    t.set_arith_mode(ptp.PTP_ARITH_EXACT)
    assert rel_l2(t.phi(), pt.phi) < 1e-11
    rng = np.random.default_rng(3)
    n = 30000
    r = rng.integers(0, 9, n).astype(np.int32)
    z = rng.uniform(0.001, pt.length - 0.001, n)
    v = rng.normal(0, 4e5, n)
    op = pt.plasma("Electrons", ptp.massE, -ptp.ePos)
    op.set_rings(r, z, v, -1e-16)
    gp = ptp.Plasma(t, "Electrons", ptp.massE, -ptp.ePos)
    gp.set_hot(hot)
    gp.upload(r, z, v, -1e-16)
    op.solve_poisson()
    gp.solvePoisson()
    assert gp.is_hot() == bool(hot)
    ids_alive = np.arange(n)
    for step in range(6):
        # lock-step: oracle potentials in
        t.set_phi(pt.phi)
        gp.set_self_potential(op.self_potential)
        en = pt.enodes()
        assert np.array_equal(t.enodes(), en)
        # oracle step on (z, v) tagged with ids: run the restated loop on a copy carrying ids through r's swaps
        z_before, v_before, r_before = op.z.copy(), op.v.copy(), op.r.copy()
        pt.move_plasmas(2e-9)
        t.push_deposit(2e-9)
        t.solve_fields()
        rg, zg, vg, idg = _by_id(gp)
        # survivors of the oracle, matched by (r, z_new, v_new) multiset since swap-pop permutes them
        ko = np.lexsort((op.v, op.z, op.r))
        kg = np.lexsort((vg, zg, rg))
        assert len(zg) == op.count()
        assert np.array_equal(rg[kg], op.r[ko]) and np.array_equal(zg[kg], op.z[ko]) and np.array_equal(vg[kg], op.v[ko])
        assert rel_l2(gp.rhs(), op.rhs) < 1e-12
        assert rel_l2(gp.selfPotential(), op.self_potential) < 1e-10
    assert op.count() < n
    t.close()
    pt.close()


# ------------------------------------------------------------------------------------------ determinism / sort / large N
def _synthetic(n, trap, seed=1):
    """Synthetic load of the default plasma's shape: ~12 rows, ~37 cells around the trap centre."""
    rng = np.random.default_rng(seed)
    occ = np.array([1.94, 1.94, 1.88, 1.68, 1.30, 0.80, 0.35, 0.10, 0.017, 0.0015])
    r = np.sort(rng.choice(len(occ), size=n, p=occ / occ.sum())).astype(np.int32)
    z = 0.03405 + 0.0021 * np.clip(rng.standard_normal(n) * 0.45, -1, 1)
    v = rng.normal(0, 4.768e4, n)
    return r, z, v


def test_fixed_point_is_bitwise_reproducible(c1_kat):
    outs = []
    for ctas in (0, 7, 40):
        t = ptp.default_trap()
        t.set_deposit_mode(ptp.PTP_DEPOSIT_FIXED64)
        t.set_tuning(ctas=ctas)
        p = ptp.Plasma(t, "Electrons", ptp.massE, -ptp.ePos)
        r, z, v = _synthetic(300000, t)
        p.upload(r, z, v, -1e-18)
        p.solvePoisson()
        t.movePlasmas(float(c1_kat["dt"]), 3)
        outs.append((p.rhs(), p.selfPotential()))
        t.close()
    for rhs, phi in outs[1:]:
        assert np.array_equal(rhs, outs[0][0]) and np.array_equal(phi, outs[0][1])


@pytest.mark.parametrize("mode", [ptp.PTP_DEPOSIT_FP64, ptp.PTP_DEPOSIT_FIXED64])
def test_large_load_properties(c1_kat, mode):
    """10 M rings (BASELINE config 3 size): size-independent properties.
    total deposited weight == live ring count; sort keeps the multiset and makes cells non-decreasing;
    sorting does not change the deposit; deposit after push == stand-alone deposit of the pushed rings."""
    n = 10_000_000
    t = ptp.default_trap()
    t.set_deposit_mode(mode)
    p = ptp.Plasma(t, "Antiprotons", ptp.massP, -ptp.ePos)
    r, z, v = _synthetic(n, t, seed=5)
    p.upload(r, z, v, -1e-19)
    scale = -p.macroChargeDensity / ptp.epsilon
    p.solvePoisson()
    w0 = p.rhs() / scale
    assert abs(w0.sum() - n) < (1e-6 if mode == ptp.PTP_DEPOSIT_FP64 else 1e-3)
    dt = float(c1_kat["dt"])
    t.movePlasmas(dt, 2)
    rhs_fused = p.rhs()
    assert p.getNumMacro() == n
    p.updateRHS()                                   # K2 on the pushed rings
    rhs_alone = p.rhs()
    if mode == ptp.PTP_DEPOSIT_FIXED64:
        assert np.array_equal(rhs_fused, rhs_alone)
    else:
        assert rel_l2(rhs_fused, rhs_alone) < 1e-13
    r1, z1, v1, id1 = p.download()
    t.sort()
    r2, z2, v2, id2 = p.download()
    assert len(id2) == n
    o1, o2 = np.argsort(id1), np.argsort(id2)
    assert np.array_equal(z1[o1], z2[o2]) and np.array_equal(v1[o1], v2[o2]) and np.array_equal(r1[o1], r2[o2])
    k2, _ = p.cell_index()
    key = r2.astype(np.int64) * 100000 + k2
    assert np.all(np.diff(key) >= 0)                # sortedness by (row, cell)
    p.updateRHS()
    if mode == ptp.PTP_DEPOSIT_FIXED64:
        assert np.array_equal(p.rhs(), rhs_alone)
    else:
        assert rel_l2(p.rhs(), rhs_alone) < 1e-13
    t.sort()                                        # idempotence
    _, z3, _, id3 = p.download()
    assert np.array_equal(np.sort(id3), np.sort(id2)) and np.array_equal(np.sort(z3), np.sort(z2))
    t.close()


@pytest.mark.parametrize("r16", ["0", "1", "1x"])
def test_fine_grid_full_size_properties(c1_kat, r16, monkeypatch):
    """(r16: inverse transform of the 4096-node rows by radix-2 pass pairs "0" / by three register-resident radix-16 rounds
    that form the rows above the plasma themselves "1" (default) / radix-16 after k_thomas_expand "1x".)
    BASELINE config 5 grid (Nz = 4096, Nr = 1024: 4.2 M unknowns - far beyond what the LU oracle can factorise) through
    size-independent properties: the direct solver's phi satisfies A phi = b for the operator applied by the independent
    stencil kernel (PenningTrap::generateSparse coefficients), for a random right-hand side and for the deposit of a
    5 M-ring load placed by the device loader; the deposit conserves the ring count; the node field is the centred
    difference of the total potential; fixed-point deposits give the same grid for 148 and 37 CTAs bit for bit."""
    from bench import density_on
    monkeypatch.setenv("PTP_FFT_R16", r16[0])
    monkeypatch.setenv("PTP_FFT_FORM_ROWS", "0" if r16 == "1x" else "1")
    Nz, Nr = 4096, 1024
    el = [ptp.Electrode(0.01322, v) for v in (0, -70, -15, -70, 0)]
    t = ptp.PenningTrap(0.01488, el, [0.0005] * 4, Nz, Nr)
    n1 = Nz + 1
    rng = np.random.default_rng(11)
    b = rng.standard_normal(t.G)
    assert rel_l2(t.apply(t.solve(b)), b) < 1e-11
    wall_rhs = t.apply(t.phi())                      # = the wall RHS: zero except on the last row
    assert np.max(np.abs(wall_rhs[: (Nr - 1) * n1])) < 1e-6 * np.max(np.abs(wall_rhs))
    dens = density_on(Nz, Nr)
    p = ptp.Plasma(t, "Antiprotons", ptp.massP, -ptp.ePos)
    n, per_row = p.loadDensity(dens, 150.0, 5_000_000)
    assert n == int(per_row.sum()) and abs(n - 5_000_000) < 2000
    scale = -p.macroChargeDensity / ptp.epsilon
    rhs = p.rhs()
    assert abs(rhs.sum() / scale - n) < 1e-6 * n
    assert rel_l2(t.apply(p.selfPotential()), rhs) < 1e-10
    dt = float(c1_kat["dt"])
    t.movePlasmas(dt, 3)
    assert p.getNumMacro() == n
    rhs = p.rhs()
    phi = p.selfPotential()
    assert abs(rhs.sum() / scale - n) < 1e-6 * n
    assert rel_l2(t.apply(phi), rhs) < 1e-10
    tot = (t.phi() + phi).reshape(Nr, n1)
    e = np.zeros_like(tot)
    e[:, 1:-1] = (tot[:, :-2] - tot[:, 2:]) / (2 * t.hz)
    assert np.array_equal(t.enodes().reshape(Nr, n1), e)            # same expression, same rounding (Source/PenningTrap.cpp:226-233)
    # fixed point: independent of how the segments are dealt to CTAs
    grids = []
    for ctas in (0, 37):
        tf = ptp.PenningTrap(0.01488, el, [0.0005] * 4, Nz, Nr)
        tf.set_deposit_mode(ptp.PTP_DEPOSIT_FIXED64)
        tf.set_tuning(ctas=ctas)
        q = ptp.Plasma(tf, "Antiprotons", ptp.massP, -ptp.ePos)
        q.loadDensity(dens, 150.0, 1_000_000)
        tf.movePlasmas(dt, 2)
        grids.append((q.rhs(), q.selfPotential()))
        tf.close()
    assert np.array_equal(grids[0][0], grids[1][0]) and np.array_equal(grids[0][1], grids[1][1])
    t.close()


def test_adaptive_resort_keeps_results_and_triggers(monkeypatch, c1_kat):
    """Electrons on a fine grid (1.6 cells of drift per step) leave the cell ranges their segments were planned for within a
    few steps. The adaptive policy (ptp_trap_set_sort_interval(-1), the default) must notice the out-of-window deposits and
    re-sort; with fixed-point deposits the whole trajectory is bitwise independent of the ring order, so the run with
    re-sorts equals the run without (interval 0) ring by ring."""
    from bench import density_on
    monkeypatch.setenv("PTP_SORT_CHECK_STEPS", "8")
    Nz, Nr = 4096, 64
    el = [ptp.Electrode(0.01322, v) for v in (0, -70, -15, -70, 0)]
    dens = density_on(Nz, Nr)
    dt = float(c1_kat["dt"])
    res = []
    for interval in (0, -1):
        t = ptp.PenningTrap(0.01488, el, [0.0005] * 4, Nz, Nr)
        t.set_deposit_mode(ptp.PTP_DEPOSIT_FIXED64)
        t.set_sort_interval(interval)
        p = ptp.Plasma(t, "Electrons", ptp.massE, -ptp.ePos)
        n, _ = p.loadDensity(dens, 150.0, 400_000)
        t.movePlasmas(dt, 60)
        r, z, v, ids = _by_id(p)
        res.append((r, z, v, ids, p.rhs(), p.selfPotential(), t.sorts_done(), p.getNumMacro()))
        t.close()
    assert res[0][6] == 0 and res[1][6] >= 1, (res[0][6], res[1][6])
    assert res[0][7] == res[1][7]
    for a, b in zip(res[0][:6], res[1][:6]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("exact", [False, True])
def test_hot_species_form_of_push_kernel(c1_kat, exact):
    """ptp_plasma_set_hot: the per-warp-bin form of K1 (one wide window, no re-sorts) against the thread-private form.
    Electrons on a fine grid (1.6 cells per step: after 40 steps the load order is gone), fixed-point deposits: positions,
    speeds, deposit grid and potential identical bit for bit at every check point, with and without re-sorts of the
    thread-private run; stand-alone deposit (K2) too. fp64 deposits: grid within the fp64 tier (1e-12), and the
    trajectories stay together at rounding level."""
    from bench import density_on
    Nz, Nr = 4096, 64
    el = [ptp.Electrode(0.01322, v) for v in (0, -70, -15, -70, 0)]
    dens = density_on(Nz, Nr)
    dt = float(c1_kat["dt"])

    def run(hot, mode, check_steps):
        t = ptp.PenningTrap(0.01488, el, [0.0005] * 4, Nz, Nr)
        t.set_deposit_mode(mode)
        t.set_arith_mode(ptp.PTP_ARITH_EXACT if exact else ptp.PTP_ARITH_FAST)
        p = ptp.Plasma(t, "Electrons", ptp.massE, -ptp.ePos)
        p.set_hot(hot)
        n, _ = p.loadDensity(dens, 150.0, 300_000)
        assert p.is_hot() == bool(hot)
        out = [(p.rhs(), p.selfPotential())]                    # the loader's first deposit (K2) and solve
        for k in check_steps:
            t.movePlasmas(dt, k)
            out.append(_by_id(p) + (p.rhs(), p.selfPotential()))
        out.append((t.sorts_done(), p.getNumMacro(), p.is_hot()))
        t.close()
        return out

    a = run(0, ptp.PTP_DEPOSIT_FIXED64, (1, 7, 40))
    b = run(1, ptp.PTP_DEPOSIT_FIXED64, (1, 7, 40))
    assert a[-1][2] is False and b[-1] == (0, a[-1][1], True)          # no re-sort of the hot form, same survivors
    for x, y in zip(a[:-1], b[:-1]):
        for u, w in zip(x, y):
            assert np.array_equal(u, w)
    c = run(0, ptp.PTP_DEPOSIT_FP64, (1, 12))
    d = run(1, ptp.PTP_DEPOSIT_FP64, (1, 12))
    assert rel_l2(d[0][0], c[0][0]) < 1e-12 and rel_l2(d[0][1], c[0][1]) < 1e-11
    assert rel_l2(d[1][1], c[1][1]) < 1e-14 and rel_l2(d[1][2], c[1][2]) < 1e-12    # (the fields of the first step differ at rounding level)
    assert rel_l2(d[1][4], c[1][4]) < 1e-12
    assert rel_l2(d[2][1], c[2][1]) < 1e-12 and rel_l2(d[2][4], c[2][4]) < 1e-9 and rel_l2(d[2][5], c[2][5]) < 1e-9


def test_hot_species_are_found_by_the_resort_policy(monkeypatch, c1_kat):
    """Adaptive policy: electrons on a fine grid need a re-sort every ~20 steps; after the second one within
    PTP_HOT_SORT_STEPS the species is handed to the per-warp-bin form and the re-sorts stop. Antiprotons on the same grid
    (0.04 cells per step) stay with the thread-private form. PTP_SCATTER=0 keeps the old behaviour (re-sorts go on)."""
    from bench import density_on
    monkeypatch.setenv("PTP_SORT_CHECK_STEPS", "8")
    Nz, Nr = 4096, 64
    el = [ptp.Electrode(0.01322, v) for v in (0, -70, -15, -70, 0)]
    dens = density_on(Nz, Nr)
    dt = float(c1_kat["dt"])
    res = {}
    for tag, mass, env in (("e", ptp.massE, None), ("p", ptp.massP, None), ("e0", ptp.massE, "0")):
        if env is not None:
            monkeypatch.setenv("PTP_SCATTER", env)
        t = ptp.PenningTrap(0.01488, el, [0.0005] * 4, Nz, Nr)
        p = ptp.Plasma(t, "Species", mass, -ptp.ePos)
        n, _ = p.loadDensity(dens, 150.0, 400_000)
        t.movePlasmas(dt, 70)
        mid = (t.sorts_done(), p.is_hot())
        t.movePlasmas(dt, 80)
        res[tag] = mid + (t.sorts_done(), p.is_hot(), p.getNumMacro(), n)
        t.close()
    assert res["e"][1] and res["e"][3] and 1 <= res["e"][0] <= 3 and res["e"][2] == res["e"][0], res
    assert not res["p"][3] and res["p"][2] <= 1, res
    assert not res["e0"][3] and res["e0"][2] > res["e"][2], res
    assert res["e"][4] == res["e"][5] and res["e0"][4] == res["e0"][5]


def test_edge_cases():
    t = ptp.default_trap()
    p = ptp.Plasma(t, "Electrons", ptp.massE, -ptp.ePos)
    # empty plasma: a step is a no-op
    p.upload(np.zeros(0, np.int32), np.zeros(0), np.zeros(0), -1e-18)
    t.movePlasmas(1e-10, 2)
    assert p.getNumMacro() == 0 and not p.rhs().any()
    # ragged: one ring in the last row, one in the first, unsorted input, positions next to both ends
    r = np.array([127, 0, 5], np.int32)
    z = np.array([t.hz * 0.25, t.lengthTrap - t.hz * 0.25, t.hz * 300.5])
    v = np.array([-1e9, 1e9, 0.0])                   # the two end rings leave on the first step
    p.upload(r, z, v, -1e-18)
    p.solvePoisson()
    rr, zz, vv, ids = p.download()
    assert sorted(ids.tolist()) == [0, 1, 2] and np.array_equal(rr[np.argsort(ids)], r)
    t.movePlasmas(1e-10, 1)
    assert p.getNumMacro() == 1
    rr, zz, vv, ids = p.download()
    assert ids.tolist() == [2] and rr.tolist() == [5]
    w = p.rhs() / (-p.macroChargeDensity / ptp.epsilon)
    assert w.sum() == pytest.approx(1.0, abs=1e-12)
    # bad input is rejected
    with pytest.raises(ptp.PtpError):
        p.upload(np.array([128], np.int32), np.array([0.01]), np.array([0.0]), -1e-18)
    with pytest.raises(ptp.PtpError):
        p.upload(np.array([1], np.int32), np.array([1.0]), np.array([0.0]), -1e-18)
    with pytest.raises(ValueError):
        ptp.PenningTrap(0.01, [ptp.Electrode(0.01, 0)], [0.001], 16, 4)
    t.close()


# ------------------------------------------------------------------------------------------ loaders on the device (f-1)
@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("num_macro,mass", [(4000, ptp.massE), (100000, ptp.massE), (1000003, ptp.massP)])
def test_device_loader_matches_reference_loader(density_files, num_macro, mass):
    """ptp_plasma_load_density against Plasma::loadDensityFile of the compiled reference (Source/Plasma.cpp:558-622):
    ring counts per row and chargeMacro exact, positions bit-identical, speeds = the reference's own deviate stream
    (minstd_rand0 + polar method reproduced in parallel) to the last bits of log(), first deposit and solve within the
    step tolerances; shards partition the load ring for ring."""
    dens = np.loadtxt(density_files[0])
    rt = ref.default_trap()
    rp = rt.plasma("Species", mass, -ref.E_POS)
    rp.load_density_file(density_files[0], 150.0, num_macro)
    r0, z0, v0 = rp.rings()
    t = ptp.default_trap()
    p = ptp.Plasma(t, "Species", mass, -ptp.ePos)
    n, per_row = p.loadDensity(dens, 150.0, num_macro)
    assert n == rp.count() == p.getNumMacro() == int(per_row.sum())
    assert np.array_equal(per_row, np.bincount(r0, minlength=t.Nr))
    par = rp.params()
    assert p.chargeMacro == par["chargeMacro"] and p.macroChargeDensity == par["macroChargeDensity"]
    r, z, v, ids = _by_id(p)
    assert np.array_equal(ids, np.arange(n))
    assert np.array_equal(r, r0)
    assert np.array_equal(z, z0)                                     # bit-identical positions
    assert np.max(np.abs(v - v0) / np.abs(v0)) < 2e-15               # same deviates; log() may differ in the last bit
    assert np.mean(v != v0) < 0.35
    assert rel_l2(p.rhs(), rp.rhs()) < 1e-12
    assert rel_l2(p.selfPotential(), rp.self_potential()) < 1e-10
    # shards: rings i = s (mod S) of every row, same values
    S = 3
    within = np.concatenate([np.arange(c) for c in per_row if c > 0])    # index of a ring inside its row, reference order
    for sh in range(S):
        q = ptp.Plasma(t, "Shard", mass, -ptp.ePos)
        ns, _ = q.loadDensity(dens, 150.0, num_macro, shard=sh, nShards=S, solve=False)
        sel = within % S == sh
        assert ns == int(sel.sum())
        rs, zs, vs, _ = _by_id(q)
        assert np.array_equal(rs, r0[sel]) and np.array_equal(zs, z0[sel]) and np.array_equal(vs, v[sel])
    t.close()
    rt.close()


# ------------------------------------------------------------------------------------------ other grid shapes
@pytest.mark.parametrize("Nz,Nr,solver,fixed", [(1024, 96, 0, 0), (2048, 16, 0, 0), (1500, 12, 0, 0), (301, 130, 0, 0), (48, 420, 0, 0),
                                                (256, 24, 2, 0), (256, 24, 2, 1), (64, 300, 2, 0), (128, 5, 2, 0), (8, 70, 2, 0)])
def test_solver_and_step_on_other_grids(Nz, Nr, solver, fixed):
    _other_grid_case(Nz, Nr, solver, fixed)


@pytest.mark.parametrize("r16", ["0", "1"])
@pytest.mark.parametrize("Nr,fixed", [(12, 0), (40, 1)])
def test_4096_node_rows_both_inverse_kernels(Nr, fixed, r16, monkeypatch):
    """Rows of config 5's length (Nz = 4096) on grids small enough for the LU oracle: k_idct_fft_field (PTP_FFT_R16=0)
    and k_idct_r16_field (=1) against the oracle - trap potential, a random right-hand side, deposits, three steps."""
    monkeypatch.setenv("PTP_FFT_R16", r16)
    _other_grid_case(4096, Nr, 2, fixed)


def test_radix16_and_radix2_inverse_agree(monkeypatch):
    """Same spectrum through both inverse-transform kernels on a 4096 x 64 grid: potentials agree to rounding."""
    args = (0.012, [0.02, 0.03, 0.02], [0.0, -50.0, 0.0], [0.001, 0.001], 4096, 64)
    rng = np.random.default_rng(5)
    out = []
    for r16 in ("0", "1"):
        monkeypatch.setenv("PTP_FFT_R16", r16)
        t = ptp.PenningTrap(args[0], [ptp.Electrode(a, b) for a, b in zip(args[1], args[2])], args[3], args[4], args[5])
        x = np.random.default_rng(5).standard_normal(t.G)
        out.append((t.phi(), t.solve(x), t.enodes()))
        t.close()
    for a, b in zip(out[0], out[1]):
        assert rel_l2(b, a) < 1e-13


@pytest.mark.parametrize("Nz,Nr,solver", [(1024, 96, 0), (301, 130, 0), (48, 420, 0), (256, 24, 2), (4096, 12, 2)])
def test_other_grids_exact_arithmetic_is_bit_exact(Nz, Nr, solver):
    """The same lock-step comparison in EXACT arithmetic (true divisions, the reference's expression order): with the oracle's
    potentials injected before every step the per-ring z and v, the cell indices and the loss flags are bit-identical -
    no index may differ, on any solver path."""
    _other_grid_case(Nz, Nr, solver, 0, exact=True)


def _other_grid_case(Nz, Nr, solver, fixed, exact=False):
    """Grid shapes that take the other code paths of the solver against the CPU oracle: odd Nz+1 (8-byte copies), pipelined
    cosine ring (1024), chunked inverse GEMM + separate node field for long rows that are not a power of two (1500), and
    the large-grid organisation (ptp_solve_wide.cu): tiled forward transform + streamed radial solves for many radial
    nodes (420, dense inverse), plus the FFT inverse fused with the node field for long power-of-two rows (2048; forced
    with solver 2 on small grids, odd and even numbers of FFT passes, more rows than one tile, fixed-point deposits)."""
    args = (0.012, [0.02, 0.03, 0.02], [0.0, -50.0, 0.0], [0.001, 0.001], Nz, Nr)
    pt = port.PortTrap(*args)
    t = ptp.PenningTrap(args[0], [ptp.Electrode(a, b) for a, b in zip(args[1], args[2])], args[3], Nz, Nr)
    if fixed:
        t.set_deposit_mode(ptp.PTP_DEPOSIT_FIXED64)
    if exact:
        t.set_arith_mode(ptp.PTP_ARITH_EXACT)
    if solver:
        t.set_solver(solver)
        t.solveLaplace()
    assert rel_l2(t.phi(), pt.phi) < 1e-10
    rng = np.random.default_rng(Nz + Nr)
    x = rng.standard_normal(t.G)
    assert rel_l2(t.solve(x), pt.solve(x)) < 1e-9
    n = 50000
    rows = min(Nr, 40)
    r = np.sort(rng.integers(0, rows, n)).astype(np.int32)
    z = pt.length * (0.5 + 0.1 * np.clip(rng.standard_normal(n) * 0.4, -1, 1))
    v = rng.normal(0, 3e4, n)
    op = pt.plasma("Electrons", ptp.massE, -ptp.ePos)
    op.set_rings(r, z, v, -2e-18)
    gp = ptp.Plasma(t, "Electrons", ptp.massE, -ptp.ePos)
    gp.upload(r, z, v, -2e-18)
    op.solve_poisson()
    gp.solvePoisson()
    assert rel_l2(gp.rhs(), op.rhs) < (1e-9 if fixed else 1e-12)
    assert rel_l2(gp.selfPotential(), op.self_potential) < 1e-9
    dt = min(0.2 * pt.hz / 3e4, 5e-10)               # keep the coarse grids out of the violently non-linear regime
    for _ in range(3):
        t.set_phi(pt.phi)
        gp.set_self_potential(op.self_potential)
        pt.move_plasmas(dt)
        t.movePlasmas(dt)
        rr, zz, vv, ids = gp.download()
        assert len(zz) == op.count()
        # rings may leave on these coarse grids, and the oracle's swap-with-back removal permutes its array:
        # compare as multisets ordered by (row, z)
        og, oo = np.lexsort((zz, rr)), np.lexsort((op.z, op.r))
        assert np.array_equal(rr[og], op.r[oo])
        assert np.max(np.abs(zz[og] - op.z[oo]) / op.z[oo]) < 1e-13
        k_g, _ = gp.cell_index()
        k_o, _ = op.cell_index()
        if exact:
            ov = np.lexsort((op.v, op.z, op.r))
            gv = np.lexsort((vv, zz, rr))
            assert np.array_equal(zz[gv], op.z[ov]) and np.array_equal(vv[gv], op.v[ov])     # bit for bit
            assert np.array_equal(k_g[og], k_o[oo])                                           # every index
        assert np.mean(k_g[og] != k_o[oo]) < 1e-3          # identical unless a 1e-14 difference in z straddles a node
        assert rel_l2(gp.rhs(), op.rhs) < (1e-9 if fixed else 1e-12)
        assert rel_l2(gp.selfPotential(), op.self_potential) < 1e-9
        assert rel_l2(t.enodes(), pt.enodes()) < 1e-7
    t.close()
    pt.close()


@pytest.mark.parametrize("cluster", ["1", "0"])
def test_step_solves_populated_rows_and_whole_grids_on_demand(c1_kat, monkeypatch, cluster):
    """A step computes potentials and node field for the populated radial rows only (rings never change their row,
    Source/Plasma.hpp:22-24) - through the one-kernel cluster solve (PTP_CLUSTER_SOLVE=1, the default for plasmas on few
    rows) or the two-kernel path restricted to those rows (=0); what the reference keeps on the whole grid
    (Plasma::selfPotential, the node field) is produced when asked for. Against PTP_FULL_SOLVE=1 (every step solves the
    whole grid). Two-kernel path: same fold row, same arithmetic per row - per-ring state, whole-grid potentials and node
    field bit-identical after 20 steps. Cluster kernel: another summation order in the inverse transform - agreement to
    rounding (z, v rel <= 1e-13, potentials <= 1e-12). A species loaded later into rows the last step left out is pushed
    with the right field (the potentials are completed first)."""
    res = []
    for full in ("1", "0"):
        monkeypatch.setenv("PTP_FULL_SOLVE", full)
        monkeypatch.setenv("PTP_CLUSTER_SOLVE", cluster)
        t, el, ap = _fresh_c1(c1_kat, ptp.PTP_DEPOSIT_FIXED64)
        el.solvePoisson()
        ap.solvePoisson()
        dt = float(c1_kat["dt"])
        t.movePlasmas(dt, 20)
        pe = el.getPotentialEnergy()
        _, z, v, ids = el.download()
        o = np.argsort(ids)
        first = (z[o], v[o], el.selfPotential(), ap.selfPotential(), t.enodes(), el.rhs(), pe)
        # a third species far out (row 100) joins: the next step needs the whole-grid potentials of the first two there
        far = ptp.Plasma(t, "Far", ptp.massP, -ptp.ePos)
        n = 5000
        rng = np.random.default_rng(3)
        far.upload(np.full(n, 100, np.int32), t.getLength() * (0.45 + 0.1 * rng.random(n)), rng.normal(0, 1e3, n), float(c1_kat["p_chargeMacro"]))
        far.solvePoisson()
        t.movePlasmas(dt, 3)
        _, zf, vf, idf = far.download()
        of = np.argsort(idf)
        res.append((first, (zf[of], vf[of], el.selfPotential(), ap.selfPotential(), far.selfPotential(), t.enodes(), el.rhs())))
        t.close()
    (a1, a2), (b1, b2) = res
    for n, (x, y) in enumerate(zip(a1[:6], b1[:6])):
        if cluster == "0":
            assert np.array_equal(x, y), n
        else:
            assert rel_l2(y, x) < (1e-13 if n < 2 else 1e-11), n
    assert a1[6] == pytest.approx(b1[6], rel=1e-12)       # (the reduction's atomic adds land in arbitrary order)
    for n, (x, y) in enumerate(zip(a2, b2)):
        assert rel_l2(y, x) < 1e-11, n


def test_graph_replay_is_bitwise_identical(c1_kat):
    """ptp_trap_set_graph: the step replayed as a CUDA graph gives exactly the stream-launched result (fixed-point deposit,
    so that the comparison is bitwise), is faster for the launch-bound default configuration, and survives a change of
    dt, an electrode change and a re-upload."""
    import time
    res = []
    for graph in (False, True):
        t, el, ap = _fresh_c1(c1_kat, ptp.PTP_DEPOSIT_FIXED64)
        t.set_graph(graph)
        el.solvePoisson()
        ap.solvePoisson()
        dt = float(c1_kat["dt"])
        t.movePlasmas(dt, 20)
        t.sync()
        t0 = time.perf_counter()
        t.movePlasmas(dt, 300)
        t.sync()
        sec = time.perf_counter() - t0
        mid = el.rhs()
        t.movePlasmas(dt * 0.5, 7)                 # new dt -> new graph
        mid2 = el.rhs()
        t.setPotential(1, -60.0)                  # same graph, new trap potential
        t.movePlasmas(dt * 0.5, 5)
        _, z, v, ids = el.download()
        o = np.argsort(ids)
        res.append((mid, mid2, z[o], v[o], el.rhs(), ap.selfPotential(), t.enodes(), sec, t.last_launches()))
        t.close()
    a, b = res
    for n, (x, y) in enumerate(zip(a[:7], b[:7])):
        assert np.array_equal(x, y), n
    assert b[8] == a[8]                            # same kernels launched, through the graph
    print("300 steps of C1: stream %.1f us/step, graph %.1f us/step" % (a[7] / 300 * 1e6, b[7] / 300 * 1e6))


# ------------------------------------------------------------------------------------------ diagnostics on the device (f-2)
@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
def test_temperature_sums_match_reference_get_temperature(c1_kat):
    """Plasma::getTemperature (Source/Plasma.cpp:212-228) from sums formed on the device at the save points - ring mass
    massMacro (8 r), mean of the speeds at two consecutive save points - against the compiled reference run through the same
    protocol (saveStates, steps, saveStates); also after a re-sort (the saved speeds travel with their rings) and for both
    species of C1. rel <= 1e-12 at the first pair (identical inputs), <= 1e-6 free-running."""
    rt = ref.default_trap()
    t, el, ap = _fresh_c1(c1_kat)
    pairs = []
    for tag, name, mass, g in (("e", "Electrons", ptp.massE, el), ("p", "Antiprotons", ptp.massP, ap)):
        o = rt.plasma(name, mass, -ref.E_POS)
        o.set_rings(c1_kat[f"{tag}_r0"], c1_kat[f"{tag}_z0"], c1_kat[f"{tag}_v0"], float(c1_kat[f"{tag}_chargeMacro"]))
        o.solve_poisson()
        g.solvePoisson()
        pairs.append((g, o))
    dt = float(c1_kat["dt"])
    rt.save_states(0.0)
    assert all(g.saveSpeeds() is None for g, _ in pairs)         # first save point: no pair yet
    for k in range(4):
        rt.move_plasmas(dt, 3)
        t.movePlasmas(dt, 3)
        if k == 2:
            t.sort()
        rt.save_states((k + 1) * 3 * dt)
        for g, o in pairs:
            assert g.saveSpeeds() == pytest.approx(o.temperature(), rel=1e-12 if k == 0 else 1e-6)
    t.close()
    rt.close()


def test_row_slice_download_and_loss_log(c1_kat):
    """ptp_plasma_download_row = the rings of one radial row (what Plasma::saveState(int indexR) needs, Source/Plasma.cpp:338-346)
    without moving the other rows; ptp_plasma_loss_log = which ring left in which step (Source/Plasma.cpp:108-118)."""
    t, el, ap = _fresh_c1(c1_kat)
    el.solvePoisson()
    ap.solvePoisson()
    dt = float(c1_kat["dt"])
    t.movePlasmas(dt, 2)
    r, z, v, ids = _by_id(el)
    for row in (0, 3, 9, 50):
        zr, vr, ir = el.downloadRow(row)
        o = np.argsort(ir)
        sel = r == row
        assert np.array_equal(ir[o], ids[sel]) and np.array_equal(zr[o], z[sel]) and np.array_equal(vr[o], v[sel])
    # losses: lower the barrier and compare the log with the oracle's per-step losses
    ot = port.default_trap()
    op = ot.plasma("Electrons", ptp.massE, -ptp.ePos)
    op.set_rings(r, z, v, float(c1_kat["e_chargeMacro"]))
    oa = ot.plasma("Antiprotons", ptp.massP, -ptp.ePos)
    ra, za, va, _ = _by_id(ap)
    oa.set_rings(ra, za, va, float(c1_kat["p_chargeMacro"]))
    op.solve_poisson()
    oa.solve_poisson()
    ot.set_potential(1, -47.0)
    t.setPotential(1, -47.0)
    lost_per_step = []
    for _ in range(70):
        before = op.count()
        ot.move_plasmas(dt, 1)
        lost_per_step.append(before - op.count())
    t.movePlasmas(dt, 70)
    lids, lsteps, total, over = el.lossLog()
    assert not over and total == len(lids) == sum(lost_per_step) > 100
    first = int(lsteps.min()) - next(i for i, x in enumerate(lost_per_step) if x)      # tag of the first of the 70 steps
    assert [int(np.sum(lsteps == first + i)) for i in range(70)] == lost_per_step
    assert len(np.unique(lids)) == len(lids) and el.getNumMacro() == op.count()
    _, _, _, alive = _by_id(el)
    assert not np.intersect1d(alive, lids).size
    t.close()
    ot.close()
