"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the dev container only (needs oracle/_ref/libptp_ref.so, i.e. /root/reference):

    python tests/golden/make_golden.py

What it writes (all small; committed):

* ``expected_density_nonzero.npz`` -- the equilibrium charge density the
  reference's driver A produces (Diagnostics/A) Grid Size and Plasma
  Period.txt:86-110): ``loadProfile(150 K, -e*250000, 3.549, 0.6, numMacro=1,
  KS=1e-13)`` on the driver-A trap, stored as the full-precision grid (indices +
  values of the non-zero nodes).  Tests rebuild driver A's two text files
  ("Z Expected Electron/Antiproton Density.csv": ratio 0.6 / 0.4, default ostream
  precision = 6 significant digits) from it.
* ``charge_density_0.txt`` -- verbatim copy of the reference's only data file
  Diagnostics/Charge_Density-0.txt (36 rows ``z n``), the known-answer vector.
* ``fixture_rings_r0.npz`` -- the r=0 rings (z only) of ``loadDensityFile(electrons, 150 K,
  100000)`` whose deposit IS that data file (lets the plain-C oracle be pinned to the
  fixture on a box without the reference).
* ``trap_kat.json`` -- scalar known answers of the default trap (hz, hr, length,
  phi_trap samples, well limits, nnz).
* ``c1_step_kat.npz`` -- C1 (4001 e- + 4001 pbar) state before/after steps of the
  reference's movePlasmas: rings, RHS, phi_self, node E (lock-step vectors).
* ``driver_d_counts.json`` -- alive counts through the driver-D e-kick protocol
  (integer KAT of the loss path, Diagnostics/D) Useless Boundary Test.txt:118-136).
"""
import json
import math
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402


def write_density_file(path, dens, ratio):
    # driver A: fileElectrons << readNumber * eRatio << '\n' with default precision (6 sig. digits);
    # readNumber itself is the 15-digit text of initialDensity (extractInitialDensity, FullPrecision).
    with open(path, "w") as f:
        for x in dens:
            x15 = float("%.15g" % x)
            f.write("%.6g\n" % (x15 * ratio))


def main():
    trap = ref.default_trap()
    kat = dict(Nz=trap.Nz, Nr=trap.Nr, hz=trap.hz, hr=trap.hr, length=trap.length, radius=trap.radius)
    phi = trap.phi()
    left, right = trap.limits()
    kat.update(phi_r0_k293=phi[293], phi_min=float(phi.min()), phi_max=float(phi.max()),
               limit_left_r0=int(left[0]), limit_right_r0=int(right[0]),
               limit_left_r127=int(left[127]), limit_right_r127=int(right[127]),
               nnz=int(len(trap.matrix()[0])))
    resid = trap.apply(phi) - trap.wall_rhs()
    kat["laplace_residual_rel"] = float(np.linalg.norm(resid) / np.linalg.norm(trap.wall_rhs()))

    # --- driver A: expected density ---------------------------------------------------------
    pl = trap.plasma("Electrons", ref.MASS_E, -ref.E_POS)
    pl.load_profile(150.0, -ref.E_POS * 250000, 3.549, 0.6, 1, 1e-13)
    dens = pl.initial_density()
    nz = np.nonzero(dens)[0].astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "expected_density_nonzero.npz"), G=np.int64(trap.G), index=nz, value=dens[nz])
    kat["expected_density_peak_r0"] = float(np.abs(dens[:trap.Nz + 1]).max() / ref.E_POS)
    trap.close()

    shutil.copyfile("/root/reference/Diagnostics/Charge_Density-0.txt", os.path.join(HERE, "charge_density_0.txt"))

    tmp = tempfile.mkdtemp()
    fe, fp = os.path.join(tmp, "e.csv"), os.path.join(tmp, "p.csv")
    write_density_file(fe, dens, 0.6)
    write_density_file(fp, dens, 1 - 0.6)

    # --- the fixture itself: 100000 electrons, RHS[275..310]*eps0/e ---------------------------
    trap = ref.default_trap()
    el = trap.plasma("Electrons", ref.MASS_E, -ref.E_POS)
    el.load_density_file(fe, 150.0, 100000)
    rhs = el.rhs()
    fixture = np.loadtxt(os.path.join(HERE, "charge_density_0.txt"))
    mine = rhs[275:311] * ref.EPSILON0 / ref.E_POS
    kat["fixture_rel_l2"] = float(np.linalg.norm(mine - fixture[:, 1]) / np.linalg.norm(fixture[:, 1]))
    kat["fixture_count_100000"] = int(el.count())
    r100k, z100k, _ = el.rings()
    np.savez_compressed(os.path.join(HERE, "fixture_rings_r0.npz"), z=z100k[r100k == 0],
                        macroChargeDensity=el.params()["macroChargeDensity"], chargeMacro=el.params()["chargeMacro"])
    print("fixture rel-L2", kat["fixture_rel_l2"])
    trap.close()

    # --- C1 lock-step vectors -----------------------------------------------------------------
    trap = ref.default_trap()
    el = trap.plasma("Electrons", ref.MASS_E, -ref.E_POS)
    el.load_density_file(fe, 150.0, 4000)
    ap = trap.plasma("Antiprotons", ref.MASS_P, -ref.E_POS)
    ap.load_density_file(fp, 150.0, 4000)
    dt = 2e-8 / 35
    out = dict(dt=dt, phi_trap=trap.phi())
    for tag, p in (("e", el), ("p", ap)):
        r, z, v = p.rings()
        out.update({f"{tag}_r0": r, f"{tag}_z0": z, f"{tag}_v0": v, f"{tag}_rhs0": p.rhs(), f"{tag}_phi0": p.self_potential(),
                    f"{tag}_mcd": p.params()["macroChargeDensity"], f"{tag}_chargeMacro": p.params()["chargeMacro"]})
    out["enodes0"] = trap.enodes()
    trap.move_plasmas(dt, 1)
    for tag, p in (("e", el), ("p", ap)):
        r, z, v = p.rings()
        out.update({f"{tag}_z1": z, f"{tag}_v1": v, f"{tag}_rhs1": p.rhs(), f"{tag}_phi1": p.self_potential()})
    out["enodes1"] = trap.enodes()
    # free-running soft KATs after 175 steps (5 periods), with histories for getTemperature
    trap.save_states(dt)
    for i in range(2, 176):
        trap.move_plasmas(dt, 1)
        if i >= 174:
            trap.save_states(i * dt)
    kat.update(c1_T_e_175=el.temperature(), c1_T_p_175=ap.temperature(), c1_PE_175=trap.last_potential_energy(),
               c1_count_e_175=int(el.count()), c1_count_p_175=int(ap.count()))
    for tag, p in (("e", el), ("p", ap)):
        r, z, v = p.rings()
        out.update({f"{tag}_z175": z, f"{tag}_v175": v, f"{tag}_rhs175": p.rhs()})
    np.savez_compressed(os.path.join(HERE, "c1_step_kat.npz"), **out)
    trap.close()

    # --- driver D: e-kick loss counts -----------------------------------------------------------
    def changed_voltage(Vi, Vf, duration, compression, t):
        return (Vf - Vi) * (1 + math.exp(-(t - duration / 2) * compression * 2 / duration)) ** -1 + Vi

    trap = ref.default_trap()
    el = trap.plasma("Electrons", ref.MASS_E, -ref.E_POS)
    el.load_density_file(fe, 150.0, 4000)
    ap = trap.plasma("Antiprotons", ref.MASS_P, -ref.E_POS)
    ap.load_density_file(fp, 150.0, 4000)
    dtD = float("%.15g" % dt)  # driver D reads deltaT back from a 15-digit text file
    counts = [(int(el.count()), int(ap.count()))]
    i = 1
    while i * dtD <= 10e-9:
        trap.set_potential(1, changed_voltage(-70, -51, 10e-9, 4.5, i * dtD))
        trap.move_plasmas(dtD, 1)
        counts.append((int(el.count()), int(ap.count())))
        i += 1
    trap.set_potential(1, -51)
    i = 1
    while i * dtD <= 80e-9:
        trap.move_plasmas(dtD, 1)
        counts.append((int(el.count()), int(ap.count())))
        i += 1
    i = 1
    while i * dtD <= 10e-9:
        trap.set_potential(1, changed_voltage(-51, -70, 10e-9, 4.5, i * dtD))
        trap.move_plasmas(dtD, 1)
        counts.append((int(el.count()), int(ap.count())))
        i += 1
    json.dump(dict(dt=dtD, counts=counts, central_well_e=el.num_central_well(), central_well_p=ap.num_central_well()),
              open(os.path.join(HERE, "driver_d_counts.json"), "w"))
    print("driver D: steps", len(counts) - 1, "final", counts[-1], "central well e", el.num_central_well())
    trap.close()

    json.dump(kat, open(os.path.join(HERE, "trap_kat.json"), "w"), indent=1)
    print(json.dumps(kat, indent=1))
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
