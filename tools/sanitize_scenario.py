"""Small scenario for compute-sanitizer: the kernels added in the third session of round 1 - radix-16 inverse (Nz = 4096),
rewritten sort (twice, so that the alternate buffers' padding logic runs), adaptive re-sort, wide field window."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PTP_SORT_CHECK_STEPS"] = "4"
ptp = importlib.import_module("pic-trapped-plasma_b200")
from bench import density_on

Nz, Nr = 4096, 16
el = [ptp.Electrode(0.01322, v) for v in (0, -70, -15, -70, 0)]
dens = density_on(Nz, Nr)
for mode in (ptp.PTP_DEPOSIT_FP64, ptp.PTP_DEPOSIT_FIXED64):
    t = ptp.PenningTrap(0.01488, el, [0.0005] * 4, Nz, Nr)
    t.set_deposit_mode(mode)
    p = ptp.Plasma(t, "Electrons", ptp.massE, -ptp.ePos)
    n, _ = p.loadDensity(dens, 150.0, 60_000)
    t.movePlasmas(2e-8 / 35, 24)
    t.sort()
    t.movePlasmas(2e-8 / 35, 8)
    t.sort()
    t.movePlasmas(2e-8 / 35, 8)
    r, z, v, ids = p.download()
    assert len(z) == p.getNumMacro() and len(np.unique(ids)) == len(ids)
    print("mode", mode, "rings", len(z), "sorts", t.sorts_done(), "phi", float(np.abs(p.selfPotential()).max()))
    t.close()
print("scenario ok")
