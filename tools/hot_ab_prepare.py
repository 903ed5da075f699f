"""Inputs of tools/hot_ab.cpp: the c5 trap (4096 x 1024) - wall potentials per axial node (PenningTrap::updateRHS boundary values)
and the non-zero nodes of the expected density on that grid - written to build/hot_ab/c5.bin."""
import importlib
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

ptp = importlib.import_module("pic-trapped-plasma_b200")
Nz, Nr = 4096, 1024
el = [ptp.Electrode(0.01322, v) for v in (0, -70, -15, -70, 0)]
gaps = [0.0005] * 4
L = 0.0
for i, e in enumerate(el):
    L += e.getLength() + (gaps[i] if i < len(gaps) else 0.0)
wall = ptp.PenningTrap.wallPotential(types.SimpleNamespace(hz=L / Nz, Nz=Nz, electrodes=el, gaps=gaps))
dens = bench.density_on(Nz, Nr)
idx = np.flatnonzero(dens).astype(np.int64)
os.makedirs(os.path.join(ROOT, "build", "hot_ab"), exist_ok=True)
with open(os.path.join(ROOT, "build", "hot_ab", "c5.bin"), "wb") as f:
    f.write(np.array([Nz, Nr, len(idx)], np.int64).tobytes())
    f.write(np.array([L / Nz, 0.01488 / Nr, L, 0.01488, bench.DT, bench.TEMPERATURE, ptp.massE, -ptp.ePos], np.float64).tobytes())
    f.write(wall.astype(np.float64).tobytes())
    f.write(idx.tobytes())
    f.write(dens[idx].astype(np.float64).tobytes())
print("wrote build/hot_ab/c5.bin:", len(idx), "non-zero density nodes")
