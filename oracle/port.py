"""TEST INFRASTRUCTURE ONLY -- ctypes view of oracle/libptp_oracle.so (ptp_oracle.c).

The plain-C restatement of the reference's PIC step, with the orchestration of
PenningTrap::movePlasmas (Source/PenningTrap.cpp:352-363) written out here.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libptp_oracle.so")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_lib = None


def build():
    subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    L = C.CDLL(LIB_PATH)
    vp, d, i, l = C.c_void_p, C.c_double, C.c_int, C.c_long
    pp = C.POINTER(C.c_void_p)

    def sig(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    sig("ptpo_trap_create", vp, d, i, _dp, _dp, i, _dp, i, i)
    sig("ptpo_trap_destroy", None, vp)
    sig("ptpo_trap_info", None, vp, C.POINTER(i), C.POINTER(i), C.POINTER(d), C.POINTER(d), C.POINTER(d), C.POINTER(d))
    sig("ptpo_set_electrode", None, vp, i, d)
    sig("ptpo_matrix_apply", None, vp, _dp, _dp)
    sig("ptpo_wall_potential", None, vp, _dp)
    sig("ptpo_wall_rhs", None, vp, _dp)
    sig("ptpo_solve", None, vp, _dp, _dp)
    sig("ptpo_well_limits", None, vp, _dp, _ip, _ip)
    sig("ptpo_node_efield", None, vp, i, pp, _dp)
    sig("ptpo_gather", d, vp, _dp, i, d)
    sig("ptpo_move_rings", l, vp, _dp, l, _ip, _dp, _dp, d, d, d)
    sig("ptpo_deposit", None, vp, l, _ip, _dp, d, _dp)
    sig("ptpo_cell_index", None, vp, l, _ip, _dp, _ip, _ip)
    sig("ptpo_potential_energy", d, vp, i, pp, l, _ip, _dp, d)
    _lib = L
    return L


def _ptr_array(arrays):
    arr = (C.c_void_p * len(arrays))()
    for n, a in enumerate(arrays):
        arr[n] = a.ctypes.data
    return arr


class PortPlasma:
    def __init__(self, trap, name, mass, charge):
        self.trap, self.name, self.mass, self.charge = trap, name, mass, charge
        self.r = np.zeros(0, np.int32)
        self.z = np.zeros(0)
        self.v = np.zeros(0)
        self.charge_macro = 0.0
        self.macro_charge_density = 0.0
        self.rhs = np.zeros(trap.G)
        self.self_potential = np.zeros(trap.G)

    def set_rings(self, r, z, v, charge_macro):
        self.r = np.array(r, dtype=np.int32)
        self.z = np.array(z, dtype=np.float64)
        self.v = np.array(v, dtype=np.float64)
        self.charge_macro = charge_macro
        # Source/Plasma.cpp:494 / :588
        self.macro_charge_density = 4 * charge_macro / (3.141592653589793238463 * self.trap.hz * self.trap.hr * self.trap.hr)

    def count(self):
        return len(self.r)

    def update_rhs(self):
        lib().ptpo_deposit(self.trap.h, len(self.r), self.r, self.z, self.macro_charge_density, self.rhs)

    def solve_poisson(self):  # Source/Plasma.cpp:95-99
        self.update_rhs()
        self.self_potential = self.trap.solve(self.rhs)

    def move_rings(self, dt, enodes):
        n = lib().ptpo_move_rings(self.trap.h, enodes, len(self.r), self.r, self.z, self.v, dt, self.charge, self.mass)
        self.r, self.z, self.v = self.r[:n].copy(), self.z[:n].copy(), self.v[:n].copy()

    def cell_index(self):
        k, idx = np.empty(len(self.r), np.int32), np.empty(len(self.r), np.int32)
        lib().ptpo_cell_index(self.trap.h, len(self.r), self.r, self.z, k, idx)
        return k, idx

    def potential_energy(self):
        phis = [self.trap.phi] + [p.self_potential for p in self.trap.plasmas]
        return lib().ptpo_potential_energy(self.trap.h, len(phis), _ptr_array(phis), len(self.r), self.r, self.z, self.charge_macro)


class PortTrap:
    def __init__(self, radius, lengths, potentials, gaps, Nz, Nr):
        L = lib()
        lengths = np.ascontiguousarray(lengths, dtype=np.float64)
        potentials = np.ascontiguousarray(potentials, dtype=np.float64)
        gaps = np.ascontiguousarray(gaps, dtype=np.float64)
        self.h = L.ptpo_trap_create(radius, len(lengths), lengths, potentials, len(gaps), gaps if len(gaps) else np.zeros(1), Nz, Nr)
        if not self.h:
            raise ValueError("Error number of gaps and electrodes; No. electrods should match No. gaps + 1")
        nz, nr, hz, hr, ln, rad = C.c_int(), C.c_int(), C.c_double(), C.c_double(), C.c_double(), C.c_double()
        L.ptpo_trap_info(self.h, nz, nr, hz, hr, ln, rad)
        self.Nz, self.Nr, self.hz, self.hr, self.length, self.radius = nz.value, nr.value, hz.value, hr.value, ln.value, rad.value
        self.G = (self.Nz + 1) * self.Nr
        self.plasmas = []
        self.solve_laplace()

    def wall_potential(self):
        out = np.empty(self.Nz + 1)
        lib().ptpo_wall_potential(self.h, out)
        return out

    def wall_rhs(self):
        out = np.empty(self.G)
        lib().ptpo_wall_rhs(self.h, out)
        return out

    def solve(self, rhs):
        out = np.empty(self.G)
        lib().ptpo_solve(self.h, np.ascontiguousarray(rhs, dtype=np.float64), out)
        return out

    def apply(self, x):
        out = np.empty(self.G)
        lib().ptpo_matrix_apply(self.h, np.ascontiguousarray(x, dtype=np.float64), out)
        return out

    def solve_laplace(self):  # Source/PenningTrap.cpp:199-203
        self.phi = self.solve(self.wall_rhs())

    def set_potential(self, idx, v):  # Source/PenningTrap.cpp:313-317
        lib().ptpo_set_electrode(self.h, idx, v)
        self.solve_laplace()

    def limits(self):
        a, b = np.empty(self.Nr, np.int32), np.empty(self.Nr, np.int32)
        lib().ptpo_well_limits(self.h, self.phi, a, b)
        return a, b

    def enodes(self):
        phis = [self.phi] + [p.self_potential for p in self.plasmas]
        out = np.empty(self.G)
        lib().ptpo_node_efield(self.h, len(phis), _ptr_array(phis), out)
        return out

    def gather(self, enodes, r, z):
        return lib().ptpo_gather(self.h, enodes, int(r), float(z))

    def plasma(self, name, mass, charge):
        p = PortPlasma(self, name, mass, charge)
        self.plasmas.append(p)
        return p

    def move_plasmas(self, dt, nsteps=1):  # Source/PenningTrap.cpp:352-363
        for _ in range(nsteps):
            enodes = self.enodes()  # all species are pushed with the pre-step field
            for p in self.plasmas:
                p.move_rings(dt, enodes)
            for p in self.plasmas:
                p.solve_poisson()

    def close(self):
        if self.h:
            lib().ptpo_trap_destroy(self.h)
            self.h = None


def default_trap(Nz=585, Nr=128):
    return PortTrap(0.01488, [0.01322] * 5, [0, -70, -15, -70, 0], [0.0005] * 4, Nz, Nr)
