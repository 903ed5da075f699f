"""TEST INFRASTRUCTURE ONLY -- ctypes view of oracle/_ref/libptp_ref.so.

That library is the reference's own Source/PenningTrap.cpp + Source/Plasma.cpp
compiled unmodified (oracle/Makefile, target ``ref``) behind the C handle API of
oracle/ref_harness.cpp.  Only tests/, __graft_entry__.smoke() and bench.py's CPU
legs may import this module; the product path never does.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libptp_ref.so")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(LIB_PATH)
    vp, d, i, l, s = C.c_void_p, C.c_double, C.c_int, C.c_long, C.c_char_p

    def sig(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    sig("ref_last_error", s)
    sig("ref_trap_create", vp, d, i, _dp, _dp, i, _dp, i, i)
    sig("ref_trap_destroy", None, vp)
    sig("ref_trap_info", None, vp, C.POINTER(i), C.POINTER(i), C.POINTER(d), C.POINTER(d), C.POINTER(d), C.POINTER(d))
    for n in ("ref_trap_get_phi", "ref_trap_get_wall_rhs", "ref_trap_get_enodes"):
        sig(n, None, vp, _dp)
    sig("ref_trap_limits", None, vp, _ip, _ip)
    sig("ref_trap_set_potential", None, vp, i, d)
    sig("ref_trap_solve", None, vp, _dp, _dp)
    sig("ref_trap_apply", None, vp, _dp, _dp)
    sig("ref_trap_matrix_nnz", l, vp)
    sig("ref_trap_matrix", None, vp, _ip, _ip, _dp)
    sig("ref_trap_efield", d, vp, i, d)
    sig("ref_trap_total_phi", d, vp, i, d)
    sig("ref_trap_move_plasmas", None, vp, d, i)
    sig("ref_trap_save_states", None, vp, d)
    sig("ref_trap_save_states_r", None, vp, d, i)
    sig("ref_trap_last_potential_energy", d, vp)
    for n in ("ref_trap_extract_trap_potential", "ref_trap_extract_trap_laplacian",
              "ref_trap_extract_trap_parameters", "ref_trap_extract_histories"):
        sig(n, None, vp, s)
    sig("ref_trap_timed_steps", d, vp, d, i, _dp)
    sig("ref_plasma_create", vp, vp, s, d, d)
    sig("ref_plasma_destroy", None, vp)
    sig("ref_plasma_load_profile", i, vp, d, d, d, d, i, d)
    sig("ref_plasma_load_density_file", i, vp, s, d, i)
    sig("ref_plasma_set_rings", None, vp, l, _ip, _dp, _dp, d)
    sig("ref_plasma_count", l, vp)
    sig("ref_plasma_num_central_well", i, vp)
    sig("ref_plasma_get_rings", None, vp, _ip, _dp, _dp)
    sig("ref_plasma_params", None, vp, C.POINTER(d), C.POINTER(d), C.POINTER(d), C.POINTER(d))
    for n in ("ref_plasma_get_rhs", "ref_plasma_get_self_potential", "ref_plasma_set_self_potential",
              "ref_plasma_get_initial_density"):
        sig(n, None, vp, _dp)
    for n in ("ref_plasma_update_rhs", "ref_plasma_solve_poisson"):
        sig(n, None, vp)
    sig("ref_plasma_move_rings", None, vp, d)
    sig("ref_plasma_temperature", d, vp)
    sig("ref_plasma_potential_energy", d, vp)
    for n in ("ref_plasma_extract_self_potential", "ref_plasma_extract_parameters",
              "ref_plasma_extract_initial_density"):
        sig(n, None, vp, s)
    _lib = L
    return L


class RefTrap:
    """PenningTrap of the reference (Source/PenningTrap.hpp:54-98)."""

    def __init__(self, radius, lengths, potentials, gaps, Nz, Nr):
        L = lib()
        lengths = np.ascontiguousarray(lengths, dtype=np.float64)
        potentials = np.ascontiguousarray(potentials, dtype=np.float64)
        gaps = np.ascontiguousarray(gaps, dtype=np.float64)
        self.h = L.ref_trap_create(radius, len(lengths), lengths, potentials, len(gaps),
                                   gaps if len(gaps) else np.zeros(1), Nz, Nr)
        if not self.h:
            raise ValueError(L.ref_last_error().decode())
        nz, nr, hz, hr, ln, rad = C.c_int(), C.c_int(), C.c_double(), C.c_double(), C.c_double(), C.c_double()
        L.ref_trap_info(self.h, nz, nr, hz, hr, ln, rad)
        self.Nz, self.Nr, self.hz, self.hr, self.length, self.radius = nz.value, nr.value, hz.value, hr.value, ln.value, rad.value
        self.G = (self.Nz + 1) * self.Nr
        self.plasmas = []

    def _grid(self, fn):
        out = np.empty(self.G)
        fn(self.h, out)
        return out

    def phi(self):
        return self._grid(lib().ref_trap_get_phi)

    def wall_rhs(self):
        return self._grid(lib().ref_trap_get_wall_rhs)

    def enodes(self):
        return self._grid(lib().ref_trap_get_enodes)

    def limits(self):
        a, b = np.empty(self.Nr, np.int32), np.empty(self.Nr, np.int32)
        lib().ref_trap_limits(self.h, a, b)
        return a, b

    def set_potential(self, idx, v):
        lib().ref_trap_set_potential(self.h, idx, v)

    def solve(self, rhs):
        out = np.empty(self.G)
        lib().ref_trap_solve(self.h, np.ascontiguousarray(rhs, dtype=np.float64), out)
        return out

    def apply(self, x):
        out = np.empty(self.G)
        lib().ref_trap_apply(self.h, np.ascontiguousarray(x, dtype=np.float64), out)
        return out

    def matrix(self):
        n = lib().ref_trap_matrix_nnz(self.h)
        r, c, v = np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n)
        lib().ref_trap_matrix(self.h, r, c, v)
        return r, c, v

    def efield(self, r, z):
        return lib().ref_trap_efield(self.h, int(r), float(z))

    def total_phi(self, r, z):
        return lib().ref_trap_total_phi(self.h, int(r), float(z))

    def move_plasmas(self, dt, nsteps=1):
        lib().ref_trap_move_plasmas(self.h, dt, nsteps)

    def save_states(self, t, r=None):
        if r is None:
            lib().ref_trap_save_states(self.h, t)
        else:
            lib().ref_trap_save_states_r(self.h, t, r)

    def last_potential_energy(self):
        return lib().ref_trap_last_potential_energy(self.h)

    def timed_steps(self, dt, nsteps):
        sec = np.zeros(4)
        ring_steps = lib().ref_trap_timed_steps(self.h, dt, nsteps, sec)
        return ring_steps, sec

    def plasma(self, name, mass, charge):
        p = RefPlasma(self, name, mass, charge)
        self.plasmas.append(p)
        return p

    def close(self):
        if self.h:
            for p in self.plasmas:
                lib().ref_plasma_destroy(p.h)
                p.h = None
            lib().ref_trap_destroy(self.h)
            self.h = None


class RefPlasma:
    """Plasma of the reference (Source/Plasma.hpp:139-197)."""

    def __init__(self, trap, name, mass, charge):
        self.trap = trap
        self.mass, self.charge = mass, charge
        self.h = lib().ref_plasma_create(trap.h, name.encode(), mass, charge)

    def load_profile(self, T, total_charge, shape, scale, num_macro, ks):
        if lib().ref_plasma_load_profile(self.h, T, total_charge, shape, scale, num_macro, ks):
            raise ValueError(lib().ref_last_error().decode())

    def load_density_file(self, path, T, num_macro):
        if lib().ref_plasma_load_density_file(self.h, str(path).encode(), T, num_macro):
            raise ValueError(lib().ref_last_error().decode())

    def set_rings(self, r, z, v, charge_macro):
        r = np.ascontiguousarray(r, dtype=np.int32)
        z = np.ascontiguousarray(z, dtype=np.float64)
        v = np.ascontiguousarray(v, dtype=np.float64)
        lib().ref_plasma_set_rings(self.h, len(r), r, z, v, charge_macro)

    def count(self):
        return lib().ref_plasma_count(self.h)

    def num_central_well(self):
        return lib().ref_plasma_num_central_well(self.h)

    def rings(self):
        n = self.count()
        r, z, v = np.empty(n, np.int32), np.empty(n), np.empty(n)
        lib().ref_plasma_get_rings(self.h, r, z, v)
        return r, z, v

    def params(self):
        a, b, c, t = C.c_double(), C.c_double(), C.c_double(), C.c_double()
        lib().ref_plasma_params(self.h, a, b, c, t)
        return dict(chargeMacro=a.value, macroChargeDensity=b.value, massMacro=c.value, temperature=t.value)

    def _grid(self, fn):
        out = np.empty(self.trap.G)
        fn(self.h, out)
        return out

    def rhs(self):
        return self._grid(lib().ref_plasma_get_rhs)

    def self_potential(self):
        return self._grid(lib().ref_plasma_get_self_potential)

    def set_self_potential(self, phi):
        lib().ref_plasma_set_self_potential(self.h, np.ascontiguousarray(phi, dtype=np.float64))

    def initial_density(self):
        return self._grid(lib().ref_plasma_get_initial_density)

    def update_rhs(self):
        lib().ref_plasma_update_rhs(self.h)

    def solve_poisson(self):
        lib().ref_plasma_solve_poisson(self.h)

    def move_rings(self, dt):
        lib().ref_plasma_move_rings(self.h, dt)

    def temperature(self):
        return lib().ref_plasma_temperature(self.h)

    def potential_energy(self):
        return lib().ref_plasma_potential_energy(self.h)


# Constants of Source/Constants.hpp:11-16 (values, not code).
E_POS = 1.602176634e-19
EPSILON0 = 8.8541878128e-12
MASS_E = 9.1093837015e-31
MASS_P = 1.67262192369e-27
PI = 3.141592653589793238463
KB = 1.380649e-23


def default_trap(Nz=585, Nr=128):
    """Driver-A trap (Diagnostics/A) Grid Size and Plasma Period.txt:57-69)."""
    return RefTrap(0.01488, [0.01322] * 5, [0, -70, -15, -70, 0], [0.0005] * 4, Nz, Nr)
