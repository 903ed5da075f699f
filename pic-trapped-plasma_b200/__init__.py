"""ctypes binding of libptp_b200.so (C ABI in include/ptp.h) for tests/ and bench.py.

The product's host side is C++ (pic-trapped-plasma_b200/host: the reference's Electrode /
PenningTrap / Plasma class surface over the same C ABI).  This module is the thin Python
mirror of that surface used by the parity tests and the benchmark: same class and method
names as the reference (Source/PenningTrap.hpp:54-98, Source/Plasma.hpp:139-197) where a
method exists on the hot path, plus the upload / download / parity hooks of the ABI.

There is no fallback: importing works without a GPU (so the ABI can be inspected), but
every compute entry point raises PtpError when the CUDA library or a device is missing.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PTP_LIB") or os.path.join(_HERE, "libptp_b200.so")  # PTP_LIB: A/B builds of the same ABI

# Source/Constants.hpp:11-16 (values)
ePos = 1.602176634e-19
epsilon = 8.8541878128e-12
massE = 9.1093837015e-31
massP = 1.67262192369e-27
PI = 3.141592653589793238463
KB = 1.380649e-23

PTP_DEPOSIT_FP64, PTP_DEPOSIT_FIXED64 = 0, 1
PTP_ARITH_FAST, PTP_ARITH_EXACT = 0, 1
PTP_SOLVER_DIRECT, PTP_SOLVER_SOR, PTP_SOLVER_DIRECT_FFT = 0, 1, 2


class PtpError(RuntimeError):
    pass


_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_lib = None

# name -> (restype, argtypes); also the list tests check against include/ptp.h
_vp, _d, _i, _i64 = C.c_void_p, C.c_double, C.c_int, C.c_int64
_pvp = C.POINTER(C.c_void_p)
SIGNATURES = {
    "ptp_last_error": (C.c_char_p, []),
    "ptp_version": (_i, []),
    "ptp_device_count": (_i, []),
    "ptp_trap_create": (_i, [_pvp, _i, _i, _d, _d, _d, _d, _i]),
    "ptp_trap_destroy": (_i, [_vp]),
    "ptp_trap_set_wall": (_i, [_vp, _vp]),
    "ptp_trap_set_wall_basis": (_i, [_vp, _i, _vp]),
    "ptp_trap_set_wall_weights": (_i, [_vp, _vp]),
    "ptp_trap_step_programme": (_i, [_vp, _d, _i, _vp]),
    "ptp_trap_solve": (_i, [_vp, _vp, _vp]),
    "ptp_trap_apply": (_i, [_vp, _vp, _vp]),
    "ptp_trap_get_phi": (_i, [_vp, _vp]),
    "ptp_trap_set_phi": (_i, [_vp, _vp]),
    "ptp_trap_get_enodes": (_i, [_vp, _vp]),
    "ptp_trap_step": (_i, [_vp, _d, _i]),
    "ptp_trap_push_deposit": (_i, [_vp, _d]),
    "ptp_trap_solve_fields": (_i, [_vp]),
    "ptp_trap_sync": (_i, [_vp]),
    "ptp_trap_last_times": (_i, [_vp, _vp]),
    "ptp_trap_set_phase_events": (_i, [_vp, _i]),
    "ptp_trap_last_launches": (_i64, [_vp]),
    "ptp_trap_sort": (_i, [_vp]),
    "ptp_trap_set_sort_interval": (_i, [_vp, _i]),
    "ptp_trap_sorts_done": (_i64, [_vp]),
    "ptp_plasma_set_hot": (_i, [_vp, _i]),
    "ptp_plasma_is_hot": (_i, [_vp]),
    "ptp_trap_set_deposit_mode": (_i, [_vp, _i]),
    "ptp_trap_set_arith_mode": (_i, [_vp, _i]),
    "ptp_trap_set_solver": (_i, [_vp, _i, _d, _i]),
    "ptp_trap_set_tuning": (_i, [_vp, _i, _i, _i, _i]),
    "ptp_trap_set_graph": (_i, [_vp, _i]),
    "ptp_comm_unique_id": (_i, [_vp]),
    "ptp_trap_comm_init": (_i, [_vp, _vp, _i, _i]),
    "ptp_trap_set_allreduce": (_i, [_vp, _i]),
    "ptp_plasma_create": (_i, [_vp, _pvp, _d, _d]),
    "ptp_plasma_destroy": (_i, [_vp]),
    "ptp_plasma_upload": (_i, [_vp, _i64, _vp, _vp, _vp, _d]),
    "ptp_plasma_load_density": (_i, [_vp, _vp, _d, _i64, _i, _i, C.POINTER(_d), _vp, C.POINTER(_i64)]),
    "ptp_plasma_deposit_solve": (_i, [_vp]),
    "ptp_plasma_deposit": (_i, [_vp]),
    "ptp_plasma_count": (_i, [_vp, C.POINTER(_i64)]),
    "ptp_plasma_download": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "ptp_plasma_cell_index": (_i, [_vp, _vp, _vp]),
    "ptp_plasma_get_rhs": (_i, [_vp, _vp]),
    "ptp_plasma_get_self_potential": (_i, [_vp, _vp]),
    "ptp_plasma_set_self_potential": (_i, [_vp, _vp]),
    "ptp_plasma_potential_energy": (_i, [_vp, _d, C.POINTER(_d)]),
    "ptp_plasma_count_central_well": (_i, [_vp, _vp, _vp, C.POINTER(_i64)]),
    "ptp_plasma_kinetic_sums": (_i, [_vp, _i, C.POINTER(_d), C.POINTER(_d), C.POINTER(_i)]),
    "ptp_plasma_download_row": (_i, [_vp, _i, _i64, _vp, _vp, _vp, C.POINTER(_i64)]),
    "ptp_plasma_loss_log": (_i, [_vp, _i64, _i64, _vp, _vp, C.POINTER(_i64), C.POINTER(_i)]),
}


def lib():
    """Load libptp_b200.so; raises PtpError (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PtpError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise PtpError("ptp error %d: %s" % (rc, lib().ptp_last_error().decode()))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Electrode:
    """Source/PenningTrap.hpp:38-51."""

    def __init__(self, aLength, aPotential):
        self.length, self.potential = float(aLength), float(aPotential)

    def getLength(self):
        return self.length

    def getPotential(self):
        return self.potential

    def setPotential(self, v):
        self.potential = float(v)


class PenningTrap:
    """Device twin of the reference's PenningTrap (Source/PenningTrap.hpp:54-98), hot-path methods only."""

    def __init__(self, radius, theElectrodes, theGaps, NumCellsZ, NumCellsR, device=0):
        self.h = None
        if len(theElectrodes) != len(theGaps) + 1:  # Source/PenningTrap.cpp:39-42
            raise ValueError("Error number of gaps and electrodes; No. electrods should match No. gaps + 1")
        self.trapRadius = float(radius)
        self.electrodes = [Electrode(e.getLength(), e.getPotential()) for e in theElectrodes]
        self.gaps = [float(g) for g in theGaps]
        self.Nz, self.Nr = int(NumCellsZ), int(NumCellsR)
        lengthTrap = 0.0
        for i, e in enumerate(self.electrodes):     # Source/PenningTrap.cpp:43-50
            lengthTrap += e.getLength()
            if i < len(self.gaps):
                lengthTrap += self.gaps[i]
        self.lengthTrap = lengthTrap
        self.hz = lengthTrap / self.Nz              # :51
        self.hr = self.trapRadius / self.Nr         # :52
        self.G = (self.Nz + 1) * self.Nr
        self.plasmas = []
        self.basis = False
        h = C.c_void_p()
        _check(lib().ptp_trap_create(C.byref(h), self.Nz, self.Nr, self.hz, self.hr, self.lengthTrap, self.trapRadius, device))
        self.h = h
        self.solveLaplace()

    # -- electrode boundary conditions ------------------------------------------------------------
    def wallPotential(self):
        """The `boundary` value per axial node of PenningTrap::updateRHS (Source/PenningTrap.cpp:169-197)."""
        hz, Nz = self.hz, self.Nz
        wall = np.zeros(Nz + 1)
        point, totalLength = 0, 0.0
        for i, e in enumerate(self.electrodes):
            boundary = e.getPotential()
            while point * hz <= e.getLength() + totalLength:
                if point <= Nz:
                    wall[point] = boundary
                point += 1
            if i < len(self.gaps):
                nxt = self.electrodes[i + 1]
                while point * hz < e.getLength() + self.gaps[i] + totalLength:
                    boundary = (point * hz - e.getLength() - totalLength) * (nxt.getPotential() - e.getPotential()) / self.gaps[i] + e.getPotential()
                    if point <= Nz:
                        wall[point] = boundary
                    point += 1
                totalLength += e.getLength() + self.gaps[i]
        while point < Nz + 1:
            wall[point] = self.electrodes[-1].getPotential()
            point += 1
        return wall

    def solveLaplace(self):
        wall = self.wallPotential()
        _check(lib().ptp_trap_set_wall(self.h, _ptr(wall)))

    def setPotential(self, indexElectrode, newPotential):  # Source/PenningTrap.cpp:313-317
        self.electrodes[indexElectrode].setPotential(newPotential)
        if self.basis:
            w = np.array([e.getPotential() for e in self.electrodes])
            _check(lib().ptp_trap_set_wall_weights(self.h, _ptr(w)))
        else:
            self.solveLaplace()

    def useElectrodeBasis(self):
        """Electrode programmes: one Laplace solution per electrode at 1 V is kept on the device, after which
        setPotential is one axpy kernel and movePlasmasProgramme runs a whole voltage schedule without host round trips."""
        saved = [e.getPotential() for e in self.electrodes]
        walls = []
        for i in range(len(self.electrodes)):
            for j, e in enumerate(self.electrodes):
                e.setPotential(1.0 if i == j else 0.0)
            walls.append(self.wallPotential())
        for e, v in zip(self.electrodes, saved):
            e.setPotential(v)
        walls = _f64(np.stack(walls))
        _check(lib().ptp_trap_set_wall_basis(self.h, len(self.electrodes), _ptr(walls)))
        self.basis = True
        _check(lib().ptp_trap_set_wall_weights(self.h, _ptr(np.array(saved, dtype=np.float64))))

    def movePlasmasProgramme(self, deltaT, potentials):
        """potentials[s][i] = potential of electrode i during step s (setPotential calls + movePlasmas of the driver-D loop)."""
        w = _f64(potentials)
        if w.ndim != 2 or w.shape[1] != len(self.electrodes) or not self.basis:
            raise ValueError("movePlasmasProgramme: needs useElectrodeBasis() and potentials[steps][electrodes]")
        _check(lib().ptp_trap_step_programme(self.h, float(deltaT), w.shape[0], _ptr(w)))
        for e, v in zip(self.electrodes, w[-1]):
            e.setPotential(float(v))

    def getLength(self):
        return self.lengthTrap

    def getRadius(self):
        return self.trapRadius

    # -- the hot path ---------------------------------------------------------------------------------
    def movePlasmas(self, deltaT, nSteps=1):  # Source/PenningTrap.cpp:352-363
        _check(lib().ptp_trap_step(self.h, float(deltaT), int(nSteps)))

    def push_deposit(self, deltaT):
        _check(lib().ptp_trap_push_deposit(self.h, float(deltaT)))

    def solve_fields(self):
        _check(lib().ptp_trap_solve_fields(self.h))

    def sync(self):
        _check(lib().ptp_trap_sync(self.h))

    def last_times(self):
        out = np.zeros(4)
        _check(lib().ptp_trap_last_times(self.h, _ptr(out)))
        return out

    def set_phase_events(self, on=True):
        _check(lib().ptp_trap_set_phase_events(self.h, 1 if on else 0))

    def last_launches(self):
        return int(lib().ptp_trap_last_launches(self.h))

    def sort(self):
        _check(lib().ptp_trap_sort(self.h))

    # -- grids -------------------------------------------------------------------------------------------
    def _get(self, fn):
        out = np.empty(self.G)
        _check(fn(self.h, _ptr(out)))
        return out

    def phi(self):
        return self._get(lib().ptp_trap_get_phi)

    def set_phi(self, phi):
        phi = _f64(phi)
        _check(lib().ptp_trap_set_phi(self.h, _ptr(phi)))

    def enodes(self):
        return self._get(lib().ptp_trap_get_enodes)

    def solve(self, rhs):
        rhs, out = _f64(rhs), np.empty(self.G)
        _check(lib().ptp_trap_solve(self.h, _ptr(rhs), _ptr(out)))
        return out

    def apply(self, x):
        x, out = _f64(x), np.empty(self.G)
        _check(lib().ptp_trap_apply(self.h, _ptr(x), _ptr(out)))
        return out

    # -- modes ---------------------------------------------------------------------------------------------
    def set_deposit_mode(self, mode):
        _check(lib().ptp_trap_set_deposit_mode(self.h, mode))

    def set_arith_mode(self, mode):
        _check(lib().ptp_trap_set_arith_mode(self.h, mode))

    def set_solver(self, solver, tol=0.0, max_iter=0):
        _check(lib().ptp_trap_set_solver(self.h, solver, tol, max_iter))

    def set_tuning(self, threads=0, window=0, ctas=-1, rings_per_thread=0):
        _check(lib().ptp_trap_set_tuning(self.h, threads, window, ctas, rings_per_thread))

    def set_graph(self, on=True):
        """True / 1: replay every step as a CUDA graph; False / 0: never; -1 or None: automatic (the library's default policy)."""
        mode = -1 if (on is None or (not isinstance(on, bool) and on < 0)) else (1 if on else 0)
        _check(lib().ptp_trap_set_graph(self.h, mode))

    def set_sort_interval(self, interval):
        """> 0: re-sort every `interval` steps; 0: never; -1 (default): adaptive (see include/ptp.h)."""
        _check(lib().ptp_trap_set_sort_interval(self.h, interval))

    def sorts_done(self):
        return int(lib().ptp_trap_sorts_done(self.h))

    def set_allreduce(self, kind):
        """0: NCCL all-reduce of the deposit grids; 1: peer-memory mode (the push kernel adds into every rank's grid)."""
        _check(lib().ptp_trap_set_allreduce(self.h, kind))

    def comm_init(self, unique_id, n_ranks, rank):
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        _check(lib().ptp_trap_comm_init(self.h, C.cast(buf, C.c_void_p), n_ranks, rank))

    def close(self):
        if self.h:
            for p in list(self.plasmas):
                p.h = None
            lib().ptp_trap_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def comm_unique_id():
    buf = (C.c_char * 128)()
    _check(lib().ptp_comm_unique_id(C.cast(buf, C.c_void_p)))
    return bytes(buf)


class Plasma:
    """Device twin of the reference's Plasma (Source/Plasma.hpp:139-197), hot-path methods only."""

    def __init__(self, trap, name, mass, charge):
        self.refTrap, self.name, self.mass, self.charge = trap, name, float(mass), float(charge)
        self.chargeMacro = 0.0
        self.macroChargeDensity = 0.0
        h = C.c_void_p()
        _check(lib().ptp_plasma_create(trap.h, C.byref(h), self.mass, self.charge))
        self.h = h
        trap.plasmas.append(self)

    def set_hot(self, mode):
        """1: push this species with the per-warp-bin form of K1 (no re-sorts); 0: never; -1: let the re-sort policy decide."""
        _check(lib().ptp_plasma_set_hot(self.h, int(mode)))

    def is_hot(self):
        return bool(lib().ptp_plasma_is_hot(self.h))

    def upload(self, r, z, v, chargeMacro):
        """Replace the rings (what both loaders end with); macro quantities as Source/Plasma.cpp:492-494."""
        r = np.ascontiguousarray(r, dtype=np.int32)
        z, v = _f64(z), _f64(v)
        t = self.refTrap
        self.chargeMacro = float(chargeMacro)
        self.massMacro = self.chargeMacro * self.mass / self.charge
        self.macroChargeDensity = 4 * self.chargeMacro / (PI * t.hz * t.hr * t.hr)
        _check(lib().ptp_plasma_upload(self.h, len(r), _ptr(r), _ptr(z), _ptr(v), self.macroChargeDensity))

    def loadDensity(self, density, temperature, numMacro, shard=0, nShards=1, solve=True):
        """Placement + speeds of Plasma::loadDensityFile / loadProfile on the device (Source/Plasma.cpp:464-528) from the
        expected density grid; returns (rings loaded on this shard, rings per row over all shards)."""
        t = self.refTrap
        density = _f64(density)
        if density.size != t.G:
            raise ValueError("The number of grid points in the file do not match this trap.")   # Source/Plasma.cpp:556
        cm, n = C.c_double(), C.c_int64()
        per_row = np.zeros(t.Nr, np.int64)
        _check(lib().ptp_plasma_load_density(self.h, _ptr(density), float(temperature), int(numMacro), int(shard), int(nShards),
                                             C.byref(cm), _ptr(per_row), C.byref(n)))
        self.chargeMacro = cm.value
        self.massMacro = self.chargeMacro * self.mass / self.charge
        self.macroChargeDensity = 4 * self.chargeMacro / (PI * t.hz * t.hr * t.hr)
        if solve:
            self.solvePoisson()
        return n.value, per_row

    def solvePoisson(self):  # Source/Plasma.cpp:95-99
        _check(lib().ptp_plasma_deposit_solve(self.h))

    def updateRHS(self):  # Source/Plasma.cpp:77-94
        _check(lib().ptp_plasma_deposit(self.h))

    def getNumMacro(self):  # Source/Plasma.cpp:147-150
        n = C.c_int64()
        _check(lib().ptp_plasma_count(self.h, C.byref(n)))
        return n.value

    def download(self):
        n = self.getNumMacro()
        r, z, v, ids = np.empty(n, np.int32), np.empty(n), np.empty(n), np.empty(n, np.int64)
        _check(lib().ptp_plasma_download(self.h, _ptr(r), _ptr(z), _ptr(v), _ptr(ids)))
        return r, z, v, ids

    def cell_index(self):
        n = self.getNumMacro()
        k, idx = np.empty(n, np.int32), np.empty(n, np.int32)
        _check(lib().ptp_plasma_cell_index(self.h, _ptr(k), _ptr(idx)))
        return k, idx

    def _get(self, fn):
        out = np.empty(self.refTrap.G)
        _check(fn(self.h, _ptr(out)))
        return out

    def rhs(self):
        return self._get(lib().ptp_plasma_get_rhs)

    def selfPotential(self):
        return self._get(lib().ptp_plasma_get_self_potential)

    def set_self_potential(self, phi):
        phi = _f64(phi)
        _check(lib().ptp_plasma_set_self_potential(self.h, _ptr(phi)))

    def getPotentialEnergy(self):  # Source/Plasma.cpp:244-252
        pe = C.c_double()
        _check(lib().ptp_plasma_potential_energy(self.h, self.chargeMacro, C.byref(pe)))
        return pe.value

    def kineticSums(self, mark=False):
        """(sum w, sum w speed^2, paired) of ptp_plasma_kinetic_sums; temperature = mass * s2 / (KB * sw)."""
        sw, s2, paired = C.c_double(), C.c_double(), C.c_int()
        _check(lib().ptp_plasma_kinetic_sums(self.h, 1 if mark else 0, C.byref(sw), C.byref(s2), C.byref(paired)))
        return sw.value, s2.value, bool(paired.value)

    def saveSpeeds(self):
        """A save point of PenningTrap::saveStates for the temperature diagnostics; returns the temperature of the pair
        (previous save point, this one) as Plasma::getTemperature would (Source/Plasma.cpp:212-228), None at the first."""
        sw, s2, paired = self.kineticSums(mark=True)
        return (self.mass * s2 / (KB * sw)) if (paired and sw > 0) else None

    def downloadRow(self, row):
        t = self.refTrap
        n = C.c_int64()
        cap = self.getNumMacro()
        z, v, ids = np.empty(cap), np.empty(cap), np.empty(cap, np.int64)
        _check(lib().ptp_plasma_download_row(self.h, int(row), cap, _ptr(z), _ptr(v), _ptr(ids), C.byref(n)))
        return z[:n.value].copy(), v[:n.value].copy(), ids[:n.value].copy()

    def lossLog(self, first=0):
        total, over = C.c_int64(), C.c_int()
        _check(lib().ptp_plasma_loss_log(self.h, int(first), 0, None, None, C.byref(total), C.byref(over)))
        n = max(0, min(total.value, 1 << 16) - first)
        ids, steps = np.empty(n, np.int64), np.empty(n, np.int64)
        if n:
            _check(lib().ptp_plasma_loss_log(self.h, int(first), n, _ptr(ids), _ptr(steps), C.byref(total), C.byref(over)))
        return ids, steps, total.value, bool(over.value)

    def getNumMacroCentralWell(self, limitLeft, limitRight):  # Source/Plasma.cpp:151-162
        a = np.ascontiguousarray(limitLeft, dtype=np.int32)
        b = np.ascontiguousarray(limitRight, dtype=np.int32)
        n = C.c_int64()
        _check(lib().ptp_plasma_count_central_well(self.h, _ptr(a), _ptr(b), C.byref(n)))
        return n.value


def default_trap(Nz=585, Nr=128, device=0):
    """Driver-A trap (Diagnostics/A) Grid Size and Plasma Period.txt:57-69)."""
    el = [Electrode(0.01322, v) for v in (0, -70, -15, -70, 0)]
    return PenningTrap(0.01488, el, [0.0005] * 4, Nz, Nr, device=device)
