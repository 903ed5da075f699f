"""Host-side ring placement: the deterministic part of the reference's loaders, vectorised.

Both Plasma::loadProfile and Plasma::loadDensityFile end with the same placement code
(reference Source/Plasma.cpp:464-526 and :558-620): per radial row the cumulative charge along z,
the number of rings per row, and equally spaced quantiles inverted by linear interpolation.
This module restates that placement with numpy so that benchmarks and multi-GPU shards can build
loads "of the named shape" without the CPU reference; speeds are drawn from numpy's generator
(the reference uses libstdc++'s minstd_rand0 + normal_distribution, Source/Plasma.cpp:508-509 -
parity tests therefore upload rings produced by the compiled reference instead).
"""
import numpy as np

PI = 3.141592653589793238463
KB = 1.380649e-23


def ring_counts(dens, Nz, Nr, hz, hr, numMacro):
    """cumulativeAtR, chargeMacro, numAtR of Source/Plasma.cpp:466-500."""
    d = np.asarray(dens, dtype=np.float64).reshape(Nr, Nz + 1)
    volume = np.empty(Nr)
    volume[0] = PI * hz * hr * hr / 4                                   # :473
    idx = np.arange(1, Nr, dtype=np.float64)
    volume[1:] = hz * hr * 2 * PI * idx * hr                            # :477
    cum = np.cumsum(volume[:, None] * d, axis=1)                        # :480-484 (same left-to-right order)
    future = cum[0, -1] + float(np.sum(cum[1:, -1] / (8 * np.arange(1, Nr))))  # :487-491
    charge_macro = future / numMacro                                    # :492
    num_at_r = np.empty(Nr, dtype=np.int64)
    num_at_r[0] = int(round(cum[0, -1] / charge_macro))                 # :496
    num_at_r[1:] = np.round(cum[1:, -1] / (8 * np.arange(1, Nr) * charge_macro)).astype(np.int64)  # :499
    return cum, charge_macro, num_at_r


def place_rings(dens, Nz, Nr, hz, hr, numMacro, rank=0, n_ranks=1):
    """Rings of rank `rank` when every row's rings are dealt round-robin to `n_ranks` ranks
    (ring i of a row goes to rank i % n_ranks). Returns r (int32), z, chargeMacro, numAtR."""
    cum, charge_macro, num_at_r = ring_counts(dens, Nz, Nr, hz, hr, numMacro)
    rs, zs = [], []
    for j in range(Nr):
        n = int(num_at_r[j])
        if n <= 0:
            continue
        i = np.arange(rank, n, n_ranks, dtype=np.float64)
        if len(i) == 0:
            continue
        c = cum[j]
        delta_q = c[-1] / (n + 1)                                       # :512
        inv = delta_q * (i + 1)                                         # :516
        ci = np.searchsorted(np.abs(c), np.abs(inv), side="left")       # :517-520 first index with |cum| >= |inv|
        ci = np.clip(ci, 1, Nz)
        z = (ci - 1) * hz + hz / 2 + hz * (inv - c[ci - 1]) / (c[ci] - c[ci - 1])  # :523
        rs.append(np.full(len(i), j, dtype=np.int32))
        zs.append(z)
    if not rs:
        return np.zeros(0, np.int32), np.zeros(0), charge_macro, num_at_r
    return np.concatenate(rs), np.concatenate(zs), charge_macro, num_at_r


def maxwellian_speeds(n, temperature, mass, seed):
    """1-D Maxwellian, sigma = sqrt(kB T / m) (Source/Plasma.cpp:509)."""
    return np.random.default_rng(seed).normal(0.0, np.sqrt(KB * temperature / mass), n)
