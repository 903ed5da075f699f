"""CPU checks of device code that needs no GPU: the radix-16 inverse-DCT kernel's index maps / bank groups (numpy model)
and the kernel's own source text compiled for the host and run as one CTA of 256 threads (tests/emu/emu_r16.cpp)."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_BUILT = {}


def test_radix16_thread_model_matches_dct_and_is_conflict_free():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import fft16_model
    fft16_model.main()            # asserts rel-L2 < 1e-14 against the DCT-I sum and one access per bank group and quarter-warp


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ (C++20 std::barrier)")
def test_radix16_kernel_text_on_host_threads():
    r = subprocess.run(["bash", os.path.join(ROOT, "tests", "emu", "emu_r16.sh")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "other rows untouched: yes" in r.stdout


# ------------------------------------------------------------------------------------------ K1 on host threads
def _run_emu_push(tmp_path, trap, enodes, r, z, v, dt, charge, mass, W, WE, fixed, exact, fixed_bits=40, seg_tiles=2, n_cta=3, scatter=0):
    import numpy as np
    exe = os.path.join(ROOT, "build", "emu", "emu_push")
    if not _BUILT.get("push"):
        b = subprocess.run(["bash", os.path.join(ROOT, "tests", "emu", "emu_push.sh")], capture_output=True, text=True, timeout=600)
        assert b.returncode == 0, b.stdout + b.stderr
        _BUILT["push"] = True
    case, out = str(tmp_path / "case.bin"), str(tmp_path / "out.bin")
    with open(case, "wb") as f:
        f.write(np.array([trap.Nz, trap.Nr, W, WE, fixed, exact, fixed_bits, seg_tiles, n_cta, scatter], np.int32).tobytes())
        f.write(np.array([len(r)], np.int64).tobytes())
        f.write(np.array([trap.hz, trap.length, dt, charge, mass], np.float64).tobytes())
        for a, t in ((enodes, np.float64), (r, np.int32), (z, np.float64), (v, np.float64)):
            f.write(np.ascontiguousarray(a, dtype=t).tobytes())
    p = subprocess.run([exe, case, out], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    raw = open(out, "rb").read()
    n, G = len(r), trap.G
    n_seg = int(np.frombuffer(raw, np.int64, 1, 0)[0])
    lost = np.frombuffer(raw, np.uint64, 2, 8)
    zo = np.frombuffer(raw, np.float64, n, 24)
    vo = np.frombuffer(raw, np.float64, n, 24 + 8 * n)
    grid = raw[24 + 16 * n: 24 + 16 * n + 8 * G]
    bnd = np.frombuffer(raw, np.uint32, 2 * trap.Nr, 24 + 16 * n + 8 * G).reshape(trap.Nr, 2)
    seg_bounds = np.frombuffer(raw, np.int32, 4 * n_seg, 24 + 16 * n + 8 * (G + trap.Nr)).reshape(n_seg, 4)
    return zo, vo, grid, bnd, lost, seg_bounds


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ (C++20 std::barrier)")
@pytest.mark.parametrize("W,WE,fixed,exact,scatter,shuffle", [(44, 256, 0, 1, 0, 0), (44, 256, 0, 0, 0, 0), (44, 256, 1, 1, 0, 0), (6, 12, 0, 1, 0, 0), (6, 6, 1, 0, 0, 0),
                                                              (300, 300, 0, 1, 1, 0), (300, 300, 0, 0, 1, 1), (300, 300, 1, 0, 1, 1), (300, 300, 1, 1, 1, 0),
                                                              (24, 24, 0, 0, 1, 1), (24, 24, 1, 0, 1, 0)])
def test_push_kernel_text_on_host_threads_matches_oracle(tmp_path, W, WE, fixed, exact, scatter, shuffle):
    """The source text of k_push_deposit (pic-trapped-plasma_b200/csrc/ptp_push.cu) compiled for the host and run CTA by CTA
    on 512 threads, against the oracle on the C1 electrons plus fast rings near both trap ends (losses): positions / speeds
    ring by ring (EXACT arithmetic: bit for bit; FAST: 1e-14), loss count, deposit (fp64 1e-12; fixed point 2^-40 per ring),
    the touched node range per row. W = 6 forces most rings through the out-of-window paths (global gather / atomics).
    scatter = 1: the per-warp-bin form of the kernel (hot species; the rings of a warp grouped by a warp sort, two rings' 16-bit keys
    per network, the scan cut short at the longest run) - rings in load order (few distinct cells per warp: the
    warp-reduction path) and shuffled within their rows (many distinct cells: the rounds path); W = 24 leaves part of the
    plasma outside the window; in fixed-point mode the deposit grid equals the thread-private form's bit for bit."""
    import numpy as np
    sys.path.insert(0, ROOT)
    from oracle import port
    kat = np.load(os.path.join(ROOT, "tests", "golden", "c1_step_kat.npz"))
    dt, mass, charge = float(kat["dt"]), 9.1093837015e-31, -1.602176634e-19
    trap = port.default_trap()
    pl = trap.plasma("Electrons", mass, charge)
    rng = np.random.default_rng(3)
    extra = 300
    r = np.concatenate([kat["e_r0"], rng.integers(0, 5, extra).astype(np.int32)])
    edge = 2e-5 * (1 + 0.1 * rng.random(extra))
    z = np.concatenate([kat["e_z0"], np.where(rng.random(extra) < 0.5, edge, trap.length - edge)])
    v = np.concatenate([kat["e_v0"], rng.normal(0, 2e5, extra)])
    order = np.argsort(r, kind="stable")
    if shuffle:
        order = np.lexsort((rng.random(len(r)), r))
    r, z, v = r[order], z[order], v[order]
    pl.set_rings(r, z, v, float(kat["e_chargeMacro"]))
    pl.solve_poisson()
    enodes = trap.enodes()
    zo, vo, grid, bnd, lost, seg_bounds = _run_emu_push(tmp_path, trap, enodes, r, z, v, dt, charge, mass, W, WE, fixed, exact, scatter=scatter)
    # the reference's ring update, expression by expression (Source/PenningTrap.cpp:328-333, Source/Plasma.cpp:105-108)
    hz, n1 = trap.hz, trap.Nz + 1
    k = np.floor(z / hz).astype(np.int64)
    w = (z - k * hz) / hz
    e = (1 - w) * enodes[n1 * r + k] + w * enodes[n1 * r + k + 1]
    vn = dt * e * charge / mass + v
    zn = dt * vn + z
    keep = (zn < trap.length) & (zn > 0)
    assert int(lost[0]) == int((~keep).sum()) and int(lost[0]) > 0
    assert np.array_equal(np.isnan(zo), ~keep)
    if exact:
        assert np.array_equal(zo[keep], zn[keep]) and np.array_equal(vo[keep], vn[keep])
    else:
        assert np.max(np.abs(zo[keep] - zn[keep]) / zn[keep]) < 1e-14
        assert np.max(np.abs(vo[keep] - vn[keep]) / np.abs(vn[keep]).max()) < 1e-14
    # ... and the same survivors as the oracle's swap-with-back loop
    pl.move_rings(dt, enodes)
    assert pl.count() == int(keep.sum())
    assert np.array_equal(np.sort(pl.z), np.sort(zn[keep]))
    # deposit at the new positions in units of one ring; the oracle's RHS carries -rho_macro / eps0 (Source/Plasma.cpp:91-92)
    pl.update_rhs()
    scale = -pl.macro_charge_density / 8.8541878128e-12
    if fixed:
        got = np.frombuffer(grid, np.int64).astype(np.float64) / 2.0 ** 40
        tol = 1e-9
    else:
        got = np.frombuffer(grid, np.float64)
        tol = 1e-12
    assert np.linalg.norm(got * scale - pl.rhs) / np.linalg.norm(pl.rhs) < tol
    assert abs(got.sum() - keep.sum()) < 1e-6
    # touched node range per row = [min cell, max cell + 1] of its survivors, encoded as (Nz + 2 - kmin, kmax + 2)
    kn = np.minimum(np.floor(zn[keep] / hz).astype(np.int64), trap.Nz - 1)
    for row in np.unique(r[keep]):
        kk = kn[r[keep] == row]
        assert tuple(bnd[row]) == (trap.Nz + 2 - kk.min(), kk.max() + 2)
    assert not bnd[np.setdiff1d(np.arange(trap.Nr), r[keep])].any()
    # out-of-window counter: nothing misses a 44-cell window around a 37-cell plasma except the rings parked at the trap ends
    if W == 6:
        assert int(lost[1]) > 1000
    if scatter:
        # rings beyond the window: only (some of) those parked at the trap ends when the window covers the plasma with its margin
        assert (int(lost[1]) > 300) == (W == 24), int(lost[1])
        if fixed:
            (tmp_path / "ref").mkdir()                     # (the case file of the scatter run stays where tests/emu/sanitize_all.sh looks for it)
            ref = _run_emu_push(tmp_path / "ref", trap, enodes, r, z, v, dt, charge, mass, 44, 256, fixed, exact)
            assert ref[2] == grid and np.array_equal(ref[0], zo, equal_nan=True) and np.array_equal(ref[1], vo)
    trap.close()


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ (C++20 std::barrier)")
def test_hot_form_count_field_holds_the_longest_segment(tmp_path):
    """The per-warp bins of the hot form pack (rings : 12 bits | weight sum : 52 bits). The planner caps a segment of a hot
    species at 31 tiles (ptp_build_segments: 4095 / (32 lanes x 4 rings per tile)); the worst case for the count field is a
    segment of that length whose rings ALL sit in one cell: 31 x 128 = 3968 rings per warp and bin. One CTA, one such segment,
    fixed-point deposits: the grid must equal the thread-private form's bit for bit and hold every ring."""
    import numpy as np
    sys.path.insert(0, ROOT)
    from oracle import port
    trap = port.default_trap()
    n = 31 * 2048
    rng = np.random.default_rng(5)
    r = np.zeros(n, np.int32)
    z = (300 + rng.random(n)) * trap.hz                                  # all in cell 300
    v = np.zeros(n)
    enodes = np.zeros(trap.G)
    args = (trap, enodes, r, z, v, 1e-10, -1.602176634e-19, 9.1093837015e-31)
    hot = _run_emu_push(tmp_path, *args, 200, 200, 1, 0, seg_tiles=31, n_cta=1, scatter=1)
    (tmp_path / "ref").mkdir()
    ref = _run_emu_push(tmp_path / "ref", *args, 44, 256, 1, 0, seg_tiles=31, n_cta=1)
    assert hot[2] == ref[2]
    grid = np.frombuffer(hot[2], np.int64)
    assert grid.sum() == n << 40 and np.count_nonzero(grid) == 2        # nodes 300 and 301 of row 0
    assert np.array_equal(hot[0], z) and int(hot[4][0]) == 0 and int(hot[4][1]) == 0
    trap.close()


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ (C++20 std::barrier)")
def test_hot_form_widest_window_on_host_threads(tmp_path):
    """The widest window the hot form can be given (1440 cells: ptp_push_scatter_window on a 227 KB SM) on a row of 2000 cells
    with rings all over it: cell offsets up to 1439 (all 11 bits of the packed sort keys), rings beyond the window on
    the global path, rings in arbitrary order. Fixed point: bit for bit the grid of the thread-private form."""
    import types
    import numpy as np
    Nz, Nr, hz = 2000, 2, 1e-5
    trap = types.SimpleNamespace(Nz=Nz, Nr=Nr, hz=hz, length=Nz * hz, G=(Nz + 1) * Nr)
    rng = np.random.default_rng(9)
    n = 6000
    r = np.sort(rng.integers(0, Nr, n)).astype(np.int32)
    z = rng.uniform(5.5, Nz - 5.5, n) * hz
    v = rng.normal(0, 3e4, n)                                            # up to ~1 cell per step
    enodes = rng.normal(0, 50.0, trap.G)
    args = (trap, enodes, r, z, v, 2e-10, -1.602176634e-19, 9.1093837015e-31)
    hot = _run_emu_push(tmp_path, *args, 1440, 1440, 1, 0, seg_tiles=2, n_cta=2, scatter=1)
    (tmp_path / "ref").mkdir()
    ref = _run_emu_push(tmp_path / "ref", *args, 44, 256, 1, 0, seg_tiles=2, n_cta=2)
    assert hot[2] == ref[2] and np.array_equal(hot[0], ref[0]) and np.array_equal(hot[1], ref[1])
    assert np.frombuffer(hot[2], np.int64).sum() == n << 40 and int(hot[4][0]) == 0
    assert 0 < int(hot[4][1]) < n // 2                                   # some rings lie beyond the window, most inside
    assert np.array_equal(hot[3], ref[3])                                # touched node range per row
    trap_cells = np.floor(hot[0] / hz).astype(np.int64)
    assert trap_cells.max() - trap_cells.min() > 1500


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ (C++20 std::barrier)")
def test_sort_kernel_text_on_host_threads():
    """K5's four kernels (count / scan / scatter / pad, pic-trapped-plasma_b200/csrc/ptp_particles.cu) on host threads, driven
    like ptp_sort_plasma over three rounds with losses in between: rows ordered by axial cell, ring multiset intact (z, v
    travel with their id), live counts, empty-slot pattern behind every live prefix of the re-used alternate buffers."""
    r = subprocess.run(["bash", os.path.join(ROOT, "tests", "emu", "emu_sort.sh")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("padding clean") == 3


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ (C++20 std::barrier)")
@pytest.mark.parametrize("Nz,Nr,rows,k_lo,k_hi", [(256, 40, range(0, 10), 100, 160),       # compact radial path, radix-2 inverse
                                                  (256, 200, [3, 180, 199], 0, 256),       # streamed radial path (deposit reaching the wall row)
                                                  (4096, 40, range(0, 7), 1900, 2200)])    # radix-16 inverse, rows 32..39 formed in the inverse
def test_large_grid_solver_text_on_host_threads_matches_lu_oracle(tmp_path, Nz, Nr, rows, k_lo, k_hi):
    """The shipped table builder (ptp_solver_build) and the kernels of ptp_solve_wide.cu - forward DCT of the touched rows,
    radial solves with the rows above the deposit folded into one pivot, block-product expansion, FFT inverse + node field -
    run on host threads in launch order, against the oracle's LU solve of the reference matrix (phi rel-L2 <= 1e-10); the
    node field must be the centred difference of phi_trap + phi bit for bit (Source/PenningTrap.cpp:226-233)."""
    import numpy as np
    sys.path.insert(0, ROOT)
    from oracle import port
    if not _BUILT.get("wide"):
        b = subprocess.run(["bash", os.path.join(ROOT, "tests", "emu", "emu_wide.sh")], capture_output=True, text=True, timeout=900)
        assert b.returncode == 0, b.stdout + b.stderr
        _BUILT["wide"] = True
    pt = port.PortTrap(0.012, [0.02, 0.03, 0.02], [0.0, -50.0, 0.0], [0.001, 0.001], Nz, Nr)
    n1 = Nz + 1
    rng = np.random.default_rng(Nz + Nr)
    rho = np.zeros((Nr, n1))
    for j in rows:
        lo = k_lo + int(rng.integers(0, 5))
        rho[j, lo:k_hi + 1] = -1e6 * rng.random(k_hi + 1 - lo)
    rho = rho.reshape(-1)
    case, out = str(tmp_path / "case.bin"), str(tmp_path / "out.bin")
    with open(case, "wb") as f:
        f.write(np.array([Nz, Nr], np.int32).tobytes())
        f.write(np.array([pt.hz, pt.hr, pt.radius], np.float64).tobytes())
        f.write(rho.tobytes())
        f.write(np.ascontiguousarray(pt.phi).tobytes())
    p = subprocess.run([os.path.join(ROOT, "build", "emu", "emu_wide"), case, out], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout + p.stderr
    raw = np.fromfile(out, np.float64)
    G = pt.G
    phi, en, phi_formed = raw[:G], raw[G:2 * G], raw[2 * G:3 * G]
    want = pt.solve(rho)
    assert np.linalg.norm(phi - want) / np.linalg.norm(want) < 1e-10
    assert np.linalg.norm(phi_formed - phi) / np.linalg.norm(phi) < 1e-13
    tot = (pt.phi + phi).reshape(Nr, n1)
    e = np.zeros_like(tot)
    e[:, 1:-1] = (tot[:, :-2] - tot[:, 2:]) / (2 * pt.hz)
    assert np.array_equal(en.reshape(Nr, n1), e)
    # the step's form: only the populated rows (rounded up to blocks of 32) are produced - bit for bit the rows of the full solve
    limit = max(rows) + 1
    out2 = str(tmp_path / "out_rows.bin")
    p = subprocess.run([os.path.join(ROOT, "build", "emu", "emu_wide"), case, out2, str(limit), str(limit)], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout + p.stderr
    got = (limit + 31) // 32 * 32
    assert ("%d rows produced" % min(got, Nr)) in p.stdout
    raw2 = np.fromfile(out2, np.float64)
    keep = min(got, Nr) * n1
    assert np.array_equal(raw2[:keep], phi[:keep]) and np.array_equal(raw2[G:G + keep], en[:keep]) and np.array_equal(raw2[2 * G:2 * G + keep], phi_formed[:keep])
    if keep < G:
        assert not raw2[keep:G].any()                                   # nothing written beyond
    pt.close()


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ (C++20 std::barrier)")
@pytest.mark.parametrize("Nz,Nr,rows,k_lo,k_hi", [(585, 128, range(0, 12), 274, 311),     # the reference's default grid: fused inverse + node field
                                                  (300, 20, [0, 5, 19], 3, 297),           # odd row length: 8-byte copies
                                                  (57, 9, [0, 2, 3], 20, 40),              # very short even rows: bulk-staged inverse, two column tiles
                                                  (1500, 12, range(0, 4), 700, 800)])      # long rows, not a power of two: chunked inverse GEMM + k_node_field
def test_default_grid_solver_text_on_host_threads_matches_lu_oracle(tmp_path, Nz, Nr, rows, k_lo, k_hi):
    """The shipped table builder and the kernels of ptp_solve.cu (row-bounds scan, forward DCT fused with the Thomas solves,
    paired-mode inverse DCT fused with the node field / chunked inverse GEMM, stencil apply, wall right-hand side) on host
    threads with ptp_solver_run's launch arithmetic, against the oracle: phi vs the LU solve (rel-L2 <= 1e-10), A phi = b through
    k_apply, node field bit for bit, wall right-hand side equal to the oracle's (Source/PenningTrap.cpp:163-198)."""
    import numpy as np
    sys.path.insert(0, ROOT)
    from oracle import port
    if not _BUILT.get("solve"):
        b = subprocess.run(["bash", os.path.join(ROOT, "tests", "emu", "emu_solve.sh")], capture_output=True, text=True, timeout=900)
        assert b.returncode == 0, b.stdout + b.stderr
        _BUILT["solve"] = True
    pt = port.PortTrap(0.01488, [0.01322] * 5, [0, -70, -15, -70, 0], [0.0005] * 4, Nz, Nr)
    n1 = Nz + 1
    rng = np.random.default_rng(Nz + Nr)
    rho = np.zeros((Nr, n1))
    for j in rows:
        lo = k_lo + int(rng.integers(0, 3))
        rho[j, lo:k_hi + 1] = -1e6 * rng.random(k_hi + 1 - lo)
    rho = rho.reshape(-1)
    wall = pt.wall_potential()
    case, out = str(tmp_path / "case.bin"), str(tmp_path / "out.bin")
    with open(case, "wb") as f:
        f.write(np.array([Nz, Nr], np.int32).tobytes())
        f.write(np.array([pt.hz, pt.hr, pt.radius], np.float64).tobytes())
        f.write(rho.tobytes())
        f.write(np.ascontiguousarray(pt.phi).tobytes())
        f.write(np.ascontiguousarray(wall).tobytes())
    p = subprocess.run([os.path.join(ROOT, "build", "emu", "emu_solve"), case, out], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout + p.stderr
    raw = np.fromfile(out, np.float64)
    G = pt.G
    phi, en, aphi, wall_rhs = raw[:G], raw[G:2 * G], raw[2 * G:3 * G], raw[3 * G:4 * G]
    want = pt.solve(rho)
    assert np.linalg.norm(phi - want) / np.linalg.norm(want) < 1e-10
    assert np.linalg.norm(aphi - rho) / np.linalg.norm(rho) < 1e-10
    assert np.linalg.norm(aphi - pt.apply(phi)) / np.linalg.norm(rho) < 1e-13
    tot = (pt.phi + phi).reshape(Nr, n1)
    e = np.zeros_like(tot)
    e[:, 1:-1] = (tot[:, :-2] - tot[:, 2:]) / (2 * pt.hz)
    assert np.array_equal(en.reshape(Nr, n1), e)
    assert np.array_equal(wall_rhs, pt.wall_rhs())
    if n1 % 2 == 0 and "k_inv_field_bulk" in p.stdout:
        # even row length: the bulk-async (TMA) staged inverse ran; the cp.async form of the same kernel gives the same bits
        out0 = str(tmp_path / "out_cpasync.bin")
        p0 = subprocess.run([os.path.join(ROOT, "build", "emu", "emu_solve"), case, out0], capture_output=True, text=True, timeout=900,
                            env=dict(os.environ, PTP_INV_BULK="0"))
        assert p0.returncode == 0 and "k_inv_field_bulk" not in p0.stdout, p0.stdout + p0.stderr
        raw0 = np.fromfile(out0, np.float64)
        assert np.array_equal(raw0[:2 * G], raw[:2 * G])
    # rows above the outermost populated row folded into its pivot (the fold row is host knowledge: rings keep their row), whole
    # grid produced: the same solution; and the step's form, which stops after the populated rows rounded up to blocks of 32
    limit = max(rows) + 1
    out2, out3 = str(tmp_path / "out_fold.bin"), str(tmp_path / "out_rows.bin")
    p = subprocess.run([os.path.join(ROOT, "build", "emu", "emu_solve"), case, out2, str(limit)], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout + p.stderr
    fold = np.fromfile(out2, np.float64)
    assert np.linalg.norm(fold[:G] - want) / np.linalg.norm(want) < 1e-10
    assert np.linalg.norm(fold[:G] - phi) / np.linalg.norm(phi) < 1e-13
    p = subprocess.run([os.path.join(ROOT, "build", "emu", "emu_solve"), case, out3, str(limit), str(limit)], capture_output=True, text=True, timeout=900)
    if "chunked inverse" in p.stdout:
        assert p.returncode == 8                                        # (that path always produces whole grids)
    else:
        assert p.returncode == 0, p.stdout + p.stderr
        part = np.fromfile(out3, np.float64)
        keep = min((limit + 31) // 32 * 32, Nr) * n1
        assert np.array_equal(part[:keep], fold[:keep]) and np.array_equal(part[G:G + keep], fold[G:G + keep])   # bit for bit
        if keep < G:
            assert not part[keep:G].any()
    pt.close()


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ (C++20 std::barrier)")
def test_loader_deviate_stream_text_on_host_threads_equals_libstdcxx():
    """k_rng_count / k_rng_scan / k_rng_emit (ptp_load.cu) on host threads: the parallel, jump-ahead reproduction of the
    reference's speed deviates - std::default_random_engine + std::normal_distribution<double>, Source/Plasma.cpp:508-509 -
    against libstdc++'s own objects, bit for bit, for 1, 2, 4097 and 100001 deviates."""
    r = subprocess.run(["bash", os.path.join(ROOT, "tests", "emu", "emu_rng.sh")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("identical to std::normal_distribution") == 4


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ (C++20 std::barrier)")
@pytest.mark.parametrize("num_macro,mass,shard,n_shards", [(4000, 9.1093837015e-31, 0, 1), (100000, 1.67262192369e-27, 0, 1), (100000, 1.67262192369e-27, 2, 3)])
def test_loader_text_on_host_threads_matches_reference_loader(tmp_path, density_files, num_macro, mass, shard, n_shards):
    """ptp_plasma_load_density's host arithmetic and its placement kernel k_place (ptp_load.cu) on host threads against
    Plasma::loadDensityFile of the compiled reference (Source/Plasma.cpp:558-622): rings per row and chargeMacro exact,
    positions and speeds bit for bit (the host shares glibc's log with the reference), shard s of S = rings i = s (mod S) of
    every row."""
    import numpy as np
    sys.path.insert(0, ROOT)
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    if not _BUILT.get("place"):
        b = subprocess.run(["bash", os.path.join(ROOT, "tests", "emu", "emu_place.sh")], capture_output=True, text=True, timeout=600)
        assert b.returncode == 0, b.stdout + b.stderr
        _BUILT["place"] = True
    dens = np.loadtxt(density_files[0])
    rt = ref.default_trap()
    rp = rt.plasma("Species", mass, -ref.E_POS)
    rp.load_density_file(density_files[0], 150.0, num_macro)
    r0, z0, v0 = rp.rings()
    par = rp.params()
    case, out = str(tmp_path / "case.bin"), str(tmp_path / "out.bin")
    with open(case, "wb") as f:
        f.write(np.array([rt.Nz, rt.Nr, shard, n_shards], np.int32).tobytes())
        f.write(np.array([num_macro], np.int64).tobytes())
        f.write(np.array([rt.hz, rt.hr, 150.0, mass], np.float64).tobytes())
        f.write(np.ascontiguousarray(dens, dtype=np.float64).tobytes())
    p = subprocess.run([os.path.join(ROOT, "build", "emu", "emu_place"), case, out], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    raw = open(out, "rb").read()
    n = int(np.frombuffer(raw, np.int64, 1, 0)[0])
    charge_macro = float(np.frombuffer(raw, np.float64, 1, 8)[0])
    per_row = np.frombuffer(raw, np.int64, rt.Nr, 16)
    o = 16 + 8 * rt.Nr
    r = np.frombuffer(raw, np.int32, n, o)
    z = np.frombuffer(raw, np.float64, n, o + 4 * n)
    v = np.frombuffer(raw, np.float64, n, o + 12 * n)
    ids = np.frombuffer(raw, np.int64, n, o + 20 * n)
    assert charge_macro == par["chargeMacro"]
    assert np.array_equal(per_row, np.bincount(r0, minlength=rt.Nr))
    within = np.concatenate([np.arange(c) for c in per_row if c > 0])        # index of a ring inside its row, reference order
    mine = within % n_shards == shard
    assert n == int(mine.sum()) and np.array_equal(ids, np.arange(n))
    assert np.array_equal(r, r0[mine]) and np.array_equal(z, z0[mine]) and np.array_equal(v, v0[mine])
    rt.close()


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ (C++20 std::barrier)")
@pytest.mark.parametrize("Nz,Nr,rows,k_lo,k_hi,n_species", [(585, 128, range(0, 12), 274, 311, 2),    # the reference's default grid, two species
                                                           (300, 24, [0, 5, 19], 3, 297, 1),          # odd row length, long touched range (5 chunks)
                                                           (57, 9, [0, 2, 3], 20, 40, 3)])            # tiny grid: one cluster, mode pairs beyond K2 idle
def test_cluster_solve_kernel_text_on_host_threads_matches_lu_oracle(tmp_path, Nz, Nr, rows, k_lo, k_hi, n_species):
    """k_solve_cluster - the one-kernel step solve on thread-block clusters (ptp_solve_cluster.cu) - on host threads: the 16
    CTAs of a cluster run concurrently, cluster.sync is a barrier over all of them, distributed shared memory a pointer
    translation. All species at once; potentials of the populated rows against the oracle's LU solve (rel-L2 <= 1e-10), the
    node field bit for bit the centred difference of phi_trap + sum of the species' potentials in registration order
    (Source/PenningTrap.cpp:226-233); rows beyond the produced ones untouched."""
    import numpy as np
    sys.path.insert(0, ROOT)
    from oracle import port
    if not _BUILT.get("cluster"):
        b = subprocess.run(["bash", os.path.join(ROOT, "tests", "emu", "emu_cluster.sh")], capture_output=True, text=True, timeout=900)
        assert b.returncode == 0, b.stdout + b.stderr
        _BUILT["cluster"] = True
    pt = port.PortTrap(0.01488, [0.01322] * 5, [0, -70, -15, -70, 0], [0.0005] * 4, Nz, Nr)
    n1 = Nz + 1
    rng = np.random.default_rng(Nz + Nr)
    rho = np.zeros((n_species, Nr, n1))
    for s in range(n_species):
        for j in rows:
            lo = k_lo + int(rng.integers(0, 3))
            rho[s, j, lo:k_hi + 1 - s] = -1e6 * rng.random(k_hi + 1 - s - lo)
    case, out = str(tmp_path / "case.bin"), str(tmp_path / "out.bin")
    with open(case, "wb") as f:
        f.write(np.array([Nz, Nr], np.int32).tobytes())
        f.write(np.array([pt.hz, pt.hr, pt.radius], np.float64).tobytes())
        f.write(rho.tobytes())
        f.write(np.ascontiguousarray(pt.phi).tobytes())
    limit = max(rows) + 1
    p = subprocess.run([os.path.join(ROOT, "build", "emu", "emu_cluster"), case, out, str(limit), str(n_species)], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout + p.stderr
    raw = np.fromfile(out, np.float64)
    G = pt.G
    got = min((limit + 15) // 16 * 16, Nr)
    phi = raw[:n_species * G].reshape(n_species, Nr, n1)
    en = raw[n_species * G:].reshape(Nr, n1)
    tot = pt.phi.reshape(Nr, n1).copy()
    for s in range(n_species):
        want = pt.solve(rho[s].reshape(-1)).reshape(Nr, n1)
        assert np.linalg.norm(phi[s, :got] - want[:got]) / np.linalg.norm(want[:got]) < 1e-10
        assert not phi[s, got:].any()
        tot[:got] = tot[:got] + phi[s, :got]
    e = np.zeros((got, n1))
    e[:, 1:-1] = (tot[:got, :-2] - tot[:got, 2:]) / (2 * pt.hz)
    assert np.array_equal(en[:got], e)
    assert np.all(en[got:] == -1.0)
    pt.close()
