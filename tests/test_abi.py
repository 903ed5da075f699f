"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/ptp.h
declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import importlib
import os
import re

import pytest

from conftest import ROOT

ptp = importlib.import_module("pic-trapped-plasma_b200")


def _declared():
    text = open(os.path.join(ROOT, "include", "ptp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ptp_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    names = _declared()
    assert len(names) >= 35
    L = C.CDLL(ptp.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_binding_covers_header():
    assert sorted(ptp.SIGNATURES) == _declared()


def test_every_declaration_cites_the_reference():
    text = open(os.path.join(ROOT, "include", "ptp.h")).read()
    assert text.count("Source/") >= 25


def test_version_and_error_string():
    L = ptp.lib()
    assert L.ptp_version() >= 100
    assert isinstance(L.ptp_last_error(), bytes)


def test_no_cpu_fallback_without_gpu():
    L = ptp.lib()
    if L.ptp_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = L.ptp_trap_create(C.byref(h), 585, 128, 1e-4, 1e-4, 0.0681, 0.01488, 0)
    assert rc == 2 and not h.value            # PTP_ECUDA
    assert b"no CPU fallback" in L.ptp_last_error()
    with pytest.raises(ptp.PtpError):
        ptp.default_trap()


def test_c_host_program_builds_against_the_abi_and_fails_loudly_without_a_gpu(tmp_path):
    """tools/hot_ab.cpp is a complete host program on the C ABI alone (trap from wall potentials, device-side load, steps,
    read-backs): it must build with nothing but include/ptp.h and the library, and - on a machine without a GPU - stop at
    ptp_trap_create with the library's error text instead of computing anything."""
    import shutil
    import subprocess
    import sys
    if shutil.which("g++") is None:
        pytest.skip("needs g++")
    exe = str(tmp_path / "hot_ab")
    pkg = os.path.join(ROOT, "pic-trapped-plasma_b200")
    b = subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", os.path.join(ROOT, "tools", "hot_ab.cpp"), "-I" + os.path.join(ROOT, "include"),
                        "-L" + pkg, "-lptp_b200", "-Wl,-rpath," + pkg, "-o", exe], capture_output=True, text=True, timeout=300)
    assert b.returncode == 0, b.stderr
    if ptp.lib().ptp_device_count() > 0:
        pytest.skip("a GPU is present: the program itself is run by tools/gpu_*.sh")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "hot_ab_prepare.py")], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert p.returncode == 0, p.stdout + p.stderr
    r = subprocess.run([exe, os.path.join(ROOT, "build", "hot_ab", "c5.bin")], capture_output=True, text=True, timeout=120, cwd=ROOT)
    assert r.returncode == 1 and "no CPU fallback" in r.stdout and "ms per step" not in r.stdout


def test_product_does_not_touch_the_oracle():
    """The product path may not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "pic-trapped-plasma_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".hpp", ".cpp", "Makefile")):
                src = open(os.path.join(base, f), errors="ignore").read()
                assert "oracle" not in src.replace("no oracle", ""), os.path.join(base, f)
                assert "libptp_ref" not in src and "ptp_oracle" not in src


@pytest.mark.skipif(not os.path.isdir("/root/reference/Diagnostics"), reason="reference only present in the dev container")
def test_reference_drivers_compile_unchanged_against_host_classes():
    """Diagnostics/A..D) *.txt are complete C++ programs; they must build against our headers without edits."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import build_drivers
    if not os.path.exists(os.path.join(ROOT, "pic-trapped-plasma_b200", "libptp_host.so")):
        pytest.skip("libptp_host.so not built")
    built = build_drivers.build()
    assert [os.path.basename(b) for b in built if "driver_" in b] == ["driver_A", "driver_B", "driver_C", "driver_D"]
