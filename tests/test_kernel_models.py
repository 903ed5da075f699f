"""CPU checks of device code that needs no GPU: the radix-16 inverse-DCT kernel's index maps / bank groups (numpy model)
and the kernel's own source text compiled for the host and run as one CTA of 256 threads (tools/emu_r16.cpp)."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_radix16_thread_model_matches_dct_and_is_conflict_free():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import fft16_model
    fft16_model.main()            # asserts rel-L2 < 1e-14 against the DCT-I sum and one access per bank group and quarter-warp


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ (C++20 std::barrier)")
def test_radix16_kernel_text_on_host_threads():
    r = subprocess.run(["bash", os.path.join(ROOT, "tools", "emu_r16.sh")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "other rows untouched: yes" in r.stdout
