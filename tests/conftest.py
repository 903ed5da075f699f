import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running CPU check (still part of the default CPU suite)")


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    nb = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / nb) if nb > 0 else float(np.linalg.norm(a - b))


def expected_density():
    """Driver-A equilibrium density grid (tests/golden/make_golden.py)."""
    d = np.load(os.path.join(GOLDEN, "expected_density_nonzero.npz"))
    dens = np.zeros(int(d["G"]))
    dens[d["index"]] = d["value"]
    return dens


def write_density_file(path, dens, ratio):
    """Driver A's text files: `file << readNumber * ratio << '\\n'` at default ostream precision
    (Diagnostics/A) Grid Size and Plasma Period.txt:103-110); readNumber is the 15-digit text of
    extractInitialDensity."""
    with open(path, "w") as f:
        for x in dens:
            f.write("%.6g\n" % (float("%.15g" % x) * ratio))


@pytest.fixture(scope="session")
def density_files(tmp_path_factory):
    d = tmp_path_factory.mktemp("density")
    dens = expected_density()
    fe, fp = str(d / "electrons.csv"), str(d / "antiprotons.csv")
    write_density_file(fe, dens, 0.6)
    write_density_file(fp, dens, 1 - 0.6)
    return fe, fp


@pytest.fixture(scope="session")
def c1_kat():
    return np.load(os.path.join(GOLDEN, "c1_step_kat.npz"))
