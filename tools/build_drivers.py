"""Compile the reference's four example programs UNCHANGED against this repo's host classes.

The reference ships its drivers as text files (Diagnostics/A..D) *.txt) that are complete C++ translation
units. They are compiled where they lie (never copied) with -x c++ against
pic-trapped-plasma_b200/host/{PenningTrap,Plasma,Constants}.hpp and linked to libptp_host.so /
libptp_b200.so -> build/drivers/driver_{A,B,C,D}. That they build at all is the source-compatibility check
of the drop-in boundary; tests/test_gpu_drivers.py runs them on the GPU box.
Only possible where /root/reference exists (the dev container); the binaries travel with the snapshot.
"""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("PTP_REFERENCE", "/root/reference")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def build(verbose=False):
    diag = os.path.join(REFERENCE, "Diagnostics")
    out_dir = os.path.join(ROOT, "build", "drivers")
    os.makedirs(out_dir, exist_ok=True)
    pkg = os.path.join(ROOT, "pic-trapped-plasma_b200")
    built = []
    # this repo's own test programs over the same class surface (tests/drivers/*.cpp)
    for src in sorted(glob.glob(os.path.join(ROOT, "tests", "drivers", "*.cpp"))):
        out = os.path.join(out_dir, os.path.splitext(os.path.basename(src))[0])
        cmd = [CXX, "-std=c++17", "-O2", src, "-I" + os.path.join(pkg, "host"), "-I" + os.path.join(ROOT, "include"),
               "-L" + pkg, "-lptp_host", "-lptp_b200", "-Wl,-rpath," + pkg, "-Wl,-rpath,$ORIGIN/../../pic-trapped-plasma_b200", "-o", out]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        built.append(out)
    if not os.path.isdir(diag):
        return built
    for letter in "ABCD":
        src = glob.glob(os.path.join(diag, letter + ")*.txt"))
        if not src:
            continue
        out = os.path.join(out_dir, "driver_" + letter)
        cmd = [CXX, "-std=c++17", "-O2", "-x", "c++", src[0], "-x", "none", "-I" + os.path.join(pkg, "host"),
               "-L" + pkg, "-lptp_host", "-lptp_b200", "-Wl,-rpath," + pkg, "-Wl,-rpath,$ORIGIN/../../pic-trapped-plasma_b200", "-o", out]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        built.append(out)
    return built


if __name__ == "__main__":
    print("\n".join(build(verbose="-v" in sys.argv)))
