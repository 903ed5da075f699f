"""bench.py's reference arm runs on the CPU (oracle/_ref or the C restatement): check the JSON line it prints against the
keys the measurement contract names. The GPU arm prints the same line plus roofline / clocks / gpu_launches (needs a B200)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--cpu-sample", "200000"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [x for x in p.stdout.splitlines() if x.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "particle-steps/s"
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"):
        assert key in d, key
    assert d["value"] > 1e6 and d["steps"] == 2 and d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "workload" in d["config"] and d["config"]["workload"].startswith("c4")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_k1_traffic_profile_belongs_to_the_shipped_push_kernel():
    """bench.py reports roofline.traffic only while profiles/k1_traffic.json was taken from the push kernel that ships: same
    source file (hash of ptp_push.cu), or - when other parts of the file have changed since the capture - the same machine code of
    the profiled instantiation inside libptp_b200.so (cuobjdump -sass). A change to the kernel without a new ncu capture must show
    up here, not as a silently stale number."""
    import hashlib
    sys.path.insert(0, ROOT)
    import bench
    prof = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))
    src = open(os.path.join(ROOT, "pic-trapped-plasma_b200", "csrc", "ptp_push.cu"), "rb").read()
    sha_src = hashlib.sha256(src).hexdigest()[:16]
    if prof["push_cu_sha16"] != sha_src:
        # a later version of the file: it must be listed as one whose K1 machine code equals the profiled one, and the shipped
        # library must really hold that machine code
        assert sha_src in prof.get("same_k1_sass_sources", []), "ptp_push.cu changed since the ncu capture: re-profile K1 or verify its SASS and list the new source hash"
        so = os.path.join(ROOT, "pic-trapped-plasma_b200", "libptp_b200.so")
        assert os.path.exists(so), "library not built: cannot compare the profiled kernel's machine code"
        sha = bench.k1_sass_sha(so)
        assert sha is not None and prof.get("k1_sass_sha16") == sha
    for wl in ("c4", "c5"):
        assert 28.0 < prof[wl]["dram_bytes_per_ring"] < 36.0        # 32 B algorithmic: no wasted re-reads
