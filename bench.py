#!/usr/bin/env python
"""Benchmark of the PIC step (PenningTrap::movePlasmas: push + deposit + all-reduce + solve).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c2|c3] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line on rank 0. Metric: particle-steps/s, whole job. Default workload "c4" =
BASELINE.json configs[3]: single-species 100 M macro-rings on the default (driver-A) trap, fp64,
sharded over the N GPUs (strong scaling, one rho all-reduce per step, solve replicated).
"""
import argparse
import importlib
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (species list [(name, mass key, share of the expected density, numMacro)], description)
    "c1": ([("Electrons", "massE", 0.6, 4000), ("Antiprotons", "massP", 0.4, 4000)],
           "the reference's own default: 4000 e- + 4000 pbar macro-rings (4001 + 4001 loaded), default trap 585x128 (BASELINE configs[0]; Diagnostics B:42-43)"),
    "c2": ([("Antiprotons", "massP", 1.0, 1_000_000)], "antiproton plasma, 1M macro-rings, default trap (BASELINE configs[1]; L2-resident)"),
    "c3": ([("Electrons", "massE", 0.5, 5_000_000), ("Antiprotons", "massP", 0.5, 5_000_000)], "e- + pbar co-trapped, 10M macro-rings (BASELINE configs[2])"),
    "c4": ([("Antiprotons", "massP", 1.0, 100_000_000)], "single-species 100M macro-rings, default trap 585x128 (BASELINE configs[3])"),
    "c5": ([("Antiprotons", "massP", 1.0, 50_000_000)], "fine-grid stress: 4096x1024 trap grid, 50M macro-rings (BASELINE configs[4])"),
    # not a BASELINE configuration: c5 with the electron mass (1.6 cells per step - rings that no re-sort can keep ordered; the
    # step hands them to the order-free form of the push kernel). Same as --workload c5 --electrons.
    "c5e": ([("Electrons", "massE", 1.0, 50_000_000)], "hot species: electrons on the 4096x1024 grid, 50M macro-rings (c5 with the electron mass; not a BASELINE configuration)"),
}
GRIDS = {"c5": (4096, 1024), "c5e": (4096, 1024)}          # (Nz, Nr); everything else runs on the reference's default 585 x 128
DT = 2e-8 / 35            # Diagnostics/C) Visualise Evolution.txt:34-36
TEMPERATURE = 150.0


def expected_density():
    d = np.load(os.path.join(ROOT, "tests", "golden", "expected_density_nonzero.npz"))
    dens = np.zeros(int(d["G"]))
    dens[d["index"]] = d["value"]
    return dens


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe). NVML through pynvml
    (a query takes ~1 ms, so even a 100 ms region gets tens of samples); nvidia-smi polling as a fallback."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.sm, self.mx, self.reasons, self.stop_flag, self.th, self.how = index, [], [], set(), False, None, "nvml"
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a list of ordinals
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                try:
                    phys = int(vis.split(",")[index])
                except Exception:
                    phys = index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
        except Exception:
            self.nv, self.how = None, "nvidia-smi"

    def _run(self):
        while not self.stop_flag:
            try:
                if self.nv:
                    nv = self.nv
                    self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                    self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                    try:
                        mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                    except Exception:
                        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                    for bit, name in self.REASONS.items():
                        if mask & bit:
                            self.reasons.add(name)
                    time.sleep(0.002)
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        r = [x.strip() for x in out.splitlines()[0].split(",")]
                        self.sm.append(float(r[0]))
                        self.mx.append(float(r[1]))
                        for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                            if val.lower().startswith("active"):
                                self.reasons.add(name)
                    time.sleep(0.02)
            except Exception:
                time.sleep(0.01)

    def start(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_flag = True
        if self.th:
            self.th.join(timeout=6)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "how": self.how}


def grid_of(workload):
    return GRIDS.get(workload, (585, 128))


def density_on(Nz, Nr):
    """The committed equilibrium density (585 x 128 grid); for other grids interpolated bilinearly in (r, z) over the
    same trap, which keeps the plasma's physical extent (the reference recomputes the equilibrium on the new grid)."""
    dens = expected_density()
    if (Nz, Nr) == (585, 128):
        return dens
    d = dens.reshape(128, 586)
    zi = np.linspace(0, 585, Nz + 1)
    ri = np.minimum(np.arange(Nr) * (128.0 / Nr), 127.0)
    z0 = np.minimum(zi.astype(int), 584)
    r0 = np.minimum(ri.astype(int), 126)
    fz, fr = zi - z0, ri - r0
    top = d[r0][:, z0] * (1 - fz) + d[r0][:, z0 + 1] * fz
    bot = d[r0 + 1][:, z0] * (1 - fz) + d[r0 + 1][:, z0 + 1] * fz
    return (top * (1 - fr)[:, None] + bot * fr[:, None]).reshape(-1)


def make_trap(mod, workload, **kw):
    Nz, Nr = grid_of(workload)
    if hasattr(mod, "PenningTrap"):
        el = [mod.Electrode(0.01322, v) for v in (0, -70, -15, -70, 0)]
        return mod.PenningTrap(0.01488, el, [0.0005] * 4, Nz, Nr, **kw)
    return mod.default_trap(Nz, Nr)


def species_of(args):
    species, desc = WORKLOADS[args.workload]
    if args.total:
        scale = args.total / float(sum(n for _, _, _, n in species))
        species = [(nm, mk, sh, max(1, int(round(n * scale)))) for nm, mk, sh, n in species]
        desc += " [--total %d]" % args.total
    if args.electrons:
        species = [("Electrons", "massE", sh, n) for _, _, sh, n in species]
        desc += " [--electrons]"
    return species, desc


def host_load(ptp, loaders, species, workload, stride, hz, hr):
    """Host-side placement of every stride-th ring of each row (the reference loader's quantile positions, vectorised in
    loaders.place_rings) + NumPy Maxwellian speeds: the load of the stand-alone reference arm, which must not touch the GPU."""
    Nz, Nr = grid_of(workload)
    dens = density_on(Nz, Nr)
    out = []
    for si, (name, mkey, share, num) in enumerate(species):
        mass = getattr(ptp, mkey)
        r, z, charge_macro, _ = loaders.place_rings(dens * share, Nz, Nr, hz, hr, num, 0, stride)
        v = loaders.maxwellian_speeds(len(r), TEMPERATURE, mass, seed=1000 * si)
        out.append((name, mass, r, z, v, charge_macro))
    return out


def sample_of_device_load(ptp, species, device_rings, stride):
    """Every stride-th ring of each row of the rings the device loader produced (downloaded): the CPU leg then runs on the
    very rings the GPU leg runs on."""
    out = []
    for (name, mkey, share, num), (r, z, v, cm) in zip(species, device_rings):
        starts = np.flatnonzero(np.r_[True, r[1:] != r[:-1]])
        within = np.arange(len(r)) - np.repeat(starts, np.diff(np.r_[starts, len(r)]))
        keep = within % stride == 0
        out.append((name, getattr(ptp, mkey), np.ascontiguousarray(r[keep]), np.ascontiguousarray(z[keep]), np.ascontiguousarray(v[keep]), cm))
    return out


def cpu_reference_run(args, species, steps, warmup, sample_rings, device_rings=None):
    """The reference's own CPU implementation (oracle/_ref: its .cpp files compiled unmodified) on a bounded
    sample of the workload: every m-th ring of each row of the same load, same trap, same dt; single thread (the
    reference has no threading)."""
    ptp = importlib.import_module("pic-trapped-plasma_b200")
    loaders = importlib.import_module("pic-trapped-plasma_b200.loaders")
    from oracle import port, ref
    kind = "reference" if ref.available() else "port"
    workload = args.workload
    total = sum(n for _, _, _, n in species)
    # bounded sample: keep the whole arm within a couple of minutes of CPU time (~2e7 ring-steps/s on one core)
    sample_rings = int(max(8_000, min(sample_rings, 1.2e9 / max(steps + warmup, 1))))
    stride = max(1, total // sample_rings)
    trap = make_trap(ref if kind == "reference" else port, workload)
    if device_rings is not None:
        load = sample_of_device_load(ptp, species, device_rings, stride)
        how = "rings downloaded from the device loader (the GPU leg's own load)"
    else:
        load = host_load(ptp, loaders, species, workload, stride, trap.hz, trap.hr)
        how = "host placement at the reference loader's quantile positions + NumPy Maxwellian speeds (same shape as the GPU leg's load, not the same deviates)"
    n = 0
    plasmas = []
    for name, mass, r, z, v, cm in load:
        p = trap.plasma(name, mass, -ptp.ePos)
        p.set_rings(r, z, v, cm * stride)
        p.solve_poisson()
        plasmas.append(p)
        n += len(r)
    if kind == "reference":
        trap.timed_steps(DT, warmup)
        ring_steps, sec = trap.timed_steps(DT, steps)
    else:
        trap.move_plasmas(DT, warmup)
        t0 = time.perf_counter()
        ring_steps = 0
        for _ in range(steps):
            ring_steps += sum(p.count() for p in plasmas)
            trap.move_plasmas(DT, 1)
        sec = [0, 0, 0, time.perf_counter() - t0]
    trap.close()
    value = ring_steps / sec[3]
    info = {"value": value, "unit": "particle-steps/s", "cores": 1, "kind": kind,
            "sample": "%d of %d rings (every %d-th ring of each row), %d steps of movePlasmas on the same trap; "
                      "1 thread of %d cores (the reference is single-threaded); solve = direct banded LU stand-in for Eigen::SparseLU; load: %s"
                      % (n, total, stride, steps, os.cpu_count(), how),
            "phases_s": {"moveRings": sec[0], "updateRHS": sec[1], "solve": sec[2], "whole": sec[3]},
            "push_deposit_particle_steps_per_s": (ring_steps / (sec[0] + sec[1])) if sec[0] + sec[1] > 0 else None}
    return info, sec[3] / steps * 1e3


def push_source_sha():
    import hashlib
    with open(os.path.join(ROOT, "pic-trapped-plasma_b200", "csrc", "ptp_push.cu"), "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()[:16]


K1_PROFILED = r"k_push_depositILi512ELi4ELb1ELb0ELb0EL[bi]0EE"   # k_push_deposit<512, 4, PUSH, fp64, FAST, thread-private bins>


def k1_sass_sha(path=None):
    """Hash of the machine code of the profiled instantiation of K1 inside the shipped library (cuobjdump -sass; addresses,
    encodings and the link-time addresses of the static shared variables masked): ties profiles/k1_traffic.json to the
    kernel that runs even when other parts of ptp_push.cu have changed since the capture. None without cuobjdump."""
    import hashlib
    import re
    import shutil
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    path = path or os.path.join(ROOT, "pic-trapped-plasma_b200", "libptp_b200.so")
    try:
        out = subprocess.run([exe, "-sass", path], capture_output=True, text=True, timeout=120).stdout
    except Exception:
        return None
    cur, body = None, []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur and re.search(K1_PROFILED, cur):
            if re.match(r"^\s*/\* 0x[0-9a-f]+ \*/\s*$", line):
                continue
            line = re.sub(r"/\*[0-9a-f]{4,}\*/", "", line)
            line = re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).strip()
            line = re.sub(r"^(U?MOV U?R\d+), 0x[0-9a-f]+ ;", r"\1, IMM ;", line)
            if line:
                body.append(line)
    return hashlib.sha256("\n".join(body).encode()).hexdigest()[:16] if body else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--deposit", default="fp64", choices=["fp64", "fixed"])
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--window", type=int, default=0)
    ap.add_argument("--ctas", type=int, default=-1)
    ap.add_argument("--rings", type=int, default=0, help="rings per thread and tile (4 or 8)")
    ap.add_argument("--sort-interval", type=int, default=-1, help="> 0: re-sort every so many steps; 0: never; -1: adaptive (library default)")
    ap.add_argument("--allreduce", default="auto", choices=["auto", "nccl", "peer", "gather"],
                    help="exchange step for N > 1 (auto: the library's choice - remote adds fused into the deposit flush up to 2^20 grid nodes, peer-memory gather above)")
    ap.add_argument("--cpu-sample", type=int, default=10_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="skip the multi-GPU parity pass (N > 1)")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"], help="replay steps as CUDA graphs (auto: the library's default policy)")
    ap.add_argument("--min-time", type=float, default=0.5, help="repeat the K-step batch until this many seconds are covered; the median batch is reported")
    ap.add_argument("--max-batches", type=int, default=400)
    ap.add_argument("--total", type=int, default=0, help="override the workload's ring count (experiments: shard-sized loads on one GPU)")
    ap.add_argument("--hot", default="auto", choices=["auto", "on", "off"],
                    help="per-warp-bin form of the push kernel for species that cannot be kept ordered by cell (auto: the re-sort policy decides)")
    ap.add_argument("--electrons", action="store_true", help="experiments: run the workload's species with the electron mass (fast rings: 0.23 cells/step on the default grid)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    species, desc = species_of(args)
    total = sum(n for _, _, _, n in species)
    config = {"workload": "%s: %s" % (args.workload, desc), "grid": "Nz=%d Nr=%d" % grid_of(args.workload), "dt_s": DT, "species": len(species),
              "rings_total": total, "deposit": args.deposit,
              "exchange": args.allreduce if world > 1 else "none (1 GPU)",
              "l2": "ring arrays %d MB per GPU vs 126 MB L2 (no flush needed)" % (total // world * 16 // 2**20) if total // world * 16 > 200e6
              else "ring arrays %d MB per GPU: L2-resident, NOT an HBM-bound measurement" % (total // world * 16 // 2**20)}

    if args.impl == "reference":
        if rank != 0:
            return
        info, ms = cpu_reference_run(args, species, args.steps, max(args.warmup, 1), min(args.cpu_sample, total))
        config_ref = dict(config)
        line = {"impl": "reference", "metric": "particle-steps/s (push+deposit+solve)", "value": info["value"], "unit": "particle-steps/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_ref,
                "cpu_baseline": info, "gpu_launches": 0,
                "e2e": {"value": info["value"], "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist

    ptp = importlib.import_module("pic-trapped-plasma_b200")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_over_ranks(x, op):
        if world == 1:
            return x
        t = torch.tensor(x, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return t.tolist() if t.dim() else float(t.item())

    def max_over_ranks(x):
        return reduce_over_ranks(x, dist.ReduceOp.MAX if world > 1 else None)

    def sum_over_ranks(x):
        return reduce_over_ranks(x, dist.ReduceOp.SUM if world > 1 else None)

    trap = make_trap(ptp, args.workload, device=local_rank)
    if world > 1:
        uid = [ptp.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        trap.comm_init(uid[0], world, rank)
        if args.allreduce == "auto":
            args.allreduce = "peer" if trap.G <= (1 << 20) else "gather"      # the library's own choice (ptp_trap_set_allreduce(t, 2))
        trap.set_allreduce({"nccl": 0, "peer": 1, "gather": 3}[args.allreduce])
        config["exchange"] = {"peer": "peer-memory adds fused into the deposit flush (system-scope atomics over NVLink) + flag barrier",
                              "gather": "peer-memory gather: each rank stores its populated rows into every rank's gather area over NVLink, flags, local sum in rank order (one kernel)",
                              "nccl": "NCCL all-reduce"}[args.allreduce]
    trap.set_deposit_mode(ptp.PTP_DEPOSIT_FIXED64 if args.deposit == "fixed" else ptp.PTP_DEPOSIT_FP64)
    if args.threads or args.window or args.ctas >= 0 or args.rings:
        trap.set_tuning(args.threads, args.window, args.ctas, args.rings)
    trap.set_sort_interval(args.sort_interval)
    if args.graph != "auto":
        trap.set_graph(args.graph == "on")

    # The load "of the named shape": the reference's own loader (Plasma::loadDensityFile, Source/Plasma.cpp:558-622) run on
    # the device by ptp_plasma_load_density from the committed equilibrium density - its quantile placement and its
    # deviate stream, this rank keeping rings i = rank (mod world) of every row.
    Nz, Nr = grid_of(args.workload)
    dens = density_on(Nz, Nr)
    dens_pinned = torch.from_numpy(dens).pin_memory().numpy()
    t_load = time.perf_counter()
    plasmas = []
    for name, mkey, share, num in species:
        p = ptp.Plasma(trap, name, getattr(ptp, mkey), -ptp.ePos)
        if args.hot != "auto":
            p.set_hot(1 if args.hot == "on" else 0)
        p.loadDensity(dens * share, TEMPERATURE, num, shard=rank, nShards=world, solve=False)
        plasmas.append(p)
    trap.sync()
    t_load = time.perf_counter() - t_load
    # pinned host staging of the same rings (the e2e leg uploads from / reads back to these; the CPU leg samples them)
    pinned = []
    if not args.no_e2e or not args.no_cpu_baseline:
        for p in plasmas:
            r, z, v, _ = p.download()
            pinned.append((torch.from_numpy(r).pin_memory().numpy(), torch.from_numpy(z).pin_memory().numpy(),
                           torch.from_numpy(v).pin_memory().numpy(), p.chargeMacro))
            del r, z, v
    n_local = sum(p.getNumMacro() for p in plasmas)
    for p in plasmas:
        p.solvePoisson()

    # ---- device-resident leg: W warm-up steps, then batches of exactly K timed steps ---------------------------
    # One batch = K steps between a barrier + synchronize on both sides, timed on the device (CUDA events on the trap's
    # stream), max over ranks. A batch of the default K is a few ms, too short for the clock sampler and for NVML to see
    # load, so the batch is repeated until --min-time is covered and the MEDIAN batch is reported.
    trap.movePlasmas(DT, args.warmup)
    trap.sync()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    sorts_before = trap.sorts_done()
    batches = []          # per batch: [ms_total, ms_push, ms_exchange, ms_solve, alive_before (global), launches]
    n_batches = 1
    wall0 = time.perf_counter()
    b = 0
    while b < n_batches:
        alive_before = sum(p.getNumMacro() for p in plasmas)
        barrier()
        trap.movePlasmas(DT, args.steps)
        trap.sync()
        barrier()
        tm = [float(x) for x in trap.last_times()]
        batches.append(tm + [float(alive_before), float(trap.last_launches())])
        if b == 0:
            first = max_over_ranks(tm[0])
            n_batches = int(min(args.max_batches, max(1, math.ceil(args.min_time * 1e3 / max(first, 1e-3)))))
        b += 1
    wall = time.perf_counter() - wall0
    alive_after = sum_over_ranks(float(sum(p.getNumMacro() for p in plasmas)))
    clocks = sampler.stop()
    sorts_timed = trap.sorts_done() - sorts_before
    hot_after = [bool(p.is_hot()) for p in plasmas]
    local = np.array(batches)
    mx = np.array(max_over_ranks(local[:, :4].tolist())).reshape(-1, 4) if world > 1 else local[:, :4]
    alive = np.array(sum_over_ranks(local[:, 4].tolist())) if world > 1 else local[:, 4]
    order = np.argsort(mx[:, 0])
    med = int(order[len(order) // 2])                      # the median batch (by whole-batch time)
    ms_total, ms_push = float(mx[med, 0]), float(mx[med, 1])
    alive_next = float(alive[med + 1]) if med + 1 < len(alive) else alive_after
    ring_steps = 0.5 * (float(alive[med]) + alive_next) * args.steps
    value = ring_steps / (ms_total * 1e-3)
    launches = int(local[med, 5])
    timing = {"batches": len(batches), "steps_per_batch": args.steps, "reported": "median batch", "timed_wall_s": wall,
              "batch_ms": {"min": float(mx[:, 0].min()), "median": ms_total, "max": float(mx[:, 0].max())}}
    # The timed steps carry no per-phase events (events between the kernels of a step cost the programmatic-launch overlap, and
    # graph replay has none): one more un-timed pass of K stream-launched steps WITH the phase events gives the kernel times
    # for the roofline and the phase table (the headline value stays the one measured above).
    phases = [float(mx[med, 1]), float(mx[med, 2]), float(mx[med, 3])]
    timing["graph_replay"] = bool(args.graph != "off" and not (world > 1 and args.allreduce == "nccl") and (args.graph == "on" or n_local <= 8_000_000 or world > 1))   # (the library's policy, ptp_api.cu want_graph)
    if ms_push == 0.0:
        trap.set_graph(False)
        trap.set_phase_events(True)
        barrier()
        trap.movePlasmas(DT, args.steps)
        trap.sync()
        barrier()
        tm = [float(x) for x in trap.last_times()]
        probe = max_over_ranks(tm)
        ms_push, phases = float(probe[1]), [float(probe[1]), float(probe[2]), float(probe[3])]
        timing["phase_probe"] = "separate un-timed pass of %d stream-launched steps (whole pass %.4f ms/step)" % (args.steps, float(probe[0]) / args.steps)
        trap.set_graph(None if args.graph == "auto" else args.graph == "on")
        trap.set_phase_events(False)
        local_ph = [float(local[med, 0])] + tm[1:4]
    else:
        local_ph = local[med, :4].tolist()
    per_rank = None
    if world > 1:
        tt = torch.tensor(local_ph, dtype=torch.float64, device="cuda")
        gathered = [torch.zeros_like(tt) for _ in range(world)]
        dist.all_gather(gathered, tt)
        per_rank = [[round(float(x) / args.steps, 5) for x in g.tolist()] for g in gathered]

    # ---- roofline of the dominant kernel (K1: 32 B per ring-step) -------------------------------------
    peak, peak_kind = peaks()
    k1_ms = ms_push / args.steps / len(species)            # one launch per species and step
    units = n_local / len(species)
    achieved = 32.0 * units / (k1_ms * 1e-3) / 1e9 if k1_ms > 0 else None   # no per-phase times under graph replay
    roofline = {"bound": "hbm", "kernel": "k_push_deposit", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "traffic": None, "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)",
                "bytes_per_unit": 32, "units_per_launch": units, "k1_ms_per_launch": k1_ms,
                "k1_share_of_step": min(1.0, ms_push / ms_total)}
    if achieved is None:
        roofline["note"] = "graph replay: per-kernel events are not recorded; run with --graph off for the roofline"
    prof = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(prof):
        # DRAM bytes per ring of ONE launch of this kernel from an `ncu --set full` capture (dram__bytes_read.sum + write.sum),
        # scaled to this launch's rings; only trusted while the kernel source is the one that was profiled
        try:
            tr = json.load(open(prof))
            entry = tr.get(args.workload)
            # same source file as profiled, or a later version of the file whose K1 machine code was found identical to the profiled
            # one (listed in the profile; tests/test_bench_contract.py re-checks the shipped library with cuobjdump), or - for a
            # version not listed - the same machine code as seen right now
            src = push_source_sha()
            same = tr.get("push_cu_sha16") == src or src in tr.get("same_k1_sass_sources", []) or \
                (rank == 0 and tr.get("k1_sass_sha16") and tr.get("k1_sass_sha16") == k1_sass_sha())
            if entry and same:
                roofline["traffic"] = entry["dram_bytes_per_ring"] * units
                roofline["traffic_source"] = "%s: %s, %.2f B/ring x rings of this launch" % (tr.get("report"), entry.get("kernel"), entry["dram_bytes_per_ring"])
            elif entry:
                roofline["traffic_source"] = "profiles/k1_traffic.json is from another version of the push kernel (source and machine code differ): not reported"
            if any(hot_after):
                # the capture is of the default (thread-private) form; a species in the order-free form runs other instantiations
                roofline["traffic"] = None
                roofline["kernel"] = "k_push_deposit (order-free form: per-warp bins, warp sort)"
                roofline["traffic_source"] = "no ncu capture of the shipped order-free form of K1 (profiles/r02_ncu_hot_form.txt holds its first version: 30.9 B/ring)"
        except Exception:
            pass

    # ---- multi-GPU parity pass (N > 1): the sharded fixed-point step against the same load on one GPU, bitwise -----
    parity = None
    if world > 1 and not args.no_verify:
        parity = verify_sharded(ptp, trap, plasmas, species, dens, rank, world, local_rank, dist, torch)
        # back to the bench load for the end-to-end leg
        trap.set_deposit_mode(ptp.PTP_DEPOSIT_FIXED64 if args.deposit == "fixed" else ptp.PTP_DEPOSIT_FP64)
        for p, (name, mkey, share, num) in zip(plasmas, species):
            p.loadDensity(dens * share, TEMPERATURE, num, shard=rank, nShards=world, solve=True)

    # ---- end-to-end leg through the C ABI with host buffers ----------------------------------------------
    e2e = None
    e2e_density = None
    if not args.no_e2e:
        barrier()
        t0 = time.perf_counter()
        h2d = 0
        for p, (pr, pz, pv, cm) in zip(plasmas, pinned):
            p.upload(pr, pz, pv, cm)                       # H2D from pinned host memory
            h2d += pr.nbytes + pz.nbytes + pv.nbytes
        t1 = time.perf_counter()
        for p in plasmas:
            p.solvePoisson()
        trap.sync()
        t2 = time.perf_counter()
        d2h = 0
        e2e_steps = args.steps
        for _ in range(e2e_steps):
            trap.movePlasmas(DT, 1)
            counts = [p.getNumMacro() for p in plasmas]    # D2H every step: the step's metric (alive rings per species)
            d2h += 8 * len(plasmas)
        t3 = time.perf_counter()
        rhs = plasmas[0].rhs()                             # D2H: density grid (the diagnostics' input)
        d2h += rhs.nbytes
        trap.sync()
        barrier()
        t4 = time.perf_counter()
        sec = max_over_ranks(t4 - t0)
        e2e_phases = {"upload_h2d": t1 - t0, "first_deposit_solve": t2 - t1, "steps_and_counts": t3 - t2, "readback_grid": t4 - t3}
        e2e = {"value": sum_over_ranks(float(sum(counts))) * e2e_steps / sec, "unit": "particle-steps/s",
               "h2d_bytes_per_step": h2d / e2e_steps, "d2h_bytes_per_step": d2h / e2e_steps,
               "protocol": "upload rings from pinned host (H2D, %d B/ring) + first deposit/solve + %d x (movePlasmas + read back of the alive counts) + read back of the density grid; "
                           "rings stay resident between steps as in the reference's API (movePlasmas(dt) takes no ring data)" % (20, e2e_steps),
               "phases_s_rank0": e2e_phases}
        # the same run started the way the reference's drivers start it: Plasma::loadDensityFile from the expected-density grid
        # (H2D of the grid only; placement + speeds on the device), beside the ring-upload variant above
        barrier()
        t0 = time.perf_counter()
        for p, (name, mkey, share, num) in zip(plasmas, species):
            p.loadDensity(dens_pinned * share, TEMPERATURE, num, shard=rank, nShards=world, solve=True)
        for _ in range(e2e_steps):
            trap.movePlasmas(DT, 1)
            counts = [p.getNumMacro() for p in plasmas]
        rhs = plasmas[0].rhs()
        trap.sync()
        barrier()
        sec = max_over_ranks(time.perf_counter() - t0)
        e2e_density = {"value": sum_over_ranks(float(sum(counts))) * e2e_steps / sec, "unit": "particle-steps/s",
                       "h2d_bytes_per_step": dens.nbytes * len(species) / e2e_steps, "d2h_bytes_per_step": d2h / e2e_steps,
                       "protocol": "loadDensityFile from the expected-density grid in pinned host memory (H2D of the grid; placement + speeds on the device) + first deposit/solve + "
                                   "%d x (movePlasmas + read back of the alive counts) + read back of the density grid" % e2e_steps}

    cpu = None
    if args.workload in GRIDS:
        args.no_cpu_baseline = True       # the stand-in LU cannot factorise the 4.2 M-node grid in reasonable time
    if not args.no_cpu_baseline:
        # rank 0's host cores, beside every N; at N > 1 the sample is taken from rank 0's shard (rings i = 0 mod N of every row)
        if rank == 0:
            sample = min(args.cpu_sample, total)
            if pinned:
                # rank 0's shard holds rings i = 0 (mod N) of every row, each standing for N rings of the full load
                dev = [(r, z, v, cm * world) for (r, z, v, cm) in pinned]
                shard_species = [(nm, mk, sh, max(1, n // world)) for nm, mk, sh, n in species]
                cpu, _ = cpu_reference_run(args, shard_species, 3, 1, sample, device_rings=dev)
            else:
                cpu, _ = cpu_reference_run(args, species, 3, 1, sample)
        barrier()

    if rank == 0:
        line = {"metric": "particle-steps/s (push+deposit+solve)", "value": value, "unit": "particle-steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
                "timing": timing, "parity": parity, "e2e_from_density": e2e_density,
                "phases_ms_per_step": {"push_deposit": phases[0] / args.steps, "allreduce": phases[1] / args.steps,
                                       "solve_node_field": phases[2] / args.steps},
                "phases_ms_per_step_per_rank[whole,push,exchange,solve]": per_rank,
                "load": {"how": "ptp_plasma_load_density (device-side Plasma::loadDensityFile placement + deviate stream)", "seconds_rank0": t_load},
                "tuning": {"threads": args.threads or 512, "window": args.window or 44, "ctas": args.ctas, "rings_per_thread": args.rings or 4,
                           "sort_interval": args.sort_interval, "sorts_in_run_rank0": sorts_timed, "graph": args.graph,
                           "hot": args.hot, "hot_form_in_use_rank0": hot_after}}
        print(json.dumps(line))
    trap.close()
    if world > 1:
        dist.destroy_process_group()


def verify_sharded(ptp, trap, plasmas, species, dens, rank, world, local_rank, dist, torch, rings=2_000_000, steps=5):
    """Reload the bench's species with `rings` rings in fixed-point deposit mode, sharded as in the timed run, step, and
    compare rank 0's grids BITWISE with the same load stepped on one GPU without a communicator (integer deposit sums do not
    depend on the summation order, so any difference is a lost or duplicated contribution of the exchange); all ranks must
    hold identical grids and the alive counts must add up."""
    total = float(sum(n for _, _, _, n in species))
    nums = [max(1, int(round(rings * n / total))) for _, _, _, n in species]
    trap.set_deposit_mode(ptp.PTP_DEPOSIT_FIXED64)
    for p, (name, mkey, share, _), n in zip(plasmas, species, nums):
        p.loadDensity(dens * share, TEMPERATURE, n, shard=rank, nShards=world, solve=True)
    trap.movePlasmas(DT, steps)
    trap.sync()
    grids = np.stack([g for p in plasmas for g in (p.rhs(), p.selfPotential())])
    count = torch.tensor([float(sum(p.getNumMacro() for p in plasmas))], dtype=torch.float64, device="cuda")
    dist.all_reduce(count)
    g = torch.from_numpy(grids).cuda()
    g0 = g.clone()
    dist.broadcast(g0, src=0)
    same = torch.tensor([int(torch.equal(g.view(torch.int64), g0.view(torch.int64)))], device="cuda")
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    out = None
    if rank == 0:
        Nz, Nr = trap.Nz, trap.Nr
        el = [ptp.Electrode(e.getLength(), e.getPotential()) for e in trap.electrodes]
        single = ptp.PenningTrap(trap.trapRadius, el, trap.gaps, Nz, Nr, device=local_rank)
        single.set_deposit_mode(ptp.PTP_DEPOSIT_FIXED64)
        sp = []
        for (name, mkey, share, _), n in zip(species, nums):
            q = ptp.Plasma(single, name, getattr(ptp, mkey), -ptp.ePos)
            q.loadDensity(dens * share, TEMPERATURE, n, solve=True)
            sp.append(q)
        single.movePlasmas(DT, steps)
        single.sync()
        ref_grids = np.stack([gr for q in sp for gr in (q.rhs(), q.selfPotential())])
        n_single = sum(q.getNumMacro() for q in sp)
        rho_ok = all(np.array_equal(grids[2 * i].view(np.int64), ref_grids[2 * i].view(np.int64)) for i in range(len(sp)))
        phi_ok = all(np.array_equal(grids[2 * i + 1].view(np.int64), ref_grids[2 * i + 1].view(np.int64)) for i in range(len(sp)))
        out = {"what": "sharded fixed-point step vs the same load on one GPU", "rings": int(sum(nums)), "steps": steps, "ranks": world,
               "rho_bitwise": bool(rho_ok), "phi_bitwise": bool(phi_ok), "replicas_identical": bool(same.item()),
               "count_sharded": int(count.item()), "count_single": int(n_single),
               "ok": bool(rho_ok and phi_ok and same.item() and int(count.item()) == int(n_single))}
        single.close()
    dist.barrier()
    return out


if __name__ == "__main__":
    main()
