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
    # name: (species list [(name, mass key, share)], total macro-rings, description)
    "c4": ([("Antiprotons", "massP", 1.0)], 100_000_000, "single-species 100M macro-rings, default trap 585x128 (BASELINE configs[3])"),
    "c2": ([("Antiprotons", "massP", 1.0)], 1_000_000, "antiproton plasma, 1M macro-rings, default trap (BASELINE configs[1]; L2-resident)"),
    "c3": ([("Electrons", "massE", 0.5), ("Antiprotons", "massP", 0.5)], 10_000_000, "e- + pbar co-trapped, 10M macro-rings (BASELINE configs[2])"),
    "c5": ([("Antiprotons", "massP", 1.0)], 50_000_000, "fine-grid stress: 4096x1024 trap grid, 50M macro-rings (BASELINE configs[4])"),
}
GRIDS = {"c5": (4096, 1024)}          # (Nz, Nr); everything else runs on the reference's default 585 x 128
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


def build_load(ptp, loaders, workload, rank, n_ranks, hz, hr):
    species, total, _ = WORKLOADS[workload]
    Nz, Nr = grid_of(workload)
    dens = density_on(Nz, Nr)
    out = []
    for si, (name, mkey, share) in enumerate(species):
        mass = getattr(ptp, mkey)
        r, z, charge_macro, _ = loaders.place_rings(dens * share, Nz, Nr, hz, hr, int(total * share), rank, n_ranks)
        v = loaders.maxwellian_speeds(len(r), TEMPERATURE, mass, seed=1000 * si + rank)
        out.append((name, mass, r, z, v, charge_macro))
    return out


def cpu_reference_run(workload, steps, warmup, sample_rings):
    """The reference's own CPU implementation (oracle/_ref: its .cpp files compiled unmodified) on a bounded
    sample of the workload: every m-th ring of the same load, same trap, same dt; single thread (the
    reference has no threading)."""
    ptp = importlib.import_module("pic-trapped-plasma_b200")
    loaders = importlib.import_module("pic-trapped-plasma_b200.loaders")
    from oracle import port, ref
    kind = "reference" if ref.available() else "port"
    species, total, _ = WORKLOADS[workload]
    # bounded sample: keep the whole arm within a couple of minutes of CPU time (~2e7 ring-steps/s on one core)
    sample_rings = int(max(100_000, min(sample_rings, 1.2e9 / max(steps + warmup, 1))))
    stride = max(1, total // sample_rings)
    trap = make_trap(ref if kind == "reference" else port, workload)
    load = build_load(ptp, loaders, workload, 0, stride, trap.hz, trap.hr)   # ring i % stride == 0 of every row
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
                      "1 thread of %d cores (the reference is single-threaded); solve = direct banded LU stand-in for Eigen::SparseLU"
                      % (n, total, stride, steps, os.cpu_count()),
            "phases_s": {"moveRings": sec[0], "updateRHS": sec[1], "solve": sec[2], "whole": sec[3]},
            "push_deposit_particle_steps_per_s": (ring_steps / (sec[0] + sec[1])) if sec[0] + sec[1] > 0 else None}
    return info, sec[3] / steps * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--deposit", default="fp64", choices=["fp64", "fixed"])
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--window", type=int, default=0)
    ap.add_argument("--ctas", type=int, default=-1)
    ap.add_argument("--rings", type=int, default=0, help="rings per thread and tile (4 or 8)")
    ap.add_argument("--sort-interval", type=int, default=-1, help="> 0: re-sort every so many steps; 0: never; -1: adaptive (library default)")
    ap.add_argument("--allreduce", default="auto", choices=["auto", "nccl", "peer"], help="exchange step for N > 1 (auto: peer memory up to 2^20 grid nodes)")
    ap.add_argument("--cpu-sample", type=int, default=10_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--graph", action="store_true", help="replay steps as CUDA graphs (no per-phase times)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    species, total, desc = WORKLOADS[args.workload]
    config = {"workload": "%s: %s" % (args.workload, desc), "grid": "Nz=%d Nr=%d" % grid_of(args.workload), "dt_s": DT, "species": len(species),
              "rings_total": total, "deposit": args.deposit,
              "exchange": ("peer-memory push fused into the deposit flush + flag barrier" if args.allreduce == "peer" else "NCCL all-reduce") if world > 1 else "none (1 GPU)",
              "l2": "ring arrays %d MB per GPU vs 126 MB L2 (no flush needed)" % (total // world * 16 // 2**20) if total // world * 16 > 200e6
              else "ring arrays %d MB per GPU: L2-resident, NOT an HBM-bound measurement" % (total // world * 16 // 2**20)}

    if args.impl == "reference":
        if rank != 0:
            return
        info, ms = cpu_reference_run(args.workload, args.steps, max(args.warmup, 1), min(args.cpu_sample, total))
        line = {"impl": "reference", "metric": "particle-steps/s (push+deposit+solve)", "value": info["value"], "unit": "particle-steps/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": info, "gpu_launches": 0,
                "e2e": {"value": info["value"], "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist

    ptp = importlib.import_module("pic-trapped-plasma_b200")
    loaders = importlib.import_module("pic-trapped-plasma_b200.loaders")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    trap = make_trap(ptp, args.workload, device=local_rank)
    if world > 1:
        uid = [ptp.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        trap.comm_init(uid[0], world, rank)
        if args.allreduce == "auto":
            args.allreduce = "peer" if trap.G <= (1 << 20) else "nccl"
        trap.set_allreduce(1 if args.allreduce == "peer" else 0)
        config["exchange"] = "peer-memory push fused into the deposit flush + flag barrier" if args.allreduce == "peer" else "NCCL all-reduce"
    trap.set_deposit_mode(ptp.PTP_DEPOSIT_FIXED64 if args.deposit == "fixed" else ptp.PTP_DEPOSIT_FP64)
    if args.threads or args.window or args.ctas >= 0 or args.rings:
        trap.set_tuning(args.threads, args.window, args.ctas, args.rings)
    trap.set_sort_interval(args.sort_interval)
    trap.set_graph(args.graph)

    # The load "of the named shape": the reference's own loader (Plasma::loadDensityFile, Source/Plasma.cpp:558-622) run on
    # the device by ptp_plasma_load_density from the committed equilibrium density - its quantile placement and its
    # deviate stream, this rank keeping rings i = rank (mod world) of every row.
    Nz, Nr = grid_of(args.workload)
    dens = density_on(Nz, Nr)
    t_load = time.perf_counter()
    plasmas = []
    for name, mkey, share in species:
        p = ptp.Plasma(trap, name, getattr(ptp, mkey), -ptp.ePos)
        p.loadDensity(dens * share, TEMPERATURE, int(total * share), shard=rank, nShards=world, solve=False)
        plasmas.append(p)
    trap.sync()
    t_load = time.perf_counter() - t_load
    # pinned host staging of the same rings (the e2e leg uploads from / reads back to these)
    pinned = []
    if not args.no_e2e:
        for p in plasmas:
            r, z, v, _ = p.download()
            pinned.append((torch.from_numpy(r).pin_memory().numpy(), torch.from_numpy(z).pin_memory().numpy(),
                           torch.from_numpy(v).pin_memory().numpy(), p.chargeMacro))
            del r, z, v
    n_local = sum(p.getNumMacro() for p in plasmas)
    for p in plasmas:
        p.solvePoisson()

    # ---- device-resident leg: W warm-up steps, then exactly K timed steps ----------------------------
    trap.movePlasmas(DT, args.warmup)
    trap.sync()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    alive_before = sum_over_ranks(sum(p.getNumMacro() for p in plasmas))
    barrier()
    sorts_before = trap.sorts_done()
    trap.movePlasmas(DT, args.steps)
    trap.sync()
    barrier()
    times = trap.last_times()
    launches = trap.last_launches()
    sorts_timed = trap.sorts_done() - sorts_before
    clocks = sampler.stop()
    ms_total = max_over_ranks(float(times[0]))
    ms_push = max_over_ranks(float(times[1]))
    per_rank = None
    if world > 1:
        tt = torch.tensor([float(x) for x in times], dtype=torch.float64, device="cuda")
        gathered = [torch.zeros_like(tt) for _ in range(world)]
        dist.all_gather(gathered, tt)
        per_rank = [[round(float(x) / args.steps, 5) for x in g.tolist()] for g in gathered]
    alive_after = sum_over_ranks(sum(p.getNumMacro() for p in plasmas))
    ring_steps = 0.5 * (alive_before + alive_after) * args.steps
    value = ring_steps / (ms_total * 1e-3)

    # ---- roofline of the dominant kernel (K1: 32 B per ring-step) -------------------------------------
    peak, peak_kind = peaks()
    k1_ms = ms_push / args.steps
    achieved = 32.0 * n_local / (k1_ms * 1e-3) / 1e9 if k1_ms > 0 else None   # no per-phase times under --graph
    roofline = {"bound": "hbm", "kernel": "k_push_deposit", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "traffic": None, "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)",
                "bytes_per_unit": 32, "units_per_launch": n_local, "k1_ms_per_launch": k1_ms,
                "k1_share_of_step": ms_push / ms_total}
    if achieved is None:
        roofline["note"] = "graph replay: per-kernel events are not recorded; run without --graph for the roofline"
    prof = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get(args.workload)
        except Exception:
            pass

    # ---- end-to-end leg through the C ABI with host buffers ----------------------------------------------
    e2e = None
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

    cpu = None
    if args.workload == "c5":
        args.no_cpu_baseline = True       # the stand-in LU cannot factorise the 4.2 M-node grid in reasonable time
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _ = cpu_reference_run(args.workload, 3, 1, min(args.cpu_sample, total))

    if rank == 0:
        line = {"metric": "particle-steps/s (push+deposit+solve)", "value": value, "unit": "particle-steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "phases_ms_per_step": {"push_deposit": ms_push / args.steps, "allreduce": float(times[2]) / args.steps,
                                       "solve_node_field": float(times[3]) / args.steps},
                "phases_ms_per_step_per_rank[whole,push,exchange,solve]": per_rank,
                "load": {"how": "ptp_plasma_load_density (device-side Plasma::loadDensityFile placement + deviate stream)", "seconds_rank0": t_load},
                "tuning": {"threads": args.threads or 512, "window": args.window or 44, "ctas": args.ctas, "rings_per_thread": args.rings or 4,
                           "sort_interval": args.sort_interval, "sorts_in_run_rank0": sorts_timed}}
        print(json.dumps(line))
    trap.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
