"""Worker of tests/test_gpu_multi.py: one process per GPU (torchrun), rings sharded round-robin per row, the deposit grids
exchanged inside ptp_trap_step (NCCL all-reduce / fused peer-memory adds / peer-memory gather), solve replicated. Rank 0
additionally runs the whole load on its own GPU without a communicator and compares. One launch covers several deposit modes
and exchange kinds (process start-up and NCCL set-up dominate a launch).
    argv: <modes, comma separated: fp64|fixed> <rings> <steps> <exchanges, comma separated: nccl|peer|gather>"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from conftest import expected_density, rel_l2

    ptp = importlib.import_module("pic-trapped-plasma_b200")
    loaders = importlib.import_module("pic-trapped-plasma_b200.loaders")
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    modes = sys.argv[1].split(",")
    n_total, steps, dt = int(sys.argv[2]), int(sys.argv[3]), 2e-8 / 35
    exchanges = (sys.argv[4] if len(sys.argv) > 4 else "nccl").split(",")
    dens = expected_density()

    def run(trap, mode, r, z, v, cm):
        trap.set_deposit_mode(ptp.PTP_DEPOSIT_FIXED64 if mode == "fixed" else ptp.PTP_DEPOSIT_FP64)
        p = ptp.Plasma(trap, "Electrons", ptp.massE, -ptp.ePos)
        p.upload(r, z, v, cm)
        p.solvePoisson()
        trap.movePlasmas(dt, steps)
        trap.sync()
        return p

    probe = ptp.default_trap(device=local)
    hz, hr = probe.hz, probe.hr
    probe.close()
    r, z, cm, _ = loaders.place_rings(dens, 585, 128, hz, hr, n_total, rank, world)
    # speeds must not depend on the sharding: draw the full row-ordered sequence and take this rank's rings
    r_all, z_all, _, num_at_r = loaders.place_rings(dens, 585, 128, hz, hr, n_total)
    v_all = loaders.maxwellian_speeds(len(r_all), 150.0, ptp.massE, seed=7)
    offs = np.concatenate([[0], np.cumsum(num_at_r)])
    idx = np.concatenate([np.arange(offs[j] + rank, offs[j + 1], world) for j in range(128) if num_at_r[j] > 0])
    assert np.array_equal(z_all[idx], z)
    v = v_all[idx]
    results = []
    for mode in modes:
        single = None
        if rank == 0:
            st = ptp.default_trap(device=local)
            ps = run(st, mode, r_all, z_all, v_all, cm)
            single = (ps.rhs(), ps.selfPotential(), ps.getNumMacro())
            st.close()
        for exchange in exchanges:
            trap = ptp.default_trap(device=local)
            uid = [ptp.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            trap.comm_init(uid[0], world, rank)
            trap.set_allreduce({"nccl": 0, "peer": 1, "gather": 3}[exchange])
            p = run(trap, mode, r, z, v, cm)
            rhs, phi = p.rhs(), p.selfPotential()
            count = torch.tensor([p.getNumMacro()], device="cuda")
            dist.all_reduce(count)
            # every rank must hold the same grids after the exchange + replicated solve
            g = torch.from_numpy(np.stack([rhs, phi])).cuda()
            g0 = g.clone()
            dist.broadcast(g0, src=0)
            flag = torch.tensor([int(torch.equal(g.view(torch.int64), g0.view(torch.int64)))], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if rank == 0:
                results.append(dict(mode=mode, exchange=exchange, ranks=world, replicas_identical=bool(flag.item()),
                                    count_sharded=int(count.item()), count_single=int(single[2]),
                                    rhs_rel=rel_l2(rhs, single[0]), phi_rel=rel_l2(phi, single[1]),
                                    rhs_bitwise=bool(np.array_equal(rhs, single[0])), phi_bitwise=bool(np.array_equal(phi, single[1]))))
            trap.close()
            dist.barrier()
    if rank == 0:
        print("RESULT " + json.dumps(results))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
