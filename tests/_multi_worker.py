"""Worker of tests/test_gpu_multi.py: one process per GPU (torchrun), rings sharded round-robin per row,
NCCL all-reduce of the deposit grids inside ptp_trap_step, solve replicated. Rank 0 additionally runs the whole
load on its own GPU without a communicator and compares."""
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
    mode = ptp.PTP_DEPOSIT_FIXED64 if sys.argv[1] == "fixed" else ptp.PTP_DEPOSIT_FP64
    n_total, steps, dt = int(sys.argv[2]), int(sys.argv[3]), 2e-8 / 35
    exchange = sys.argv[4] if len(sys.argv) > 4 else "nccl"
    dens = expected_density()

    def run(trap, r, z, v, cm):
        trap.set_deposit_mode(mode)
        p = ptp.Plasma(trap, "Electrons", ptp.massE, -ptp.ePos)
        p.upload(r, z, v, cm)
        p.solvePoisson()
        trap.movePlasmas(dt, steps)
        trap.sync()
        return p

    trap = ptp.default_trap(device=local)
    uid = [ptp.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    trap.comm_init(uid[0], world, rank)
    trap.set_allreduce({"nccl": 0, "peer": 1, "gather": 3}[exchange])
    r, z, cm, _ = loaders.place_rings(dens, 585, 128, trap.hz, trap.hr, n_total, rank, world)
    # speeds must not depend on the sharding: draw the full row-ordered sequence and take this rank's rings
    r_all, z_all, _, num_at_r = loaders.place_rings(dens, 585, 128, trap.hz, trap.hr, n_total)
    v_all = loaders.maxwellian_speeds(len(r_all), 150.0, ptp.massE, seed=7)
    offs = np.concatenate([[0], np.cumsum(num_at_r)])
    idx = np.concatenate([np.arange(offs[j] + rank, offs[j + 1], world) for j in range(128) if num_at_r[j] > 0])
    assert np.array_equal(z_all[idx], z)
    p = run(trap, r, z, v_all[idx], cm)
    rhs, phi = p.rhs(), p.selfPotential()
    count = torch.tensor([p.getNumMacro()], device="cuda")
    dist.all_reduce(count)
    result = {"ok": True}
    # every rank must hold the same grids after the all-reduce + replicated solve
    g = torch.from_numpy(np.stack([rhs, phi])).cuda()
    g0 = g.clone()
    dist.broadcast(g0, src=0)
    same = bool(torch.equal(g, g0))
    flag = torch.tensor([int(same)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        single = ptp.default_trap(device=local)
        ps = run(single, r_all, z_all, v_all, cm)
        result.update(replicas_identical=bool(flag.item()), count_sharded=int(count.item()), count_single=ps.getNumMacro(),
                      rhs_rel=rel_l2(rhs, ps.rhs()), phi_rel=rel_l2(phi, ps.selfPotential()),
                      rhs_bitwise=bool(np.array_equal(rhs, ps.rhs())), phi_bitwise=bool(np.array_equal(phi, ps.selfPotential())),
                      ms=trap.last_times().tolist())
        single.close()
        print("RESULT " + json.dumps(result))
    trap.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
