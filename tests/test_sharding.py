"""CPU checks of the multi-GPU decomposition (world_size 2, gloo): the round-robin ring shards partition the
load exactly, and summing rank-local deposit grids with an all-reduce reproduces the single-rank grid -- the
one exchange step of the sharded PIC step. The deposit here is the CPU oracle's; the CUDA path is
checked against the same property on the GPU box (tests/test_gpu_multi.py)."""
import importlib
import os
import socket

import numpy as np
import pytest

from conftest import expected_density, rel_l2

loaders = importlib.import_module("pic-trapped-plasma_b200.loaders")

HZ, HR = 0.00011641025641025642, 0.00011625
NZ, NR = 585, 128


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_shards_partition_the_load():
    dens = expected_density()
    r_all, z_all, cm, num_at_r = loaders.place_rings(dens, NZ, NR, HZ, HR, 40000)
    assert len(r_all) == int(num_at_r.sum())
    for world in (2, 3, 8):
        parts = [loaders.place_rings(dens, NZ, NR, HZ, HR, 40000, rank, world) for rank in range(world)]
        assert all(p[2] == cm for p in parts)
        assert sum(len(p[0]) for p in parts) == len(r_all)
        z_cat = np.sort(np.concatenate([p[1] for p in parts]))
        assert np.array_equal(z_cat, np.sort(z_all))
        # every rank sees the same row occupancy shape to within one ring per row
        counts = np.array([np.bincount(p[0], minlength=NR) for p in parts])
        assert (counts.max(axis=0) - counts.min(axis=0)).max() <= 1


def test_placement_matches_reference_loader(c1_kat):
    """The vectorised placement reproduces the reference's loadDensityFile ring for ring (golden C1 load)."""
    dens = expected_density()
    ratio_file = np.array([float("%.6g" % (float("%.15g" % x) * 0.6)) for x in dens])
    r, z, cm, num_at_r = loaders.place_rings(ratio_file, NZ, NR, HZ, HR, 4000)
    assert list(num_at_r[:10]) == [777, 775, 751, 672, 518, 318, 141, 41, 7, 1]
    assert np.array_equal(r, c1_kat["e_r0"])
    assert cm == pytest.approx(float(c1_kat["e_chargeMacro"]), rel=1e-14)
    assert np.max(np.abs(z - c1_kat["e_z0"])) < 1e-15


def _worker(rank, world, port, out):
    import torch.distributed as dist
    import torch

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    dens = expected_density()
    r, z, cm, _ = loaders.place_rings(dens, NZ, NR, HZ, HR, 20000, rank, world)
    # rank-local deposit in units of one ring (what the CUDA kernel accumulates), then the exchange step
    k = np.floor(z / HZ).astype(np.int64)
    w = (z - k * HZ) / HZ
    grid = np.zeros((NZ + 1) * NR)
    np.add.at(grid, (NZ + 1) * r + k, 1 - w)
    np.add.at(grid, (NZ + 1) * r + k + 1, w)
    t = torch.from_numpy(grid)
    dist.all_reduce(t)
    # fixed-point flavour: int64 sums are independent of the reduction order
    fixed = np.zeros((NZ + 1) * NR, dtype=np.int64)
    wq = np.rint(w * 2.0 ** 40).astype(np.int64)
    np.add.at(fixed, (NZ + 1) * r + k, (1 << 40) - wq)
    np.add.at(fixed, (NZ + 1) * r + k + 1, wq)
    tf = torch.from_numpy(fixed)
    dist.all_reduce(tf)
    if rank == 0:
        np.savez(out, grid=t.numpy(), fixed=tf.numpy(), n=len(r))
    dist.destroy_process_group()


def test_allreduce_of_rank_local_deposits_equals_global_deposit(tmp_path):
    import torch.multiprocessing as mp

    out = str(tmp_path / "rank0.npz")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = np.load(out)
    dens = expected_density()
    r, z, cm, _ = loaders.place_rings(dens, NZ, NR, HZ, HR, 20000)
    k = np.floor(z / HZ).astype(np.int64)
    w = (z - k * HZ) / HZ
    grid = np.zeros((NZ + 1) * NR)
    np.add.at(grid, (NZ + 1) * r + k, 1 - w)
    np.add.at(grid, (NZ + 1) * r + k + 1, w)
    assert rel_l2(got["grid"], grid) < 1e-13
    assert abs(got["grid"].sum() - len(r)) < 1e-8
    fixed = np.zeros((NZ + 1) * NR, dtype=np.int64)
    wq = np.rint(w * 2.0 ** 40).astype(np.int64)
    np.add.at(fixed, (NZ + 1) * r + k, (1 << 40) - wq)
    np.add.at(fixed, (NZ + 1) * r + k + 1, wq)
    assert np.array_equal(got["fixed"], fixed)                     # bitwise, whatever the rank count
    assert int(got["fixed"].sum()) == len(r) << 40
