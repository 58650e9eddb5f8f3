"""Host logic of the row-slab multi-GPU path, on CPU with two gloo ranks: the slab plan
tiles the grid in whole strips, and the rendezvous passes file descriptors between the
processes (the mechanism that carries the CUDA memory handles on a GPU box)."""
import ctypes
import importlib
import os
import sys
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ifl = importlib.import_module("incremental-fluids_b200")


def plan(L, h, world, rank):
    a, b = ctypes.c_int(), ctypes.c_int()
    rc = L.ifl_dist_plan(h, world, rank, ctypes.byref(a), ctypes.byref(b))
    return rc, a.value, b.value


def test_plan_tiles_the_grid_in_whole_strips():
    L = ifl.load_library()
    for h in (32, 64, 100, 128, 136, 1000, 4096, 16384):
        for world in range(1, 9):
            if (h + 63) // 64 < world:
                assert plan(L, h, world, 0)[0] != 0
                continue
            prev = 0
            for r in range(world):
                rc, a, b = plan(L, h, world, r)
                assert rc == 0 and a == prev and a % 64 == 0 and b > a
                prev = b
            assert prev == h
    rc, a, b = plan(L, 4096, 8, 3)
    assert (a, b) == (1536, 2048)


def _worker(rank, world, port, rdv, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = ifl.load_library()
    rc = L.ifl_dist_selftest(rank, world, rdv.encode())
    mine = torch.tensor(list(plan(L, 200, world, rank)) + [rc], dtype=torch.int64)
    alls = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(alls, mine)
    if rank == 0:
        q.put([t.tolist() for t in alls])
    dist.barrier()
    dist.destroy_process_group()


def test_two_gloo_ranks_rendezvous_and_plan():
    world = 2
    rdv = os.path.join(tempfile.mkdtemp(prefix="ifl_rdv_"), "sock")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, rdv, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # [plan rc, row0, row1, selftest rc] per rank: 200 rows = 4 strips of 64 rows -> 2 + 2
    assert res == [[0, 0, 128, 0], [0, 128, 200, 0]], res


def test_create_dist_without_gpu_fails_loudly():
    L = ifl.load_library()
    ctx = ctypes.c_void_p()
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    assert L.ifl_create_dist(ctypes.byref(ctx), 64, 64, 3, 0, 0, 2, b"/tmp/ifl_none") != 0
    assert b"no CUDA device" in L.ifl_last_error()
