"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: partition plans, slab broadcast protocol, framebuffer
reduce / row gather.  The GPU kernels are not involved; the same functions drive NCCL in bench.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gi_raytracer_b200 import dist as gd


def test_partition_plans_cover_everything_once():
    for world in (1, 2, 4, 8):
        r = gd.sample_ranges(8, world)
        assert r[0][0] == 0 and all(r[i][1] == r[i + 1][0] for i in range(world - 1)) and r[-1][1] == 8 * world
        for h in (1, 17, 1080, 2160):
            blocks = gd.row_blocks(h, world)
            rows = sorted(y for b in blocks for y0, y1 in b for y in range(y0, y1))
            assert rows == list(range(h))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dev = torch.device("cpu")
        # slab broadcast: root owns 1000 bytes, the others reserve + adopt
        state = {}
        root_slab = torch.arange(1000, dtype=torch.int64).to(torch.uint8)

        def make():
            return root_slab, root_slab.numel()

        def reserve(n):
            state["buf"] = torch.zeros(n, dtype=torch.uint8)
            return state["buf"]

        def adopt(n):
            state["adopted"] = n

        n = gd.broadcast_slab(make, reserve, adopt, rank, 0, dev)
        ok = n == 1000 and (rank == 0 or (state["adopted"] == 1000 and torch.equal(state["buf"], root_slab)))
        # sample split: reduce of partial sums
        acc = torch.full((6, 3), float(rank + 1), dtype=torch.float64)
        gd.reduce_accum(acc, 0)
        if rank == 0:
            ok = ok and bool((acc == sum(range(1, world + 1))).all())
        # tile split: row gather
        h, w = 37, 5
        blocks = gd.row_blocks(h, world, block=4)
        frame_ref = torch.arange(h * w * 3, dtype=torch.float64).reshape(h, w, 3)
        local = torch.cat([frame_ref[y0:y1] for y0, y1 in blocks[rank]]) if blocks[rank] else torch.zeros((0, w, 3), dtype=torch.float64)
        frame = gd.gather_rows(local, blocks, h, w)
        ok = ok and torch.equal(frame, frame_ref)
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_world2_gloo_protocols():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out.get(r) for r in range(world)), dict(out)
