"""Two GPUs, one process each, nothing but the C ABI between them (gi_comm_init / gi_photon_map_bcast / gi_framebuffer_gather /
gi_framebuffer_reduce — NCCL from C; the 128-byte communicator id travels through a file).  Checks the claims of SURVEY 8e:
  * the photon map built on rank 0 and broadcast as one slab answers gather queries on rank 1 exactly like on rank 0;
  * tile split: the ranks' interleaved row blocks, gathered on rank 0, equal the one-GPU frame BIT FOR BIT;
  * sample split: the ranks' partial sums, reduced onto rank 0, equal the one-GPU frame to 1e-13 (fp64 reassociation only).
Skipped on boxes with fewer than two GPUs (run it with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`)."""
import os
import time

import numpy as np
import pytest

from conftest import ROOT, have_assets, scene_path

pytestmark = pytest.mark.gpu

W = H = 200          # 13 blocks of 16 rows: an odd count, so the ranks own different numbers of rows and the last block is short
SPP, DEPTH, PHOTONS, BLOCK = 4, 6, 60000, 16


def _worker(rank, world, tmp, out):
    import torch
    from gi_raytracer_b200 import capi, host
    from gi_raytracer_b200.abi import render_params
    torch.cuda.set_device(rank)
    ctx = capi.Context(rank)
    try:
        idf = os.path.join(tmp, "nccl_id")
        if rank == 0:
            with open(idf + ".tmp", "wb") as f:
                f.write(capi.comm_unique_id())
            os.rename(idf + ".tmp", idf)
        else:
            t0 = time.time()
            while not os.path.exists(idf):
                if time.time() - t0 > 60:
                    raise TimeoutError("no communicator id from rank 0")
                time.sleep(0.01)
        ctx.comm_init(open(idf, "rb").read(), rank, world)
        info = ctx.comm_info()
        assert info["rank"] == rank and info["nranks"] == world and info["nccl_version"] > 20000
        sc = host.load_scene(scene_path("caustics"))
        ctx.upload_scene(sc)
        if rank == 0:
            ctx.photon_trace(PHOTONS, 5, seed=3)
            ctx.photon_map_build(None)
        ctx.photon_map_bcast(0)
        ctx.synchronize()
        pm = ctx.photon_map_info()
        # the adopted map answers like the original: same queries on both ranks, results compared on rank 0 through files
        rng = np.random.RandomState(5)
        lo, hi = sc.root_box[:3], sc.root_box[3:]
        q = lo + rng.rand(4000, 3) * (hi - lo)
        d = rng.randn(4000, 3)
        rgb, knn, nc = ctx.gather(q, d, 32)
        np.save(os.path.join(tmp, f"gather_{rank}.npy"), np.concatenate([rgb.ravel(), knn.ravel().astype(np.float64), nc.astype(np.float64)]))
        P = render_params(W, H, SPP, max_depth=DEPTH, seed=11)
        dev = torch.device("cuda", rank)
        # -- tile split: own row blocks in one wavefront, fp64 rows gathered on rank 0
        rows = ctx.rows_of_part(H, BLOCK, world, rank)
        local = torch.zeros((rows * W, 3), dtype=torch.float64, device=dev)
        ctx.render_rows_dev(P, BLOCK, world, rank, 0, SPP, local.data_ptr())
        frame = torch.zeros((H * W, 3), dtype=torch.float64, device=dev)
        ctx.framebuffer_gather(local.data_ptr(), W * 24, H, BLOCK, frame.data_ptr(), 0)
        # ... and the resolved 8-bit rows (row_bytes = 3 * W is not a multiple of 16 here: the byte-wise placement path)
        rgb_local = torch.zeros((rows * W, 3), dtype=torch.uint8, device=dev)
        ctx.resolve_dev(rows * W, local.data_ptr(), SPP, rgb_local.data_ptr())
        frame8 = torch.zeros((H * W, 3), dtype=torch.uint8, device=dev)
        ctx.framebuffer_gather(rgb_local.data_ptr(), W * 3, H, BLOCK, frame8.data_ptr(), 0)
        # -- sample split: own half of the samples of every pixel, partial sums added onto rank 0
        s0, s1 = rank * (SPP // world), (rank + 1) * (SPP // world)
        part = torch.zeros((H * W, 3), dtype=torch.float64, device=dev)
        ctx.render_tile_dev(P, 0, 0, W, H, s0, s1, part.data_ptr())
        ctx.framebuffer_reduce(part.data_ptr(), H * W * 3, 0)
        ctx.synchronize()
        if rank == 0:
            full, _ = ctx.render_tile(P, 0, 0, W, H, 0, SPP)
            g0 = np.load(os.path.join(tmp, "gather_0.npy"))
            t0 = time.time()
            while not os.path.exists(os.path.join(tmp, "gather_1.npy")) and time.time() - t0 < 60:
                time.sleep(0.01)
            time.sleep(0.2)
            g1 = np.load(os.path.join(tmp, "gather_1.npy"))
            tile_frame = frame.cpu().numpy()
            res = dict(
                map_kept=pm["n_kept"],
                gather_equal=bool(g0.tobytes() == g1.tobytes()),
                tile_bits_equal=bool(tile_frame.tobytes() == full.tobytes()),
                tile8_equal=bool(np.array_equal(frame8.cpu().numpy(), ctx.resolve(full, SPP))),
                sample_max_rel=float(np.max(np.abs(part.cpu().numpy() - full) / (np.abs(full) + 1e-300))),
                mean=float(full.mean()),
            )
            out.update(res)
        else:
            out["rank1_map_kept"] = pm["n_kept"]
        ctx.comm_destroy()
    finally:
        ctx.close()


@pytest.mark.skipif(not have_assets("caustics"), reason="assets not staged")
def test_two_gpu_tile_and_sample_split_equal_one_gpu_frame(lib_built, tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, str(tmp_path), out), nprocs=2, join=True)
    out = dict(out)
    assert out["map_kept"] == out["rank1_map_kept"] > 0.9 * PHOTONS
    assert out["gather_equal"], "the broadcast photon map answers differently on rank 1"
    assert out["tile_bits_equal"], "tile split + gi_framebuffer_gather differs from the one-GPU frame"
    assert out["tile8_equal"]
    assert out["sample_max_rel"] < 1e-13, out["sample_max_rel"]
    assert out["mean"] > 0
