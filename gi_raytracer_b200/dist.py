"""Multi-GPU plumbing: one process per GPU.

The path shards without a data-path collective (pixels, samples and photons are independent).  Only two
exchanges exist: the photon map is built once on the root rank and broadcast as one slab, and the per-rank framebuffer
partial sums are reduced (sample split) or gathered (tile split) at the end of a frame.  On GPUs both are the C ABI's own
NCCL calls (gi_comm_init / gi_photon_map_bcast / gi_framebuffer_reduce / gi_framebuffer_gather, csrc/gi_comm.inc); torch.distributed
only carries the 128-byte communicator id between the processes (init_comm).  The torch.distributed forms of the same protocols
(broadcast_slab, reduce_accum, gather_rows) remain for the world-2 gloo tests on CPU."""
import numpy as np
import torch
import torch.distributed as dist


class _DevMem:
    """A raw device range exposed through __cuda_array_interface__ so torch can wrap it without a copy."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def wrap_device_bytes(ptr, nbytes, device):
    return torch.as_tensor(_DevMem(ptr, nbytes), device=device)


def sample_ranges(spp_per_rank, world):
    """Sample-index split (weak scaling): rank r renders samples [r*spp, (r+1)*spp) of every pixel."""
    return [(r * spp_per_rank, (r + 1) * spp_per_rank) for r in range(world)]


def row_blocks(height, world, block=16):
    """Tile split (strong scaling): interleaved blocks of `block` rows, round-robin over ranks, like the reference's
    `schedule(dynamic, 10)` row loop.  Returns per-rank lists of (y0, y1)."""
    out = [[] for _ in range(world)]
    for k, y0 in enumerate(range(0, height, block)):
        out[k % world].append((y0, min(height, y0 + block)))
    return out


def init_comm(ctx, rank, world, device):
    """gi_comm_init on every rank: rank 0 makes the NCCL unique id (gi_comm_unique_id), torch.distributed carries its 128 bytes."""
    from . import capi
    buf = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(buf, src=0)
    ctx.comm_init(bytes(buf.cpu().numpy().tobytes()), rank, world)


def broadcast_slab(make_root_slab, reserve, adopt, rank, root, device):
    """Broadcast a photon-map slab.  make_root_slab() -> (ptr, nbytes) on the root; reserve(nbytes) -> ptr and
    adopt(nbytes) on the others.  Pointers are device pointers (or torch CPU uint8 tensors under gloo)."""
    size = torch.zeros(1, dtype=torch.int64, device=device)
    buf = None
    if rank == root:
        buf, nbytes = make_root_slab()
        size[0] = nbytes
    dist.broadcast(size, src=root)
    nbytes = int(size.item())
    if rank != root:
        buf = reserve(nbytes)
    t = buf if isinstance(buf, torch.Tensor) else wrap_device_bytes(buf, nbytes, device)
    dist.broadcast(t, src=root)
    if rank != root:
        adopt(nbytes)
    return nbytes


def share_photon_map(ctx, rank, world, root=0):
    """Root has a built map; every other rank receives it over NCCL and adopts it.  With a gi_comm on the context this is the C ABI's
    own gi_photon_map_bcast; the torch.distributed form below is what the gloo tests drive."""
    if world == 1:
        return ctx.photon_map_slab()[1]
    if ctx.comm_info()["nranks"] == world:
        ctx.synchronize()
        ctx.photon_map_bcast(root)
        ctx.synchronize()
        return ctx.photon_map_slab()[1]
    device = torch.device("cuda", ctx.device)
    ctx.synchronize()
    n = broadcast_slab(ctx.photon_map_slab, ctx.photon_map_reserve_slab, ctx.photon_map_adopt_slab, rank, root, device)
    torch.cuda.synchronize(device)
    return n


def reduce_accum(accum, root=0):
    """Sum per-rank framebuffer partial sums (sample split) onto the root."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum, dst=root, op=dist.ReduceOp.SUM)
    return accum


def gather_rows(local_rows, blocks, height, width, root=0):
    """Tile split: all-gather the per-rank row blocks and reassemble the frame (every rank gets the full frame).
    local_rows: tensor [n_local_rows, width, 3]; blocks: the row_blocks() plan."""
    world = dist.get_world_size()
    counts = [sum(y1 - y0 for y0, y1 in b) for b in blocks]
    pad = max(counts)
    send = torch.zeros((pad, width, 3), dtype=local_rows.dtype, device=local_rows.device)
    send[: local_rows.shape[0]] = local_rows
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send)
    frame = torch.empty((height, width, 3), dtype=local_rows.dtype, device=local_rows.device)
    for r in range(world):
        k = 0
        for y0, y1 in blocks[r]:
            frame[y0:y1] = recv[r][k:k + (y1 - y0)]
            k += y1 - y0
    return frame
