"""N>1 host logic on CPU: world_size-2 (and 3) gloo processes each own the statically interleaved tiles
tile_id % world == rank, pack them tile-major exactly like gort_render_shard_device's slab, gather to
rank 0 (what bench.py does with NCCL) and un-swizzle into the row-major frame."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import common as Cm


def synthetic_frame(width, height):
    y, x = np.mgrid[0:height, 0:width]
    img = np.zeros((height, width, 4), dtype=np.uint8)
    img[..., 0] = (x * 7 + y * 3) % 256
    img[..., 1] = (x ^ y) % 256
    img[..., 2] = (x * y) % 251
    img[..., 3] = 255
    return img


def pack_slab(gort, frame, rank, world):
    """What a rank's device slab holds: its tiles in order, each 32x32x4, zero-padded outside the image."""
    h, w = frame.shape[:2]
    per = gort.shard_slab_bytes(w, h, world) // (32 * 32 * 4)
    slab = np.zeros((per, 32, 32, 4), dtype=np.uint8)
    tx = (w + 31) // 32
    for j, t in enumerate(gort.tiles_of_shard(w, h, rank, world)):
        x0, y0 = (t % tx) * 32, (t // tx) * 32
        tile = frame[y0:y0 + 32, x0:x0 + 32]
        slab[j, :tile.shape[0], :tile.shape[1]] = tile
    return slab


def _worker(rank, world, port, width, height, q):
    sys.path.insert(0, Cm.ROOT)
    import importlib
    gort = importlib.import_module("concurrent-raytracer-go_b200")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frame = synthetic_frame(width, height)
    slab = torch.from_numpy(pack_slab(gort, frame, rank, world))
    gathered = [torch.empty_like(slab) for _ in range(world)] if rank == 0 else None
    dist.gather(slab, gathered, dst=0)
    if rank == 0:
        out = gort.unswizzle_host(torch.stack(gathered).numpy(), world, width, height)
        q.put(bool((out == frame).all()))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,width,height", [(2, 800, 600), (3, 70, 45)])
def test_interleaved_tiles_gather_gloo(world, width, height):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, width, height, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


def test_tiles_partition(gort):
    for (w, h, n) in [(800, 600, 8), (1200, 900, 4), (33, 65, 2), (1, 1, 8)]:
        tiles = [gort.tiles_of_shard(w, h, r, n) for r in range(n)]
        flat = sorted(t for ts in tiles for t in ts)
        assert flat == list(range(((w + 31) // 32) * ((h + 31) // 32)))
        assert max(len(t) for t in tiles) - min(len(t) for t in tiles) <= 1
