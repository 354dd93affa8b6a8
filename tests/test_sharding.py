"""Row-tile partition logic and the one-process-per-GPU handle exchange, on CPU (gloo, world_size 2)."""
import os
import socket

import numpy as np
import pytest

from meteoros_b200 import sharding


@pytest.mark.parametrize("h,tile_rows,world", [(4320, 32, 8), (2160, 32, 4), (1080, 32, 2), (1080, 8, 8), (70, 16, 3), (36, 64, 2)])
def test_tiles_cover_every_row_exactly_once(h, tile_rows, world):
    owner = np.full(h, -1)
    for rank in range(world):
        for t in sharding.tiles_of_rank(h, tile_rows, world, rank):
            r0, r1 = sharding.rows_of_tile(h, tile_rows, t)
            assert r0 % 4 == 0 and (owner[r0:r1] == -1).all()
            owner[r0:r1] = rank
    assert (owner >= 0).all()
    n = sharding.num_tiles(h, tile_rows)
    counts = [len(sharding.tiles_of_rank(h, tile_rows, world, r)) for r in range(world)]
    assert sum(counts) == n and max(counts) - min(counts) <= 1
    # cyclic assignment: the marched (upper) half of the frame is spread over all ranks
    if n >= 2 * world:
        upper = owner[: h // 2]
        assert len(set(upper.tolist())) == world


def test_bad_tile_rows_rejected():
    with pytest.raises(ValueError):
        sharding.num_tiles(1080, 12)
    with pytest.raises(ValueError):
        sharding.tiles_of_rank(1080, 32, 2, 2)


def test_views_round_robin():
    got = sorted(v for r in range(8) for v in sharding.views_of_rank(256, 8, r))
    assert got == list(range(256))


class StubRenderer:
    """Stands in for CloudRenderer on a machine without a GPU: records what ShardedFrame asks of it."""

    def __init__(self, rank, height):
        self.rank, self.height, self.width = rank, height, 64
        self.calls = []
        self.output = (None, None)

    def export_image_handle(self, which):
        self.calls.append(("export", which))
        return bytes([which + 1]) * 64

    def open_peer_image(self, handle):
        assert len(handle) == 64
        self.calls.append(("open", handle[0]))
        return 0x1000 * handle[0]

    def set_cloud_output(self, hdr, mask):
        self.output = (hdr, mask)

    def set_cloud_forward(self, hdr):
        self.calls.append(("forward", hdr))

    def dispatch_cloud_tiles(self, tile_rows, begin, end, stride):
        self.calls.append(("tiles", tile_rows, begin, end, stride))

    def synchronize(self):
        self.calls.append(("sync",))

    def close_peer_image(self, ptr):
        self.calls.append(("close", ptr))


def _worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        r = StubRenderer(rank, 270)
        sf = sharding.ShardedFrame(r, dist, tile_rows=16, with_mask=True)
        sf.dispatch()
        sf.finish()
        out_during = r.output
        sf.close()
        # gather by forwarding: stores stay local, the peer image is handed to the side kernel instead
        r2 = StubRenderer(rank, 270)
        sf2 = sharding.ShardedFrame(r2, dist, tile_rows=16, with_mask=False, mode="forward")
        sf2.dispatch()
        sf2.finish()
        sf2.close()
        q.put((rank, r.calls + [("second",)] + r2.calls + [("out2", r2.output)], out_during, r.output))
    finally:
        dist.destroy_process_group()


def test_sharded_frame_handle_exchange_gloo():
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in procs:
        rank, calls, during, after = q.get(timeout=120)
        res[rank] = (calls, during, after)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    calls0, during0, _ = res[0]
    calls1, during1, after1 = res[1]
    # rank 0 exports HDR (image 0) and mask (image 2) and keeps writing locally
    assert ("export", 0) in calls0 and ("export", 2) in calls0 and during0 == (None, None)
    # rank 1 maps both and redirects its stores to rank 0's memory, then restores
    assert ("open", 1) in calls1 and ("open", 3) in calls1
    assert during1 == (0x1000, 0x3000) and after1 == (None, None)
    assert ("close", 0x1000) in calls1 and ("close", 0x3000) in calls1
    n = sharding.num_tiles(270, 16)
    assert ("tiles", 16, 0, n, 2) in calls0 and ("tiles", 16, 1, n, 2) in calls1
    # forward mode: rank 1 arms the forwarder with rank 0's HDR image, never redirects its stores, disarms on close
    second1 = calls1[calls1.index(("second",)):]
    assert ("forward", 0x1000) in second1 and second1.index(("forward", 0x1000)) < second1.index(("tiles", 16, 1, n, 2))
    assert ("forward", None) in second1 and ("out2", (None, None)) in second1
    assert not any(c[0] == "forward" for c in calls0)
    with pytest.raises(ValueError):
        sharding.ShardedFrame(StubRenderer(0, 270), None, with_mask=True, mode="forward")
