"""Row-tile sharding of one frame across the GPUs of a box, one process per GPU.

The Cloud pass is embarrassingly parallel per ray: noise volumes and uniforms are replicated
on every GPU, the ray grid is cut into tiles of `tile_rows` pixel rows (a multiple of 8, so 4x4 pixel blocks and
the kernel's 16x8 CTAs stay intact) and tile t belongs to rank t % world -- cyclic, because the lower half of a
frame is horizon-culled and contiguous bands would leave half the GPUs idle.

There is no collective on the data path.  Rank 0 exports its HDR / mask images as CUDA IPC handles, the other ranks
map them (mtOpenPeerImage) and point the kernel's output there (mtSetCloudOutput): finished rays are stored
straight into GPU 0's memory over NVLink by the ray-march kernel's own float4 stores, overlapping the march of the
rays still in flight.  torch.distributed carries only the 64-byte handles and the barriers.
"""
from __future__ import annotations

from dataclasses import dataclass

from .api import IMAGE_CLOUD_CUR, IMAGE_GODRAY_MASK


def num_tiles(height: int, tile_rows: int) -> int:
    if tile_rows < 8 or tile_rows % 8:
        raise ValueError("tile_rows must be a positive multiple of 8")
    return (height + tile_rows - 1) // tile_rows


def tiles_of_rank(height: int, tile_rows: int, world: int, rank: int) -> range:
    """Tiles owned by `rank`: rank, rank + world, ... (cyclic)."""
    if not (0 <= rank < world):
        raise ValueError("rank outside world")
    return range(rank, num_tiles(height, tile_rows), world)


def rows_of_tile(height: int, tile_rows: int, tile: int) -> tuple[int, int]:
    return tile * tile_rows, min(height, (tile + 1) * tile_rows)


def views_of_rank(n_views: int, world: int, rank: int) -> range:
    """Batch-of-views mode (BASELINE config 5): whole frames round-robin over ranks, no communication."""
    return range(rank, n_views, world)


@dataclass
class PeerOutput:
    hdr_ptr: int = 0
    mask_ptr: int = 0


class ShardedFrame:
    """Drives one CloudRenderer per rank so that rank 0 ends up holding the whole frame.

    `dist` is torch.distributed (or any object with get_rank/get_world_size/broadcast_object_list/barrier), passed in
    so the host logic can be exercised on CPU with the gloo backend and a stub renderer.
    """

    def __init__(self, renderer, dist, tile_rows: int = 32, with_mask: bool = True, mode: str = "peer_store", groups: int = 4):
        """mode "peer_store": the march kernel's own stores land in GPU 0's image (fused compute + transfer).
        mode "bulk_store": the same, but a warp's 16x2 pixels leave through shared memory and two bulk asynchronous copies
        (cp.async.bulk, mtSetCloudStoreMode): the marching warp never waits for the remote write.
        mode "copy": tiles are rendered locally in `groups` launches and each finished group is pushed to GPU 0 by the
        copy engine (mtCopyTilesToPeer) while the next group renders.
        mode "forward": the march kernel stores locally and counts finished CTAs per tile; a small side kernel on a
        high-priority stream pushes each finished tile to GPU 0 (mtSetCloudForward) -- no marching warp waits on NVLink.
        mode "local": measurement only -- every rank keeps its tiles (no gather), to separate compute share from transfer."""
        if mode not in ("peer_store", "bulk_store", "copy", "forward", "local"):
            raise ValueError("mode must be 'peer_store', 'bulk_store', 'copy', 'forward' or 'local'")
        if mode == "forward" and with_mask:
            raise ValueError("mode 'forward' gathers the HDR image only (with_mask=False)")
        self.mode, self.groups = mode, max(1, int(groups))
        self.r = renderer
        self.dist = dist
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()
        self.tile_rows = tile_rows
        self.with_mask = with_mask
        self.peer = PeerOutput()
        self._connect()

    def _connect(self):
        handles = [None, None]
        if self.rank == 0:
            handles[0] = self.r.export_image_handle(IMAGE_CLOUD_CUR)
            handles[1] = self.r.export_image_handle(IMAGE_GODRAY_MASK) if self.with_mask else b""
        if self.world > 1:
            self.dist.broadcast_object_list(handles, src=0)
        if self.rank != 0:
            self.peer.hdr_ptr = self.r.open_peer_image(handles[0])
            # without god rays the mask never leaves the GPU that made it: stores stay local
            self.peer.mask_ptr = self.r.open_peer_image(handles[1]) if self.with_mask else 0
            if self.mode in ("peer_store", "bulk_store"):
                self.r.set_cloud_output(self.peer.hdr_ptr, self.peer.mask_ptr or None)
                if self.mode == "bulk_store":
                    self.r.set_cloud_store_mode(1)
            elif self.mode == "forward":
                self.r.set_cloud_forward(self.peer.hdr_ptr)
        if self.world > 1:
            self.dist.barrier()

    def dispatch(self):
        """Launch this rank's tiles (asynchronous on the renderer's stream)."""
        n = num_tiles(self.r.height, self.tile_rows)
        if self.mode in ("peer_store", "bulk_store", "forward", "local") or self.rank == 0:
            self.r.dispatch_cloud_tiles(self.tile_rows, self.rank, n, self.world)
            return
        mine = len(range(self.rank, n, self.world))
        per = (mine + self.groups - 1) // self.groups
        for g in range(self.groups):
            first = self.rank + g * per * self.world
            last = min(n, self.rank + (g + 1) * per * self.world)
            if first >= last:
                break
            self.r.dispatch_cloud_tiles(self.tile_rows, first, last, self.world)
            self.r.copy_tiles_to_peer(IMAGE_CLOUD_CUR, self.tile_rows, first, last, self.world, self.peer.hdr_ptr)
            if self.with_mask:
                self.r.copy_tiles_to_peer(IMAGE_GODRAY_MASK, self.tile_rows, first, last, self.world, self.peer.mask_ptr)

    def finish(self):
        """Frame boundary: every rank's stores have landed in rank 0's image."""
        self.r.synchronize()
        if self.world > 1:
            self.dist.barrier()

    def close(self):
        if self.rank != 0:
            if self.mode in ("peer_store", "bulk_store"):
                self.r.set_cloud_output(None, None)
                if self.mode == "bulk_store":
                    self.r.set_cloud_store_mode(0)
            elif self.mode == "forward":
                self.r.set_cloud_forward(None)
            if self.peer.hdr_ptr:
                self.r.close_peer_image(self.peer.hdr_ptr)
            if self.peer.mask_ptr:
                self.r.close_peer_image(self.peer.mask_ptr)
            self.peer = PeerOutput()
        if self.world > 1:
            self.dist.barrier()
