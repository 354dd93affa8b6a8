#!/usr/bin/env python
"""Multi-GPU correctness check (run under torchrun with 2+ ranks): the row-tile sharded frame gathered on rank 0 --
by kernel peer stores (direct and through cp.async.bulk), by copy-engine pushes and by the tile-forwarding side kernel -- must be bit-identical to the
single-GPU frame.  Each mode runs three frames in a row so that frame-to-frame hazards would show."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from meteoros_b200 import api, scene, sharding, textures  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w, h = 1920, 1080
cam, sc = scene.Camera(w, h), scene.Scene()
sc.update_time(1 / 60)
ok = True
for mode in ("peer_store", "bulk_store", "copy", "forward"):
    for with_mask in ((False,) if mode == "forward" else (True, False)):
        with api.CloudRenderer(w, h, device=local) as r:
            r.upload_noise(textures.load_noise())
            r.set_camera(cam.ubo()); r.set_time(sc.ubo())
            sf = sharding.ShardedFrame(r, dist, tile_rows=8, with_mask=with_mask, mode=mode)
            for _ in range(3):
                sf.dispatch()
                sf.finish()
            if rank == 0:
                got_hdr, got_mask = r.read_image(api.IMAGE_CLOUD_CUR), r.read_image(api.IMAGE_GODRAY_MASK)
            sf.close()
            if rank == 0:
                r.clear_images()
                r.dispatch_cloud_full()
                want_hdr, want_mask = r.read_image(api.IMAGE_CLOUD_CUR), r.read_image(api.IMAGE_GODRAY_MASK)
                same = np.array_equal(got_hdr, want_hdr)
                if with_mask:
                    same = same and np.array_equal(got_mask, want_mask)
                print(f"mode={mode} with_mask={with_mask} world={world}: gathered frame bit-identical to single-GPU: {same}")
                ok = ok and same
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
