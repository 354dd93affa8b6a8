import sys; sys.path.insert(0,'/root/repo')
from meteoros_b200 import api, scene, textures
w,h=7680,4320
cam, sc = scene.Camera(w,h), scene.Scene(); sc.update_time(1/60)
with api.CloudRenderer(w,h) as r:
    r.upload_noise(textures.load_noise()); r.set_camera(cam.ubo()); r.set_time(sc.ubo())
    def timeit(fn, n=5):
        for _ in range(2): fn()
        r.synchronize(); ts=[]
        for _ in range(n):
            r.event_record(0); fn(); r.event_record(1); ts.append(r.event_elapsed_ms(0,1))
        return min(ts)
    full = timeit(lambda: r.dispatch_cloud_full())
    print('full 8K', full, 'ideal 1/8', full/8)
    for tr in (8, 16, 32, 64, 128, 256):
        n=(h+tr-1)//tr
        for rank in (0, 3, 7):
            t = timeit(lambda: r.dispatch_cloud_tiles(tr, rank, n, 8))
            print('tile_rows', tr, 'rank', rank, 'ms', round(t,3))
