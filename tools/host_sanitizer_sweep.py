#!/usr/bin/env python
"""Random scenes through the host build of the kernel cores (tests/hostsim) loaded from a sanitizer-instrumented library.

  cd tests/hostsim && g++ -mfma -O1 -g -std=c++17 -fPIC -shared -ffp-contract=off -fsanitize=undefined \
      -fno-sanitize-recover=undefined -x c++ hostsim.cpp -o /tmp/libhostsim_ubsan.so && cd ../..
  LD_PRELOAD=$(gcc -print-file-name=libubsan.so) python tools/host_sanitizer_sweep.py /tmp/libhostsim_ubsan.so
(same with -fsanitize=address and libasan.so, ASAN_OPTIONS=detect_leaks=0).  Device code cannot run under these tools; the
cores are the same source the kernels compile (csrc/*_core.cuh), so index arithmetic and struct accesses are covered."""
import sys, ctypes as C, numpy as np
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
import hostsim
hostsim._lib = C.CDLL(sys.argv[1])
import oracle
from meteoros_b200 import scene, textures
noise=textures.load_noise()
rng=np.random.default_rng(77)
for trial in range(40):
    w,h=int(rng.integers(9,70)),int(rng.integers(9,50))
    ey=float(-rng.choice([0.0,10.0,5e3,7.4e3,7.6e3,9e3,1.9e4,2.1e4,3e4]))
    eye=(float(rng.uniform(-3e3,3e3)),ey,float(rng.uniform(-3e3,3e3)))
    cam=scene.Camera(w,h,eye=eye,ref=(eye[0],eye[1],eye[2]-1.0),fovy=float(rng.uniform(20,100)))
    cam.rotate_about_up(float(rng.uniform(-180,180))); cam.rotate_about_right(float(rng.uniform(-60,85)))
    old=cam.ubo(); cam.rotate_about_up(float(rng.uniform(-3,3))); new=cam.ubo()
    sc,sky,tun=scene.Scene(),scene.Sky().ubo(),scene.default_tuning()
    if trial%5==4: tun["use_weather"],tun["weather_scale"]=1,1e-4
    sc.time["time"]=(0.016,float(rng.uniform(0,500))); sc.time["frameCountMod16"]=int(rng.integers(0,16)); tm=sc.ubo()
    hostsim.cloud(new,tm,tun,noise,w,h,True,oracle.RAY_DEBUG_DTYPE)
    prev=rng.random((h,w,4),dtype=np.float32)
    hostsim.reproject(new,old,tm,prev)
    hdr=rng.random((h,w,4),dtype=np.float32)*4
    hostsim.godrays(new,sky["lightColor"][:3],prev,hdr); hostsim.tonemap(tm,hdr)
    ldr=rng.integers(0,256,(h,w,4),dtype=np.uint8); hostsim.txaa(new,old,tm,ldr,ldr[::-1].copy())
print("sanitizer sweep: kernel cores clean over 40 random scenes")
