"""meteoros_b200 -- B200-native (sm_100a CUDA) implementation of the Meteoros cloud-rendering hot path.

  meteoros_b200.api       CloudRenderer: the reference's dispatch surface over the C ABI (include/meteoros_b200.h)
  meteoros_b200.scene     Camera / Scene / Sky uniform producers (camera.cpp, Scene.cpp, Sky.cpp)
  meteoros_b200.textures  the reference's noise inputs
  meteoros_b200.sharding  row-tile partition + peer-mapped gather for one-process-per-GPU runs
  meteoros_b200/csrc      the kernels and the C ABI (libmeteoros_b200.so)
"""
from . import scene, textures  # noqa: F401

__all__ = ["scene", "textures", "api"]
