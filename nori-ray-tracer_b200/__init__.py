"""nori-ray-tracer_b200: the B200-native rendering hot path of Nori behind a C ABI.

Only what the path needs lives here:
  csrc/        hand-written sm_100a CUDA kernels + the extern "C" nori_gpu_* boundary (libnori_gpu.so)
  abi.py       ctypes mirror of include/nori_gpu.h
  nscene.py    flat scene container <-> nori_gpu_scene
  gpu.py       thin binding over libnori_gpu.so (fails loudly when the library or a GPU is missing)
  render.py    host-side mirror of the reference's RenderThread (render.cpp:135-290)
The directory name carries a hyphen (it is the repo's layout contract), so import it through
`__graft_entry__.import_package()` which registers it as `nori_ray_tracer_b200`.
"""
from . import abi, nscene  # noqa: F401

__all__ = ["abi", "nscene"]
