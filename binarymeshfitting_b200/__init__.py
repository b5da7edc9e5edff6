"""binarymeshfitting_b200 -- B200-native (sm_100a) implementation of BinaryMeshFitting's per-chunk
extraction hot path behind a C ABI (include/bmf_b200.h, libbmf_b200.so).

  csrc/   hand-written CUDA kernels + the C ABI
  host/   C++ mirror of the reference's Sampler / DMCChunk / ChunkGenerator / MeshProcessor classes
  capi.py ctypes view of the C ABI for the test / bench harness
  world.py LOD leaf enumeration + multi-GPU partitioning (host side)
"""
from .capi import BmfError, Context, load_library  # noqa: F401
