"""nsvf_b200 — B200-native (sm_100a) sparse-voxel ray-marching hot path, drop-in for NSVF's fairnr.clib
and the torch-level interpolation / compositing / prune / split stages around it.

Layout:
  csrc/            hand-written CUDA kernels + the C ABI (include/nsvf_b200.h)
  _lib.py          ctypes binding of lib/libnsvf_b200.so (no fallback)
  clib/            mirror of fairnr/clib: `_ext` (Level-1, the 7 pybind functions) and the
                   autograd.Function callables of fairnr/clib/__init__.py (Level-2)
  ops.py           differentiable fused ops: trilinear_embed, composite, fill_in_blend, and the field's non-GEMM passes
                   (linear_layernorm_relu, posenc, narrow_linear)
  encoder.py renderer.py geometry.py split.py pipeline.py dist.py
                   host mirror of SparseVoxelEncoder / VolumeRenderer / the model's ray-marching pipeline
  field.py         the nsvf_base field MLP: cuBLAS contractions + hand-written passes, CUDA-graph replay per chunk
  blas.py          selects cuBLAS 12.9's fp32-accurate BF16x9 tensor-core algorithm (call before `import torch`)
"""
__version__ = "0.1.0"
