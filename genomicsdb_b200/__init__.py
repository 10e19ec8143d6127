"""gdb-b200: B200-native GenomicsDB variant-combine hot path.

The product is native: CUDA kernels (csrc/gdb_kernels*.cuh) behind the C ABI of include/gdbgpu.h, with a host-side C++
mirror of the reference's plugin surface (csrc/gdb_host_api.*).  This Python package only binds the C ABI for tests,
bench.py and tooling; it holds no compute path and no CPU fallback.
"""
from . import capi  # noqa: F401
