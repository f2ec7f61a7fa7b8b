"""uvrt-b200: B200-native wavefront hot path of the UV-robot dose ray tracer.

The product is the two shared libraries under _build/ (built by `make` / __graft_entry__.build()):

  libuvrt.so       sm_100a CUDA kernels behind the C ABI of include/uvrt.h
  libuvrt_host.so  C++ Mesh / BVH / RayTracer (the reference's interfaces) + include/uvrt_host.h

This Python package is only a ctypes view of those ABIs for the tests and bench.py.  It has no
compute path of its own and no CPU fallback: without the built libraries it raises, and without a
CUDA device `Context()` / `Sim.init()` raise UvrtError.
"""
from .binding import (  # noqa: F401
    Context, Sim, SimParams, UvrtError, BUF, STAGE, lib, host, build, build_dir, include_dir,
    declared_symbols, RAY_DTYPE, NODE_DTYPE,
)
