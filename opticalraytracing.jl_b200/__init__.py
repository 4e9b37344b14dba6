"""ort_b200 -- host side of the B200-native batched sequential ray tracer.

Mirrors the public API of Sagnac/OpticalRayTracing.jl for its data-parallel hot path (solve,
raytrace, full_trace, transfer, reverse_transfer, wavegrad, TSA ...) on top of the C ABI of
libort_b200.so (include/ort_b200.h).  The Julia shim a maintainer would add is in
julia/OpticalRayTracingB200.jl; Julia is not installed in this environment, so this Python
mirror is what drives the same ABI in tests and benchmarks.
"""
from . import _lib, prescriptions
from ._lib import (FAST, STRICT, FLAG_CLIP, FLAG_DOMAIN, FLAG_MISS, FLAG_TIR, Context, OrtError,
                   PinnedArray, STATS_DTYPE)

__all__ = ["_lib", "prescriptions", "Context", "OrtError", "PinnedArray", "STATS_DTYPE", "FAST",
           "STRICT", "FLAG_MISS", "FLAG_TIR", "FLAG_DOMAIN", "FLAG_CLIP"]
