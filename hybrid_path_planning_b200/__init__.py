"""hybrid_path_planning_b200 -- B200-native (sm_100a) core of the Hybrid A* planners of
armando-genis/hybrid_path_planning, behind a C ABI (include/hpp_b200.h).

The product is ``libhpp_b200.so`` (hand-written CUDA, built in-tree by ``csrc/build.sh``) plus the C++ host
shells under ``host/`` that keep the reference's class API.  This Python package is plumbing only: a ctypes
binding of the C ABI used by the tests and by bench.py.  There is no CPU fallback: importing works anywhere,
but every call needs the shared library and a B200.
"""
from .capi import (Context, Map, Comm, HppError, lib, lib_path, plan_batch_multi, comm_unique_id, shard_range,  # noqa: F401
                   DEFAULT_PARAMS, VARIANT_RS, VARIANT_DUBINS)

__all__ = ["Context", "Map", "Comm", "HppError", "lib", "lib_path", "plan_batch_multi", "comm_unique_id", "shard_range",
           "DEFAULT_PARAMS", "VARIANT_RS", "VARIANT_DUBINS"]
