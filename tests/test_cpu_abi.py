"""CPU (no GPU): the C-ABI library loads and exports every symbol include/hpp_b200.h declares; the product has no CPU
fallback (creating a context without a B200 is an error); the Python host mirror refuses to run without the library."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "hpp_b200.h")
LIB = os.path.join(ROOT, "hybrid_path_planning_b200", "libhpp_b200.so")


def declared_symbols():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hpp_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_whole_path():
    names = declared_symbols()
    for must in ("hpp_map_create", "hpp_map_layers", "hpp_inflate", "hpp_collide_circles", "hpp_collide_sat",
                 "hpp_heuristic_rs", "hpp_heuristic_dubins", "hpp_heuristic_batch", "hpp_expand", "hpp_rs_shot",
                 "hpp_dubins_shot", "hpp_plan_batch"):
        assert must in names


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "libhpp_b200.so missing: run __graft_entry__.build()"
    L = ctypes.CDLL(LIB)
    missing = [n for n in declared_symbols() if not hasattr(L, n)]
    assert not missing, f"declared in include/hpp_b200.h but not exported: {missing}"
    L.hpp_version.restype = ctypes.c_char_p
    assert b"sm_100a" in L.hpp_version()


def test_no_cpu_fallback():
    """Without a usable sm_100 device the context cannot be created -- and nothing else can run without a context."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the fallback check is for CPU-only hosts")
    L = ctypes.CDLL(LIB)
    h = ctypes.c_void_p()
    assert L.hpp_ctx_create(0, ctypes.byref(h)) == 2  # HPP_ERR_CUDA
    assert not h.value
    import hybrid_path_planning_b200 as hpp
    with pytest.raises(hpp.HppError):
        hpp.Context(0)


def test_header_cites_the_reference_interface():
    src = open(HDR).read()
    for cite in ("RS/include/Grid_map.h:85", "RS/include/HybridAstar.h:62", "RS/include/HybridAstar.h:52",
                 "RS/include/reeds_shepp_path.h:51", "DB/include/dubins_path.h:57", "MI/src/map_inflate_node.cpp"):
        assert cite in src


def _classify_c(param):
    p = param.strip()
    if "*" in p or "[" in p:
        return "p"
    t = re.sub(r"\b(const|volatile)\b", "", p).split()
    t = " ".join(t[:-1]) if len(t) > 1 else t[0]
    if t == "float":
        return "f"
    if t == "double":
        return "d"
    if t in ("size_t", "uint64_t", "int64_t"):
        return "z"
    if t in ("int", "int32_t", "uint32_t", "hpp_status"):
        return "i"
    raise AssertionError(f"unclassified parameter type: {param!r}")


def _classify_ctypes(t):
    if t is ctypes.c_void_p or t is ctypes.c_char_p or hasattr(t, "contents") or issubclass(t, ctypes._Pointer):
        return "p"
    if t is ctypes.c_float:
        return "f"
    if t is ctypes.c_double:
        return "d"
    if ctypes.sizeof(t) == 8:
        return "z"
    return "i"


def test_ctypes_signatures_match_the_header():
    """every prototype of include/hpp_b200.h against the argtypes the Python binding declares: same parameter count, and
    pointer / float / double / 32-bit / 64-bit class per parameter (a float passed as a double is silent garbage)"""
    import hybrid_path_planning_b200.capi as capi
    L = capi.lib()
    src = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)
    protos = re.findall(r"\b(hpp_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S)
    assert len(protos) >= 30
    checked = 0
    for name, params in protos:
        f = getattr(L, name)
        if f.argtypes is None:
            continue
        want = [] if params.strip() in ("", "void") else [_classify_c(p) for p in params.split(",")]
        got = [_classify_ctypes(t) for t in f.argtypes]
        assert got == want, f"{name}: header {want} vs ctypes {got}"
        checked += 1
    assert checked >= 30
