"""ctypes loader for the CPU oracle (oracle/vkrs_oracle.c + vkrs_oracle_host.cpp).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / ``--impl reference`` legs of bench.py.  The product package
(vkradixsort_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libvkrs_oracle.so")

_u32p = ctypes.POINTER(ctypes.c_uint32)
_u64p = ctypes.POINTER(ctypes.c_uint64)


class PushConstants(ctypes.Structure):
    """multiradixsort/include/MultiRadixSortPass.h:17-31 (16 bytes, std430)."""

    _fields_ = [
        ("g_num_elements", ctypes.c_uint32),
        ("g_shift", ctypes.c_uint32),
        ("g_num_workgroups", ctypes.c_uint32),
        ("g_num_blocks_per_workgroup", ctypes.c_uint32),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc/g++ (a few seconds); returns the .so path."""
    srcs = [os.path.join(_HERE, f) for f in ("vkrs_oracle.c", "vkrs_oracle_host.cpp", "Makefile")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    if stale:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.vkrs_oracle_global_invocation_size.restype = ctypes.c_uint32
        L.vkrs_oracle_global_invocation_size.argtypes = [ctypes.c_uint32, ctypes.c_uint32]
        L.vkrs_oracle_workgroup_count.restype = ctypes.c_uint32
        L.vkrs_oracle_workgroup_count.argtypes = [ctypes.c_uint32, ctypes.c_uint32]
        L.vkrs_oracle_multi_histograms.restype = None
        L.vkrs_oracle_multi_histograms.argtypes = [_u32p, _u32p, ctypes.POINTER(PushConstants)]
        L.vkrs_oracle_multi_scatter.restype = None
        L.vkrs_oracle_multi_scatter.argtypes = [_u32p, _u32p, _u32p, ctypes.POINTER(PushConstants), _u32p, _u32p]
        L.vkrs_oracle_multi_sort.restype = None
        L.vkrs_oracle_multi_sort.argtypes = [_u32p, _u32p, _u32p, ctypes.c_uint32, ctypes.c_uint32,
                                             ctypes.c_uint32, _u32p, _u32p]
        L.vkrs_oracle_single_sort.restype = None
        L.vkrs_oracle_single_sort.argtypes = [_u32p, _u32p, ctypes.c_uint32, _u32p, _u32p]
        L.vkrs_oracle_multi_sort64.restype = None
        L.vkrs_oracle_multi_sort64.argtypes = [_u64p, _u64p, _u32p, ctypes.c_uint32, ctypes.c_uint32,
                                               ctypes.c_uint32]
        L.vkrs_oracle_generate_random.restype = None
        L.vkrs_oracle_generate_random.argtypes = [_u32p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32]
        L.vkrs_oracle_generate_random64.restype = None
        L.vkrs_oracle_generate_random64.argtypes = [_u64p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64]
        L.vkrs_oracle_std_sort.restype = ctypes.c_double
        L.vkrs_oracle_std_sort.argtypes = [_u32p, ctypes.c_uint64]
        L.vkrs_oracle_std_sort64.restype = ctypes.c_double
        L.vkrs_oracle_std_sort64.argtypes = [_u64p, ctypes.c_uint64]
        L.vkrs_oracle_parallel_sort.restype = ctypes.c_double
        L.vkrs_oracle_parallel_sort.argtypes = [_u32p, ctypes.c_uint64]
        L.vkrs_oracle_num_threads.restype = ctypes.c_int
        L.vkrs_oracle_num_threads.argtypes = []
        L.vkrs_oracle_set_num_threads.restype = None
        L.vkrs_oracle_set_num_threads.argtypes = [ctypes.c_int]
        L.vkrs_oracle_stable_sort_pairs.restype = ctypes.c_double
        L.vkrs_oracle_stable_sort_pairs.argtypes = [_u32p, _u32p, ctypes.c_uint64]
        L.vkrs_oracle_test_sort.restype = ctypes.c_int64
        L.vkrs_oracle_test_sort.argtypes = [_u32p, ctypes.c_uint64, _u32p, ctypes.c_uint64]
        _lib = L
    return _lib


def _p32(a: np.ndarray | None):
    if a is None:
        return None
    assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u32p)


def _p64(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u64p)


def global_invocation_size(n: int, nb: int) -> int:
    return int(lib().vkrs_oracle_global_invocation_size(n, nb))


def workgroup_count(n: int, nb: int) -> int:
    return int(lib().vkrs_oracle_workgroup_count(n, nb))


def push_constants(n: int, shift: int, nb: int, W: int | None = None) -> PushConstants:
    return PushConstants(n, shift, workgroup_count(n, nb) if W is None else W, nb)


def multi_histograms(keys: np.ndarray, pc: PushConstants) -> np.ndarray:
    hist = np.zeros(max(1, pc.g_num_workgroups) * 256, dtype=np.uint32)
    lib().vkrs_oracle_multi_histograms(_p32(keys), _p32(hist), ctypes.byref(pc))
    return hist[: pc.g_num_workgroups * 256]


def multi_scatter(keys: np.ndarray, hist: np.ndarray, pc: PushConstants, values: np.ndarray | None = None):
    out = np.zeros_like(keys)
    vout = np.zeros_like(values) if values is not None else None
    lib().vkrs_oracle_multi_scatter(_p32(keys), _p32(out), _p32(np.ascontiguousarray(hist)), ctypes.byref(pc),
                                    _p32(values), _p32(vout))
    return (out, vout) if values is not None else out


def multi_sort(keys: np.ndarray, nb: int = 32, values: np.ndarray | None = None):
    """Returns (buf0, buf1, hist[, val0, val1]) after the reference's 4-pass loop."""
    n = keys.shape[0]
    buf0 = keys.copy()
    buf1 = np.zeros_like(buf0)
    W = workgroup_count(n, nb)
    hist = np.zeros(max(1, W) * 256, dtype=np.uint32)
    v0 = values.copy() if values is not None else None
    v1 = np.zeros_like(v0) if values is not None else None
    lib().vkrs_oracle_multi_sort(_p32(buf0), _p32(buf1), _p32(hist), n, nb, 4, _p32(v0), _p32(v1))
    if values is not None:
        return buf0, buf1, hist[: W * 256], v0, v1
    return buf0, buf1, hist[: W * 256]


def single_sort(keys: np.ndarray, values: np.ndarray | None = None):
    n = keys.shape[0]
    buf0 = keys.copy()
    buf1 = np.zeros_like(buf0)
    v0 = values.copy() if values is not None else None
    v1 = np.zeros_like(v0) if values is not None else None
    lib().vkrs_oracle_single_sort(_p32(buf0), _p32(buf1), n, _p32(v0), _p32(v1))
    return (buf0, v0) if values is not None else buf0


def multi_sort64(keys: np.ndarray, nb: int = 32) -> np.ndarray:
    n = keys.shape[0]
    buf0 = keys.copy()
    buf1 = np.zeros_like(buf0)
    hist = np.zeros(max(1, workgroup_count(n, nb)) * 256, dtype=np.uint32)
    lib().vkrs_oracle_multi_sort64(_p64(buf0), _p64(buf1), _p32(hist), n, nb, 8)
    return buf0


def generate_random(n: int, seed: int, max_value: int = 0xFFFFFFFF) -> np.ndarray:
    out = np.empty(n, dtype=np.uint32)
    lib().vkrs_oracle_generate_random(_p32(out), n, seed, max_value)
    return out


def generate_random64(n: int, seed: int, max_value: int = 0x0FFFFFFFFFFF) -> np.ndarray:
    out = np.empty(n, dtype=np.uint64)
    lib().vkrs_oracle_generate_random64(_p64(out), n, seed, max_value)
    return out


def std_sort(keys: np.ndarray):
    """In-place single-thread std::sort; returns milliseconds (MultiRadixSort.cpp:141-146)."""
    return float(lib().vkrs_oracle_std_sort(_p32(keys), keys.shape[0]))


def parallel_sort(keys: np.ndarray):
    return float(lib().vkrs_oracle_parallel_sort(_p32(keys), keys.shape[0]))


def num_threads() -> int:
    return int(lib().vkrs_oracle_num_threads())


def set_num_threads(threads: int) -> None:
    """OpenMP threads of the restated shaders / parallel sort (overrides an inherited OMP_NUM_THREADS)."""
    lib().vkrs_oracle_set_num_threads(int(threads))


def stable_sort_pairs(keys: np.ndarray, values: np.ndarray):
    return float(lib().vkrs_oracle_stable_sort_pairs(_p32(keys), _p32(values), keys.shape[0]))


def test_sort(reference: np.ndarray, out: np.ndarray) -> int:
    """-1 equal, -2 size mismatch, else first differing index (MultiRadixSort.cpp:148-161)."""
    return int(lib().vkrs_oracle_test_sort(_p32(reference), reference.shape[0], _p32(out), out.shape[0]))
