"""ctypes loader for oracle/_ref/libvkrs_ref.so: the reference's OWN three compute shaders, compiled from their text
under /root/reference as C++ (oracle/ref_shim/, recipe `make -C oracle ref`).

TEST INFRASTRUCTURE ONLY -- it pins the hand-written restatement (oracle/vkrs_oracle.c) and the golden fixtures to the
reference's source.  /root/reference exists only in the build container; the built library travels with the repo
(oracle/_ref/ is git-ignored, not gpurun-ignored).  ``available()`` says whether it can be used here.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libvkrs_ref.so")
REFERENCE_ROOT = "/root/reference"

_u32p = ctypes.POINTER(ctypes.c_uint32)


class PushConstants(ctypes.Structure):
    """multiradixsort/include/MultiRadixSortPass.h:17-31 (16 bytes, std430)."""

    _fields_ = [("g_num_elements", ctypes.c_uint32), ("g_shift", ctypes.c_uint32),
                ("g_num_workgroups", ctypes.c_uint32), ("g_num_blocks_per_workgroup", ctypes.c_uint32)]


def build() -> bool:
    """Compiles the shaders where they lie (only possible where /root/reference exists).  True if the library exists afterwards."""
    if os.path.isdir(REFERENCE_ROOT):
        subprocess.run(["make", "-C", _HERE, "-s", "ref"], check=True)
    return os.path.exists(LIB_PATH)


def available() -> bool:
    return os.path.exists(LIB_PATH) or build()


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref is not built and /root/reference is absent")
        L = ctypes.CDLL(LIB_PATH)
        L.vkrs_ref_multi_histograms.restype = None
        L.vkrs_ref_multi_histograms.argtypes = [_u32p, _u32p, ctypes.POINTER(PushConstants)]
        L.vkrs_ref_multi_scatter.restype = None
        L.vkrs_ref_multi_scatter.argtypes = [_u32p, _u32p, _u32p, ctypes.POINTER(PushConstants)]
        L.vkrs_ref_multi_sort.restype = None
        L.vkrs_ref_multi_sort.argtypes = [_u32p, _u32p, _u32p, ctypes.c_uint32, ctypes.c_uint32]
        L.vkrs_ref_single_sort.restype = None
        L.vkrs_ref_single_sort.argtypes = [_u32p, _u32p, ctypes.c_uint32]
        L.vkrs_ref_describe.restype = ctypes.c_char_p
        _lib = L
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u32p)


def workgroup_count(n: int, nb: int) -> int:
    gis = n // nb + (1 if n % nb else 0)  # MultiRadixSort.cpp:13-15
    return (gis + 255) // 256             # ComputePass.h:24-29


def push_constants(n: int, shift: int, nb: int) -> PushConstants:
    return PushConstants(n, shift, workgroup_count(n, nb), nb)


def multi_histograms(keys: np.ndarray, pc: PushConstants) -> np.ndarray:
    """One dispatch of multi_radixsort_histograms.comp: the histogram matrix [work group][256]."""
    hist = np.zeros(max(1, pc.g_num_workgroups) * 256, dtype=np.uint32)
    lib().vkrs_ref_multi_histograms(_p(keys), _p(hist), ctypes.byref(pc))
    return hist[: pc.g_num_workgroups * 256]


def multi_scatter(keys: np.ndarray, hist: np.ndarray, pc: PushConstants) -> np.ndarray:
    """One dispatch of multi_radixsort.comp: the keys after this digit pass."""
    out = np.zeros_like(keys)
    lib().vkrs_ref_multi_scatter(_p(keys), _p(out), _p(np.ascontiguousarray(hist)), ctypes.byref(pc))
    return out


def multi_sort(keys: np.ndarray, nb: int = 32):
    """MultiRadixSort::execute's four iterations: returns (buf0 = sorted, buf1 = scratch, hist of the last pass)."""
    n = keys.shape[0]
    buf0, buf1 = keys.copy(), np.zeros_like(keys)
    W = workgroup_count(n, nb)
    hist = np.zeros(max(1, W) * 256, dtype=np.uint32)
    lib().vkrs_ref_multi_sort(_p(buf0), _p(buf1), _p(hist), n, nb)
    return buf0, buf1, hist[: W * 256]


def single_sort(keys: np.ndarray) -> np.ndarray:
    """One work group of single_radixsort.comp; the result is in buffer 0."""
    buf0, buf1 = keys.copy(), np.zeros_like(keys)
    lib().vkrs_ref_single_sort(_p(buf0), _p(buf1), keys.shape[0])
    return buf0
