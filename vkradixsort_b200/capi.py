"""ctypes binding of the C-ABI in include/vkradixsort_b200.h.

The shared library (vkradixsort_b200/lib/libvkradixsort_b200.so, built by
``__graft_entry__.build()`` / ``make -C vkradixsort_b200/csrc``) is the product; this module is
only the thinnest possible way to reach it from Python tests and bench.py.  There is no CPU
fallback: if the library is missing, loading raises.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libvkradixsort_b200.so")
if os.environ.get("VKRS_LIB_PATH"):  # tuning runs: a differently configured build of the same library
    LIB_PATH = os.environ["VKRS_LIB_PATH"]

VKRS_OK = 0
VKRS_ERR_INVALID_ARGUMENT = -1
VKRS_ERR_CUDA = -2
VKRS_ERR_UNSUPPORTED = -3
VKRS_ERR_INTERNAL = -4

KEY_U32, KEY_I32, KEY_F32 = 0, 1, 2

WORKGROUP_SIZE = 256
RADIX_SORT_BINS = 256

# Every symbol include/vkradixsort_b200.h declares (tests check that all of them are exported).
EXPORTED_SYMBOLS = [
    "vkrs_create", "vkrs_destroy", "vkrs_last_error", "vkrs_version",
    "vkrs_global_invocation_size", "vkrs_workgroup_count",
    "vkrs_multi_histograms", "vkrs_multi_scatter", "vkrs_multi_pass",
    "vkrs_multi_sort", "vkrs_multi_sort_pairs", "vkrs_multi_sort_u64", "vkrs_multi_sort_staged", "vkrs_multi_sort_typed",
    "vkrs_single_sort", "vkrs_sort_auto", "vkrs_multi_sort_host", "vkrs_key_range", "vkrs_partition",
    "vkrs_partition_count", "vkrs_partition_scatter_p2p", "vkrs_ipc_alloc", "vkrs_ipc_open", "vkrs_ipc_close", "vkrs_ipc_free",
    "vkrs_check_device_error", "vkrs_num_variants", "vkrs_variant_name", "vkrs_set_variant", "vkrs_get_variant",
    "vkrs_set_profiling", "vkrs_profile_collect", "vkrs_profile_entry", "vkrs_debug_counters",
    "vkrs_launch_count", "vkrs_tile_size",
    "vkrs_set_schedule", "vkrs_get_schedule", "vkrs_schedule_name", "vkrs_bucket_stats",
    "vkrs_debug_bucket_stop", "vkrs_set_key_span_hint", "vkrs_host_timings", "vkrs_resolve_schedule", "vkrs_exchange_plan", "vkrs_peer_barrier",
]

# vkrs_schedule (include/vkradixsort_b200.h)
SCHEDULE_AUTO, SCHEDULE_LSD, SCHEDULE_LSD_UNSTABLE_FIRST, SCHEDULE_BUCKET = 0, 1, 2, 3
NUM_SCHEDULES = 4


class MultiPushConstants(ctypes.Structure):
    """vkrs_multi_push_constants == MultiRadixSortPass::PushConstants (MultiRadixSortPass.h:17-31)."""

    _fields_ = [
        ("g_num_elements", ctypes.c_uint32),
        ("g_shift", ctypes.c_uint32),
        ("g_num_workgroups", ctypes.c_uint32),
        ("g_num_blocks_per_workgroup", ctypes.c_uint32),
    ]


class SinglePushConstants(ctypes.Structure):
    """vkrs_single_push_constants == SingleRadixSortPass::PushConstants (SingleRadixSortPass.h:16-18)."""

    _fields_ = [("g_num_elements", ctypes.c_uint32)]


class VkrsError(RuntimeError):
    """Raised for any non-zero status, as the reference throws std::runtime_error (ComputePass.h:51-53)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"vkrs status {status}: {message}")
        self.status = status


_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C vkradixsort_b200/csrc`. There is no CPU fallback."
        )
    L = ctypes.CDLL(LIB_PATH)
    vp, u32, u64, i32 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int
    mpc, spc = ctypes.POINTER(MultiPushConstants), ctypes.POINTER(SinglePushConstants)
    sig = {
        "vkrs_create": (i32, [ctypes.POINTER(vp), i32, u64]),
        "vkrs_destroy": (i32, [vp]),
        "vkrs_last_error": (ctypes.c_char_p, [vp]),
        "vkrs_version": (ctypes.c_char_p, []),
        "vkrs_global_invocation_size": (u32, [u32, u32]),
        "vkrs_workgroup_count": (u32, [u32]),
        "vkrs_multi_histograms": (i32, [vp, vp, vp, mpc, vp]),
        "vkrs_multi_scatter": (i32, [vp, vp, vp, vp, mpc, vp, vp, vp]),
        "vkrs_multi_pass": (i32, [vp, vp, vp, vp, mpc, vp]),
        "vkrs_multi_sort": (i32, [vp, vp, vp, vp, mpc, vp]),
        "vkrs_multi_sort_pairs": (i32, [vp, vp, vp, vp, vp, vp, mpc, vp]),
        "vkrs_multi_sort_u64": (i32, [vp, vp, vp, vp, mpc, vp]),
        "vkrs_multi_sort_staged": (i32, [vp, vp, vp, vp, mpc, vp]),
        "vkrs_multi_sort_typed": (i32, [vp, vp, vp, vp, mpc, i32, vp]),
        "vkrs_single_sort": (i32, [vp, vp, vp, spc, vp]),
        "vkrs_sort_auto": (i32, [vp, vp, vp, u32, vp]),
        "vkrs_multi_sort_host": (i32, [vp, vp, u32, vp]),
        "vkrs_key_range": (i32, [vp, vp, u32, vp, vp]),
        "vkrs_partition": (i32, [vp, vp, vp, vp, vp, u32, u32, u32, vp, vp]),
        "vkrs_partition_count": (i32, [vp, vp, u32, u32, u32, i32, vp, vp]),
        "vkrs_partition_scatter_p2p": (i32, [vp, vp, vp, u32, u32, u32, vp, vp, vp]),
        "vkrs_ipc_alloc": (i32, [vp, u64, ctypes.POINTER(vp), ctypes.c_char_p]),
        "vkrs_ipc_open": (i32, [vp, ctypes.c_char_p, ctypes.POINTER(vp)]),
        "vkrs_ipc_close": (i32, [vp, vp]),
        "vkrs_ipc_free": (i32, [vp, vp]),
        "vkrs_check_device_error": (i32, [vp, vp]),
        "vkrs_num_variants": (i32, []),
        "vkrs_variant_name": (ctypes.c_char_p, [i32]),
        "vkrs_set_variant": (i32, [vp, i32]),
        "vkrs_get_variant": (i32, [vp]),
        "vkrs_debug_counters": (i32, [vp, i32, ctypes.POINTER(u64)]),
        "vkrs_set_profiling": (i32, [vp, i32]),
        "vkrs_profile_collect": (i32, [vp]),
        "vkrs_profile_entry": (i32, [vp, i32, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_double),
                                     ctypes.POINTER(u64)]),
        "vkrs_launch_count": (u64, [vp]),
        "vkrs_tile_size": (u32, []),
        "vkrs_set_schedule": (i32, [vp, i32]),
        "vkrs_get_schedule": (i32, [vp]),
        "vkrs_schedule_name": (ctypes.c_char_p, [i32]),
        "vkrs_bucket_stats": (i32, [vp, ctypes.POINTER(u32), vp]),
        "vkrs_debug_bucket_stop": (i32, [vp, i32]),
        "vkrs_set_key_span_hint": (i32, [vp, u32, u32]),
        "vkrs_host_timings": (i32, [vp, ctypes.POINTER(ctypes.c_double)]),
        "vkrs_resolve_schedule": (i32, [vp, u32]),
        "vkrs_exchange_plan": (i32, [vp, vp, u32, u32, vp, vp, vp, vp, u32, u32, vp]),
        "vkrs_peer_barrier": (i32, [vp, vp, vp, u32, u32, u32, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _ptr(x) -> int | None:
    """Device/host address of a torch tensor, numpy array, int or None."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "ctypes"):
        return x.ctypes.data
    raise TypeError(f"cannot take the address of {type(x)}")


def _stream(stream) -> int | None:
    if stream is None:
        try:
            import torch

            if torch.cuda.is_available():
                return torch.cuda.current_stream().cuda_stream or None
        except ImportError:
            pass
        return None
    if isinstance(stream, int):
        return stream or None
    return stream.cuda_stream or None


def global_invocation_size(num_elements: int, nb: int) -> int:
    return int(load().vkrs_global_invocation_size(num_elements, nb))


def workgroup_count(global_invocation_size_: int) -> int:
    return int(load().vkrs_workgroup_count(global_invocation_size_))


def multi_push_constants(num_elements: int, nb: int = 32, shift: int = 0) -> MultiPushConstants:
    """Sizing exactly as MultiRadixSort::execute does it (MultiRadixSort.cpp:12-27)."""
    W = workgroup_count(global_invocation_size(num_elements, nb))
    return MultiPushConstants(num_elements, shift, W, nb)


class Handle:
    """Owns one vkrs_handle (Pass::create / Pass::release)."""

    def __init__(self, device: int = 0, max_num_elements_hint: int = 0):
        self._lib = load()
        h = ctypes.c_void_p()
        st = self._lib.vkrs_create(ctypes.byref(h), device, max_num_elements_hint)
        if st != VKRS_OK:
            raise VkrsError(st, (self._lib.vkrs_last_error(None) or b"").decode())
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._lib.vkrs_destroy(self._h)
            self._h = None

    release = close  # reference name: Pass::release

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, st: int):
        if st != VKRS_OK:
            raise VkrsError(st, (self._lib.vkrs_last_error(self._h) or b"").decode())

    # ---- per stage ----
    def multi_histograms(self, elements_in, histograms, pc: MultiPushConstants, stream=None):
        self._check(self._lib.vkrs_multi_histograms(self._h, _ptr(elements_in), _ptr(histograms), ctypes.byref(pc),
                                                    _stream(stream)))

    def multi_scatter(self, elements_in, elements_out, histograms, pc: MultiPushConstants, values_in=None,
                      values_out=None, stream=None):
        self._check(self._lib.vkrs_multi_scatter(self._h, _ptr(elements_in), _ptr(elements_out), _ptr(histograms),
                                                 ctypes.byref(pc), _ptr(values_in), _ptr(values_out),
                                                 _stream(stream)))

    def multi_pass(self, elements_in, elements_out, histograms, pc: MultiPushConstants, stream=None):
        self._check(self._lib.vkrs_multi_pass(self._h, _ptr(elements_in), _ptr(elements_out), _ptr(histograms),
                                              ctypes.byref(pc), _stream(stream)))

    # ---- whole sorts ----
    def multi_sort(self, buf0, buf1, histograms, pc: MultiPushConstants, stream=None):
        self._check(self._lib.vkrs_multi_sort(self._h, _ptr(buf0), _ptr(buf1), _ptr(histograms), ctypes.byref(pc),
                                              _stream(stream)))

    def multi_sort_pairs(self, keys0, keys1, values0, values1, histograms, pc: MultiPushConstants, stream=None):
        self._check(self._lib.vkrs_multi_sort_pairs(self._h, _ptr(keys0), _ptr(keys1), _ptr(values0), _ptr(values1),
                                                    _ptr(histograms), ctypes.byref(pc), _stream(stream)))

    def multi_sort_u64(self, buf0, buf1, histograms, pc: MultiPushConstants, stream=None):
        self._check(self._lib.vkrs_multi_sort_u64(self._h, _ptr(buf0), _ptr(buf1), _ptr(histograms),
                                                  ctypes.byref(pc), _stream(stream)))

    def multi_sort_typed(self, buf0, buf1, histograms, pc: MultiPushConstants, key_type: int, stream=None):
        """key_type: KEY_U32 / KEY_I32 / KEY_F32."""
        self._check(self._lib.vkrs_multi_sort_typed(self._h, _ptr(buf0), _ptr(buf1), _ptr(histograms), ctypes.byref(pc),
                                                    key_type, _stream(stream)))

    def multi_sort_staged(self, buf0, buf1, histograms, pc: MultiPushConstants, stream=None):
        self._check(self._lib.vkrs_multi_sort_staged(self._h, _ptr(buf0), _ptr(buf1), _ptr(histograms),
                                                     ctypes.byref(pc), _stream(stream)))

    def single_sort(self, buf0, buf1, pc: SinglePushConstants, stream=None):
        self._check(self._lib.vkrs_single_sort(self._h, _ptr(buf0), _ptr(buf1), ctypes.byref(pc), _stream(stream)))

    def sort_auto(self, buf0, buf1, num_elements: int, stream=None):
        self._check(self._lib.vkrs_sort_auto(self._h, _ptr(buf0), _ptr(buf1), num_elements, _stream(stream)))

    def multi_sort_host(self, host_keys, num_elements: int, stream=None):
        self._check(self._lib.vkrs_multi_sort_host(self._h, _ptr(host_keys), num_elements, _stream(stream)))

    def host_timings(self) -> dict:
        """Device ms of the last multi_sort_host call: host-to-device copy, sort, device-to-host copy."""
        out = (ctypes.c_double * 3)()
        self._check(self._lib.vkrs_host_timings(self._h, out))
        return {"h2d_ms": out[0], "sort_ms": out[1], "d2h_ms": out[2]}

    def resolve_schedule(self, num_elements: int) -> int:
        """The schedule `auto` (or the explicitly set one) runs for a keys-only sort of num_elements keys."""
        r = self._lib.vkrs_resolve_schedule(self._h, num_elements)
        if r < 0:
            self._check(r)
        return int(r)

    # ---- multi-GPU partition step ----
    def key_range(self, keys, num_elements: int, min_max_out, stream=None):
        self._check(self._lib.vkrs_key_range(self._h, _ptr(keys), num_elements, _ptr(min_max_out), _stream(stream)))

    def partition(self, keys_in, keys_out, num_elements: int, key_base: int, shift: int, bucket_counts,
                  values_in=None, values_out=None, stream=None):
        self._check(self._lib.vkrs_partition(self._h, _ptr(keys_in), _ptr(keys_out), _ptr(values_in), _ptr(values_out),
                                             num_elements, key_base, shift, _ptr(bucket_counts), _stream(stream)))

    def exchange_plan(self, all_counts, world: int, rank: int, peer_key_ptrs, peer_value_ptrs, dst_tables, summary,
                      capacity: int = 0xFFFFFFFF, max_imbalance_permille: int = 0, stream=None):
        self._check(self._lib.vkrs_exchange_plan(self._h, _ptr(all_counts), world, rank, _ptr(peer_key_ptrs), _ptr(peer_value_ptrs),
                                                 _ptr(dst_tables), _ptr(summary), capacity, max_imbalance_permille, _stream(stream)))

    def peer_barrier(self, flags_local, peer_flag_ptrs, world: int, rank: int, epoch: int, stream=None):
        self._check(self._lib.vkrs_peer_barrier(self._h, _ptr(flags_local), _ptr(peer_flag_ptrs), world, rank, epoch, _stream(stream)))

    def partition_count(self, keys_in, num_elements: int, key_base: int, shift: int, bucket_counts, with_values=False,
                        stream=None):
        self._check(self._lib.vkrs_partition_count(self._h, _ptr(keys_in), num_elements, key_base, shift,
                                                   1 if with_values else 0, _ptr(bucket_counts), _stream(stream)))

    def partition_scatter_p2p(self, keys_in, num_elements: int, key_base: int, shift: int, dst_tables, values_in=None,
                              gate=None, stream=None):
        self._check(self._lib.vkrs_partition_scatter_p2p(self._h, _ptr(keys_in), _ptr(values_in), num_elements, key_base,
                                                         shift, _ptr(dst_tables), _ptr(gate), _stream(stream)))

    def ipc_alloc(self, nbytes: int):
        """-> (device pointer, 64-byte IPC handle) of a cudaMalloc'ed buffer other ranks can map."""
        ptr = ctypes.c_void_p()
        hd = ctypes.create_string_buffer(64)
        self._check(self._lib.vkrs_ipc_alloc(self._h, nbytes, ctypes.byref(ptr), hd))
        return int(ptr.value), hd.raw

    def ipc_open(self, ipc_handle: bytes) -> int:
        ptr = ctypes.c_void_p()
        self._check(self._lib.vkrs_ipc_open(self._h, ipc_handle, ctypes.byref(ptr)))
        return int(ptr.value)

    def ipc_close(self, ptr: int):
        self._check(self._lib.vkrs_ipc_close(self._h, ptr))

    def ipc_free(self, ptr: int):
        self._check(self._lib.vkrs_ipc_free(self._h, ptr))

    # ---- misc ----
    def check_device_error(self, stream=None):
        self._check(self._lib.vkrs_check_device_error(self._h, _stream(stream)))

    def set_variant(self, variant: int):
        self._check(self._lib.vkrs_set_variant(self._h, variant))

    @property
    def variant(self) -> int:
        return int(self._lib.vkrs_get_variant(self._h))

    def set_schedule(self, schedule: int):
        """SCHEDULE_AUTO / SCHEDULE_LSD / SCHEDULE_LSD_UNSTABLE_FIRST / SCHEDULE_BUCKET (keys-only whole sort)."""
        self._check(self._lib.vkrs_set_schedule(self._h, schedule))

    @property
    def schedule(self) -> int:
        return int(self._lib.vkrs_get_schedule(self._h))

    def bucket_stats(self, stream=None) -> dict:
        """Control words of the last bucket-schedule sort (synchronises the stream)."""
        out = (ctypes.c_uint32 * 16)()
        self._check(self._lib.vkrs_bucket_stats(self._h, out, _stream(stream)))
        names = ("shift1", "shift2", "fallback", "recount", "key_min", "max_bucket", "pieces1", "pieces2", "key_max", "_", "base",
                 "big_buckets", "big_items")
        return {k: int(v) for k, v in zip(names, out) if k != "_"}

    def set_key_span_hint(self, lo_key: int = 0, hi_key: int = 0xFFFFFFFF):
        """All keys of the following keys-only sorts lie in [lo_key, hi_key]; defaults = no hint."""
        self._check(self._lib.vkrs_set_key_span_hint(self._h, lo_key, hi_key))

    def debug_bucket_stop(self, stage: int):
        """Test aid: end the bucket schedule after stage 1 / 2 / 3 (0 = whole schedule)."""
        self._check(self._lib.vkrs_debug_bucket_stop(self._h, stage))

    def debug_counters(self, enable: bool) -> list:
        out = (ctypes.c_uint64 * 32)()
        self._check(self._lib.vkrs_debug_counters(self._h, 1 if enable else 0, out))
        return list(out)

    def set_profiling(self, enable: bool):
        self._check(self._lib.vkrs_set_profiling(self._h, 1 if enable else 0))

    def profile(self) -> dict:
        """{kernel name: {"ms": total device ms, "launches": count}} since set_profiling(True)."""
        k = self._lib.vkrs_profile_collect(self._h)
        if k < 0:
            self._check(k)
        out = {}
        for i in range(k):
            name, ms, cnt = ctypes.c_char_p(), ctypes.c_double(), ctypes.c_uint64()
            self._check(self._lib.vkrs_profile_entry(self._h, i, ctypes.byref(name), ctypes.byref(ms),
                                                     ctypes.byref(cnt)))
            out[name.value.decode()] = {"ms": ms.value, "launches": cnt.value}
        return out

    @property
    def launch_count(self) -> int:
        return int(self._lib.vkrs_launch_count(self._h))


def num_variants() -> int:
    return int(load().vkrs_num_variants())


def variant_name(v: int) -> str:
    return (load().vkrs_variant_name(v) or b"").decode()


def schedule_name(s: int) -> str:
    return (load().vkrs_schedule_name(s) or b"").decode()


def tile_size() -> int:
    return int(load().vkrs_tile_size())


def version() -> str:
    return load().vkrs_version().decode()
