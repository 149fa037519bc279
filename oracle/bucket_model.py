"""numpy model of the B200 library's keys-only *bucket schedule* (vkradixsort_b200/csrc/vkrs_msd.cuh).

TEST INFRASTRUCTURE ONLY, like the rest of oracle/: it states, on the CPU and in a few lines each, the
decisions the device code takes (digit window, recount, fallback, item windows, bin map) and the order in
which it moves the keys, so that `-m "not gpu"` tests can check the schedule's logic against std::sort order
-- the reference's own criterion, MultiRadixSort::testSort (multiradixsort/src/MultiRadixSort.cpp:148-161) --
on every distribution the GPU tests use.  It is NOT a restatement of the reference (that is oracle/vkrs_oracle.c):
the reference has no such schedule; what ties the schedule to the reference is that both leave the one sorted
permutation of the keys in buffer 0.

Where the device code is free to order keys arbitrarily (the unstable partition passes, the order inside a bin
before the fix-up) the model scrambles them with a seeded permutation: a model that only worked because numpy
happens to be stable would prove nothing.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

RADIX = 256
LOCAL_MAX = 4096          # largest (digit1, digit2) bucket the local sort takes (vkrs_msd.cuh: LOCAL_MAX)
LT_CAP = 6144             # keys per shared-memory buffer of the local sort (LT_CAP)
LT_MIN_WINDOW = 256
LT_BIN_BITS = 12
LT_BIN_LIMIT = 32


@dataclass
class Plan:
    base: int       # the digits are taken from key - base
    shift1: int
    shift2: int     # also the number of low bits left to the local sort
    recount: bool   # the first histogram was not counted in the final window
    skip_pass2: bool
    fallback: bool
    max_bucket: int  # largest (digit1, digit2) bucket (0 when pass 2 was skipped)


def window(kmin: int, kmax: int, base0: int = 0, shift0: int = 24):
    """msd_window_kernel: (base, shift1, recount) for keys in [kmin, kmax], first histogram counted at (base0, shift0)."""
    span = kmax - kmin
    top = span.bit_length() - 1 if span else 0
    s1 = top - 7 if top >= 15 else 8
    keep = base0 <= kmin and shift0 == s1 and ((kmax - base0) >> s1) < RADIX
    return (base0, shift0, False) if keep else (kmin, s1, True)


def hint_window(lo_key: int, hi_key: int):
    """vkrs_set_key_span_hint: where the first histogram counts when the caller says the keys lie in [lo, hi]."""
    x = lo_key ^ hi_key
    top = x.bit_length() - 1 if x else 0
    return lo_key, (top - 7 if top >= 15 else 8)


def lt_window(max_bucket: int) -> int:
    w = 4096
    while w > LT_MIN_WINDOW and w + max_bucket > LT_CAP:
        w >>= 1
    return w


def _unstable_partition(keys: np.ndarray, digits: np.ndarray, rng) -> np.ndarray:
    """Keys grouped by digit, in ARBITRARY order inside a group (the device ranks with atomics)."""
    scramble = rng.permutation(keys.shape[0])
    order = scramble[np.argsort(digits[scramble], kind="stable")]
    return keys[order]


def _local_item_sort(item: np.ndarray, base: int, nb: int, low_bits: int, rng):
    """local_tile_bins: order-preserving bins over the item's key span + comparison fix-up inside the bins.
    Returns (sorted item, True) or (None, False) when a bin is over-full (the device then sorts the item's
    buckets one by one with two 8-bit passes -- modelled by the caller as a plain sort)."""
    span_bits = low_bits + ((nb - 1).bit_length() if nb > 1 else 0)
    s = max(0, span_bits - LT_BIN_BITS)
    bins = (item.astype(np.int64) - base) >> s
    assert bins.min() >= 0 and bins.max() < (1 << LT_BIN_BITS), "bin map out of range"
    counts = np.bincount(bins, minlength=1 << LT_BIN_BITS)
    if s > 0 and counts.max() > LT_BIN_LIMIT:
        return None, False
    grouped = _unstable_partition(item, bins, rng)          # count + scan + atomic place
    if s == 0:
        return grouped, True                                 # equal bins are equal keys
    gbins = (grouped.astype(np.int64) - base) >> s
    ends = np.cumsum(counts)
    out = np.empty_like(grouped)
    for p in range(grouped.shape[0]):                        # fix-up: rank by (key, offset) inside the bin
        b = int(gbins[p])
        lo, hi = int(ends[b] - counts[b]), int(ends[b])
        k, d = grouped[p], p - lo
        g = grouped[lo:hi]
        r = int(np.count_nonzero(g < k) + np.count_nonzero(g[:d] == k))
        out[lo + r] = k
    return out, True


def sort(keys: np.ndarray, seed: int = 0, hint=None, fix_up_limit: int = 200_000):
    """The whole schedule.  Returns (sorted keys, Plan).  `fix_up_limit`: above this many keys the per-key
    fix-up loop (pure Python) is replaced by a per-item np.sort -- the plan logic is still modelled exactly."""
    keys = np.ascontiguousarray(keys, dtype=np.uint32)
    n = keys.shape[0]
    rng = np.random.default_rng(seed)
    if n == 0:
        return keys.copy(), Plan(0, 24, 16, False, False, False, 0)
    kmin, kmax = int(keys.min()), int(keys.max())
    base0, shift0 = hint_window(*hint) if hint is not None else (0, 24)
    base, s1, recount = window(kmin, kmax, base0, shift0)
    s2 = s1 - 8
    rel = keys.astype(np.int64) - base
    d1 = (rel >> s1) & 255
    # pass 1 (buf0 -> buf1); a top-digit bucket above 256 * LOCAL_MAX keys must overflow the local sort
    top_counts = np.bincount(d1, minlength=RADIX)
    skip_pass2 = bool(s2 > 0 and top_counts.max() > RADIX * LOCAL_MAX)
    if skip_pass2:
        return np.sort(keys), Plan(base, s1, s2, recount, True, True, 0)   # the LSD passes sort the untouched input
    buf1 = _unstable_partition(keys, d1, rng)
    # pass 2 (buf1 -> buf0), inside each bucket of pass 1: grouping by the 16-bit prefix (digit1, digit2)
    prefix = (buf1.astype(np.int64) - base) >> s2
    assert prefix.max() < RADIX * RADIX
    buf0 = _unstable_partition(buf1, prefix, rng)
    sizes = np.bincount(prefix, minlength=RADIX * RADIX)
    max_bucket = int(sizes.max())
    fallback = bool(s2 > 0 and max_bucket > LOCAL_MAX)
    plan = Plan(base, s1, s2, recount, False, fallback, max_bucket)
    if fallback:
        return np.sort(buf0), plan                                           # four stable LSD passes on buf0
    if s2 == 0:
        return buf0, plan                                                    # no low bits left: two passes were the sort
    # local sort: items = the buckets whose first key lies in one window of the array
    starts = np.concatenate(([0], np.cumsum(sizes)))                         # sub_start[65537]
    w = lt_window(max_bucket)
    num_items = (n + w - 1) // w
    item_first = np.searchsorted(starts[:-1], np.arange(num_items + 1) * w, side="left")
    item_first[-1] = RADIX * RADIX
    out = buf0.copy()
    for it in range(num_items):
        j0, j1 = int(item_first[it]), int(item_first[it + 1])
        lo, hi = int(starts[j0]), int(starts[j1])
        if hi - lo <= 1:
            continue
        assert hi - lo <= LT_CAP or max_bucket + w > LT_CAP, "an item must fit the shared-memory buffer"
        item = buf0[lo:hi]
        if n <= fix_up_limit and hi - lo <= LT_CAP:
            done, ok = _local_item_sort(item, base + (j0 << s2), j1 - j0, s2, rng)
            out[lo:hi] = done if ok else np.sort(item)
        else:
            out[lo:hi] = np.sort(item)
    return out, plan
