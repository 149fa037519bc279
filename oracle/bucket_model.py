"""numpy model of the B200 library's keys-only *bucket schedule* (vkradixsort_b200/csrc/vkrs_msd.cuh).

TEST INFRASTRUCTURE ONLY, like the rest of oracle/: it states, on the CPU and in a few lines each, the
decisions the device code takes (digit window, recount, big buckets and fallback, item windows, bin map) and the order in
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
LT_CAP = 7680             # keys per shared-memory buffer of the local sort (LT_CAP)
LT_MIN_WINDOW = 256
LT_BIN_BITS = 12
LT_BINS = 1 << LT_BIN_BITS
LT_BIN_LIMIT = 32
BIG_MAX = 1024            # big buckets (above LOCAL_MAX keys) one sort finishes by counting ...
BIG_POOL_WORDS = 16 << 20  # ... as long as their 2^low_bits counters fit this pool


@dataclass
class Plan:
    base: int       # the digits are taken from key - base
    shift1: int
    shift2: int     # also the number of low bits left to the local sort
    recount: bool   # the first histogram was not counted in the final window
    fallback: bool  # more big buckets than the counting path takes: the stable LSD passes sort the array
    max_bucket: int  # largest (digit1, digit2) bucket the shared-memory sort sees (big buckets excluded)
    big_buckets: int = 0  # buckets above LOCAL_MAX keys: sorted by counting their low bits
    redo_items: int = 0   # items the bins path refused (over-full bin, no room): sorted bucket by bucket


def window(kmin: int, kmax: int, base0: int = 0, shift0: int = 24):
    """msd_window_kernel: (base, shift1, recount) for keys in [kmin, kmax], first histogram counted at (base0, shift0)."""
    span = kmax - kmin
    top = span.bit_length() - 1 if span else 0
    s1 = top - 7 if top >= 15 else 8
    keep = base0 <= kmin and shift0 == s1 and ((kmax - base0) >> s1) < RADIX
    return (base0, shift0, False) if keep else (kmin, s1, True)


def hint_window(lo_key: int, hi_key: int):
    """vkrs_set_key_span_hint: where the first histogram counts when the caller says the keys lie in [lo, hi]."""
    x = hi_key - lo_key  # the span, as msd_window_kernel takes it from (largest - smallest key)
    top = x.bit_length() - 1 if x else 0
    return lo_key, (top - 7 if top >= 15 else 8)


def lt_window(max_bucket: int) -> int:
    """The largest multiple of 512 with window + max_bucket <= LT_CAP - 4 (an item then fits a buffer), at least LT_MIN_WINDOW."""
    room = LT_CAP - 4
    if max_bucket + LT_MIN_WINDOW >= room:
        return LT_MIN_WINDOW
    return min(16384, max(LT_MIN_WINDOW, (room - max_bucket) & ~511))


def bin_mult(num_buckets: int, low_bits: int) -> int:
    """lt_bin_mult: 0 = exact mode (the span fits the bins), else about floor(LT_BINS * 2^32 / span), never above it
    (the device rounds a single-precision quotient down: modelled with the same arithmetic)."""
    span = num_buckets << low_bits
    if span <= LT_BINS:
        return 0
    q = np.float32(np.float32(LT_BINS * 4294967296.0) / np.float32(span)) * np.float32(1.0 - 1.0 / 2097152.0)
    return int(np.float32(q))


def _unstable_partition(keys: np.ndarray, digits: np.ndarray, rng) -> np.ndarray:
    """Keys grouped by digit, in ARBITRARY order inside a group (the device ranks with atomics)."""
    scramble = rng.permutation(keys.shape[0])
    order = scramble[np.argsort(digits[scramble], kind="stable")]
    return keys[order]


def _local_item_sort(item: np.ndarray, base: int, nb: int, low_bits: int, rng):
    """local_tile_bins: order-preserving bins over the item's key span, bin = umulhi(key - base, mult), + comparison
    fix-up inside the bins.  Returns (sorted item, True) or (None, False) when a bin is over-full (the item then goes
    to msd_local_redo_kernel: its buckets one by one with two 8-bit passes -- modelled by the caller as a plain sort)."""
    mult = bin_mult(nb, low_bits)
    rel = item.astype(np.int64) - base
    bins = rel if mult == 0 else (rel * mult) >> 32
    assert bins.min() >= 0 and bins.max() < LT_BINS, "bin map out of range"
    assert np.all(np.diff(bins[np.argsort(rel, kind="stable")]) >= 0), "bin map not monotone"
    counts = np.bincount(bins, minlength=LT_BINS)
    if mult != 0 and counts.max() >= LT_BIN_LIMIT:
        return None, False
    grouped = _unstable_partition(item, bins, rng)          # count + scan + atomic place
    if mult == 0:
        return grouped, True                                 # equal bins are equal keys
    grel = grouped.astype(np.int64) - base
    gbins = (grel * mult) >> 32
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


def _counting_sort(bucket: np.ndarray, value_base: int, low_bits: int) -> np.ndarray:
    """msd_big_hist_kernel + big_scan_bucket + msd_big_fill_kernel: histogram of the low bits, exclusive scan, every
    value written as often as it was counted -- no key moves."""
    low = (bucket.astype(np.int64) - value_base) & ((1 << low_bits) - 1)
    assert np.array_equal(low + value_base, bucket.astype(np.int64)), "a bucket's keys agree above the low bits"
    counts = np.bincount(low, minlength=1 << low_bits)
    return (np.repeat(np.arange(1 << low_bits, dtype=np.int64), counts) + value_base).astype(np.uint32)


GUESS_SAMPLES = 16384  # MSD_GUESS_SAMPLES


def guess_window(keys: np.ndarray):
    """msd_init_kernel / msd_guess_window(): without a key-span hint the first histogram counts in the digit window
    of 16384 sample keys (16 blocks of 1024 consecutive keys, evenly spread, both ends of the array included) -- pushed down by an eighth of a top-level bucket, or to 0 -- unless the largest
    sample would not fit it.  Returns (base0, shift0)."""
    n = keys.shape[0]
    t = np.arange(GUESS_SAMPLES, dtype=np.int64) % 1024          # thread
    i = np.arange(GUESS_SAMPLES, dtype=np.int64) // 1024         # block of 1024 consecutive keys, 16 of them
    idx = np.minimum(t, n - 1)
    if n >= 1024:
        idx = idx + (i * (n - 1024)) // 15
    smp = keys[idx]
    smin, smax = int(smp.min()), int(smp.max())
    span = smax - smin
    top = span.bit_length() - 1 if span else 0
    s1 = top - 7 if top >= 15 else 8
    margin = 1 << (s1 - 3)
    b0 = smin - margin if smin > margin else 0
    if ((smax - b0) >> s1) >= RADIX:
        return 0, 24
    return b0, s1


def sort(keys: np.ndarray, seed: int = 0, hint=None, fix_up_limit: int = 200_000, guess: bool = True):
    """The whole schedule.  Returns (sorted keys, Plan).  `fix_up_limit`: above this many keys the per-key
    fix-up loop (pure Python) is replaced by a per-item np.sort -- the plan logic is still modelled exactly.
    `guess` = the sampled guess of the digit window is on (the library's default; VKRS_GUESS_WINDOW=0 turns it off)."""
    keys = np.ascontiguousarray(keys, dtype=np.uint32)
    n = keys.shape[0]
    rng = np.random.default_rng(seed)
    if n == 0:
        return keys.copy(), Plan(0, 24, 16, False, False, 0)
    kmin, kmax = int(keys.min()), int(keys.max())
    base0, shift0 = hint_window(*hint) if hint is not None else (guess_window(keys) if guess else (0, 24))
    base, s1, recount = window(kmin, kmax, base0, shift0)
    s2 = s1 - 8
    rel = keys.astype(np.int64) - base
    d1 = (rel >> s1) & 255
    # pass 1 (buf0 -> buf1)
    buf1 = _unstable_partition(keys, d1, rng)
    # pass 2 (buf1 -> buf0), inside each bucket of pass 1: grouping by the 16-bit prefix (digit1, digit2)
    prefix = (buf1.astype(np.int64) - base) >> s2
    assert prefix.max() < RADIX * RADIX
    buf0 = _unstable_partition(buf1, prefix, rng)
    sizes = np.bincount(prefix, minlength=RADIX * RADIX)
    if s2 == 0:
        return buf0, Plan(base, s1, s2, recount, False, int(sizes.max()))    # no low bits left: two passes were the sort
    # buckets above LOCAL_MAX keys are "big": sorted by counting, unless there are more than the counter pool takes
    big = np.flatnonzero(sizes > LOCAL_MAX)
    limit = min(BIG_MAX, BIG_POOL_WORDS >> s2)
    max_bucket = int(sizes[sizes <= LOCAL_MAX].max())
    if big.shape[0] > limit:
        return np.sort(buf0), Plan(base, s1, s2, recount, True, max_bucket, int(big.shape[0]))  # four stable LSD passes on buf0
    plan = Plan(base, s1, s2, recount, False, max_bucket, int(big.shape[0]))
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
        item = buf0[lo:hi]
        if hi - lo <= LT_CAP - 4:
            if n <= fix_up_limit:
                done, ok = _local_item_sort(item, base + (j0 << s2), j1 - j0, s2, rng)
            else:
                done, ok = np.sort(item), True
            if ok:
                out[lo:hi] = done
                continue
        # msd_local_redo_kernel: the item's buckets one by one; big buckets are left alone
        plan.redo_items += 1
        assert hi - lo <= LT_CAP - 4 or any(sizes[j] > LOCAL_MAX for j in range(j0, j1)), "only a big bucket makes an item too large"
        for j in range(j0, j1):
            if 1 < sizes[j] <= LOCAL_MAX:
                out[starts[j]:starts[j + 1]] = np.sort(buf0[starts[j]:starts[j + 1]])
    for j in big:                                                            # the counting path
        out[starts[j]:starts[j + 1]] = _counting_sort(buf0[starts[j]:starts[j + 1]], base + (int(j) << s2), s2)
    return out, plan
