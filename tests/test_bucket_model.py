"""CPU tests of the bucket schedule's logic through its numpy model (oracle/bucket_model.py): the digit window on
the occupied key range, the recount rule, the big-bucket (counting) and fallback rules, the item windows and the bin map
+ fix-up of the local sort -- against std::sort order (MultiRadixSort::testSort, multiradixsort/src/MultiRadixSort.cpp:148-161).
The same inputs and the same expected control words are asserted on the device in tests/test_gpu_bucket.py."""
import numpy as np
import pytest

from oracle import bucket_model as M


def distributions(oracle, n, seed):
    ar = np.arange(n, dtype=np.uint64)
    rng = np.random.default_rng(seed)
    return {
        "uniform32": oracle.generate_random(n, seed, 0xFFFFFFFF),
        "reference28": oracle.generate_random(n, seed + 1, 0x0FFFFFFF),  # MultiRadixSort.cpp:126
        "bits20": oracle.generate_random(n, seed + 2, 0x000FFFFF),
        "bits12": oracle.generate_random(n, seed + 3, 0x00000FFF),
        "sorted": (ar * 4294967295 // max(n, 1)).astype(np.uint32),
        "all_equal": np.full(n, 0xDEADBEEF, dtype=np.uint32),
        "two_valued": ((ar % 2) * 0xFFFFFFFF).astype(np.uint32),
        "hot_prefix": (np.where(rng.random(n) < 0.5, np.uint32(0xABCD0000), 0).astype(np.uint32)
                       | rng.integers(0, 1 << 16, n, dtype=np.uint32)).astype(np.uint32),
        "around_2^31": (np.uint32(0x7FFFF000) + rng.integers(0, 0x2000, n, dtype=np.uint32)).astype(np.uint32),
        "offset_range": (np.uint32(123456789) + rng.integers(0, 3_000_000, n, dtype=np.uint32)).astype(np.uint32),
        "clustered_low_bits": ((rng.integers(0, 1 << 14, n, dtype=np.uint32) << 18) | rng.integers(0, 4, n, dtype=np.uint32)).astype(np.uint32),
    }


@pytest.mark.parametrize("n", [1, 2, 255, 6145, 70_001])
def test_model_sorts_every_distribution(oracle, n):
    for name, keys in distributions(oracle, n, 4242 + n).items():
        out, plan = M.sort(keys, seed=n)
        assert oracle.test_sort(np.sort(keys), out) == -1, (n, name, plan)
        assert plan.shift2 == plan.shift1 - 8 and 8 <= plan.shift1 <= 24, (name, plan)
        assert int(keys.min()) >= plan.base, (name, plan)
        assert ((int(keys.max()) - plan.base) >> plan.shift1) < 256, (name, plan)


def test_digit_window_and_recount(oracle):
    n = 100_003
    # full-range keys: counted at (0, 24) from the start, nothing to redo
    _, p = M.sort(oracle.generate_random(n, 1, 0xFFFFFFFF))
    assert (p.base, p.shift1, p.shift2, p.recount, p.fallback, p.big_buckets) == (0, 24, 16, False, False, 0)
    # the reference's 28-bit keys: the window moves under bit 27.  The sampled guess of the window (16384 evenly spread
    # keys) finds it before the first histogram; without the guess the first histogram is counted again
    keys = oracle.generate_random(n, 2, 0x0FFFFFFF)
    _, p = M.sort(keys)
    assert (p.shift1, p.shift2, p.recount, p.base) == (20, 12, False, 0)
    _, p = M.sort(keys, guess=False)
    assert (p.shift1, p.shift2, p.recount, p.base) == (20, 12, True, int(keys.min()))
    assert M.guess_window(keys) == (0, 20) and M.guess_window(oracle.generate_random(n, 1, 0xFFFFFFFF)) == (0, 24)
    # a key range that does not start near 0 and nearly fills a power of two: no base the samples suggest can hold the
    # largest key -- the guess is dropped, the recount does the job
    assert M.guess_window((np.uint32(0xC0000000) | (oracle.generate_random(n, 9, 0xFFFFFFFF) >> np.uint32(2))).astype(np.uint32)) == (0, 24)
    # small signed integers after the sign-flip map sit around 2^31: guessed
    ints = (np.random.default_rng(3).integers(-5000, 5001, n).astype(np.int32).view(np.uint32) ^ np.uint32(0x80000000))
    b0, s0 = M.guess_window(ints)
    assert s0 == 8 and b0 <= int(ints.min()) and ((int(ints.max()) - b0) >> s0) < 256
    # one rank's key range after the multi-GPU exchange: shared top bits
    keys = (np.uint32(0xC0000000) | (oracle.generate_random(n, 9, 0xFFFFFFFF) >> np.uint32(2))).astype(np.uint32)
    _, p = M.sort(keys)
    assert (p.shift1, p.recount) == (22, True)
    _, p = M.sort(keys, hint=(0xC0000000, 0xFFFFFFFF))   # the exchange tells the local sort its key range
    assert (p.shift1, p.recount, p.base) == (22, False, 0xC0000000)
    _, p = M.sort(keys, hint=(0, 0xFFFF))                # a wrong hint only costs the recount
    assert (p.shift1, p.recount) == (22, True)
    # 16 varying bits: two passes are the whole sort, no local sort, never a fallback
    keys = (np.uint32(0x12340000) | np.random.default_rng(7).integers(0, 1 << 16, 300_000, dtype=np.uint32)).astype(np.uint32)
    out, p = M.sort(keys)
    assert (p.shift1, p.shift2, p.fallback) == (8, 0, False) and np.array_equal(out, np.sort(keys))
    # small signed integers around zero after the sign-flip map of vkrs_multi_sort_typed
    ints = np.random.default_rng(8).integers(-100, 101, 50_000).astype(np.int32)
    mapped = (ints.view(np.uint32) ^ np.uint32(0x80000000))
    out, p = M.sort(mapped)
    assert (p.shift2, p.fallback) == (0, False) and np.array_equal(out, np.sort(mapped))


def test_big_bucket_and_fallback_rules(oracle):
    rng = np.random.default_rng(7)
    n = 300_000
    # four 16-bit-prefix buckets of 75,000 keys: too large for shared memory, sorted by counting -- no fallback
    keys = ((rng.integers(0, 4, n, dtype=np.uint32) << 30) | rng.integers(0, 1 << 16, n, dtype=np.uint32)).astype(np.uint32)
    out, p = M.sort(keys)
    assert not p.fallback and p.big_buckets == 4 and p.shift1 == 24, p
    assert np.array_equal(out, np.sort(keys))
    # half of the keys under one prefix, the rest uniform: one big bucket, its neighbours' item goes bucket by bucket
    m = 400_000
    keys = np.where(rng.random(m) < 0.5, np.uint32(0x2BCD0000) | rng.integers(0, 1 << 16, m, dtype=np.uint32),
                    rng.integers(0, 1 << 32, m, dtype=np.uint64).astype(np.uint32)).astype(np.uint32)
    out, p = M.sort(keys, fix_up_limit=0)
    assert not p.fallback and p.big_buckets == 1 and p.redo_items >= 1 and p.max_bucket < 100, p
    assert np.array_equal(out, np.sort(keys))
    # 512 buckets of ~5,900 keys: more than the 256 the counter pool takes at 16 low bits -> the LSD passes
    m = 3_000_000
    keys = ((rng.integers(0, 2, m, dtype=np.uint32) * np.uint32(0xFF000000)) | rng.integers(0, 1 << 24, m, dtype=np.uint32)).astype(np.uint32)
    out, p = M.sort(keys)
    assert p.fallback and p.big_buckets == 512 and p.shift1 == 24, p
    assert np.array_equal(out, np.sort(keys))
    # uniform keys never get there below 2.2e8 keys: the largest of 65536 buckets stays far below 4096
    _, p = M.sort(oracle.generate_random(2_000_003, 5, 0xFFFFFFFF))
    assert not p.fallback and p.big_buckets == 0 and p.max_bucket < 100


def test_item_windows_and_bin_map():
    assert M.lt_window(0) == 7676 & ~511 == 7168 and M.lt_window(1716) == 5632 and M.lt_window(2044) == 5632
    assert M.lt_window(2045) == 5120 and M.lt_window(4096) == 3072 and M.lt_window(5000) == 2560
    assert M.lt_window(7000) == 512 and M.lt_window(7500) == 256  # the last cannot fit: such items go to the per-bucket path
    assert M.hint_window(0, 0xFFFFFFFF) == (0, 24) and M.hint_window(0xC0000000, 0xFFFFFFFF) == (0xC0000000, 22)
    assert M.hint_window(0x60000000, 0x9FFFFFFF) == (0x60000000, 22)  # the span decides, not the differing bits
    # the multiplicative bin map uses all bins whatever the span, stays below LT_BINS and is monotone
    for nb, low_bits in ((1, 16), (3, 16), (5, 16), (700, 8), (65536, 16), (1, 13), (2, 12)):
        span = nb << low_bits
        mult = M.bin_mult(nb, low_bits)
        x = np.unique(np.concatenate([np.arange(0, min(span, 5000)), np.linspace(0, span - 1, 5000).astype(np.int64), [span - 1]]))
        bins = (x * mult) >> 32
        assert bins.max() < M.LT_BINS and np.all(np.diff(bins) >= 0)
        assert bins.max() >= M.LT_BINS - 2, (nb, low_bits, int(bins.max()))
    assert M.bin_mult(1, 12) == 0 and M.bin_mult(16, 8) == 0  # the span fits the bins: exact mode


def test_model_property_random_masks_and_offsets():
    """Random sizes, bit masks, offsets and duplicate levels: the model's output is the sorted input and its
    digit window always covers the occupied key range."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=120, deadline=None)
    @given(n=st.integers(1, 3000), mask_bits=st.integers(0, 32), offset=st.integers(0, 0xFFFFFFFF), distinct=st.integers(1, 5000),
           seed=st.integers(0, 2**31))
    def check(n, mask_bits, offset, distinct, seed):
        rng = np.random.default_rng(seed)
        mask = (1 << mask_bits) - 1
        pool = (rng.integers(0, 1 << 32, size=distinct, dtype=np.uint64) & np.uint64(mask))
        keys = ((pool[rng.integers(0, distinct, size=n)] + np.uint64(offset)) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        out, plan = M.sort(keys, seed=seed)
        assert np.array_equal(out, np.sort(keys)), plan
        assert plan.base <= int(keys.min()) and ((int(keys.max()) - plan.base) >> plan.shift1) < 256

    check()
