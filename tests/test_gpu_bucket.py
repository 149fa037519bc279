"""GPU parity tests of the keys-only schedules that rank without stability (run with -m gpu):
the bucket schedule (two unstable top-digit partition passes + shared-memory sort of every
16-bit-prefix bucket, vkrs_msd.cuh) and the LSD schedule with an unstable first pass.

Bar: bit-exact against std::sort / numpy, the reference's own criterion
(multiradixsort/src/MultiRadixSort.cpp:148-161); the intermediate stages are checked through the
properties that define them (a permutation of the input, grouped by the partition digits).
Nothing here reads /root/reference.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    return torch.device("cuda:0")


@pytest.fixture()
def handle(built_lib):
    from vkradixsort_b200 import Handle

    h = Handle(0, 1 << 20)
    yield h
    h.close()


def to_dev(a, dev):
    return torch.from_numpy(a.view(np.int32)).to(dev)


def to_host(t):
    return t.cpu().numpy().view(np.uint32)


def run_sort(handle, keys, dev, schedule, stop=0):
    from vkradixsort_b200 import capi

    n = keys.shape[0]
    b0 = to_dev(keys, dev)
    b1 = torch.full_like(b0, 0x5A5A5A5A)  # poisoned scratch
    handle.set_schedule(schedule)
    handle.debug_bucket_stop(stop)
    handle.multi_sort(b0, b1, None, capi.multi_push_constants(n, 32))
    handle.check_device_error()
    handle.debug_bucket_stop(0)
    return to_host(b0), to_host(b1)


def distributions(oracle, n, seed):
    ar = np.arange(n, dtype=np.uint64)
    rng = np.random.default_rng(seed)
    return {
        "uniform32": oracle.generate_random(n, seed, 0xFFFFFFFF),
        "reference28": oracle.generate_random(n, seed + 1, 0x0FFFFFFF),  # MultiRadixSort.cpp:126
        "bits20": oracle.generate_random(n, seed + 2, 0x000FFFFF),
        "bits12": oracle.generate_random(n, seed + 3, 0x00000FFF),
        "bits5": oracle.generate_random(n, seed + 4, 31),
        "sorted": (ar * 4294967295 // max(n, 1)).astype(np.uint32),
        "descending": (0xFFFFFFFF - ar * 4294967295 // max(n, 1)).astype(np.uint32),
        "all_equal": np.full(n, 0xDEADBEEF, dtype=np.uint32),
        "zeros": np.zeros(n, dtype=np.uint32),
        "max": np.full(n, 0xFFFFFFFF, dtype=np.uint32),
        "two_valued": ((ar % 2) * 0xFFFFFFFF).astype(np.uint32),
        "hot_prefix": np.where(rng.random(n) < 0.5, np.uint32(0xABCD0000), 0).astype(np.uint32)
        | rng.integers(0, 1 << 16, n, dtype=np.uint32),  # half of the keys in one 16-bit-prefix bucket
        "low16_only_high_random": (rng.integers(0, 1 << 16, n, dtype=np.uint32) << 16).astype(np.uint32),
        "around_2^31": (np.uint32(0x7FFFF000) + rng.integers(0, 0x2000, n, dtype=np.uint32)).astype(np.uint32),  # narrow range across a bit boundary
        "offset_range": (np.uint32(123456789) + rng.integers(0, 3_000_000, n, dtype=np.uint32)).astype(np.uint32),
    }


BUCKET_SIZES = [1, 2, 33, 255, 4096, 6143, 6144, 6145, 12289, 100_003, 1_000_001, 5_000_011]


@pytest.mark.parametrize("n", BUCKET_SIZES)
def test_bucket_schedule_matches_std_sort(handle, dev, oracle, n):
    from vkradixsort_b200 import capi

    for name, keys in distributions(oracle, n, 4242 + n).items():
        keys = np.ascontiguousarray(keys, dtype=np.uint32)
        out, _ = run_sort(handle, keys, dev, capi.SCHEDULE_BUCKET)
        assert oracle.test_sort(np.sort(keys), out) == -1, f"n={n} {name} {handle.bucket_stats()}"


@pytest.mark.parametrize("n", [1, 6145, 100_003, 1 << 20, 3_000_001])
def test_lsd_unstable_first_pass_matches_std_sort(handle, dev, oracle, n):
    from vkradixsort_b200 import capi

    for name, keys in distributions(oracle, n, 99 + n).items():
        keys = np.ascontiguousarray(keys, dtype=np.uint32)
        out, _ = run_sort(handle, keys, dev, capi.SCHEDULE_LSD_UNSTABLE_FIRST)
        assert oracle.test_sort(np.sort(keys), out) == -1, f"n={n} {name}"


def test_bucket_stages(handle, dev, oracle):
    """Stage by stage: pass 1 leaves a permutation of the input grouped by the top digit in buffer 1,
    pass 2 one grouped by the top two digits in buffer 0, the local sort finishes it."""
    from vkradixsort_b200 import capi

    n = 2_000_003
    for name, max_value, s1 in (("uniform32", 0xFFFFFFFF, 24), ("reference28", 0x0FFFFFFF, 20)):
        keys = oracle.generate_random(n, 555, max_value)
        _, b1 = run_sort(handle, keys, dev, capi.SCHEDULE_BUCKET, stop=1)
        st = handle.bucket_stats()
        assert st["shift1"] == s1 and st["shift2"] == s1 - 8, (name, st)
        assert st["recount"] == 0, (name, st)  # 28-bit keys: the sampled guess of the digit window saves the recount
        base = np.uint32(st["key_min"] if st["recount"] else 0)  # the digits are taken from key - base (vkrs_msd.cuh)
        d1 = ((b1 - base) >> np.uint32(s1)) & np.uint32(255)
        assert np.all(np.diff(d1.astype(np.int64)) >= 0), f"{name}: pass 1 output is not grouped by the top digit"
        assert np.array_equal(np.sort(b1), np.sort(keys)), f"{name}: pass 1 output is not a permutation of the input"
        b0, _ = run_sort(handle, keys, dev, capi.SCHEDULE_BUCKET, stop=2)
        d12 = ((b0 - base) >> np.uint32(s1 - 8)).astype(np.int64)
        assert np.all(np.diff(d12) >= 0), f"{name}: pass 2 output is not grouped by the top two digits"
        assert np.array_equal(np.sort(b0), np.sort(keys)), f"{name}: pass 2 output is not a permutation of the input"
        b0, _ = run_sort(handle, keys, dev, capi.SCHEDULE_BUCKET, stop=3)
        assert handle.bucket_stats()["fallback"] == 0, name
        assert oracle.test_sort(np.sort(keys), b0) == -1, f"{name}: local sort"


def test_digit_window_guess_and_recount(dev, oracle, built_lib, monkeypatch):
    """Without a key-span hint the first histogram counts in the digit window of 16384 sample keys (msd_init_kernel);
    with the guess off (VKRS_GUESS_WINDOW=0) or wrong, msd_window_kernel moves the window and the histogram is recounted.
    The control words match the numpy model of the same decisions (oracle/bucket_model.py)."""
    from oracle import bucket_model as M
    from vkradixsort_b200 import Handle, capi

    n = 300_001
    rng = np.random.default_rng(21)
    cases = {
        "uniform32": oracle.generate_random(n, 1, 0xFFFFFFFF),
        "reference28": oracle.generate_random(n, 2, 0x0FFFFFFF),
        "bits20": oracle.generate_random(n, 3, 0x000FFFFF),
        "rank_range": (np.uint32(0xC0000000) | (oracle.generate_random(n, 9, 0xFFFFFFFF) >> np.uint32(2))).astype(np.uint32),
        "offset_range": (np.uint32(123456789) + rng.integers(0, 3_000_000, n, dtype=np.uint32)).astype(np.uint32),
        "around_2^31": (np.uint32(0x7FFFF000) + rng.integers(0, 0x2000, n, dtype=np.uint32)).astype(np.uint32),
        "all_equal": np.full(n, 0xDEADBEEF, dtype=np.uint32),
        # the samples miss the outliers: the guess is wrong, the recount repairs it
        "outliers": np.concatenate([oracle.generate_random(n - 2, 4, 0x00FFFFFF), np.array([0xFFFFFFF0, 0x80000000], dtype=np.uint32)]),
    }
    for guess in (True, False):
        monkeypatch.setenv("VKRS_GUESS_WINDOW", "1" if guess else "0")
        h = Handle(0, n)
        try:
            for name, keys in cases.items():
                out, _ = run_sort(h, keys, dev, capi.SCHEDULE_BUCKET)
                st = h.bucket_stats()
                assert np.array_equal(out, np.sort(keys)), (name, guess)
                _, p = M.sort(keys, guess=guess, fix_up_limit=0)
                assert (st["shift1"], st["shift2"], st["recount"], st["base"]) == (p.shift1, p.shift2, int(p.recount), p.base), (name, guess, st, p)
        finally:
            h.close()
    keys = cases["reference28"]
    assert M.sort(keys, guess=True, fix_up_limit=0)[1].recount is False and M.sort(keys, guess=False, fix_up_limit=0)[1].recount is True


def test_big_buckets_are_counted_not_sorted(handle, dev, oracle):
    """A 16-bit-prefix bucket above 4096 keys cannot be finished in shared memory.  All its keys agree above the low
    16 bits, so it is sorted by counting (histogram of the low bits, scan, fill) whatever its size -- no whole-array
    fallback for skewed keys."""
    from vkradixsort_b200 import capi

    rng = np.random.default_rng(7)
    n = 300_000
    cases = {
        # four 16-bit-prefix buckets of 75,000 keys each
        "four_prefixes": ((rng.integers(0, 4, n, dtype=np.uint32) << 30) | rng.integers(0, 1 << 16, n, dtype=np.uint32)).astype(np.uint32),
        # half of the keys under one 16-bit prefix, the rest uniform
        "hot_prefix": np.where(rng.random(2_000_003) < 0.5, np.uint32(0x2BCD0000) | rng.integers(0, 1 << 16, 2_000_003, dtype=np.uint32),
                               rng.integers(0, 1 << 32, 2_000_003, dtype=np.uint64).astype(np.uint32)).astype(np.uint32),
        # two narrow clusters far apart (bell-shaped, heavy duplicates inside)
        "two_clusters": np.concatenate([(0x20000000 + rng.normal(0, 3000, 700_000)).astype(np.int64),
                                        (0xD0000000 + rng.normal(0, 40000, 800_000)).astype(np.int64)]).astype(np.uint32),
        # one giant bucket spanning several histogram work items, few distinct low values (same-address atomics)
        "giant_few_values": np.concatenate([(np.uint32(0x77770000) | (rng.integers(0, 5, 1_500_000, dtype=np.uint32) * np.uint32(13001))).astype(np.uint32),
                                            np.array([3, 0xFFFFFF00], dtype=np.uint32)]),
    }
    cases["two_clusters"] = rng.permutation(cases["two_clusters"])
    for name, keys in cases.items():
        out, _ = run_sort(handle, keys, dev, capi.SCHEDULE_BUCKET)
        st = handle.bucket_stats()
        assert np.array_equal(out, np.sort(keys)), (name, st)
        assert st["fallback"] == 0 and st["big_buckets"] >= 1, (name, st)
        # ... and the counters are clean for the next sort on this handle
        out, _ = run_sort(handle, keys[::-1].copy(), dev, capi.SCHEDULE_BUCKET)
        assert np.array_equal(out, np.sort(keys)), (name, "second sort")
    st = handle.bucket_stats()


def test_big_bucket_inside_an_item_that_fits_typed_keys(handle, dev, oracle):
    """A bucket of 4097 ... 7676 keys is "big" (counted) but small enough for its whole item to fit a shared-memory buffer:
    the item is then sorted -- and, for typed keys, mapped back -- by the local sort as well.  The counting histogram must
    see the bucket's keys as pass 2 left them (it runs before the local sort).  Found by tools/fuzz_bucket.py: negative
    float32 keys came out wrong (the inverse float map flips the low bits the histogram counts; the int32 map does not).
    Narrow clusters put a few thousand keys into each of several ADJACENT buckets, so that such an item spans few
    buckets and the bins path takes it."""
    from vkradixsort_b200 import capi

    def float_order(bits):
        key = np.where(bits >= np.uint32(1 << 31), ~bits, bits | np.uint32(1 << 31))
        return bits[np.argsort(key, kind="stable")]

    saw_big = 0
    for seed in range(16):
        rng = np.random.default_rng(1000 + seed)
        n = int(rng.integers(30_000, 90_000))
        centres = rng.integers(-(1 << 31), 1 << 31, int(rng.integers(2, 12)), dtype=np.int64)
        sigma = float(10 ** rng.uniform(4.3, 5.3))
        bits = ((centres[rng.integers(0, centres.shape[0], n)] + (rng.normal(0, sigma, n)).astype(np.int64)) & 0xFFFFFFFF).astype(np.uint32)
        bits = np.where(np.isnan(bits.view(np.float32)), np.uint32(0), bits).astype(np.uint32)  # no NaNs: their place is a convention
        for kt, want in ((capi.KEY_F32, float_order(bits)), (capi.KEY_I32, np.sort(bits.view(np.int32)).view(np.uint32)), (capi.KEY_U32, np.sort(bits))):
            b0 = torch.from_numpy(bits.view(np.int32).copy()).to(dev)
            handle.set_schedule(capi.SCHEDULE_BUCKET)
            handle.multi_sort_typed(b0, torch.empty_like(b0), None, capi.multi_push_constants(n, 32), kt)
            handle.check_device_error()
            st = handle.bucket_stats()
            saw_big += st["big_buckets"]
            assert np.array_equal(b0.cpu().numpy().view(np.uint32), want), (seed, n, kt, st)
    assert saw_big > 0, "no case produced a big bucket: the test no longer covers what it is for"


def test_item_table_end_marker_when_n_is_a_multiple_of_the_window(handle, dev, oracle):
    """The item table ends with (number of buckets, n).  With empty buckets at the top of the key window AND n a multiple
    of the item window, the thread that writes the marker computed an empty range (found by tools/fuzz_bucket.py: the
    last item kept a stale end and the tail of the array stayed unsorted)."""
    from vkradixsort_b200 import capi

    rng = np.random.default_rng(17)
    for window in (7168, 6656, 5632, 3072, 512):
        for k in (1, 2, 9, 41):
            n = window * k
            # keys in [0, 0.75 * 2^32): the top quarter of the 65536 buckets is empty
            keys = rng.integers(0, 3 << 30, n, dtype=np.uint64).astype(np.uint32)
            keys[0], keys[1] = 0, (3 << 30) - 1
            for seed_sort in range(2):  # twice: a stale marker of the previous sort must not help
                out, _ = run_sort(handle, keys, dev, capi.SCHEDULE_BUCKET)
                assert np.array_equal(out, np.sort(keys)), (window, k)
            # clusters: larger buckets, smaller windows
            c = (rng.integers(0, 1 << 31, 7, dtype=np.int64)[rng.integers(0, 7, n)] + rng.normal(0, 20000, n).astype(np.int64)) & 0xFFFFFFFF
            out, _ = run_sort(handle, c.astype(np.uint32), dev, capi.SCHEDULE_BUCKET)
            assert np.array_equal(out, np.sort(c.astype(np.uint32))), (window, k, "clusters")
            run_sort(handle, oracle.generate_random(n + 1537, 3, 0xFFFFFFFF), dev, capi.SCHEDULE_BUCKET)  # another size in between


def test_skew_aware_ranking_loops(handle, dev, oracle):
    """The unstable scatter picks its ranking loop per tile and warp from the FIRST round (32 keys) of the warp's
    640-key chunk: plain atomics, "all one digit", or "hot digit" (vkrs_msd.cuh).  Chunks whose later rounds do not look
    like the first one must come out right whichever loop was picked."""
    from vkradixsort_b200 import capi

    rng = np.random.default_rng(99)
    n = 1_920_000  # a whole number of 7680-key tiles: every tile takes the full-tile path
    chunk = 640
    pos = np.arange(n) % chunk
    rnd = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    one = (np.uint32(0x5A000000) | (rnd & np.uint32(0x00FFFFFF))).astype(np.uint32)     # top digit 0x5A
    lane = np.arange(n) % 32
    cases = {
        "first_round_uniform_rest_random": np.where(pos < 32, one, rnd),
        "first_round_random_rest_uniform": np.where(pos < 32, rnd, one),
        "first_round_hot_rest_random": np.where((pos < 32) & (lane < 12), one, rnd),        # 12 of 32 lanes on the hot digit
        "hot_lanes_every_round": np.where(lane % 3 == 0, one, rnd),                       # 11 of 32, every round
        "hot_digit_changes_per_chunk": np.where(lane < 20, (np.uint32(0x01000000) * ((np.arange(n) // chunk) % 256).astype(np.uint32))
                                                | (rnd & np.uint32(0x00FFFFFF)), rnd),
        "all_one_digit": one,
    }
    for name, keys in cases.items():
        keys = keys.astype(np.uint32)
        for schedule in (capi.SCHEDULE_BUCKET, capi.SCHEDULE_LSD_UNSTABLE_FIRST):
            out, _ = run_sort(handle, keys, dev, schedule)
            assert np.array_equal(out, np.sort(keys)), (name, schedule)
    # the same through int32 keys (the typed transform is applied when a key is read from the tile)
    ints = cases["first_round_hot_rest_random"].view(np.int32)
    b0 = torch.from_numpy(ints.copy()).to(dev)
    handle.set_schedule(capi.SCHEDULE_BUCKET)
    handle.multi_sort_typed(b0, torch.empty_like(b0), None, capi.multi_push_constants(n, 32), capi.KEY_I32)
    handle.check_device_error()
    assert np.array_equal(b0.cpu().numpy(), np.sort(ints))


def test_bucket_fallback_is_taken_and_correct(handle, dev, oracle):
    """More big buckets than the counting path takes (256 at 16 low bits): the device raises the fallback word and
    the stable LSD passes behind the schedule sort the array."""
    from vkradixsort_b200 import capi

    n = 300_000
    rng = np.random.default_rng(7)
    # 512 16-bit-prefix buckets of ~5,900 keys each
    m = 3_000_000
    keys0 = ((rng.integers(0, 2, m, dtype=np.uint32) * np.uint32(0xFF000000)) | rng.integers(0, 1 << 24, m, dtype=np.uint32)).astype(np.uint32)
    out0, _ = run_sort(handle, keys0, dev, capi.SCHEDULE_BUCKET)
    st = handle.bucket_stats()
    assert st["fallback"] == 1 and st["big_buckets"] > 256 and st["shift1"] == 24, st
    assert np.array_equal(out0, np.sort(keys0))
    # the next sort on the handle is not disturbed by the unfinished counting
    keys9 = ((rng.integers(0, 4, n, dtype=np.uint32) << 30) | rng.integers(0, 1 << 16, n, dtype=np.uint32)).astype(np.uint32)
    out9, _ = run_sort(handle, keys9, dev, capi.SCHEDULE_BUCKET)
    st = handle.bucket_stats()
    assert st["fallback"] == 0 and st["big_buckets"] == 4, st
    assert np.array_equal(out9, np.sort(keys9))
    # only the occupied key range is sort work: the digit window sits on (key - smallest key), two passes finish the job
    keys1 = (np.uint32(0x12340000) | rng.integers(0, 1 << 16, n, dtype=np.uint32)).astype(np.uint32)
    out1, _ = run_sort(handle, keys1, dev, capi.SCHEDULE_BUCKET)
    st = handle.bucket_stats()
    assert st["fallback"] == 0 and st["shift1"] == 8 and st["shift2"] == 0, st
    assert np.array_equal(out1, np.sort(keys1))
    # one rank's key range after the multi-GPU exchange: top bits constant, the rest uniform
    keys4 = (np.uint32(0xC0000000) | (oracle.generate_random(n, 9, 0xFFFFFFFF) >> np.uint32(2))).astype(np.uint32)
    out4, _ = run_sort(handle, keys4, dev, capi.SCHEDULE_BUCKET)
    st = handle.bucket_stats()
    assert st["fallback"] == 0 and st["shift1"] == 22 and st["recount"] == 1, st
    assert np.array_equal(out4, np.sort(keys4))
    # the same keys spread over many prefixes: no fallback
    keys2 = oracle.generate_random(n, 8, 0xFFFFFFFF)
    out2, _ = run_sort(handle, keys2, dev, capi.SCHEDULE_BUCKET)
    assert handle.bucket_stats()["fallback"] == 0
    assert np.array_equal(out2, np.sort(keys2))
    # all keys equal and no low bits left to sort: nothing to fall back for
    keys3 = np.full(n, 0xFF00, dtype=np.uint32)
    out3, _ = run_sort(handle, keys3, dev, capi.SCHEDULE_BUCKET)
    assert np.array_equal(out3, keys3)


def test_key_span_hint(handle, dev, oracle):
    """vkrs_set_key_span_hint: a right hint saves the histogram recount, a wrong one only costs it."""
    from vkradixsort_b200 import capi

    n = 400_000
    keys = (np.uint32(0xC0000000) | (oracle.generate_random(n, 11, 0xFFFFFFFF) >> np.uint32(2))).astype(np.uint32)
    handle.set_key_span_hint(0xC0000000, 0xFFFFFFFF)
    out, _ = run_sort(handle, keys, dev, capi.SCHEDULE_BUCKET)
    st = handle.bucket_stats()
    assert st["shift1"] == 22 and st["recount"] == 0, st
    assert np.array_equal(out, np.sort(keys))
    handle.set_key_span_hint(0, 0xFFFF)  # wrong: the keys are far outside
    out, _ = run_sort(handle, keys, dev, capi.SCHEDULE_BUCKET)
    st = handle.bucket_stats()
    assert st["shift1"] == 22 and st["recount"] == 1, st
    assert np.array_equal(out, np.sort(keys))
    handle.set_key_span_hint()  # none
    out, _ = run_sort(handle, keys, dev, capi.SCHEDULE_BUCKET)
    assert handle.bucket_stats()["recount"] == 1
    assert np.array_equal(out, np.sort(keys))


def test_typed_keys_through_the_bucket_schedule(handle, dev, oracle):
    """int32 / float32 keys on the bucket schedule: pass 1 reads the keys through the
    order-preserving map, the local sort -- or the last fallback pass, or the plain map-back when no low bits
    are left -- writes them back through the inverse.  Exact against numpy on the same bits."""
    from vkradixsort_b200 import capi

    rng = np.random.default_rng(5)

    def sort_typed(arr, key_type, schedule):
        b0 = torch.from_numpy(arr.view(np.int32).copy()).to(dev)
        b1 = torch.full_like(b0, 0x5A5A5A5A)
        handle.set_schedule(schedule)
        handle.multi_sort_typed(b0, b1, None, capi.multi_push_constants(arr.shape[0], 32), key_type)
        handle.check_device_error()
        return b0.cpu().numpy()

    def float_order(f):
        bits = f.view(np.uint32)
        ordered = np.where(bits >> 31, ~bits, bits | np.uint32(0x80000000))
        return f[np.argsort(ordered, kind="stable")]

    for n, schedule in ((1, capi.SCHEDULE_BUCKET), (6145, capi.SCHEDULE_BUCKET), (300_001, capi.SCHEDULE_BUCKET),
                        (5_000_011, capi.SCHEDULE_AUTO)):
        ints = rng.integers(-(1 << 31), 1 << 31, size=n, dtype=np.int64).astype(np.int32)
        assert np.array_equal(sort_typed(ints, capi.KEY_I32, schedule), np.sort(ints)), (n, "int32 uniform")
        assert handle.bucket_stats()["fallback"] == 0
        small = rng.integers(-100, 101, size=n, dtype=np.int64).astype(np.int32)  # 201 values around zero: two passes, no local sort
        assert np.array_equal(sort_typed(small, capi.KEY_I32, schedule), np.sort(small)), (n, "int32 in [-100, 100]")
        st = handle.bucket_stats()
        assert st["fallback"] == 0 and st["shift2"] == 0, st
        hot = np.where(rng.random(n) < 0.5, np.int32(5), ints).astype(np.int32)  # half of the keys equal: one bucket is finished by counting, its fill undoes the map
        assert np.array_equal(sort_typed(hot, capi.KEY_I32, schedule), np.sort(hot)), (n, "int32, one hot value")
        if n > 20_000:
            st = handle.bucket_stats()
            assert st["fallback"] == 0 and st["big_buckets"] == 1, st
        many = (rng.integers(0, 600, size=n, dtype=np.int64) * 7_000_001 - (1 << 31)).astype(np.int32)  # 600 values: too many big buckets, the fallback passes undo the map
        assert np.array_equal(sort_typed(many, capi.KEY_I32, schedule), np.sort(many)), (n, "int32, 600 values")
        if n > 3_000_000:
            assert handle.bucket_stats()["fallback"] == 1
        nonneg = rng.integers(0, 1 << 16, size=n, dtype=np.int64).astype(np.int32)  # 16 varying bits: no local sort, plain map-back
        assert np.array_equal(sort_typed(nonneg, capi.KEY_I32, schedule), np.sort(nonneg)), (n, "int32 in [0, 65535]")
        st = handle.bucket_stats()
        assert st["shift2"] == 0 and st["fallback"] == 0, st
        f = (rng.standard_normal(n) * 1e6).astype(np.float32)
        f[:: 7] = 0.0
        f[1:: 11] = -0.0
        f[2:: 13] = np.inf
        f[3:: 17] = -np.inf
        got = sort_typed(f, capi.KEY_F32, schedule).view(np.float32)
        assert np.array_equal(got.view(np.uint32), float_order(f).view(np.uint32)), (n, "float32")
        u = oracle.generate_random(n, 31 + n, 0xFFFFFFFF)
        assert np.array_equal(sort_typed(u, capi.KEY_U32, schedule).view(np.uint32), np.sort(u)), (n, "uint32")
    handle.set_schedule(capi.SCHEDULE_AUTO)


def test_bucket_workspace_reuse_and_unaligned(handle, dev, oracle):
    """Sizes going up and down on one handle (the pass-1 piece table is cached per N) and key buffers
    that start at a 4-byte, not 16-byte, aligned address (no TMA: the workers copy the tiles in)."""
    from vkradixsort_b200 import capi

    handle.set_schedule(capi.SCHEDULE_BUCKET)
    for i, n in enumerate([700_001, 9_000, 700_001, 1_500_000, 3, 700_001]):
        keys = oracle.generate_random(n, 300 + i, 0xFFFFFFFF)
        out, _ = run_sort(handle, keys, dev, capi.SCHEDULE_BUCKET)
        assert np.array_equal(out, np.sort(keys)), n
    n = 250_001
    for off in (1, 2, 3):
        keys = oracle.generate_random(n, 70 + off, 0xFFFFFFFF)
        big0 = torch.zeros(n + 8, dtype=torch.int32, device=dev)
        big1 = torch.zeros(n + 8, dtype=torch.int32, device=dev)
        big0[off:off + n] = to_dev(keys, dev)
        handle.multi_sort(big0[off:], big1[off:], None, capi.multi_push_constants(n, 32))
        handle.check_device_error()
        assert np.array_equal(to_host(big0[off:off + n]), np.sort(keys)), off
        assert int(big0[:off].abs().sum()) == 0 and int(big0[off + n:].abs().sum()) == 0, "wrote outside the buffer"


def test_auto_schedule_picks_bucket_for_large_n(handle, dev, oracle):
    """6*10^7 keys through the default (auto) schedule: sortedness + multiset checksum on the device,
    then element-wise against numpy."""
    from vkradixsort_b200 import capi

    n = 60_000_000
    g = torch.Generator(device=dev)
    g.manual_seed(2024)
    b0 = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int32, device=dev, generator=g)
    keys = to_host(b0)
    b1 = torch.empty_like(b0)
    handle.set_schedule(capi.SCHEDULE_AUTO)
    launches0 = handle.launch_count
    handle.multi_sort(b0, b1, None, capi.multi_push_constants(n, 32))
    handle.check_device_error()
    st = handle.bucket_stats()
    assert st["shift1"] == 24 and st["fallback"] == 0 and st["pieces2"] > 0, st
    assert handle.launch_count - launches0 >= 8
    assert np.array_equal(to_host(b0), np.sort(keys))
