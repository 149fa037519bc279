"""CPU tests: pin the oracle (the CPU restatement of the reference shaders) against the only
criterion the reference itself holds -- output == std::sort(input), element-wise
(MultiRadixSort.cpp:148-161) -- and against the committed fixtures."""
import glob
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SIZES = [0, 1, 2, 100, 255, 256, 257, 1000, 8191, 8192, 8193, 100000]
NBS = [1, 3, 32]


def test_dispatch_sizing(oracle):
    # MultiRadixSort.cpp:13-15 + ComputePass.h:24-29; README.md:261 quotes 123 work groups
    assert oracle.global_invocation_size(1000000, 32) == 31250
    assert oracle.workgroup_count(1000000, 32) == 123
    assert oracle.workgroup_count(100000000, 32) == 12208
    assert oracle.workgroup_count(100000000, 4096) == 96
    assert oracle.workgroup_count(1000, 32) == 1
    assert oracle.workgroup_count(0, 32) == 0
    assert oracle.workgroup_count(1, 32) == 1
    assert oracle.global_invocation_size(33, 32) == 2


@pytest.mark.parametrize("nb", NBS)
@pytest.mark.parametrize("max_value", [0xFFFFFFFF, 0x0FFFFFFF, 3])
def test_multi_sort_equals_std_sort(oracle, nb, max_value):
    for n in SIZES:
        keys = oracle.generate_random(n, 1234 + n, max_value)
        buf0, _, _ = oracle.multi_sort(keys, nb)
        ref = keys.copy()
        oracle.std_sort(ref)  # the reference's own baseline + verifier
        assert oracle.test_sort(ref, buf0) == -1
        assert np.array_equal(buf0, np.sort(keys))


@pytest.mark.parametrize("nb", NBS)
def test_multi_sort_is_stable(oracle, nb):
    for n in SIZES:
        keys = oracle.generate_random(n, 99 + n, 15)  # many duplicates
        vals = np.arange(n, dtype=np.uint32)
        buf0, _, _, v0, _ = oracle.multi_sort(keys, nb, vals)
        assert np.array_equal(v0, np.argsort(keys, kind="stable").astype(np.uint32))
        k2, v2 = keys.copy(), vals.copy()
        oracle.stable_sort_pairs(k2, v2)
        assert np.array_equal(buf0, k2) and np.array_equal(v0, v2)


def test_single_sort(oracle):
    for n in SIZES:
        for mx in (0xFFFFFFFF, 0x0FFFFFFF, 1):
            keys = oracle.generate_random(n, 7 + n, mx)
            assert np.array_equal(oracle.single_sort(keys), np.sort(keys))
    keys = oracle.generate_random(3000, 5, 3)
    vals = np.arange(3000, dtype=np.uint32)
    k, v = oracle.single_sort(keys, vals)
    assert np.array_equal(v, np.argsort(keys, kind="stable").astype(np.uint32))


def test_adversarial_inputs(oracle):
    n = 20000
    cases = {
        "all_equal": np.full(n, 0xDEADBEEF, dtype=np.uint32),
        "sorted": np.arange(n, dtype=np.uint32) * 7919,
        "descending": (n - np.arange(n, dtype=np.uint32)),  # SingleRadixSort.cpp:96 (commented-out input)
        "two_valued": (np.arange(n, dtype=np.uint32) % 2) * 0xFFFFFFFF,
        "zeros": np.zeros(n, dtype=np.uint32),
        "max": np.full(n, 0xFFFFFFFF, dtype=np.uint32),
    }
    for name, keys in cases.items():
        keys = np.ascontiguousarray(keys.astype(np.uint32))
        assert np.array_equal(oracle.multi_sort(keys, 32)[0], np.sort(keys)), name
        assert np.array_equal(oracle.single_sort(keys), np.sort(keys)), name


def test_histogram_rows_and_scatter_stage(oracle):
    n, nb, shift = 10000, 3, 8
    keys = oracle.generate_random(n, 42)
    pc = oracle.push_constants(n, shift, nb)
    hist = oracle.multi_histograms(keys, pc).reshape(-1, 256)
    slab = nb * 256
    for w in range(pc.g_num_workgroups):
        d = (keys[w * slab:(w + 1) * slab] >> shift) & 255
        assert np.array_equal(hist[w], np.bincount(d, minlength=256).astype(np.uint32))
    out = oracle.multi_scatter(keys, hist.reshape(-1), pc)
    order = np.argsort((keys >> shift) & 255, kind="stable")
    assert np.array_equal(out, keys[order])


def test_sort64(oracle):
    for n in (0, 1, 1000, 8193):
        keys = oracle.generate_random64(n, 5 + n)
        assert np.array_equal(oracle.multi_sort64(keys), np.sort(keys))


def test_generator_is_deterministic_and_in_range(oracle):
    a = oracle.generate_random(1000, 0x5EED0001, 0x0FFFFFFF)
    b = oracle.generate_random(1000, 0x5EED0001, 0x0FFFFFFF)
    assert np.array_equal(a, b) and a.max() <= 0x0FFFFFFF
    # std::mt19937 known answer: the 10000th output of a default-seeded engine is 4123659995
    # (C++ standard [rand.predef]); with a full-range uint32 distribution libstdc++ passes the
    # engine output through unchanged.
    c = oracle.generate_random(10000, 5489, 0xFFFFFFFF)
    assert int(c[9999]) == 4123659995


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))))
def test_golden_fixtures(oracle, path):
    g = np.load(path)
    keys, nb = g["keys"], int(g["nb"])
    n = keys.shape[0]
    buf0, buf1, hist = oracle.multi_sort(keys, nb)
    assert np.array_equal(buf0, g["sorted"])
    assert np.array_equal(buf1, g["final_buf1"])
    assert np.array_equal(hist, g["final_hist"])
    pc = oracle.push_constants(n, 8, nb)
    h8 = oracle.multi_histograms(keys, pc)
    assert np.array_equal(h8, g["hist_shift8"])
    assert np.array_equal(oracle.multi_scatter(keys, h8, pc), g["pass_shift8"])
    assert np.array_equal(oracle.single_sort(keys), g["sorted"])
    vals = np.arange(n, dtype=np.uint32)
    assert np.array_equal(oracle.multi_sort(keys, nb, vals)[3], g["stable_order"])
