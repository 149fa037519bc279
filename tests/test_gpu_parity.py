"""GPU parity tests (run on the B200 box with -m gpu): every C-ABI compute entry point against the
oracle (CPU restatement of the reference shaders) on identical seeded inputs, bit-exact.

Bar (integer work): element-wise equality, as MultiRadixSort::testSort demands
(multiradixsort/src/MultiRadixSort.cpp:148-161).  Nothing here reads /root/reference.
"""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def handle(built_lib):
    from vkradixsort_b200 import Handle

    h = Handle(0, 1 << 20)
    yield h
    h.close()


def to_dev(a, dev):
    if a.dtype == np.uint32:
        return torch.from_numpy(a.view(np.int32)).to(dev)
    if a.dtype == np.uint64:
        return torch.from_numpy(a.view(np.int64)).to(dev)
    raise TypeError(a.dtype)


def to_host(t):
    a = t.cpu().numpy()
    return a.view(np.uint32) if a.dtype == np.int32 else a.view(np.uint64)


def scratch_like(t):
    # poison the scratch buffer so that a path which forgets to write something cannot pass by luck
    return torch.full_like(t, 0x5A5A5A5A if t.dtype == torch.int32 else 0x5A5A5A5A5A5A5A5A)


SIZES = [1, 2, 31, 100, 255, 256, 257, 1000, 4095, 4096, 4097, 8191, 8192, 8193, 16385, 100000, 1 << 20, 3000001]


def gpu_multi_sort(handle, keys, dev, nb=32, variant=None):
    from vkradixsort_b200 import capi

    n = keys.shape[0]
    b0 = to_dev(keys, dev)
    b1 = scratch_like(b0)
    pc = capi.multi_push_constants(n, nb)
    hist = torch.zeros(max(1, pc.g_num_workgroups) * 256, dtype=torch.int32, device=dev)
    if variant is not None:
        handle.set_variant(variant)
    handle.multi_sort(b0, b1, hist, pc)
    handle.check_device_error()
    return to_host(b0)


@pytest.mark.parametrize("max_value", [0xFFFFFFFF, 0x0FFFFFFF, 3])
def test_fused_multi_sort_matches_oracle(handle, dev, oracle, max_value):
    handle.set_variant(0)
    for n in SIZES:
        keys = oracle.generate_random(n, 1234 + n, max_value)
        out = gpu_multi_sort(handle, keys, dev)
        if n <= 100000:
            expect = oracle.multi_sort(keys, 32)[0]  # the restated reference shaders
        else:
            expect = np.sort(keys)  # == std::sort, the reference's own criterion
        assert oracle.test_sort(expect, out) == -1, f"n={n} max_value={max_value:#x}"


def test_fused_all_variants(handle, dev, oracle):
    from vkradixsort_b200 import capi

    keys = oracle.generate_random(777_777, 31337, 0xFFFFFFFF)
    dups = oracle.generate_random(300_001, 31338, 255)
    for v in range(capi.num_variants()):
        for k in (keys, dups):
            out = gpu_multi_sort(handle, k, dev, variant=v)
            assert np.array_equal(out, np.sort(k)), capi.variant_name(v)
    handle.set_variant(0)


def test_fused_zero_elements_is_a_noop(handle, dev):
    from vkradixsort_b200 import capi

    b0 = torch.zeros(4, dtype=torch.int32, device=dev)
    b1 = torch.zeros(4, dtype=torch.int32, device=dev)
    pc = capi.multi_push_constants(0, 32)
    assert pc.g_num_workgroups == 0
    handle.multi_sort(b0, b1, None, pc)
    handle.sort_auto(b0, b1, 0)


def test_adversarial_inputs(handle, dev, oracle):
    n = 250_007
    ar = np.arange(n, dtype=np.uint32)
    cases = {
        "all_equal": np.full(n, 0xDEADBEEF, dtype=np.uint32),
        "zeros": np.zeros(n, dtype=np.uint32),
        "max": np.full(n, 0xFFFFFFFF, dtype=np.uint32),  # collides with the tail padding value
        "sorted": ar * np.uint32(7919),
        "descending": (n - ar).astype(np.uint32),  # SingleRadixSort.cpp:96
        "two_valued": (ar % 2) * np.uint32(0xFFFFFFFF),
        "one_hot_digit": ((ar % 3) << 16).astype(np.uint32),
        "mostly_max": np.where(ar % 97 == 0, ar, np.uint32(0xFFFFFFFF)).astype(np.uint32),
    }
    for name, keys in cases.items():
        keys = np.ascontiguousarray(keys, dtype=np.uint32)
        out = gpu_multi_sort(handle, keys, dev)
        assert np.array_equal(out, np.sort(keys)), name


def test_repeated_sorts_reuse_workspace(handle, dev, oracle):
    """Tile status is recycled between passes and between calls; sizes going up and down must
    not see stale status words."""
    for i, n in enumerate([500_000, 20_000, 1_000_003, 9000, 500_000, 1]):
        keys = oracle.generate_random(n, 900 + i, 0xFFFFFFFF)
        assert np.array_equal(gpu_multi_sort(handle, keys, dev), np.sort(keys)), n


def test_unaligned_sub_buffers(handle, dev, oracle):
    """Keys starting at a 4-byte (not 16-byte) aligned address: the multi-GPU path sorts slices."""
    from vkradixsort_b200 import capi

    n = 123_457
    for off in (1, 2, 3):
        keys = oracle.generate_random(n, 77 + off, 0xFFFFFFFF)
        big0 = torch.zeros(n + 8, dtype=torch.int32, device=dev)
        big1 = torch.zeros(n + 8, dtype=torch.int32, device=dev)
        big0[off:off + n] = to_dev(keys, dev)
        handle.multi_sort(big0[off:], big1[off:], None, capi.multi_push_constants(n, 32))
        handle.check_device_error()
        assert np.array_equal(to_host(big0[off:off + n]), np.sort(keys))
        assert int(big0[:off].abs().sum()) == 0 and int(big0[off + n:].abs().sum()) == 0  # no out-of-range writes


@pytest.mark.parametrize("nb", [1, 3, 32, 512])
def test_staged_stages_match_oracle(handle, dev, oracle, nb):
    """Stage by stage: the histogram matrix and one scatter pass, bit-exact against the restated
    multi_radixsort_histograms.comp / multi_radixsort.comp."""
    from vkradixsort_b200 import capi

    for n in (1, 255, 1000, 8193, 100_000):
        keys = oracle.generate_random(n, 5150 + n, 0xFFFFFFFF)
        for shift in (0, 8, 24):
            pc = capi.multi_push_constants(n, nb, shift)
            opc = oracle.push_constants(n, shift, nb)
            assert opc.g_num_workgroups == pc.g_num_workgroups
            d_in = to_dev(keys, dev)
            d_hist = torch.full((pc.g_num_workgroups * 256,), -1, dtype=torch.int32, device=dev)
            handle.multi_histograms(d_in, d_hist, pc)
            exp_hist = oracle.multi_histograms(keys, opc)
            assert np.array_equal(to_host(d_hist), exp_hist), (n, nb, shift)
            d_out = scratch_like(d_in)
            handle.multi_scatter(d_in, d_out, d_hist, pc)
            assert np.array_equal(to_host(d_out), oracle.multi_scatter(keys, exp_hist, opc)), (n, nb, shift)


@pytest.mark.parametrize("nb", [1, 32, 4096])
def test_staged_sort_through_reference_shaped_passes(dev, oracle, built_lib, nb):
    from vkradixsort_b200 import GPUContext, MultiRadixSort

    for n in (1000, 100_003, 1_000_000):
        if nb == 1 and n > 200_000:
            continue
        keys = oracle.generate_random(n, 4242 + n, 0x0FFFFFFF)  # the reference's 28-bit distribution
        app = MultiRadixSort(nb)
        bufs = [to_dev(keys, dev), torch.zeros(n, dtype=torch.int32, device=dev),
                torch.zeros(app.histogram_buffer_elements(n), dtype=torch.int32, device=dev)]
        app.execute(GPUContext(0), bufs, n)
        exp0, exp1, exp_hist = oracle.multi_sort(keys, nb)
        assert oracle.test_sort(exp0, to_host(bufs[0])) == -1
        # the staged path reproduces even the scratch contents the reference leaves behind
        assert np.array_equal(to_host(bufs[1]), exp1)
        assert np.array_equal(to_host(bufs[2])[: exp_hist.shape[0]], exp_hist)


def test_single_sort_matches_oracle(dev, oracle, built_lib):
    from vkradixsort_b200 import GPUContext, SingleRadixSort

    for n in (1, 2, 100, 255, 256, 257, 1000, 8191, 8192, 8193, 50_000):
        for mx in (0xFFFFFFFF, 0x0FFFFFFF, 1):
            keys = oracle.generate_random(n, 7 + n, mx)
            bufs = [to_dev(keys, dev), torch.zeros(n, dtype=torch.int32, device=dev)]
            SingleRadixSort().execute(GPUContext(0), bufs, n)
            assert oracle.test_sort(oracle.single_sort(keys), to_host(bufs[0])) == -1, (n, mx)


def test_small_sorts_one_launch_and_its_gated_second_kernel(handle, dev, oracle):
    """vkrs_single_sort below 6141 keys: the keys are one item of the local sort (small_sort_kernel); only an
    over-full bin -- a few distinct values far apart -- leaves the work to the four-pass kernel behind it."""
    from vkradixsort_b200 import capi

    rng = np.random.default_rng(11)

    def run(keys):
        b0, b1 = to_dev(keys, dev), torch.zeros(keys.shape[0], dtype=torch.int32, device=dev)
        handle.single_sort(b0, b1, capi.SinglePushConstants(keys.shape[0]))
        handle.check_device_error()
        return to_host(b0)

    for n in (1, 2, 3, 31, 32, 33, 255, 256, 1000, 4095, 4096, 6141, 7000, 7675, 7676, 7677, 8000):  # 7676 = the largest one-launch sort
        cases = {
            "uniform32": oracle.generate_random(n, 70 + n, 0xFFFFFFFF),
            "reference28": oracle.generate_random(n, 71 + n, 0x0FFFFFFF),
            "narrow": (np.uint32(0xFFFFF000) + rng.integers(0, 4000, n, dtype=np.uint32)).astype(np.uint32),  # exact mode, up to 2^32 - 1
            "all_equal": np.full(n, 0x80000000, dtype=np.uint32),
            "extremes": rng.choice(np.array([0, 0xFFFFFFFF], dtype=np.uint32), n),                 # two values, the whole key range apart
            "few_far_apart": rng.choice(np.array([5, 1 << 20, 3 << 30, 0xFFFFFFF0], dtype=np.uint32), n),  # over-full bins: the gated kernel sorts
            "sorted": np.sort(oracle.generate_random(n, 72 + n, 0xFFFFFFFF)),
            "descending": np.sort(oracle.generate_random(n, 73 + n, 0xFFFFFFFF))[::-1].copy(),
        }
        for name, keys in cases.items():
            assert np.array_equal(run(keys), np.sort(keys)), (n, name)
    # unaligned sub-buffer
    base = to_dev(oracle.generate_random(3003, 5, 0xFFFFFFFF), dev)
    view = base[3:]
    want = np.sort(to_host(view))
    handle.single_sort(view, torch.zeros(3000, dtype=torch.int32, device=dev), capi.SinglePushConstants(3000))
    assert np.array_equal(to_host(view), want)


def test_sort_auto_crossover(handle, dev, oracle):
    for n in (1, 1000, 4096, 4097, 12_288, 12_289, 100_000):
        keys = oracle.generate_random(n, 17 + n, 0xFFFFFFFF)
        b0 = to_dev(keys, dev)
        b1 = scratch_like(b0)
        handle.sort_auto(b0, b1, n)
        handle.check_device_error()
        assert np.array_equal(to_host(b0), np.sort(keys)), n


def test_pairs_are_stable(handle, dev, oracle):
    from vkradixsort_b200 import capi

    for n, mx in ((1, 3), (1000, 15), (8193, 0xFFFFFFFF), (300_001, 15), (1_000_003, 0xFFFFFFFF), (500_000, 0)):
        keys = oracle.generate_random(n, 99 + n, mx)
        vals = np.arange(n, dtype=np.uint32)
        k0, v0 = to_dev(keys, dev), to_dev(vals, dev)
        k1, v1 = scratch_like(k0), scratch_like(v0)
        handle.multi_sort_pairs(k0, k1, v0, v1, None, capi.multi_push_constants(n, 32))
        handle.check_device_error()
        ek, ev = keys.copy(), vals.copy()
        oracle.stable_sort_pairs(ek, ev)  # std::stable_sort by key
        assert np.array_equal(to_host(k0), ek), (n, mx)
        assert np.array_equal(to_host(v0), ev), (n, mx)


def test_pairs_2e7_exact_against_stable_sort(handle, dev, oracle):
    """BASELINE.json config 3 at a size std::stable_sort finishes in seconds: keys AND payloads bit-exact."""
    from vkradixsort_b200 import capi

    n = 20_000_000
    keys = oracle.generate_random(n, 0x5EED0003, 0x0FFFFFFF)  # the reference's distribution; thousands of duplicates
    keys[::7] &= np.uint32(0xFFFF)  # ... and heavy duplication in a seventh of them
    vals = np.arange(n, dtype=np.uint32)
    k0, v0 = to_dev(keys, dev), to_dev(vals, dev)
    k1, v1 = scratch_like(k0), scratch_like(v0)
    handle.multi_sort_pairs(k0, k1, v0, v1, None, capi.multi_push_constants(n, 32))
    handle.check_device_error()
    ek, ev = keys.copy(), vals.copy()
    oracle.stable_sort_pairs(ek, ev)
    assert np.array_equal(to_host(k0), ek)
    assert np.array_equal(to_host(v0), ev)


def test_pairs_1e8_properties(handle, dev):
    """BASELINE.json config 3 at its stated size, 10^8 key + payload pairs, checked on the device through
    size-independent properties: keys ascending; the payload (= original index) leads back to the key; among equal
    keys the payloads ascend (the stable order the reference's ranking defines, multi_radixsort.comp:107-122)."""
    from vkradixsort_b200 import capi

    n = 100_000_000
    gen = torch.Generator(device=dev)
    gen.manual_seed(0x5EED0033)
    keys = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int32, device=dev, generator=gen)
    keys[: n // 4] &= 0xFFFFF  # a quarter of the keys share 2^20 values: ~24 duplicates each
    k0, v0 = keys.clone(), torch.arange(n, dtype=torch.int32, device=dev)
    k1, v1 = torch.empty_like(k0), torch.empty_like(v0)
    handle.multi_sort_pairs(k0, k1, v0, v1, None, capi.multi_push_constants(n, 32))
    handle.check_device_error()
    del k1, v1
    flip = torch.tensor(-(1 << 31), dtype=torch.int32, device=dev)
    o = k0 ^ flip
    assert bool((o[1:] >= o[:-1]).all()), "keys not ascending"
    del o
    assert bool((keys[v0.long()] == k0).all()), "payload does not lead back to its key"
    eq = k0[1:] == k0[:-1]
    assert int(eq.sum()) > 1_000_000
    assert bool((v0[1:][eq] > v0[:-1][eq]).all()), "equal keys out of input order: not stable"
    # the payloads are a permutation of 0 .. n-1
    assert int(v0.to(torch.int64).sum()) == n * (n - 1) // 2


def test_staged_scatter_with_values(handle, dev, oracle):
    from vkradixsort_b200 import capi

    n, nb, shift = 50_001, 32, 8
    keys = oracle.generate_random(n, 606, 0xFFFFFFFF)
    vals = np.arange(n, dtype=np.uint32)
    pc = capi.multi_push_constants(n, nb, shift)
    opc = oracle.push_constants(n, shift, nb)
    d_in, d_vin = to_dev(keys, dev), to_dev(vals, dev)
    d_hist = torch.zeros(pc.g_num_workgroups * 256, dtype=torch.int32, device=dev)
    d_out, d_vout = scratch_like(d_in), scratch_like(d_vin)
    handle.multi_histograms(d_in, d_hist, pc)
    handle.multi_scatter(d_in, d_out, d_hist, pc, d_vin, d_vout)
    ek, ev = oracle.multi_scatter(keys, oracle.multi_histograms(keys, opc), opc, vals)
    assert np.array_equal(to_host(d_out), ek) and np.array_equal(to_host(d_vout), ev)


def test_u64_sort(handle, dev, oracle):
    from vkradixsort_b200 import capi

    for n, mx in ((1, 0x0FFFFFFFFFFF), (1000, 0x0FFFFFFFFFFF), (8193, 0xFFFFFFFFFFFFFFFF), (700_001, 0x0FFFFFFFFFFF)):
        keys = oracle.generate_random64(n, 5 + n, mx)  # MultiRadixSort.cpp:128 uses the 44-bit range
        b0 = to_dev(keys, dev)
        b1 = scratch_like(b0)
        handle.multi_sort_u64(b0, b1, None, capi.multi_push_constants(n, 32))
        handle.check_device_error()
        expect = oracle.multi_sort64(keys) if n <= 10_000 else np.sort(keys)
        assert np.array_equal(to_host(b0), expect), (n, mx)


def test_host_buffer_entry(handle, oracle):
    """vkrs_multi_sort_host: prepareBuffers + loop + download (MultiRadixSort.cpp:83-102)."""
    for n in (1000, 1_000_003):
        keys = oracle.generate_random(n, 2024 + n, 0xFFFFFFFF)
        host = torch.from_numpy(keys.view(np.int32).copy()).pin_memory()
        handle.multi_sort_host(host, n)
        assert np.array_equal(host.numpy().view(np.uint32), np.sort(keys))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))))
def test_golden_fixtures(handle, dev, oracle, path):
    from vkradixsort_b200 import GPUContext, SingleRadixSort, capi

    g = np.load(path)
    keys, nb = g["keys"], int(g["nb"])
    n = keys.shape[0]
    assert np.array_equal(gpu_multi_sort(handle, keys, dev, nb), g["sorted"])
    pc = capi.multi_push_constants(n, nb, 8)
    d_in = to_dev(keys, dev)
    d_hist = torch.zeros(pc.g_num_workgroups * 256, dtype=torch.int32, device=dev)
    d_out = scratch_like(d_in)
    handle.multi_histograms(d_in, d_hist, pc)
    handle.multi_scatter(d_in, d_out, d_hist, pc)
    assert np.array_equal(to_host(d_hist), g["hist_shift8"])
    assert np.array_equal(to_host(d_out), g["pass_shift8"])
    bufs = [to_dev(keys, dev), torch.zeros(n, dtype=torch.int32, device=dev)]
    SingleRadixSort().execute(GPUContext(0), bufs, n)
    assert np.array_equal(to_host(bufs[0]), g["sorted"])
    vals = np.arange(n, dtype=np.uint32)
    k0, v0 = to_dev(keys, dev), to_dev(vals, dev)
    handle.multi_sort_pairs(k0, scratch_like(k0), v0, scratch_like(v0), None, capi.multi_push_constants(n, nb))
    assert np.array_equal(to_host(v0), g["stable_order"])


def test_full_size_1e8_properties(handle, dev, oracle):
    """BASELINE.json config 2 at full size, through size-independent properties computed on the
    device (sortedness + multiset checksums), plus an exact comparison of a strided sample of
    order statistics against numpy's sort of the same input."""
    from vkradixsort_b200 import capi

    n = 100_000_000
    gen = torch.Generator(device=dev)
    gen.manual_seed(0x5EED0002)
    b0 = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int32, device=dev, generator=gen)
    keys_host = b0.cpu().numpy().view(np.uint32)
    sum_in = int(b0.to(torch.int64).sum())
    b1 = scratch_like(b0)
    handle.multi_sort(b0, b1, None, capi.multi_push_constants(n, 32))
    handle.check_device_error()
    # unsigned order == signed order after flipping the sign bit
    flipped = b0 ^ torch.tensor(-(1 << 31), dtype=torch.int32, device=dev)
    assert bool((flipped[1:] >= flipped[:-1]).all()), "output is not sorted"
    assert int(b0.to(torch.int64).sum()) == sum_in, "multiset changed (sum)"
    del flipped
    expect = np.sort(keys_host)
    out = to_host(b0)
    assert oracle.test_sort(expect, out) == -1


def test_key_range_and_partition(handle, dev, oracle):
    """Device side of the multi-GPU exchange: bucket(key) = min(255, (key - base) >> shift), stable."""
    from vkradixsort_b200 import dist as D

    for n, mx in ((1, 0xFFFFFFFF), (1000, 0xFFFFFFFF), (300_001, 0x0FFFFFFF), (1_000_003, 0xFFFFFFFF), (50_000, 1000)):
        keys = oracle.generate_random(n, 4000 + n, mx)
        vals = np.arange(n, dtype=np.uint32)
        d_k, d_v = to_dev(keys, dev), to_dev(vals, dev)
        mm = torch.zeros(2, dtype=torch.int32, device=dev)
        handle.key_range(d_k, n, mm)
        lo, hi = (int(x) & 0xFFFFFFFF for x in mm.tolist())
        assert (lo, hi) == (int(keys.min()), int(keys.max()))
        base, shift = D.choose_bucket_map(lo, hi)
        bucket = np.minimum(255, (keys - np.uint32(base)) >> np.uint32(shift))
        order = np.argsort(bucket, kind="stable")
        counts = torch.zeros(256, dtype=torch.int32, device=dev)
        for with_vals in (False, True):
            o_k, o_v = scratch_like(d_k), scratch_like(d_v)
            handle.partition(d_k, o_k, n, base, shift, counts, d_v if with_vals else None, o_v if with_vals else None)
            assert np.array_equal(to_host(o_k), keys[order]), (n, mx, with_vals)
            assert np.array_equal(to_host(counts), np.bincount(bucket, minlength=256).astype(np.uint32))
            if with_vals:
                assert np.array_equal(to_host(o_v), vals[order])


def test_example_programs(built_lib):
    """The two example executables are the reference's own end-to-end tests: same
    stdout lines, exit code 0 and "Test passed." on success."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bindir = os.path.join(root, "vkradixsort_b200", "bin")
    if not os.path.exists(os.path.join(bindir, "multiradixsortexample")):
        subprocess.run(["make", "-C", os.path.join(root, "examples"), "-s"], check=True)
    for args in (["multiradixsortexample"], ["multiradixsortexample", "100003", "3", "--seed", "7", "--bits", "32"],
                 ["multiradixsortexample", "1e7", "32", "--fast", "--seed", "9"], ["singleradixsortexample", "1000", "--seed", "1"],
                 ["singleradixsortexample"]):
        r = subprocess.run([os.path.join(bindir, args[0])] + args[1:], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        tag = "[MultiRadixSort] " if args[0].startswith("multi") else "[SingleRadixSort] "
        lines = r.stdout.strip().splitlines()
        assert lines[0].startswith(tag + "Sorting ") and lines[0].endswith("32bit numbers.")
        assert any(l.startswith(tag + "GPU sort finished in ") and l.endswith("[ms].") for l in lines)
        assert any(l.startswith(tag + "CPU sort finished in ") for l in lines)
        assert lines[-1] == tag + "Test passed."


def test_two_gpu_bucket_exchange():
    """BASELINE.json config 5 in miniature: 2 ranks over NCCL, checked against numpy on rank 0."""
    import subprocess
    import sys

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541",
                        os.path.join(root, "tests", "dist_gpu_worker.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "DIST_OK" in r.stdout


def test_config5_single_gpu_size_properties(handle, dev):
    """BASELINE.json config 5 at G=1: 8 x 10^8 keys on one GPU (3.2 GB per buffer), checked through
    size-independent properties on the device: sortedness and the multiset checksum."""
    from vkradixsort_b200 import capi

    n = 800_000_000
    gen = torch.Generator(device=dev)
    gen.manual_seed(0x5EED0005)
    b0 = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int32, device=dev, generator=gen)
    sum_in = int(b0.to(torch.int64).sum())
    b1 = torch.empty_like(b0)
    handle.multi_sort(b0, b1, None, capi.multi_push_constants(n, 32))
    handle.check_device_error()
    del b1
    flip = torch.tensor(-(1 << 31), dtype=torch.int32, device=dev)
    for c in range(0, n, 1 << 28):
        x = b0[c:min(n, c + (1 << 28) + 1)] ^ flip
        assert bool((x[1:] >= x[:-1]).all()), f"not sorted in chunk starting at {c}"
        del x
    assert int(b0.to(torch.int64).sum()) == sum_in


def test_typed_keys_signed_and_float(handle, dev, oracle):
    """Fused key transforms: int32 and float32 keys sort by value with the same
    four passes; exact against numpy on the same bits."""
    from vkradixsort_b200 import capi

    rng = np.random.default_rng(77)
    for n in (1, 1000, 8193, 300_001, 2_000_003):
        pc = capi.multi_push_constants(n, 32)
        ints = rng.integers(-(1 << 31), 1 << 31, size=n, dtype=np.int64).astype(np.int32)
        b0 = torch.from_numpy(ints.copy()).to(dev)
        handle.multi_sort_typed(b0, scratch_like(b0), None, pc, capi.KEY_I32)
        assert np.array_equal(b0.cpu().numpy(), np.sort(ints)), n
        f = rng.standard_normal(n).astype(np.float32) * np.float32(1e6)
        f[:: 7] = 0.0
        f[1:: 11] = -0.0
        f[2:: 13] = np.inf
        f[3:: 17] = -np.inf
        b0 = torch.from_numpy(f.copy()).to(dev)
        handle.multi_sort_typed(b0, torch.empty_like(b0), None, pc, capi.KEY_F32)
        got = b0.cpu().numpy()
        bits = f.view(np.uint32)
        ordered = np.where(bits >> 31, ~bits, bits | np.uint32(0x80000000))  # the same order-preserving map, on the host
        expect = f[np.argsort(ordered, kind="stable")]
        assert np.array_equal(got.view(np.uint32), expect.view(np.uint32)), n
        assert np.array_equal(got, np.sort(f))  # and it is numpy's value order (no NaNs here)
        u = oracle.generate_random(n, 31 + n, 0xFFFFFFFF)
        b0 = to_dev(u, dev)
        handle.multi_sort_typed(b0, scratch_like(b0), None, pc, capi.KEY_U32)
        assert np.array_equal(to_host(b0), np.sort(u))
    handle.check_device_error()
