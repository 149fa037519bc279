"""CPU tests of the multi-GPU bucket exchange (vkradixsort_b200/dist.py): the host arithmetic, and the
whole exchange protocol on 2 `gloo` ranks with a numpy test double standing in for the three device
calls (key range, stable partition, local sort) -- the GPU kernels themselves are covered by the -m gpu
tests; nothing here is a product fallback."""
import os
import socket

import numpy as np
import pytest

from vkradixsort_b200 import dist as D


def test_choose_bucket_map():
    assert D.choose_bucket_map(0, 0xFFFFFFFF) == (0, 24)
    assert D.choose_bucket_map(0, 0x0FFFFFFF) == (0, 20)  # the reference's 28-bit range
    assert D.choose_bucket_map(1000, 1000) == (1000, 0)
    assert D.choose_bucket_map(5, 260) == (5, 0)
    assert D.choose_bucket_map(5, 261) == (5, 1)
    assert D.choose_bucket_map(10, 5) == (0, 0)  # empty input everywhere
    for lo, hi in ((0, 255), (7, 0xFFFFFFFF), (123456, 987654321), (0x80000000, 0x80000001)):
        base, shift = D.choose_bucket_map(lo, hi)
        assert 0 <= (hi - base) >> shift <= 255 and (lo - base) >> shift == 0


def test_plan_exchange_balances_and_is_consistent():
    rng = np.random.default_rng(3)
    for world in (1, 2, 4, 8):
        counts = rng.integers(0, 1000, size=(world, 256))
        plans = [D.plan_exchange(counts, r) for r in range(world)]
        b = plans[0].boundaries
        assert b[0] == 0 and b[-1] == 256 and all(x <= y for x, y in zip(b, b[1:]))
        assert all(p.boundaries == b for p in plans)
        for r in range(world):
            assert sum(plans[r].send_counts) == counts[r].sum()
            for d in range(world):
                assert plans[r].send_counts[d] == plans[d].recv_counts[r]
        assert plans[0].imbalance < 1.05
    # everything in one bucket: one rank takes it all, the others get empty ranges
    counts = np.zeros((4, 256), dtype=np.int64)
    counts[:, 17] = 100
    totals = sorted(sum(D.plan_exchange(counts, r).recv_counts) for r in range(4))
    assert totals == [0, 0, 0, 400]
    # no keys at all
    p = D.plan_exchange(np.zeros((2, 256), dtype=np.int64), 0)
    assert p.send_counts == [0, 0] and p.recv_counts == [0, 0]


def test_destination_offsets():
    """Fused exchange: every (source, owner) part of a receive buffer starts where the parts of the lower
    source ranks end, and the parts tile the buffer exactly."""
    rng = np.random.default_rng(5)
    world = 4
    counts = rng.integers(0, 50, size=(world, 256))
    counts[:, 100:140] = 0  # empty buckets
    plans = [D.plan_exchange(counts, r) for r in range(world)]
    b = plans[0].boundaries
    for s_ in range(world):
        owner, offset, first, end = D.destination_offsets(counts, b, s_)
        for bucket in range(256):
            r = int(owner[bucket])
            assert b[r] <= bucket < b[r + 1] and (int(first[bucket]), int(end[bucket])) == (b[r], b[r + 1])
            assert int(offset[bucket]) == sum(plans[r].recv_counts[:s_])
    for r in range(world):
        assert sum(plans[r].recv_counts) == int(counts[:, b[r]:b[r + 1]].sum())


class NumpyOps:
    """Test double for DeviceOps on CPU tensors: same contracts as the C-ABI calls it replaces."""

    def __init__(self):
        import torch

        self.torch = torch

    @staticmethod
    def _u32(t):
        return t.numpy().view(np.uint32)

    def key_range(self, keys, n):
        k = self._u32(keys)[:n]
        lo, hi = (int(k.min()), int(k.max())) if n else (0xFFFFFFFF, 0)
        return self.torch.tensor([lo, hi], dtype=self.torch.int64)

    def partition(self, keys_in, keys_out, n, key_base, shift, values_in=None, values_out=None):
        k = self._u32(keys_in)[:n]
        bucket = np.minimum(255, (k - np.uint32(key_base)) >> np.uint32(shift))
        order = np.argsort(bucket, kind="stable")
        self._u32(keys_out)[:n] = k[order]
        if values_in is not None:
            self._u32(values_out)[:n] = self._u32(values_in)[:n][order]
        return self.torch.from_numpy(np.bincount(bucket, minlength=256).astype(np.int32))

    def local_sort(self, buf0, buf1, n, val0=None, val1=None, key_span=None):
        k = self._u32(buf0)[:n]
        if key_span is not None and n:  # the hint handed to the device sort must cover what this rank received
            assert key_span[0] <= int(k.min()) and int(k.max()) <= key_span[1], (key_span, int(k.min()), int(k.max()))
        order = np.argsort(k, kind="stable")
        if val0 is not None:
            self._u32(val0)[:n] = self._u32(val0)[:n][order]
        k[:] = k[order]

    def empty(self, n):
        return self.torch.zeros(max(1, n), dtype=self.torch.int32)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, out_dir):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 20_000
        rng = np.random.default_rng(100 + rank)
        if case == "u32":
            keys = rng.integers(0, 1 << 32, size=n, dtype=np.uint32)
        elif case == "ref28":
            keys = rng.integers(0, 1 << 28, size=n, dtype=np.uint32)
        elif case == "dups":
            keys = rng.integers(0, 5, size=n, dtype=np.uint32) * np.uint32(0x01000193)
        else:  # "skewed": rank 0 holds small keys only, rank 1 large ones
            keys = (rng.integers(0, 1 << 16, size=n, dtype=np.uint32) + np.uint32(rank * 0xF0000000))
        vals = (np.arange(n, dtype=np.uint32) + np.uint32(rank * n))  # global original index
        sorter = D.DistributedSorter(None, n, world, rank, torch.device("cpu"), pairs=True, ops=NumpyOps())
        k_t = torch.from_numpy(keys.view(np.int32).copy())
        v_t = torch.from_numpy(vals.view(np.int32).copy())
        out_k, out_v = sorter.sort(k_t, torch.zeros_like(k_t), v_t, torch.zeros_like(v_t))
        np.savez(os.path.join(out_dir, f"{case}_{rank}.npz"), keys=keys, vals=vals,
                 out_k=out_k.numpy().view(np.uint32), out_v=out_v.numpy().view(np.uint32),
                 boundaries=np.array(sorter.last_plan.boundaries))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["u32", "ref28", "dups", "skewed"])
def test_two_rank_exchange_on_gloo(tmp_path, case):
    import torch.multiprocessing as mp

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), case, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"{case}_{r}.npz") for r in range(world)]
    all_keys = np.concatenate([p["keys"] for p in parts])
    all_vals = np.concatenate([p["vals"] for p in parts])
    order = np.argsort(all_keys, kind="stable")  # ranks hold consecutive chunks: global stable order
    got_k = np.concatenate([p["out_k"] for p in parts])
    got_v = np.concatenate([p["out_v"] for p in parts])
    assert np.array_equal(got_k, all_keys[order])
    assert np.array_equal(got_v, all_vals[order])
    assert np.array_equal(parts[0]["boundaries"], parts[1]["boundaries"])
    if case in ("u32", "ref28"):
        sizes = [len(p["out_k"]) for p in parts]
        assert max(sizes) / (sum(sizes) / world) < 1.05  # buckets follow the occupied range
