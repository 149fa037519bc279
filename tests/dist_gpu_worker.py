"""Worker of test_two_gpu_bucket_exchange (run under torchrun, one rank per GPU)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vkradixsort_b200 import Handle  # noqa: E402
from vkradixsort_b200.dist import DistributedSorter  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    h = Handle(local, 1 << 20)

    def check(sorter, keys, vals, pairs, tag):
        n = keys.shape[0]
        k = torch.from_numpy(keys.view(np.int32).copy()).to(dev)
        v = torch.from_numpy(vals.view(np.int32).copy()).to(dev)
        if pairs:
            out_k, out_v = sorter.sort(k, torch.empty_like(k), v, torch.empty_like(v))
        else:
            out_k, out_v = sorter.sort(k, torch.empty_like(k)), None
        h.check_device_error()
        # gather everything on rank 0 and compare with numpy's stable sort of the concatenation
        sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([out_k.numel()], dtype=torch.int64, device=dev))
        sizes = [int(s) for s in sizes]
        pad = max(max(sizes), 1)

        def gather(t):
            buf = torch.zeros(pad, dtype=torch.int32, device=dev)
            buf[: t.numel()] = t
            parts = [torch.empty_like(buf) for _ in range(world)]
            dist.all_gather(parts, buf)
            return np.concatenate([p[:s].cpu().numpy().view(np.uint32) for p, s in zip(parts, sizes)])

        def gather_in(a):
            parts = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(world)]
            dist.all_gather(parts, torch.from_numpy(a.view(np.int32).copy()).to(dev))
            return np.concatenate([p.cpu().numpy().view(np.uint32) for p in parts])

        got_k, all_k = gather(out_k), gather_in(keys)
        order = np.argsort(all_k, kind="stable")
        assert np.array_equal(got_k, all_k[order]), tag
        if pairs:
            assert np.array_equal(gather(out_v), gather_in(vals)[order]), (tag, "values")
        return sizes

    cases = (("u32", 1_000_003, 1 << 32), ("ref28", 777_777, 1 << 28), ("dups", 300_000, 7),
             ("all_equal", 200_001, 1), ("three_values", 250_000, 3))
    for case, n, hi in cases:
        rng = np.random.default_rng(17 + rank)
        keys = rng.integers(0, hi, size=n, dtype=np.uint64).astype(np.uint32)
        if case == "three_values":
            keys = keys * np.uint32(0x40000001)
        vals = (np.arange(n, dtype=np.uint32) + np.uint32(rank * n))
        for pairs, p2p in ((False, False), (True, False), (False, True), (True, True)):
            sorter = DistributedSorter(h, n, world, rank, dev, pairs=pairs, p2p=p2p)
            sizes = check(sorter, keys, vals, pairs, (case, pairs, p2p, "first sort: planned on the host"))
            assert sorter.used_p2p == p2p, (case, pairs, p2p)
            if case in ("u32", "ref28"):
                assert max(sizes) / (sum(sizes) / world) < 1.05, sizes
            # again on the same sorter: when the top-byte buckets balanced, the exchange is now planned on the device
            keys2 = np.random.default_rng(99 + rank).integers(0, hi, size=n, dtype=np.uint64).astype(np.uint32)
            check(sorter, keys2, vals, pairs, (case, pairs, p2p, "second sort"))
            if p2p and case == "u32":
                assert sorter._speculate, "full-range keys balance with the top-byte buckets: the device-planned exchange should be on"
                # ... and an input that does NOT balance closes the gate of the speculative exchange on every rank
                skew = (keys2 >> np.uint32(12)).astype(np.uint32)
                check(sorter, skew, vals, pairs, (case, pairs, p2p, "third sort: speculation refused on the device"))
                assert sorter.used_key_range and not sorter._speculate
                check(sorter, keys, vals, pairs, (case, pairs, p2p, "fourth sort: back on the host-planned path"))
            sorter.close()
    h.close()
    dist.barrier()
    if rank == 0:
        print("DIST_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
