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
    for case, n, hi in (("u32", 1_000_003, 1 << 32), ("ref28", 777_777, 1 << 28), ("dups", 300_000, 7)):
        rng = np.random.default_rng(17 + rank)
        keys = rng.integers(0, hi, size=n, dtype=np.uint64).astype(np.uint32)
        vals = (np.arange(n, dtype=np.uint32) + np.uint32(rank * n))
        for pairs, p2p in ((False, False), (True, False), (False, True), (True, True)):
            sorter = DistributedSorter(h, n, world, rank, dev, pairs=pairs, p2p=p2p)
            k = torch.from_numpy(keys.view(np.int32).copy()).to(dev)
            v = torch.from_numpy(vals.view(np.int32).copy()).to(dev)
            if pairs:
                out_k, out_v = sorter.sort(k, torch.empty_like(k), v, torch.empty_like(v))
            else:
                out_k, out_v = sorter.sort(k, torch.empty_like(k)), None
            h.check_device_error()
            assert sorter.used_p2p == p2p, (case, pairs, p2p)
            # gather everything on rank 0 and compare with numpy's stable sort of the concatenation
            sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
            dist.all_gather(sizes, torch.tensor([out_k.numel()], dtype=torch.int64, device=dev))
            sizes = [int(s) for s in sizes]
            pad = max(sizes)

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
            assert np.array_equal(got_k, all_k[order]), (case, pairs, p2p)
            if pairs:
                assert np.array_equal(gather(out_v), gather_in(vals)[order]), (case, "values")
            if case != "dups":
                assert max(sizes) / (sum(sizes) / world) < 1.05, sizes
            sorter.close()
    h.close()
    dist.barrier()
    if rank == 0:
        print("DIST_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
