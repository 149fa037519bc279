"""One-process-per-GPU global sort: bucket exchange around the local B200 sort.

The reference is single-device; BASELINE.json config 5 defines this extension:
N GPUs hold n keys each, the result is globally sorted with rank r holding the r-th key range.

    1. key range      vkrs_key_range + all-reduce(min, max)            -> (key_base, shift)
    2. partition      vkrs_partition: bucket(key) = min(255, (key - key_base) >> shift); the local
                      keys come out stably grouped by bucket, with the 256 bucket counts
    3. plan           all-gather of the counts; contiguous bucket ranges are dealt to the ranks so
                      that the loads are as even as the bucket granularity allows (pure host
                      arithmetic: plan_exchange)
    4. exchange       ONE all-to-all-v (torch.distributed.all_to_all_single; NCCL over NVLink /
                      NVSwitch on GPUs): the slices are already contiguous, nothing is packed
    5. local sort     vkrs_multi_sort of what arrived.  Stable end to end: pieces arrive in source-rank
                      order, each in original order, and the LSD sort keeps ties in place.

Partition-first (instead of sort -> exchange -> merge) needs no merge kernel and moves every key
across NVLink at most once.  Buckets follow the OCCUPIED key range, so the reference's 28-bit
distribution (keys < 2^28, MultiRadixSort.cpp:126) balances as well as full-range keys.

Device work goes through an `ops` object: DeviceOps (the C-ABI, default) or a test double that the
CPU `gloo` tests provide.  There is no CPU fallback in this module.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

NUM_BUCKETS = 256


def choose_bucket_map(key_min: int, key_max: int) -> tuple[int, int]:
    """(key_base, shift) such that (key - key_base) >> shift is in [0, 255] for every key."""
    if key_max < key_min:  # no keys anywhere
        return 0, 0
    span = key_max - key_min
    shift = max(0, span.bit_length() - 8)
    return key_min, shift


@dataclass
class ExchangePlan:
    boundaries: list  # world+1 bucket indices: rank r owns buckets [boundaries[r], boundaries[r+1])
    send_counts: list  # keys this rank sends to each rank
    recv_counts: list  # keys this rank receives from each rank
    imbalance: float   # max load / mean load over the ranks


def plan_exchange(all_counts: np.ndarray, rank: int) -> ExchangePlan:
    """all_counts[r][b] = keys of rank r in bucket b.  Deals contiguous bucket ranges to the ranks:
    rank r's range ends at the first bucket where the running total reaches (r+1)/world of all keys."""
    all_counts = np.asarray(all_counts, dtype=np.int64)
    world = all_counts.shape[0]
    per_bucket = all_counts.sum(axis=0)
    total = int(per_bucket.sum())
    cum = np.cumsum(per_bucket)
    boundaries = [0]
    for r in range(1, world):
        target = total * r / world
        b = 0
        if total > 0:
            idx = int(np.searchsorted(cum, target, side="left"))  # first bucket whose running total reaches the target
            idx = min(idx, NUM_BUCKETS - 1)
            before = int(cum[idx - 1]) if idx > 0 else 0
            # cut before or after that bucket, whichever lands closer to the target
            b = idx if (target - before) <= (int(cum[idx]) - target) else idx + 1
        boundaries.append(min(NUM_BUCKETS, max(boundaries[-1], b)))  # monotone; empty ranges are allowed
    boundaries.append(NUM_BUCKETS)
    send = [int(all_counts[rank, boundaries[d]:boundaries[d + 1]].sum()) for d in range(world)]
    recv = [int(all_counts[s, boundaries[rank]:boundaries[rank + 1]].sum()) for s in range(world)]
    loads = [int(per_bucket[boundaries[d]:boundaries[d + 1]].sum()) for d in range(world)]
    mean = total / world if total else 1.0
    return ExchangePlan(boundaries, send, recv, (max(loads) / mean) if total else 1.0)


class DeviceOps:
    """The product path: every step is a C-ABI call on this rank's GPU."""

    def __init__(self, handle, device):
        import torch

        self.torch = torch
        self.handle = handle
        self.device = device
        self.minmax = torch.zeros(2, dtype=torch.int32, device=device)
        self.counts = torch.zeros(NUM_BUCKETS, dtype=torch.int32, device=device)

    def key_range(self, keys, n):
        """-> int64 tensor [min, max] (unsigned values) on the device."""
        self.handle.key_range(keys, n, self.minmax)
        return self.minmax.to(self.torch.int64) & 0xFFFFFFFF

    def partition(self, keys_in, keys_out, n, key_base, shift, values_in=None, values_out=None):
        """-> int32 tensor [256] of bucket counts on the device; keys_out grouped by bucket."""
        self.handle.partition(keys_in, keys_out, n, key_base, shift, self.counts, values_in, values_out)
        return self.counts

    def local_sort(self, buf0, buf1, n, val0=None, val1=None):
        from . import capi

        pc = capi.multi_push_constants(n, 32)
        if val0 is not None:
            self.handle.multi_sort_pairs(buf0, buf1, val0, val1, None, pc)
        else:
            self.handle.multi_sort(buf0, buf1, None, pc)

    def empty(self, n):
        return self.torch.empty(max(1, n), dtype=self.torch.int32, device=self.device)


class DistributedSorter:
    """Keys are int32-typed torch tensors holding uint32 bit patterns (torch has no uint32 math)."""

    def __init__(self, handle, n_local: int, world: int, rank: int, device, pairs: bool = False, ops=None,
                 group=None, slack: float = 1.25):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.world, self.rank, self.group = world, rank, group
        self.ops = ops if ops is not None else DeviceOps(handle, device)
        self.pairs = pairs
        self.capacity = int(n_local * slack) + 1024
        self.recv = [self.ops.empty(self.capacity), self.ops.empty(self.capacity)]
        self.recv_vals = [self.ops.empty(self.capacity), self.ops.empty(self.capacity)] if pairs else None
        self.last_plan: ExchangePlan | None = None
        self.exchange_bytes = 0

    def _ensure(self, n):
        if n > self.capacity:
            self.capacity = int(n * 1.1) + 1024
            self.recv = [self.ops.empty(self.capacity), self.ops.empty(self.capacity)]
            if self.pairs:
                self.recv_vals = [self.ops.empty(self.capacity), self.ops.empty(self.capacity)]

    def sort(self, keys, scratch, values=None, values_scratch=None):
        """keys/scratch: n-element device buffers of this rank (scratch is overwritten).  Returns the
        sorted keys this rank owns after the exchange (a view into an internal buffer), or
        (keys, values) when payloads are given."""
        torch, dist = self.torch, self.dist
        n = int(keys.numel())
        if self.world == 1:
            self.ops.local_sort(keys, scratch, n, values, values_scratch)
            self.last_plan = None
            return (keys, values) if values is not None else keys

        # 1. occupied key range over all ranks
        mm = self.ops.key_range(keys, n)
        t = torch.stack([mm[0], -mm[1]])
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        kmin, neg_kmax = (int(x) for x in t.tolist())
        key_base, shift = choose_bucket_map(kmin, -neg_kmax)

        # 2. stable partition by bucket
        counts = self.ops.partition(keys, scratch, n, key_base, shift, values, values_scratch)

        # 3. plan from everybody's bucket counts
        gathered = torch.empty(self.world * NUM_BUCKETS, dtype=counts.dtype, device=counts.device)
        dist.all_gather_into_tensor(gathered, counts, group=self.group)
        plan = plan_exchange(gathered.view(self.world, NUM_BUCKETS).cpu().numpy(), self.rank)
        self.last_plan = plan
        total_recv = sum(plan.recv_counts)
        self._ensure(total_recv)

        # 4. one all-to-all-v; the send slices are contiguous in `scratch`
        recv = self.recv[0][:total_recv]
        dist.all_to_all_single(recv, scratch[:n], plan.recv_counts, plan.send_counts, group=self.group)
        self.exchange_bytes = 4 * (n - plan.send_counts[self.rank])
        recv_v = None
        if values is not None:
            recv_v = self.recv_vals[0][:total_recv]
            dist.all_to_all_single(recv_v, values_scratch[:n], plan.recv_counts, plan.send_counts, group=self.group)

        # 5. local sort of this rank's key range
        if total_recv > 0:
            self.ops.local_sort(recv, self.recv[1][:total_recv], total_recv,
                                recv_v, self.recv_vals[1][:total_recv] if recv_v is not None else None)
        return (recv, recv_v) if values is not None else recv
