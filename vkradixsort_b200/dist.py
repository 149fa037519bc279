"""One-process-per-GPU global sort: bucket exchange around the local B200 sort.

The reference is single-device; BASELINE.json config 5 defines this extension:
N GPUs hold n keys each, the result is globally sorted with rank r holding the r-th key range.

    1. bucket map     bucket(key) = min(255, (key - key_base) >> shift).  First try the top byte
                      (key_base 0, shift 24: valid for any input); only if that balances badly
                      (narrow or skewed keys) vkrs_key_range + all-reduce(min, max) map the OCCUPIED
                      key range onto the 256 buckets
    2. partition      vkrs_partition: the local keys come out stably grouped by bucket, with the 256
                      bucket counts
    3. plan           all-gather of the counts; contiguous bucket ranges are dealt to the ranks so
                      that the loads are as even as the bucket granularity allows (pure host
                      arithmetic: plan_exchange)
    4. exchange       fused (default on GPUs of one node): the partition kernel itself stores every
                      bucket straight into the owning rank's receive buffer through CUDA-IPC peer
                      mappings (vkrs_partition_count -> plan -> vkrs_partition_scatter_p2p), so the keys
                      cross NVLink while other tiles are still being ranked and no separate collective
                      moves data; or staged: vkrs_partition + ONE all-to-all-v (all_to_all_single; the
                      slices are already contiguous, nothing is packed) -- the portable path, also used
                      when a receive buffer would overflow
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


def destination_offsets(all_counts: np.ndarray, boundaries: list, rank: int) -> tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """For the fused exchange, per bucket b: (owner[b], offset[b], first[b], end[b]) = the rank that owns
    b, the element of that rank's receive buffer where THIS rank's part starts, and the owner's bucket
    range [first, end).  The receive buffer of rank r is laid out source rank by source rank; inside a
    source's part the keys come tile by tile in the sender's order (the receiver sorts anyway)."""
    all_counts = np.asarray(all_counts, dtype=np.int64)
    world = all_counts.shape[0]
    owner = np.zeros(NUM_BUCKETS, dtype=np.int64)
    offset = np.zeros(NUM_BUCKETS, dtype=np.int64)
    first = np.zeros(NUM_BUCKETS, dtype=np.int64)
    end = np.zeros(NUM_BUCKETS, dtype=np.int64)
    for r in range(world):
        lo, hi = boundaries[r], boundaries[r + 1]
        owner[lo:hi] = r
        offset[lo:hi] = int(all_counts[:rank, lo:hi].sum())  # parts of the lower source ranks
        first[lo:hi] = lo
        end[lo:hi] = hi
    return owner, offset, first, end


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
        b = 0
        if total > 0:
            # first bucket whose running total reaches total * r / world -- in integers, exactly as the device-side
            # plan (vkrs_exchange.cuh: exchange_plan_kernel) computes it
            tr = total * r
            idx = int(np.searchsorted(cum * world, tr, side="left"))
            idx = min(idx, NUM_BUCKETS - 1)
            before = int(cum[idx - 1]) if idx > 0 else 0
            # cut before or after that bucket, whichever lands closer to the target
            b = idx if (tr - before * world) <= (int(cum[idx]) * world - tr) else idx + 1
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

    def partition_count(self, keys_in, n, key_base, shift, with_values=False):
        self.handle.partition_count(keys_in, n, key_base, shift, self.counts, with_values)
        return self.counts

    def partition_scatter_p2p(self, keys_in, n, key_base, shift, dst_tables, values_in=None, gate=None):
        self.handle.partition_scatter_p2p(keys_in, n, key_base, shift, dst_tables, values_in, gate)

    def local_sort(self, buf0, buf1, n, val0=None, val1=None, key_span=None):
        """key_span = (lo, hi): the key range this rank owns after the exchange (a hint for the local sort)."""
        from . import capi

        pc = capi.multi_push_constants(n, 32)
        if val0 is not None:
            self.handle.multi_sort_pairs(buf0, buf1, val0, val1, None, pc)
        else:
            if key_span is not None:
                self.handle.set_key_span_hint(*key_span)
            self.handle.multi_sort(buf0, buf1, None, pc)
            if key_span is not None:
                self.handle.set_key_span_hint()

    def empty(self, n):
        return self.torch.empty(max(1, n), dtype=self.torch.int32, device=self.device)


class DistributedSorter:
    """Keys are int32-typed torch tensors holding uint32 bit patterns (torch has no uint32 math).

    p2p=None picks the fused peer-to-peer exchange when the device path is in use and world > 1."""

    def __init__(self, handle, n_local: int, world: int, rank: int, device, pairs: bool = False, ops=None,
                 group=None, slack: float = 1.25, p2p: bool | None = None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.world, self.rank, self.group = world, rank, group
        self.handle, self.device = handle, device
        self.ops = ops if ops is not None else DeviceOps(handle, device)
        self.pairs = pairs
        self.p2p = (ops is None and world > 1) if p2p is None else p2p
        self.capacity = int(n_local * slack) + 1024
        if world > 1 and dist.is_initialized():
            # every rank must size (and later re-size) its receive buffers identically: decisions that depend on the
            # capacity are taken without communication
            t = torch.tensor([self.capacity], dtype=torch.int64, device=device if dist.get_backend(group) == "nccl" else "cpu")
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
            self.capacity = int(t.item())
        self.last_plan: ExchangePlan | None = None
        self._epoch = 0
        self._speculate = False  # the device-planned exchange is used once a host-planned sort has shown that the top-byte buckets balance
        self.exchange_bytes = 0
        self.used_p2p = False
        self.used_key_range = False
        self.max_imbalance = 1.10  # above this the top-byte buckets are replaced by buckets over the occupied key range
        self._ipc_owned, self._ipc_opened = [], []
        self._alloc()

    # ---- buffers ----
    def _alloc(self):
        self.recv = [self.ops.empty(self.capacity), self.ops.empty(self.capacity)]
        self.recv_vals = [self.ops.empty(self.capacity), self.ops.empty(self.capacity)] if self.pairs else None
        if self.p2p:
            self._alloc_p2p()

    def _share(self, nbytes):
        """cudaMalloc + CUDA IPC: a buffer of this rank every rank of the node can store into.  Returns
        (own pointer, [the buffer of rank r as this process sees it])."""
        torch, dist = self.torch, self.dist
        ptr, hd = self.handle.ipc_alloc(nbytes)
        self._ipc_owned.append(ptr)
        mine = torch.frombuffer(bytearray(hd), dtype=torch.uint8).to(self.device)
        everyone = torch.empty(self.world * 64, dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(everyone, mine, group=self.group)
        handles = everyone.cpu().numpy().tobytes()
        ptrs = []
        for r in range(self.world):
            if r == self.rank:
                ptrs.append(ptr)
            else:
                p = self.handle.ipc_open(handles[64 * r:64 * (r + 1)])
                self._ipc_opened.append(p)
                ptrs.append(p)
        return ptr, ptrs

    def _alloc_p2p(self):
        """Receive buffers (and the flag array of the peer barrier) every rank of the node can store into, wrapped as
        torch tensors for the local sort.  Collective: every rank calls it at the same point."""
        torch = self.torch
        self._release_p2p()
        self.peer_ptrs, self.peer_ptrs_dev = [], []
        for which in range(2 if self.pairs else 1):  # keys, payloads
            ptr, ptrs = self._share(4 * self.capacity)
            self.peer_ptrs.append(ptrs)
            self.peer_ptrs_dev.append(torch.tensor(ptrs, dtype=torch.int64, device=self.device))
            t = _tensor_from_pointer(torch, ptr, self.capacity, self.device)
            if which == 0:
                self.recv[0] = t
            else:
                self.recv_vals[0] = t
        fptr, fptrs = self._share(4 * 64)
        self.flags = _tensor_from_pointer(torch, fptr, 64, self.device)
        self.flags.zero_()
        self.peer_flags_dev = torch.tensor(fptrs, dtype=torch.int64, device=self.device)
        self._epoch = 0
        torch.cuda.synchronize()
        self.dist.barrier(group=self.group)  # every flag array is zero before anybody signals
        if not hasattr(self, "_gathered"):  # (kept across re-allocations: a sort in progress holds its counts here)
            self.dst_tables = torch.zeros(4 * NUM_BUCKETS, dtype=torch.int64, device=self.device)
            self.summary = torch.zeros(4 + 64, dtype=torch.int32, device=self.device)
            self._gathered = torch.zeros(self.world * NUM_BUCKETS, dtype=torch.int32, device=self.device)
            self._counts_host = torch.zeros(self.world * NUM_BUCKETS, dtype=torch.int32).pin_memory()
            self._summary_host = torch.zeros(4 + 64, dtype=torch.int32).pin_memory()
            self._counts_event = torch.cuda.Event()

    def _release_p2p(self):
        if self._ipc_opened or self._ipc_owned:
            self.torch.cuda.synchronize()
            if self.world > 1:
                self.dist.barrier(group=self.group)  # nobody may still be storing into a buffer that goes away
        for p in self._ipc_opened:
            self.handle.ipc_close(p)
        for p in self._ipc_owned:
            self.handle.ipc_free(p)
        self._ipc_opened, self._ipc_owned = [], []

    def close(self):
        if self.p2p:
            self._release_p2p()

    def _ensure(self, n):
        if n > self.capacity:
            self.capacity = int(n * 1.1) + 1024
            self._alloc()

    # ---- the sort ----
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

        def count_and_gather(key_base, shift):
            # bucket counts (fused exchange) or the whole stable partition (staged exchange), then
            # everybody's counts.  The all-gather also orders this step after every rank's previous
            # local sort, so the receive buffers are free to be overwritten.
            if self.p2p:
                counts = self.ops.partition_count(keys, n, key_base, shift, values is not None)
                gathered = self._gathered
            else:
                counts = self.ops.partition(keys, scratch, n, key_base, shift, values, values_scratch)
                gathered = torch.empty(self.world * NUM_BUCKETS, dtype=counts.dtype, device=counts.device)
            dist.all_gather_into_tensor(gathered, counts, group=self.group)
            return gathered

        def fused_exchange(gathered, key_base, shift, gate):
            """plan on the device -> partition kernel that stores into the owners' receive buffers -> peer barrier;
            nothing here waits for the host.  gate: the plan kernel decides whether the scatter runs (balance, capacity)
            and the decision is copied to the host between the two kernels."""
            self.handle.exchange_plan(gathered, self.world, self.rank, self.peer_ptrs_dev[0],
                                      self.peer_ptrs_dev[1] if values is not None else None, self.dst_tables, self.summary,
                                      self.capacity, int(round(self.max_imbalance * 1000)) if gate else 0)
            if gate:
                self._summary_host.copy_(self.summary, non_blocking=True)
                self._counts_event.record()
            self.ops.partition_scatter_p2p(keys, n, key_base, shift, self.dst_tables, values, self.summary[2:3] if gate else None)
            self._epoch += 1
            self.handle.peer_barrier(self.flags, self.peer_flags_dev, self.world, self.rank, self._epoch)

        # 1+2+3. Speculate that the keys use the full 32-bit range: bucket = top byte is valid for ANY
        # input (it only may balance badly), and it saves the key-range pass and one host round trip.
        key_base, shift = 0, 24
        self.used_key_range = False
        gathered = count_and_gather(key_base, shift)
        speculated = False
        if self.p2p and self._speculate:
            # The last sort balanced with the top-byte buckets and fitted the buffers: run the whole exchange on the
            # device right away and look at the counts on the host WHILE it runs.  The plan kernel itself closes the
            # gate of the scatter if this input does not balance or would overflow a receive buffer (every rank
            # sees the same counts, so every rank takes the same decision); then the host-planned path below repeats it.
            self._counts_host.copy_(gathered, non_blocking=True)
            fused_exchange(gathered, key_base, shift, gate=True)
            self._counts_event.synchronize()  # counts and the plan kernel's verdict are on the host; the scatter is running
            all_counts = self._counts_host.numpy().reshape(self.world, NUM_BUCKETS).astype(np.int64)
            speculated = int(self._summary_host[2]) != 0  # closed on every rank alike: nothing has been stored anywhere
            self._speculate = speculated
        else:
            all_counts = gathered.view(self.world, NUM_BUCKETS).cpu().numpy()
        plan = plan_exchange(all_counts, self.rank)
        largest = max(int(all_counts[:, plan.boundaries[r]:plan.boundaries[r + 1]].sum()) for r in range(self.world))
        if not speculated and plan.imbalance > self.max_imbalance:
            # narrow or skewed keys: map the OCCUPIED range onto the 256 buckets and count again
            mm = self.ops.key_range(keys, n)
            t = torch.stack([mm[0], -mm[1]])
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
            kmin, neg_kmax = (int(x) for x in t.tolist())
            key_base, shift = choose_bucket_map(kmin, -neg_kmax)
            all_counts = count_and_gather(key_base, shift).view(self.world, NUM_BUCKETS).cpu().numpy()
            self.used_key_range = True
            plan = plan_exchange(all_counts, self.rank)
            largest = max(int(all_counts[:, plan.boundaries[r]:plan.boundaries[r + 1]].sum()) for r in range(self.world))
        self.last_plan = plan
        total_recv = sum(plan.recv_counts)
        self.exchange_bytes = 4 * (n - plan.send_counts[self.rank]) * (2 if values is not None else 1)
        if largest > self.capacity:
            # `largest` and `capacity` are the same numbers on every rank: the re-allocation (a collective when the
            # buffers are shared between the ranks) is entered by all of them together
            self._ensure(largest)
        fused = self.p2p
        self.used_p2p = fused
        recv = self.recv[0][:total_recv]
        recv_v = self.recv_vals[0][:total_recv] if values is not None else None

        # 4. exchange
        if fused:
            if not speculated:
                fused_exchange(self._gathered, key_base, shift, gate=False)  # the counts of the bucket map that was settled on
                self._speculate = not self.used_key_range
        else:
            dist.all_to_all_single(recv, scratch[:n], plan.recv_counts, plan.send_counts, group=self.group)
            if values is not None:
                dist.all_to_all_single(recv_v, values_scratch[:n], plan.recv_counts, plan.send_counts, group=self.group)

        # 5. local sort of this rank's key range
        if total_recv > 0:
            # the keys this rank received: buckets [b_lo, b_hi) of bucket(key) = min(255, (key - key_base) >> shift)
            b_lo, b_hi = int(plan.boundaries[self.rank]), int(plan.boundaries[self.rank + 1])
            span_lo = min(0xFFFFFFFF, key_base + (b_lo << shift))
            span_hi = 0xFFFFFFFF if b_hi >= NUM_BUCKETS else min(0xFFFFFFFF, key_base + (b_hi << shift) - 1)
            if b_lo == 0:
                span_lo = 0  # keys below key_base cannot occur, but the hint need not rely on it
            self.ops.local_sort(recv, self.recv[1][:total_recv], total_recv,
                                recv_v, self.recv_vals[1][:total_recv] if recv_v is not None else None,
                                key_span=(span_lo, max(span_lo, span_hi)))
        return (recv, recv_v) if values is not None else recv


def _tensor_from_pointer(torch, ptr: int, numel: int, device):
    """int32 torch tensor over `numel` elements of device memory owned elsewhere (no copy, no free)."""

    class _Ext:
        pass

    ext = _Ext()
    ext.__cuda_array_interface__ = {"shape": (numel,), "typestr": "<i4", "data": (ptr, False), "version": 3}
    return torch.as_tensor(ext, device=device)
