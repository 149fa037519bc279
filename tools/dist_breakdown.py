"""Stage-by-stage wall time of the bucket exchange (run under torchrun, one rank per GPU)."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vkradixsort_b200 import Handle, capi
from vkradixsort_b200 import dist as D

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
h = Handle(local, n)
keys = torch.from_numpy(np.random.default_rng(rank).integers(0, 1 << 32, size=n, dtype=np.uint32).view(np.int32)).to(dev)
buf0, buf1 = torch.empty_like(keys), torch.empty_like(keys)
sorter = D.DistributedSorter(h, n, world, rank, dev)
ops = sorter.ops
def sync(): torch.cuda.synchronize()
acc = {}
for it in range(6):
    buf0.copy_(keys); dist.barrier(); sync()
    t = [time.perf_counter()]
    mm = ops.key_range(buf0, n); sync(); t.append(time.perf_counter())
    tt = torch.stack([mm[0], -mm[1]]); dist.all_reduce(tt, op=dist.ReduceOp.MIN); kmin, nk = (int(x) for x in tt.tolist()); t.append(time.perf_counter())
    base, shift = D.choose_bucket_map(kmin, -nk)
    counts = ops.partition(buf0, buf1, n, base, shift); sync(); t.append(time.perf_counter())
    g = torch.empty(world * 256, dtype=counts.dtype, device=dev); dist.all_gather_into_tensor(g, counts)
    plan = D.plan_exchange(g.view(world, 256).cpu().numpy(), rank); t.append(time.perf_counter())
    tot = sum(plan.recv_counts); recv = sorter.recv[0][:tot]
    dist.all_to_all_single(recv, buf1[:n], plan.recv_counts, plan.send_counts); sync(); t.append(time.perf_counter())
    ops.local_sort(recv, sorter.recv[1][:tot], tot); sync(); t.append(time.perf_counter())
    if it >= 2:
        for name, a, b in zip(["key_range", "allreduce+sync", "partition", "allgather+plan", "all_to_all", "local_sort"], t, t[1:]):
            acc[name] = acc.get(name, 0) + (b - a) * 1e3 / 4
if rank == 0:
    print("staged", {k: round(v, 3) for k, v in acc.items()}, "total", round(sum(acc.values()), 3), "ms; sent MB", 4 * (n - plan.send_counts[rank]) / 1e6)
# fused peer-to-peer exchange
sorter.close()
sorter = D.DistributedSorter(h, n, world, rank, dev, p2p=True)
ops = sorter.ops
acc = {}
for it in range(6):
    buf0.copy_(keys); dist.barrier(); sync()
    t = [time.perf_counter()]
    mm = ops.key_range(buf0, n); sync(); t.append(time.perf_counter())
    tt = torch.stack([mm[0], -mm[1]]); dist.all_reduce(tt, op=dist.ReduceOp.MIN); kmin, nk = (int(x) for x in tt.tolist()); t.append(time.perf_counter())
    base, shift = D.choose_bucket_map(kmin, -nk)
    counts = ops.partition_count(buf0, n, base, shift); sync(); t.append(time.perf_counter())
    g = torch.empty(world * 256, dtype=counts.dtype, device=dev); dist.all_gather_into_tensor(g, counts)
    ac = g.view(world, 256).cpu().numpy(); plan = D.plan_exchange(ac, rank); t.append(time.perf_counter())
    owner, offset, first, end = D.destination_offsets(ac, plan.boundaries, rank)
    tab = sorter._dst_host.numpy(); tab[:256] = np.array(sorter.peer_ptrs[0], dtype=np.int64)[owner] + 4 * offset
    tab[512:768] = first; tab[768:1024] = end
    sorter.dst_tables.copy_(sorter._dst_host, non_blocking=True)
    ops.partition_scatter_p2p(buf0, n, base, shift, sorter.dst_tables); sync(); t.append(time.perf_counter())
    dist.barrier(); sync(); t.append(time.perf_counter())
    tot = sum(plan.recv_counts); recv = sorter.recv[0][:tot]
    ops.local_sort(recv, sorter.recv[1][:tot], tot); sync(); t.append(time.perf_counter())
    if it >= 2:
        for name, a, b in zip(["key_range", "allreduce+sync", "count", "allgather+plan", "scatter_p2p", "barrier", "local_sort"], t, t[1:]):
            acc[name] = acc.get(name, 0) + (b - a) * 1e3 / 4
ok = bool((recv[1:] ^ -(1 << 31) >= recv[:-1] ^ -(1 << 31)).all())
if rank == 0:
    print("fused ", {k: round(v, 3) for k, v in acc.items()}, "total", round(sum(acc.values()), 3), "ms; sorted:", ok)
sorter.close()
dist.destroy_process_group()
