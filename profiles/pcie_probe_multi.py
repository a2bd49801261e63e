"""Concurrent D2H / H2D of 23.6 MB on every rank of a node, before and after binding the process to the CPUs of its
GPU's NUMA node (page-locked buffers are first-touched where the process runs).
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/pcie_probe_multi.py"""
import os
import time
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def numa_of_gpu(i):
    pr = torch.cuda.get_device_properties(i)
    try:
        bus = "%04x:%02x:%02x.0" % (int(pr.pci_domain_id), int(pr.pci_bus_id), int(pr.pci_device_id))
    except (AttributeError, TypeError, ValueError):
        return -1, "?"
    p = f"/sys/bus/pci/devices/{bus}/numa_node"
    if os.path.exists(p):
        return int(open(p).read().strip()), bus
    return -1, bus


def cpus_of_node(n):
    out = set()
    for part in open(f"/sys/devices/system/node/node{n}/cpulist").read().strip().split(","):
        a, _, b = part.partition("-")
        out.update(range(int(a), int(b or a) + 1))
    return out


def measure(tag):
    n = int(23.6e6 / 4)
    t = torch.randn(n, device=dev)
    buf = torch.empty(n).pin_memory()          # a fresh page-locked block (not the caching allocator's)
    buf.zero_()
    res = []
    for direction in ("D2H", "H2D"):
        for _ in range(3):
            (buf.copy_(t, non_blocking=True) if direction == "D2H" else t.copy_(buf, non_blocking=True))
            torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            (buf.copy_(t, non_blocking=True) if direction == "D2H" else t.copy_(buf, non_blocking=True))
            torch.cuda.synchronize()
        res.append(23.6e-3 / ((time.perf_counter() - t0) / 10))
    print(f"rank {rank} {tag}: D2H {res[0]:.1f} GB/s, H2D {res[1]:.1f} GB/s", flush=True)


node, bus = numa_of_gpu(local)
aff = sorted(os.sched_getaffinity(0))
print(f"rank {rank}: GPU {local} bus {bus} numa {node}; affinity {len(aff)} cpus [{aff[0]}..{aff[-1]}]; nodes "
      f"{sorted(d for d in os.listdir('/sys/devices/system/node') if d.startswith('node'))}", flush=True)
measure("unbound")
if node >= 0:
    try:
        os.sched_setaffinity(0, cpus_of_node(node) & set(aff) or cpus_of_node(node))
        measure(f"bound to node {node}")
    except OSError as e:
        print(f"rank {rank}: cannot bind: {e}", flush=True)
if world > 1:
    dist.destroy_process_group()
