"""Probe (2+ GPUs): torch symmetric memory on this box -- P2P buffer pointers, NVSwitch multicast support, barrier under CUDA-graph
capture, and the library's own multimem / one-shot all-reduce timings on the student's flat gradient size (reference points for the
fused all-reduce + SGD kernel).   torchrun --nproc-per-node N tools/symm_probe.py"""
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as sm

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 13_100_000
t = sm.empty(n, dtype=torch.float32, device="cuda")
t.fill_(rank + 1.0)
h = sm.rendezvous(t, dist.group.WORLD)
if rank == 0:
    print("world", h.world_size, "buffer_ptrs", [hex(p) for p in h.buffer_ptrs], "multicast_ptr", hex(h.multicast_ptr),
          "signal_pad_size", h.signal_pad_size, flush=True)
    try:
        print("has_multicast_support", sm._SymmetricMemory.has_multicast_support(torch.device("cuda").type, local), flush=True)
    except Exception as e:
        print("has_multicast_support raised", e, flush=True)
h.barrier(channel=0)
torch.cuda.synchronize()


def timeit(name, fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print("%-40s %.3f ms" % (name, e0.elapsed_time(e1) / iters), flush=True)


timeit("nccl all_reduce 52 MB", lambda: dist.all_reduce(t))
timeit("barrier", lambda: h.barrier(channel=0))
for op in ("multimem_all_reduce_", "one_shot_all_reduce", "two_shot_all_reduce_"):
    try:
        f = getattr(torch.ops.symm_mem, op)
        timeit("symm_mem." + op, lambda: f(t, "sum", dist.group.WORLD.group_name))
    except Exception as e:
        if rank == 0:
            print("symm_mem.%s failed: %s" % (op, str(e)[:200]), flush=True)
# capture the barrier in a CUDA graph
try:
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        h.barrier(channel=1)
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            h.barrier(channel=1)
            t.mul_(1.0)
            h.barrier(channel=1)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
    if rank == 0:
        print("barrier under graph capture: ok", flush=True)
except Exception as e:
    if rank == 0:
        print("barrier under graph capture failed:", str(e)[:300], flush=True)
dist.barrier()
dist.destroy_process_group()
