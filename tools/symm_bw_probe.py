"""torchrun target (2+ GPUs): is local access to peer-mapped symmetric memory slower than to ordinary device memory,
and what does a plain push / pull over NVLink reach?  torch copies of the 9.6 MB gradient block, CUDA-event timed.
Measurement helper (gpurun), not part of the product."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
n = 2408716 + 12            # floats of [out5 | ds | db | d table | dW]
n = n // 4 * 4
s_t = symm.empty((n,), dtype=torch.float32, device=dev)
hdl = symm.rendezvous(s_t, dist.group.WORLD)
peer = hdl.get_buffer((rank + 1) % world, (n,), torch.float32)
a = torch.randn(n, device=dev); b = torch.empty(n, device=dev)
s_t.copy_(a)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, cold, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    ts = []
    for _ in range(reps):
        if cold:
            flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


tests = [("normal -> normal", lambda: b.copy_(a)), ("symm -> normal (local)", lambda: b.copy_(s_t)),
         ("normal -> symm (local)", lambda: s_t.copy_(a)), ("normal -> PEER symm (push)", lambda: peer.copy_(a)),
         ("PEER symm -> normal (pull)", lambda: b.copy_(peer)),
         ("half: normal -> PEER (push 4.8 MB)", lambda: peer[:n // 2].copy_(a[:n // 2]))]
for name, fn in tests:
    for cold in (False, True):
        t = timeit(fn, cold)
        nb = n * 4 if "half" not in name else n * 2
        if rank == 0:
            print("%-36s %s: %.1f us  (%.0f GB/s payload)" % (name, "cold" if cold else "warm", t, nb / t / 1e3), flush=True)
if rank == 0:
    print("multicast_ptr:", hex(hdl.multicast_ptr) if getattr(hdl, "multicast_ptr", 0) else None)
torch.cuda.synchronize(); dist.barrier(); os._exit(0)
