"""Probe (torchrun, >= 2 GPUs): symmetric memory rendezvous, TMA loads from PEER memory through the
tcgen05 engine, and the latency of a symmetric-memory barrier vs an NCCL all-gather."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
import multimodal_baby_b200 as m
from multimodal_baby_b200 import _cabi
b, E = 512, 512
t = symm.empty((b, 2 * E), dtype=torch.bfloat16, device=dev)
hdl = symm.rendezvous(t, dist.group.WORLD)
print(rank, "rendezvous ok; ptrs", [hex(p) for p in hdl.buffer_ptrs], "multicast_ptr", hex(hdl.multicast_ptr) if hdl.multicast_ptr else None, flush=True)
g = torch.Generator().manual_seed(rank)
t.copy_(torch.randn(b, 2 * E, generator=g).to(torch.bfloat16))
hdl.barrier()
peer_rank = (rank + 1) % world
peer = hdl.get_buffer(peer_rank, (b, 2 * E), torch.bfloat16)
# reference: NCCL all-gather
allb = torch.empty((world * b, 2 * E), dtype=torch.bfloat16, device=dev)
dist.all_gather_into_tensor(allb, t)
ref_peer = allb[peer_rank * b:(peer_rank + 1) * b]
assert torch.equal(peer, ref_peer), "peer view mismatch"
# tcgen05 GEMM with the B operand in PEER memory (TMA over NVLink)
x = t[:, :E].contiguous()
C1 = torch.empty((b, b), dtype=torch.float32, device=dev); C2 = torch.empty_like(C1)
st = torch.cuda.current_stream().cuda_stream
_cabi.call("cvcl_gemm_f32out", x.data_ptr(), E, 0, peer.data_ptr() + 2 * E, 2 * E, 0, b, b, E, 1.0, C1.data_ptr(), b, st)
_cabi.call("cvcl_gemm_f32out", x.data_ptr(), E, 0, ref_peer.data_ptr() + 2 * E, 2 * E, 0, b, b, E, 1.0, C2.data_ptr(), b, st)
torch.cuda.synchronize()
print(rank, "TMA-from-peer GEMM max diff", float((C1 - C2).abs().max()), "ref max", float(C2.abs().max()), flush=True)
# latency: symm barrier vs NCCL all-gather of the features
def timeit(fn, n=200):
    for _ in range(20): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
tb = timeit(lambda: hdl.barrier())
tg = timeit(lambda: dist.all_gather_into_tensor(allb, t))
flat = torch.zeros(2400000, device=dev)
ta = timeit(lambda: dist.all_reduce(flat), 50)
if rank == 0:
    print("us per op: symm barrier %.1f, nccl all_gather(1MB/rank) %.1f, nccl all_reduce(9.6MB) %.1f" % (tb, tg, ta), flush=True)
torch.cuda.synchronize(); dist.barrier(); os._exit(0)
