"""torchrun target: the sharded one-kernel step (csrc/fused_step.cuh, world > 1) timed in place -- CUDA-event time of
a graph replay (L2 flushed, max over ranks) and the in-kernel globaltimer stamps of CTA 0 of every rank.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/fused_sharded_probe.py [pairs_per_gpu]
Measurement helper (gpurun), not part of the product."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import statistics
import torch, torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
import multimodal_baby_b200 as m
import bench
from bench import build_model, synth_batch, S_FIXED, E, K, V, L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
_, model = build_model(dev, dist.group.WORLD)
f, ids, lens = synth_batch(1234 + rank, B)
x = torch.from_numpy(f).to(dev).to(torch.bfloat16); ids_d = torch.from_numpy(ids).to(dev); lens_d = torch.from_numpy(lens).to(dev)
fcw, fcb = model.image_embed.model.fc.weight, model.image_embed.model.fc.bias
table = model.text_embed.embedding.weight
m.ops.register_weight_shadow(fcw, fcw.detach().to(torch.bfloat16).contiguous())
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def raw():
    return m.ops.flat_step_sharded(x, ids_d, lens_d, fcw, fcb, table, S_FIXED, True, True, False, dist.group.WORLD)


side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        raw()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize(); dist.barrier()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    raw()
for mode in ("cold", "warm"):
    ts = []
    for _ in range(30):
        if mode == "cold":
            flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    t = torch.tensor([statistics.mean(ts[5:]), min(ts)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ws = m.ops._FUSED_WS.get(("sharded", lr, B, L, E, K, V, world))
    tm = ws[128:128 + 8 * 48].view(torch.int64).cpu().numpy() if ws is not None else None
    names = ["P0 head+text", "P1 norm (+features land)", "P2 similarity (+partials land)", "P3 dlogits+dQ", "P4 norm-bwd",
             "P5 dW+dtable (+tiles land)", "P6 owner sums+broadcast", "?", "?"]
    line = ""
    if tm is not None:
        prev = tm[0]
        for k in range(1, 10):
            if tm[k]:
                line += "%s %.1f | " % (names[k - 1], (tm[k] - prev) / 1e3); prev = tm[k]
        line += "total %.1f" % ((tm[15] - tm[0]) / 1e3)
        line += "\n      stamps (us from start): barriers " + " ".join("%.1f" % ((tm[k] - tm[0]) / 1e3) for k in range(1, 10) if tm[k])
        line += " | in-phase " + " ".join("%d:%.1f" % (k, (tm[k] - tm[0]) / 1e3) for k in range(16, 30) if tm[k])
    for r in range(world):
        dist.barrier()
        if r == rank:
            if r == 0:
                print("%s: step (event, max over ranks) mean %.1f us, min %.1f us" % (mode, t[0].item(), t[1].item()))
            print("  rank %d in-kernel: %s" % (rank, line), flush=True)
torch.cuda.synchronize(); dist.barrier()
os._exit(0)
