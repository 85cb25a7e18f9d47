"""BASELINE.json config 3 under torchrun: global-batch contrastive loss, B pairs sharded over N GPUs,
NCCL feature all-gather over NVLink, fwd + bwd at feature level (img_r, txt_r bf16 [b,512] per rank).
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_global_batch.py [B]
Device-timed with CUDA events, max over ranks; prints one JSON line on rank 0."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
group = None
if world > 1:
    dist.init_process_group("nccl", device_id=dev); group = dist.group.WORLD
import multimodal_baby_b200 as m
from bench import S_FIXED, load_peaks
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
E = 512; b = B // world
g = torch.Generator().manual_seed(100 + rank)
img = torch.nn.functional.normalize(torch.randn(b, E, generator=g), dim=1).to(dev).to(torch.bfloat16)
txt = torch.nn.functional.normalize(torch.randn(b, E, generator=g), dim=1).to(dev).to(torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def step():
    i = img.detach().requires_grad_(True); t = txt.detach().requires_grad_(True)
    out = m.ops.sim_infonce(i, t, S_FIXED, group)
    out[0].backward()
    return out[0]
for _ in range(3): step()
torch.cuda.synchronize()
if world > 1: dist.barrier()
ts = []
for _ in range(8):
    flush.zero_()
    if world > 1: dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); loss = step(); e1.record(); e1.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = torch.tensor([float(np.mean(ts))], device=dev)
if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
ms = float(ms.item())
if rank == 0:
    peaks = load_peaks()
    flops_rank = (12 if world > 1 else 8) * b * B * E
    print(json.dumps({"config": "global-batch contrastive loss fwd+bwd, %d pairs over %d GPU(s)" % (B, world),
                      "n_gpus": world, "ms_per_step": ms, "pairs_per_s": B / (ms * 1e-3), "loss": float(loss.item()),
                      "tflop_s_per_gpu": flops_rank / (ms * 1e-3) / 1e12,
                      "tensor_frac_per_gpu": flops_rank / (ms * 1e-3) / 1e12 / peaks["bf16_tflops"]}))
sys.stdout.flush()
if world > 1:
    torch.cuda.synchronize(); dist.barrier(); os._exit(0)
