"""torchrun target: in-situ cost of each phase of the sharded flat train step (512 pairs per GPU).
For k = 1..7 the step is captured as a CUDA graph that stops after phase k (ops.flat_step_sharded
`phase_limit`), replayed with an L2 flush in between and event-timed (max over ranks); the difference
between consecutive k is what that phase adds to the step where it actually runs (with PDL overlap,
side-stream branches and the other ranks' skew), which isolated per-collective timings do not show.
Three exchange back ends are measured: NCCL collectives, the pull and the push peer-memory kernels.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/sharded_phases.py [reps] [nccl,pull,push]
"""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
import multimodal_baby_b200 as m
from bench import build_model, synth_batch, S_FIXED

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
B = 512
PHASES = ["1 local encoders", "2 + feature exchange", "3 + similarity/InfoNCE fwd", "4 + LSE exchange", "5 + Gs",
          "6 + dI,dT,dW,scatter", "7 + gradient all-reduce (full step)"]
_, model = build_model(dev, dist.group.WORLD)
f, ids, lens = synth_batch(1234 + rank, B)
x = torch.from_numpy(f).to(dev).to(torch.bfloat16); ids = torch.from_numpy(ids).to(dev); lens = torch.from_numpy(lens).to(dev)
w, b = model.image_embed.model.fc.weight, model.image_embed.model.fc.bias
table = model.text_embed.embedding.weight
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = {}
MODES = sys.argv[2].split(",") if len(sys.argv) > 2 else ["nccl", "pull", "push"]
for mode in MODES:
    os.environ["CVCL_B200_SYMM"] = "0" if mode == "nccl" else "1"
    os.environ["CVCL_B200_PEER_MODE"] = "pull" if mode == "pull" else "push"
    res = []
    for k in range(1, 8):
        lim = None if k == 7 else k
        def step():
            return m.ops.flat_step_sharded(x, ids, lens, w, b, table, S_FIXED, True, True, False, dist.group.WORLD,
                                           phase_limit=lim)
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(); dist.barrier()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step()
        for _ in range(5):
            flush.zero_(); g.replay()
        torch.cuda.synchronize(); dist.barrier()
        evs = []
        for _ in range(reps):
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); evs.append((e0, e1))
        torch.cuda.synchronize()
        t = torch.tensor([statistics.mean(a.elapsed_time(c) for a, c in evs)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res.append(float(t.item()) * 1e3)
        if rank == 0:
            print("%s  %-38s %8.1f us  (+%.1f)" % (mode, PHASES[k - 1], res[-1], res[-1] - (res[-2] if k > 1 else 0)), flush=True)
        del g
    used = [v is not None for v in m.sharding.PeerExchange._cache.values()]
    out[mode] = dict(cumulative_us=res, peer_exchange_active=bool(used and all(used)) if mode != "nccl" else False)
for px in m.sharding.PeerExchange._cache.values():
    if px is not None:
        px.check()
torch.cuda.synchronize(); dist.barrier()
if rank == 0:
    print(json.dumps(dict(world=world, pairs_per_gpu=B, phases=PHASES, **out)))
sys.stdout.flush()
os._exit(0)
