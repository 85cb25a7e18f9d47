"""In-situ cost of each phase of the single-GPU fused flat train step (B = 512 by default): the step is
captured as a CUDA graph that stops after phase k (CVCL_B200_STEP_PHASES, read by
cvcl_flat_contrastive_step), replayed with an L2 flush in between and event-timed.  Differences between
consecutive k show what each phase adds where it actually runs (PDL overlap, parallel graph branches).
    python tools/step_phases.py [B] [reps] [phase list, e.g. 11,1,0]"""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_baby_b200 as m
from bench import build_model, synth_batch, S_FIXED

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
dev = torch.device("cuda:0")
_, model = build_model(dev, None)
f, ids, lens = synth_batch(1234, B)
x = torch.from_numpy(f).to(dev).to(torch.bfloat16); ids = torch.from_numpy(ids).to(dev); lens = torch.from_numpy(lens).to(dev)
w, b = model.image_embed.model.fc.weight, model.image_embed.model.fc.bias
table = model.text_embed.embedding.weight
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
NAMES = {11: "1a head chain only (cast W, split-K GEMM, bias+norm) + memsets", 12: "1b text encoder only + memsets",
         1: "1 encoders (both branches)", 2: "2 + similarity / InfoNCE (+ merge)", 3: "3 + Gs",
         4: "4 + dI, dT", 0: "5 + dW, embedding scatter (full step)", -1: "empty graph (one memset node)"}
out = {}
KS = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [-1, 11, 12, 1, 2, 3, 4, 0]
for k in KS:
    os.environ["CVCL_B200_STEP_PHASES"] = str(k)
    def step():
        if k == -1:
            return flush[:256].zero_()
        return m.ops.flat_contrastive_step(x, ids, lens, w, b, table, S_FIXED, True, True, False)
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    for _ in range(5):
        flush.zero_(); g.replay()
    torch.cuda.synchronize()
    evs = []
    for _ in range(reps):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); evs.append((e0, e1))
    torch.cuda.synchronize()
    ts = [a.elapsed_time(c) * 1e3 for a, c in evs]
    out[NAMES[k]] = dict(mean_us=statistics.mean(ts), median_us=statistics.median(ts), min_us=min(ts))
    print("%-70s mean %7.1f  median %7.1f  min %7.1f us" % (NAMES[k], statistics.mean(ts), statistics.median(ts), min(ts)), flush=True)
os.environ.pop("CVCL_B200_STEP_PHASES", None)
print(json.dumps(dict(pairs=B, reps=reps, phases=out)))
