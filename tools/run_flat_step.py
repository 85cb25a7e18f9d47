"""Run the fused flat contrastive step eagerly a few times (target for ncu captures).
usage: python tools/run_flat_step.py [B] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import multimodal_baby_b200 as m
from bench import synth_batch, synth_weights, S_FIXED

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
f, ids, lens = synth_batch(1234, B)
W, b, table = synth_weights()
x = torch.from_numpy(f).to(dev).to(torch.bfloat16)
ids = torch.from_numpy(ids).to(dev); lens = torch.from_numpy(lens).to(dev)
W = torch.from_numpy(W).to(dev); b = torch.from_numpy(b).to(dev); table = torch.from_numpy(table).to(dev)
for _ in range(iters):
    out = m.ops.flat_contrastive_step(x, ids, lens, W, b, table, S_FIXED, True, True, False)
torch.cuda.synchronize()
print("loss", out[0][0].item())
