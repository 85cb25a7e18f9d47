"""Run the 4-way evaluation kernel on 100 000 synthetic frames a few times (target for ncu captures).
usage: python tools/run_eval.py [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_baby_b200 as m
from bench import S_FIXED

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(5)
N, C, E = 25000, 22, 512
frames = torch.randn(N * 4, E, generator=g).to(dev); cats = torch.randn(C, E, generator=g).to(dev)
idx = torch.randint(0, C, (N,), generator=g).to(torch.int32).to(dev)
for _ in range(iters):
    pred, _ = m.ops.eval_nway(frames, cats, idx, 4, True, S_FIXED, False)
torch.cuda.synchronize()
print("pred[:8]", pred[:8].tolist())
