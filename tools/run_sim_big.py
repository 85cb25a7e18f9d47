"""ncu target: similarity + InfoNCE forward/backward at a large global batch (default 8192)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_baby_b200 as m
from bench import S_FIXED
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
img = torch.nn.functional.normalize(torch.randn(B, 512, generator=g), dim=1).to(dev).to(torch.bfloat16)
txt = torch.nn.functional.normalize(torch.randn(B, 512, generator=g), dim=1).to(dev).to(torch.bfloat16)
for _ in range(2):
    i = img.detach().requires_grad_(True); t = txt.detach().requires_grad_(True)
    out = m.ops.sim_infonce(i, t, S_FIXED)
    out[0].backward()
torch.cuda.synchronize()
print("loss", out[0].item())
