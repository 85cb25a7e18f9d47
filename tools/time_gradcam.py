"""Grad-CAM maps for 64 images incl. the 224x224 bicubic resize as one CUDA graph (L2 flushed between replays)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_baby_b200 as m
dev = torch.device("cuda:0"); g = torch.Generator().manual_seed(1)
N = 64
act = torch.relu(torch.randn(N, 2048, 7, 7, generator=g)).to(dev); W = (torch.randn(512, 2048, generator=g) / 45).to(dev)
b = torch.zeros(512, device=dev)
t = torch.nn.functional.normalize(torch.randn(N, 512, generator=g), dim=1).to(dev)
f = lambda: m.ops.bicubic_upsample(m.ops.gradcam_flat(act, W, b, t, True), 224, 224)
for _ in range(3):
    f()
torch.cuda.synchronize()
gph = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    f()
torch.cuda.synchronize()
with torch.cuda.graph(gph):
    out = f()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for _ in range(30):
    flush.zero_(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); gph.replay(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
print("gradcam N=64 (6 kernels, graph): mean %.1f us min %.1f us; activation %.1f MB" % (
    1e3 * sum(ts) / len(ts), 1e3 * min(ts), act.numel() * 4 / 1e6))
