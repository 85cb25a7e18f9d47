"""Is the slow H2D state tied to a buffer or to time / driver state?  Measurement helper (gpurun)."""
import os, statistics, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
N = 2203648
dst = torch.empty(N, dtype=torch.uint8, device=dev)
dst2 = torch.empty(N, dtype=torch.uint8, device=dev)
bufs = {}
for name in "ABCD":
    t = torch.empty(N, dtype=torch.uint8).pin_memory(); t.fill_(ord(name)); bufs[name] = t
m, model = bench.build_model(dev)


def med(buf, d=dst, n=12):
    ts = []
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); d.copy_(buf, non_blocking=True); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


def row(label):
    print("%-58s " % label + " ".join("%s:%5.0f" % (k, med(v)) for k, v in bufs.items()), flush=True)


row("baseline 0"); row("baseline 1")
time.sleep(1.0); row("after sleep 1 s")
E = torch.empty(N, dtype=torch.uint8).pin_memory(); E.fill_(9); bufs["E"] = E; row("after new pinned alloc E")
va = torch.empty(0, dtype=torch.uint8).set_(bufs["A"].untyped_storage(), 0, (N,))
print("   copy from set_() view of A: %.0f" % med(va)); row("after copying from a set_() view of A")
print("   copy A -> other device buffer: %.0f" % med(bufs["A"], dst2))
vb = bufs["B"][:N]
print("   copy from slice view of B: %.0f" % med(vb)); row("after copying from a slice view of B")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    print("   copy C on a side stream: %.0f" % med(bufs["C"]))
row("after copying C on a side stream")
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    dst.copy_(bufs["D"], non_blocking=True)
for _ in range(20):
    g.replay()
torch.cuda.synchronize(); row("after capturing + replaying a memcpy node from D")
del g; torch.cuda.synchronize(); row("after deleting that graph")
x = torch.randn(4096, 4096, device=dev)
for _ in range(50):
    y = x @ x
torch.cuda.synchronize(); row("after 50 big matmuls")
row("baseline again")
