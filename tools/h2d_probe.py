"""Which event makes H2D copies from one pinned staging buffer slow (176 us instead of 46 us for 2.2 MB, seen in the
lagged mode of GraphedContrastiveStep on some boxes)?  Measurement helper (gpurun)."""
import os, statistics, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
N = 2203648
dst = torch.empty(N, dtype=torch.uint8, device=dev)


def h2d(buf, label):
    ts = []
    for _ in range(30):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); dst.copy_(buf, non_blocking=True); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print("%-70s %.1f us  (ptr %x)" % (label, statistics.median(ts), buf.data_ptr()), flush=True)


A = torch.zeros(N, dtype=torch.uint8).pin_memory()
h2d(A, "A = zeros().pin_memory(), fresh")
A.fill_(3); h2d(A, "A after CPU fill_")
m, model = bench.build_model(dev)
h2d(A, "A after model build")
B = torch.empty(N, dtype=torch.uint8).pin_memory(); B.copy_(A)
h2d(A, "A after CPU read (B.copy_(A))"); h2d(B, "B (fresh, CPU-written)")
# views / set_ based byte view
bv = torch.empty(0, dtype=torch.uint8).set_(A.untyped_storage(), 0, (N,))
h2d(bv, "byte view of A via set_()")
# graph capture of a copy from A
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    dst.copy_(A, non_blocking=True)
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
with torch.cuda.graph(g):
    dst.copy_(A, non_blocking=True)
for _ in range(20):
    g.replay()
torch.cuda.synchronize()
h2d(A, "A after being the source of a captured memcpy node (graph alive)")
ts = []
for _ in range(30):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
print("%-70s %.1f us" % ("   replay of that graph (memcpy node)", statistics.median(ts)))
del g
torch.cuda.synchronize()
h2d(A, "A after the graph was destroyed")
# the real thing
xh, ih, lh = m.staging.packed_buffers([((512, bench.K), torch.bfloat16), ((512, bench.L), torch.int64), ((512,), torch.int64)])
f, ids, lens = bench.synth_batch(1, 512)
xh.copy_(torch.from_numpy(f)); ih.copy_(torch.from_numpy(ids)); lh.copy_(torch.from_numpy(lens))
C = torch.empty(0, dtype=torch.uint8).set_(xh.untyped_storage(), 0, (N,))
h2d(C, "C = packed staging arena, fresh")
for kw in (dict(prefetch=True, lagged_loss=False), dict(prefetch=True, lagged_loss=True)):
    st = m.GraphedContrastiveStep(model, xh, ih, lh, **kw)
    st.prime()
    for _ in range(50):
        st()
    st.flush(); torch.cuda.synchronize()
    h2d(C, "C after GraphedContrastiveStep(%s) ran 50 calls" % kw)
    for k, ha in enumerate(st._host_arenas):
        h2d(ha, "   its staging set %d" % k)
    del st
    torch.cuda.synchronize()
    h2d(C, "C after that step object was deleted")

# write-combined staging memory from the library
print("---- write-combined")
W = m.staging.host_arena(N, True)
print("is_pinned:", W.is_pinned())
W.fill_(5); h2d(W, "WC arena after fill_")
tmp = W.clone(); h2d(W, "WC arena after a CPU read (clone)")
W[: N // 2].copy_(B[: N // 2]); h2d(W, "WC arena after half overwritten from B")
small = torch.ones(4096, dtype=torch.uint8)
t0 = time.perf_counter()
for i in range(0, N - 4096, 4096):
    W[i:i + 4096].copy_(small)
t1 = time.perf_counter()
h2d(W, "WC arena after 4 KB-chunk writes (%.0f us of CPU)" % ((t1 - t0) * 1e6))
P = torch.empty(N, dtype=torch.uint8).pin_memory()
for i in range(0, N - 4096, 4096):
    P[i:i + 4096].copy_(small)
h2d(P, "ordinary pinned after 4 KB-chunk writes (cached, dirty)")
P.copy_(B); h2d(P, "ordinary pinned after one big copy_ (memcpy)")
q = P.sum().item(); h2d(P, "ordinary pinned after CPU read (sum)")
P.fill_(1); h2d(P, "ordinary pinned after fill_")
NP = m.staging.host_arena(N, False)
NP.fill_(2); h2d(NP, "library pinned (not WC) after fill_")
q = NP.sum().item(); h2d(NP, "library pinned (not WC) after CPU read")
t0 = time.perf_counter(); W.copy_(B); t1 = time.perf_counter(); P.copy_(B); t2 = time.perf_counter()
print("CPU copy of 2.2 MB into WC: %.0f us, into ordinary pinned: %.0f us" % ((t1 - t0) * 1e6, (t2 - t1) * 1e6))
