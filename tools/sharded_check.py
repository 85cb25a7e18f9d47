"""torchrun target: sharded global-batch InfoNCE (exchange over NVLink: peer-memory kernels or NCCL) == the single-GPU
global-batch result, and the full sharded model step == single-GPU step on the concatenated batch.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/sharded_check.py [B_global]
Prints `SHARDED_OK ...` on rank 0 on success, raises otherwise."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
import multimodal_baby_b200 as m
from bench import build_model, synth_batch, step_api, S_FIXED, E

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
b = B // world
g = torch.Generator().manual_seed(11)
img = torch.nn.functional.normalize(torch.randn(B, E, generator=g), dim=1).to(dev)
txt = torch.nn.functional.normalize(torch.randn(B, E, generator=g), dim=1).to(dev)
sl = slice(rank * b, (rank + 1) * b)

# (1) feature-level op: sharded vs single-device
def run(i, t, group):
    i = i.clone().requires_grad_(True); t = t.clone().requires_grad_(True)
    s = torch.tensor(S_FIXED, device=dev, requires_grad=True)
    out = m.ops.sim_infonce(i, t, s, group)
    out[0].backward()
    return [o.detach() for o in out[:5]], i.grad, t.grad, s.grad, out[5]
ref5, rdi, rdt, rds, rarg = run(img, txt, None)
got5, gdi, gdt, gds, garg = run(img[sl], txt[sl], dist.group.WORLD)
for a, c in zip(ref5, got5):
    assert abs(a.item() - c.item()) <= 2e-6 * max(1.0, abs(a.item())), (a.item(), c.item())
def rel(a, c): return float((a - c).norm() / c.norm())
assert rel(gdi, rdi[sl]) <= 1e-5, rel(gdi, rdi[sl])
assert rel(gdt, rdt[sl]) <= 1e-5, rel(gdt, rdt[sl])
assert abs(gds.item() - rds.item()) <= 1e-4 * abs(rds.item()) + 1e-6, (gds.item(), rds.item())
assert torch.equal(garg, rarg[sl])

# (2) whole model step through the public API: sharded (512*world pairs) vs single GPU
_, model_s = build_model(dev, dist.group.WORLD)
_, model_1 = build_model(dev, None)
fs, ids, lens = zip(*[synth_batch(1234 + r, 512) for r in range(world)])
x_all = torch.from_numpy(np.concatenate(fs)).to(dev).to(torch.bfloat16)
ids_all = torch.from_numpy(np.concatenate(ids)).to(dev); lens_all = torch.from_numpy(np.concatenate(lens)).to(dev)
l1 = step_api(model_1, x_all, ids_all, lens_all, 1)
ls = step_api(model_s, x_all[rank * 512:(rank + 1) * 512], ids_all[rank * 512:(rank + 1) * 512],
              lens_all[rank * 512:(rank + 1) * 512], world)
assert abs(l1.item() - ls.item()) <= 2e-6 * abs(l1.item()), (l1.item(), ls.item())
for (n, p1), (_, ps) in zip(model_1.named_parameters(), model_s.named_parameters()):
    if p1.grad is None:
        continue
    r = rel(ps.grad, p1.grad)
    assert r <= 5e-3, (n, r)
# (2b) the gradient sum runs in a fixed rank order: every rank holds bit-identical gradients
sums = torch.stack([ps.grad.double().sum() for ps in model_s.parameters() if ps.grad is not None])
allsums = [torch.empty_like(sums) for _ in range(world)]
dist.all_gather(allsums, sums)
for q in range(world):
    assert torch.equal(allsums[q], allsums[0]), ("gradients differ between ranks", q)
# (2c) forward-only sharded step (only the five scalars are summed over the ranks), then a training step again: the
# two forms alternate on the same exchange buffers
fw = model_s.image_embed.model.fc
xs, is_, ls_ = (x_all[rank * 512:(rank + 1) * 512], ids_all[rank * 512:(rank + 1) * 512], lens_all[rank * 512:(rank + 1) * 512])
st5, _, _ = m.ops.flat_step_sharded(xs, is_, ls_, fw.weight, fw.bias, model_s.text_embed.embedding.weight, S_FIXED,
                                    True, False, False, dist.group.WORLD)
assert abs(st5[0].item() - l1.item()) <= 2e-6 * abs(l1.item()), (st5[0].item(), l1.item())
ls2 = step_api(model_s, xs, is_, ls_, world)
assert abs(ls2.item() - ls.item()) <= 2e-6 * abs(ls.item()), (ls2.item(), ls.item())
for (n, p1), (_, ps) in zip(model_1.named_parameters(), model_s.named_parameters()):
    if p1.grad is not None:
        assert rel(ps.grad, p1.grad) <= 5e-3, (n, rel(ps.grad, p1.grad))
# (3) op-by-op sharded path (features -> sim_infonce with the group) + explicit gradient all-reduce
_, model_o = build_model(dev, dist.group.WORLD)
model_o.train_path = "ops"
for p in model_o.parameters():
    p.grad = None
lo = model_o.calculate_contrastive_loss(x_all[rank * 512:(rank + 1) * 512], ids_all[rank * 512:(rank + 1) * 512],
                                        lens_all[rank * 512:(rank + 1) * 512])[0]
lo.backward()
m.sharding.allreduce_gradients(model_o.parameters(), dist.group.WORLD)
assert abs(l1.item() - lo.item()) <= 2e-6 * abs(l1.item()), (l1.item(), lo.item())
for (n, p1), (_, po) in zip(model_1.named_parameters(), model_o.named_parameters()):
    if p1.grad is None:
        continue
    assert rel(po.grad, p1.grad) <= 5e-3, (n, rel(po.grad, p1.grad))
# which exchange back end ran: peer-memory kernels by default, NCCL collectives with CVCL_B200_SYMM=0
pxs = list(m.sharding.PeerExchange._cache.values())
peer = bool(pxs) and all(v is not None for v in pxs)
if os.environ.get("CVCL_B200_SYMM", "1") != "0":
    assert peer, "peer-memory exchange expected but not active (symmetric memory unavailable?)"
    for v in pxs:
        v.check()                      # no cross-rank barrier timed out
else:
    assert not pxs
# (4) repeated steps reuse the persistent symmetric blocks: same loss every time (split-K atomics: to rounding)
for _ in range(3):
    l_again = step_api(model_s, x_all[rank * 512:(rank + 1) * 512], ids_all[rank * 512:(rank + 1) * 512],
                       lens_all[rank * 512:(rank + 1) * 512], world)
    assert abs(l_again.item() - ls.item()) <= 2e-6 * abs(ls.item()), (l_again.item(), ls.item())
torch.cuda.synchronize(); dist.barrier()
if rank == 0:
    print("SHARDED_OK world=%d B=%d loss=%.6f model_loss=%.6f exchange=%s" % (
        world, B, got5[0].item(), ls.item(), "peer" if peer else "nccl"))
dist.destroy_process_group()
