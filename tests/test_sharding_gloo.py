"""world_size-2/4 `gloo` tests (CPU) of the sharded global-batch InfoNCE orchestration
(multimodal-baby_b200/sharding.py): collectives, diagonal offsets and gradient reductions, with
the arithmetic injected as a torch restatement (the CUDA ops replace it on the GPU).  The result
must equal the single-process reference computation on the concatenated batch."""
import math
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
S = float(-np.log(0.07))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def cpu_compute_fwd(img_q, txt_k, txt_q, img_k, ls, diag_off, inv_rows):
    """torch restatement of cvcl_sim_infonce_fwd (same outputs, same scaling)."""
    sc = math.exp(ls)
    outs = []
    for q, k in ((img_q, txt_k), (txt_q, img_k)):
        x = sc * q.double() @ k.double().T
        lse = torch.logsumexp(x, 1)
        gt = torch.arange(q.shape[0]) + diag_off
        ce = (lse - x[torch.arange(q.shape[0]), gt]).sum()
        ent = (lse - (torch.softmax(x, 1) * x).sum(1)).sum()
        arg = x.argmax(1)
        outs.append((lse, ce, ent, (arg == gt).sum().double(), arg))
    out5 = torch.zeros(8, dtype=torch.float64)
    out5[0] = (outs[0][1] + outs[1][1]) * 0.5 * inv_rows
    out5[1] = outs[0][3] * inv_rows; out5[2] = outs[1][3] * inv_rows
    out5[3] = outs[0][2] * inv_rows; out5[4] = outs[1][2] * inv_rows
    return out5, outs[0][0], outs[1][0], outs[0][4].int(), outs[1][4].int()


def cpu_compute_bwd(img_q, txt_k, txt_q, img_k, ls, diag_off, coef, lse_q0, lse_k0, lse_q1, lse_k1):
    sc = math.exp(ls)
    res = []
    ds = torch.zeros(1, dtype=torch.float64)
    for z, (q, k, lq, lk) in enumerate(((img_q, txt_k, lse_q0, lse_k0), (txt_q, img_k, lse_q1, lse_k1))):
        x = sc * q.double() @ k.double().T
        G = torch.exp(x - lq[:, None]) + torch.exp(x - lk[None, :])
        idx = torch.arange(q.shape[0])
        G[idx, idx + diag_off] -= 2
        G = G * coef
        res.append(sc * G @ k.double())
        if z == 0:
            ds += (G * x).sum()
    return res[0], res[1], ds


def _worker(rank, world, port, B, E, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from multimodal_baby_b200 import sharding
    g = torch.Generator().manual_seed(7)
    img = torch.nn.functional.normalize(torch.randn(B, E, generator=g, dtype=torch.float64), dim=1)
    txt = torch.nn.functional.normalize(torch.randn(B, E, generator=g, dtype=torch.float64), dim=1)
    b = B // world
    sl = slice(rank * b, (rank + 1) * b)
    out5, saved, (a0, a1) = sharding.infonce_forward(img[sl].contiguous(), txt[sl].contiguous(), S,
                                                     dist.group.WORLD, cpu_compute_fwd)
    dimg, dtxt, ds = sharding.infonce_backward(saved, dist.group.WORLD, cpu_compute_bwd)
    # replicated-parameter gradient reduction: SUM over ranks
    p = torch.nn.Parameter(torch.zeros(3, dtype=torch.float64))
    p.grad = torch.full((3,), float(rank + 1), dtype=torch.float64)
    sharding.allreduce_gradients([p], dist.group.WORLD)
    q.put((rank, out5.numpy(), dimg.numpy(), dtxt.numpy(), ds.numpy(), a0.numpy(), p.grad.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_infonce_equals_single_process(world):
    B, E = 24, 32
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, E, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference on the concatenated batch (autograd, fp64)
    g = torch.Generator().manual_seed(7)
    img = torch.nn.functional.normalize(torch.randn(B, E, generator=g, dtype=torch.float64), dim=1)
    txt = torch.nn.functional.normalize(torch.randn(B, E, generator=g, dtype=torch.float64), dim=1)
    img.requires_grad_(True); txt.requires_grad_(True)
    s = torch.tensor(S, dtype=torch.float64, requires_grad=True)
    sys.path.insert(0, ROOT)
    from oracle import cvcl_oracle as O
    lpi, lpt = O.logits_from_match(O.similarity_flat(img, txt), s)
    ref = O.infonce(lpi, lpt)
    ref.loss.backward()
    b = B // world
    for rank, out5, dimg, dtxt, ds, a0, pg in results:
        assert abs(out5[0] - ref.loss.item()) < 1e-12
        assert abs(out5[1] - ref.image_accuracy.item()) < 1e-7   # reference accuracy is fp32
        assert abs(out5[3] - ref.image_entropy.item()) < 1e-12
        assert abs(out5[4] - ref.text_entropy.item()) < 1e-12
        sl = slice(rank * b, (rank + 1) * b)
        np.testing.assert_allclose(dimg, img.grad[sl].numpy(), atol=1e-13)
        np.testing.assert_allclose(dtxt, txt.grad[sl].numpy(), atol=1e-13)
        assert abs(ds[0] - s.grad.item()) < 1e-12
        assert np.array_equal(a0, ref.image_pred[sl].numpy().astype(np.int32))
        assert np.allclose(pg, world * (world + 1) / 2)
