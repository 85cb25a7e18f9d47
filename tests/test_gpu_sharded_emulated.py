"""-m gpu, ONE device: the sharded global-batch loss (BASELINE config 3) checked against the oracle by
running every rank's share on the same GPU.  Rank r calls the same C-ABI entry points a real rank calls
(`cvcl_sim_infonce_fwd` / `_bwd_g` / `cvcl_feat_grad_norm_bwd` through ops.sim_infonce_fwd / sim_infonce_bwd) on the
gathered features with diag_off = r*b; loss parts, dI, dT and ds are compared with
`oracle.sharded_contrastive_loss` and the single-process autograd / fp64 closed form.  The multi-GPU runs of the
same code are in tests/test_gpu_sharded.py (skipped on a 1-GPU box)."""
import math

import numpy as np
import pytest
import torch

from _util import O, S_DEFAULT, assert_grad_close, t

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def cv():
    import multimodal_baby_b200 as m
    m._cabi.load()
    return m


def _features(seed, B, E=512):
    gen = torch.Generator().manual_seed(seed)
    img = torch.nn.functional.normalize(torch.randn(B, E, generator=gen), dim=1)
    txt = torch.nn.functional.normalize(torch.randn(B, E, generator=gen) + 0.5 * img, dim=1)   # correlated pairs
    return img, txt


def _reference(img, txt, s):
    """single-process global-batch loss + feature gradients (the reference's own definition, multimodal.py:796-818)
    on the bf16-rounded operands, fp32 autograd."""
    i = img.to(torch.bfloat16).float().requires_grad_(True)
    tt = txt.to(torch.bfloat16).float().requires_grad_(True)
    sv = torch.tensor(s, requires_grad=True)
    lpi, lpt = O.logits_from_match(O.similarity_flat(i, tt), sv)
    res = O.infonce(lpi, lpt)
    res.loss.backward()
    return res, i.grad, tt.grad, sv.grad.item()


@pytest.mark.parametrize("world,b", [(2, 256), (4, 128), (8, 64), (8, 512)])
def test_emulated_shards_match_oracle(cv, world, b):
    B = world * b
    img, txt = _features(100 + world, B)
    s = S_DEFAULT
    ref, dI_ref, dT_ref, ds_ref = _reference(img, txt, s)
    parts = O.sharded_contrastive_loss(img.to(torch.bfloat16).float(), txt.to(torch.bfloat16).float(), s, world)
    assert abs(float(sum(parts)) - float(ref.loss)) <= 1e-5 * float(ref.loss)      # oracle identity
    i16 = img.to(DEV).to(torch.bfloat16); t16 = txt.to(DEV).to(torch.bfloat16)
    fwd = cv.ops._raw(cv.ops.sim_infonce_fwd); bwd = cv.ops._raw(cv.ops.sim_infonce_bwd)
    outs = []
    lse0 = torch.empty(B, device=DEV); lse1 = torch.empty(B, device=DEV)
    for r in range(world):
        sl = slice(r * b, (r + 1) * b)
        out5, l0, l1, a0, a1 = fwd(i16[sl].contiguous(), t16, t16[sl].contiguous(), i16, s, r * b, 1.0 / B)
        outs.append(out5.clone()); lse0[sl] = l0; lse1[sl] = l1
        # this rank's share of the loss: its row block and column block (both halves carry the factor 1/2)
        assert abs(float(out5[0]) - float(parts[r])) <= 1e-3 * abs(float(parts[r])), (r, float(out5[0]), float(parts[r]))
    tot = torch.stack(outs).sum(0)
    assert abs(float(tot[0]) - float(ref.loss)) <= 1e-3 * float(ref.loss)
    assert abs(float(tot[3]) - float(ref.image_entropy)) <= 2e-3 and abs(float(tot[4]) - float(ref.text_entropy)) <= 2e-3
    assert abs(float(tot[1]) - float(ref.image_accuracy)) <= 0.02 and abs(float(tot[2]) - float(ref.text_accuracy)) <= 0.02
    # LSEs of every rank's rows against fp64 on the bf16 operands
    S64 = math.exp(s) * (i16.double() @ t16.double().t())
    assert float((lse0.double() - torch.logsumexp(S64, 1)).abs().max()) <= 2e-3
    assert float((lse1.double() - torch.logsumexp(S64, 0)).abs().max()) <= 2e-3
    # backward: each rank's dI / dT for its local pairs, ds summed over ranks (the all-reduce of the real run)
    dI = torch.empty(B, img.shape[1], device=DEV); dT = torch.empty_like(dI); ds = 0.0
    for r in range(world):
        sl = slice(r * b, (r + 1) * b)
        di, dt, dsr = bwd(i16[sl].contiguous(), t16, t16[sl].contiguous(), i16, s, r * b, 0.5 / B,
                          lse0[sl].contiguous(), lse1, lse1[sl].contiguous(), lse0)
        dI[sl] = di; dT[sl] = dt; ds += float(dsr[0])
    assert_grad_close(dI.cpu().numpy(), dI_ref.numpy(), "dI")
    assert_grad_close(dT.cpu().numpy(), dT_ref.numpy(), "dT")
    assert abs(ds - ds_ref) <= 2e-2 * abs(ds_ref) + 1e-3


def test_config3_size_forward_and_backward(cv):
    """BASELINE config 3 at its stated size on one GPU: B = 32768 forward (loss / LSE against fp64 on a 256-row
    subset of both directions) and B = 8192 gradients against the fp64 closed form (G = (P_row + P_col - 2I)/2B)."""
    B, E = 32768, 512
    img, txt = _features(33, B)
    i16 = img.to(DEV).to(torch.bfloat16); t16 = txt.to(DEV).to(torch.bfloat16)
    s = S_DEFAULT
    out5, l0, l1, a0, a1 = cv.ops._raw(cv.ops.sim_infonce_fwd)(i16, t16, t16, i16, s, 0, 1.0 / B)
    rows = torch.arange(0, B, B // 256, device=DEV)
    S_r = math.exp(s) * (i16[rows].double() @ t16.double().t())            # [256, B]
    S_c = math.exp(s) * (t16[rows].double() @ i16.double().t())
    assert float((l0[rows].double() - torch.logsumexp(S_r, 1)).abs().max()) <= 2e-3
    assert float((l1[rows].double() - torch.logsumexp(S_c, 1)).abs().max()) <= 2e-3
    assert (a0[rows].long() == S_r.argmax(1)).float().mean().item() >= 0.98
    # the loss from the kernel's own LSEs and fp64 positives == the kernel's loss
    diag = math.exp(s) * (i16.double() * t16.double()).sum(1)
    loss64 = 0.5 * ((l0.double() - diag).mean() + (l1.double() - diag).mean())
    assert abs(float(out5[0]) - float(loss64)) <= 1e-4 * float(loss64)
    assert float(out5[0]) < math.log(B)                                    # correlated pairs: below chance level
    del S_r, S_c
    # gradients at 8192 against the closed form in fp64 (on the device: 8192^2 doubles = 0.5 GB)
    B2 = 8192
    i2 = i16[:B2].contiguous(); t2 = t16[:B2].contiguous()
    out5, l0, l1, _, _ = cv.ops._raw(cv.ops.sim_infonce_fwd)(i2, t2, t2, i2, s, 0, 1.0 / B2)
    dI, dT, ds = cv.ops._raw(cv.ops.sim_infonce_bwd)(i2, t2, t2, i2, s, 0, 0.5 / B2, l0, l1, l1, l0)
    I = i2.double(); T = t2.double()
    S = math.exp(s) * (I @ T.t())
    G = (torch.softmax(S, 1) + torch.softmax(S, 0)) / (2 * B2)
    G.diagonal().sub_(1.0 / B2)
    dI64 = math.exp(s) * (G @ T); dT64 = math.exp(s) * (G.t() @ I); ds64 = float((G * S).sum())
    assert_grad_close(dI.cpu().numpy(), dI64.cpu().numpy(), "dI@8192")
    assert_grad_close(dT.cpu().numpy(), dT64.cpu().numpy(), "dT@8192")
    assert abs(float(ds[0]) - ds64) <= 2e-2 * abs(ds64) + 1e-3


def test_flat_step_zipf_ids(cv):
    """vocabulary ids are frequency ranked, so real batches are Zipf distributed: a few rows of the
    embedding table receive most of the gradient.  The one-kernel step (token-count GEMM) and the oracle agree."""
    from _util import case_inputs, oracle_flat_step
    B, E, V, L = 512, 512, 2350, 25
    inp = case_inputs(2024, B, E, "flat")
    rng = np.random.RandomState(7)
    lens = rng.randint(3, L + 1, size=B).astype(np.int64)
    ranks = np.arange(4, V)
    pz = 1.0 / (ranks - 3.0); pz /= pz.sum()                       # Zipf(1.0) over the non-special ids
    ids = np.zeros((B, L), np.int64)
    for r in range(B):
        ids[r, 0] = 2; ids[r, lens[r] - 1] = 3
        if lens[r] > 2:
            ids[r, 1:lens[r] - 1] = rng.choice(ranks, size=lens[r] - 2, p=pz)
    inp["ids"], inp["lens"] = ids, lens
    ref = oracle_flat_step(inp)
    d = {k: t(v, DEV) for k, v in inp.items()}
    W = d["W"].requires_grad_(True); b = d["b"].requires_grad_(True); table = d["table"].requires_grad_(True)
    s = torch.tensor(S_DEFAULT, device=DEV, requires_grad=True)
    out = cv.ops.flat_contrastive_loss(d["f"], d["ids"], d["lens"], W, b, table, s, True)
    out[0].backward()
    assert abs(out[0].item() - ref["loss"].item()) <= 1e-3 * abs(ref["loss"].item())
    assert_grad_close(table.grad.cpu().numpy(), ref["dtable"].numpy(), "dtable (Zipf)")
    assert_grad_close(W.grad.cpu().numpy(), ref["dW"].numpy(), "dW")
    hot = np.argsort(-np.abs(ref["dtable"].numpy()).sum(1))[:8]    # the heavy rows individually
    for v in hot:
        assert_grad_close(table.grad[v].cpu().numpy(), ref["dtable"][v].numpy(), "dtable[%d]" % v, cos_min=0.999, rel_max=2e-2)
    assert not table.grad[0].any()
