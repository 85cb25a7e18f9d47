"""-m gpu: spatial-embedding path (7x7 layer4 map, per-token text features) with "mean" and "max"
text-to-location similarity (multimodal.py:757-780), forward + backward, against the reference
golden vectors and the CPU oracle.  The model is driven through the drop-in MultiModalModel API."""
import argparse

import numpy as np
import pytest
import torch

from _util import O, S_DEFAULT, assert_grad_close, assert_logits_close, case_inputs, golden, t, rel_fro

pytestmark = pytest.mark.gpu
DEV = "cuda"


def build(cv, E, sim, inp, fix_temperature=False):
    args = argparse.Namespace(embedding_type="spatial", embedding_dim=E, normalize_features=True,
                              fix_temperature=fix_temperature, temperature=0.07, text_encoder="embedding",
                              sim=sim)
    vocab = {str(i): i for i in range(2350)}
    m = cv.MultiModalModel(cv.VisionEncoder(args, trunk="pooled"), cv.TextEncoder(vocab, 2048, args), args)
    with torch.no_grad():
        conv = m.image_embed.model[-1]
        conv.weight.copy_(t(inp["W"])[:, :, None, None]); conv.bias.copy_(t(inp["b"]))
        m.text_embed.embedding.weight.copy_(t(inp["table"]))
    return m.to(DEV).train()


@pytest.fixture(scope="module")
def cv():
    import multimodal_baby_b200 as m
    m._cabi.load()
    return m


def run_model(cv, E, sim, inp, precision=None):
    m = build(cv, E, sim, inp)
    if precision is not None:
        m.spatial_max_precision = precision
    out = m.calculate_contrastive_loss(t(inp["f"], DEV), t(inp["ids"], DEV), t(inp["lens"], DEV))
    out[0].backward()
    conv = m.image_embed.model[-1]
    return out, dict(dW=conv.weight.grad.reshape(E, -1).cpu().numpy(), db=conv.bias.grad.cpu().numpy(),
                     dtable=m.text_embed.embedding.weight.grad.cpu().numpy(),
                     ds=m.logit_neg_log_temperature.grad.item())


@pytest.mark.parametrize("name,sim", [("spatial_mean_e64_b6", "mean"), ("spatial_mean_e512_b12", "mean"),
                                      ("spatial_max_e64_b6", "max"), ("spatial_max_e512_b12", "max")])
def test_spatial_vs_reference_golden(cv, name, sim):
    g = golden(name)
    E, B = int(g["E"]), int(g["B"])
    inp = case_inputs(int(g["seed"]), B, E, "spatial")
    out, gr = run_model(cv, E, sim, inp)
    assert abs(out[0].item() - float(g["loss"])) <= 1e-3 * abs(float(g["loss"]))
    assert_logits_close(out[5].cpu().numpy(), g["logits_per_image"])
    assert_logits_close(out[6].cpu().numpy(), g["logits_per_text"])
    assert out[7].shape == (B, E, 7, 7) and out[8].shape == (B, 2048, 7, 7)
    assert float((out[7][:2].detach().cpu() - t(g["image_features_head"])).abs().max()) <= 6e-3
    # "max": head and scores run with two-term bf16 operands (model.spatial_max_precision = "split_bf16"), so the
    # arg-max locations agree with the reference's fp32 arithmetic and the north_star gradient gate applies
    # (cosine >= 0.999, rel-Frobenius <= 2e-2); single-term bf16 flipped near-tied locations (round 1: 0.9 / 0.9)
    cm, rm = (0.998, 5e-2) if sim == "mean" else (0.999, 2e-2)
    assert_grad_close(gr["db"], g["db"], "db", cos_min=cm, rel_max=rm)
    assert abs(gr["ds"] - float(g["ds"])) <= rm * abs(float(g["ds"])) + 2e-3
    assert rel_fro(gr["dW"][:8, :64], g["dW_slice"]) <= rm
    assert abs(np.linalg.norm(gr["dW"]) - float(g["dW_norm"])) <= rm * float(g["dW_norm"])
    assert rel_fro(gr["dtable"][:8], g["dtable_rows"]) <= rm
    assert abs(np.linalg.norm(gr["dtable"]) - float(g["dtable_norm"])) <= rm * float(g["dtable_norm"])
    assert not gr["dtable"][0].any()
    if "dW" in g:
        assert_grad_close(gr["dW"], g["dW"], "dW", cos_min=cm, rel_max=rm)


@pytest.mark.parametrize("sim,B,fr", [("mean", 40, None), ("max", 40, None), ("max", 37, None),
                                      ("max", 24, "bf16"), ("max", 6, "bf16")])
def test_spatial_vs_oracle(cv, sim, B, fr):
    """fr="bf16": the oracle rounds the encoded features to bf16 (the kernels' operand precision),
    which removes most arg-max flips, so the tight gradient gate applies to the max path too."""
    E = 512
    inp = case_inputs(900 + B, B, E, "spatial")
    ref = O.contrastive_step(t(inp["f"]), t(inp["ids"]), t(inp["lens"]), t(inp["W"]), t(inp["b"]),
                             t(inp["table"]), S_DEFAULT, "spatial", sim, feature_round=fr)
    # fr="bf16" compares against an oracle whose features were rounded to bf16: that is the fast single-term mode
    out, gr = run_model(cv, E, sim, inp, precision="bf16" if fr else None)
    assert abs(out[0].item() - ref["loss"].item()) <= 1e-3 * abs(ref["loss"].item())
    assert_logits_close(out[5].cpu().numpy(), ref["logits_per_image"].numpy())
    cm, rm = (0.998, 5e-2) if sim == "mean" else ((0.997, 8e-2) if fr else (0.999, 2e-2))
    assert_grad_close(gr["dW"], ref["dW"].reshape(E, -1).numpy(), "dW", cos_min=cm, rel_max=rm)
    assert_grad_close(gr["dtable"], ref["dtable"].numpy(), "dtable", cos_min=cm, rel_max=rm)
    assert abs(gr["ds"] - ref["ds"].item()) <= rm * abs(ref["ds"].item()) + 2e-3


def test_spatial_max_kernel_exact_on_bf16_inputs(cv):
    """kernel-level: with bf16-representable inputs the only difference to the fp64 oracle is the
    fp32 accumulation order, so match agrees to ~1e-6 and the saved argmax is exact (no near-ties)."""
    rng = np.random.RandomState(3)
    Bi, Bt, L, HW, E = 23, 17, 25, 49, 512
    img = torch.nn.functional.normalize(t(rng.standard_normal((Bi, HW, E)).astype(np.float32)), dim=-1)
    tok = torch.nn.functional.normalize(t(rng.standard_normal((Bt, L, E)).astype(np.float32)), dim=-1)
    img = img.to(torch.bfloat16).float(); tok = tok.to(torch.bfloat16).float()
    lens = t(rng.randint(3, L + 1, size=Bt).astype(np.int64))
    for b in range(Bt):
        tok[b, lens[b]:] = 0
    ref = O.similarity_spatial_max(img.double().permute(0, 2, 1).reshape(Bi, E, 7, 7), tok.double(), lens)
    match, a_it, a_ti = cv.ops.spatial_max_fwd(img.to(DEV).to(torch.bfloat16), tok.to(DEV).to(torch.bfloat16),
                                               lens.to(DEV))
    assert float((match.cpu().double() - ref).abs().max()) <= 2e-6
    mm = torch.einsum('ihe,tle->itlh', img.double(), tok.double())
    ref_arg = mm.argmax(-1)                                     # [Bi,Bt,L]
    got = a_it.cpu().view(Bi, Bt, L).long()
    valid = (torch.arange(L)[None, :] < lens[:, None])[None].expand(Bi, -1, -1)
    assert torch.equal(got[valid], ref_arg[valid])
    assert torch.equal(a_ti.cpu().view(Bt, L, Bi).permute(2, 0, 1), a_it.cpu().view(Bi, Bt, L))


def test_spatial_max_backward_exact_given_argmax(cv):
    """backward of the max similarity on bf16-representable inputs (kernel argmax == fp64 argmax):
    must equal torch autograd of einsum+amax+sum+div (multimodal.py:775-780) to fp32 accuracy."""
    rng = np.random.RandomState(4)
    Bi, Bt, L, HW, E = 19, 21, 25, 49, 256
    img = torch.nn.functional.normalize(t(rng.standard_normal((Bi, HW, E)).astype(np.float32)), dim=-1)
    tok = torch.nn.functional.normalize(t(rng.standard_normal((Bt, L, E)).astype(np.float32)), dim=-1)
    img = img.to(torch.bfloat16).float(); tok = tok.to(torch.bfloat16).float()
    ids, lens = O.synth_tokens(rng, Bt, L, 2350)
    for b in range(Bt):
        tok[b, lens[b]:] = 0
    g = t(rng.standard_normal((Bi, Bt)).astype(np.float32))
    ir = img.double().clone().requires_grad_(True); tr = tok.double().clone().requires_grad_(True)
    ref = O.similarity_spatial_max(ir.permute(0, 2, 1).reshape(Bi, E, 7, 7), tr, t(lens))
    (ref * g.double()).sum().backward()
    idv = img.to(DEV).requires_grad_(True); tdv = tok.to(DEV).requires_grad_(True)
    got = cv.ops.spatial_max_similarity(idv, tdv, t(lens, DEV), t(ids, DEV))
    (got * g.to(DEV)).sum().backward()
    assert float((got.detach().cpu().double() - ref.detach()).abs().max()) <= 2e-6
    assert rel_fro(idv.grad.cpu().numpy(), ir.grad.numpy()) <= 1e-5
    valid = (torch.arange(L)[None, :] < t(lens)[:, None])
    assert rel_fro(tdv.grad.cpu()[valid].numpy(), tr.grad[valid].numpy()) <= 1e-5


def test_spatial_max_backward_mma_form_equals_gather_form(cv):
    """tensor-core backward (P expansion + two GEMMs, bf16 P) == SIMT gather backward == autograd."""
    rng = np.random.RandomState(6)
    Bi, Bt, L, HW, E = 70, 66, 25, 49, 512
    img = torch.nn.functional.normalize(t(rng.standard_normal((Bi, HW, E)).astype(np.float32)), dim=-1)
    tok = torch.nn.functional.normalize(t(rng.standard_normal((Bt, L, E)).astype(np.float32)), dim=-1)
    img = img.to(torch.bfloat16).float(); tok = tok.to(torch.bfloat16).float()
    ids, lens = O.synth_tokens(rng, Bt, L, 2350)
    for b in range(Bt):
        tok[b, lens[b]:] = 0
    g = t(rng.standard_normal((Bi, Bt)).astype(np.float32))
    res = {}
    for flag in (True, False):
        cv.ops.SPATIAL_MAX_BWD_MMA = flag
        idv = img.to(DEV).requires_grad_(True); tdv = tok.to(DEV).requires_grad_(True)
        got = cv.ops.spatial_max_similarity(idv, tdv, t(lens, DEV), t(ids, DEV))
        (got * g.to(DEV)).sum().backward()
        res[flag] = (idv.grad.cpu().numpy(), tdv.grad.cpu().numpy())
    cv.ops.SPATIAL_MAX_BWD_MMA = True
    valid = (torch.arange(L)[None, :] < t(lens)[:, None]).numpy()
    assert rel_fro(res[True][0], res[False][0]) <= 4e-3          # P coefficients are bf16
    assert rel_fro(res[True][1][valid], res[False][1][valid]) <= 4e-3
    ir = img.double().clone().requires_grad_(True); tr = tok.double().clone().requires_grad_(True)
    ref = O.similarity_spatial_max(ir.permute(0, 2, 1).reshape(Bi, E, 7, 7), tr, t(lens))
    (ref * g.double()).sum().backward()
    assert rel_fro(res[True][0], ir.grad.numpy()) <= 4e-3
    assert rel_fro(res[True][1][valid], tr.grad.numpy()[valid]) <= 4e-3


@pytest.mark.parametrize("sim", ["mean", "max"])
def test_graphed_spatial_step_equals_eager(cv, sim):
    """GraphedLossStep: the whole spatial train step (head on the 7x7 map, text tokens, similarity, InfoNCE, backward)
    captured as one CUDA graph == the eager module call, and a replay picks up new inputs in the static buffers."""
    E, B = 512, 24
    inps = [case_inputs(900 + k, B, E, "spatial") for k in range(2)]
    m = build(cv, E, sim, inps[0], fix_temperature=True)
    m.materialize_logits = m.materialize_text_outputs = False
    conv = m.image_embed.model[-1]
    params = [conv.weight, conv.bias, m.text_embed.embedding.weight]

    def eager(inp):
        for p in params:
            p.grad = None
        out = m.calculate_contrastive_loss(t(inp["f"], DEV), t(inp["ids"], DEV), t(inp["lens"], DEV))
        out[0].backward()
        return out[0].item(), [p.grad.clone() for p in params]
    ref = [eager(i) for i in inps]
    x = t(inps[0]["f"], DEV).clone(); ids = t(inps[0]["ids"], DEV).clone(); lens = t(inps[0]["lens"], DEV).clone()
    step = cv.GraphedLossStep(lambda: m.calculate_contrastive_loss(x, ids, lens)[0], params)
    for k in (0, 1, 0):
        x.copy_(t(inps[k]["f"], DEV)); ids.copy_(t(inps[k]["ids"], DEV)); lens.copy_(t(inps[k]["lens"], DEV))
        loss = step()
        assert abs(loss.item() - ref[k][0]) <= 1e-5 * abs(ref[k][0]), (k, loss.item(), ref[k][0])
        for p, g in zip(params, ref[k][1]):
            assert rel_fro(p.grad.cpu().numpy(), g.cpu().numpy()) <= 1e-4, k       # float-atomic order only


def test_spatial_config4_size_subsampled(cv):
    """BASELINE config 4 (B = 1024, 7x7 map, L = 25): the spatial max / mean paths at their stated size.  The oracle
    would need the [B, B, L, 49] score tensor, so the checks are on sub-blocks against fp64 torch (computed on the
    GPU, independent of the kernels): match values and arg-max locations for 24 sampled images x all texts, the
    tensor-core backward (compacted token rows) against the SIMT gather backward on the whole problem, d img rows of
    the sampled images against the closed form from the fp64 arg-max, and the mean-path loss against fp64."""
    rng = np.random.RandomState(41)
    B, L, HW, E = 1024, 25, 49, 512
    gen = torch.Generator().manual_seed(41)
    img = torch.nn.functional.normalize(torch.randn(B, HW, E, generator=gen), dim=-1).to(torch.bfloat16).float()
    tok = torch.nn.functional.normalize(torch.randn(B, L, E, generator=gen), dim=-1).to(torch.bfloat16).float()
    ids, lens = O.synth_tokens(rng, B, L, 2350)
    lens_t = t(lens)
    valid = torch.arange(L)[None, :] < lens_t[:, None]                  # [B, L]
    tok = tok * valid[:, :, None]
    g = torch.randn(B, B, generator=gen) / B
    imd, tkd, ld, idd, gd = img.to(DEV), tok.to(DEV), lens_t.to(DEV), t(ids, DEV), g.to(DEV)
    res = {}
    for flag in (True, False):
        cv.ops.SPATIAL_MAX_BWD_MMA = flag
        iv = imd.clone().requires_grad_(True); tv = tkd.clone().requires_grad_(True)
        match = cv.ops.spatial_max_similarity(iv, tv, ld, idd)
        (match * gd).sum().backward()
        res[flag] = (match.detach(), iv.grad, tv.grad)
    cv.ops.SPATIAL_MAX_BWD_MMA = True
    torch.cuda.synchronize()
    # forward on sampled images, fp64
    sel = torch.from_numpy(rng.choice(B, 24, replace=False)).to(DEV)
    sc = torch.einsum('ihe,tle->itlh', imd[sel].double(), tkd.double())          # [24, B, L, 49]
    mx, am = sc.max(dim=-1)
    vd = valid.to(DEV)
    ref_match = (mx * vd[None]).sum(-1) / ld[None].double()
    assert float((res[True][0][sel].double() - ref_match).abs().max()) <= 2e-6
    # the two backward forms agree on the whole problem (P coefficients are bf16 in the tensor-core form)
    assert rel_fro(res[True][1].cpu().numpy(), res[False][1].cpu().numpy()) <= 4e-3
    vm = valid.numpy()
    assert rel_fro(res[True][2].cpu().numpy()[vm], res[False][2].cpu().numpy()[vm]) <= 4e-3
    assert float(res[True][2].cpu()[~valid].abs().max()) == 0.0                  # pad rows: exact zeros
    # d img of the sampled images from the fp64 arg-max: dimg[i, h] = sum_{t,l valid} [am = h] g[i,t]/len[t] tok[t,l]
    coef = (gd[sel].double() / ld[None].double())[:, :, None] * vd[None]         # [24, B, L]
    onehot = torch.nn.functional.one_hot(am, HW).double() * coef[..., None]       # [24, B, L, 49]
    ref_dimg = torch.einsum('itlh,tle->ihe', onehot, tkd.double())
    assert rel_fro(res[False][1][sel].cpu().numpy(), ref_dimg.cpu().numpy()) <= 1e-5
    assert rel_fro(res[True][1][sel].cpu().numpy(), ref_dimg.cpu().numpy()) <= 4e-3
    # mean path: pooled factors -> InfoNCE loss vs fp64 (features rounded to bf16 as the kernels' operands)
    table = torch.randn(2350, E, generator=gen) * 0.1
    table[0] = 0
    tb = table.to(DEV).requires_grad_(True); iv = imd.clone().requires_grad_(True)
    _, tp = cv.ops.text_features_spatial(idd, ld, tb, True, 1.0 / HW, want_tok=False)
    out = cv.ops.sim_infonce(cv.ops.spatial_pool(iv), tp, S_DEFAULT)
    out[0].backward()
    rows = torch.nn.functional.normalize(tb.detach().double()[idd], dim=-1) * (idd != 0)[..., None]
    tp64 = rows.sum(1) / ld[:, None].double() / HW
    ip64 = imd.double().sum(1)
    logits = float(np.exp(S_DEFAULT)) * ip64.to(torch.bfloat16).double() @ tp64.to(torch.bfloat16).double().T
    lab = torch.arange(B, device=DEV)
    ref_loss = 0.5 * (torch.nn.functional.cross_entropy(logits, lab) + torch.nn.functional.cross_entropy(logits.T, lab))
    assert abs(out[0].item() - ref_loss.item()) <= 1e-3 * abs(ref_loss.item())
    assert iv.grad is not None and tb.grad is not None and not tb.grad[0].any()
