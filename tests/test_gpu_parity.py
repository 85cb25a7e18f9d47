"""GPU parity suite (-m gpu): the CUDA path, called through the C ABI (ctypes -> libcvcl_b200.so),
against the CPU oracle on identical seeded inputs and against the committed golden vectors that the
unmodified reference produced.  Tolerances (BASELINE.json north_star / SURVEY 8d):
  token / length handling, eval argmax ....... bit-exact
  fp32 kernels (K1, K7) ........................ <= 2e-6 abs on unit-norm features
  loss ......................................... <= 1e-3 relative
  logits (bf16 operands, fp32 accumulate) ...... <= 1e-2 of max|logit|
  gradients .................................... cosine >= 0.999 and rel-Frobenius <= 2e-2
"""
import math

import numpy as np
import pytest
import torch

from _util import (O, S_DEFAULT, assert_grad_close, assert_logits_close, case_inputs, golden, t,
                   oracle_flat_step, rel_fro)

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def cv():
    import multimodal_baby_b200 as m
    m._cabi.load()
    return m


def dev_inputs(inp):
    return {k: t(v, DEV) for k, v in inp.items()}


# ------------------------------------------------------------------------------- K1
@pytest.mark.parametrize("B,L,E", [(8, 25, 64), (33, 25, 512), (512, 25, 512), (5, 7, 128), (3, 25, 1024)])
def test_text_encoder_flat_matches_oracle(cv, B, L, E):
    rng = np.random.RandomState(B + E)
    _, _, table = O.synth_weights(rng, E, 8, 2350)
    ids, lens = O.synth_tokens(rng, B, L, 2350, min_len=min(3, L))
    for norm in (True, False):
        ref, _ = O.encode_text(t(ids), t(lens), t(table), "flat", norm)
        got = cv.ops.text_features_flat(t(ids, DEV), t(lens, DEV), t(table, DEV), norm)
        tol = 2e-6 if norm else 2e-6 * float(ref.abs().max())
        assert float((got.cpu() - ref).abs().max()) <= tol


def test_text_encoder_edge_cases(cv):
    rng = np.random.RandomState(11)
    _, _, table = O.synth_weights(rng, 512, 8, 2350)
    L = 25
    ids = np.zeros((4, L), np.int64)
    ids[0, :L] = rng.randint(4, 2350, L)            # full length, no padding
    ids[1, 0] = 17                                  # single token
    ids[2, :3] = [2, 2349, 3]                       # last vocabulary row
    ids[3, :2] = [1, 1]                             # repeated <unk>
    lens = np.array([L, 1, 3, 2], np.int64)
    ref, _ = O.encode_text(t(ids), t(lens), t(table), "flat", True)
    got = cv.ops.text_features_flat(t(ids, DEV), t(lens, DEV), t(table, DEV), True)
    assert float((got.cpu() - ref).abs().max()) <= 2e-6
    # the divisor is len, not the number of non-pad tokens (multimodal.py:503)
    lens2 = np.array([L, 5, 3, 2], np.int64)
    ref2, _ = O.encode_text(t(ids), t(lens2), t(table), "flat", False)
    got2 = cv.ops.text_features_flat(t(ids, DEV), t(lens2, DEV), t(table, DEV), False)
    assert float((got2.cpu() - ref2).abs().max()) <= 1e-6
    # empty batch
    e = cv.ops.text_features_flat(torch.zeros((0, L), dtype=torch.int64, device=DEV),
                                  torch.zeros((0,), dtype=torch.int64, device=DEV), t(table, DEV), True)
    assert e.shape == (0, 512)


def test_text_outputs_gather_bit_exact(cv):
    rng = np.random.RandomState(12)
    _, _, table = O.synth_weights(rng, 512, 8, 2350)
    ids, lens = O.synth_tokens(rng, 16)
    got = cv.ops.text_outputs(t(ids, DEV), t(table, DEV))
    assert torch.equal(got.cpu(), O.embedding_lookup(t(ids), t(table)))


def test_text_encoder_spatial_tokens(cv):
    rng = np.random.RandomState(13)
    _, _, table = O.synth_weights(rng, 512, 8, 2350)
    ids, lens = O.synth_tokens(rng, 9)
    ref, _ = O.encode_text(t(ids), t(lens), t(table), "spatial", True)
    tok, pooled = cv.ops.text_features_spatial(t(ids, DEV), t(lens, DEV), t(table, DEV), True, 1.0 / 49)
    assert float((tok.cpu() - ref).abs().max()) <= 2e-6
    refp = ref.sum(1) / (49 * t(lens)[:, None])
    assert float((pooled.cpu() - refp).abs().max()) <= 2e-6


def test_text_encoder_backward_matches_oracle(cv):
    rng = np.random.RandomState(14)
    _, _, table = O.synth_weights(rng, 256, 8, 500)
    ids, lens = O.synth_tokens(rng, 32, 25, 500)
    g = rng.standard_normal((32, 256)).astype(np.float32)
    tr = t(table).requires_grad_(True)
    ref, _ = O.encode_text(t(ids), t(lens), tr, "flat", True)
    (ref * t(g)).sum().backward()
    ref_grad = tr.grad.clone(); ref_grad[0] = 0
    td = t(table, DEV).requires_grad_(True)
    got = cv.ops.text_features_flat(t(ids, DEV), t(lens, DEV), td, True)
    (got * t(g, DEV)).sum().backward()
    assert rel_fro(td.grad.cpu().numpy(), ref_grad.numpy()) <= 1e-5
    assert not td.grad[0].any()


# ------------------------------------------------------------------------------- K2
@pytest.mark.parametrize("M,E", [(8, 64), (32, 512), (200, 512), (512, 512), (4, 512), (130, 320)])
def test_head_features_matches_oracle(cv, M, E):
    rng = np.random.RandomState(M + E)
    W, b, _ = O.synth_weights(rng, E, 2048, 8)
    f = O.synth_trunk_features(rng, (M, 2048))
    ref = O.encode_image(t(f), t(W), t(b), "flat", True)
    got = cv.ops.head_features(t(f, DEV), t(W, DEV), t(b, DEV), True).cpu()
    # bf16 operands: unit-norm rows agree to ~3e-3 per element, cosine per row ~1
    assert float((got - ref).abs().max()) <= 6e-3
    cos = (got * ref).sum(1)
    assert float(cos.min()) >= 0.9999
    assert float((got.norm(dim=1) - 1).abs().max()) <= 1e-5
    refu = O.encode_image(t(f), t(W), t(b), "flat", False)
    gotu = cv.ops.head_features(t(f, DEV), t(W, DEV), t(b, DEV), False).cpu()
    assert float((gotu - refu).abs().max()) <= 1e-2 * float(refu.abs().max())


def test_head_backward_matches_oracle(cv):
    rng = np.random.RandomState(21)
    W, b, _ = O.synth_weights(rng, 512, 2048, 8)
    f = O.synth_trunk_features(rng, (96, 2048))
    g = rng.standard_normal((96, 512)).astype(np.float32)
    Wr, br, fr = t(W).requires_grad_(True), t(b).requires_grad_(True), t(f).requires_grad_(True)
    (O.encode_image(fr, Wr, br, "flat", True) * t(g)).sum().backward()
    Wd, bd, fd = (t(W, DEV).requires_grad_(True), t(b, DEV).requires_grad_(True),
                  t(f, DEV).requires_grad_(True))
    (cv.ops.head_features(fd, Wd, bd, True) * t(g, DEV)).sum().backward()
    assert_grad_close(Wd.grad.cpu().numpy(), Wr.grad.numpy(), "dW")
    assert_grad_close(bd.grad.cpu().numpy(), br.grad.numpy(), "db")
    assert_grad_close(fd.grad.cpu().numpy(), fr.grad.numpy(), "df")


# ------------------------------------------------------------------------------- K3
@pytest.mark.parametrize("name", ["forward_4x3_e512", "forward_4x1_e512"])
def test_forward_logits_golden_ni_ne_nt(cv, name):
    g = golden(name)
    inp = case_inputs(int(g["seed"]), int(g["Ni"]), int(g["E"]), "flat", Bt=int(g["Nt"]))
    d = dev_inputs(inp)
    img = cv.ops.head_features(d["f"], d["W"], d["b"], True)
    txt = cv.ops.text_features_flat(d["ids"], d["lens"], d["table"], True)
    lpi, lpt = cv.ops.sim_logits(img, txt, S_DEFAULT)
    assert lpi.shape == (int(g["Ni"]), int(g["Nt"])) and lpt.shape == (int(g["Nt"]), int(g["Ni"]))
    assert_logits_close(lpi.cpu().numpy(), g["logits_per_image"])
    assert_logits_close(lpt.cpu().numpy(), g["logits_per_text"])
    assert torch.equal(lpi.t().contiguous(), lpt)


def test_sim_logits_backward(cv):
    rng = np.random.RandomState(31)
    Ni, Nt, E = 70, 45, 512
    img = torch.nn.functional.normalize(t(rng.standard_normal((Ni, E)).astype(np.float32)), dim=1)
    txt = torch.nn.functional.normalize(t(rng.standard_normal((Nt, E)).astype(np.float32)), dim=1)
    g1 = t(rng.standard_normal((Ni, Nt)).astype(np.float32)); g2 = t(rng.standard_normal((Nt, Ni)).astype(np.float32))
    ir, tr = img.clone().requires_grad_(True), txt.clone().requires_grad_(True)
    sr = torch.tensor(S_DEFAULT, requires_grad=True)
    a, b = O.logits_from_match(O.similarity_flat(ir, tr), sr)
    ((a * g1).sum() + (b * g2).sum()).backward()
    idv, tdv = img.to(DEV).requires_grad_(True), txt.to(DEV).requires_grad_(True)
    sd = torch.tensor(S_DEFAULT, device=DEV, requires_grad=True)
    a2, b2 = cv.ops.sim_logits(idv, tdv, sd)
    ((a2 * g1.to(DEV)).sum() + (b2 * g2.to(DEV)).sum()).backward()
    assert_grad_close(idv.grad.cpu().numpy(), ir.grad.numpy(), "dimg")
    assert_grad_close(tdv.grad.cpu().numpy(), tr.grad.numpy(), "dtxt")
    # ds = sum g*logits with random-sign g: compare against the magnitude of the summands
    mag = float((g1 * a.detach()).abs().sum() + (g2 * b.detach()).abs().sum())
    assert abs(sd.grad.item() - sr.grad.item()) <= 2e-3 * mag


# ------------------------------------------------------------------------------- fused step
def check_step(cv, inp, ref, normalize=True, loss_tol=1e-3):
    d = dev_inputs(inp)
    W = d["W"].requires_grad_(True); b = d["b"].requires_grad_(True)
    table = d["table"].requires_grad_(True)
    s = torch.tensor(S_DEFAULT, device=DEV, requires_grad=True)
    loss, iacc, tacc, ient, tent, img_f, txt_f = cv.ops.flat_contrastive_loss(
        d["f"], d["ids"], d["lens"], W, b, table, s, normalize, want_features=True)
    loss.backward()
    torch.cuda.synchronize()
    B = inp["f"].shape[0]
    assert abs(loss.item() - float(ref["loss"])) <= loss_tol * abs(float(ref["loss"]))
    assert abs(ient.item() - float(ref["image_entropy"])) <= 2e-3 * max(1.0, float(ref["image_entropy"]))
    assert abs(tent.item() - float(ref["text_entropy"])) <= 2e-3 * max(1.0, float(ref["text_entropy"]))
    # in-batch accuracy is not bit-stable under bf16 at random init
    assert abs(iacc.item() - float(ref["image_accuracy"])) <= max(2.0 / B, 0.02)
    assert abs(tacc.item() - float(ref["text_accuracy"])) <= max(2.0 / B, 0.02)
    if "image_features" in ref:
        assert float((img_f.cpu() - torch.as_tensor(np.asarray(ref["image_features"]))).abs().max()) <= 6e-3
        assert float((txt_f.cpu() - torch.as_tensor(np.asarray(ref["text_features"]))).abs().max()) <= 1e-5
    return dict(dW=W.grad.cpu().numpy(), db=b.grad.cpu().numpy(), dtable=table.grad.cpu().numpy(),
                ds=s.grad.item(), img=img_f, txt=txt_f)


@pytest.mark.parametrize("name", ["flat_e64_b8", "flat_e512_b32", "flat_e512_b160"])
def test_flat_step_vs_reference_golden(cv, name):
    g = golden(name)
    inp = case_inputs(int(g["seed"]), int(g["B"]), int(g["E"]), "flat")
    got = check_step(cv, inp, g)
    E = int(g["E"])
    assert_grad_close(got["db"], g["db"], "db")
    assert abs(got["ds"] - float(g["ds"])) <= 2e-2 * abs(float(g["ds"])) + 1e-3
    assert rel_fro(got["dW"][:8, :64], g["dW_slice"]) <= 3e-2
    assert abs(np.linalg.norm(got["dW"]) - float(g["dW_norm"])) <= 2e-2 * float(g["dW_norm"])
    assert rel_fro(got["dtable"][:8], g["dtable_rows"]) <= 3e-2
    assert abs(np.linalg.norm(got["dtable"]) - float(g["dtable_norm"])) <= 2e-2 * float(g["dtable_norm"])
    assert not got["dtable"][0].any()
    if "dW" in g:
        assert_grad_close(got["dW"], g["dW"], "dW")
    # logits from the returned features
    lpi, lpt = cv.ops.sim_logits(got["img"], got["txt"], S_DEFAULT)
    assert_logits_close(lpi.cpu().numpy(), g["logits_per_image"])
    assert_logits_close(lpt.cpu().numpy(), g["logits_per_text"])


@pytest.mark.parametrize("B,E", [(512, 512), (300, 512), (128, 256), (1024, 512)])
def test_flat_step_vs_oracle(cv, B, E):
    inp = case_inputs(1000 + B, B, E, "flat")
    ref = oracle_flat_step(inp)
    got = check_step(cv, inp, {k: (v.numpy() if torch.is_tensor(v) else v) for k, v in ref.items()})
    assert_grad_close(got["dW"], ref["dW"].numpy(), "dW")
    assert_grad_close(got["db"], ref["db"].numpy(), "db")
    assert_grad_close(got["dtable"], ref["dtable"].numpy(), "dtable")
    assert abs(got["ds"] - ref["ds"].item()) <= 2e-2 * abs(ref["ds"].item()) + 1e-3


def test_flat_step_unnormalized(cv):
    inp = case_inputs(77, 64, 128, "flat")
    inp["table"] *= 0.05                      # keep un-normalised logits in a sane range
    ref = oracle_flat_step(inp, s=0.0, normalize=False)
    d = dev_inputs(inp)
    W = d["W"].requires_grad_(True); table = d["table"].requires_grad_(True)
    b = d["b"].requires_grad_(True)
    loss, *_ = cv.ops.flat_contrastive_loss(d["f"], d["ids"], d["lens"], W, b, table, 0.0, False)
    loss.backward()
    assert abs(loss.item() - ref["loss"].item()) <= 2e-3 * abs(ref["loss"].item())
    assert_grad_close(W.grad.cpu().numpy(), ref["dW"].numpy(), "dW", rel_max=3e-2)
    assert_grad_close(table.grad.cpu().numpy(), ref["dtable"].numpy(), "dtable", rel_max=3e-2)


def test_ops_path_equals_fused_path(cv):
    """op-by-op autograd path (head -> text -> sim_infonce) == the single fused C call."""
    inp = case_inputs(55, 256, 512, "flat")
    d = dev_inputs(inp)
    res = []
    for fused in (True, False):
        W = d["W"].clone().requires_grad_(True); b = d["b"].clone().requires_grad_(True)
        table = d["table"].clone().requires_grad_(True)
        s = torch.tensor(S_DEFAULT, device=DEV, requires_grad=True)
        if fused:
            loss = cv.ops.flat_contrastive_loss(d["f"], d["ids"], d["lens"], W, b, table, s, True)[0]
        else:
            img = cv.ops.head_features(d["f"], W, b, True)
            txt = cv.ops.text_features_flat(d["ids"], d["lens"], table, True)
            loss = cv.ops.sim_infonce(img, txt, s)[0]
        loss.backward()
        res.append((loss.item(), W.grad.cpu().numpy(), b.grad.cpu().numpy(), table.grad.cpu().numpy(),
                    s.grad.item()))
    assert abs(res[0][0] - res[1][0]) <= 1e-5 * abs(res[0][0])
    for i, n in ((1, "dW"), (2, "db"), (3, "dtable")):
        assert_grad_close(res[0][i], res[1][i], n, cos_min=0.9999, rel_max=1e-2)
    assert abs(res[0][4] - res[1][4]) <= 1e-2 * abs(res[1][4]) + 1e-4


def test_sim_infonce_properties_full_size(cv):
    """size-independent properties at B = 8192 (the oracle would need B^2 fp32 on the host):
    (1) identical pairs with orthogonal-ish features -> loss known in closed form,
    (2) permuting the pairs leaves loss / entropy unchanged,
    (3) uniform logits -> loss = ln B, entropy = ln B, ds = 0."""
    B, E = 8192, 512
    gen = torch.Generator(device="cpu").manual_seed(5)
    x = torch.nn.functional.normalize(torch.randn(B, E, generator=gen), dim=1).to(DEV)
    y = torch.nn.functional.normalize(torch.randn(B, E, generator=gen), dim=1).to(DEV)
    l1 = cv.ops.sim_infonce(x, y, S_DEFAULT)
    perm = torch.randperm(B, generator=gen).to(DEV)
    l2 = cv.ops.sim_infonce(x[perm], y[perm], S_DEFAULT)
    assert abs(l1[0].item() - l2[0].item()) <= 1e-5 * l1[0].item()
    assert abs(l1[3].item() - l2[3].item()) <= 1e-5 * l1[3].item()
    # uniform: all features identical -> all logits equal
    u = torch.nn.functional.normalize(torch.ones(1, E), dim=1).expand(B, E).contiguous().to(DEV)
    s = torch.tensor(S_DEFAULT, device=DEV, requires_grad=True)
    lu = cv.ops.sim_infonce(u, u.clone(), s)
    assert abs(lu[0].item() - math.log(B)) <= 1e-4 * math.log(B)
    assert abs(lu[3].item() - math.log(B)) <= 1e-4 * math.log(B)
    lu[0].backward()
    assert abs(s.grad.item()) <= 1e-3
    # fp64 reference on a row subset: lse of 64 sampled rows
    xs = x[:64].double().cpu(); ya = y.double().cpu()
    xb = x[:64].to(torch.bfloat16).double().cpu(); yb = y.to(torch.bfloat16).double().cpu()
    ref_rows = torch.logsumexp(math.exp(S_DEFAULT) * xb @ yb.T, dim=1)
    out5, lse0, lse1, a0, a1 = cv.ops.sim_infonce_fwd(x.to(torch.bfloat16), y.to(torch.bfloat16),
                                                      y.to(torch.bfloat16), x.to(torch.bfloat16),
                                                      S_DEFAULT, 0, 1.0 / B)
    assert float((lse0[:64].double().cpu() - ref_rows).abs().max()) <= 1e-4
    ref_arg = torch.argmax(xb @ yb.T, dim=1)
    assert (a0[:64].cpu().long() == ref_arg).float().mean().item() >= 0.98


def test_sim_infonce_one_pass_equals_two_pass(cv):
    """unit_norm=True at >= 4096 pairs: ONE similarity pass yields the statistics of both directions (row partials
    per thread, column partials through the lane butterfly, fixed-reference exponent) == the two-GEMM form, and both
    == fp64 on sampled rows and columns; structured inputs (planted positives, duplicated rows -> arg-max ties)."""
    B, E = 8192, 512
    gen = torch.Generator(device="cpu").manual_seed(17)
    x = torch.nn.functional.normalize(torch.randn(B, E, generator=gen), dim=1)
    y = torch.nn.functional.normalize(0.6 * x + 0.8 * torch.nn.functional.normalize(torch.randn(B, E, generator=gen), dim=1), dim=1)
    y[100] = y[7]; x[4000] = x[4001]                     # exact ties for the arg-max of text 7/100 and image rows
    xb, yb = x.to(torch.bfloat16).to(DEV), y.to(torch.bfloat16).to(DEV)
    one = cv.ops.sim_infonce_fwd(xb, yb, yb, xb, S_DEFAULT, 0, 1.0 / B, True)
    import os
    os.environ["CVCL_B200_SIM_TWO_PASS"] = "1"
    try:
        two = cv.ops.sim_infonce_fwd(xb, yb, yb, xb, S_DEFAULT, 0, 1.0 / B, True)
    finally:
        del os.environ["CVCL_B200_SIM_TWO_PASS"]
    torch.cuda.synchronize()
    for a, b_ in zip(one[0][:5].tolist(), two[0][:5].tolist()):
        assert abs(a - b_) <= 2e-5 * max(1.0, abs(b_)), (one[0], two[0])
    assert float((one[1] - two[1]).abs().max()) <= 2e-4 and float((one[2] - two[2]).abs().max()) <= 2e-4
    assert torch.equal(one[3], two[3]) and torch.equal(one[4], two[4])
    # fp64 on a subset: rows (images) and columns (texts)
    S64 = math.exp(S_DEFAULT) * (xb[:96].double().cpu() @ yb.double().cpu().T)
    assert float((one[1][:96].double().cpu() - torch.logsumexp(S64, dim=1)).abs().max()) <= 1e-4
    assert torch.equal(one[3][:96].cpu().long(), torch.argmax(S64, dim=1))
    S64t = math.exp(S_DEFAULT) * (yb[:128].double().cpu() @ xb.double().cpu().T)
    assert float((one[2][:128].double().cpu() - torch.logsumexp(S64t, dim=1)).abs().max()) <= 1e-4
    assert torch.equal(one[4][:128].cpu().long(), torch.argmax(S64t, dim=1))
    # the whole op through autograd (the model's flat "ops" path passes unit_norm for normalised features)
    xi = x.to(DEV).requires_grad_(True); yi = y.to(DEV).requires_grad_(True)
    l1 = cv.ops.sim_infonce(xi, yi, S_DEFAULT, None, True); l1[0].backward()
    xj = x.to(DEV).requires_grad_(True); yj = y.to(DEV).requires_grad_(True)
    l2 = cv.ops.sim_infonce(xj, yj, S_DEFAULT, None, False); l2[0].backward()
    assert abs(l1[0].item() - l2[0].item()) <= 2e-5 * abs(l2[0].item())
    assert rel_fro(xi.grad.cpu().numpy(), xj.grad.cpu().numpy()) <= 1e-3
    assert rel_fro(yi.grad.cpu().numpy(), yj.grad.cpu().numpy()) <= 1e-3


# ------------------------------------------------------------------------------- K7
def test_eval_nway_golden_bit_exact(cv):
    from test_oracle_golden import eval_case_inputs
    g = golden("eval_4way_e512")
    W, b, table, f = eval_case_inputs(g)
    # fp32 end to end: head on the host oracle (exact fp32), kernel does normalise + dot + argmax
    img = O.head_flat(t(f).reshape(-1, 2048), t(W), t(b))
    txt, _ = O.text_encoder_flat(t(g["ids"]), t(g["lens"]), t(table))
    pred, logits = cv.ops.eval_nway(img.to(DEV), txt.to(DEV), None, 4, True, S_DEFAULT)
    assert np.array_equal(pred.cpu().numpy(), g["pred"])
    np.testing.assert_allclose(logits.cpu().numpy(), g["logits"], atol=2e-5)


def test_eval_nway_100k_frames_vs_oracle(cv):
    """config 5: 25 000 trials x 4 frames, 22 categories; argmax must equal the fp32 oracle except
    where the oracle's own top-2 gap is below fp32 resolution (reported, not tolerated silently)."""
    rng = np.random.RandomState(2024)
    N, C, E = 25000, 22, 512
    img = rng.standard_normal((N * 4, E)).astype(np.float32)
    cat = rng.standard_normal((C, E)).astype(np.float32)
    idx = rng.randint(0, C, size=N).astype(np.int32)
    pred, logits = cv.ops.eval_nway(t(img, DEV), t(cat, DEV), t(idx, DEV), 4, True, S_DEFAULT)
    ref_pred, ref_logits = O.eval_nway(t(img).reshape(N, 4, E), t(cat)[t(idx).long()], S_DEFAULT)
    mism = np.nonzero(pred.cpu().numpy() != ref_pred.numpy().astype(np.int32))[0]
    if len(mism):
        l64 = O.eval_nway(t(img).double().reshape(N, 4, E), t(cat).double()[t(idx).long()], S_DEFAULT)[1]
        top2 = torch.topk(l64[mism], 2, dim=1).values
        gaps = (top2[:, 0] - top2[:, 1]).abs() / top2[:, 0].abs()
        assert float(gaps.max()) < 1e-6, f"{len(mism)} argmax mismatches, fp64 top-2 rel gap {gaps.max():.3e}"
    np.testing.assert_allclose(logits.cpu().numpy(), ref_logits.numpy(), atol=2e-5)
    # predictions only (screen on raw dots, near-ties through the reference arithmetic): identical
    pred2, empty = cv.ops.eval_nway(t(img, DEV), t(cat, DEV), t(idx, DEV), 4, True, S_DEFAULT, False)
    assert empty.numel() == 0 and torch.equal(pred2, pred)


@pytest.mark.parametrize("n_trials", [4099, 4097, 64, 3])
def test_eval_nway_ragged_counts_and_ties(cv, n_trials):
    """trial counts that are not a multiple of the streaming kernel's stage (4 trials), with exact ties
    (duplicated candidate rows -> first index wins, torch.argmax semantics) and zero rows."""
    rng = np.random.RandomState(n_trials)
    E = 512
    img = rng.standard_normal((n_trials * 4, E)).astype(np.float32)
    img[5] = img[4]                       # trial 1: candidates 0 and 1 identical
    img[8:12] = 0.0                       # trial 2: all-zero candidates (norm clamp 1e-12)
    cat = rng.standard_normal((n_trials, E)).astype(np.float32)
    cat[1] = img[4]                       # ... and they are the best match: an exact tie at the maximum
    ref_pred, ref_logits = O.eval_nway(t(img).reshape(n_trials, 4, E), t(cat), S_DEFAULT)
    for want in (True, False):
        pred, logits = cv.ops.eval_nway(t(img, DEV), t(cat, DEV), None, 4, True, S_DEFAULT, want)
        p = pred.cpu().numpy()
        mism = np.nonzero(p != ref_pred.numpy().astype(np.int32))[0]
        l64 = O.eval_nway(t(img).double().reshape(n_trials, 4, E), t(cat).double(), S_DEFAULT)[1]
        for i in mism:                    # only fp32-unresolvable gaps may differ
            top2 = torch.topk(l64[i], 2).values
            assert float((top2[0] - top2[1]).abs()) <= 1e-6 * max(1.0, float(top2[0].abs())), (i, l64[i])
        if n_trials > 2:
            assert p[1] == 0 and ref_pred[1].item() == 0 and p[2] == 0
        if want:
            np.testing.assert_allclose(logits.cpu().numpy(), ref_logits.numpy(), atol=2e-5)


# ------------------------------------------------------------------------------- error behaviour
def test_errors_are_loud(cv):
    x = torch.randn(4, 2048); ids = torch.zeros(4, 25, dtype=torch.int64); lens = torch.ones(4, dtype=torch.int64)
    W = torch.randn(64, 2048); b = torch.randn(64); table = torch.randn(100, 64)
    with pytest.raises(RuntimeError, match="no CPU"):
        cv.ops.flat_contrastive_loss(x, ids, lens, W, b, table, 0.0)
    with pytest.raises(TypeError):
        cv.ops.text_features_flat(ids.int().to(DEV), lens.to(DEV), table.to(DEV))
    with pytest.raises(cv.CvclError):           # E not a multiple of 4
        cv.ops.text_features_flat(ids.to(DEV), lens.to(DEV), torch.randn(100, 30, device=DEV))


def test_graphed_step_equals_eager(cv):
    """GraphedContrastiveStep (H2D + fused step + D2H captured in one CUDA graph) == the eager module
    call, and replays pick up new staged batches."""
    import argparse
    E = 512
    inp = case_inputs(321, 256, E, "flat")
    args = argparse.Namespace(embedding_type="flat", embedding_dim=E, normalize_features=True,
                              fix_temperature=True, temperature=0.07, text_encoder="embedding")
    vocab = {str(i): i for i in range(2350)}
    m = cv.MultiModalModel(cv.VisionEncoder(args, trunk="pooled"), cv.TextEncoder(vocab, 2048, args), args)
    with torch.no_grad():
        m.image_embed.model.fc.weight.copy_(t(inp["W"])); m.image_embed.model.fc.bias.copy_(t(inp["b"]))
        m.text_embed.embedding.weight.copy_(t(inp["table"]))
    m.to(DEV).train()
    m.materialize_logits = m.materialize_text_outputs = m.materialize_features = False
    out = m.calculate_contrastive_loss(t(inp["f"], DEV), t(inp["ids"], DEV), t(inp["lens"], DEV))
    out[0].backward()
    ref = [p.grad.clone() for p in (m.image_embed.model.fc.weight, m.image_embed.model.fc.bias,
                                    m.text_embed.embedding.weight)]
    xh = t(inp["f"]).pin_memory(); ih = t(inp["ids"]).pin_memory(); lh = t(inp["lens"]).pin_memory()
    step = cv.GraphedContrastiveStep(m, xh, ih, lh)
    loss = step()
    assert abs(loss - out[0].item()) <= 1e-6 * abs(loss)
    got = [m.image_embed.model.fc.weight.grad, m.image_embed.model.fc.bias.grad, m.text_embed.embedding.weight.grad]
    for a, b in zip(got, ref):
        assert rel_fro(a.cpu().numpy(), b.cpu().numpy()) <= 1e-4       # float-atomic order only
    # a new batch staged in the pinned buffers is picked up by the next replay
    inp2 = case_inputs(322, 256, E, "flat")
    xh.copy_(t(inp2["f"])); ih.copy_(t(inp2["ids"])); lh.copy_(t(inp2["lens"]))
    loss2 = step()
    out2 = m.calculate_contrastive_loss(t(inp2["f"], DEV), t(inp2["ids"], DEV), t(inp2["lens"], DEV))
    assert abs(loss2 - out2[0].item()) <= 1e-6 * abs(loss2) and abs(loss2 - loss) > 1e-6


def test_graphed_step_prefetch_pipeline(cv):
    """prefetch=True: call k computes on the batch copied during call k-1 and copies the batch staged
    now; losses therefore trail the staged batches by one call."""
    import argparse
    E = 512
    args = argparse.Namespace(embedding_type="flat", embedding_dim=E, normalize_features=True,
                              fix_temperature=True, temperature=0.07, text_encoder="embedding")
    vocab = {str(i): i for i in range(2350)}
    m = cv.MultiModalModel(cv.VisionEncoder(args, trunk="pooled"), cv.TextEncoder(vocab, 2048, args), args)
    inps = [case_inputs(400 + k, 128, E, "flat") for k in range(3)]
    with torch.no_grad():
        m.image_embed.model.fc.weight.copy_(t(inps[0]["W"])); m.image_embed.model.fc.bias.copy_(t(inps[0]["b"]))
        m.text_embed.embedding.weight.copy_(t(inps[0]["table"]))
    m.to(DEV).train()
    m.materialize_logits = m.materialize_text_outputs = m.materialize_features = False
    ref = [m.calculate_contrastive_loss(t(i["f"], DEV), t(i["ids"], DEV), t(i["lens"], DEV))[0].item() for i in inps]
    xh = t(inps[0]["f"]).pin_memory(); ih = t(inps[0]["ids"]).pin_memory(); lh = t(inps[0]["lens"]).pin_memory()
    step = cv.GraphedContrastiveStep(m, xh, ih, lh, prefetch=True)
    step.prime()                                        # batch 0 on the device
    got = []
    for k in (1, 2, 2):                                 # stage batch k, run: computes batch k-1
        xh.copy_(t(inps[k]["f"])); ih.copy_(t(inps[k]["ids"])); lh.copy_(t(inps[k]["lens"]))
        got.append(step())
    for a, b in zip(got, ref):
        assert abs(a - b) <= 1e-6 * abs(b), (got, ref)
    assert m.image_embed.model.fc.weight.grad is not None


def test_graphed_step_packed_staging(cv):
    """staging buffers that are views of one pinned arena (PinnedBatchStager) are moved by ONE H2D copy per step:
    same losses as the three-copy form, in every mode, and new batches are picked up."""
    import argparse
    E = 512
    args = argparse.Namespace(embedding_type="flat", embedding_dim=E, normalize_features=True,
                              fix_temperature=True, temperature=0.07, text_encoder="embedding")
    vocab = {str(i): i for i in range(2350)}
    m = cv.MultiModalModel(cv.VisionEncoder(args, trunk="pooled"), cv.TextEncoder(vocab, 2048, args), args)
    inps = [case_inputs(700 + k, 128, E, "flat") for k in range(4)]
    with torch.no_grad():
        m.image_embed.model.fc.weight.copy_(t(inps[0]["W"])); m.image_embed.model.fc.bias.copy_(t(inps[0]["b"]))
        m.text_embed.embedding.weight.copy_(t(inps[0]["table"]))
    m.to(DEV).train()
    m.materialize_logits = m.materialize_text_outputs = m.materialize_features = False
    ref = [m.calculate_contrastive_loss(t(i["f"], DEV), t(i["ids"], DEV), t(i["lens"], DEV))[0].item() for i in inps]
    st = cv.PinnedBatchStager(128, feat_shape=(2048,), feat_dtype=torch.float32, max_len=inps[0]["ids"].shape[1])
    assert cv.staging.packed_span((st.x_host, st.ids_host, st.lens_host)) is not None

    def stage(step, k):
        step.x_host.copy_(t(inps[k]["f"])); step.ids_host.copy_(t(inps[k]["ids"])); step.lens_host.copy_(t(inps[k]["lens"]))
    stage(st, 0)
    step = cv.GraphedContrastiveStep(m, st.x_host, st.ids_host, st.lens_host)
    assert step.packed
    for k in (0, 1, 2):
        stage(step, k)
        loss = step()
        assert abs(loss - ref[k]) <= 1e-6 * abs(ref[k]), (k, loss, ref[k])
    step = cv.GraphedContrastiveStep(m, st.x_host, st.ids_host, st.lens_host, prefetch=True, lagged_loss=True,
                                     own_staging=True)        # all staging sets allocated (and timed) by the step
    assert step.packed and len(step.staging_probe_us["kept"]) == 2 and step.x_host is not st.x_host
    stage(step, 0)
    step.prime()
    got = []
    for k in (1, 2, 3):
        stage(step, k)
        got.append(step())
    got.append(step.flush())
    assert got[0] != got[0]                             # nothing finished at the first call
    for a, b in zip(got[1:], ref[:3]):
        assert abs(a - b) <= 1e-6 * abs(b), (got, ref)
    # staging health check (re-times the sets, replaces slow ones) and a forced re-capture keep the step intact
    assert step.check_staging() in (0, 1, 2)
    step._capture(1); step._capture(0)
    stage(step, 0); step.prime()
    got = []
    for k in (1, 2):
        stage(step, k)
        got.append(step())
    got.append(step.flush())
    for a, b in zip(got[1:], ref[:2]):
        assert abs(a - b) <= 1e-6 * abs(b), (got, ref)


def test_head_backward_large_m_split_k(cv):
    """M = 128*49 rows (spatial-head shape): the weight gradient takes the split-K path
    (contraction split over blockIdx.z, fp32 vector atomics) and the head GEMM the 2-CTA/SM config."""
    rng = np.random.RandomState(23)
    M, E = 128 * 49, 512
    W, b, _ = O.synth_weights(rng, E, 2048, 8)
    f = O.synth_trunk_features(rng, (M, 2048))
    g = rng.standard_normal((M, E)).astype(np.float32)
    Wr, br = t(W).requires_grad_(True), t(b).requires_grad_(True)
    ref = O.encode_image(t(f), Wr, br, "flat", True)
    (ref * t(g)).sum().backward()
    Wd, bd = t(W, DEV).requires_grad_(True), t(b, DEV).requires_grad_(True)
    got = cv.ops.head_features(t(f, DEV), Wd, bd, True)
    (got * t(g, DEV)).sum().backward()
    assert float((got.detach().cpu() - ref.detach()).abs().max()) <= 6e-3
    assert_grad_close(Wd.grad.cpu().numpy(), Wr.grad.numpy(), "dW")
    assert_grad_close(bd.grad.cpu().numpy(), br.grad.numpy(), "db")


def test_fused_adamw_matches_torch(cv):
    """FusedAdamW == torch.optim.AdamW (the reference's optimiser, multimodal_lit.py:112-128) over a
    few steps on head-shaped tensors, including a length that is not a multiple of 4."""
    gen = torch.Generator().manual_seed(9)
    shapes = [(512, 2048), (512,), (2350, 512), (7,)]
    ref_p = [torch.randn(s, generator=gen).to(DEV).requires_grad_(True) for s in shapes]
    got_p = [p.detach().clone().requires_grad_(True) for p in ref_p]
    ref_opt = torch.optim.AdamW(ref_p, lr=1e-2, weight_decay=0.1)
    got_opt = cv.FusedAdamW(got_p, lr=1e-2, weight_decay=0.1)
    for it in range(4):
        for a, b in zip(ref_p, got_p):
            g = torch.randn(a.shape, generator=gen).to(DEV)
            a.grad = g.clone(); b.grad = g.clone()
        ref_opt.step(); got_opt.step()
    for a, b in zip(ref_p, got_p):
        assert float((a - b).abs().max()) <= 2e-6 * max(1.0, float(a.abs().max()))


def test_graphed_step_lagged_loss(cv):
    """lagged_loss=True: call k returns the loss of replay k-1; flush() returns the last one."""
    import argparse
    E = 512
    args = argparse.Namespace(embedding_type="flat", embedding_dim=E, normalize_features=True,
                              fix_temperature=True, temperature=0.07, text_encoder="embedding")
    vocab = {str(i): i for i in range(2350)}
    m = cv.MultiModalModel(cv.VisionEncoder(args, trunk="pooled"), cv.TextEncoder(vocab, 2048, args), args)
    inps = [case_inputs(500 + k, 128, E, "flat") for k in range(3)]
    with torch.no_grad():
        m.image_embed.model.fc.weight.copy_(t(inps[0]["W"])); m.image_embed.model.fc.bias.copy_(t(inps[0]["b"]))
        m.text_embed.embedding.weight.copy_(t(inps[0]["table"]))
    m.to(DEV).train()
    m.materialize_logits = m.materialize_text_outputs = m.materialize_features = False
    ref = [m.calculate_contrastive_loss(t(i["f"], DEV), t(i["ids"], DEV), t(i["lens"], DEV))[0].item() for i in inps]
    xh = t(inps[0]["f"]).pin_memory(); ih = t(inps[0]["ids"]).pin_memory(); lh = t(inps[0]["lens"]).pin_memory()
    step = cv.GraphedContrastiveStep(m, xh, ih, lh, prefetch=True, lagged_loss=True)
    step.prime()                                        # batch 0 on the device
    got = []
    for k in (1, 2, 2):                                 # replay j computes batch j (copied during replay j-1)
        # the pinned staging buffers are double-buffered in lagged mode: step.x_host / ids_host / lens_host name
        # the set the NEXT call copies, which no replay in flight reads -- no synchronize needed here
        step.x_host.copy_(t(inps[k]["f"])); step.ids_host.copy_(t(inps[k]["ids"])); step.lens_host.copy_(t(inps[k]["lens"]))
        got.append(step())
        step.x_host.fill_(float("nan"))                 # scribbling on the OTHER set cannot disturb the replay in flight
    got.append(step.flush())
    assert got[0] != got[0]                             # NaN: nothing had finished at the first call
    for a, b in zip(got[1:], ref):
        assert abs(a - b) <= 1e-6 * abs(b), (got, ref)


def _flat_model(cv, inp, E=512, fix_temperature=True):
    import argparse
    args = argparse.Namespace(embedding_type="flat", embedding_dim=E, normalize_features=True,
                              fix_temperature=fix_temperature, temperature=0.07, text_encoder="embedding")
    vocab = {str(i): i for i in range(2350)}
    m = cv.MultiModalModel(cv.VisionEncoder(args, trunk="pooled"), cv.TextEncoder(vocab, 2048, args), args)
    with torch.no_grad():
        m.image_embed.model.fc.weight.copy_(t(inp["W"])); m.image_embed.model.fc.bias.copy_(t(inp["b"]))
        m.text_embed.embedding.weight.copy_(t(inp["table"]))
    m.to(DEV).train()
    m.materialize_logits = m.materialize_text_outputs = m.materialize_features = False
    return m


def _head_params(mm):
    return [mm.image_embed.model.fc.weight, mm.image_embed.model.fc.bias, mm.text_embed.embedding.weight]


def test_graphed_train_step_with_optimizer_in_graph(cv):
    """GraphedContrastiveStep(optimizer=FusedAdamW): forward + backward + AdamW + bf16 weight shadow refresh as ONE
    graph == the eager loop (calculate_contrastive_loss, backward, torch.optim.AdamW) on the same batches."""
    inps = [case_inputs(700 + k, 256, 512, "flat") for k in range(4)]
    ref_m = _flat_model(cv, inps[0]); got_m = _flat_model(cv, inps[0])
    ref_opt = torch.optim.AdamW(_head_params(ref_m), lr=1e-3, weight_decay=0.1)
    ref_losses = []
    for i in inps:
        ref_opt.zero_grad(set_to_none=True)
        loss = ref_m.calculate_contrastive_loss(t(i["f"], DEV), t(i["ids"], DEV), t(i["lens"], DEV))[0]
        loss.backward(); ref_opt.step(); ref_losses.append(loss.item())
    opt = cv.FusedAdamW(_head_params(got_m), lr=1e-3, weight_decay=0.1)
    xh = t(inps[0]["f"]).pin_memory(); ih = t(inps[0]["ids"]).pin_memory(); lh = t(inps[0]["lens"]).pin_memory()
    step = cv.GraphedContrastiveStep(got_m, xh, ih, lh, optimizer=opt)
    got_losses = []
    for i in inps:
        step.x_host.copy_(t(i["f"])); step.ids_host.copy_(t(i["ids"])); step.lens_host.copy_(t(i["lens"]))
        got_losses.append(step())
    assert opt.device_step_count() == len(inps)
    for a, b in zip(got_losses, ref_losses):            # later losses depend on the earlier updates
        assert abs(a - b) <= 2e-3 * abs(b), (got_losses, ref_losses)
    for pa, pb in zip(_head_params(got_m), _head_params(ref_m)):
        assert float((pa - pb).abs().max()) <= 2e-3 * max(1.0, float(pb.abs().max()))
    # the shadow the head GEMM reads is the bf16 image of the updated master weight
    w = got_m.image_embed.model.fc.weight
    assert torch.equal(cv.ops.weight_shadow(w), w.detach().to(torch.bfloat16))


def test_graphed_step_trainable_temperature(cv):
    """fix_temperature=False (the reference default, multimodal.py:711-715): s lives on the device, the graph
    reads it every replay and writes ds into s.grad -- no host sync, no rebuild when s changes."""
    inp = case_inputs(810, 256, 512, "flat")
    m = _flat_model(cv, inp, fix_temperature=False)
    xh = t(inp["f"]).pin_memory(); ih = t(inp["ids"]).pin_memory(); lh = t(inp["lens"]).pin_memory()
    step = cv.GraphedContrastiveStep(m, xh, ih, lh)
    for sval in (S_DEFAULT, 2.0):
        with torch.no_grad():
            m.logit_neg_log_temperature.fill_(sval)
        loss = step()
        ref = oracle_flat_step(inp, s=sval)
        assert abs(loss - ref["loss"].item()) <= 1e-3 * abs(ref["loss"].item())
        ds = m.logit_neg_log_temperature.grad.item()
        assert abs(ds - ref["ds"].item()) <= 2e-2 * abs(ref["ds"].item()) + 1e-3
