"""CPU suite (-m "not gpu"): the oracle restatement against the golden vectors produced by the
UNMODIFIED reference (oracle/make_golden.py), plus internal consistency of the oracle
(closed-form fp64 backward vs autograd, sharded loss vs global loss, batched eval vs trial loop)."""
import json
import os

import numpy as np
import pytest
import torch

from _util import O, S_DEFAULT, case_inputs, golden, t, rel_fro, GOLD

TRAIN_CASES = [
    ("flat_e64_b8", "flat", "mean"), ("flat_e512_b32", "flat", "mean"), ("flat_e512_b160", "flat", "mean"),
    ("spatial_mean_e64_b6", "spatial", "mean"), ("spatial_max_e64_b6", "spatial", "max"),
    ("spatial_max_e512_b12", "spatial", "max"), ("spatial_mean_e512_b12", "spatial", "mean"),
]


@pytest.mark.parametrize("name,etype,sim", TRAIN_CASES)
def test_oracle_matches_reference_golden(name, etype, sim):
    g = golden(name)
    inp = case_inputs(int(g["seed"]), int(g["B"]), int(g["E"]), etype)
    out = O.contrastive_step(t(inp["f"]), t(inp["ids"]), t(inp["lens"]), t(inp["W"]), t(inp["b"]),
                             t(inp["table"]), S_DEFAULT, etype, sim)
    assert abs(out["loss"].item() - float(g["loss"])) <= 2e-6 * abs(float(g["loss"]))
    np.testing.assert_allclose(out["logits_per_image"].numpy(), g["logits_per_image"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(out["logits_per_text"].numpy(), g["logits_per_text"], rtol=0, atol=2e-5)
    assert np.array_equal(out["image_pred"].numpy(), g["image_pred"])
    assert np.array_equal(out["text_pred"].numpy(), g["text_pred"])
    for k in ("image_accuracy", "text_accuracy", "image_entropy", "text_entropy"):
        assert abs(float(out[k]) - float(g[k])) <= 1e-5
    assert abs(out["ds"].item() - float(g["ds"])) <= 1e-4 * max(abs(float(g["ds"])), 1e-3)
    np.testing.assert_allclose(out["db"].numpy(), g["db"], rtol=1e-4, atol=1e-7)
    dW = out["dW"].reshape(int(g["E"]), -1).numpy()
    assert rel_fro(dW[:8, :64], g["dW_slice"]) <= 1e-4
    assert abs(np.linalg.norm(dW.astype(np.float64)) - float(g["dW_norm"])) <= 1e-4 * float(g["dW_norm"])
    dtab = out["dtable"].numpy()
    assert rel_fro(dtab[:8], g["dtable_rows"]) <= 1e-4
    assert abs(np.linalg.norm(dtab.astype(np.float64)) - float(g["dtable_norm"])) <= 1e-4 * float(g["dtable_norm"])
    assert not dtab[0].any()                       # padding row never receives a gradient
    if "dW" in g:
        assert rel_fro(dW, g["dW"]) <= 1e-4
        assert rel_fro(dtab[g["dtable_nz_ids"]], g["dtable_nz"]) <= 1e-4
    if etype == "flat":
        np.testing.assert_allclose(out["image_features"].numpy(), g["image_features"], atol=1e-6)
        np.testing.assert_allclose(out["text_features"].numpy(), g["text_features"], atol=1e-6)


@pytest.mark.parametrize("name", ["forward_4x3_e512", "forward_4x1_e512"])
def test_oracle_forward_ni_ne_nt(name):
    g = golden(name)
    inp = case_inputs(int(g["seed"]), int(g["Ni"]), int(g["E"]), "flat", Bt=int(g["Nt"]))
    lpi, lpt, _, _ = O.forward(t(inp["f"]), t(inp["ids"]), t(inp["lens"]), t(inp["W"]), t(inp["b"]),
                               t(inp["table"]), S_DEFAULT)
    np.testing.assert_allclose(lpi.numpy(), g["logits_per_image"], atol=2e-5)
    np.testing.assert_allclose(lpt.numpy(), g["logits_per_text"], atol=2e-5)
    assert lpi.shape == (int(g["Ni"]), int(g["Nt"])) and lpt.shape == (int(g["Nt"]), int(g["Ni"]))


def eval_case_inputs(g):
    rng = np.random.RandomState(int(g["seed"]))
    W, b, table = O.synth_weights(rng, int(g["E"]))
    f = O.synth_trunk_features(rng, (int(g["n_trials"]), int(g["n_way"]), 2048))
    return W, b, table, f


def test_oracle_eval_matches_reference_trial_loop():
    g = golden("eval_4way_e512")
    W, b, table, f = eval_case_inputs(g)
    preds = O.eval_trial_loop(t(f), t(g["ids"]), t(g["lens"]), t(W), t(b), t(table), S_DEFAULT)
    assert np.array_equal(preds.numpy().astype(np.int32), g["pred"])
    # batched form == per-trial loop
    img = O.head_flat(t(f).reshape(-1, 2048), t(W), t(b)).reshape(len(f), 4, -1)
    txt, _ = O.text_encoder_flat(t(g["ids"]), t(g["lens"]), t(table))
    pred2, logits = O.eval_nway(img, txt, S_DEFAULT)
    assert np.array_equal(pred2.numpy().astype(np.int32), g["pred"])
    np.testing.assert_allclose(logits.numpy(), g["logits"], atol=2e-5)


def test_closed_form_backward_matches_autograd_fp64():
    inp = case_inputs(7, 24, 64, "flat")
    out = O.contrastive_step(t(inp["f"]), t(inp["ids"]), t(inp["lens"]), t(inp["W"]), t(inp["b"]),
                             t(inp["table"]), S_DEFAULT, "flat", dtype=torch.float64)
    cf = O.closed_form_flat_backward(inp["f"], inp["ids"], inp["lens"], inp["W"], inp["b"], inp["table"],
                                     S_DEFAULT)
    assert abs(cf["loss"] - out["loss"].item()) < 1e-12
    assert abs(cf["ds"] - out["ds"].item()) < 1e-12
    assert rel_fro(cf["dW"], out["dW"].numpy()) < 1e-12
    assert rel_fro(cf["db"], out["db"].numpy()) < 1e-12
    assert rel_fro(cf["dtable"], out["dtable"].numpy()) < 1e-12


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_loss_equals_global(world):
    rng = np.random.RandomState(3)
    B, E = 32, 64
    img = torch.nn.functional.normalize(t(rng.standard_normal((B, E))), dim=1)
    txt = torch.nn.functional.normalize(t(rng.standard_normal((B, E))), dim=1)
    lpi, lpt = O.logits_from_match(O.similarity_flat(img, txt), S_DEFAULT)
    ref = O.infonce(lpi, lpt).loss
    parts = O.sharded_contrastive_loss(img, txt, S_DEFAULT, world)
    assert abs(sum(p.item() for p in parts) - ref.item()) < 1e-12


def test_entropy_identity():
    """get_entropy == lse - sum p*logit (what the fused kernel computes)."""
    x = torch.randn(16, 16, dtype=torch.float64) * 5
    lse = torch.logsumexp(x, -1)
    alt = lse - (torch.softmax(x, -1) * x).sum(-1)
    assert torch.allclose(O.get_entropy(x), alt, atol=1e-12)


def test_spatial_mean_factorises():
    """multimodal.py:765-770 == pooled-image . pooled-text / (HW * len) (used by the kernels)."""
    rng = np.random.RandomState(5)
    img = t(rng.standard_normal((5, 16, 7, 7)))
    txt = t(rng.standard_normal((4, 6, 16)))
    lens = torch.tensor([6, 3, 1, 4])
    ref = O.similarity_spatial_mean(img, txt, lens)
    alt = img.sum((2, 3)) @ (txt.sum(1) / (49 * lens[:, None])).T
    assert torch.allclose(ref, alt, atol=1e-12)


def test_tokenize_golden_layout():
    with open(os.path.join(GOLD, "tokenize.json")) as fh:
        g = json.load(fh)
    ids = np.array(g["ids"]); lens = np.array(g["lens"])
    assert ids.shape == (len(g["texts"]), 25)
    for row, n in zip(ids, lens):
        assert row[0] == O.SOS_TOKEN_ID and row[n - 1] == O.EOS_TOKEN_ID and not row[n:].any()
    assert lens[3] == 25                            # 40 words truncated to 23 + sos + eos
    assert ids[2][4] == O.UNK_TOKEN_ID              # out-of-vocabulary word


def test_synth_tokens_contract():
    ids, lens = O.synth_tokens(np.random.RandomState(0), 64)
    assert ids.dtype == np.int64 and lens.dtype == np.int64
    assert lens.min() >= 3 and lens.max() <= 25
    for row, n in zip(ids, lens):
        assert row[0] == 2 and row[n - 1] == 3 and (row[1:n - 1] >= 4).all() and not row[n:].any()


# ---------------------------------------------------------------- Grad-CAM
def test_oracle_gradcam_matches_reference_golden():
    """oracle restatement (autograd and closed form) of attention_maps.gradCAM vs the maps the unmodified
    reference produced (oracle/make_golden.py: run_gradcam_case), both normalisation modes + the resize."""
    from oracle.make_golden import gradcam_inputs
    g = golden("gradcam_e512_n3")
    inp = gradcam_inputs(int(g["seed"]), int(g["N"]), int(g["E"]))
    for norm, key in ((True, "norm"), (False, "raw")):
        cam, big = O.gradcam_flat(t(inp["act"]), t(inp["W"]), t(inp["b"]), t(inp["target"]), norm, (224, 224))
        scale = float(np.abs(g["cam_" + key]).max())
        assert scale > 0 and (g["cam_" + key] > 0).any() and (g["cam_" + key] == 0).any()     # the clamp is exercised
        assert np.abs(cam.numpy() - g["cam_" + key]).max() <= 1e-5 * scale
        assert np.abs(big.numpy()[:, :, ::3, ::3] - g["resized_" + key]).max() <= 1e-5 * scale
        cf = O.gradcam_flat_closed_form(t(inp["act"]), t(inp["W"]), t(inp["b"]), t(inp["target"]), norm)
        assert float((cf - cam).abs().max()) <= 1e-5 * scale
    # fp64: autograd and closed form agree to rounding
    a64 = t(inp["act"]).double(); W64 = t(inp["W"]).double(); b64 = t(inp["b"]).double(); t64 = t(inp["target"]).double()
    cam64, _ = O.gradcam_flat(a64, W64, b64, t64, True)
    cf64 = O.gradcam_flat_closed_form(a64, W64, b64, t64, True)
    assert float((cam64 - cf64).abs().max()) <= 1e-12
