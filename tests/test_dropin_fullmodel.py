"""BASELINE config 1 and the true drop-in case.

CPU (build container only, skipped where /root/reference is absent): the drop-in MultiModalModel is constructed with
the REFERENCE's own VisionEncoder / TextEncoder objects -- exactly what multimodal_lit.py:61-65 does after the import
swap -- and must expose the reference's state_dict keys, attributes and trunk boundary.

GPU: the full model (seeded ResNeXt-50 trunk as the stock torch module + this library's head) through
calculate_contrastive_loss / forward / encode_image against goldens produced by the unmodified reference
(oracle/make_golden_fullmodel.py), flat and spatial (mean, max)."""
import argparse
import os

import numpy as np
import pytest
import torch

from _util import ROOT, assert_logits_close, golden, t
from oracle import ref_import as R
from oracle.make_golden import case_inputs
from oracle.make_golden_fullmodel import load_trunk, seeded_images, seeded_trunk

needs_ref = pytest.mark.skipif(not R.reference_available(), reason="reference tree not present (GPU box)")


def _load_head(model, inp, embedding_type):
    with torch.no_grad():
        if embedding_type == "flat":
            model.image_embed.model.fc.weight.copy_(torch.from_numpy(inp["W"]))
            model.image_embed.model.fc.bias.copy_(torch.from_numpy(inp["b"]))
        else:
            conv = list(model.image_embed.model.children())[-1]
            conv.weight.copy_(torch.from_numpy(inp["W"])[:, :, None, None])
            conv.bias.copy_(torch.from_numpy(inp["b"]))
        model.text_embed.embedding.weight.copy_(torch.from_numpy(inp["table"]))


@needs_ref
@pytest.mark.parametrize("embedding_type,sim,fix", [("flat", "mean", False), ("flat", "mean", True), ("spatial", "max", False)])
def test_dropin_with_reference_encoders_cpu(embedding_type, sim, fix):
    import contextlib, io, warnings
    import multimodal_baby_b200 as cv
    ref_mm, _ = R.load_reference()
    args = R.make_args(embedding_type, sim, 512, fix)
    vocab = R.reference_vocab()
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ve = ref_mm.VisionEncoder(args)
        te = ref_mm.TextEncoder(vocab, ve.last_cnn_out_dim, args)
        ref_model = ref_mm.MultiModalModel(ve, te, args)
    mine = cv.MultiModalModel(ve, te, args)               # the reference's OWN encoder objects
    assert list(mine.state_dict().keys()) == list(ref_model.state_dict().keys())
    assert len(mine.state_dict()) > 300                    # the whole ResNeXt trunk is in there
    for name in ("normalize_features", "sim", "embedding_type", "fix_temperature", "image_embed", "text_embed",
                 "logit_neg_log_temperature"):
        assert hasattr(mine, name)
    assert isinstance(mine.logit_neg_log_temperature, torch.nn.Parameter) == (not fix)
    assert float(mine.logit_neg_log_temperature) == pytest.approx(float(ref_model.logit_neg_log_temperature))
    # the head parameters are found where the reference keeps them
    w, b = mine._head()
    assert tuple(w.shape) == (512, 2048) and tuple(b.shape) == (512,)
    # trunk boundary on CPU (the trunk is a stock torch module): pooled / layer4 activations equal the reference's
    ve.eval()
    x = seeded_images(5, n=2)
    with torch.no_grad():
        boundary, fmap = cv.split_trunk_forward(mine.image_embed, x, run_head=False)
        ref_feat, ref_fmap = ve(x)
    assert torch.equal(fmap, ref_fmap)
    if embedding_type == "flat":
        assert tuple(boundary.shape) == (2, 2048)
        assert torch.allclose(boundary, torch.flatten(torch.nn.functional.adaptive_avg_pool2d(ref_fmap, 1), 1), atol=1e-6)
        # head applied by torch on CPU reproduces the reference's features: the split is exact
        assert torch.allclose(torch.nn.functional.linear(boundary, w, b), ref_feat, atol=1e-5)
    else:
        assert torch.equal(boundary, fmap)
    # the CUDA-only contract: CPU tensors raise instead of silently taking another path
    with pytest.raises(RuntimeError):
        mine.calculate_contrastive_loss(x, torch.zeros(2, 25, dtype=torch.int64), torch.full((2,), 3, dtype=torch.int64))


def _mirror_model(name, dev):
    import multimodal_baby_b200 as cv
    g = golden(name)
    et, sim = str(g["embedding_type"]), str(g["sim"])
    args = argparse.Namespace(embedding_type=et, embedding_dim=int(g["E"]), normalize_features=True, fix_temperature=True,
                              temperature=0.07, text_encoder="embedding", cnn_model="resnext50_32x4d", finetune_cnn=False,
                              sim=sim)
    vocab = {str(i): i for i in range(2350)}
    m = cv.MultiModalModel(cv.VisionEncoder(args, trunk="resnext"), cv.TextEncoder(vocab, 2048, args), args)
    load_trunk(m.image_embed, seeded_trunk(), et)
    inp = case_inputs(int(g["seed"]), int(g["B"]), int(g["E"]), et)
    _load_head(m, inp, et)
    return g, inp, m.to(dev).eval()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["fullmodel_flat_b8", "fullmodel_spatial_mean_b8", "fullmodel_spatial_max_b8"])
def test_full_model_matches_reference_golden(name):
    torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
    dev = "cuda"
    g, inp, m = _mirror_model(name, dev)
    x = seeded_images(int(g["seed"]) + 1).to(dev)
    ids, lens = t(inp["ids"], dev), t(inp["lens"], dev)
    with torch.no_grad():
        out = m.calculate_contrastive_loss(x, ids, lens)
        lpi, lpt = m(x, ids, lens)
        feats, fmap = m.encode_image(x)
    # the trunk itself (stock torch on the GPU vs the reference on CPU)
    assert tuple(fmap.shape) == tuple(g["feature_map_shape"])
    assert abs(float(fmap.mean()) - float(g["feature_map_mean"])) <= 1e-3 * abs(float(g["feature_map_mean"]))
    assert abs(out[0].item() - float(g["loss"])) <= 1e-3 * abs(float(g["loss"]))
    assert abs(out[3].item() - float(g["image_entropy"])) <= 2e-3 and abs(out[4].item() - float(g["text_entropy"])) <= 2e-3
    assert abs(out[1].item() - float(g["image_accuracy"])) <= 0.125 + 1e-6       # 8 rows, near-tied random-init logits
    assert_logits_close(out[5].cpu().numpy(), g["logits_per_image"])
    assert_logits_close(out[6].cpu().numpy(), g["logits_per_text"])
    assert_logits_close(lpi.cpu().numpy(), g["logits_per_image"])
    assert_logits_close(lpt.cpu().numpy(), g["logits_per_text"])
    if "image_features" in g:
        assert float((feats.cpu() - torch.from_numpy(g["image_features"])).abs().max()) <= 6e-3
        assert float((fmap.mean((2, 3))[:, :16].cpu() - torch.from_numpy(g["pooled_head"])).abs().max()) <= \
            1e-3 * float(np.abs(g["pooled_head"]).max())
    else:
        assert float((feats[:2, :, :2, :2].cpu() - torch.from_numpy(g["image_features_slice"])).abs().max()) <= 6e-3


@pytest.mark.gpu
def test_full_model_training_step_updates_only_the_head():
    """frozen trunk (finetune_cnn=False, multimodal.py:175-177): calculate_contrastive_loss + backward leaves gradients on
    fc / embedding only, and they equal the head-only step on the trunk-boundary features."""
    import multimodal_baby_b200 as cv
    dev = "cuda"
    g, inp, m = _mirror_model("fullmodel_flat_b8", dev)
    m.train(); m.image_embed.model.eval()                  # BatchNorm on its running statistics, as in the golden
    x = seeded_images(int(g["seed"]) + 1).to(dev)
    ids, lens = t(inp["ids"], dev), t(inp["lens"], dev)
    out = m.calculate_contrastive_loss(x, ids, lens)
    out[0].backward()
    with_grad = sorted(n for n, p in m.named_parameters() if p.grad is not None)
    assert with_grad == ["image_embed.model.fc.bias", "image_embed.model.fc.weight", "text_embed.embedding.weight"]
    with torch.no_grad():
        pooled, _ = cv.split_trunk_forward(m.image_embed, x, run_head=False)
    W = m.image_embed.model.fc.weight.detach().clone().requires_grad_(True)
    b = m.image_embed.model.fc.bias.detach().clone().requires_grad_(True)
    tab = m.text_embed.embedding.weight.detach().clone().requires_grad_(True)
    loss2 = cv.ops.flat_contrastive_loss(pooled, ids, lens, W, b, tab, m.logit_neg_log_temperature, True)[0]
    loss2.backward()
    assert abs(loss2.item() - out[0].item()) <= 1e-6 * abs(out[0].item())
    assert torch.equal(W.grad, m.image_embed.model.fc.weight.grad)
    assert torch.equal(tab.grad, m.text_embed.embedding.weight.grad)
