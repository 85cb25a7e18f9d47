"""Import the UNMODIFIED reference (wkvong/multimodal-baby) from /root/reference.

TEST INFRASTRUCTURE ONLY.  Works only in the build container (the GPU box has no
/root/reference); used by oracle/make_golden.py to generate tests/golden/*.npz and by
the `-m "not gpu"` tests that pin oracle/cvcl_oracle.py against the live reference.

Five third-party modules that the reference imports but this image lacks are stubbed in
sys.modules before the import (pytorch_lightning, clip, spacy, matplotlib, pycocoevalcap);
nothing under /root/reference is modified or copied.
"""
import argparse
import json
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("CVCL_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "multimodal", "multimodal.py"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _LightningModule(torch.nn.Module):          # multimodal_lit.py:35 subclasses this
    def save_hyperparameters(self, *a, **k):      # multimodal_lit.py:74
        pass

    def log(self, *a, **k):                       # multimodal_lit.py:247-253
        pass


class _LightningDataModule:                       # multimodal_data_module.py:217
    def __init__(self, *a, **k):
        pass


_loaded = None


def load_reference():
    """Returns (multimodal.multimodal, multimodal.multimodal_lit) of the reference."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if "pytorch_lightning" not in sys.modules:
        _stub("pytorch_lightning", LightningModule=_LightningModule,
              LightningDataModule=_LightningDataModule, seed_everything=torch.manual_seed)
    if "clip" not in sys.modules:
        _stub("clip")
    if "spacy" not in sys.modules:
        _stub("spacy", load=lambda name: (
            lambda text: [types.SimpleNamespace(text=t) for t in text.split()]))
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        _stub("matplotlib")
        _stub("matplotlib.pyplot")
    for sub in ("", ".bleu", ".bleu.bleu", ".meteor", ".meteor.meteor", ".rouge",
                ".rouge.rouge", ".cider", ".cider.cider", ".spice", ".spice.spice"):
        if "pycocoevalcap" + sub not in sys.modules:
            _stub("pycocoevalcap" + sub, Bleu=None, Meteor=None, Rouge=None, Cider=None,
                  Spice=None)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from multimodal import multimodal as ref_mm          # noqa: E402
    from multimodal import multimodal_lit as ref_lit     # noqa: E402
    _loaded = (ref_mm, ref_lit)
    return _loaded


def reference_vocab():
    with open(os.path.join(REFERENCE_ROOT, "multimodal", "vocab.json")) as f:
        return json.load(f)


def make_args(embedding_type="flat", sim="mean", embedding_dim=512, fix_temperature=False,
              temperature=0.07, normalize_features=True, **extra):
    ns = argparse.Namespace(
        embedding_type=embedding_type, embedding_dim=embedding_dim, pretrained_cnn=False,
        cnn_model="resnext50_32x4d", finetune_cnn=False, text_encoder="embedding",
        normalize_features=normalize_features, fix_temperature=fix_temperature,
        temperature=temperature, dropout_i=0.5, dropout_o=0.0, crange=1, sim=sim)
    for k, v in extra.items():
        setattr(ns, k, v)
    return ns


class PooledTrunk(torch.nn.Module):
    """Head-only shim: keeps the reference's own fc, fed pooled [B,2048] features, and
    satisfies the layer4 forward hook at multimodal.py:96-102."""

    def __init__(self, fc):
        super().__init__()
        self.layer4 = torch.nn.Identity()
        self.fc = fc

    def forward(self, x):
        return self.fc(self.layer4(x))


def build_reference_model(embedding_type="flat", sim="mean", embedding_dim=512,
                          fix_temperature=False, vocab=None, head_only=True, lit=False,
                          **extra):
    """Construct the reference MultiModalModel (random init, no network).  With
    head_only=True the ResNeXt trunk is replaced after construction by the shim so the
    reference forward/loss code runs unchanged on trunk-boundary features."""
    import contextlib
    import io
    ref_mm, ref_lit = load_reference()
    args = make_args(embedding_type, sim, embedding_dim, fix_temperature, **extra)
    vocab = vocab if vocab is not None else reference_vocab()
    with contextlib.redirect_stdout(io.StringIO()):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ve = ref_mm.VisionEncoder(args)
        te = ref_mm.TextEncoder(vocab, ve.last_cnn_out_dim, args)
    if head_only:
        if embedding_type == "flat":
            ve.model = PooledTrunk(ve.model.fc)
        else:
            ve.model = torch.nn.Sequential(torch.nn.Identity(), ve.model[-1])
    if lit:
        m = ref_lit.MultiModalLitModel(ve, te, args)
    else:
        m = ref_mm.MultiModalModel(ve, te, args)
    return m
