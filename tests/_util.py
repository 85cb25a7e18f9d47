"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import cvcl_oracle as O                      # noqa: E402
from oracle.make_golden import case_inputs               # noqa: E402,F401

GOLD = os.path.join(ROOT, "tests", "golden")
S_DEFAULT = float(-np.log(0.07))                         # multimodal.py:711


def golden(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False))


def t(x, device="cpu", dtype=None):
    y = torch.from_numpy(np.ascontiguousarray(x))
    if dtype is not None:
        y = y.to(dtype)
    return y.to(device)


def cosine(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    return float(a @ b / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-300))


def rel_fro(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def assert_grad_close(got, ref, name, cos_min=0.999, rel_max=2e-2):
    """gradient gate of SURVEY section 8d: cosine >= 0.999 and rel-Frobenius <= 2e-2."""
    c, r = cosine(got, ref), rel_fro(got, ref)
    assert c >= cos_min and r <= rel_max, f"{name}: cosine={c:.6f} rel_fro={r:.3e}"


def assert_logits_close(got, ref, tol=1e-2):
    """logit gate: max |delta| <= tol * max |logit| (bf16 operands, fp32 accumulate)."""
    ref = np.asarray(ref, np.float64); got = np.asarray(got, np.float64)
    err = np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30)
    assert err <= tol, f"logits: max|delta|/max|logit| = {err:.3e} > {tol}"


def oracle_flat_step(inp, s=S_DEFAULT, dtype=torch.float32, normalize=True):
    return O.contrastive_step(t(inp["f"]), t(inp["ids"]), t(inp["lens"]), t(inp["W"]), t(inp["b"]),
                              t(inp["table"]), s, "flat", normalize=normalize, dtype=dtype)
