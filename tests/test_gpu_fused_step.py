"""The one-kernel flat train step (csrc/fused_step.cuh, cvcl_flat_step_fused): every phase against a torch
restatement on the same bf16-rounded operands, the whole step against the oracle (through
ops.flat_contrastive_loss, which routes the covered shapes to it: test_gpu_parity.py::test_flat_step_vs_oracle),
against the multi-kernel step, replay stability and the device-side temperature."""
import os
import sys

import numpy as np
import pytest
import torch

from _util import O, S_DEFAULT, assert_grad_close, case_inputs, oracle_flat_step, t

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def cv():
    import multimodal_baby_b200 as m
    m._cabi.load()
    return m


def dev_inputs(inp):
    return {k: t(v, DEV) for k, v in inp.items()}


@pytest.mark.parametrize("B,E,K", [(512, 512, 2048), (300, 512, 2048), (128, 256, 512), (1024, 512, 2048),
                                   (640, 384, 1024)])
def test_fused_phases_vs_restatement(cv, B, E, K):
    import fused_check
    rep = fused_check.run(B, E, K, verbose=False)
    # operands are identical bf16 values on both sides: only the accumulation order differs
    assert rep["P0 head partial sum vs x.W^T"] <= 1e-5, rep
    assert rep["P0 txt16"] <= 2e-3 and rep["P0 txt_f32 max abs"] <= 2e-6, rep     # bf16 rounding flips at most
    assert rep["P1 img16"] <= 2e-3 and rep["P1 invn_i"] <= 1e-5 and rep["P0 invn_t"] <= 1e-5, rep
    assert rep["P2 lse0 max abs"] <= 2e-3 and rep["P2 lse1 max abs"] <= 2e-3, rep
    assert rep["P3 dI (no diag term)"] <= 1e-2 and rep["P3 dT (no diag term)"] <= 1e-2, rep
    assert rep["P4 du16"] <= 2e-2 and rep["P4 dm16"] <= 2e-2 and rep["P1 token counts exact"], rep
    assert abs(rep["loss"][0] - rep["loss"][1]) <= 1e-3 * abs(rep["loss"][1]), rep
    assert abs(rep["ent"][0] - rep["ent"][1]) <= 2e-3 and abs(rep["ent"][2] - rep["ent"][3]) <= 2e-3, rep
    assert abs(rep["acc"][0] - rep["acc"][1]) <= 0.02 and abs(rep["acc"][2] - rep["acc"][3]) <= 0.02, rep
    assert rep["dW"] <= 2e-2 and rep["db"] <= 2e-2 and rep["dtable"] <= 2e-2 and rep["dtable row0 zero"] == 0.0, rep
    assert abs(rep["ds"][0] - rep["ds"][1]) <= 2e-2 * abs(rep["ds"][1]) + 1e-3, rep
    assert not rep["nan in grads"], rep
    # no atomics on data anywhere: replays are bit-identical, the embedding gradient included
    assert rep["replay: out5 bit-identical"] and rep["replay: all gradients bit-identical"], rep
    assert rep["control block after run"][:3] == [0, 0, 0], rep


def _step(cv, d, s, fused, normalize=True, want=False):
    cv.ops.FUSED_STEP = fused
    try:
        W = d["W"].clone().requires_grad_(True); b = d["b"].clone().requires_grad_(True)
        table = d["table"].clone().requires_grad_(True)
        out = cv.ops.flat_contrastive_loss(d["f"], d["ids"], d["lens"], W, b, table, s, normalize, want_features=want)
        out[0].backward()
        torch.cuda.synchronize()
    finally:
        cv.ops.FUSED_STEP = True
    return out, W.grad, b.grad, table.grad


@pytest.mark.parametrize("B", [512, 200])
def test_fused_equals_multikernel_step(cv, B):
    inp = case_inputs(4000 + B, B, 512, "flat")
    d = dev_inputs(inp)
    sa = torch.tensor(S_DEFAULT, device=DEV, requires_grad=True)
    sb = torch.tensor(S_DEFAULT, device=DEV, requires_grad=True)
    oa, dWa, dba, dta = _step(cv, d, sa, True, want=True)
    ob, dWb, dbb, dtb = _step(cv, d, sb, False, want=True)
    assert abs(oa[0].item() - ob[0].item()) <= 1e-5 * abs(ob[0].item())
    for i in (1, 2, 3, 4):
        assert abs(oa[i].item() - ob[i].item()) <= 1e-3
    assert float((oa[5] - ob[5]).abs().max()) <= 1e-6 and float((oa[6] - ob[6]).abs().max()) <= 1e-6
    assert_grad_close(dWa.cpu().numpy(), dWb.cpu().numpy(), "dW", cos_min=0.9999, rel_max=1e-2)
    assert_grad_close(dba.cpu().numpy(), dbb.cpu().numpy(), "db", cos_min=0.9999, rel_max=1e-2)
    assert_grad_close(dta.cpu().numpy(), dtb.cpu().numpy(), "dtable", cos_min=0.9999, rel_max=1e-2)
    assert abs(sa.grad.item() - sb.grad.item()) <= 1e-2 * abs(sb.grad.item()) + 1e-4


def test_fused_device_side_temperature(cv):
    """trainable temperature (multimodal.py:711-715): s is read on the device, ds matches the oracle, and a
    changed s is picked up without rebuilding anything."""
    inp = case_inputs(4321, 256, 512, "flat")
    d = dev_inputs(inp)
    for sval in (S_DEFAULT, float(np.log(1.0 / 0.1))):
        ref = oracle_flat_step(inp, s=sval)
        s = torch.tensor(sval, device=DEV, requires_grad=True)
        out, dW, db, dt = _step(cv, d, s, True)
        assert abs(out[0].item() - ref["loss"].item()) <= 1e-3 * abs(ref["loss"].item())
        assert abs(s.grad.item() - ref["ds"].item()) <= 2e-2 * abs(ref["ds"].item()) + 1e-3
        assert_grad_close(dW.cpu().numpy(), ref["dW"].numpy(), "dW")


def test_fused_forward_only_and_unnormalized(cv):
    inp = case_inputs(99, 384, 256, "flat")
    inp["table"] *= 0.05
    d = dev_inputs(inp)
    ref = oracle_flat_step(inp, s=0.0, normalize=False)
    with torch.no_grad():
        out = cv.ops.flat_contrastive_loss(d["f"], d["ids"], d["lens"], d["W"], d["b"], d["table"], 0.0, False)
    assert abs(out[0].item() - ref["loss"].item()) <= 2e-3 * abs(ref["loss"].item())
    W = d["W"].clone().requires_grad_(True); table = d["table"].clone().requires_grad_(True)
    b = d["b"].clone().requires_grad_(True)
    loss = cv.ops.flat_contrastive_loss(d["f"], d["ids"], d["lens"], W, b, table, 0.0, False)[0]
    loss.backward()
    assert_grad_close(W.grad.cpu().numpy(), ref["dW"].numpy(), "dW", rel_max=3e-2)
    assert_grad_close(table.grad.cpu().numpy(), ref["dtable"].numpy(), "dtable", rel_max=3e-2)


def test_weight_shadow_tracks_parameter_updates(cv):
    """the bf16 shadow of W is recast when (and only when) the parameter changed."""
    inp = case_inputs(5, 128, 512, "flat")
    d = dev_inputs(inp)
    W = torch.nn.Parameter(d["W"].clone()); b = torch.nn.Parameter(d["b"].clone())
    table = torch.nn.Parameter(d["table"].clone())
    n0 = cv._cabi.load().cvcl_launch_count()
    l1 = cv.ops.flat_contrastive_loss(d["f"], d["ids"], d["lens"], W, b, table, S_DEFAULT, True)[0].item()
    n1 = cv._cabi.load().cvcl_launch_count()
    l2 = cv.ops.flat_contrastive_loss(d["f"], d["ids"], d["lens"], W, b, table, S_DEFAULT, True)[0].item()
    n2 = cv._cabi.load().cvcl_launch_count()
    assert l1 == l2
    assert n2 - n1 == (n1 - n0) - 1            # second call: no weight cast
    with torch.no_grad():
        W.mul_(0.5)
    l3 = cv.ops.flat_contrastive_loss(d["f"], d["ids"], d["lens"], W, b, table, S_DEFAULT, True)[0].item()
    ref = oracle_flat_step({**inp, "W": inp["W"] * 0.5})
    assert abs(l3 - ref["loss"].item()) <= 1e-3 * abs(ref["loss"].item())


def test_two_tiles_per_cta_layout(cv, monkeypatch):
    """T = 2 (two similarity tiles of one row block per CTA: the layout 8 ranks x 512 pairs need) forced on one GPU at
    1024 pairs: every phase against the restatement."""
    import fused_check
    monkeypatch.setenv("CVCL_B200_FUSED_FORCE_T", "2")
    lay = cv.ops.fused_layout(1024, 25, 512, 2048, 2350)
    assert lay["nPart"] * 2 == lay["nCB"]
    rep = fused_check.run(1024, 512, 2048, verbose=False)
    assert rep["P2 lse0 max abs"] <= 2e-3 and rep["P2 lse1 max abs"] <= 2e-3, rep
    assert rep["P3 dI (no diag term)"] <= 1e-2 and rep["P3 dT (no diag term)"] <= 1e-2, rep
    assert abs(rep["loss"][0] - rep["loss"][1]) <= 1e-3 * abs(rep["loss"][1]), rep
    assert rep["dW"] <= 2e-2 and rep["db"] <= 2e-2 and rep["dtable"] <= 2e-2, rep
    assert abs(rep["ds"][0] - rep["ds"][1]) <= 2e-2 * abs(rep["ds"][1]) + 1e-3, rep
    assert rep["replay: all gradients bit-identical"], rep


def test_eight_rank_tile_layout_on_one_gpu(cv, monkeypatch):
    """T = 2 AND one CTA per similarity tile (QS = 1: the layout of 8 ranks x 512 pairs, where the dQ GEMM of P3 needs
    eight operand slabs through a four-deep ring -- the producer must not wait for a slot before its share of the Gs
    pass), forced on one GPU at 1024 pairs."""
    import fused_check
    monkeypatch.setenv("CVCL_B200_FUSED_FORCE_T", "2")
    monkeypatch.setenv("CVCL_B200_FUSED_FORCE_QS", "1")
    lay = cv.ops.fused_layout(1024, 25, 512, 2048, 2350)
    assert lay["nPart"] * 2 == lay["nCB"] and lay["QS"] == 1
    rep = fused_check.run(1024, 512, 2048, verbose=False)
    assert rep["P3 dI (no diag term)"] <= 1e-2 and rep["P3 dT (no diag term)"] <= 1e-2, rep
    assert abs(rep["loss"][0] - rep["loss"][1]) <= 1e-3 * abs(rep["loss"][1]), rep
    assert rep["dW"] <= 2e-2 and rep["db"] <= 2e-2 and rep["dtable"] <= 2e-2, rep
    assert rep["replay: all gradients bit-identical"], rep


def test_sharded_entry_with_one_rank_equals_single_gpu_entry(cv):
    """cvcl_flat_step_fused_sharded with world = 1 (the peer tables name local buffers): the pointer plumbing of the
    sharded form -- gathered buffers, local slices, flag words, epoch -- gives bit-identical results."""
    import ctypes
    from multimodal_baby_b200 import _cabi
    B, E, K, V, L = 256, 512, 2048, 2350, 25
    inp = case_inputs(31337, B, E, "flat")
    d = dev_inputs(inp)
    x16 = d["f"].to(torch.bfloat16).contiguous(); w16 = d["W"].to(torch.bfloat16).contiguous()
    p = cv.ops._p
    st = torch.cuda.current_stream().cuda_stream
    lib = _cabi.load()
    f32 = dict(dtype=torch.float32, device=DEV)

    def outputs():
        flat = torch.zeros(4 + E + V * E + E * K, **f32)
        return torch.zeros(8, **f32), flat, cv.ops.split_flat_grads(flat, E, K, V)
    o1, flat1, (ds, db, dt, dW) = outputs()
    ws1 = torch.zeros(int(lib.cvcl_flat_fused_workspace_bytes(B, L, E, K, V)), dtype=torch.uint8, device=DEV)
    _cabi.call("cvcl_flat_step_fused", p(x16), p(w16), p(d["ids"]), p(d["lens"]), p(d["b"]), p(d["table"]), B, L, E, K, V,
               1, S_DEFAULT, None, 1, p(ws1), p(o1), None, None, p(dW), p(db), p(dt), p(ds), None, 0, st)
    o2, flat2, (ds2, db2, dt2, dW2) = outputs()
    ws2 = torch.zeros(int(lib.cvcl_flat_fused_sharded_workspace_bytes(B, L, E, K, V, 1)), dtype=torch.uint8, device=DEV)
    txt_all = torch.zeros(B, E, dtype=torch.bfloat16, device=DEV); img_all = torch.zeros_like(txt_all)
    lse_all = torch.zeros(int(lib.cvcl_flat_fused_sharded_part_bytes(B, 1)) // 4, **f32); flags = torch.zeros(32, dtype=torch.int32, device=DEV)
    epoch = torch.zeros(1, dtype=torch.int32, device=DEV)
    arr = ctypes.c_void_p * 1
    for _ in range(2):                                   # twice: the epoch advances, the control block is reusable
        _cabi.call("cvcl_flat_step_fused_sharded", p(x16), p(w16), p(d["ids"]), p(d["lens"]), p(d["b"]), p(d["table"]),
                   B, L, E, K, V, 1, S_DEFAULT, None, 1, p(ws2), p(o2), None, None, p(dW2), p(db2), p(dt2), p(ds2), None, 0,
                   1, 0, arr(txt_all.data_ptr()), arr(img_all.data_ptr()), arr(lse_all.data_ptr()), arr(flags.data_ptr()),
                   epoch.data_ptr(), None, None, 0, st)
    torch.cuda.synchronize()
    assert int(epoch.item()) == 2
    assert torch.equal(o1[:5], o2[:5]) and torch.equal(flat1, flat2)
    assert float(txt_all.float().norm(dim=1).min()) > 0.99               # the features landed in the gathered buffers


def test_out_of_range_token_id_is_reported(cv):
    """a token id outside [0, V) is treated as <pad> by the kernels and flagged in the device status word:
    ops.check_token_ids() raises IndexError (the reference's nn.Embedding fails with a device-side assert)."""
    inp = case_inputs(99, 128, 512, "flat")
    d = dev_inputs(inp)
    try:                                                       # start from a clean status word whatever ran before
        cv.ops.check_token_ids()
    except IndexError:
        pass
    ids = d["ids"].clone(); ids[3, 0] = 2350 + 7
    cv.ops.flat_contrastive_step(d["f"], ids, d["lens"], d["W"], d["b"], d["table"], S_DEFAULT, True, True, False)
    with pytest.raises(IndexError):
        cv.ops.check_token_ids()
    cv.ops.check_token_ids()                                   # the word was reset
    ids[3, 0] = -1
    cv.ops.text_features_flat(ids, d["lens"], d["table"])
    with pytest.raises(IndexError):
        cv.ops.check_token_ids()
