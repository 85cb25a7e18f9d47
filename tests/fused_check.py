"""Phase-by-phase check of the one-kernel flat step (csrc/fused_step.cuh) on a GPU.

    python tests/fused_check.py [B] [E] [K]          (also used by tests/test_gpu_fused_step.py)

Runs cvcl_flat_step_fused with phase_limit = 1..5 and 0 and compares every workspace block with a
torch restatement ON THE SAME bf16-ROUNDED OPERANDS (so the differences are accumulation order only),
then prints the in-kernel phase timeline (globaltimer stamps of CTA 0).  Test infrastructure.
"""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import multimodal_baby_b200 as cv
from multimodal_baby_b200 import _cabi
from oracle import cvcl_oracle as O

DEV = torch.device("cuda", 0)
S = float(-np.log(0.07))


def reference(x16, w16, b, table, ids, lens, s, normalize=True):
    """fp32 torch restatement on the bf16-rounded operands (device)."""
    B = x16.shape[0]
    u = x16.float() @ w16.float().t() + b
    invn_i = 1.0 / u.norm(dim=1).clamp_min(1e-12) if normalize else torch.ones(B, device=u.device)
    img = u * invn_i[:, None]
    emb = table[ids]                                   # [B, L, E]
    m = emb.sum(1) / lens[:, None].float()
    invn_t = 1.0 / m.norm(dim=1).clamp_min(1e-12) if normalize else torch.ones(B, device=u.device)
    txt = m * invn_t[:, None]
    img16, txt16 = img.bfloat16(), txt.bfloat16()
    scale = math.exp(s)
    Sm = scale * (img16.float() @ txt16.float().t())
    lse0 = torch.logsumexp(Sm, 1); lse1 = torch.logsumexp(Sm, 0)
    diag = Sm.diagonal()
    loss = 0.5 * ((lse0 - diag).mean() + (lse1 - diag).mean())
    P0 = torch.softmax(Sm, 1); P1 = torch.softmax(Sm, 0)
    ent0 = (lse0 - (P0 * Sm).sum(1)).mean(); ent1 = (lse1 - (P1 * Sm).sum(0)).mean()
    acc0 = (Sm.argmax(1) == torch.arange(B, device=u.device)).float().mean()
    acc1 = (Sm.argmax(0) == torch.arange(B, device=u.device)).float().mean()
    coef = 0.5 / B
    Gs = scale * coef * (P0 + P1)
    Gs16 = Gs.bfloat16().float()
    dcoef = -2.0 * scale * coef
    dI = Gs16 @ txt16.float() + dcoef * txt16.float()
    dT = Gs16.t() @ img16.float() + dcoef * img16.float()
    if normalize:
        du = (dI - img16.float() * (img16.float() * dI).sum(1, keepdim=True)) * invn_i[:, None]
        dm = (dT - txt16.float() * (txt16.float() * dT).sum(1, keepdim=True)) * invn_t[:, None]
    else:
        du, dm = dI, dT
    dm = dm / lens[:, None].float()
    du16 = du.bfloat16()
    dW = du16.float().t() @ x16.float()
    db = du.sum(0)
    dtable = torch.zeros_like(table)
    L = ids.shape[1]
    dtable.index_add_(0, ids.reshape(-1), dm[:, None, :].expand(B, L, dm.shape[1]).reshape(B * L, -1))
    dtable[0] = 0
    G = coef * (P0 + P1) - 2 * coef * torch.eye(B, device=u.device)
    ds = (G * Sm).sum()
    return dict(u=u, img16=img16, txt16=txt16, invn_i=invn_i, invn_t=invn_t, lse0=lse0, lse1=lse1, loss=loss,
                ent0=ent0, ent1=ent1, acc0=acc0, acc1=acc1, dI=dI - dcoef * txt16.float(), dT=dT - dcoef * img16.float(),
                du16=du16, dm=dm, dW=dW, db=db, dtable=dtable, ds=ds, txt=txt)


def relerr(a, b):
    a = a.double(); b = b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def run(B, E, K, V=2350, L=25, normalize=True, need_grads=True, verbose=True):
    rng = np.random.RandomState(7 + B)
    f = O.synth_trunk_features(rng, (B, K))
    ids_np, lens_np = O.synth_tokens(rng, B, L, V)
    W, b, table = O.synth_weights(np.random.RandomState(0), E, K, V)
    x16 = torch.from_numpy(f).to(DEV).bfloat16().contiguous()
    w16 = torch.from_numpy(W).to(DEV).bfloat16().contiguous()
    b_d = torch.from_numpy(b).to(DEV); tab = torch.from_numpy(table).to(DEV)
    ids = torch.from_numpy(ids_np).to(DEV); lens = torch.from_numpy(lens_np).to(DEV)
    ref = reference(x16, w16, b_d, tab, ids, lens, S, normalize)
    lay = cv.ops.fused_layout(B, L, E, K, V)
    Bp, KS, nPart, nCB = lay["Bp"], lay["KS"], lay["nPart"], lay["nCB"]
    ws = torch.zeros(lay["bytes"], dtype=torch.uint8, device=DEV)
    f32 = dict(dtype=torch.float32, device=DEV)
    out5 = torch.zeros(8, **f32)
    img_f = torch.zeros(B, E, **f32); txt_f = torch.zeros(B, E, **f32)
    flat = torch.full((4 + E + V * E + E * K,), float("nan"), **f32)
    ds, db, dtable, dW = cv.ops.split_flat_grads(flat, E, K, V)
    p = cv.ops._p
    st = torch.cuda.current_stream().cuda_stream

    def launch(limit):
        _cabi.call("cvcl_flat_step_fused", p(x16), p(w16), p(ids), p(lens), p(b_d), p(tab), B, L, E, K, V,
                   int(normalize), S, None, int(need_grads), p(ws), p(out5), p(img_f), p(txt_f), p(dW), p(db), p(dtable),
                   p(ds), None, limit, st)
        torch.cuda.synchronize()

    def blk(name, nbytes, dtype):
        return ws[lay[name]:lay[name] + nbytes].view(dtype)

    rep = {}
    launch(1)
    hp = blk("hpart", 4 * KS * Bp * E, torch.float32).view(KS, Bp, E)[:, :B].sum(0)
    rep["P0 head partial sum vs x.W^T"] = relerr(hp + b_d, ref["u"])
    txt16 = blk("txt16", 2 * Bp * E, torch.bfloat16).view(Bp, E)[:B]
    rep["P0 txt16"] = relerr(txt16.float(), ref["txt16"].float())
    rep["P0 txt_f32 max abs"] = float((txt_f - ref["txt"]).abs().max())
    launch(2)
    img16 = blk("img16", 2 * Bp * E, torch.bfloat16).view(Bp, E)[:B]
    rep["P1 img16"] = relerr(img16.float(), ref["img16"].float())
    invn = blk("invn", 4 * 2 * Bp, torch.float32).view(2, Bp)[:, :B]
    rep["P1 invn_i"] = relerr(invn[0], ref["invn_i"]); rep["P0 invn_t"] = relerr(invn[1], ref["invn_t"])
    launch(3)
    part = blk("part", 16 * 2 * 2 * nCB * Bp, torch.float32).view(2, 2 * nCB, Bp, 4)[:, :, :B]
    mx = part[..., 0]; l = part[..., 1]
    gm = mx.max(1).values
    lse = gm + torch.log((l * torch.exp(mx - gm[:, None])).sum(1))
    rep["P2 lse0 max abs"] = float((lse[0] - ref["lse0"]).abs().max())
    rep["P2 lse1 max abs"] = float((lse[1] - ref["lse1"]).abs().max())
    if need_grads:
        launch(4)
        dq = blk("dqpart", 4 * 2 * nPart * Bp * E, torch.float32).view(2, nPart, Bp, E)[:, :, :B].sum(1)
        rep["P3 dI (no diag term)"] = relerr(dq[0], ref["dI"]); rep["P3 dT (no diag term)"] = relerr(dq[1], ref["dT"])
        lse_k = blk("lse", 4 * 2 * Bp, torch.float32).view(2, Bp)[:, :B]
        rep["P3 lse0 max abs"] = float((lse_k[0] - ref["lse0"]).abs().max())
        launch(5)
        du16 = blk("du16", 2 * Bp * E, torch.bfloat16).view(Bp, E)[:B]
        rep["P4 du16"] = relerr(du16.float(), ref["du16"].float())
        dm16 = blk("dm16", 2 * Bp * E, torch.bfloat16).view(Bp, E)[:B]
        rep["P4 dm16"] = relerr(dm16.float(), ref["dm"].bfloat16().float())
        cm = blk("cmat", 2 * Bp * lay["Vp"], torch.bfloat16).view(Bp, lay["Vp"])[:B, :V].float()
        cref = torch.zeros(B, V, device=DEV)
        cref.scatter_add_(1, ids, torch.ones_like(ids, dtype=torch.float32)); cref[:, 0] = 0
        rep["P1 token counts exact"] = bool(torch.equal(cm, cref))
    flat.fill_(float("nan")); out5.zero_()
    launch(0)
    rep["loss"] = (float(out5[0]), float(ref["loss"]))
    rep["acc"] = (float(out5[1]), float(ref["acc0"]), float(out5[2]), float(ref["acc1"]))
    rep["ent"] = (float(out5[3]), float(ref["ent0"]), float(out5[4]), float(ref["ent1"]))
    if need_grads:
        rep["dW"] = relerr(dW, ref["dW"]); rep["db"] = relerr(db, ref["db"]); rep["dtable"] = relerr(dtable, ref["dtable"])
        rep["dtable row0 zero"] = float(dtable[0].abs().max())
        rep["ds"] = (float(ds[0]), float(ref["ds"]))
        rep["nan in grads"] = bool(torch.isnan(flat[0:1]).any() or torch.isnan(flat[4:]).any())
    # replay stability (the barrier counter must come back to zero) + timeline
    o1 = out5.clone(); g1 = flat.clone()
    for _ in range(5):
        launch(0)
    rep["replay: out5 bit-identical"] = bool(torch.equal(o1[:5], out5[:5]))
    if need_grads:
        rep["replay: all gradients bit-identical"] = bool(torch.equal(g1[0:1], flat[0:1]) and torch.equal(g1[4:], flat[4:]))
    tm = ws[lay["ctrl"] + 128:lay["ctrl"] + 640].view(torch.int64).cpu().numpy()
    names = ["start", "P0 head+text", "P1 normalise", "P2 similarity", "P3 Gs+dQ", "P4 finish"]
    line = []
    for k in range(1, 6):
        if tm[k]:
            line.append("%s %.2f us" % (names[k], (tm[k] - tm[k - 1]) / 1e3))
    if tm[15]:
        last = max(k for k in range(6) if tm[k])
        line.append("P5 dW+dtable+final %.2f us" % ((tm[15] - tm[last]) / 1e3))
        line.append("total %.2f us" % ((tm[15] - tm[0]) / 1e3))
    rep["timeline (CTA 0, warm L2)"] = "; ".join(line)
    fine = {16: ("P0 accumulator ready", 0), 17: ("P0 slab stored", 0), 18: ("P0 text queue drained", 0), 19: ("P1 rows done", 1),
            20: ("P2 accumulator ready", 2), 21: ("P2 statistics done", 2), 22: ("P3 LSEs merged", 3), 23: ("P3 Gs written", 3),
            24: ("P3 accumulator ready", 3), 25: ("P3 slab stored", 3), 26: ("P4 rows done", 4), 27: ("P5 accumulator ready", 5),
            28: ("P5 tile stored", 5)}
    rep["in-phase stamps (us after the phase began)"] = "; ".join(
        "%s +%.2f" % (nm, (tm[i] - tm[ph]) / 1e3) for i, (nm, ph) in fine.items() if tm[i] and tm[ph])
    rep["control block after run"] = ws[:16].view(torch.int32).cpu().tolist()
    ws[lay["ctrl"] + 128:lay["ctrl"] + 640].zero_()
    launch(100)
    tb = ws[lay["ctrl"] + 128:lay["ctrl"] + 640].view(torch.int64).cpu().numpy()
    rep["grid barrier alone (us each, 6 in a row)"] = " ".join("%.2f" % ((tb[k + 1] - tb[k]) / 1e3) for k in range(1, 6))
    if verbose:
        for k, v in rep.items():
            print("%-36s %s" % (k, v))
    return rep


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    E = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    K = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
    rep = run(B, E, K)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "fused_debug_%d_%d_%d.json" % (B, E, K)), "w") as fh:
        json.dump(rep, fh, indent=1)
