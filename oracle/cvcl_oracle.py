"""CPU oracle for the CVCL contrastive hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

A restatement, op for op, of the reference's algorithm for the path named by
BASELINE.json `north_star` (wkvong/multimodal-baby, `multimodal/multimodal.py`).  Every
function cites the reference file:line it follows (paths relative to /root/reference).
The reference computes this path with stock ATen ops in fp32; the oracle issues the same
ATen ops on CPU tensors (dtype selectable: float32 to mirror the reference, float64 for
closed-form checks), plus an independent numpy float64 closed form of the backward
 so that autograd itself is cross-checked.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` legs may import this module, and only as the checker / reported baseline.
The product (`multimodal-baby_b200/`) never imports it.

PARITY PINNING: the reference's own tests hold no golden vectors for this path
(SURVEY.md section 4).  The oracle is instead pinned against outputs of the *unmodified
reference itself*, run in the build container by `oracle/make_golden.py`
(-> `tests/golden/*.npz`, committed) and, when /root/reference is present, live in
`tests/test_oracle_vs_reference.py`.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F

# multimodal/multimodal_data_module.py:37-50
PAD_TOKEN_ID, UNK_TOKEN_ID, SOS_TOKEN_ID, EOS_TOKEN_ID = 0, 1, 2, 3
MAX_LEN_UTTERANCE = 25


# --------------------------------------------------------------------------------------
# text side
# --------------------------------------------------------------------------------------
def embedding_lookup(ids: torch.Tensor, table: torch.Tensor) -> torch.Tensor:
    """multimodal.py:496 `embedding = self.embedding(x)` -> [B, L, E].
    nn.Embedding(padding_idx=0) (multimodal.py:311-312) is a plain row gather in the
    forward; the padding row is zero by construction and receives zero gradient."""
    return F.embedding(ids, table, padding_idx=PAD_TOKEN_ID)


def text_encoder_flat(ids, lens, table):
    """multimodal.py:499-503: mean embedding per utterance, `sum(dim=1) / x_len[:,None]`.
    ALL L positions are summed (pads hit the zero row); the divisor is x_len.
    Returns (ret [B,E], output [B,L,E]) as multimodal.py:575-584 with dropout_o = 0."""
    emb = embedding_lookup(ids, table)
    ret = torch.sum(emb, dim=1) / lens.unsqueeze(1)
    return ret, emb


def text_encoder_spatial(ids, lens, table):
    """multimodal.py:498,579-580: `ret = output = embedding` (no pooling)."""
    emb = embedding_lookup(ids, table)
    return emb, emb


def l2_normalize(x, dim):
    """multimodal.py:736,743: F.normalize(x, p=2, dim) = x / max(||x||_2, 1e-12)."""
    return F.normalize(x, p=2, dim=dim)


# --------------------------------------------------------------------------------------
# image side (projection head only; the ResNeXt trunk is outside the path)
# --------------------------------------------------------------------------------------
def head_flat(f, W, b):
    """multimodal.py:186-192 (`model.fc = nn.Linear(2048, E)`) applied at :101."""
    return F.linear(f, W, b)


def head_spatial(fmap, W, b):
    """multimodal.py:181-185: 1x1 Conv2d(2048 -> E) on the layer4 map, NCHW.
    W may be [E,K] or [E,K,1,1]."""
    if W.dim() == 2:
        W = W[:, :, None, None]
    return F.conv2d(fmap, W, b)


# --------------------------------------------------------------------------------------
# similarity + temperature
# --------------------------------------------------------------------------------------
def similarity_flat(img, txt):
    """multimodal.py:755 `match = image_features @ text_features.T`."""
    return img @ txt.T


def similarity_spatial_mean(img, txt, lens):
    """multimodal.py:765-770; img [Bi,E,H,W], txt [Bt,L,E], lens broadcast over columns."""
    match_sum = torch.einsum('iehw,tle->it', [img, txt])
    return match_sum / (img.size(-2) * img.size(-1) * lens)


def similarity_spatial_max(img, txt, lens):
    """multimodal.py:775-780: max over the H*W locations per word, sum over words / len."""
    match_max = torch.einsum('iehw,tle->itlhw', [img, txt])
    match_max = torch.amax(match_max, dim=(3, 4))
    return torch.sum(match_max, dim=2) / lens


def logits_from_match(match, s):
    """multimodal.py:783-787: logit_scale = exp(s), s = -log(temperature)."""
    scale = s.exp() if torch.is_tensor(s) else math.exp(s)
    return match * scale, match.t() * scale


def get_entropy(logits, dim=-1):
    """multimodal/utils.py:106-108."""
    log_p = F.log_softmax(logits, dim=dim)
    return (F.softmax(log_p, dim=dim) * -log_p).sum(dim=dim)


@dataclass
class InfoNCE:
    loss: torch.Tensor
    image_accuracy: torch.Tensor
    text_accuracy: torch.Tensor
    image_entropy: torch.Tensor
    text_entropy: torch.Tensor
    image_pred: torch.Tensor
    text_pred: torch.Tensor


def infonce(lpi, lpt, label_offset: int = 0) -> InfoNCE:
    """multimodal.py:801-818.  labels = arange(B) (+label_offset for a row shard whose
    diagonal sits at a column offset -- the sharded extension of SURVEY section 8e)."""
    B = lpi.size(0)
    gt = torch.arange(B, dtype=torch.long) + label_offset
    loss = (F.cross_entropy(lpi, gt) + F.cross_entropy(lpt, gt)).div(2)
    ip = torch.argmax(lpi, dim=-1)
    tp = torch.argmax(lpt, dim=-1)
    return InfoNCE(loss, (ip == gt).sum() / B, (tp == gt).sum() / B,
                   get_entropy(lpi).mean(), get_entropy(lpt).mean(), ip, tp)


# --------------------------------------------------------------------------------------
# whole-path restatements (forward + autograd backward)
# --------------------------------------------------------------------------------------
def encode_image(f, W, b, embedding_type="flat", normalize=True):
    """multimodal.py:732-737.  flat: f [B,K] -> [B,E]; spatial: f [B,K,H,W] -> [B,E,H,W],
    normalised over the channel dim (dim=1) in both cases."""
    u = head_flat(f, W, b) if embedding_type == "flat" else head_spatial(f, W, b)
    return l2_normalize(u, 1) if normalize else u


def encode_text(ids, lens, table, embedding_type="flat", normalize=True):
    """multimodal.py:739-744 -> (text_features, text_outputs)."""
    if embedding_type == "flat":
        ret, out = text_encoder_flat(ids, lens, table)
    else:
        ret, out = text_encoder_spatial(ids, lens, table)
    return (l2_normalize(ret, -1) if normalize else ret), out


def _round_bf16_ste(x):
    """round to bf16 with a straight-through gradient: emulates the kernels' operand precision
    (bf16 features, fp32 accumulation) inside the otherwise unchanged reference algorithm."""
    return x + (x.detach().to(torch.bfloat16).to(x.dtype) - x.detach())


def forward(f, ids, lens, W, b, table, s, embedding_type="flat", sim="mean", normalize=True,
            feature_round=None):
    """multimodal.py:746-794 -> (logits_per_image, logits_per_text, img_feat, txt_feat).
    feature_round="bf16" (not in the reference) rounds the encoded features to bf16 before the
    similarity, which is the operand precision BASELINE.json north_star prescribes."""
    img = encode_image(f, W, b, embedding_type, normalize)
    txt, _ = encode_text(ids, lens, table, embedding_type, normalize)
    if feature_round == "bf16":
        img, txt = _round_bf16_ste(img), _round_bf16_ste(txt)
    if embedding_type == "flat":
        match = similarity_flat(img, txt)
    elif sim == "mean":
        match = similarity_spatial_mean(img, txt, lens)
    else:
        match = similarity_spatial_max(img, txt, lens)
    lpi, lpt = logits_from_match(match, s)
    return lpi, lpt, img, txt


def contrastive_step(f, ids, lens, W, b, table, s, embedding_type="flat", sim="mean",
                     normalize=True, dtype=torch.float32, need_df=False, feature_round=None):
    """multimodal.py:796-822 + loss.backward(): returns a dict with the forward outputs
    and the gradients of the four trainable tensors (+ df if need_df)."""
    W = W.detach().to(dtype).clone().requires_grad_(True)
    b = b.detach().to(dtype).clone().requires_grad_(True)
    table = table.detach().to(dtype).clone().requires_grad_(True)
    s = torch.as_tensor(s, dtype=torch.float64).detach().to(dtype).clone().requires_grad_(True)
    f = f.detach().to(dtype).clone().requires_grad_(need_df)
    lpi, lpt, img, txt = forward(f, ids, lens, W, b, table, s, embedding_type, sim, normalize,
                                 feature_round)
    res = infonce(lpi, lpt)
    res.loss.backward()
    dtab = table.grad.clone()
    dtab[PAD_TOKEN_ID].zero_()          # nn.Embedding(padding_idx=0): row 0 gets no grad
    out = dict(loss=res.loss.detach(), image_accuracy=res.image_accuracy,
               text_accuracy=res.text_accuracy, image_entropy=res.image_entropy.detach(),
               text_entropy=res.text_entropy.detach(), image_pred=res.image_pred,
               text_pred=res.text_pred, logits_per_image=lpi.detach(),
               logits_per_text=lpt.detach(), image_features=img.detach(),
               text_features=txt.detach(), dW=W.grad, db=b.grad, dtable=dtab, ds=s.grad)
    if need_df:
        out["df"] = f.grad
    return out


def sharded_contrastive_loss(img_all, txt_all, s, world_size):
    """SURVEY section 8e restated on CPU: rank r owns pairs [r*b,(r+1)*b); it computes
    its row block (image->text CE) and column block (text->image CE) against the
    all-gathered features.  Returns per-rank partial sums whose total equals the
    single-process global-batch loss of multimodal.py:808-810."""
    B = img_all.size(0)
    b = B // world_size
    scale = s.exp() if torch.is_tensor(s) else math.exp(s)
    parts = []
    for r in range(world_size):
        sl = slice(r * b, (r + 1) * b)
        rows = img_all[sl] @ txt_all.T * scale          # [b, B]
        cols = txt_all[sl] @ img_all.T * scale          # [b, B]
        gt = torch.arange(b) + r * b
        parts.append((F.cross_entropy(rows, gt, reduction="sum")
                      + F.cross_entropy(cols, gt, reduction="sum")) / (2 * B))
    return parts


# --------------------------------------------------------------------------------------
# independent closed-form backward, numpy float64
# --------------------------------------------------------------------------------------
def closed_form_flat_backward(f, ids, lens, W, b, table, s):
    """G = (P_row + P_col - 2I)/(2B); dI = e^s G T; dT = e^s G^T I; ds = sum(G*S);
    normalise-bwd du = (dI - I<I,dI>)/||u||; dW = du^T f; db = sum du;
    dm = (dT - T<T,dT>)/||m||; dE[v] += dm[b]/len[b] for each l with ids[b,l]=v; dE[0]=0."""
    f = np.asarray(f, np.float64); W = np.asarray(W, np.float64); b = np.asarray(b, np.float64)
    table = np.asarray(table, np.float64); ids = np.asarray(ids); lens = np.asarray(lens)
    s = float(s)
    B = f.shape[0]
    u = f @ W.T + b
    nu = np.maximum(np.linalg.norm(u, axis=1, keepdims=True), 1e-12)
    I = u / nu
    m = table[ids].sum(1) / lens[:, None]
    nm = np.maximum(np.linalg.norm(m, axis=1, keepdims=True), 1e-12)
    T = m / nm
    S = math.exp(s) * (I @ T.T)

    def lse(x, axis):
        mx = x.max(axis=axis, keepdims=True)
        return mx + np.log(np.exp(x - mx).sum(axis=axis, keepdims=True))
    lr, lc = lse(S, 1), lse(S, 0)
    loss = (-(np.diag(S) - lr[:, 0]).mean() - (np.diag(S) - lc[0]).mean()) / 2
    G = (np.exp(S - lr) + np.exp(S - lc) - 2 * np.eye(B)) / (2 * B)
    dI = math.exp(s) * G @ T
    dT = math.exp(s) * G.T @ I
    ds = float((G * S).sum())
    du = (dI - I * (I * dI).sum(1, keepdims=True)) / nu
    dm = (dT - T * (T * dT).sum(1, keepdims=True)) / nm
    dW = du.T @ f
    db = du.sum(0)
    dtab = np.zeros_like(table)
    np.add.at(dtab, ids.reshape(-1), np.repeat(dm / lens[:, None], ids.shape[1], axis=0))
    dtab[PAD_TOKEN_ID] = 0
    return dict(loss=loss, dW=dW, db=db, dtable=dtab, ds=ds, dI=dI, dT=dT, G=G,
                image_features=I, text_features=T, logits_per_image=S,
                row_lse=lr[:, 0], col_lse=lc[0])


# --------------------------------------------------------------------------------------
# Labeled-S style n-way evaluation
# --------------------------------------------------------------------------------------
def eval_nway(img_feat, txt_feat, s=None, normalize=True):
    """multimodal_lit.py:466-511 / eval.py:196-214 for a batch of trials.
    img_feat [N, n_way, E] (target first), txt_feat [N, E] (already pooled, un-normalised).
    Per trial: logits_per_text = scale * T.I^T -> [1, n_way]; pred = argmax (first max)."""
    if normalize:
        img_feat = l2_normalize(img_feat, -1)
        txt_feat = l2_normalize(txt_feat, -1)
    scale = 1.0 if s is None else math.exp(float(s))
    logits = torch.einsum('nwe,ne->nw', img_feat, txt_feat) * scale
    return torch.argmax(logits, dim=-1), logits


def classify_ncat(img_feat, txt_feat, s=None, normalize=True):
    """n-category classification of frames (multimodal_saycam_data_module.py:545-606 / forward(), multimodal.py:
    782-794): logits_per_image = scale * I.T^T over the C category texts; pred = argmax (first max)."""
    if normalize:
        img_feat = l2_normalize(img_feat, -1)
        txt_feat = l2_normalize(txt_feat, -1)
    scale = 1.0 if s is None else math.exp(float(s))
    logits = img_feat @ txt_feat.t() * scale
    return torch.argmax(logits, dim=-1), logits


def cosine_nearest(queries, keys):
    """analysis_cvcl/duplicates.py:561-607: F.normalize both sets, all-pairs cosine, per query (evaluation frame)
    np.max / np.argmax over the keys (training frames)."""
    sim = l2_normalize(queries, -1) @ l2_normalize(keys, -1).t()
    return sim.max(dim=1).values, torch.argmax(sim, dim=1)


def eval_trial_loop(f_trials, ids, lens, W, b, table, s):
    """The reference's literal per-trial loop (eval.py:196-214): one model() call per
    trial on 4 frames + 1 label; used on small N to pin eval_nway()."""
    preds = []
    for t in range(f_trials.shape[0]):
        lpi, lpt, _, _ = forward(f_trials[t], ids[t:t + 1], lens[t:t + 1], W, b, table, s)
        preds.append(int(torch.argmax(lpt[0])))
    return torch.tensor(preds)


# --------------------------------------------------------------------------------------
# synthetic inputs shared by tests, bench and golden generation
# --------------------------------------------------------------------------------------
# ------------------------------------------------------------------------------------------------
# Grad-CAM attention maps: multimodal/attention_maps.py:111-165 restated for the flat
# head (saliency layer = layer4, followed by the trunk's global average pool and `fc`).
# ------------------------------------------------------------------------------------------------
def gradcam_flat(act, W, b, target, normalize=True, out_hw=None):
    """act [N,K,H,W] = layer4 activation, W [E,K], b [E] = fc, target [N,E] (text features).
    The reference's op sequence: output = fc(avgpool(act)) (attention_maps.py:143, torchvision ResNet
    forward), F.normalize (:144-145), output.backward(target) (:146), alpha = grad.mean((2,3)) (:114),
    cam = clamp(sum_c act*alpha, min=0) (:116-119), optional bicubic resize (:158-163).
    -> (cam [N,1,H,W], resized [N,1,*out_hw] | None)."""
    act = act.detach().clone().requires_grad_(True)
    out = F.linear(F.adaptive_avg_pool2d(act, 1).flatten(1), W, b)
    if normalize:
        out = F.normalize(out, p=2, dim=1)
    out.backward(target)
    grad = act.grad
    alpha = grad.mean(dim=(2, 3), keepdim=True)
    cam = torch.clamp(torch.sum(act.detach() * alpha, dim=1, keepdim=True), min=0)
    resized = None
    if out_hw is not None:
        resized = F.interpolate(cam, tuple(out_hw), mode="bicubic", align_corners=False)
    return cam, resized


def gradcam_flat_closed_form(act, W, b, target, normalize=True):
    """The same map without autograd: the head is linear in the pooled activation, so the gradient
    w.r.t. act[n,c,h,w] does not depend on (h,w): alpha[n,c] = (W^T g[n])[c] / (H*W) with
    g = target (no normalisation) or (target - y <y,target>) / ||u|| (F.normalize backward)."""
    N, K, H, Wd = act.shape
    u = F.linear(act.mean(dim=(2, 3)), W, b)
    g = target
    if normalize:
        nrm = u.norm(dim=1, keepdim=True).clamp_min(1e-12)
        y = u / nrm
        g = (target - y * (y * target).sum(1, keepdim=True)) / nrm
    alpha = (g @ W) / float(H * Wd)
    return torch.clamp(torch.einsum("nchw,nc->nhw", act, alpha), min=0).unsqueeze(1)


def synth_tokens(rng: np.random.RandomState, B, L=MAX_LEN_UTTERANCE, V=2350, min_len=3):
    """len ~ U{3..L}; ids[b,0]=<sos>, ids[b,len-1]=<eos>, interior ~ U{4..V-1}, pads 0
    (collate layout of multimodal_data_module.py:98-109)."""
    lens = rng.randint(min_len, L + 1, size=B).astype(np.int64)
    ids = np.zeros((B, L), np.int64)
    for i in range(B):
        n = lens[i]
        ids[i, 0] = SOS_TOKEN_ID
        ids[i, 1:n - 1] = rng.randint(4, V, size=n - 2)
        ids[i, n - 1] = EOS_TOKEN_ID
    return ids, lens


def synth_weights(rng: np.random.RandomState, E=512, K=2048, V=2350):
    """Linear: U(-1/sqrt(K), 1/sqrt(K)) (kaiming-uniform a=sqrt(5)); Embedding: N(0,1),
    row 0 zero -- the reference's init distributions, drawn from numpy for portability."""
    bound = 1.0 / math.sqrt(K)
    W = rng.uniform(-bound, bound, size=(E, K)).astype(np.float32)
    b = rng.uniform(-bound, bound, size=(E,)).astype(np.float32)
    table = rng.standard_normal((V, E)).astype(np.float32)
    table[PAD_TOKEN_ID] = 0
    return W, b, table


def synth_trunk_features(rng: np.random.RandomState, shape):
    """f = relu(N(0,1)): post-ReLU pooled trunk activations are non-negative."""
    return np.maximum(rng.standard_normal(shape), 0).astype(np.float32)
