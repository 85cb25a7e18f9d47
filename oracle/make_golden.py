"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on
seeded synthetic inputs.  TEST INFRASTRUCTURE; runs only in the build container.

    python oracle/make_golden.py            # rewrites tests/golden/

Inputs and weights are drawn from numpy RandomState(seed) (bit-stable across versions) by
oracle.cvcl_oracle.synth_*, copied into the reference modules' own parameters, and the
reference's MultiModalModel.forward / calculate_contrastive_loss (+ loss.backward()) are
run as they are.  Only seeds, shapes and OUTPUTS are stored; tests regenerate the inputs.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cvcl_oracle as O          # noqa: E402
from oracle import ref_import as R           # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def case_inputs(seed, B, E, embedding_type, L=25, K=2048, V=2350, Bt=None):
    """The single source of truth for golden-case inputs (tests call this too)."""
    rng = np.random.RandomState(seed)
    W, b, table = O.synth_weights(rng, E, K, V)
    shape = (B, K) if embedding_type == "flat" else (B, K, 7, 7)
    f = O.synth_trunk_features(rng, shape)
    ids, lens = O.synth_tokens(rng, B if Bt is None else Bt, L, V)
    return dict(W=W, b=b, table=table, f=f, ids=ids, lens=lens)


def load_into_reference(model, inp, embedding_type):
    with torch.no_grad():
        if embedding_type == "flat":
            model.image_embed.model.fc.weight.copy_(torch.from_numpy(inp["W"]))
            model.image_embed.model.fc.bias.copy_(torch.from_numpy(inp["b"]))
        else:
            conv = model.image_embed.model[-1]
            conv.weight.copy_(torch.from_numpy(inp["W"])[:, :, None, None])
            conv.bias.copy_(torch.from_numpy(inp["b"]))
        model.text_embed.embedding.weight.copy_(torch.from_numpy(inp["table"]))


def run_train_case(name, seed, B, E, embedding_type, sim="mean", store_full_grads=True):
    inp = case_inputs(seed, B, E, embedding_type)
    m = R.build_reference_model(embedding_type, sim=sim, embedding_dim=E,
                                fix_temperature=False)
    load_into_reference(m, inp, embedding_type)
    m.train()
    f = torch.from_numpy(inp["f"]); ids = torch.from_numpy(inp["ids"])
    lens = torch.from_numpy(inp["lens"])
    out = m.calculate_contrastive_loss(f, ids, lens)
    loss = out[0]
    loss.backward()
    head = m.image_embed.model.fc if embedding_type == "flat" else m.image_embed.model[-1]
    dW = head.weight.grad.reshape(E, -1).numpy()
    dtab = m.text_embed.embedding.weight.grad.numpy()
    rec = dict(
        seed=seed, B=B, E=E, embedding_type=embedding_type, sim=sim,
        loss=loss.item(), image_accuracy=out[1].item(), text_accuracy=out[2].item(),
        image_entropy=out[3].item(), text_entropy=out[4].item(),
        logits_per_image=out[5].detach().numpy(), logits_per_text=out[6].detach().numpy(),
        image_pred=torch.argmax(out[5], -1).numpy(), text_pred=torch.argmax(out[6], -1).numpy(),
        db=head.bias.grad.numpy(), ds=m.logit_neg_log_temperature.grad.item(),
        dW_norm=np.linalg.norm(dW.astype(np.float64)),
        dtable_norm=np.linalg.norm(dtab.astype(np.float64)),
        dW_slice=dW[:8, :64].copy(), dtable_rows=dtab[:8].copy(),
    )
    if embedding_type == "flat":
        rec["image_features"] = out[7].detach().numpy()
        tf, _ = m.encode_text(ids, lens)
        rec["text_features"] = tf.detach().numpy()
    else:
        rec["image_features_head"] = out[7].detach().numpy()[:2]       # [2,E,7,7]
    if store_full_grads:
        rec["dW"] = dW
        nz = np.unique(inp["ids"])
        rec["dtable_nz_ids"] = nz
        rec["dtable_nz"] = dtab[nz]
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **rec)
    print(name, "loss", rec["loss"], "ds", rec["ds"])


def run_forward_case(name, seed, Ni, Nt, E):
    """README usage (4 images x 3 texts): Ni != Nt through MultiModalModel.forward."""
    inp = case_inputs(seed, Ni, E, "flat", Bt=Nt)
    m = R.build_reference_model("flat", embedding_dim=E, fix_temperature=True)
    load_into_reference(m, inp, "flat")
    m.eval()
    with torch.no_grad():
        lpi, lpt = m(torch.from_numpy(inp["f"]), torch.from_numpy(inp["ids"]),
                     torch.from_numpy(inp["lens"]))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), seed=seed, Ni=Ni, Nt=Nt, E=E,
                        logits_per_image=lpi.numpy(), logits_per_text=lpt.numpy())
    print(name, lpi.shape, lpt.shape)


def run_eval_case(name, seed, n_trials, E, n_way=4):
    """Labeled-S trials through the reference's literal loop (eval.py:196-214):
    per trial model(img[4], label[1,L], len[1]) -> argmax(logits_per_text[0])."""
    rng = np.random.RandomState(seed)
    W, b, table = O.synth_weights(rng, E)
    f = O.synth_trunk_features(rng, (n_trials, n_way, 2048))
    vocab = R.reference_vocab()
    cats = ["ball", "car", "kitty", "chair", "door", "hand", "sand", "window"]
    word_ids = np.array([vocab[c] for c in cats], np.int64)
    which = rng.randint(0, len(cats), size=n_trials)
    ids = np.zeros((n_trials, 3), np.int64)
    ids[:, 0] = O.SOS_TOKEN_ID; ids[:, 1] = word_ids[which]; ids[:, 2] = O.EOS_TOKEN_ID
    lens = np.full(n_trials, 3, np.int64)
    m = R.build_reference_model("flat", embedding_dim=E, fix_temperature=True)
    load_into_reference(m, dict(W=W, b=b, table=table), "flat")
    m.eval()
    preds, logits = [], []
    with torch.no_grad():
        for t in range(n_trials):
            _, lpt = m(torch.from_numpy(f[t]), torch.from_numpy(ids[t:t + 1]),
                       torch.from_numpy(lens[t:t + 1]))
            preds.append(int(torch.argmax(lpt[0])))
            logits.append(lpt[0].numpy())
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), seed=seed, n_trials=n_trials, E=E,
                        n_way=n_way, ids=ids, lens=lens, pred=np.array(preds, np.int32),
                        logits=np.stack(logits))
    print(name, "acc", np.mean(np.array(preds) == 0))


def gradcam_inputs(seed, N, E, K=2048, HW=7):
    """inputs of the Grad-CAM golden case (tests call this too): post-ReLU layer4 activations, the
    reference's fc init, L2-normalised text features as the backward target."""
    rng = np.random.RandomState(seed)
    W, b, _ = O.synth_weights(rng, E, K, 8)
    act = O.synth_trunk_features(rng, (N, K, HW, HW))
    tgt = rng.standard_normal((N, E)).astype(np.float32)
    tgt /= np.linalg.norm(tgt, axis=1, keepdims=True)
    return dict(W=W, b=b, act=act, target=tgt.astype(np.float32))


def run_gradcam_case(name, seed, N, E):
    """attention_maps.gradCAM of the UNMODIFIED reference on a head-only vision model whose saliency layer
    (`layer4`) is an identity fed the activation map (the trunk stays outside the product); the resize is
    the reference's own F.interpolate call (attention_maps.py:158-163) applied to 224 x 224."""
    import collections
    R.load_reference()
    from multimodal import attention_maps as ref_am
    inp = gradcam_inputs(seed, N, E)
    fc = torch.nn.Linear(inp["W"].shape[1], E)
    with torch.no_grad():
        fc.weight.copy_(torch.from_numpy(inp["W"])); fc.bias.copy_(torch.from_numpy(inp["b"]))
    model = torch.nn.Sequential(collections.OrderedDict(
        layer4=torch.nn.Identity(), avgpool=torch.nn.AdaptiveAvgPool2d((1, 1)), flatten=torch.nn.Flatten(1), fc=fc))
    out = {}
    for norm in (True, False):
        cams = []
        for i in range(N):                      # the reference calls it image by image (generate_attention_maps.py:103-110)
            cam = ref_am.gradCAM(model, torch.from_numpy(inp["act"][i:i + 1]).clone(),
                                 torch.from_numpy(inp["target"][i:i + 1]), model.layer4,
                                 normalize_features=norm, resize=False)
            cams.append(cam.detach())
        cam = torch.cat(cams)
        big = torch.nn.functional.interpolate(cam, (224, 224), mode="bicubic", align_corners=False)
        key = "norm" if norm else "raw"
        out["cam_" + key] = cam.numpy()
        out["resized_" + key] = big.numpy()[:, :, ::3, ::3].copy()       # every third pixel: 75 x 75
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), seed=seed, N=N, E=E, **out)
    print(name, "cam max", float(out["cam_norm"].max()), "nonzero frac", float((out["cam_norm"] > 0).mean()))


def collate_batch(seed=112):
    """a batch as the reference's datasets yield it (img, token row, length, raw utterance) with lengths
    around and beyond MAX_LEN_UTTERANCE = 25; the tests rebuild it from the seed."""
    rng = np.random.RandomState(seed)
    lens = [3, 25, 31, 1, 26, 12, 40]
    batch = []
    for i, n in enumerate(lens):
        row = rng.randint(4, 2350, size=n).astype(np.int64)
        row[0] = O.SOS_TOKEN_ID; row[-1] = O.EOS_TOKEN_ID
        batch.append((torch.full((3, 4, 4), float(i)), torch.from_numpy(row), n, "utt %d" % i))
    return batch


def run_collate_case(name, long=True):
    """multiModalDataset_collate_fn of the UNMODIFIED reference (multimodal_data_module.py:98-109)."""
    R.load_reference()
    from multimodal import multimodal_data_module as ref_dm
    batch = collate_batch()
    if not long:
        batch = [b for b in batch if b[2] <= 12]
    img, ids, lens, raw = ref_dm.multiModalDataset_collate_fn(batch)
    with open(os.path.join(GOLD, name + ".json"), "w") as fh:
        json.dump(dict(long=long, img_shape=list(img.shape), img_first=[float(v) for v in img[:, 0, 0, 0]],
                       ids=ids.tolist(), lens=lens.tolist(), raw=raw, ids_dtype=str(ids.dtype),
                       lens_dtype=str(lens.dtype)), fh)
    print(name, tuple(ids.shape), lens.tolist())


def run_tokenize_case(name):
    """MultiModalLitModel.tokenize (multimodal_lit.py:161-190) on pre-tokenised text
    (whitespace split; spaCy itself is not available offline)."""
    m = R.build_reference_model("flat", embedding_dim=64, fix_temperature=True, lit=True)
    texts = ["ball", "look at the kitty", "where is the zzzunknownzzz ball ?",
             " ".join(["car"] * 40), "a"]
    ids, lens = m.tokenize(texts)
    with open(os.path.join(GOLD, name + ".json"), "w") as fh:
        json.dump(dict(texts=texts, ids=ids.tolist(), lens=lens.tolist()), fh)
    print(name, lens.tolist())


def run_state_dict_case(name):
    """parameter attribute paths the drop-in must keep (load_from_checkpoint compatibility)."""
    rec = {}
    for et in ("flat", "spatial"):
        for fix in (False, True):
            m = R.build_reference_model(et, embedding_dim=64, fix_temperature=fix, lit=True)
            rec["%s_fix%d" % (et, int(fix))] = {
                "state_dict": sorted(m.state_dict().keys()),
                "trainable": sorted(n for n, p in m.named_parameters() if p.requires_grad),
                "temperature_is_parameter": isinstance(m.model.logit_neg_log_temperature, torch.nn.Parameter),
                "temperature_value": float(m.model.logit_neg_log_temperature),
            }
    with open(os.path.join(GOLD, name + ".json"), "w") as fh:
        json.dump(rec, fh, indent=1)
    print(name, list(rec))


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    run_train_case("flat_e64_b8", 101, 8, 64, "flat")
    run_train_case("flat_e512_b32", 102, 32, 512, "flat", store_full_grads=False)
    run_train_case("flat_e512_b160", 103, 160, 512, "flat", store_full_grads=False)
    run_train_case("spatial_mean_e64_b6", 104, 6, 64, "spatial", "mean")
    run_train_case("spatial_max_e64_b6", 105, 6, 64, "spatial", "max")
    run_train_case("spatial_max_e512_b12", 106, 12, 512, "spatial", "max",
                   store_full_grads=False)
    run_train_case("spatial_mean_e512_b12", 107, 12, 512, "spatial", "mean",
                   store_full_grads=False)
    run_forward_case("forward_4x3_e512", 108, 4, 3, 512)
    run_forward_case("forward_4x1_e512", 109, 4, 1, 512)
    run_eval_case("eval_4way_e512", 110, 64, 512)
    run_gradcam_case("gradcam_e512_n3", 111, 3, 512)
    run_collate_case("collate_long", True)
    run_collate_case("collate_short", False)
    run_tokenize_case("tokenize")
    run_state_dict_case("state_dict_keys")


if __name__ == "__main__":
    main()
