"""-m gpu: the reference's own test strategy (tests/test_batching.py:59-106): a batched forward must
equal the per-example ("unbatched") forward.  The reference file is stale against its current
signatures, so the property is restated here on the drop-in modules with the same
shapes (B=4, max_len=16, ids in [1,16), lens in [1,16)) and tolerances (atol 1e-5)."""
import argparse

import numpy as np
import pytest
import torch

from _util import O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def random_padded_tensor(rng, batch_size, max_seq_len, lo, hi):
    """tests/test_batching.py:45-57: random ids in [lo,hi), random lens in [1,max_seq_len), zero padded."""
    x = rng.randint(lo, hi, size=(batch_size, max_seq_len)).astype(np.int64)
    x_len = rng.randint(1, max_seq_len, size=batch_size).astype(np.int64)
    for i, n in enumerate(x_len):
        x[i, n:] = 0
    return torch.from_numpy(x), torch.from_numpy(x_len)


def make_text_encoder(cv, embedding_type, E=128, V=16):
    args = argparse.Namespace(embedding_type=embedding_type, embedding_dim=E, text_encoder="embedding",
                              dropout_i=0.0, dropout_o=0.0, crange=1)
    vocab = {str(i): i for i in range(V)}
    return cv.TextEncoder(vocab, 2048, args).to(DEV)


@pytest.fixture(scope="module")
def cv():
    import multimodal_baby_b200 as m
    m._cabi.load()
    return m


def forward_unbatched(model, x, x_len):
    """multimodal.py:586-600 (_forward_unbatched, embedding branch) through the same module."""
    outs = []
    for i in range(x.shape[0]):
        ret, _, _ = model(x[i:i + 1], x_len[i:i + 1])
        outs.append(ret[0])
    return torch.stack(outs)


@pytest.mark.parametrize("embedding_type", ["flat", "spatial"])
def test_text_encoder_batched_equals_unbatched(cv, embedding_type):
    rng = np.random.RandomState(0)
    model = make_text_encoder(cv, embedding_type)
    x, x_len = random_padded_tensor(rng, 4, 16, 1, 16)
    x, x_len = x.to(DEV), x_len.to(DEV)
    batched, _, _ = model(x, x_len)
    unbatched = forward_unbatched(model, x, x_len)
    assert torch.allclose(batched, unbatched, atol=1e-5)
    # and both equal the definition: sum_l E[x_l] / len (flat) or E[x] (spatial)
    emb = model.embedding.weight[x]
    ref = emb.sum(1) / x_len[:, None] if embedding_type == "flat" else emb
    assert torch.allclose(batched, ref, atol=1e-5)


def test_head_batched_equals_unbatched(cv):
    """the image side of the same property (tests/test_batching.py:21-42, test_cnn): rows of a batch
    through the projection head == each row alone (different tile occupancy, same numbers)."""
    rng = np.random.RandomState(1)
    W = torch.from_numpy((rng.standard_normal((128, 2048)) / 45).astype(np.float32)).to(DEV)
    b = torch.from_numpy(rng.standard_normal(128).astype(np.float32) * 0.01).to(DEV)
    x = torch.from_numpy(np.maximum(rng.standard_normal((4, 2048)), 0).astype(np.float32)).to(DEV)
    batched = cv.ops.head_features(x, W, b, True)
    single = torch.cat([cv.ops.head_features(x[i:i + 1], W, b, True) for i in range(4)])
    assert torch.allclose(batched, single, atol=1e-5)


def test_model_forward_batched_equals_per_pair(cv):
    """logits[i, t] of a batched forward == the 1x1 forward of image i with text t."""
    rng = np.random.RandomState(2)
    E = 128
    args = argparse.Namespace(embedding_type="flat", embedding_dim=E, normalize_features=True,
                              fix_temperature=True, temperature=0.07, text_encoder="embedding")
    vocab = {str(i): i for i in range(64)}
    m = cv.MultiModalModel(cv.VisionEncoder(args, trunk="pooled"), cv.TextEncoder(vocab, 2048, args), args).to(DEV).eval()
    x = torch.from_numpy(np.maximum(rng.standard_normal((4, 2048)), 0).astype(np.float32)).to(DEV)
    ids, lens = random_padded_tensor(rng, 3, 16, 1, 64)
    ids, lens = ids.to(DEV), lens.to(DEV)
    with torch.no_grad():
        lpi, lpt = m(x, ids, lens)
        assert lpi.shape == (4, 3) and lpt.shape == (3, 4)
        for i in range(4):
            for j in range(3):
                one, _ = m(x[i:i + 1], ids[j:j + 1], lens[j:j + 1])
                assert abs(one.item() - lpi[i, j].item()) <= 1e-4


def test_write_combined_staging_arena():
    """staging.host_arena: page-locked (optionally write-combined) host memory from the library's own allocator,
    usable as the source of an asynchronous H2D copy and as a packed staging set."""
    import multimodal_baby_b200 as cv
    a = cv.staging.host_arena(1 << 20, True)
    assert a.dtype == torch.uint8 and a.numel() == 1 << 20 and a.is_pinned()
    src = torch.arange(1 << 20, dtype=torch.int64).to(torch.uint8)
    a.copy_(src)
    d = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    d.copy_(a, non_blocking=True)
    torch.cuda.synchronize()
    assert torch.equal(d.cpu(), src)
    x, ids, lens = cv.staging.packed_buffers([((4, 8), torch.bfloat16), ((4, 25), torch.int64), ((4,), torch.int64)],
                                             write_combined=True)
    assert x.is_pinned() and cv.staging.packed_span((x, ids, lens)) is not None
    assert float(x.float().abs().sum()) == 0.0 and int(ids.sum()) == 0
    del a, x, ids, lens                                # the finalizer frees the allocation with the last view


def _lit_model(cv, E=512, fix_temperature=False):
    args = argparse.Namespace(embedding_type="flat", embedding_dim=E, normalize_features=True,
                              fix_temperature=fix_temperature, temperature=0.07, text_encoder="embedding")
    vocab = {"<pad>": 0, "<unk>": 1, "<sos>": 2, "<eos>": 3, **{"w%d" % i: i for i in range(4, 2350)}}
    lit = cv.MultiModalLitModel(cv.VisionEncoder(args, trunk="pooled"), cv.TextEncoder(vocab, 2048, args), args, vocab=vocab)
    return lit.to(DEV)


def test_linear_f32_matches_torch(cv):
    rng = np.random.RandomState(5)
    for M, N, K in ((256, 512, 2048), (70, 130, 64), (1, 512, 2048)):
        x = torch.from_numpy(rng.standard_normal((M, K)).astype(np.float32)).to(DEV)
        w = torch.from_numpy((rng.standard_normal((N, K)) / 45).astype(np.float32)).to(DEV)
        b = torch.from_numpy(rng.standard_normal(N).astype(np.float32)).to(DEV)
        got = cv.ops.linear_f32(x, w, b)
        ref = (x.double() @ w.double().t() + b.double()).float()
        assert float((got - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))


def test_lit_training_and_trial_steps_on_gpu(cv):
    """the reference's Lightning entry points with this library behind them (multimodal_lit.py:227-266, 445-511):
    training_step logs the reference's metric names incl. `temperature` and returns a loss that backpropagates into
    the head; the Labeled-S trial step (dataloader 1) scores one 4-way trial, target first."""
    from _util import case_inputs
    lit = _lit_model(cv)
    inp = case_inputs(77, 64, 512, "flat")
    with torch.no_grad():
        lit.model.image_embed.model.fc.weight.copy_(torch.from_numpy(inp["W"]))
        lit.model.image_embed.model.fc.bias.copy_(torch.from_numpy(inp["b"]))
        lit.model.text_embed.embedding.weight.copy_(torch.from_numpy(inp["table"]))
    logged = {}
    lit.log = lambda name, value, *a, **k: logged.__setitem__(name, float(value))
    batch = (torch.from_numpy(inp["f"]).to(DEV), torch.from_numpy(inp["ids"]).to(DEV), torch.from_numpy(inp["lens"]).to(DEV), None)
    lit.train()
    ret = lit.training_step(batch, 0)
    ret["loss"].backward()
    ref = O.contrastive_step(torch.from_numpy(inp["f"]), torch.from_numpy(inp["ids"]), torch.from_numpy(inp["lens"]),
                             torch.from_numpy(inp["W"]), torch.from_numpy(inp["b"]), torch.from_numpy(inp["table"]),
                             float(-np.log(0.07)))
    assert abs(ret["loss"].item() - ref["loss"].item()) <= 1e-3 * abs(ref["loss"].item())
    for name in ("train_infonce_loss", "train_image_accuracy", "train_text_accuracy", "train_image_entropy",
                 "train_text_entropy", "train_loss", "temperature"):
        assert name in logged, name
    assert logged["temperature"] == pytest.approx(0.07, rel=1e-5) and ret["batch_size"] == 64
    g = lit.model.image_embed.model.fc.weight.grad
    assert g is not None and float(g.abs().sum()) > 0 and lit.model.logit_neg_log_temperature.grad is not None
    # one trial: 4 candidate "images" (trunk-boundary rows, the pooled trunk) and one label; the matching row first
    lit.eval()
    logged.clear()
    with torch.no_grad():
        txt = lit.encode_text(batch[1][:1], batch[2][:1])
        imgs = lit.encode_image(batch[0][:4])
        want = int(torch.argmax(imgs @ txt[0]).item() == 0)
        # (the reference's x.view(-1, *x.shape[-3:]) flattens [1, n_way, C, H, W]; a pooled-trunk "image" is one
        # 2048-row, so the trial is handed over as [1, n_way, 1, 1, 2048] -> [n_way, 1, 1, 2048] -> rows)
        x = batch[0][:4].reshape(1, 4, 1, 1, 2048)
        out = lit.validation_step((x, batch[1][:1], batch[2][:1], [["ball"]]), 0, 1)
    assert out["accuracy"] == want
    assert set(logged) == {"val_accuracy", "val_entropy", "val_accuracy_ball"} and 0.0 <= logged["val_entropy"] <= np.log(4) + 1e-5


def test_evaluate_trials_equals_golden(cv):
    """batch_trials-style inputs -> MultiModalLitModel.evaluate_trials (own fp32 head + K7) == the predictions of the
    reference's per-trial loop (tests/golden/eval_4way_e512.npz), bit-exact argmax."""
    from _util import golden
    from test_oracle_golden import eval_case_inputs
    g = golden("eval_4way_e512")
    W, b, table, f = eval_case_inputs(g)
    lit = _lit_model(cv, fix_temperature=True)
    with torch.no_grad():
        lit.model.image_embed.model.fc.weight.copy_(torch.from_numpy(W))
        lit.model.image_embed.model.fc.bias.copy_(torch.from_numpy(b))
        lit.model.text_embed.embedding.weight.copy_(torch.from_numpy(table))
    feats = torch.from_numpy(f).to(DEV)                                  # [N, 4, 2048] trunk-boundary rows
    pred, logits = lit.evaluate_trials(feats, torch.from_numpy(g["ids"]).to(DEV), torch.from_numpy(g["lens"]).to(DEV))
    assert np.array_equal(pred.cpu().numpy(), g["pred"])
    np.testing.assert_allclose(logits.cpu().numpy(), g["logits"], atol=5e-5)


def test_classify_ncat_and_cosine_nearest_vs_oracle(cv):
    """the two other evaluation forms (22-category classification, cosine nearest-neighbour search): fp32 scores,
    first-maximum arg-max; chunked key sets give the same answer as one pass, ties go to the lowest index."""
    rng = np.random.RandomState(11)
    N, C, E = 3000, 22, 512
    img = torch.from_numpy(rng.standard_normal((N, E)).astype(np.float32))
    cat = torch.from_numpy(rng.standard_normal((C, E)).astype(np.float32))
    img[5] = 0.0                                                       # a zero row: all scores 0 -> category 0
    pred, logits = cv.ops.classify_ncat(img.to(DEV), cat.to(DEV), True, float(-np.log(0.07)))
    rp, rl = O.classify_ncat(img, cat, float(-np.log(0.07)))
    assert np.array_equal(pred.cpu().numpy(), rp.numpy().astype(np.int32)) and int(pred[5]) == 0
    np.testing.assert_allclose(logits.cpu().numpy(), rl.numpy(), atol=2e-5)
    # nearest neighbour: 700 queries against 5000 keys, in one pass and in chunks of 1024 / 999 keys
    q = torch.from_numpy(rng.standard_normal((700, E)).astype(np.float32))
    k = torch.from_numpy(rng.standard_normal((5000, E)).astype(np.float32))
    k[4000] = k[17]; k[1024] = k[17] * 3.0                              # exact cosine ties across chunks
    q[0] = k[17] * 0.5                                                 # its nearest key is 17 (first of the tied three)
    rb, ra = O.cosine_nearest(q, k)
    for chunk in (16384, 1024, 999):
        best, arg = cv.ops.cosine_nearest(q.to(DEV), k.to(DEV), chunk=chunk)
        got = arg.cpu().numpy()
        mism = np.nonzero(got != ra.numpy())[0]
        # the only admissible differences are fp32 near-ties of the oracle itself
        for i in mism:
            sims = (torch.nn.functional.normalize(q[i].double(), dim=0) @ torch.nn.functional.normalize(k.double(), dim=1).t())
            assert abs(float(sims[got[i]] - sims[ra[i]])) < 1e-6, (chunk, i)
        assert int(arg[0]) == 17, chunk
        np.testing.assert_allclose(best.cpu().numpy(), rb.numpy(), atol=2e-6)


def test_lit_classify_frames(cv):
    from _util import case_inputs
    lit = _lit_model(cv, fix_temperature=True)
    inp = case_inputs(78, 40, 512, "flat")
    with torch.no_grad():
        lit.model.image_embed.model.fc.weight.copy_(torch.from_numpy(inp["W"]))
        lit.model.image_embed.model.fc.bias.copy_(torch.from_numpy(inp["b"]))
        lit.model.text_embed.embedding.weight.copy_(torch.from_numpy(inp["table"]))
    ids, lens = torch.from_numpy(inp["ids"][:22]), torch.from_numpy(inp["lens"][:22])
    pred, logits = lit.classify_frames(torch.from_numpy(inp["f"]).to(DEV), ids.to(DEV), lens.to(DEV))
    lpi, _, _, _ = O.forward(torch.from_numpy(inp["f"]), ids, lens, torch.from_numpy(inp["W"]), torch.from_numpy(inp["b"]),
                             torch.from_numpy(inp["table"]), float(-np.log(0.07)))
    assert logits.shape == (40, 22)
    np.testing.assert_allclose(logits.cpu().numpy(), lpi.numpy(), atol=5e-5)
    assert np.array_equal(pred.cpu().numpy(), torch.argmax(lpi, dim=1).numpy().astype(np.int32))
