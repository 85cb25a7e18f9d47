"""-m gpu: Grad-CAM attention maps -- `ops.gradcam_flat` / `ops.bicubic_upsample` and the
`attention_maps.gradCAM` drop-in against the maps of the unmodified reference (tests/golden/gradcam_*.npz,
multimodal/attention_maps.py:111-165) and against the CPU oracle on other shapes.  fp32 end to end:
tolerance 2e-5 of the largest map value (summation order only)."""
import collections

import numpy as np
import pytest
import torch

from _util import O, golden, t

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 2e-5


@pytest.fixture(scope="module")
def cv():
    import multimodal_baby_b200 as m
    return m


def test_gradcam_matches_reference_golden(cv):
    from oracle.make_golden import gradcam_inputs
    g = golden("gradcam_e512_n3")
    inp = gradcam_inputs(int(g["seed"]), int(g["N"]), int(g["E"]))
    for norm, key in ((True, "norm"), (False, "raw")):
        cam = cv.ops.gradcam_flat(t(inp["act"], DEV), t(inp["W"], DEV), t(inp["b"], DEV), t(inp["target"], DEV), norm)
        scale = float(np.abs(g["cam_" + key]).max())
        assert np.abs(cam.cpu().numpy() - g["cam_" + key]).max() <= TOL * scale
        big = cv.ops.bicubic_upsample(cam, 224, 224)
        assert tuple(big.shape) == (int(g["N"]), 1, 224, 224)
        assert np.abs(big.cpu().numpy()[:, :, ::3, ::3] - g["resized_" + key]).max() <= TOL * scale


@pytest.mark.parametrize("N,K,E,H,W,norm", [(5, 2048, 512, 7, 7, True), (1, 256, 64, 7, 7, True), (4, 512, 128, 8, 8, False),
                                            (2, 64, 32, 3, 5, True)])
def test_gradcam_matches_oracle(cv, N, K, E, H, W, norm):
    rng = np.random.RandomState(N * 1000 + K)
    act = np.maximum(rng.standard_normal((N, K, H, W)), 0).astype(np.float32)
    Wt = (rng.standard_normal((E, K)) / np.sqrt(K)).astype(np.float32)
    b = (0.1 * rng.standard_normal(E)).astype(np.float32)
    tg = rng.standard_normal((N, E)).astype(np.float32)
    ref, ref_big = O.gradcam_flat(t(act), t(Wt), t(b), t(tg), norm, (37, 53))
    cam = cv.ops.gradcam_flat(t(act, DEV), t(Wt, DEV), t(b, DEV), t(tg, DEV), norm)
    scale = float(ref.abs().max())
    assert scale > 0
    assert float((cam.cpu() - ref).abs().max()) <= TOL * scale
    big = cv.ops.bicubic_upsample(cam, 37, 53)                   # non-square, non-integer scale factors
    assert float((big.cpu() - ref_big).abs().max()) <= TOL * scale


def test_bicubic_matches_torch_on_random_maps(cv):
    """up- and down-scaling of signed maps against F.interpolate on the CPU (the reference's call)."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn(3, 1, 7, 7, generator=g)
    for hw in ((224, 224), (7, 7), (5, 3), (13, 29)):
        ref = torch.nn.functional.interpolate(x, hw, mode="bicubic", align_corners=False)
        got = cv.ops.bicubic_upsample(x.to(DEV), hw[0], hw[1]).cpu()
        assert float((got - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


def test_gradcam_dropin_signature(cv):
    """attention_maps.gradCAM(model, input, target, layer, normalize_features, resize) on a ResNet-shaped module:
    == the oracle on the activation the saliency layer produced; unsupported layers are loud."""
    torch.manual_seed(0)
    model = torch.nn.Sequential(collections.OrderedDict(
        conv=torch.nn.Conv2d(3, 64, 3, stride=4, padding=1), relu=torch.nn.ReLU(),
        layer4=torch.nn.Conv2d(64, 128, 3, stride=2, padding=1),
        avgpool=torch.nn.AdaptiveAvgPool2d((1, 1)), flatten=torch.nn.Flatten(1), fc=torch.nn.Linear(128, 32))).to(DEV)
    x = torch.randn(2, 3, 56, 56, device=DEV)
    tgt = torch.nn.functional.normalize(torch.randn(2, 32, device=DEV), dim=1)
    got = cv.attention_maps.gradCAM(model, x, tgt, model.layer4, normalize_features=True)
    assert tuple(got.shape) == (2, 1, 56, 56)
    with torch.no_grad():
        act = model.layer4(model.relu(model.conv(x)))
    ref, ref_big = O.gradcam_flat(act.cpu(), model.fc.weight.detach().cpu(), model.fc.bias.detach().cpu(), tgt.cpu(),
                                  True, (56, 56))
    assert float((got.cpu() - ref_big).abs().max()) <= 5e-5 * float(ref_big.abs().max())
    small = cv.attention_maps.gradCAM(model, x, tgt, model.layer4, normalize_features=True, resize=False)
    assert float((small.cpu() - ref).abs().max()) <= 5e-5 * float(ref.abs().max())
    with pytest.raises(NotImplementedError):
        cv.attention_maps.gradCAM(model, x, tgt, model.conv)
