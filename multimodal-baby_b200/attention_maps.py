"""Grad-CAM attention maps with the reference's call signature (multimodal/attention_maps.py:111-165;
callers: eval.py:206, eval_shuffled.py:206, analysis_cvcl/generate_attention_maps.py:104).

    gradCAM(model.vision_encoder.model, image, text_features, model.vision_encoder.model.layer4,
            normalize_features=model.model.normalize_features)        -> [N, 1, H_in, W_in]

The reference runs the trunk forward AND backward under autograd to obtain d output / d layer4.  For the
saliency layer it uses (layer4, followed only by the global average pool and `fc`) that gradient has a
closed form -- the head is linear in the pooled activation -- so here the trunk (the reference's torch
module, outside the product) runs forward once without autograd and `ops.gradcam_flat` /
`ops.bicubic_upsample` (hand-written CUDA, fp32) do the rest; images are processed as one batch instead of
one call per image.  Other saliency layers / non-linear heads raise NotImplementedError (no fallback).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


def gradCAM_with_act_and_grad(act: torch.Tensor, grad: torch.Tensor) -> torch.Tensor:
    """attention_maps.py:111-122 for an explicit gradient: alpha = grad.mean((2,3)), clamp(sum_c act*alpha, 0).
    Plain torch (used by callers that already hold a gradient, analysis_tools/multimodal_visualization.py:39)."""
    alpha = grad.mean(dim=(2, 3), keepdim=True)
    return torch.clamp(torch.sum(act * alpha, dim=1, keepdim=True), min=0)


@torch.no_grad()
def gradCAM(model: nn.Module, input: torch.Tensor, target: torch.Tensor, layer: nn.Module,
            normalize_features: bool = False, resize: bool = True) -> torch.Tensor:
    """Same arguments and result as the reference's gradCAM.  `model` is the vision trunk + head
    (torchvision ResNet layout: ... layer4 -> avgpool -> fc), `layer` must be its `layer4`."""
    if not isinstance(layer, nn.Module):
        raise TypeError("layer must be an nn.Module")
    fc = getattr(model, "fc", None)
    if layer is not getattr(model, "layer4", None) or not isinstance(fc, nn.Linear):
        raise NotImplementedError("gradCAM: only the reference's configuration is implemented "
                                  "(saliency layer = model.layer4 followed by avgpool and a Linear `fc`)")
    grabbed = {}
    hook = layer.register_forward_hook(lambda mod, inp, out: grabbed.__setitem__("act", out))
    try:
        model(input)                                   # trunk forward only; its fc output is not used
    finally:
        hook.remove()
    act = grabbed["act"].float()
    cam = ops.gradcam_flat(act, fc.weight, fc.bias, target.to(act.device), bool(normalize_features))
    if resize:
        cam = ops.bicubic_upsample(cam, int(input.shape[2]), int(input.shape[3]))
    return cam
