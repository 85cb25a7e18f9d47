"""CUDA-graph train step for the flat contrastive path: H2D staging, the fused step and the D2H read
of the loss are captured ONCE and replayed per batch (B200-first: graphs instead of per-op Python
dispatch; the captured body is exactly `ops.flat_contrastive_step` / `ops.flat_step_sharded`, i.e.
MultiModalModel.calculate_contrastive_loss + backward, multimodal.py:796-822).

    step = GraphedContrastiveStep(model, x_host, ids_host, lens_host)   # pinned host staging buffers
    loss = step()            # copies the staged batch, runs fwd+bwd, returns the loss (python float)
    # gradients are in p.grad of the head parameters (static views of one flat buffer)

The loader writes the next batch into `step.x_host / ids_host / lens_host` (pinned) between calls.
"""
from __future__ import annotations

import torch

from . import ops


class GraphedContrastiveStep:
    def __init__(self, model, x_host, ids_host, lens_host, warmup=3):
        if model.embedding_type != "flat":
            raise NotImplementedError("GraphedContrastiveStep covers the flat-embedding train step")
        for t in (x_host, ids_host, lens_host):
            if t.is_cuda or not t.is_pinned():
                raise ValueError("staging buffers must be pinned host tensors")
        self.model = model
        self.group = model.process_group
        self.x_host, self.ids_host, self.lens_host = x_host, ids_host, lens_host
        w, b = model._head()
        table = model.text_embed.embedding.weight
        dev = table.device
        self.dev = dev
        self.x = torch.empty_like(x_host, device=dev)
        self.ids = torch.empty_like(ids_host, device=dev)
        self.lens = torch.empty_like(lens_host, device=dev)
        self.stats_host = torch.zeros(8, dtype=torch.float32).pin_memory()
        s = model.logit_neg_log_temperature
        if isinstance(s, torch.nn.Parameter):
            raise NotImplementedError("graphed step needs fix_temperature=True (s is baked into the graph)")
        ls = ops._scalar(s)
        norm = bool(model.normalize_features)
        E, K, V = table.shape[1], w.shape[1], table.shape[0]

        @torch.no_grad()
        def body():
            self.x.copy_(self.x_host, non_blocking=True)
            self.ids.copy_(self.ids_host, non_blocking=True)
            self.lens.copy_(self.lens_host, non_blocking=True)
            if self.group is None:
                out5, _, _, flat = ops.flat_contrastive_step(self.x, self.ids, self.lens, w, b, table, ls,
                                                             norm, True, False)
            else:
                stats, _, _ = ops.flat_step_sharded(self.x, self.ids, self.lens, w, b, table, ls, norm,
                                                    True, False, self.group)
                out5, flat = stats[:8], stats[8:]
            self.stats_host.copy_(out5, non_blocking=True)
            return flat

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.flat = body()
        ds, db, dtable, dW = ops.split_flat_grads(self.flat, E, K, V)
        # static gradient views: replay overwrites them in place
        w_param, b_param = self._head_params()
        w_param.grad = dW.view_as(w_param)
        if b_param is not None:
            b_param.grad = db
        table.grad = dtable

    def _head_params(self):
        m = self.model.image_embed.model
        return m.fc.weight, m.fc.bias

    def __call__(self):
        self.graph.replay()
        torch.cuda.current_stream(self.dev).synchronize()
        return float(self.stats_host[0])

    def stats(self):
        """(loss, image_accuracy, text_accuracy, image_entropy, text_entropy) of the last step."""
        return tuple(float(v) for v in self.stats_host[:5])
