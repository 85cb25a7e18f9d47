"""CUDA-graph train step for the flat contrastive path: H2D staging, the fused step, (optionally) the
optimizer update and the D2H read of the loss are captured ONCE and replayed per batch (B200-first: graphs
instead of per-op Python dispatch; the captured body is exactly `ops.flat_contrastive_step` /
`ops.flat_step_sharded`, i.e. MultiModalModel.calculate_contrastive_loss + backward, multimodal.py:796-822).

    step = GraphedContrastiveStep(model, x_host, ids_host, lens_host)   # pinned host staging buffers
    loss = step()            # copies the staged batch, runs fwd+bwd, returns the loss (python float)
    # gradients are in p.grad of the head parameters (static views of one flat buffer)

The loader writes the next batch into `step.x_host / ids_host / lens_host` (pinned) between calls.

`optimizer=FusedAdamW(...)` puts the parameter update (and the refresh of the bf16 weight shadow the head
GEMM reads) into the same graph: one replay = one complete training step, and because replays are
stream-ordered every step sees the weights the previous step produced.

`prefetch=True` double-buffers the device inputs: replay k computes on the batch that replay k-1
copied while it was computing, and copies the batch now staged in the pinned buffers for replay k+1
(H2D overlaps the kernels on a second graph branch).  Every call still moves one full batch host ->
device; the returned loss belongs to the batch staged one call earlier (call `prime()` once after
staging the first batch).

`own_staging=True` (packed staging only): the step allocates all of its pinned staging sets itself -- a few
candidates are timed and the fastest kept -- and the caller's tensors only provide the initial contents.

`lagged_loss=True` (needs prefetch) additionally keeps one replay in flight: a call enqueues replay k and
waits only for replay k-1, whose loss it returns (each graph writes its own pinned stats buffer).  The
host never idles behind the GPU; every step's loss is still read back, one call later.  `flush()` waits
for the replay in flight and returns its loss.  Two consequences, both handled here:
  * the H2D copies of the replay still in flight read the pinned staging buffers, so the staging buffers
    are DOUBLE-BUFFERED too: `x_host / ids_host / lens_host` always name the set the next call will copy,
    which no replay in flight is reading (the set alternates with every call);
  * without `optimizer=`, an optimizer step issued by the caller after call k lands behind replay k in the
    stream, so replay k computed its gradients with the weights of step k-2: gradients are one step stale
    (asynchronous-SGD semantics).  Pass `optimizer=` to keep the update inside the graph and exact.
"""
from __future__ import annotations

import torch

from . import ops


class GraphedContrastiveStep:
    def __init__(self, model, x_host, ids_host, lens_host, warmup=3, prefetch=False, lagged_loss=False,
                 optimizer=None, own_staging=False):
        if model.embedding_type != "flat":
            raise NotImplementedError("GraphedContrastiveStep covers the flat-embedding train step")
        for t in (x_host, ids_host, lens_host):
            if t.is_cuda or not t.is_pinned():
                raise ValueError("staging buffers must be pinned host tensors")
        self.model = model
        self.group = model.process_group
        w, b = model._head()
        table = model.text_embed.embedding.weight
        dev = table.device
        self.dev = dev
        self.prefetch = bool(prefetch)
        nbuf = 2 if self.prefetch else 1
        if lagged_loss and not prefetch:
            raise ValueError("lagged_loss=True needs prefetch=True (two alternating graphs)")
        self.lagged = bool(lagged_loss)
        # Staging buffers that are views of ONE pinned arena (staging.PinnedBatchStager / staging.packed_buffers)
        # cross PCIe as a single copy per step; separately allocated tensors take three (each small copy adds
        # ~4.5 us of latency: measured 53.7 us vs 44.5 us for the 2 MB of features alone).
        from . import staging
        span = staging.packed_span((x_host, ids_host, lens_host))

        def byte_view(t, lo, n):
            return torch.empty(0, dtype=torch.uint8, device=t.device).set_(t.untyped_storage(), lo, (n,))

        def carve(arena, like, off):
            nb = like.numel() * like.element_size()
            return arena[off:off + nb].view(like.dtype).view(like.shape)

        def new_set(device=None):
            """a set of (x, ids, lens) buffers shaped like the caller's, on `device` or pinned; packed when the
            caller's are.  -> (views, arena | None)"""
            if span is None:
                ts = tuple(torch.empty_like(t, device=device) if device is not None else torch.empty_like(t).pin_memory()
                           for t in (x_host, ids_host, lens_host))
                return ts, None
            arena = torch.empty((span[1],), dtype=torch.uint8, device=device) if device is not None else \
                torch.empty((span[1],), dtype=torch.uint8).pin_memory()
            return tuple(carve(arena, t, o) for t, o in zip((x_host, ids_host, lens_host), span[2])), arena

        self.bufs, self._dev_arenas = [], []
        for _ in range(nbuf):
            ds_, da = new_set(dev)
            self.bufs.append(ds_); self._dev_arenas.append(da)
        self.packed = span is not None
        # pinned staging: one set per graph in lagged mode (the replay in flight may still be reading its set)
        n_sets = 2 if self.lagged else 1
        self.staging_probe_us = None
        if own_staging and span is not None:
            # Every staging set is allocated HERE (the caller's tensors only provide the initial contents; write the
            # next batch into `step.x_host / ids_host / lens_host`).  On the hosts this was developed on, pinned
            # allocations come in a fast and a slow flavour (2.2 MB in 46 us or in 70-200 us, profiles/
            # r02_h2d_staging_probe.txt), so a few candidates are timed and the fastest ones kept.
            cands = [new_set() for _ in range(n_sets + 3)]
            times = []
            for hs, ha in cands:
                for dst, src in zip(hs, (x_host, ids_host, lens_host)):
                    dst.copy_(src)
                ts = []
                for _ in range(9):
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    e0.record(); self._dev_arenas[0].copy_(ha, non_blocking=True); e1.record(); e1.synchronize()
                    ts.append(e0.elapsed_time(e1) * 1e3)
                times.append(sorted(ts)[len(ts) // 2])
            order = sorted(range(len(cands)), key=lambda i: times[i])[:n_sets]
            self.host_sets = [cands[i][0] for i in order]
            self._host_arenas = [cands[i][1] for i in order]
            self.staging_probe_us = {"candidates": [round(v, 1) for v in times], "kept": [round(times[i], 1) for i in order]}
        else:
            self.host_sets = [(x_host, ids_host, lens_host)]
            self._host_arenas = [byte_view(x_host, span[0], span[1]) if span is not None else None]
            if self.lagged:
                hs, ha = new_set()
                self.host_sets.append(hs); self._host_arenas.append(ha)
                for dst, src in zip(self.host_sets[1], self.host_sets[0]):
                    dst.copy_(src)
        self.x, self.ids, self.lens = self.bufs[0]
        self.stats_bufs = [torch.zeros(8, dtype=torch.float32).pin_memory() for _ in range(nbuf)]
        self.stats_host = self.stats_bufs[0]
        self.events = [torch.cuda.Event() for _ in range(nbuf)]
        self.pending = None
        self.copy_stream = torch.cuda.Stream(device=dev) if self.prefetch else None
        self.calls = 0
        self.optimizer = optimizer
        s = model.logit_neg_log_temperature
        self.s_param = s if isinstance(s, torch.nn.Parameter) else None
        if self.s_param is not None and (self.group is not None or not s.is_cuda):
            raise NotImplementedError("a trainable temperature in the graphed step needs the single-GPU "
                                      "one-kernel path with the parameter on the device")
        ls = 0.0 if self.s_param is not None else ops._scalar(s)
        s_dev = self.s_param.detach().reshape(1) if self.s_param is not None else None
        norm = bool(model.normalize_features)
        E, K, V = table.shape[1], w.shape[1], table.shape[0]
        if self.s_param is not None and not ops.fused_supported(x_host.shape[0], ids_host.shape[1], E, K, V):
            raise NotImplementedError("trainable temperature: shape not covered by the one-kernel step")
        if optimizer is not None:
            optimizer.attach_shadow(self._head_params()[0])      # the update kernel keeps bf16(W) current
        # with the update inside the graph every graph binds the .grad views of ITS gradient buffer while it
        # is captured (the optimizer launch bakes those pointers in); nothing outside reads older gradients
        self._opt_in_graph = optimizer is not None

        def h2d(buf, host):
            if self.packed:                                # the whole batch in one copy
                kb = next(i for i, b_ in enumerate(self.bufs) if b_ is buf)
                kh = next(i for i, h_ in enumerate(self.host_sets) if h_ is host)
                self._dev_arenas[kb].copy_(self._host_arenas[kh], non_blocking=True)
                return
            buf[0].copy_(host[0], non_blocking=True)
            buf[1].copy_(host[1], non_blocking=True)
            buf[2].copy_(host[2], non_blocking=True)

        @torch.no_grad()
        def body(cur, nxt, stats_host, slot, host):
            main = torch.cuda.current_stream(dev)
            if nxt is None:
                h2d(cur, host)                             # copy, then compute (serial)
            else:
                self.copy_stream.wait_stream(main)         # copy the NEXT batch beside the kernels
                with torch.cuda.stream(self.copy_stream):
                    h2d(nxt, host)
            x_d, ids_d, lens_d = cur
            if self.group is None:
                out5, _, _, flat = ops.flat_contrastive_step(x_d, ids_d, lens_d, w, b, table, ls, norm, True, False,
                                                             s_dev)
            else:
                stats, _, _ = ops.flat_step_sharded(x_d, ids_d, lens_d, w, b, table, ls, norm, True, False,
                                                    self.group, stats_slot=slot)
                out5, flat = stats[:8], stats[8:]
            if self._opt_in_graph:
                self._bind_views(flat, E, K, V, table)
                self.optimizer.graph_step()
            stats_host.copy_(out5, non_blocking=True)
            if nxt is not None:
                main.wait_stream(self.copy_stream)
            return flat

        pairs = [(self.bufs[0], None, self.stats_bufs[0])] if not self.prefetch else [
            (self.bufs[0], self.bufs[1], self.stats_bufs[0]), (self.bufs[1], self.bufs[0], self.stats_bufs[1])]
        self._dims = (E, K, V, table)
        backup = None
        if optimizer is not None:            # the warm-up iterations run real updates: undone below
            backup = [(p, p.detach().clone()) for g in optimizer.param_groups for p in g["params"]]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                for k, (cur, nxt, sh) in enumerate(pairs):
                    body(cur, nxt, sh, k, self.host_sets[k % len(self.host_sets)])
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        if optimizer is not None:
            self._reset_optimizer_after_warmup(backup)
        self._body, self._pairs, self._new_set = body, pairs, new_set
        self.graphs, self.flats, self._grad_views = [None] * len(pairs), [None] * len(pairs), {}
        for k in range(len(pairs)):
            self._capture(k)
        self._bind_grads(0, E, K, V, table)

    def _capture(self, k):
        cur, nxt, sh = self._pairs[k]
        gph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gph):
            flat = self._body(cur, nxt, sh, k, self.host_sets[k % len(self.host_sets)])
        self.graphs[k], self.flats[k] = gph, flat
        self._grad_views.pop(k, None)
        if k == 0:
            self.graph, self.flat = gph, flat

    def _time_set(self, host_arena, n=7):
        if getattr(self, "_probe_dev", None) is None:      # scratch target: the input buffers hold staged batches
            self._probe_dev = torch.empty_like(host_arena, device=self.dev)
        ts = []
        for _ in range(n):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); self._probe_dev.copy_(host_arena, non_blocking=True); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        return sorted(ts)[len(ts) // 2]

    def check_staging(self, slow_factor=1.4):
        """own_staging=True only.  Re-time the H2D copy of every staging set; a set that has become slow (a pinned
        allocation can drop from 46 us to 70-200 us per 2.2 MB on some hosts, profiles/r02_h2d_staging_probe.txt) is
        replaced by a freshly allocated fast one (contents preserved) and the graphs that copy from it are captured
        again.  Waits for the replay in flight; no step is executed.  -> number of sets replaced.  A training loop may
        call it every few thousand steps."""
        if not self.packed or self.staging_probe_us is None:
            return 0
        self.flush()
        torch.cuda.synchronize(self.dev)
        ref = min(self.staging_probe_us["kept"])
        replaced = 0
        for j in range(len(self.host_sets)):
            t = self._time_set(self._host_arenas[j])
            if t <= slow_factor * ref:
                continue
            best = None
            for _ in range(4):
                hs, ha = self._new_set()
                ha.copy_(self._host_arenas[j])
                tt = self._time_set(ha)
                if best is None or tt < best[0]:
                    best = (tt, hs, ha)
                if tt <= slow_factor * ref:
                    break
            if best[0] < t:
                self.host_sets[j], self._host_arenas[j] = best[1], best[2]
                replaced += 1
                for k in range(len(self.graphs)):
                    if k % len(self.host_sets) == j:
                        self._capture(k)
        if replaced:
            torch.cuda.synchronize(self.dev)
            self._bind_grads(0, *self._dims)
        return replaced

    # the warm-up iterations ran real optimizer steps on whatever was staged: undo them
    def _reset_optimizer_after_warmup(self, backup):
        opt = self.optimizer
        with torch.no_grad():
            for p, v in backup:
                p.copy_(v)
        opt.reset_state()
        torch.cuda.synchronize(self.dev)

    def _bind_views(self, flat, E, K, V, table):
        ds, db, dtable, dW = ops.split_flat_grads(flat, E, K, V)
        w_param, b_param = self._head_params()
        w_param.grad = dW.view_as(w_param)
        if b_param is not None:
            b_param.grad = db
        table.grad = dtable
        if self.s_param is not None:
            self.s_param.grad = ds.view(())

    def _bind_grads(self, which, E, K, V, table):
        if self._opt_in_graph:
            return
        views = self._grad_views.get(which)
        if views is None:
            ds, db, dtable, dW = ops.split_flat_grads(self.flats[which], E, K, V)
            w_param, b_param = self._head_params()
            views = self._grad_views[which] = (w_param, dW.view_as(w_param), b_param, db, table, dtable, ds.view(()))
        w_param, dW, b_param, db, table, dtable, ds = views
        w_param.grad = dW
        if b_param is not None:
            b_param.grad = db
        table.grad = dtable
        if self.s_param is not None:
            self.s_param.grad = ds

    # pinned staging buffers the NEXT call will copy (never read by a replay in flight)
    @property
    def x_host(self):
        return self.host_sets[self.calls % len(self.host_sets)][0]

    @property
    def ids_host(self):
        return self.host_sets[self.calls % len(self.host_sets)][1]

    @property
    def lens_host(self):
        return self.host_sets[self.calls % len(self.host_sets)][2]

    def prime(self):
        """prefetch mode: copy the batch currently staged in the pinned buffers into device buffer 0."""
        k = self.calls % len(self.host_sets)
        if self.packed:
            self._dev_arenas[0].copy_(self._host_arenas[k], non_blocking=True)
        else:
            for d, h in zip(self.bufs[0], self.host_sets[k]):
                d.copy_(h, non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        self.calls = 0

    def _head_params(self):
        m = self.model.image_embed.model
        return m.fc.weight, m.fc.bias

    def __call__(self):
        which = self.calls % len(self.graphs)
        self.graphs[which].replay()
        self.calls += 1
        if not self.lagged:
            torch.cuda.current_stream(self.dev).synchronize()
            if len(self.graphs) > 1:
                self._bind_grads(which, *self._dims)       # .grad views of the buffer just written
            self.stats_host = self.stats_bufs[which]
            return float(self.stats_host[0])
        # one replay stays in flight: wait for the previous one only and return ITS loss
        self.events[which].record(torch.cuda.current_stream(self.dev))
        prev, self.pending = self.pending, which
        if prev is None:
            return float("nan")                            # nothing finished yet (first call)
        self.events[prev].synchronize()
        self._bind_grads(prev, *self._dims)
        self.stats_host = self.stats_bufs[prev]
        return float(self.stats_host[0])

    def flush(self):
        """lagged mode: wait for the replay still in flight, bind its gradients, return its loss."""
        if self.pending is None:
            return float(self.stats_host[0])
        prev, self.pending = self.pending, None
        self.events[prev].synchronize()
        self._bind_grads(prev, *self._dims)
        self.stats_host = self.stats_bufs[prev]
        return float(self.stats_host[0])

    def stats(self):
        """(loss, image_accuracy, text_accuracy, image_entropy, text_entropy) of the last step."""
        return tuple(float(v) for v in self.stats_host[:5])


class GraphedLossStep:
    """Whole-step CUDA graph of an arbitrary loss closure built from this package's ops: `loss_fn()` (forward through
    the autograd wrappers of ops.py) and `loss.backward()` are captured once and replayed per batch.  This is how the
    spatial paths (multimodal.py:757-780: `sim = mean` and `sim = max` over the 7x7 map), which run as a sequence of
    eight to twelve kernels, get rid of their per-op host cost (B200-first: a graph, not Python dispatch; spatial mean
    at 1024 pairs: 470 us eager -> see profiles/README.md).

        x = torch.empty(B, 49, E, device=dev); ...                 # static input buffers, refilled between calls
        step = GraphedLossStep(lambda: loss_of(x, ids, lens), params=[table, x_leaf, ...])
        loss = step()            # replays forward + backward; returns the static 0-dim loss tensor (no sync)
        # gradients: p.grad of every tensor in `params` (static buffers, overwritten by the next call)

    `loss_fn` must not synchronise (no .item(), no host-side reads) and must read its inputs from fixed buffers.
    Every tensor in `params` must be a leaf that requires grad and is used by `loss_fn`."""

    def __init__(self, loss_fn, params, warmup=3):
        self.params = list(params)
        if not self.params:
            raise ValueError("GraphedLossStep: no parameters")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("GraphedLossStep runs on CUDA tensors only")
        self.dev = dev
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                for p in self.params:
                    p.grad = None
                loss_fn().backward()
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        for p in self.params:
            p.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = loss_fn()
            self.loss.backward()
        self.grads = [p.grad for p in self.params]
        if any(g is None for g in self.grads):
            raise RuntimeError("GraphedLossStep: a tensor in `params` received no gradient from loss_fn")

    def __call__(self):
        self.graph.replay()
        for p, g in zip(self.params, self.grads):      # the static gradient buffers (in case the caller reset .grad)
            p.grad = g
        return self.loss
