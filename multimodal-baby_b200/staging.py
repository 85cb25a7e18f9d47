"""Host -> device input staging for the contrastive step.

The reference collates a batch with `multiModalDataset_collate_fn`
(multimodal/multimodal_data_module.py:98-109): pad the token rows of the batch to the longest utterance,
truncate to MAX_LEN_UTTERANCE = 25, clamp the lengths.  The padded width therefore changes from batch to
batch, and every step pays pageable-host H2D copies of freshly allocated tensors.

`multiModalDataset_collate_fn` below returns exactly what the reference's function returns (drop-in for the
data module).  `PinnedBatchStager` is what the CUDA-graph train step wants instead: fixed-shape
[B, 25] int64 ids / [B] int64 lengths (and optionally the trunk-boundary features) in PINNED host buffers
that `GraphedContrastiveStep` copies from inside its graph; positions >= len hold PAD (0), which is the
invariant the text-encoder kernel relies on (SURVEY 8a note on implicit masking).
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
from torch.nn.utils.rnn import pad_sequence

MAX_LEN_UTTERANCE = 25          # multimodal_data_module.py:30
PAD_TOKEN_ID = 0                # multimodal_data_module.py:24


def multiModalDataset_collate_fn(batch):
    """multimodal_data_module.py:98-109, same outputs: (img [B,...], ids [B, min(max_len, 25)] int64,
    lengths [B] int64 clamped to 25, raw utterances list)."""
    img, utterance_idxs, utterance_length, raw_utterance = zip(*batch)
    img = torch.stack(img, 0)
    utterance_idxs = pad_sequence(utterance_idxs, batch_first=True, padding_value=PAD_TOKEN_ID)
    utterance_length = torch.tensor(utterance_length, dtype=torch.long)
    if utterance_idxs.size(1) > MAX_LEN_UTTERANCE:
        utterance_idxs = utterance_idxs[:, :MAX_LEN_UTTERANCE]
        utterance_length = torch.minimum(utterance_length, torch.tensor(MAX_LEN_UTTERANCE, dtype=torch.long))
    return img, utterance_idxs, utterance_length, list(raw_utterance)


def host_arena(nbytes: int, write_combined: bool = True) -> torch.Tensor:
    """uint8 [nbytes] page-locked host memory from the library (cvcl_host_alloc), WRITE-COMBINED by default: the
    staging buffer of the per-step H2D copy.  The CPU never caches write-combined lines, so the DMA engine does not
    snoop CPU caches (an ordinary pinned buffer whose lines sit in a CPU cache -- after a CPU read or a cached write
    -- copies 4x slower on this pool's hosts: ~200 us instead of 46 us for 2.2 MB).  Write it, do not read it (reads
    are uncached).  Freed when the tensor and all its views are gone."""
    import ctypes
    import weakref
    import numpy as np
    from . import _cabi
    lib = _cabi.load()
    ptr = lib.cvcl_host_alloc(int(nbytes), 1 if write_combined else 0)
    if not ptr:
        _cabi.check(-3)
    buf = (ctypes.c_uint8 * int(nbytes)).from_address(ptr)
    arr = np.frombuffer(buf, dtype=np.uint8)
    t = torch.from_numpy(arr)
    weakref.finalize(buf, lib.cvcl_host_free, ctypes.c_void_p(ptr))     # `buf` lives as long as a view of it does
    return t


def packed_buffers(specs, pin=True, device=None, align=256, write_combined=False):
    """[(shape, dtype), ...] -> zero-initialised tensors that are views of ONE byte arena (each segment `align`-byte
    aligned), pinned host memory by default or on `device`: a batch laid out like this crosses PCIe as one copy.
    write_combined=True: the arena comes from `host_arena` (CPU-write-only staging memory)."""
    offs, n = [], 0
    for shape, dtype in specs:
        n = (n + align - 1) // align * align
        offs.append(n)
        n += int(torch.Size(shape).numel()) * torch.empty((), dtype=dtype).element_size()
    if device is not None:
        arena = torch.zeros((n,), dtype=torch.uint8, device=device)
    elif pin and write_combined:
        arena = host_arena(n, True)
        arena.zero_()
    else:
        arena = torch.zeros((n,), dtype=torch.uint8)
        if pin:
            arena = arena.pin_memory()
    out = []
    for (shape, dtype), o in zip(specs, offs):
        nb = int(torch.Size(shape).numel()) * torch.empty((), dtype=dtype).element_size()
        out.append(arena[o:o + nb].view(dtype).view(tuple(shape)))
    return out


def packed_span(tensors):
    """If `tensors` are contiguous views of one storage, in ascending order and without big holes: (storage byte
    offset of the first, total span in bytes, per-tensor offsets relative to the span); else None."""
    try:
        base = tensors[0].untyped_storage().data_ptr()
        if any(t.untyped_storage().data_ptr() != base or not t.is_contiguous() for t in tensors):
            return None
        lo = [t.data_ptr() - base for t in tensors]
        hi = [o + t.numel() * t.element_size() for o, t in zip(lo, tensors)]
        if any(lo[i + 1] < hi[i] for i in range(len(lo) - 1)) or any((o - lo[0]) % 16 for o in lo):
            return None
        span = hi[-1] - lo[0]
        if span > sum(h - l for l, h in zip(lo, hi)) + 4096 * len(tensors):
            return None
        return lo[0], span, [o - lo[0] for o in lo]
    except (RuntimeError, IndexError):
        return None


class PinnedBatchStager:
    """Fixed-shape pinned staging buffers for `GraphedContrastiveStep(model, x_host, ids_host, lens_host)`.

        stager = PinnedBatchStager(batch_size=512, feat_shape=(2048,))
        step = GraphedContrastiveStep(model, stager.x_host, stager.ids_host, stager.lens_host, prefetch=True)
        for feats, token_rows, lengths in loader:
            stager.stage(token_rows, lengths, feats)      # host-side writes only, no allocation
            loss = step()

    `stage` accepts the per-sample token rows (a sequence of 1-D int64 tensors, as the dataset yields them)
    or an already padded [B, L'] tensor (as the reference's collate produces), pads / truncates to
    `max_len`, forces PAD beyond each length and clamps the lengths -- the same ids / lengths the reference's
    collate would hand to the model, in a layout that never changes shape."""

    def __init__(self, batch_size: int, feat_shape: Optional[Sequence[int]] = None,
                 feat_dtype: torch.dtype = torch.bfloat16, max_len: int = MAX_LEN_UTTERANCE,
                 pin: Optional[bool] = None):
        pin = torch.cuda.is_available() if pin is None else bool(pin)

        self.batch_size, self.max_len = int(batch_size), int(max_len)
        # ONE pinned arena [x | ids | lens]: GraphedContrastiveStep recognises views of one storage and moves the
        # whole batch with a single H2D copy (three separate copies cost ~4.5 us of latency each on top of the bytes)
        specs = [((batch_size,) + tuple(feat_shape), feat_dtype)] if feat_shape is not None else []
        specs += [((batch_size, max_len), torch.int64), ((batch_size,), torch.int64)]
        views = packed_buffers(specs, pin=pin)
        self.x_host = views[0] if feat_shape is not None else None
        self.ids_host, self.lens_host = views[-2], views[-1]

    @torch.no_grad()
    def stage(self, utterance_idxs, utterance_length, feats: Optional[torch.Tensor] = None):
        B, L = self.batch_size, self.max_len
        lens = torch.as_tensor(utterance_length, dtype=torch.int64).reshape(-1)
        if lens.numel() != B:
            raise ValueError("PinnedBatchStager: got %d utterances for a batch of %d" % (lens.numel(), B))
        self.ids_host.fill_(PAD_TOKEN_ID)
        if torch.is_tensor(utterance_idxs) and utterance_idxs.dim() == 2:
            if utterance_idxs.shape[0] != B:
                raise ValueError("PinnedBatchStager: padded ids have %d rows, batch is %d" % (utterance_idxs.shape[0], B))
            w = min(L, utterance_idxs.shape[1])
            self.ids_host[:, :w].copy_(utterance_idxs[:, :w])
        else:
            if len(utterance_idxs) != B:
                raise ValueError("PinnedBatchStager: got %d token rows for a batch of %d" % (len(utterance_idxs), B))
            for i, row in enumerate(utterance_idxs):
                row = torch.as_tensor(row, dtype=torch.int64).reshape(-1)
                w = min(L, row.numel())
                self.ids_host[i, :w].copy_(row[:w])
        torch.clamp(lens, max=L, out=self.lens_host)
        # the invariant the kernels rely on: every position >= len holds PAD
        pos = torch.arange(L, dtype=torch.int64)[None, :]
        self.ids_host.masked_fill_(pos >= self.lens_host[:, None], PAD_TOKEN_ID)
        if feats is not None:
            if self.x_host is None:
                raise ValueError("PinnedBatchStager was built without a feature buffer")
            self.x_host.copy_(feats)              # casts (e.g. fp32 -> bf16) on the host
        return self.ids_host, self.lens_host


def batch_trials(items, max_len: Optional[int] = None):
    """Host front-end of the batched Labeled-S evaluation.  `items` are what the reference's
    `LabeledSEvalDataset.__getitem__` returns (multimodal_data_module.py:124-156): (imgs [n_way,3,H,W] with the
    target first, label row [L] int64, label_len, [raw_label]).  The reference feeds them to the model one trial at
    a time (eval.py:196-214, multimodal_lit.py:466-511: 2200 trunk passes of 4 images and 2200 text encodings of
    22 distinct labels).  Here the frames are stacked for ONE trunk pass and identical label rows are shared:

        frames [N*n_way, 3, H, W], label_ids [C, L] int64 (zero padded), label_lens [C] int64,
        label_index [N] int32 (row of label_ids per trial), raw_labels [N]

    which is the input of `MultiModalLitModel.evaluate_trials` (`ops.eval_nway`, K7)."""
    if len(items) == 0:
        raise ValueError("batch_trials: no trials")
    n_way = items[0][0].shape[0]
    rows, lens, index, raw, seen = [], [], [], [], {}
    for imgs, label, label_len, raw_label in items:
        if imgs.shape[0] != n_way:
            raise ValueError("batch_trials: trials with %d and %d candidates cannot share a batch" % (n_way, imgs.shape[0]))
        label = torch.as_tensor(label, dtype=torch.int64).reshape(-1)
        n = int(label_len)
        if n < 1 or n > label.numel():
            raise ValueError("batch_trials: label length %d does not fit a row of %d ids" % (n, label.numel()))
        key = tuple(label[:n].tolist())
        if key not in seen:
            seen[key] = len(rows)
            rows.append(label[:n])
            lens.append(n)
        index.append(seen[key])
        raw.append(raw_label[0] if isinstance(raw_label, (list, tuple)) else raw_label)
    L = max(lens) if max_len is None else int(max_len)
    if max(lens) > L:
        raise ValueError("batch_trials: a label of %d ids does not fit max_len=%d" % (max(lens), L))
    label_ids = torch.full((len(rows), L), PAD_TOKEN_ID, dtype=torch.int64)
    for i, r in enumerate(rows):
        label_ids[i, :r.numel()] = r
    frames = torch.cat([it[0] for it in items], dim=0)
    return (frames, label_ids, torch.tensor(lens, dtype=torch.int64), torch.tensor(index, dtype=torch.int32), raw)
