#!/usr/bin/env python
"""bench.py -- contrastive fwd+bwd pairs/s of the CVCL hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload ...]

Workload (default `flat512`, BASELINE.json configs[1]): CVCL flat-embedding contrastive train
step, 512 synthetic pairs per GPU (max utterance length 25, E=512, K=2048, V=2350), bf16 operands
with fp32 accumulation.  A "step" = text encoder + projection head + similarity/InfoNCE forward
and the full backward (dW, db, d table, d s) for one batch.  With N > 1 GPUs every rank holds
512 pairs and the InfoNCE runs over the global batch of 512*N pairs (NCCL all-gather of the
features): weak scaling.

One JSON line on stdout (rank 0):
  value        pairs/s, inputs resident in HBM, CUDA-event timed per step, L2 flushed between steps
  e2e          pairs/s through MultiModalModel.calculate_contrastive_loss(...) + backward() with
               pinned HOST inputs (H2D inside the timed region) and a D2H read of the loss
  roofline     the step kernel (one persistent kernel = the whole step at N = 1): algorithmic bytes and
               flops / the same CUDA-event duration the value is computed from, vs MEASURED_PEAKS.json;
               `phases` = in-kernel globaltimer stamps of its six phases
  cpu_baseline the oracle port (same ATen fp32 op sequence as the reference) on the host cores

`--impl reference` times that CPU path alone (rank 0 only) and prints the same line shape.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

def _baseline_metric():
    try:
        with open(os.path.join(ROOT, "BASELINE.json")) as fh:
            return json.load(fh)["metric"]
    except (OSError, KeyError, ValueError):
        return "contrastive fwd+bwd pairs/sec at 1/2/4/8 B200; % HBM / tensor-pipe roofline"


METRIC = _baseline_metric()          # BASELINE.json's metric; `value` is its pairs/sec part
E, K, V, L = 512, 2048, 2350, 25
S_FIXED = float(-np.log(0.07))


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"],
                    bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


# ----------------------------------------------------------------------------------------------
def synth_batch(seed, B):
    from oracle import cvcl_oracle as O
    rng = np.random.RandomState(seed)
    f = O.synth_trunk_features(rng, (B, K))
    ids, lens = O.synth_tokens(rng, B, L, V)
    return f, ids, lens


def synth_weights():
    from oracle import cvcl_oracle as O
    return O.synth_weights(np.random.RandomState(0), E, K, V)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            c = [x.strip() for x in r.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0])); mx.append(float(c[1]))
            except ValueError:
                continue
            for n, v in zip(names, c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------
def build_model(dev, group=None):
    import multimodal_baby_b200 as m
    args = argparse.Namespace(embedding_type="flat", embedding_dim=E, normalize_features=True,
                              fix_temperature=True, temperature=0.07, text_encoder="embedding")
    vocab = {str(i): i for i in range(V)}
    model = m.MultiModalModel(m.VisionEncoder(args, trunk="pooled"), m.TextEncoder(vocab, K, args), args)
    W, b, table = synth_weights()
    with torch.no_grad():
        model.image_embed.model.fc.weight.copy_(torch.from_numpy(W))
        model.image_embed.model.fc.bias.copy_(torch.from_numpy(b))
        model.text_embed.embedding.weight.copy_(torch.from_numpy(table))
    model.to(dev).train()
    model.materialize_logits = False          # the trainer ignores them (multimodal_lit.py:241-266)
    model.materialize_text_outputs = False
    model.materialize_features = False
    model.process_group = group
    return m, model


def m_staging():
    from multimodal_baby_b200 import staging
    return staging


def step_api(model, x, ids, lens, world):
    """the call a user makes: loss + backward (+ DDP-style gradient sum across ranks)."""
    params = getattr(model, "_bench_params", None)      # what optimizer.zero_grad(set_to_none=True) does, without
    if params is None:                                  # re-walking the module tree every step
        params = model._bench_params = list(model.parameters())
    for p in params:
        p.grad = None
    out = model.calculate_contrastive_loss(x, ids, lens)
    out[0].backward()          # sharded fused path: gradients are already summed over the ranks
    return out[0]


def step_work(B, world, sum_len):
    """algorithmic work of one rank's step: bytes every operand / result
    has to cross HBM at least once, flops of the contractions the reference performs."""
    Bg = B * world
    bytes_ = (2 * B * K            # x (bf16), read by the head GEMM; the dW GEMM re-reads it from L2
              + 2 * E * K          # bf16 weight shadow
              + 8 * B * L + 8 * B  # token ids, lengths
              + sum_len * E * 4    # embedding rows gathered (the 4.8 MB table is L2 resident across steps)
              + 4 * E * K + 4 * V * E + 4 * E)      # dW, d table, d bias written
    flops = 2 * B * K * E * 2 + 2 * B * Bg * E * (4 if world == 1 else 6)   # head + dW; S, dI, dT (+ column block)
    return bytes_, flops


def ncu_traffic(label):
    """DRAM bytes per launch of that kernel from the committed `ncu --set full` capture
    (profiles/r02_ncu_traffic.json, written by tools/ncu_traffic.py); None if not captured."""
    for name in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as fh:
                v = json.load(fh).get(label, {}).get("traffic")
            if v is not None:
                return v
        except (OSError, ValueError):
            pass
    return None


def step_roofline(ms, B, world, sum_len, peaks, kernel, phases=None):
    """roofline of the dominant kernel.  At N = 1 the whole step is ONE kernel, so its duration is the
    CUDA-event step time itself (no separate, differently-timed measurement).  The step is classified by
    the slower of its two ceilings."""
    bytes_, flops = step_work(B, world, sum_len)
    t_h = bytes_ / (peaks["hbm_gbs"] * 1e9)
    t_t = flops / (peaks["bf16_tflops"] * 1e12)
    sec = ms * 1e-3
    r = dict(kernel=kernel, ms=ms, algorithmic_bytes=bytes_, algorithmic_flops=flops,
             hbm_gb_s=bytes_ / sec / 1e9, hbm_frac=t_h / sec, tflop_s=flops / sec / 1e12, tensor_frac=t_t / sec,
             ceiling_us=max(t_h, t_t) * 1e6, peak_source=peaks["source"], traffic=ncu_traffic(kernel))
    if t_h >= t_t:
        r.update(bound="hbm", achieved=r["hbm_gb_s"], peak=peaks["hbm_gbs"], unit="GB/s", frac=r["hbm_frac"])
    else:
        r.update(bound="tensor", achieved=r["tflop_s"], peak=peaks["bf16_tflops"], unit="TFLOP/s", frac=r["tensor_frac"])
    if phases:
        r["phases_us"] = phases
    return r


def fused_phase_timeline(m, B):
    """in-kernel phase durations (globaltimer stamps of CTA 0, see include/cvcl_b200.h) of the last launch."""
    try:
        lay = m.ops.fused_layout(B, L, E, K, V)
        ws = m.ops._FUSED_WS.get((torch.cuda.current_device(), B, L, E, K, V))
        if ws is None:
            return None
        tm = ws[lay["ctrl"] + 128:lay["ctrl"] + 256].view(torch.int64).cpu().numpy()
        names = ["P0 head GEMM + text encoder", "P1 slab sum + normalise + token counts", "P2 similarity + row statistics",
                 "P3 dL/dlogits + dQ partials", "P4 normalise-backward", "P5 dW + d table + final sums"]
        out, prev = {}, tm[0]
        for k in range(1, 6):
            if tm[k]:
                out[names[k - 1]] = round(float(tm[k] - prev) / 1e3, 2); prev = tm[k]
        if tm[15]:
            out[names[5]] = round(float(tm[15] - prev) / 1e3, 2)
            out["kernel total"] = round(float(tm[15] - tm[0]) / 1e3, 2)
        return out
    except Exception:                              # noqa: BLE001
        return None


def _event_time(fn, reps, warm, flush):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.mean(ts), min(ts)


def config3_object(m, dev, peaks, flush, world, rank, group, B3=32768):
    """BASELINE configs[2]: global-batch contrastive loss over 32 768 pairs, fwd + bwd at the feature level
    (d img, d txt, d s).  N = 1: the whole batch on one GPU.  N > 1: rank r owns B3/N pairs, the features are
    all-gathered and every rank evaluates its row block and column block; max over ranks."""
    b = B3 // world
    g = torch.Generator().manual_seed(77 + rank)
    img = torch.nn.functional.normalize(torch.randn(b, E, generator=g), dim=1).to(dev)
    txt = torch.nn.functional.normalize(torch.randn(b, E, generator=g), dim=1).to(dev)

    def step():
        i = img.detach().requires_grad_(True); t = txt.detach().requires_grad_(True)
        out = m.ops.sim_infonce(i, t, S_FIXED, group, True)          # unit-norm rows (F.normalize above)
        out[0].backward()
    ms_eager, _ = _event_time(step, 5, 2, flush)
    ms, form = ms_eager, "eager autograd (op by op)"
    try:       # the same fwd+bwd as ONE CUDA graph (collectives included): no per-op host cost between the kernels
        i_leaf = img.clone().requires_grad_(True); t_leaf = txt.clone().requires_grad_(True)
        gstep3 = m.GraphedLossStep(lambda: m.ops.sim_infonce(i_leaf, t_leaf, S_FIXED, group, True)[0], [i_leaf, t_leaf])
        ms_g, _ = _event_time(gstep3, 5, 2, flush)
        del gstep3
    except Exception as exc:                               # noqa: BLE001
        ms_g = None
        form = "eager autograd (graph capture failed: %s)" % (repr(exc)[:80],)
    ok = torch.tensor([1 if (ms_g is not None and ms_g < ms_eager) else 0], device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()):
        ms, form = ms_g, "one CUDA graph (GraphedLossStep)"
    tt = torch.tensor([ms], device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt.item())
    flops_alg = 8.0 * B3 * B3 * E                       # one S for both directions + recompute + dI + dT
    flops_exec_rank = (10.0 if world == 1 else 12.0) * b * B3 * E
    sec = ms * 1e-3
    return {"workload": "global-batch InfoNCE, %d pairs, features in, fwd+bwd (d img, d txt, d s)" % B3,
            "pairs_per_gpu": b, "ms": ms, "ms_eager": ms_eager, "launch_form": form, "pairs_per_s": B3 / sec,
            "algorithmic_flops": flops_alg, "tflop_s_algorithmic": flops_alg / sec / 1e12,
            "tensor_frac_per_gpu": flops_alg / world / sec / 1e12 / peaks["bf16_tflops_sustained"],
            "executed_flops_per_gpu": flops_exec_rank,
            "tensor_frac_executed_per_gpu": flops_exec_rank / sec / 1e12 / peaks["bf16_tflops_sustained"],
            "peak": "bf16_tflops_sustained (%s)" % peaks["source"],
            "exchange": "NCCL all-gather of the bf16 features + LSEs (feature-level op)" if world > 1 else None}


def secondary_configs(m, dev, peaks, flush):
    """BASELINE configs 1, 4, 5 on one GPU, bounded to a few seconds each (config 3: config3_object)."""
    out = {}
    # ---- config 1: full model (ResNeXt-50 trunk as the stock torch module + this library's head), B = 8, 224^2
    try:
        args = argparse.Namespace(embedding_type="flat", embedding_dim=E, normalize_features=True,
                                  fix_temperature=True, temperature=0.07, text_encoder="embedding",
                                  cnn_model="resnext50_32x4d", finetune_cnn=False)
        vocab = {str(i): i for i in range(V)}
        torch.manual_seed(0)
        model = m.MultiModalModel(m.VisionEncoder(args, trunk="resnext"), m.TextEncoder(vocab, K, args), args)
        model.materialize_logits = model.materialize_text_outputs = model.materialize_features = False
        B1 = 8
        from oracle import cvcl_oracle as O
        ids, lens = O.synth_tokens(np.random.RandomState(3), B1, L, V)
        imgs_h = torch.rand(B1, 3, 224, 224).pin_memory()
        ids_h = torch.from_numpy(ids).pin_memory(); lens_h = torch.from_numpy(lens).pin_memory()
        # CPU baseline first (the same modules on the host cores: torch trunk + the oracle's op sequence for the head)
        model.train()
        torch.set_num_threads(os.cpu_count() or 1)
        fc = model.image_embed.model.fc
        tab = model.text_embed.embedding.weight

        def cpu_step():
            with torch.no_grad():
                pooled, _ = m.split_trunk_forward(model.image_embed, imgs_h, run_head=False)
            return O.contrastive_step(pooled, ids_h, lens_h, fc.weight.detach().clone(), fc.bias.detach().clone(),
                                      tab.detach().clone(), S_FIXED)["loss"]
        cpu_step()
        t0 = time.perf_counter(); n_cpu = 3
        for _ in range(n_cpu):
            cpu_loss = float(cpu_step())
        cpu_dt = (time.perf_counter() - t0) / n_cpu
        model.to(dev)
        x_d = torch.empty_like(imgs_h, device=dev); i_d = ids_h.to(dev); l_d = lens_h.to(dev)

        def gpu_step():
            x_d.copy_(imgs_h, non_blocking=True); i_d.copy_(ids_h, non_blocking=True); l_d.copy_(lens_h, non_blocking=True)
            for prm in model.parameters():
                prm.grad = None
            o = model.calculate_contrastive_loss(x_d, i_d, l_d)
            o[0].backward()
            return o[0].item()
        for _ in range(3):
            gpu_loss = gpu_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter(); n_gpu = 20
        for _ in range(n_gpu):
            gpu_loss = gpu_step()
        torch.cuda.synchronize()
        gpu_dt = (time.perf_counter() - t0) / n_gpu
        out["config1"] = {"workload": "full CVCL model, random init, B=8, 224x224: ResNeXt-50 trunk (stock torch module, frozen) "
                                      "+ contrastive head fwd+bwd; e2e with H2D of the images and D2H of the loss",
                          "gpu_ms": gpu_dt * 1e3, "gpu_pairs_per_s": B1 / gpu_dt, "gpu_loss": gpu_loss,
                          "cpu_ms": cpu_dt * 1e3, "cpu_pairs_per_s": B1 / cpu_dt, "cpu_loss": cpu_loss,
                          "cpu_cores": os.cpu_count(), "cpu_kind": "port (torch trunk + oracle head on the host cores)",
                          "loss_rel_diff": abs(gpu_loss - cpu_loss) / max(abs(cpu_loss), 1e-9),
                          "h2d_bytes_per_step": imgs_h.numel() * 4 + ids_h.numel() * 8 + lens_h.numel() * 8}
        del model
    except Exception as exc:                               # noqa: BLE001
        out["config1"] = {"error": repr(exc)[:300]}
    # ---- config 4: spatial 7x7 embeddings, max and mean, B = 1024, fwd+bwd from projected features
    try:
        from oracle import cvcl_oracle as O
        B4, HW = 1024, 49
        rng = np.random.RandomState(4)
        ids, lens = O.synth_tokens(rng, B4, L, V)
        table = O.synth_weights(rng, E, 8, V)[2]
        ids_d = torch.from_numpy(ids).to(dev); lens_d = torch.from_numpy(lens).to(dev); table_d = torch.from_numpy(table).to(dev)
        g = torch.Generator().manual_seed(4)
        imgs = torch.nn.functional.normalize(torch.randn(B4, HW, E, generator=g), dim=-1).to(dev)

        def sp(sim):
            def fn():
                i = imgs.detach().requires_grad_(True); tab = table_d.detach().requires_grad_(True)
                if sim == "max":
                    tok, _ = m.ops.text_features_spatial(ids_d, lens_d, tab, True)
                    loss = m.ops.infonce_from_match(m.ops.spatial_max_similarity(i, tok, lens_d, ids_d), S_FIXED)[0]
                else:
                    ip, tp = m.ops.spatial_mean_factors(i, ids_d, lens_d, tab, True)
                    loss = m.ops.sim_infonce(ip, tp, S_FIXED)[0]
                loss.backward()
            return fn
        # the same closures as ONE CUDA graph each (GraphedLossStep: forward + backward captured once): the form a
        # training loop uses; the eager op-by-op time is reported beside it
        i_leaf = imgs.clone().requires_grad_(True); tab_leaf = table_d.clone().requires_grad_(True)

        def spg(sim):
            def fn():
                if sim == "max":
                    tok, _ = m.ops.text_features_spatial(ids_d, lens_d, tab_leaf, True)
                    return m.ops.infonce_from_match(m.ops.spatial_max_similarity(i_leaf, tok, lens_d, ids_d), S_FIXED)[0]
                ip, tp = m.ops.spatial_mean_factors(i_leaf, ids_d, lens_d, tab_leaf, True)
                return m.ops.sim_infonce(ip, tp, S_FIXED)[0]
            return fn
        ms_eager, _ = _event_time(sp("max"), 5, 2, flush)
        gmax = m.GraphedLossStep(spg("max"), [i_leaf, tab_leaf])
        ms, _ = _event_time(gmax, 5, 2, flush)
        fl = 2.0 * B4 * B4 * HW * L * E + 4.0 * B4 * B4 * L * E
        out["config4_max"] = {"workload": "spatial 7x7, sim=max, B=1024, fwd+bwd (one CUDA graph)", "ms": ms,
                              "ms_eager": ms_eager, "pairs_per_s": B4 / (ms * 1e-3),
                              "algorithmic_flops": fl, "tensor_frac": fl / (ms * 1e-3) / 1e12 / peaks["bf16_tflops"]}
        del gmax
        ms_eager, _ = _event_time(sp("mean"), 10, 3, flush)
        gmean = m.GraphedLossStep(spg("mean"), [i_leaf, tab_leaf])
        ms, _ = _event_time(gmean, 10, 3, flush)
        by = 2.0 * (B4 * HW * E + B4 * L * E) * 4
        out["config4_mean"] = {"workload": "spatial 7x7, sim=mean, B=1024, fwd+bwd (one CUDA graph)", "ms": ms,
                               "ms_eager": ms_eager, "pairs_per_s": B4 / (ms * 1e-3),
                               "algorithmic_bytes": by, "hbm_frac": by / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}
        del gmean
    except Exception as exc:                               # noqa: BLE001
        out["config4"] = {"error": repr(exc)[:300]}
    # ---- config 5: Labeled-S style 4-way eval, 100k frames, 22 categories, fp32
    try:
        N5, C5 = 25000, 22
        g = torch.Generator().manual_seed(5)
        frames = torch.randn(N5 * 4, E, generator=g).to(dev); cats = torch.randn(C5, E, generator=g).to(dev)
        idx = torch.randint(0, C5, (N5,), generator=g).to(torch.int32).to(dev)
        ms, mn = _event_time(lambda: m.ops.eval_nway(frames, cats, idx, 4, True, S_FIXED, False), 20, 3, flush)
        by = N5 * 4 * E * 4.0
        out["config5"] = {"workload": "4-way eval, 100 000 frames, 22 categories, fp32 -> predictions", "ms": ms, "ms_min": mn,
                          "frames_per_s": N5 * 4 / (ms * 1e-3), "algorithmic_bytes": by,
                          "hbm_frac": by / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}
    except Exception as exc:                               # noqa: BLE001
        out["config5"] = {"error": repr(exc)[:300]}
    return out


def cpu_reference_step_fn(B):
    """the oracle port: same ATen fp32 op sequence as the reference's calculate_contrastive_loss +
    backward, on trunk-boundary features, all host threads."""
    from oracle import cvcl_oracle as O
    f, ids, lens = synth_batch(1234, B)
    W, b, table = synth_weights()
    tf, tids, tl = torch.from_numpy(f), torch.from_numpy(ids), torch.from_numpy(lens)
    tW, tb, tt = torch.from_numpy(W), torch.from_numpy(b), torch.from_numpy(table)

    def fn():
        return O.contrastive_step(tf, tids, tl, tW, tb, tt, S_FIXED)["loss"]
    return fn


def time_cpu(B, steps, warmup, budget_s=None):
    """-> (seconds per step, steps timed).  budget_s bounds the timed region: the number of timed steps
    is cut to what fits (estimated from the warm-up steps), never below 3."""
    torch.set_num_threads(os.cpu_count() or 1)
    fn = cpu_reference_step_fn(B)
    t0 = time.perf_counter()
    for _ in range(warmup):
        fn()
    est = (time.perf_counter() - t0) / max(warmup, 1)
    if budget_s is not None and est > 0:
        steps = max(3, min(steps, int(budget_s / est)))
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    dt = (time.perf_counter() - t0) / steps
    return dt, steps


# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs-per-gpu", type=int, default=512)
    ap.add_argument("--no-breakdown", action="store_true", help="(kept for old command lines; ignored)")
    ap.add_argument("--no-configs", action="store_true", help="skip the bounded runs of BASELINE configs 1, 3, 4, 5")
    a = ap.parse_args()
    if a.warmup < 3:
        a.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    B = a.pairs_per_gpu
    config = {"workload": "CVCL flat-embedding contrastive train step (fwd+bwd), %d synthetic pairs per GPU, "
                          "global InfoNCE batch %d, L<=25, E=512, K=2048, V=2350" % (B, B * max(world, 1)),
              "pairs_per_gpu": B, "global_batch": B * max(world, 1), "max_len": L,
              "parallelism": ("pairs sharded over %d ranks, global-batch InfoNCE (weak scaling: %d pairs per rank)"
                              % (world, B)) if world > 1 else "single GPU",
              "l2": "256 MiB write between timed steps (inputs are smaller than L2)"}
    # how THIS arm runs the workload (kept out of `config`, which names the workload and is identical for both arms)
    implementation = {"parallelism": ("one persistent kernel per rank: features, softmax partials and gradient tiles cross "
                                      "NVLink as stores from the producing phases, cross-rank grid barriers, owner-sums in "
                                      "rank order (no NCCL, no exchange kernels)") if world > 1 else
                                     "the whole step is one persistent kernel"}

    if a.impl == "reference":
        if rank != 0:
            return
        # the same workload as the B200 arm at this N: the global batch of B * world pairs in one process
        # (the reference is single-process); timed steps are bounded to about 90 s of CPU work
        Bref = B * max(world, a.gpus, 1)
        config["global_batch"] = Bref
        config["workload"] = config["workload"].replace("global InfoNCE batch %d" % (B * max(world, 1)),
                                                        "global InfoNCE batch %d" % Bref)
        dt, steps = time_cpu(Bref, a.steps, a.warmup, budget_s=90.0)
        val = Bref / dt
        cores = os.cpu_count() or 1
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": val, "unit": "pairs/s", "n_gpus": a.gpus,
            "steps": steps, "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": cores, "kind": "port",
                             "sample": "%d of the %d requested steps (bounded to ~90 s) of the B=%d flat train step "
                                       "(oracle port of the reference's ATen fp32 op sequence, torch %d threads)"
                                       % (steps, a.steps, Bref, cores)},
            "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    assert torch.cuda.is_available(), "bench.py --impl b200 needs a CUDA device (no CPU fallback)"
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    peaks = load_peaks()
    m, model = build_model(dev, group)
    from multimodal_baby_b200 import _cabi
    lib = _cabi.load()

    f, ids, lens = synth_batch(1234 + rank, B)
    # pinned staging as ONE arena [x | ids | lens] (what staging.PinnedBatchStager hands out): one H2D copy per step
    x_host, ids_host, lens_host = m_staging().packed_buffers(
        [((B, K), torch.bfloat16), ((B, L), torch.int64), ((B,), torch.int64)])
    x_host.copy_(torch.from_numpy(f)); ids_host.copy_(torch.from_numpy(ids)); lens_host.copy_(torch.from_numpy(lens))
    x = x_host.to(dev); ids_d = ids_host.to(dev); lens_d = lens_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # the graphed e2e step is built FIRST: it allocates its second pinned staging set, and freshly pinned host memory
    # copies at a fraction of the PCIe rate for the first second or two on these hosts (measured: 2.2 MB in 160-200 us
    # right after cudaHostAlloc, 46 us one second later; tools/h2d_probe2.py) -- by the time the e2e loop runs, every
    # staging buffer is in its steady state, as it is in a training run
    gstep, n_replaced, restaged_dt = None, 0, None
    # bf16 shadow of the projection weight: cast ONCE after loading; in training the optimizer kernel (FusedAdamW /
    # cvcl_adamw_multi_step) rewrites it together with the fp32 master, so the step itself never casts W (round 1
    # spent 4-5 us per step on that cast)
    _fcw = model.image_embed.model.fc.weight
    m.ops.register_weight_shadow(_fcw, _fcw.detach().to(torch.bfloat16).contiguous())
    t_gstep = time.perf_counter()
    try:
        gstep = m.GraphedContrastiveStep(model, x_host, ids_host, lens_host, prefetch=True, lagged_loss=True,
                                         own_staging=True)
    except Exception as exc:                 # noqa: BLE001
        if rank == 0:
            print("graphed e2e step unavailable: %s" % exc, file=sys.stderr)

    # ------------------------------------------------------------------ device-timed value
    fcw, fcb = model.image_embed.model.fc.weight, model.image_embed.model.fc.bias
    table = model.text_embed.embedding.weight
    fused = world == 1 and m.ops.fused_supported(B, L, E, K, V)
    fused_sh = world > 1 and m.ops.FUSED_STEP and bool(lib.cvcl_flat_fused_sharded_supported(B, L, E, K, V, world))
    implementation["step_kernel"] = ("one persistent cooperative kernel (csrc/fused_step.cuh)" if (fused or fused_sh) else
                             "multi-kernel sequence (cvcl_flat_contrastive_step / flat_step_sharded)")
    graph = None
    if world == 1:
        def raw_step():
            return m.ops.flat_contrastive_step(x, ids_d, lens_d, fcw, fcb, table, S_FIXED, True, True, False)
    else:
        def raw_step():
            return m.ops.flat_step_sharded(x, ids_d, lens_d, fcw, fcb, table, S_FIXED, True, True, False, group)
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                raw_step()
        torch.cuda.current_stream().wait_stream(side)
        barrier()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            g_out = raw_step()
        run_step = graph.replay
        run_step(); barrier()
    except Exception as exc:                 # e.g. NCCL capture unsupported: time the eager step
        if rank == 0:
            print("CUDA graph capture failed (%s); timing the eager step" % exc, file=sys.stderr)
        graph = None
        run_step = (lambda: raw_step()) if world == 1 else (lambda: step_api(model, x, ids_d, lens_d, world))

    sampler = ClockSampler(local_rank)
    for _ in range(a.warmup):
        flush.zero_(); run_step()
    barrier()
    if rank == 0:
        sampler.start()
    n0 = lib.cvcl_launch_count()
    evs = []
    barrier()
    t_wall0 = time.perf_counter()
    for _ in range(a.steps):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); run_step(); e1.record()
        evs.append((e0, e1))
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    ms = statistics.mean(step_ms)
    n_launch = lib.cvcl_launch_count() - n0
    # the same K steps launched directly (one C-ABI call per step, no graph): with a single kernel per step a graph
    # replay only adds its own launch latency; the faster of the two forms is the value, both are reported
    launch_form = "cuda graph replay" if graph is not None else "direct launch"
    ms_graph = ms if graph is not None else None
    ms_direct = None
    if graph is not None and world == 1:
        for _ in range(a.warmup):
            flush.zero_(); raw_step()
        barrier()
        evs2 = []
        for _ in range(a.steps):
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); raw_step(); e1.record()
            evs2.append((e0, e1))
        barrier()
        direct_ms = [e0.elapsed_time(e1) for e0, e1 in evs2]
        ms_direct = statistics.mean(direct_ms)
        if ms_direct < ms:
            ms, step_ms, launch_form = ms_direct, direct_ms, "direct launch"
    # The clock samples belong to the device-timed region.  K steps can be shorter than the poller's 50 ms period, so
    # the same step keeps running (untimed) until two samples exist; then the poller is stopped: its driver queries
    # must not sit on the host path of the wall-clock e2e loops below.
    clocks = None
    if world == 1:
        t_s = time.perf_counter()
        while len(sampler.rows) < 2 and time.perf_counter() - t_s < 1.5:
            flush.zero_(); run_step()
    else:
        for _ in range(1000):               # sharded steps are collective: the same fixed count on every rank
            flush.zero_(); run_step()
    barrier()
    if rank == 0:
        clocks = sampler.stop()
    if graph is not None:       # replays launch the captured kernels without passing through the C ABI
        n1 = lib.cvcl_launch_count(); raw_step(); per = lib.cvcl_launch_count() - n1
        n_launch = per * a.steps
        torch.cuda.synchronize()
    t = torch.tensor([ms], device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = B * world / (ms * 1e-3)

    # ------------------------------------------------------------------ e2e through the public API
    # (a) eager: MultiModalModel.calculate_contrastive_loss(...) + backward(), pinned host inputs
    e2e_steps = max(500, min(a.steps, 2000))     # wall-clock timed: enough calls that start-up effects do not dominate
    for _ in range(3):
        x.copy_(x_host, non_blocking=True); step_api(model, x, ids_d, lens_d, world).item()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        x.copy_(x_host, non_blocking=True)
        ids_d.copy_(ids_host, non_blocking=True)
        lens_d.copy_(lens_host, non_blocking=True)
        loss_host = step_api(model, x, ids_d, lens_d, world).item()      # D2H read of the loss
    barrier()
    eager_dt = (time.perf_counter() - t0) / e2e_steps
    # (b) the same step as a captured CUDA graph (GraphedContrastiveStep): H2D of the staged batch,
    #     fwd+bwd, D2H of the loss all inside the replay; wall clock per call incl. the stream sync
    e2e_dt, e2e_api = eager_dt, "MultiModalModel.calculate_contrastive_loss + backward (eager)"
    try:
        if gstep is None:
            raise RuntimeError("not built")
        # the pinned staging buffers are double-buffered (a replay in flight may still be reading its set): stage a
        # second, different batch in the other set so that consecutive steps copy different data
        f2, ids2, lens2 = synth_batch(4321 + rank, B)
        for dst, src in zip(gstep.host_sets[-1], (torch.from_numpy(f2).to(torch.bfloat16), torch.from_numpy(ids2),
                                                  torch.from_numpy(lens2))):
            dst.copy_(src)
        gstep.prime()
        # warm-up to the steady state: blocks of 200 calls until the staging memory is at least 3 s old AND two
        # consecutive blocks agree within 5 % (6 s of warm-up at most)
        prev_blk, t_w0 = None, time.perf_counter()
        while True:
            tb = time.perf_counter()
            for _ in range(200):
                gstep()
            gstep.flush()
            blk = time.perf_counter() - tb
            aged = time.perf_counter() - t_gstep >= 3.0
            done = aged and prev_blk is not None and abs(blk - prev_blk) <= 0.05 * prev_blk
            done = done or time.perf_counter() - t_w0 >= 6.0
            if world > 1:                    # sharded steps are collective: every rank leaves after the same block
                import torch.distributed as dist
                flag = torch.tensor([0 if done else 1], device=dev)
                dist.all_reduce(flag, op=dist.ReduceOp.MAX)
                done = int(flag.item()) == 0
            if done:
                break
            prev_blk = blk
        # a staging set that has turned slow since it was allocated is replaced (GraphedContrastiveStep.check_staging);
        # the same two rounds on every rank (sharded steps are collective)
        n_replaced = 0
        for _ in range(5):
            n_fix = gstep.check_staging()
            n_replaced += n_fix
            if n_fix:
                time.sleep(1.0)              # newly pinned memory needs a moment before it copies at full rate
            tb = time.perf_counter()
            for _ in range(200):
                gstep()
            gstep.flush()
            blk_ms = (time.perf_counter() - tb) / 200 * 1e3
            again = 1 if (n_fix or blk_ms > 1.25 * max(ms, 0.052)) else 0
            if world > 1:                    # sharded steps are collective: every rank takes the same number of rounds
                import torch.distributed as dist
                flag = torch.tensor([again], device=dev)
                dist.all_reduce(flag, op=dist.ReduceOp.MAX)
                again = int(flag.item())
            if not again:
                break
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            loss_host = gstep()            # loss of the previous replay (read back every step)
        loss_host = gstep.flush()          # ... and of the last one, inside the timed region
        barrier()
        g_dt = (time.perf_counter() - t0) / e2e_steps
        # the same loop with the loader's part on the same thread: every step first WRITES the next batch (pageable
        # source, alternating between two batches) into the pinned staging set the call will copy -- reported beside
        # the headline e2e, which starts from pinned memory as the contract says
        src_sets = [(torch.from_numpy(f).to(torch.bfloat16), torch.from_numpy(ids), torch.from_numpy(lens)),
                    (torch.from_numpy(f2).to(torch.bfloat16), torch.from_numpy(ids2), torch.from_numpy(lens2))]
        n_rs = max(100, e2e_steps // 2)
        barrier()
        t0 = time.perf_counter()
        for k in range(n_rs):
            xs, is_, ls_ = src_sets[k & 1]
            gstep.x_host.copy_(xs); gstep.ids_host.copy_(is_); gstep.lens_host.copy_(ls_)
            gstep()
        gstep.flush()
        barrier()
        restaged_dt = (time.perf_counter() - t0) / n_rs
        if g_dt < e2e_dt:
            e2e_dt, e2e_api = g_dt, ("GraphedContrastiveStep(model, prefetch=True, lagged_loss=True)() = "
                                     "calculate_contrastive_loss + backward as one CUDA graph; every call copies one full "
                                     "batch H2D (overlapped with the kernels of the previously copied batch) and reads back "
                                     "the loss of the previous replay (one replay in flight, flushed inside the timed region)")
    except Exception as exc:
        if rank == 0:
            print("graphed e2e step unavailable: %s" % exc, file=sys.stderr)
    t = torch.tensor([e2e_dt, eager_dt], device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_dt, eager_dt = float(t[0].item()), float(t[1].item())
    h2d = x_host.numel() * 2 + ids_host.numel() * 8 + lens_host.numel() * 8

    # ------------------------------------------------------------------ roofline of the step kernel
    roof = None
    if rank == 0:
        sum_len = int(lens.sum())
        if world == 1 and fused:
            run_step(); torch.cuda.synchronize()
            roof = step_roofline(ms, B, 1, sum_len, peaks, "flat_step_kernel", fused_phase_timeline(m, B))
        else:
            roof = step_roofline(ms, B, world, sum_len, peaks,
                                 ("flat_step_kernel (sharded: one kernel per rank incl. the exchange and the gradient sum)"
                                  if fused_sh else "sharded step (multi-kernel sequence incl. the peer-memory collectives)")
                                 if world > 1 else "flat step (multi-kernel sequence)")
    # ------------------------------------------------------------------ the other BASELINE configurations (bounded)
    configs = None
    if not a.no_configs:
        try:
            configs = {"config3": config3_object(m, dev, peaks, flush, world, rank, group)}    # every rank takes part
        except Exception as exc:                           # noqa: BLE001
            configs = {"config3": {"error": repr(exc)[:300]}}
        if rank == 0 and world == 1:
            configs.update(secondary_configs(m, dev, peaks, flush))

    def finish():
        # leave without tearing the communicator down: destroying an NCCL process group while a
        # captured CUDA graph still references its kernels can hang
        sys.stdout.flush(); sys.stderr.flush()
        if world > 1:
            torch.cuda.synchronize()
            import torch.distributed as dist
            dist.barrier()
            os._exit(0)

    if rank != 0:
        finish()
        return

    cpu = None
    if world == 1:
        n_cpu = 40
        cdt, n_cpu = time_cpu(B, n_cpu, 3, budget_s=30.0)
        cores = os.cpu_count() or 1
        cpu = {"value": B / cdt, "unit": "pairs/s", "cores": cores, "kind": "port", "ms_per_step": cdt * 1e3,
               "sample": "%d steps of the same B=%d flat train step (oracle port: the reference's ATen fp32 op "
                         "sequence + autograd on CPU, %d torch threads)" % (n_cpu, B, cores)}
    if world > 1:
        from multimodal_baby_b200 import sharding as _sh
        used = [v is not None for v in _sh.PeerExchange._cache.values()]
        implementation["exchange"] = (("inside the step kernel: posted stores over NVLink peer memory from the producing phases, "
                               "cross-rank grid barriers on flag words (csrc/fused_step.cuh)") if fused_sh else
                              ("single kernels over NVLink peer memory with in-kernel cross-rank barriers: feature "
                               "all-gather, LSE all-gather, two-shot in-place gradient all-reduce (csrc/peer_collectives.cuh)")
                              ) if used and all(used) else "NCCL all-gather / all-reduce"
    line = {
        "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config, "implementation": implementation,
        "e2e": {"value": B * world / e2e_dt, "unit": "pairs/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 32, "ms_per_step": e2e_dt * 1e3, "steps": e2e_steps, "api": e2e_api,
                "staging_h2d_probe_us": getattr(gstep, "staging_probe_us", None), "staging_sets_replaced": n_replaced,
                "restaged_ms_per_step": restaged_dt * 1e3 if restaged_dt is not None else None,
                "eager_module_api": {"value": B * world / eager_dt, "ms_per_step": eager_dt * 1e3}},
        "gpu_launches": int(n_launch), "gpu_launches_per_step": int(n_launch // max(a.steps, 1)),
        "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "configs": configs,
        "timing": {"ms_min": min(step_ms), "ms_median": statistics.median(step_ms),
                   "wall_s_incl_flush": t_wall, "cuda_graph": graph is not None, "launch_form": launch_form,
                   "ms_graph_replay": ms_graph, "ms_direct_launch": ms_direct},
    }
    print(json.dumps(line))
    finish()


if __name__ == "__main__":
    main()
