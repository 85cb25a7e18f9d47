"""Where the end-to-end step time goes at 512 pairs (one GPU): the one-kernel step cold (L2 flushed) vs warm
from its in-kernel stamps, the H2D copies alone, and the per-call wall-clock distribution of
GraphedContrastiveStep in its three modes.  Measurement helper (gpurun), not part of the product."""
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def pct(v, q):
    v = sorted(v)
    return v[min(len(v) - 1, int(q * len(v)))]


def main():
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    m, model = bench.build_model(dev)
    B = 512
    f, ids, lens = bench.synth_batch(1234, B)
    x_host, ids_host, lens_host = m.staging.packed_buffers(
        [((B, bench.K), torch.bfloat16), ((B, bench.L), torch.int64), ((B,), torch.int64)])
    x_host.copy_(torch.from_numpy(f)); ids_host.copy_(torch.from_numpy(ids)); lens_host.copy_(torch.from_numpy(lens))
    x = x_host.to(dev); ids_d = ids_host.to(dev); lens_d = lens_host.to(dev)
    fcw, fcb = model.image_embed.model.fc.weight, model.image_embed.model.fc.bias
    table = model.text_embed.embedding.weight
    m.ops.register_weight_shadow(fcw, fcw.detach().to(torch.bfloat16).contiguous())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def raw():
        return m.ops.flat_contrastive_step(x, ids_d, lens_d, fcw, fcb, table, bench.S_FIXED, True, True, False)
    for _ in range(5):
        raw()
    torch.cuda.synchronize()
    for mode in ("warm", "cold"):
        tl = []
        for _ in range(5):
            if mode == "cold":
                flush.zero_()
            raw(); torch.cuda.synchronize()
            tl.append(bench.fused_phase_timeline(m, B))
        print(mode, "kernel total us:", [t["kernel total"] for t in tl])
        print(mode, "phases:", tl[-1])
        lay = m.ops.fused_layout(B, bench.L, bench.E, bench.K, bench.V)
        ws = m.ops._FUSED_WS.get((torch.cuda.current_device(), B, bench.L, bench.E, bench.K, bench.V))
        tm = ws[lay["ctrl"] + 128:lay["ctrl"] + 128 + 8 * 48].view(torch.int64).cpu().numpy()
        print(mode, "stamps us from start: barriers", " ".join("%.1f" % ((tm[k] - tm[0]) / 1e3) for k in range(1, 7) if tm[k]),
              "| end %.1f |" % ((tm[15] - tm[0]) / 1e3),
              " ".join("%d:%.1f" % (k, (tm[k] - tm[0]) / 1e3) for k in range(16, 30) if tm[k]))

    # H2D copies alone
    s = torch.cuda.current_stream()
    for name, bufs in (("x only", [(x, x_host)]), ("x+ids+lens", [(x, x_host), (ids_d, ids_host), (lens_d, lens_host)])):
        ts = []
        for _ in range(50):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for d, h in bufs:
                d.copy_(h, non_blocking=True)
            e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
        nb = sum(h.numel() * h.element_size() for _, h in bufs)
        print("H2D %s: %d bytes, median %.1f us (%.1f GB/s), min %.1f" % (name, nb, statistics.median(ts),
                                                                      nb / statistics.median(ts) / 1e3, min(ts)))

    for kw in (dict(prefetch=False, lagged_loss=False), dict(prefetch=True, lagged_loss=False),
               dict(prefetch=True, lagged_loss=True)):
        g = m.GraphedContrastiveStep(model, x_host, ids_host, lens_host, **kw)
        if kw["prefetch"]:
            g.prime()
        for _ in range(20):
            g()
        g.flush(); torch.cuda.synchronize()
        per = []
        t00 = time.perf_counter()
        for _ in range(2000):
            t0 = time.perf_counter(); g(); per.append((time.perf_counter() - t0) * 1e6)
        g.flush(); torch.cuda.synchronize()
        tot = (time.perf_counter() - t00) / 2000 * 1e6
        print("graphed %s: mean %.1f us/call; p10 %.1f p50 %.1f p90 %.1f p99 %.1f max %.1f" % (
            kw, tot, pct(per, .1), pct(per, .5), pct(per, .9), pct(per, .99), max(per)))
        print("   calls 1000..1047:", " ".join("%d" % v for v in per[1000:1048]))
        if kw["lagged_loss"]:
            for k, ha in enumerate(g._host_arenas):          # the two pinned staging sets: is one of them slow?
                ts = []
                for _ in range(30):
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    e0.record(); g._dev_arenas[0].copy_(ha, non_blocking=True); e1.record(); e1.synchronize()
                    ts.append(e0.elapsed_time(e1) * 1e3)
                print("   H2D from staging set %d (%d bytes, pinned=%s, ptr %x): median %.1f us" % (
                    k, ha.numel(), ha.is_pinned(), ha.data_ptr(), statistics.median(ts)))
            fresh = [torch.empty(2203648, dtype=torch.uint8).pin_memory() for _ in range(3)]
            big = torch.empty(3 * 2203648, dtype=torch.uint8).pin_memory()
            fresh += [big[i * 2203648:(i + 1) * 2203648] for i in range(3)]
            for k, ha in enumerate(fresh):
                ha.fill_(1)
                ts = []
                for _ in range(30):
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    e0.record(); g._dev_arenas[0].copy_(ha, non_blocking=True); e1.record(); e1.synchronize()
                    ts.append(e0.elapsed_time(e1) * 1e3)
                print("   H2D from fresh pinned buffer %d (ptr %x): median %.1f us" % (k, ha.data_ptr(), statistics.median(ts)))
            # where does a slow call spend its time: the replay (cudaGraphLaunch) or the wait for the previous replay?
            tr, tw = [], []
            for _ in range(400):
                which = g.calls % len(g.graphs)
                t0 = time.perf_counter(); g.graphs[which].replay(); t1 = time.perf_counter()
                g.calls += 1
                g.events[which].record(torch.cuda.current_stream(dev))
                prev, g.pending = g.pending, which
                if prev is not None:
                    g.events[prev].synchronize()
                t2 = time.perf_counter()
                tr.append((t1 - t0) * 1e6); tw.append((t2 - t1) * 1e6)
            g.flush(); torch.cuda.synchronize()
            print("   lagged split: replay() p10 %.1f p50 %.1f p90 %.1f | record+wait p10 %.1f p50 %.1f p90 %.1f" % (
                pct(tr, .1), pct(tr, .5), pct(tr, .9), pct(tw, .1), pct(tw, .5), pct(tw, .9)))
            print("   replay us:", " ".join("%d" % v for v in tr[200:232]))
            print("   wait   us:", " ".join("%d" % v for v in tw[200:232]))

    # eager module API
    per = []
    for i in range(300):
        t0 = time.perf_counter()
        x.copy_(x_host, non_blocking=True); ids_d.copy_(ids_host, non_blocking=True); lens_d.copy_(lens_host, non_blocking=True)
        bench.step_api(model, x, ids_d, lens_d, 1).item()
        per.append((time.perf_counter() - t0) * 1e6)
    per = per[50:]
    print("eager module API: p10 %.1f p50 %.1f p90 %.1f" % (pct(per, .1), pct(per, .5), pct(per, .9)))
    # host cost alone (no sync)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(200):
        bench.step_api(model, x, ids_d, lens_d, 1)
    host = (time.perf_counter() - t0) / 200 * 1e6
    torch.cuda.synchronize()
    print("eager module API host enqueue cost (no sync): %.1f us/step" % host)
    import cProfile, pstats, io
    pr = cProfile.Profile()
    pr.enable()
    for i in range(300):
        bench.step_api(model, x, ids_d, lens_d, 1)
    pr.disable()
    torch.cuda.synchronize()
    sio = io.StringIO()
    pstats.Stats(pr, stream=sio).sort_stats("cumulative").print_stats(45)
    print(sio.getvalue())
    sio = io.StringIO()
    pstats.Stats(pr, stream=sio).sort_stats("tottime").print_stats(30)
    print(sio.getvalue())


if __name__ == "__main__":
    main()
