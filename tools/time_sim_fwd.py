"""Forward similarity + InfoNCE statistics at large batch: one similarity pass (unit-norm features) vs one GEMM per
direction (CVCL_B200_SIM_TWO_PASS=1).  usage: python tools/time_sim_fwd.py [B ...]"""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_baby_b200 as m
dev = torch.device("cuda:0")
S = float(-math.log(0.07))
for B in [int(a) for a in sys.argv[1:]] or [8192, 32768]:
    g = torch.Generator().manual_seed(1)
    x = torch.nn.functional.normalize(torch.randn(B, 512, generator=g), dim=1).to(dev).to(torch.bfloat16)
    y = torch.nn.functional.normalize(torch.randn(B, 512, generator=g), dim=1).to(dev).to(torch.bfloat16)
    for mode in ("one", "two"):
        if mode == "two":
            os.environ["CVCL_B200_SIM_TWO_PASS"] = "1"
        else:
            os.environ.pop("CVCL_B200_SIM_TWO_PASS", None)
        f = lambda: m.ops.sim_infonce_fwd(x, y, y, x, S, 0, 1.0 / B, True)
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); f(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
        print(B, mode, "pass: %.3f ms (min %.3f)  -> %.0f TFLOP/s algorithmic (2*B*B*E)" % (
            sum(ts) / len(ts), min(ts), 2.0 * B * B * 512 / min(ts) / 1e9))
