"""Run the spatial (7x7) contrastive step, sim = mean or max, B = 1024, eagerly a few times (target for ncu launch
lists / captures).  usage: python tools/run_spatial.py mean|max [B] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import multimodal_baby_b200 as m
from bench import S_FIXED, E, L, V
from oracle import cvcl_oracle as O

sim = sys.argv[1] if len(sys.argv) > 1 else "mean"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda:0"); HW = 49
rng = np.random.RandomState(4)
ids, lens = O.synth_tokens(rng, B, L, V)
table = O.synth_weights(rng, E, 8, V)[2]
ids_d = torch.from_numpy(ids).to(dev); lens_d = torch.from_numpy(lens).to(dev)
g = torch.Generator().manual_seed(4)
imgs = torch.nn.functional.normalize(torch.randn(B, HW, E, generator=g), dim=-1).to(dev)
i_leaf = imgs.clone().requires_grad_(True); tab = torch.from_numpy(table).to(dev).requires_grad_(True)
for _ in range(iters):
    i_leaf.grad = None; tab.grad = None
    if sim == "max":
        tok, _ = m.ops.text_features_spatial(ids_d, lens_d, tab, True)
        loss = m.ops.infonce_from_match(m.ops.spatial_max_similarity(i_leaf, tok, lens_d, ids_d), S_FIXED)[0]
    else:
        _, tp = m.ops.text_features_spatial(ids_d, lens_d, tab, True, 1.0 / HW, want_tok=False)
        loss = m.ops.sim_infonce(m.ops.spatial_pool(i_leaf), tp, S_FIXED)[0]
    loss.backward()
torch.cuda.synchronize()
print("loss", loss.item())
