"""Headline rollout (Rossler 65,536 x 256, use_cache=True) as one launch or as k launches over trajectory slices
(each slice gets the host's warps-per-CTA choice for its size): same box, alternating."""
import sys
sys.path.insert(0, "transformer-physx_b200"); sys.path.insert(0, ".")
import torch
from bench import build_modules
dev = torch.device("cuda:0")
cfg, emb, tr = build_modules("rossler", dev)
B, S = 65536, 256
g0 = torch.nn.functional.layer_norm(torch.randn(B, 1, cfg.n_embd, device=dev), (cfg.n_embd,))
def run(k):
    outs = []
    for i in range(k):
        lo, hi = B * i // k, B * (i + 1) // k
        outs.append(tr.generate(g0[lo:hi], max_length=S + 1, use_cache=True, return_presents=False)[0])
    return outs
for rep in range(2):
    for k in (1, 2, 3, 4, 7):
        run(k)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            run(k)
        b.record()
        torch.cuda.synchronize()
        print(f"rep {rep}: {k} launch(es): {a.elapsed_time(b) / 3:.2f} ms  last launch {tr.last_launch()}", flush=True)
