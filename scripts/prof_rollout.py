"""Tiny driver for ncu: runs generate() a few times with fixed tuning.  GPU only."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "transformer-physx_b200")):
    sys.path.insert(0, p)
import torch
from bench import build_modules
ap = argparse.ArgumentParser()
ap.add_argument("--kind", default="rossler")
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--warps", type=int, default=0)
ap.add_argument("--group", type=int, default=0)
ap.add_argument("--batch", type=int, default=16384)
ap.add_argument("--horizon", type=int, default=64)
ap.add_argument("--nocache", action="store_true")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda:0")
if a.kind == "cylinder":
    from tests import synth
    from tests.util import make_gpt2
    cfg = synth.make_config("cylinder")
    tr = make_gpt2(cfg, 2025, False, dev)
else:
    cfg, emb, tr = build_modules(a.kind, dev)
tr.set_rollout_tuning(variant=a.variant, warps_per_cta=a.warps, traj_per_warp=a.group)
g0 = torch.nn.functional.layer_norm(torch.randn(a.batch, 1, cfg.n_embd, device=dev), (cfg.n_embd,))
for _ in range(a.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = tr.generate(g0, max_length=a.horizon + 1, use_cache=not a.nocache, return_presents=False)[0]
    e1.record()
    torch.cuda.synchronize()
    print(f"{a.kind} variant={a.variant} B={a.batch} S={a.horizon} cache={not a.nocache}: {e0.elapsed_time(e1):.3f} ms  {tr.last_launch()}")
