"""Cylinder decoder stack: generate() time for a batch under different rollout tunings.
    python scripts/time_cyl_generate.py B [S]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "transformer-physx_b200"))
import torch
from tests import synth
from tests.util import make_gpt2

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = int(sys.argv[2]) if len(sys.argv) > 2 else 400
dev = torch.device("cuda:0")
cfg = synth.make_config("cylinder")
m = make_gpt2(cfg, 11, False, dev)
x0 = torch.nn.functional.layer_norm(torch.randn(B, 1, cfg.n_embd, device=dev), (cfg.n_embd,))
for variant, g in ((0, 0), (7, 1), (7, 2), (7, 4), (1, 0), (1, 2), (1, 4), (1, 8)):
    m.set_rollout_tuning(variant=variant, traj_per_warp=g)
    try:
        for _ in range(2):
            m.generate(x0, max_length=S + 1, use_cache=True, return_presents=False)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            m.generate(x0, max_length=S + 1, use_cache=True, return_presents=False)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        print(f"B={B} S={S} tuning=({variant},{g}) {m.last_launch()}: {ms:.2f} ms = {ms * 1e3 / S:.1f} us/step, {B * S / ms:.0f} k traj-steps/s")
    except Exception as e:
        print(f"B={B} tuning=({variant},{g}): {type(e).__name__} {str(e)[:100]}")
