"""ncu driver: a few fused transformer training steps (GPU only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "transformer-physx_b200")):
    sys.path.insert(0, p)
import torch
from tests import synth
from tests.util import make_gpt2
kind = sys.argv[1] if len(sys.argv) > 1 else "lorenz"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device("cuda:0")
cfg = synth.make_config(kind)
m = make_gpt2(cfg, 2025, False, dev)
x = torch.randn(B, cfg.n_ctx, cfg.n_embd, device=dev)
y = torch.randn(B, cfg.n_ctx, cfg.n_embd, device=dev)
for _ in range(3):
    l = m.fused_loss_and_grads(x, y)[0]
torch.cuda.synchronize()
print(float(l))
