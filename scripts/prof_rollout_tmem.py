"""One launch of the tensor-memory window rollout (variant 10) for ncu: python scripts/prof_rollout_tmem.py [kind] [B] [S] [variant]"""
import sys
import torch
sys.path.insert(0, "transformer-physx_b200")
sys.path.insert(0, ".")
from tests import synth
from tests.util import make_gpt2

kind = sys.argv[1] if len(sys.argv) > 1 else "rossler"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 8 * 4
S = int(sys.argv[3]) if len(sys.argv) > 3 else 96
variant = int(sys.argv[4]) if len(sys.argv) > 4 else 10
dev = torch.device("cuda:0")
cfg = synth.make_config(kind)
m = make_gpt2(cfg, 903, True, dev, sigma=0.1)
m.set_rollout_tuning(variant=variant)
x0 = torch.nn.functional.layer_norm(torch.randn(B, 1, cfg.n_embd, device=dev), (cfg.n_embd,))
for _ in range(2):
    out = m.generate(x0, max_length=S + 1, use_cache=True, return_presents=False)[0]
torch.cuda.synchronize()
print(m.last_launch(), float(out.abs().mean()))
