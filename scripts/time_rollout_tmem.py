"""Time the on-chip (tensor-memory) window rollout (variant 10) against the HBM-window kernel (variant 2).
    python scripts/time_rollout_tmem.py [kind] [B] [S]"""
import ctypes
import sys
import torch
sys.path.insert(0, "transformer-physx_b200")
sys.path.insert(0, ".")
from tests import synth
from tests.util import make_gpt2
from trphysx_b200 import _native as N

kind = sys.argv[1] if len(sys.argv) > 1 else "rossler"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
S = int(sys.argv[3]) if len(sys.argv) > 3 else 256
UC = (sys.argv[5] != "0") if len(sys.argv) > 5 else True
dev = torch.device("cuda:0")
cfg = synth.make_config(kind)
m = make_gpt2(cfg, 903, True, dev, sigma=0.1)
x0 = torch.nn.functional.layer_norm(torch.randn(B, 1, cfg.n_embd, device=dev), (cfg.n_embd,))
bpc = ctypes.c_float()
N.check(N.lib().trpx_probe_tmem_ld(0, ctypes.byref(bpc)))
print(f"tcgen05.ld: {bpc.value:.1f} B/clk/SM (8 warps, 32x32b.x32)")
res = {}
for variant in [int(v) for v in (sys.argv[4].split(",") if len(sys.argv) > 4 else ["10", "2"])]:
    m.set_rollout_tuning(variant=variant)
    for _ in range(2):
        out = m.generate(x0, max_length=S + 1, use_cache=UC, return_presents=False)[0]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        out = m.generate(x0, max_length=S + 1, use_cache=UC, return_presents=False)[0]
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    res[variant] = out
    print(f"{kind} use_cache={UC} B={B} S={S} variant {variant} {m.last_launch()}: {ms:.2f} ms = {B * S / ms / 1e3:.1f} M traj-steps/s")
ks = sorted(res); a, b = res[ks[-1]][:, :65].double(), res[ks[0]][:, :65].double()
print("rel-L2 variant 10 vs 2 (first 64 steps):", float((a - b).norm() / b.norm()))
