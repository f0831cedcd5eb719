"""Lorenz / Rossler `recover` at N samples: FP32 SIMT kernel (mode 1), 3xTF32 mma.sync kernel (mode 2), tcgen05 kernel
(mode 0 = automatic for large N).  Parity of each against the SIMT kernel and, on a slice, against the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "transformer-physx_b200"))
import torch
from bench import build_modules
from oracle import trpx_oracle as O
dev = torch.device("cuda:0")
cfg, emb, tr = build_modules("rossler", dev)
for N in [int(a) for a in (sys.argv[1:] or ["100003", "16842752"])]:
    torch.manual_seed(1)
    g = torch.nn.functional.layer_norm(torch.randn(N, 32, device=dev), (32,)) * 1.5
    outs = {}
    for mode in (1, 2, 0):
        emb.set_recover_mode(mode)
        for _ in range(2):
            y = emb.recover(g)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            y = emb.recover(g)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        outs[mode] = y
        print(f"N={N} mode {mode}: {ms:.3f} ms = {N * 35000 / ms / 1e9:.1f} TFLOP/s", flush=True)
    for mode in (2, 0):
        d = (outs[mode].double() - outs[1].double())
        print(f"  mode {mode} vs SIMT: rel-L2 {float(d.norm() / outs[1].double().norm()):.3e}, max-abs {float(d.abs().max()):.3e}")
    emb.set_recover_mode(0)
