"""Quick GPU check of the tcgen05 decoder (conv mode 3) against the FP32 SIMT path (mode 1) and the oracle."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "transformer-physx_b200"))
import torch
from tests import synth
from tests.util import make_cylinder
from oracle import trpx_oracle as O

dev = torch.device("cuda:0")
cfg = synth.make_config("cylinder")
m = make_cylinder(cfg, 55, dev)
sd = O.to_torch(synth.cylinder_embedding_weights(cfg, 55))
for N in (1, 3, 70):
    g = torch.from_numpy(synth.normal_embeds((N, 128), 5 + N))
    ref = O.cylinder_recover(sd, cfg, g) if N <= 3 else None
    outs = {}
    for mode in (1, 3):
        m.set_conv_mode(mode)
        with torch.no_grad():
            outs[mode] = m.recover(g.to(dev))
        torch.cuda.synchronize()
    print(f"N={N}: mode3 vs mode1 rel-L2 {O.rel_l2(outs[3].cpu(), outs[1].cpu()):.3e}", end="")
    if ref is not None:
        print(f"; vs oracle: mode1 {O.rel_l2(outs[1].cpu(), ref):.3e} mode3 {O.rel_l2(outs[3].cpu(), ref):.3e}", end="")
    d = (outs[3] - outs[1]).abs()
    print(f"; max abs diff {float(d.max()):.3e} at {tuple(int(v) for v in torch.unravel_index(d.argmax(), d.shape))}")
# timing
g = torch.from_numpy(synth.normal_embeds((1024, 128), 99)).to(dev)
for mode in (2, 3):
    m.set_conv_mode(mode)
    with torch.no_grad():
        for _ in range(2):
            m.recover(g)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            m.recover(g)
        b.record()
        torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"mode {mode}: recover 1024 frames {ms:.3f} ms = {ms * 1e3 / 1024:.2f} us/frame = {234.75e6 * 1024 / ms / 1e9:.1f} TFLOP/s")
