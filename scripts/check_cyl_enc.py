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
    x, visc = synth.cylinder_fields(N, seed=3)
    x, visc = torch.from_numpy(x), torch.from_numpy(visc)
    ref = O.cylinder_embed(sd, cfg, x, visc)
    outs = {}
    for mode in (1, 3):
        m.set_conv_mode(mode)
        with torch.no_grad():
            outs[mode] = m.embed(x.to(dev), visc.to(dev))
        torch.cuda.synchronize()
    print(f"N={N}: embed mode3 vs oracle {O.rel_l2(outs[3].cpu(), ref):.3e}, mode1 vs oracle {O.rel_l2(outs[1].cpu(), ref):.3e}")
x, visc = synth.cylinder_fields(1024, seed=4)
x, visc = torch.from_numpy(x).to(dev), torch.from_numpy(visc).to(dev)
for mode in (1, 3):
    m.set_conv_mode(mode)
    with torch.no_grad():
        for _ in range(2):
            m.embed(x, visc)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            m.embed(x, visc)
        b.record()
        torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"mode {mode}: embed 1024 frames {ms:.3f} ms = {ms*1e3/1024:.2f} us/frame = {16.81e6*1024/ms/1e9:.1f} TFLOP/s")
