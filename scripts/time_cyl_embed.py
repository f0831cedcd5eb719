"""Times CylinderEmbedding.embed (encoder) for one conv mode; used under ncu for the per-kernel launch list."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "transformer-physx_b200"))
import torch
from tests import synth
from tests.util import make_cylinder
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 3
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dev = torch.device("cuda:0")
cfg = synth.make_config("cylinder")
m = make_cylinder(cfg, 55, dev)
m.set_conv_mode(mode)
x, visc = synth.cylinder_fields(N, seed=4)
x, visc = torch.from_numpy(x).to(dev), torch.from_numpy(visc).to(dev)
with torch.no_grad():
    for _ in range(2):
        m.embed(x, visc)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        m.embed(x, visc)
    b.record()
    torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
print(f"mode {mode}: embed {N} frames {ms:.3f} ms = {ms*1e3/N:.2f} us/frame = {16.81e6*N/ms/1e9:.1f} TFLOP/s")
