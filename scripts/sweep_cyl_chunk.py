"""Decoder chunk-size sweep: frames per chunk decide whether a layer's output is still in L2 when the next layer reads it.
    python scripts/sweep_cyl_chunk.py [N]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "transformer-physx_b200"))
import torch
from tests import synth
from tests.util import make_cylinder

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
dev = torch.device("cuda:0")
cfg = synth.make_config("cylinder")
m = make_cylinder(cfg, 55, dev)
m.set_conv_mode(3)
g = torch.from_numpy(synth.normal_embeds((N, 128), 99)).to(dev)
with torch.no_grad():
    for chunk in (0, 32, 64, 96, 128, 148, 192, 256, 384):
        if chunk:
            m.set_chunk(chunk)
        for _ in range(2):
            m.recover(g)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            m.recover(g)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        print(f"chunk {chunk or 'default'}: recover {N} frames {ms:.3f} ms = {ms * 1e3 / N:.3f} us/frame = {234.75e6 * N / ms / 1e9:.1f} TFLOP/s")
