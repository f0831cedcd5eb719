import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "transformer-physx_b200")):
    sys.path.insert(0, p)
import torch
from tests import synth
from tests.util import make_cylinder
dev = torch.device("cuda:0")
mode, chunk, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
cfg = synth.make_config("cylinder")
cyl = make_cylinder(cfg, 3, dev)
cyl.set_conv_mode(mode); cyl.set_chunk(chunk)
g = torch.randn(n, 128, device=dev)
with torch.no_grad():
    for _ in range(2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); cyl.recover(g); b.record(); torch.cuda.synchronize()
        print(f"mode {mode} chunk {chunk} N {n}: {a.elapsed_time(b):.3f} ms  {n*234.75e6/a.elapsed_time(b)/1e9:.2f} TFLOP/s")
