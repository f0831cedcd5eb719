"""Teacher-forced forward (PhysformerGPT2.forward) at B x T tokens: tensor-core kernel vs the block-GEMV kernel."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "transformer-physx_b200"))
import torch
from tests import synth
from tests.util import make_gpt2
from oracle import trpx_oracle as O
dev = torch.device("cuda:0")
B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 64)
cfg = synth.make_config("lorenz")
m = make_gpt2(cfg, 17, True, dev, sigma=0.2)
x = torch.from_numpy(synth.normal_embeds((B, T, 32), 3)).to(dev)
res = {}
for variant, name in ((1, "forward_kernel (block GEMV)"), (0, "forward_e32_tc_kernel (3xTF32 mma.sync)")):
    m.set_rollout_tuning(variant=variant)
    with torch.no_grad():
        for _ in range(2):
            out = m(x, use_cache=False)[0]
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            out = m(x, use_cache=False)[0]
        b.record()
        torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    flops = B * T * 4 * (24 * 32 * 32 + 4 * (T / 2) * 32) + B * T * 2 * 32 * 32
    res[variant] = out
    print(f"{name}: {ms:.3f} ms for {B} x {T} tokens = {B*T/ms/1e3:.1f} M tokens/s, {flops/ms/1e9:.1f} TFLOP/s")
print("tc vs block-GEMV rel-L2:", O.rel_l2(res[0].cpu(), res[1].cpu()))
ref = O.gpt2_forward(O.to_torch(synth.gpt2_weights(cfg, 17, stress=True, stress_sigma=0.2)), cfg, x[:8].cpu(), use_cache=False)[0]
print("tc vs oracle (8 sequences) rel-L2:", O.rel_l2(res[0][:8].cpu(), ref))
