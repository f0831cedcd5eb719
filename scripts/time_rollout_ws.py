"""Warp-specialised streaming rollout (variant 11) against the tensor-core kernel it shares its arithmetic with
(variant 6: expected bit-identical) and the production window kernel (variant 2): parity on a small ragged batch
(states and presents), then timing at full size.
    python scripts/time_rollout_ws.py [kind] [B] [S] [variants]"""
import sys
import torch
sys.path.insert(0, "transformer-physx_b200")
sys.path.insert(0, ".")
from tests import synth
from tests.util import make_gpt2

kind = sys.argv[1] if len(sys.argv) > 1 else "rossler"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
S = int(sys.argv[3]) if len(sys.argv) > 3 else 256
variants = [int(v) for v in (sys.argv[4].split(",") if len(sys.argv) > 4 else ["2", "11"])]
dev = torch.device("cuda:0")
cfg = synth.make_config(kind)
m = make_gpt2(cfg, 903, True, dev, sigma=0.1)
for b, s in ((37, 3 * cfg.n_ctx + 5), (2500, 40)):
    xs = torch.nn.functional.layer_norm(torch.randn(b, 1, cfg.n_embd, device=dev), (cfg.n_embd,))
    outs = {}
    for v in (6, 11, 2):
        m.set_rollout_tuning(variant=v)
        outs[v] = m.generate(xs, max_length=s + 1, use_cache=True)
        assert m.last_launch()["variant"] == v, m.last_launch()
    torch.cuda.synchronize()
    a, c, d = outs[6][0], outs[11][0], outs[2][0]
    print(f"parity B={b} S={s}: 11 vs 6 max-abs {float((a - c).abs().max()):.3e} (bit-equal: {bool(torch.equal(a, c))}), "
          f"11 vs 2 rel-L2 {float((c.double() - d.double()).norm() / d.double().norm()):.3e}, presents equal: "
          f"{all(bool(torch.equal(x, y)) for x, y in zip(outs[6][1], outs[11][1]))}", flush=True)
x0 = torch.nn.functional.layer_norm(torch.randn(B, 1, cfg.n_embd, device=dev), (cfg.n_embd,))
for variant in variants:
    m.set_rollout_tuning(variant=variant)
    for _ in range(2):
        out = m.generate(x0, max_length=S + 1, use_cache=True, return_presents=False)[0]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        out = m.generate(x0, max_length=S + 1, use_cache=True, return_presents=False)[0]
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{kind} B={B} S={S} variant {variant} {m.last_launch()}: {ms:.2f} ms = {B * S / ms / 1e3:.1f} M traj-steps/s", flush=True)
