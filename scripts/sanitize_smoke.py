"""Tiny invocations of every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):
    compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py
GPU only; results are not checked here (the parity tests do that)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "transformer-physx_b200")):
    sys.path.insert(0, p)
import torch
from tests import synth
from tests.util import make_gpt2, make_lorenz, make_cylinder, make_grayscott

dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "all"

if which in ("all", "e32"):
    for kind in ("rossler", "lorenz"):
        cfg = synth.make_config(kind)
        m = make_gpt2(cfg, 1, True, dev)
        x0 = torch.randn(37, 1, 32, device=dev)
        for variant, g in ((2, 1), (2, 4), (2, 8), (3, 4), (5, 0), (6, 0), (8, 4), (1, 2), (10, 0)):
            m.set_rollout_tuning(variant=variant, traj_per_warp=g)
            for uc in (True, False):
                out = m.generate(x0, max_length=cfg.n_ctx + 9, use_cache=uc)
        m.set_rollout_tuning()
        hid = m(torch.randn(3, cfg.n_ctx, 32, device=dev), use_cache=True, output_attentions=True)
    torch.cuda.synchronize()
    print("e32 ok")
if which in ("all", "cluster"):
    cfg = synth.make_config("cylinder")
    m = make_gpt2(cfg, 2, True, dev)
    for B, g in ((1, 1), (7, 2), (9, 4), (20, 16)):
        m.set_rollout_tuning(variant=7, traj_per_warp=g)
        for uc in (True, False):
            m.generate(torch.randn(B, 1, 128, device=dev), max_length=22, use_cache=uc)
    m.set_rollout_tuning(variant=1)
    m.generate(torch.randn(5, 1, 128, device=dev), max_length=20, use_cache=True)
    torch.cuda.synchronize()
    print("cluster ok")
if which in ("all", "train"):
    from trphysx_b200.optim import FusedTransformerStep, FusedClipAdam, FusedKoopmanStep
    for kind, T in (("rossler", 32), ("cylinder", 16), ("lorenz", 11)):
        cfg = synth.make_config(kind)
        m = make_gpt2(cfg, 3, True, dev).train()
        step = FusedTransformerStep(m, max_norm=0.1)
        x = torch.randn(3, T, cfg.n_embd, device=dev)
        step(x, torch.randn_like(x))
        step(x, torch.randn_like(x))
    cfg = synth.make_config("lorenz")
    emb = make_lorenz(cfg, 4, "lorenz", dev)
    from trphysx_b200.embedding.embedding_lorenz import LorenzEmbeddingTrainer
    head = LorenzEmbeddingTrainer(cfg).to(dev)
    opt = FusedClipAdam(head.parameters(), max_norm=0.1)
    FusedKoopmanStep(head, opt)(torch.randn(24, 6, 3, device=dev))
    torch.cuda.synchronize()
    print("train ok")
if which in ("all", "emb"):
    from trphysx_b200.evaluation import eval_states
    cfg = synth.make_config("lorenz")
    emb = make_lorenz(cfg, 5, "lorenz", dev)
    x = torch.randn(1001, 3, device=dev)
    g = emb.embed(x)
    emb.recover(g); emb.koopmanOperation(g); emb.koopmanOperator
    eval_states(emb, g.view(7, 143, 32), x.view(7, 143, 3))
    ccfg = synth.make_config("cylinder")
    cyl = make_cylinder(ccfg, 6, dev)
    xs = torch.randn(3, 3, 64, 128, device=dev)
    visc = torch.rand(3, 1, device=dev) * 0.01 + 0.003
    gc = cyl.embed(xs, visc)
    for mode in (2, 1, 0):
        cyl.set_conv_mode(mode)
        cyl.recover(gc)
    cyl.koopmanOperation(gc, visc)
    torch.cuda.synchronize()
    print("emb ok")
if which in ("all", "gs"):
    gcfg = synth.make_config("grayscott")
    gs = make_grayscott(gcfg, 8, dev)
    xg = torch.rand(3, 2, 32, 32, 32, device=dev)
    with torch.no_grad():
        outs = gs(xg)
        gs.koopmanOperation(outs[0])
        gs.koopmanOperator
    tr = make_gpt2(gcfg, 9, True, dev)
    tr.generate(outs[0].unsqueeze(1), max_length=6, use_cache=True)
    torch.cuda.synchronize()
    print("gs ok")
if which in ("all", "cyl3"):
    ccfg = synth.make_config("cylinder")
    cyl = make_cylinder(ccfg, 6, dev)
    xs = torch.randn(5, 3, 64, 128, device=dev)
    visc = torch.rand(5, 1, device=dev) * 0.01 + 0.003
    cyl.set_conv_mode(3)
    cyl.recover(cyl.embed(xs, visc))
    torch.cuda.synchronize()
    print("cyl3 ok")
