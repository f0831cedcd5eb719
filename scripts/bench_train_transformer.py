"""Teacher-forced transformer training step: fused forward+backward (csrc/gpt2_train.cu) timed through
PhysformerTrain.forward + loss.backward() + clip_grad_norm_ + Adam, the step body of utils/trainer.py:193-207,244-277.
GPU only.  python scripts/bench_train_transformer.py [--kind lorenz] [--batches 16,64,256,1024] [--cpu]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "transformer-physx_b200")):
    sys.path.insert(0, p)
import torch
from tests import synth
from tests.util import make_gpt2

ap = argparse.ArgumentParser()
ap.add_argument("--kind", default="lorenz")
ap.add_argument("--batches", default="16,64,256,1024")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--cpu", action="store_true", help="also time the oracle port (torch CPU autograd) at the first batch size")
a = ap.parse_args()
from trphysx_b200.transformer import PhysformerTrain
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:                                   # data parallel: one process per GPU, ONE all-reduce of the gradient blob
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
cfg = synth.make_config(a.kind)
T = cfg.n_ctx
for B in [int(b) for b in a.batches.split(",")]:
    m = make_gpt2(cfg, 2025, False, dev)
    head = PhysformerTrain(cfg, m).to(dev).train()
    opt = torch.optim.Adam(head.parameters(), lr=1e-3, weight_decay=1e-10)
    x = torch.randn(B, T, cfg.n_embd, device=dev)
    y = torch.randn(B, T, cfg.n_embd, device=dev)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = head(inputs_embeds=x, labels_embeds=y)[0]
        loss.backward()
        torch.nn.utils.clip_grad_norm_(head.parameters(), 0.1)
        opt.step()
        return loss

    def fwd_bwd_only():
        return m.fused_loss_and_grads(x, y)[0]

    from trphysx_b200.optim import FusedTransformerStep
    m2 = make_gpt2(cfg, 2025, False, dev)
    fstep = FusedTransformerStep(m2.train(), lr=1e-3, weight_decay=1e-10, max_norm=0.1, world_size=world)

    def fused_step():
        return fstep(x, y)

    res = {"kind": a.kind, "B_per_gpu": B, "T": T, "n_gpus": world}
    for name, fn in (("step_ms", step), ("fused_fwd_bwd_ms", fwd_bwd_only), ("fused_step_ms", fused_step)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            l = fn()
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / a.iters
    res["tokens_per_s"] = world * B * T / (res["fused_step_ms"] * 1e-3)
    res["loss_after"] = float(l)
    if a.cpu and B == int(a.batches.split(",")[0]):
        from oracle import trpx_oracle as O
        sd = O.to_torch(synth.gpt2_weights(cfg, 2025, stress=False))
        xc, yc = x.cpu(), y.cpu()
        torch.set_num_threads(os.cpu_count())
        O.gpt2_train_grads(sd, cfg, xc, yc)
        t0 = time.perf_counter()
        for _ in range(5):
            O.gpt2_train_grads(sd, cfg, xc, yc)
        res["cpu_port_fwd_bwd_ms"] = (time.perf_counter() - t0) / 5 * 1e3
        res["cpu_threads"] = os.cpu_count()
    if rank == 0:
        print(json.dumps(res), flush=True)
if world > 1:
    dist.destroy_process_group()
