"""BASELINE config 4: Cylinder flow — CylinderEmbedding conv enc/dec (64x128 grid, viscosity input) +
PhysformerGPT2 (n_embd=128, 6 layers, n_ctx=16), 400-step rollout.  One JSON line per batch size.  GPU only."""
import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "transformer-physx_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from tests import synth
from oracle import trpx_oracle as O  # noqa: E402
from tests.util import make_cylinder, make_gpt2  # noqa: E402

DEC_FLOPS = 234_749_952      # SURVEY.md Appendix B, per frame
ENC_FLOPS = 16_809_984


def ev():
    return torch.cuda.Event(enable_timing=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="1,6,64,512")
    ap.add_argument("--horizon", type=int, default=400)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--cpu", action="store_true", help="also time the oracle port on the host (B=4, short horizon)")
    ap.add_argument("--conv-mode", type=int, default=2)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    cfg = synth.make_config("cylinder")
    emb = make_cylinder(cfg, 2030, dev)
    if args.chunk:
        emb.set_chunk(args.chunk)
    emb.set_conv_mode(args.conv_mode)
    tr = make_gpt2(cfg, 2031, False, dev)
    S = args.horizon
    for B in [int(b) for b in args.batches.split(",")]:
        x_np, v_np = synth.cylinder_fields(B, seed=0)
        x, visc = torch.from_numpy(x_np).to(dev), torch.from_numpy(v_np).to(dev)
        res = {}
        with torch.no_grad():
            for rep in range(3):
                e = [ev() for _ in range(4)]
                e[0].record()
                g0 = emb.embed(x, visc)
                e[1].record()
                out = tr.generate(g0.unsqueeze(1), max_length=S + 1, use_cache=True, return_presents=False)[0]
                e[2].record()
                states = emb.recover(out.view(B * (S + 1), cfg.n_embd))
                e[3].record()
                torch.cuda.synchronize()
                res = {"embed_ms": e[0].elapsed_time(e[1]), "generate_ms": e[1].elapsed_time(e[2]),
                       "recover_ms": e[2].elapsed_time(e[3]), "total_ms": e[0].elapsed_time(e[3])}
                del states
        frames = B * (S + 1)
        line = {"workload": f"cylinder B={B} horizon={S}", "B": B, "S": S, **{k: round(v, 3) for k, v in res.items()},
                "traj_steps_per_s_e2e": round(B * S / res["total_ms"] * 1e3, 1),
                "traj_steps_per_s_generate": round(B * S / res["generate_ms"] * 1e3, 1),
                "recover_frames_per_s": round(frames / res["recover_ms"] * 1e3, 1),
                "recover_tflops": round(frames * DEC_FLOPS / res["recover_ms"] / 1e9, 2),
                "embed_tflops": round(B * ENC_FLOPS / res["embed_ms"] / 1e9, 3),
                "conv_mode": args.conv_mode, "rollout_launch": tr.last_launch()}
        print(json.dumps(line), flush=True)
    if args.cpu:
        Bc, Sc = 4, 40
        sd_e = O.to_torch(synth.cylinder_embedding_weights(cfg, 2030))
        sd_t = O.to_torch(synth.gpt2_weights(cfg, 2031, stress=False))
        x_np, v_np = synth.cylinder_fields(Bc, seed=0)
        xc, vc = torch.from_numpy(x_np), torch.from_numpy(v_np)
        torch.set_num_threads(os.cpu_count())
        with torch.no_grad():
            t0 = time.perf_counter()
            g0 = O.cylinder_embed(sd_e, cfg, xc, vc)
            out, _ = O.gpt2_generate(sd_t, cfg, g0.unsqueeze(1), Sc + 1, use_cache=True)
            st = O.cylinder_recover(sd_e, cfg, out.reshape(-1, cfg.n_embd))
            dt = time.perf_counter() - t0
        print(json.dumps({"cpu_baseline": {"value": Bc * Sc / dt, "unit": "trajectory-steps/s", "cores": os.cpu_count(),
                                           "kind": "port", "sample": f"cylinder B={Bc} horizon={Sc} embed+generate+recover"}}))


if __name__ == "__main__":
    main()
