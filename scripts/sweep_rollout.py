"""Time the rollout kernel alone over tuning knobs (warps per CTA, trajectories per warp) and batch sizes.
Prints one JSON line per point.  GPU only."""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "transformer-physx_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from bench import algorithmic_flops_per_traj_step, build_modules  # noqa: E402


def time_rollout(tr, g0, S, use_cache, reps=3):
    for _ in range(1):
        tr.generate(g0, max_length=S + 1, use_cache=use_cache, return_presents=False)
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        tr.generate(g0, max_length=S + 1, use_cache=use_cache, return_presents=False)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kinds", default="rossler,lorenz")
    ap.add_argument("--batches", default="8192")
    ap.add_argument("--horizon", type=int, default=256)
    ap.add_argument("--warps", default="2,4,6,8,11")
    ap.add_argument("--groups", default="8,4")
    ap.add_argument("--variants", default="0")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    for kind in args.kinds.split(","):
        cfg, emb, tr = build_modules(kind, dev)
        for B in [int(b) for b in args.batches.split(",")]:
            g0 = torch.nn.functional.layer_norm(torch.randn(B, 1, cfg.n_embd, device=dev), (cfg.n_embd,))
            S = args.horizon
            for variant in [int(v) for v in args.variants.split(",")]:
                for G in [int(g) for g in args.groups.split(",")]:
                    for W in [int(w) for w in args.warps.split(",")]:
                        tr.set_rollout_tuning(variant=variant, warps_per_cta=W, traj_per_warp=0 if variant in (5, 6) else G)
                        for uc in (True, False):
                            ms = time_rollout(tr, g0, S, uc)
                            fl = algorithmic_flops_per_traj_step(cfg, S, uc) * B * S
                            print(json.dumps({"kind": kind, "B": B, "S": S, "use_cache": uc, "variant": variant,
                                              "G": G, "W": W, "ms": round(ms, 3),
                                              "traj_steps_per_s": round(B * S / ms * 1e3),
                                              "tflops": round(fl / ms / 1e9, 2), "launch": tr.last_launch()}),
                                  flush=True)


if __name__ == "__main__":
    main()
