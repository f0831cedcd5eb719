"""BASELINE config 5: Lorenz rollout throughput sweep, batch 1K-1M trajectories x horizon 64-1024, vs the host-CPU
reference (oracle port; measured up to B*S <= ~4M, larger points not measured).  One JSON line per point."""
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
from bench import algorithmic_flops_per_traj_step, build_modules  # noqa: E402
from tests import synth
from oracle import trpx_oracle as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="1024,4096,16384,65536,262144,1048576")
    ap.add_argument("--horizons", default="64,128,256,512,1024")
    ap.add_argument("--cpu-budget", type=float, default=4e6)
    ap.add_argument("--max-units", type=float, default=3e8, help="skip GPU points with B*S above this")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    cfg, emb, tr = build_modules("lorenz", dev)
    sd_t = O.to_torch(synth.gpt2_weights(cfg, 2025, stress=False))
    E = cfg.n_embd
    for B in [int(b) for b in args.batches.split(",")]:
        g0 = torch.nn.functional.layer_norm(torch.randn(B, 1, E, device=dev), (E,))
        for S in [int(s) for s in args.horizons.split(",")]:
            if B * S > args.max_units or B * (S + 1) * E * 4 > 120e9:
                continue
            row = {"B": B, "S": S}
            for uc in (True, False):
                with torch.no_grad():
                    out = tr.generate(g0, max_length=S + 1, use_cache=uc, return_presents=False)
                    del out
                    ts = []
                    for _ in range(2):
                        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        a.record()
                        out = tr.generate(g0, max_length=S + 1, use_cache=uc, return_presents=False)
                        b.record()
                        torch.cuda.synchronize()
                        ts.append(a.elapsed_time(b))
                        del out
                ms = min(ts)
                key = "cache" if uc else "nocache"
                row[f"gpu_{key}_ms"] = round(ms, 3)
                row[f"gpu_{key}_traj_steps_per_s"] = round(B * S / ms * 1e3)
                row[f"gpu_{key}_tflops"] = round(algorithmic_flops_per_traj_step(cfg, S, uc) * B * S / ms / 1e9, 2)
                row[f"launch_{key}"] = tr.last_launch()
            if B * S <= args.cpu_budget:
                torch.set_num_threads(os.cpu_count())
                x = g0.cpu()
                with torch.no_grad():
                    for uc in (True, False):
                        t0 = time.perf_counter()
                        O.gpt2_generate(sd_t, cfg, x, S + 1, use_cache=uc)
                        dt = time.perf_counter() - t0
                        row[f"cpu_{'cache' if uc else 'nocache'}_traj_steps_per_s"] = round(B * S / dt)
                row["cpu_cores"] = os.cpu_count()
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
