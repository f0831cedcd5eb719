"""Does a fresh box speed up over the first seconds of load?  Times the headline rollout launch repeatedly and samples
the SM clock / power between launches."""
import subprocess, sys, time
sys.path.insert(0, "transformer-physx_b200"); sys.path.insert(0, ".")
import torch
from bench import build_modules
dev = torch.device("cuda:0")
cfg, emb, tr = build_modules("rossler", dev)
g0 = torch.nn.functional.layer_norm(torch.randn(65536, 1, cfg.n_embd, device=dev), (cfg.n_embd,))
t0 = time.time()
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 60):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    tr.generate(g0, max_length=257, use_cache=True, return_presents=False)
    b.record()
    torch.cuda.synchronize()
    if i % 4 == 0:
        q = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu,clocks_throttle_reasons.active",
                            "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
        print(f"t={time.time() - t0:6.1f}s launch {i:3d}: {a.elapsed_time(b):7.2f} ms   [{q}]", flush=True)
