"""KV-window streaming probe (GPU only): the memory side of the cache-mode rollout without its arithmetic.
python scripts/probe_window.py [--traj 65536] [--ctx 32] [--layers 4] [--steps 64]"""
import argparse, ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "transformer-physx_b200"))
from trphysx_b200 import _native as N

ap = argparse.ArgumentParser()
ap.add_argument("--traj", type=int, default=65536)
ap.add_argument("--ctx", type=int, default=32)
ap.add_argument("--layers", type=int, default=4)
ap.add_argument("--steps", type=int, default=64)
a = ap.parse_args()
lib = N.lib()
for warps in (4, 8, 12, 16, 24, 32):
    for rows in (4, 8, 16):
        for pf in (0, 1):
            g, ms = ctypes.c_float(), ctypes.c_float()
            N.check(lib.trpx_probe_window_stream(0, a.traj, a.ctx, a.layers, a.steps, warps, rows, pf,
                                                 ctypes.byref(g), ctypes.byref(ms)))
            print(json.dumps({"n_traj": a.traj, "n_ctx": a.ctx, "n_layer": a.layers, "steps": a.steps, "warps_per_cta": warps,
                              "rows_in_flight": rows, "l2_prefetch": pf, "GBps": round(g.value, 1), "ms": round(ms.value, 3),
                              "resident_window_MB": round(148 * warps * 4 * a.layers * 2 * a.ctx * 128 / 1e6, 1)}), flush=True)
