#!/usr/bin/env python
"""Benchmark of the rollout hot path (BASELINE.json metric: rollout trajectory-steps/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is ONE pass of the hot path over one batch of synthetic initial states:
    embed(x0) -> generate(max_length = horizon+1, use_cache=True) -> recover(all tokens)
Workload at N=1: BASELINE config[1] — the 65,536-trajectory batched Rossler rollout, horizon 256 (it fits one GPU).
N>1: `value` is weak scaling (every GPU runs its own 65,536-trajectory shard, no collective on the data path); the
literal config[1] (65,536 trajectories in total, 65,536/N per GPU) is reported beside it under "strong".

  value    trajectory-steps/s, whole job, inputs resident in HBM (device-timed with CUDA events, max over ranks)
  e2e      same metric through the public module API with HOST buffers: pinned x0 -> H2D, ..., states -> D2H
           (generate + recover run per slice of one round of resident trajectories; a slice's states travel to the
           pinned host buffer on a copy stream while the next slice decodes; the step ends when the last state has
           reached the host)
  roofline the rollout kernel alone (CUDA events around it): KV-window stream bytes / duration vs the measured HBM peak
  configs  the other BASELINE configs, each device-timed in this run with its own roofline record: Lorenz 256 x 128
           (config 0), Cylinder 400-step rollouts (config 3), the Koopman training step with its all-reduce (config 2),
           three points of the Lorenz sweep (config 4)
  cpu_baseline   the reference's own modules (baseline/_ref, unmodified) on the host cores, bounded sample, rank 0 only;
           the oracle port is used only if the reference install is absent

--impl reference times the CPU path alone and prints the same line with "impl": "reference".
"""
import argparse
import contextlib
import io
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "transformer-physx_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "rollout_trajectory_steps_per_sec"
UNIT = "trajectory-steps/s"
CYL_DEC_FLOPS = 234_749_952      # SURVEY.md section 8d / Appendix B, per frame
CYL_ENC_FLOPS = 16_809_984

WORKLOADS = {
    # name: (config kind, per-GPU trajectories, horizon)
    "rossler": ("rossler", 65536, 256),     # BASELINE config[1]: the 65,536-trajectory batched Rossler rollout
    "rossler8k": ("rossler", 8192, 256),    # its 1/8 shard
    "lorenz": ("lorenz", 8192, 128),        # BASELINE config[0] architecture at a GPU-sized batch
    "lorenz256": ("lorenz", 256, 128),      # BASELINE config[0] as stated (256 trajectories, 128 steps)
}


def algorithmic_flops_per_traj_step(cfg, horizon, use_cache=True):
    """n_layer*(24 E^2 + 4 L E) + 2 E^2 with L keys in the window, averaged over the horizon."""
    E, nl, C = cfg.n_embd, cfg.n_layer, cfg.n_ctx
    tot = 0.0
    for s in range(horizon):
        L = min(s + 1, C) if use_cache else 1
        tot += nl * (24 * E * E + 4 * L * E) + 2 * E * E
    return tot / horizon


def window_bytes_per_traj(cfg, horizon):
    """HBM traffic model of the cache-mode rollout for ONE trajectory over the horizon: every step reads the
    KV window of every layer (L keys x 2 x E floats), appends one key/value row per layer and writes one token.
    This is SURVEY.md section 8d's "if the window spills" figure: with >= 4 resident warps per SM needed to hide
    latency, the in-flight windows (32-64 KiB per trajectory) exceed the on-chip storage by an order of magnitude
    (DESIGN.md section 3), so the window stream IS the kernel's necessary traffic; ncu confirms it (profiles/)."""
    E, nl, C = cfg.n_embd, cfg.n_layer, cfg.n_ctx
    tot = 0
    for s in range(horizon):
        L = min(s + 1, C)
        tot += nl * (2 * L * E * 4 + 2 * E * 4) + E * 4
    return tot


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_modules(kind, device):
    from tests import synth
    from tests.util import make_gpt2, make_lorenz
    cfg = synth.make_config(kind)
    emb = make_lorenz(cfg, 2024, "lorenz" if kind == "lorenz" else "rossler", device)
    tr = make_gpt2(cfg, 2025, False, device)        # random init like the reference's _init_weights
    return cfg, emb, tr


def initial_states(kind, n, seed):
    from tests import synth
    gen = synth.lorenz_trajectories if kind == "lorenz" else synth.rossler_trajectories
    return torch.from_numpy(gen(n, 1, seed=seed)[:, 0].copy())


def ev():
    return torch.cuda.Event(enable_timing=True)


# ----------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own modules (baseline/_ref) or, if that install is absent, the oracle port
# ----------------------------------------------------------------------------------------------------------
def cpu_reference_pass(kind, B, S, use_cache=True):
    """Returns (run, how): run() is one CPU pass embed -> generate -> recover on B trajectories over S steps with the
    bench's weights; `how` is "reference" (unmodified trphysx modules, stock generate() loop) or "port" (oracle)."""
    from tests import synth
    cfg_ns = synth.make_config(kind)
    w_e = synth.lorenz_embedding_weights(cfg_ns, 2024)
    w_t = synth.gpt2_weights(cfg_ns, 2025, stress=False)
    x0 = initial_states(kind, B, 7)
    try:
        from oracle import ref_shim
        if not ref_shim.available():
            raise ImportError("no reference install")
        ref_shim.load()
        from trphysx.config.configuration_phys import PhysConfig
        from trphysx.embedding import LorenzEmbedding      # the Rossler module (examples/) has the same arithmetic
        from trphysx.transformer import PhysformerGPT2
        cfg = PhysConfig(**vars(cfg_ns))
        emb = LorenzEmbedding(cfg).eval()
        tr = PhysformerGPT2(cfg, kind).eval()
        emb.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in w_e.items()}, strict=True)
        tr.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in w_t.items()}, strict=True)

        def run():
            with torch.no_grad():
                g0 = emb.embed(x0)
                out = tr.generate(g0.unsqueeze(1), max_length=S + 1, use_cache=use_cache)[0]
                return emb.recover(out.reshape(-1, cfg.n_embd))
        return run, "reference"
    except Exception as exc:          # noqa: BLE001 — any import problem falls back to the port, and says so
        print(f"[bench] reference install unusable ({exc!r}); timing the oracle port instead", file=sys.stderr)
        from oracle import trpx_oracle as O
        sd_e, sd_t = O.to_torch(w_e), O.to_torch(w_t)

        def run():
            with torch.no_grad():
                g0 = O.lorenz_embed(sd_e, cfg_ns, x0)
                out, _ = O.gpt2_generate(sd_t, cfg_ns, g0.unsqueeze(1), S + 1, use_cache=use_cache)
                return O.lorenz_recover(sd_e, cfg_ns, out.reshape(-1, cfg_ns.n_embd))
        return run, "port"


def time_cpu(kind, B, S, threads, use_cache=True, budget_s=12.0, min_reps=2, max_reps=20):
    torch.set_num_threads(threads)
    run, how = cpu_reference_pass(kind, B, S, use_cache)
    run()
    t0 = time.perf_counter()
    reps = 0
    while reps < min_reps or (time.perf_counter() - t0 < budget_s and reps < max_reps):
        run()
        reps += 1
    dt = (time.perf_counter() - t0) / reps
    return B * S / dt, how, reps


def cpu_sample_text(kind, B, S, how, use_cache=True):
    src = ("unmodified reference modules from baseline/_ref (LorenzEmbedding + PhysformerGPT2.generate, stock loop)"
           if how == "reference" else "oracle port (torch CPU fp32)")
    return f"{kind} B={B} horizon={S} use_cache={use_cache} embed+generate+recover, {src}"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind, _, horizon = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    B, S = args.cpu_batch, min(args.cpu_horizon, horizon)
    torch.set_num_threads(threads)
    run, how = cpu_reference_pass(kind, B, S, True)
    for _ in range(max(1, min(args.warmup, 2))):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    dt = (time.perf_counter() - t0) / args.steps
    value = B * S / dt
    sample = cpu_sample_text(kind, B, S, how)
    variants = {}
    if not args.no_variants:
        # the other reference modes, smaller samples (bounded): default use_cache=False, and one thread
        v, _, _ = time_cpu(kind, B, S, threads, use_cache=False, budget_s=6.0)
        variants["use_cache_false_all_threads"] = {"value": v, "sample": cpu_sample_text(kind, B, S, how, False)}
        Bs = max(64, B // 8)
        v, _, _ = time_cpu(kind, Bs, S, 1, use_cache=True, budget_s=6.0, min_reps=1)
        variants["use_cache_true_1_thread"] = {"value": v, "sample": cpu_sample_text(kind, Bs, S, how, True)}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "sample": sample,
                   "note": "a rate on a bounded batch of the same model, horizon and cache mode as the GPU arm"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": how, "sample": sample, "variants": variants},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_name(args):
    kind, B, S = WORKLOADS[args.workload]
    B = args.batch or B
    S = args.horizon or S
    return f"{kind}_rollout B={B}/GPU horizon={S} use_cache=True (embed+generate+recover)"


# ----------------------------------------------------------------------------------------------------------
# the other BASELINE configs, device-timed in the same run
# ----------------------------------------------------------------------------------------------------------
def timed_ms(fn, reps, warm=1, flush=None):
    with torch.no_grad():
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            if flush is not None:
                flush.fill_(3.0)
            a, b = ev(), ev()
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def record_lorenz(dev, kind, B, S, flush, peaks, fp32_peak, e2e=True):
    """Lorenz-architecture rollout at (B, S): device pass, kernel alone, optionally the host-buffer pass."""
    from trphysx_b200 import _native as N
    from trphysx_b200.evaluation import rollout_to_host
    cfg, emb, tr = build_modules(kind, dev)
    E = cfg.n_embd
    x0_host = initial_states(kind, B, 4321).pin_memory()
    x0 = x0_host.to(dev)

    def device_pass():
        g0 = emb.embed(x0)
        out = tr.generate(g0.unsqueeze(1), max_length=S + 1, use_cache=True, return_presents=False)[0]
        return emb.recover(out.view(B * (S + 1), E))

    big = B * S >= (1 << 22)
    ms = timed_ms(device_pass, 3 if big else 7, 2, flush if big else None)
    with torch.no_grad():
        g0 = emb.embed(x0).unsqueeze(1).contiguous()
    k_ms = timed_ms(lambda: tr.generate(g0, max_length=S + 1, use_cache=True, return_presents=False), 3 if big else 7, 1,
                    flush if big else None)
    launch = tr.last_launch()
    rec = {"workload": f"{kind} B={B} horizon={S} use_cache=True", "ms_per_pass": ms, "traj_steps_per_s": B * S / (ms * 1e-3),
           "generate_ms": k_ms, "rollout_launch": launch}
    wbytes = window_bytes_per_traj(cfg, S) * B
    flops = algorithmic_flops_per_traj_step(cfg, S) * B * S
    hbm = peaks.get("hbm_gbs", 6650.0)
    gbs, tf = wbytes / (k_ms * 1e-3) / 1e9, flops / (k_ms * 1e-3) / 1e12
    sm_fill = launch["grid"] * launch["block"] / 32.0 / (148 * 8)
    if launch["variant"] == 10:
        # KV window on chip (tensor memory): HBM sees the tokens only; the launch is a latency chain of tile phases
        tok = B * S * cfg.n_embd * 4
        rec["roofline"] = {"bound": "latency", "achieved": tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": tf / fp32_peak,
                           "hbm_bytes_model": tok, "hbm_GBps": tok / (k_ms * 1e-3) / 1e9,
                           "note": f"rollout_e32t_kernel: KV window in tensor memory, {launch['grid']} CTAs x {launch['group']} resident "
                                   "trajectories, 8 warps per tile; HBM traffic = tokens only (algorithmic minimum); a small batch "
                                   "cannot fill the machine, so the step latency (4 layers x ~6 barrier phases) is what counts"}
    else:
        rec["roofline"] = {"bound": "hbm" if B >= 4096 else "latency", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                           "fp32_view": {"achieved": tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": tf / fp32_peak},
                           "note": ("KV-window stream model (bench.py window_bytes_per_traj)" if B >= 4096 else
                                    f"{launch['grid']} CTAs x {launch['block'] // 32} warps = {sm_fill:.2f} of the 148 x 8 warp slots: "
                                    "a small batch cannot fill the machine, the launch is latency bound")}
    if e2e:
        host = torch.empty((B, S + 1, 3), dtype=torch.float32).pin_memory()
        cs = torch.cuda.Stream(device=dev)
        e_ms = timed_ms(lambda: rollout_to_host(emb, tr, x0_host, S + 1, host, device=dev, use_cache=True,
                                                chunks=4 if B >= 4096 else 1, copy_stream=cs), 3 if big else 7, 1,
                        flush if big else None)
        rec["e2e"] = {"value": B * S / (e_ms * 1e-3), "unit": UNIT, "ms_per_pass": e_ms, "h2d_bytes": B * 12,
                      "d2h_bytes": B * (S + 1) * 12}
    return rec


def record_cylinder(dev, B, S, umma_peak, host_e2e):
    """BASELINE config 3 (Cylinder 400-step rollout): embed -> generate -> recover; decoder roofline against the measured
    tcgen05 TF32 rate / 3 (error-compensated 3xTF32 product)."""
    from tests import synth
    from tests.util import make_cylinder, make_gpt2
    from trphysx_b200.evaluation import rollout_to_host
    cfg = synth.make_config("cylinder")
    emb = make_cylinder(cfg, 2030, dev)
    tr = make_gpt2(cfg, 2031, False, dev)
    x_np, v_np = synth.cylinder_fields(B, seed=0)
    x_host, v_host = torch.from_numpy(x_np).pin_memory(), torch.from_numpy(v_np).pin_memory()
    x, visc = x_host.to(dev), v_host.to(dev)
    res = {}
    with torch.no_grad():
        for _ in range(3 if B <= 64 else 2):
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
    dec_tf = frames * CYL_DEC_FLOPS / (res["recover_ms"] * 1e-3) / 1e12
    rec = {"workload": f"cylinder B={B} horizon={S} use_cache=True (embed+generate+recover)", **res,
           "traj_steps_per_s": B * S / (res["total_ms"] * 1e-3), "rollout_launch": tr.last_launch(),
           "recover_us_per_frame": res["recover_ms"] * 1e3 / frames,
           "roofline": {"bound": "tensor", "kernel": "cyl_umma_conv_kernel x4 (tcgen05 kind::tf32, 3 passes) + cyl_dec0_kernel",
                        "achieved": dec_tf, "peak": umma_peak / 3.0, "unit": "TFLOP/s", "frac": dec_tf / (umma_peak / 3.0),
                        "flops_per_frame": CYL_DEC_FLOPS,
                        "peak_source": "tcgen05 TF32 rate measured in this run (trpx_probe_umma_tf32) / 3 for the "
                                       "error-compensated 3xTF32 product"}}
    if host_e2e:
        host = torch.empty((B, S + 1, 3, 64, 128), dtype=torch.float32).pin_memory()
        cs = torch.cuda.Stream(device=dev)
        e_ms = timed_ms(lambda: rollout_to_host(emb, tr, x_host, S + 1, host, device=dev, use_cache=True,
                                                chunks=4 if B >= 4 else 1, copy_stream=cs, visc_host=v_host), 2, 1)
        rec["e2e"] = {"value": B * S / (e_ms * 1e-3), "unit": UNIT, "ms_per_pass": e_ms, "h2d_bytes": x_host.numel() * 4 + B * 4,
                      "d2h_bytes": host.numel() * 4}
    return rec


def record_koopman_train(dev, world, rank):
    """BASELINE config 2: Lorenz Koopman training step, states [512,16,3] per GPU: fused fwd+bwd -> (N>1: one NCCL
    all-reduce of the 144,768-byte flat gradient) -> clip + Adam."""
    from tests import synth
    from tests.util import torch_sd
    from trphysx_b200.config import LorenzConfig
    from trphysx_b200.embedding import LorenzEmbeddingTrainer
    from trphysx_b200.optim import FusedClipAdam, FusedKoopmanStep
    cfg_ns = synth.make_config("lorenz")
    head = LorenzEmbeddingTrainer(LorenzConfig())
    head.embedding_model.load_state_dict(torch_sd(synth.lorenz_embedding_weights(cfg_ns, 77)))
    head.to(dev)
    B, T = 512, 16
    states = torch.from_numpy(synth.lorenz_trajectories(B, T - 1, seed=100 + rank)).to(dev)
    opt = FusedClipAdam(head.parameters(), lr=1e-3, weight_decay=1e-8, max_norm=0.1)
    fused = FusedKoopmanStep(head, opt, world)
    for _ in range(5):
        fused(states)
    torch.cuda.synchronize()
    K = 30
    a, b = ev(), ev()
    a.record()
    for _ in range(K):
        loss = fused(states)[0]
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / K
    flops = 3 * 107e3 * B * T * 2           # ~3 x 107 kFLOP per (b, t), forward + backward ~ 2x
    return {"workload": f"lorenz_koopman_train_step states=[{B},{T},3]/GPU fwd+bwd{'+allreduce' if world > 1 else ''}+clip+adam",
            "ms_per_step": ms, "samples_per_s": world * B / (ms * 1e-3), "loss": float(loss),
            "roofline": {"bound": "latency", "achieved": flops / (ms * 1e-3) / 1e12, "unit": "TFLOP/s",
                         "note": "8 launches over ~0.1 GFLOP of tiny tensors (+ a 145 KB all-reduce at N>1): launch/latency "
                                 "bound by construction (SURVEY 8d K7)"}}, ms


def record_forward(dev, fp32_peak):
    """PhysformerGPT2.forward teacher-forced over 4,096 Lorenz sequences of 64 tokens: the tensor-core kernel
    (3xTF32 mma.sync) beside the block-GEMV kernel it replaces."""
    from tests import synth
    from tests.util import make_gpt2
    cfg = synth.make_config("lorenz")
    m = make_gpt2(cfg, 17, False, dev)
    B, T = 4096, 64
    x = torch.randn(B, T, cfg.n_embd, device=dev)
    flops = B * T * cfg.n_layer * (24 * 32 * 32 + 4 * (T / 2) * 32) + B * T * 2 * 32 * 32
    out = {"workload": f"lorenz forward B={B} T={T} (teacher forced, use_cache=False)"}
    for variant, key in ((0, "forward_e32_tc_kernel"), (1, "forward_kernel")):
        m.set_rollout_tuning(variant=variant)
        ms = timed_ms(lambda: m(x, use_cache=False), 5, 2)
        out[key] = {"ms": ms, "tokens_per_s": B * T / (ms * 1e-3), "tflops": flops / (ms * 1e-3) / 1e12}
    tf = out["forward_e32_tc_kernel"]["tflops"]
    out["roofline"] = {"bound": "fp32", "achieved": tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": tf / fp32_peak,
                       "note": "algorithmic FLOPs (n_layer*(24E^2 + 4*(T/2)*E) + 2E^2 per token) against the FP32 FMA peak measured "
                               "in this run; the dense part runs as 3xTF32 mma.sync, the causal attention on the FP32 pipe"}
    out["speedup_vs_block_gemv"] = out["forward_kernel"]["ms"] / out["forward_e32_tc_kernel"]["ms"]
    return out


def record_transformer_train(dev, mma_peak):
    """Teacher-forced transformer training step: fused forward + backward of PhysformerTrain.forward for 1,024
    Lorenz sequences of 64 tokens; the tensor-core kernel (every contraction a 3xTF32 mma.sync tile product) beside the
    FP32 SIMT kernel it replaces (TRPX_TRAIN_SIMT=1)."""
    from tests import synth
    from tests.util import make_gpt2
    cfg = synth.make_config("lorenz")
    m = make_gpt2(cfg, 19, False, dev)
    B, T = 1024, 64
    x = torch.randn(B, T, cfg.n_embd, device=dev)
    y = torch.randn(B, T, cfg.n_embd, device=dev)
    fwd = B * T * cfg.n_layer * (24 * 32 * 32 + 4 * (T / 2) * 32) + B * T * 2 * 32 * 32
    flops = 3 * fwd                                     # forward + the two backward products of every contraction
    out = {"workload": f"lorenz transformer training step B={B} T={T}: fused forward + backward (loss and all gradients)"}
    for env, key in (("0", "train_e32_tc_kernel"), ("1", "train_fwd_bwd_kernel")):
        os.environ["TRPX_TRAIN_SIMT"] = env
        ms = timed_ms(lambda: m.fused_loss_and_grads(x, y), 5, 2)
        out[key] = {"ms": ms, "tokens_per_s": B * T / (ms * 1e-3), "tflops": flops / (ms * 1e-3) / 1e12}
    os.environ["TRPX_TRAIN_SIMT"] = "0"
    tf = out["train_e32_tc_kernel"]["tflops"]
    out["roofline"] = {"bound": "tensor", "achieved": tf, "peak": mma_peak / 3.0, "unit": "TFLOP/s", "frac": tf / (mma_peak / 3.0),
                       "peak_source": "measured mma.sync TF32 rate in this run (trpx_probe_tf32_mma) / 3 for the error-compensated "
                                      "3xTF32 product",
                       "note": "algorithmic FLOPs = 3 x forward (n_layer*(24E^2 + 4*(T/2)*E) + 2E^2 per token); the kernel is a "
                               "chain of ~35 small tile products per layer and sequence separated by CTA barriers, two sequences "
                               "per SM: latency bound (ncu: profiles/README.md)"}
    out["speedup_vs_simt"] = out["train_fwd_bwd_kernel"]["ms"] / out["train_e32_tc_kernel"]["ms"]
    return out


def extra_records(args, dev, world, rank, flush, peaks, fp32_peak, umma_peak):
    """Every other BASELINE config, measured in this run.  All ranks execute this (the training step has a collective)."""
    import torch.distributed as dist
    rec = {}
    rec["koopman_train_step"], ms = record_koopman_train(dev, world, rank)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rec["koopman_train_step"]["ms_per_step"] = float(t)
        rec["koopman_train_step"]["samples_per_s"] = world * 512 / (float(t) * 1e-3)
    if world > 1:
        # the literal BASELINE config[1]: 65,536 trajectories in total, sharded over the GPUs (strong scaling)
        kind, Btot, S = WORKLOADS["rossler"]
        Bl = Btot // world
        cfg, emb, tr = build_modules(kind, dev)
        x0 = initial_states(kind, Bl, 2000 + rank).to(dev)

        def pass_device():
            g0 = emb.embed(x0)
            out = tr.generate(g0.unsqueeze(1), max_length=S + 1, use_cache=True, return_presents=False)[0]
            return emb.recover(out.view(Bl * (S + 1), cfg.n_embd))
        with torch.no_grad():
            for _ in range(3):
                pass_device()
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            evs = []
            for _ in range(5):
                flush.fill_(5.0)
                a, b = ev(), ev()
                a.record()
                pass_device()
                b.record()
                evs.append((a, b))
            torch.cuda.synchronize()
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in evs) / len(evs)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        launch = tr.last_launch()
        groups = (Bl + launch["group"] - 1) // launch["group"]
        slots = launch["grid"] * (launch["block"] // 32)
        rec["strong"] = {"workload": f"rossler 65,536 trajectories in total = {Bl}/GPU x {world} GPUs, horizon {S}",
                         "scaling": "strong", "ms_per_step": float(t), "value": Btot * S / (float(t) * 1e-3), "unit": UNIT,
                         "rollout_launch": launch,
                         "limiter": f"{groups} warp groups of {launch['group']} trajectories on {slots} resident warp slots "
                                    f"({groups / max(slots, 1):.2f} waves) + fixed embed/recover launches"}
    if rank == 0 and not args.no_configs:
        rec["forward_teacher_forced"] = record_forward(dev, fp32_peak)
        rec["transformer_train_step"] = record_transformer_train(dev, peaks.get("_mma_sync_tf32", 278.0))
        rec["lorenz256"] = record_lorenz(dev, "lorenz", 256, 128, flush, peaks, fp32_peak)
        rec["lorenz_sweep"] = [record_lorenz(dev, "lorenz", B, S, flush, peaks, fp32_peak, e2e=False)
                               for B, S in ((1024, 64), (65536, 256), (262144, 128))]
        rec["cylinder"] = [record_cylinder(dev, 6, 400, umma_peak, True), record_cylinder(dev, 64, 400, umma_peak, True),
                           record_cylinder(dev, 512, 400, umma_peak, False)]
    if world > 1:
        dist.barrier()
    return rec


def reference_eager_on_gpu(dev, kind, B, S):
    """Informational: the unmodified reference modules moved to the B200 (PyTorch eager, FP32, TF32 off) on a bounded
    sample of the same workload — the number a user of the reference would see before switching."""
    from oracle import ref_shim
    from tests import synth
    if not ref_shim.available():
        return {"unavailable": "baseline/_ref not installed"}
    ref_shim.load()
    from trphysx.config.configuration_phys import PhysConfig
    from trphysx.embedding import LorenzEmbedding
    from trphysx.transformer import PhysformerGPT2
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg_ns = synth.make_config(kind)
    cfg = PhysConfig(**vars(cfg_ns))
    emb, tr = LorenzEmbedding(cfg).eval(), PhysformerGPT2(cfg, kind).eval()
    emb.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in synth.lorenz_embedding_weights(cfg_ns, 2024).items()})
    tr.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in synth.gpt2_weights(cfg_ns, 2025).items()})
    emb.to(dev)
    tr.to(dev)
    x0 = initial_states(kind, B, 7).to(dev)

    def run():
        with torch.no_grad():
            g0 = emb.embed(x0)
            out = tr.generate(g0.unsqueeze(1), max_length=S + 1, use_cache=True)[0]
            return emb.recover(out.reshape(-1, cfg.n_embd))
    run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"value": B * S / dt, "unit": UNIT, "sample": f"{kind} B={B} horizon={S} use_cache=True, reference modules "
            "(baseline/_ref) in PyTorch eager on cuda:0, FP32 (TF32 off), wall clock"}


# ----------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from trphysx_b200 import _native as N
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    kind, B, S = WORKLOADS[args.workload]
    B = args.batch or B
    S = args.horizon or S
    cfg, emb, tr = build_modules(kind, dev)
    if args.tune:
        tr.set_rollout_tuning(**{k: int(v) for k, v in (kv.split("=") for kv in args.tune.split(","))})
    E = cfg.n_embd
    x0_host = initial_states(kind, B, 1000 + rank).pin_memory()
    x0_dev = x0_host.to(dev)
    states_host = torch.empty((B, S + 1, 3), dtype=torch.float32).pin_memory()
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)     # > 126 MB L2

    def pass_device():
        g0 = emb.embed(x0_dev)
        out = tr.generate(g0.unsqueeze(1), max_length=S + 1, use_cache=True, return_presents=False)[0]
        return emb.recover(out.view(B * (S + 1), E))

    from trphysx_b200.evaluation import rollout_to_host
    copy_stream = torch.cuda.Stream(device=dev)

    def pass_e2e():
        # ONE public call: pinned x0 -> H2D -> embed -> [generate -> recover] per slice of one round of resident
        # trajectories, each slice's states going device -> host on the copy stream while the next slice decodes
        rollout_to_host(emb, tr, x0_host, S + 1, states_host, device=dev, use_cache=True, chunks=4,
                        copy_stream=copy_stream)

    def timed(fn, steps, warmup, sampler=None):
        with torch.no_grad():
            for _ in range(warmup):
                fn()
                flush.fill_(1.0)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            if sampler:
                sampler.start()
            evs = [(ev(), ev()) for _ in range(steps)]
            for a, b in evs:
                flush.fill_(float(len(evs)))          # evict L2 between timed iterations (outside the event pair)
                a.record()
                fn()
                b.record()
            torch.cuda.synchronize()
            clocks = sampler.stop() if sampler else None
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in evs) / steps
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), clocks

    sampler = ClockSampler(local) if rank == 0 else None
    ms_dev, clocks = timed(pass_device, args.steps, args.warmup, sampler)
    ms_e2e, _ = timed(pass_e2e, args.steps, max(1, args.warmup // 2))
    units = world * B * S
    value = units / (ms_dev * 1e-3)
    e2e = units / (ms_e2e * 1e-3)

    # ---- the dominant kernel alone: rollout -------------------------------------------------------------
    with torch.no_grad():
        g0 = emb.embed(x0_dev).unsqueeze(1).contiguous()
        kev = []
        for _ in range(3):
            tr.generate(g0, max_length=S + 1, use_cache=True, return_presents=False)
        for _ in range(max(3, args.steps)):
            flush.fill_(2.0)
            out = torch.empty((B, S + 1, E), device=dev)
            out[:, :1] = g0
            a, b = ev(), ev()
            h = tr._handle(dev)
            a.record()
            N.check(N.lib().trpx_gpt2_rollout(h, out.data_ptr(), out.stride(0), out[:, 1:].data_ptr(), out.stride(0),
                                              B, S, 1, None, N.stream_ptr(dev)))
            b.record()
            kev.append((a, b))
        torch.cuda.synchronize()
        k_ms = statistics.median(a.elapsed_time(b) for a, b in kev)
        launch = tr.last_launch()
        # Markov (reference default use_cache=False) mode, for the record
        for _ in range(2):
            tr.generate(g0, max_length=S + 1, use_cache=False)
        a, b = ev(), ev()
        a.record()
        tr.generate(g0, max_length=S + 1, use_cache=False)
        b.record()
        torch.cuda.synchronize()
        nocache_ms = a.elapsed_time(b)
        nocache_launch = tr.last_launch()
        # the same workload with the KV window ON CHIP (tensor memory, rollout variant 10): the measured prototype of
        # BASELINE north_star's "KV cache in shared memory or registers" (profiles/README.md)
        onchip = None
        if rank == 0 and cfg.n_ctx in (32, 64) and cfg.n_layer <= 4:
            tr.set_rollout_tuning(variant=10)
            for _ in range(2):
                tr.generate(g0, max_length=S + 1, use_cache=True, return_presents=False)
            a, b = ev(), ev()
            a.record()
            tr.generate(g0, max_length=S + 1, use_cache=True, return_presents=False)
            b.record()
            torch.cuda.synchronize()
            onchip = {"kernel": "rollout_e32t_kernel (KV window in tensor memory, tcgen05.st/ld; the automatic choice only up to two rounds of resident tiles, i.e. B <= 2,368 here)",
                      "kernel_ms": a.elapsed_time(b), "traj_steps_per_s": B * S / (a.elapsed_time(b) * 1e-3),
                      "rollout_launch": tr.last_launch(),
                      "hbm_bytes_per_launch_model": B * S * E * 4 + B * E * 4,
                      "note": "HBM traffic = tokens only (ncu: profiles/r02_ncu_rollout_tmem.json); at this batch slower than "
                              "the HBM-window kernel because 8 resident trajectories per SM force all warps onto one tile"}
            tr.set_rollout_tuning()
        # embed / recover alone (median of 5)
        em, rc = [], []
        for _ in range(5):
            a, b, c = (ev() for _ in range(3))
            a.record()
            emb.embed(x0_dev)
            b.record()
            emb.recover(out.view(B * (S + 1), E))
            c.record()
            torch.cuda.synchronize()
            em.append(a.elapsed_time(b))
            rc.append(b.elapsed_time(c))
        embed_ms, recover_ms = statistics.median(em), statistics.median(rc)
        del out

    fp32, mma, umma = N.ctypes.c_float(), N.ctypes.c_float(), N.ctypes.c_float()
    N.check(N.lib().trpx_probe_fp32_fma(local, N.ctypes.byref(fp32)))
    N.check(N.lib().trpx_probe_tf32_mma(local, N.ctypes.byref(mma)))
    N.check(N.lib().trpx_probe_umma_tf32(local, N.ctypes.byref(umma)))
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    peaks = dict(peaks)
    peaks["_mma_sync_tf32"] = mma.value
    extras = extra_records(args, dev, world, rank, flush, peaks, fp32.value, umma.value)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # the window access pattern alone (no arithmetic), same warps / rows in flight as the rollout: what the memory
    # system delivers for this pattern on this box
    probe_gbs = None
    if cfg.n_embd == 32 and cfg.n_ctx % 32 == 0:
        pg, pm = N.ctypes.c_float(), N.ctypes.c_float()
        N.check(N.lib().trpx_probe_window_stream(local, min(B - B % 4, 65536), cfg.n_ctx, cfg.n_layer, 32, 8, 16, 0,
                                                 N.ctypes.byref(pg), N.ctypes.byref(pm)))
        probe_gbs = pg.value
    flops_step = algorithmic_flops_per_traj_step(cfg, S)
    k_flops = flops_step * B * S
    achieved_tf = k_flops / (k_ms * 1e-3) / 1e12
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    model_bytes = window_bytes_per_traj(cfg, S) * B
    achieved_gbs = model_bytes / (k_ms * 1e-3) / 1e9
    kname = {1: "rollout_generic_kernel", 2: "rollout_e32_kernel", 5: "rollout_e32v2_kernel", 6: "rollout_e32v4_kernel",
             7: "rollout_cluster_kernel", 8: "rollout_e32s_kernel", 10: "rollout_e32t_kernel"}
    # dram__bytes of the same launch from a committed ncu capture: reported only when the capture was taken with the very
    # launch configuration this run used (kernel variant, grid, block, group), with its provenance beside it
    traffic, traffic_source = None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_rollout_traffic.json")
    if os.path.exists(tpath):
        for rec in json.load(open(tpath)):
            if rec.get("kind") == kind and rec.get("B") == B and rec.get("S") == S:
                same = all(rec.get("launch", {}).get(k) == launch[k] for k in ("variant", "grid", "block", "group"))
                if same:
                    traffic = rec["dram_bytes_per_launch"]
                    traffic_source = {"file": "profiles/ncu_rollout_traffic.json", "capture": rec.get("capture"),
                                      "commit": rec.get("commit"), "launch": rec.get("launch")}
                else:
                    traffic_source = {"omitted": "the committed ncu capture was taken with another launch configuration",
                                      "captured_launch": rec.get("launch"), "this_launch": launch}
    nc_flops = algorithmic_flops_per_traj_step(cfg, S, use_cache=False) * B * S
    nc_tf = nc_flops / (nocache_ms * 1e-3) / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args), "weights": "random init (reference _init_weights)",
                   "l2": "flushed between timed iterations (256 MiB write)", "parallelism": f"traj-shard x{world}",
                   "rollout_launch": launch},
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": B * 3 * 4,
                "d2h_bytes_per_step": B * (S + 1) * 3 * 4},
        "gpu_launches": 3 * args.steps,
        "roofline": {"bound": "hbm", "kernel": kname.get(launch["variant"], "?"),
                     "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                     "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)",
                     "traffic": traffic, "traffic_source": traffic_source, "kernel_ms": k_ms, "bytes_per_launch": model_bytes,
                     "bytes_model": "KV-window stream: per step and layer 2*L*E*4 B read + 2*E*4 B appended, + E*4 B token "
                                    "written (SURVEY 8d spill case; see bench.py window_bytes_per_traj and DESIGN.md 3)",
                     "kernel_traj_steps_per_s": B * S / (k_ms * 1e-3),
                     "window_stream_probe_GBps": probe_gbs,
                     "fp32_view": {"achieved": achieved_tf, "peak": fp32.value, "unit": "TFLOP/s",
                                   "frac": achieved_tf / fp32.value, "flops_per_traj_step": flops_step,
                                   "peak_source": "FP32 FMA microbenchmark in this run (trpx_probe_fp32_fma)"},
                     "token_only_view": {"algorithmic_GBps": B * S * E * 4 / (k_ms * 1e-3) / 1e9,
                                         "note": "tokens written only (window treated as on-chip): not reachable, see DESIGN.md"}},
        "use_cache_false": {"note": "the reference's DEFAULT generate() mode (1-token forward at position 0, no window)",
                            "kernel": kname.get(nocache_launch["variant"], "?"), "kernel_ms": nocache_ms,
                            "traj_steps_per_s": B * S / (nocache_ms * 1e-3),
                            "roofline": {"bound": "tensor", "achieved": nc_tf, "peak": mma.value / 3.0, "unit": "TFLOP/s",
                                         "frac": nc_tf / (mma.value / 3.0),
                                         "peak_source": "measured mma.sync TF32 rate in this run (trpx_probe_tf32_mma) / 3 "
                                                        "for the error-compensated 3xTF32 product",
                                         "vs_fp32_fma_peak": nc_tf / fp32.value}},
        "onchip_window": onchip,
        "breakdown_ms": {"embed": embed_ms, "rollout": k_ms, "recover": recover_ms, "rollout_use_cache_false": nocache_ms},
        "probes": {"fp32_fma_tflops": fp32.value, "tf32_mma_sync_tflops": mma.value, "tf32_tcgen05_tflops": umma.value},
        "configs": extras,
    }
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        Bc, Sc = args.cpu_batch, min(args.cpu_horizon, S)
        v, how, reps = time_cpu(kind, Bc, Sc, threads, use_cache=True, budget_s=12.0)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": how,
                                "sample": cpu_sample_text(kind, Bc, Sc, how) + f", {reps} passes"}
        if not args.no_configs:
            line["reference_eager_on_b200"] = reference_eager_on_gpu(dev, kind, 8192, min(S, 128))
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="rossler", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU trajectories (default: workload's)")
    ap.add_argument("--horizon", type=int, default=0)
    ap.add_argument("--tune", default="", help="e.g. warps_per_cta=8,traj_per_warp=8")
    ap.add_argument("--cpu-batch", type=int, default=2048, help="trajectories of the bounded CPU sample")
    ap.add_argument("--cpu-horizon", type=int, default=256, help="horizon of the CPU sample (capped at the workload's)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the records of the other BASELINE configs")
    ap.add_argument("--no-variants", action="store_true", help="reference arm: skip the use_cache=False / 1-thread timings")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints its version banner on
    # stdout at communicator creation), so file descriptor 1 points at stderr while the benchmark runs and the JSON
    # line is written to the saved descriptor.
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            if args.impl == "reference":
                run_reference(args)
            else:
                run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    out = [l for l in buf.getvalue().splitlines() if l.strip()]
    for l in out[:-1]:
        print(l, file=sys.stderr)
    if out:
        print(out[-1], flush=True)


if __name__ == "__main__":
    main()
