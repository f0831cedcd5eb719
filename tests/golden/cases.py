"""Table of golden cases shared by make_golden.py (reference side) and the parity tests."""

# kind/over -> config; seed -> weights (seed), forward input (seed+1), second chunk (seed+2),
# generation context (seed+3).  `ctx` > 1 exercises "only the last context token is used" (fact 2);
# max_length > n_ctx + ctx exercises the saturated sliding window (fact 5).
GPT2_CASES = {
    # Lorenz shape (E=32, 4 layers, 4 heads, n_ctx=64), reference-style init
    "lorenz_init":   dict(kind="lorenz", seed=101, stress=False, B=3, T=64, Tp=20, T2=30, ctx=1, max_length=150),
    "lorenz_stress": dict(kind="lorenz", seed=102, stress=True,  B=3, T=64, Tp=33, T2=31, ctx=5, max_length=150),
    # Rossler shape (n_ctx=32)
    "rossler_stress": dict(kind="rossler", seed=103, stress=True, B=4, T=32, Tp=7, T2=25, ctx=1, max_length=100),
    # tiny context so the window is saturated almost immediately
    "tiny_ctx8":     dict(kind="lorenz", over=dict(n_ctx=8), seed=104, stress=True, B=5, T=8, Tp=3, T2=5, ctx=3, max_length=40),
    # odd shapes: 2 layers, 2 heads (hd=16)
    "two_heads":     dict(kind="lorenz", over=dict(n_ctx=16, n_layer=2, n_head=2), seed=105, stress=True, B=2, T=11, Tp=4, T2=9, ctx=2, max_length=40),
    # Cylinder shape (E=128, 6 layers, hd=32, n_ctx=16)
    "cylinder_init":   dict(kind="cylinder", seed=106, stress=False, B=2, T=16, Tp=6, T2=10, ctx=1, max_length=48),
    "cylinder_stress": dict(kind="cylinder", over=dict(n_layer=3), seed=107, stress=True, B=2, T=16, Tp=5, T2=11, ctx=1, max_length=40),
    # sigma = 0.1 stress weights: the cached rollout stays well conditioned over the WHOLE horizon (3 x n_ctx + a few
    # steps) without collapsing to a fixed point (consecutive tokens still differ by ~1 % at 2 x n_ctx), so every token
    # produced with a full / wrapped KV ring is compared.  B >= 32 so the production trajectories-per-warp groupings
    # (8 for n_ctx 32, 4 for n_ctx 64) get several full groups plus a ragged one.  fwdB: batch of the teacher-forced part (keeps the fixture small).
    "rossler_sigma01": dict(kind="rossler", seed=108, stress=True, sigma=0.1, B=40, fwdB=3, T=32, Tp=9, T2=23, ctx=1, max_length=101),
    "lorenz_sigma01":  dict(kind="lorenz", seed=109, stress=True, sigma=0.1, B=36, fwdB=3, T=64, Tp=21, T2=43, ctx=2, max_length=200),
}


def weights_kw(c):
    """Keyword arguments of tests.synth.gpt2_weights for a case (stress weights default to sigma = 0.3)."""
    kw = dict(stress=c["stress"])
    if "sigma" in c:
        kw["stress_sigma"] = c["sigma"]
    return kw

LORENZ_CASES = {
    "lorenz":  dict(kind="lorenz", seed=201, N=37),
    "rossler": dict(kind="rossler", seed=202, N=16),
}

CYL_CASES = {
    "cyl": dict(seed=301, N=2),
}

GS_CASES = {
    "gs": dict(seed=351, N=2),
}

TRAIN_CASES = {
    "lorenz":  dict(kind="lorenz", seed=401, B=24, T=6, iters=2),
    "rossler": dict(kind="rossler", seed=402, B=16, T=4, iters=1),
}

# Teacher-forced transformer training step: PhysformerTrain.forward + loss.backward() of the reference
# (phys_transformer_helpers.py:36-69).  Stored: loss, hidden states, every parameter gradient (state_dict names).
GPT2_TRAIN_CASES = {
    "lorenz_stress": dict(kind="lorenz", seed=501, stress=True, B=5, T=64),
    "rossler_init": dict(kind="rossler", seed=502, stress=False, B=3, T=32),
    "rossler_short": dict(kind="rossler", seed=503, stress=True, B=4, T=7),
    "cylinder_stress": dict(kind="cylinder", seed=504, stress=True, B=2, T=16,
                            keep=("h.0.ln_1.", "h.0.attn.c_attn.", "h.2.attn.c_proj.bias", "h.3.ln_2.", "h.5.mlp.c_fc.bias",
                                  "h.5.mlp.c_proj.bias", "ln_f.", "mlp_f.")),
}

# Cylinder Koopman training loss (embedding_cylinder.py:236-281), value only: [B, T, 3, 64, 128] series built from
# synth.cylinder_fields(B*T) frames, viscosity per sample.
CYL_TRAIN_CASES = {
    "cyl": dict(seed=601, B=2, T=3),
}
