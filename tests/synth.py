"""Deterministic synthetic weights and inputs (numpy only).

TEST / BENCH INFRASTRUCTURE.  Shared by tests/golden/make_golden.py (which feeds the same arrays
to the imported reference), the parity tests, __graft_entry__.smoke() and bench.py.  Nothing here is
imported by the product package.

Everything is derived from ``numpy.random.default_rng(seed)`` so the golden fixtures only have to store
the reference's OUTPUTS: weights and inputs are regenerated bit-identically wherever they are needed.

State-dict key names/shapes follow the reference modules:
  transformer  trphysx/transformer/phys_transformer_gpt2.py:126-138, attention.py:49-71, utils.py:37-40
  lorenz       trphysx/embedding/embedding_lorenz.py:41-71
  cylinder     trphysx/embedding/embedding_cylinder.py:41-108
"""
from collections import OrderedDict
from types import SimpleNamespace

import numpy as np


# --------------------------------------------------------------------------------------------------
# configurations (plain namespaces with the PhysConfig field names, configuration_phys.py:52-74)
# --------------------------------------------------------------------------------------------------
def make_config(kind: str, **over):
    """kind: 'lorenz' (configuration_lorenz.py:23-29), 'rossler' (examples/rossler/rossler_module/
    configuration_rossler.py:24-30), 'cylinder' (configuration_cylinder.py:23-27)."""
    base = dict(activation_function="gelu_new", resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0,
                layer_norm_epsilon=1e-5, output_hidden_states=False, output_attentions=False,
                use_cache=True, max_length=8, min_length=0, k_size=1)
    if kind == "lorenz":
        base.update(n_ctx=64, n_embd=32, n_layer=4, n_head=4, state_dims=[3], initializer_range=0.05)
    elif kind == "rossler":
        base.update(n_ctx=32, n_embd=32, n_layer=4, n_head=4, state_dims=[3], initializer_range=0.02)
    elif kind == "cylinder":
        base.update(n_ctx=16, n_embd=128, n_layer=6, n_head=4, state_dims=[3, 64, 128],
                    initializer_range=0.01)
    elif kind == "grayscott":                       # configuration_grayscott.py:15-38
        base.update(n_ctx=128, n_embd=512, n_layer=2, n_head=32, state_dims=[2, 32, 32, 32],
                    initializer_range=0.01)
    else:
        raise ValueError(kind)
    base.update(over)
    return SimpleNamespace(**base)


# --------------------------------------------------------------------------------------------------
# weights
# --------------------------------------------------------------------------------------------------
def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def gpt2_weights(cfg, seed: int, stress: bool = False, stress_sigma: float = 0.3) -> "OrderedDict[str, np.ndarray]":
    """Random transformer state_dict.  stress=False mimics `_init_weights`
    (phys_transformer_base.py:46-57: N(0, initializer_range), zero bias, LN = (1, 0)).
    stress=True draws larger weights, non-zero biases and random LN affine so that attention logits,
    the softmax and the GELU non-linearity are actually exercised (SURVEY.md §8c degeneracy warning)."""
    rng = np.random.default_rng(seed)
    E, L, C = cfg.n_embd, cfg.n_layer, cfg.n_ctx
    sw = stress_sigma if stress else cfg.initializer_range      # stress_sigma < 0.3: milder, better conditioned
    sd = OrderedDict()

    def w(*shape):
        return _f32(rng.normal(0.0, sw, size=shape))

    def b(n):
        return _f32(rng.normal(0.0, 0.2, size=n)) if stress else np.zeros(n, np.float32)

    def ln(prefix):
        if stress:
            sd[prefix + ".weight"] = _f32(1.0 + 0.3 * rng.normal(size=E))
            sd[prefix + ".bias"] = _f32(0.2 * rng.normal(size=E))
        else:
            sd[prefix + ".weight"] = np.ones(E, np.float32)
            sd[prefix + ".bias"] = np.zeros(E, np.float32)

    tril = np.tril(np.ones((C, C), np.uint8)).reshape(1, 1, C, C)
    for l in range(L):
        p = f"h.{l}."
        ln(p + "ln_1")
        sd[p + "attn.bias"] = tril.copy()
        sd[p + "attn.masked_bias"] = np.array(-1e4, np.float32)
        sd[p + "attn.c_attn.weight"] = w(E, 3 * E)
        sd[p + "attn.c_attn.bias"] = b(3 * E)
        sd[p + "attn.c_proj.weight"] = w(E, E)
        sd[p + "attn.c_proj.bias"] = b(E)
        ln(p + "ln_2")
        sd[p + "mlp.c_fc.weight"] = w(E, 4 * E)
        sd[p + "mlp.c_fc.bias"] = b(4 * E)
        sd[p + "mlp.c_proj.weight"] = w(4 * E, E)
        sd[p + "mlp.c_proj.bias"] = b(E)
    ln("ln_f")
    sd["mlp_f.weight"] = w(E, E)
    sd["mlp_f.bias"] = b(E)
    sd["wpe.weight"] = w(C, E)          # present in the state_dict, never used by forward()
    return sd


def lorenz_embedding_weights(cfg, seed: int) -> "OrderedDict[str, np.ndarray]":
    """Random Lorenz/Rossler embedding state_dict (embedding_lorenz.py:41-71).  Uniform(-1/sqrt(fan_in),
    +1/sqrt(fan_in)) like torch's Linear default; non-trivial LN affine, mu/std so every term is live."""
    rng = np.random.default_rng(seed)
    E, H, S = cfg.n_embd, 500, cfg.state_dims[0]
    sd = OrderedDict()

    def lin(prefix, fan_out, fan_in):
        k = 1.0 / np.sqrt(fan_in)
        sd[prefix + ".weight"] = _f32(rng.uniform(-k, k, size=(fan_out, fan_in)))
        sd[prefix + ".bias"] = _f32(rng.uniform(-k, k, size=fan_out))

    sd["kMatrixDiag"] = _f32(np.linspace(1, 0, E) + 0.01 * rng.normal(size=E))
    sd["kMatrixUT"] = _f32(0.1 * rng.uniform(size=(E - 1) + (E - 2)))
    sd["mu"] = _f32([0.3, -0.2, 24.0][:S])
    sd["std"] = _f32([8.0, 9.0, 8.5][:S])
    lin("observableNet.0", H, S)
    lin("observableNet.2", E, H)
    sd["observableNet.3.weight"] = _f32(1.0 + 0.1 * rng.normal(size=E))
    sd["observableNet.3.bias"] = _f32(0.1 * rng.normal(size=E))
    lin("recoveryNet.0", H, E)
    lin("recoveryNet.2", S, H)
    return sd


def cylinder_embedding_weights(cfg, seed: int) -> "OrderedDict[str, np.ndarray]":
    """Random Cylinder embedding state_dict (embedding_cylinder.py:41-108)."""
    rng = np.random.default_rng(seed)
    E = cfg.n_embd
    sd = OrderedDict()

    def conv(prefix, cout, cin):
        k = 1.0 / np.sqrt(cin * 9)
        sd[prefix + ".weight"] = _f32(rng.uniform(-k, k, size=(cout, cin, 3, 3)))
        sd[prefix + ".bias"] = _f32(rng.uniform(-k, k, size=cout))

    def lin(prefix, fan_out, fan_in):
        k = 1.0 / np.sqrt(fan_in)
        sd[prefix + ".weight"] = _f32(rng.uniform(-k, k, size=(fan_out, fan_in)))
        sd[prefix + ".bias"] = _f32(rng.uniform(-k, k, size=fan_out))

    sd["mu"] = _f32([0.1, -0.05, 0.02, 0.008])
    sd["std"] = _f32([0.9, 1.1, 1.2, 0.005])
    conv("observableNet.0", 16, 4)
    conv("observableNet.2", 32, 16)
    conv("observableNet.4", 64, 32)
    conv("observableNet.6", 128, 64)
    conv("observableNet.8", E // 32, 128)
    sd["observableNetFC.0.weight"] = _f32(1.0 + 0.1 * rng.normal(size=E))
    sd["observableNetFC.0.bias"] = _f32(0.1 * rng.normal(size=E))
    conv("recoveryNet.1", 128, E // 32)
    conv("recoveryNet.4", 64, 128)
    conv("recoveryNet.7", 32, 64)
    conv("recoveryNet.10", 16, 32)
    conv("recoveryNet.12", 3, 16)
    nband = sum(E - i for i in range(1, 5))
    lin("kMatrixDiagNet.0", 50, 1)
    lin("kMatrixDiagNet.2", E, 50)
    lin("kMatrixUT.0", 50, 1)
    lin("kMatrixUT.2", nband, 50)
    lin("kMatrixLT.0", 50, 1)
    lin("kMatrixLT.2", nband, 50)
    return sd


def grayscott_embedding_weights(cfg, seed: int) -> "OrderedDict[str, np.ndarray]":
    """Random Gray-Scott embedding state_dict (embedding_grayscott.py:41-108) with non-trivial BatchNorm3d running
    statistics; key order = the module's own state_dict order."""
    rng = np.random.default_rng(seed)
    E = cfg.n_embd
    sd = OrderedDict()

    def conv(prefix, cout, cin, k):
        b = 1.0 / np.sqrt(cin * k ** 3)
        sd[prefix + ".weight"] = _f32(rng.uniform(-b, b, size=(cout, cin, k, k, k)))
        sd[prefix + ".bias"] = _f32(rng.uniform(-b, b, size=cout))

    def bn(prefix, c):
        sd[prefix + ".weight"] = _f32(1.0 + 0.1 * rng.normal(size=c))
        sd[prefix + ".bias"] = _f32(0.1 * rng.normal(size=c))
        sd[prefix + ".running_mean"] = _f32(0.1 * rng.normal(size=c))
        sd[prefix + ".running_var"] = _f32(rng.uniform(0.5, 1.5, size=c))
        sd[prefix + ".num_batches_tracked"] = np.asarray(7, dtype=np.int64)

    def lin(prefix, fan_out, fan_in):
        b = 1.0 / np.sqrt(fan_in)
        sd[prefix + ".weight"] = _f32(rng.uniform(-b, b, size=(fan_out, fan_in)))
        sd[prefix + ".bias"] = _f32(rng.uniform(-b, b, size=fan_out))

    nband = sum(E - i for i in range(1, 10))
    sd["kMatrixDiag"] = _f32(rng.uniform(0.5, 1.0, size=E))
    sd["kMatrixUT"] = _f32(0.01 * rng.uniform(0, 1, size=nband))
    sd["mu"] = np.asarray(0.3, dtype=np.float32)
    sd["std"] = np.asarray(0.8, dtype=np.float32)
    for i, (cout, cin, k) in zip((0, 3, 6, 9), ((64, 2, 5), (128, 64, 3), (128, 128, 3), (64, 128, 1))):
        conv(f"observableNet.{i}", cout, cin, k)
        bn(f"observableNet.{i + 1}", cout)
    lin("observableNetFC.0", 512, 4096)
    lin("observableNetFC.2", E, 512)
    sd["observableNetFC.3.weight"] = _f32(1.0 + 0.1 * rng.normal(size=E))
    sd["observableNetFC.3.bias"] = _f32(0.1 * rng.normal(size=E))
    lin("recoveryNetFC.0", 512, E)
    lin("recoveryNetFC.2", 4096, 512)
    for i, (cout, cin, k) in zip((0, 4, 8, 12), ((128, 64, 1), (128, 128, 3), (64, 128, 3), (64, 64, 3))):
        conv(f"recoveryNet.{i}", cout, cin, k)
        bn(f"recoveryNet.{i + 1}", cout)
    conv("recoveryNet.15", 2, 64, 1)
    return sd


def grayscott_fields(n: int, seed: int = 0) -> np.ndarray:
    """Synthetic [n, 2, 32, 32, 32] concentration fields in [0, 1] (smooth modes + noise)."""
    rng = np.random.default_rng(seed)
    ax = np.linspace(0, 2 * np.pi, 32, endpoint=False)
    Z, Y, X = np.meshgrid(ax, ax, ax, indexing="ij")
    out = np.empty((n, 2, 32, 32, 32))
    for i in range(n):
        for c in range(2):
            k = rng.integers(1, 4, size=3)
            ph = rng.uniform(0, 2 * np.pi, size=3)
            out[i, c] = 0.5 + 0.3 * np.sin(k[0] * Z + ph[0]) * np.cos(k[1] * Y + ph[1]) * np.sin(k[2] * X + ph[2]) \
                + 0.05 * rng.normal(size=(32, 32, 32))
    return _f32(out)


# --------------------------------------------------------------------------------------------------
# synthetic trajectories (the Zenodo data is not available offline)
# --------------------------------------------------------------------------------------------------
def _rk4(f, x, dt, steps):
    out = np.empty((steps + 1,) + x.shape, np.float64)
    out[0] = x
    for s in range(steps):
        k1 = f(x)
        k2 = f(x + 0.5 * dt * k1)
        k3 = f(x + 0.5 * dt * k2)
        k4 = f(x + dt * k3)
        x = x + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
        out[s + 1] = x
    return out


def lorenz_trajectories(n: int, steps: int, seed: int = 12345, dt: float = 0.01) -> np.ndarray:
    """[n, steps+1, 3] float32 Lorenz-63 (sigma=10, rho=28, beta=8/3), classic RK4."""
    rng = np.random.default_rng(seed)
    x0 = np.stack([rng.uniform(-20, 20, n), rng.uniform(-20, 20, n), rng.uniform(10, 40, n)], -1)

    def f(x):
        return np.stack([10.0 * (x[:, 1] - x[:, 0]),
                         x[:, 0] * (28.0 - x[:, 2]) - x[:, 1],
                         x[:, 0] * x[:, 1] - (8.0 / 3.0) * x[:, 2]], -1)

    return _f32(np.transpose(_rk4(f, x0, dt, steps), (1, 0, 2)))


def rossler_trajectories(n: int, steps: int, seed: int = 12345, dt: float = 0.05) -> np.ndarray:
    """[n, steps+1, 3] float32 Rossler (a=b=0.2, c=5.7), classic RK4."""
    rng = np.random.default_rng(seed)
    x0 = rng.normal(size=(n, 3)) * np.array([5.0, 5.0, 1.0]) + np.array([0.0, 0.0, 1.0])

    def f(x):
        return np.stack([-x[:, 1] - x[:, 2],
                         x[:, 0] + 0.2 * x[:, 1],
                         0.2 + x[:, 2] * (x[:, 0] - 5.7)], -1)

    return _f32(np.transpose(_rk4(f, x0, dt, steps), (1, 0, 2)))


def cylinder_fields(n: int, seed: int = 0):
    """x [n,3,64,128] ~ N(0,1), visc [n,1] ~ U(2/750, 2/100) (visc = 2/Re, dataset_cylinder.py:36)."""
    rng = np.random.default_rng(seed)
    x = _f32(rng.normal(size=(n, 3, 64, 128)))
    visc = _f32(rng.uniform(2.0 / 750.0, 2.0 / 100.0, size=(n, 1)))
    return x, visc


def normal_embeds(shape, seed: int, scale: float = 1.0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return _f32(scale * rng.normal(size=shape))
