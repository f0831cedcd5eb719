"""CPU oracle: a functional restatement of the reference's rollout hot path.

TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this file.  The product package
(transformer-physx_b200/) never imports anything under oracle/ and has no CPU path at all.

Parity status: PINNED.  The reference's own tests hold no golden vectors for this path
(SURVEY.md §4, §8c), so the oracle is pinned against outputs of the reference itself: the fixtures in
tests/golden/*.npz were produced by tests/golden/make_golden.py, which imports the unmodified reference
from /root/reference, loads it with the tests/synth.py weights and records its outputs.
tests/test_oracle_golden.py checks every function below against those fixtures (and, where
/root/reference is present, tests/test_oracle_vs_reference.py repeats the comparison live).

The arithmetic lives in PyTorch (torch >= 1.7, reference setup.py:138), which is installed here, so the
restatement uses torch CPU tensors in float32 (or float64 when `dtype=torch.float64` is requested to
get a higher-precision yardstick).  All functions take a plain dict name -> tensor using the
reference's state_dict key names; there are no nn.Modules.

Reference lines restated (relative to /root/reference):
  gpt2_forward        trphysx/transformer/phys_transformer_gpt2.py:147-269 (+ Block :100-118, MLP :53-55)
  _attention          trphysx/transformer/attention.py:153-200, 75-119 ; Conv1D utils.py:42-54
  gelu_new            trphysx/transformer/utils.py:57-60
  gpt2_generate       trphysx/transformer/generate_utils.py:25-56, 66-110, 112-169
  lorenz_*            trphysx/embedding/embedding_lorenz.py:74-167 (Rossler: examples/rossler/
                      rossler_module/embedding_rossler.py:71-161, identical arithmetic)
  cylinder_*          trphysx/embedding/embedding_cylinder.py:112-222 (training loss :236-281)
  lorenz_train_loss   trphysx/embedding/embedding_lorenz.py:181-221 (weights: Rossler :195-215)
  clip_adam_step      trphysx/embedding/training/enn_trainer.py:93-102 + torch.optim.Adam semantics
"""
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


def to_torch(sd_np, dtype=torch.float32) -> SD:
    """numpy state dict (tests/synth.py) -> torch CPU tensors; float arrays are cast to `dtype`."""
    out = {}
    for k, v in sd_np.items():
        t = torch.from_numpy(np.array(v))
        out[k] = t.to(dtype) if t.is_floating_point() else t
    return out


# ==================================================================================================
# transformer
# ==================================================================================================
def gelu_new(x: Tensor) -> Tensor:
    # utils.py:57-60  (tanh form, cubic via pow)
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


def position_embedding(cfg, position_ids: Tensor, like: Tensor) -> Tensor:
    """phys_transformer_gpt2.py:215-219.  Even channels sin(p / 10000^(2j/E)); the index
    `i[:, :, n_embd % 2]` collapses the frequency vector to its element 0, so for even n_embd EVERY odd
    channel is cos(p)."""
    E = cfg.n_embd
    pe = torch.zeros_like(like)
    i = torch.arange(0, E // 2, dtype=torch.float, device=like.device).unsqueeze(0).unsqueeze(0)
    i = i.to(like.dtype)
    pe[:, :, ::2] = torch.sin(position_ids.unsqueeze(-1) / 10000 ** (2 * i / E))
    i0 = i[:, :, E % 2]
    pe[:, :, 1::2] = torch.cos(position_ids.unsqueeze(-1) / 10000 ** (2 * i0 / E))
    return pe


def _conv1d(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    # utils.py:50-53: y = addmm(bias, x2d, W) with W stored [in, out]
    shp = x.shape[:-1] + (w.shape[1],)
    return torch.addmm(b, x.reshape(-1, x.shape[-1]), w).view(shp)


def _attention(sd: SD, p: str, cfg, x: Tensor, layer_past: Optional[Tensor], use_cache: bool,
               want_attn: bool = False):
    """attention.py:176-200 and _attn :97-116.  Returns (a, present or None, attn or None)."""
    B, T, E = x.shape
    H = cfg.n_head
    hd = E // H
    qkv = _conv1d(x, sd[p + "c_attn.weight"], sd[p + "c_attn.bias"])
    q, k, v = qkv.split(E, dim=2)
    q = q.view(B, T, H, hd).permute(0, 2, 1, 3)          # [B,H,T,hd]
    k = k.view(B, T, H, hd).permute(0, 2, 3, 1)          # [B,H,hd,T]
    v = v.view(B, T, H, hd).permute(0, 2, 1, 3)          # [B,H,T,hd]
    if layer_past is not None:
        k = torch.cat((layer_past[0].transpose(-2, -1), k), dim=-1)
        v = torch.cat((layer_past[1], v), dim=-2)
    present = torch.stack((k.transpose(-2, -1), v)) if use_cache else None   # [2,B,H,Tk,hd]
    w = torch.matmul(q, k)
    w = w / (float(v.size(-1)) ** 0.5)                                       # scale=True (gpt2:134)
    nd, ns = w.size(-2), w.size(-1)
    mask = sd[p + "bias"][:, :, ns - nd: ns, :ns]
    w = torch.where(mask.bool(), w, sd[p + "masked_bias"].to(w.dtype))       # fill -1e4, not -inf
    w = torch.softmax(w, dim=-1)
    a = torch.matmul(w, v)
    a = a.permute(0, 2, 1, 3).contiguous().view(B, T, E)
    a = _conv1d(a, sd[p + "c_proj.weight"], sd[p + "c_proj.bias"])
    return a, present, (w if want_attn else None)


def gpt2_forward(sd: SD, cfg, inputs_embeds: Tensor, past: Optional[Sequence[Tensor]] = None,
                 use_cache: bool = True, output_attentions: bool = False):
    """PhysformerGPT2.forward with position_ids/prop_embeds/attention_mask/head_mask = None.
    Returns (hidden [B,T,E], presents tuple | None, attentions tuple | None)."""
    B, T, E = inputs_embeds.shape
    eps = cfg.layer_norm_epsilon
    if past is None:
        past_length = 0
        past = [None] * cfg.n_layer
    else:
        past_length = past[0][0].size(-2)
    pos = torch.arange(past_length, T + past_length, dtype=torch.float).to(inputs_embeds.dtype)
    pos = pos.unsqueeze(0).repeat(B, 1)
    h = inputs_embeds + position_embedding(cfg, pos, inputs_embeds) + torch.zeros_like(inputs_embeds)
    presents, attns = [], []
    for l in range(cfg.n_layer):
        p = f"h.{l}."
        a_in = F.layer_norm(h, (E,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], eps)
        a, present, w = _attention(sd, p + "attn.", cfg, a_in, past[l], use_cache, output_attentions)
        h = h + a
        m_in = F.layer_norm(h, (E,), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], eps)
        m = _conv1d(gelu_new(_conv1d(m_in, sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"])),
                    sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])
        h = h + m
        presents.append(present)
        attns.append(w)
    h = F.linear(F.layer_norm(h, (E,), sd["ln_f.weight"], sd["ln_f.bias"], eps),
                 sd["mlp_f.weight"], sd["mlp_f.bias"])                       # nn.Linear: W is [out,in]
    return (h, tuple(presents) if use_cache else None, tuple(attns) if output_attentions else None)


def gpt2_generate(sd: SD, cfg, inputs_embeds: Tensor, max_length: int, use_cache: bool = False):
    """GenerationMixin.generate/_generate_time_series.  Note: because
    `past=past` is always passed — even as None — only the LAST context token is ever fed to the model;
    with use_cache=False every step is therefore a one-token forward at position 0 (a Markov map);
    with use_cache=True the cache is cut to its last n_ctx-1 entries after every step and the position
    id is the number of cached keys.  Returns (embeds [B,max_length,E], presents | None)."""
    assert isinstance(max_length, int) and max_length > 0
    assert isinstance(use_cache, bool)
    cur_len = inputs_embeds.shape[1]
    assert cur_len < max_length
    out = inputs_embeds
    past = None
    presents = None
    while cur_len < max_length:
        x_last = out[:, -cfg.n_ctx:][:, -1].unsqueeze(1)
        hidden, presents, _ = gpt2_forward(sd, cfg, x_last, past=past, use_cache=use_cache)
        if use_cache:
            past = [pr[:, :, :, -(cfg.n_ctx - 1):] for pr in presents]
        out = torch.cat([out, hidden[:, -1:]], dim=1)
        cur_len += 1
    return out, presents


# ==================================================================================================
# Lorenz / Rossler embedding
# ==================================================================================================
def gpt2_train_grads(sd: SD, cfg, inputs_embeds: Tensor, labels_embeds: Tensor):
    """PhysformerTrain.forward + loss.backward() (phys_transformer_helpers.py:36-69, utils/trainer.py:244-256):
    loss = MSELoss()(transformer.forward(inputs_embeds)[0], labels_embeds).  Returns (loss, hidden, {name: grad})
    for every tensor of `sd` that the forward uses (wpe, attn.bias and masked_bias get no gradient)."""
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()
              if v.dtype.is_floating_point and not k.endswith("masked_bias")}      # buffers are not parameters
    full = dict(sd)
    full.update(leaves)
    hidden, _, _ = gpt2_forward(full, cfg, inputs_embeds, use_cache=False)
    loss = torch.nn.functional.mse_loss(hidden, labels_embeds)
    loss.backward()
    grads = {k: v.grad for k, v in leaves.items() if v.grad is not None}
    return loss.detach(), hidden.detach(), grads


def lorenz_embed(sd: SD, cfg, x: Tensor) -> Tensor:
    # embedding_lorenz.py:94-105 ; observableNet :41-47
    x = (x - sd["mu"].unsqueeze(0)) / sd["std"].unsqueeze(0)
    h = F.relu(F.linear(x, sd["observableNet.0.weight"], sd["observableNet.0.bias"]))
    g = F.linear(h, sd["observableNet.2.weight"], sd["observableNet.2.bias"])
    return F.layer_norm(g, (cfg.n_embd,), sd["observableNet.3.weight"], sd["observableNet.3.bias"],
                        cfg.layer_norm_epsilon)


def lorenz_recover(sd: SD, cfg, g: Tensor) -> Tensor:
    # embedding_lorenz.py:107-118 ; recoveryNet :49-53
    h = F.relu(F.linear(g, sd["recoveryNet.0.weight"], sd["recoveryNet.0.bias"]))
    y = F.linear(h, sd["recoveryNet.2.weight"], sd["recoveryNet.2.bias"])
    return sd["std"].unsqueeze(0) * y + sd["mu"].unsqueeze(0)


def lorenz_koopman_matrix(sd: SD, cfg) -> Tensor:
    """Dense skew-symmetric-plus-diagonal operator, embedding_lorenz.py:59-66, 130-137."""
    E = cfg.n_embd
    xidx = np.concatenate([np.arange(0, E - i) for i in range(1, 3)])
    yidx = np.concatenate([np.arange(i, E) for i in range(1, 3)])
    K = torch.zeros(E, E, dtype=sd["kMatrixUT"].dtype)
    K[xidx, yidx] = sd["kMatrixUT"]
    K[yidx, xidx] = -sd["kMatrixUT"]
    d = np.arange(E)
    K[d, d] = sd["kMatrixDiag"]
    return K


def lorenz_koopman(sd: SD, cfg, g: Tensor) -> Tensor:
    # embedding_lorenz.py:140-142
    K = lorenz_koopman_matrix(sd, cfg)
    return torch.bmm(K.expand(g.size(0), K.size(0), K.size(0)), g.unsqueeze(-1)).squeeze(-1)


def lorenz_train_loss(sd: SD, cfg, states: Tensor, w_recon: float = 1e4, w_dyn: float = 1.0,
                      w_reg: float = 1e-1) -> Tuple[Tensor, Tensor]:
    """LorenzEmbeddingTrainer.forward (embedding_lorenz.py:181-221).  Rossler uses w_recon=1e3
    (embedding_rossler.py:197,209).  Differentiable w.r.t. every floating tensor in `sd`."""
    mse = F.mse_loss
    x0 = states[:, 0]
    g0 = lorenz_embed(sd, cfg, x0)
    xr0 = lorenz_recover(sd, cfg, g0)
    loss = w_recon * mse(x0, xr0)
    loss_rec = mse(x0, xr0).detach()
    g_old = g0
    for t in range(1, states.shape[1]):
        xt = states[:, t]
        xr = lorenz_recover(sd, cfg, lorenz_embed(sd, cfg, xt))
        g_pred = lorenz_koopman(sd, cfg, g_old)
        xg = lorenz_recover(sd, cfg, g_pred)
        K = lorenz_koopman_matrix(sd, cfg)
        loss = loss + w_dyn * mse(xg, xt) + w_recon * mse(xr, xt) + w_reg * torch.sum(K ** 2)
        loss_rec = loss_rec + mse(xr, xt).detach()
        g_old = g_pred
    return loss, loss_rec


LORENZ_PARAM_ORDER = ["kMatrixDiag", "kMatrixUT",
                      "observableNet.0.weight", "observableNet.0.bias",
                      "observableNet.2.weight", "observableNet.2.bias",
                      "observableNet.3.weight", "observableNet.3.bias",
                      "recoveryNet.0.weight", "recoveryNet.0.bias",
                      "recoveryNet.2.weight", "recoveryNet.2.bias"]     # nn.Module.parameters() order


def lorenz_train_grads(sd: SD, cfg, states: Tensor, **kw):
    """loss, loss_reconstruct and the flat gradient in nn.Module.parameters() order."""
    params = {k: sd[k].clone().requires_grad_(True) for k in LORENZ_PARAM_ORDER}
    full = dict(sd)
    full.update(params)
    loss, loss_rec = lorenz_train_loss(full, cfg, states, **kw)
    grads = torch.autograd.grad(loss, [params[k] for k in LORENZ_PARAM_ORDER])
    return loss.detach(), loss_rec, torch.cat([g.reshape(-1) for g in grads])


def clip_adam_step(param: Tensor, grad: Tensor, m: Tensor, v: Tensor, step: int, lr: float = 1e-3,
                   beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, weight_decay: float = 1e-8,
                   max_norm: float = 0.1):
    """enn_trainer.py:97-99 on flat buffers: clip_grad_norm_(max_norm) (scale = max_norm/(norm+1e-6),
    clamped to 1) then torch.optim.Adam (L2 weight decay added to the gradient, bias correction).
    Returns (param', m', v', total_norm)."""
    total_norm = torch.linalg.vector_norm(grad, 2)
    coef = torch.clamp(max_norm / (total_norm + 1e-6), max=1.0)
    g = grad * coef
    g = g + weight_decay * param
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    param = param - (lr / bc1) * (m / denom)
    return param, m, v, total_norm


# ==================================================================================================
# Cylinder embedding
# ==================================================================================================
def _conv3x3_rep(x: Tensor, w: Tensor, b: Tensor, stride: int) -> Tensor:
    # nn.Conv2d(k=3, padding=1, padding_mode='replicate')  embedding_cylinder.py:43-59
    return F.conv2d(F.pad(x, (1, 1, 1, 1), mode="replicate"), w, b, stride=stride)


def cylinder_mask() -> Tensor:
    # embedding_cylinder.py:38-39
    X, Y = np.meshgrid(np.linspace(-2, 14, 128), np.linspace(-4, 4, 64))
    return torch.tensor(np.sqrt(X ** 2 + Y ** 2) < 1, dtype=torch.bool)


def cylinder_embed(sd: SD, cfg, x: Tensor, visc: Tensor) -> Tensor:
    # embedding_cylinder.py:138-154   (visc is [B,1])
    x = torch.cat([x, visc.unsqueeze(-1).unsqueeze(-1) * torch.ones_like(x[:, :1])], dim=1)
    x = (x - sd["mu"].view(1, -1, 1, 1)) / sd["std"].view(1, -1, 1, 1)
    for i, s in ((0, 2), (2, 2), (4, 2), (6, 2)):
        x = F.relu(_conv3x3_rep(x, sd[f"observableNet.{i}.weight"], sd[f"observableNet.{i}.bias"], s))
    x = _conv3x3_rep(x, sd["observableNet.8.weight"], sd["observableNet.8.bias"], 1)
    g = x.reshape(x.size(0), -1)
    return F.layer_norm(g, (cfg.n_embd,), sd["observableNetFC.0.weight"], sd["observableNetFC.0.bias"],
                        cfg.layer_norm_epsilon)


def cylinder_recover(sd: SD, cfg, g: Tensor, apply_mask: bool = True) -> Tensor:
    # embedding_cylinder.py:156-170 ; recoveryNet :70-88
    x = g.view(-1, cfg.n_embd // 32, 4, 8)
    for i in (1, 4, 7, 10):
        x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
        x = F.relu(_conv3x3_rep(x, sd[f"recoveryNet.{i}.weight"], sd[f"recoveryNet.{i}.bias"], 1))
    x = _conv3x3_rep(x, sd["recoveryNet.12.weight"], sd["recoveryNet.12.bias"], 1)
    x = sd["std"][:3].view(1, -1, 1, 1) * x + sd["mu"][:3].view(1, -1, 1, 1)
    if apply_mask:                      # recover() masks; forward() does NOT
        x = x.clone()
        x[cylinder_mask().repeat(x.size(0), x.size(1), 1, 1)] = 0
    return x


def _mlp1(sd: SD, p: str, v: Tensor) -> Tensor:
    return F.linear(F.relu(F.linear(v, sd[p + ".0.weight"], sd[p + ".0.bias"])),
                    sd[p + ".2.weight"], sd[p + ".2.bias"])


def cylinder_koopman_matrix(sd: SD, cfg, visc: Tensor) -> Tuple[Tensor, Tensor]:
    """Per-sample dense operator [B,E,E] and its diagonal [B,E] (embedding_cylinder.py:95-101,182-191)."""
    E = cfg.n_embd
    xidx = np.concatenate([np.arange(0, E - i) for i in range(1, 5)])
    yidx = np.concatenate([np.arange(i, E) for i in range(1, 5)])
    B = visc.size(0)
    K = torch.zeros(B, E, E, dtype=visc.dtype)
    K[:, xidx, yidx] = _mlp1(sd, "kMatrixUT", 100 * visc)
    K[:, yidx, xidx] = _mlp1(sd, "kMatrixLT", 100 * visc)
    diag = _mlp1(sd, "kMatrixDiagNet", 100 * visc)
    d = np.arange(E)
    K[:, d, d] = diag
    return K, diag


def cylinder_koopman(sd: SD, cfg, g: Tensor, visc: Tensor) -> Tensor:
    K, _ = cylinder_koopman_matrix(sd, cfg, visc)
    return torch.bmm(K, g.unsqueeze(-1)).squeeze(-1)


CYL_PARAM_ORDER = ([f"observableNet.{i}.{k}" for i in (0, 2, 4, 6, 8) for k in ("weight", "bias")] +
                   ["observableNetFC.0.weight", "observableNetFC.0.bias"] +
                   [f"recoveryNet.{i}.{k}" for i in (1, 4, 7, 10, 12) for k in ("weight", "bias")] +
                   [f"{net}.{i}.{k}" for net in ("kMatrixDiagNet", "kMatrixUT", "kMatrixLT") for i in (0, 2)
                    for k in ("weight", "bias")])                                  # nn.Module.parameters() order


def cylinder_train_loss(sd: SD, cfg, states: Tensor, visc: Tensor, w_rec: float = 1e1, w_dyn: float = 1e1,
                        w_reg: float = 1e-2) -> Tuple[Tensor, Tensor]:
    """CylinderEmbeddingTrainer.forward (embedding_cylinder.py:236-281).  states [B,T,3,64,128], visc [B,1].
    The per-step reconstruction comes from forward() (decoder WITHOUT the cylinder mask, fact 8), the Koopman
    prediction is decoded by recover() (WITH it); the regulariser sums K^2 over the per-sample operators."""
    mse = F.mse_loss
    x0 = states[:, 0]
    g0 = cylinder_embed(sd, cfg, x0, visc)
    xr0 = cylinder_recover(sd, cfg, g0, apply_mask=False)
    loss = w_rec * mse(x0, xr0)
    loss_rec = mse(x0, xr0).detach()
    g_old = g0
    for t in range(1, states.shape[1]):
        xt = states[:, t]
        xr = cylinder_recover(sd, cfg, cylinder_embed(sd, cfg, xt, visc), apply_mask=False)
        K, _ = cylinder_koopman_matrix(sd, cfg, visc)
        g_pred = torch.bmm(K, g_old.unsqueeze(-1)).squeeze(-1)
        xg = cylinder_recover(sd, cfg, g_pred, apply_mask=True)
        loss = loss + w_dyn * mse(xg, xt) + w_rec * mse(xr, xt) + w_reg * torch.sum(K ** 2)
        loss_rec = loss_rec + mse(xr, xt).detach()
        g_old = g_pred
    return loss, loss_rec


def cylinder_train_grads(sd: SD, cfg, states: Tensor, visc: Tensor, **kw):
    """loss, loss_reconstruct and the flat gradient in nn.Module.parameters() order."""
    params = {k: sd[k].clone().requires_grad_(True) for k in CYL_PARAM_ORDER}
    full = dict(sd)
    full.update(params)
    loss, loss_rec = cylinder_train_loss(full, cfg, states, visc, **kw)
    grads = torch.autograd.grad(loss, [params[k] for k in CYL_PARAM_ORDER], allow_unused=True)   # T = 1: no Koopman step
    grads = [g if g is not None else torch.zeros_like(params[k]) for g, k in zip(grads, CYL_PARAM_ORDER)]
    return loss.detach(), loss_rec, torch.cat([g.reshape(-1) for g in grads])


# ==================================================================================================
# helpers used by the tests
# ==================================================================================================
# ==================================================================================================
# Gray-Scott embedding (eval mode)
# ==================================================================================================
def _gs_block(sd: SD, p: str, i: int, x: Tensor, k: int, stride: int) -> Tensor:
    # nn.Conv3d(padding=k//2, padding_mode='circular') + nn.BatchNorm3d (running stats) + LeakyReLU(0.02)
    # embedding_grayscott.py:46-58,76-96
    pad = k // 2
    if pad:
        x = F.pad(x, (pad,) * 6, mode="circular")
    x = F.conv3d(x, sd[f"{p}.{i}.weight"], sd[f"{p}.{i}.bias"], stride=stride)
    b = f"{p}.{i + 1}"
    x = F.batch_norm(x, sd[b + ".running_mean"], sd[b + ".running_var"], sd[b + ".weight"], sd[b + ".bias"], False, 0.1, 1e-5)
    return F.leaky_relu(x, 0.02)


def grayscott_latent(sd: SD, cfg, x: Tensor) -> Tensor:
    # observableNet of embedding_grayscott.py:46-59 on the normalised state (:212-214)
    x = (x - sd["mu"]) / sd["std"]
    for i, k, s in ((0, 5, 2), (3, 3, 2), (6, 3, 2), (9, 1, 1)):
        x = _gs_block(sd, "observableNet", i, x, k, s)
    return x


def grayscott_embed(sd: SD, cfg, x: Tensor) -> Tensor:
    # embedding_grayscott.py:141-153, observableNetFC :61-66
    g0 = grayscott_latent(sd, cfg, x)
    h = F.leaky_relu(F.linear(g0.reshape(g0.size(0), -1), sd["observableNetFC.0.weight"], sd["observableNetFC.0.bias"]), 0.02)
    h = F.linear(h, sd["observableNetFC.2.weight"], sd["observableNetFC.2.bias"])
    return F.layer_norm(h, (cfg.n_embd,), sd["observableNetFC.3.weight"], sd["observableNetFC.3.bias"], cfg.layer_norm_epsilon)


def grayscott_decode_latent(sd: SD, cfg, z: Tensor) -> Tensor:
    # recoveryNet of embedding_grayscott.py:75-98 + _unnormalize (:216-218)
    x = _gs_block(sd, "recoveryNet", 0, z, 1, 1)
    for i in (4, 8, 12):
        x = F.interpolate(x, scale_factor=2, mode="trilinear", align_corners=False)
        x = _gs_block(sd, "recoveryNet", i, x, 3, 1)
    x = F.conv3d(x, sd["recoveryNet.15.weight"], sd["recoveryNet.15.bias"])
    return sd["std"] * x + sd["mu"]


def grayscott_recover_latent(sd: SD, cfg, g: Tensor) -> Tensor:
    # recoveryNetFC of embedding_grayscott.py:68-73 (LeakyReLU(1.0) is the identity)
    h = F.linear(g, sd["recoveryNetFC.0.weight"], sd["recoveryNetFC.0.bias"])
    h = F.leaky_relu(F.linear(h, sd["recoveryNetFC.2.weight"], sd["recoveryNetFC.2.bias"]), 0.02)
    return h.view(-1, 64, 4, 4, 4)


def grayscott_recover(sd: SD, cfg, g: Tensor) -> Tensor:
    # embedding_grayscott.py:155-167
    return grayscott_decode_latent(sd, cfg, grayscott_recover_latent(sd, cfg, g))


def grayscott_koopman_matrix(sd: SD, cfg) -> Tensor:
    # embedding_grayscott.py:100-107,179-186: skew-symmetric bands 1..9 + diagonal
    E = cfg.n_embd
    xidx = np.concatenate([np.arange(0, E - i) for i in range(1, 10)])
    yidx = np.concatenate([np.arange(i, E) for i in range(1, 10)])
    K = torch.zeros(E, E, dtype=sd["kMatrixUT"].dtype)
    K[xidx, yidx] = sd["kMatrixUT"]
    K[yidx, xidx] = -sd["kMatrixUT"]
    d = np.arange(E)
    K[d, d] = sd["kMatrixDiag"]
    return K


def grayscott_koopman(sd: SD, cfg, g: Tensor) -> Tensor:
    K = grayscott_koopman_matrix(sd, cfg)
    return torch.bmm(K.unsqueeze(0).expand(g.size(0), -1, -1), g.unsqueeze(-1)).squeeze(-1)


def rel_l2(a, b) -> float:
    """|| a - b ||_2 / || b ||_2 over the whole tensor (SURVEY.md §8d parity thresholds)."""
    a = torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a).double().reshape(-1)
    b = torch.as_tensor(np.asarray(b) if not torch.is_tensor(b) else b).double().reshape(-1)
    return float(torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b).clamp_min(1e-300))
