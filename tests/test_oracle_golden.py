"""The oracle (oracle/trpx_oracle.py) against the reference's recorded outputs (tests/golden/*.npz,
made by tests/golden/make_golden.py from the unmodified reference).  CPU only."""
import os

import numpy as np
import pytest
import torch

from tests import synth
from oracle import trpx_oracle as O
from tests.golden.cases import (weights_kw, GPT2_CASES, LORENZ_CASES, CYL_CASES, TRAIN_CASES, GPT2_TRAIN_CASES,
                                CYL_TRAIN_CASES, GS_CASES)

TOL = 2e-6          # both sides are torch-CPU fp32 executing the same aten ops; only threading differs


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize("name", list(GPT2_CASES))
def test_gpt2_oracle_matches_reference(golden_dir, name):
    c = GPT2_CASES[name]
    gold = _load(golden_dir, f"gpt2_{name}.npz")
    cfg = synth.make_config(c["kind"], **c.get("over", {}))
    sd = O.to_torch(synth.gpt2_weights(cfg, c["seed"], **weights_kw(c)))
    B, T, E = c.get("fwdB", c["B"]), c["T"], cfg.n_embd
    x = torch.from_numpy(synth.normal_embeds((B, T, E), c["seed"] + 1))
    hid, presents, _ = O.gpt2_forward(sd, cfg, x, use_cache=True)
    assert O.rel_l2(hid, gold["fwd_hidden"]) < TOL
    assert O.rel_l2(presents[0], gold["fwd_present_first"]) < TOL
    assert O.rel_l2(presents[-1], gold["fwd_present_last"]) < TOL
    x2 = torch.from_numpy(synth.normal_embeds((B, c["T2"], E), c["seed"] + 2))
    past = [p[:, :, :, :c["Tp"]] for p in presents]
    hid2, pres2, _ = O.gpt2_forward(sd, cfg, x2, past=past, use_cache=True)
    assert O.rel_l2(hid2, gold["fwd2_hidden"]) < TOL
    assert O.rel_l2(pres2[-1], gold["fwd2_present_last"]) < TOL
    ctx = torch.from_numpy(synth.normal_embeds((c["B"], c["ctx"], E), c["seed"] + 3))
    g1, pres = O.gpt2_generate(sd, cfg, ctx, c["max_length"], use_cache=True)
    assert g1.shape == gold["gen_cache"].shape
    assert O.rel_l2(g1, gold["gen_cache"]) < 20 * TOL
    assert O.rel_l2(pres[-1], gold["gen_cache_present_last"]) < 20 * TOL
    g0, none = O.gpt2_generate(sd, cfg, ctx, c["max_length"], use_cache=False)
    assert none is None
    assert O.rel_l2(g0, gold["gen_nocache"]) < 20 * TOL


def test_generate_uses_only_last_context_token():
    """contexts differing in all but the last token generate identical output."""
    cfg = synth.make_config("lorenz", n_ctx=8)
    sd = O.to_torch(synth.gpt2_weights(cfg, 7, stress=True))
    a = torch.from_numpy(synth.normal_embeds((2, 4, 32), 1))
    b = torch.from_numpy(synth.normal_embeds((2, 4, 32), 2))
    b[:, -1] = a[:, -1]
    for uc in (False, True):
        ga, _ = O.gpt2_generate(sd, cfg, a, 20, use_cache=uc)
        gb, _ = O.gpt2_generate(sd, cfg, b, 20, use_cache=uc)
        assert torch.equal(ga[:, 4:], gb[:, 4:])


def test_position_embedding_quirk():
    """fact 4: odd channels are cos(p) for every channel pair."""
    cfg = synth.make_config("lorenz")
    pos = torch.arange(0, 64, dtype=torch.float).unsqueeze(0)
    pe = O.position_embedding(cfg, pos, torch.zeros(1, 64, 32))
    for j in range(1, 32, 2):
        assert torch.allclose(pe[0, :, j], torch.cos(pos[0]), atol=1e-6)
    assert torch.allclose(pe[0, :, 2], torch.sin(pos[0] / 10000 ** (2.0 / 32)), atol=1e-6)


@pytest.mark.parametrize("name", list(LORENZ_CASES))
def test_lorenz_embedding_oracle(golden_dir, name):
    c = LORENZ_CASES[name]
    gold = _load(golden_dir, f"emb_{name}.npz")
    cfg = synth.make_config(c["kind"])
    sd = O.to_torch(synth.lorenz_embedding_weights(cfg, c["seed"]))
    gen = synth.lorenz_trajectories if c["kind"] == "lorenz" else synth.rossler_trajectories
    x = torch.from_numpy(gen(c["N"], 4, seed=c["seed"])[:, -1])
    g = O.lorenz_embed(sd, cfg, x)
    assert O.rel_l2(g, gold["embed"]) < TOL
    assert O.rel_l2(g, gold["fwd_g"]) < TOL
    assert O.rel_l2(O.lorenz_recover(sd, cfg, g), gold["recover"]) < TOL
    assert O.rel_l2(O.lorenz_recover(sd, cfg, g), gold["fwd_xhat"]) < TOL
    assert O.rel_l2(O.lorenz_koopman(sd, cfg, g), gold["koopman"]) < TOL
    assert np.array_equal(O.lorenz_koopman_matrix(sd, cfg).numpy(), gold["kmatrix"])


@pytest.mark.parametrize("name", list(CYL_CASES))
def test_cylinder_embedding_oracle(golden_dir, name):
    c = CYL_CASES[name]
    gold = _load(golden_dir, f"cyl_{name}.npz")
    cfg = synth.make_config("cylinder")
    sd = O.to_torch(synth.cylinder_embedding_weights(cfg, c["seed"]))
    x, visc = synth.cylinder_fields(c["N"], seed=c["seed"])
    x, visc = torch.from_numpy(x), torch.from_numpy(visc)
    g = O.cylinder_embed(sd, cfg, x, visc)
    assert O.rel_l2(g, gold["embed"]) < 5 * TOL
    rec = O.cylinder_recover(sd, cfg, g)
    assert O.rel_l2(rec, gold["recover"]) < 5 * TOL
    assert int((rec[0, 0] == 0).sum()) >= 196           # the masked cylinder cells (fact 8)
    nomask = O.cylinder_recover(sd, cfg, g, apply_mask=False)
    assert O.rel_l2(nomask[:, 0], gold["fwd_xhat_ch0"]) < 5 * TOL
    assert O.rel_l2(O.cylinder_koopman(sd, cfg, g, visc), gold["koopman"]) < 5 * TOL
    K, diag = O.cylinder_koopman_matrix(sd, cfg, visc)
    assert O.rel_l2(diag, gold["kdiag"]) < TOL
    assert O.rel_l2(K[0], gold["kmatrix0"]) < TOL


@pytest.mark.parametrize("name", list(GS_CASES))
def test_grayscott_embedding_oracle(golden_dir, name):
    """Gray-Scott embedding (eval mode) restated in oracle/ against the unmodified reference's outputs."""
    c = GS_CASES[name]
    gold = _load(golden_dir, f"gs_{name}.npz")
    cfg = synth.make_config("grayscott")
    sd = O.to_torch(synth.grayscott_embedding_weights(cfg, c["seed"]))
    x = torch.from_numpy(synth.grayscott_fields(c["N"], seed=c["seed"]))
    g0 = O.grayscott_latent(sd, cfg, x)
    assert O.rel_l2(g0, gold["fwd_g0"]) < 5 * TOL
    g = O.grayscott_embed(sd, cfg, x)
    assert O.rel_l2(g, gold["embed"]) < 5 * TOL
    assert O.rel_l2(O.grayscott_recover_latent(sd, cfg, g), gold["fwd_out0"]) < 5 * TOL
    rec = O.grayscott_recover(sd, cfg, g)
    assert O.rel_l2(rec, gold["recover"]) < 5 * TOL
    assert O.rel_l2(rec[:, 0], gold["fwd_xhat_ch0"]) < 5 * TOL
    assert O.rel_l2(O.grayscott_decode_latent(sd, cfg, g0)[:, 1], gold["fwd_x2_ch1"]) < 5 * TOL
    assert O.rel_l2(O.grayscott_koopman(sd, cfg, g), gold["koopman"]) < 5 * TOL
    assert np.array_equal(O.grayscott_koopman_matrix(sd, cfg).numpy(), gold["kmatrix0"])


@pytest.mark.parametrize("name", list(TRAIN_CASES))
def test_lorenz_train_step_oracle(golden_dir, name):
    c = TRAIN_CASES[name]
    gold = _load(golden_dir, f"train_{name}.npz")
    cfg = synth.make_config(c["kind"])
    sd = O.to_torch(synth.lorenz_embedding_weights(cfg, c["seed"]))
    gen = synth.lorenz_trajectories if c["kind"] == "lorenz" else synth.rossler_trajectories
    states = torch.from_numpy(gen(c["B"], c["T"] - 1, seed=c["seed"]))
    w_recon = 1e4 if c["kind"] == "lorenz" else 1e3
    flat = torch.cat([sd[k].reshape(-1) for k in O.LORENZ_PARAM_ORDER])
    m = torch.zeros_like(flat)
    v = torch.zeros_like(flat)
    sizes = [sd[k].numel() for k in O.LORENZ_PARAM_ORDER]
    for it in range(c["iters"]):
        loss, loss_rec, grad = O.lorenz_train_grads(sd, cfg, states, w_recon=w_recon)
        assert abs(float(loss) - float(gold[f"loss{it}"])) <= 1e-5 * abs(float(gold[f"loss{it}"]))
        assert abs(float(loss_rec) - float(gold[f"loss_rec{it}"])) <= 1e-5 * abs(float(gold[f"loss_rec{it}"]))
        assert O.rel_l2(grad, gold[f"grad{it}"]) < 1e-5
        flat, m, v, norm = O.clip_adam_step(flat, grad, m, v, it + 1)
        assert abs(float(norm) - float(gold[f"norm{it}"])) <= 1e-5 * float(gold[f"norm{it}"])
        for k, t in zip(O.LORENZ_PARAM_ORDER, torch.split(flat, sizes)):
            sd[k] = t.view_as(sd[k]).clone()
    assert O.rel_l2(flat, gold["params_after"]) < 1e-6


@pytest.mark.parametrize("name", list(GPT2_TRAIN_CASES))
def test_gpt2_train_step_oracle(golden_dir, name):
    """Teacher-forced loss and parameter gradients against the reference's recorded backward()."""
    c = GPT2_TRAIN_CASES[name]
    gold = _load(golden_dir, f"gpt2train_{name}.npz")
    cfg = synth.make_config(c["kind"])
    sd = O.to_torch(synth.gpt2_weights(cfg, c["seed"], stress=c["stress"]))
    x = torch.from_numpy(synth.normal_embeds((c["B"], c["T"], cfg.n_embd), c["seed"] + 1))
    y = torch.from_numpy(synth.normal_embeds((c["B"], c["T"], cfg.n_embd), c["seed"] + 2))
    loss, hidden, grads = O.gpt2_train_grads(sd, cfg, x, y)
    assert abs(float(loss) - float(gold["loss"])) <= 2e-6 * abs(float(gold["loss"]))
    assert O.rel_l2(hidden, gold["hidden"]) < TOL
    keys = [k[5:] for k in gold.files if k.startswith("grad.")]
    assert keys and all(k in grads for k in keys)
    if c.get("keep") is None:
        assert sorted(keys) == sorted(grads)              # same set of tensors receives a gradient
    for k in keys:
        assert O.rel_l2(grads[k], gold["grad." + k]) < 5 * TOL, k


@pytest.mark.parametrize("name", list(CYL_TRAIN_CASES))
def test_cylinder_train_step_oracle(golden_dir, name):
    """Cylinder Koopman training loss and flat gradient against the reference's recorded backward()."""
    c = CYL_TRAIN_CASES[name]
    gold = _load(golden_dir, f"cyltrain_{name}.npz")
    cfg = synth.make_config("cylinder")
    sd = O.to_torch(synth.cylinder_embedding_weights(cfg, c["seed"]))
    B, T = c["B"], c["T"]
    x, visc = synth.cylinder_fields(B * T, seed=c["seed"])
    states = torch.from_numpy(x).view(B, T, 3, 64, 128)
    visc = torch.from_numpy(visc).view(B, T, 1)[:, 0]
    loss, loss_rec, grad = O.cylinder_train_grads(sd, cfg, states, visc)
    assert abs(float(loss) - float(gold["loss"])) <= 2e-6 * float(gold["loss"])
    assert abs(float(loss_rec) - float(gold["loss_rec"])) <= 2e-6 * float(gold["loss_rec"])
    assert grad.numel() == gold["grad"].size == 262535
    assert O.rel_l2(grad, gold["grad"]) < 5e-6
