"""Helpers shared by the parity tests: build the B200 modules from the oracle/synth.py weights."""
import numpy as np
import torch

from oracle import synth


def torch_sd(sd_np):
    return {k: torch.from_numpy(np.array(v)) for k, v in sd_np.items()}


def make_gpt2(cfg_ns, seed, stress, device=None):
    from trphysx_b200.config import PhysConfig
    from trphysx_b200.transformer import PhysformerGPT2
    cfg = PhysConfig(**vars(cfg_ns))
    m = PhysformerGPT2(cfg, "test").eval()
    m.load_state_dict(torch_sd(synth.gpt2_weights(cfg_ns, seed, stress=stress)), strict=True)
    return m.to(device) if device is not None else m


def make_lorenz(cfg_ns, seed, kind="lorenz", device=None):
    from trphysx_b200.config import PhysConfig
    from trphysx_b200.embedding import LorenzEmbedding, RosslerEmbedding
    cfg = PhysConfig(**vars(cfg_ns))
    m = (LorenzEmbedding if kind == "lorenz" else RosslerEmbedding)(cfg).eval()
    m.load_state_dict(torch_sd(synth.lorenz_embedding_weights(cfg_ns, seed)), strict=True)
    return m.to(device) if device is not None else m


def make_cylinder(cfg_ns, seed, device=None):
    from trphysx_b200.config import PhysConfig
    from trphysx_b200.embedding import CylinderEmbedding
    cfg = PhysConfig(**vars(cfg_ns))
    m = CylinderEmbedding(cfg).eval()
    m.load_state_dict(torch_sd(synth.cylinder_embedding_weights(cfg_ns, seed)), strict=True)
    return m.to(device) if device is not None else m
