"""The oracle against the UNMODIFIED reference imported live from /root/reference (skipped where the tree is absent,
e.g. on the GPU box).  tests/test_oracle_golden.py pins the oracle on recorded reference outputs; this file repeats
the comparison on fresh seeds and shapes that have no fixture, so a fixture can never be the only witness."""
import numpy as np
import pytest
import torch

from oracle import ref_shim
from oracle import trpx_oracle as O
from tests import synth

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")
TOL = 2e-6


def _t(sd):
    return {k: torch.from_numpy(np.array(v)) for k, v in sd.items()}


@pytest.mark.parametrize("kind,over,sigma,B,ctx,max_length", [
    ("rossler", {}, 0.1, 5, 1, 70), ("lorenz", {"n_ctx": 16}, 0.2, 3, 2, 40), ("cylinder", {"n_layer": 2}, 0.05, 2, 1, 36)])
def test_generate_and_forward_live(kind, over, sigma, B, ctx, max_length):
    ref_shim.load()
    from trphysx.config.configuration_phys import PhysConfig
    from trphysx.transformer import PhysformerGPT2
    cfg = synth.make_config(kind, **over)
    w = synth.gpt2_weights(cfg, 4711, stress=True, stress_sigma=sigma)
    model = PhysformerGPT2(PhysConfig(**vars(cfg)), "live").eval()
    model.load_state_dict(_t(w), strict=True)
    sd = O.to_torch(w)
    x = torch.from_numpy(synth.normal_embeds((B, cfg.n_ctx, cfg.n_embd), 4712))
    c = torch.from_numpy(synth.normal_embeds((B, ctx, cfg.n_embd), 4713))
    with torch.no_grad():
        hid, pres = model(x, use_cache=True)
        o_hid, o_pres, _ = O.gpt2_forward(sd, cfg, x, use_cache=True)
        assert O.rel_l2(o_hid, hid) < TOL
        for a, b in zip(o_pres, pres):
            assert O.rel_l2(a, b) < TOL
        for uc in (True, False):
            r = model.generate(c, max_length=max_length, use_cache=uc)
            o, op = O.gpt2_generate(sd, cfg, c, max_length, use_cache=uc)
            assert O.rel_l2(o, r[0]) < 20 * TOL
            if uc:
                assert O.rel_l2(op[-1], r[1][-1]) < 20 * TOL


def test_embeddings_live():
    ref_shim.load()
    from trphysx.config.configuration_phys import PhysConfig
    from trphysx.embedding import CylinderEmbedding, LorenzEmbedding
    cfg = synth.make_config("lorenz")
    w = synth.lorenz_embedding_weights(cfg, 4720)
    m = LorenzEmbedding(PhysConfig(**vars(cfg))).eval()
    m.load_state_dict(_t(w), strict=True)
    x = torch.from_numpy(synth.lorenz_trajectories(33, 2, seed=4721)[:, -1])
    sd = O.to_torch(w)
    with torch.no_grad():
        g = m.embed(x)
        assert O.rel_l2(O.lorenz_embed(sd, cfg, x), g) < TOL
        assert O.rel_l2(O.lorenz_recover(sd, cfg, g), m.recover(g)) < TOL
        assert O.rel_l2(O.lorenz_koopman(sd, cfg, g), m.koopmanOperation(g)) < TOL
    ccfg = synth.make_config("cylinder")
    cw = synth.cylinder_embedding_weights(ccfg, 4722)
    cm = CylinderEmbedding(PhysConfig(**vars(ccfg))).eval()
    cm.load_state_dict(_t(cw), strict=True)
    fx, visc = synth.cylinder_fields(2, seed=4723)
    fx, visc = torch.from_numpy(fx), torch.from_numpy(visc)
    csd = O.to_torch(cw)
    with torch.no_grad():
        g = cm.embed(fx, visc)
        assert O.rel_l2(O.cylinder_embed(csd, ccfg, fx, visc), g) < 5 * TOL
        assert O.rel_l2(O.cylinder_recover(csd, ccfg, g), cm.recover(g)) < 5 * TOL
        assert O.rel_l2(O.cylinder_koopman(csd, ccfg, g, visc), cm.koopmanOperation(g, visc)) < 5 * TOL
