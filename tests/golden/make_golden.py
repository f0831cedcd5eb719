"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference).

Run here (CPU container, reference present):   python tests/golden/make_golden.py
Weights and inputs come from tests/synth.py (numpy default_rng, seeded) so only the reference's
outputs are stored.  Every case is listed in CASES below; tests/test_oracle_golden.py and the -m gpu
parity tests iterate over the same table.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import synth
from oracle import ref_shim  # noqa: E402
from tests.golden.cases import (weights_kw, GPT2_CASES, LORENZ_CASES, CYL_CASES, TRAIN_CASES, GPT2_TRAIN_CASES,  # noqa: E402
                                CYL_TRAIN_CASES, GS_CASES)

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(1)          # deterministic reduction order for the fixtures
torch.manual_seed(0)


def _t(sd):
    return {k: torch.from_numpy(np.array(v)) for k, v in sd.items()}


def gpt2_cases(trphysx, only=None):
    from trphysx.transformer import PhysformerGPT2
    from trphysx.config.configuration_phys import PhysConfig
    for name, c in GPT2_CASES.items():
        if only and name not in only:
            continue
        cfg_ns = synth.make_config(c["kind"], **c.get("over", {}))
        cfg = PhysConfig(**vars(cfg_ns))
        model = PhysformerGPT2(cfg, "golden").eval()
        model.load_state_dict(_t(synth.gpt2_weights(cfg_ns, c["seed"], **weights_kw(c))), strict=True)
        out = {}
        B, T = c.get("fwdB", c["B"]), c["T"]
        x = torch.from_numpy(synth.normal_embeds((B, T, cfg.n_embd), c["seed"] + 1))
        with torch.no_grad():
            hid, presents = model(x, use_cache=True)
            out["fwd_hidden"] = hid.numpy()
            out["fwd_present_first"] = presents[0].numpy()
            out["fwd_present_last"] = presents[-1].numpy()
            # continuation with a past (teacher-forced second chunk), when it fits the context
            if c.get("T2"):
                T2 = c["T2"]
                x2 = torch.from_numpy(synth.normal_embeds((B, T2, cfg.n_embd), c["seed"] + 2))
                past = [p[:, :, :, :c["Tp"]] for p in presents]
                hid2, pres2 = model(x2, past=past, use_cache=True)
                out["fwd2_hidden"] = hid2.numpy()
                out["fwd2_present_last"] = pres2[-1].numpy()
            ctx = torch.from_numpy(synth.normal_embeds((c["B"], c["ctx"], cfg.n_embd), c["seed"] + 3))
            g1 = model.generate(ctx, max_length=c["max_length"], use_cache=True)
            out["gen_cache"] = g1[0].numpy()
            out["gen_cache_present_last"] = g1[1][-1].numpy()
            g0 = model.generate(ctx, max_length=c["max_length"], use_cache=False)
            out["gen_nocache"] = g0[0].numpy()
        np.savez_compressed(os.path.join(OUT, f"gpt2_{name}.npz"), **out)
        print("gpt2", name, {k: v.shape for k, v in out.items()})


def lorenz_cases(trphysx):
    from trphysx.embedding import LorenzEmbedding
    from trphysx.config.configuration_phys import PhysConfig
    RosslerConfig, RosslerEmbedding, _ = ref_shim.rossler_modules()
    for name, c in LORENZ_CASES.items():
        cfg_ns = synth.make_config(c["kind"])
        cfg = PhysConfig(**vars(cfg_ns))
        model = (LorenzEmbedding if c["kind"] == "lorenz" else RosslerEmbedding)(cfg).eval()
        model.load_state_dict(_t(synth.lorenz_embedding_weights(cfg_ns, c["seed"])), strict=True)
        traj = (synth.lorenz_trajectories if c["kind"] == "lorenz" else synth.rossler_trajectories)(
            c["N"], 4, seed=c["seed"])
        x = torch.from_numpy(traj[:, -1])
        with torch.no_grad():
            g = model.embed(x)
            g2, xhat = model(x)
            out = dict(embed=g.numpy(), recover=model.recover(g).numpy(), fwd_g=g2.numpy(),
                       fwd_xhat=xhat.numpy(), koopman=model.koopmanOperation(g).numpy(),
                       kmatrix=model.koopmanOperator.detach().numpy())
        np.savez_compressed(os.path.join(OUT, f"emb_{name}.npz"), **out)
        print("emb", name, {k: v.shape for k, v in out.items()})


def cyl_cases(trphysx):
    from trphysx.embedding import CylinderEmbedding
    from trphysx.config.configuration_phys import PhysConfig
    for name, c in CYL_CASES.items():
        cfg_ns = synth.make_config("cylinder")
        cfg = PhysConfig(**vars(cfg_ns))
        model = CylinderEmbedding(cfg).eval()
        model.load_state_dict(_t(synth.cylinder_embedding_weights(cfg_ns, c["seed"])), strict=True)
        x, visc = synth.cylinder_fields(c["N"], seed=c["seed"])
        x, visc = torch.from_numpy(x), torch.from_numpy(visc)
        with torch.no_grad():
            g = model.embed(x, visc)
            g2, xhat = model(x, visc)
            rec = model.recover(g)
            kop = model.koopmanOperation(g, visc)
            out = dict(embed=g.numpy(), recover=rec.numpy().astype(np.float32),
                       fwd_xhat_ch0=xhat[:, 0].numpy(),       # forward() does NOT mask (fact 8)
                       koopman=kop.numpy(), kdiag=model.koopmanDiag.numpy(),
                       kmatrix0=model.koopmanOperator[0].detach().numpy())
        np.savez_compressed(os.path.join(OUT, f"cyl_{name}.npz"), **out)
        print("cyl", name, {k: v.shape for k, v in out.items()})


def gs_cases(trphysx):
    from trphysx.embedding.embedding_grayscott import GrayScottEmbedding
    from trphysx.config.configuration_phys import PhysConfig
    for name, c in GS_CASES.items():
        cfg_ns = synth.make_config("grayscott")
        cfg = PhysConfig(**vars(cfg_ns))
        model = GrayScottEmbedding(cfg).eval()
        model.load_state_dict(_t(synth.grayscott_embedding_weights(cfg_ns, c["seed"])), strict=True)
        x = torch.from_numpy(synth.grayscott_fields(c["N"], seed=c["seed"]))
        with torch.no_grad():
            g = model.embed(x)
            f_g, f_xhat, f_g0, f_out0, f_x2 = model(x)
            rec = model.recover(g)
            kop = model.koopmanOperation(g)
            out = dict(embed=g.numpy(), recover=rec.numpy(), fwd_g0=f_g0.numpy(), fwd_out0=f_out0.numpy(),
                       fwd_xhat_ch0=f_xhat[:, 0].numpy(), fwd_x2_ch1=f_x2[:, 1].numpy(), koopman=kop.numpy(),
                       kmatrix0=model.koopmanOperator[0].detach().numpy())
        np.savez_compressed(os.path.join(OUT, f"gs_{name}.npz"), **out)
        print("gs", name, {k: v.shape for k, v in out.items()})


def train_cases(trphysx):
    from trphysx.embedding.embedding_lorenz import LorenzEmbeddingTrainer
    from trphysx.config.configuration_phys import PhysConfig
    RosslerConfig, RosslerEmbedding, RosslerEmbeddingTrainer = ref_shim.rossler_modules()
    for name, c in TRAIN_CASES.items():
        cfg_ns = synth.make_config(c["kind"])
        cfg = PhysConfig(**vars(cfg_ns))
        head = (LorenzEmbeddingTrainer if c["kind"] == "lorenz" else RosslerEmbeddingTrainer)(cfg)
        head.embedding_model.load_state_dict(_t(synth.lorenz_embedding_weights(cfg_ns, c["seed"])),
                                             strict=True)
        gen = synth.lorenz_trajectories if c["kind"] == "lorenz" else synth.rossler_trajectories
        states = torch.from_numpy(gen(c["B"], c["T"] - 1, seed=c["seed"]))
        opt = torch.optim.Adam(head.parameters(), lr=1e-3, weight_decay=1e-8)
        out = {}
        for it in range(c["iters"]):          # the step body of enn_trainer.py:91-102
            loss, loss_rec = head(states)
            loss = loss.sum()
            loss.backward()
            flat = torch.cat([p.grad.reshape(-1) for p in head.parameters()])
            norm = torch.nn.utils.clip_grad_norm_(head.parameters(), 0.1)
            opt.step()
            opt.zero_grad()
            out[f"loss{it}"] = loss.detach().numpy()
            out[f"loss_rec{it}"] = loss_rec.numpy()
            out[f"grad{it}"] = flat.numpy()
            out[f"norm{it}"] = norm.numpy()
        out["params_after"] = torch.cat([p.detach().reshape(-1) for p in head.parameters()]).numpy()
        np.savez_compressed(os.path.join(OUT, f"train_{name}.npz"), **out)
        print("train", name, {k: v.shape for k, v in out.items()})


def gpt2_train_cases(trphysx):
    from trphysx.transformer import PhysformerGPT2, PhysformerTrain
    from trphysx.config.configuration_phys import PhysConfig
    for name, c in GPT2_TRAIN_CASES.items():
        cfg_ns = synth.make_config(c["kind"])
        cfg = PhysConfig(**vars(cfg_ns))
        head = PhysformerTrain(cfg, PhysformerGPT2(cfg, "golden"))
        head.transformer.load_state_dict(_t(synth.gpt2_weights(cfg_ns, c["seed"], stress=c["stress"])), strict=True)
        head.train()
        B, T = c["B"], c["T"]
        x = torch.from_numpy(synth.normal_embeds((B, T, cfg.n_embd), c["seed"] + 1))
        y = torch.from_numpy(synth.normal_embeds((B, T, cfg.n_embd), c["seed"] + 2))
        outputs = head(inputs_embeds=x, labels_embeds=y)          # the step body of utils/trainer.py:244-256
        loss = outputs[0]
        loss.backward()
        out = {"loss": loss.detach().numpy(), "hidden": outputs[1].detach().numpy()}
        keep = c.get("keep")                                       # wide models: a subset keeps the fixture small
        for k, p_ in head.transformer.named_parameters():
            if p_.grad is not None and (keep is None or any(k.startswith(pre) for pre in keep)):
                out["grad." + k] = p_.grad.numpy()
        np.savez_compressed(os.path.join(OUT, f"gpt2train_{name}.npz"), **out)
        print("gpt2train", name, float(loss), len(out))


def cyl_train_cases(trphysx):
    from trphysx.embedding.embedding_cylinder import CylinderEmbeddingTrainer
    from trphysx.config.configuration_phys import PhysConfig
    for name, c in CYL_TRAIN_CASES.items():
        cfg_ns = synth.make_config("cylinder")
        head = CylinderEmbeddingTrainer(PhysConfig(**vars(cfg_ns)))
        head.embedding_model.load_state_dict(_t(synth.cylinder_embedding_weights(cfg_ns, c["seed"])), strict=True)
        B, T = c["B"], c["T"]
        x, visc = synth.cylinder_fields(B * T, seed=c["seed"])
        states = torch.from_numpy(x).view(B, T, 3, 64, 128)
        visc = torch.from_numpy(visc).view(B, T, 1)[:, 0]
        loss, loss_rec = head(states, visc)                   # the step body of enn_trainer.py:93-96
        loss = loss.sum()
        loss.backward()
        flat = torch.cat([p.grad.reshape(-1) for p in head.parameters()])
        norm = torch.nn.utils.clip_grad_norm_(head.parameters(), 0.1)
        np.savez_compressed(os.path.join(OUT, f"cyltrain_{name}.npz"), loss=loss.detach().numpy(), loss_rec=loss_rec.numpy(),
                            grad=flat.numpy(), norm=norm.numpy())
        print("cyltrain", name, float(loss), float(loss_rec), flat.shape, float(norm))


if __name__ == "__main__":
    t = ref_shim.load()
    if len(sys.argv) > 1 and sys.argv[1] == "gpt2train":
        gpt2_train_cases(t)                                        # added later: leaves the other fixtures untouched
    elif len(sys.argv) > 1 and sys.argv[1] == "cyltrain":
        cyl_train_cases(t)
    elif len(sys.argv) > 1 and sys.argv[1] == "gs":
        gs_cases(t)
    elif len(sys.argv) > 2 and sys.argv[1] == "gpt2":              # python make_golden.py gpt2 <case> [<case> ...]
        gpt2_cases(t, only=sys.argv[2:])
    else:
        cyl_train_cases(t)
        gpt2_cases(t)
        lorenz_cases(t)
        cyl_cases(t)
        train_cases(t)
        gpt2_train_cases(t)
        gs_cases(t)
