"""-m gpu: embedding kernels (Lorenz/Rossler MLP, Cylinder convs, Koopman operators, training step) against the
reference's recorded outputs and the CPU oracle.  FP32 tolerance rel-L2 <= 1e-5."""
import os

import numpy as np
import pytest
import torch

from tests import synth
from oracle import trpx_oracle as O
from tests.golden.cases import LORENZ_CASES, CYL_CASES, TRAIN_CASES, CYL_TRAIN_CASES, GS_CASES
from tests.util import make_lorenz, make_cylinder, make_grayscott, torch_sd

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("name", list(LORENZ_CASES))
def test_lorenz_embedding_matches_reference(golden_dir, dev, name):
    c = LORENZ_CASES[name]
    gold = np.load(os.path.join(golden_dir, f"emb_{name}.npz"))
    cfg = synth.make_config(c["kind"])
    m = make_lorenz(cfg, c["seed"], c["kind"], dev)
    gen = synth.lorenz_trajectories if c["kind"] == "lorenz" else synth.rossler_trajectories
    x = torch.from_numpy(gen(c["N"], 4, seed=c["seed"])[:, -1]).to(dev)
    with torch.no_grad():
        g = m.embed(x)
        assert O.rel_l2(g.cpu(), gold["embed"]) < TOL
        assert O.rel_l2(m.recover(g).cpu(), gold["recover"]) < TOL
        g2, xhat = m(x)
        assert O.rel_l2(g2.cpu(), gold["fwd_g"]) < TOL and O.rel_l2(xhat.cpu(), gold["fwd_xhat"]) < TOL
        assert O.rel_l2(m.koopmanOperation(g).cpu(), gold["koopman"]) < TOL
        assert np.array_equal(m.koopmanOperator.cpu().numpy(), gold["kmatrix"])
        assert m.koopmanDiag is m.kMatrixDiag


@pytest.mark.parametrize("N", [1, 2, 255, 100_003])
def test_lorenz_embed_recover_vs_oracle_sizes(dev, N):
    cfg = synth.make_config("lorenz")
    m = make_lorenz(cfg, 31, "lorenz", dev)
    sd = O.to_torch(synth.lorenz_embedding_weights(cfg, 31))
    x = torch.from_numpy(synth.normal_embeds((N, 3), 8, scale=9.0))
    with torch.no_grad():
        g = m.embed(x.to(dev))
        g_ref = O.lorenz_embed(sd, cfg, x)
        assert O.rel_l2(g.cpu(), g_ref) < TOL
        assert O.rel_l2(m.recover(g).cpu(), O.lorenz_recover(sd, cfg, g_ref)) < TOL
        assert O.rel_l2(m.koopmanOperation(g).cpu(), O.lorenz_koopman(sd, cfg, g_ref)) < TOL
        assert m.embed(x[:0].to(dev)).shape == (0, 32)


def test_embed_small_and_large_batch_kernels_agree(dev):
    """`embed` switches from the warp-per-8-samples kernel to the thread-per-sample kernel at N = SMs x 256: the same rows
    through both must agree (same contraction order; only the LayerNorm reduction differs), and the boundary sizes
    (N not a multiple of 8, N = threshold, threshold + 1) must be exact against the oracle."""
    cfg = synth.make_config("lorenz")
    m = make_lorenz(cfg, 31, "lorenz", dev)
    sd = O.to_torch(synth.lorenz_embedding_weights(cfg, 31))
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    thr = sms * 256
    x = torch.from_numpy(synth.normal_embeds((thr + 1, 3), 9, scale=9.0)).to(dev)
    with torch.no_grad():
        big = m.embed(x)                         # thread-per-sample kernel
        small = m.embed(x[:thr])                 # warp-per-8-samples kernel
        assert O.rel_l2(small.cpu(), big[:thr].cpu()) < 1e-6
        for n in (1, 7, 8, 9, 1001):
            assert O.rel_l2(m.embed(x[:n]).cpu(), O.lorenz_embed(sd, cfg, x[:n].cpu())) < TOL
        ref = O.lorenz_embed(sd, cfg, x[-4096:].cpu())
        assert O.rel_l2(big[-4096:].cpu(), ref) < TOL


def test_mu_std_assignment_is_picked_up(dev):
    """Scripts overwrite model.mu / model.std by attribute assignment (examples/lorenz/train_lorenz_enn.py:55-57)."""
    cfg = synth.make_config("lorenz")
    m = make_lorenz(cfg, 32, "lorenz", dev)
    sd = O.to_torch(synth.lorenz_embedding_weights(cfg, 32))
    x = torch.from_numpy(synth.normal_embeds((64, 3), 3, scale=5.0))
    with torch.no_grad():
        m.embed(x.to(dev))
        m.mu = torch.tensor([1.0, 2.0, 3.0]).to(dev)
        m.std = torch.tensor([4.0, 5.0, 6.0]).to(dev)
        sd["mu"], sd["std"] = m.mu.cpu(), m.std.cpu()
        assert O.rel_l2(m.embed(x.to(dev)).cpu(), O.lorenz_embed(sd, cfg, x)) < TOL
        m.observableNet[0].weight.mul_(1.5)              # in-place parameter update -> repack
        sd["observableNet.0.weight"] = m.observableNet[0].weight.cpu()
        assert O.rel_l2(m.embed(x.to(dev)).cpu(), O.lorenz_embed(sd, cfg, x)) < TOL


@pytest.mark.parametrize("name", list(CYL_CASES))
def test_cylinder_embedding_matches_reference(golden_dir, dev, name):
    c = CYL_CASES[name]
    gold = np.load(os.path.join(golden_dir, f"cyl_{name}.npz"))
    cfg = synth.make_config("cylinder")
    m = make_cylinder(cfg, c["seed"], dev)
    x, visc = synth.cylinder_fields(c["N"], seed=c["seed"])
    x, visc = torch.from_numpy(x).to(dev), torch.from_numpy(visc).to(dev)
    with torch.no_grad():
        g = m.embed(x, visc)
        assert O.rel_l2(g.cpu(), gold["embed"]) < TOL
        rec = m.recover(g)
        assert O.rel_l2(rec.cpu(), gold["recover"]) < TOL
        assert torch.equal(rec.cpu() == 0, torch.from_numpy(gold["recover"]) == 0) or \
            int((rec[0, 0] == 0).sum()) >= 196
        g2, xhat = m(x, visc)
        assert O.rel_l2(xhat[:, 0].cpu(), gold["fwd_xhat_ch0"]) < TOL          # forward() does not mask
        assert O.rel_l2(m.koopmanOperation(g, visc).cpu(), gold["koopman"]) < TOL
        assert O.rel_l2(m.koopmanDiag.cpu(), gold["kdiag"]) < TOL
        assert O.rel_l2(m.koopmanOperator[0].cpu(), gold["kmatrix0"]) < TOL


@pytest.mark.parametrize("conv_mode", [0, 1, 2, 3])  # 0: 3xTF32 mma.sync (in-kernel build), 1: FP32 SIMT, 2: pipelined 3xTF32 mma.sync, 3: tcgen05 (default)
@pytest.mark.parametrize("N,chunk", [(1, 64), (5, 2), (70, 64)])
def test_cylinder_vs_oracle_chunking(dev, N, chunk, conv_mode):
    cfg = synth.make_config("cylinder")
    m = make_cylinder(cfg, 55, dev)
    m.set_conv_mode(conv_mode)
    sd = O.to_torch(synth.cylinder_embedding_weights(cfg, 55))
    x, visc = synth.cylinder_fields(N, seed=3)
    x, visc = torch.from_numpy(x), torch.from_numpy(visc)
    with torch.no_grad():
        g = m.embed(x.to(dev), visc.to(dev))
        m.set_chunk(chunk)
        g_ref = O.cylinder_embed(sd, cfg, x, visc)
        assert O.rel_l2(g.cpu(), g_ref) < TOL
        assert O.rel_l2(m.embed(x.to(dev), visc.to(dev)).cpu(), g_ref) < TOL
        rec_err = O.rel_l2(m.recover(g).cpu(), O.cylinder_recover(sd, cfg, g_ref))
        print(f"cylinder recover conv_mode={conv_mode} N={N}: rel-L2 {rec_err:.2e}")
        assert rec_err < TOL
        assert O.rel_l2(m.koopmanOperation(g, visc.to(dev)).cpu(), O.cylinder_koopman(sd, cfg, g_ref, visc)) < TOL


@pytest.mark.parametrize("name", list(TRAIN_CASES))
def test_koopman_train_step_matches_reference(golden_dir, dev, name):
    """loss, loss_reconstruct, flat gradient, clip norm and post-Adam parameters vs. the reference step body
    (embedding_lorenz.py:181-221 + enn_trainer.py:93-102)."""
    from trphysx_b200.config import PhysConfig
    from trphysx_b200.embedding import LorenzEmbeddingTrainer, RosslerEmbeddingTrainer
    from trphysx_b200.optim import FusedClipAdam
    c = TRAIN_CASES[name]
    gold = np.load(os.path.join(golden_dir, f"train_{name}.npz"))
    cfg_ns = synth.make_config(c["kind"])
    head = (LorenzEmbeddingTrainer if c["kind"] == "lorenz" else RosslerEmbeddingTrainer)(PhysConfig(**vars(cfg_ns)))
    head.embedding_model.load_state_dict(torch_sd(synth.lorenz_embedding_weights(cfg_ns, c["seed"])))
    head.to(dev)
    gen = synth.lorenz_trajectories if c["kind"] == "lorenz" else synth.rossler_trajectories
    states = torch.from_numpy(gen(c["B"], c["T"] - 1, seed=c["seed"])).to(dev)
    opt = FusedClipAdam(head.parameters(), lr=1e-3, weight_decay=1e-8, max_norm=0.1)
    for it in range(c["iters"]):
        loss, loss_rec = head(states)
        loss = loss.sum()
        loss.backward()
        flat = torch.cat([p.grad.reshape(-1) for p in head.parameters()])
        assert abs(float(loss) - float(gold[f"loss{it}"])) <= 2e-5 * abs(float(gold[f"loss{it}"]))
        assert abs(float(loss_rec) - float(gold[f"loss_rec{it}"])) <= 2e-5 * abs(float(gold[f"loss_rec{it}"]))
        assert O.rel_l2(flat.cpu(), gold[f"grad{it}"]) < TOL
        norm = opt.step()
        opt.zero_grad()
        assert abs(float(norm) - float(gold[f"norm{it}"])) <= 2e-5 * float(gold[f"norm{it}"])
    after = torch.cat([p.detach().reshape(-1) for p in head.parameters()])
    assert O.rel_l2(after.cpu(), gold["params_after"]) < 1e-6


def test_train_step_config3_vs_oracle(dev):
    """BASELINE config 3 shape: states [512, 16, 3]."""
    from trphysx_b200.config import PhysConfig
    from trphysx_b200.embedding import LorenzEmbeddingTrainer
    cfg_ns = synth.make_config("lorenz")
    head = LorenzEmbeddingTrainer(PhysConfig(**vars(cfg_ns)))
    head.embedding_model.load_state_dict(torch_sd(synth.lorenz_embedding_weights(cfg_ns, 77)))
    head.to(dev)
    states = torch.from_numpy(synth.lorenz_trajectories(512, 15, seed=77))
    loss, loss_rec = head(states.to(dev))
    loss.backward()
    flat = torch.cat([p.grad.reshape(-1) for p in head.parameters()])
    sd = O.to_torch(synth.lorenz_embedding_weights(cfg_ns, 77))
    l_ref, lr_ref, g_ref = O.lorenz_train_grads(sd, cfg_ns, states)
    assert abs(float(loss) - float(l_ref)) <= 2e-5 * abs(float(l_ref))
    assert abs(float(loss_rec) - float(lr_ref)) <= 2e-5 * abs(float(lr_ref))
    assert O.rel_l2(flat.cpu(), g_ref) < TOL


def test_train_step_tensor_core_decoder_opt_in(dev, monkeypatch):
    """TRPX_KOOPMAN_TC=1: the decoder forward+backward of the Koopman training step as 3xTF32 tile products.  Faster
    (0.19 vs 0.25 ms at config 3) but this gradient is ill conditioned (1e4 loss weights, ~50x cancellation in the row sums):
    the 2^-22 per-product error lands it at ~1.1e-5 from the oracle, over the 1e-5 bar, which is why it is opt-in.  Pinned
    here at 3e-5 (loss at 2e-5) on the config-3 shape and on a ragged small batch (partial tiles, idle CTAs)."""
    from trphysx_b200.config import PhysConfig
    from trphysx_b200.embedding import LorenzEmbeddingTrainer
    monkeypatch.setenv("TRPX_KOOPMAN_TC", "1")
    cfg_ns = synth.make_config("lorenz")
    for B, T in ((512, 16), (21, 5)):
        head = LorenzEmbeddingTrainer(PhysConfig(**vars(cfg_ns)))
        head.embedding_model.load_state_dict(torch_sd(synth.lorenz_embedding_weights(cfg_ns, 77)))
        head.to(dev)
        states = torch.from_numpy(synth.lorenz_trajectories(B, T - 1, seed=77))
        loss, loss_rec = head(states.to(dev))
        loss.backward()
        flat = torch.cat([p.grad.reshape(-1) for p in head.parameters()])
        sd = O.to_torch(synth.lorenz_embedding_weights(cfg_ns, 77))
        l_ref, lr_ref, g_ref = O.lorenz_train_grads(sd, cfg_ns, states)
        assert abs(float(loss.detach()) - float(l_ref)) <= 2e-5 * abs(float(l_ref))
        assert abs(float(loss_rec) - float(lr_ref)) <= 2e-5 * abs(float(lr_ref))
        assert O.rel_l2(flat.cpu(), g_ref) < 3e-5


def test_fused_koopman_step_equals_autograd_step(dev):
    """FusedKoopmanStep (no autograd graph) gives the same parameters as head(states); backward(); opt.step()."""
    from trphysx_b200.config import PhysConfig
    from trphysx_b200.embedding import LorenzEmbeddingTrainer
    from trphysx_b200.optim import FusedClipAdam, FusedKoopmanStep
    cfg_ns = synth.make_config("lorenz")
    states = torch.from_numpy(synth.lorenz_trajectories(64, 7, seed=9)).to(dev)
    outs = []
    for fused in (False, True):
        head = LorenzEmbeddingTrainer(PhysConfig(**vars(cfg_ns)))
        head.embedding_model.load_state_dict(torch_sd(synth.lorenz_embedding_weights(cfg_ns, 21)))
        head.to(dev)
        opt = FusedClipAdam(head.parameters(), lr=1e-3, weight_decay=1e-8, max_norm=0.1)
        stepper = FusedKoopmanStep(head, opt)
        for _ in range(3):
            if fused:
                loss = stepper(states)[0]
            else:
                loss, _ = head(states)
                loss.backward()
                opt.step()
                opt.zero_grad()
        outs.append((float(loss), torch.cat([p.detach().reshape(-1) for p in head.parameters()]).clone()))
    assert abs(outs[0][0] - outs[1][0]) <= 1e-6 * abs(outs[0][0])
    assert torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("kind", ["lorenz", "rossler"])
def test_eval_states_fused_recover_mse(dev, kind):
    """Trainer.eval_states (utils/trainer.py:380-390): MSELoss()(recover(pred).view(B,T,3), states), here in one
    launch without writing the states; against the oracle's recover + torch MSE, ragged N, and the optional states."""
    from trphysx_b200.evaluation import eval_states, embed_series
    cfg = synth.make_config(kind)
    w = synth.lorenz_embedding_weights(cfg, 812)
    emb = make_lorenz(cfg, 812, kind, dev)
    B, T = 37, 29                                     # N = 1073: not a multiple of anything
    g = torch.from_numpy(synth.normal_embeds((B, T, cfg.n_embd), 813))
    states = torch.from_numpy(synth.normal_embeds((B, T, 3), 814)) * 3.0
    ref_states = O.lorenz_recover(O.to_torch(w), cfg, g.view(-1, cfg.n_embd)).view(B, T, 3)
    ref = torch.nn.functional.mse_loss(ref_states, states)
    err = eval_states(emb, g.to(dev), states.to(dev))
    assert err.dim() == 0 and abs(float(err) - float(ref)) <= 2e-6 * float(ref)
    err2, rec = eval_states(emb, g.to(dev), states.to(dev), return_states=True)
    assert float(err2) == float(err) and rec.shape == (B, T, 3)
    assert O.rel_l2(rec.cpu(), ref_states) < TOL
    # bulk embedding of a whole series in one call (dataset_lorenz.py:22-47)
    series = torch.from_numpy(synth.lorenz_trajectories(8, 40, seed=5) if kind == "lorenz"
                              else synth.rossler_trajectories(8, 40, seed=5))
    ge = embed_series(emb, series.to(dev))
    assert ge.shape == (8, 41, cfg.n_embd)
    assert O.rel_l2(ge.cpu().view(-1, cfg.n_embd), O.lorenz_embed(O.to_torch(w), cfg, series.view(-1, 3))) < TOL


def _cyl_head(cfg, seed, dev):
    from trphysx_b200.embedding.embedding_cylinder import CylinderEmbeddingTrainer
    head = CylinderEmbeddingTrainer(cfg)
    head.embedding_model.load_state_dict(torch_sd(synth.cylinder_embedding_weights(cfg, seed)))
    return head.to(dev)


@pytest.mark.parametrize("name", list(CYL_TRAIN_CASES))
def test_cylinder_training_step_matches_reference(golden_dir, dev, name):
    """CylinderEmbeddingTrainer.forward + loss.backward() (embedding_cylinder.py:236-281, enn_trainer.py:93-96) against
    the reference's recorded loss, loss_reconstruct, flat gradient (all 262,535 entries, parameters() order) and
    clip_grad_norm_ value; the loss VALUE under no_grad (inference kernels) must agree with the fused step."""
    c = CYL_TRAIN_CASES[name]
    gold = np.load(os.path.join(golden_dir, f"cyltrain_{name}.npz"))
    cfg = synth.make_config("cylinder")
    head = _cyl_head(cfg, c["seed"], dev)
    B, T = c["B"], c["T"]
    x, visc = synth.cylinder_fields(B * T, seed=c["seed"])
    states = torch.from_numpy(x).view(B, T, 3, 64, 128).to(dev)
    visc = torch.from_numpy(visc).view(B, T, 1)[:, 0].to(dev)
    with torch.no_grad():
        loss_v, loss_rec_v = head(states, visc)
    assert abs(float(loss_v) - float(gold["loss"])) <= 1e-5 * float(gold["loss"])
    assert abs(float(loss_rec_v) - float(gold["loss_rec"])) <= 1e-5 * float(gold["loss_rec"])
    loss, loss_rec = head(states, visc)
    assert not loss_rec.requires_grad
    loss.backward()
    assert abs(float(loss) - float(gold["loss"])) <= 1e-5 * float(gold["loss"])
    assert abs(float(loss_rec) - float(gold["loss_rec"])) <= 1e-5 * float(gold["loss_rec"])
    flat = torch.cat([p.grad.reshape(-1) for p in head.parameters()]).cpu()
    err = O.rel_l2(flat, gold["grad"])
    off, worst = 0, (0.0, "")
    for k, p in head.embedding_model.named_parameters():       # every tensor on its own: small ones must not hide
        n = p.numel()
        e = O.rel_l2(flat[off:off + n], gold["grad"][off:off + n])
        worst = max(worst, (e, k))
        off += n
    print(f"cylinder training step: loss {float(loss):.5f}, flat gradient rel-L2 {err:.2e}, worst tensor {worst[1]} {worst[0]:.2e}")
    assert err < TOL and worst[0] < 3 * TOL, worst
    norm = torch.nn.utils.clip_grad_norm_(head.parameters(), 0.1)
    assert abs(float(norm) - float(gold["norm"])) <= 1e-5 * float(gold["norm"])


@pytest.mark.parametrize("B,T", [(3, 2), (5, 2), (2, 1)])
def test_cylinder_training_step_vs_oracle_and_descent(dev, B, T):
    """Sizes without a fixture against the oracle's autograd, determinism (two calls give the same bits), and the loss goes
    down under plain SGD through the same handle.  The yardstick is the oracle in FLOAT64: for some shapes (B=3, T=2) the
    FP32 CPU backward of the reference is itself 1e-4..3e-4 away from it per tensor (oneDNN picks a less accurate
    algorithm), so each tensor is allowed max(1e-5, 4x the FP32 reference's own deviation), as in the transformer
    training-step test."""
    cfg = synth.make_config("cylinder")
    head = _cyl_head(cfg, 612, dev)
    w_np = synth.cylinder_embedding_weights(cfg, 612)
    sd = O.to_torch(w_np)
    x, visc = synth.cylinder_fields(B * T, seed=77)
    states_c = torch.from_numpy(x).view(B, T, 3, 64, 128)
    visc_c = torch.from_numpy(visc).view(B, T, 1)[:, 0]
    ref_loss, ref_rec, ref_grad32 = O.cylinder_train_grads(sd, cfg, states_c, visc_c)
    _, _, ref_grad = O.cylinder_train_grads(O.to_torch(w_np, torch.float64), cfg, states_c.double(), visc_c.double())
    states, visc = states_c.to(dev), visc_c.to(dev)
    losses, first = [], None
    for it in range(3):
        head.zero_grad(set_to_none=True)
        loss, rec = head(states, visc)
        loss.backward()
        flat = torch.cat([p.grad.reshape(-1) for p in head.parameters()])
        if it == 0:
            assert abs(float(loss) - float(ref_loss)) <= 1e-5 * float(ref_loss)
            assert abs(float(rec) - float(ref_rec)) <= 1e-5 * float(ref_rec)
            off, fc = 0, flat.cpu().double()
            for k, p in head.embedding_model.named_parameters():
                n = p.numel()
                if float(ref_grad[off:off + n].norm()) > 0:
                    noise = O.rel_l2(ref_grad32[off:off + n].double(), ref_grad[off:off + n])
                    e = O.rel_l2(fc[off:off + n], ref_grad[off:off + n])
                    assert e < max(TOL, 4 * noise), (k, e, noise)
                else:
                    assert float(fc[off:off + n].abs().max()) == 0.0, k
                off += n
            head.zero_grad(set_to_none=True)
            loss2, _ = head(states, visc)
            loss2.backward()
            assert torch.equal(flat, torch.cat([p.grad.reshape(-1) for p in head.parameters()])) and float(loss2) == float(loss)
        losses.append(float(loss))
        with torch.no_grad():
            for p in head.parameters():
                p.add_(p.grad, alpha=-2e-3)
    assert losses[2] < losses[1] < losses[0], losses


@pytest.mark.parametrize("kind,n", [("lorenz", 16384), ("rossler", 50001)])
def test_recover_tensor_core_kernel(dev, kind, n):
    """recover for N >= 16384 runs on the 3xTF32 tensor-core kernel: against the oracle at the FP32 bar, against the
    FP32 SIMT kernel (mode 1), ragged last tile."""
    cfg = synth.make_config(kind)
    w = synth.lorenz_embedding_weights(cfg, 911)
    emb = make_lorenz(cfg, 911, kind, dev)
    g = torch.from_numpy(synth.normal_embeds((n, cfg.n_embd), 912)) * 1.5
    ref = O.lorenz_recover(O.to_torch(w), cfg, g)
    tc = emb.recover(g.to(dev))
    emb.set_recover_mode(1)
    simt = emb.recover(g.to(dev))
    emb.set_recover_mode(0)
    assert not torch.equal(tc, simt)                      # really two different kernels
    e_tc, e_simt = O.rel_l2(tc.cpu(), ref), O.rel_l2(simt.cpu(), ref)
    print(f"recover {kind} N={n}: rel-L2 tensor-core {e_tc:.2e}, SIMT {e_simt:.2e}")
    assert e_tc < TOL and e_simt < TOL
    assert float((tc - simt).abs().max()) < 1e-4 * float(ref.abs().max())
    # the fused recover + MSE takes the same kernel at this size
    target = ref + 0.1 * torch.from_numpy(synth.normal_embeds((n, 3), 913))
    want = torch.nn.functional.mse_loss(ref, target)
    got = emb.recover_mse(g.to(dev), target.to(dev))
    assert abs(float(got) - float(want)) <= 5e-6 * float(want)


@pytest.mark.parametrize("kind,n", [("rossler", 75776), ("lorenz", 200001), ("rossler", 75775)])
def test_recover_tcgen05_kernel(dev, kind, n):
    """recover for N >= 4 x 128 x (SM count) samples runs on the tcgen05 kernel (csrc/mlpemb_umma.cu: layer 1 as
    error-compensated kind::tf32 MMAs with accumulators in tensor memory, layer 2 in FP32 in the epilogue): against the
    oracle at the FP32 bar and against the other two kernels; exact tile multiple, ragged last tile, and one sample below
    the switch (which must take the mma.sync kernel)."""
    cfg = synth.make_config(kind)
    w = synth.lorenz_embedding_weights(cfg, 921)
    emb = make_lorenz(cfg, 921, kind, dev)
    g = torch.from_numpy(synth.normal_embeds((n, cfg.n_embd), 922)) * 1.5
    ref = O.lorenz_recover(O.to_torch(w), cfg, g)
    outs = {}
    for mode in (0, 2, 1):
        emb.set_recover_mode(mode)
        outs[mode] = emb.recover(g.to(dev))
    emb.set_recover_mode(0)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    if n >= 4 * 128 * sms:
        assert not torch.equal(outs[0], outs[2])          # really the tcgen05 kernel
    else:
        assert torch.equal(outs[0], outs[2])
    for mode in (0, 2, 1):
        e = O.rel_l2(outs[mode].cpu(), ref)
        print(f"recover {kind} N={n} mode {mode}: rel-L2 {e:.2e}")
        assert e < TOL
    assert float((outs[0] - outs[1]).abs().max()) < 1e-4 * float(ref.abs().max())
    # rows of the last (partial) tile and the first rows of the next tile of another CTA
    assert O.rel_l2(outs[0][-200:].cpu(), ref[-200:]) < TOL and O.rel_l2(outs[0][:300].cpu(), ref[:300]) < TOL
    # the fused recover + MSE (Trainer.eval_states) takes the same kernel at this size: value, and the states it can return
    target = ref + 0.1 * torch.from_numpy(synth.normal_embeds((n, 3), 923))
    want = torch.nn.functional.mse_loss(ref, target)
    got = emb.recover_mse(g.to(dev), target.to(dev))
    assert abs(float(got) - float(want)) <= 5e-6 * float(want)
    got2, states = emb.recover_mse(g.to(dev), target.to(dev), return_states=True)
    assert abs(float(got2) - float(want)) <= 5e-6 * float(want) and torch.equal(states, outs[0])


def test_rollout_to_host_pipeline(dev):
    """evaluation.rollout_to_host: pinned x0 -> embed -> generate -> chunked recover with overlapped D2H equals the
    three public calls done one after the other."""
    from trphysx_b200.evaluation import rollout_to_host
    from tests.util import make_gpt2
    cfg = synth.make_config("rossler")
    emb = make_lorenz(cfg, 921, "rossler", dev)
    tr = make_gpt2(cfg, 922, False, dev)
    B, S = 1003, 20
    x0 = torch.from_numpy(synth.rossler_trajectories(B, 0, seed=923)[:, 0].copy()).pin_memory()
    host = torch.empty((B, S + 1, 3), dtype=torch.float32).pin_memory()
    rollout_to_host(emb, tr, x0, S + 1, host, device=dev, chunks=4)
    torch.cuda.synchronize()
    g0 = emb.embed(x0.to(dev))
    out = tr.generate(g0.unsqueeze(1), max_length=S + 1, use_cache=True)[0]
    ref = emb.recover(out.view(B * (S + 1), cfg.n_embd)).view(B, S + 1, 3)
    # chunks of 5,271 samples take the FP32 recover kernel, the single call (21,063 samples) the tensor-core one
    assert O.rel_l2(host, ref.cpu()) < 1e-6
    # large batch: generate() itself runs slice by slice (one round of resident trajectories each) under the copies
    B2, S2 = 2 * tr.rollout_round_trajectories(dev) + 1234, 6
    x1 = torch.from_numpy(synth.rossler_trajectories(B2, 0, seed=924)[:, 0].copy()).pin_memory()
    host2 = torch.empty((B2, S2 + 1, 3), dtype=torch.float32).pin_memory()
    rollout_to_host(emb, tr, x1, S2 + 1, host2, device=dev, chunks=4)
    torch.cuda.synchronize()
    assert tr.last_launch()["grid"] * 64 >= (B2 + 2) // 3          # the last generate() call saw a third of the batch
    host3 = torch.empty_like(host2).pin_memory()
    rollout_to_host(emb, tr, x1, S2 + 1, host3, device=dev, chunks=4, pipeline=False)
    torch.cuda.synchronize()
    assert O.rel_l2(host2, host3) < 1e-6
    out1 = tr.generate(emb.embed(x1.to(dev)).unsqueeze(1), max_length=S2 + 1, use_cache=True)[0]
    ref1 = emb.recover(out1.view(B2 * (S2 + 1), cfg.n_embd)).view(B2, S2 + 1, 3)
    assert O.rel_l2(host2, ref1.cpu()) < 1e-6


@pytest.mark.parametrize("B,S,auto_variant", [(6, 400, 7), (64, 400, 7), (160, 200, 1)])
def test_cylinder_end_to_end_config4(dev, B, S, auto_variant):
    """BASELINE config 4 end to end through the public modules: CylinderEmbedding.embed(x0, visc) -> PhysformerGPT2.
    generate(max_length = 401, use_cache=True) -> CylinderEmbedding.recover, against the oracle's embed / generate /
    recover.  B = 6 and B = 64 take the cluster rollout kernel (one and two waves of single-trajectory clusters), B = 160
    rollout_generic_kernel<G>1> (automatic selection).  The
    embedded rollout is compared over the whole horizon; the recovered fields for every trajectory at a spread of time
    steps (the oracle's decoder costs 0.23 GFLOP per frame on the CPU)."""
    from tests.util import make_gpt2, stable_prefix
    cfg = synth.make_config("cylinder")
    emb = make_cylinder(cfg, 931, dev)
    w_t = synth.gpt2_weights(cfg, 932, stress=True, stress_sigma=0.05)
    tr = make_gpt2(cfg, 932, True, dev, sigma=0.05)
    sd_e = O.to_torch(synth.cylinder_embedding_weights(cfg, 931))
    x, visc = synth.cylinder_fields(B, seed=933)
    x, visc = torch.from_numpy(x), torch.from_numpy(visc)
    with torch.no_grad():
        g0 = emb.embed(x.to(dev), visc.to(dev))
        out = tr.generate(g0.unsqueeze(1), max_length=S + 1, use_cache=True, return_presents=False)[0]
        info = tr.last_launch()
        assert info["variant"] == auto_variant and (auto_variant == 7 or info["group"] > 1), info
        fields = emb.recover(out.reshape(-1, cfg.n_embd)).view(B, S + 1, 3, 64, 128)
    g0_ref = O.cylinder_embed(sd_e, cfg, x, visc)
    assert O.rel_l2(g0.cpu(), g0_ref) < TOL
    n, ref = stable_prefix(cfg, w_t, g0_ref.unsqueeze(1), S + 1, True)
    assert n == S + 1, n
    got = out.cpu()
    err = O.rel_l2(got, ref)
    per_step = (got - ref).norm(dim=2).max(0).values / ref.norm(dim=2).max(0).values
    print(f"cylinder end-to-end B={B}: launch {info}, embedded rollout rel-L2 {err:.2e}, worst step {float(per_step.max()):.2e}")
    assert err < TOL and float(per_step.max()) < 3 * TOL
    steps = [0, 1, 15, 16, 17, 100, 250, S] if B <= 8 else [0, 17, S]
    ref_fields = O.cylinder_recover(sd_e, cfg, ref[:, steps].reshape(-1, cfg.n_embd)).view(B, len(steps), 3, 64, 128)
    ferr = O.rel_l2(fields[:, steps].cpu(), ref_fields)
    print(f"cylinder end-to-end B={B}: recovered fields rel-L2 {ferr:.2e} over {B * len(steps)} frames")
    assert ferr < TOL
    assert int((fields[:, steps][:, :, 0] == 0).sum()) >= 196 * B * len(steps)      # cylinder cells are masked


def test_feature_cache_blocks_and_round_trip(dev, tmp_path):
    """data_utils.build_feature_cache (PhysicalDataset's constructor logic, dataset_phys.py:53-86 + LorenzDataset /
    CylinderDataset.embed_data): blocks of the bulk-embedded series equal the oracle's embedding of the same windows,
    eval=True keeps the states, ndata stops early, the second call is served from the pickle."""
    from trphysx_b200.data_utils import build_feature_cache, cached_features_file
    cfg = synth.make_config("lorenz")
    emb = make_lorenz(cfg, 812, "lorenz", dev)
    sd = O.to_torch(synth.lorenz_embedding_weights(cfg, 812))
    series = {str(k): synth.lorenz_trajectories(1, 40, seed=60 + k)[0] for k in range(3)}      # three [41, 3] series
    data = tmp_path / "lorenz.hdf5"
    data.write_bytes(b"")
    ex, st = build_feature_cache(emb, str(data), series, block_size=16, stride=8, ndata=2, eval=True)
    assert len(ex) == 2 * 4 and len(st) == len(ex)                    # (41 - 16)//8 + 1 = 4 blocks per series, 2 series
    assert ex[0].device.type == "cpu" and ex[0].shape == (16, 32) and st[0].shape == (16, 3)
    s1 = torch.from_numpy(series["1"])
    assert torch.equal(st[5], s1[8:24])
    assert O.rel_l2(ex[5], O.lorenz_embed(sd, cfg, s1[8:24])) < TOL
    assert os.path.isfile(cached_features_file(str(data), emb, 16, ndata=2))
    ex2, st2 = build_feature_cache(emb, str(data), None, block_size=16, stride=8, ndata=2, eval=True)   # from the cache
    assert all(torch.equal(a, b) for a, b in zip(ex, ex2)) and all(torch.equal(a, b) for a, b in zip(st, st2))
    # Cylinder series carry their viscosity (2 / Re, dataset_cylinder.py:36)
    ccfg = synth.make_config("cylinder")
    cemb = make_cylinder(ccfg, 55, dev)
    csd = O.to_torch(synth.cylinder_embedding_weights(ccfg, 55))
    x, _ = synth.cylinder_fields(6, seed=9)
    cex, _ = build_feature_cache(cemb, str(tmp_path / "cyl.hdf5"), {"400": x}, block_size=4, stride=2,
                                 visc_of_key=lambda k: 2.0 / float(k))
    assert len(cex) == 2 and cex[0].shape == (4, 128)
    ref = O.cylinder_embed(csd, ccfg, torch.from_numpy(x), torch.full((6, 1), 2.0 / 400.0))
    assert O.rel_l2(torch.cat([cex[0], cex[1][2:]]), ref) < TOL


def test_fused_cylinder_step_equals_reference_loop(dev):
    """FusedKoopmanStep on the Cylinder head (fused fwd+bwd -> clip + Adam on the flat master buffer, no autograd graph)
    against the reference's step body run by torch on the oracle graph (enn_trainer.py:93-102: loss.backward();
    clip_grad_norm_(0.1); Adam(lr, weight_decay).step()), two steps; afterwards embed() sees the updated weights."""
    from trphysx_b200.optim import FusedClipAdam, FusedKoopmanStep
    cfg = synth.make_config("cylinder")
    head = _cyl_head(cfg, 731, dev)
    w_np = synth.cylinder_embedding_weights(cfg, 731)
    sd = {k: (v.clone().requires_grad_(True) if k in O.CYL_PARAM_ORDER else v) for k, v in O.to_torch(w_np).items()}
    leaves = [sd[k] for k in O.CYL_PARAM_ORDER]
    ref_opt = torch.optim.Adam(leaves, lr=1e-3, weight_decay=1e-8)
    B, T = 2, 3
    x, visc = synth.cylinder_fields(B * T, seed=732)
    states_c = torch.from_numpy(x).view(B, T, 3, 64, 128)
    visc_c = torch.from_numpy(visc).view(B, T, 1)[:, 0]
    ref_losses = []
    for _ in range(2):
        ref_opt.zero_grad()
        loss, _ = O.cylinder_train_loss(sd, cfg, states_c, visc_c)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(leaves, 0.1)
        ref_opt.step()
        ref_losses.append(float(loss))
    opt = FusedClipAdam(head.parameters(), lr=1e-3, weight_decay=1e-8, max_norm=0.1)
    step = FusedKoopmanStep(head, opt)
    states, visc = states_c.to(dev), visc_c.to(dev)
    for i in range(2):
        l2 = step(states, visc)
        assert abs(float(l2[0]) - ref_losses[i]) <= 2e-5 * ref_losses[i], (i, float(l2[0]), ref_losses[i])
    # Adam divides by sqrt(v): entries whose gradient is at the FP32 noise level step by +-lr on that noise on both
    # sides, so the parameters agree to lr-sized absolute errors on a few entries; the UPDATES must agree as a whole
    got = dict(head.embedding_model.named_parameters())
    w0 = O.to_torch(w_np)
    for k in O.CYL_PARAM_ORDER:
        a, b = got[k].detach().cpu(), sd[k].detach()
        assert O.rel_l2(a, b) < 1e-4, k
        assert O.rel_l2(a - w0[k], b - w0[k]) < 5e-2, k
    m = head.embedding_model.eval()
    with torch.no_grad():
        g = m.embed(states[:, 0], visc)
    cur = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    assert O.rel_l2(g.cpu(), O.cylinder_embed(cur, cfg, states_c[:, 0], visc_c)) < TOL        # the kernels use the UPDATED weights


# ------------------------------------------------------------------------------------------------------------------
# Gray-Scott embedding: eval-mode network on the kernels of csrc/gs.cu
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(GS_CASES))
def test_grayscott_embedding_matches_reference(golden_dir, dev, name):
    c = GS_CASES[name]
    gold = np.load(os.path.join(golden_dir, f"gs_{name}.npz"))
    cfg = synth.make_config("grayscott")
    m = make_grayscott(cfg, c["seed"], dev)
    x = torch.from_numpy(synth.grayscott_fields(c["N"], seed=c["seed"])).to(dev)
    with torch.no_grad():
        g = m.embed(x)
        assert O.rel_l2(g.cpu(), gold["embed"]) < TOL
        rec = m.recover(g)
        assert O.rel_l2(rec.cpu(), gold["recover"]) < TOL
        fg, xhat, g0, out0, x2 = m(x)
        assert torch.equal(fg, g)
        assert O.rel_l2(g0.cpu(), gold["fwd_g0"]) < TOL
        assert O.rel_l2(out0.cpu(), gold["fwd_out0"]) < TOL
        assert O.rel_l2(xhat[:, 0].cpu(), gold["fwd_xhat_ch0"]) < TOL
        assert O.rel_l2(x2[:, 1].cpu(), gold["fwd_x2_ch1"]) < TOL
        assert O.rel_l2(m.koopmanOperation(g).cpu(), gold["koopman"]) < TOL
        K = m.koopmanOperator
        assert K.shape == (c["N"], cfg.n_embd, cfg.n_embd)
        assert np.array_equal(K[0].cpu().numpy(), gold["kmatrix0"]) and torch.equal(K[0], K[-1])


@pytest.mark.parametrize("N", [1, 19])
def test_grayscott_vs_oracle_batches(dev, N):
    """Ragged batch (more than one 16-frame chunk, a partial 8-sample Linear pass), a different seed, mu/std reassigned."""
    cfg = synth.make_config("grayscott")
    w = synth.grayscott_embedding_weights(cfg, 77)
    m = make_grayscott(cfg, 77, dev)
    m.mu = torch.tensor(0.45, device=dev)
    m.std = torch.tensor(0.6, device=dev)
    sd = O.to_torch(w)
    sd["mu"], sd["std"] = torch.tensor(0.45), torch.tensor(0.6)
    x = torch.from_numpy(synth.grayscott_fields(N, seed=5))
    with torch.no_grad():
        g = m.embed(x.to(dev))
        g_ref = O.grayscott_embed(sd, cfg, x)
        assert O.rel_l2(g.cpu(), g_ref) < TOL
        assert O.rel_l2(m.recover(g).cpu(), O.grayscott_recover(sd, cfg, g_ref)) < TOL
        assert O.rel_l2(m.koopmanOperation(g).cpu(), O.grayscott_koopman(sd, cfg, g_ref)) < TOL
    m.train()
    with pytest.raises(NotImplementedError):
        m.embed(x.to(dev))
