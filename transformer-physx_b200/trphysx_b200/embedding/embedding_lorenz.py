"""Drop-in `LorenzEmbedding` / `LorenzEmbeddingTrainer` (reference: trphysx/embedding/embedding_lorenz.py).

The module tree owns the parameters under the reference's state_dict names; embed / recover /
koopmanOperation / the training step run in libtrpx (fused MLP kernels, banded Koopman matvec, fused
forward+backward of the Koopman loss)."""
import ctypes
from typing import Tuple

import numpy as np
import torch
from torch import nn

from .. import _native as N
from .embedding_model import EmbeddingModel, EmbeddingTrainingHead

Tensor = torch.Tensor
HIDDEN = 500            # hard-coded in the reference (embedding_lorenz.py:38-39)


class LorenzEmbedding(EmbeddingModel):
    model_name = "embedding_lorenz"

    def __init__(self, config):
        super().__init__(config)
        S, E = config.state_dims[0], config.n_embd
        self.observableNet = nn.Sequential(
            nn.Linear(S, HIDDEN), nn.ReLU(), nn.Linear(HIDDEN, E),
            nn.LayerNorm(E, eps=config.layer_norm_epsilon), nn.Dropout(config.embd_pdrop))
        self.recoveryNet = nn.Sequential(nn.Linear(E, HIDDEN), nn.ReLU(), nn.Linear(HIDDEN, S))
        self.obsdim = E
        self.kMatrixDiag = nn.Parameter(torch.linspace(1, 0, E))
        # offsets +1 and +2 of the skew-symmetric part (embedding_lorenz.py:59-67)
        self.xidx = torch.LongTensor(np.concatenate([np.arange(0, E - i) for i in range(1, 3)]))
        self.yidx = torch.LongTensor(np.concatenate([np.arange(i, E) for i in range(1, 3)]))
        self.kMatrixUT = nn.Parameter(0.1 * torch.rand(self.xidx.size(0)))
        self.register_buffer("mu", torch.tensor([0.0, 0.0, 0.0]))
        self.register_buffer("std", torch.tensor([1.0, 1.0, 1.0]))
        self._handles = N.HandleCache("trpx_mlpemb_destroy")
        self.kMatrix = None

    # ------------------------------------------------------------------ native plumbing
    def _weight_list(self):
        o, r = self.observableNet, self.recoveryNet
        return [self.mu, self.std, o[0].weight, o[0].bias, o[2].weight, o[2].bias, o[3].weight, o[3].bias,
                r[0].weight, r[0].bias, r[2].weight, r[2].bias, self.kMatrixDiag, self.kMatrixUT]

    def _handle(self, device: torch.device):
        lib = N.lib()
        if self.training and self.config.embd_pdrop:
            raise NotImplementedError("dropout > 0 in train mode is not implemented in the kernels")
        idx = device.index if device.index is not None else torch.cuda.current_device()
        entry = self._handles.get(idx)
        if entry is None:
            cfg = N.MlpEmbConfig(self.config.state_dims[0], HIDDEN, self.config.n_embd, self.config.layer_norm_epsilon)
            h = ctypes.c_void_p()
            N.check(lib.trpx_mlpemb_create(ctypes.byref(cfg), idx, ctypes.byref(h)))
            entry = (h, N.WeightTracker())
            self._handles[idx] = entry
            N.check(lib.trpx_mlpemb_set_recover_mode(h, getattr(self, "_recover_mode", 0)))
        h, tracker = entry
        ws = self._weight_list()
        if tracker.changed(ws):
            tensors = []
            for w in ws:
                if w.device != device:
                    raise NotImplementedError(f"parameter on {w.device} but input on {device}: move the model first")
                tensors.append(N.require_cuda_f32(w.detach(), "parameter"))
            N.check(lib.trpx_mlpemb_load_weights(h, N.ptr_array(tensors), len(tensors), N.stream_ptr(device)))
            self._keepalive = tensors
        return h


    def invalidate_weights(self):
        """Force the packed native copy of the weights to be rebuilt on the next call.  Needed only after writes that
        torch's version counters cannot see (`p.data.mul_()`, EMA through `.data`, a custom kernel): parameter updates
        through autograd-visible in-place ops, optimizers, `load_state_dict`, `.to()` and the fused optimizers of
        `trphysx_b200.optim` are detected automatically."""
        self._handles.invalidate()

    # ------------------------------------------------------------------ reference API
    def embed(self, x: Tensor) -> Tensor:
        """[B, 3] -> [B, n_embd]  (embedding_lorenz.py:94-105)"""
        self._inference_only("embed")
        x = N.require_cuda_f32(x, "x")
        assert x.dim() == 2 and x.size(1) == self.config.state_dims[0]
        g = torch.empty((x.size(0), self.obsdim), device=x.device, dtype=torch.float32)
        N.check(N.lib().trpx_mlpemb_embed(self._handle(x.device), x.data_ptr(), g.data_ptr(), x.size(0),
                                          N.stream_ptr(x.device)))
        return g

    def recover(self, g: Tensor) -> Tensor:
        """[B, n_embd] -> [B, 3]  (embedding_lorenz.py:107-118)"""
        self._inference_only("recover")
        g = N.require_cuda_f32(g, "g")
        assert g.dim() == 2 and g.size(1) == self.obsdim
        x = torch.empty((g.size(0), self.config.state_dims[0]), device=g.device, dtype=torch.float32)
        N.check(N.lib().trpx_mlpemb_recover(self._handle(g.device), g.data_ptr(), x.data_ptr(), g.size(0),
                                            N.stream_ptr(g.device)))
        return x

    def set_recover_mode(self, mode: int = 0):
        """0: automatic (tcgen05 kernel for N >= 4 x 128 x SM count samples, 3xTF32 mma.sync kernel for N >= 16384),
        1: FP32 SIMT kernel only, 2: never the tcgen05 kernel (A/B)."""
        self._recover_mode = mode
        for h, _ in self._handles.values():
            N.check(N.lib().trpx_mlpemb_set_recover_mode(h, mode))

    def recover_mse(self, g: Tensor, states: Tensor, return_states: bool = False):
        """MSELoss()(recover(g), states) in one launch without materialising the recovered states — the computation
        of `Trainer.eval_states` (trphysx/utils/trainer.py:380-390).  g: [N, n_embd]; states: [N, 3] (any leading
        shape with N*3 values).  Returns the MSE (device scalar) [and the recovered states [N, 3]]."""
        self._inference_only("recover_mse")
        g = N.require_cuda_f32(g, "g")
        states = N.require_cuda_f32(states, "states")
        assert g.dim() == 2 and g.size(1) == self.obsdim and states.numel() == g.size(0) * self.config.state_dims[0]
        mse = torch.empty((), device=g.device, dtype=torch.float32)
        out = torch.empty((g.size(0), self.config.state_dims[0]), device=g.device) if return_states else None
        N.check(N.lib().trpx_mlpemb_recover_mse(self._handle(g.device), g.data_ptr(), states.data_ptr(), g.size(0),
                                                out.data_ptr() if return_states else None, mse.data_ptr(),
                                                N.stream_ptr(g.device)))
        return (mse, out) if return_states else mse

    def forward(self, x: Tensor) -> Tuple[Tensor, Tensor]:
        """(g, xhat) = (embed(x), recover(embed(x)))  (embedding_lorenz.py:74-92)"""
        g = self.embed(x)
        return g, self.recover(g)

    def koopmanOperation(self, g: Tensor) -> Tensor:
        """One Koopman step as a banded matvec; also refreshes `koopmanOperator` (embedding_lorenz.py:120-142)."""
        self._inference_only("koopmanOperation")
        g = N.require_cuda_f32(g, "g")
        assert g.dim() == 2 and g.size(1) == self.obsdim
        out = torch.empty_like(g)
        h = self._handle(g.device)
        N.check(N.lib().trpx_mlpemb_koopman(h, g.data_ptr(), out.data_ptr(), g.size(0), N.stream_ptr(g.device)))
        self.kMatrix = self._dense_operator(g.device)
        return out

    def _dense_operator(self, device) -> Tensor:
        K = torch.empty((self.obsdim, self.obsdim), device=device, dtype=torch.float32)
        N.check(N.lib().trpx_mlpemb_koopman_matrix(self._handle(device), K.data_ptr(), N.stream_ptr(device)))
        return K

    @property
    def koopmanOperator(self, requires_grad: bool = True) -> Tensor:
        """Dense [n_embd, n_embd] operator of the last koopmanOperation call (embedding_lorenz.py:144-157)."""
        return self.kMatrix

    @property
    def koopmanDiag(self):
        return self.kMatrixDiag


class _FusedKoopmanLoss(torch.autograd.Function):
    """loss, loss_reconstruct = f(states; params): forward AND backward run in one native call; backward()
    only hands the stored flat gradient (nn.Module.parameters() order) back to autograd."""

    @staticmethod
    def forward(ctx, model, states, w_recon, w_dyn, w_reg, *params):
        lib = N.lib()
        dev = states.device
        h = model._handle(dev)
        B, T = states.size(0), states.size(1)
        nparam = int(lib.trpx_mlpemb_num_params(h))
        loss2 = torch.empty(2, device=dev, dtype=torch.float32)
        flat = torch.empty(nparam, device=dev, dtype=torch.float32)
        N.check(lib.trpx_mlpemb_train_fwd_bwd(h, states.data_ptr(), B, T, w_recon, w_dyn, w_reg, 1.0,
                                              loss2.data_ptr(), flat.data_ptr(), N.stream_ptr(dev)))
        ctx.flat = flat
        ctx.shapes = [p.shape for p in params]
        loss, loss_rec = loss2[0].clone(), loss2[1].clone()
        ctx.mark_non_differentiable(loss_rec)       # the reference returns loss_reconstruct detached (:221)
        return loss, loss_rec

    @staticmethod
    def backward(ctx, gloss, _grec):
        grads, off = [], 0
        for shp in ctx.shapes:
            n = int(np.prod(shp))
            grads.append((ctx.flat[off:off + n] * gloss).view(shp))
            off += n
        return (None, None, None, None, None) + tuple(grads)


class LorenzEmbeddingTrainer(EmbeddingTrainingHead):
    """Training head: forward(states[B,T,3]) -> (loss, loss_reconstruct) (embedding_lorenz.py:169-221).
    `loss.backward()` fills the parameters' .grad from the fused native forward+backward."""
    W_RECON, W_DYN, W_REG = 1e4, 1.0, 1e-1

    def __init__(self, config):
        super().__init__()
        self.embedding_model = LorenzEmbedding(config)

    def forward(self, states: Tensor):
        self.embedding_model.train()
        m = self.embedding_model
        device = m.devices[0]
        states = N.require_cuda_f32(states.to(device), "states")
        assert states.dim() == 3 and states.size(2) == m.config.state_dims[0]
        params = list(m.parameters())
        loss, loss_rec = _FusedKoopmanLoss.apply(m, states, self.W_RECON, self.W_DYN, self.W_REG, *params)
        m.kMatrix = m._dense_operator(device)
        return loss, loss_rec

    @torch.no_grad()
    def evaluate(self, states: Tensor):
        """One-step reconstruction error (embedding_lorenz.py:223-251)."""
        self.embedding_model.eval()
        m = self.embedding_model
        device = m.devices[0]
        y_target = states[:, 1:].to(device)
        x_input = states[:, :-1].to(device)
        B, T = x_input.size(0), x_input.size(1)
        flat = N.require_cuda_f32(x_input.reshape(B * T, -1), "states")
        y_pred = m.recover(m.embed(flat)).view(B, T, -1)
        return nn.functional.mse_loss(y_target, y_pred), y_pred, y_target
