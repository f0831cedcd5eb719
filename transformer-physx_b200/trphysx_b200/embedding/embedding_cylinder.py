"""Drop-in `CylinderEmbedding` (reference: trphysx/embedding/embedding_cylinder.py:25-222).
embed = fused normalise + 5 replicate-padded 3x3 convs + LayerNorm; recover = 4 x (bilinear x2 + conv + ReLU)
+ conv + un-normalise + cylinder mask; koopmanOperation = viscosity-conditioned banded matvec — all in libtrpx."""
import ctypes
from typing import Tuple

import numpy as np
import torch
from torch import nn

from .. import _native as N
from .embedding_model import EmbeddingModel, EmbeddingTrainingHead

Tensor = torch.Tensor


def _conv(cin, cout, stride):
    return nn.Conv2d(cin, cout, kernel_size=(3, 3), stride=stride, padding=1, padding_mode="replicate")


def _up():
    return nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)


class CylinderEmbedding(EmbeddingModel):
    model_name = "embedding_cylinder"

    def __init__(self, config) -> None:
        super().__init__(config)
        E = config.n_embd
        X, Y = np.meshgrid(np.linspace(-2, 14, 128), np.linspace(-4, 4, 64))
        self.mask = torch.tensor(np.sqrt(X ** 2 + Y ** 2) < 1, dtype=torch.bool)
        self.observableNet = nn.Sequential(
            _conv(4, 16, 2), nn.ReLU(True), _conv(16, 32, 2), nn.ReLU(True), _conv(32, 64, 2), nn.ReLU(True),
            _conv(64, 128, 2), nn.ReLU(True), _conv(128, E // 32, 1))
        self.observableNetFC = nn.Sequential(nn.LayerNorm(E, eps=config.layer_norm_epsilon), nn.Dropout(config.embd_pdrop))
        self.recoveryNet = nn.Sequential(
            _up(), _conv(E // 32, 128, 1), nn.ReLU(), _up(), _conv(128, 64, 1), nn.ReLU(),
            _up(), _conv(64, 32, 1), nn.ReLU(), _up(), _conv(32, 16, 1), nn.ReLU(), _conv(16, 3, 1))
        self.obsdim = E
        self.kMatrixDiagNet = nn.Sequential(nn.Linear(1, 50), nn.ReLU(), nn.Linear(50, E))
        self.xidx = torch.LongTensor(np.concatenate([np.arange(0, E - i) for i in range(1, 5)]))
        self.yidx = torch.LongTensor(np.concatenate([np.arange(i, E) for i in range(1, 5)]))
        nband = self.xidx.size(0)
        self.kMatrixUT = nn.Sequential(nn.Linear(1, 50), nn.ReLU(), nn.Linear(50, nband))
        self.kMatrixLT = nn.Sequential(nn.Linear(1, 50), nn.ReLU(), nn.Linear(50, nband))
        self.register_buffer("mu", torch.tensor([0.0, 0.0, 0.0, 0.0]))
        self.register_buffer("std", torch.tensor([1.0, 1.0, 1.0, 1.0]))
        self._handles = N.HandleCache("trpx_cyl_destroy")
        self._last_visc = None
        self.kMatrixDiag = None

    def _weight_list(self):
        ws = [self.mu, self.std]
        for i in (0, 2, 4, 6, 8):
            ws += [self.observableNet[i].weight, self.observableNet[i].bias]
        ws += [self.observableNetFC[0].weight, self.observableNetFC[0].bias]
        for i in (1, 4, 7, 10, 12):
            ws += [self.recoveryNet[i].weight, self.recoveryNet[i].bias]
        for net in (self.kMatrixDiagNet, self.kMatrixUT, self.kMatrixLT):
            ws += [net[0].weight, net[0].bias, net[2].weight, net[2].bias]
        return ws

    def _handle(self, device: torch.device):
        lib = N.lib()
        if self.training and self.config.embd_pdrop:
            raise NotImplementedError("dropout > 0 in train mode is not implemented in the kernels")
        idx = device.index if device.index is not None else torch.cuda.current_device()
        entry = self._handles.get(idx)
        if entry is None:
            h = ctypes.c_void_p()
            N.check(lib.trpx_cyl_create(self.config.n_embd, self.config.layer_norm_epsilon, idx, ctypes.byref(h)))
            entry = (h, N.WeightTracker())
            self._handles[idx] = entry
            if getattr(self, "_conv_mode", None) is not None:
                N.check(lib.trpx_cyl_set_conv_mode(h, self._conv_mode))
            if getattr(self, "_chunk", None) is not None:
                N.check(lib.trpx_cyl_set_chunk(h, self._chunk))
        h, tracker = entry
        ws = self._weight_list()
        if tracker.changed(ws):
            tensors = []
            for w in ws:
                if w.device != device:
                    raise NotImplementedError(f"parameter on {w.device} but input on {device}: move the model first")
                tensors.append(N.require_cuda_f32(w.detach(), "parameter"))
            N.check(lib.trpx_cyl_load_weights(h, N.ptr_array(tensors), len(tensors), N.stream_ptr(device)))
            self._keepalive = tensors
        return h

    def invalidate_weights(self):
        """Force the packed native copy of the weights to be rebuilt on the next call.  Needed only after writes that
        torch's version counters cannot see (`p.data.mul_()`, EMA through `.data`, a custom kernel): parameter updates
        through autograd-visible in-place ops, optimizers, `load_state_dict`, `.to()` and the fused optimizers of
        `trphysx_b200.optim` are detected automatically."""
        self._handles.invalidate()

    def set_conv_mode(self, mode: int):
        """3 = tcgen05 (UMMA) implicit-GEMM decoder (default), 2 = pipelined 3xTF32 mma.sync decoder, 1 = FP32 SIMT direct
        convs, 0 = 3xTF32 mma.sync with in-kernel patch build."""
        self._conv_mode = mode
        for h, _ in self._handles.values():
            N.check(N.lib().trpx_cyl_set_conv_mode(h, mode))

    def set_chunk(self, frames: int):
        self._chunk = frames
        for h, _ in self._handles.values():
            N.check(N.lib().trpx_cyl_set_chunk(h, frames))


    @staticmethod
    def _visc(visc: Tensor, B: int) -> Tensor:
        visc = N.require_cuda_f32(visc, "visc")
        if visc.numel() != B:
            raise ValueError(f"visc must hold one viscosity per sample ([{B}, 1]), got {tuple(visc.shape)}")
        return visc.reshape(B)

    def embed(self, x: Tensor, visc: Tensor) -> Tensor:
        """[B,3,64,128], [B,1] -> [B, n_embd]  (embedding_cylinder.py:138-154)"""
        self._inference_only("embed")
        x = N.require_cuda_f32(x, "x")
        assert x.dim() == 4 and tuple(x.shape[1:]) == (3, 64, 128), "x must be [B, 3, 64, 128]"
        v = self._visc(visc, x.size(0))
        g = torch.empty((x.size(0), self.obsdim), device=x.device, dtype=torch.float32)
        N.check(N.lib().trpx_cyl_embed(self._handle(x.device), x.data_ptr(), v.data_ptr(), g.data_ptr(), x.size(0),
                                       N.stream_ptr(x.device)))
        return g

    def _decode(self, g: Tensor, apply_mask: bool) -> Tensor:
        g = N.require_cuda_f32(g, "g").reshape(-1, self.obsdim)
        x = torch.empty((g.size(0), 3, 64, 128), device=g.device, dtype=torch.float32)
        N.check(N.lib().trpx_cyl_recover(self._handle(g.device), g.data_ptr(), x.data_ptr(), g.size(0),
                                         1 if apply_mask else 0, N.stream_ptr(g.device)))
        return x

    def recover(self, g: Tensor) -> Tensor:
        """[B, n_embd] -> [B,3,64,128] with the cylinder cells zeroed (embedding_cylinder.py:156-170)"""
        self._inference_only("recover")
        return self._decode(g, True)

    def forward(self, x: Tensor, visc: Tensor) -> Tuple[Tensor, Tensor]:
        """(g, xhat).  As in the reference the mask is NOT applied here: `mask0 = ... is True` is the Python
        bool False (embedding_cylinder.py:133-134)."""
        g = self.embed(x, visc)
        return g, self._decode(g, False)

    def koopmanOperation(self, g: Tensor, visc: Tensor) -> Tensor:
        """[B, n_embd], [B,1] -> [B, n_embd]  (embedding_cylinder.py:172-196)"""
        self._inference_only("koopmanOperation")
        g = N.require_cuda_f32(g, "g")
        v = self._visc(visc, g.size(0))
        out = torch.empty_like(g)
        kdiag = torch.empty_like(g)
        N.check(N.lib().trpx_cyl_koopman(self._handle(g.device), g.data_ptr(), v.data_ptr(), out.data_ptr(),
                                         kdiag.data_ptr(), None, g.size(0), N.stream_ptr(g.device)))
        self.kMatrixDiag = kdiag
        self._last_visc = v
        return out

    @property
    def koopmanOperator(self, requires_grad: bool = True) -> Tensor:
        """Dense per-sample operator [B, n_embd, n_embd] of the last koopmanOperation call, materialised on demand."""
        if self._last_visc is None:
            return None
        v = self._last_visc
        B, E = v.numel(), self.obsdim
        K = torch.empty((B, E, E), device=v.device, dtype=torch.float32)
        dummy_in = torch.zeros((B, E), device=v.device, dtype=torch.float32)
        dummy_out = torch.empty_like(dummy_in)
        N.check(N.lib().trpx_cyl_koopman(self._handle(v.device), dummy_in.data_ptr(), v.data_ptr(), dummy_out.data_ptr(),
                                         None, K.data_ptr(), B, N.stream_ptr(v.device)))
        return K

    @property
    def koopmanDiag(self):
        return self.kMatrixDiag


class _FusedCylinderKoopmanLoss(torch.autograd.Function):
    """loss, loss_reconstruct = f(states, visc; params): forward AND backward run in one native call
    (csrc/cyl_train.cu); backward() only hands the stored flat gradient (nn.Module.parameters() order) to autograd."""

    @staticmethod
    def forward(ctx, model, states, visc, w_rec, w_dyn, w_reg, *params):
        lib = N.lib()
        dev = states.device
        h = model._handle(dev)
        B, T = states.size(0), states.size(1)
        nparam = int(lib.trpx_cyl_num_params(h))
        loss2 = torch.empty(2, device=dev, dtype=torch.float32)
        flat = torch.empty(nparam, device=dev, dtype=torch.float32)
        N.check(lib.trpx_cyl_train_fwd_bwd(h, states.data_ptr(), visc.data_ptr(), B, T, w_rec, w_dyn, w_reg, 1.0,
                                           loss2.data_ptr(), flat.data_ptr(), N.stream_ptr(dev)))
        ctx.flat = flat
        ctx.shapes = [p.shape for p in params]
        loss, loss_rec = loss2[0].clone(), loss2[1].clone()
        ctx.mark_non_differentiable(loss_rec)       # the reference accumulates loss_reconstruct detached (:250,275)
        return loss, loss_rec

    @staticmethod
    def backward(ctx, gloss, _grec):
        grads, off = [], 0
        for shp in ctx.shapes:
            n = int(np.prod(shp))
            grads.append((ctx.flat[off:off + n] * gloss).view(shp))
            off += n
        return (None, None, None, None, None, None) + tuple(grads)


class CylinderEmbeddingTrainer(EmbeddingTrainingHead):
    """Training head: forward(states [B,T,3,64,128], viscosity [B]) -> (loss, loss_reconstruct)
    (embedding_cylinder.py:224-281).  With gradients enabled the loss and the gradient of every parameter come from ONE
    fused native forward+backward (`loss.backward()` only scatters it); under torch.no_grad() the value is computed
    with the inference kernels."""
    W_REC, W_DYN, W_REG = 1e1, 1e1, 1e-2

    def __init__(self, config) -> None:
        super().__init__()
        self.embedding_model = CylinderEmbedding(config)

    def forward(self, states: Tensor, viscosity: Tensor):
        """(loss, loss_reconstruct) of embedding_cylinder.py:236-281: 1e1 MSE(x, recover(embed(x))) per step (decoder
        without mask, as `forward()` of the reference) + 1e1 MSE(x_t, recover(K g_{t-1})) (with mask)
        + 1e-2 sum(K^2) over the per-sample operators."""
        assert states.size(0) == viscosity.size(0), \
            "State variable and viscosity tensor should have the same batch dimensions."
        m = self.embedding_model
        device = m.devices[0]
        if torch.is_grad_enabled() and any(p.requires_grad for p in m.parameters()):
            m.train()
            st = N.require_cuda_f32(states.to(device), "states")
            assert st.dim() == 5 and tuple(st.shape[2:]) == (3, 64, 128), "states must be [B, T, 3, 64, 128]"
            v = m._visc(viscosity.to(device), st.size(0))
            loss, loss_rec = _FusedCylinderKoopmanLoss.apply(m, st, v, self.W_REC, self.W_DYN, self.W_REG, *list(m.parameters()))
            m._last_visc = v
            return loss, loss_rec
        mse = nn.functional.mse_loss
        viscosity = viscosity.to(device)
        xin0 = states[:, 0].to(device)
        g0, xrec0 = m(xin0, viscosity)
        loss = 1e1 * mse(xin0, xrec0)
        loss_reconstruct = mse(xin0, xrec0)
        g_old = g0
        for t0 in range(1, states.shape[1]):
            xin0 = states[:, t0].to(device)
            _, xrec1 = m(xin0, viscosity)
            g_pred = m.koopmanOperation(g_old, viscosity)
            xgrec1 = m.recover(g_pred)
            loss = loss + 1e1 * mse(xgrec1, xin0) + 1e1 * mse(xrec1, xin0) \
                + 1e-2 * torch.sum(torch.pow(m.koopmanOperator, 2))
            loss_reconstruct = loss_reconstruct + mse(xrec1, xin0)
            g_old = g_pred
        return loss, loss_reconstruct

    @torch.no_grad()
    def evaluate(self, states: Tensor, viscosity: Tensor):
        """One-step reconstruction error (embedding_cylinder.py:283-313)."""
        self.embedding_model.eval()
        m = self.embedding_model
        device = m.devices[0]
        y_target = states[:, 1:].to(device)
        x_input = states[:, :-1].to(device)
        viscosity = viscosity.to(device)
        B, T = x_input.size(0), x_input.size(1)
        flat = N.require_cuda_f32(x_input.reshape(B * T, 3, 64, 128), "states")
        visc = viscosity.reshape(B, 1).repeat_interleave(T, dim=0)
        y_pred = m.recover(m.embed(flat, visc)).view(B, T, 3, 64, 128)
        return nn.functional.mse_loss(y_target, y_pred), y_pred, y_target
