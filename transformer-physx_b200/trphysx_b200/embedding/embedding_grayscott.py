"""Drop-in `GrayScottEmbedding` (reference: trphysx/embedding/embedding_grayscott.py:25-218).

embed = normalise + 4 x [circular Conv3d + BatchNorm3d + LeakyReLU] + 2 Linear + LayerNorm; recover = 2 Linear + 1^3 conv +
3 x [trilinear x2 + circular Conv3d + BatchNorm3d + LeakyReLU] + 1^3 conv + un-normalise; koopmanOperation = skew-symmetric
banded matvec (9 off-diagonals) — all in libtrpx (csrc/gs.cu).  The kernels implement the EVAL-mode network (BatchNorm3d
with its running statistics); train mode (batch statistics) raises."""
import ctypes
from typing import Tuple

import numpy as np
import torch
from torch import nn

from .. import _native as N
from .embedding_model import EmbeddingModel, EmbeddingTrainingHead

Tensor = torch.Tensor


def _conv(cin, cout, k, stride):
    return nn.Conv3d(cin, cout, kernel_size=(k, k, k), stride=stride, padding=k // 2, padding_mode="circular")


def _up():
    return nn.Upsample(scale_factor=2, mode="trilinear", align_corners=False)


class GrayScottEmbedding(EmbeddingModel):
    model_name = "embedding_grayscott"

    def __init__(self, config) -> None:
        super().__init__(config)
        E = config.n_embd
        act = lambda: nn.LeakyReLU(0.02, inplace=True)          # noqa: E731
        self.observableNet = nn.Sequential(
            _conv(2, 64, 5, 2), nn.BatchNorm3d(64), act(), _conv(64, 128, 3, 2), nn.BatchNorm3d(128), act(),
            _conv(128, 128, 3, 2), nn.BatchNorm3d(128), act(), _conv(128, 64, 1, 1), nn.BatchNorm3d(64), act())
        self.observableNetFC = nn.Sequential(
            nn.Linear(64 * 4 * 4 * 4, 8 * 4 * 4 * 4), act(), nn.Linear(8 * 4 * 4 * 4, E),
            nn.LayerNorm(E, eps=config.layer_norm_epsilon))
        self.recoveryNetFC = nn.Sequential(
            nn.Linear(E, 8 * 4 * 4 * 4), nn.LeakyReLU(1.0, inplace=True), nn.Linear(8 * 4 * 4 * 4, 64 * 4 * 4 * 4), act())
        self.recoveryNet = nn.Sequential(
            _conv(64, 128, 1, 1), nn.BatchNorm3d(128), act(), _up(), _conv(128, 128, 3, 1), nn.BatchNorm3d(128), act(),
            _up(), _conv(128, 64, 3, 1), nn.BatchNorm3d(64), act(), _up(), _conv(64, 64, 3, 1), nn.BatchNorm3d(64), act(),
            _conv(64, 2, 1, 1))
        self.kMatrixDiag = nn.Parameter(torch.ones(E))
        self.xidx = torch.LongTensor(np.concatenate([np.arange(0, E - i) for i in range(1, 10)]))
        self.yidx = torch.LongTensor(np.concatenate([np.arange(i, E) for i in range(1, 10)]))
        self.kMatrixUT = nn.Parameter(0.01 * torch.rand(self.xidx.size(0)))
        self.register_buffer("mu", torch.tensor(0.0))
        self.register_buffer("std", torch.tensor(1.0))
        self._handles = N.HandleCache("trpx_gs_destroy")
        self._last_B = None

    def _weight_list(self):
        ws = [self.mu, self.std]
        for net, idxs in ((self.observableNet, (0, 3, 6, 9)), (self.recoveryNet, (0, 4, 8, 12))):
            for i in idxs:
                bn = net[i + 1]
                ws += [net[i].weight, net[i].bias, bn.weight, bn.bias, bn.running_mean, bn.running_var]
        fc, rf = self.observableNetFC, self.recoveryNetFC
        ws += [fc[0].weight, fc[0].bias, fc[2].weight, fc[2].bias, fc[3].weight, fc[3].bias]
        ws += [rf[0].weight, rf[0].bias, rf[2].weight, rf[2].bias]
        ws += [self.recoveryNet[15].weight, self.recoveryNet[15].bias, self.kMatrixDiag, self.kMatrixUT]
        return ws

    def _handle(self, device: torch.device):
        lib = N.lib()
        if self.training:
            raise NotImplementedError("GrayScottEmbedding kernels implement eval mode (BatchNorm3d running statistics); "
                                      "call model.eval() first")
        idx = device.index if device.index is not None else torch.cuda.current_device()
        entry = self._handles.get(idx)
        if entry is None:
            h = ctypes.c_void_p()
            N.check(lib.trpx_gs_create(self.config.n_embd, self.config.layer_norm_epsilon, idx, ctypes.byref(h)))
            entry = (h, N.WeightTracker())
            self._handles[idx] = entry
        h, tracker = entry
        ws = self._weight_list()
        if tracker.changed(ws):
            tensors = []
            for w in ws:
                if w.device != device:
                    raise NotImplementedError(f"parameter on {w.device} but input on {device}: move the model first")
                tensors.append(N.require_cuda_f32(w.detach().reshape(-1), "parameter"))
            N.check(lib.trpx_gs_load_weights(h, N.ptr_array(tensors), len(tensors), N.stream_ptr(device)))
            self._keepalive = tensors
        return h

    def invalidate_weights(self):
        self._handles.invalidate()

    def _check_x(self, x: Tensor) -> Tensor:
        x = N.require_cuda_f32(x, "x")
        assert x.dim() == 5 and tuple(x.shape[1:]) == (2, 32, 32, 32), "x must be [B, 2, 32, 32, 32]"
        return x

    def _embed(self, x: Tensor, want_g0: bool):
        B = x.size(0)
        g = torch.empty((B, self.config.n_embd), device=x.device, dtype=torch.float32)
        g0 = torch.empty((B, 64, 4, 4, 4), device=x.device, dtype=torch.float32) if want_g0 else None
        N.check(N.lib().trpx_gs_embed(self._handle(x.device), x.data_ptr(), g.data_ptr(),
                                      g0.data_ptr() if want_g0 else None, B, N.stream_ptr(x.device)))
        return g, g0

    def _recover(self, g: Tensor, want_out0: bool):
        B = g.size(0)
        x = torch.empty((B, 2, 32, 32, 32), device=g.device, dtype=torch.float32)
        out0 = torch.empty((B, 64, 4, 4, 4), device=g.device, dtype=torch.float32) if want_out0 else None
        N.check(N.lib().trpx_gs_recover(self._handle(g.device), g.data_ptr(), x.data_ptr(),
                                        out0.data_ptr() if want_out0 else None, B, N.stream_ptr(g.device)))
        return x, out0

    def embed(self, x: Tensor) -> Tensor:
        """[B,2,32,32,32] -> [B, n_embd]  (embedding_grayscott.py:127-139)"""
        self._inference_only("embed")
        return self._embed(self._check_x(x), False)[0]

    def recover(self, g: Tensor) -> Tensor:
        """[B, n_embd] -> [B,2,32,32,32]  (embedding_grayscott.py:141-153)"""
        self._inference_only("recover")
        g = N.require_cuda_f32(g, "g").reshape(-1, self.config.n_embd)
        return self._recover(g, False)[0]

    def forward(self, x: Tensor) -> Tuple[Tensor, ...]:
        """(g, xhat, g0, out0, unnormalize(recoveryNet(g0)))  (embedding_grayscott.py:112-125)"""
        self._inference_only("forward")
        x = self._check_x(x)
        g, g0 = self._embed(x, True)
        xhat, out0 = self._recover(g, True)
        x2 = torch.empty_like(xhat)
        N.check(N.lib().trpx_gs_decode_latent(self._handle(x.device), g0.data_ptr(), x2.data_ptr(), x.size(0),
                                              N.stream_ptr(x.device)))
        return g, xhat, g0, out0, x2

    def koopmanOperation(self, g: Tensor) -> Tensor:
        """[B, n_embd] -> [B, n_embd]  (embedding_grayscott.py:155-175)"""
        self._inference_only("koopmanOperation")
        g = N.require_cuda_f32(g, "g")
        out = torch.empty_like(g)
        N.check(N.lib().trpx_gs_koopman(self._handle(g.device), g.data_ptr(), out.data_ptr(), None, g.size(0),
                                        N.stream_ptr(g.device)))
        self._last_B = (g.size(0), g.device)
        return out

    @property
    def koopmanOperator(self, requires_grad: bool = True) -> Tensor:
        """Dense operator [B, n_embd, n_embd] of the last koopmanOperation call (the reference expands it per sample),
        materialised on demand."""
        if self._last_B is None:
            return None
        B, dev = self._last_B
        E = self.config.n_embd
        K = torch.empty((B, E, E), device=dev, dtype=torch.float32)
        zin = torch.zeros((B, E), device=dev, dtype=torch.float32)
        zout = torch.empty_like(zin)
        N.check(N.lib().trpx_gs_koopman(self._handle(dev), zin.data_ptr(), zout.data_ptr(), K.data_ptr(), B,
                                        N.stream_ptr(dev)))
        return K

    @property
    def koopmanDiag(self):
        return self.kMatrixDiag


class GrayScottEmbeddingTrainer(EmbeddingTrainingHead):
    """Training head (embedding_grayscott.py:220-306).  `evaluate` (an extension: the reference head has none) runs on the
    kernels.  `forward` (the training loss) needs BatchNorm3d batch statistics and their gradients, which the kernels do
    not implement: it raises instead of silently computing something else."""

    def __init__(self, config) -> None:
        super().__init__()
        self.embedding_model = GrayScottEmbedding(config)

    def forward(self, states: Tensor):
        raise NotImplementedError("GrayScottEmbeddingTrainer.forward: train-mode BatchNorm3d (batch statistics) and the "
                                  "3-D conv backward are not implemented in the kernels")

    @torch.no_grad()
    def evaluate(self, states: Tensor):
        """MSE of recover(embed(x_t)) against x_t in eval mode; returns (loss, y_pred, y_target) like the other heads."""
        self.embedding_model.eval()
        m = self.embedding_model
        device = m.devices[0]
        mse = nn.functional.mse_loss
        x = states.to(device)
        B, T = x.size(0), x.size(1)
        flat = N.require_cuda_f32(x.reshape(B * T, 2, 32, 32, 32), "states")
        y = m.recover(m.embed(flat)).view(B, T, 2, 32, 32, 32)
        return mse(x, y), y, x
