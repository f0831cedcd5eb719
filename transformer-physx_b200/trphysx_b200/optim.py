"""Fused clip_grad_norm_ + Adam for the embedding training step (reference step epilogue:
trphysx/embedding/training/enn_trainer.py:97-99 with torch.optim.Adam(lr, weight_decay) from
examples/lorenz/train_lorenz_enn.py:61).  Parameters and Adam state live in ONE flat device buffer; the
nn.Parameters are re-pointed at views of it, so `clip + step` is two kernel launches in total."""
import torch

from . import _native as N


class FusedClipAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_norm=0.1):
        self.params = [p for p in params]
        assert self.params, "no parameters"
        dev = self.params[0].device
        if dev.type != "cuda":
            raise NotImplementedError("FusedClipAdam runs on CUDA devices only (no CPU fallback)")
        self.lr, self.betas, self.eps, self.weight_decay, self.max_norm = lr, betas, eps, weight_decay, max_norm
        self.sizes = [p.numel() for p in self.params]
        n = sum(self.sizes)
        self.flat = torch.empty(n, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(n, device=dev, dtype=torch.float32)
        self.m = torch.zeros(n, device=dev, dtype=torch.float32)
        self.v = torch.zeros(n, device=dev, dtype=torch.float32)
        self.norm = torch.zeros(1, device=dev, dtype=torch.float32)
        off = 0
        with torch.no_grad():
            for p, sz in zip(self.params, self.sizes):
                self.flat[off:off + sz].copy_(p.detach().reshape(-1))
                p.data = self.flat[off:off + sz].view(p.shape)            # parameters become views of the flat buffer
                p.grad = self.grad[off:off + sz].view(p.shape)
                off += sz
        self.step_count = 0

    def flat_grad(self) -> torch.Tensor:
        """The flat gradient buffer (all-reduce THIS for data parallel training)."""
        off = 0
        for p, sz in zip(self.params, self.sizes):
            if p.grad is None:
                self.grad[off:off + sz].zero_()
            elif p.grad.data_ptr() != self.grad[off:off + sz].data_ptr():
                self.grad[off:off + sz].copy_(p.grad.reshape(-1))
            off += sz
        return self.grad

    @torch.no_grad()
    def step(self) -> torch.Tensor:
        """clip to max_norm, then Adam.  Returns the pre-clip global gradient norm (device scalar)."""
        g = self.flat_grad()
        self.step_count += 1
        dev = self.flat.device
        N.check(N.lib().trpx_clip_adam(self.flat.data_ptr(), g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                       self.flat.numel(), self.step_count, self.lr, self.betas[0], self.betas[1],
                                       self.eps, self.weight_decay, self.max_norm, self.norm.data_ptr(),
                                       dev.index if dev.index is not None else torch.cuda.current_device(),
                                       N.stream_ptr(dev)))
        N.bump_weight_epoch()                       # packed-weight caches must notice the in-place update
        return self.norm[0]

    def zero_grad(self):
        self.grad.zero_()
        off = 0
        for p, sz in zip(self.params, self.sizes):
            p.grad = self.grad[off:off + sz].view(p.shape)
            off += sz
