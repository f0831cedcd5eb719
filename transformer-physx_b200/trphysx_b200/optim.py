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


class FusedKoopmanStep:
    """One Koopman-embedding training step with minimal host work: the fused forward+backward writes the flat
    gradient straight into the optimizer's buffer, then (data parallel) ONE all-reduce, then clip + Adam.
    Equivalent to the step body of trphysx/embedding/training/enn_trainer.py:93-102
    (`loss = model(states); loss.backward(); clip_grad_norm_(0.1); optimizer.step(); zero_grad()`), but without
    building an autograd graph: 3 native calls, ~8 kernel launches."""

    def __init__(self, head, optimizer: FusedClipAdam, world_size: int = 1):
        self.head, self.opt, self.world = head, optimizer, world_size
        dev = optimizer.flat.device
        self.loss2 = torch.zeros(2, device=dev, dtype=torch.float32)

    @torch.no_grad()
    def __call__(self, states: torch.Tensor, viscosity: torch.Tensor = None):
        """states [B, T, 3] (Lorenz / Rossler) or [B, T, 3, 64, 128] with viscosity [B] (Cylinder) on the model's
        device.  Returns the device tensor [loss, loss_reconstruct]."""
        m = self.head.embedding_model
        m.train()
        states = N.require_cuda_f32(states, "states")
        dev = states.device
        h = m._handle(dev)
        if viscosity is not None:                   # Cylinder head (embedding_cylinder.py:236-281)
            v = m._visc(viscosity.to(dev), states.size(0))
            N.check(N.lib().trpx_cyl_train_fwd_bwd(
                h, states.data_ptr(), v.data_ptr(), states.size(0), states.size(1), self.head.W_REC, self.head.W_DYN,
                self.head.W_REG, 1.0, self.loss2.data_ptr(), self.opt.grad.data_ptr(), N.stream_ptr(dev)))
        else:
            N.check(N.lib().trpx_mlpemb_train_fwd_bwd(
                h, states.data_ptr(), states.size(0), states.size(1), self.head.W_RECON, self.head.W_DYN,
                self.head.W_REG, 1.0, self.loss2.data_ptr(), self.opt.grad.data_ptr(), N.stream_ptr(dev)))
        if self.world > 1:
            from .parallel import average_flat_grad
            average_flat_grad(self.opt.grad, self.world)
        self.opt.step_count += 1
        o = self.opt
        N.check(N.lib().trpx_clip_adam(o.flat.data_ptr(), o.grad.data_ptr(), o.m.data_ptr(), o.v.data_ptr(),
                                       o.flat.numel(), o.step_count, o.lr, o.betas[0], o.betas[1], o.eps,
                                       o.weight_decay, o.max_norm, o.norm.data_ptr(),
                                       dev.index if dev.index is not None else torch.cuda.current_device(),
                                       N.stream_ptr(dev)))
        N.bump_weight_epoch()
        return self.loss2


class FusedTransformerStep:
    """One teacher-forced transformer training step with minimal host work.  Equivalent to the step
    body of trphysx/utils/trainer.py:193-207,244-277 (`loss = model(**inputs)[0]; loss.backward();
    clip_grad_norm_(max_grad_norm); optimizer.step(); zero_grad()` with torch.optim.Adam), but: the fused
    forward+backward writes the gradient as one blob, (data parallel) ONE all-reduce averages it, clip + Adam update
    the master parameters — kept in the same packed layout — in two launches, and the updated blob is handed to the
    kernels with one device copy.  The module's parameters are views of the master buffer, so state_dict(), save_model()
    and evaluate()/generate() see the current weights."""

    def __init__(self, transformer, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_norm=1.0,
                 world_size: int = 1):
        self.model, self.world = transformer, world_size
        self.lr, self.betas, self.eps, self.weight_decay, self.max_norm = lr, betas, eps, weight_decay, max_norm
        params = transformer._weight_list()
        dev = params[0].device
        if dev.type != "cuda":
            raise NotImplementedError("FusedTransformerStep runs on CUDA devices only (no CPU fallback)")
        n = sum(p.numel() for p in params)
        self.flat = torch.empty(n, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(n, device=dev, dtype=torch.float32)
        self.m = torch.zeros(n, device=dev, dtype=torch.float32)
        self.v = torch.zeros(n, device=dev, dtype=torch.float32)
        self.norm = torch.zeros(1, device=dev, dtype=torch.float32)
        self.loss = torch.zeros((), device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p, view in zip(params, transformer._blob_views(self.flat)):
                view.copy_(p.detach())
                p.data = view                      # mlp_f.weight becomes a transposed view: same values, blob strides
        self.step_count = 0
        N.bump_weight_epoch()

    @torch.no_grad()
    def __call__(self, inputs_embeds: torch.Tensor, labels_embeds: torch.Tensor) -> torch.Tensor:
        """inputs_embeds, labels_embeds: [B, T, n_embd] on the model's device.  Returns the loss (device scalar)."""
        x = N.require_cuda_f32(inputs_embeds, "inputs_embeds")
        y = N.require_cuda_f32(labels_embeds, "labels_embeds")
        dev = x.device
        lib = N.lib()
        h = self.model._handle(dev)
        N.check(lib.trpx_gpt2_train_fwd_bwd(h, x.data_ptr(), y.data_ptr(), x.size(0), x.size(1), self.loss.data_ptr(),
                                            None, self.grad.data_ptr(), N.stream_ptr(dev)))
        if self.world > 1:
            from .parallel import average_flat_grad
            average_flat_grad(self.grad, self.world)
        self.step_count += 1
        N.check(lib.trpx_clip_adam(self.flat.data_ptr(), self.grad.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                   self.flat.numel(), self.step_count, self.lr, self.betas[0], self.betas[1], self.eps,
                                   self.weight_decay, self.max_norm, self.norm.data_ptr(),
                                   dev.index if dev.index is not None else torch.cuda.current_device(),
                                   N.stream_ptr(dev)))
        N.check(lib.trpx_gpt2_load_blob(h, self.flat.data_ptr(), N.stream_ptr(dev)))
        return self.loss
