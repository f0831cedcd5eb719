"""Drop-in `PhysformerGPT2` whose forward()/generate() run hand-written sm_100a kernels (libtrpx).

The nn.Module tree below exists only to own the parameters under the reference's state_dict names
(trphysx/transformer/phys_transformer_gpt2.py:126-138, attention.py:49-71, utils.py:37-40) so that
reference checkpoints load unchanged and optimizers / save_model / .to() keep working.  No arithmetic
is done in Python: forward() -> trpx_gpt2_forward, generate() -> trpx_gpt2_rollout (one launch for the
whole horizon).  Unsupported reference features raise NotImplementedError — there is no fallback.
"""
import ctypes
import logging
import os
from typing import List, Optional, Tuple

import torch
from torch import nn

from .. import _native as N

logger = logging.getLogger(__name__)
Tensor = torch.Tensor


class Conv1D(nn.Module):
    """Parameter holder for the GPT-style linear layer: weight [nx, nf] (in, out), bias [nf] (utils.py:37-40)."""

    def __init__(self, nf: int, nx: int):
        super().__init__()
        self.nf = nf
        self.weight = nn.Parameter(torch.empty(nx, nf).normal_(std=0.02))
        self.bias = nn.Parameter(torch.zeros(nf))


class MaskedAttention(nn.Module):
    def __init__(self, nx: int, n_ctx: int, config, scale: bool = False, mask: str = "tril"):
        super().__init__()
        assert nx % config.n_head == 0
        if mask != "tril":
            raise NotImplementedError(f"mask type {mask!r}: only the causal 'tril' mask is implemented in the kernels")
        self.register_buffer("bias", torch.tril(torch.ones((n_ctx, n_ctx), dtype=torch.uint8)).view(1, 1, n_ctx, n_ctx))
        self.register_buffer("masked_bias", torch.tensor(-1e4))
        self.n_head = config.n_head
        self.split_size = nx
        self.scale = scale
        self.c_attn = Conv1D(nx * 3, nx)
        self.c_proj = Conv1D(nx, nx)


class MLP(nn.Module):
    def __init__(self, n_state: int, config):
        super().__init__()
        self.c_fc = Conv1D(n_state, config.n_embd)
        self.c_proj = Conv1D(config.n_embd, n_state)


class Block(nn.Module):
    def __init__(self, n_ctx: int, config, scale: bool = False):
        super().__init__()
        nx = config.n_embd
        self.ln_1 = nn.LayerNorm(nx, eps=config.layer_norm_epsilon)
        self.attn = MaskedAttention(nx, n_ctx, config, scale)
        self.ln_2 = nn.LayerNorm(nx, eps=config.layer_norm_epsilon)
        self.mlp = MLP(4 * nx, config)


class PhysformerBase(nn.Module):
    """save/load and initialisation shared by the transformer modules (phys_transformer_base.py:20-137)."""
    model_name: str = "transformer_model"

    def __init__(self, config):
        super().__init__()
        self.config = config

    def _init_weights(self, module):
        # `.data` writes do not bump torch's version counters, so tell the packed-weight caches explicitly
        N.bump_weight_epoch()
        if isinstance(module, (nn.Linear, nn.Embedding, Conv1D)):
            module.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
            if isinstance(module, (nn.Linear, Conv1D)) and module.bias is not None:
                module.bias.data.zero_()
        elif isinstance(module, nn.LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)

    def _num_parameters(self) -> int:
        return sum(p.numel() for p in self.parameters())

    def save_model(self, save_directory: str, epoch: int = 0) -> None:
        if os.path.isfile(save_directory):
            raise AssertionError(f"Provided path ({save_directory}) should be a directory, not a file")
        os.makedirs(save_directory, exist_ok=True)
        torch.save(self.state_dict(), os.path.join(save_directory, f"{self.model_name}{epoch:d}.pth"))

    def load_model(self, file_or_path_directory: str, epoch: int = 0) -> None:
        if os.path.isfile(file_or_path_directory):
            path = file_or_path_directory
        elif os.path.isdir(file_or_path_directory):
            path = os.path.join(file_or_path_directory, f"{self.model_name}{epoch:d}.pth")
        else:
            raise FileNotFoundError(f"Provided path or file ({file_or_path_directory}) does not exist")
        self.load_state_dict(torch.load(path, map_location=lambda storage, loc: storage))


def position_table(n_ctx: int, n_embd: int) -> Tensor:
    """[n_ctx, n_embd] table of the reference's computed position embedding (phys_transformer_gpt2.py:215-219),
    evaluated with the same torch float32 ops.  Even channels sin(p/10000^(2j/E)); because of the reference's
    `i[:, :, n_embd % 2]` indexing every odd channel is cos(p / 10000^(2*i[n_embd%2]/E)) = cos(p) for even E."""
    pos = torch.arange(0, n_ctx, dtype=torch.float).view(1, n_ctx)
    i = torch.arange(0, n_embd // 2, dtype=torch.float).unsqueeze(0).unsqueeze(0)
    pe = torch.zeros(1, n_ctx, n_embd)
    pe[:, :, ::2] = torch.sin(pos.unsqueeze(-1) / 10000 ** (2 * i / n_embd))
    i0 = i[:, :, n_embd % 2]
    pe[:, :, 1::2] = torch.cos(pos.unsqueeze(-1) / 10000 ** (2 * i0 / n_embd))
    return pe[0].contiguous()


class PhysformerGPT2(PhysformerBase):
    """Same constructor, parameters and call signatures as the reference class
    (trphysx/transformer/phys_transformer_gpt2.py:120-269 + GenerationMixin, generate_utils.py:21-173)."""

    def __init__(self, config, model_name: str = None) -> None:
        super().__init__(config)
        if getattr(config, "activation_function", "gelu_new") != "gelu_new":
            raise NotImplementedError("only activation_function='gelu_new' is implemented in the kernels")
        self.output_hidden_states = config.output_hidden_states
        self.drop = nn.Dropout(config.embd_pdrop)
        self.h = nn.ModuleList([Block(config.n_ctx, config, scale=True) for _ in range(config.n_layer)])
        self.ln_f = nn.LayerNorm(config.n_embd, eps=config.layer_norm_epsilon)
        self.mlp_f = nn.Linear(config.n_embd, config.n_embd)
        self.wpe = nn.Embedding(config.n_ctx, config.n_embd)      # in the state_dict, unused by forward()
        self.apply(self._init_weights)
        self.n_embd = config.n_embd
        if model_name is not None:
            self.model_name = "transformer_" + model_name
        self._handles = N.HandleCache("trpx_gpt2_destroy")            # device index -> (native handle, WeightTracker)
        self._pe = None
        self._tuning = (0, 0, 0, 0)
        logger.info("Number of parameters: {}".format(self._num_parameters()))

    # ------------------------------------------------------------------ native plumbing
    def _weight_list(self) -> List[Tensor]:
        ws = []
        for blk in self.h:
            ws += [blk.ln_1.weight, blk.ln_1.bias, blk.attn.c_attn.weight, blk.attn.c_attn.bias,
                   blk.attn.c_proj.weight, blk.attn.c_proj.bias, blk.ln_2.weight, blk.ln_2.bias,
                   blk.mlp.c_fc.weight, blk.mlp.c_fc.bias, blk.mlp.c_proj.weight, blk.mlp.c_proj.bias]
        return ws + [self.ln_f.weight, self.ln_f.bias, self.mlp_f.weight, self.mlp_f.bias]

    def _check_supported(self):
        c = self.config
        if self.training and (c.resid_pdrop or c.embd_pdrop or c.attn_pdrop):
            raise NotImplementedError("dropout > 0 in train mode is not implemented in the kernels")
        if self.output_hidden_states:
            raise NotImplementedError("output_hidden_states is not implemented in the kernels")

    def _handle(self, device: torch.device):
        lib = N.lib()
        idx = device.index if device.index is not None else torch.cuda.current_device()
        entry = self._handles.get(idx)
        if entry is None:
            c = self.config
            cfg = N.Gpt2Config(c.n_ctx, c.n_embd, c.n_layer, c.n_head, c.layer_norm_epsilon)
            h = ctypes.c_void_p()
            N.check(lib.trpx_gpt2_create(ctypes.byref(cfg), idx, ctypes.byref(h)))
            entry = (h, N.WeightTracker())
            self._handles[idx] = entry
            N.check(lib.trpx_gpt2_set_rollout_tuning(h, *self._tuning))
        h, tracker = entry
        ws = self._weight_list()
        if tracker.changed(ws):
            tensors = []
            for w in ws:
                if w.device != device:
                    raise NotImplementedError(f"parameter on {w.device} but input on {device}: move the model first")
                tensors.append(N.require_cuda_f32(w.detach(), "parameter"))
            if self._pe is None or self._pe.device != device:
                self._pe = position_table(self.config.n_ctx, self.config.n_embd).to(device)
            N.check(lib.trpx_gpt2_load_weights(h, N.ptr_array(tensors), len(tensors), self._pe.data_ptr(),
                                               N.stream_ptr(device)))
            self._keepalive = tensors
        return h

    def invalidate_weights(self):
        """Force the packed native copy of the weights to be rebuilt on the next call.  Needed only after writes that
        torch's version counters cannot see (`p.data.mul_()`, EMA through `.data`, a custom kernel): parameter updates
        through autograd-visible in-place ops, optimizers, `load_state_dict`, `.to()` and the fused optimizers of
        `trphysx_b200.optim` are detected automatically."""
        self._handles.invalidate()

    # ------------------------------------------------------------------ fused training step
    def _blob_views(self, blob: Tensor) -> List[Tensor]:
        """Views of a packed blob (csrc/gpt2_layout.cuh) in `_weight_list()` order and parameter shapes."""
        E = self.config.n_embd
        sizes = []
        for _ in self.h:
            sizes += [(E,), (E,), (E, 3 * E), (3 * E,), (E, E), (E,), (E,), (E,), (E, 4 * E), (4 * E,), (4 * E, E), (E,)]
        sizes += [(E,), (E,), (E, E), (E,)]
        views, o = [], 0
        for shp in sizes:
            n = 1
            for d in shp:
                n *= d
            views.append(blob[o:o + n].view(*shp))
            o += n
        assert o == blob.numel()
        views[-2] = views[-2].t()          # mlp_f.weight is stored transposed ([in][out]) in the blob
        return views

    def fused_loss_and_grads(self, inputs_embeds: Tensor, labels_embeds: Tensor):
        """loss = MSE(forward(inputs_embeds), labels_embeds) and its gradient for every parameter of the stack in ONE
        launch pair (csrc/gpt2_train.cu).  Returns (loss [scalar tensor], hidden [B,T,E], grads in `_weight_list()`
        order)."""
        self._check_supported()
        x = N.require_cuda_f32(inputs_embeds, "inputs_embeds")
        y = N.require_cuda_f32(labels_embeds, "labels_embeds")
        if x.dim() != 3 or x.shape != y.shape or x.shape[-1] != self.config.n_embd:
            raise ValueError("inputs_embeds and labels_embeds must both be [B, T, n_embd]")
        B, T, E = x.shape
        h = self._handle(x.device)
        total = N.lib().trpx_gpt2_blob_size(h)
        loss = torch.empty((), device=x.device, dtype=torch.float32)
        hidden = torch.empty_like(x)
        gblob = torch.empty(total, device=x.device, dtype=torch.float32)
        N.check(N.lib().trpx_gpt2_train_fwd_bwd(h, x.data_ptr(), y.data_ptr(), B, T, loss.data_ptr(), hidden.data_ptr(),
                                                gblob.data_ptr(), N.stream_ptr(x.device)))
        return loss, hidden, self._blob_views(gblob)

    def rollout_round_trajectories(self, device, use_cache: bool = True):
        """Trajectories the window kernel keeps resident per round on `device` (one persistent CTA per SM, 8 warps, 8
        trajectories per warp for n_ctx <= 32, 4 for n_ctx = 64: csrc/gpt2.cu, trpx_gpt2_rollout), or None where the rollout
        is not organised in such rounds.  Callers that pipeline work around generate() slice large batches by it."""
        c = self.config
        if not use_cache or c.n_embd != 32 or c.n_head != 4 or c.n_ctx > 64 or getattr(self, "_tuning", (0, 0, 0, 0)) != (0, 0, 0, 0):
            return None
        sms = torch.cuda.get_device_properties(device).multi_processor_count
        return sms * 8 * (8 if c.n_ctx <= 32 else 4)

    def set_rollout_tuning(self, variant: int = 0, warps_per_cta: int = 0, traj_per_warp: int = 0, max_ctas: int = 0):
        """Kernel tuning knobs (0 = automatic); see trpx_gpt2_set_rollout_tuning in include/trpx.h."""
        self._tuning = (variant, warps_per_cta, traj_per_warp, max_ctas)
        for h, _ in self._handles.values():
            N.check(N.lib().trpx_gpt2_set_rollout_tuning(h, *self._tuning))

    def last_launch(self, device=None) -> dict:
        idx = torch.cuda.current_device() if device is None else torch.device(device).index
        h, _ = self._handles[idx]
        vals = [ctypes.c_int() for _ in range(5)]
        N.check(N.lib().trpx_gpt2_last_launch(h, *[ctypes.byref(v) for v in vals]))
        keys = ("variant", "grid", "block", "smem", "group")
        return {k: v.value for k, v in zip(keys, vals)}


    # ------------------------------------------------------------------ reference API
    def forward(self, inputs_embeds: Tensor, position_ids: Tensor = None, prop_embeds: Tensor = None,
                past: List[Tensor] = None, attention_mask=None, head_mask=None, use_cache: bool = True,
                output_attentions: bool = False) -> Tuple:
        for name, val in (("position_ids", position_ids), ("prop_embeds", prop_embeds),
                          ("attention_mask", attention_mask), ("head_mask", head_mask)):
            if val is not None:
                raise NotImplementedError(f"{name} is not implemented in the kernels (reference default None only)")
        self._check_supported()
        x = N.require_cuda_f32(inputs_embeds, "inputs_embeds")
        assert x.dim() == 3 and x.size(-1) == self.config.n_embd, "inputs_embeds must be [B, T, n_embd]"
        B, T, E = x.shape
        c = self.config
        H, hd = c.n_head, E // c.n_head
        dev = x.device
        h = self._handle(dev)
        Tp = 0
        past_arr = None
        keep = []
        if past is not None:
            assert len(past) == c.n_layer
            Tp = past[0][0].size(-2)
            if Tp > 0:
                for p_l in past:
                    pl = N.require_cuda_f32(p_l, "past")
                    assert tuple(pl.shape) == (2, B, H, Tp, hd), f"past layer must be [2,{B},{H},{Tp},{hd}]"
                    keep.append(pl)
                past_arr = N.ptr_array(keep)
        if Tp + T > c.n_ctx:
            raise ValueError(f"past ({Tp}) + tokens ({T}) exceeds n_ctx ({c.n_ctx})")
        hidden = torch.empty_like(x)
        presents = attn = None
        pres_arr = attn_arr = None
        if use_cache is True:
            presents = tuple(torch.empty((2, B, H, Tp + T, hd), device=dev, dtype=torch.float32) for _ in range(c.n_layer))
            pres_arr = N.ptr_array(presents)
        if output_attentions:
            attn = tuple(torch.empty((B, H, T, Tp + T), device=dev, dtype=torch.float32) for _ in range(c.n_layer))
            attn_arr = N.ptr_array(attn)
        N.check(N.lib().trpx_gpt2_forward(h, x.data_ptr(), past_arr, Tp, hidden.data_ptr(), pres_arr, attn_arr,
                                          B, T, N.stream_ptr(dev)))
        outputs = (hidden,)
        if use_cache is True:
            outputs = outputs + (presents,)
        if output_attentions:
            outputs = outputs + (attn,)
        return outputs

    @torch.no_grad()
    def generate(self, inputs_embeds: Tensor, position_ids: Tensor = None, prop_embeds: Tensor = None,
                 max_length: int = None, attention_mask=None, use_cache: bool = False,
                 return_presents: bool = True, **model_specific_kwargs) -> Tuple:
        """[B, c, E] context -> ([B, max_length, E],) (+ (presents,) when use_cache).  As in the reference only the
        LAST context token conditions the rollout (generate_utils.py:51-54).  `return_presents=False` (an extension)
        skips materialising the final KV windows."""
        max_length = max_length if max_length is not None else self.config.max_length
        use_cache = use_cache if use_cache is not None else self.config.use_cache
        assert isinstance(max_length, int) and max_length > 0, "`max_length` should be a strictly positive integer."
        assert isinstance(use_cache, bool), "`use_cache` should be a boolean."
        for name, val in (("position_ids", position_ids), ("prop_embeds", prop_embeds),
                          ("attention_mask", attention_mask)):
            if val is not None:
                raise NotImplementedError(f"{name} is not implemented in the kernels (reference default None only)")
        if model_specific_kwargs:
            raise NotImplementedError(f"unsupported generate() arguments: {sorted(model_specific_kwargs)}")
        self._check_supported()
        x = N.require_cuda_f32(inputs_embeds, "inputs_embeds")
        assert x.dim() == 3 and x.size(-1) == self.config.n_embd, "inputs_embeds must be [B, T, n_embd]"
        B, cur_len, E = x.shape
        assert cur_len < max_length, (f"The input context is {cur_len}, but `max_length` is only {max_length}. "
                                      "Please make sure that `max_length` larger than the input")
        c = self.config
        dev = x.device
        h = self._handle(dev)
        S = max_length - cur_len
        out = torch.empty((B, max_length, E), device=dev, dtype=torch.float32)
        out[:, :cur_len] = x
        presents_buf = None
        Tk = min(S, c.n_ctx)
        if use_cache and return_presents:
            presents_buf = torch.empty((c.n_layer, 2, B, c.n_head, Tk, E // c.n_head), device=dev, dtype=torch.float32)
        x0 = out[:, cur_len - 1]
        N.check(N.lib().trpx_gpt2_rollout(
            h, x0.data_ptr(), out.stride(0), out[:, cur_len:].data_ptr(), out.stride(0), B, S, 1 if use_cache else 0,
            presents_buf.data_ptr() if presents_buf is not None else None, N.stream_ptr(dev)))
        if use_cache and return_presents:
            return (out, tuple(presents_buf[l] for l in range(c.n_layer)))
        return (out,)
