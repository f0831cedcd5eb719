"""`PhysformerTrain`: the training / evaluation head (reference:
trphysx/transformer/phys_transformer_helpers.py:22-122).  evaluate()/generate() are on the rollout hot path;
the teacher-forced training forward() returns a loss whose backward() fills the `.grad` of every
parameter of the stack from ONE fused forward+backward launch (csrc/gpt2_train.cu) — there is no autograd graph
through the decoder blocks."""
import torch
from torch import nn

from .phys_transformer_gpt2 import PhysformerBase


class _FusedStackLoss(torch.autograd.Function):
    """loss(params) with the gradient computed eagerly by the fused kernel; backward only scales it."""

    @staticmethod
    def forward(ctx, model, inputs_embeds, labels_embeds, *params):
        loss, hidden, grads = model.fused_loss_and_grads(inputs_embeds, labels_embeds)
        ctx.grads = grads
        ctx.mark_non_differentiable(hidden)
        return loss, hidden

    @staticmethod
    def backward(ctx, gloss, _ghidden):
        return (None, None, None) + tuple(g * gloss for g in ctx.grads)


class PhysformerTrain(PhysformerBase):
    def __init__(self, config, transformer_model: PhysformerBase = None) -> None:
        super().__init__(config)
        self.transformer = transformer_model
        self.transformer.apply(self._init_weights)

    def forward(self, inputs_embeds, labels_embeds, **kwargs):
        params = self.transformer._weight_list()
        if labels_embeds is not None and torch.is_grad_enabled() and any(p.requires_grad for p in params):
            if any(kwargs.get(k) for k in ("past", "attention_mask", "head_mask", "prop_embeds", "position_ids",
                                           "output_attentions")):
                raise NotImplementedError("the fused training step takes inputs_embeds and labels_embeds only")
            if inputs_embeds.requires_grad:
                raise NotImplementedError("gradients with respect to inputs_embeds are not computed by the fused step")
            loss, hidden = _FusedStackLoss.apply(self.transformer, inputs_embeds, labels_embeds, *params)
            # NOTE: the reference appends `presents` here when use_cache=True (its forward() default); the fused step
            # does not materialise the KV tensors (nothing in the training loop reads them, utils/trainer.py:244-256)
            return (loss, hidden, labels_embeds)
        outputs = self.transformer.forward(inputs_embeds=inputs_embeds, **kwargs)
        if labels_embeds is not None:
            hidden = outputs[0]
            loss = nn.functional.mse_loss(hidden, labels_embeds)
            outputs = (loss, hidden, labels_embeds) + outputs[1:]
        return outputs

    def evaluate(self, inputs_embeds, labels_embeds, **kwargs):
        """generate() + MSE against the target embeddings (phys_transformer_helpers.py:71-104)."""
        max_length = labels_embeds.size(1)
        outputs = self.transformer.generate(inputs_embeds=inputs_embeds, max_length=max_length, **kwargs)
        pred = outputs[0]
        error = nn.functional.mse_loss(pred, labels_embeds)
        return (error, pred, labels_embeds) + outputs[1:]

    def generate(self, *args, **kwargs):
        return self.transformer.generate(*args, **kwargs)

    def save_model(self, *args, **kwargs):
        self.transformer.save_model(*args, **kwargs)

    def load_model(self, *args, **kwargs):
        self.transformer.load_model(*args, **kwargs)
