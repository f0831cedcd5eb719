"""`PhysformerTrain`: the evaluation head that drives generate() (reference:
trphysx/transformer/phys_transformer_helpers.py:22-122).  evaluate()/generate() are on the rollout hot path;
the teacher-forced training forward() needs autograd through the decoder stack, which is a "next" row
 and raises until it is built."""
import torch
from torch import nn

from .phys_transformer_gpt2 import PhysformerBase


class PhysformerTrain(PhysformerBase):
    def __init__(self, config, transformer_model: PhysformerBase = None) -> None:
        super().__init__(config)
        self.transformer = transformer_model
        self.transformer.apply(self._init_weights)

    def forward(self, inputs_embeds, labels_embeds, **kwargs):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.transformer.parameters()):
            raise NotImplementedError(
                "teacher-forced training (autograd through the decoder stack) is not built yet; "
                "call under torch.no_grad() for the loss value only")
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
