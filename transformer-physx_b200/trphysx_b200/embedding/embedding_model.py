"""Common parents of the Koopman embedding modules and of their training heads.

They carry the part of the reference interface that is not arithmetic (trphysx/embedding/embedding_model.py:19-147):
the `model_name` / checkpoint convention, the shape and device properties the trainers query, and the rule that the
compute methods of this package run CUDA kernels which do not record an autograd graph."""
from typing import List

import torch
from torch import nn

from ..checkpoint import read_state_dict, write_state_dict


class EmbeddingModel(nn.Module):
    """Parent of LorenzEmbedding / RosslerEmbedding / CylinderEmbedding.

    Subclasses provide `embed`, `recover`, `koopmanOperation`, `koopmanOperator` and `koopmanDiag`; all of them are
    ctypes calls into libtrpx on the current CUDA stream."""

    model_name: str = "embedding_model"

    def __init__(self, config):
        super().__init__()
        self.config = config

    # -- what subclasses must supply -----------------------------------------------------------------------------
    def embed(self, x):
        raise NotImplementedError(type(self).__name__ + " does not implement embed()")

    def recover(self, x):
        raise NotImplementedError(type(self).__name__ + " does not implement recover()")

    @property
    def koopmanOperator(self):
        raise NotImplementedError(type(self).__name__ + " does not implement koopmanOperator")

    @property
    def koopmanDiag(self):
        raise NotImplementedError(type(self).__name__ + " does not implement koopmanDiag")

    # -- bookkeeping the trainers rely on ------------------------------------------------------------------------
    @property
    def embed_dims(self) -> int:
        return self.config.n_embd

    @property
    def input_dims(self):
        return self.config.state_dims

    @property
    def num_parameters(self) -> int:
        total = 0
        for p in self.parameters():
            total += p.numel()
        return total

    @property
    def devices(self) -> List[torch.device]:
        """Distinct devices holding parameters or buffers, in first-seen order."""
        seen = {}
        for t in self.parameters():
            seen.setdefault(t.device, None)
        for t in self.buffers():
            seen.setdefault(t.device, None)
        return list(seen)

    def _inference_only(self, what: str):
        """The kernels behind embed/recover/koopmanOperation do not record an autograd graph."""
        if torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError(
                f"{what}: autograd through the native kernels is not provided. Use the fused training head "
                "(e.g. LorenzEmbeddingTrainer) for gradients, or call under torch.no_grad() / model.eval().")

    # -- checkpoints -----------------------------------------------------------------------------------------------
    def save_model(self, save_directory: str, epoch: int = 0) -> None:
        write_state_dict(self, save_directory, epoch)

    def load_model(self, file_or_path_directory: str, epoch: int = 0) -> None:
        read_state_dict(self, file_or_path_directory, epoch)


class EmbeddingTrainingHead(nn.Module):
    """Parent of the `*EmbeddingTrainer` heads: owns `embedding_model` and forwards checkpoint I/O to it."""

    def forward(self, *args, **kwargs):
        raise NotImplementedError(type(self).__name__ + " does not implement forward()")

    def evaluate(self, *args, **kwargs):
        raise NotImplementedError(type(self).__name__ + " does not implement evaluate()")

    def _model(self) -> EmbeddingModel:
        model = getattr(self, "embedding_model", None)
        assert model is not None, "the training head has no embedding_model yet"
        return model

    def save_model(self, *args, **kwargs):
        self._model().save_model(*args, **kwargs)

    def load_model(self, *args, **kwargs):
        self._model().load_model(*args, **kwargs)
