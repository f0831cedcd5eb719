"""Base classes of the drop-in embedding modules (reference: trphysx/embedding/embedding_model.py:19-147)."""
import os
from abc import abstractmethod

import torch
from torch import nn


class EmbeddingModel(nn.Module):
    model_name: str = "embedding_model"

    def __init__(self, config):
        super().__init__()
        self.config = config

    @abstractmethod
    def embed(self, x):
        raise NotImplementedError("embed function has not been properly overridden")

    @abstractmethod
    def recover(self, x):
        raise NotImplementedError("recover function has not been properly overridden")

    @property
    @abstractmethod
    def koopmanOperator(self):
        pass

    @property
    @abstractmethod
    def koopmanDiag(self):
        pass

    @property
    def input_dims(self):
        return self.config.state_dims

    @property
    def embed_dims(self):
        return self.config.n_embd

    @property
    def num_parameters(self):
        return sum(p.numel() for p in self.parameters())

    @property
    def devices(self):
        devs = []
        for t in list(self.parameters()) + list(self.buffers()):
            if t.device not in devs:
                devs.append(t.device)
        return devs

    def _inference_only(self, what: str):
        """The kernels behind embed/recover/koopmanOperation do not record an autograd graph."""
        if torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError(
                f"{what}: autograd through the native kernels is not provided. Use the fused training head "
                "(e.g. LorenzEmbeddingTrainer) for gradients, or call under torch.no_grad() / model.eval().")

    def save_model(self, save_directory: str, epoch: int = 0) -> None:
        if os.path.isfile(save_directory):
            raise FileNotFoundError(f"Provided path ({save_directory}) should be a directory, not a file")
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


class EmbeddingTrainingHead(nn.Module):
    @abstractmethod
    def forward(self, *args, **kwargs):
        raise NotImplementedError("forward function has not been properly overridden")

    @abstractmethod
    def evaluate(self, *args, **kwargs):
        raise NotImplementedError("evaluate function has not been properly overridden")

    def save_model(self, *args, **kwargs):
        assert self.embedding_model is not None, "Must initialize embedding model before saving."
        self.embedding_model.save_model(*args, **kwargs)

    def load_model(self, *args, **kwargs):
        assert self.embedding_model is not None, "Must initialize embedding model before loading."
        self.embedding_model.load_model(*args, **kwargs)
