from .phys_transformer_gpt2 import PhysformerGPT2, Conv1D, MaskedAttention, Block, MLP
from .phys_transformer_helpers import PhysformerTrain
from .auto import AutoPhysTransformer

__all__ = ["PhysformerGPT2", "PhysformerTrain", "AutoPhysTransformer", "Conv1D", "MaskedAttention", "Block", "MLP"]
