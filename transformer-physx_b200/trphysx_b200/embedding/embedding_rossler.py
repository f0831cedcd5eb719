"""Drop-in `RosslerEmbedding` (reference: examples/rossler/rossler_module/embedding_rossler.py:23-246).
Arithmetic identical to the Lorenz embedding; only the loss weights differ (1e3 instead of 1e4, :197,209)."""
from .embedding_lorenz import LorenzEmbedding, LorenzEmbeddingTrainer


class RosslerEmbedding(LorenzEmbedding):
    model_name = "embedding_rossler"


class RosslerEmbeddingTrainer(LorenzEmbeddingTrainer):
    W_RECON, W_DYN, W_REG = 1e3, 1.0, 1e-1

    def __init__(self, config):
        super(LorenzEmbeddingTrainer, self).__init__()
        self.embedding_model = RosslerEmbedding(config)
