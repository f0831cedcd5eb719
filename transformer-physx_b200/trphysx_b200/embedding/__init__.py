from .embedding_model import EmbeddingModel, EmbeddingTrainingHead
from .embedding_lorenz import LorenzEmbedding, LorenzEmbeddingTrainer
from .embedding_rossler import RosslerEmbedding, RosslerEmbeddingTrainer
from .embedding_cylinder import CylinderEmbedding, CylinderEmbeddingTrainer
from .embedding_grayscott import GrayScottEmbedding, GrayScottEmbeddingTrainer
from .embedding_auto import AutoEmbeddingModel, MODEL_MAPPING, TRAINING_MAPPING

__all__ = ["EmbeddingModel", "EmbeddingTrainingHead", "LorenzEmbedding", "LorenzEmbeddingTrainer",
           "RosslerEmbedding", "RosslerEmbeddingTrainer", "CylinderEmbedding", "CylinderEmbeddingTrainer", "GrayScottEmbedding", "GrayScottEmbeddingTrainer",
           "AutoEmbeddingModel", "MODEL_MAPPING", "TRAINING_MAPPING"]
