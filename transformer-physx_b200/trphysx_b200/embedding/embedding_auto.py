"""name -> embedding model registry (reference: trphysx/embedding/embedding_auto.py:19-130)."""
from collections import OrderedDict
from typing import Optional

from .embedding_cylinder import CylinderEmbedding, CylinderEmbeddingTrainer
from .embedding_grayscott import GrayScottEmbedding, GrayScottEmbeddingTrainer
from .embedding_lorenz import LorenzEmbedding, LorenzEmbeddingTrainer
from .embedding_rossler import RosslerEmbedding, RosslerEmbeddingTrainer

MODEL_MAPPING = OrderedDict([("lorenz", LorenzEmbedding), ("cylinder", CylinderEmbedding), ("grayscott", GrayScottEmbedding),
                             ("rossler", RosslerEmbedding)])
TRAINING_MAPPING = OrderedDict([("lorenz", LorenzEmbeddingTrainer), ("cylinder", CylinderEmbeddingTrainer),
                                ("grayscott", GrayScottEmbeddingTrainer), ("rossler", RosslerEmbeddingTrainer)])


class AutoEmbeddingModel:
    def __init__(self):
        raise EnvironmentError("AutoEmbeddingModel should not be initiated directly. The class methods should be used instead.")

    @classmethod
    def init_model(cls, model_name: str, config):
        if model_name not in MODEL_MAPPING:
            raise ValueError("Provided model name, {:s}, not found in existing embedding models.".format(model_name))
        return MODEL_MAPPING[model_name](config)

    @classmethod
    def init_trainer(cls, model_name: str, config):
        if model_name not in TRAINING_MAPPING:
            raise KeyError("Provided model name, {:s}, not found in existing training models.".format(model_name))
        return TRAINING_MAPPING[model_name](config)

    @classmethod
    def load_model(cls, model_name: str, config, file_or_path_directory: Optional[str] = None, epoch: int = 0):
        model = cls.init_model(model_name, config)
        if file_or_path_directory is not None:
            model.load_model(file_or_path_directory, epoch)
        return model
