"""`AutoPhysTransformer`: named in the north star; the reference constructs `PhysformerGPT2(config, name)`
directly (examples/lorenz/train_lorenz_transformer.py:64), so this is a thin registry mirroring
AutoEmbeddingModel (trphysx/embedding/embedding_auto.py:37-130)."""
from collections import OrderedDict
from typing import Optional

from ..config import AutoPhysConfig
from .phys_transformer_gpt2 import PhysformerGPT2

TRANSFORMER_MAPPING = OrderedDict([("lorenz", PhysformerGPT2), ("rossler", PhysformerGPT2), ("cylinder", PhysformerGPT2)])


class AutoPhysTransformer:
    def __init__(self):
        raise EnvironmentError("AutoPhysTransformer should not be initiated directly. Use its class methods.")

    @classmethod
    def init_model(cls, model_name: str, config=None) -> PhysformerGPT2:
        if model_name not in TRANSFORMER_MAPPING:
            raise ValueError(f"Provided model name, {model_name}, not found in existing transformer models.")
        config = config if config is not None else AutoPhysConfig.load_config(model_name)
        return TRANSFORMER_MAPPING[model_name](config, model_name)

    @classmethod
    def load_model(cls, model_name: str, config=None, file_or_path_directory: Optional[str] = None, epoch: int = 0):
        model = cls.init_model(model_name, config)
        if file_or_path_directory is not None:
            model.load_model(file_or_path_directory, epoch)
        return model
