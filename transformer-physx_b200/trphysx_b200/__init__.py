"""trphysx_b200 — B200-native rollout path behind the `trphysx` module API.

    from trphysx_b200.config import LorenzConfig
    from trphysx_b200.embedding import LorenzEmbedding
    from trphysx_b200.transformer import PhysformerGPT2

All compute runs in libtrpx.so (hand-written sm_100a CUDA, include/trpx.h); there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import config, data_utils, embedding, transformer  # noqa: F401
from .config import AutoPhysConfig, PhysConfig  # noqa: F401
from .embedding import AutoEmbeddingModel  # noqa: F401
from .transformer import AutoPhysTransformer, PhysformerGPT2, PhysformerTrain  # noqa: F401
from .parallel import shard_bounds, ShardedRollout, average_flat_grad  # noqa: F401
from ._native import bump_weight_epoch as invalidate_weights  # noqa: F401,E402
"""`invalidate_weights()`: call after modifying parameters through `.data` (or any other way torch's version
counters cannot see); optimizer steps, load_state_dict, .to() and attribute assignment are detected automatically."""
