"""The data format either side of the rollout path: the reference's pickle feature cache."""
from .feature_cache import (cached_features_file, read_cache, write_cache, embed_series_blocks,  # noqa: F401
                            build_feature_cache)
